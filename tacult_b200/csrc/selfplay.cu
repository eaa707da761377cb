// Device-resident self-play step (see selfplay.cuh).  One warp per env.
#include "selfplay.cuh"

namespace tc {

__device__ __forceinline__ void unpack_hdr2(const uint4 &h0, const uint4 &h1, Pos &p, uint32_t &n, uint32_t &off)
{
    p.mine[0] = h0.x; p.mine[1] = h0.y; p.mine[2] = h0.z;
    p.occ[0] = h0.w;  p.occ[1] = h1.x;  p.occ[2] = h1.y;
    p.last = h1.z & 0xFu;
    n = (h1.z >> 8) & 0xFFu;
    off = h1.w;
}

__global__ void k_selfplay_init(SelfPlayStore sp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.E) return;
    EnvState e;
    memset(&e, 0, sizeof e);
    e.player = 1;
    sp.env[i] = e;
    if (i == 0) {
        *sp.finished_count = 0;
        for (int k = 0; k < 8; ++k) sp.stats[k] = 0ull;
    }
}

__global__ void k_selfplay_roots(SelfPlayStore sp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sp.E) return;
    const EnvState &e = sp.env[i];
#pragma unroll
    for (int k = 0; k < 8; ++k) sp.root_states[8 * (size_t)i + k] = e.state[k];
    sp.root_active[i] = e.done ? 0 : 1;
}

// After the ply's simulations: visit counts -> move (coach.py:124-143), trajectory record (coach.py:136-138),
// game end bookkeeping (coach.py:143-155).
__global__ void __launch_bounds__(128) k_selfplay_move(TreeStore ts, SelfPlayStore sp, tc_selfplay_config cfg)
{
    const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (env >= sp.E) return;
    EnvState e = sp.env[env];
    if (e.done) return;
    const TreeCtl *c = ts.ctl + env;
    const int root = c->root_idx;
    if (!c->active || root < 0) {            // pools exhausted (error flag already raised): park the env
        if (lane == 0) sp.env[env].done = 1;
        return;
    }
    const uint4 *hdr = ts.node_hdr + 2ull * ((size_t)env * ts.capN + root);
    Pos p;
    uint32_t n, off;
    unpack_hdr2(hdr[0], hdr[1], p, n, off);
    Summary s = summarize(p);
    uint32_t lw[3], am[3];
    legal_words(p, s, lw);
    bool lg[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        int a = lane + 32 * q;
        lg[q] = (a < 81) && action_bit(lw, a);
        am[q] = __ballot_sync(FULL, lg[q]);
    }
    const Edge *blk = tree_edges(ts, env) + off;
    uint32_t cnt[3];
    uint32_t total = 0, maxc = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        cnt[q] = 0;
        if (lg[q]) {
            int j = __popc(am[q] & ((1u << lane) - 1u)) + (q >= 1 ? __popc(am[0]) : 0) + (q >= 2 ? __popc(am[1]) : 0);
            cnt[q] = blk[j].N;
        }
        total += cnt[q];
        maxc = max(maxc, cnt[q]);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        total += __shfl_xor_sync(FULL, total, d);
        maxc = max(maxc, __shfl_xor_sync(FULL, maxc, d));
    }
    const bool tau_one = (e.ply + 1) < cfg.temp_threshold;          // coach.py:124-125
    const uint64_t r = mix64(cfg.seed ^ mix64(((uint64_t)env * 0x9E3779B97F4A7C15ull) ^
                                              ((uint64_t)e.game * 0xD1B54A32D192ED03ull) ^
                                              ((uint64_t)e.ply * 0x8CB92BA72F3D8DD7ull)));
    // weight of every action: its count (tau = 1) or 1 for the arg-max ties (tau = 0, mcts.py:206-209)
    uint32_t wgt[3];
    const bool proportional = tau_one && total > 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) wgt[q] = proportional ? cnt[q] : ((lg[q] && cnt[q] == maxc) ? 1u : 0u);
    uint32_t wsum = wgt[0] + wgt[1] + wgt[2];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) wsum += __shfl_xor_sync(FULL, wsum, d);
    const uint32_t target = (uint32_t)(r % (uint64_t)wsum);
    // first action (in action order) whose inclusive prefix weight exceeds target
    int action = -1;
    uint32_t before = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        uint32_t incl = wgt[q];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += o;
        }
        uint32_t hit = __ballot_sync(FULL, wgt[q] > 0 && before + incl > target);
        if (action < 0 && hit) action = 32 * q + (__ffs(hit) - 1);
        before += __shfl_sync(FULL, incl, 31);
    }

    // trajectory record of this ply
    tc_example *rec = sp.pending + (size_t)env * 81 + e.n_records;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        int a = lane + 32 * q;
        if (a < 81) rec->counts[a] = (uint16_t)min(cnt[q], 65535u);
    }
    if (lane < 8) rec->state[lane] = e.state[lane];
    if (lane == 0) {
        rec->action = (int8_t)action;
        rec->tau_one = tau_one ? 1 : 0;
        rec->player = (int8_t)e.player;
        rec->z_sign = 0;
        reinterpret_cast<uint16_t *>(rec)[99] = 0;      // the two padding bytes at offset 198: records compare bytewise
        rec->env = env;
        rec->game = e.game;
        rec->ply = e.ply;
        rec->finished = 0;
    }
    e.n_records += 1;

    Pos nx = play(p, action);
    const int st = status_of(summarize(nx));
    if (lane == 0) {
        atomicAdd(sp.stats + 0, 1ull);
        atomicAdd(sp.stats + 1, (unsigned long long)cfg.num_sims);
    }
    if (st == 0) {
        uint32_t w[8];
        pack(nx, w);
#pragma unroll
        for (int k = 0; k < 8; ++k) e.state[k] = w[k];
        e.player = -e.player;
        e.ply += 1;
    } else {
        // st == 1: the side to move (-player) lost, i.e. `player` won.  st == 2: draw, scored +-1e-4 with the
        // sign convention of coach.py:143-149 (r is taken for the player to move after the last move).
        __syncwarp();
        tc_example *game_recs = sp.pending + (size_t)env * 81;
        const int n_rec = e.n_records;
        for (int i = lane; i < n_rec; i += 32) {
            int who = game_recs[i].player;
            int z = (st == 2) ? (who == -e.player ? 1 : -1) : (st == 1 ? (who == e.player ? 1 : -1)
                                                                        : (who == e.player ? -1 : 1));
            game_recs[i].z_sign = (int8_t)z;
            game_recs[i].finished = (st == 2) ? 2 : 1;
        }
        __syncwarp();
        int base = 0;
        const bool keep = !(cfg.flags & TC_SP_DISCARD_RECORDS);
        if (keep && lane == 0) base = atomicAdd(sp.finished_count, n_rec);
        base = __shfl_sync(FULL, base, 0);
        if (!keep) {
            // records are not wanted (warm-up / throughput runs)
        } else if (base + n_rec <= sp.finished_cap) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(game_recs);
            uint32_t *dst = reinterpret_cast<uint32_t *>(sp.finished + base);
            const int words = n_rec * (int)(sizeof(tc_example) / 4);
            for (int i = lane; i < words; i += 32) dst[i] = src[i];
        } else {
            // no room: the records are dropped and the shortfall reported (tc_selfplay_drain -> TC_ECAPACITY);
            // slots of a partially fitting game are marked invalid
            for (int i = lane; i < n_rec; i += 32)
                if (base + i < sp.finished_cap) { sp.finished[base + i].finished = 0; sp.finished[base + i].env = -1; }
            if (lane == 0) atomicAdd(sp.stats + 6, (unsigned long long)n_rec);
        }
        if (lane == 0) {
            atomicAdd(sp.stats + 2, 1ull);
            int winner = (st == 2) ? 0 : (st == 1 ? e.player : -e.player);
            atomicAdd(sp.stats + (winner == 1 ? 3 : (winner == -1 ? 4 : 5)), 1ull);
        }
        e.n_records = 0;
        if (cfg.auto_reset) {
#pragma unroll
            for (int k = 0; k < 8; ++k) e.state[k] = 0u;
            e.player = 1;
            e.ply = 0;
            e.game += 1;
            e.needs_reset = 1;
        } else {
            e.done = 1;
        }
    }
    if (lane == 0) sp.env[env] = e;
}

// clears the tree of every env whose game just restarted (VectorizedMCTS.reset(i), mcts.py:165-172)
__global__ void k_reset_marked(TreeStore ts, SelfPlayStore sp)
{
    const int tree = blockIdx.x;
    if (!sp.env[tree].needs_reset) return;
    uint32_t *table = ts.hash + (size_t)tree * ts.hashSize;
    for (uint32_t i = threadIdx.x; i < ts.hashSize; i += blockDim.x) table[i] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        TreeCtl *c = ts.ctl + tree;
        c->root_idx = -1;
        c->active = 0;
        c->n_nodes = 0;
        c->n_edges = 0;
        c->leaf_kind = LEAF_NONE;
        c->path_len = 0;
        sp.env[tree].needs_reset = 0;
    }
}

// Training examples from trajectory records (coach.py:136-149): one warp per output example.
// Image k = 2 i + flip of getSymmetries (utac_game.py:188-190): np.rot90(x, i) then, if flip, np.fliplr.
// rot90 once maps out[r][c] = in[c][8 - r]; the source cell of output cell (r, c) is found by undoing flip, then i turns.
__global__ void __launch_bounds__(256) k_examples_from_records(int n_out, const tc_example *__restrict__ records,
                                                               int n_records, const int32_t *__restrict__ index,
                                                               float *__restrict__ obs, float *__restrict__ pi,
                                                               float *__restrict__ z)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_out) return;
    const int ex = index ? index[row] : row;
    const int rec_i = ex >> 3, image = ex & 7;
    float *o = obs + 324 * (size_t)row, *po = pi + 81 * (size_t)row;
    if (rec_i < 0 || rec_i >= n_records) {          // out-of-range index: a zero example, never a wild read
        for (int a = lane; a < 324; a += 32) o[a] = 0.f;
        for (int a = lane; a < 81; a += 32) po[a] = 0.f;
        if (lane == 0) z[row] = 0.f;
        return;
    }
    const tc_example *rec = records + rec_i;
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = rec->state[k];
    Pos p = unpack(w);
    uint32_t lw[3];
    legal_words(p, summarize(p), lw);
    const int turns = image >> 1;
    const bool flip = (image & 1) != 0;
    uint32_t total = 0;
    for (int a = lane; a < 81; a += 32) total += rec->counts[a];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) total += __shfl_xor_sync(FULL, total, d);
    const bool proportional = rec->tau_one != 0 && total > 0;
    const int played = rec->action;
    for (int a = lane; a < 81; a += 32) {
        int r = a / 9, c = a - 9 * r;
        if (flip) c = 8 - c;
        for (int t = 0; t < turns; ++t) {           // out[r][c] = in[c][8 - r]
            const int nr = c, nc = 8 - r;
            r = nr;
            c = nc;
        }
        const int src = 9 * r + c;
        const bool mine = action_bit(p.mine, src), occ = action_bit(p.occ, src);
        o[a] = mine ? 1.f : 0.f;
        o[81 + a] = (occ && !mine) ? 1.f : 0.f;
        o[162 + a] = action_bit(lw, src) ? 1.f : 0.f;
        o[243 + a] = 1.f;
        // float(float64(count) / float64(total)): the reference divides in float64 (mcts.py:211-213) and converts to
        // float32 once (coach.py:138)
        po[a] = proportional ? (float)((double)rec->counts[src] / (double)total) : (src == played ? 1.f : 0.f);
    }
    if (lane == 0) z[row] = (float)((double)rec->z_sign * (rec->finished == 2 ? 1e-4 : 1.0));
}

void launch_examples_from_records(int n_out, const tc_example *records, int n_records, const int32_t *index, float *obs,
                                  float *pi, float *z, cudaStream_t st)
{
    if (n_out <= 0) return;
    k_examples_from_records<<<ceil_div(n_out, 8), 256, 0, st>>>(n_out, records, n_records, index, obs, pi, z);
}

void launch_selfplay_init(const SelfPlayStore &sp, cudaStream_t st)
{
    k_selfplay_init<<<ceil_div(sp.E, 256), 256, 0, st>>>(sp);
}

void launch_selfplay_roots(const SelfPlayStore &sp, cudaStream_t st)
{
    k_selfplay_roots<<<ceil_div(sp.E, 256), 256, 0, st>>>(sp);
}

void launch_selfplay_move(const TreeStore &ts, const SelfPlayStore &sp, tc_selfplay_config cfg, cudaStream_t st)
{
    k_selfplay_move<<<ceil_div(sp.E, 4), 128, 0, st>>>(ts, sp, cfg);
}

void launch_reset_marked(const TreeStore &ts, const SelfPlayStore &sp, cudaStream_t st)
{
    k_reset_marked<<<sp.E, 256, 0, st>>>(ts, sp);
}

}  // namespace tc
