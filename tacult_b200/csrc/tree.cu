// Warp-per-tree PUCT selection, leaf expansion and backup (see tree.cuh for layout and citations).
#include "tree.cuh"

#include "rules_warp.cuh"

namespace tc {

// ---------------------------------------------------------------------------------------------
// small warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    return h ^ (h >> 16);
}

// 32-bit arithmetic only (the table needs a slot index and an independent 12-bit tag, nothing cryptographic):
// low word -> slot, bits 40..51 -> tag
__device__ __forceinline__ uint64_t key_hash(const Pos &p)
{
    const uint32_t K = 0x9E3779B1u;
    uint32_t x = p.mine[0];
    x = x * K + p.mine[1];
    x = x * K + p.mine[2];
    x = x * K + p.occ[0];
    x = x * K + p.occ[1];
    x = x * K + p.occ[2];
    x = x * K + p.last;
    return ((uint64_t)fmix32(x ^ 0x68E31DA4u) << 32) | (uint64_t)fmix32(x);
}

__device__ __forceinline__ void unpack_hdr(const uint4 &h0, const uint4 &h1, Pos &p, uint32_t &n_edges,
                                           uint32_t &edge_off)
{
    p.mine[0] = h0.x; p.mine[1] = h0.y; p.mine[2] = h0.z;
    p.occ[0] = h0.w;  p.occ[1] = h1.x;  p.occ[2] = h1.y;
    p.last = h1.z & 0xFu;
    n_edges = (h1.z >> 8) & 0xFFu;
    edge_off = h1.w;
}

__device__ __forceinline__ int edge_index(const uint32_t am[3], int q, int lane)
{
    int j = __popc(am[q] & ((1u << lane) - 1u));
    if (q >= 1) j += __popc(am[0]);
    if (q >= 2) j += __popc(am[1]);
    return j;
}

// Looks the position up in the tree's table.  Returns the node index or -1; when absent `free_slot`
// is where it would be inserted (0xFFFFFFFF if the table is full).
__device__ int table_find(const TreeStore &ts, int tree, const Pos &p, uint64_t h, int lane, uint32_t &free_slot)
{
    const uint32_t mask = ts.hashSize - 1u;
    const uint32_t *table = ts.hash + (size_t)tree * ts.hashSize;
    const uint4 *hdr = ts.node_hdr + 2ull * (size_t)tree * ts.capN;
    const uint32_t tag = (uint32_t)(h >> 40) & 0xFFFu;
    const uint32_t start = (uint32_t)h & mask;
    for (uint32_t probe = 0; probe < ts.hashSize; probe += 32) {
        uint32_t slot = (start + probe + (uint32_t)lane) & mask;
        uint32_t e = table[slot];
        unsigned empties = __ballot_sync(FULL, e == 0u);
        unsigned cands = __ballot_sync(FULL, e != 0u && (e >> 20) == tag);
        int first_empty = empties ? (__ffs(empties) - 1) : 32;
        if (first_empty < 32) cands &= (1u << first_empty) - 1u;
        while (cands) {
            int src = __ffs(cands) - 1;
            cands &= cands - 1u;
            uint32_t ce = __shfl_sync(FULL, e, src);
            uint32_t idx = (ce & 0xFFFFFu) - 1u;
            uint4 h0 = hdr[2ull * idx], h1 = hdr[2ull * idx + 1];
            if (h0.x == p.mine[0] && h0.y == p.mine[1] && h0.z == p.mine[2] && h0.w == p.occ[0] &&
                h1.x == p.occ[1] && h1.y == p.occ[2] && (h1.z & 0xFu) == p.last)
                return (int)idx;
        }
        if (first_empty < 32) {
            free_slot = (start + probe + (uint32_t)first_empty) & mask;
            return -1;
        }
    }
    free_slot = 0xFFFFFFFFu;
    return -1;
}

// TC_TRACK_POS=1: the descent carries the position down the path (one `play` per level) instead of reading the parent's
// header back when it reaches the leaf.  Measured on B200: 0 = 35.8 us, 1 = 37.0 us per launch at 4096 trees (the extra
// registers spill), so the header read stays.
#ifndef TC_TRACK_POS
#define TC_TRACK_POS 0
#endif
// TC_PUCT_PREFILTER=1: decide the PUCT arg-max from float32 intervals when they separate the candidates, and only
// fall back to the float64 evaluation on (near) ties; see select_tree.
#ifndef TC_PUCT_PREFILTER
#define TC_PUCT_PREFILTER 1
#endif

// Creates a node for `p` (action mask am) and requests its evaluation.  Returns the node index, or -1 if a pool is
// exhausted (error flag raised, tree parked).  off/n receive the new node's edge block.  Everything the rest of the
// simulation needs (bump pointers, the leaf's row / edge block / legal mask, §8(d) counters) goes straight into the
// control block: nothing of it is carried in registers through the descent loop (the kernel is register- and issue-bound).
__device__ int create_leaf(const TreeStore &ts, int tree, TreeCtl *cp, const Pos &p, uint64_t h, uint32_t free_slot,
                           const uint32_t am[3], int lane, uint32_t &off, uint32_t &n)
{
    n = (uint32_t)(__popc(am[0]) + __popc(am[1]) + __popc(am[2]));
    const uint32_t n_nodes = cp->n_nodes, n_edges = cp->n_edges;
    if (n_nodes >= ts.capN || n_edges + n > ts.capE || free_slot == 0xFFFFFFFFu) {
        if (lane == 0) {
            atomicExch(ts.error_flag, 1);
            cp->active = 0;
        }
        return -1;
    }
    const uint32_t idx = n_nodes;
    off = n_edges;
    size_t g = (size_t)tree * ts.capN + idx;
    if (lane == 0) {
        ts.node_hdr[2 * g] = make_uint4(p.mine[0], p.mine[1], p.mine[2], p.occ[0]);
        ts.node_hdr[2 * g + 1] = make_uint4(p.occ[1], p.occ[2], p.last | (n << 8), off);
        ts.hash[(size_t)tree * ts.hashSize + free_slot] = (((uint32_t)(h >> 40) & 0xFFFu) << 20) | (idx + 1u);
    }
    Edge *blk = tree_edges(ts, tree) + off;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (am[q] == 0u) continue;                     // warp-uniform: most leaves have moves in one or two thirds only
        if ((am[q] >> lane) & 1u) {
            uint4 *rec = reinterpret_cast<uint4 *>(blk + edge_index(am, q, lane));
            rec[0] = edge_pack(0.0, 0u, CHILD_UNRESOLVED);                       // Q, N, child
            rec[1] = edge_pack(0.0, 0u, (uint32_t)(lane + 32 * q));              // P (set by the expansion), coff, meta
        }
    }
    if (lane == 0) {
        const int row = atomicAdd(ts.eval_count, 1);
        // the packed position of the leaf: one 32-byte row, two vector stores
        uint4 *dst = reinterpret_cast<uint4 *>(ts.eval_state + 8 * (size_t)row);
        dst[0] = make_uint4(p.mine[0], p.mine[1], p.mine[2], p.occ[0]);
        dst[1] = make_uint4(p.occ[1], p.occ[2], p.last, 0u);
        cp->n_nodes = n_nodes + 1u;
        cp->n_edges = n_edges + n;
        cp->leaf_kind = LEAF_NN;
        cp->leaf_node = (int)idx;
        cp->leaf_row = row;
        cp->leaf_off = off;
        cp->leaf_n = n;
        cp->leaf_am[0] = am[0]; cp->leaf_am[1] = am[1]; cp->leaf_am[2] = am[2];
        atomicAdd(&cp->nn_leaves, 1ull);               // reductions without a return value: one RED each
        atomicAdd(&cp->sum_leaf_legal, (unsigned long long)n);
    }
    return (int)idx;
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_pack_fields(int n, const uint16_t *__restrict__ board, const uint16_t *__restrict__ occ,
                              const int8_t *__restrict__ last, uint32_t *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[8];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        w[g] = 0; w[3 + g] = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            w[g] |= ((uint32_t)board[9 * (size_t)i + 3 * g + j] & F9) << (9 * j);
            w[3 + g] |= ((uint32_t)occ[9 * (size_t)i + 3 * g + j] & F9) << (9 * j);
        }
    }
    w[6] = (uint32_t)((int)last[i] + 1) & 0xFu;
    w[7] = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) out[8 * (size_t)i + k] = w[k];
}

__global__ void k_reset(TreeStore ts, int env, const int32_t *__restrict__ list)
{
    // one block per tree: clear the hash table and the control block
    int tree = list ? list[blockIdx.x] : (env >= 0 ? env : (int)blockIdx.x);
    if (tree < 0 || tree >= ts.E) return;
    uint32_t *table = ts.hash + (size_t)tree * ts.hashSize;
    for (uint32_t i = threadIdx.x; i < ts.hashSize; i += blockDim.x) table[i] = 0u;
    if (threadIdx.x == 0) {
        TreeCtl *c = ts.ctl + tree;          // the §8(d) counters survive a reset (tc_reset_stats clears them)
        memset(c->root_state, 0, sizeof c->root_state);
        c->root_idx = -1;
        c->active = 0;
        c->n_nodes = 0;
        c->n_edges = 0;
        c->leaf_kind = LEAF_NONE;
        c->leaf_node = -1;
        c->leaf_row = 0;
        c->path_len = 0;
        c->leaf_value = 0.0;
        c->root_off = 0;
        c->root_n = 0;
    }
}

__global__ void k_set_roots(TreeStore ts, const uint32_t *__restrict__ states, const uint8_t *__restrict__ active)
{
    int tree = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (tree >= ts.E) return;
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = states[8 * (size_t)tree + k];
    Pos p = unpack(w);
    int status = status_of(summarize(p));
    bool on = (active == nullptr || active[tree] != 0) && status == 0;
    uint32_t free_slot;
    int idx = -1;
    uint32_t off = 0, n = 0;
    if (on) {
        idx = table_find(ts, tree, p, key_hash(p), lane, free_slot);
        if (idx >= 0) {
            const uint4 *hdr = ts.node_hdr + 2ull * ((size_t)tree * ts.capN + idx);
            Pos q;
            unpack_hdr(hdr[0], hdr[1], q, n, off);
        }
    }
    if (lane == 0) {
        TreeCtl *c = ts.ctl + tree;
        uint32_t packed[8];
        pack(p, packed);
#pragma unroll
        for (int k = 0; k < 8; ++k) c->root_state[k] = packed[k];
        c->root_idx = idx;
        c->root_off = off;
        c->root_n = n;
        c->active = on ? 1 : 0;
        c->leaf_kind = LEAF_NONE;
        c->path_len = 0;
    }
}

// One ply for every env on the device (tc_advance): the stored root + the chosen action -> canonical successor, which is
// looked up in the tree (statistics persist across plies, coach.py:162) and becomes the next root.  A finished game
// either parks the env or, with `restart`, gets a fresh tree and the initial position (mcts.py:165-172 + getInitBoard).
__global__ void k_advance(TreeStore ts, const int8_t *__restrict__ actions, int restart, uint32_t *__restrict__ states_out,
                          int8_t *__restrict__ status_out, int *__restrict__ illegal)
{
    const int tree = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tree >= ts.E) return;
    TreeCtl *c = ts.ctl + tree;
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = c->root_state[k];
    Pos p = unpack(w);
    const int a = actions[tree];
    int status = status_of(summarize(p));
    if (a >= 0) {
        uint32_t lw[3];
        legal_words(p, summarize(p), lw);
        if (a > 80 || !action_bit(lw, a)) {
            if (lane == 0) atomicAdd(illegal, 1);
        } else {
            p = play(p, a);
            status = status_of(summarize(p));
        }
    }
    if (lane == 0) {
        pack(p, w);
        if (states_out) {
            uint4 *dst = reinterpret_cast<uint4 *>(states_out + 8 * (size_t)tree);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        if (status_out) status_out[tree] = (int8_t)status;
    }
    const bool fresh = status != 0 && restart != 0;
    if (fresh) {                                   // new game: empty tree, initial position
        uint32_t *table = ts.hash + (size_t)tree * ts.hashSize;
        for (uint32_t i = lane; i < ts.hashSize; i += 32) table[i] = 0u;
        p.mine[0] = p.mine[1] = p.mine[2] = p.occ[0] = p.occ[1] = p.occ[2] = 0u;
        p.last = 0u;
    }
    const bool on = fresh || status == 0;
    uint32_t free_slot, off = 0, n = 0;
    int idx = -1;
    if (on && !fresh) {
        idx = table_find(ts, tree, p, key_hash(p), lane, free_slot);
        if (idx >= 0) {
            const uint4 *hdr = ts.node_hdr + 2ull * ((size_t)tree * ts.capN + idx);
            Pos q;
            unpack_hdr(hdr[0], hdr[1], q, n, off);
        }
    }
    if (lane == 0) {
        pack(p, w);
#pragma unroll
        for (int k = 0; k < 8; ++k) c->root_state[k] = w[k];
        c->root_idx = idx;
        c->root_off = off;
        c->root_n = n;
        c->active = on ? 1 : 0;
        c->leaf_kind = LEAF_NONE;
        c->path_len = 0;
        if (fresh) {
            c->n_nodes = 0;
            c->n_edges = 0;
        }
    }
}

// Order-preserving map of a double to an unsigned key (-0.0 folded into +0.0 first, so equal values give equal keys
// exactly as the reference's strict '>' sees them).
__device__ __forceinline__ unsigned long long double_key(double u)
{
    long long bits = __double_as_longlong(__dadd_rn(u, 0.0));
    return (unsigned long long)bits ^ (bits < 0 ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
}

// Warp arg-max in three REDUX steps instead of a 5-round shuffle butterfly: max of the high words, max of the low words
// among the lanes that hold that high word, then the lowest edge index among the lanes that hold the maximum (edges are in
// action order: "lowest action index wins ties", mcts.py:314).  Lanes without a candidate pass j = 1 << 20.
__device__ __forceinline__ int warp_argmax(double best_u, int best_j)
{
    unsigned long long key = 0ull;                 // lanes without a candidate can never win
    if (best_j < (1 << 20)) key = double_key(best_u);
    const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
    const uint32_t mhi = __reduce_max_sync(FULL, hi);
    const uint32_t mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
    const bool top = best_j < (1 << 20) && hi == mhi && lo == mlo;
    return (int)__reduce_min_sync(FULL, top ? (uint32_t)best_j : (uint32_t)(1 << 20));
}

// PUCT score of one edge in float64, operation by operation as mcts.py:309-312 (no FMA contraction)
__device__ __forceinline__ double puct_exact(double cpuct, double P, double Q, uint32_t N, double sq_visited, double sq_fresh)
{
    const double cpp = __dmul_rn(cpuct, P);
    return N ? __dadd_rn(Q, __ddiv_rn(__dmul_rn(cpp, sq_visited), (double)(1u + N))) : __dmul_rn(cpp, sq_fresh);
}

// Selection at a node wider than a warp (free moves: up to 81 edges), three edges per lane, float64 throughout.
__device__ __noinline__ int select_wide(const Edge *blk, uint32_t n, int lane, double cpuct)
{
    double P[3], Q[3];
    uint32_t N[3], nsum = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const uint32_t j = (uint32_t)lane + 32u * q;
        N[q] = 0;
        P[q] = Q[q] = 0.0;
        if (j < n) {
            const uint4 st = edge_load_stats(blk + j), lk = edge_load_link(blk + j);
            Q[q] = u2d(st.x, st.y);
            N[q] = st.z;
            P[q] = u2d(lk.x, lk.y);
            nsum += N[q];
        }
    }
    nsum = __reduce_add_sync(FULL, nsum);
    const double sq_visited = nsum ? __dsqrt_rn((double)nsum) : 0.0;
    const double sq_fresh = __dsqrt_rn(__dadd_rn((double)nsum, 1e-8));
    double best_u = -INFINITY;
    int best_j = 1 << 20;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const uint32_t j = (uint32_t)lane + 32u * q;
        if (j < n) {
            const double u = puct_exact(cpuct, P[q], Q[q], N[q], sq_visited, sq_fresh);
            if (u > best_u) { best_u = u; best_j = (int)j; }       // ascending j: first maximum wins; NaN never does
        }
    }
    return warp_argmax(best_u, best_j);
}

// One simulation's descent for every tree: mcts.py:243-325 (selection part of `_search`).
// Every level costs one dependent memory round trip: the lanes load the whole edge block (priors, values,
// counts and where each child lives), Ns is the warp sum of the counts, the arg-max lane broadcasts the child.
__device__ __forceinline__ void select_tree(const TreeStore &ts, const int tree, const int lane, const double cpuct)
{
    TreeCtl *cp = ts.ctl + tree;
    if (!cp->active) {
        if (lane == 0) { cp->leaf_kind = LEAF_NONE; cp->path_len = 0; }
        return;
    }
    // the simulation's outcome is written to the control block where it is decided (terminal leaf, new leaf, exhausted
    // pool); until then "no leaf"
    if (lane == 0) cp->leaf_kind = LEAF_NONE;
    const int root_idx = cp->root_idx;
    uint32_t depth = 0, sum_legal = 0;
    uint2 *path = ts.path + (size_t)tree * MAX_DEPTH;
    const uint4 *hdr = ts.node_hdr + 2ull * (size_t)tree * ts.capN;

    Pos p;
    if (TC_TRACK_POS || root_idx < 0) {
        uint32_t rs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) rs[k] = cp->root_state[k];
        p = unpack(rs);
    }
    if (root_idx < 0) {
        // root never expanded: it is this simulation's leaf (mcts.py:268-290 at recursion depth 0)
        uint32_t am[3];
        int status;
        warp_legal_mask(p, lane, am, status);
        uint64_t h = key_hash(p);
        uint32_t free_slot;
        int idx = table_find(ts, tree, p, h, lane, free_slot);
        uint32_t off = 0, n = 0;
        if (idx < 0) idx = create_leaf(ts, tree, cp, p, h, free_slot, am, lane, off, n);
        else {
            Pos q;
            unpack_hdr(hdr[2ull * idx], hdr[2ull * idx + 1], q, n, off);
        }
        if (lane == 0) {
            cp->root_idx = idx;
            cp->root_off = off;
            cp->root_n = n;
        }
    } else {
        uint32_t off = cp->root_off, n = cp->root_n;
#if !TC_TRACK_POS
        int cur = root_idx;
#endif
        Edge *const edges = tree_edges(ts, tree);
        while (true) {
            Edge *const blk = edges + off;
            int best_j;
            uint32_t child, coff, meta;
            if (n <= 32u) {
                // ---- the common node: one edge per lane, two 16-byte loads each ----
                const bool valid = (uint32_t)lane < n;
                uint4 st = make_uint4(0u, 0u, 0u, 0u), lk = make_uint4(0u, 0u, 0u, 0u);
                if (valid) {
                    st = edge_load_stats(blk + lane);
                    lk = edge_load_link(blk + lane);
                }
                const uint32_t N = st.z;
                const uint32_t nsum = __reduce_add_sync(FULL, N);                          // Ns[s]
                const double P = u2d(lk.x, lk.y), Q = u2d(st.x, st.y);
                best_j = 1 << 20;
                bool decided = false;
#if TC_PUCT_PREFILTER
                // PUCT over the legal edges (mcts.py:301-316).  The arg-max only needs the float64 values when two
                // candidates are close.  A float32 evaluation of the same formula is within 1e-6 * (|Q| + |U|) of the exact
                // value (conversions, approximate sqrt and division, one rounding per operation); with a ten-fold margin
                // every edge gets an interval [u - eps, u + eps] that contains its exact score.  If only ONE edge's upper
                // end reaches the best lower end, that edge is the unique exact maximum — no tie can exist, so no tie-break
                // is needed — and the ~70 float64 instructions per level (two software square roots and a division, the
                // kernel's issue-slot hogs) are skipped.  Otherwise (exact ties of equal priors, near ties) the exact
                // path below decides, as it does for nodes wider than a warp.
                {
                    // sqrt(Ns) for visited edges; sqrt(Ns + 1e-8) for fresh ones is the same float32 once Ns >= 1 (1e-8 is
                    // below half an ulp) and 1e-4 at Ns = 0; sqrt.approx is within 2^-23
                    float s0;
                    asm("sqrt.approx.f32 %0, %1;" : "=f"(s0) : "f"((float)nsum));    // (float)nsum is exact below 2^24
                    const float s_fresh = nsum ? s0 : 1e-4f;
                    float u = 0.f, mag = 0.f;
                    if (valid) {
                        const float cpf = (float)cpuct * (float)P;
                        if (N) {
                            const float qf = (float)Q;
                            const float uf = __fdividef(cpf * s0, (float)(1u + N));
                            u = qf + uf;
                            mag = fabsf(qf) + fabsf(uf);
                        } else {
                            u = cpf * s_fresh;
                            mag = fabsf(u);
                        }
                    }
                    const float eps = 1e-5f * mag + 1e-30f;
                    auto okey = [](float x) {                                        // order-preserving float -> uint32
                        const uint32_t bits = __float_as_uint(x);
                        return bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
                    };
                    const bool finite = valid && fabsf(u) < 1e30f;                   // also false for NaN
                    const uint32_t klo = finite ? okey(u - eps) : 0u;
                    const uint32_t best_lo = __reduce_max_sync(FULL, klo);
                    const unsigned any_bad = __ballot_sync(FULL, valid && !finite);
                    const unsigned cand = __ballot_sync(FULL, finite && okey(u + eps) >= best_lo);
                    if (!any_bad && __popc(cand) == 1) {
                        best_j = __ffs(cand) - 1;
                        decided = true;
                    }
                }
#endif
                if (!decided) {
                    const double sq_visited = nsum ? __dsqrt_rn((double)nsum) : 0.0;      // only used by edges with N > 0
                    const double sq_fresh = __dsqrt_rn(__dadd_rn((double)nsum, 1e-8));
                    double best_u = -INFINITY;
                    if (valid) {
                        best_u = puct_exact(cpuct, P, Q, N, sq_visited, sq_fresh);
                        best_j = lane;
                        if (!(best_u > -INFINITY)) { best_u = -INFINITY; best_j = 1 << 20; }   // NaN / -inf is never selected (mcts.py:314)
                    }
                    best_j = warp_argmax(best_u, best_j);
                }
                // the winning lane broadcasts where the child lives
                const int src = best_j & 31;
                child = __shfl_sync(FULL, st.w, src);
                coff = __shfl_sync(FULL, lk.z, src);
                meta = __shfl_sync(FULL, lk.w, src);
            } else {
                // ---- wide node (free move): rare; the winner's record is read back with a uniform load ----
                best_j = select_wide(blk, n, lane, cpuct);
                child = coff = meta = 0u;
                if (best_j < (1 << 20)) {
                    const uint4 st = edge_load_stats(blk + best_j), lk = edge_load_link(blk + best_j);
                    child = st.w;
                    coff = lk.z;
                    meta = lk.w;
                }
            }
            sum_legal += n;
            if (best_j >= (1 << 20)) {       // cannot happen for a live position; park the tree loudly
                if (lane == 0) {
                    atomicExch(ts.error_flag, 2);
                    cp->active = 0;
                }
                depth = 0;
                break;
            }
            if (lane == 0) path[depth] = make_uint2(off, (uint32_t)best_j | (n << 8));
            ++depth;

            if (child >= CHILD_TERM_BASE && child != CHILD_UNRESOLVED) {
                if (lane == 0) {
                    const int st = (int)(child - CHILD_TERM_BASE);
                    cp->leaf_kind = LEAF_TERMINAL;
                    cp->leaf_value = st == 2 ? 1e-4 : (st == 3 ? 1.0 : -1.0);   // utac_game.py:64-81
                    atomicAdd(&cp->terminal_leaves, 1ull);
                }
                break;
            }
            const int action = (int)(meta & 0xFFu);
            if (child != CHILD_UNRESOLVED) {
#if TC_TRACK_POS
                p = play(p, action);
#else
                cur = (int)child;
#endif
                off = coff;
                n = (meta >> 8) & 0xFFu;
                continue;
            }

            // first traversal of this edge: successor position (mcts.py:319-320)
#if !TC_TRACK_POS
            {
                uint32_t pn, poff;
                unpack_hdr(hdr[2ull * cur], hdr[2ull * cur + 1], p, pn, poff);
            }
#endif
            Pos nx = play(p, action);
            uint32_t nam[3];
            int nst;
            warp_legal_mask(nx, lane, nam, nst);
            if (nst != 0) {
                if (lane == 0) {
                    blk[best_j].child = CHILD_TERM_BASE + (uint32_t)nst;
                    cp->leaf_kind = LEAF_TERMINAL;
                    cp->leaf_value = nst == 2 ? 1e-4 : (nst == 3 ? 1.0 : -1.0);
                    atomicAdd(&cp->terminal_leaves, 1ull);
                }
                break;
            }
            uint64_t h = key_hash(nx);
            uint32_t free_slot;
            int idx = table_find(ts, tree, nx, h, lane, free_slot);
            if (idx >= 0) {             // transposition: the position is already a node (mcts.py:268 is false)
                Pos q;
                uint32_t cn, co;
                unpack_hdr(hdr[2ull * idx], hdr[2ull * idx + 1], q, cn, co);
                if (lane == 0) {
                    blk[best_j].child = (uint32_t)idx;
                    blk[best_j].coff = co;
                    blk[best_j].meta = (uint32_t)action | (cn << 8);
                    atomicAdd(&cp->transpositions, 1ull);
                }
#if TC_TRACK_POS
                p = nx;
#else
                cur = idx;
#endif
                off = co;
                n = cn;
                continue;
            }
            uint32_t co = 0, cn = 0;
            idx = create_leaf(ts, tree, cp, nx, h, free_slot, nam, lane, co, cn);
            if (idx >= 0) {
                if (lane == 0) {
                    blk[best_j].child = (uint32_t)idx;
                    blk[best_j].coff = co;
                    blk[best_j].meta = (uint32_t)action | (cn << 8);
                }
            } else {
                depth = 0;              // pool exhausted: drop this simulation (leaf_kind stays LEAF_NONE)
            }
            break;
        }
    }
    if (lane == 0) {
        cp->path_len = (int)depth;
        // §8(d) counters: reductions without a return value (one RED each, no load on the warp's critical path)
        atomicAdd(&cp->sims, 1ull);
        atomicAdd(&cp->sum_depth, (unsigned long long)depth);
        atomicAdd(&cp->sum_legal, (unsigned long long)sum_legal);
    }
}

// numpy's float64 np.sum over 81 contiguous values (8 running sums, pairwise tree, tail), mcts.py:273
__device__ __forceinline__ double numpy_sum81(const double x[3], int lane)
{
    const int j8 = lane & 7;
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        // element a = 8k + j8 (< 80) lives in x[a / 32] of lane a % 32; a / 32 == k / 4 because j8 < 8
        double v = __shfl_sync(FULL, x[k >> 2], 8 * (k & 3) + j8);
        r = (k == 0) ? v : __dadd_rn(r, v);
    }
    double r0 = __shfl_sync(FULL, r, 0), r1 = __shfl_sync(FULL, r, 1), r2 = __shfl_sync(FULL, r, 2),
           r3 = __shfl_sync(FULL, r, 3), r4 = __shfl_sync(FULL, r, 4), r5 = __shfl_sync(FULL, r, 5),
           r6 = __shfl_sync(FULL, r, 6), r7 = __shfl_sync(FULL, r, 7);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    double tail = __shfl_sync(FULL, x[2], 16);      // element 80
    return __dadd_rn(res, tail);
}

// Expansion of the evaluated leaf (mcts.py:268-290) and backup along the path (mcts.py:327-341).
__device__ __forceinline__ void expand_backup_tree(const TreeStore &ts, const int tree, const int lane)
{
    const TreeCtl *cp = ts.ctl + tree;
    int kind = cp->leaf_kind;
    if (kind == LEAF_NONE) return;
    int path_len = cp->path_len;
    double x;
    if (kind == LEAF_NN) {
        const int row = cp->leaf_row;
        const uint32_t n = cp->leaf_n, off = cp->leaf_off;
        const uint32_t am[3] = {cp->leaf_am[0], cp->leaf_am[1], cp->leaf_am[2]};
        const float *pi = ts.eval_pi + 81 * (size_t)row;
        double m[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            int a = lane + 32 * q;
            m[q] = 0.0;
            if (a < 81 && ((am[q] >> lane) & 1u)) m[q] = (double)pi[a];
        }
        // production mode: the policy head already masked and renormalised (float32); parity mode: mcts.py:272-275
        const double total = ts.priors_masked ? 1.0 : numpy_sum81(m, lane);
        Edge *blk = tree_edges(ts, tree) + off;
        const bool ok = total > 0.0;
        const double uniform = __ddiv_rn(1.0, (double)n);     // (0 + valids) / np.sum(valids), mcts.py:282-283
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (am[q] == 0u) continue;                        // warp-uniform: skips the division where no lane has a move
            if ((am[q] >> lane) & 1u) {
                int j = edge_index(am, q, lane);
                blk[j].P = ts.priors_masked ? m[q] : (ok ? __ddiv_rn(m[q], total) : uniform);
            }
        }
        x = (double)ts.eval_v[row];
    } else {
        x = cp->leaf_value;
    }
    // edge `dist` levels above the leaf receives 0.0 + -(value one level below)  (mcts.py:325,342)
    const uint2 *path = ts.path + (size_t)tree * MAX_DEPTH;
    for (int e = lane; e < path_len; e += 32) {
        uint2 pe = path[e];
        int dist = path_len - e;
        double v = (dist & 1) ? -x : x;
        v = __dadd_rn(0.0, v);                                  // -0.0 normalises to +0.0 as in the reference
        const uint32_t j = pe.y & 0xFFu;
        uint4 *rec = reinterpret_cast<uint4 *>(tree_edges(ts, tree) + pe.x + j);      // the {Q, N, child} half
        const uint4 st = *rec;
        const uint32_t cnt = st.z;
        double q = v;
        if (cnt) q = __ddiv_rn(__dadd_rn(__dmul_rn((double)cnt, u2d(st.x, st.y)), v), (double)(cnt + 1u));
        *rec = edge_pack(q, cnt + 1u, st.w);
    }
}

__global__ void __launch_bounds__(128, 7) k_select(TreeStore ts, double cpuct)
{
    griddep_launch(ts.pdl_trigger == 1 || ts.pdl_trigger == 3);
    griddep_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0 && ts.eval_count_next) *ts.eval_count_next = 0;
    const int tree = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tree >= ts.E) return;
    select_tree(ts, tree, threadIdx.x & 31, cpuct);
}

__global__ void __launch_bounds__(128) k_expand_backup(TreeStore ts)
{
    griddep_wait();
    const int tree = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tree >= ts.E) return;
    expand_backup_tree(ts, tree, threadIdx.x & 31);
}

// Backup of simulation k and descent of simulation k+1 in one launch: the warp that owns the tree finishes the
// expansion/backup, then walks down again while the path's nodes are still in cache.
__global__ void __launch_bounds__(128, 7) k_backup_select(TreeStore ts, double cpuct)
{
    griddep_launch(ts.pdl_trigger == 1 || ts.pdl_trigger == 3);
    griddep_wait();
    // the other counter of the pair was last read by the evaluator kernels of the previous simulation, which have
    // completed; the next simulation's descent starts counting from zero without a memset node in the stream
    if (blockIdx.x == 0 && threadIdx.x == 0 && ts.eval_count_next) *ts.eval_count_next = 0;
    const int tree = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tree >= ts.E) return;
    const int lane = threadIdx.x & 31;
    expand_backup_tree(ts, tree, lane);
    __syncwarp();          // the lanes' writes to the edge records are ordered before the descent's reads
    select_tree(ts, tree, lane, cpuct);
}

// Deterministic evaluator, specification in oracle/evaluators.py (HashEvaluator)
__global__ void k_hash_eval(TreeStore ts, HashEvalParams hp)
{
    griddep_wait();
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= *ts.eval_count) return;
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = ts.eval_state[8 * (size_t)row + k];
    Pos p = unpack(w);
    uint32_t am[3];
    int status;
    warp_legal_mask(p, lane, am, status);
    uint64_t h = mix64(hp.salt);
    h = mix64(h ^ p.mine[0]); h = mix64(h ^ p.mine[1]); h = mix64(h ^ p.mine[2]);
    h = mix64(h ^ p.occ[0]);  h = mix64(h ^ p.occ[1]);  h = mix64(h ^ p.occ[2]);
    h = mix64(h ^ am[0]);     h = mix64(h ^ am[1]);     h = mix64(h ^ am[2]);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        int a = lane + 32 * q;
        if (a < 81) {
            uint64_t r = mix64(h ^ ((uint64_t)(a + 1) * 0x9E3779B97F4A7C15ull));
            float v;
            if (hp.pi_levels == 0) v = (float)((uint32_t)(r >> 40) + 1u) * 5.9604644775390625e-08f;   // 2^-24
            else v = (float)(uint32_t)(r % (uint64_t)hp.pi_levels) * 0.125f;
            ts.eval_pi[81 * (size_t)row + a] = v;
        }
    }
    if (lane == 0) {
        uint64_t r = mix64(h ^ 0xA5A5A5A5DEADBEEFull);
        float v;
        if (hp.v_levels == 0) v = (float)((int)(uint32_t)(r >> 40) - (1 << 23)) * 1.1920928955078125e-07f;  // 2^-23
        else v = (float)((int)(r % (uint64_t)(2 * hp.v_levels + 1)) - hp.v_levels) * 0.25f;
        ts.eval_v[row] = v;
    }
}

__global__ void k_root_counts(TreeStore ts, uint32_t *__restrict__ counts, double *__restrict__ qout)
{
    int tree = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (tree >= ts.E) return;
    const TreeCtl *c = ts.ctl + tree;
    int root = c->root_idx;
    bool have = c->active && root >= 0;
    Pos p;
    uint32_t n = 0, off = 0;
    uint32_t am[3] = {0, 0, 0};
    if (have) {
        const uint4 *hdr = ts.node_hdr + 2ull * ((size_t)tree * ts.capN + root);
        unpack_hdr(hdr[0], hdr[1], p, n, off);
        int status;
        warp_legal_mask(p, lane, am, status);
    }
    const Edge *blk = tree_edges(ts, tree) + off;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        int a = lane + 32 * q;
        if (a < 81) {
            uint32_t cnt = 0;
            double qv = nan("");
            if (have && ((am[q] >> lane) & 1u)) {
                int j = edge_index(am, q, lane);
                cnt = blk[j].N;
                if (cnt) qv = blk[j].Q;
            }
            if (counts) counts[81 * (size_t)tree + a] = cnt;
            if (qout) qout[81 * (size_t)tree + a] = qv;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline int warp_grid(int n_warps, int threads) { return ceil_div(n_warps, threads / 32); }

void launch_pack_fields(int n, const uint16_t *board, const uint16_t *occ, const int8_t *last, uint32_t *out,
                        cudaStream_t st)
{
    k_pack_fields<<<ceil_div(n, 256), 256, 0, st>>>(n, board, occ, last, out);
}

void launch_reset(const TreeStore &ts, int env, cudaStream_t st)
{
    k_reset<<<env >= 0 ? 1 : ts.E, 256, 0, st>>>(ts, env, nullptr);
}

void launch_reset_list(const TreeStore &ts, int n, const int32_t *envs_dev, cudaStream_t st)
{
    if (n > 0) k_reset<<<n, 256, 0, st>>>(ts, -1, envs_dev);
}

void launch_set_roots(const TreeStore &ts, const uint32_t *states_dev, const uint8_t *active_dev, cudaStream_t st)
{
    k_set_roots<<<warp_grid(ts.E, 128), 128, 0, st>>>(ts, states_dev, active_dev);
}

void launch_advance(const TreeStore &ts, const int8_t *actions_dev, int restart, uint32_t *states_dev, int8_t *status_dev,
                    int *illegal_dev, cudaStream_t st)
{
    k_advance<<<warp_grid(ts.E, 128), 128, 0, st>>>(ts, actions_dev, restart, states_dev, status_dev, illegal_dev);
}

void launch_select(const TreeStore &ts, double cpuct, cudaStream_t st)
{
    prefer_max_shared(k_select);
    launch_chain(k_select, warp_grid(ts.E, 128), 128, 0, st, ts, cpuct);
}

void launch_backup_select(const TreeStore &ts, double cpuct, cudaStream_t st)
{
    prefer_max_shared(k_backup_select);
    launch_chain(k_backup_select, warp_grid(ts.E, 128), 128, 0, st, ts, cpuct);
}

void launch_expand_backup(const TreeStore &ts, cudaStream_t st)
{
    prefer_max_shared(k_expand_backup);
    launch_chain(k_expand_backup, warp_grid(ts.E, 128), 128, 0, st, ts);
}

void launch_hash_eval(const TreeStore &ts, HashEvalParams hp, cudaStream_t st)
{
    launch_chain(k_hash_eval, warp_grid(ts.E, 256), 256, 0, st, ts, hp);
}

void launch_root_counts(const TreeStore &ts, uint32_t *counts_dev, double *q_dev, cudaStream_t st)
{
    k_root_counts<<<warp_grid(ts.E, 128), 128, 0, st>>>(ts, counts_dev, q_dev);
}

}  // namespace tc
