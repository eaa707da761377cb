// C ABI of the search engine and evaluator (include/tacult_b200.h).
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "nn.cuh"
#include "nn_bf16.cuh"
#include "selfplay.cuh"
#include "tree.cuh"

namespace tc {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

__global__ void k_log_append(TreeStore ts, uint32_t *log_states, float *log_pi, float *log_v, int *log_count,
                             int capacity)
{
    const int rows = *ts.eval_count;
    const int base = *log_count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * 81; i += gridDim.x * blockDim.x) {
        int row = i / 81, a = i % 81;
        if (base + row >= capacity) continue;
        log_pi[(size_t)(base + row) * 81 + a] = ts.eval_pi[(size_t)row * 81 + a];
        if (a < 8) log_states[(size_t)(base + row) * 8 + a] = ts.eval_state[(size_t)row * 8 + a];
        if (a == 8) log_v[base + row] = ts.eval_v[row];
    }
}

__global__ void k_log_advance(TreeStore ts, int *log_count) { *log_count += *ts.eval_count; }

__global__ void k_tree_totals(TreeStore ts, unsigned long long *out)
{
    // out[0..6] counters, [7] nodes, [8] edges, [9] max nodes, [10] max edges
    unsigned long long acc[11];
    for (int k = 0; k < 11; ++k) acc[k] = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ts.E; t += gridDim.x * blockDim.x) {
        const TreeCtl &c = ts.ctl[t];
        acc[0] += c.sims; acc[1] += c.sum_depth; acc[2] += c.sum_legal; acc[3] += c.nn_leaves;
        acc[4] += c.sum_leaf_legal; acc[5] += c.terminal_leaves; acc[6] += c.transpositions;
        acc[7] += c.n_nodes; acc[8] += c.n_edges;
        acc[9] = max(acc[9], (unsigned long long)c.n_nodes);
        acc[10] = max(acc[10], (unsigned long long)c.n_edges);
    }
    for (int k = 0; k < 9; ++k) atomicAdd(out + k, acc[k]);
    atomicMax(out + 9, acc[9]);
    atomicMax(out + 10, acc[10]);
}

__global__ void k_zero_counters(TreeStore ts)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ts.E) return;
    TreeCtl &c = ts.ctl[t];
    c.sims = c.sum_depth = c.sum_legal = c.nn_leaves = c.sum_leaf_legal = c.terminal_leaves = c.transpositions = 0;
}

}  // namespace tc

using namespace tc;

struct tc_engine {
    tc_config cfg{};
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    TreeStore ts{};
    DevBuf<uint4> node_hdr;
    DevBuf<uint2> path;
    DevBuf<uint32_t> hash, eval_state, root_states, counts;
    DevBuf<unsigned long long> edge_mem, totals;
    DevBuf<TreeCtl> ctl;
    DevBuf<float> eval_pi, eval_v;
    DevBuf<int> eval_count, error_flag, log_count;
    DevBuf<uint8_t> root_active;
    DevBuf<uint16_t> fld_board, fld_occ;
    DevBuf<int8_t> fld_last;
    DevBuf<double> root_q;
    DevBuf<int32_t> reset_list;
    DevBuf<int8_t> adv_actions, adv_status;
    DevBuf<int> adv_illegal;
    // evaluation log
    DevBuf<uint32_t> log_states;
    DevBuf<float> log_pi, log_v;
    // evaluator
    int evaluator = TC_EVAL_HASH;
    HashEvalParams hash_params{};
    FoldedNet folded;
    NetF32 net32, net32x[3];       // one evaluator instance per part (scratch buffers are per instance)
    NetBF16 net16, net16x[3];
    bool have_weights = false;
    // The trees are searched as two independent halves on two streams, so that one half's tree kernels overlap
    // the other half's network kernels (different SM resources: no shared memory vs. all of it).
    int n_halves = 1;              // number of parts (1, 2 or 4), each on its own stream
    TreeStore hts[4]{};
    cudaStream_t side_stream[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    // scratch for tc_forward
    DevBuf<float> fwd_obs, fwd_pi, fwd_v, fwd_logits;
    DevBuf<uint32_t> fwd_states;
    int fwd_capacity = 0;
    // self-play
    SelfPlayStore sp{};
    tc_selfplay_config sp_cfg{};
    bool sp_ready = false;
    DevBuf<EnvState> sp_env;
    DevBuf<tc_example> sp_pending, sp_finished;
    DevBuf<int> sp_finished_count;
    DevBuf<unsigned long long> sp_stats;
    unsigned long long launches = 0;
    Profiler prof;
};

static int check_engine(tc_engine *e)
{
    TC_CHECK(e != nullptr, TC_EINVAL, "null engine handle");
    TC_CUDA(cudaSetDevice(e->cfg.device));
    return TC_OK;
}

static int check_device_errors(tc_engine *e)
{
    int flag = 0;
    TC_CUDA(cudaMemcpyAsync(&flag, e->error_flag.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    if (flag != 0) {
        cudaMemsetAsync(e->error_flag.p, 0, sizeof(int), e->stream);
        TC_CHECK(flag != 1, TC_ECAPACITY,
                 "a tree exhausted its node/edge/hash pool (max_nodes_per_tree=%d, max_edges_per_tree=%u); "
                 "raise the capacities in tc_config", e->cfg.max_nodes_per_tree, e->ts.capE);
        TC_CHECK(false, TC_ESTATE, "search kernel found a live position without legal moves (flag %d)", flag);
    }
    return TC_OK;
}

extern "C" int tc_abi_version(void) { return TC_ABI_VERSION; }
extern "C" const char *tc_last_error(void) { return g_error; }

extern "C" int tc_host_alloc(size_t bytes, void **out)
{
    TC_CHECK(out != nullptr, TC_EINVAL, "tc_host_alloc: null output");
    *out = nullptr;
    if (bytes == 0) return TC_OK;
    TC_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return TC_OK;
}

extern "C" int tc_host_free(void *p)
{
    if (p) TC_CUDA(cudaFreeHost(p));
    return TC_OK;
}

extern "C" int tc_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes)
{
    TC_CHECK(free_bytes && total_bytes, TC_EINVAL, "tc_device_memory: null output");
    TC_CUDA(cudaSetDevice(device));
    size_t f = 0, t = 0;
    TC_CUDA(cudaMemGetInfo(&f, &t));
    *free_bytes = f;
    *total_bytes = t;
    return TC_OK;
}

extern "C" int tc_create(const tc_config *cfg, tc_engine **out)
{
    TC_CHECK(cfg && out, TC_EINVAL, "tc_create: null argument");
    TC_CHECK(cfg->num_envs >= 1, TC_EINVAL, "num_envs must be >= 1");
    TC_CHECK(cfg->max_nodes_per_tree >= 2 && (uint32_t)cfg->max_nodes_per_tree <= MAX_NODES_PER_TREE, TC_EINVAL,
             "max_nodes_per_tree out of range [2, %u]", MAX_NODES_PER_TREE);
    TC_CHECK(cfg->evaluator >= TC_EVAL_NET_F32 && cfg->evaluator <= TC_EVAL_NET_BF16_MASKED, TC_EINVAL,
             "unknown evaluator %d", cfg->evaluator);
    int ndev = 0;
    TC_CUDA(cudaGetDeviceCount(&ndev));
    TC_CHECK(cfg->device >= 0 && cfg->device < ndev, TC_EINVAL, "device %d not present (%d visible)", cfg->device, ndev);
    TC_CUDA(cudaSetDevice(cfg->device));
    tc_engine *e = new tc_engine();
    e->cfg = *cfg;
    e->evaluator = cfg->evaluator;
    e->hash_params.salt = cfg->hash_salt;
    e->hash_params.pi_levels = cfg->hash_pi_levels;
    e->hash_params.v_levels = cfg->hash_v_levels;
    const size_t E = cfg->num_envs, capN = cfg->max_nodes_per_tree;
    const size_t capE = cfg->max_edges_per_tree > 0 ? (size_t)cfg->max_edges_per_tree : 12 * capN;
    size_t hs = 64;
    while (hs < 2 * capN) hs <<= 1;
    auto fail = [&](int code) {
        delete e;
        return code;
    };
#define TC_ALLOC(buf, count)                                                                           \
    do {                                                                                               \
        cudaError_t _e = (buf).alloc(count);                                                           \
        if (_e != cudaSuccess) {                                                                       \
            set_error("tc_create: cudaMalloc of %zu bytes failed: %s", (size_t)(buf).bytes(),          \
                      cudaGetErrorString(_e));                                                         \
            return fail(TC_ECUDA);                                                                     \
        }                                                                                              \
    } while (0)
    TC_ALLOC(e->node_hdr, 2 * E * capN);
    TC_ALLOC(e->edge_mem, 4 * E * capE);
    TC_ALLOC(e->hash, E * hs);
    TC_ALLOC(e->ctl, E);
    TC_ALLOC(e->path, E * MAX_DEPTH);
    TC_ALLOC(e->eval_state, 8 * E);
    TC_ALLOC(e->eval_pi, 81 * E);
    TC_ALLOC(e->eval_v, E);
    TC_ALLOC(e->eval_count, 8);          // a pair per partition: simulation s counts its leaves in [s & 1]
    TC_ALLOC(e->error_flag, 1);
    TC_ALLOC(e->root_states, 8 * E);
    TC_ALLOC(e->root_active, E);
    TC_ALLOC(e->counts, 81 * E);
    TC_ALLOC(e->root_q, 81 * E);
    TC_ALLOC(e->fld_board, 9 * E);
    TC_ALLOC(e->fld_occ, 9 * E);
    TC_ALLOC(e->fld_last, E);
    TC_ALLOC(e->totals, 16);
    TC_ALLOC(e->log_count, 1);
    if (cfg->eval_log_capacity > 0) {
        TC_ALLOC(e->log_states, 8 * (size_t)cfg->eval_log_capacity);
        TC_ALLOC(e->log_pi, 81 * (size_t)cfg->eval_log_capacity);
        TC_ALLOC(e->log_v, (size_t)cfg->eval_log_capacity);
    }
#undef TC_ALLOC
    if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("tc_create: cudaStreamCreate failed");
        return fail(TC_ECUDA);
    }
    e->stream = e->own_stream;
    TreeStore &ts = e->ts;
    ts.E = (int)E;
    ts.capN = (uint32_t)capN;
    ts.capE = (uint32_t)capE;
    ts.hashSize = (uint32_t)hs;
    ts.node_hdr = e->node_hdr.p;
    ts.edge_mem = e->edge_mem.p;
    ts.hash = e->hash.p;
    ts.ctl = e->ctl.p;
    ts.path = e->path.p;
    ts.eval_state = e->eval_state.p;
    ts.eval_pi = e->eval_pi.p;
    ts.eval_v = e->eval_v.p;
    ts.eval_count = e->eval_count.p;
    ts.error_flag = e->error_flag.p;
    {
        const char *hv = getenv("TACULT_HALVES");
        // One partition by default: with programmatic dependent launches the single chain select -> trunk -> fc1 ->
        // heads runs back to back, and full-batch kernels are more efficient than two half-batch chains overlapped on
        // two streams (measured 30.5 M vs 30.3 M sims/s at 4096 trees).  TACULT_HALVES=2..4 splits the trees.
        e->n_halves = 1;
        if (hv) {
            const int want = atoi(hv);
            e->n_halves = (want >= 1 && want <= 4 && E >= want) ? want : 1;
        }
        // the evaluation log is appended in stream order: one partition only, or log_count would race across streams
        if (cfg->eval_log_capacity > 0) e->n_halves = 1;
        // partition sizes are multiples of 32 trees: per-partition row blocks of the leaf buffers stay 16-byte aligned
        const int per = (((int)((E + e->n_halves - 1) / e->n_halves)) + 31) & ~31;
        e->n_halves = (int)((E + per - 1) / per);
        for (int h = 0; h < e->n_halves; ++h) {
            const size_t t0 = (size_t)h * per;
            TreeStore v = ts;
            v.E = std::min(per, (int)E - (int)t0);
            v.node_hdr = ts.node_hdr + 2 * capN * t0;
            v.edge_mem = ts.edge_mem + 4 * capE * t0;
            v.hash = ts.hash + hs * t0;
            v.ctl = ts.ctl + t0;
            v.path = ts.path + (size_t)MAX_DEPTH * t0;
            v.eval_state = ts.eval_state + 8 * t0;
            v.eval_pi = ts.eval_pi + 81 * t0;
            v.eval_v = ts.eval_v + t0;
            v.eval_count = ts.eval_count + 2 * h;
            e->hts[h] = v;
        }
        if (e->n_halves > 1) {
            bool ok = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) == cudaSuccess;
            for (int h = 1; h < e->n_halves && ok; ++h)
                ok = cudaStreamCreateWithFlags(&e->side_stream[h - 1], cudaStreamNonBlocking) == cudaSuccess &&
                     cudaEventCreateWithFlags(&e->ev_join[h - 1], cudaEventDisableTiming) == cudaSuccess;
            if (!ok) {
                set_error("tc_create: side streams / events could not be created");
                return fail(TC_ECUDA);
            }
        }
    }
    cudaMemsetAsync(e->ctl.p, 0, e->ctl.bytes(), e->stream);
    cudaMemsetAsync(e->error_flag.p, 0, sizeof(int), e->stream);
    cudaMemsetAsync(e->eval_count.p, 0, 8 * sizeof(int), e->stream);
    cudaMemsetAsync(e->log_count.p, 0, sizeof(int), e->stream);
    cudaMemsetAsync(e->root_active.p, 0, e->root_active.bytes(), e->stream);
    launch_reset(ts, -1, e->stream);
    e->launches += 1;
    if (cudaStreamSynchronize(e->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        set_error("tc_create: device initialisation failed");
        return fail(TC_ECUDA);
    }
    *out = e;
    return TC_OK;
}

extern "C" int tc_destroy(tc_engine *e)
{
    if (!e) return TC_OK;
    cudaSetDevice(e->cfg.device);
    cudaStreamSynchronize(e->stream);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    for (int h = 0; h < 3; ++h) {
        if (e->side_stream[h]) { cudaStreamSynchronize(e->side_stream[h]); cudaStreamDestroy(e->side_stream[h]); }
        if (e->ev_join[h]) cudaEventDestroy(e->ev_join[h]);
    }
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    delete e;
    return TC_OK;
}

extern "C" int tc_set_stream(tc_engine *e, void *cuda_stream)
{
    TC_TRY(check_engine(e));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return TC_OK;
}

extern "C" int tc_sync(tc_engine *e)
{
    TC_TRY(check_engine(e));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    return TC_OK;
}

extern "C" int tc_reset(tc_engine *e, int env)
{
    TC_TRY(check_engine(e));
    TC_CHECK(env < e->ts.E, TC_EINVAL, "tc_reset: env %d out of range", env);
    launch_reset(e->ts, env, e->stream);
    e->launches += 1;
    TC_CUDA(cudaGetLastError());
    return TC_OK;
}

extern "C" int tc_reset_envs(tc_engine *e, int n, const int32_t *envs_host)
{
    TC_TRY(check_engine(e));
    TC_CHECK(n >= 0 && (n == 0 || envs_host), TC_EINVAL, "tc_reset_envs: null list");
    if (n == 0) return TC_OK;
    for (int i = 0; i < n; ++i)
        TC_CHECK(envs_host[i] >= 0 && envs_host[i] < e->ts.E, TC_EINVAL, "tc_reset_envs: env %d out of range", envs_host[i]);
    if ((size_t)n > e->reset_list.n) TC_CUDA(e->reset_list.alloc(std::max<size_t>((size_t)n, (size_t)e->ts.E)));
    TC_CUDA(cudaMemcpyAsync(e->reset_list.p, envs_host, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    launch_reset_list(e->ts, n, e->reset_list.p, e->stream);
    e->launches += 1;
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(e->stream));      // the host list may be reused by the caller
    return TC_OK;
}

extern "C" int tc_set_hash_evaluator(tc_engine *e, uint64_t salt, int pi_levels, int v_levels)
{
    TC_TRY(check_engine(e));
    TC_CHECK(pi_levels >= 0 && v_levels >= 0 && v_levels <= 4, TC_EINVAL, "bad hash evaluator levels");
    e->hash_params.salt = salt;
    e->hash_params.pi_levels = pi_levels;
    e->hash_params.v_levels = v_levels;
    return TC_OK;
}

extern "C" int tc_set_evaluator(tc_engine *e, int evaluator)
{
    TC_TRY(check_engine(e));
    TC_CHECK(evaluator >= TC_EVAL_NET_F32 && evaluator <= TC_EVAL_NET_BF16_MASKED, TC_EINVAL, "unknown evaluator %d",
             evaluator);
    e->evaluator = evaluator;
    return TC_OK;
}

static int set_roots_dev(tc_engine *e, const uint8_t *active_host)
{
    if (active_host)
        TC_CUDA(cudaMemcpyAsync(e->root_active.p, active_host, e->ts.E, cudaMemcpyHostToDevice, e->stream));
    launch_set_roots(e->ts, e->root_states.p, active_host ? e->root_active.p : nullptr, e->stream);
    e->launches += 1;
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(e->stream));     // the host buffers may be reused by the caller
    return TC_OK;
}

extern "C" int tc_set_roots(tc_engine *e, const uint16_t *board_host, const uint16_t *occ_host,
                            const int8_t *last_square_host, const uint8_t *active_host)
{
    TC_TRY(check_engine(e));
    TC_CHECK(board_host && occ_host && last_square_host, TC_EINVAL, "tc_set_roots: null buffer");
    const int E = e->ts.E;
    TC_CUDA(cudaMemcpyAsync(e->fld_board.p, board_host, 9 * (size_t)E * sizeof(uint16_t), cudaMemcpyHostToDevice,
                            e->stream));
    TC_CUDA(cudaMemcpyAsync(e->fld_occ.p, occ_host, 9 * (size_t)E * sizeof(uint16_t), cudaMemcpyHostToDevice,
                            e->stream));
    TC_CUDA(cudaMemcpyAsync(e->fld_last.p, last_square_host, (size_t)E, cudaMemcpyHostToDevice, e->stream));
    launch_pack_fields(E, e->fld_board.p, e->fld_occ.p, e->fld_last.p, e->root_states.p, e->stream);
    e->launches += 1;
    return set_roots_dev(e, active_host);
}

extern "C" int tc_set_roots_packed(tc_engine *e, const uint32_t *states_host, const uint8_t *active_host)
{
    TC_TRY(check_engine(e));
    TC_CHECK(states_host, TC_EINVAL, "tc_set_roots_packed: null buffer");
    TC_CUDA(cudaMemcpyAsync(e->root_states.p, states_host, 8 * (size_t)e->ts.E * sizeof(uint32_t),
                            cudaMemcpyHostToDevice, e->stream));
    return set_roots_dev(e, active_host);
}

extern "C" int tc_advance(tc_engine *e, const int8_t *actions_host, int restart, uint32_t *states_host, int8_t *status_host)
{
    TC_TRY(check_engine(e));
    TC_CHECK(actions_host, TC_EINVAL, "tc_advance: null actions");
    const size_t E = (size_t)e->ts.E;
    if (e->adv_status.n < E) {
        TC_CUDA(e->adv_actions.alloc(E));
        TC_CUDA(e->adv_status.alloc(E));
        TC_CUDA(e->adv_illegal.alloc(1));
    }
    TC_CUDA(cudaMemcpyAsync(e->adv_actions.p, actions_host, E, cudaMemcpyHostToDevice, e->stream));
    TC_CUDA(cudaMemsetAsync(e->adv_illegal.p, 0, sizeof(int), e->stream));
    launch_advance(e->ts, e->adv_actions.p, restart, e->root_states.p, e->adv_status.p, e->adv_illegal.p, e->stream);
    e->launches += 1;
    TC_CUDA(cudaGetLastError());
    int bad = 0;
    if (states_host)
        TC_CUDA(cudaMemcpyAsync(states_host, e->root_states.p, 32 * E, cudaMemcpyDeviceToHost, e->stream));
    if (status_host) TC_CUDA(cudaMemcpyAsync(status_host, e->adv_status.p, E, cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaMemcpyAsync(&bad, e->adv_illegal.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    TC_CHECK(bad == 0, TC_EILLEGAL, "tc_advance: %d illegal move(s) requested", bad);
    return TC_OK;
}

// One lock-step simulation for every tree of part `h`: select -> evaluate leaves -> expand + backup.
// Inside a run of simulations the backup of simulation k and the descent of simulation k+1 share a launch
// (`first` = no pending backup, `last` = no following descent).
static int enqueue_simulation(tc_engine *e, double cpuct, int h, cudaStream_t st, int sim, bool last)
{
    // leaf counter: simulation `sim` counts in slot sim & 1 of the partition's pair; its descent kernel zeroes the other
    // slot for simulation sim + 1, so no memset node sits between the kernels of the chain (they are launched with
    // programmatic stream serialisation, common.cuh)
    TreeStore ts = e->hts[h];
    ts.priors_masked = e->evaluator == TC_EVAL_NET_BF16_MASKED;
    ts.pdl_trigger = pdl_trigger_enabled();
    int *pair = ts.eval_count;
    ts.eval_count = pair + (sim & 1);
    ts.eval_count_next = pair + ((sim + 1) & 1);
    const bool first = sim == 0;
    NetF32 &n32 = h == 0 ? e->net32 : e->net32x[h - 1];
    NetBF16 &n16 = h == 0 ? e->net16 : e->net16x[h - 1];
    e->prof.begin(TC_K_SELECT, st);
    if (first) launch_select(ts, cpuct, st);
    else launch_backup_select(ts, cpuct, st);
    e->prof.end(st);
    e->launches += 1;
    if (e->evaluator == TC_EVAL_HASH) {
        e->prof.begin(TC_K_EVAL_HASH, st);
        launch_hash_eval(ts, e->hash_params, st);
        e->prof.end(st);
        e->launches += 1;
    } else if (e->evaluator == TC_EVAL_NET_F32) {
        e->launches += n32.forward(ts.E, ts.eval_count, ts.eval_state, nullptr, ts.eval_pi, ts.eval_v, nullptr, st, &e->prof);
    } else {
        int n = n16.forward(ts.E, ts.eval_count, ts.eval_state, nullptr, ts.eval_pi, ts.eval_v, nullptr, st, &e->prof,
                            ts.priors_masked);
        if (n < 0) return n;
        e->launches += n;
    }
    if (e->cfg.eval_log_capacity > 0) {      // single-half mode only (tc_create)
        k_log_append<<<64, 256, 0, st>>>(ts, e->log_states.p, e->log_pi.p, e->log_v.p, e->log_count.p,
                                         e->cfg.eval_log_capacity);
        k_log_advance<<<1, 1, 0, st>>>(ts, e->log_count.p);
        e->launches += 2;
    }
    if (last) {
        e->prof.begin(TC_K_EXPAND_BACKUP, st);
        launch_expand_backup(ts, st);
        e->prof.end(st);
        e->launches += 1;
    }
    return TC_OK;
}

// `num_sims` lock-step simulations of every tree; with two halves the second runs on the side stream
static int enqueue_search(tc_engine *e, int num_sims, double cpuct)
{
    if (num_sims <= 0) return TC_OK;
    TC_CUDA(cudaMemsetAsync(e->eval_count.p, 0, 8 * sizeof(int), e->stream));      // slot 0 of every pair for simulation 0
    if (e->n_halves > 1) {
        TC_CUDA(cudaEventRecord(e->ev_fork, e->stream));
        for (int h = 1; h < e->n_halves; ++h) TC_CUDA(cudaStreamWaitEvent(e->side_stream[h - 1], e->ev_fork, 0));
    }
    for (int s = 0; s < num_sims; ++s)
        for (int h = 0; h < e->n_halves; ++h)
            TC_TRY(enqueue_simulation(e, cpuct, h, h == 0 ? e->stream : e->side_stream[h - 1], s, s == num_sims - 1));
    for (int h = 1; h < e->n_halves; ++h) {
        TC_CUDA(cudaEventRecord(e->ev_join[h - 1], e->side_stream[h - 1]));
        TC_CUDA(cudaStreamWaitEvent(e->stream, e->ev_join[h - 1], 0));
    }
    return TC_OK;
}

extern "C" int tc_profile_enable(tc_engine *e, int on)
{
    TC_TRY(check_engine(e));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    double ms[TC_K_CLASSES] = {0};
    unsigned long long n[TC_K_CLASSES] = {0};
    e->prof.collect(ms, n, TC_K_CLASSES);
    e->prof.on = on != 0;
    return TC_OK;
}

extern "C" int tc_profile_read(tc_engine *e, double *ms_by_class, uint64_t *launches_by_class)
{
    TC_TRY(check_engine(e));
    TC_CHECK(ms_by_class && launches_by_class, TC_EINVAL, "tc_profile_read: null output");
    TC_CUDA(cudaStreamSynchronize(e->stream));
    unsigned long long n[TC_K_CLASSES] = {0};
    for (int k = 0; k < TC_K_CLASSES; ++k) ms_by_class[k] = 0.0;
    e->prof.collect(ms_by_class, n, TC_K_CLASSES);
    for (int k = 0; k < TC_K_CLASSES; ++k) launches_by_class[k] = n[k];
    return TC_OK;
}

static int check_evaluator_ready(tc_engine *e)
{
    if (e->evaluator == TC_EVAL_HASH) return TC_OK;
    TC_CHECK(e->have_weights, TC_ESTATE, "network evaluator selected but tc_load_weights has not been called");
    if (e->evaluator == TC_EVAL_NET_BF16 || e->evaluator == TC_EVAL_NET_BF16_MASKED)
        TC_CHECK(e->net16.loaded, TC_ESTATE, "bf16 evaluator not available for this network shape: %s",
                 e->net16.why_not.c_str());
    return TC_OK;
}

extern "C" int tc_search(tc_engine *e, int num_sims, double cpuct)
{
    TC_TRY(check_engine(e));
    TC_CHECK(num_sims >= 0, TC_EINVAL, "tc_search: negative num_sims");
    TC_TRY(check_evaluator_ready(e));
    TC_TRY(enqueue_search(e, num_sims, cpuct));
    TC_CUDA(cudaGetLastError());
    return check_device_errors(e);
}

extern "C" int tc_root_counts(tc_engine *e, uint32_t *counts_host)
{
    TC_TRY(check_engine(e));
    TC_CHECK(counts_host, TC_EINVAL, "tc_root_counts: null buffer");
    launch_root_counts(e->ts, e->counts.p, nullptr, e->stream);
    e->launches += 1;
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaMemcpyAsync(counts_host, e->counts.p, e->counts.bytes(), cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    return TC_OK;
}

extern "C" int tc_root_q(tc_engine *e, double *q_host)
{
    TC_TRY(check_engine(e));
    TC_CHECK(q_host, TC_EINVAL, "tc_root_q: null buffer");
    launch_root_counts(e->ts, nullptr, e->root_q.p, e->stream);
    e->launches += 1;
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaMemcpyAsync(q_host, e->root_q.p, e->root_q.bytes(), cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    return TC_OK;
}

extern "C" int tc_get_stats(tc_engine *e, tc_stats *out)
{
    TC_TRY(check_engine(e));
    TC_CHECK(out, TC_EINVAL, "tc_get_stats: null output");
    TC_CUDA(cudaMemsetAsync(e->totals.p, 0, e->totals.bytes(), e->stream));
    k_tree_totals<<<std::min(ceil_div(e->ts.E, 256), 64), 256, 0, e->stream>>>(e->ts, e->totals.p);
    e->launches += 1;
    unsigned long long h[16];
    TC_CUDA(cudaMemcpyAsync(h, e->totals.p, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    out->sims = h[0]; out->sum_depth = h[1]; out->sum_legal = h[2]; out->nn_leaves = h[3];
    out->sum_leaf_legal = h[4]; out->terminal_leaves = h[5]; out->transpositions = h[6];
    out->nodes_in_use = h[7]; out->edges_in_use = h[8]; out->max_nodes_one_tree = h[9]; out->max_edges_one_tree = h[10];
    out->kernel_launches = e->launches;
    return TC_OK;
}

extern "C" int tc_reset_stats(tc_engine *e)
{
    TC_TRY(check_engine(e));
    k_zero_counters<<<ceil_div(e->ts.E, 256), 256, 0, e->stream>>>(e->ts);
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(e->stream));
    e->launches = 0;
    return TC_OK;
}

extern "C" int tc_eval_log(tc_engine *e, int max, uint32_t *states_host, float *pi_host, float *v_host, int *count_out)
{
    TC_TRY(check_engine(e));
    TC_CHECK(count_out, TC_EINVAL, "tc_eval_log: null count_out");
    int n = 0;
    TC_CUDA(cudaMemcpyAsync(&n, e->log_count.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    *count_out = n;
    n = std::min(std::min(n, max), e->cfg.eval_log_capacity);
    if (n > 0) {
        if (states_host) TC_CUDA(cudaMemcpy(states_host, e->log_states.p, 8 * (size_t)n * 4, cudaMemcpyDeviceToHost));
        if (pi_host) TC_CUDA(cudaMemcpy(pi_host, e->log_pi.p, 81 * (size_t)n * 4, cudaMemcpyDeviceToHost));
        if (v_host) TC_CUDA(cudaMemcpy(v_host, e->log_v.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    }
    return TC_OK;
}

// ---- evaluator ---------------------------------------------------------------------------------
extern "C" size_t tc_weight_blob_floats(const tc_net_desc *desc) { return desc ? blob_floats(*desc) : 0; }

extern "C" int tc_load_weights(tc_engine *e, const tc_net_desc *desc, const float *blob_host, size_t n_floats)
{
    TC_TRY(check_engine(e));
    TC_CHECK(desc && blob_host, TC_EINVAL, "tc_load_weights: null argument");
    TC_CUDA(cudaStreamSynchronize(e->stream));
    TC_TRY(fold_blob(*desc, blob_host, n_floats, e->folded));
    TC_CHECK(desc->channels < 32 || desc->channels % 32 == 0, TC_EINVAL,
             "channels must be < 32 or a multiple of 32, got %d", desc->channels);
    for (int h = 0; h < 3; ++h)
        if (e->side_stream[h]) TC_CUDA(cudaStreamSynchronize(e->side_stream[h]));
    const int cap = std::max(e->hts[0].E, 1);
    TC_TRY(e->net32.upload(e->folded, cap));
    e->net16.upload(e->folded, cap, e->n_halves > 1);       // optional: records why_not when the shape is unsupported
    for (int h = 1; h < e->n_halves; ++h) {
        TC_TRY(e->net32x[h - 1].upload(e->folded, std::max(e->hts[h].E, 1)));
        e->net16x[h - 1].upload(e->folded, std::max(e->hts[h].E, 1), true);
    }
    // the uploads above are blocking cudaMemcpy / cudaMemset calls on the legacy stream, while every engine stream is
    // non-blocking: make sure the last DMA and zero-fill have landed before a search can be enqueued
    TC_CUDA(cudaDeviceSynchronize());
    e->have_weights = true;
    return TC_OK;
}

static int ensure_fwd_scratch(tc_engine *e, int batch)
{
    if (batch <= e->fwd_capacity) return TC_OK;
    TC_CUDA(e->fwd_obs.alloc((size_t)batch * e->folded.d.in_planes * 81));
    TC_CUDA(e->fwd_states.alloc((size_t)batch * 8));
    TC_CUDA(e->fwd_pi.alloc((size_t)batch * 81));
    TC_CUDA(e->fwd_v.alloc((size_t)batch));
    TC_CUDA(e->fwd_logits.alloc((size_t)batch * 81));
    e->fwd_capacity = batch;
    return TC_OK;
}

static int run_forward(tc_engine *e, int evaluator, int batch, const uint32_t *states_dev, const float *obs_dev,
                       float *pi_dev, float *v_dev, float *logits_dev)
{
    TC_CHECK(e->have_weights, TC_ESTATE, "tc_forward: tc_load_weights has not been called");
    TC_CHECK(evaluator == TC_EVAL_NET_F32 || evaluator == TC_EVAL_NET_BF16 || evaluator == TC_EVAL_NET_BF16_MASKED, TC_EINVAL,
             "tc_forward: evaluator must be TC_EVAL_NET_F32, TC_EVAL_NET_BF16 or TC_EVAL_NET_BF16_MASKED");
    TC_CHECK(evaluator != TC_EVAL_NET_BF16_MASKED || (states_dev && !logits_dev), TC_EINVAL,
             "the masked policy needs packed positions (tc_forward_states*) and no logits output");
    TC_CHECK(states_dev == nullptr || e->folded.d.in_planes == 4, TC_EINVAL,
             "packed positions imply 4 input planes, the loaded network has %d", e->folded.d.in_planes);
    // the network scratch is sized for num_envs rows: run larger batches in slices
    const int cap = e->net32.capacity;
    for (int r0 = 0; r0 < batch; r0 += cap) {
        const int nb = std::min(cap, batch - r0);
        const uint32_t *s = states_dev ? states_dev + (size_t)r0 * 8 : nullptr;
        const float *o = obs_dev ? obs_dev + (size_t)r0 * e->folded.d.in_planes * 81 : nullptr;
        float *lg = logits_dev ? logits_dev + (size_t)r0 * 81 : nullptr;
        if (evaluator == TC_EVAL_NET_F32) {
            e->launches += e->net32.forward(nb, nullptr, s, o, pi_dev + (size_t)r0 * 81, v_dev + r0, lg, e->stream, &e->prof);
        } else {
            TC_CHECK(e->net16.loaded, TC_ESTATE, "bf16 evaluator not available for this network shape: %s",
                     e->net16.why_not.c_str());
            int n = e->net16.forward(nb, nullptr, s, o, pi_dev + (size_t)r0 * 81, v_dev + r0, lg, e->stream, &e->prof,
                                     evaluator == TC_EVAL_NET_BF16_MASKED);
            if (n < 0) return n;
            e->launches += n;
        }
    }
    TC_CUDA(cudaGetLastError());
    return TC_OK;
}

extern "C" int tc_forward(tc_engine *e, int evaluator, int batch, const float *obs_host, float *pi_host, float *v_host,
                          float *logits_host)
{
    TC_TRY(check_engine(e));
    TC_CHECK(batch >= 0 && (batch == 0 || (obs_host && pi_host && v_host)), TC_EINVAL, "tc_forward: null buffer");
    if (batch == 0) return TC_OK;
    TC_CHECK(e->have_weights, TC_ESTATE, "tc_forward: tc_load_weights has not been called");
    TC_TRY(ensure_fwd_scratch(e, batch));
    const size_t in_floats = (size_t)batch * e->folded.d.in_planes * 81;
    TC_CUDA(cudaMemcpyAsync(e->fwd_obs.p, obs_host, in_floats * 4, cudaMemcpyHostToDevice, e->stream));
    TC_TRY(run_forward(e, evaluator, batch, nullptr, e->fwd_obs.p, e->fwd_pi.p, e->fwd_v.p,
                       logits_host ? e->fwd_logits.p : nullptr));
    TC_CUDA(cudaMemcpyAsync(pi_host, e->fwd_pi.p, (size_t)batch * 81 * 4, cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaMemcpyAsync(v_host, e->fwd_v.p, (size_t)batch * 4, cudaMemcpyDeviceToHost, e->stream));
    if (logits_host)
        TC_CUDA(cudaMemcpyAsync(logits_host, e->fwd_logits.p, (size_t)batch * 81 * 4, cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    return TC_OK;
}

extern "C" int tc_forward_states(tc_engine *e, int evaluator, int batch, const uint32_t *states_host, float *pi_host,
                                 float *v_host)
{
    TC_TRY(check_engine(e));
    TC_CHECK(batch >= 0 && (batch == 0 || (states_host && pi_host && v_host)), TC_EINVAL,
             "tc_forward_states: null buffer");
    if (batch == 0) return TC_OK;
    TC_CHECK(e->have_weights, TC_ESTATE, "tc_forward_states: tc_load_weights has not been called");
    TC_TRY(ensure_fwd_scratch(e, batch));
    TC_CUDA(cudaMemcpyAsync(e->fwd_states.p, states_host, (size_t)batch * 32, cudaMemcpyHostToDevice, e->stream));
    TC_TRY(run_forward(e, evaluator, batch, e->fwd_states.p, nullptr, e->fwd_pi.p, e->fwd_v.p, nullptr));
    TC_CUDA(cudaMemcpyAsync(pi_host, e->fwd_pi.p, (size_t)batch * 81 * 4, cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaMemcpyAsync(v_host, e->fwd_v.p, (size_t)batch * 4, cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    return TC_OK;
}

extern "C" int tc_forward_states_dev(tc_engine *e, int evaluator, int batch, const uint32_t *states_dev, float *pi_dev,
                                     float *v_dev)
{
    TC_TRY(check_engine(e));
    TC_CHECK(batch >= 0 && (batch == 0 || (states_dev && pi_dev && v_dev)), TC_EINVAL,
             "tc_forward_states_dev: null buffer");
    if (batch == 0) return TC_OK;
    return run_forward(e, evaluator, batch, states_dev, nullptr, pi_dev, v_dev, nullptr);
}

extern "C" int tc_forward_obs_dev(tc_engine *e, int evaluator, int batch, const float *obs_dev, float *pi_dev, float *v_dev)
{
    TC_TRY(check_engine(e));
    TC_CHECK(batch >= 0 && (batch == 0 || (obs_dev && pi_dev && v_dev)), TC_EINVAL, "tc_forward_obs_dev: null buffer");
    if (batch == 0) return TC_OK;
    return run_forward(e, evaluator, batch, nullptr, obs_dev, pi_dev, v_dev, nullptr);
}

extern "C" int tc_debug_trunk(tc_engine *e, int evaluator, int batch, const uint32_t *states_host, int n_convs,
                              float *out_host)
{
    TC_TRY(check_engine(e));
    TC_CHECK(e->have_weights, TC_ESTATE, "tc_debug_trunk: tc_load_weights has not been called");
    TC_CHECK(batch >= 1 && batch <= e->net32.capacity && states_host && out_host && n_convs >= 0, TC_EINVAL,
             "tc_debug_trunk: bad argument (batch must be <= %d)", e->net32.capacity);
    const size_t n = (size_t)batch * 81 * e->folded.d.channels;
    DevBuf<uint32_t> st;
    DevBuf<float> out;
    TC_CUDA(st.alloc((size_t)batch * 8));
    TC_CUDA(out.alloc(n));
    TC_CUDA(cudaMemcpyAsync(st.p, states_host, (size_t)batch * 32, cudaMemcpyHostToDevice, e->stream));
    if (evaluator == TC_EVAL_NET_F32) TC_TRY(e->net32.debug_trunk(batch, st.p, n_convs, out.p, e->stream));
    else TC_TRY(debug_trunk_bf16(e->net16, batch, st.p, n_convs, out.p, e->stream));
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaMemcpyAsync(out_host, out.p, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    return TC_OK;
}

// ---- self-play driver --------------------------------------------------------------------------
extern "C" int tc_selfplay_begin(tc_engine *e, const tc_selfplay_config *cfg)
{
    TC_TRY(check_engine(e));
    TC_CHECK(cfg && cfg->num_sims >= 1 && cfg->temp_threshold >= 0, TC_EINVAL, "tc_selfplay_begin: bad config");
    TC_TRY(check_evaluator_ready(e));
    const size_t E = e->ts.E;
    if (!e->sp_ready) {
        TC_CUDA(e->sp_env.alloc(E));
        TC_CUDA(e->sp_pending.alloc(E * 81));
        TC_CUDA(e->sp_finished.alloc(E * 256));
        TC_CUDA(e->sp_finished_count.alloc(1));
        TC_CUDA(e->sp_stats.alloc(8));
        e->sp.E = (int)E;
        e->sp.env = e->sp_env.p;
        e->sp.pending = e->sp_pending.p;
        e->sp.finished = e->sp_finished.p;
        e->sp.finished_cap = (int)(E * 256);
        e->sp.finished_count = e->sp_finished_count.p;
        e->sp.stats = e->sp_stats.p;
        e->sp.root_states = e->root_states.p;
        e->sp.root_active = e->root_active.p;
        e->sp_ready = true;
    }
    e->sp_cfg = *cfg;
    launch_selfplay_init(e->sp, e->stream);
    launch_reset(e->ts, -1, e->stream);
    e->launches += 2;
    TC_CUDA(cudaGetLastError());
    return TC_OK;
}

extern "C" int tc_selfplay_configure(tc_engine *e, const tc_selfplay_config *cfg)
{
    TC_TRY(check_engine(e));
    TC_CHECK(e->sp_ready, TC_ESTATE, "tc_selfplay_configure before tc_selfplay_begin");
    TC_CHECK(cfg && cfg->num_sims >= 1 && cfg->temp_threshold >= 0, TC_EINVAL, "tc_selfplay_configure: bad config");
    e->sp_cfg = *cfg;
    return TC_OK;
}

extern "C" int tc_selfplay_step(tc_engine *e, int plies)
{
    TC_TRY(check_engine(e));
    TC_CHECK(e->sp_ready, TC_ESTATE, "tc_selfplay_step before tc_selfplay_begin");
    TC_CHECK(plies >= 0, TC_EINVAL, "negative plies");
    for (int p = 0; p < plies; ++p) {
        e->prof.begin(TC_K_OTHER, e->stream);
        launch_selfplay_roots(e->sp, e->stream);
        launch_set_roots(e->ts, e->root_states.p, e->root_active.p, e->stream);
        e->prof.end(e->stream);
        e->launches += 2;
        TC_TRY(enqueue_search(e, e->sp_cfg.num_sims, e->sp_cfg.cpuct));
        e->prof.begin(TC_K_OTHER, e->stream);
        launch_selfplay_move(e->ts, e->sp, e->sp_cfg, e->stream);
        e->launches += 1;
        if (e->sp_cfg.auto_reset) {
            launch_reset_marked(e->ts, e->sp, e->stream);
            e->launches += 1;
        }
        e->prof.end(e->stream);
    }
    TC_CUDA(cudaGetLastError());
    return TC_OK;
}

extern "C" int tc_selfplay_sync(tc_engine *e)
{
    TC_TRY(check_engine(e));
    return check_device_errors(e);
}

extern "C" int tc_selfplay_get_stats(tc_engine *e, tc_selfplay_stats *out)
{
    TC_TRY(check_engine(e));
    TC_CHECK(out && e->sp_ready, TC_ESTATE, "tc_selfplay_get_stats before tc_selfplay_begin");
    unsigned long long h[8];
    int nfin = 0;
    TC_CUDA(cudaMemcpyAsync(h, e->sp_stats.p, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaMemcpyAsync(&nfin, e->sp_finished_count.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    out->plies = h[0]; out->sims = h[1]; out->games_finished = h[2];
    out->first_player_wins = h[3]; out->second_player_wins = h[4]; out->draws = h[5];
    out->examples = (uint64_t)std::min(nfin, e->sp.finished_cap);
    return TC_OK;
}

static int drain_common(tc_engine *e, int max, tc_example *out, bool to_host, int *count_out)
{
    TC_CHECK(e->sp_ready, TC_ESTATE, "tc_selfplay_drain before tc_selfplay_begin");
    TC_CHECK(count_out && (max == 0 || out), TC_EINVAL, "tc_selfplay_drain: null argument");
    int nfin = 0;
    unsigned long long dropped = 0;
    TC_CUDA(cudaMemcpyAsync(&nfin, e->sp_finished_count.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaMemcpyAsync(&dropped, e->sp_stats.p + 6, sizeof dropped, cudaMemcpyDeviceToHost, e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    if (dropped != 0) {      // fail loudly, then start over with an empty buffer
        cudaMemsetAsync(e->sp_finished_count.p, 0, sizeof(int), e->stream);
        cudaMemsetAsync(e->sp_stats.p + 6, 0, sizeof(unsigned long long), e->stream);
        cudaStreamSynchronize(e->stream);
        TC_CHECK(false, TC_ECAPACITY, "self-play record buffer (%d records) overflowed: %llu records dropped, the "
                 "buffer was discarded; call tc_selfplay_drain more often", e->sp.finished_cap, dropped);
    }
    TC_CHECK(nfin <= max, TC_EINVAL, "tc_selfplay_drain: %d records pending, buffer holds %d", nfin, max);
    if (nfin > 0)
        TC_CUDA(cudaMemcpyAsync(out, e->sp_finished.p, (size_t)nfin * sizeof(tc_example),
                                to_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, e->stream));
    TC_CUDA(cudaMemsetAsync(e->sp_finished_count.p, 0, sizeof(int), e->stream));
    TC_CUDA(cudaStreamSynchronize(e->stream));
    *count_out = nfin;
    return TC_OK;
}

extern "C" int tc_selfplay_drain(tc_engine *e, int max, tc_example *out_host, int *count_out)
{
    TC_TRY(check_engine(e));
    return drain_common(e, max, out_host, true, count_out);
}

extern "C" int tc_selfplay_drain_dev(tc_engine *e, int max, tc_example *out_dev, int *count_out)
{
    TC_TRY(check_engine(e));
    return drain_common(e, max, out_dev, false, count_out);
}

extern "C" int tc_examples_from_records_dev(tc_engine *e, int n_out, const tc_example *records_dev, int n_records,
                                            const int32_t *index_dev, float *obs_dev, float *pi_dev, float *z_dev)
{
    TC_TRY(check_engine(e));
    TC_CHECK(n_out >= 0 && n_records >= 0, TC_EINVAL, "tc_examples_from_records_dev: negative count");
    if (n_out == 0) return TC_OK;
    TC_CHECK(records_dev && obs_dev && pi_dev && z_dev, TC_EINVAL, "tc_examples_from_records_dev: null buffer");
    TC_CHECK(index_dev || n_out <= 8 * (long long)n_records, TC_EINVAL,
             "tc_examples_from_records_dev: %d outputs from %d records without an index list", n_out, n_records);
    launch_examples_from_records(n_out, records_dev, n_records, index_dev, obs_dev, pi_dev, z_dev, e->stream);
    e->launches += 1;
    TC_CUDA(cudaGetLastError());
    return TC_OK;
}
