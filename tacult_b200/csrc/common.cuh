// Shared host/device helpers for the tacult_b200 CUDA library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <cstdlib>
#include <utility>
#include <vector>

#include "../../include/tacult_b200.h"

namespace tc {

void set_error(const char *fmt, ...);

#define TC_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            tc::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return TC_ECUDA;                                                                       \
        }                                                                                          \
    } while (0)

#define TC_CHECK(cond, code, ...)                                                                  \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            tc::set_error(__VA_ARGS__);                                                            \
            return (code);                                                                         \
        }                                                                                          \
    } while (0)

#define TC_TRY(expr)                                                                               \
    do {                                                                                           \
        int _r = (expr);                                                                           \
        if (_r != TC_OK) return _r;                                                                \
    } while (0)

constexpr int SM_COUNT = 148;      // B200: 2 dies x 74 SMs
constexpr int WARP = 32;

template <typename T>
struct DevBuf {                     // RAII device allocation
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void **)&p, count * sizeof(T));
    }
    size_t bytes() const { return n * sizeof(T); }
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Optional per-kernel-class timing with CUDA events on the launching stream (tc_profile_*).
// Programmatic dependent launch: a kernel launched with this attribute may start (and run its prologue) while the
// previous kernel in the stream is still draining; it must call griddep_wait() before it reads anything the previous
// kernels wrote and before it writes anything they may still read.  The wait returns once every prerequisite grid has
// completed and its memory is visible; in a normally launched kernel it is a no-op.  No kernel here triggers its
// dependents early: waiting CTAs would hold shared memory the other partition's kernels need.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next kernel of the chain start launching now: its CTAs become resident as soon as an SM has room for them, run
// their prologue and park in griddep_wait() until THIS grid has completed (correctness never depends on the trigger).
// Opt-in (TACULT_PDL_TRIGGER=1): parked CTAs hold registers / shared memory.
__device__ __forceinline__ void griddep_launch(int on)
{
    if (on) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline int pdl_trigger_enabled()
{
    // 3: only the tree kernels trigger their dependents at entry - the trunk's CTAs, which need a whole SM, become resident
    // as the descent's last trees finish and run their prologue (shared-memory set-up, TMEM allocation, weight TMA) behind
    // that tail.  Every kernel launched with the programmatic attribute calls griddep_wait() before it reads anything a
    // predecessor wrote.
    static const int on = [] { const char *v = getenv("TACULT_PDL_TRIGGER"); return v ? atoi(v) : 3; }();
    return on;
}

inline bool pdl_enabled()
{
    static const int on = [] { const char *v = getenv("TACULT_PDL"); return v ? atoi(v) : 1; }();
    return on != 0;
}

// TACULT_CARVEOUT=1: every kernel of the simulation chain asks for the maximum shared-memory carve-out, so that an SM does
// not have to re-partition its L1 / shared memory between the tree kernels (no shared memory) and the network kernels
// (87 - 228 KB) four times per simulation.
inline int carveout_enabled()
{
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("TACULT_CARVEOUT");
        on = (e && atoi(e) != 0) ? 1 : 0;
    }
    return on;
}
template <typename K>
inline void prefer_max_shared(K kernel)
{
    if (carveout_enabled()) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

struct Profiler {
    bool on = false;
    struct Span { int cls; cudaEvent_t a, b; cudaStream_t st; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t open_a = nullptr;
    int open_cls = -1;
    cudaEvent_t get()
    {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void begin(int cls, cudaStream_t st)
    {
        if (!on) return;
        open_a = get();
        open_cls = cls;
        cudaEventRecord(open_a, st);
    }
    void end(cudaStream_t st)
    {
        if (!on || open_cls < 0) return;
        cudaEvent_t b = get();
        cudaEventRecord(b, st);
        spans.push_back({open_cls, open_a, b, st});
        open_cls = -1;
    }
    // sums elapsed ms / launches per class and recycles the events; call after a stream synchronise
    void collect(double *ms, unsigned long long *launches, int n_cls)
    {
        if (getenv("TACULT_TRACE") && spans.size() >= 64) {
            // start / duration of the last launches on each stream, relative to a common origin (overlap picture)
            const size_t first = spans.size() - 64;
            for (size_t i = first; i < spans.size(); ++i) {
                float t0 = 0.f, dt = 0.f;
                cudaEventElapsedTime(&t0, spans[first].a, spans[i].a);
                cudaEventElapsedTime(&dt, spans[i].a, spans[i].b);
                fprintf(stderr, "[trace] stream %p class %d start %8.1f us dur %6.1f us\n", (void *)spans[i].st, spans[i].cls,
                        t0 * 1e3f, dt * 1e3f);
            }
        }
        for (auto &s : spans) {
            float t = 0.f;
            cudaEventElapsedTime(&t, s.a, s.b);
            if (s.cls < n_cls) { ms[s.cls] += t; launches[s.cls] += 1; }
            pool.push_back(s.a);
            pool.push_back(s.b);
        }
        spans.clear();
    }
    ~Profiler()
    {
        for (auto &s : spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
        for (auto e : pool) cudaEventDestroy(e);
    }
};

}  // namespace tc
