// bf16 tcgen05/TMEM forward — placeholder until the tensor-core trunk lands (next milestone).
#include "nn_bf16.cuh"

namespace tc {

struct NetBF16Impl {};

NetBF16::~NetBF16() { delete impl; }

int NetBF16::upload(const FoldedNet &, int)
{
    loaded = false;
    why_not = "bf16 tensor-core path not built yet";
    return TC_OK;
}

int NetBF16::forward(int, const int *, const uint32_t *, const float *, float *, float *, float *, cudaStream_t,
                     Profiler *)
{
    set_error("bf16 evaluator: %s", why_not.c_str());
    return TC_ESTATE;
}

}  // namespace tc
