// bf16 tensor-core forward (production mode, within 2e-2 of the reference net): the conv trunk and fc1 as
// tcgen05/TMEM implicit GEMMs fed by TMA.  See nn_bf16.cu.
#pragma once
#include <string>

#include "nn.cuh"

namespace tc {

struct NetBF16Impl;

struct NetBF16 {
    bool loaded = false;
    std::string why_not = "tc_load_weights has not been called";
    NetBF16Impl *impl = nullptr;
    NetBF16() = default;
    NetBF16(const NetBF16 &) = delete;
    NetBF16 &operator=(const NetBF16 &) = delete;
    ~NetBF16();
    // returns TC_OK and sets `loaded`, or leaves loaded == false with the reason in why_not
    int upload(const FoldedNet &f, int capacity);
    // same contract as NetF32::forward; returns the number of kernels launched or a negative TC_E* code
    int forward(int batch, const int *count_dev, const uint32_t *states, const float *obs, float *pi, float *v,
                float *logits, cudaStream_t st, Profiler *prof = nullptr);
};

int debug_trunk_bf16(NetBF16 &net, int batch, const uint32_t *states_dev, int n_convs, float *out_dev, cudaStream_t st);

}  // namespace tc
