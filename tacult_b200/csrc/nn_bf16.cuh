// bf16 tensor-core forward (production mode, within 2e-2 of the reference net): the conv trunk and fc1 as
// tcgen05/TMEM implicit GEMMs fed by TMA.  See nn_bf16.cu.
#pragma once
#include <string>

#include <cuda.h>
#include <cuda_bf16.h>

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
    // share_sm: other partitions' kernels run next to this instance's (limits fc1 / heads shared memory)
    int upload(const FoldedNet &f, int capacity, bool share_sm = false);
    // same contract as NetF32::forward; returns the number of kernels launched or a negative TC_E* code
    // masked != 0 (packed states only): the policy comes out masked with the leaf's legal moves and renormalised
    int forward(int batch, const int *count_dev, const uint32_t *states, const float *obs, float *pi, float *v,
                float *logits, cudaStream_t st, Profiler *prof = nullptr, int masked = 0);
};

// fc1 with K split over a cluster of two CTAs (nn_fc1.cu); a = activations [rows][K] (box 128 x 64), b = weights [E][K]
// (box 64 x 64), out = bf16 embedding [rows][E] with bias and ReLU applied
size_t fc1_splitk_smem(int stages);
int launch_fc1_splitk(const CUtensorMap &a, const CUtensorMap &b, int batch, const int *count_dev, int K, int E, int stages,
                      const float *bias, __nv_bfloat16 *out, cudaStream_t st);

// fc1 and the policy / value heads in one launch (nn_fc1.cu): emb_map = the bf16 embedding scratch [rows][E] (box 128 x 64),
// head_w = [96][E] (box 96 x 64); `ready` = one zeroed int per 128-row tile; masked != 0 masks the soft-max with the legal
// moves of `states` and renormalises (production mode)
size_t fc1_heads_smem(int stages);
int launch_fc1_heads(const CUtensorMap &a, const CUtensorMap &b, const CUtensorMap &emb_map, const CUtensorMap &head_w,
                     int batch, const int *count_dev, int K, int E, int stages, const float *bias, __nv_bfloat16 *emb,
                     int *ready, const float *head_bias, float *pi, float *v, const uint32_t *states, int masked,
                     cudaStream_t st);

int debug_trunk_bf16(NetBF16 &net, int batch, const uint32_t *states_dev, int n_convs, float *out_dev, cudaStream_t st);

}  // namespace tc
