// Batched dual-head ResNet forward (evaluator).  Replaces `_UtacNNet.forward`
// (/root/reference/src/tacult/network.py:81-102) behind `NNetWrapper.predict`
// (/root/reference/src/tacult/base/nn_wrapper.py:125-136).
#pragma once
#include <vector>

#include "common.cuh"

namespace tc {

// BatchNorm-folded parameters on the host, in the reference's own tensor layouts
struct FoldedNet {
    tc_net_desc d;
    std::vector<float> stem_w, stem_b;                 // [C][in][3][3], [C]
    std::vector<std::vector<float>> conv_w, conv_b;    // 2*blocks x ([C][C][3][3], [C])
    std::vector<float> fc1_w, fc1_b;                   // [E][81*C] (k = c*81 + p), [E]
    std::vector<float> pi_w, pi_b, v_w, v_b;           // [81][E],[81],[1][E],[1]
};

// fp32 CUDA-core path (parity mode)
struct NetF32 {
    tc_net_desc d{};
    int capacity = 0;                                   // max batch rows
    DevBuf<float> stem_w, stem_b;                       // [9][in][C], [C]
    std::vector<DevBuf<float>> conv_w, conv_b;          // [9][C][C] (tap, cin, cout)
    DevBuf<float> fc1_w, fc1_b;                         // [81*C][E] with k = p*C + c
    DevBuf<float> head_w, head_b;                       // [E][82] (81 policy columns then value), [82]
    DevBuf<float> act[3];                               // [cap][81][C]
    DevBuf<float> emb;                                  // [cap][E]
    bool loaded = false;
    int upload(const FoldedNet &f, int capacity);
    // input: packed states (states != null) or dense planes (obs != null); count_dev (optional) holds the
    // live batch size on the device, `batch` is the launch bound.  Returns the number of kernels launched.
    int forward(int batch, const int *count_dev, const uint32_t *states, const float *obs, float *pi, float *v,
                float *logits, cudaStream_t st, Profiler *prof = nullptr);
    int debug_trunk(int batch, const uint32_t *states, int n_convs, float *out, cudaStream_t st);
};

size_t blob_floats(const tc_net_desc &d);
int fold_blob(const tc_net_desc &d, const float *blob, size_t n, FoldedNet &out);

}  // namespace tc
