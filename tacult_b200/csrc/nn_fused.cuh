// Fused shared-memory-resident trunk for narrow nets; see nn_fused.cu.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <vector>

#include "common.cuh"

namespace tc {

// largest tile count T (and positions per group G) that fits shared memory and TMEM for C channels
bool fused_plan(int C, int &G, int &T);
size_t fused_smem_bytes(int C, int T);            // dynamic shared memory of k_trunk_fused for T tiles per group
void fused_bias_tiles(const std::vector<float> &bias_all, std::vector<__nv_bfloat16> &out);
int launch_trunk_fused(int C, const CUtensorMap &w_stem, const CUtensorMap &w_conv, const CUtensorMap &bias_tiles, int G,
                       int T, int n_conv, int in_planes, int batch, const int *count_dev, const uint32_t *states,
                       const float *obs, __nv_bfloat16 *out_flat, int sm_count, cudaStream_t st);

// column-tiled variant for the 32-channel net (nn_trunk_col.cu): half the shared-memory operand traffic per layer
bool trunk_col_supported(int C, int in_planes);
void trunk_col_pack(const float *w, int c_out, int c_in, int k_pad, __nv_bfloat16 *dst);
int launch_trunk_col(const CUtensorMap &w_stem, const CUtensorMap &w_conv, const float *bias_all, int n_conv,
                     int in_planes, int batch, const int *count_dev, const uint32_t *states, const float *obs,
                     __nv_bfloat16 *out_flat, int sm_count, cudaStream_t st);

}  // namespace tc
