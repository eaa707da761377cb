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

// second formulation (nn_fused_dx.cu): one A fetch feeds the three horizontal taps, 3C accumulator columns per tile
bool fused_dx_plan(int C, int &G, int &T);
void fused_dx_pack_weights(const std::vector<std::vector<float>> &conv_w, int C, std::vector<__nv_bfloat16> &out);
void fused_dx_pack_stem(const std::vector<float> &stem_w, int C, int in_planes, std::vector<__nv_bfloat16> &out);
void fused_dx_bias_tiles(const std::vector<float> &bias_all, int C, std::vector<__nv_bfloat16> &out);
int launch_trunk_dx(int C, const CUtensorMap &w_stem, const CUtensorMap &w_conv, const CUtensorMap &bias_tiles, int G, int T,
                    int n_conv, int in_planes, int batch, const int *count_dev, const uint32_t *states, const float *obs,
                    __nv_bfloat16 *out_flat, int sm_count, cudaStream_t st);

// third formulation (nn_fused_p16.cu): dx taps stacked in N on a pitch-16 layout; shares the weight / bias packing of the
// second one (fused_dx_pack_*, fused_dx_bias_tiles)
bool fused_p16_plan(int C, int &G, int &T);
size_t fused_p16_smem_bytes(int C, int T);
int launch_trunk_p16(int C, const CUtensorMap &w_stem, const CUtensorMap &w_conv, const CUtensorMap &bias_tiles, int G, int T,
                     int n_conv, int in_planes, int batch, const int *count_dev, const uint32_t *states, const float *obs,
                     __nv_bfloat16 *out_flat, int sm_count, cudaStream_t st);

}  // namespace tc
