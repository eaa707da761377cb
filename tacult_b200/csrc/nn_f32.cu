// fp32 CUDA-core forward of the dual-head ResNet ("parity mode": within 1e-3 of the reference
// PyTorch model on the same weights).  Graph: /root/reference/src/tacult/network.py:81-102,
// residual block network.py:35-41.  BatchNorm (eval mode, eps 1e-5) is folded on the host.
//
// Activations are NHWC fp32 [row][81][C]; every kernel takes the live row count from device
// memory so the search loop never synchronises with the host.
#include <math.h>

#include "nn.cuh"
#include "rules.cuh"

namespace tc {

// ------------------------------------------------------------------------------------------------
// host: state_dict blob -> folded parameters
// ------------------------------------------------------------------------------------------------
size_t blob_floats(const tc_net_desc &d)
{
    size_t C = d.channels, B = d.num_residual_blocks, E = d.embedding_size, I = d.in_planes;
    size_t n = C * I * 9 + C + 4 * C;
    n += 2 * B * (C * C * 9 + C + 4 * C);
    n += E * 81 * C + E + 4 * E;
    n += 81 * E + 81 + E + 1;
    return n;
}

namespace {
struct Cursor {
    const float *p;
    const float *take(size_t n)
    {
        const float *r = p;
        p += n;
        return r;
    }
};

// w' = w * g/sqrt(var+eps), b' = (b-mean)*g/sqrt(var+eps)+beta   (per output channel)
void fold_bn(const float *w, const float *b, const float *gamma, const float *beta, const float *mean,
             const float *var, size_t n_out, size_t per_out, std::vector<float> &wf, std::vector<float> &bf)
{
    wf.resize(n_out * per_out);
    bf.resize(n_out);
    for (size_t o = 0; o < n_out; ++o) {
        double s = (double)gamma[o] / sqrt((double)var[o] + 1e-5);
        for (size_t k = 0; k < per_out; ++k) wf[o * per_out + k] = (float)((double)w[o * per_out + k] * s);
        bf[o] = (float)(((double)b[o] - (double)mean[o]) * s + (double)beta[o]);
    }
}
}  // namespace

int fold_blob(const tc_net_desc &d, const float *blob, size_t n, FoldedNet &out)
{
    TC_CHECK(d.channels >= 4 && d.channels % 4 == 0 && d.channels <= 512, TC_EINVAL,
             "channels must be a multiple of 4 in [4,512], got %d", d.channels);
    TC_CHECK(d.num_residual_blocks >= 0 && d.num_residual_blocks <= 64, TC_EINVAL, "bad num_residual_blocks %d",
             d.num_residual_blocks);
    TC_CHECK(d.embedding_size >= 4 && d.embedding_size % 4 == 0 && d.embedding_size <= 1024, TC_EINVAL,
             "embedding_size must be a multiple of 4 in [4,1024], got %d", d.embedding_size);
    TC_CHECK(d.in_planes >= 1 && d.in_planes <= 32, TC_EINVAL, "bad in_planes %d", d.in_planes);
    TC_CHECK(n == blob_floats(d), TC_EINVAL, "weight blob has %zu floats, expected %zu", n, blob_floats(d));
    const size_t C = d.channels, E = d.embedding_size, I = d.in_planes;
    Cursor c{blob};
    out.d = d;
    {
        const float *w = c.take(C * I * 9), *b = c.take(C), *g = c.take(C), *be = c.take(C), *m = c.take(C),
                    *v = c.take(C);
        fold_bn(w, b, g, be, m, v, C, I * 9, out.stem_w, out.stem_b);
    }
    out.conv_w.assign(2 * d.num_residual_blocks, {});
    out.conv_b.assign(2 * d.num_residual_blocks, {});
    for (int l = 0; l < 2 * d.num_residual_blocks; ++l) {
        const float *w = c.take(C * C * 9), *b = c.take(C), *g = c.take(C), *be = c.take(C), *m = c.take(C),
                    *v = c.take(C);
        fold_bn(w, b, g, be, m, v, C, C * 9, out.conv_w[l], out.conv_b[l]);
    }
    {
        const float *w = c.take(E * 81 * C), *b = c.take(E), *g = c.take(E), *be = c.take(E), *m = c.take(E),
                    *v = c.take(E);
        fold_bn(w, b, g, be, m, v, E, 81 * C, out.fc1_w, out.fc1_b);
    }
    out.pi_w.assign(c.p, c.p + 81 * E); c.take(81 * E);
    out.pi_b.assign(c.p, c.p + 81);     c.take(81);
    out.v_w.assign(c.p, c.p + E);       c.take(E);
    out.v_b.assign(c.p, c.p + 1);       c.take(1);
    return TC_OK;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int live_rows(const int *count_dev, int batch)
{
    int n = batch;
    if (count_dev) n = min(n, *count_dev);
    return n;
}

// conv_in (in_planes -> C, 3x3, pad 1) + folded BN + ReLU; one block per position.
// Input is either dense planes [row][in][81] or a packed canonical position (planes built here:
// stones of the side to move, opponent stones, legal cells, ones — see oracle/utac_rules.h).
__global__ void __launch_bounds__(256) k_stem_f32(int batch, const int *__restrict__ count_dev,
                                                  const uint32_t *__restrict__ states, const float *__restrict__ obs,
                                                  int in_planes, int C, const float *__restrict__ w,
                                                  const float *__restrict__ bias, float *__restrict__ out)
{
    extern __shared__ float s_in[];        // [in_planes][121], zero border
    const int row = blockIdx.x;
    if (row >= live_rows(count_dev, batch)) return;
    for (int i = threadIdx.x; i < in_planes * 121; i += blockDim.x) s_in[i] = 0.f;
    __syncthreads();
    if (states) {
        if (threadIdx.x < 81) {
            uint32_t wd[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) wd[k] = states[8 * (size_t)row + k];
            Pos p = unpack(wd);
            Summary s = summarize(p);
            uint32_t lw[3];
            legal_words(p, s, lw);
            int a = threadIdx.x, y = a / 9, x = a % 9, cell = (y + 1) * 11 + x + 1;
            bool mine = action_bit(p.mine, a), occ = action_bit(p.occ, a);
            s_in[cell] = mine ? 1.f : 0.f;
            s_in[121 + cell] = (occ && !mine) ? 1.f : 0.f;
            s_in[242 + cell] = action_bit(lw, a) ? 1.f : 0.f;
            s_in[363 + cell] = 1.f;
        }
    } else {
        for (int i = threadIdx.x; i < in_planes * 81; i += blockDim.x) {
            int pl = i / 81, a = i % 81;
            s_in[pl * 121 + (a / 9 + 1) * 11 + a % 9 + 1] = obs[(size_t)row * in_planes * 81 + i];
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 81 * C; idx += blockDim.x) {
        int p = idx / C, c = idx % C;
        int base = (p / 9) * 11 + p % 9;           // top-left of the 3x3 window in the padded plane
        float acc = bias[c];
        for (int pl = 0; pl < in_planes; ++pl) {
#pragma unroll
            for (int t = 0; t < 9; ++t)
                acc = fmaf(s_in[pl * 121 + base + (t / 3) * 11 + t % 3], w[(t * in_planes + pl) * C + c], acc);
        }
        out[(size_t)row * 81 * C + idx] = fmaxf(acc, 0.f);
    }
}

// 3x3 conv C->C + folded BN (+ residual) + ReLU.  A block owns CONV_G positions x CT output channels and
// walks the input channels in chunks of CK staged in shared memory with a zero halo.
constexpr int CONV_G = 4;
constexpr int CONV_THREADS = 256;
constexpr int CONV_PPT = 11;                 // board cells per thread, upper bound

__global__ void __launch_bounds__(CONV_THREADS) k_conv3x3_f32(int batch, const int *__restrict__ count_dev, int C,
                                                              int CK, int CT, const float *__restrict__ in,
                                                              const float *__restrict__ w,
                                                              const float *__restrict__ bias,
                                                              const float *__restrict__ residual,
                                                              float *__restrict__ out)
{
    extern __shared__ float smem[];
    const int n_rows = live_rows(count_dev, batch);
    const int row0 = blockIdx.x * CONV_G;
    if (row0 >= n_rows) return;
    const int cout0 = blockIdx.y * CT;
    const int stride = CK + 1;
    float *s_in = smem;                                   // [CONV_G*121][CK+1]
    float *s_w = smem + CONV_G * 121 * stride;            // [9][CK][CT]
    const int tq = CT / 4, cq = threadIdx.x % tq, ps = threadIdx.x / tq, nslots = CONV_THREADS / tq;
    const int cells = CONV_G * 81;

    int cellbase[CONV_PPT];
#pragma unroll
    for (int i = 0; i < CONV_PPT; ++i) {
        int pos = ps + i * nslots;
        int b = pos / 81, p = pos % 81;
        cellbase[i] = pos < cells ? (b * 121 + (p / 9) * 11 + p % 9) * stride : 0;
    }
    float acc[CONV_PPT][4];
#pragma unroll
    for (int i = 0; i < CONV_PPT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    for (int i = threadIdx.x; i < CONV_G * 121 * stride; i += CONV_THREADS) s_in[i] = 0.f;

    for (int c0 = 0; c0 < C; c0 += CK) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < cells * CK; idx += CONV_THREADS) {
            int bp = idx / CK, ci = idx % CK;
            int b = bp / 81, p = bp % 81;
            float v = 0.f;
            if (row0 + b < n_rows) v = in[((size_t)(row0 + b) * 81 + p) * C + c0 + ci];
            s_in[(b * 121 + (p / 9 + 1) * 11 + p % 9 + 1) * stride + ci] = v;
        }
        for (int idx = threadIdx.x; idx < 9 * CK * CT; idx += CONV_THREADS) {
            int co = idx % CT, ci = (idx / CT) % CK, t = idx / (CT * CK);
            s_w[idx] = w[((size_t)t * C + c0 + ci) * C + cout0 + co];
        }
        __syncthreads();
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
            const int toff = ((t / 3) * 11 + t % 3) * stride;
            const float *wrow = s_w + (size_t)t * CK * CT + 4 * cq;
#pragma unroll 4
            for (int ci = 0; ci < CK; ++ci) {
                float4 wv = *reinterpret_cast<const float4 *>(wrow + ci * CT);
#pragma unroll
                for (int i = 0; i < CONV_PPT; ++i) {
                    float a = s_in[cellbase[i] + toff + ci];
                    acc[i][0] = fmaf(a, wv.x, acc[i][0]);
                    acc[i][1] = fmaf(a, wv.y, acc[i][1]);
                    acc[i][2] = fmaf(a, wv.z, acc[i][2]);
                    acc[i][3] = fmaf(a, wv.w, acc[i][3]);
                }
            }
        }
    }
    const float4 bv = *reinterpret_cast<const float4 *>(bias + cout0 + 4 * cq);
#pragma unroll
    for (int i = 0; i < CONV_PPT; ++i) {
        int pos = ps + i * nslots;
        if (pos >= cells) continue;
        int b = pos / 81, p = pos % 81;
        if (row0 + b >= n_rows) continue;
        size_t o = ((size_t)(row0 + b) * 81 + p) * C + cout0 + 4 * cq;
        float4 r = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
        if (residual) {
            float4 rs = *reinterpret_cast<const float4 *>(residual + o);
            r.x += rs.x; r.y += rs.y; r.z += rs.z; r.w += rs.w;
        }
        r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
        *reinterpret_cast<float4 *>(out + o) = r;
    }
}

// out[M][N] = relu(A[M][K] * W[K][N] + bias): fc1 + folded BatchNorm1d + ReLU (network.py:95)
__global__ void __launch_bounds__(256) k_linear_relu_f32(int batch, const int *__restrict__ count_dev, int K, int N,
                                                         const float *__restrict__ A, const float *__restrict__ W,
                                                         const float *__restrict__ bias, float *__restrict__ out)
{
    __shared__ __align__(16) float As[16][64 + 4];
    __shared__ __align__(16) float Bs[16][64];
    const int M = live_rows(count_dev, batch);
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    if (m0 >= M) return;
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int a_row = t / 4, a_k = (t % 4) * 4;
    const int b_k = t / 16, b_n = (t % 16) * 4;
    for (int k0 = 0; k0 < K; k0 += 16) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + a_row < M && k0 + a_k < K)
            av = *reinterpret_cast<const float4 *>(A + (size_t)(m0 + a_row) * K + k0 + a_k);
        if (k0 + b_k < K && n0 + b_n < N) bv = *reinterpret_cast<const float4 *>(W + (size_t)(k0 + b_k) * N + n0 + b_n);
        __syncthreads();
        As[a_k + 0][a_row] = av.x; As[a_k + 1][a_row] = av.y; As[a_k + 2][a_row] = av.z; As[a_k + 3][a_row] = av.w;
        *reinterpret_cast<float4 *>(&Bs[b_k][b_n]) = bv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n < N) out[(size_t)m * N + n] = fmaxf(acc[i][j] + bias[n], 0.f);
        }
    }
}

// policy logits + softmax (unmasked, network.py:99) and tanh value (network.py:100); one warp per row
__global__ void __launch_bounds__(256) k_heads_f32(int batch, const int *__restrict__ count_dev, int E,
                                                   const float *__restrict__ emb, const float *__restrict__ W,
                                                   const float *__restrict__ bias, float *__restrict__ pi,
                                                   float *__restrict__ v, float *__restrict__ logits)
{
    extern __shared__ float s_emb[];       // [8][E]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    const int M = live_rows(count_dev, batch);
    if (row >= M) return;                  // whole warp leaves together; no block barrier below
    float *e = s_emb + warp * E;
    for (int k = lane; k < E; k += 32) e[k] = emb[(size_t)row * E + k];
    __syncwarp();
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    const bool has2 = lane + 64 < 82;
    for (int k = 0; k < E; ++k) {
        float h = e[k];
        const float *wr = W + (size_t)k * 82;
        acc0 = fmaf(h, wr[lane], acc0);
        acc1 = fmaf(h, wr[lane + 32], acc1);
        if (has2) acc2 = fmaf(h, wr[lane + 64], acc2);
    }
    acc0 += bias[lane];
    acc1 += bias[lane + 32];
    if (has2) acc2 += bias[lane + 64];
    // lane 17 of the third group is the value head (column 81)
    const bool pol2 = lane + 64 < 81;
    float m = fmaxf(acc0, acc1);
    if (pol2) m = fmaxf(m, acc2);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
    float e0 = expf(acc0 - m), e1 = expf(acc1 - m), e2 = pol2 ? expf(acc2 - m) : 0.f;
    float s = e0 + e1 + e2;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    float inv = 1.f / s;
    float *pr = pi + (size_t)row * 81;
    pr[lane] = e0 * inv;
    pr[lane + 32] = e1 * inv;
    if (pol2) pr[lane + 64] = e2 * inv;
    if (logits) {
        float *lr = logits + (size_t)row * 81;
        lr[lane] = acc0;
        lr[lane + 32] = acc1;
        if (pol2) lr[lane + 64] = acc2;
    }
    if (lane == 17) v[row] = tanhf(acc2);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int upload_vec(DevBuf<float> &dst, const std::vector<float> &src)
{
    TC_CUDA(dst.alloc(src.size()));
    TC_CUDA(cudaMemcpy(dst.p, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice));
    return TC_OK;
}

static size_t conv_smem_bytes(int C)
{
    int CK = C < 32 ? C : 32, CT = C < 32 ? C : 32;
    return sizeof(float) * ((size_t)CONV_G * 121 * (CK + 1) + (size_t)9 * CK * CT);
}

int NetF32::upload(const FoldedNet &f, int cap)
{
    d = f.d;
    capacity = cap;
    const int C = d.channels, E = d.embedding_size, I = d.in_planes;
    std::vector<float> t;
    // stem [C][I][3][3] -> [tap][I][C]
    t.assign((size_t)9 * I * C, 0.f);
    for (int co = 0; co < C; ++co)
        for (int ci = 0; ci < I; ++ci)
            for (int tap = 0; tap < 9; ++tap) t[((size_t)tap * I + ci) * C + co] = f.stem_w[((size_t)co * I + ci) * 9 + tap];
    TC_TRY(upload_vec(stem_w, t));
    TC_TRY(upload_vec(stem_b, f.stem_b));
    conv_w.clear();
    conv_b.clear();
    conv_w.resize(f.conv_w.size());
    conv_b.resize(f.conv_b.size());
    for (size_t l = 0; l < f.conv_w.size(); ++l) {
        t.assign((size_t)9 * C * C, 0.f);
        for (int co = 0; co < C; ++co)
            for (int ci = 0; ci < C; ++ci)
                for (int tap = 0; tap < 9; ++tap)
                    t[((size_t)tap * C + ci) * C + co] = f.conv_w[l][((size_t)co * C + ci) * 9 + tap];
        TC_TRY(upload_vec(conv_w[l], t));
        TC_TRY(upload_vec(conv_b[l], f.conv_b[l]));
    }
    // fc1 [E][c*81+p] -> [p*C+c][E]   (activations are NHWC, the reference flattens NCHW: network.py:92)
    t.assign((size_t)81 * C * E, 0.f);
    for (int e = 0; e < E; ++e)
        for (int c = 0; c < C; ++c)
            for (int p = 0; p < 81; ++p) t[((size_t)p * C + c) * E + e] = f.fc1_w[(size_t)e * 81 * C + (size_t)c * 81 + p];
    TC_TRY(upload_vec(fc1_w, t));
    TC_TRY(upload_vec(fc1_b, f.fc1_b));
    // heads: [E][82]
    t.assign((size_t)E * 82, 0.f);
    for (int k = 0; k < E; ++k) {
        for (int o = 0; o < 81; ++o) t[(size_t)k * 82 + o] = f.pi_w[(size_t)o * E + k];
        t[(size_t)k * 82 + 81] = f.v_w[k];
    }
    TC_TRY(upload_vec(head_w, t));
    std::vector<float> hb(f.pi_b);
    hb.push_back(f.v_b[0]);
    TC_TRY(upload_vec(head_b, hb));
    for (int i = 0; i < 3; ++i) TC_CUDA(act[i].alloc((size_t)cap * 81 * C));
    TC_CUDA(emb.alloc((size_t)cap * E));
    TC_CUDA(cudaFuncSetAttribute(k_conv3x3_f32, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)conv_smem_bytes(C)));
    loaded = true;
    return TC_OK;
}

int NetF32::forward(int batch, const int *count_dev, const uint32_t *states, const float *obs, float *pi, float *v,
                    float *logits, cudaStream_t st, Profiler *prof)
{
    Profiler none;
    if (!prof) prof = &none;
    const int C = d.channels, E = d.embedding_size, I = d.in_planes;
    int launches = 0;
    prof->begin(TC_K_NN_STEM, st);
    k_stem_f32<<<batch, 256, sizeof(float) * I * 121, st>>>(batch, count_dev, states, obs, I, C, stem_w.p, stem_b.p,
                                                            act[0].p);
    prof->end(st);
    ++launches;
    const int CK = C < 32 ? C : 32, CT = CK;
    dim3 cgrid(ceil_div(batch, CONV_G), C / CT);
    const size_t csm = conv_smem_bytes(C);
    float *x = act[0].p, *y = act[1].p, *z = act[2].p;
    for (int b = 0; b < d.num_residual_blocks; ++b) {
        prof->begin(TC_K_NN_CONV, st);
        k_conv3x3_f32<<<cgrid, CONV_THREADS, csm, st>>>(batch, count_dev, C, CK, CT, x, conv_w[2 * b].p,
                                                        conv_b[2 * b].p, nullptr, y);
        prof->end(st);
        prof->begin(TC_K_NN_CONV, st);
        k_conv3x3_f32<<<cgrid, CONV_THREADS, csm, st>>>(batch, count_dev, C, CK, CT, y, conv_w[2 * b + 1].p,
                                                        conv_b[2 * b + 1].p, x, z);
        prof->end(st);
        launches += 2;
        float *tmp = x;
        x = z;
        z = tmp;
    }
    dim3 lgrid(ceil_div(batch, 64), ceil_div(E, 64));
    prof->begin(TC_K_NN_FC1, st);
    k_linear_relu_f32<<<lgrid, 256, 0, st>>>(batch, count_dev, 81 * C, E, x, fc1_w.p, fc1_b.p, emb.p);
    prof->end(st);
    prof->begin(TC_K_NN_HEADS, st);
    k_heads_f32<<<ceil_div(batch, 8), 256, sizeof(float) * 8 * E, st>>>(batch, count_dev, E, emb.p, head_w.p,
                                                                        head_b.p, pi, v, logits);
    prof->end(st);
    launches += 2;
    return launches;
}

void launch_heads_f32(int batch, const int *count_dev, int E, const float *emb, const float *W, const float *bias,
                      float *pi, float *v, float *logits, cudaStream_t st)
{
    k_heads_f32<<<ceil_div(batch, 8), 256, sizeof(float) * 8 * E, st>>>(batch, count_dev, E, emb, W, bias, pi, v, logits);
}

// test hook (tc_debug_trunk): activations after the stem and n_convs convolutions, f32 [batch][81][C]
int NetF32::debug_trunk(int batch, const uint32_t *states, int n_convs, float *out, cudaStream_t st)
{
    const int C = d.channels, I = d.in_planes;
    k_stem_f32<<<batch, 256, sizeof(float) * I * 121, st>>>(batch, nullptr, states, nullptr, I, C, stem_w.p, stem_b.p, act[0].p);
    const int CK = C < 32 ? C : 32, CT = CK;
    dim3 cgrid(ceil_div(batch, CONV_G), C / CT);
    const size_t csm = conv_smem_bytes(C);
    float *x = act[0].p, *y = act[1].p, *z = act[2].p, *cur = x;
    for (int l = 0; l < n_convs && l < 2 * d.num_residual_blocks; ++l) {
        if ((l & 1) == 0) {
            k_conv3x3_f32<<<cgrid, CONV_THREADS, csm, st>>>(batch, nullptr, C, CK, CT, x, conv_w[l].p, conv_b[l].p, nullptr, y);
            cur = y;
        } else {
            k_conv3x3_f32<<<cgrid, CONV_THREADS, csm, st>>>(batch, nullptr, C, CK, CT, y, conv_w[l].p, conv_b[l].p, x, z);
            cur = z;
            float *t = x; x = z; z = t;
        }
    }
    TC_CUDA(cudaMemcpyAsync(out, cur, (size_t)batch * 81 * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return TC_OK;
}

}  // namespace tc
