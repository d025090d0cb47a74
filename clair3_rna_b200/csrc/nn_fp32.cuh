// Pileup network forward, fp32 on CUDA cores ("nn_impl = 0").
//
// Exact-arithmetic companion of the tensor-core path in nn_tc.cuh: same layer
// decomposition, plain FFMA.  It exists so probabilities can be checked at fp32
// precision on the device itself and so the integer half can be brought up
// independently; the tcgen05 path is the one the bench measures.
//
// Math restated from /root/reference/clair3_rna/model.py:126-216 (Keras LSTM,
// gate order i,f,c,o, sigmoid/tanh; Bidirectional concat; Flatten; Dense+SELU;
// SELU-then-softmax heads) — SURVEY.md Appendix D.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace c3r {

constexpr int U1 = 128, U2 = 160, NT = 33;
constexpr int G1 = 4 * U1, G2 = 4 * U2;          // gate widths 512 / 640
constexpr int H1W = 2 * U1, H2W = 2 * U2;        // 256 / 320
constexpr int L4_IN = NT * H2W;                  // 10560
constexpr int DENSE = 128;

struct NetF32 {                                  // device pointers, Keras layouts
    int C;
    const float* w1;    // [C, 1024]   = [fwd kernel | bwd kernel]
    const float* b1;    // [1024]
    const float* u1;    // [2][128, 512]
    const float* w2;    // [256, 1280]
    const float* b2;    // [1280]
    const float* u2;    // [2][160, 640]
    const float* k4; const float* b4;            // [10560,128]
    const float* k51; const float* b51;          // [128,128]
    const float* k52; const float* b52;
    const float* ky1; const float* by1;          // [128,21]
    const float* ky2; const float* by2;          // [128,3]
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float seluf_(float x) {
    return 1.0507009873554805f * (x > 0.0f ? x : 1.6732632423543772f * expm1f(x));
}

__global__ void k_i32_to_f32(const int32_t* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}

// C[M,N] = act(A[M,K] * B[K,N] + bias[N]);  64x64 tile, 256 threads, 4x4 per thread
template <int ACT>
__global__ void __launch_bounds__(256) k_sgemm(const float* __restrict__ A, const float* __restrict__ B,
                                               const float* __restrict__ bias, float* __restrict__ Cm,
                                               int64_t M, int N, int K) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * 64;
    const int n0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int k0 = 0; k0 < K; k0 += 16) {
        // A tile: 64 rows x 16 k  (thread loads 4 elements)
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int r = e >> 4, kk = e & 15;
            const int64_t m = m0 + r;
            As[kk][r] = (m < M && k0 + kk < K) ? A[m * K + k0 + kk] : 0.0f;
        }
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int kk = e >> 6, c = e & 63;
            Bs[kk][c] = (k0 + kk < K && n0 + c < N) ? B[(int64_t)(k0 + kk) * N + n0 + c] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.0f);
            if (ACT == 1) v = seluf_(v);
            Cm[m * N + n] = v;
        }
    }
}

// One LSTM time step for both directions.
//   zx  [n, 33, 2*4u]   input projection + bias, (dir, gate, unit) along the last axis
//   U   [2][u, 4u]      recurrent kernels
//   hprev/hnext [2][n, u], c [2][n, u]
//   hout [n, 33, 2u]    layer output, forward time order
// block = 32 units x 8 site-groups, each thread 4 sites; grid (ceil(n/32), u/32, 2)
template <int U>
__global__ void __launch_bounds__(256) k_lstm_step(const float* __restrict__ zx, const float* __restrict__ Uk,
                                                   const float* __restrict__ hprev, float* __restrict__ hnext,
                                                   float* __restrict__ c, float* __restrict__ hout,
                                                   int64_t n, int step) {
    __shared__ float hs[32][U + 1];
    const int dir = blockIdx.z;
    const int t = dir == 0 ? step : NT - 1 - step;
    const int unit = blockIdx.y * 32 + (threadIdx.x & 31);
    const int sg = threadIdx.x >> 5;                         // 0..7
    const int64_t s0 = (int64_t)blockIdx.x * 32;
    const float* Ud = Uk + (int64_t)dir * U * 4 * U;
    const float* hp = hprev + (int64_t)dir * n * U;
    for (int e = threadIdx.x; e < 32 * U; e += 256) {
        const int r = e / U, k = e % U;
        hs[r][k] = (step > 0 && s0 + r < n) ? hp[(s0 + r) * U + k] : 0.0f;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[i][g] = 0.0f;
    if (step > 0) {
        for (int k = 0; k < U; ++k) {
            float w[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) w[g] = Ud[(int64_t)k * 4 * U + g * U + unit];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float h = hs[sg * 4 + i][k];
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[i][g] = fmaf(h, w[g], acc[i][g]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t s = s0 + sg * 4 + i;
        if (s >= n) continue;
        const float* z = zx + (s * NT + t) * (8 * U) + dir * 4 * U;
        const float zi = acc[i][0] + z[unit], zf = acc[i][1] + z[U + unit];
        const float zg = acc[i][2] + z[2 * U + unit], zo = acc[i][3] + z[3 * U + unit];
        const int64_t ci = (int64_t)dir * n * U + s * U + unit;
        const float cp = step > 0 ? c[ci] : 0.0f;
        const float cn = sigmoidf_(zf) * cp + sigmoidf_(zi) * tanhf(zg);
        const float h = sigmoidf_(zo) * tanhf(cn);
        c[ci] = cn;
        hnext[ci] = h;
        hout[(s * NT + t) * (2 * U) + dir * U + unit] = h;
    }
}

// L5_1/L5_2 + the two SELU heads + softmax.  Persistent blocks of 256 threads keep both 128x128 kernels in shared
// memory (128 KB, read from L2 once per block instead of once per 16 sites); each half of the block (one thread per
// hidden unit) takes HS sites at a time.
constexpr int HS = 16;
constexpr int HEADS_THREADS = 256;
constexpr int HEADS_SMEM = (2 * DENSE * DENSE + 3 * 2 * HS * DENSE + 2 * HS * 24 + DENSE * 24) * 4;
// l4: SELU'd L4 activations [n][128] (n_part == 0), or n_part planes of raw partial sums `stride` floats apart
// (split-K GEMM): then x = selu(sum + b4) is formed here and, if l4_out is given, stored for inspection.
__global__ void __launch_bounds__(HEADS_THREADS) k_heads(NetF32 w, const float* __restrict__ l4, int n_part, size_t stride,
                                                         float* __restrict__ l4_out, float* __restrict__ probs, int64_t n) {
    extern __shared__ __align__(16) float wsm[];             // [2][DENSE][DENSE]: k51, k52, then the activations
    float (*x)[HS][DENSE] = (float (*)[HS][DENSE])(wsm + 2 * DENSE * DENSE);
    float (*a1)[HS][DENSE] = x + 2;
    float (*a2)[HS][DENSE] = x + 4;
    float (*y)[HS][24] = (float (*)[HS][24])(wsm + 2 * DENSE * DENSE + 3 * 2 * HS * DENSE);
    const int hf = threadIdx.x >> 7, j = threadIdx.x & 127;
    for (int i = threadIdx.x; i < DENSE * DENSE / 4; i += HEADS_THREADS) {
        ((float4*)wsm)[i] = ((const float4*)w.k51)[i];
        ((float4*)wsm)[DENSE * DENSE / 4 + i] = ((const float4*)w.k52)[i];
    }
    const float* w1s = wsm;
    const float* w2s = wsm + DENSE * DENSE;
    const float b1 = w.b51[j], b2 = w.b52[j];
    const float b4j = n_part > 0 ? w.b4[j] : 0.0f;
    // the two small head kernels [128][21] and [128][3] side by side as [128][24]
    float (*wy)[24] = (float (*)[24])(wsm + 2 * DENSE * DENSE + 3 * 2 * HS * DENSE + 2 * HS * 24);
    for (int i = threadIdx.x; i < DENSE * 24; i += HEADS_THREADS) {
        const int k = i / 24, o = i % 24;
        wy[k][o] = o < 21 ? w.ky1[k * 21 + o] : w.ky2[k * 3 + (o - 21)];
    }
    __syncthreads();
    for (int64_t g0 = (int64_t)blockIdx.x * 2 * HS; g0 < n; g0 += (int64_t)gridDim.x * 2 * HS) {
        const int64_t s0 = g0 + hf * HS;
        const int ns = (int)(n - s0 < HS ? (n - s0 < 0 ? 0 : n - s0) : HS);
        if (n_part > 0) {
            for (int i = 0; i < HS; ++i) {
                float v = 0.0f;
                if (i < ns) {
                    v = b4j;
                    for (int q = 0; q < n_part; ++q) v += l4[(size_t)q * stride + (s0 + i) * DENSE + j];
                    v = seluf_(v);
                    if (l4_out) l4_out[(s0 + i) * DENSE + j] = v;
                }
                x[hf][i][j] = v;
            }
        } else {
            for (int i = 0; i < HS; ++i) x[hf][i][j] = i < ns ? l4[(s0 + i) * DENSE + j] : 0.0f;
        }
        __syncthreads();
        float s1[HS], s2[HS];
#pragma unroll
        for (int i = 0; i < HS; ++i) { s1[i] = b1; s2[i] = b2; }
        for (int k = 0; k < DENSE; k += 4) {             // 4 k per step: one 16-byte shared load feeds 8 FMAs
            float w1[4], w2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { w1[q] = w1s[(k + q) * DENSE + j]; w2[q] = w2s[(k + q) * DENSE + j]; }
#pragma unroll
            for (int i = 0; i < HS; ++i) {
                const float4 xv = *(const float4*)&x[hf][i][k];
                s1[i] = fmaf(xv.x, w1[0], s1[i]); s2[i] = fmaf(xv.x, w2[0], s2[i]);
                s1[i] = fmaf(xv.y, w1[1], s1[i]); s2[i] = fmaf(xv.y, w2[1], s2[i]);
                s1[i] = fmaf(xv.z, w1[2], s1[i]); s2[i] = fmaf(xv.z, w2[2], s2[i]);
                s1[i] = fmaf(xv.w, w1[3], s1[i]); s2[i] = fmaf(xv.w, w2[3], s2[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < HS; ++i) { a1[hf][i][j] = seluf_(s1[i]); a2[hf][i][j] = seluf_(s2[i]); }
        __syncthreads();
        // 24 outputs x HS sites = 384 dot products over the half's 128 threads
        for (int e = j; e < 24 * HS; e += 128) {
            const int i = e / 24, o = e % 24;
            const float* av = o < 21 ? a1[hf][i] : a2[hf][i];
            float v0 = o < 21 ? w.by1[o] : w.by2[o - 21], v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;   // four chains in flight
            for (int k = 0; k < DENSE; k += 4) {
                const float4 a4 = *(const float4*)&av[k];
                v0 = fmaf(a4.x, wy[k][o], v0); v1 = fmaf(a4.y, wy[k + 1][o], v1);
                v2 = fmaf(a4.z, wy[k + 2][o], v2); v3 = fmaf(a4.w, wy[k + 3][o], v3);
            }
            y[hf][i][o] = seluf_((v0 + v1) + (v2 + v3));
        }
        __syncthreads();
        for (int e = j; e < 24 * HS; e += 128) {
            const int i = e / 24, o = e % 24;
            if (i >= ns) continue;
            const int lo = o < 21 ? 0 : 21, hi = o < 21 ? 21 : 24;
            float mx = y[hf][i][lo];
            for (int k = lo + 1; k < hi; ++k) mx = fmaxf(mx, y[hf][i][k]);
            float sum = 0.0f;
            for (int k = lo; k < hi; ++k) sum += expf(y[hf][i][k] - mx);
            probs[(s0 + i) * 24 + o] = expf(y[hf][i][o] - mx) / sum;
        }
        __syncthreads();
    }
}
inline cudaError_t launch_heads(const NetF32& w, const float* l4, int n_part, size_t stride, float* l4_out, float* probs,
                                int64_t n, int sm_count, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, HEADS_SMEM);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    if (n <= 0) return cudaSuccess;
    int64_t groups = (n + 2 * HS - 1) / (2 * HS);
    const unsigned grid = (unsigned)(groups < sm_count ? groups : sm_count);
    k_heads<<<grid, HEADS_THREADS, HEADS_SMEM, st>>>(w, l4, n_part, stride, l4_out, probs, n);
    return cudaGetLastError();
}

struct NetF32Scratch {
    float *x, *zx1, *h1, *zx2, *h2, *l4, *hbuf, *cbuf;      // sized for `cap` sites
    int64_t cap;
};

inline size_t netf32_scratch_bytes(int64_t cap, int C) {
    size_t f = (size_t)cap * NT * C + (size_t)cap * NT * 2 * G1 + (size_t)cap * NT * H1W + (size_t)cap * NT * 2 * G2 +
               (size_t)cap * NT * H2W + (size_t)cap * DENSE + (size_t)cap * 2 * U2 * 2 + (size_t)cap * 2 * U2;
    return f * sizeof(float) + 4096;
}

// forward of n <= scratch.cap sites; tensor int32 [n,33,C] on the device -> probs [n,24].  returns launches.
inline int netf32_forward(const NetF32& w, const NetF32Scratch& s, const int32_t* tensor, int64_t n, float* probs,
                          cudaStream_t st) {
    if (n <= 0) return 0;
    int launches = 0;
    const int C = w.C;
    const int64_t rows = n * NT;
    k_i32_to_f32<<<(unsigned)((rows * C + 1023) / 1024 < 4096 ? (rows * C + 1023) / 1024 : 4096), 256, 0, st>>>(tensor, s.x, rows * C);
    ++launches;
    dim3 g1((2 * G1 + 63) / 64, (unsigned)((rows + 63) / 64));
    k_sgemm<0><<<g1, 256, 0, st>>>(s.x, w.w1, w.b1, s.zx1, rows, 2 * G1, C);
    ++launches;
    float* hA = s.hbuf;
    float* hB = s.hbuf + (size_t)s.cap * 2 * U2;
    for (int step = 0; step < NT; ++step) {
        dim3 g((unsigned)((n + 31) / 32), U1 / 32, 2);
        k_lstm_step<U1><<<g, 256, 0, st>>>(s.zx1, w.u1, hA, hB, s.cbuf, s.h1, n, step);
        ++launches;
        float* tmp = hA; hA = hB; hB = tmp;
    }
    dim3 g2((2 * G2 + 63) / 64, (unsigned)((rows + 63) / 64));
    k_sgemm<0><<<g2, 256, 0, st>>>(s.h1, w.w2, w.b2, s.zx2, rows, 2 * G2, H1W);
    ++launches;
    for (int step = 0; step < NT; ++step) {
        dim3 g((unsigned)((n + 31) / 32), U2 / 32, 2);
        k_lstm_step<U2><<<g, 256, 0, st>>>(s.zx2, w.u2, hA, hB, s.cbuf, s.h2, n, step);
        ++launches;
        float* tmp = hA; hA = hB; hB = tmp;
    }
    dim3 g4((DENSE + 63) / 64, (unsigned)((n + 63) / 64));
    k_sgemm<1><<<g4, 256, 0, st>>>(s.h2, w.k4, w.b4, s.l4, n, DENSE, L4_IN);
    ++launches;
    launch_heads(w, s.l4, 0, 0, nullptr, probs, n, 148, st);
    ++launches;
    return launches;
}

}  // namespace c3r
