// mimi_rvq.cuh — Mimi's split residual vector quantiser on the GPU: the codes <-> latent boundary either side of the LM step
// (first slice).  Replaces the ggml graphs of
//   mimi_quantizer_encode / mimi_decode_latent          src/moshi/models/compression.h:93-99, 216-222
//   moshi_split_rvq_encode / _decode, moshi_rvq_*       src/moshi/quantization/vq.h:18-117
//   moshi_residual_vq_*, moshi_EuclideanCodebook_*      src/moshi/quantization/core_vq.h:14-193
// Index work: the codes must be bit-exact, so the distance of a latent to a centroid keeps ggml's arithmetic and ORDER:
// f32 (b - a), f32 square, summed over the 256 dimensions in index order in double, rounded once, r = 1 / (c + 1), first maximum.
// One thread owns one centroid (sequential sum); the codebooks are stored transposed [D][bins] so that the threads of a warp read
// consecutive addresses.  The 1 x 1 convolutions run on cond_linear_kernel (tts_kernels.cuh: F16 weights, activation rounded to
// F16 like ggml's im2col, exact products summed in double).
#pragma once
#include "common.cuh"

namespace msx {
namespace rvq {

constexpr int kNearestThreads = 256;

// key of (r, j): larger r wins, ties -> smaller j.  r = 1 / (c + 1) is in (0, 1]: its bit pattern is monotone.
__device__ __forceinline__ unsigned long long nearest_key(float r, int j) {
    return ((unsigned long long)__float_as_uint(r) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)j);
}

// grid (bins / 256, T): every thread one centroid of frame t; block maximum -> atomicMax(keys[t])
__global__ void __launch_bounds__(kNearestThreads) nearest_kernel(const float *cb_t /*[D][bins]*/, int bins, int D, const float *x /*[T][D]*/,
                                                                   unsigned long long *keys /*[T]*/) {
    extern __shared__ float a_s[];                     // [D] the frame's residual
    const int t = blockIdx.y, j = blockIdx.x * kNearestThreads + threadIdx.x;
    for (int d = threadIdx.x; d < D; d += kNearestThreads) a_s[d] = x[(size_t)t * D + d];
    __syncthreads();
    unsigned long long key = 0ull;
    if (j < bins) {
        double s = 0.0;
#pragma unroll 8
        for (int d = 0; d < D; d++) {
            const float diff = __fsub_rn(cb_t[(size_t)d * bins + j], a_s[d]);
            s += (double)__fmul_rn(diff, diff);
        }
        const float c = (float)s;
        key = nearest_key(__fdiv_rn(1.0f, __fadd_rn(c, 1.0f)), j);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o); key = k2 > key ? k2 : key; }
    __shared__ unsigned long long sb[kNearestThreads / 32];
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kNearestThreads / 32; w++) key = sb[w] > key ? sb[w] : key;
        atomicMax(keys + t, key);
    }
}
// code of frame t = index in its key; residual -= centroid; key reset for the next layer.  grid (T), D threads
__global__ void apply_kernel(const float *cb /*[bins][D]*/, int D, float *x /*[T][D]*/, unsigned long long *keys, int32_t *codes /*[T]*/) {
    const int t = blockIdx.x;
    const int j = (int)(0xffffffffu - (uint32_t)(keys[t] & 0xffffffffull));
    for (int d = threadIdx.x; d < D; d += blockDim.x) x[(size_t)t * D + d] = __fsub_rn(x[(size_t)t * D + d], cb[(size_t)j * D + d]);
    __syncthreads();
    if (threadIdx.x == 0) { codes[t] = j; keys[t] = 0ull; }
}
// out[t][d] = sum over layers q < n_q of cb[q][codes[q][t]][d], added in layer order.  One thread per (t, d)
__global__ void decode_kernel(const float *cb /*[n][bins][D]*/, int n_q, int bins, int D, const int32_t *codes /*[n_q][T]*/, int T, float *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)T * D) return;
    const int t = (int)(i / D), d = (int)(i % D);
    float s = 0.f;
    for (int q = 0; q < n_q; q++) {
        const int j = min(max(codes[(size_t)q * T + t], 0), bins - 1);
        const float v = cb[((size_t)q * bins + j) * D + d];
        s = q == 0 ? v : __fadd_rn(s, v);
    }
    out[i] = s;
}
// [n][bins][D] -> [n][D][bins]
__global__ void transpose_kernel(const float *src, float *dst, int bins, int D, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long per = (long long)bins * D, q = i / per, r = i % per;
    const int j = (int)(r / D), d = (int)(r % D);
    dst[q * per + (long long)d * bins + j] = src[i];
}

}  // namespace rvq
}  // namespace msx
