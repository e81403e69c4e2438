// sample.cuh — temperature / top-k sampling on the device (sm_100a).
//
// Replaces the graph the reference builds in moshi_sample_token (src/moshi/utils/sampling.h:46-64) for
// use_sampling && temp > 0:
//     ggml_scale(logits, 1/temp) -> ggml_soft_max -> ggml_argsort_top_k(k) -> get_rows (the k probabilities)
//     -> moshi_multinomial: q_j = p_j / e_j, e ~ Exp(1) drawn ON THE HOST with libc rand()
//        (src/context.h:464-480) -> ggml_argmax (first maximum) -> get_rows(indices)
// The Exp(1) noise is an input (k floats per call, candidate order = descending probability), so host and
// oracle can feed identical numbers.  Ties in probability are ordered by ascending token id.
// One CTA of 1024 threads per call; the chosen token is published as an arg-max key (common.cuh), which is
// what the greedy path publishes, so the rest of the step is unchanged.
#pragma once
#include "common.cuh"

namespace msx {

constexpr int kSampleThreads = 1024;
constexpr int kSampleMaxK = 256;

struct SampleArgs {
    const float *logits = nullptr;   // [n]
    int32_t n = 0;
    int32_t k = 0;                   // min(top_k, n) <= kSampleMaxK
    float inv_temp = 1.f;            // 1 / temperature
    const float *noise = nullptr;    // [k] Exp(1) draws, candidate j = j-th largest probability
    unsigned long long *key = nullptr;
    float *probs = nullptr;          // [n] scratch: the probabilities are computed once and re-read by the selection passes
    // batched steps: one CTA per stream (blockIdx.x); per-stream strides in elements, key = a field of the stream's Ctrl
    int32_t logits_stride = 0, noise_stride = 0, probs_stride = 0;
    Ctrl *ctrl = nullptr;            // non-null: key = ctrl[stream].text_key (key_index < 0) or .audio_key[key_index]
    int32_t key_index = -1;
};

__device__ __forceinline__ double block_sum_d(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kSampleThreads / 32; w++) t += scratch[w];
    return t;
}

__global__ void __launch_bounds__(kSampleThreads) sample_kernel(const SampleArgs a0) {
    griddep_launch();
    griddep_wait();
    SampleArgs a = a0;
    {
        const int b = blockIdx.x;
        a.logits += (size_t)b * a.logits_stride; a.noise += (size_t)b * a.noise_stride; a.probs += (size_t)b * a.probs_stride;
        if (a.ctrl) a.key = a.key_index < 0 ? &a.ctrl[b].text_key : &a.ctrl[b].audio_key[a.key_index];
    }
    __shared__ double s_d[kSampleThreads / 32];
    __shared__ float s_f[kSampleThreads / 32];
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_scan[kSampleThreads / 32];
    __shared__ unsigned long long s_keys[kSampleMaxK];
    __shared__ unsigned s_cnt, s_prefix, s_need;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (a.n + kSampleThreads - 1) / kSampleThreads;      // contiguous index range per thread
    const int i0 = tid * per, i1 = min(a.n, i0 + per);

    // ---- softmax(logits / temp): max, exp (through double), sum in double, * (float)(1/sum) ----
    float mx = -INFINITY;
    for (int i = i0; i < i1; i++) mx = fmaxf(mx, a.logits[i] * a.inv_temp);
    mx = warp_max(mx);
    if (lane == 0) s_f[warp] = mx;
    __syncthreads();
    mx = s_f[0];
    for (int w = 1; w < kSampleThreads / 32; w++) mx = fmaxf(mx, s_f[w]);
    double lsum = 0.0;
    for (int i = i0; i < i1; i++) lsum += (double)(float)exp((double)(a.logits[i] * a.inv_temp - mx));
    const double tot = block_sum_d(lsum, s_d);
    const float inv = (float)(1.0 / tot);
    for (int i = i0; i < i1; i++) a.probs[i] = (float)exp((double)(a.logits[i] * a.inv_temp - mx)) * inv;
    auto prob_bits = [&](int i) { return __float_as_uint(a.probs[i]); };   // p >= 0: bits are monotonic; each thread re-reads only what it wrote

    // ---- k-th largest probability by radix select on the 32 probability bits (4 passes x 8 bits) ----
    unsigned prefix = 0, need = (unsigned)a.k;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        const unsigned mask_hi = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = i0; i < i1; i++) {
            const unsigned b = prob_bits(i);
            if ((b & mask_hi) == prefix) atomicAdd(&s_hist[(b >> shift) & 0xff], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned acc = 0; int bin = 255;
            for (; bin > 0; bin--) { if (acc + s_hist[bin] >= need) break; acc += s_hist[bin]; }
            s_prefix = prefix | ((unsigned)bin << shift);
            s_need = need - acc;
        }
        __syncthreads();
        prefix = s_prefix; need = s_need;
        __syncthreads();
    }
    // prefix = bits of the k-th largest probability; `need` of the elements equal to it are taken, lowest ids first
    const unsigned thr = prefix;
    if (tid == 0) s_cnt = 0;
    unsigned ties = 0;
    for (int i = i0; i < i1; i++) ties += prob_bits(i) == thr;
    // exclusive scan of the per-thread tie counts (threads own ascending contiguous index ranges)
    unsigned incl = ties;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < warp; w++) base += s_scan[w];
    unsigned tie_rank = base + incl - ties;
    for (int i = i0; i < i1; i++) {
        const unsigned b = prob_bits(i);
        bool take = b > thr;
        if (b == thr) { take = tie_rank < need; tie_rank++; }
        if (take) {
            const unsigned slot = atomicAdd(&s_cnt, 1u);
            if (slot < (unsigned)kSampleMaxK) s_keys[slot] = ((unsigned long long)b << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
        }
    }
    __syncthreads();
    for (int j = tid; j < kSampleMaxK; j += kSampleThreads) if (j >= a.k) s_keys[j] = 0ull;
    __syncthreads();
    // ---- bitonic sort of the (<= 256) candidates, descending key = descending probability, ascending id ----
    for (int size = 2; size <= kSampleMaxK; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (tid < kSampleMaxK / 2) {
                const int lo = 2 * tid - (tid & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long x = s_keys[lo], y = s_keys[hi];
                if ((x < y) == desc) { s_keys[lo] = y; s_keys[hi] = x; }
            }
            __syncthreads();
        }
    }
    // ---- multinomial: argmax_j p_j / e_j (first maximum) ----
    if (warp == 0) {
        float best = -INFINITY; int bestj = 0;      // (all-NaN candidates fall back to the most probable one)
        for (int j = lane; j < a.k; j += 32) {
            const float p = __uint_as_float((unsigned)(s_keys[j] >> 32));
            const float q = p / a.noise[j];
            if (q > best) { best = q; bestj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
            if (ob > best || (ob == best && oj < bestj)) { best = ob; bestj = oj; }
            bestj = min(bestj, a.k - 1);
        }
        if (lane == 0) {
            const int token = (int)(0xffffffffu - (unsigned)(s_keys[bestj] & 0xffffffffull));
            *a.key = argmax_key(1.0f, token);
        }
    }
}

}  // namespace msx
