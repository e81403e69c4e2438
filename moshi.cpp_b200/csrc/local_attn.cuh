// local_attn.cuh — ring attention of tiny rings fused into the out_proj GEMV of the PDL-chained path (sm_100a).
//
// Depformer layers (transformer.h:973-1039, ring of depformer_context <= 64 slots): every CTA of the out_proj GEMV first
// recomputes the attention of ALL heads into shared memory (one warp per head), then runs the GEMV on it — one launch instead of
// attention + out_proj.  Arithmetic = attention.cuh (bf16-rounded q and p, products summed in double, exact softmax).
#pragma once
#include "attention.cuh"
#include "common.cuh"
#include "gemv.cuh"

namespace msx {

// ---- ring attention of all heads, recomputed by every CTA (depformer) ---------------------------------
// Same arithmetic as attn_kernel (bf16-rounded q and p, products summed in double, exact softmax), one
// warp per head, results to ctx_s[dim] in shared memory.  CTA `writer` also performs the ring insert.
template <int DH>
__device__ __forceinline__ void attn_local(const AttnArgs &a, int heads, float *ctx_s, float *scratch, bool writer, int nwarps) {
    constexpr int LPS = DH / 8;              // lanes per slot
    constexpr int SPI = 32 / LPS;            // slots per warp iteration
    const int lane = threadIdx.x & 31, warp = uniform_warp_id();
    const int cap = a.cap;
    const int pos = a.pos_const >= 0 ? a.pos_const : a.ctrl->offset;
    const int slot = pos % cap;
    const int n_valid = (pos >= cap - 1) ? cap : pos + 1;
    // per-warp scratch: q[DH] f32 | knew[DH] bf16 | vnew[DH] bf16 | scores[64] f32
    float *q_s = scratch + warp * (DH + DH + 64);
    uint16_t *knew = reinterpret_cast<uint16_t *>(q_s + DH);
    uint16_t *vnew = knew + DH;
    float *sc_s = q_s + DH + DH;
    const float scale = 1.f / sqrtf((float)DH);
    constexpr int KPRE = 4, VPRE = 16, NV = DH / 64;       // ring rows kept in registers: KPRE * SPI (K), VPRE (V) slots
    const bool pre = n_valid <= KPRE * SPI && n_valid <= VPRE;
    for (int h = warp; h < heads; h += nwarps) {
        const float *q = a.qkv + h * DH, *k = a.qkv + a.dim + h * DH, *v = a.qkv + 2 * a.dim + h * DH;
        const int g = lane / LPS, sl = lane % LPS;
        // Every load of this head — q / k / v of the step and the K and V rows of ALL valid slots — is requested here, before the
        // first use: one L2 round trip (~0.5 us) instead of one per loop iteration (the tiny rings of the depformer made this
        // kernel a chain of ~10 dependent round trips).
        uint4 kpre[KPRE];
        uint32_t vpre[NV][VPRE];
        if (pre) {
#pragma unroll
            for (int u = 0; u < KPRE; u++) {
                const int i = u * SPI + g;
                kpre[u] = make_uint4(0, 0, 0, 0);
                if (i < n_valid && i != slot) kpre[u] = __ldcg(reinterpret_cast<const uint4 *>(a.kc + ((size_t)h * cap + i) * DH + sl * 8));
            }
#pragma unroll
            for (int s = 0; s < NV; s++)
#pragma unroll
                for (int i = 0; i < VPRE; i++) {
                    vpre[s][i] = 0u;
                    if (i < n_valid && i != slot) vpre[s][i] = __ldcg(reinterpret_cast<const uint32_t *>(a.vc + ((size_t)h * cap + i) * DH + 2 * lane + 64 * s));
                }
        }
        for (int j = lane; j < DH / 2; j += 32) {
            const float2 qq = __ldcg(reinterpret_cast<const float2 *>(q + 2 * j));
            const float2 kk = __ldcg(reinterpret_cast<const float2 *>(k + 2 * j));
            const float2 vv = __ldcg(reinterpret_cast<const float2 *>(v + 2 * j));
            if (a.max_period) {
                const float arg = (float)pos * a.rope_freq[j];
                const float cs = (float)cos((double)arg), sn = (float)sin((double)arg);
                q_s[j] = bf16_round(__fsub_rn(__fmul_rn(qq.x, cs), __fmul_rn(qq.y, sn)));
                q_s[DH / 2 + j] = bf16_round(__fadd_rn(__fmul_rn(qq.x, sn), __fmul_rn(qq.y, cs)));
                knew[j] = f32_to_bf16_bits(__fsub_rn(__fmul_rn(kk.x, cs), __fmul_rn(kk.y, sn)));
                knew[DH / 2 + j] = f32_to_bf16_bits(__fadd_rn(__fmul_rn(kk.x, sn), __fmul_rn(kk.y, cs)));
            } else {
                q_s[2 * j] = bf16_round(qq.x); q_s[2 * j + 1] = bf16_round(qq.y);
                knew[2 * j] = f32_to_bf16_bits(kk.x); knew[2 * j + 1] = f32_to_bf16_bits(kk.y);
            }
            vnew[2 * j] = f32_to_bf16_bits(vv.x); vnew[2 * j + 1] = f32_to_bf16_bits(vv.y);
        }
        __syncwarp();
        if (writer) {
            const size_t o = ((size_t)h * cap + slot) * DH;
            for (int t = lane; t < DH / 4; t += 32) {
                reinterpret_cast<uint2 *>(a.kc + o)[t] = reinterpret_cast<const uint2 *>(knew)[t];
                reinterpret_cast<uint2 *>(a.vc + o)[t] = reinterpret_cast<const uint2 *>(vnew)[t];
            }
        }
        float qv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) qv[i] = q_s[sl * 8 + i];
        float lmax = -INFINITY;
        auto score = [&](int i, uint4 kk) {
            const bool valid = i < n_valid;
            if (valid && i == slot) kk = reinterpret_cast<const uint4 *>(knew)[sl];
            double d = 0.0;
            d += (double)(bf16_bits_to_f32(kk.x & 0xffff) * qv[0]); d += (double)(bf16_bits_to_f32(kk.x >> 16) * qv[1]);
            d += (double)(bf16_bits_to_f32(kk.y & 0xffff) * qv[2]); d += (double)(bf16_bits_to_f32(kk.y >> 16) * qv[3]);
            d += (double)(bf16_bits_to_f32(kk.z & 0xffff) * qv[4]); d += (double)(bf16_bits_to_f32(kk.z >> 16) * qv[5]);
            d += (double)(bf16_bits_to_f32(kk.w & 0xffff) * qv[6]); d += (double)(bf16_bits_to_f32(kk.w >> 16) * qv[7]);
#pragma unroll
            for (int o = LPS / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            const float s = (float)d * scale + 0.0f;
            if (valid) { if (sl == 0) sc_s[i] = s; lmax = fmaxf(lmax, s); }
        };
        if (pre) {
#pragma unroll
            for (int u = 0; u < KPRE; u++) if (u * SPI < n_valid) score(u * SPI + g, kpre[u]);     // warp-uniform trip count
        } else {
            for (int i0 = 0; i0 < n_valid; i0 += SPI) {
                const int i = i0 + g;
                uint4 kk = make_uint4(0, 0, 0, 0);
                if (i < n_valid && i != slot) kk = __ldcg(reinterpret_cast<const uint4 *>(a.kc + ((size_t)h * cap + i) * DH + sl * 8));
                score(i, kk);
            }
        }
        lmax = warp_max(lmax);
        __syncwarp();
        double lsum = 0.0;
        for (int i = lane; i < n_valid; i += 32) { const float e = (float)exp((double)(sc_s[i] - lmax)); sc_s[i] = e; lsum += (double)e; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        const float inv = (float)(1.0 / lsum);
        __syncwarp();
        // context: lane owns dims {2*lane, 2*lane+1} (+64 for DH = 128)
#pragma unroll
        for (int s = 0; s < NV; s++) {
            const int d0 = 2 * lane + 64 * s;
            double acc0 = 0.0, acc1 = 0.0;
            if (pre) {
#pragma unroll
                for (int i = 0; i < VPRE; i++) if (i < n_valid) {
                    const float p = bf16_round(sc_s[i] * inv);
                    const uint32_t vv = i == slot ? *reinterpret_cast<const uint32_t *>(vnew + d0) : vpre[s][i];
                    acc0 += (double)(bf16_bits_to_f32(vv & 0xffff) * p);
                    acc1 += (double)(bf16_bits_to_f32(vv >> 16) * p);
                }
            } else {
                for (int i = 0; i < n_valid; i++) {
                    const float p = bf16_round(sc_s[i] * inv);
                    uint32_t vv;
                    if (i == slot) vv = *reinterpret_cast<const uint32_t *>(vnew + d0);
                    else vv = __ldcg(reinterpret_cast<const uint32_t *>(a.vc + ((size_t)h * cap + i) * DH + d0));
                    acc0 += (double)(bf16_bits_to_f32(vv & 0xffff) * p);
                    acc1 += (double)(bf16_bits_to_f32(vv >> 16) * p);
                }
            }
            ctx_s[h * DH + d0] = (float)acc0; ctx_s[h * DH + d0 + 1] = (float)acc1;
        }
        __syncwarp();
    }
}

// ---- standalone fused kernel: local ring attention (all heads, every CTA) + out_proj GEMV -----------------
// Used by the PDL-chained path for transformers whose ring is tiny (depformer: <= 64 slots): one launch
// instead of attention + out_proj.  smem: [gemv region][ctx_s dim floats][attn scratch]
template <int WT, int LANES, int DH>
__global__ void __launch_bounds__(kGemvThreads, 1) dq_matvec_local_attn_kernel(const MatvecArgs g, const AttnArgs a, const int heads, const int pro,
                                                                          const int epi, const int gemv_region) {
    extern __shared__ __align__(16) uint8_t smem[];
    griddep_launch();
    const BlockGeom bg{kGemvThreads, kGemvThreads / 32};
    float *ctx_s = reinterpret_cast<float *>(smem + gemv_region);
    float *scratch = ctx_s + a.dim;
    // the out_proj weights do not depend on qkv: gemv_body requests its first steps, THEN waits for the previous kernel and
    // runs the attention (the whole 0.6 MB matrix of a depformer layer is in flight under it)
    auto mid = [&]() {
        griddep_wait();                              // qkv comes from the previous kernel
        attn_local<DH>(a, heads, ctx_s, scratch, blockIdx.x == 0, bg.nwarps);
        block_sync(bg);
    };
    gemv_body<WT, LANES, false, false, true, decltype(mid)>(g, ctx_s, false, pro, epi, smem, blockIdx.x, gridDim.x, bg, mid);   // lean body: residual epilogue only
}
__host__ __device__ inline int local_attn_smem_bytes(int gemv_bytes, int dim, int dh) {
    return (gemv_bytes + 15) / 16 * 16 + dim * 4 + (kGemvThreads / 32) * (2 * dh + 64) * 4 + 64;
}

}  // namespace msx
