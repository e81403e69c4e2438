// mma_gemm.cuh — batched-stream dequant-GEMM: one Q4_K weight matrix times up to 8 activation columns
// (one column per conversation stream) on the int8 tensor-core path (sm_100a).
//
// Replaces, for a batch of independent streams stepped in lock-step,
//     torch_nn_linear = ggml_mul_mat(W_q4_K, x)              src/torch.h:79-87
// with the same surrounding fusions as gemv.cuh (residual, silu-gate, arg-max, +embedding); the weights are
// read from HBM ONCE per step for all streams (SURVEY.md §8e "independent streams ... batch b = 8").
//
// Numerics are those of gemv.cuh / ggml's CPU mul_mat, per column: Q8_K activations, exact integer
// sub-block dots (here: mma.sync m16n8k32 u8 x s8 -> s32, one MMA per 32-weight sub-block of 16 rows),
// integer scaling by the 6-bit scales / mins, fp32 block scales, DOUBLE accumulation of the exact block
// terms, one rounding.  A batched step therefore produces, stream by stream, the single-stream result.
//
// Weight layout ("units"): the matrix is cut into tiles of 16 rows; one unit = one tile x one 256-weight
// super-block = 2368 contiguous bytes, already in MMA A-fragment order:
//     [pair p 0..3][lane 0..31] 16 B = raw nibble words {row g: k-word t, row g+8: k-word t, row g: k-word 4+t,
//                                      row g+8: k-word 4+t}   (g = lane/4, t = lane%4; low nibbles = sub-block
//                                      2p, high nibbles = sub-block 2p+1)
//     [row 0..15][pair 0..3]    4 B  = {sc_lo, sc_hi, m_lo, m_hi}
//     [row 0..15]               4 B  = {fp16 d, fp16 dmin}
// A warp streams whole units with ONE bulk-copy (TMA) instruction each into its private shared-memory ring,
// completion signalled on an mbarrier.  For the gated MLP a tile holds 8 gate rows (tile rows 0-7) and the 8
// matching up rows (8-15), so a thread's accumulator pair (row g, row g+8) is exactly one silu-gate output.
#pragma once
#include "common.cuh"
#include "gemv.cuh"

namespace msx {

constexpr int kUnitBytes = 2368;
constexpr int kMmaCols = 8;                 // activation columns of one MMA (= streams per batch)
constexpr int kGemmThreads = 512;
constexpr int kGemmWarps = kGemmThreads / 32;
constexpr int kPartBytes = 2 * kGemmWarps * 4 * 32 * 8;   // double-buffered per-warp partial accumulators
constexpr int kGemmMaxStages = 6;
constexpr int kGemmSmemMax = 227 * 1024;

struct QTiles {
    const uint8_t *units = nullptr;
    int32_t K = 0, rows = 0, n_tiles = 0, nsb = 0;   // nsb = units per tile row (K / 256 for Q4_K, K / 128 for Q8_0)
    int32_t type = 12;
};
// Q8_0 units: 16 rows x 128 weights (4 blocks of 32) = [4 blocks][32 lanes] 16 B of int8 in A-fragment order
// {row g: k-word t, row g+8: k-word t, row g: k-word 4+t, row g+8: k-word 4+t} + [16 rows][4 blocks] fp16 d = 2176 bytes
constexpr int kUnitBytesQ8 = 2176;
__host__ __device__ constexpr int unit_bytes_of(int wt) { return wt == 12 ? 2368 : kUnitBytesQ8; }
__host__ __device__ constexpr int unit_weights_of(int wt) { return wt == 12 ? 256 : 128; }

// Quantised activation image of one GEMM input: written by quant_q8k_kernel, bulk-copied verbatim into
// shared memory by the GEMM.
//     xq  [K/64 pairs][32 lanes] 16 B : B-fragments {sub-block 2p: k 4t.., k 16+4t..; sub-block 2p+1: same}, col = lane/4
//     bsw [K/64 pairs][8 cols]   4 B  : int16 sums of the two 32-element sub-blocks (for the dmin term)
//     dx  [K/256][8 cols]        f32  : Q8_K block scales
// Q8_0 image: xq as above (sub-block = 32-block) + dx [K/32][8 cols] f32 (fp16-rounded Q8_0 block scales)
__host__ __device__ inline int act_image_bytes(int K, int wt = 12) { return wt == 12 ? K * 8 + K / 2 + K / 8 : K * 8 + K; }
__host__ __device__ inline int gemm_fixed_smem(int K, int wt = 12) { return ((act_image_bytes(K, wt) + 127) & ~127) + kPartBytes + 1024; }
__host__ inline int gemm_stages_for(int K, int wt = 12) {
    const int s = (kGemmSmemMax - gemm_fixed_smem(K, wt)) / (kGemmWarps * unit_bytes_of(wt));
    return s > 4 ? 4 : s;
}
__host__ inline int gemm_smem_bytes(int K, int stages, int wt = 12) { return gemm_fixed_smem(K, wt) + stages * kGemmWarps * unit_bytes_of(wt); }

// ---- load-time: QLinear planes (common.cuh) -> units ------------------------------------------------
// gate != 0: the planes hold interleaved (gate j, up j) rows; tile row r <- virtual row 16*tile + 2*(r&7) + (r>>3)
__global__ void tile_q4k_kernel(const QLinear w, int gate, uint8_t *units, int n_tiles) {
    const int nsb = w.K >> 8, P = w.K >> 6;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (tile, sb, 16-byte chunk)
    if (idx >= (long long)n_tiles * nsb * 148) return;
    const int chunk = (int)(idx % 148);
    const long long u = idx / 148;
    const int sb = (int)(u % nsb), tile = (int)(u / nsb);
    auto vrow = [&](int r) { return gate ? 16 * tile + 2 * (r & 7) + (r >> 3) : 16 * tile + r; };
    uint4 o = make_uint4(0, 0, 0, 0);
    if (chunk < 128) {
        const int p = chunk >> 5, lane = chunk & 31, g = lane >> 2, t = lane & 3;
        const int pp = sb * 4 + p;
        const int G = pp / w.gs, q = pp % w.gs, gsz = min(w.gs, P - G * w.gs);
        uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int row = vrow(g + 8 * h);
            if (row < w.rows) {
                const uint32_t *c0 = reinterpret_cast<const uint32_t *>(w.qs + (size_t)row * (w.K >> 1) + (size_t)G * w.gs * 32 + q * 16);
                v[h] = c0[t];
                v[2 + h] = c0[gsz * 4 + t];
            }
        }
        o = make_uint4(v[0], v[1], v[2], v[3]);
    } else if (chunk < 144) {
        const int row = vrow(chunk - 128);
        if (row < w.rows) o = *reinterpret_cast<const uint4 *>(w.sc + (size_t)row * P + sb * 4);
    } else {
        uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int row = vrow((chunk - 144) * 4 + i);
            if (row < w.rows) v[i] = reinterpret_cast<const uint32_t *>(w.dd)[(size_t)row * nsb + sb];
        }
        o = make_uint4(v[0], v[1], v[2], v[3]);
    }
    reinterpret_cast<uint4 *>(units + (size_t)u * kUnitBytes)[chunk] = o;
}

// Q8_0 planes (common.cuh) -> Q8_0 units
__global__ void tile_q8_0_kernel(const QLinear w, int gate, uint8_t *units, int n_tiles) {
    const int nsb = w.K >> 7, P = w.K >> 5;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (tile, unit, 16-byte chunk): 136 chunks per unit
    if (idx >= (long long)n_tiles * nsb * 136) return;
    const int chunk = (int)(idx % 136);
    const long long u = idx / 136;
    const int sb = (int)(u % nsb), tile = (int)(u / nsb);
    auto vrow = [&](int r) { return gate ? 16 * tile + 2 * (r & 7) + (r >> 3) : 16 * tile + r; };
    uint4 o = make_uint4(0, 0, 0, 0);
    if (chunk < 128) {
        const int bi = chunk >> 5, lane = chunk & 31, g = lane >> 2, t = lane & 3;
        const int pp = sb * 4 + bi;                                            // 32-block index inside the row
        const int G = pp / w.gs, q = pp % w.gs, gsz = min(w.gs, P - G * w.gs);
        uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int row = vrow(g + 8 * h);
            if (row < w.rows) {
                const uint32_t *c0 = reinterpret_cast<const uint32_t *>(w.qs + (size_t)row * w.K + (size_t)G * w.gs * 32 + q * 16);
                v[h] = c0[t];
                v[2 + h] = c0[gsz * 4 + t];
            }
        }
        o = make_uint4(v[0], v[1], v[2], v[3]);
    } else {
        // dd: [16 rows][4 blocks] fp16 = 8 bytes per row -> chunk c holds rows 2c, 2c+1
        uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int row = vrow((chunk - 128) * 2 + i);
            if (row < w.rows) {
                const uint16_t *d = reinterpret_cast<const uint16_t *>(w.dd) + (size_t)row * P + sb * 4;
                v[2 * i] = (uint32_t)d[0] | ((uint32_t)d[1] << 16);
                v[2 * i + 1] = (uint32_t)d[2] | ((uint32_t)d[3] << 16);
            }
        }
        o = make_uint4(v[0], v[1], v[2], v[3]);
    }
    reinterpret_cast<uint4 *>(units + (size_t)u * kUnitBytesQ8)[chunk] = o;
}

// ---- activation quantisation: one CTA per stream ------------------------------------------------------
struct QuantArgs {
    const float *x = nullptr; int32_t ld = 0;        // x[b * ld + i]
    const float *alpha = nullptr; float eps = 0.f;   // alpha != null: RMSNorm first (transformer.h:15-23)
    float *norm_out = nullptr; int32_t norm_ld = 0;  // optional copy of the normalised vector (transformer_out)
    uint8_t *img = nullptr;
    int32_t K = 0;
    int32_t plain = 0;                               // 1: image for the tcgen05 GEMM (tc_gemm.cuh: one record per super-block, operand layout)
};

constexpr int kQChunk = 3;
// CTAs per column: every CTA reduces the whole column (RMSNorm) but quantises only every gridDim.y-th block of each warp
__host__ inline int quant_parts_for(int K) { const int per_warp = ((K >> 8) + kGemmWarps - 1) / kGemmWarps; return per_warp < 1 ? 1 : (per_warp > kQChunk ? kQChunk : per_warp); }
__global__ void __launch_bounds__(kGemmThreads) quant_q8k_kernel(const QuantArgs a) {
    __shared__ double red[kGemmWarps];
    griddep_launch();
    griddep_wait();
    const int col = blockIdx.x, lane = threadIdx.x & 31, warp = uniform_warp_id();
    const int K = a.K, nblk = K >> 8;
    const float *x = a.x + (size_t)col * a.ld;
    const int nb_w = warp < nblk ? (nblk - warp + kGemmWarps - 1) / kGemmWarps : 0;
    const bool keep = nb_w <= kQChunk;
    const bool norm = a.alpha != nullptr;
    float v[kQChunk][8], al[kQChunk][8];
    float scale = 1.f;
    auto load8 = [&](int e0, float (&d)[8], const float *src) {
        const float4 p0 = __ldcg(reinterpret_cast<const float4 *>(src + e0)), p1 = __ldcg(reinterpret_cast<const float4 *>(src + e0 + 4));
        d[0] = p0.x; d[1] = p0.y; d[2] = p0.z; d[3] = p0.w; d[4] = p1.x; d[5] = p1.y; d[6] = p1.z; d[7] = p1.w;
    };
    auto e0_of = [&](int i) { return (warp + i * kGemmWarps) * 256 + lane * 8; };
    if (norm) {
        double ss = 0.0;
#pragma unroll 1
        for (int c0 = 0; c0 < nb_w; c0 += kQChunk) {
#pragma unroll
            for (int j = 0; j < kQChunk; j++) {
                if (c0 + j < nb_w) { load8(e0_of(c0 + j), v[j], x); if (keep) load8(e0_of(c0 + j), al[j], a.alpha); }
                else {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[j][i] = 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < kQChunk; j++)
#pragma unroll
                for (int i = 0; i < 8; i++) ss += (double)(v[j][i] * v[j][i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        double tot = 0.0;
#pragma unroll 1
        for (int w = 0; w < kGemmWarps; w++) tot += red[w];
        const float mean = (K & (K - 1)) == 0 ? (float)scalbn(tot, -(31 - __clz(K))) : (float)(tot / K);
        scale = 1.0f / sqrtf(mean + a.eps);
    }
    uint32_t *bsw = reinterpret_cast<uint32_t *>(a.img + (size_t)K * 8);
    float *dx = reinterpret_cast<float *>(a.img + (size_t)K * 8 + (K >> 1));
#pragma unroll 1
    for (int c0 = 0; c0 < nb_w; c0 += kQChunk) {
        if (!(norm && keep)) {
#pragma unroll
            for (int j = 0; j < kQChunk; j++)
                if (c0 + j < nb_w) { load8(e0_of(c0 + j), v[j], x); if (norm) load8(e0_of(c0 + j), al[j], a.alpha); }
        }
#pragma unroll
        for (int j = 0; j < kQChunk; j++) {
            if (c0 + j >= nb_w) continue;                    // warp-uniform
            if ((c0 + j) % (int)gridDim.y != (int)blockIdx.y) continue;   // long columns: the warp's blocks are dealt over gridDim.y CTAs
            const int blk = warp + (c0 + j) * kGemmWarps, e0 = e0_of(c0 + j);
            if (norm) {
#pragma unroll
                for (int i = 0; i < 8; i++) v[j][i] = __fmul_rn(al[j][i], __fmul_rn(v[j][i], scale));
                if (a.norm_out) {
                    float *no = a.norm_out + (size_t)col * a.norm_ld + e0;
                    *reinterpret_cast<float4 *>(no) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
                    *reinterpret_cast<float4 *>(no + 4) = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
                }
            }
            // quantize_row_q8_K (see gemv.cuh quantize_block_q8k)
            float amax = 0.f, mx = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) { const float ax = fabsf(v[j][i]); if (ax > amax) { amax = ax; mx = v[j][i]; } }
            const float wmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(amax)));
            const unsigned hit = __ballot_sync(0xffffffffu, amax == wmax);
            const float carrier = __shfl_sync(0xffffffffu, mx, __ffs(hit) - 1);
            int q[8];
            float d = 0.f;
            if (wmax == 0.f) {
#pragma unroll
                for (int i = 0; i < 8; i++) q[i] = 0;
            } else {
                const float iscale = -127.f / carrier;
#pragma unroll
                for (int i = 0; i < 8; i++) { const int t = __float2int_rn(iscale * v[j][i]); q[i] = t < 127 ? t : 127; }
                d = 1.f / iscale;
            }
            int s = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) s += q[i];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);                 // sum of the lane's 32-element sub-block
            const int s_next = __shfl_down_sync(0xffffffffu, s, 4);    // the following sub-block
            if (a.plain) {             // record of super-block blk: x8 tile in operand layout | sums per 32 [64][8] | scales [64]
                uint8_t *rec = a.img + (size_t)blk * (64 * 256 + 64 * 16 + 64 * 4);
                if ((lane & 3) == 0) reinterpret_cast<int16_t *>(rec + 64 * 256)[col * 8 + (lane >> 2)] = (int16_t)s;
                if (lane == 0) reinterpret_cast<float *>(rec + 64 * 256 + 64 * 16)[col] = d;
                *reinterpret_cast<uint2 *>(rec + (col >> 3) * 2048 + (lane >> 1) * 128 + (col & 7) * 16 + (lane & 1) * 8) = pack8(q);
                continue;
            }
            const int sbk = lane >> 2, jj = lane & 3;
            const int p = 4 * blk + (sbk >> 1);
            if ((lane & 7) == 0) bsw[(size_t)p * 8 + col] = (uint32_t)(s & 0xffff) | ((uint32_t)(s_next & 0xffff) << 16);
            if (lane == 0) dx[(size_t)blk * 8 + col] = d;
            const uint2 pk = pack8(q);
            // fragment order: word of k-offset 4t' (t' = 2jj, 2jj+1) -> which = t'/4, t = t'%4
            uint32_t *dst = reinterpret_cast<uint32_t *>(a.img + ((size_t)p * 32 + col * 4) * 16) + (sbk & 1) * 2 + (jj >> 1);
            dst[(2 * (jj & 1)) * 4] = pk.x;
            dst[(2 * (jj & 1) + 1) * 4] = pk.y;
        }
    }
}

// Q8_0 activations (ggml quantize_row_q8_0 per 32: d = amax / 127, q = roundf(x / d), d kept as fp16) in the same fragment
// order; dx [K/32][8 cols] f32.  One CTA per stream, same structure as quant_q8k_kernel.
__global__ void __launch_bounds__(kGemmThreads) quant_q8_0_kernel(const QuantArgs a) {
    __shared__ double red[kGemmWarps];
    griddep_launch();
    griddep_wait();
    const int col = blockIdx.x, lane = threadIdx.x & 31, warp = uniform_warp_id();
    const int K = a.K, nblk = (K + 255) >> 8;
    const float *x = a.x + (size_t)col * a.ld;
    const int nb_w = warp < nblk ? (nblk - warp + kGemmWarps - 1) / kGemmWarps : 0;
    const bool norm = a.alpha != nullptr;
    float scale = 1.f;
    auto load8 = [&](int e0, float (&d)[8], const float *src) {
        if (e0 < K) {
            const float4 p0 = __ldcg(reinterpret_cast<const float4 *>(src + e0)), p1 = __ldcg(reinterpret_cast<const float4 *>(src + e0 + 4));
            d[0] = p0.x; d[1] = p0.y; d[2] = p0.z; d[3] = p0.w; d[4] = p1.x; d[5] = p1.y; d[6] = p1.z; d[7] = p1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) d[i] = 0.f;
        }
    };
    if (norm) {
        double ss = 0.0;
        for (int j = 0; j < nb_w; j++) {
            float v[8];
            load8((warp + j * kGemmWarps) * 256 + lane * 8, v, x);
#pragma unroll
            for (int i = 0; i < 8; i++) ss += (double)(v[i] * v[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        double tot = 0.0;
        for (int w = 0; w < kGemmWarps; w++) tot += red[w];
        const float mean = (K & (K - 1)) == 0 ? (float)scalbn(tot, -(31 - __clz(K))) : (float)(tot / K);
        scale = 1.0f / sqrtf(mean + a.eps);
    }
    float *dx = reinterpret_cast<float *>(a.img + (size_t)K * 8);
    for (int j = 0; j < nb_w; j++) {
        if (j % (int)gridDim.y != (int)blockIdx.y) continue;
        const int e0 = (warp + j * kGemmWarps) * 256 + lane * 8;
        float v[8], al[8];
        load8(e0, v, x);
        const bool act = e0 < K;
        if (norm) {
            load8(e0, al, a.alpha);
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __fmul_rn(al[i], __fmul_rn(v[i], scale));
            if (a.norm_out && act) {
                float *no = a.norm_out + (size_t)col * a.norm_ld + e0;
                *reinterpret_cast<float4 *>(no) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(no + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));        // the lane's 32-element block
        const float d = amax / 127.f;
        const float id = d ? 1.0f / d : 0.0f;
        if (!act) continue;
        int q[8];
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = (int)roundf(v[i] * id);
        const int blk32 = e0 >> 5, jj = lane & 3;                       // 32-block index in the row
        const int p = blk32 >> 1;                                         // fragment "pair" of two 32-blocks
        if (jj == 0) dx[(size_t)blk32 * 8 + col] = __half2float(__float2half_rn(d));
        const uint2 pk = pack8(q);
        uint32_t *dst = reinterpret_cast<uint32_t *>(a.img + ((size_t)p * 32 + col * 4) * 16) + (blk32 & 1) * 2 + (jj >> 1);
        dst[(2 * (jj & 1)) * 4] = pk.x;
        dst[(2 * (jj & 1) + 1) * 4] = pk.y;
    }
}

// ---- the GEMM -----------------------------------------------------------------------------------------
struct MatmulArgs {
    QTiles w;
    const uint8_t *img = nullptr;     // activation image of this input (act_image_bytes(K))
    float *out = nullptr; int32_t ld = 0;   // out[col * ld + row]
    int32_t nb = 0;                   // live columns (streams)
    int32_t epi = 0;
    int32_t stages = 4;
    Ctrl *ctrl = nullptr;             // [nb] control blocks
    int32_t key_index = -1;           // EPI_ARGMAX: -1 = text_key, k = audio_key[k]
    int32_t emb_step = 0;             // EPI_ADD_EMB
    EmbTable emb;
    long long *stamps = nullptr;      // debug: [grid][8] globaltimer stamps
    // fused activation prologue for short inner dimensions (K * nb <= 16384): every CTA normalises + quantises the nb
    // columns itself instead of copying an image produced by a separate quant_q8k_kernel launch (one launch less in
    // the depformer's dependent chain)
    const float *xsrc = nullptr; int32_t xld = 0;
    const float *alpha = nullptr; float eps = 0.f;
};
constexpr int kFusedMaxPairs = 4;     // (column, 256-block) pairs per warp in the fused prologue
__host__ __device__ inline bool gemm_can_fuse_quant(int K, int nb) { return (K >> 8) * nb <= kGemmWarps * kFusedMaxPairs && K <= 1024; }

__device__ __forceinline__ void mma_u8s8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}
// c + a.s16[0] * b.u8[2] + a.s16[1] * b.u8[3]
__device__ __forceinline__ int dp2a_hi_su(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// one unit (16 rows x 256 weights) against the 8 activation columns: acc[] = {(row g, col 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)}
__device__ __forceinline__ void mma_s8s8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}
// Q8_0 unit (16 rows x 4 blocks of 32) against the 8 columns: ggml_vec_dot_q8_0_q8_0 per block: sumi * (fp16 d_w * fp16 d_x),
// the block terms accumulated in double
__device__ __forceinline__ void unit_compute_q8(const uint8_t *slot, const uint8_t *img, int K, int sb, int lane, double (&acc)[4]) {
    const int g = lane >> 2, t = lane & 3;
    const uint2 ddg2 = *reinterpret_cast<const uint2 *>(slot + 2048 + g * 8);
    const uint2 ddh2 = *reinterpret_cast<const uint2 *>(slot + 2048 + (g + 8) * 8);
    const uint32_t ddg[2] = {ddg2.x, ddg2.y}, ddh[2] = {ddh2.x, ddh2.y};
    const uint8_t *xq = img + (size_t)sb * 1024 + lane * 16;
    const float *dx = reinterpret_cast<const float *>(img + (size_t)K * 8) + (size_t)sb * 32 + 2 * t;
#pragma unroll
    for (int p = 0; p < 2; p++) {
        const uint4 w0 = *reinterpret_cast<const uint4 *>(slot + (2 * p) * 512 + lane * 16);
        const uint4 w1 = *reinterpret_cast<const uint4 *>(slot + (2 * p + 1) * 512 + lane * 16);
        const uint4 xb = *reinterpret_cast<const uint4 *>(xq + p * 512);
        int c[4], d[4];
        mma_s8s8(c, w0.x, w0.y, w0.z, w0.w, xb.x, xb.y);
        mma_s8s8(d, w1.x, w1.y, w1.z, w1.w, xb.z, xb.w);
        const float2 wg = __half22float2(*reinterpret_cast<const __half2 *>(&ddg[p]));      // d of blocks 2p, 2p+1, row g
        const float2 wh = __half22float2(*reinterpret_cast<const __half2 *>(&ddh[p]));      // row g+8
        const float2 x0 = *reinterpret_cast<const float2 *>(dx + (2 * p) * 8);               // block 2p: columns 2t, 2t+1
        const float2 x1 = *reinterpret_cast<const float2 *>(dx + (2 * p + 1) * 8);
        acc[0] = fma((double)__fmul_rn(wg.x, x0.x), (double)c[0], acc[0]); acc[1] = fma((double)__fmul_rn(wg.x, x0.y), (double)c[1], acc[1]);
        acc[2] = fma((double)__fmul_rn(wh.x, x0.x), (double)c[2], acc[2]); acc[3] = fma((double)__fmul_rn(wh.x, x0.y), (double)c[3], acc[3]);
        acc[0] = fma((double)__fmul_rn(wg.y, x1.x), (double)d[0], acc[0]); acc[1] = fma((double)__fmul_rn(wg.y, x1.y), (double)d[1], acc[1]);
        acc[2] = fma((double)__fmul_rn(wh.y, x1.x), (double)d[2], acc[2]); acc[3] = fma((double)__fmul_rn(wh.y, x1.y), (double)d[3], acc[3]);
    }
}

__device__ __forceinline__ void unit_compute(const uint8_t *slot, const uint8_t *img, int K, int sb, int lane, double (&acc)[4]) {
    const int g = lane >> 2, t = lane & 3;
    const uint4 scg4 = *reinterpret_cast<const uint4 *>(slot + 2048 + g * 16);
    const uint4 sch4 = *reinterpret_cast<const uint4 *>(slot + 2048 + (g + 8) * 16);
    const uint32_t ddg = *reinterpret_cast<const uint32_t *>(slot + 2304 + g * 4);
    const uint32_t ddh = *reinterpret_cast<const uint32_t *>(slot + 2304 + (g + 8) * 4);
    const uint32_t scg[4] = {scg4.x, scg4.y, scg4.z, scg4.w}, sch[4] = {sch4.x, sch4.y, sch4.z, sch4.w};
    int lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, mn[4] = {0, 0, 0, 0};
    const uint8_t *xq = img + (size_t)sb * 2048 + lane * 16;
    const uint32_t *bsw = reinterpret_cast<const uint32_t *>(img + (size_t)K * 8) + sb * 32 + 2 * t;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const uint4 w = *reinterpret_cast<const uint4 *>(slot + p * 512 + lane * 16);
        const uint4 xb = *reinterpret_cast<const uint4 *>(xq + p * 512);
        int c[4], d[4];
        mma_u8s8(c, w.x & 0x0F0F0F0Fu, w.y & 0x0F0F0F0Fu, w.z & 0x0F0F0F0Fu, w.w & 0x0F0F0F0Fu, xb.x, xb.y);
        mma_u8s8(d, w.x & 0xF0F0F0F0u, w.y & 0xF0F0F0F0u, w.z & 0xF0F0F0F0u, w.w & 0xF0F0F0F0u, xb.z, xb.w);   // 16 x the high-nibble dots
        const int slg = (int)__byte_perm(scg[p], 0, 0x4440), shg = (int)__byte_perm(scg[p], 0, 0x4441);
        const int slh = (int)__byte_perm(sch[p], 0, 0x4440), shh = (int)__byte_perm(sch[p], 0, 0x4441);
        lo[0] += slg * c[0]; lo[1] += slg * c[1]; lo[2] += slh * c[2]; lo[3] += slh * c[3];
        hi[0] += shg * d[0]; hi[1] += shg * d[1]; hi[2] += shh * d[2]; hi[3] += shh * d[3];
        const uint2 bw = *reinterpret_cast<const uint2 *>(bsw + p * 8);
        mn[0] = dp2a_hi_su(bw.x, scg[p], mn[0]); mn[1] = dp2a_hi_su(bw.y, scg[p], mn[1]);
        mn[2] = dp2a_hi_su(bw.x, sch[p], mn[2]); mn[3] = dp2a_hi_su(bw.y, sch[p], mn[3]);
    }
    const float2 dxv = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(img + (size_t)K * 8 + (K >> 1)) + sb * 8 + 2 * t);
    const float2 dmg = __half22float2(*reinterpret_cast<const __half2 *>(&ddg));
    const float2 dmh = __half22float2(*reinterpret_cast<const __half2 *>(&ddh));
    acc[0] = fma((double)__fmul_rn(dmg.x, dxv.x), (double)(lo[0] + (hi[0] >> 4)), acc[0]); acc[0] = fma(-(double)__fmul_rn(dmg.y, dxv.x), (double)mn[0], acc[0]);
    acc[1] = fma((double)__fmul_rn(dmg.x, dxv.y), (double)(lo[1] + (hi[1] >> 4)), acc[1]); acc[1] = fma(-(double)__fmul_rn(dmg.y, dxv.y), (double)mn[1], acc[1]);
    acc[2] = fma((double)__fmul_rn(dmh.x, dxv.x), (double)(lo[2] + (hi[2] >> 4)), acc[2]); acc[2] = fma(-(double)__fmul_rn(dmh.y, dxv.x), (double)mn[2], acc[2]);
    acc[3] = fma((double)__fmul_rn(dmh.x, dxv.y), (double)(lo[3] + (hi[3] >> 4)), acc[3]); acc[3] = fma(-(double)__fmul_rn(dmh.y, dxv.y), (double)mn[3], acc[3]);
}

// fused prologue: (column, block) pairs p = warp, warp + 16, ... ; same arithmetic and image layout as quant_q8k_kernel
__device__ __forceinline__ void fused_quant_columns(const MatmulArgs &a, uint8_t *img, double *sscratch, int warp, int lane) {
    const int K = a.w.K, nblk = K >> 8, npairs = a.nb * nblk;
    const bool norm = a.alpha != nullptr;
    float v[kFusedMaxPairs][8], al[kFusedMaxPairs][8];
    auto load8 = [&](const float *src, float (&d)[8]) {
        const float4 p0 = __ldcg(reinterpret_cast<const float4 *>(src)), p1 = __ldcg(reinterpret_cast<const float4 *>(src + 4));
        d[0] = p0.x; d[1] = p0.y; d[2] = p0.z; d[3] = p0.w; d[4] = p1.x; d[5] = p1.y; d[6] = p1.z; d[7] = p1.w;
    };
#pragma unroll
    for (int j = 0; j < kFusedMaxPairs; j++) {
        const int p = warp + j * kGemmWarps;
        if (p < npairs) {
            const int col = p / nblk, blk = p - col * nblk, e0 = blk * 256 + lane * 8;
            load8(a.xsrc + (size_t)col * a.xld + e0, v[j]);
            if (norm) load8(a.alpha + e0, al[j]);
        }
    }
    if (norm) {
#pragma unroll
        for (int j = 0; j < kFusedMaxPairs; j++) {
            const int p = warp + j * kGemmWarps;
            if (p < npairs) {
                double ss = 0.0;
#pragma unroll
                for (int i = 0; i < 8; i++) ss += (double)(v[j][i] * v[j][i]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                if (lane == 0) sscratch[p] = ss;
            }
        }
        __syncthreads();
    }
    uint32_t *bsw = reinterpret_cast<uint32_t *>(img + (size_t)K * 8);
    float *dx = reinterpret_cast<float *>(img + (size_t)K * 8 + (K >> 1));
#pragma unroll
    for (int j = 0; j < kFusedMaxPairs; j++) {
        const int p = warp + j * kGemmWarps;
        if (p >= npairs) continue;                          // warp-uniform
        const int col = p / nblk, blk = p - col * nblk;
        if (norm) {
            double tot = 0.0;
            for (int b = 0; b < nblk; b++) tot += sscratch[col * nblk + b];
            const float mean = (K & (K - 1)) == 0 ? (float)scalbn(tot, -(31 - __clz(K))) : (float)(tot / K);
            const float scale = 1.0f / sqrtf(mean + a.eps);
#pragma unroll
            for (int i = 0; i < 8; i++) v[j][i] = __fmul_rn(al[j][i], __fmul_rn(v[j][i], scale));
        }
        float amax = 0.f, mx = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) { const float ax = fabsf(v[j][i]); if (ax > amax) { amax = ax; mx = v[j][i]; } }
        const float wmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(amax)));
        const unsigned hit = __ballot_sync(0xffffffffu, amax == wmax);
        const float carrier = __shfl_sync(0xffffffffu, mx, __ffs(hit) - 1);
        int q[8];
        float d = 0.f;
        if (wmax == 0.f) {
#pragma unroll
            for (int i = 0; i < 8; i++) q[i] = 0;
        } else {
            const float iscale = -127.f / carrier;
#pragma unroll
            for (int i = 0; i < 8; i++) { const int t = __float2int_rn(iscale * v[j][i]); q[i] = t < 127 ? t : 127; }
            d = 1.f / iscale;
        }
        int s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += q[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const int s_next = __shfl_down_sync(0xffffffffu, s, 4);
        const int sbk = lane >> 2, jj = lane & 3;
        const int pr = 4 * blk + (sbk >> 1);
        if ((lane & 7) == 0) bsw[(size_t)pr * 8 + col] = (uint32_t)(s & 0xffff) | ((uint32_t)(s_next & 0xffff) << 16);
        if (lane == 0) dx[(size_t)blk * 8 + col] = d;
        const uint2 pk = pack8(q);
        uint32_t *dst = reinterpret_cast<uint32_t *>(img + ((size_t)pr * 32 + col * 4) * 16) + (sbk & 1) * 2 + (jj >> 1);
        dst[(2 * (jj & 1)) * 4] = pk.x;
        dst[(2 * (jj & 1) + 1) * 4] = pk.y;
    }
}

// Work schedule of one CTA.  The CTA owns a contiguous range of R tiles and walks it in "rounds": a round
// takes n = the largest power of two <= min(remaining, 8) tiles and gives each of them wpt = 16 / n warps,
// which split the tile's super-blocks (sb = wq, wq + wpt, ...).  Every round therefore costs each warp
// nsb * n / 16 units whatever R is (perfect balance inside the CTA), needs one CTA barrier, and its
// partial accumulators always fill exactly one 16 KB buffer.
struct WarpRound {
    int tile0, rem;        // first tile of the round, tiles left including this round
    int n, lw;             // tiles in this round, log2(warps per tile)
    int tile, wq, upr;     // this warp: its tile, its first super-block, its units in this round
};
__device__ __forceinline__ void round_setup(WarpRound &w, int warp, int nsb) {
    if (w.rem >= 8) { w.n = 8; w.lw = 1; } else if (w.rem >= 4) { w.n = 4; w.lw = 2; } else if (w.rem >= 2) { w.n = 2; w.lw = 3; } else { w.n = 1; w.lw = 4; }
    const int wpt = 1 << w.lw;
    w.tile = w.tile0 + (warp >> w.lw);
    w.wq = warp & (wpt - 1);
    w.upr = (w.rem > 0 && w.wq < nsb) ? (nsb - w.wq + wpt - 1) >> w.lw : 0;
}
__device__ __forceinline__ void round_next(WarpRound &w, int warp, int nsb) {
    w.tile0 += w.n; w.rem -= w.n;
    round_setup(w, warp, nsb);
}

// VAR: 0 = everything (heads, depformer_in, debug stamps), 1 = lean (store / residual / gate epilogues, image from the quantise
// kernel), 2 = lean with the fused short-K activation prologue.  One small body per kernel: consecutive launches of a frame
// alternate between variants, and kernel size costs instruction-cache misses (see gemv.cuh).
template <int WT, int VAR = 0>
__global__ void __launch_bounds__(kGemmThreads, 1) dq_matmul_mma_kernel(const MatmulArgs a) {
    constexpr bool LEAN = VAR != 0, FUSED = VAR == 2 || (VAR == 0 && WT == 12);
    constexpr int kUB = unit_bytes_of(WT);
    extern __shared__ __align__(16) uint8_t smem[];
    griddep_launch();
    long long *stamp = (!LEAN && a.stamps) ? a.stamps + (size_t)blockIdx.x * 8 : nullptr;
    if (stamp && threadIdx.x == 0) stamp[0] = global_ns();
    const int lane = threadIdx.x & 31, warp = uniform_warp_id();
    const int g = lane >> 2, t = lane & 3;
    const int K = a.w.K, nsb = a.w.nsb, S = a.stages;
    const int img_sz = act_image_bytes(K, WT);
    uint8_t *img = smem;
    double *part = reinterpret_cast<double *>(smem + ((img_sz + 127) & ~127));            // [2][16 warps][4][32]
    uint8_t *bar_base = reinterpret_cast<uint8_t *>(part) + kPartBytes;                 // mbarriers: [0] image, [1 + warp*4 + s]
    uint8_t *ring = bar_base + 1024 + (size_t)warp * S * kUB;
    const uint32_t bar_u32 = (uint32_t)__cvta_generic_to_shared(bar_base);
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t my_bar = bar_u32 + 8 + warp * (kGemmMaxStages * 8);

    const int t_begin = (int)((long long)blockIdx.x * a.w.n_tiles / gridDim.x);
    const int t_end = (int)((long long)(blockIdx.x + 1) * a.w.n_tiles / gridDim.x);

    if (lane == 0) {
        for (int s = 0; s < S; s++) mbar_init(my_bar + s * 8, 1);
        if (warp == 0) mbar_init(bar_u32, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // producer cursor (lane 0): walks the same schedule as the consumer, S units ahead
    WarpRound pc{t_begin, t_end - t_begin};
    round_setup(pc, warp, nsb);
    int pi = 0, ps = 0;                 // unit inside the round, ring slot
    const uint8_t *psrc = a.w.units + ((size_t)pc.tile * nsb + pc.wq) * kUB;
    auto issue = [&]() -> bool {       // lane 0 only; false when the warp's work list is exhausted
        if (pi >= pc.upr) {
            do { round_next(pc, warp, nsb); } while (pc.rem > 0 && pc.upr == 0);
            if (pc.rem <= 0) return false;
            pi = 0;
            psrc = a.w.units + ((size_t)pc.tile * nsb + pc.wq) * kUB;
        }
        mbar_expect_tx(my_bar + ps * 8, kUB);
        bulk_g2s(ring_u32 + ps * kUB, psrc, kUB, my_bar + ps * 8);
        psrc += (size_t)kUB << pc.lw;
        pi++;
        if (++ps == S) ps = 0;
        return true;
    };
    // weights never depend on the previous kernel: fill the ring before waiting for it (PDL)
    if (lane == 0) for (int u = 0; u < S; u++) if (!issue()) break;
    __syncthreads();                  // image barrier initialised
    if (stamp && threadIdx.x == 0) stamp[1] = global_ns();
    griddep_wait();
    if (stamp && threadIdx.x == 0) stamp[2] = global_ns();
    if (FUSED && (VAR == 2 || a.xsrc)) {
        fused_quant_columns(a, img, part, warp, lane);      // partial buffers double as the sum-of-squares scratch
        __syncthreads();
    } else {
        // activation image: 16 bulk copies (one per warp) so the L2 -> shared transfer is requested in parallel
        if (threadIdx.x == 0) mbar_expect_tx(bar_u32, (uint32_t)img_sz);
        if (lane == 0) {
            const int piece = ((img_sz / kGemmWarps) + 15) & ~15;
            const int off = warp * piece;
            const int len = min(piece, img_sz - off);
            if (len > 0) bulk_g2s((uint32_t)__cvta_generic_to_shared(img) + off, a.img + off, (uint32_t)len, bar_u32);
        }
    }
    // per-thread epilogue constants: as reducer this thread owns column 2t + (warp & 1)
    const int col = 2 * t + (warp & 1);
    const bool col_live = col < a.nb;
    float *const ocol = a.out + (size_t)col * a.ld;
    int emb_token = 0;
    if (!LEAN && a.epi == EPI_ADD_EMB && col_live) emb_token = depformer_prev_token(a.ctrl + col, a.emb_step);
    unsigned long long best = 0ull;
    if (!(FUSED && (VAR == 2 || a.xsrc))) mbar_wait(bar_u32, 0);
    if (stamp && threadIdx.x == 0) stamp[3] = global_ns();

    WarpRound cr{t_begin, t_end - t_begin};
    round_setup(cr, warp, nsb);
    int cs = 0; uint32_t cphase = 0;    // consumer ring slot and its mbarrier parity
    int rr = 0;
#pragma unroll 1
    for (; cr.rem > 0; round_next(cr, warp, nsb), rr++) {
        double *pbuf = part + (size_t)(rr & 1) * (kGemmWarps * 128);
        // reducer role of this warp in this round: two rotating warps per tile (parity = column parity)
        const int idx = (warp - ((rr * 2 * cr.n) & (kGemmWarps - 1))) & (kGemmWarps - 1);
        const bool reducer = idx < 2 * cr.n && col_live;
        const int red_tile = cr.tile0 + (idx >> 1), e = idx & 1;
        float old0 = 0.f, old1 = 0.f;
        if (reducer && a.epi == EPI_RESID) {      // residual: old values requested now, consumed after the round's dot products
            const int row0 = red_tile * 16 + g;
            if (row0 < a.w.rows) old0 = __ldcg(ocol + row0);
            if (row0 + 8 < a.w.rows) old1 = __ldcg(ocol + row0 + 8);
        }
        if (cr.upr > 0) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
            for (int i = 0; i < cr.upr; i++) {
                mbar_wait(my_bar + cs * 8, cphase);
                if (WT == 12) unit_compute(ring + (size_t)cs * kUB, img, K, cr.wq + (i << cr.lw), lane, acc);
                else unit_compute_q8(ring + (size_t)cs * kUB, img, K, cr.wq + (i << cr.lw), lane, acc);
                __syncwarp();
                if (lane == 0) issue();
                if (++cs == S) { cs = 0; cphase ^= 1; }
            }
            double *pw = pbuf + warp * 128 + lane;
            pw[0] = acc[0]; pw[32] = acc[1]; pw[64] = acc[2]; pw[96] = acc[3];
        }
        __syncthreads();
        if (reducer) {
            const int nw = min(1 << cr.lw, nsb);
            double s0 = 0.0, s1 = 0.0;
            const double *pj = pbuf + ((idx >> 1) << cr.lw) * 128 + lane;
            for (int j = 0; j < nw; j++, pj += 128) { s0 += pj[e * 32]; s1 += pj[(e + 2) * 32]; }
            const float v0 = (float)s0, v1 = (float)s1;
            const int row0 = red_tile * 16 + g, row1 = row0 + 8;
            if (a.epi == EPI_GATE) {
                const int h = red_tile * 8 + g;
                if (2 * h < a.w.rows) ocol[h] = (v0 / (1.0f + (float)exp((double)(-v0)))) * v1;
            } else if (a.epi == EPI_RESID) {
                if (row0 < a.w.rows) ocol[row0] = old0 + v0;
                if (row1 < a.w.rows) ocol[row1] = old1 + v1;
            } else if (LEAN) {
                if (row0 < a.w.rows) ocol[row0] = v0;
                if (row1 < a.w.rows) ocol[row1] = v1;
            } else {
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    const int row = hh ? row1 : row0;
                    const float v = hh ? v1 : v0;
                    if (row >= a.w.rows) continue;
                    if (a.epi == EPI_STORE) ocol[row] = v;
                    else if (a.epi == EPI_ARGMAX) {
                        ocol[row] = v;
                        const unsigned long long k = argmax_key(v, row);
                        best = k > best ? k : best;
                    } else if (a.epi == EPI_ADD_EMB) {
                        float em;
                        if (a.emb_step == 0) { em = emb_element(a.emb, emb_token < 0 ? 0 : emb_token, row); em = em * (emb_token == -1 ? 0.f : 1.f); }
                        else em = emb_element(a.emb, emb_token, row);
                        ocol[row] = v + em;
                    }
                }
            }
        }
    }
    if (stamp && threadIdx.x == 0) stamp[4] = global_ns();
    if (!LEAN && a.epi == EPI_ARGMAX) {
        // best over the 8 row groups of the warp (lanes sharing t), then over the warps of equal parity
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, best, o); best = x > best ? x : best; }
        unsigned long long *sbest = reinterpret_cast<unsigned long long *>(part);     // partial buffers are free after the last barrier
        __syncthreads();
        if (lane < 4) sbest[warp * 4 + lane] = best;
        __syncthreads();
        if (threadIdx.x < a.nb) {
            const int c = threadIdx.x, tt = c >> 1, ee = c & 1;
            unsigned long long bb = 0;
            for (int w = ee; w < kGemmWarps; w += 2) { const unsigned long long x = sbest[w * 4 + tt]; bb = x > bb ? x : bb; }
            if (bb) atomicMax(a.key_index < 0 ? &a.ctrl[c].text_key : &a.ctrl[c].audio_key[a.key_index], bb);
        }
    }
}

#define dq_matmul_q4k_kernel dq_matmul_mma_kernel<12, 0>
#define dq_matmul_q8_0_kernel dq_matmul_mma_kernel<8, 0>
__host__ inline int gemm_grid_for(int n_tiles, int num_sms) { return n_tiles < num_sms ? n_tiles : num_sms; }

}  // namespace msx
