// tc_gemm.cuh — Q4_K x Q8_K dequant-GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in tensor
// memory) for the dense contraction of the path: the batched-T prompt prefill, 64 prompt positions per weight pass.
//
// Replaces, for T columns at once, what the reference runs T times: ggml_mul_mat of a Q4_K matrix with one Q8_K-quantised
// activation column (torch_nn_linear src/torch.h:79-87 -> ggml_vec_dot_q4_K_q8_K) inside moshi_lmgen_step_system_prompts /
// _voice_prompt (src/moshi/models/lm.h:983-1134).
//
// Exact arithmetic on tensor cores.  ggml's block dot is  sum_s sc_s * (sum_k q_k x_k)  with 6-bit sub-block scales sc_s: the
// integer part cannot be one int8 MMA because q * sc needs 10 bits.  Split sc = sc_lo + 8 sc_hi (3 bits each): q * sc_lo and
// q * sc_hi are <= 15 * 7 = 105 and fit s8, so TWO MMAs per super-block (K = 256) give the exact int32
//     isum = sum_k (q sc_lo)_k x_k + 8 sum_k (q sc_hi)_k x_k
// and everything after that is the arithmetic of gemv.cuh: mins term with dp2a, the two block terms d * dx * isum and
// dmin * dx * imin (fp32 scale products, exact in double) accumulated in double, one rounding at the end.
//
// One CTA = one tile of 128 weight rows x all 64 columns; per super-block: (a) the 128 raw GGUF blocks (18 KB, contiguous in the
// "tc layout" built at load) and the 64 x 256 int8 activations go to shared memory by TMA, (b) every thread expands half a block into
// the two s8 operand tiles in the canonical K-major no-swizzle layout (8-row x 16-byte core matrices), (c) one thread issues 16
// tcgen05.mma (M 128, N 64, K 32) into two 64-column accumulators and commits to an mbarrier, (d) sixteen warps read the
// accumulators back with tcgen05.ld (warp = 32 TMEM lanes x 16 columns) and fold them into 16 double accumulators per thread.
// The accumulators are double-buffered in tensor memory, so the tensor cores work on super-block i + 1 while the CUDA cores fold
// super-block i and expand super-block i + 2 (pipeline at the shared-memory map below).  Descriptor encodings pinned by
// scripts/tcgen05_probe.cu.
#pragma once
#include "common.cuh"

namespace msx {
namespace tc {

constexpr int kM = 128, kN = 64, kThreads = 512;
constexpr int kRawBytes = kM * 144;                    // 128 Q4_K blocks of one super-block column
constexpr int kOperandBytes = kM * 256;                // s8 [128][256]
constexpr int kSBO = (256 / 16) * 128;                 // bytes between 8-row groups of an operand tile
// activation image (quant_q8k_kernel with QuantArgs::plain): one record per super-block,
//   s8 x8 [64 cols][256] already in the canonical operand layout (16384 B) | int16 sums per 32 [64][8] (1024 B) | f32 scale [64] (256 B)
constexpr int kImgX8 = kN * 256, kImgAux = kN * 16 + kN * 4, kImgRec = kImgX8 + kImgAux;
__host__ __device__ inline size_t image_bytes(int K) { return (size_t)(K >> 8) * kImgRec; }
__host__ __device__ inline size_t image_x8_offset(int col, int k) {      // byte of activation k of column col
    return (size_t)(k >> 8) * kImgRec + (size_t)(col >> 3) * kSBO + (size_t)((k & 255) >> 4) * 128 + (size_t)(col & 7) * 16 + (k & 15);
}
// shared memory.  The step of super-block i has three phases on three engines: (1) TMA brings the raw blocks and the activation
// record, (2) all threads expand the blocks into the two s8 operand tiles, (3) the tensor cores multiply, (4) all threads fold the
// accumulators into their doubles.  Steady state of iteration i:  MMA(i + 1) on the tensor cores  ||  fold(i) + expand(i + 2) on the
// CUDA cores  ||  raw(i + 3), raw(i + 4), record(i + 2), record(i + 3) in flight.  That takes two raw buffers, two operand buffers,
// two accumulator sets in tensor memory, and three buffers of what the fold reads (block headers, activation sums / scales) and of
// the activation tile (its record is one TMA target).  One CTA per SM.
constexpr int kOffRaw = 0;                             // 2 x 18432 raw blocks
constexpr int kOffHdr = kOffRaw + 2 * kRawBytes;       // 3 x {d | dmin, scales[12]} of the 128 rows
constexpr int kOffAlo = kOffHdr + 3 * kM * 16;         // 2 x s8 [128][256]
constexpr int kOffAhi = kOffAlo + 2 * kOperandBytes;
constexpr int kOffB = kOffAhi + 2 * kOperandBytes;     // 3 x s8 [64][256]
constexpr int kOffAux = kOffB + 3 * kImgX8;            // 3 x {sums, scales}
constexpr int kOffMisc = kOffAux + 3 * kImgAux;        // mbarriers, TMEM base
constexpr int kSmemBytes = kOffMisc + 96;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
constexpr int kTmemCols = 256;                         // set s at 128 s: [0, 64) sc_lo product, [64, 128) sc_hi product
// instruction descriptor, kind::i8 (cute/arch/mma_sm100_desc.hpp): D = S32, A = B = signed 8 bit, both K-major, N >> 3, M >> 4
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);

struct TcGemmArgs {
    const uint8_t *w = nullptr;       // tc layout: [rows / 128][K / 256][128 rows][144 B]
    int32_t K = 0, rows = 0;          // stored rows (gate / up interleaved for EPI_GATE)
    const uint8_t *img = nullptr;     // activation image (see image_bytes)
    float *out = nullptr;             // out[col * ld + row]
    int32_t ld = 0, nb = 0, epi = 0;  // live columns, EPI_STORE / EPI_RESID / EPI_GATE
    // split-K: `parts` CTAs share a tile (each takes a range of super-blocks) so that every SM holds two CTAs also for the
    // 4096-row matrices; they leave their double partial sums in `partial` [tile][part][64][128] and the last one to arrive
    // (ticket in `tickets[tile]`) adds them in part order — a fixed order, so the result does not depend on timing
    int32_t parts = 1;
    double *partial = nullptr;
    unsigned int *tickets = nullptr;
};
// parts that minimise (waves of one CTA per SM) x (fixed cost + super-blocks per CTA), plus the cost of the ordered reduction
__host__ inline int parts_for(int n_tiles, int nsb, int num_sms) {
    int best = 1; double best_t = 1e30;
    for (int p = 1; p <= 8 && nsb / p >= 2; p++) {
        const int ctas = n_tiles * p, waves = (ctas + num_sms - 1) / num_sms, per = (nsb + p - 1) / p;
        const double t = waves * (4.0 + per * 2.2) + (p > 1 ? 3.0 + 0.5 * p : 0.0);
        if (t < best_t) { best_t = t; best = p; }
    }
    return best;
}

// shared-memory matrix descriptor: start >> 4 | LBO (128 B between the two core matrices of a K = 32 step) | SBO | version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
// The epilogue converts two integers and two fp32 scale products per (row, column, super-block) to double; the conversion
// instructions run on the 8-lane XU pipe and would bound the kernel.  Integers go through the 2^52 trick (one XOR + one DADD),
// the d * dx product is widened with integer operations when it is a normal number (else converted), only dmin * dx uses F2F.
__device__ __forceinline__ double int_to_double(int i) {
    return __hiloint2double(0x43300000, (int)((uint32_t)i ^ 0x80000000u)) - 4503601774854144.0;      // 2^52 + 2^31
}
__device__ __forceinline__ double widen_f32(float p) {
    const uint32_t b = __float_as_uint(p), e = b & 0x7f800000u;
    if (e == 0u || e == 0x7f800000u) return (double)p;                                              // zero, denormal, inf, nan
    return __hiloint2double((int)((((b & 0x7fffffffu) >> 3) + 0x38000000u) | (b & 0x80000000u)), (int)(b << 29));
}

__global__ void __launch_bounds__(kThreads, 1) tc_gemm_q4k_kernel(const TcGemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_u = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bar_mma = smem_u + kOffMisc /* [2] */, bar_raw = bar_mma + 16 /* [2] */, bar_b = bar_mma + 32 /* [3] */;
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + kOffMisc + 64);
    griddep_launch();
    if (tid == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(bar_mma + 8 * i, 1); mbar_init(bar_raw + 8 * i, 1); }
        for (int i = 0; i < 3; i++) mbar_init(bar_b + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;

    const int K = a.K, nsb = K >> 8, P = a.parts;
    const int tile = blockIdx.x / P, part = blockIdx.x - tile * P;
    const int sb0 = (int)((long long)nsb * part / P), total = (int)((long long)nsb * (part + 1) / P) - sb0;   // super-blocks of this CTA
    auto issue_raw = [&](int it) {                    // weights do not depend on the predecessor kernel
        mbar_expect_tx(bar_raw + 8 * (it & 1), kRawBytes);
        bulk_g2s(smem_u + kOffRaw + (it & 1) * kRawBytes, a.w + ((size_t)tile * nsb + sb0 + it) * kRawBytes, kRawBytes, bar_raw + 8 * (it & 1));
    };
    auto issue_b = [&](int it) {
        const uint8_t *rec = a.img + (size_t)(sb0 + it) * kImgRec;
        const int b3 = it % 3;
        mbar_expect_tx(bar_b + 8 * b3, kImgRec);
        bulk_g2s(smem_u + kOffB + b3 * kImgX8, rec, kImgX8, bar_b + 8 * b3);
        bulk_g2s(smem_u + kOffAux + b3 * kImgAux, rec + kImgX8, kImgAux, bar_b + 8 * b3);
    };
    // expansion of step `it`: thread = (row, 64-weight group): nibbles x 3-bit halves of the two sub-block scales -> s8 operand tiles
    const int er = tid & 127, ej = tid >> 7;
    auto expand = [&](int it) {
        const int buf = it & 1;
        mbar_wait(bar_raw + 8 * buf, (uint32_t)((it >> 1) & 1));
        const uint8_t *blk = smem + kOffRaw + buf * kRawBytes + er * 144;
        const uint4 hd = *reinterpret_cast<const uint4 *>(blk);                             // {d | dmin, scales[12]}
        if (ej == 0) *reinterpret_cast<uint4 *>(smem + kOffHdr + (it % 3) * (kM * 16) + er * 16) = hd;
        const uint32_t sc_lo = hd.y & 0x3f3f3f3fu, sc_hi = (hd.w & 0x0f0f0f0fu) | ((hd.y >> 2) & 0x30303030u);   // get_scale_min_k4
        uint8_t *alo = smem + kOffAlo + buf * kOperandBytes + (er >> 3) * kSBO + (er & 7) * 16;
        uint8_t *ahi = smem + kOffAhi + buf * kOperandBytes + (er >> 3) * kSBO + (er & 7) * 16;
        const int j = ej;                                            // low nibbles = sub-block 2j, high = 2j + 1
        const uint32_t scw = j < 2 ? sc_lo : sc_hi;
        const uint32_t sa = (scw >> ((j & 1) * 16)) & 0xffu, sb2 = (scw >> ((j & 1) * 16 + 8)) & 0xffu;
        const uint32_t la = sa & 7u, ha = sa >> 3, lb = sb2 & 7u, hb = sb2 >> 3;
        const uint4 w0 = *reinterpret_cast<const uint4 *>(blk + 16 + 32 * j), w1 = *reinterpret_cast<const uint4 *>(blk + 32 + 32 * j);
        const uint32_t lo[8] = {w0.x & 0x0f0f0f0fu, w0.y & 0x0f0f0f0fu, w0.z & 0x0f0f0f0fu, w0.w & 0x0f0f0f0fu,
                                w1.x & 0x0f0f0f0fu, w1.y & 0x0f0f0f0fu, w1.z & 0x0f0f0f0fu, w1.w & 0x0f0f0f0fu};
        const uint32_t hi[8] = {(w0.x >> 4) & 0x0f0f0f0fu, (w0.y >> 4) & 0x0f0f0f0fu, (w0.z >> 4) & 0x0f0f0f0fu, (w0.w >> 4) & 0x0f0f0f0fu,
                                (w1.x >> 4) & 0x0f0f0f0fu, (w1.y >> 4) & 0x0f0f0f0fu, (w1.z >> 4) & 0x0f0f0f0fu, (w1.w >> 4) & 0x0f0f0f0fu};
        // bytes <= 15 * 7: the packed multiply never carries between bytes.  16-byte piece p holds k in [16 p, 16 p + 16)
        *reinterpret_cast<uint4 *>(alo + (4 * j + 0) * 128) = make_uint4(lo[0] * la, lo[1] * la, lo[2] * la, lo[3] * la);
        *reinterpret_cast<uint4 *>(alo + (4 * j + 1) * 128) = make_uint4(lo[4] * la, lo[5] * la, lo[6] * la, lo[7] * la);
        *reinterpret_cast<uint4 *>(alo + (4 * j + 2) * 128) = make_uint4(hi[0] * lb, hi[1] * lb, hi[2] * lb, hi[3] * lb);
        *reinterpret_cast<uint4 *>(alo + (4 * j + 3) * 128) = make_uint4(hi[4] * lb, hi[5] * lb, hi[6] * lb, hi[7] * lb);
        *reinterpret_cast<uint4 *>(ahi + (4 * j + 0) * 128) = make_uint4(lo[0] * ha, lo[1] * ha, lo[2] * ha, lo[3] * ha);
        *reinterpret_cast<uint4 *>(ahi + (4 * j + 1) * 128) = make_uint4(lo[4] * ha, lo[5] * ha, lo[6] * ha, lo[7] * ha);
        *reinterpret_cast<uint4 *>(ahi + (4 * j + 2) * 128) = make_uint4(hi[0] * hb, hi[1] * hb, hi[2] * hb, hi[3] * hb);
        *reinterpret_cast<uint4 *>(ahi + (4 * j + 3) * 128) = make_uint4(hi[4] * hb, hi[5] * hb, hi[6] * hb, hi[7] * hb);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy stores -> visible to the tensor-core proxy
    };
    // 2 x 8 MMAs of K = 32 on step `it` into accumulator set it & 1 (one thread; operands expanded and block-synchronised before)
    auto issue_mma = [&](int it) {
        const int buf = it & 1, b3 = it % 3;
        mbar_wait(bar_b + 8 * b3, (uint32_t)((it / 3) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t td = tmem + buf * 128;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            const uint64_t db = make_desc(smem_u + kOffB + b3 * kImgX8 + ks * 256);
            mma_i8(td, make_desc(smem_u + kOffAlo + buf * kOperandBytes + ks * 256), db, ks > 0 ? 1u : 0u);
            mma_i8(td + kN, make_desc(smem_u + kOffAhi + buf * kOperandBytes + ks * 256), db, ks > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma + 8 * buf) : "memory");
    };
    if (tid == 0) { issue_raw(0); if (total > 1) issue_raw(1); }
    griddep_wait();                                   // the activation image comes from the quantise kernel (PDL)
    if (tid == 0) for (int i = 0; i < 3 && i < total; i++) issue_b(i);
    expand(0);
    __syncthreads();
    if (tid == 0) { if (total > 2) issue_raw(2); issue_mma(0); }
    if (total > 1) expand(1);
    __syncthreads();
    if (tid == 0 && total > 3) issue_raw(3);

    const int q = warp & 3, cg = warp >> 2;           // fold: TMEM lane quadrant, group of 16 columns
    const int row = q * 32 + lane;
    double acc[16];
#pragma unroll
    for (int c = 0; c < 16; c++) acc[c] = 0.0;

    for (int it = 0; it < total; it++) {
        const int buf = it & 1, b3 = it % 3;
        if (tid == 0 && it + 1 < total) issue_mma(it + 1);       // runs while this iteration folds step `it` and expands step it + 2
        mbar_wait(bar_mma + 8 * buf, (uint32_t)((it >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- accumulators -> registers, block terms in double (arithmetic of gemv.cuh compute_step) ----
        {
            const uint4 hd = *reinterpret_cast<const uint4 *>(smem + kOffHdr + b3 * (kM * 16) + row * 16);
            const uint32_t m_lo = hd.z & 0x3f3f3f3fu, m_hi = ((hd.w >> 4) & 0x0f0f0f0fu) | ((hd.z >> 2) & 0x30303030u);
            const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&hd.x));
            const uint8_t *aux = smem + kOffAux + b3 * kImgAux;
            const int4 *bs_s = reinterpret_cast<const int4 *>(aux);
            const float *dx_s = reinterpret_cast<const float *>(aux + kN * 16);
            uint32_t plo[16], phi[16];
            const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + buf * 128 + cg * 16;
            tmem_ld16(ta, plo);
            tmem_ld16(ta + kN, phi);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const int col = cg * 16 + c;
                const int4 b4 = bs_s[col];
                const float dxv = dx_s[col];
                const int isum = (int)plo[c] + 8 * (int)phi[c];
                int imin = __dp2a_lo(b4.x, (int)m_lo, 0);
                imin = __dp2a_hi(b4.y, (int)m_lo, imin);
                imin = __dp2a_lo(b4.z, (int)m_hi, imin);
                imin = __dp2a_hi(b4.w, (int)m_hi, imin);
                acc[c] = fma(widen_f32(dm.x * dxv), int_to_double(isum), acc[c]);
                acc[c] = fma(-(double)(dm.y * dxv), int_to_double(imin), acc[c]);
            }
        }
        // ---- operands of step it + 2 into the buffers step `it` has released ----
        if (it + 2 < total) expand(it + 2);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                              // accumulator set and record buffer of step `it`, raw buffer of step it + 2: free
        if (tid == 0) {
            if (it + 3 < total) issue_b(it + 3);
            if (it + 4 < total) issue_raw(it + 4);
        }
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");

    // ---- tile epilogue: thread = row, 32 columns ----
    if (P > 1) {
        double *mine = a.partial + ((size_t)tile * P + part) * (kN * kM);
#pragma unroll
        for (int c = 0; c < 16; c++) mine[(cg * 16 + c) * kM + row] = acc[c];
        __threadfence();
        __syncthreads();
        unsigned int *flag = reinterpret_cast<unsigned int *>(smem + kOffMisc + 72);
        if (tid == 0) *flag = atomicAdd(a.tickets + tile, 1u);
        __syncthreads();
        if (*flag != (unsigned)(P - 1)) return;       // not the last part of this tile
        __threadfence();
        const double *all = a.partial + (size_t)tile * P * (kN * kM) + (cg * 16) * kM + row;
#pragma unroll
        for (int c = 0; c < 16; c++) acc[c] = 0.0;
        for (int pp = 0; pp < P; pp++) {              // part order; the 16 loads of a part are independent and in flight together
            double t[16];
#pragma unroll
            for (int c = 0; c < 16; c++) t[c] = __ldcg(all + (size_t)pp * (kN * kM) + c * kM);
#pragma unroll
            for (int c = 0; c < 16; c++) acc[c] += t[c];
        }
        if (tid == 0) a.tickets[tile] = 0u;           // ready for the next launch
    }
    const int grow = tile * kM + row;
#pragma unroll
    for (int c = 0; c < 16; c++) {
        const int col = cg * 16 + c;
        const float v = (float)acc[c];
        if (a.epi == EPI_GATE) {
            const float u = __shfl_down_sync(0xffffffffu, v, 1);                                     // (gate, up) rows are adjacent
            if (col < a.nb && (lane & 1) == 0) a.out[(size_t)col * a.ld + (grow >> 1)] = (v / (1.0f + (float)exp((double)(-v)))) * u;
        } else if (col < a.nb) {
            float *o = a.out + (size_t)col * a.ld + grow;
            *o = a.epi == EPI_RESID ? *o + v : v;
        }
    }
}

// GGUF row-major Q4_K blocks -> tc layout.  One thread per 16-byte piece; perm_half > 0 interleaves rows for the gated MLP
// (stored row v <- source row (v & 1 ? perm_half + v / 2 : v / 2)), like the other repack kernels.
__global__ void tc_layout_kernel(const uint8_t *src, uint8_t *dst, int rows, int nsb, int perm_half) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * nsb * 9) return;
    const int piece = (int)(idx % 9);
    const long long blk = idx / 9;
    const int sb = (int)(blk % nsb), v = (int)(blk / nsb);
    const int srow = perm_half > 0 ? ((v & 1) ? perm_half + (v >> 1) : (v >> 1)) : v;
    const uint4 t = *reinterpret_cast<const uint4 *>(src + ((size_t)srow * nsb + sb) * 144 + piece * 16);
    *reinterpret_cast<uint4 *>(dst + (((size_t)(v / kM) * nsb + sb) * kM + (v % kM)) * 144 + piece * 16) = t;
}

}  // namespace tc
}  // namespace msx
