// tc_gemm.cuh — Q4_K x Q8_K dequant-GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in tensor
// memory) for the dense contractions of the path: the batched-T prompt prefill (64 prompt positions per weight pass) and lock-step
// batches of 9..64 conversations (16 / 32 / 64 live columns).
//
// Replaces, for up to 64 columns at once, what the reference runs once per column: ggml_mul_mat of a Q4_K matrix with one
// Q8_K-quantised activation column (torch_nn_linear src/torch.h:79-87 -> ggml_vec_dot_q4_K_q8_K), T times inside
// moshi_lmgen_step_system_prompts / _voice_prompt (src/moshi/models/lm.h:983-1134), once per process and frame for concurrent streams.
//
// Exact arithmetic on tensor cores.  ggml's block dot is  sum_s sc_s * (sum_k q_k x_k)  with 6-bit sub-block scales sc_s: the
// integer part cannot be one int8 MMA because q * sc needs 10 bits.  Split sc = sc_lo + 8 sc_hi (3 bits each): q * sc_lo and
// q * sc_hi are <= 15 * 7 = 105 and fit s8, so TWO MMAs per super-block (K = 256) give the exact int32
//     isum = sum_k (q sc_lo)_k x_k + 8 sum_k (q sc_hi)_k x_k
// and everything after that is the arithmetic of gemv.cuh: mins term with dp2a, the two block terms d * dx * isum and
// dmin * dx * imin (fp32 scale products, exact in double) accumulated in double, one rounding at the end.
//
// A step = one tile of 128 weight rows x one super-block (256 weights) x all live columns: (a) the 128 raw GGUF blocks (18 KB,
// contiguous in the "tc layout" built at load) and the columns' 256 int8 activations go to shared memory by TMA, (b) the compute
// warps expand the blocks into the two s8 operand tiles in the canonical K-major no-swizzle layout (8-row x 16-byte core
// matrices), (c) one thread issues 16 tcgen05.mma (M 128, N = columns, K 32) into two accumulators and commits to an mbarrier,
// (d) sixteen warps read the accumulators back with tcgen05.ld (warp = 32 TMEM lanes x columns / 4) and fold them into their double
// accumulators.  The accumulators are double-buffered in tensor memory, so the tensor cores work on step i + 1 while the CUDA
// cores fold step i and expand step i + 2 (pipeline at the shared-memory map below).  The launch is persistent: one CTA per SM
// walks a contiguous range of the matrix's step sequence (stream-K, see TcMatmulArgs).  Descriptor encodings pinned by
// scripts/tcgen05_probe.cu; measurements and the optimisation log in profiles/r2_tcgen05_prefill.md.
#pragma once
#include <algorithm>
#include "common.cuh"
#include "gemv.cuh"      // depformer_prev_token (the embedding-add follow-up kernel)

namespace msx {
namespace tc {

constexpr int kM = 128, kN = 64;
constexpr int kComputeThreads = 512;                   // 16 warps expand and fold; warp 16 issues the MMAs, warp 17 the TMA copies
constexpr int kThreads = kComputeThreads + 64;
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
// shared memory.  The step of super-block i has four phases on three engines: (1) TMA brings the raw blocks and the activation
// record, (2) the compute warps expand the blocks into the two s8 operand tiles, (3) the tensor cores multiply, (4) the compute warps
// fold the accumulators into their doubles.  Warp-specialised: 16 compute warps, one MMA-issuing thread (warp 16), one TMA-issuing
// thread (warp 17) — both instruction kinds block their issuing thread for ~1.5 us per step, which a compute thread cannot afford.
// Steady state of iteration i:  MMA(i + 1) on the tensor cores  ||  fold(i) + expand(i + 2) on the CUDA cores  ||  raw(i + 3),
// raw(i + 4), record(i + 2), record(i + 3) in flight.  That takes two raw buffers, two operand buffers, two accumulator sets in tensor
// memory, three buffers of the activation record (tile + sums / scales, one TMA target) and four of the block headers (a compute
// warp may run one iteration ahead of another: they meet only at the mbarrier that releases iteration i's buffers).  One CTA per SM.
constexpr int kOffRaw = 0;                             // 2 x 18432 raw blocks
constexpr int kOffHdr = kOffRaw + 2 * kRawBytes;       // 4 x {d | dmin, scales[12]} of the 128 rows
constexpr int kOffAlo = kOffHdr + 4 * kM * 16;         // 2 x s8 [128][256]
constexpr int kOffAhi = kOffAlo + 2 * kOperandBytes;
constexpr int kOffB = kOffAhi + 2 * kOperandBytes;     // 3 x s8 [64][256]
constexpr int kOffAux = kOffB + 3 * kImgX8;            // 3 x {sums, scales}
constexpr int kOffMisc = kOffAux + 3 * kImgAux;        // mbarriers, TMEM base
constexpr int kSmemBytes = kOffMisc + 96;             // 9 mbarriers, TMEM base, ticket flag
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
constexpr int kTmemCols = 256;                         // set s at 128 s: [0, 64) sc_lo product, [64, 128) sc_hi product
// instruction descriptor, kind::i8 (cute/arch/mma_sm100_desc.hpp): D = S32, A = B = signed 8 bit, both K-major, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t idesc_for(int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kM >> 4) << 24); }

struct TcMatmulArgs {
    const uint8_t *w = nullptr;       // tc layout: [rows / 128][K / 256][128 rows][144 B]
    int32_t K = 0, rows = 0;          // stored rows (gate / up interleaved for EPI_GATE)
    const uint8_t *img = nullptr;     // activation image (see image_bytes)
    float *out = nullptr;             // out[col * ld + row]
    int32_t ld = 0, nb = 0, epi = 0;  // live columns, EPI_STORE / EPI_RESID / EPI_GATE
    // stream-K: the launch is one CTA per SM; the tiles x super-blocks steps of the matrix are one sequence (contiguous in the tc
    // layout) cut into gridDim.x equal ranges, so every SM does the same number of steps whatever the row count.  A tile whose
    // steps lie in several ranges is finished by the last CTA to arrive (ticket in `tickets[tile]`): every contributor leaves its
    // double partial sums in `partial` [cta][first / later tile of the range][64][128] and the finisher adds them in range order
    // — a fixed order, so the result does not depend on timing
    double *partial = nullptr;
    unsigned int *tickets = nullptr;
};
// CTAs of a launch.  More tiles than SMs: one range per SM (stream-K proper).  Fewer: either one range per SM again, or `p` aligned
// ranges per tile (plain split-K: every CTA leaves exactly one partial tile and no tile is cut at an odd place) — whichever is
// shorter by the measured costs of scripts/tc_gemm_probe.cu (2.1 us per step, ~3 us per partial tile left, ~0.7 us per part added up)
__host__ inline int grid_for(int n_tiles, int nsb, int num_sms) {
    const long long S = (long long)n_tiles * nsb;
    if (S <= num_sms) return (int)S;
    if (n_tiles >= num_sms) return num_sms;
    int p = num_sms / n_tiles;
    while (p > 1 && nsb % p) p--;
    const double step = 2.1, left = 3.0, add = 0.7;
    const double aligned = (nsb / p) * step + (p > 1 ? left + add * p : 0.0);
    const double spread = (double)((S + num_sms - 1) / num_sms);
    const double stream = spread * step + 2 * left + add * (nsb / spread + 2);
    return aligned <= stream ? n_tiles * p : num_sms;
}
__host__ inline size_t partial_bytes(int num_sms) { return (size_t)num_sms * 2 * kN * kM * sizeof(double); }

// shared-memory matrix descriptor: start >> 4 | LBO (128 B between the two core matrices of a K = 32 step) | SBO | version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
template <int N> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tmem_ld<4>(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
// The fold converts two integers and two fp32 scale products per (row, column, super-block) to double.  Integers go through the
// 2^52 trick (one XOR + one DADD on the 64-lane double-precision pipe); the products use F2F.F64.F32 (16 lanes / clk / SM, measured
// with scripts/fp64_rate_probe.cu) — building them with integer operations instead costs 8 issue slots each and measured 12 % slower
// once the fold was the only thing left on the critical path (scripts/tc_gemm_probe.cu).
__device__ __forceinline__ double int_to_double(int i) {
    return __hiloint2double(0x43300000, (int)((uint32_t)i ^ 0x80000000u)) - 4503601774854144.0;      // 2^52 + 2^31
}

#ifdef MSX_TC_TIMELINE      // scripts/tc_gemm_probe.cu: per-iteration stamps of CTA 0, thread 0 {loop top, accumulators ready, fold done, expand done}
__device__ long long *g_tc_timeline;
#define TC_STAMP(i, k) do { if (blockIdx.x == 0 && tid == 0 && (i) < 64) g_tc_timeline[(i) * 4 + (k)] = clock64(); } while (0)
#else
#define TC_STAMP(i, k) do { } while (0)
#endif

__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory"); }

// NC = live activation columns rounded up to 16 / 32 / 64: the MMA's N, the columns a thread folds (NC / 4) and the bytes of the
// activation tile a step copies all scale with it, so a batch of 16 streams does not pay for 64
template <int NC>
__global__ void __launch_bounds__(kThreads, 1) tc_matmul_q4k_kernel(const TcMatmulArgs a) {
    constexpr int FC = NC / 4;                        // columns per compute thread
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    TC_STAMP(60, 0);
    const uint32_t smem_u = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bar_mma = smem_u + kOffMisc /* [2] */, bar_raw = bar_mma + 16 /* [2] */, bar_b = bar_mma + 32 /* [3] */, bar_done = bar_mma + 56 /* [2] */;
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + kOffMisc + 72);
    unsigned int *flag = reinterpret_cast<unsigned int *>(smem + kOffMisc + 80);
    griddep_launch();
    if (tid == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(bar_mma + 8 * i, 1); mbar_init(bar_raw + 8 * i, 1); mbar_init(bar_done + 8 * i, kComputeThreads / 32); }
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

    const int K = a.K, nsb = K >> 8, G = (int)gridDim.x, cta = (int)blockIdx.x;
    const long long S = (long long)(a.rows / kM) * nsb;                   // steps of the whole matrix
    const int g0 = (int)(S * cta / G), total = (int)(S * (cta + 1) / G) - g0;       // this CTA: steps [g0, g0 + total), total >= 1
    const bool is_mma = tid == kComputeThreads, is_tma = tid == kComputeThreads + 32, is_compute = tid < kComputeThreads;
    auto issue_raw = [&](int it) {                    // weights do not depend on the predecessor kernel
        mbar_expect_tx(bar_raw + 8 * (it & 1), kRawBytes);
        bulk_g2s(smem_u + kOffRaw + (it & 1) * kRawBytes, a.w + (size_t)(g0 + it) * kRawBytes, kRawBytes, bar_raw + 8 * (it & 1));
    };
    auto issue_b = [&](int it) {
        const uint8_t *rec = a.img + (size_t)((g0 + it) % nsb) * kImgRec;
        const int b3 = it % 3;
        mbar_expect_tx(bar_b + 8 * b3, NC * 256 + kImgAux);
        bulk_g2s(smem_u + kOffB + b3 * kImgX8, rec, NC * 256, bar_b + 8 * b3);          // columns [0, NC): NC / 8 groups of 2048 B
        bulk_g2s(smem_u + kOffAux + b3 * kImgAux, rec + kImgX8, kImgAux, bar_b + 8 * b3);
    };
    // expansion of step `it`: thread = (row, 64-weight group): nibbles x 3-bit halves of the two sub-block scales -> s8 operand tiles
    const int er = tid & 127, ej = (tid >> 7) & 3;
    auto expand = [&](int it) {
        const int buf = it & 1;
        mbar_wait(bar_raw + 8 * buf, (uint32_t)((it >> 1) & 1));
        const uint8_t *blk = smem + kOffRaw + buf * kRawBytes + er * 144;
        const uint4 hd = *reinterpret_cast<const uint4 *>(blk);                             // {d | dmin, scales[12]}
        if (ej == 0) *reinterpret_cast<uint4 *>(smem + kOffHdr + (it & 3) * (kM * 16) + er * 16) = hd;
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
    // 2 x 8 MMAs of K = 32 on step `it` into accumulator set it & 1 (one thread; operands expanded and synchronised before)
    auto issue_mma = [&](int it) {
        const int buf = it & 1, b3 = it % 3;
        mbar_wait(bar_b + 8 * b3, (uint32_t)((it / 3) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t td = tmem + buf * 128;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            const uint64_t db = make_desc(smem_u + kOffB + b3 * kImgX8 + ks * 256);
            mma_i8(td, make_desc(smem_u + kOffAlo + buf * kOperandBytes + ks * 256), db, idesc_for(NC), ks > 0 ? 1u : 0u);
            mma_i8(td + kN, make_desc(smem_u + kOffAhi + buf * kOperandBytes + ks * 256), db, idesc_for(NC), ks > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma + 8 * buf) : "memory");
    };
    // ---- start-up (all warps): two steps expanded, the first multiplication under way ----
    if (is_tma) { issue_raw(0); if (total > 1) issue_raw(1); }
    griddep_wait();                                   // the activation image comes from the quantise kernel (PDL)
    if (is_tma) for (int i = 0; i < 3 && i < total; i++) issue_b(i);
    if (is_compute) expand(0);
    __syncthreads();
    if (is_tma && total > 2) issue_raw(2);
    if (is_mma) issue_mma(0);
    if (is_compute && total > 1) expand(1);
    __syncthreads();
    if (is_tma && total > 3) issue_raw(3);

    if (is_mma) {
        // ---- tensor-core issuer: step it + 1 as soon as the compute warps are through iteration it - 1 (operands of it + 1 expanded,
        //      accumulator set of it - 1 folded) ----
        for (int it = 0; it + 1 < total; it++) {
            if (it >= 1) mbar_wait(bar_done + 8 * ((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1));
            issue_mma(it + 1);
        }
    } else if (is_tma) {
        // ---- loader: the buffers iteration `it` released take the record of step it + 3 and the raw blocks of step it + 4 ----
        for (int it = 0; it + 3 < total; it++) {
            mbar_wait(bar_done + 8 * (it & 1), (uint32_t)((it >> 1) & 1));
            issue_b(it + 3);
            if (it + 4 < total) issue_raw(it + 4);
        }
    } else if (is_compute) {
        const int q = warp & 3, cg = warp >> 2;       // fold: TMEM lane quadrant, group of FC columns
        const int row = q * 32 + lane;
        double acc[FC];
#pragma unroll
        for (int c = 0; c < FC; c++) acc[c] = 0.0;
        // block terms in double (arithmetic of gemv.cuh compute_step) from accumulator set it & 1
        auto fold = [&](int it) {
            const int buf = it & 1, b3 = it % 3;
            const uint4 hd = *reinterpret_cast<const uint4 *>(smem + kOffHdr + (it & 3) * (kM * 16) + row * 16);
            const uint32_t m_lo = hd.z & 0x3f3f3f3fu, m_hi = ((hd.w >> 4) & 0x0f0f0f0fu) | ((hd.z >> 2) & 0x30303030u);
            const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&hd.x));
            const uint8_t *aux = smem + kOffAux + b3 * kImgAux;
            const int4 *bs_s = reinterpret_cast<const int4 *>(aux);
            const float *dx_s = reinterpret_cast<const float *>(aux + kN * 16);
            uint32_t plo[FC], phi[FC];
            const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + buf * 128 + cg * FC;
            tmem_ld<FC>(ta, plo);
            tmem_ld<FC>(ta + kN, phi);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < FC; c++) {
                const int col = cg * FC + c;
                const int4 b4 = bs_s[col];
                const float dxv = dx_s[col];
                const int isum = (int)plo[c] + 8 * (int)phi[c];
                int imin = __dp2a_lo(b4.x, (int)m_lo, 0);
                imin = __dp2a_hi(b4.y, (int)m_lo, imin);
                imin = __dp2a_lo(b4.z, (int)m_hi, imin);
                imin = __dp2a_hi(b4.w, (int)m_hi, imin);
                acc[c] = fma((double)(dm.x * dxv), int_to_double(isum), acc[c]);
                acc[c] = fma(-(double)(dm.y * dxv), int_to_double(imin), acc[c]);
            }
        };
        // range of the step sequence that holds step g: the largest c with floor(S c / G) <= g
        auto owner = [&](long long g) { return (int)(((g + 1) * G - 1) / S); };
        // end of this CTA's steps of `tile`: finish it (alone, or as the last of its contributors) or leave the partial sums
        auto flush = [&](int tile) {
            const int c_first = owner((long long)tile * nsb), c_last = owner((long long)tile * nsb + nsb - 1);
            if (c_first != c_last) {
                double *mine = a.partial + ((size_t)cta * 2 + (tile != g0 / nsb ? 1 : 0)) * (kN * kM) + (cg * FC) * kM + row;
#pragma unroll
                for (int c = 0; c < FC; c++) __stcg(mine + c * kM, acc[c]);
                compute_barrier();
                if (tid == 0) { __threadfence(); *flag = atomicAdd(a.tickets + tile, 1u); }
                compute_barrier();
                if (*flag != (unsigned)(c_last - c_first)) return;           // another contributor finishes this tile
                __threadfence();
#pragma unroll
                for (int c = 0; c < FC; c++) acc[c] = 0.0;
                auto part_of = [&](int cc) {
                    return a.partial + ((size_t)cc * 2 + (tile != (int)(S * cc / G) / nsb ? 1 : 0)) * (kN * kM) + (cg * FC) * kM + row;
                };
                for (int cc = c_first; cc <= c_last; cc += 2) {               // range order = super-block order; two ranges' loads in flight
                    const double *s0 = part_of(cc), *s1 = part_of(cc + 1 <= c_last ? cc + 1 : cc);
                    double t0[FC], t1[FC];
#pragma unroll
                    for (int c = 0; c < FC; c++) { t0[c] = __ldcg(s0 + c * kM); t1[c] = __ldcg(s1 + c * kM); }
                    const bool two = cc + 1 <= c_last;
#pragma unroll
                    for (int c = 0; c < FC; c++) { acc[c] += t0[c]; if (two) acc[c] += t1[c]; }
                }
                if (tid == 0) a.tickets[tile] = 0u;                            // ready for the next launch
            }
            const int grow = tile * kM + row;
            if (a.epi == EPI_GATE) {
                // (gate, up) rows are adjacent lanes.  The even lane finishes the first half of the pair's columns, the odd lane the
                // second half, so every lane evaluates FC / 2 double-precision exponentials instead of FC with half of the lanes idle
#pragma unroll
                for (int c = 0; c < FC / 2; c++) {
                    const float lo = (float)acc[c], hi = (float)acc[c + FC / 2];
                    const float recv = __shfl_xor_sync(0xffffffffu, (lane & 1) ? lo : hi, 1);
                    const float gte = (lane & 1) ? recv : lo, up = (lane & 1) ? hi : recv;
                    const int col = cg * FC + c + ((lane & 1) ? FC / 2 : 0);
                    if (col < a.nb) a.out[(size_t)col * a.ld + (grow >> 1)] = (gte / (1.0f + (float)exp((double)(-gte)))) * up;
                }
            } else {
#pragma unroll
                for (int c = 0; c < FC; c++) {
                    const int col = cg * FC + c;
                    if (col < a.nb) {
                        float *o = a.out + (size_t)col * a.ld + grow;
                        const float v = (float)acc[c];
                        *o = a.epi == EPI_RESID ? *o + v : v;
                    }
                }
            }
        };
        TC_STAMP(60, 1);
        for (int it = 0; it < total; it++) {
            const int buf = it & 1;
            TC_STAMP(it, 0);
            mbar_wait(bar_mma + 8 * buf, (uint32_t)((it >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            TC_STAMP(it, 1);
            // fold step `it`; operands of step it + 2 into the buffers step `it` has released.  Half of the warps take the two in the
            // other order, so that the conversion / double-precision pipes and the integer / shared-memory pipes are busy at the same time
            if (cg & 1) {
                if (it + 2 < total) expand(it + 2);
                fold(it);
            } else {
                fold(it);
                TC_STAMP(it, 2);
                if (it + 2 < total) expand(it + 2);
            }
            TC_STAMP(it, 3);
            // this warp is through iteration `it` (no block barrier: the warps drift apart, which spreads their fold and expand phases
            // over time); when all 16 have arrived the accumulator set and record buffer of step `it` and the raw buffer of step
            // it + 2 are free, and the operands of step it + 2 are complete
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_done + 8 * buf);
            const int g = g0 + it;
            if (g % nsb == nsb - 1 || it == total - 1) {
                TC_STAMP(61 + (it == total - 1 ? 1 : 0), 0);
                flush(g / nsb);
                TC_STAMP(61 + (it == total - 1 ? 1 : 0), 1);
#pragma unroll
                for (int c = 0; c < FC; c++) acc[c] = 0.0;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_STAMP(60, 2);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// ---- the two epilogues of the batched decode step that the GEMM does not carry (wide batches only: 9 + dep_q launches per frame) ----
// arg-max key of every live column's logits row (EPI_ARGMAX of gemv.cuh / mma_gemm.cuh: first maximum wins)
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float *logits, int ld, int n, Ctrl *ctrl, int key_index) {
    griddep_launch();
    griddep_wait();
    const float *row = logits + (size_t)blockIdx.x * ld;
    unsigned long long best = 0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const unsigned long long k = argmax_key(__ldcg(row + i), i); best = k > best ? k : best; }
#pragma unroll
    for (int o = 16; o; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, best, o); best = x > best ? x : best; }
    __shared__ unsigned long long sb[8];
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) best = sb[w] > best ? sb[w] : best;
        Ctrl *c = ctrl + blockIdx.x;
        atomicMax(key_index < 0 ? &c->text_key : &c->audio_key[key_index], best);
    }
}
// depformer step input of every live column: x[col] += embedding of the column's previous token (EPI_ADD_EMB; lm.h:464-467, 494-516)
__global__ void __launch_bounds__(256) dep_embed_add_cols_kernel(const Ctrl *ctrl, const EmbTable emb, int step, float *x, int ld, int n) {
    griddep_launch();
    griddep_wait();
    const int token = depformer_prev_token(ctrl + blockIdx.y, step);
    float *xc = x + (size_t)blockIdx.y * ld;
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
        float e;
        if (step == 0) {
            e = emb_element(emb, token < 0 ? 0 : token, row);
            e = e * (token == -1 ? 0.f : 1.f);
        } else e = emb_element(emb, token, row);
        xc[row] = __ldcg(xc + row) + e;
    }
}

__host__ inline int columns_for(int nb) { return nb <= 16 ? 16 : nb <= 32 ? 32 : 64; }
__host__ inline const void *kernel_for(int nc) {
    return nc == 16 ? (const void *)tc_matmul_q4k_kernel<16> : nc == 32 ? (const void *)tc_matmul_q4k_kernel<32> : (const void *)tc_matmul_q4k_kernel<64>;
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
