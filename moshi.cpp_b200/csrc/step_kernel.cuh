// step_kernel.cuh — one transformer stack of the decode step as ONE persistent kernel (sm_100a).
//
// Replaces, per frame, the reference's two ggml graphs
//     moshi_lmmodel_forward_text_build/_step          src/moshi/models/lm.h:659-690  (temporal transformer + text head)
//     moshi_lmmodel_depformer_step                    src/moshi/models/lm.h:446-553  (dep_q codebook steps)
// i.e. what the PDL-chained path (gemv.cuh / attention.cuh) runs as 5 L + 3 and dep_q (4 Ld + 2) + 2 launches.
//
// Why: round 1 measured the single-stream frame as latency-structured — every one of the 373 dependent launches paid
// ~3 us of hand-over (tail of the producer kernel, activation prologue repeated by every CTA, ramp of the weight stream)
// around 2-8 us of HBM streaming.  Here a stack is a list of *phases* executed by 148 co-resident CTAs:
//
//   * one PRODUCER warp per CTA walks the whole phase list ahead of the consumers and streams this CTA's share of every
//     weight matrix through a 8-slot (144 KB) shared-memory ring with cp.async.bulk (TMA) + mbarrier completion.  Weights
//     do not depend on activations, so the HBM stream never stops at a phase boundary: while the consumers exchange
//     activations the ring absorbs ~3.7 us worth of the next matrices.
//   * 16 CONSUMER warps: lane = weight row, unit = (<= 32 rows) x (one 256-weight super-block), x broadcast from shared
//     memory; ~0.8 instructions per weight instead of 2.9 (no per-lane cp.async issue, scales decoded once per
//     super-block, no cross-lane reduction) — the consumers drain the ring ~3x faster than HBM fills it.
//   * activations move between CTAs as 8-byte {value, sequence} words ("LL" protocol): the flag travels with the
//     datum, so a consumer that needs a vector simply polls the vector — no grid barrier, no fence, no kernel boundary.
//     sequence = launch epoch << 12 | producing phase + 1; a buffer is only rewritten by a CTA that has since consumed
//     a full vector which transitively depends on every reader of the old generation.
//
// Arithmetic is exactly the reference-faithful arithmetic of gemv.cuh / attention.cuh (Q8_K / Q8_0 activation
// re-quantisation, exact integer block dots, block terms accumulated in double, bf16 ring cache, softmax normalised
// before the bf16 rounding), so the oracle parity tests hold bit for bit.
#pragma once
#include "common.cuh"

namespace msx {
namespace sk {

// ---- geometry ------------------------------------------------------------------------------------------------------
constexpr int kConsumerWarps = 16;
constexpr int kConsumers = kConsumerWarps * 32;      // 512
constexpr int kThreads = kConsumers + 32;            // + one producer warp
constexpr int kSlotBytes = 18432;                    // 4 Q4_K units (32 rows x 144 B) or 2 Q8_0 units (32 rows x 272 B)
constexpr int kSlots = 8;                            // a multiple of the consumer-group count (4 or 8): a slot is always drained by the
                                                     // same warps, so no waiter can be more than one mbarrier phase away from its barrier
constexpr int kMaxRowsCta = 224;                     // rows of one matrix owned by one CTA (text head: 32000 / 148 = 217)
constexpr int kMaxK = 16384;
constexpr int kMaxSplit = 8;                         // split-KV factor of the ring attention
constexpr int kMaxCta = 160;

// shared-memory map (bytes)
constexpr int kOffRing = 0;
constexpr int kOffBars = kSlots * kSlotBytes;                    // full[kSlots], empty[kSlots]
constexpr int kOffX8 = kOffBars + 256;                           // int8 activations [kMaxK]
constexpr int kOffBs = kOffX8 + kMaxK;                           // int16 sums per 32 [kMaxK / 32]
constexpr int kOffDx = kOffBs + kMaxK / 32 * 2;                  // f32 scales: per 256 (Q8_K) or per 32 (Q8_0)
constexpr int kOffPart = kOffDx + kMaxK / 32 * 4;                // f64 [kConsumerWarps][kMaxRowsCta]
constexpr int kOffRed = kOffPart + kConsumerWarps * kMaxRowsCta * 8;   // 32 x 8 B reduction scratch
constexpr int kOffRope = kOffRed + 256;                          // f32 [128]: cos | sin of this step's position
constexpr int kOffMisc = kOffRope + 512;                         // abort flag, token, ...
constexpr int kSmemBytes = kOffMisc + 64;
constexpr int kAttnScratch = kOffRed - kOffX8;                   // attention phases reuse the GEMV operand area
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct __align__(8) LL { float v; uint32_t seq; };

enum PhaseType : int { PH_EMBED = 0, PH_GEMV = 1, PH_ATTN = 2, PH_DEP_EMBED = 3, PH_FINALIZE_T = 4, PH_FINALIZE_D = 5 };

// One phase.  Vectors exchanged between CTAs are LL arrays; `*_src` is the index of the phase (of the same launch)
// that wrote the vector, which fixes the sequence number a reader waits for.
struct __align__(16) StepPhase {
    int32_t type = 0, pro = 0, epi = 0, gran = 1;
    int32_t fam = 0, pad_[3] = {};    // kernel family of the phase (engine.cu Family), for the timeline only
    // ---- PH_GEMV ----
    const uint8_t *w = nullptr;       // matrix in stream layout (see repack_stream_kernel)
    int32_t K = 0, rows = 0;          // stored rows (gate / up rows interleaved for EPI_GATE)
    const LL *x_ll = nullptr;         // input vector [K] written by phase x_src ...
    const float *x_plain = nullptr;   // ... or plain floats written by an earlier launch
    int32_t x_src = 0, resid_src = 0;
    const float *alpha = nullptr;     // PRO_RMS
    float eps = 0.f;
    int32_t step = 0;                 // PH_DEP_EMBED / PH_ATTN(depformer): codebook step k
    LL *out = nullptr;                // output vector (LL)
    const LL *resid = nullptr;        // EPI_RESID: out[r] = resid[r] + acc (ping-pong residual buffers)
    float *norm_out = nullptr;        // PRO_RMS: CTA 0 also stores rms_norm(x) * alpha as plain floats (transformer_out)
    float *out_plain = nullptr;       // EPI_ARGMAX: logits as plain floats (read by the host / the sampler)
    LL *keys = nullptr;               // EPI_ARGMAX: per-CTA arg-max keys [n_cta][2] = {hi, lo}
    // ---- PH_ATTN ----
    uint16_t *kc = nullptr, *vc = nullptr;   // this layer's ring [H][cap][DH] bf16
    int32_t heads = 0, dh = 0, cap = 0, split = 1;
    int32_t pos_const = -1, max_period = 0;
    LL *scores = nullptr;             // split exchange: the scores of every head [H][cap]
    // ---- PH_EMBED / PH_DEP_EMBED ----
    const EmbTable *tables = nullptr; // PH_EMBED: device array [n_tables]
    int32_t n_tables = 0, dim = 0;
    EmbTable emb;                     // PH_DEP_EMBED: table of the previous token
    const float *embed_in = nullptr;  // PH_EMBED: voice-embedding prompt row (ctrl->embed_override)
    const LL *prev_keys = nullptr;    // PH_DEP_EMBED (k > 0): keys written by phase prev_src; FINALIZE: first key block
    int32_t prev_src = 0, dep_q = 0, has_depformer = 0, keys_stride = 0;
    int32_t key_src[40] = {};         // PH_FINALIZE_*: phase that wrote keys block k
};

struct StepArgs {
    const StepPhase *phases = nullptr;
    int32_t n_phases = 0;
    int32_t rope_dh = 0;              // > 0: cos / sin table of ctrl->offset for head dim rope_dh
    const float *rope_freq = nullptr;
    Ctrl *ctrl = nullptr;
    uint32_t *epoch = nullptr;        // launch counter of the stream (device); CTA 0 advances it in its finalize phase
    long long *dbg = nullptr;         // optional timeline: [n_phases][n_cta][8] globaltimer ns = {start, end, after prologue, after main loop, input loaded, rms scale known, epilogue stored, -}
};

// ---- small PTX helpers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ long long gtime_ns() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t addr, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
// weights are read exactly once per frame: L2 evict-first, so the 4 GB stream does not push the activation vectors and
// the KV ring out of the 126 MB L2
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "l"(pol) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// every wait in the kernel is bounded: a missing peer turns into an error flag, never into a hung device
struct Watch {
    volatile int *abort_flag;
    Ctrl *ctrl;
    long long t0 = 0;
    uint32_t spins = 0;
    __device__ __forceinline__ bool expired() {
        if ((++spins & 1023u) == 0u) {
            if (*abort_flag) return true;
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 6000000000ll) { *abort_flag = 1; ctrl->error = 3; return true; }   // ~3 s
        }
        return false;
    }
    __device__ __forceinline__ void reset() { t0 = 0; spins = 0; }
};

__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity, Watch &wd) {
    wd.reset();
    while (!mbar_try(addr, parity)) { if (wd.expired()) return; }
}

// ---- LL vectors ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(LL *p, float v, uint32_t seq) {
    const unsigned long long bits = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(bits) : "memory");
}
__device__ __forceinline__ void ll_store_u32(LL *p, uint32_t v, uint32_t seq) { ll_store(p, __uint_as_float(v), seq); }
__device__ __forceinline__ unsigned long long ll_ld1(const LL *p) {
    unsigned long long r;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ ulonglong2 ll_ld2(const LL *p) {
    ulonglong2 r;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float ll_wait1(const LL *p, uint32_t seq, Watch &wd) {
    wd.reset();
    unsigned long long r = ll_ld1(p);
    while ((uint32_t)(r >> 32) != seq) { if (wd.expired()) break; r = ll_ld1(p); }
    return __uint_as_float((uint32_t)r);
}
__device__ __forceinline__ uint32_t ll_wait1_u32(const LL *p, uint32_t seq, Watch &wd) { return __float_as_uint(ll_wait1(p, seq, wd)); }
// 8 consecutive entries (64-byte aligned)
__device__ __forceinline__ void ll_wait8(const LL *p, uint32_t seq, float (&v)[8], Watch &wd) {
    wd.reset();
    for (;;) {
        const ulonglong2 a = ll_ld2(p), b = ll_ld2(p + 2), c = ll_ld2(p + 4), d = ll_ld2(p + 6);
        const bool ok = (uint32_t)(a.x >> 32) == seq && (uint32_t)(a.y >> 32) == seq && (uint32_t)(b.x >> 32) == seq && (uint32_t)(b.y >> 32) == seq &&
                        (uint32_t)(c.x >> 32) == seq && (uint32_t)(c.y >> 32) == seq && (uint32_t)(d.x >> 32) == seq && (uint32_t)(d.y >> 32) == seq;
        if (ok || wd.expired()) {
            v[0] = __uint_as_float((uint32_t)a.x); v[1] = __uint_as_float((uint32_t)a.y); v[2] = __uint_as_float((uint32_t)b.x); v[3] = __uint_as_float((uint32_t)b.y);
            v[4] = __uint_as_float((uint32_t)c.x); v[5] = __uint_as_float((uint32_t)c.y); v[6] = __uint_as_float((uint32_t)d.x); v[7] = __uint_as_float((uint32_t)d.y);
            return;
        }
    }
}
// one attempt at 8 consecutive entries (no waiting): all four loads are in flight together
__device__ __forceinline__ bool ll_try8(const LL *p, uint32_t seq, float (&v)[8]) {
    const ulonglong2 a = ll_ld2(p), b = ll_ld2(p + 2), c = ll_ld2(p + 4), d = ll_ld2(p + 6);
    v[0] = __uint_as_float((uint32_t)a.x); v[1] = __uint_as_float((uint32_t)a.y); v[2] = __uint_as_float((uint32_t)b.x); v[3] = __uint_as_float((uint32_t)b.y);
    v[4] = __uint_as_float((uint32_t)c.x); v[5] = __uint_as_float((uint32_t)c.y); v[6] = __uint_as_float((uint32_t)d.x); v[7] = __uint_as_float((uint32_t)d.y);
    return (uint32_t)(a.x >> 32) == seq && (uint32_t)(a.y >> 32) == seq && (uint32_t)(b.x >> 32) == seq && (uint32_t)(b.y >> 32) == seq &&
           (uint32_t)(c.x >> 32) == seq && (uint32_t)(c.y >> 32) == seq && (uint32_t)(d.x >> 32) == seq && (uint32_t)(d.y >> 32) == seq;
}
__device__ __forceinline__ void ll_store_f64(LL *p, double v, uint32_t seq) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    ll_store_u32(p, (uint32_t)b, seq); ll_store_u32(p + 1, (uint32_t)(b >> 32), seq);
}
__device__ __forceinline__ double ll_wait_f64(const LL *p, uint32_t seq, Watch &wd) {
    wd.reset();
    ulonglong2 r = ll_ld2(p);
    while ((uint32_t)(r.x >> 32) != seq || (uint32_t)(r.y >> 32) != seq) { if (wd.expired()) break; r = ll_ld2(p); }
    return __longlong_as_double((long long)(((unsigned long long)(uint32_t)r.y << 32) | (unsigned long long)(uint32_t)r.x));
}

// ---- work split ------------------------------------------------------------------------------------------------------
// Rows of a matrix are cut into n_cta contiguous ranges (multiples of `gran`); a CTA's range is cut into tiles of 32 rows
// (the last one shorter); a unit = one tile x one 256-weight super-block.  Stream layout of a matrix = CTA spans in row
// order; inside a span units in [tile][super-block] order; inside a unit of th rows: [N chunks][th rows] 16 B of quants,
// then [th rows] 16 B of block header — for Q4_K the 16 header bytes {d, dmin, scales[12]} and the 8 x 16 quant bytes of
// the GGUF block as they are (144 B per row and super-block = GGUF bytes exactly); for Q8_0 the 8 fp16 scales of the
// 8 blocks and 16 x 16 int8 (272 B).
__host__ __device__ inline int row_begin(int rows, int gran, int n_cta, int cta) { return (int)((long long)cta * (rows / gran) / n_cta) * gran; }
template <int WT> struct Fmt;
template <> struct Fmt<12> { static constexpr int kChunks = 8, kRowBytes = 144, kUnitsPerSlot = 4; };
template <> struct Fmt<8> { static constexpr int kChunks = 16, kRowBytes = 272, kUnitsPerSlot = 2; };

struct Span {            // this CTA's share of one matrix
    int r0, n_rows, nsb, n_tiles, n_units, n_chunks;
    size_t base;         // byte offset of the span in the matrix
};
template <int WT>
__host__ __device__ inline Span span_of(int K, int rows, int gran, int n_cta, int cta) {
    Span s;
    s.r0 = row_begin(rows, gran, n_cta, cta);
    s.n_rows = row_begin(rows, gran, n_cta, cta + 1) - s.r0;
    s.nsb = K >> 8;
    s.n_tiles = (s.n_rows + 31) >> 5;
    s.n_units = s.n_tiles * s.nsb;
    s.n_chunks = (s.n_units + Fmt<WT>::kUnitsPerSlot - 1) / Fmt<WT>::kUnitsPerSlot;
    s.base = (size_t)s.r0 * s.nsb * Fmt<WT>::kRowBytes;
    return s;
}
// byte offset of unit u inside the span (u == n_units: end of the span)
template <int WT>
__host__ __device__ inline uint32_t unit_off(const Span &s, int u) {
    if (u >= s.n_units) return (uint32_t)s.n_rows * s.nsb * Fmt<WT>::kRowBytes;
    const int tile = u / s.nsb, sb = u - tile * s.nsb;
    const int th = min(32, s.n_rows - tile * 32);
    return (uint32_t)(tile * 32 * s.nsb + sb * th) * Fmt<WT>::kRowBytes;
}

// ---- producer warp ---------------------------------------------------------------------------------------------------
// A slot holds kUnitsPerSlot units at fixed offsets (32 rows apart, whatever their height), so the consumers need no
// address arithmetic beyond (tile, super-block); the producer issues one bulk copy per unit and one expect_tx per slot.
//
// (An L2 look-ahead — a second cursor issuing cp.async.bulk.prefetch.L2 16 chunks ahead of the loads — was measured and removed:
// the main loops became 30-40 % slower, see profiles/r2_step_kernel.md.)
template <int WT>
struct ChunkCursor {
    static constexpr int UPS = Fmt<WT>::kUnitsPerSlot, RB = Fmt<WT>::kRowBytes;
    const StepArgs &a;
    int p = -1, c = 0, tile = 0, sb = 0, u = 0;
    Span s{};
    const uint8_t *src = nullptr;
    bool new_phase = false;
    __device__ __forceinline__ explicit ChunkCursor(const StepArgs &a_) : a(a_) { s.n_chunks = 0; }
    // moves to the next chunk; false at the end of the program.  After a true return: src / n_u / ub[] describe the chunk.
    int n_u = 0;
    uint32_t ub[UPS], bytes = 0;
    __device__ __forceinline__ bool next() {
        new_phase = false;
        while (c >= s.n_chunks) {
            if (++p >= a.n_phases) return false;
            const StepPhase *ph = a.phases + p;
            if (__ldg(&ph->type) != PH_GEMV) continue;
            s = span_of<WT>(__ldg(&ph->K), __ldg(&ph->rows), __ldg(&ph->gran), (int)gridDim.x, (int)blockIdx.x);
            src = reinterpret_cast<const uint8_t *>(__ldg(reinterpret_cast<const unsigned long long *>(&ph->w))) + s.base;
            c = 0; tile = 0; sb = 0; u = 0; new_phase = true; bytes = 0;
        }
        src += bytes;                                            // past the previous chunk
        n_u = min(UPS, s.n_units - u);
        bytes = 0;
#pragma unroll
        for (int k = 0; k < UPS; k++) {
            ub[k] = 0;
            if (k < n_u) {
                ub[k] = (uint32_t)min(32, s.n_rows - tile * 32) * RB;
                bytes += ub[k];
                if (++sb == s.nsb) { sb = 0; tile++; }
            }
        }
        u += n_u; c++;
        return true;
    }
};
template <int WT>
__device__ __forceinline__ void producer_loop(const StepArgs &a, uint8_t *smem, Watch &wd) {
    constexpr int UPS = Fmt<WT>::kUnitsPerSlot, UB = 32 * Fmt<WT>::kRowBytes;
    const uint32_t ring = smem_u32(smem + kOffRing), bars = smem_u32(smem + kOffBars);
    uint32_t slot = 0, lap = 0;                                 // position of the next chunk in the ring (whole launch)
    const unsigned long long pol = policy_evict_first();
    ChunkCursor<WT> ld(a);
    while (ld.next()) {
        mbar_wait(bars + (kSlots + slot) * 8, (lap & 1u) ^ 1u, wd);            // slot drained by its consumers
        if (*wd.abort_flag) return;
        mbar_expect_tx(bars + slot * 8, ld.bytes);
        const uint8_t *src = ld.src;
#pragma unroll
        for (int k = 0; k < UPS; k++) if (k < ld.n_u) {
            bulk_g2s(ring + slot * kSlotBytes + k * UB, src, ld.ub[k], bars + slot * 8, pol);
            src += ld.ub[k];
        }
        if (++slot == kSlots) { slot = 0; lap++; }
    }
}

// ---- activation prologue: (RMSNorm) + re-quantisation into shared memory ---------------------------------------------
// Same arithmetic as gemv.cuh (ggml rms_norm with the sum of squares in double; quantize_row_q8_K / quantize_row_q8_0).
// Warp w owns the 256-element blocks w, w + 16, ...; lane = 8 consecutive elements.
__device__ __forceinline__ void quant_q8k_block(int b, int lane, const float (&v)[8], int8_t *x8, int16_t *bs, float *dx) {
    float amax = 0.f, mx = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { const float ax = fabsf(v[i]); if (ax > amax) { amax = ax; mx = v[i]; } }
    const float wmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(amax)));
    const unsigned hit = __ballot_sync(0xffffffffu, amax == wmax);          // carrier = first element attaining the max
    const float carrier = __shfl_sync(0xffffffffu, mx, __ffs(hit) - 1);
    int q[8];
    float d = 0.f;
    if (wmax == 0.f) {
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = 0;
    } else {
        const float iscale = -127.f / carrier;
#pragma unroll
        for (int i = 0; i < 8; i++) { const int t = __float2int_rn(iscale * v[i]); q[i] = t < 127 ? t : 127; }
        d = 1.f / iscale;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += q[i];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if ((lane & 3) == 0) bs[b * 8 + (lane >> 2)] = (int16_t)s;
    if (lane == 0) dx[b] = d;
    uint2 pk;
    pk.x = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
    pk.y = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
    *reinterpret_cast<uint2 *>(x8 + b * 256 + lane * 8) = pk;
}
__device__ __forceinline__ void quant_q8_0_block(int b, int lane, const float (&v)[8], int8_t *x8, float *dx) {
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    const float d = amax / 127.f;
    const float id = d ? 1.0f / d : 0.0f;
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = (int)roundf(v[i] * id);
    if ((lane & 3) == 0) dx[b * 8 + (lane >> 2)] = __half2float(__float2half_rn(d));
    uint2 pk;
    pk.x = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
    pk.y = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
    *reinterpret_cast<uint2 *>(x8 + b * 256 + lane * 8) = pk;
}

template <int WT>
__device__ __forceinline__ void gemv_prologue(const StepPhase &ph, uint32_t epoch_bits, uint8_t *smem, Watch &wd, long long *stamp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = ph.K, nblk = K >> 8;
    int8_t *x8 = reinterpret_cast<int8_t *>(smem + kOffX8);
    int16_t *bs = reinterpret_cast<int16_t *>(smem + kOffBs);
    float *dx = reinterpret_cast<float *>(smem + kOffDx);
    double *red = reinterpret_cast<double *>(smem + kOffRed);
    const uint32_t seq = epoch_bits | (uint32_t)(ph.x_src + 1);
    constexpr int kKeep = 4;                                    // blocks of one warp kept in registers (K <= 16384)
    const int nb_w = warp < nblk ? (nblk - warp + kConsumerWarps - 1) / kConsumerWarps : 0;
    // Warp w takes the blocks (w + 16 i + rot) mod nblk.  rot differs from CTA to CTA, so that the 148 CTAs — which all read the
    // same vector at the same moment — start on different L2 lines instead of queueing on one slice after the other.
    const int rot = (int)(blockIdx.x % (unsigned)nblk);
    int blk[kKeep];
#pragma unroll
    for (int i = 0; i < kKeep; i++) { int b = warp + i * kConsumerWarps + rot; if (b >= nblk) b -= nblk; blk[i] = b; }
    float v[kKeep][8];
    // the warp's blocks are requested together (one L2 round trip); a block whose producer has not stored yet is re-polled
    if (ph.x_ll) {
        bool ok[kKeep];
#pragma unroll
        for (int i = 0; i < kKeep; i++) ok[i] = i < nb_w ? ll_try8(ph.x_ll + blk[i] * 256 + lane * 8, seq, v[i]) : true;
#pragma unroll
        for (int i = 0; i < kKeep; i++) if (!ok[i]) ll_wait8(ph.x_ll + blk[i] * 256 + lane * 8, seq, v[i], wd);
    } else {
#pragma unroll
        for (int i = 0; i < kKeep; i++) if (i < nb_w) {
            const int e0 = blk[i] * 256 + lane * 8;
            const float4 p0 = __ldcg(reinterpret_cast<const float4 *>(ph.x_plain + e0)), p1 = __ldcg(reinterpret_cast<const float4 *>(ph.x_plain + e0 + 4));
            v[i][0] = p0.x; v[i][1] = p0.y; v[i][2] = p0.z; v[i][3] = p0.w; v[i][4] = p1.x; v[i][5] = p1.y; v[i][6] = p1.z; v[i][7] = p1.w;
        }
    }
    if (stamp && threadIdx.x == 0) stamp[4] = gtime_ns();
    float scale = 1.f;
    if (ph.pro == PRO_RMS) {
        double ss = 0.0;
#pragma unroll
        for (int i = 0; i < kKeep; i++) if (i < nb_w) {
#pragma unroll
            for (int j = 0; j < 8; j++) ss += (double)(v[i][j] * v[i][j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        consumer_sync();
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; w++) tot += red[w];
        const float mean = (K & (K - 1)) == 0 ? (float)scalbn(tot, -(31 - __clz(K))) : (float)(tot / K);
        scale = 1.0f / sqrtf(mean + ph.eps);
    }
    if (stamp && threadIdx.x == 0) stamp[5] = gtime_ns();
#pragma unroll
    for (int i = 0; i < kKeep; i++) if (i < nb_w) {
        const int b = blk[i], e0 = b * 256 + lane * 8;
        if (ph.pro == PRO_RMS) {
            const float4 a0 = __ldg(reinterpret_cast<const float4 *>(ph.alpha + e0)), a1 = __ldg(reinterpret_cast<const float4 *>(ph.alpha + e0 + 4));
            const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int j = 0; j < 8; j++) v[i][j] = __fmul_rn(al[j], __fmul_rn(v[i][j], scale));      // alpha * (x * scale)
            if (ph.norm_out && blockIdx.x == 0) {
                *reinterpret_cast<float4 *>(ph.norm_out + e0) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
                *reinterpret_cast<float4 *>(ph.norm_out + e0 + 4) = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
            }
        }
        if (WT == 12) quant_q8k_block(b, lane, v[i], x8, bs, dx);
        else quant_q8_0_block(b, lane, v[i], x8, dx);
    }
}

// ---- one unit: th rows x one super-block, lane = row -------------------------------------------------------------------
// Q4_K: exact integer sums per super-block like ggml_vec_dot_q4_K_q8_K (isum = sum_s sc_s * dot_s, imin = sum_s m_s * bsum_s),
// then the two exact products d*dx*isum and dmin*dx*imin accumulated in double.
__device__ __forceinline__ double unit_q4k(const uint8_t *ub, int th, int lane, const int8_t *xs, const int16_t *bsb, float dxv) {
    const uint4 hd = *reinterpret_cast<const uint4 *>(ub + 8 * th * 16 + lane * 16);       // {d | dmin, scales[12]}
    const uint32_t s0 = hd.y, s1 = hd.z, s2 = hd.w;
    // get_scale_min_k4 for all eight sub-blocks at once (bytes of sc_lo / sc_hi = scales 0-3 / 4-7, same for the mins)
    const uint32_t sc_lo = s0 & 0x3f3f3f3fu, m_lo = s1 & 0x3f3f3f3fu;
    const uint32_t sc_hi = (s2 & 0x0f0f0f0fu) | ((s0 >> 2) & 0x30303030u);
    const uint32_t m_hi = ((s2 >> 4) & 0x0f0f0f0fu) | ((s1 >> 2) & 0x30303030u);
    const int4 b4 = *reinterpret_cast<const int4 *>(bsb);                                   // 8 x int16 sums per 32
    int imin = __dp2a_lo(b4.x, (int)m_lo, 0);
    imin = __dp2a_hi(b4.y, (int)m_lo, imin);
    imin = __dp2a_lo(b4.z, (int)m_hi, imin);
    imin = __dp2a_hi(b4.w, (int)m_hi, imin);
    int isum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int4 a0 = *reinterpret_cast<const int4 *>(ub + (2 * j) * th * 16 + lane * 16);
        const int4 a1 = *reinterpret_cast<const int4 *>(ub + (2 * j + 1) * th * 16 + lane * 16);
        const int4 xa0 = *reinterpret_cast<const int4 *>(xs + 64 * j);
        const int4 xa1 = *reinterpret_cast<const int4 *>(xs + 64 * j + 16);
        const int4 xb0 = *reinterpret_cast<const int4 *>(xs + 64 * j + 32);
        const int4 xb1 = *reinterpret_cast<const int4 *>(xs + 64 * j + 48);
        int dl = 0, dh = 0;   // dh accumulates 16 * (high nibble) products: an exact multiple of 16
        dl = __dp4a(a0.x & 0x0F0F0F0F, xa0.x, dl); dh = dp4a_us((unsigned)a0.x & 0xF0F0F0F0u, xb0.x, dh);
        dl = __dp4a(a0.y & 0x0F0F0F0F, xa0.y, dl); dh = dp4a_us((unsigned)a0.y & 0xF0F0F0F0u, xb0.y, dh);
        dl = __dp4a(a0.z & 0x0F0F0F0F, xa0.z, dl); dh = dp4a_us((unsigned)a0.z & 0xF0F0F0F0u, xb0.z, dh);
        dl = __dp4a(a0.w & 0x0F0F0F0F, xa0.w, dl); dh = dp4a_us((unsigned)a0.w & 0xF0F0F0F0u, xb0.w, dh);
        dl = __dp4a(a1.x & 0x0F0F0F0F, xa1.x, dl); dh = dp4a_us((unsigned)a1.x & 0xF0F0F0F0u, xb1.x, dh);
        dl = __dp4a(a1.y & 0x0F0F0F0F, xa1.y, dl); dh = dp4a_us((unsigned)a1.y & 0xF0F0F0F0u, xb1.y, dh);
        dl = __dp4a(a1.z & 0x0F0F0F0F, xa1.z, dl); dh = dp4a_us((unsigned)a1.z & 0xF0F0F0F0u, xb1.z, dh);
        dl = __dp4a(a1.w & 0x0F0F0F0F, xa1.w, dl); dh = dp4a_us((unsigned)a1.w & 0xF0F0F0F0u, xb1.w, dh);
        const uint32_t scw = j < 2 ? sc_lo : sc_hi;
        const int sh = (j & 1) * 16;
        isum += (int)((scw >> sh) & 0xffu) * dl + (int)((scw >> (sh + 8)) & 0xffu) * (dh >> 4);
    }
    const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&hd.x));
    double acc = (double)(dm.x * dxv) * (double)isum;
    acc = fma(-(double)(dm.y * dxv), (double)imin, acc);
    return acc;
}
// Q8_0: per 32-block sumi * (fp16(d_w) * fp16(d_x)) like ggml_vec_dot_q8_0_q8_0, block terms accumulated in double
__device__ __forceinline__ double unit_q8_0(const uint8_t *ub, int th, int lane, const int8_t *xs, const float *dxb) {
    const uint4 hd = *reinterpret_cast<const uint4 *>(ub + 16 * th * 16 + lane * 16);      // 8 fp16 block scales
    const uint32_t hw[4] = {hd.x, hd.y, hd.z, hd.w};
    double acc = 0.0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const int4 a0 = *reinterpret_cast<const int4 *>(ub + (2 * b) * th * 16 + lane * 16);
        const int4 a1 = *reinterpret_cast<const int4 *>(ub + (2 * b + 1) * th * 16 + lane * 16);
        const int4 x0 = *reinterpret_cast<const int4 *>(xs + 32 * b);
        const int4 x1 = *reinterpret_cast<const int4 *>(xs + 32 * b + 16);
        int sum = 0;
        sum = __dp4a(a0.x, x0.x, sum); sum = __dp4a(a0.y, x0.y, sum); sum = __dp4a(a0.z, x0.z, sum); sum = __dp4a(a0.w, x0.w, sum);
        sum = __dp4a(a1.x, x1.x, sum); sum = __dp4a(a1.y, x1.y, sum); sum = __dp4a(a1.z, x1.z, sum); sum = __dp4a(a1.w, x1.w, sum);
        const float dw = __half2float(__ushort_as_half((unsigned short)((b & 1) ? (hw[b >> 1] >> 16) : (hw[b >> 1] & 0xffffu))));
        acc = fma((double)(dw * dxb[b]), (double)sum, acc);
    }
    return acc;
}

// ---- GEMV phase (consumer side) ----------------------------------------------------------------------------------------
template <int WT>
__device__ __forceinline__ void gemv_phase(const StepPhase &ph, uint32_t epoch_bits, int p, uint32_t &q_base, uint8_t *smem, Watch &wd, long long *stamp) {
    constexpr int UPS = Fmt<WT>::kUnitsPerSlot, NG = kConsumerWarps / UPS;
    const int cta = blockIdx.x, n_cta = gridDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const Span s = span_of<WT>(ph.K, ph.rows, ph.gran, n_cta, cta);
    const uint32_t seq_out = epoch_bits | (uint32_t)(p + 1);
    const uint32_t bars = smem_u32(smem + kOffBars);
    double *part = reinterpret_cast<double *>(smem + kOffPart);
    // the old residual values of this CTA's rows were complete phases ago: request them now, use them in the epilogue
    unsigned long long resid_raw = 0ull;
    if (ph.epi == EPI_RESID && tid < s.n_rows) resid_raw = ll_ld1(ph.resid + s.r0 + tid);
    if (s.n_rows > 0) {
        gemv_prologue<WT>(ph, epoch_bits, smem, wd, stamp);
        consumer_sync();
        if (stamp && tid == 0) stamp[2] = gtime_ns();
        const int8_t *x8 = reinterpret_cast<const int8_t *>(smem + kOffX8);
        const int16_t *bs = reinterpret_cast<const int16_t *>(smem + kOffBs);
        const float *dx = reinterpret_cast<const float *>(smem + kOffDx);
        // chunk c of this phase is global chunk q_base + c; it is consumed by warp group (q % NG): warp w takes unit w % UPS of it
        const int grp = warp / UPS, sub = warp % UPS;
        int c = (int)((grp + NG - (q_base % NG)) % NG);
        uint32_t slot = (q_base + (uint32_t)c) % kSlots, lap = (q_base + (uint32_t)c) / kSlots;
        int u = c * UPS + sub;
        int tile = u / s.nsb, sb = u - tile * s.nsb;                         // advanced incrementally below (16 units per round)
        for (; c < s.n_chunks; c += NG) {
            mbar_wait(bars + slot * 8, lap & 1u, wd);
            if (u < s.n_units) {
                const int th = min(32, s.n_rows - tile * 32);
                const uint8_t *ub = smem + kOffRing + slot * kSlotBytes + sub * (32 * Fmt<WT>::kRowBytes);
                if (lane < th) {
                    double acc;
                    if (WT == 12) acc = unit_q4k(ub, th, lane, x8 + sb * 256, bs + sb * 8, dx[sb]);
                    else acc = unit_q8_0(ub, th, lane, x8 + sb * 256, dx + sb * 8);
                    double *pp = part + warp * kMaxRowsCta + tile * 32 + lane;
                    *pp += acc;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + (kSlots + slot) * 8);
            u += kConsumerWarps; sb += kConsumerWarps;
            while (sb >= s.nsb) { sb -= s.nsb; tile++; }
            slot += NG; if (slot >= kSlots) { slot -= kSlots; lap++; }
        }
        consumer_sync();
        if (stamp && tid == 0) stamp[3] = gtime_ns();
    }
    q_base += (uint32_t)s.n_chunks;

    // ---- cross-warp reduction (fixed order) + epilogue ----
    const int n_out = ph.epi == EPI_GATE ? s.n_rows >> 1 : s.n_rows;
    unsigned long long best = 0ull;
    if (tid < n_out) {
        auto total = [&](int r) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kConsumerWarps; w++) { t += part[w * kMaxRowsCta + r]; part[w * kMaxRowsCta + r] = 0.0; }
            return (float)t;
        };
        if (ph.epi == EPI_GATE) {
            const float g = total(2 * tid), u = total(2 * tid + 1);
            ll_store(ph.out + (s.r0 >> 1) + tid, (g / (1.0f + (float)exp((double)(-g)))) * u, seq_out);
        } else {
            const int row = s.r0 + tid;
            const float v = total(tid);
            if (ph.epi == EPI_RESID) {
                const uint32_t rseq = epoch_bits | (uint32_t)(ph.resid_src + 1);
                const float old = (uint32_t)(resid_raw >> 32) == rseq ? __uint_as_float((uint32_t)resid_raw) : ll_wait1(ph.resid + row, rseq, wd);
                ll_store(ph.out + row, old + v, seq_out);
            } else if (ph.epi == EPI_ARGMAX) {
                if (ph.out) ll_store(ph.out + row, v, seq_out);
                if (ph.out_plain) ph.out_plain[row] = v;
                best = argmax_key(v, row);
            } else ll_store(ph.out + row, v, seq_out);
        }
    }
    if (stamp && tid == 0) stamp[6] = gtime_ns();
    if (ph.epi == EPI_ARGMAX) {
        // CTA maximum -> this CTA's key entry (every CTA writes one, also those without rows: readers poll all of them)
        unsigned long long *sb = reinterpret_cast<unsigned long long *>(smem + kOffRed) + 16;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); best = t > best ? t : best; }
        if (lane == 0) sb[warp] = best;
        consumer_sync();
        if (tid == 0) {
            unsigned long long bb = 0ull;
            for (int w = 0; w < kConsumerWarps; w++) bb = sb[w] > bb ? sb[w] : bb;
            ll_store_u32(ph.keys + 2 * cta, (uint32_t)(bb >> 32), seq_out);
            ll_store_u32(ph.keys + 2 * cta + 1, (uint32_t)bb, seq_out);
        }
    }
}

// arg-max over the per-CTA keys of phase `src` (all consumer threads call; result in every thread)
__device__ __forceinline__ unsigned long long gather_key(const LL *keys, int src, uint32_t epoch_bits, uint8_t *smem, Watch &wd) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_cta = gridDim.x;
    const uint32_t seq = epoch_bits | (uint32_t)(src + 1);
    unsigned long long *sb = reinterpret_cast<unsigned long long *>(smem + kOffRed) + 16;
    unsigned long long k = 0ull;
    if (tid < n_cta) {
        wd.reset();
        ulonglong2 r = ll_ld2(keys + 2 * tid);
        while ((uint32_t)(r.x >> 32) != seq || (uint32_t)(r.y >> 32) != seq) { if (wd.expired()) break; r = ll_ld2(keys + 2 * tid); }
        k = ((unsigned long long)(uint32_t)r.x << 32) | (unsigned long long)(uint32_t)r.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, k, o); k = t > k ? t : k; }
    consumer_sync();                 // scratch may still be read by the previous user
    if (lane == 0) sb[warp] = k;
    consumer_sync();
    unsigned long long bb = 0ull;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; w++) bb = sb[w] > bb ? sb[w] : bb;
    return bb;
}

// ---- embedding sum (lm.h:555-584, lm_utils.h:157-182): tables added left to right like the graph --------------------------
__device__ __forceinline__ void embed_phase(const StepPhase &ph, const StepArgs &a, uint32_t epoch_bits, int p, uint8_t *smem) {
    const int cta = blockIdx.x, n_cta = gridDim.x, tid = threadIdx.x;
    const Ctrl *c = a.ctrl;
    const uint32_t seq_out = epoch_bits | (uint32_t)(p + 1);
    const int i0 = row_begin(ph.dim, 1, n_cta, cta), n = row_begin(ph.dim, 1, n_cta, cta + 1) - i0;
    float *tmp = reinterpret_cast<float *>(smem + kOffPart);          // [n_tables][n] (part is all zero between GEMV phases)
    if (ph.embed_in && c->embed_override) {
        if (tid < n) ll_store(ph.out + i0 + tid, __ldcg(ph.embed_in + i0 + tid), seq_out);
        return;
    }
    const int32_t *toks = c->feed_n ? c->feed + (size_t)(c->frame % c->feed_n) * c->n_in : c->tokens;
    for (int j = tid; j < ph.n_tables * n; j += kConsumers) {
        const int t = j / n, i = j - t * n;
        const int tok = toks[t];
        float e = emb_element(ph.tables[t], tok < 0 ? 0 : tok, i0 + i);
        e = e * (tok == -1 ? 0.f : 1.f);
        tmp[j] = e;
    }
    consumer_sync();
    if (tid < n) {
        float acc = tmp[tid];
        for (int t = 1; t < ph.n_tables; t++) acc = acc + tmp[t * n + tid];
        ll_store(ph.out + i0 + tid, acc, seq_out);
    }
    consumer_sync();
    for (int j = tid; j < ph.n_tables * n; j += kConsumers) tmp[j] = 0.f;      // leave `part` zeroed
    consumer_sync();
}

// depformer step input: depformer_in[k] . t_out (hoisted GEMV) + embedding of the previous token (lm.h:464-467, 494-516)
__device__ __forceinline__ void dep_embed_phase(const StepPhase &ph, const StepArgs &a, uint32_t epoch_bits, int p, uint8_t *smem, Watch &wd) {
    const int cta = blockIdx.x, n_cta = gridDim.x, tid = threadIdx.x;
    const Ctrl *c = a.ctrl;
    const int k = ph.step;
    int token;
    if (k == 0) { const int o = c->text_override; token = o != INT32_MIN ? o : c->out_tokens[0]; }
    else {
        const int f = c->force[k - 1];
        const unsigned long long key = gather_key(ph.prev_keys, ph.prev_src, epoch_bits, smem, wd);    // every thread takes part
        token = f != INT32_MIN ? f : argmax_key_index(key);
    }
    const uint32_t seq_out = epoch_bits | (uint32_t)(p + 1), seq_in = epoch_bits | (uint32_t)(ph.x_src + 1);
    const int i0 = row_begin(ph.dim, 1, n_cta, cta), n = row_begin(ph.dim, 1, n_cta, cta + 1) - i0;
    if (tid < n) {
        const int row = i0 + tid;
        float e;
        if (k == 0) { e = emb_element(ph.emb, token < 0 ? 0 : token, row); e = e * (token == -1 ? 0.f : 1.f); }
        else e = emb_element(ph.emb, token, row);
        const float d = ll_wait1(ph.x_ll + row, seq_in, wd);
        ll_store(ph.out + row, d + e, seq_out);
    }
}

// ---- RoPE + ring insert + single-query attention over the bf16 ring (attention.cuh arithmetic) ------------------------------
// CTA (h, c) = (cta / S, cta % S) of head h:  scores of slot range c  ->  ONE exchange (the scores of the head, as LL words)  ->
// every split has all scores, so max, exp, sum and the bf16-rounded probabilities are computed locally and identically  ->
// context dims [c DH/S, (c+1) DH/S) over ALL slots (V is cut by dims, so no partial contexts have to be merged).
// Short rings (<= kSoloCtx valid slots: the depformer, the first frames of a conversation) are handled by split 0 alone.
// The first batch of K and V rows is requested BEFORE q / k / v of this step are polled: the ring rows of earlier steps do
// not depend on the predecessor phase, so their round trip overlaps the wait.
constexpr int kSoloCtx = 64;
template <int DH>
__device__ __forceinline__ void attn_phase(const StepPhase &ph, const StepArgs &a, uint32_t epoch_bits, int p, uint8_t *smem, Watch &wd) {
    constexpr int LPS = DH / 8;                   // lanes per slot in the score pass (8 dims = 16 B of bf16 each)
    constexpr int NG = kConsumers / LPS;          // slots in flight per CTA iteration
    constexpr int U = 4;
    const int cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (cta >= ph.heads * ph.split) return;
    const int h = cta / ph.split, c = cta - h * ph.split;
    const int cap = ph.cap, dim = ph.heads * DH;
    const int pos = ph.pos_const >= 0 ? ph.pos_const : a.ctrl->offset;
    const int slot = pos % cap;
    const int n_valid = (pos >= cap - 1) ? cap : pos + 1;
    const bool solo = n_valid <= kSoloCtx;
    if (solo && c != 0) return;
    const int S = solo ? 1 : ph.split;
    const uint32_t seq_in = epoch_bits | (uint32_t)(ph.x_src + 1), seq_out = epoch_bits | (uint32_t)(p + 1);

    uint8_t *scr = smem + kOffX8;
    double *part = reinterpret_cast<double *>(scr);                          // [512 / (DS/8)][DS] = 32 KB
    float *q_s = reinterpret_cast<float *>(scr + 32768);                     // [DH] bf16-rounded q'
    uint16_t *knew = reinterpret_cast<uint16_t *>(q_s + DH);                 // [DH]
    uint16_t *vnew = knew + DH;                                              // [DH]
    float *sc_s = reinterpret_cast<float *>(vnew + DH);                      // [n_valid]
    double *dred = reinterpret_cast<double *>(smem + kOffRed);
    float *fred = reinterpret_cast<float *>(smem + kOffRed) + 32;
    const float *rope = reinterpret_cast<const float *>(smem + kOffRope);

    const int lo = (int)((long long)n_valid * c / S), hi = (int)((long long)n_valid * (c + 1) / S);
    const int DS = DH / S, LPV = DS / 8, NGV = kConsumers / LPV;              // context pass: lanes per slot, slots in flight
    const int g = tid / LPS, sl = tid % LPS;                                  // score pass
    const int gv = tid / LPV, slv = tid - gv * LPV;                           // context pass
    const uint16_t *kbase = ph.kc + (size_t)h * cap * DH + sl * 8, *vbase = ph.vc + (size_t)h * cap * DH + c * DS + slv * 8;
    uint4 kk[U], vv[U];
    auto load_k = [&](int i0) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = i0 + u * NG + g;
            kk[u] = make_uint4(0, 0, 0, 0);
            if (i < hi && i != slot) kk[u] = __ldcg(reinterpret_cast<const uint4 *>(kbase + (size_t)i * DH));
        }
    };
    auto load_v = [&](int i0) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = i0 + u * NGV;
            vv[u] = make_uint4(0, 0, 0, 0);
            if (i < n_valid && i != slot) vv[u] = __ldcg(reinterpret_cast<const uint4 *>(vbase + (size_t)i * DH));
        }
    };
    load_k(lo);
    load_v(gv);

    // ---- 1. q / k / v of this head (all six words requested together), RoPE (interleaved pairs -> [re half | im half]) ----
    if (tid < DH / 2) {
        const int j = tid;
        const LL *pq = ph.x_ll + h * DH + 2 * j, *pk = pq + dim, *pv = pk + dim;
        wd.reset();
        ulonglong2 rq = ll_ld2(pq), rk = ll_ld2(pk), rv = ll_ld2(pv);
        while ((uint32_t)(rq.x >> 32) != seq_in || (uint32_t)(rq.y >> 32) != seq_in || (uint32_t)(rk.x >> 32) != seq_in ||
               (uint32_t)(rk.y >> 32) != seq_in || (uint32_t)(rv.x >> 32) != seq_in || (uint32_t)(rv.y >> 32) != seq_in) {
            if (wd.expired()) break;
            rq = ll_ld2(pq); rk = ll_ld2(pk); rv = ll_ld2(pv);
        }
        const float qr = __uint_as_float((uint32_t)rq.x), qi = __uint_as_float((uint32_t)rq.y);
        const float kr = __uint_as_float((uint32_t)rk.x), ki = __uint_as_float((uint32_t)rk.y);
        const float vr = __uint_as_float((uint32_t)rv.x), vi = __uint_as_float((uint32_t)rv.y);
        if (ph.max_period) {
            const float cs = rope[j], sn = rope[DH / 2 + j];
            q_s[j] = bf16_round(__fsub_rn(__fmul_rn(qr, cs), __fmul_rn(qi, sn)));
            q_s[DH / 2 + j] = bf16_round(__fadd_rn(__fmul_rn(qr, sn), __fmul_rn(qi, cs)));
            knew[j] = f32_to_bf16_bits(__fsub_rn(__fmul_rn(kr, cs), __fmul_rn(ki, sn)));
            knew[DH / 2 + j] = f32_to_bf16_bits(__fadd_rn(__fmul_rn(kr, sn), __fmul_rn(ki, cs)));
        } else {
            q_s[2 * j] = bf16_round(qr); q_s[2 * j + 1] = bf16_round(qi);
            knew[2 * j] = f32_to_bf16_bits(kr); knew[2 * j + 1] = f32_to_bf16_bits(ki);
        }
        vnew[2 * j] = f32_to_bf16_bits(vr); vnew[2 * j + 1] = f32_to_bf16_bits(vi);
    }
    consumer_sync();
    if (c == 0) {                                                             // ring insert (moshi_kv_cache_insert_kv)
        const size_t o = ((size_t)h * cap + slot) * DH;
        if (tid < DH / 4) reinterpret_cast<uint2 *>(ph.kc + o)[tid] = reinterpret_cast<const uint2 *>(knew)[tid];
        else if (tid >= 64 && tid < 64 + DH / 4) reinterpret_cast<uint2 *>(ph.vc + o)[tid - 64] = reinterpret_cast<const uint2 *>(vnew)[tid - 64];
    }

    // ---- 2. scores over this split's share of the valid slots; published to the other splits of the head ----
    {
        const float scale = 1.f / sqrtf((float)DH);
        float qv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) qv[i] = q_s[sl * 8 + i];
        for (int i0 = lo; i0 < hi; i0 += NG * U) {
            if (i0 != lo) load_k(i0);
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * NG + g;
                if (i == slot && i < hi) kk[u] = reinterpret_cast<const uint4 *>(knew)[sl];
                // bf16 x bf16 products are exact in fp32; they are summed in double (order-independent)
                double d = 0.0;
                d += (double)(bf16_bits_to_f32(kk[u].x & 0xffff) * qv[0]); d += (double)(bf16_bits_to_f32(kk[u].x >> 16) * qv[1]);
                d += (double)(bf16_bits_to_f32(kk[u].y & 0xffff) * qv[2]); d += (double)(bf16_bits_to_f32(kk[u].y >> 16) * qv[3]);
                d += (double)(bf16_bits_to_f32(kk[u].z & 0xffff) * qv[4]); d += (double)(bf16_bits_to_f32(kk[u].z >> 16) * qv[5]);
                d += (double)(bf16_bits_to_f32(kk[u].w & 0xffff) * qv[6]); d += (double)(bf16_bits_to_f32(kk[u].w >> 16) * qv[7]);
#pragma unroll
                for (int o = LPS / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                const float sv = (float)d * scale + 0.0f;
                if (i < hi && sl == 0) {
                    sc_s[i] = sv;
                    if (S > 1) ll_store(ph.scores + (size_t)h * cap + i, sv, seq_out);
                }
            }
        }
    }
    if (S > 1) {       // the other splits' scores: up to 8 words per thread requested together, missing ones re-polled
        const LL *src = ph.scores + (size_t)h * cap;
        for (int i0 = tid; i0 < n_valid; i0 += 8 * kConsumers) {
            unsigned long long r[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { const int i = i0 + u * kConsumers; r[u] = (i < n_valid && (i < lo || i >= hi)) ? ll_ld1(src + i) : 0ull; }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * kConsumers;
                if (i < n_valid && (i < lo || i >= hi)) {
                    wd.reset();
                    while ((uint32_t)(r[u] >> 32) != seq_out) { if (wd.expired()) break; r[u] = ll_ld1(src + i); }
                    sc_s[i] = __uint_as_float((uint32_t)r[u]);
                }
            }
        }
    }
    consumer_sync();

    // ---- 3. max, exp and row sum over ALL slots (ggml soft_max: expf(x - max), sum in double, scale by 1 / sum) ----
    float lmax = -INFINITY;
    for (int i = tid; i < n_valid; i += kConsumers) lmax = fmaxf(lmax, sc_s[i]);
    lmax = warp_max(lmax);
    if (lane == 0) fred[warp] = lmax;
    consumer_sync();
    float gmax = fred[0];
#pragma unroll
    for (int w = 1; w < kConsumerWarps; w++) gmax = fmaxf(gmax, fred[w]);
    double lsum = 0.0;
    for (int i = tid; i < n_valid; i += kConsumers) { const float e = (float)exp((double)(sc_s[i] - gmax)); sc_s[i] = e; lsum += (double)e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) dred[warp] = lsum;
    consumer_sync();
    double gsum = 0.0;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; w++) gsum += dred[w];
    const float inv = (float)(1.0 / gsum);

    // ---- 4. context dims [c DS, (c+1) DS) = sum_i bf16(p_i) * V_i over all slots ----
    {
        double acc[8];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = 0.0;
        for (int i0 = gv; i0 < n_valid; i0 += NGV * U) {
            if (i0 != gv) load_v(i0);
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * NGV;
                if (i < n_valid) {
                    if (i == slot) vv[u] = reinterpret_cast<const uint4 *>(vnew)[c * LPV + slv];
                    const float pr = bf16_round(sc_s[i] * inv);
                    acc[0] += (double)(bf16_bits_to_f32(vv[u].x & 0xffff) * pr); acc[1] += (double)(bf16_bits_to_f32(vv[u].x >> 16) * pr);
                    acc[2] += (double)(bf16_bits_to_f32(vv[u].y & 0xffff) * pr); acc[3] += (double)(bf16_bits_to_f32(vv[u].y >> 16) * pr);
                    acc[4] += (double)(bf16_bits_to_f32(vv[u].z & 0xffff) * pr); acc[5] += (double)(bf16_bits_to_f32(vv[u].z >> 16) * pr);
                    acc[6] += (double)(bf16_bits_to_f32(vv[u].w & 0xffff) * pr); acc[7] += (double)(bf16_bits_to_f32(vv[u].w >> 16) * pr);
                }
            }
        }
        // only the groups that saw a slot write their partials (n_valid is small in the depformer and early in a conversation)
        if (gv < n_valid) {
#pragma unroll
            for (int i = 0; i < 8; i++) part[gv * DS + slv * 8 + i] = acc[i];
        }
    }
    consumer_sync();
    {   // reduce over the groups: 4 threads per dim take a quarter of the groups each, then a fixed-order sum of the four
        const int ng = min(NGV, n_valid);
        double *qsum = reinterpret_cast<double *>(scr + 32768);               // [4][DS]: q / k / v rows and the probabilities are dead now
        if (tid < 4 * DS) {
            const int qd = tid / DS, d = tid - qd * DS;
            const int g0 = (ng * qd) / 4, g1 = (ng * (qd + 1)) / 4;
            double t = 0.0;
            for (int gg = g0; gg < g1; gg++) t += part[gg * DS + d];
            qsum[qd * DS + d] = t;
        }
        consumer_sync();
        if (tid < DS) {
            const double t = ((qsum[tid] + qsum[DS + tid]) + qsum[2 * DS + tid]) + qsum[3 * DS + tid];
            ll_store(ph.out + h * DH + c * DS + tid, (float)t, seq_out);
        }
    }
    consumer_sync();
    // the scratch aliases the GEMV accumulation area, which must read zero again (only what this phase can have touched)
    {
        const int used = 32768 + max(DH * 8 + n_valid * 4 + 16, 4 * DS * 8);  // bytes of scratch written, from kOffX8
        const int z0 = kOffPart - kOffX8, z1 = min(used, kOffRed - kOffX8);
        for (int j = z0 / 16 + tid; j < (z1 + 15) / 16; j += kConsumers) reinterpret_cast<uint4 *>(scr)[j] = make_uint4(0, 0, 0, 0);
    }
    consumer_sync();
}

// ---- the kernel ----------------------------------------------------------------------------------------------------------
template <int WT>
__global__ void __launch_bounds__(kThreads, 1) step_kernel(const StepArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(16) StepPhase s_ph[2];
    volatile int *abort_flag = reinterpret_cast<volatile int *>(smem + kOffMisc);
    const int tid = threadIdx.x, cta = blockIdx.x;
    const uint32_t bars = smem_u32(smem + kOffBars);
    if (tid == 0) {
        for (int s = 0; s < kSlots; s++) { mbar_init(bars + s * 8, 1); mbar_init(bars + (kSlots + s) * 8, Fmt<WT>::kUnitsPerSlot); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *abort_flag = 0;
    }
    {   // zero the accumulation area; cos / sin of this step's position (ggml_timestep_embedding on the f32 position,
        // through double: correctly rounded irrespective of the libm — rope.h:8-20)
        for (int j = tid; j < (kOffRed - kOffPart) / 16; j += kThreads) reinterpret_cast<uint4 *>(smem + kOffPart)[j] = make_uint4(0, 0, 0, 0);
        float *rope = reinterpret_cast<float *>(smem + kOffRope);
        if (a.rope_dh && tid < a.rope_dh / 2) {
            const float arg = (float)a.ctrl->offset * a.rope_freq[tid];
            rope[tid] = (float)cos((double)arg);
            rope[a.rope_dh / 2 + tid] = (float)sin((double)arg);
        }
        const int n4 = (int)(sizeof(StepPhase) / 4);
        for (int i = tid; i < n4; i += kThreads) reinterpret_cast<uint32_t *>(&s_ph[0])[i] = __ldg(reinterpret_cast<const uint32_t *>(a.phases) + i);
    }
    const uint32_t epoch_bits = (*reinterpret_cast<volatile uint32_t *>(a.epoch) & 0xfffffu) << 12;
    __syncthreads();
    Watch wd{abort_flag, a.ctrl};

    if (tid >= kConsumers) {                                   // ---- producer warp ----
        if (tid == kConsumers) producer_loop<WT>(a, smem, wd);
        return;
    }
    // ---- consumer warps ----
    uint32_t q_base = 0;
    constexpr int n4 = (int)(sizeof(StepPhase) / 4);
    for (int p = 0; p < a.n_phases; p++) {
        const StepPhase &ph = s_ph[p & 1];
        // descriptor of the next phase travels while this one runs
        uint32_t nw[(n4 + kConsumers - 1) / kConsumers];
        const bool more = p + 1 < a.n_phases;
#pragma unroll
        for (int i = 0; i < (n4 + kConsumers - 1) / kConsumers; i++) {
            const int j = tid + i * kConsumers;
            nw[i] = (more && j < n4) ? __ldg(reinterpret_cast<const uint32_t *>(a.phases + p + 1) + j) : 0u;
        }
        if (a.dbg && tid == 0) a.dbg[((size_t)p * gridDim.x + cta) * 8] = gtime_ns();
        switch (ph.type) {
            case PH_GEMV: gemv_phase<WT>(ph, epoch_bits, p, q_base, smem, wd, a.dbg ? a.dbg + ((size_t)p * gridDim.x + cta) * 8 : nullptr); break;
            case PH_ATTN: if (ph.dh == 128) attn_phase<128>(ph, a, epoch_bits, p, smem, wd); else attn_phase<64>(ph, a, epoch_bits, p, smem, wd); break;
            case PH_EMBED: embed_phase(ph, a, epoch_bits, p, smem); break;
            case PH_DEP_EMBED: dep_embed_phase(ph, a, epoch_bits, p, smem, wd); break;
            case PH_FINALIZE_T: {
                // greedy text token out of the per-CTA keys; position advances (states->offset += T, transformer.h:1269-1270)
                if (cta == 0) {
                    const unsigned long long key = gather_key(ph.prev_keys, ph.prev_src, epoch_bits, smem, wd);
                    if (tid == 0) {
                        Ctrl *c = a.ctrl;
                        c->out_tokens[0] = argmax_key_index(key);
                        c->offset += 1;
                        if (!ph.has_depformer && c->feed_n) { if (c->trace) c->trace[c->frame] = c->out_tokens[0]; c->frame += 1; }
                        *a.epoch += 1u;
                    }
                }
                break;
            }
            case PH_FINALIZE_D: {
                if (cta == 0) {
                    Ctrl *c = a.ctrl;
                    for (int k = 0; k < ph.dep_q; k++) {
                        const unsigned long long key = gather_key(ph.prev_keys + (size_t)k * ph.keys_stride, ph.key_src[k], epoch_bits, smem, wd);
                        if (tid == 0) c->out_tokens[1 + k] = argmax_key_index(key);
                    }
                    consumer_sync();
                    if (c->feed_n && c->trace && tid <= ph.dep_q) c->trace[(size_t)c->frame * (ph.dep_q + 1) + tid] = c->out_tokens[tid];
                    consumer_sync();
                    if (tid == 0) { if (c->feed_n) c->frame += 1; *a.epoch += 1u; }
                }
                break;
            }
        }
        if (a.dbg && tid == 0) a.dbg[((size_t)p * gridDim.x + cta) * 8 + 1] = gtime_ns();
        if (more) {
#pragma unroll
            for (int i = 0; i < (n4 + kConsumers - 1) / kConsumers; i++) {
                const int j = tid + i * kConsumers;
                if (j < n4) reinterpret_cast<uint32_t *>(&s_ph[(p + 1) & 1])[j] = nw[i];
            }
        }
        consumer_sync();
        if (*abort_flag) break;
    }
}

// ---- load-time repack: GGUF row-major blocks -> stream layout (see "work split" above) ----------------------------------------
// One thread per (stored row, super-block).  perm_half > 0 interleaves rows for the gated MLP:
// stored row v <- source row (v & 1 ? perm_half + v / 2 : v / 2).
__global__ void repack_stream_kernel(const uint8_t *src, uint8_t *dst, int type, int rows, int K, int gran, int n_cta, int perm_half) {
    const int nsb = K >> 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * nsb) return;
    const int v = (int)(idx / nsb), sb = (int)(idx % nsb);
    const int srow = perm_half > 0 ? ((v & 1) ? perm_half + (v >> 1) : (v >> 1)) : v;
    // owner CTA of stored row v: the largest c with row_begin(c) <= v
    int c = (int)(((long long)(v / gran) * n_cta) / (rows / gran));
    while (c + 1 < n_cta && row_begin(rows, gran, n_cta, c + 1) <= v) c++;
    while (c > 0 && row_begin(rows, gran, n_cta, c) > v) c--;
    const int r0 = row_begin(rows, gran, n_cta, c), n_rows = row_begin(rows, gran, n_cta, c + 1) - r0;
    const int lr = v - r0, tile = lr >> 5, r = lr & 31, th = min(32, n_rows - tile * 32);
    if (type == 12) {
        const uint8_t *blk = src + ((size_t)srow * nsb + sb) * 144;
        uint8_t *ub = dst + ((size_t)r0 * nsb + (size_t)tile * 32 * nsb + (size_t)sb * th) * 144;
        *reinterpret_cast<uint4 *>(ub + 8 * th * 16 + r * 16) = *reinterpret_cast<const uint4 *>(blk);          // {d, dmin, scales[12]}
#pragma unroll
        for (int j = 0; j < 8; j++) *reinterpret_cast<uint4 *>(ub + j * th * 16 + r * 16) = *reinterpret_cast<const uint4 *>(blk + 16 + j * 16);
    } else {
        const uint8_t *blk = src + ((size_t)srow * nsb + sb) * 8 * 34;                                            // 8 blocks of 34 B
        uint8_t *ub = dst + ((size_t)r0 * nsb + (size_t)tile * 32 * nsb + (size_t)sb * th) * 272;
        uint16_t *hd = reinterpret_cast<uint16_t *>(ub + 16 * th * 16 + r * 16);
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const uint16_t *s16 = reinterpret_cast<const uint16_t *>(blk + b * 34);                               // 2-byte aligned
            hd[b] = s16[0];
            uint16_t *o0 = reinterpret_cast<uint16_t *>(ub + (2 * b) * th * 16 + r * 16), *o1 = reinterpret_cast<uint16_t *>(ub + (2 * b + 1) * th * 16 + r * 16);
#pragma unroll
            for (int i = 0; i < 8; i++) { o0[i] = s16[1 + i]; o1[i] = s16[9 + i]; }
        }
    }
}

// LL vector -> plain floats (host reads of intermediate vectors; test hook)
__global__ void ll_unpack_kernel(const LL *in, float *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i].v;
}

}  // namespace sk
}  // namespace msx
