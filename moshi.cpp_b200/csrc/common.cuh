// common.cuh — shared device helpers and argument structs for the msx kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

namespace msx {

constexpr int kThreads = 256;   // every msx kernel runs 8 warps per CTA
constexpr int kWarps = kThreads / 32;

// ---- control block: per-stream device-resident step parameters --------------------------------
// CUDA graphs are captured once per stream; everything that changes per frame (position, tokens,
// overrides) lives here and is read by the kernels, never passed as a kernel argument.
struct Ctrl {
    int32_t offset;                 // temporal position of the step being computed
    int32_t frame;                  // frames run in resident mode (indexes `feed`)
    int32_t feed_n;                 // 0 = host mode (tokens[] written by the host)
    int32_t n_in;                   // n_q + 1
    const int32_t *feed;            // resident mode: [feed_n][n_q+1]
    int32_t *trace;                 // resident mode: [n_steps][1+dep_q] outputs or null
    // ---- host-written input block: one H2D copy per call (kCtrlInOffset / kCtrlInBytes) ----
    int32_t text_override;          // INT32_MIN = use greedy text token
    int32_t tokens[40];             // inputs of this frame
    int32_t force[40];              // depformer feed-forward overrides (INT32_MIN = greedy)
    int32_t embed_override;         // != 0: the embedding sum is replaced by the stream's embed_in row (voice-embedding prompt)
    int32_t pad0_[2];
    // ---- device-written output block: one D2H copy per call ----
    int32_t out_tokens[41];         // {text, audio[dep_q]}
    int32_t pad1_[3];
    unsigned long long text_key;    // arg-max keys (see argmax_key())
    unsigned long long audio_key[40];
    int32_t error;                  // set by a device-side watchdog (step kernel waits, tensor-parallel inbox polls)
    int32_t pad2_;
};
constexpr size_t kCtrlInOffset = 32;
constexpr size_t kCtrlInBytes = 84 * 4;
constexpr size_t kCtrlOutOffset = kCtrlInOffset + kCtrlInBytes;
constexpr size_t kCtrlOutBytes = 44 * 4;
static_assert(offsetof(Ctrl, text_override) == kCtrlInOffset, "Ctrl layout");
static_assert(offsetof(Ctrl, out_tokens) == kCtrlOutOffset, "Ctrl layout");
static_assert(offsetof(Ctrl, text_key) == kCtrlOutOffset + kCtrlOutBytes, "Ctrl layout");

// ---- tensor descriptors -------------------------------------------------------------------------
// Repacked quantised linear [rows][K] (see DESIGN.md "Data layout in HBM").
//   Q4_K: qs  rows*(K/2) bytes, per row groups of `gs` pairs: [gs first-chunks][gs second-chunks]
//         sc  rows*(K/64) u32   {sc_lo, sc_hi, m_lo, m_hi} per 64-weight pair
//         dd  rows*(K/256) u32  {fp16 d, fp16 dmin} per super-block
//   Q8_0: qs  rows*K int8, per row groups of `gs` blocks: [gs first-halves][gs second-halves]
//         dd  rows*(K/32) fp16 d
struct QLinear {
    const uint8_t *qs = nullptr;
    const uint32_t *sc = nullptr;
    const void *dd = nullptr;
    int32_t type = 0;     // GgmlType
    int32_t K = 0;
    int32_t rows = 0;     // stored ("virtual") rows
    int32_t gs = 32;      // lanes per row group: 32 or 16
    int32_t gate = 0;     // rows interleaved (gate j, up j) for the gated MLP
};

// Embedding table kept in its GGUF row format (single rows are gathered, nothing to coalesce).
struct EmbTable {
    const uint8_t *data = nullptr;
    int32_t type = 0;
    int32_t K = 0;
    int32_t rows = 0;
    int32_t row_bytes = 0;
};

enum Prologue : int { PRO_PLAIN = 0, PRO_RMS = 1 };
enum Epilogue : int {
    EPI_STORE = 0,     // out[r] = acc
    EPI_RESID = 1,     // out[r] += acc                       (residual stream, in place)
    EPI_GATE = 2,      // out[r/2] = silu(acc[r]) * acc[r+1]  (rows interleaved at repack)
    EPI_ARGMAX = 3,    // out[r] = acc and atomicMax(key)
    EPI_ADD_EMB = 4,   // out[r] = acc + emb_table[token][r]  (depformer_in + last-token embedding)
    EPI_STORE_F64 = 6, // out_f64[r] = the un-rounded double accumulator (tensor-parallel partial sums, all-reduced in double)
    EPI_ADD_VEC = 5,   // out[r] = acc + addvec[r]            (depformer_in + low-rank / demux embedding computed by small_linear_kernel)
};

// Tensor-parallel peer-memory context (device resident).  Every rank owns an inbox [2 parities][world][dim] of 16-byte
// entries {value.lo, seq, value.hi, seq} that its peers map through CUDA IPC.  A GEMV with EPI_STORE_F64 pushes every
// un-rounded fp64 partial sum straight into every rank's inbox from its epilogue (one 16-byte store per row and rank over
// NVLink); the sequence number travels WITH the data in each 8-byte half (the "LL" idea: no fence, no separate flag, no
// completion counting), so the consumer (tp_apply_p2p_kernel) just polls the entries it needs and adds them in rank
// order.  Two parities: a rank can be at most one reduce ahead of its slowest peer.
struct TpCtx {
    int32_t rank = 0, world = 1, dim = 0, pad = 0;
    uint4 *inbox[8] = {};         // inbox base of every rank (own entry = local pointer)
    uint32_t *frame_ctr = nullptr;// local: temporal graphs completed (incremented by finalize_temporal_kernel); the sequence
                                  // number of reduce i of a frame is frame_ctr * reduces_per_frame + i + 1: stable for the whole
                                  // graph, so producers and consumers may read it at any time (no hand-over through a counter)
    int32_t reduces_per_frame = 0, pad2 = 0;
    int32_t *error = nullptr;     // Ctrl::error (spin watchdog)
};
__device__ __forceinline__ uint32_t tp_seq(const TpCtx *tp, int idx) {
    return *reinterpret_cast<const volatile uint32_t *>(tp->frame_ctr) * (uint32_t)tp->reduces_per_frame + (uint32_t)idx + 1u;
}

struct MatvecArgs {
    QLinear w;
    const float *x = nullptr;       // [K] activations (f32)
    const double *xparts = nullptr; // alternative input: x = sum of nparts double vectors (split-KV attention partials)
    int32_t nparts = 0, part_stride = 0;
    const float *alpha = nullptr;   // PRO_RMS: [K]
    float eps = 0.f;
    float *norm_out = nullptr;      // PRO_RMS: one CTA also stores rms_norm(x)*alpha here (transformer_out)
    float *out = nullptr;
    unsigned long long *key = nullptr;  // EPI_ARGMAX
    // EPI_ADD_EMB: token = force/override >= 0 ? that : decoded key / out_tokens
    EmbTable emb;
    Ctrl *ctrl = nullptr;
    int32_t emb_step = 0;           // 0: text token (scaled embedding), k>0: audio token of step k-1 (chained)
    const float *addvec = nullptr;  // EPI_ADD_VEC
    double *out_f64 = nullptr;      // EPI_STORE_F64
    const TpCtx *tp = nullptr;      // EPI_STORE_F64: push to the peers' inboxes instead (see TpCtx)
    int32_t tp_idx = 0;             // which reduce of the frame this launch feeds
};

// warp index broadcast from lane 0: tells the compiler the value is warp-uniform, so loops and branches on it
// are known to be convergent (no WARPSYNC.COLLECTIVE wrappers around the shuffles inside them)
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor is still running; it must not touch the predecessor's outputs
// (or overwrite its inputs) before griddep_wait().  griddep_launch() lets the NEXT kernel start early.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ---- TMA bulk copies + mbarrier completion (cp.async.bulk -> SASS UBLKCP; mbarrier -> SYNCS) ----------------------------
__device__ __forceinline__ void mbar_init(uint32_t addr, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}

// ---- small device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// fp32 -> bf16 bits, round to nearest even (ggml_compute_fp32_to_bf16)
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float f) {
    uint32_t u = __float_as_uint(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 64);
    return (uint16_t)((u + (0x7fffu + ((u >> 16) & 1u))) >> 16);
}
__device__ __forceinline__ float bf16_bits_to_f32(uint32_t h) { return __uint_as_float(h << 16); }
__device__ __forceinline__ float bf16_round(float f) { return bf16_bits_to_f32(f32_to_bf16_bits(f)); }

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t *p) { uint32_t v; asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// 128-bit streaming load (weights are read exactly once per step: do not allocate in L1)
__device__ __forceinline__ int4 ldg_stream(const void *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// arg-max key: larger value wins, ties -> smaller index (ggml argmax = first maximum)
__device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}
__device__ __forceinline__ int argmax_key_index(unsigned long long key) {
    return (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
}

// one element of a GGUF-format row (Q4_0 / Q8_0 / F32 / F16 / BF16): ggml dequantize_row_* per element
// (the host validates token ids — check_tokens in engine.cu; the clamp keeps a stray id inside the table regardless)
__device__ __forceinline__ float emb_element(const EmbTable &t, int row, int i) {
    row = max(0, min(row, t.rows - 1));
    const uint8_t *r = t.data + (size_t)row * t.row_bytes;
    switch (t.type) {
        case 0: return reinterpret_cast<const float *>(r)[i];
        case 1: return __half2float(reinterpret_cast<const __half *>(r)[i]);
        case 30: return bf16_bits_to_f32(reinterpret_cast<const uint16_t *>(r)[i]);
        case 8: {  // Q8_0: 34-byte blocks {fp16 d, int8 qs[32]}
            const uint8_t *b = r + (size_t)(i >> 5) * 34;
            float d = __half2float(*reinterpret_cast<const __half *>(b));
            return (float)((const int8_t *)(b + 2))[i & 31] * d;
        }
        case 2: {  // Q4_0: 18-byte blocks {fp16 d, qs[16]}: low nibbles = elems 0..15, high = 16..31
            const uint8_t *b = r + (size_t)(i >> 5) * 18;
            float d = __half2float(*reinterpret_cast<const __half *>(b));
            int j = i & 31;
            int q = (j < 16) ? (b[2 + j] & 0x0F) : (b[2 + j - 16] >> 4);
            return (float)(q - 8) * d;
        }
    }
    return 0.f;
}

}  // namespace msx
