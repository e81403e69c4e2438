// engine.cu — the translation unit behind the C ABI of include/moshi_b200.h.  The kernels are templates in the .cuh headers;
// the host side is split by concern into .inl files that share this unit's internal types (see DESIGN.md for the mapping to
// the reference):
//   model_loader.inl     GGUF -> device arenas, repack / on-load quantisation, tensor-parallel shards, msx_model_*
//   launcher.inl         kernel families, PDL launches, graph-capture bookkeeping
//   stream.inl           per-conversation state, launch chains of the two stacks, CUDA graphs, msx_stream_* / msx_step_*
//     step_program.inl   phase programs of the persistent step kernel
//   tensor_parallel.inl  peer-memory all-reduce plumbing (CUDA IPC)
//   conditioning.inl     TTS conditioning memory, voice conditioners, VAD
//   resident.inl         device-resident frame loops, profiling, timeline
//   generator.inl        LMGen host logic (msx_gen_*)
//   test_hooks.inl       unit-level entry points for tests / bench
//   converters.inl       GGUF / safetensors -> quantised GGUF
//   batch.inl            lock-step batches of streams and the batched-T prompt prefill (msx_batch_*, msx_stream_prefill)
// This file keeps the error channel, the run-time NCCL binding and the includes.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/moshi_b200.h"
#include "attention.cuh"
#include "common.cuh"
#include "gemv.cuh"
#include "gguf_file.h"
#include "local_attn.cuh"
#include "misc_kernels.cuh"
#include "mma_gemm.cuh"
#include "safetensors_file.h"
#include "sample.cuh"
#include "step_kernel.cuh"
#include "tts_kernels.cuh"
#include "tc_gemm.cuh"
#include "mimi_rvq.cuh"

using namespace msx;

// -------------------------------------------------------------------------------------------------
// errors
// -------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(MSX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));               \
    } while (0)

extern "C" const char *msx_last_error(void) { return g_err.c_str(); }

// ---- NCCL, bound at run time (tensor-parallel streams only) -------------------------------------------
// dlopen by soname: inside a torch process this resolves to the NCCL torch has already loaded, so one
// process never holds two NCCL copies; a process that never asks for tensor parallelism never loads it.
namespace {
struct Nccl {
    typedef struct { char internal[128]; } UniqueId;
    typedef void *Comm;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};
Nccl &nccl() {
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { n.why = std::string("cannot load libnccl: ") + dlerror(); return n; }
    n.GetUniqueId = (int (*)(Nccl::UniqueId *))dlsym(h, "ncclGetUniqueId");
    n.CommInitRank = (int (*)(Nccl::Comm *, int, Nccl::UniqueId, int))dlsym(h, "ncclCommInitRank");
    n.CommDestroy = (int (*)(Nccl::Comm))dlsym(h, "ncclCommDestroy");
    n.AllReduce = (int (*)(const void *, void *, size_t, int, int, Nccl::Comm, cudaStream_t))dlsym(h, "ncclAllReduce");
    n.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.GetErrorString;
    if (!n.ok) n.why = "libnccl lacks an expected symbol";
    return n;
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0;       // ncclDataType_t / ncclRedOp_t values (nccl.h)
}  // namespace
extern "C" const char *msx_version(void) { return "moshi_b200 0.1 (sm_100a)"; }
extern "C" int msx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

#include "model_loader.inl"
#include "launcher.inl"
#include "stream.inl"
#include "tensor_parallel.inl"
#include "conditioning.inl"
#include "resident.inl"
#include "generator.inl"
#include "test_hooks.inl"
#include "converters.inl"
#include "batch.inl"
#include "mimi_rvq.inl"

static void free_prefill(struct msx_batch *b) { delete b; }
