// tcgen05_probe.cu — dev probe for the open item "tcgen05 kind::i8 GEMM for the dense contractions" (DESIGN.md section 6):
// one 128 x 64 x 256 signed-int8 tile product on the 5th-generation tensor cores — operands in shared memory in the canonical
// K-major no-swizzle layout (8-row x 16-byte core matrices), accumulator in tensor memory, read back with tcgen05.ld — checked
// against a CPU loop.  It pins the descriptor encodings (shared-memory matrix descriptor, kind::i8 instruction descriptor) that a
// Q4_K prefill GEMM would build on: with the 6-bit sub-block scales split as sc = sc_lo + 8 sc_hi, the operands q * sc_lo and
// q * sc_hi (<= 15 * 7 = 105) fit s8 and two MMAs per super-block give the exact integer sum_k sc * q * x.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -o /tmp/tcgen05_probe scripts/tcgen05_probe.cu && /tmp/tcgen05_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 256;
constexpr int kSBO = (K / 16) * 128;      // bytes between 8-row groups: a row group holds K / 16 core matrices of 128 B
constexpr int kLBO = 128;                 // bytes between the two core matrices of one K = 32 step

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::i8: c = S32 (2) << 4 | a = INT8 (1) << 7 | b = INT8 (1) << 10 | K-major both | N >> 3 << 17 | M >> 4 << 24
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__global__ void __launch_bounds__(128, 1) probe_kernel(const int8_t *A, const int8_t *B, int32_t *D, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem, *sB = smem + M * K;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + M * K + N * K);
    uint32_t *slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // row-major [rows][K] -> canonical layout: (row / 8) * SBO + (k / 16) * 128 + (row % 8) * 16 + k % 16
    for (int i = tid; i < M * K / 16; i += 128) { const int r = i / (K / 16), c = i % (K / 16); *reinterpret_cast<uint4 *>(sA + (r / 8) * kSBO + c * 128 + (r % 8) * 16) = *reinterpret_cast<const uint4 *>(A + (size_t)r * K + c * 16); }
    for (int i = tid; i < N * K / 16; i += 128) { const int r = i / (K / 16), c = i % (K / 16); *reinterpret_cast<uint4 *>(sB + (r / 8) * kSBO + c * 128 + (r % 8) * 16) = *reinterpret_cast<const uint4 *>(B + (size_t)r * K + c * 16); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor-core (async) proxy
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    long long t0 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int ks = 0; ks < K / 32; ks++) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * kLBO), db = make_desc(smem_u32(sB) + ks * 2 * kLBO);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(kIdesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    {   // wait for the MMAs (phase 0 of the mbarrier)
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(mbar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) cycles[0] = clock64() - t0;
    // warp w reads TMEM lanes [32 w, 32 w + 32): thread = accumulator row, 64 columns in two loads of 32
    uint32_t r[64];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + h * 32;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[h * 32 + 0]), "=r"(r[h * 32 + 1]), "=r"(r[h * 32 + 2]), "=r"(r[h * 32 + 3]), "=r"(r[h * 32 + 4]), "=r"(r[h * 32 + 5]), "=r"(r[h * 32 + 6]), "=r"(r[h * 32 + 7]),
                       "=r"(r[h * 32 + 8]), "=r"(r[h * 32 + 9]), "=r"(r[h * 32 + 10]), "=r"(r[h * 32 + 11]), "=r"(r[h * 32 + 12]), "=r"(r[h * 32 + 13]), "=r"(r[h * 32 + 14]), "=r"(r[h * 32 + 15]),
                       "=r"(r[h * 32 + 16]), "=r"(r[h * 32 + 17]), "=r"(r[h * 32 + 18]), "=r"(r[h * 32 + 19]), "=r"(r[h * 32 + 20]), "=r"(r[h * 32 + 21]), "=r"(r[h * 32 + 22]), "=r"(r[h * 32 + 23]),
                       "=r"(r[h * 32 + 24]), "=r"(r[h * 32 + 25]), "=r"(r[h * 32 + 26]), "=r"(r[h * 32 + 27]), "=r"(r[h * 32 + 28]), "=r"(r[h * 32 + 29]), "=r"(r[h * 32 + 30]), "=r"(r[h * 32 + 31])
                     : "r"(ta) : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int row = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 64; j++) D[(size_t)row * N + j] = (int32_t)r[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
    std::vector<int8_t> hA(M * K), hB(N * K);
    srand(7);
    for (auto &v : hA) v = (int8_t)(rand() % 211 - 105);       // the range of q * sc_lo
    for (auto &v : hB) v = (int8_t)(rand() % 255 - 127);       // Q8_K activations
    int8_t *dA, *dB; int32_t *dD; long long *dc;
    cudaMalloc(&dA, hA.size()); cudaMalloc(&dB, hB.size()); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, M * N * 4);
    const int smem = M * K + N * K + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, dc);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<int32_t> hD(M * N); long long cyc = 0;
    cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            int ref = 0;
            for (int k = 0; k < K; k++) ref += (int)hA[i * K + k] * (int)hB[j * K + k];
            if (ref != hD[i * N + j] && bad++ < 5) printf("mismatch D[%d][%d] = %d, expected %d\n", i, j, hD[i * N + j], ref);
        }
    printf("tcgen05.mma kind::i8 %dx%dx%d: %s (%d mismatches), 8 MMAs + commit + wait = %lld cycles\n", M, N, K, bad ? "MISMATCH" : "exact", bad, cyc);
    return bad ? 2 : 0;
}
