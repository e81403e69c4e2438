// sm_stream_probe.cu — dev probe: how fast can N SMs (N = 1 ... 148) pull bytes through cp.async.bulk (TMA) + mbarrier rings?
// One CTA per SM, one issuing thread, a ring of 8 x 18 KB slots refilled as soon as a copy lands (no compute).  Reports GB/s per
// SM and in total, for a buffer far larger than L2 (HBM) and one that fits L2.  Used to size cluster-resident designs
// (DESIGN.md section 6): a depformer layer on a 16-CTA cluster is bounded by what 16 SMs can stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -I moshi.cpp_b200/csrc -o /tmp/sm_stream_probe scripts/sm_stream_probe.cu
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

using namespace msx;
constexpr int kSlots = 8, kSlotBytes = 18432;

__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t *src, size_t bytes_per_cta, size_t stride, int chunks, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem), bars = ring + kSlots * kSlotBytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; s++) mbar_init(bars + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint8_t *p = src + (size_t)blockIdx.x * stride;
        const long long t0 = global_ns();
        int issued = 0;
        for (; issued < kSlots && issued < chunks; issued++) {
            mbar_expect_tx(bars + issued * 8, kSlotBytes);
            bulk_g2s(ring + issued * kSlotBytes, p + ((size_t)issued * kSlotBytes) % bytes_per_cta, kSlotBytes, bars + issued * 8);
        }
        for (int c = 0; c < chunks; c++) {
            const int s = c % kSlots;
            mbar_wait(bars + s * 8, (uint32_t)((c / kSlots) & 1));
            if (issued < chunks) {
                mbar_expect_tx(bars + s * 8, kSlotBytes);
                bulk_g2s(ring + s * kSlotBytes, p + ((size_t)issued * kSlotBytes) % bytes_per_cta, kSlotBytes, bars + s * 8);
                issued++;
            }
        }
        out[blockIdx.x] = global_ns() - t0;
    }
}

int main() {
    const size_t big = (size_t)148 * 64 * 1024 * 1024;        // 9.9 GB: 64 MB per CTA, never re-read inside L2
    uint8_t *buf; long long *d_out;
    if (cudaMalloc(&buf, big) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 1, big);
    cudaMalloc(&d_out, 148 * 8);
    const int smem = kSlots * kSlotBytes + 128;
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[148];
    for (int pass = 0; pass < 2; pass++) {
        const size_t per = pass == 0 ? (size_t)64 * 1024 * 1024 : (size_t)288 * 1024;      // HBM : L2-resident (288 KB x 148 = 42 MB)
        const int chunks = pass == 0 ? 1800 : 4000;                                        // 33 MB / 74 MB per CTA
        for (int n : {1, 2, 4, 8, 16, 32, 37, 74, 148}) {
            stream_kernel<<<n, 64, smem>>>(buf, per - per % kSlotBytes, (size_t)64 * 1024 * 1024, chunks, d_out);   // warm (L2 pass: fills L2)
            stream_kernel<<<n, 64, smem>>>(buf, per - per % kSlotBytes, (size_t)64 * 1024 * 1024, chunks, d_out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(h, d_out, n * 8, cudaMemcpyDeviceToHost);
            long long worst = 0;
            for (int i = 0; i < n; i++) worst = h[i] > worst ? h[i] : worst;
            const double gb = (double)chunks * kSlotBytes / 1e9;
            printf("%s  %3d CTAs: %6.1f GB/s per SM (slowest CTA), %7.1f GB/s total\n", pass == 0 ? "HBM" : "L2 ", n, gb / (worst * 1e-9), n * gb / (worst * 1e-9));
        }
    }
    return 0;
}
