// attn_probe.cu — dev harness: the ring attention kernel (csrc/attention.cuh) alone, at 7B shapes with a full ring, with its
// optional in-kernel timeline.  Prints per-stage times (median over CTAs / slowest CTA) and the launch time from CUDA events with
// the ring colder than L2 (32 layers of rings rotate, 1.57 GB).
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -I moshi.cpp_b200/csrc -o /tmp/attn_probe scripts/attn_probe.cu && /tmp/attn_probe [n_valid] [split] [small_ctx]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "attention.cuh"

using namespace msx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char **argv) {
    const int n_valid = argc > 1 ? atoi(argv[1]) : 3000, split = argc > 2 ? atoi(argv[2]) : 8, small_ctx = argc > 3 ? atoi(argv[3]) : 32;
    const int H = 32, DH = 128, cap = 3000, dim = H * DH, L = 32;
    const size_t ring = (size_t)H * cap * DH;
    uint16_t *kc, *vc; float *qkv, *ctx; Ctrl *ctrl; long long *dbg;
    CK(cudaMalloc(&kc, ring * 2 * L)); CK(cudaMalloc(&vc, ring * 2 * L));
    CK(cudaMalloc(&qkv, 3 * dim * 4)); CK(cudaMalloc(&ctx, dim * 4)); CK(cudaMalloc(&ctrl, sizeof(Ctrl)));
    CK(cudaMalloc(&dbg, (size_t)H * split * 8 * 8));
    {   // bf16 noise in [-1, 1): 0x3f80 = 1.0
        std::vector<uint16_t> h(ring);
        for (size_t i = 0; i < ring; i++) h[i] = (uint16_t)(0x3c00 + (rand() & 0x3ff) + ((rand() & 1) << 15));
        for (int l = 0; l < L; l++) { CK(cudaMemcpy(kc + l * ring, h.data(), ring * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(vc + l * ring, h.data(), ring * 2, cudaMemcpyHostToDevice)); }
        std::vector<float> q(3 * dim);
        for (auto &v : q) v = (rand() & 0xffff) / 65536.f - 0.5f;
        CK(cudaMemcpy(qkv, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
        Ctrl c; memset(&c, 0, sizeof(c)); c.offset = n_valid - 1 + (n_valid >= cap ? cap : 0);
        CK(cudaMemcpy(ctrl, &c, sizeof(c), cudaMemcpyHostToDevice));
    }
    CK(cudaFuncSetAttribute(attn_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<128>(cap, split)));
    AttnArgs a; a.qkv = qkv; a.ctx = ctx; a.ctrl = ctrl; a.cap = cap; a.dim = dim; a.max_period = 0; a.small_ctx = small_ctx;
    auto launch = [&](int layer, long long *d) {
        a.kc = kc + layer * ring; a.vc = vc + layer * ring; a.dbg = d;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(split, H, 1); cfg.blockDim = dim3(kThreads, 1, 1); cfg.dynamicSmemBytes = attn_smem_bytes<128>(cap, split);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = split; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, attn_kernel<128, true>, a);
    };
    for (int i = 0; i < 64; i++) CK(launch(i % L, nullptr));
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 320;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) launch(i % L, nullptr);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = 2.0 * H * (double)std::min(n_valid, cap) * DH * 2;
    printf("n_valid %d split %d small_ctx %d smem %d B: %.2f us / launch back to back (%.0f GB/s of K+V)\n", n_valid, split, small_ctx, attn_smem_bytes<128>(cap, split), ms * 1e3 / reps, bytes / (ms * 1e-3 / reps) / 1e9);
    CK(cudaMemset(dbg, 0, (size_t)H * split * 64));
    CK(launch(7, dbg)); CK(cudaDeviceSynchronize());
    std::vector<long long> h((size_t)H * split * 8);
    CK(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
    long long t0 = h[0];
    for (size_t c = 0; c < (size_t)H * split; c++) t0 = std::min(t0, h[c * 8]);
    const char *names[6] = {"start", "scores done", "max known", "probs known", "context done", "end"};
    for (int s = 0; s < 6; s++) {
        std::vector<double> v;
        for (size_t c = 0; c < (size_t)H * split; c++) if (h[c * 8 + s]) v.push_back((h[c * 8 + s] - t0) * 1e-3);
        if (v.empty()) continue;
        std::sort(v.begin(), v.end());
        printf("  %-13s first %6.2f  median %6.2f  last %6.2f us\n", names[s], v.front(), v[v.size() / 2], v.back());
    }
    return 0;
}
