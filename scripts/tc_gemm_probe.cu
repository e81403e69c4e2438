// tc_gemm_probe.cu — the tcgen05 prefill GEMM (moshi.cpp_b200/csrc/tc_gemm.cuh) alone on synthetic blocks: launch time per shape and a
// per-iteration timeline of CTA 0 (clock64 stamps compiled in with MSX_TC_TIMELINE).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DMSX_TC_TIMELINE -I moshi.cpp_b200/csrc -o scripts/tc_gemm_probe scripts/tc_gemm_probe.cu
#include <cstdio>
#include <vector>
#include <algorithm>
#include "common.cuh"
#include "tc_gemm.cuh"
using namespace msx;
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(tc::tc_matmul_q4k_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
    long long *tl; cudaMalloc(&tl, 64 * 4 * 8); cudaMemset(tl, 0, 64 * 4 * 8);
    cudaMemcpyToSymbol(tc::g_tc_timeline, &tl, sizeof(tl));
    struct Shape { const char *name; int rows, K, epi; } shapes[] = {
        {"in_proj 12288 x 4096", 12288, 4096, EPI_STORE}, {"out_proj 4096 x 4096", 4096, 4096, EPI_RESID},
        {"linear_in 22528 x 4096", 22528, 4096, EPI_GATE}, {"linear_out 4096 x 11264", 4096, 11264, EPI_RESID}};
    for (const Shape &sh : shapes) {
        const int nsb = sh.K / 256, tiles = sh.rows / tc::kM;
        const size_t wbytes = (size_t)sh.rows * nsb * 144;
        uint8_t *w, *img; float *out; double *partial; unsigned int *tickets;
        cudaMalloc(&w, wbytes * 4); cudaMemset(w, 0x11, wbytes * 4);               // 4 rotating copies (> L2 for the big shapes)
        cudaMalloc(&img, tc::image_bytes(sh.K)); cudaMemset(img, 0x01, tc::image_bytes(sh.K));
        cudaMalloc(&out, (size_t)64 * sh.rows * 4); cudaMemset(out, 0, (size_t)64 * sh.rows * 4);
        cudaMalloc(&partial, tc::partial_bytes(sms)); cudaMalloc(&tickets, tiles * 4); cudaMemset(tickets, 0, tiles * 4);
        tc::TcMatmulArgs g;
        g.K = sh.K; g.rows = sh.rows; g.img = img; g.out = out; g.ld = sh.epi == EPI_GATE ? sh.rows / 2 : sh.rows; g.nb = 64; g.epi = sh.epi;
        g.partial = partial; g.tickets = tickets;
        const int grid = tc::grid_for(tiles, nsb, sms);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            for (int i = 0; i < 20; i++) { g.w = w + (size_t)(i & 3) * wbytes; tc::tc_matmul_q4k_kernel<64><<<grid, tc::kThreads, tc::kSmemBytes>>>(g); }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        printf("%-26s grid %3d, %5.2f steps per CTA: %7.2f us per launch (%s)\n", sh.name, grid, (double)tiles * nsb / grid, ms * 1e3 / 20, cudaGetErrorString(err));
        long long h[64 * 4]; cudaMemcpy(h, tl, sizeof(h), cudaMemcpyDeviceToHost);
        const int n = std::min(4, (int)((long long)tiles * nsb / grid));
        for (int i = 0; i < n; i++)
            printf("   it %2d: top->acc ready %6lld  fold %6lld  expand %6lld  barrier + loop end %6lld  | iteration %6lld clk\n", i, h[i * 4 + 1] - h[i * 4],
                   h[i * 4 + 2] - h[i * 4 + 1], h[i * 4 + 3] - h[i * 4 + 2], i + 1 < n ? h[(i + 1) * 4] - h[i * 4 + 3] : 0LL, i + 1 < n ? h[(i + 1) * 4] - h[i * 4] : 0LL);
        printf("   CTA 0: start-up %lld clk, loop + flushes %lld clk; flush inside the loop %lld clk, last flush %lld clk\n", h[60 * 4 + 1] - h[60 * 4], h[60 * 4 + 2] - h[60 * 4 + 1],
               h[61 * 4 + 1] - h[61 * 4], h[62 * 4 + 1] - h[62 * 4]);
        cudaFree(w); cudaFree(img); cudaFree(out); cudaFree(partial); cudaFree(tickets);
    }
    return 0;
}
