// fp64_rate_probe.cu — sustained issue rate of the double-precision instructions the fold of tc_gemm.cuh uses (DFMA, DADD, F2F.F64.F32),
// in lanes per clock per SM, 1 CTA of 512 or 1024 threads per SM.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate_probe fp64_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double *out, float *fin, int iters, long long *clk) {
    double a[8]; float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-3 + i; f[i] = fin[(threadIdx.x + i) & 255]; }
    const double m = out[0], c = out[1];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = fma(a[i], m, c);
            if (OP == 1) a[i] = a[i] + c;
            if (OP == 2) { a[i] += (double)f[i]; f[i] = f[i] * 1.0001f; }          // F2F + DADD + FMUL
            if (OP == 3) { f[i] = fmaf(f[i], 1.0001f, 0.5f); }                      // reference: FFMA
            if (OP == 4) { a[i] = __hiloint2double(0x43300000, __float_as_int(f[i])) - c; f[i] = f[i] * 1.0001f; }   // DADD with integer-built input
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; i++) s += a[i] + f[i];
    out[2 + blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
    double *out; float *fin; long long *clk;
    cudaMalloc(&out, (2 + 148 * 1024) * 8); cudaMalloc(&fin, 1024); cudaMalloc(&clk, 148 * 8);
    double h[2] = {1.0000001, 1e-9}; cudaMemcpy(out, h, 16, cudaMemcpyHostToDevice); cudaMemset(fin, 0x3f, 1024);
    const int iters = 4096;
    const char *names[5] = {"DFMA", "DADD", "F2F.F64.F32 + DADD + FMUL", "FFMA", "int-built DADD + FMUL"};
    for (int threads : {512, 1024})
        for (int op = 0; op < 5; op++) {
            for (int rep = 0; rep < 2; rep++) {
                if (op == 0) k<0><<<148, threads>>>(out, fin, iters, clk);
                if (op == 1) k<1><<<148, threads>>>(out, fin, iters, clk);
                if (op == 2) k<2><<<148, threads>>>(out, fin, iters, clk);
                if (op == 3) k<3><<<148, threads>>>(out, fin, iters, clk);
                if (op == 4) k<4><<<148, threads>>>(out, fin, iters, clk);
                cudaDeviceSynchronize();
            }
            long long c0; cudaMemcpy(&c0, clk, 8, cudaMemcpyDeviceToHost);
            printf("%4d threads  %-28s %8.2f thread-iterations (x8 ops) per clock per SM -> %6.1f op-lanes / clk / SM\n", threads, names[op],
                   (double)threads * iters / c0, (double)threads * iters * 8 / c0);
        }
    return 0;
}
