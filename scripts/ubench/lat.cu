// Dependent-chain latencies of the FP64 / select instructions the exact fill is made of (one warp).
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* cyc, double a0, double b0, double c0)
{
    double a = a0 + threadIdx.x, b = b0, c = c0;
    long long t0, t1;
    // DADD chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) a = a + b;
    t1 = clock64(); cyc[0] = t1 - t0;
    // DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) a = __fma_rn(a, c, b);
    t1 = clock64(); cyc[1] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) a = a * c;
    t1 = clock64(); cyc[2] = t1 - t0;
    // compare+select chain: a = (a > x_i) ? a : x_i with x_i = b + i (independent values)
    double x = b;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) { x = x + 1.0; a = (a > x) ? a : x; }
    t1 = clock64(); cyc[3] = t1 - t0;
    // DADD + compare+select (the fill's horizontal chain)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) { double s = a + c; a = (s >= b) ? s : b; }
    t1 = clock64(); cyc[4] = t1 - t0;
    // integer-key compare + select on the same data
    long long ka = __double_as_longlong(a), kb = __double_as_longlong(b);
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) { kb += 3; ka = (ka > kb) ? ka : kb; }
    t1 = clock64(); cyc[5] = t1 - t0;
    // 4 independent DFMA chains (throughput with ILP 4)
    double p = a, q = a + 1, r = a + 2, s = a + 3;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) { p = __fma_rn(p, c, b); q = __fma_rn(q, c, b); r = __fma_rn(r, c, b); s = __fma_rn(s, c, b); }
    t1 = clock64(); cyc[6] = t1 - t0;
    // 8 independent DFMA chains
    double p2 = a + 4, q2 = a + 5, r2 = a + 6, s2 = a + 7;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; i++) { p = __fma_rn(p, c, b); q = __fma_rn(q, c, b); r = __fma_rn(r, c, b); s = __fma_rn(s, c, b);
                                  p2 = __fma_rn(p2, c, b); q2 = __fma_rn(q2, c, b); r2 = __fma_rn(r2, c, b); s2 = __fma_rn(s2, c, b); }
    t1 = clock64(); cyc[7] = t1 - t0;
    out[threadIdx.x + blockIdx.x * blockDim.x] = a + x + __longlong_as_double(ka) + p + q + r + s + p2 + q2 + r2 + s2;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
    const char* names[] = {"DADD dep", "DFMA dep", "DMUL dep", "cmp+sel dep (DSETP+2FSEL, +indep DADD)", "DADD+cmp+sel dep", "int64 cmp+sel dep", "4xDFMA indep (per group)", "8xDFMA indep (per group)"};
    for (int warps = 1; warps <= 16; warps *= 2)
    {
        k<<<1, 32 * warps>>>(out, cyc, 1.0, 1e-3, 0.999);
        cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("warps per SM = %d (per SMSP %g)\n", warps, warps / 4.0);
        for (int i = 0; i < 8; i++) printf("  %-45s %.2f cycles/iter\n", names[i], (double)h[i] / N);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
