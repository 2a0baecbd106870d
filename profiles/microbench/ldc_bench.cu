// Microbenchmark: throughput of indexed constant loads (LDC.64 c[bank][R + imm]) when the lanes of a warp use
// 1, 2 or 4 distinct indices -- would the constant cache be a second path for the per-net W2 of the acting nets?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldc_bench ldc_bench.cu && ./ldc_bench
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float2 c_tab[4][512];
__global__ void k(int distinct, int iters, float *out, long long *cycles) {
    const int net = (threadIdx.x & 31) % distinct;
    float a = 0.f, b = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 96; ++q) {
            const float2 u = c_tab[net][(q + it) & 511];
            a = fmaf(u.x, 1.0001f, a);
            b = fmaf(u.y, 0.9999f, b);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {1, 8, 32}) {
        for (int distinct : {1, 2, 4, 32}) {
            const int iters = 200;
            k<<<148, warps * 32>>>(distinct, iters, out, cyc);
            cudaDeviceSynchronize();
            k<<<148, warps * 32>>>(distinct, iters, out, cyc);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            printf("warps/SM %2d distinct %2d : %.2f cycles per LDC.64 per warp, %.2f SM-cycles per warp-LDC\n", warps, distinct,
                   avg / (iters * 96.0), avg / (iters * 96.0 * warps));
        }
    }
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
}
