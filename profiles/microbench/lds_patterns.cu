// Microbenchmark: SM-cycles per warp-wide LDS.128 at saturation (32 warps per SM) for the access patterns the rollout
// kernels use.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

// pattern -> byte offset of the lane's first quad; every iteration reads 16 quads at +16 B steps (immediate offsets)
__device__ __forceinline__ uint32_t lane_base(int pattern, uint32_t lane, uint32_t warp) {
    const uint32_t h = (lane * 2654435761u + warp * 40503u) >> 7;  // fixed pseudo-random per lane
    switch (pattern) {
        case 0: return lane * 16u;                                     // contiguous 512 B, conflict-free (stride 32 quads per read)
        case 1: return 0u;                                             // warp-uniform address
        case 2: return (lane & 7u) * 16u;                              // rotated, one row for the whole warp (8 distinct quads)
        case 3: return (h & 3u) * 1152u + (lane & 7u) * 16u;           // rotated, 4 rows mixed over the lanes (W2 of 4 nets)
        case 4: return (h % 111u) * 384u + (lane & 7u) * 16u;          // rotated, a random row of 111 per lane (X / Y gathers)
        case 5: return (h % 111u) * 384u;                              // straight (no rotation), random rows
        case 6: return (lane >> 4) * 1152u + (lane & 7u) * 16u;        // rotated, 2 rows (half-warps)
        case 7: return (h % 12u) * 384u + (lane & 7u) * 16u;           // rotated, 12 distinct rows (X rows of the new layout)
        default: return (h % 25u) * 384u + (lane & 7u) * 16u;          // rotated, 25 distinct rows (Y rows of the new layout)
    }
}

__global__ void __launch_bounds__(1024) bench(int pattern, int iters, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ long long t_w[32];
    for (int e = threadIdx.x; e < (64 * 1024) / 16; e += blockDim.x) reinterpret_cast<uint4 *>(smem)[e] = make_uint4(e, 0, 0, 0);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 *p = reinterpret_cast<const uint4 *>(smem + lane_base(pattern, lane, warp));
    const int step = pattern == 0 ? 32 : 1;
    uint4 s = make_uint4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"((uint32_t)__cvta_generic_to_shared(p + i * step)));
            s.x ^= v.x; s.y ^= v.y; s.z ^= v.z; s.w ^= v.w;
        }
    }
    const long long dt = clock64() - t0;
    if (lane == 0) t_w[warp] = dt;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long m = 0;
        for (int k = 0; k < 32; ++k) m = t_w[k] > m ? t_w[k] : m;
        out[0] = m;
    }
    if ((s.x ^ s.y ^ s.z ^ s.w) == 0x12345u) out[1] = 1;
}

int main() {
    long long *d, r[2];
    cudaMalloc(&d, sizeof(r));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char *names[] = {"contiguous 512 B", "warp-uniform", "rotated, 1 row", "rotated, 4 rows mixed", "rotated, random of 111 rows",
                           "straight, random of 111 rows", "rotated, 2 rows by half-warp", "rotated, random of 12 rows", "rotated, random of 25 rows"};
    const int iters = 400;
    for (int rep = 0; rep < 2; ++rep)
        for (int pat = 0; pat < 9; ++pat) {
            cudaMemset(d, 0, sizeof(r));
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            bench<<<148, 1024, 64 * 1024>>>(pat, iters, d);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(r, d, sizeof(r), cudaMemcpyDeviceToHost);
            printf("%-32s %s  %.2f SM-cycles per warp LDS.128 (32 warps x %d x 16 loads in %lld cycles; kernel %.1f us = %.0f cycles at 1.965 GHz)\n", names[pat],
                   cudaGetErrorString(e), (double)r[0] / (32.0 * iters * 16), iters, r[0], ms * 1e3, ms * 1.965e6);
        }
    return 0;
}
