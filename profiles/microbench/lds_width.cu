// Microbenchmark: what an LDS costs at saturation (32 warps per SM) as a function of the access width and of the number
// of active lanes -- is the bound per instruction, per byte returned to the register file, or per wavefront?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_width lds_width.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int W>  // W = bytes per lane: 4, 8, 16
__global__ void __launch_bounds__(1024) bench(int active, int iters, long long *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ long long t_w[32];
    for (int e = threadIdx.x; e < (64 * 1024) / 16; e += blockDim.x) reinterpret_cast<uint4 *>(smem)[e] = make_uint4(e, 0, 0, 0);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + (lane & 7u) * 16u + ((lane * 2654435761u + warp * 40503u) >> 7) % 111u * 384u;
    const bool on = (int)lane < active;
    uint32_t s = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (on) {
                if (W == 16) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + i * 16));
                    s ^= v.x ^ v.y ^ v.z ^ v.w;
                } else if (W == 8) {
                    uint2 v;
                    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(base + i * 16));
                    s ^= v.x ^ v.y;
                } else {
                    uint32_t v;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + i * 16));
                    s ^= v;
                }
            }
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
    if (s == 0x12345u) out[1] = 1;
}

template <int W>
static void run(long long *d) {
    long long r[2];
    cudaFuncSetAttribute(bench<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 400;
    const int actives[] = {32, 16, 8, 4, 1};
    for (int a : actives) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(d, 0, sizeof(r));
            bench<W><<<148, 1024, 64 * 1024>>>(a, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(r, d, sizeof(r), cudaMemcpyDeviceToHost);
            if (rep) printf("LDS.%-3d %2d active lanes (rotated random rows)  %s  %.2f SM-cycles per warp instruction = %.0f B/clk/SM\n", W * 8, a,
                            cudaGetErrorString(e), (double)r[0] / (32.0 * iters * 16), (double)a * W * 32.0 * iters * 16 / (double)r[0]);
        }
    }
}

int main() {
    long long *d;
    cudaMalloc(&d, 16);
    run<16>(d);
    run<8>(d);
    run<4>(d);
    return 0;
}
