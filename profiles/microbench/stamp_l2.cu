// Can a compact stamp array stay in L2 across a streaming kernel?  600k random 32-bit atomicMax into a 2 x 33.5 MB array
// (u32 stamps of two 2^23-slot reservoirs), timed (a) right after the array was touched, (b) after a kernel streamed 128 MB
// of writes through L2 (what the rollout does between two inserts), (c) the same with the array under a persisting
// access-policy window.  For comparison: the same atomics as 64-bit atomicMax into 32-byte slots (2 x 268 MB, today's layout).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stamp_l2 stamp_l2.cu && ./stamp_l2
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void red32(uint32_t *a, uint32_t n_slots, uint32_t n, uint32_t salt) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicMax(a + mix(i * 2654435761u + salt) % n_slots, i + salt);
}
__global__ void red64(unsigned long long *a, uint32_t n_slots, uint32_t n, uint32_t salt) {  // stamp = third word pair of a 32-byte slot
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicMax(a + 4ull * (mix(i * 2654435761u + salt) % n_slots) + 2, (unsigned long long)(i + salt));
}
// the same random slots: plain loads / L2 prefetches of the stamp words, and prefetch followed by the atomic
__global__ void ld64(const unsigned long long *a, uint32_t n_slots, uint32_t n, uint32_t salt, unsigned long long *out) {
    unsigned long long acc = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        acc += __ldcg(a + 4ull * (mix(i * 2654435761u + salt) % n_slots) + 2);
    if (acc == 0x123456789ull) *out = acc;
}
__global__ void pf64(const unsigned long long *a, uint32_t n_slots, uint32_t n, uint32_t salt) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 4ull * (mix(i * 2654435761u + salt) % n_slots) + 2));
}
__global__ void pf_red64(unsigned long long *a, uint32_t n_slots, uint32_t n, uint32_t salt) {
    const uint32_t stride = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = i0; i < n; i += stride)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 4ull * (mix(i * 2654435761u + salt) % n_slots) + 2));
    for (uint32_t i = i0; i < n; i += stride)
        atomicMax(a + 4ull * (mix(i * 2654435761u + salt) % n_slots) + 2, (unsigned long long)(i + salt));
}
__global__ void stream(uint4 *p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(i, 1, 2, 3);
}
__global__ void touch(uint32_t *a, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] += 1;
}

int main() {
    const uint32_t slots = 2u << 23, n = 600000;
    uint32_t *s32; unsigned long long *s64; uint4 *big;
    cudaMalloc(&s32, slots * 4ull); cudaMalloc(&s64, slots * 32ull); cudaMalloc(&big, 128ull << 20);
    cudaMemset(s32, 0, slots * 4ull); cudaMemset(s64, 0, slots * 32ull);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int maxp = 0, l2 = 0, maxw = 0;
    cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, 0);
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0);
    cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, 0);
    printf("L2 %d MB, max persisting %d MB, max window %d MB\n", l2 >> 20, maxp >> 20, maxw >> 20);
    auto time = [&](const char *what, auto f, bool streamed, bool warm) {
        float best = 1e9f, sum = 0;
        for (int r = 0; r < 6; ++r) {
            if (warm) touch<<<592, 256, 0, st>>>(s32, slots);
            if (streamed) stream<<<592, 256, 0, st>>>(big, (128ull << 20) / 16);
            cudaEventRecord(a, st); f(r); cudaEventRecord(b, st); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (r) { sum += ms; best = ms < best ? ms : best; }
        }
        printf("%-64s mean %.1f us  best %.1f us\n", what, sum / 5 * 1e3, best * 1e3);
    };
    time("u32 stamps, array just touched", [&](int r) { red32<<<592, 256, 0, st>>>(s32, slots, n, r * 7919u + 1); }, false, true);
    time("u32 stamps, 128 MB streamed since the touch", [&](int r) { red32<<<592, 256, 0, st>>>(s32, slots, n, r * 7919u + 1); }, true, true);
    time("u32 stamps, nothing in between (previous launch's lines)", [&](int r) { red32<<<592, 256, 0, st>>>(s32, slots, n, r * 7919u + 1); }, false, false);
    time("u32 stamps, 128 MB streamed, no touch", [&](int r) { red32<<<592, 256, 0, st>>>(s32, slots, n, r * 7919u + 1); }, true, false);
    time("u64 stamps in 32-byte slots (today), 128 MB streamed", [&](int r) { red64<<<592, 256, 0, st>>>(s64, slots, n, r * 7919u + 1); }, true, false);
    unsigned long long *sink; cudaMalloc(&sink, 8);
    time("u64 stamps in slots: plain loads of the same sectors, streamed", [&](int r) { ld64<<<592, 256, 0, st>>>(s64, slots, n, r * 7919u + 1, sink); }, true, false);
    time("u64 stamps in slots: L2 prefetches only, streamed", [&](int r) { pf64<<<592, 256, 0, st>>>(s64, slots, n, r * 7919u + 1); }, true, false);
    time("u64 stamps in slots: prefetch all, then the atomics, streamed", [&](int r) { pf_red64<<<592, 256, 0, st>>>(s64, slots, n, r * 7919u + 1); }, true, false);
    // persisting window over the u32 array
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp);
    cudaStreamAttrValue v = {};
    v.accessPolicyWindow.base_ptr = s32;
    v.accessPolicyWindow.num_bytes = (size_t)slots * 4 < (size_t)maxw ? (size_t)slots * 4 : (size_t)maxw;
    v.accessPolicyWindow.hitRatio = 1.0f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    printf("set window: %s\n", cudaGetErrorString(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v)));
    time("u32 stamps, persisting window, touched then 128 MB streamed", [&](int r) { red32<<<592, 256, 0, st>>>(s32, slots, n, r * 7919u + 1); }, true, true);
    time("u32 stamps, persisting window, 128 MB streamed, no touch", [&](int r) { red32<<<592, 256, 0, st>>>(s32, slots, n, r * 7919u + 1); }, true, false);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
