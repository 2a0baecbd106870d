// Second layer (64 -> 3) and output heads of the acting nets, shared by every forward variant.
// Uses Blackwell's packed fp32 FMA (fma.rn.f32x2 -> FFMA2): even and odd hidden units accumulate into
// the two halves of a 64-bit register, 96 FFMA2 instead of 192 FFMA per decision.
#pragma once
#include <cstdint>

namespace nfsp {

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void fma2(unsigned long long &acc, unsigned long long x, unsigned long long y) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(x), "l"(y));
}
__device__ __forceinline__ void unpack2(unsigned long long a, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
}
__device__ __forceinline__ float hsum2(unsigned long long a) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
    return lo + hi;
}

struct Layer2Acc {
    unsigned long long z0 = 0ull, z1 = 0ull, z2 = 0ull;  // (even, odd) partial sums of the three outputs

    // four pre-activations of hidden units 4q..4q+3 (relu applied here, agent.py:102,111) against the three
    // float4 weight quads W2[4q..4q+3][c]
    __device__ __forceinline__ void quad(float a, float b, float c, float d, const float4 &u0, const float4 &u1,
                                         const float4 &u2) {
        const unsigned long long hxy = pack2(fmaxf(a, 0.f), fmaxf(b, 0.f)), hzw = pack2(fmaxf(c, 0.f), fmaxf(d, 0.f));
        fma2(z0, hxy, pack2(u0.x, u0.y)); fma2(z0, hzw, pack2(u0.z, u0.w));
        fma2(z1, hxy, pack2(u1.x, u1.y)); fma2(z1, hzw, pack2(u1.z, u1.w));
        fma2(z2, hxy, pack2(u2.x, u2.y)); fma2(z2, hzw, pack2(u2.z, u2.w));
    }

    // same with the pre-activations given as the sum of two table rows (packed adds: x * 1 + y)
    __device__ __forceinline__ void quad_sum(const float4 &x, const float4 &y, const float4 &u0, const float4 &u1,
                                             const float4 &u2) {
        const unsigned long long one = pack2(1.f, 1.f);
        unsigned long long sxy = pack2(y.x, y.y), szw = pack2(y.z, y.w);
        fma2(sxy, pack2(x.x, x.y), one);
        fma2(szw, pack2(x.z, x.w), one);
        float a, b, c, d;
        unpack2(sxy, a, b);
        unpack2(szw, c, d);
        quad(a, b, c, d, u0, u1, u2);
    }

    // + b2, then relu head (best response, agent.py:103) or softmax head (average policy, agent.py:112)
    __device__ __forceinline__ void head(const float4 &b2, bool is_br, float &o0, float &o1, float &o2) const {
        head_of_sums(hsum2(z0), hsum2(z1), hsum2(z2), b2, is_br, o0, o1, o2);
    }
    __device__ __forceinline__ static void head_of_sums(float s0, float s1, float s2, const float4 &b2, bool is_br, float &o0,
                                                        float &o1, float &o2) {
        const float y0 = s0 + b2.x, y1 = s1 + b2.y, y2 = s2 + b2.z;
        if (is_br) {
            o0 = fmaxf(y0, 0.f); o1 = fmaxf(y1, 0.f); o2 = fmaxf(y2, 0.f);
        } else {
            const float m = fmaxf(y0, fmaxf(y1, y2));
            const float e0 = expf(y0 - m), e1 = expf(y1 - m), e2 = expf(y2 - m);
            const float inv = 1.0f / (e0 + e1 + e2);
            o0 = e0 * inv; o1 = e1 * inv; o2 = e2 * inv;
        }
    }
};

}  // namespace nfsp
