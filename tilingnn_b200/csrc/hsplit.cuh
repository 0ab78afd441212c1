// fp16 two-term split of fp32 activations for k_conv_h (conv_h.cu):  x = hi + lo * 2^-11.
#pragma once

#include <cuda_fp16.h>

#include "tgnn_internal.h"

namespace tgnn {

// Packed fp32 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2, two IEEE round-to-nearest operations per instruction -- the same bits
// as the scalar forms, half the issue slots).  The GIN kernels are bound by instruction issue, not by a data pipe.
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else   // (host-side layout checks include this header for the index maps and compile for the default architecture)
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y)); }
#endif
__device__ __forceinline__ void f4add(float4& a, const float4& b) {
    const float2 lo = f2add(make_float2(a.x, a.y), make_float2(b.x, b.y)), hi = f2add(make_float2(a.z, a.w), make_float2(b.z, b.w));
    a = make_float4(lo.x, lo.y, hi.x, hi.y);
}
// c[i] = fma(s[i], k, a[i] + b[i]) for four accumulator registers
__device__ __forceinline__ void f4_fma_add(float (&c)[4], const float (&s)[4], float k, const float (&a)[4], const float (&b)[4]) {
    const float2 kk = make_float2(k, k);
    const float2 r0 = f2fma(make_float2(s[0], s[1]), kk, f2add(make_float2(a[0], a[1]), make_float2(b[0], b[1])));
    const float2 r1 = f2fma(make_float2(s[2], s[3]), kk, f2add(make_float2(a[2], a[3]), make_float2(b[2], b[3])));
    c[0] = r0.x; c[1] = r0.y; c[2] = r1.x; c[3] = r1.y;
}

// hi = fp16(x), lo = fp16((x - hi) * 2^11); the subtraction and the scaling are exact in fp32.
__device__ __forceinline__ void split_h(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * 2048.f);
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// Position of the uint4 holding channels 4q..4q+3 inside a split row: lane t of k_conv_h needs the pieces q = t (k16
// step 0) and q = 4 + t (step 1), so they are stored next to each other and come in with ONE 256-bit load.
__host__ __device__ __forceinline__ int xh_pos(int q) { return 2 * (q & 3) + (q >> 2); }

// four consecutive channels -> {hi(c0,c1), hi(c2,c3), lo(c0,c1), lo(c2,c3)}; `bad` is set when a value is outside
// the fp16 range (or NaN)
__device__ __forceinline__ uint4 split_h4(const float4& v, bool& bad) {
    __half h0, h1, h2, h3, l0, l1, l2, l3;
    split_h(v.x, h0, l0); split_h(v.y, h1, l1); split_h(v.z, h2, l2); split_h(v.w, h3, l3);
    bad |= !(fabsf(v.x) <= TG_H_LIMIT) | !(fabsf(v.y) <= TG_H_LIMIT) | !(fabsf(v.z) <= TG_H_LIMIT) | !(fabsf(v.w) <= TG_H_LIMIT);
    return make_uint4(pack_h2(h0, h1), pack_h2(h2, h3), pack_h2(l0, l1), pack_h2(l2, l3));
}

// mma.sync m16n8k16, fp16 operands, fp32 accumulate
__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// two values -> packed fp16 pairs {hi(v0), hi(v1)} and {lo(v0), lo(v1)}
__device__ __forceinline__ void split_h2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 f = __half22float2(h);
    // (v - hi) * 2^11 = fma(v, 2^11, -2^11 hi): both products and the difference are exact, so the bits are those of the scalar form
    const float2 d = f2fma(make_float2(v0, v1), make_float2(2048.f, 2048.f), f2mul(f, make_float2(-2048.f, -2048.f)));
    const __half2 l = __floats2half2_rn(d.x, d.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

}  // namespace tgnn
