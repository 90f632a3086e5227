// mf_device.cuh -- device-side scalar semantics shared by every kernel.
//
// The reference's epilogue (src/ops/conv_2d.rs:93-98, depthwise_conv_2d.rs:90-95, fully_connected.rs:68-73) is
//     y = sat_T( roundf( (f32(out_zp) + c0[ch]) + (c1[ch] * f32(acc_i32)) ) )   then relu / relu6 on y
// in IEEE f32 with separate multiply and add (no FMA) and roundf = round half AWAY from zero.
// On the device:
//   * i32 -> f32:  __int2float_rn   (RN-even; |acc| can exceed 2^24 for K >= 1024)
//   * mul / add:   __fmul_rn / __fadd_rn (never contracted to FMA, whatever -fmad says)
//   * c0z = f32(out_zp) + c0[ch] is the same single f32 add, done once on the host at load time
//   * roundf + saturate + activation clamp:  trunc( clamp( t + copysign(0x3EFFFFFF, t), lo, hi ) ).
//     0x3EFFFFFF is the largest float below 0.5; trunc(t + copysign(0.49999997, t)) == roundf(t) for every finite
//     float t (exhaustively checked over all 2^32 bit patterns against glibc roundf), clamping commutes
//     with the monotone trunc, and sat_T followed by max(.,zp) / min(.,q6) is one clamp to [lo, hi] (lo/hi integers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mf {

__device__ __forceinline__ float round_bias(float t) {
    return __int_as_float(0x3EFFFFFF | (__float_as_int(t) & (int)0x80000000));
}

// acc -> quantized output value (as int in [lo, hi])
__device__ __forceinline__ int requant(int acc, float c0z, float c1, float lo, float hi) {
    float t = __fadd_rn(c0z, __fmul_rn(c1, __int2float_rn(acc)));
    float s = __fadd_rn(t, round_bias(t));
    s = fminf(fmaxf(s, lo), hi);
    return __float2int_rz(s);
}

// roundf(t) then saturate/clamp, for the f32 paths (quantize, pool, softmax)
__device__ __forceinline__ int round_clamp(float t, float lo, float hi) {
    float s = __fadd_rn(t, round_bias(t));
    s = fminf(fmaxf(s, lo), hi);
    return __float2int_rz(s);
}

template <bool U8> __device__ __forceinline__ int ld_elem(const uint8_t *p) {
    return U8 ? (int)(*p) : (int)(*(const int8_t *)p);
}

// pack four values already clamped to the int8 (or uint8) range into one word: 3 PRMT
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
    const uint32_t lo = __byte_perm((uint32_t)a, (uint32_t)b, 0x0040), hi = __byte_perm((uint32_t)c, (uint32_t)d, 0x0040);
    return __byte_perm(lo, hi, 0x5410);
}

// requant for layers whose clamp is the whole int8 range (every conv of person_detect: ReLU6's upper bound quantizes
// to 127 and the lower bound is the zero point -128): the saturating F2I.S8 conversion replaces both FMNMX.
// cvt.rzi.sat.s8.f32 == trunc then saturate to [-128,127], NaN -> 0 (exactly Rust's `as i8`).
__device__ __forceinline__ int requant_full_i8(int acc, float c0z, float c1) {
    float t = __fadd_rn(c0z, __fmul_rn(c1, __int2float_rn(acc)));
    float s = __fadd_rn(t, round_bias(t));
    int y;
    asm("cvt.rzi.sat.s8.f32 %0, %1;" : "=r"(y) : "f"(s));
    return y;
}
template <bool FULL> __device__ __forceinline__ int requant_t(int acc, float c0z, float c1, float lo, float hi) {
    return FULL ? requant_full_i8(acc, c0z, c1) : requant(acc, c0z, c1, lo, hi);
}

}  // namespace mf
