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

// ------------------------------------------------------------------------------------------------------------
// XU-free variants.  On sm_100 I2F / F2I issue to the XU pipe (measured: 4 and 8 clocks per warp instruction per sub-partition,
// tools/ubench/pipes.cu; IDP.4A, IMAD, FADD2 share the FMA-heavy pipe at 2).  The same results are obtained on the FMA / ALU
// pipes with exact bit tricks (each verified exhaustively on the host against (float)x / roundf -- see DESIGN.md):
//   * i2f_exact<false>(x), |x| <= 2^22 : as_float(x + 0x4B400000) - 1.5*2^23   (ulp is 1 in [2^23, 2^24))
//   * i2f_exact<true>(x), any int32    : x = hi*4096 + lo; hi*4096 and lo are built the same way in two binades and the
//                                        single RN add of the two exact parts is the correctly rounded value of x
//   * round_clamp_nx(t): clamp first (commutes with rounding), then u = RZ(|tc| + (1.5*2^22 + 0.5)) lands on the 0.5-grid
//     of [2^22, 2^23): its mantissa holds floor(2|tc| + 1), whose half is floor(|tc| + 0.5) = |roundf(tc)|.
// ------------------------------------------------------------------------------------------------------------
template <bool BIG> __device__ __forceinline__ float i2f_exact(int x) {
    if (!BIG) return __fadd_rn(__int_as_float(x + 0x4B400000), -12582912.0f);
    const float fh = __fadd_rn(__int_as_float((x >> 12) + 0x51400000), -51539607552.0f);   // hi * 4096 exactly
    const float fl = __fadd_rn(__int_as_float((x & 0xFFF) | 0x4B400000), -12582912.0f);     // lo exactly
    return __fadd_rn(fh, fl);
}
__device__ __forceinline__ int round_clamp_nx(float t, float lo, float hi) {
    const float tc = fminf(fmaxf(t, lo), hi);
    const float u = __fadd_rz(fabsf(tc), 6291456.5f);
    const int n = (__float_as_int(u) >> 1) & 0x1FF;
    return tc < 0.0f ? -n : n;
}
template <bool BIG> __device__ __forceinline__ int requant_nx(int acc, float c0z, float c1, float lo, float hi) {
    const float t = __fadd_rn(c0z, __fmul_rn(c1, i2f_exact<BIG>(acc)));
    return round_clamp_nx(t, lo, hi);
}
// XU variant for int8 outputs: I2F + F2I (two XU-pipe instructions, ~7 issue slots per value instead of 13-17).
// FULL: the clamp is the whole int8 range, so the saturating F2I.S8 replaces both FMNMX
// (cvt.rzi.sat.s8.f32 == trunc then saturate to [-128,127], NaN -> 0, exactly Rust's `as i8`).
template <bool FULL> __device__ __forceinline__ int requant_xu(int acc, float c0z, float c1, float lo, float hi) {
    if (!FULL) return requant(acc, c0z, c1, lo, hi);
    const float t = __fadd_rn(c0z, __fmul_rn(c1, __int2float_rn(acc)));
    const float s = __fadd_rn(t, round_bias(t));
    int y;
    asm("cvt.rzi.sat.s8.f32 %0, %1;" : "=r"(y) : "f"(s));
    return y;
}

// ------------------------------------------------------------------------------------------------------------
// Four values per call, fewest issue slots (the hot kernels are bound by issue slots, profiles/r01h_*):
//   * the accumulators arrive PRE-BIASED: a = 0x4B400000 + acc (the bias rides in the dp4a accumulator init or in the
//     correction term that is added anyway), |acc| <= 2^22, so as_float(a) - 1.5*2^23 == float(acc) exactly -- one packed
//     FADD2 per two values instead of one I2F (XU pipe) each;
//   * the two epilogue adds are packed FADD2 (add.rn.f32x2: two independent IEEE RN adds, one issue slot).  The multiply
//     stays scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false (checked in SASS),
//     which would round once instead of twice and break bit-exactness; FMUL + FADD2 is never contracted.
//   4 values: 2 FADD2 + 4 FMUL + 2 FADD2 + 4 LOP3 + 2 FADD2 + 2 F2IP.S8.F32 (f2i_pack4 below) = 16 issue slots and no XU-pipe
//   instruction (round 1: 4 F2I.S8 + 3 PRMT in place of the 2 F2IP = 21 slots and 32 XU clocks).
// ------------------------------------------------------------------------------------------------------------
constexpr int kAccBias = 0x4B400000;   // bits of 1.5 * 2^23 = 12582912.0f

__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
template <bool FULL> __device__ __forceinline__ int round_sat_s8(float s, float lo, float hi) {
    if (!FULL) s = fminf(fmaxf(s, lo), hi);
    int y;
    asm("cvt.rzi.sat.s8.f32 %0, %1;" : "=r"(y) : "f"(s));      // trunc, saturate to int8, NaN -> 0 (Rust's `as i8`)
    return y;
}
// Four floats -> four saturated bytes in one word (byte k = value k): trunc toward zero, saturate, NaN -> 0 -- bit for bit what four
// cvt.rzi.sat.s8.f32 (F2I.S8) and three PRMT produce (checked over all 2^32 float patterns, tools/ubench/f2ip.cu).  ptxas fuses each
// {cvt.rzi.s32.f32 x2, cvt.pack.sat.s8.s32.b32} into ONE F2IP.S8.F32.TRUNC.NTZ, which does not run on the quarter-rate XU pipe:
// 2 issue slots per four values instead of 4 F2I.S8 (8 clocks each per sub-partition) + 3 PRMT (tests/test_abi_host.py pins the SASS).
template <bool U8 = false> __device__ __forceinline__ uint32_t f2i_pack4(float a, float b, float c, float d) {
    int ia, ib, ic, id;
    uint32_t hi, r;
    asm("cvt.rzi.s32.f32 %0, %1;" : "=r"(ia) : "f"(a));
    asm("cvt.rzi.s32.f32 %0, %1;" : "=r"(ib) : "f"(b));
    asm("cvt.rzi.s32.f32 %0, %1;" : "=r"(ic) : "f"(c));
    asm("cvt.rzi.s32.f32 %0, %1;" : "=r"(id) : "f"(d));
    if (U8) {
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(id), "r"(ic), "r"(0));
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(ib), "r"(ia), "r"(hi));
    } else {
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(id), "r"(ic), "r"(0));
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(ib), "r"(ia), "r"(hi));
    }
    return r;
}
// two pre-biased accumulators -> two rounded values (still float; the caller converts four at a time with f2i_pack4)
template <bool FULL> __device__ __forceinline__ float2 requant2_biased_f(int a0, int a1, float z0, float z1, float s0, float s1, float lo, float hi) {
    const float2 f = fadd2(make_float2(__int_as_float(a0), __int_as_float(a1)), make_float2(-12582912.0f, -12582912.0f));
    const float2 t = fadd2(make_float2(z0, z1), make_float2(__fmul_rn(s0, f.x), __fmul_rn(s1, f.y)));
    float2 r = fadd2(t, make_float2(round_bias(t.x), round_bias(t.y)));
    if (!FULL) { r.x = fminf(fmaxf(r.x, lo), hi); r.y = fminf(fmaxf(r.y, lo), hi); }
    return r;
}
template <bool FULL> __device__ __forceinline__ uint32_t requant4_biased(int a0, int a1, int a2, int a3, float4 z, float4 s, float lo, float hi) {
    const float2 r01 = requant2_biased_f<FULL>(a0, a1, z.x, z.y, s.x, s.y, lo, hi);
    const float2 r23 = requant2_biased_f<FULL>(a2, a3, z.z, z.w, s.z, s.w, lo, hi);
    return f2i_pack4(r01.x, r01.y, r23.x, r23.y);
}

// same packed adds for accumulators of any magnitude (I2F on the XU pipe instead of the bias trick); full int8 clamp only
__device__ __forceinline__ uint32_t requant4_i2f(int a0, int a1, int a2, int a3, float4 z, float4 s) {
    const float2 t01 = fadd2(make_float2(z.x, z.y), make_float2(__fmul_rn(s.x, __int2float_rn(a0)), __fmul_rn(s.y, __int2float_rn(a1))));
    const float2 t23 = fadd2(make_float2(z.z, z.w), make_float2(__fmul_rn(s.z, __int2float_rn(a2)), __fmul_rn(s.w, __int2float_rn(a3))));
    const float2 r01 = fadd2(t01, make_float2(round_bias(t01.x), round_bias(t01.y)));
    const float2 r23 = fadd2(t23, make_float2(round_bias(t23.x), round_bias(t23.y)));
    return f2i_pack4(r01.x, r01.y, r23.x, r23.y);
}

// general form: any clamp range [lo, hi] inside the output type's range, int8 or uint8 outputs (u8 is a warp-uniform run-time flag),
// accumulators of any magnitude (BIG: I2F on the now otherwise idle XU pipe; else the exact bias trick).  Same arithmetic as requant():
// t = c0z + c1 * f32(acc); s = t + copysign(0.49999997, t); clamp; truncate.  ~8 issue slots per value less than requant_nx + pack4.
template <bool BIG> __device__ __forceinline__ uint32_t requant4_clamp(int a0, int a1, int a2, int a3, float4 z, float4 s, float lo, float hi, bool u8) {
    const float f0 = BIG ? __int2float_rn(a0) : i2f_exact<false>(a0), f1 = BIG ? __int2float_rn(a1) : i2f_exact<false>(a1);
    const float f2 = BIG ? __int2float_rn(a2) : i2f_exact<false>(a2), f3 = BIG ? __int2float_rn(a3) : i2f_exact<false>(a3);
    const float2 t01 = fadd2(make_float2(z.x, z.y), make_float2(__fmul_rn(s.x, f0), __fmul_rn(s.y, f1)));
    const float2 t23 = fadd2(make_float2(z.z, z.w), make_float2(__fmul_rn(s.z, f2), __fmul_rn(s.w, f3)));
    const float2 r01 = fadd2(t01, make_float2(round_bias(t01.x), round_bias(t01.y)));
    const float2 r23 = fadd2(t23, make_float2(round_bias(t23.x), round_bias(t23.y)));
    const float c0 = fminf(fmaxf(r01.x, lo), hi), c1 = fminf(fmaxf(r01.y, lo), hi), c2 = fminf(fmaxf(r23.x, lo), hi), c3 = fminf(fmaxf(r23.y, lo), hi);
    return u8 ? f2i_pack4<true>(c0, c1, c2, c3) : f2i_pack4<false>(c0, c1, c2, c3);
}

// sign-extended byte k of a packed word: one PRMT (selector msb = replicate the sign of the selected byte)
template <int K> __device__ __forceinline__ int sx8(uint32_t w) {
    int r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "r"((uint32_t)(((8 | K) * 0x1110) | K)));
    return r;
}

// zero-extended byte k of a packed word (uint8 tensors, `T = u8` of the reference's `Quantized` trait, src/quantize.rs:6-7)
template <int K> __device__ __forceinline__ int zx8(uint32_t w) {
    int r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "r"((uint32_t)(0x4440 | K)));
    return r;
}
template <bool U8, int K> __device__ __forceinline__ int ext8(uint32_t w) { return U8 ? zx8<K>(w) : sx8<K>(w); }
// four byte products summed into acc, for int8 or uint8 operands
template <bool U8> __device__ __forceinline__ int dot4(uint32_t a, uint32_t b, int acc) {
    return U8 ? (int)__dp4a(a, b, (unsigned)acc) : __dp4a((int)a, (int)b, acc);
}

// ------------------------------------------------------------------------------------------------------------
// byte shuffles of the depthwise kernels (mf_kernels.cu, mf_fused.cu)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <uint32_t SEL> __device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "n"(SEL));
    return r;
}
// four words (one per column, 4 channels each; the 4th is ignored) -> per-channel (col0, col1, col2, don't-care) registers
__device__ __forceinline__ void transpose_3x4(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t (&t)[4]) {
    const uint32_t lo = prmt<0x5140>(v0, v1), hi = prmt<0x7362>(v0, v1);   // (v0.0 v1.0 v0.1 v1.1), (v0.2 v1.2 v0.3 v1.3)
    t[0] = prmt<0x4410>(lo, v2);
    t[1] = prmt<0x5532>(lo, v2);
    t[2] = prmt<0x6610>(hi, v2);
    t[3] = prmt<0x7732>(hi, v2);
}
__device__ __forceinline__ void transpose_4x4(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t (&t)[4]) {
    const uint32_t lo01 = prmt<0x5140>(v0, v1), hi01 = prmt<0x7362>(v0, v1);
    const uint32_t lo23 = prmt<0x5140>(v2, v3), hi23 = prmt<0x7362>(v2, v3);
    t[0] = prmt<0x5410>(lo01, lo23);
    t[1] = prmt<0x7632>(lo01, lo23);
    t[2] = prmt<0x5410>(hi01, hi23);
    t[3] = prmt<0x7632>(hi01, hi23);
}

}  // namespace mf
