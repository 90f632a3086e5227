// Issue-rate microbenchmark for the instruction mixes the depthwise / epilogue code is built from (sm_100a).
// Prints clocks per warp-instruction group per SM sub-partition for each mix.  Build: nvcc -arch=sm_100a -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { FFMA2, FFMA, FADD2, PRMT, LOP3, IMAD, IADD, DP4A, I2F, F2I, FMNMX, SHF, ISETP_SEL, FADD, LDS,
       FFMA2_FFMA, FFMA2_LOP3, PRMT_FFMA, DP4A_PRMT, DP4A_FFMA, DP4A_IMAD, I2F_F2I, F2I_FFMA, I2F_PRMT, PRMT_IADD, DP4A_LOP3, FFMA_IMAD, DP4A_F2I, F2I32, I2IP, F2I32_I2IP, FADD2_DP4A, FADD2_PRMT, FADD2_FADD, F2I_DP4A_PRMT, NMODES };
static const char *names[NMODES] = {"FFMA2", "FFMA", "FADD2", "PRMT", "LOP3", "IMAD", "IADD3", "IDP4A", "I2F.S32", "F2I.S8", "FMNMX", "SHF", "ISETP+SEL", "FADD", "LDS",
       "FFMA2+FFMA", "FFMA2+LOP3", "PRMT+FFMA", "IDP4A+PRMT", "IDP4A+FFMA", "IDP4A+IMAD", "I2F+F2I", "F2I+FFMA", "I2F+PRMT", "PRMT+IADD3", "IDP4A+LOP3", "FFMA+IMAD", "IDP4A+F2I", "F2I.S32", "I2IP", "F2I.S32+I2IP", "FADD2+IDP4A", "FADD2+PRMT", "FADD2+FADD", "F2I.S8+IDP4A+PRMT"};
static const int group[NMODES] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 2, 2, 2, 2, 3};

template <int MODE> __global__ void k(uint64_t *out, long long *cyc, int iters, float seed) {
    __shared__ uint32_t sm[1024];
    uint64_t a[8]; float f[8]; uint32_t u[8]; int32_t v[8];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
    for (int i = 0; i < 8; ++i) { float2 t = make_float2(seed + i, seed - i); a[i] = *reinterpret_cast<uint64_t *>(&t); f[i] = seed * i; u[i] = (uint32_t)(seed * 1000) + i * 0x01020304u; v[i] = i; }
    float2 m2 = make_float2(1.0000001f, 0.9999999f), c2 = make_float2(1e-9f, -1e-9f);
    uint64_t M = *reinterpret_cast<uint64_t *>(&m2), C = *reinterpret_cast<uint64_t *>(&c2);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#define IS(m) (MODE == (m))
                if (IS(FFMA2) || IS(FFMA2_FFMA) || IS(FFMA2_LOP3)) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(M), "l"(C));
                if (IS(FFMA) || IS(FFMA2_FFMA) || IS(PRMT_FFMA) || IS(DP4A_FFMA) || IS(F2I_FFMA) || IS(FFMA_IMAD)) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(m2.x), "f"(c2.x));
                if (IS(FADD2) || IS(FADD2_DP4A) || IS(FADD2_PRMT) || IS(FADD2_FADD)) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(C));
                if (IS(FADD) || IS(FADD2_FADD)) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2.x));
                if (IS(PRMT) || IS(FADD2_PRMT) || IS(F2I_DP4A_PRMT) || IS(PRMT_FFMA) || IS(DP4A_PRMT) || IS(I2F_PRMT) || IS(PRMT_IADD)) asm volatile("prmt.b32 %0, %0, %1, 0x7650;" : "+r"(u[i]) : "r"(0x4B000000u + i));
                if (IS(LOP3) || IS(FFMA2_LOP3) || IS(DP4A_LOP3)) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(0x80000000u), "r"(0x3EFFFFFFu + it));
                if (IS(IMAD) || IS(DP4A_IMAD) || IS(FFMA_IMAD)) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(it | 3), "r"(i));
                if (IS(IADD) || IS(PRMT_IADD)) asm volatile("add.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(it));
                if (IS(DP4A) || IS(FADD2_DP4A) || IS(F2I_DP4A_PRMT) || IS(DP4A_PRMT) || IS(DP4A_FFMA) || IS(DP4A_IMAD) || IS(DP4A_LOP3) || IS(DP4A_F2I)) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(u[(i + 1) & 7]), "r"(0x01020304 + it));
                if (IS(I2F) || IS(I2F_F2I) || IS(I2F_PRMT)) { float t; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(v[i])); f[i] = t; if (IS(I2F)) v[i] = __float_as_int(t) >> 1; }
                if (IS(F2I) || IS(I2F_F2I) || IS(F2I_FFMA) || IS(DP4A_F2I)) { int t; asm volatile("cvt.rzi.sat.s8.f32 %0, %1;" : "=r"(t) : "f"(f[i])); v[i] = IS(DP4A_F2I) ? v[i] : t; if (IS(F2I)) f[i] = __int_as_float(t + 0x3f800000); if (IS(DP4A_F2I)) u[i] = t; }
                if (IS(F2I_DP4A_PRMT)) { int t; asm volatile("cvt.rzi.sat.s8.f32 %0, %1;" : "=r"(t) : "f"(f[i])); f[i] = __int_as_float(t + 0x3f800000); }
                if (IS(F2I32) || IS(F2I32_I2IP)) { int t; asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(t) : "f"(f[i])); if (IS(F2I32)) f[i] = __int_as_float(t + 0x3f800000); else v[i] = t; }
                if (IS(I2IP) || IS(F2I32_I2IP)) asm volatile("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(v[i]), "r"(v[(i + 1) & 7]));
                if (IS(FMNMX)) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2.x + it));
                if (IS(SHF)) asm volatile("shf.r.wrap.b32 %0, %0, %1, 3;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
                if (IS(ISETP_SEL)) { asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(u[i]) : "r"(0x7fffffffu - it), "r"(i + it)); }
                if (IS(LDS)) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u[i]) : "r"(sbase + ((u[i] & 7) << 7)));
            }
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + (uint64_t)__float_as_uint(f[i]) + u[i] + v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(int warps_per_smsp) {
    const int threads = 128 * warps_per_smsp, blocks = 148, iters = 1000;
    uint64_t *o; long long *c;
    cudaMalloc(&o, sizeof(uint64_t) * blocks * threads); cudaMalloc(&c, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads>>>(o, c, 10, 1.5f);
    k<MODE><<<blocks, threads>>>(o, c, iters, 1.5f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    printf("%-12s w/SMSP=%d  %6.2f clk per group of %d warp-instr  %s\n", names[MODE], warps_per_smsp, avg / ((double)iters * 32 * warps_per_smsp), group[MODE], e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(o); cudaFree(c);
}
template <int M> void all(int w) { run<M>(w); if constexpr (M + 1 < NMODES) all<M + 1>(w); }
int main() {
    for (int w : {2, 6}) { all<0>(w); printf("\n"); }
    return 0;
}
