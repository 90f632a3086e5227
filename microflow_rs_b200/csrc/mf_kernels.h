// mf_kernels.h -- launch interface of the hand-written CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mf {

// n / d for n < 2^31 with one mul.hi + shift (d fixed per launch; verified exhaustively on the host)
struct FastDiv {
    uint32_t d = 1, mul = 0, shr = 0;
    FastDiv() = default;
    explicit FastDiv(uint32_t div) : d(div ? div : 1) {
        if (d == 1) return;
        uint32_t l = 0;
        while ((1ull << l) < d) ++l;
        const uint32_t p = 31 + l;
        mul = (uint32_t)(((1ull << p) + d - 1) / d);
        shr = p - 32;
    }
#ifdef __CUDACC__
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (__umulhi(n, mul) >> shr); }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t &q, uint32_t &r) const { q = div(n); r = n - q * d; }
#endif
};

// Programmatic dependent launch (PDL).  A kernel launched with pdl != 0 may become resident while the previous kernel of the
// stream is still running: its prologue (barrier init, TMEM allocation, weights and constants) overlaps the predecessor's
// tail, and it executes pdl_wait() before its first access to activation memory -- that returns only when the predecessor
// grid has completed and its writes are visible, so both the RAW on this layer's input and the WAR on the ping-pong output
// buffer are ordered.  Every kernel that can be a predecessor calls pdl_trigger() at its start so the dependent may be
// scheduled as SM resources free up.  Without the launch attribute both instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*fn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int pdl, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, fn, KArgs(args)...);
}
#endif

// One quantized conv / depthwise-conv layer on `batch` independent NHWC samples.
struct ConvArgs {
    const uint8_t *in = nullptr;
    uint8_t *out = nullptr;
    const uint8_t *w = nullptr;     // conv: OHWI [Cout][KH][KW][Cin]; depthwise: [KH][KW][Cout]
    const int32_t *w_zp = nullptr;  // [Cout] (per-tensor values are replicated)
    const float *c0z = nullptr;     // [Cout]  f32(out_zp) + c0[ch]
    const float *c1 = nullptr;      // [Cout]
    const int32_t *kcorr = nullptr; // [Cout]  in_zp * sum over ALL taps (and input channels) of w -- fast kernels only
    int H = 1, W = 1, Cin = 1, OH = 1, OW = 1, Cout = 1, KH = 1, KW = 1, sh = 1, sw = 1;
    int off_r = 0, off_c = 0;       // SAME: (K-1)/2 (src/tensor.rs:193), VALID: 0
    int in_zp = 0;
    float lo = -128.f, hi = 127.f;
    int is_u8 = 0, depthwise = 0;
    int wzp_nonzero = 0;            // some weight zero-point != 0 (the general formula with the view-sum term)
    int big_acc = 0;                // 1 if |acc - kcorr| can exceed 2^22 (selects the general exact int->float)
    long long batch = 0;
    int pdl = 0;                    // launch with programmatic stream serialization (see launch_pdl); smem / tcgen05 kernels only
};

struct FcArgs {
    const uint8_t *in = nullptr;   // [batch][K]
    uint8_t *out = nullptr;        // [batch][N]
    const uint8_t *w = nullptr;    // [N][K]
    const float *c0z = nullptr;    // [N]
    const int32_t *c2 = nullptr;   // [N]
    float c1 = 0.f;
    int32_t c3 = 0, w_zp = 0;
    int K = 1, N = 1;
    float lo = -128.f, hi = 127.f;
    int is_u8 = 0;
    long long batch = 0;
    // fc_warp_kernel only: a trailing softmax over the N outputs of each sample run by the same warp (softmax_kernel's arithmetic).
    // sm_out == nullptr: plain fully_connected.  `out` then receives the FC result (the softmax input = "logits") if non-null.
    uint8_t *sm_out = nullptr;
    const float *exp_lut = nullptr;
    int sm_rows = 1, sm_cols = 1;
    float sm_out_scale = 1.f, sm_out_zp = 0.f, sm_lo = -128.f, sm_hi = 127.f;
    int pdl = 0;
    float *out_f32 = nullptr;      // with sm_out: also the model's final dequantize of the softmax output
    float dq_scale = 1.f, dq_zp = 0.f;
};

struct PoolArgs {
    const uint8_t *in = nullptr;
    uint8_t *out = nullptr;
    int H = 1, W = 1, C = 1, OH = 1, OW = 1, KH = 1, KW = 1, sh = 1, sw = 1, off_r = 0, off_c = 0;
    float c0 = 0.f, c1 = 0.f, lo = -128.f, hi = 127.f;
    int is_u8 = 0;
    long long batch = 0;
};

struct SoftmaxArgs {
    const uint8_t *in = nullptr;
    uint8_t *out = nullptr;
    const float *exp_lut = nullptr;  // [256] expf(f32(q) * in_scale), indexed by the raw byte
    int rows = 1, cols = 1;
    float out_scale = 1.f, out_zp = 0.f, lo = -128.f, hi = 127.f;
    long long batch = 0;
};

// Classifier tail in one launch: global average pool -> 1x1 conv with <= 8 outputs -> (reshape) -> softmax.
struct TailArgs {
    const uint8_t *in = nullptr;     // [batch][HW][C] int8
    uint8_t *out = nullptr;          // [batch][N] softmax output
    uint8_t *logits = nullptr;       // optional [batch][N]: the conv output (= the softmax input)
    int HW = 1, C = 128, N = 2;      // C % 128 == 0, N <= 8
    float inv_len = 1.f, pool_c0 = 0.f, pool_c1 = 0.f, pool_lo = -128.f, pool_hi = 127.f;   // average_pool_2d.rs:52-56
    const uint8_t *w = nullptr;      // conv weights [N][C]
    const float *c0z = nullptr, *c1 = nullptr;
    const int32_t *kcorr = nullptr;
    float conv_lo = -128.f, conv_hi = 127.f;
    const float *exp_lut = nullptr;
    int sm_rows = 1, sm_cols = 1;
    float out_scale = 1.f, out_zp = 0.f, sm_lo = -128.f, sm_hi = 127.f;
    long long batch = 0;
    int pdl = 0;
    // optional: the model's final dequantize (src/tensor.rs:89-92) of the softmax output in the same launch
    float *out_f32 = nullptr;
    float dq_scale = 1.f, dq_zp = 0.f;
};
cudaError_t launch_tail_fused(const TailArgs &a, cudaStream_t s);

// ---- generic direct kernels: any shape / zero point / dtype; the cross-check path --------------------
cudaError_t launch_conv_generic(const ConvArgs &a, cudaStream_t s);
cudaError_t launch_fc_generic(const FcArgs &a, cudaStream_t s);
cudaError_t launch_pool_generic(const PoolArgs &a, cudaStream_t s);
cudaError_t launch_softmax(const SoftmaxArgs &a, cudaStream_t s);
cudaError_t launch_quantize(const float *in, uint8_t *out, size_t n, float scale, float zp, int is_u8, cudaStream_t s);
cudaError_t launch_dequantize(const uint8_t *in, float *out, size_t n, float scale, float zp, int is_u8, cudaStream_t s, int pdl = 0);

// MF_LAYOUT_NALGEBRA <-> NHWC: per sample, elements of `elem` bytes, src [C][R][elem] -> dst [R][C][elem] (pass R and C
// swapped for the opposite direction).  Buffer4D / Buffer2D of the reference are column-major (src/buffer.rs:5-16).
cudaError_t launch_layout_transpose(const uint8_t *src, uint8_t *dst, long long batch, int R, int C, int elem, cudaStream_t s);

// ---- SIMT fast kernels (int8, weight zero-point 0): coalesced NHWC, dp4a -------------------------------
bool dwconv_c4_eligible(const ConvArgs &a);      // depthwise, Cin == Cout, C % 4 == 0, int8, weight zero-points 0
bool dwconv_c4_general_eligible(const ConvArgs &a);   // the same kernel also takes uint8 and non-zero weight zero-points
cudaError_t launch_dwconv_c4(const ConvArgs &a, cudaStream_t s);
bool dwconv3x3_rows_eligible(const ConvArgs &a); // + 3x3, stride 1x1 or 2x2: sliding 3x3 window down a column strip
cudaError_t launch_dwconv3x3_rows(const ConvArgs &a, cudaStream_t s);
bool dwconv3x3_smem_eligible(const ConvArgs &a); // + whole sample staged in shared memory by cp.async.bulk (large batches)
cudaError_t launch_dwconv3x3_smem(const ConvArgs &a, int num_sms, cudaStream_t s);
bool dwconv3x3_uses_pair(const ConvArgs &a);     // stride 1: launch_dwconv3x3_smem runs dwconv3x3_pair_kernel (two output columns per thread)
bool dwconv_cin1_eligible(const ConvArgs &a);    // depthwise with Cin == 1 (depth multiplier), Cout % 4 == 0, Cout <= 16
cudaError_t launch_dwconv_cin1(const ConvArgs &a, cudaStream_t s);
bool dwconv_cin1_smem_eligible(const ConvArgs &a);   // 3x3, Cout == 8, whole input image staged by cp.async.bulk (large batches)
cudaError_t launch_dwconv_cin1_smem(const ConvArgs &a, int num_sms, cudaStream_t s);
bool dwconv_cin1_taps_eligible(const ConvArgs &a);   // Cin == 1 -> 8 channels, kernels of 10..128 taps (speech layer 1): dp4a along the kernel rows
cudaError_t launch_dwconv_cin1_taps(const ConvArgs &a, int num_sms, cudaStream_t s);
bool pwconv_dp4a_eligible(const ConvArgs &a);    // 1x1 conv (any stride), Cin % 4 == 0, any Cout
cudaError_t launch_pwconv_dp4a(const ConvArgs &a, cudaStream_t s);
bool fc_warp_eligible(const FcArgs &a);          // K % 16 == 0, N <= 8
cudaError_t launch_fc_warp(const FcArgs &a, cudaStream_t s);

}  // namespace mf
