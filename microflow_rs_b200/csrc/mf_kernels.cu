// mf_kernels.cu -- SIMT kernels (CUDA cores) of the MicroFlow B200 backend.
//
//  * generic direct kernels: one thread per output element, every term of the reference formula evaluated as
//    written (dot, view-sum * w_zp, masked filter-sum * in_zp, len * Cin * in_zp * w_zp).  Any shape, stride,
//    padding, zero point, int8 or uint8.  They are the cross-check path (MF_FLAG_FORCE_GENERIC) and the
//    fallback for shapes no fast kernel takes.
//  * fast kernels: int8 with weight zero-point 0 (every shipped model), NHWC-coalesced 32-bit / 128-bit
//    accesses, dp4a.  Out-of-bounds taps read the input zero-point instead of 0, which turns the reference's
//    per-pixel "masked filter-sum" correction into a per-channel constant (kcorr = in_zp * sum_all w):
//       sum_valid (v - iz) * w  ==  sum_all v' * w - iz * sum_all w ,  v' = v inside, iz outside   (integer-exact)
//
// Reference semantics: src/ops/conv_2d.rs:50-107, depthwise_conv_2d.rs:50-104, fully_connected.rs:42-81,
// average_pool_2d.rs:46-65, softmax.rs:20-26, src/tensor.rs:180-228 (view), src/quantize.rs:16-29.
#include <algorithm>
#include <cstdlib>

#include "mf_device.cuh"
#include "mf_kernels.h"

namespace mf {

static inline unsigned grid_for(long long total, int block) { return (unsigned)((total + block - 1) / block); }

// ================================================================================================
// generic conv / depthwise conv
// ================================================================================================
template <bool U8>
__global__ void __launch_bounds__(256) conv_generic_kernel(ConvArgs a, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int co = (int)(idx % a.Cout);
    long long p = idx / a.Cout;
    const int j = (int)(p % a.OW); p /= a.OW;
    const int i = (int)(p % a.OH);
    const long long b = p / a.OH;
    const uint8_t *in = a.in + (size_t)b * a.H * a.W * a.Cin;
    const int fz = a.w_zp[co];
    int dot = 0, vsum = 0, fsum = 0, len = 0;
    for (int m = 0; m < a.KH; ++m) {
        const int r = a.sh * i + m - a.off_r;
        for (int n = 0; n < a.KW; ++n) {
            const int c = a.sw * j + n - a.off_c;
            if (r < 0 || r >= a.H || c < 0 || c >= a.W) continue;       // tensor.rs:196-218: value 0, mask false, len -= 1
            ++len;
            const uint8_t *px = in + ((size_t)r * a.W + c) * a.Cin;
            if (a.depthwise) {
                const int ci = co < a.Cin ? co : 0;                       // depthwise_conv_2d.rs:67,:72
                const int v = ld_elem<U8>(px + ci);
                const int f = ld_elem<U8>(a.w + ((size_t)m * a.KW + n) * a.Cout + co);
                dot += v * f; vsum += v; fsum += f;
            } else {
                const uint8_t *fw = a.w + (((size_t)co * a.KH + m) * a.KW + n) * a.Cin;
                for (int ch = 0; ch < a.Cin; ++ch) {
                    const int v = ld_elem<U8>(px + ch), f = ld_elem<U8>(fw + ch);
                    dot += v * f; vsum += v; fsum += f;
                }
            }
        }
    }
    const int cin_f = a.depthwise ? 1 : a.Cin;                            // conv_2d.rs:90 vs depthwise_conv_2d.rs:87
    const int acc = dot - vsum * fz - a.in_zp * fsum + len * cin_f * a.in_zp * fz;
    a.out[idx] = (uint8_t)requant(acc, a.c0z[co], a.c1[co], a.lo, a.hi);
}

cudaError_t launch_conv_generic(const ConvArgs &a, cudaStream_t s) {
    const long long total = a.batch * a.OH * a.OW * a.Cout;
    if (total <= 0) return cudaSuccess;
    if (a.is_u8) conv_generic_kernel<true><<<grid_for(total, 256), 256, 0, s>>>(a, total);
    else conv_generic_kernel<false><<<grid_for(total, 256), 256, 0, s>>>(a, total);
    return cudaGetLastError();
}

// ================================================================================================
// generic fully connected
// ================================================================================================
template <bool U8>
__global__ void __launch_bounds__(256) fc_generic_kernel(FcArgs a, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int j = (int)(idx % a.N);
    const long long b = idx / a.N;
    const uint8_t *x = a.in + (size_t)b * a.K, *w = a.w + (size_t)j * a.K;
    int dot = 0, rowsum = 0;
    for (int k = 0; k < a.K; ++k) {
        const int v = ld_elem<U8>(x + k);
        dot += v * ld_elem<U8>(w + k);
        rowsum += v;
    }
    const int acc = dot - rowsum * a.w_zp - a.c2[j] + a.c3;               // fully_connected.rs:71
    a.out[idx] = (uint8_t)requant(acc, a.c0z[j], a.c1, a.lo, a.hi);
}

cudaError_t launch_fc_generic(const FcArgs &a, cudaStream_t s) {
    const long long total = a.batch * a.N;
    if (total <= 0) return cudaSuccess;
    if (a.is_u8) fc_generic_kernel<true><<<grid_for(total, 256), 256, 0, s>>>(a, total);
    else fc_generic_kernel<false><<<grid_for(total, 256), 256, 0, s>>>(a, total);
    return cudaGetLastError();
}

// ================================================================================================
// average pool (average_pool_2d.rs:52-56): x = (1/f32(len)) * f32(sum);  y = roundf(c0 * x + c1)
// ================================================================================================
template <bool U8>
__global__ void __launch_bounds__(256) pool_generic_kernel(PoolArgs a, long long total) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ch = (int)(idx % a.C);
    long long p = idx / a.C;
    const int j = (int)(p % a.OW); p /= a.OW;
    const int i = (int)(p % a.OH);
    const long long b = p / a.OH;
    const uint8_t *in = a.in + (size_t)b * a.H * a.W * a.C;
    int sum = 0, len = 0;
    for (int m = 0; m < a.KH; ++m) {
        const int r = a.sh * i + m - a.off_r;
        for (int n = 0; n < a.KW; ++n) {
            const int c = a.sw * j + n - a.off_c;
            if (r < 0 || r >= a.H || c < 0 || c >= a.W) continue;
            ++len;
            sum += ld_elem<U8>(in + ((size_t)r * a.W + c) * a.C + ch);
        }
    }
    const float x = __fmul_rn(__fdiv_rn(1.0f, __int2float_rn(len)), __int2float_rn(sum));
    const float t = __fadd_rn(__fmul_rn(a.c0, x), a.c1);
    a.out[idx] = (uint8_t)round_clamp(t, a.lo, a.hi);
}

cudaError_t launch_pool_generic(const PoolArgs &a, cudaStream_t s) {
    const long long total = a.batch * a.OH * a.OW * a.C;
    if (total <= 0) return cudaSuccess;
    if (a.is_u8) pool_generic_kernel<true><<<grid_for(total, 256), 256, 0, s>>>(a, total);
    else pool_generic_kernel<false><<<grid_for(total, 256), 256, 0, s>>>(a, total);
    return cudaGetLastError();
}

// ================================================================================================
// softmax tail (softmax.rs:20-26; activation.rs:44-46): one thread per sample.
// exp values come from the host-built 256-entry table (exact port of libm expf); the sum runs in nalgebra's
// column-major order starting from 0.0; y = quantize(e / sum).
// ================================================================================================
__global__ void __launch_bounds__(128) softmax_kernel(SoftmaxArgs a) {
    __shared__ float lut[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) lut[t] = a.exp_lut[t];
    __syncthreads();
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    const int n = a.rows * a.cols;
    const uint8_t *x = a.in + (size_t)b * n;
    uint8_t *y = a.out + (size_t)b * n;
    float sum = 0.0f;
    for (int j = 0; j < a.cols; ++j)
        for (int i = 0; i < a.rows; ++i) sum = __fadd_rn(sum, lut[x[(size_t)i * a.cols + j]]);
    for (int k = 0; k < n; ++k) {
        const float q = __fadd_rn(__fdiv_rn(__fdiv_rn(lut[x[k]], sum), a.out_scale), a.out_zp);   // quantize.rs:17
        y[k] = (uint8_t)round_clamp(q, a.lo, a.hi);
    }
}

cudaError_t launch_softmax(const SoftmaxArgs &a, cudaStream_t s) {
    if (a.batch <= 0) return cudaSuccess;
    softmax_kernel<<<grid_for(a.batch, 128), 128, 0, s>>>(a);
    return cudaGetLastError();
}

// ================================================================================================
// quantize / dequantize (quantize.rs:16-29)
// ================================================================================================
__global__ void __launch_bounds__(256) quantize_kernel(const float *in, uint8_t *out, size_t n, float scale, float zp, float lo, float hi) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float x = in[idx];
    float t = __fadd_rn(__fdiv_rn(x, scale), zp);
    if (t != t) t = 0.0f;                                                 // Rust `NaN as i8` == 0
    out[idx] = (uint8_t)round_clamp(t, lo, hi);
}
template <bool U8>
__global__ void __launch_bounds__(256) dequantize_kernel(const uint8_t *in, float *out, size_t n, float scale, float zp) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (idx >= n) return;
    out[idx] = __fmul_rn(scale, __fsub_rn(__int2float_rn(ld_elem<U8>(in + idx)), zp));
}
cudaError_t launch_quantize(const float *in, uint8_t *out, size_t n, float scale, float zp, int is_u8, cudaStream_t s) {
    if (!n) return cudaSuccess;
    quantize_kernel<<<grid_for((long long)n, 256), 256, 0, s>>>(in, out, n, scale, zp, is_u8 ? 0.f : -128.f, is_u8 ? 255.f : 127.f);
    return cudaGetLastError();
}
cudaError_t launch_dequantize(const uint8_t *in, float *out, size_t n, float scale, float zp, int is_u8, cudaStream_t s, int pdl) {
    if (!n) return cudaSuccess;
    const dim3 grid(grid_for((long long)n, 256));
    if (is_u8) return launch_pdl(dequantize_kernel<true>, grid, dim3(256), 0, s, pdl, in, out, n, scale, zp);
    return launch_pdl(dequantize_kernel<false>, grid, dim3(256), 0, s, pdl, in, out, n, scale, zp);
}

// ================================================================================================
// host-buffer layout conversion (MF_LAYOUT_NALGEBRA): dst[b][r][c][e] = src[b][c][r][e].  One thread per destination byte
// (writes coalesced; an input image is a few KB, so the strided reads stay in L1/L2).
// ================================================================================================
__global__ void __launch_bounds__(256) layout_transpose_kernel(const uint8_t *src, uint8_t *dst, uint32_t per_sample, FastDiv fd_ce, FastDiv fd_e, int R, int C,
                                                               int elem, long long batch) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_sample) return;
    uint32_t r, rem, c, e;
    fd_ce.divmod(idx, r, rem);
    fd_e.divmod(rem, c, e);
    const uint32_t sidx = (c * (uint32_t)R + r) * (uint32_t)elem + e;
    for (long long b = blockIdx.y; b < batch; b += gridDim.y) dst[(size_t)b * per_sample + idx] = src[(size_t)b * per_sample + sidx];
}
cudaError_t launch_layout_transpose(const uint8_t *src, uint8_t *dst, long long batch, int R, int C, int elem, cudaStream_t s) {
    const long long per = (long long)R * C * elem;
    if (per <= 0 || batch <= 0) return cudaSuccess;
    if (per >= (1ll << 31)) return cudaErrorInvalidValue;
    const long long gx = (per + 255) / 256;
    long long gy = (148ll * 16 + gx - 1) / gx;
    if (gy > batch) gy = batch;
    if (gy > 65535) gy = 65535;
    layout_transpose_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(src, dst, (uint32_t)per, FastDiv((uint32_t)(C * elem)), FastDiv((uint32_t)elem), R, C, elem,
                                                                            batch);
    return cudaGetLastError();
}

// ================================================================================================
// FAST kernels.  Grid: blockIdx.y = sample (grid-stride), blockIdx.x * blockDim.x + threadIdx.x = 32-bit index inside
// the sample, decoded with FastDiv (no 64-bit div/mod on the device).
// ================================================================================================
// gridDim.y is capped so that the launch is ~16 CTAs per SM: every CTA then loops over many samples and its prologue
// (weights to registers / shared memory, index decode, epilogue constants) is paid once, not once per sample
static inline dim3 grid2(long long per_sample, int block, long long batch) {
    const long long gx = (per_sample + block - 1) / block;
    long long gy = (148ll * 16 + gx - 1) / gx;
    if (gy > batch) gy = batch;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

// ------------------------------------------------------------------------------------------------
// depthwise conv, Cin == Cout = C, C % 4 == 0, int8, w_zp == 0, any kernel / stride / padding.
// One thread = 4 consecutive channels (one 32-bit word) of one output pixel; a warp covers 128 contiguous
// output bytes and reads 128 contiguous input bytes per tap (stride 1).  Bytes are sign-extended with one PRMT each
// and multiplied with IMAD (IDP.4A issues to the slow XU pipe on sm_100 and is avoided in every hot loop).
// ------------------------------------------------------------------------------------------------
// U8: uint8 tensors (zero-extended bytes).  WZP: non-zero weight zero-points -- with zero-point padding every term of the reference
// formula (depthwise_conv_2d.rs:66-87) is uniform over the image: sum_valid (v - iz)(w - wz) = sum_all v'w - wz * sum_all v' - iz * sum_all w
// + taps * iz * wz  (v' = v inside, iz outside), so the only extra work is the window sum S = sum_all v' per channel.
template <int KH_T, int KW_T, bool U8, bool WZP>
__global__ void __launch_bounds__(256) dwconv_c4_kernel(ConvArgs a, uint32_t words_per_sample, FastDiv fd_g, FastDiv fd_ow) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= words_per_sample) return;
    const int KH = KH_T ? KH_T : a.KH, KW = KW_T ? KW_T : a.KW;
    const int G = a.Cout >> 2;
    uint32_t p, g, i, j;
    fd_g.divmod(idx, p, g);
    fd_ow.divmod(p, i, j);
    const uint32_t *ww = reinterpret_cast<const uint32_t *>(a.w);
    const uint32_t izw = (uint32_t)(a.in_zp & 0xff) * 0x01010101u;
    int4 kc = __ldg(reinterpret_cast<const int4 *>(a.kcorr) + g);
    int4 wz = make_int4(0, 0, 0, 0);
    if (WZP) {                                                    // fold the constant term taps * iz * wz into the correction
        wz = __ldg(reinterpret_cast<const int4 *>(a.w_zp) + g);
        const int t = KH * KW * a.in_zp;
        kc.x -= t * wz.x; kc.y -= t * wz.y; kc.z -= t * wz.z; kc.w -= t * wz.w;
    }
    const float4 z = __ldg(reinterpret_cast<const float4 *>(a.c0z) + g);
    const float4 s = __ldg(reinterpret_cast<const float4 *>(a.c1) + g);
    for (long long b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const uint32_t *inw = reinterpret_cast<const uint32_t *>(a.in) + (size_t)b * a.H * a.W * G;
        int acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
        for (int m = 0; m < KH; ++m) {
            const int r = a.sh * (int)i + m - a.off_r;
            const bool rok = (unsigned)r < (unsigned)a.H;
#pragma unroll
            for (int n = 0; n < KW; ++n) {
                const int c = a.sw * (int)j + n - a.off_c;
                const bool ok = rok && (unsigned)c < (unsigned)a.W;
                const uint32_t v = ok ? __ldg(inw + ((size_t)r * a.W + c) * G + g) : izw;
                const uint32_t wv = __ldg(ww + (size_t)(m * KW + n) * G + g);
                const int v0 = ext8<U8, 0>(v), v1 = ext8<U8, 1>(v), v2 = ext8<U8, 2>(v), v3 = ext8<U8, 3>(v);
                acc0 += v0 * ext8<U8, 0>(wv);
                acc1 += v1 * ext8<U8, 1>(wv);
                acc2 += v2 * ext8<U8, 2>(wv);
                acc3 += v3 * ext8<U8, 3>(wv);
                if (WZP) { s0 += v0; s1 += v1; s2 += v2; s3 += v3; }
            }
        }
        if (WZP) { acc0 -= wz.x * s0; acc1 -= wz.y * s1; acc2 -= wz.z * s2; acc3 -= wz.w * s3; }
        reinterpret_cast<uint32_t *>(a.out)[(size_t)b * words_per_sample + idx] =
            requant4_clamp<false>(acc0 - kc.x, acc1 - kc.y, acc2 - kc.z, acc3 - kc.w, z, s, a.lo, a.hi, U8);
    }
}

// any depthwise shape with Cin == Cout, C % 4 == 0 and accumulators within 2^22: int8 or uint8, any weight zero-points
bool dwconv_c4_general_eligible(const ConvArgs &a) {
    return a.depthwise && a.Cin == a.Cout && (a.Cout % 4) == 0 && a.kcorr != nullptr && a.w_zp != nullptr && !a.big_acc;
}
// the int8 / weight-zero-point-0 subset every other depthwise fast kernel builds on
bool dwconv_c4_eligible(const ConvArgs &a) { return dwconv_c4_general_eligible(a) && !a.is_u8 && !a.wzp_nonzero; }
cudaError_t launch_dwconv_c4(const ConvArgs &a, cudaStream_t s) {
    const long long per = (long long)a.OH * a.OW * (a.Cout / 4);
    if (per <= 0 || a.batch <= 0) return cudaSuccess;
    const FastDiv fg((uint32_t)(a.Cout / 4)), fow((uint32_t)a.OW);
    const dim3 grid = grid2(per, 256, a.batch);
    const bool k33 = a.KH == 3 && a.KW == 3;
#define MF_C4(U, Z)                                                                                          \
    do {                                                                                                     \
        if (k33) dwconv_c4_kernel<3, 3, U, Z><<<grid, 256, 0, s>>>(a, (uint32_t)per, fg, fow);               \
        else dwconv_c4_kernel<0, 0, U, Z><<<grid, 256, 0, s>>>(a, (uint32_t)per, fg, fow);                   \
    } while (0)
    if (a.is_u8) { if (a.wzp_nonzero) MF_C4(true, true); else MF_C4(true, false); }
    else { if (a.wzp_nonzero) MF_C4(false, true); else MF_C4(false, false); }
#undef MF_C4
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 (every person_detect depthwise layer but the first), stride 1x1 or 2x2.
// One thread = one 4-channel word of one output column, walking DOWN a strip of output rows with the 3x3 input
// window held UNPACKED in 36 registers (three rotating row arrays): each new output row loads 3 (stride 1) or 6
// (stride 2) words instead of 9 and unpacks every byte once (PRMT sign-extend); the 36 sign-extended weights, the
// epilogue constants and all index math are hoisted out of the row loop; MACs are IMAD, the epilogue is XU-free.
// Lanes run along (column, channel-word), i.e. along contiguous NHWC memory: every load and store is coalesced.
// ------------------------------------------------------------------------------------------------
template <int S, int XU>   // XU = how many of the 4 channels of each output word take the XU (I2F + F2I.S8) epilogue
__global__ void __launch_bounds__(128, 4) dwconv3x3_rows_kernel(ConvArgs a, uint32_t threads_per_sample, uint32_t rows_per_strip, FastDiv fd_xw, FastDiv fd_g) {
    // Input rows travel global -> shared through a private cp.async ring per thread (DEPTH rows x 3 words): the loads of
    // rows r+2 .. r+DEPTH+1 are in flight while row r is being multiplied, at zero register cost, so the DRAM latency
    // (~1 us under load) is covered even at 16 resident warps per SM.  Every thread only reads what it copied itself:
    // cp.async.wait_group is the only synchronisation.
    constexpr int DEPTH = (S == 1) ? 4 : 6;
    __shared__ uint32_t ring[DEPTH][3][128];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= threads_per_sample) return;
    const int G = a.Cout >> 2;
    uint32_t strip, x, j, g;
    fd_xw.divmod(t, strip, x);
    fd_g.divmod(x, j, g);
    const uint32_t *ww = reinterpret_cast<const uint32_t *>(a.w);
    int wi[9][4];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const uint32_t wv = __ldg(ww + (size_t)k * G + g);
        wi[k][0] = sx8<0>(wv); wi[k][1] = sx8<1>(wv); wi[k][2] = sx8<2>(wv); wi[k][3] = sx8<3>(wv);
    }
    const int4 kc = __ldg(reinterpret_cast<const int4 *>(a.kcorr) + g);
    const float4 z = __ldg(reinterpret_cast<const float4 *>(a.c0z) + g);
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(a.c1) + g);
    const uint32_t izw = (uint32_t)(a.in_zp & 0xff) * 0x01010101u;
    const int c0 = S * (int)j - a.off_c;
    const bool cok0 = (unsigned)c0 < (unsigned)a.W, cok1 = (unsigned)(c0 + 1) < (unsigned)a.W, cok2 = (unsigned)(c0 + 2) < (unsigned)a.W;
    const int i0 = (int)(strip * rows_per_strip);
    const int i1 = min(a.OH, i0 + (int)rows_per_strip);
    const int row_words = a.W * G, out_row_words = a.OW * G;
    const float lo = a.lo, hi = a.hi;
    const int H = a.H;
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(&ring[0][0][threadIdx.x]);

    for (long long b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const int r_first = S * i0 - a.off_r;
        const int total = (S == 1) ? (i1 - i0) + 2 : 2 * (i1 - i0) + 1;       // input rows this strip consumes
        const uint32_t *p = reinterpret_cast<const uint32_t *>(a.in) + (size_t)b * H * row_words + (ptrdiff_t)r_first * row_words + (ptrdiff_t)c0 * G + g;
        uint32_t *o = reinterpret_cast<uint32_t *>(a.out) + ((size_t)b * a.OH + i0) * out_row_words + (size_t)j * G + g;
        int issued = 0, taken = 0;
        auto issue = [&]() {      // request input row r_first + issued (one commit group per row, empty past the end)
            if (issued < total && (unsigned)(r_first + issued) < (unsigned)H) {
                const uint32_t dst = ring0 + (uint32_t)(issued % DEPTH) * (3 * 128 * 4);
                if (cok0) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(p) : "memory");
                if (cok1) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 512), "l"(p + G) : "memory");
                if (cok2) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 1024), "l"(p + 2 * G) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            p += row_words;
            ++issued;
        };
        auto take = [&](int (&d)[12]) {   // d[n * 4 + k] = channel k of window column n of input row r_first + taken
            asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
            const bool rok = (unsigned)(r_first + taken) < (unsigned)H;
            const int slot = taken % DEPTH;
            const uint32_t v0 = (rok && cok0) ? ring[slot][0][threadIdx.x] : izw;
            const uint32_t v1 = (rok && cok1) ? ring[slot][1][threadIdx.x] : izw;
            const uint32_t v2 = (rok && cok2) ? ring[slot][2][threadIdx.x] : izw;
            d[0] = sx8<0>(v0); d[1] = sx8<1>(v0); d[2] = sx8<2>(v0); d[3] = sx8<3>(v0);
            d[4] = sx8<0>(v1); d[5] = sx8<1>(v1); d[6] = sx8<2>(v1); d[7] = sx8<3>(v1);
            d[8] = sx8<0>(v2); d[9] = sx8<1>(v2); d[10] = sx8<2>(v2); d[11] = sx8<3>(v2);
            ++taken;
        };
        auto emit = [&](const int (&r0)[12], const int (&r1)[12], const int (&r2)[12]) {
            int acc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int s = r0[k] * wi[0][k];
                s += r0[4 + k] * wi[1][k]; s += r0[8 + k] * wi[2][k];
                s += r1[k] * wi[3][k]; s += r1[4 + k] * wi[4][k]; s += r1[8 + k] * wi[5][k];
                s += r2[k] * wi[6][k]; s += r2[4 + k] * wi[7][k]; s += r2[8 + k] * wi[8][k];
                acc[k] = s;
            }
            if (XU >= 4) *o = requant4_i2f(acc[0] - kc.x, acc[1] - kc.y, acc[2] - kc.z, acc[3] - kc.w, z, sc);      // full int8 clamp: I2F + packed F2IP
            else
            *o = pack4(XU > 0 ? requant_xu<true>(acc[0] - kc.x, z.x, sc.x, lo, hi) : requant_nx<false>(acc[0] - kc.x, z.x, sc.x, lo, hi),
                       XU > 1 ? requant_xu<true>(acc[1] - kc.y, z.y, sc.y, lo, hi) : requant_nx<false>(acc[1] - kc.y, z.y, sc.y, lo, hi),
                       XU > 2 ? requant_xu<true>(acc[2] - kc.z, z.z, sc.z, lo, hi) : requant_nx<false>(acc[2] - kc.z, z.z, sc.z, lo, hi),
                       XU > 3 ? requant_xu<true>(acc[3] - kc.w, z.w, sc.w, lo, hi) : requant_nx<false>(acc[3] - kc.w, z.w, sc.w, lo, hi));
            o += out_row_words;
        };
        int ra[12], rb[12], rc[12];
        int left = i1 - i0;                                         // output rows still to produce
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) issue();
        // every take() is followed by one issue(): exactly DEPTH groups stay in flight, so wait_group<DEPTH-1> == "my row landed"
        if (S == 1) {
            take(ra); issue();
            take(rb); issue();
            while (true) {
                take(rc); issue(); emit(ra, rb, rc); if (--left == 0) break;
                take(ra); issue(); emit(rb, rc, ra); if (--left == 0) break;
                take(rb); issue(); emit(rc, ra, rb); if (--left == 0) break;
            }
        } else {
            take(ra); issue();
            while (true) {
                take(rb); issue(); take(rc); issue(); emit(ra, rb, rc); if (--left == 0) break;
                take(rb); issue(); take(ra); issue(); emit(rc, rb, ra); if (--left == 0) break;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // drain before the ring is reused for the next sample
    }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3, sample-resident variant (the fast path at large batch).
// A sample's whole input tensor is contiguous in HBM (1-36 KB for every layer of person_detect), so ONE cp.async.bulk
// (TMA 1-D bulk copy, mbarrier-tracked) brings it into shared memory; a ring of NB sample buffers per CTA keeps the next
// samples in flight while the current one is computed.  Compute threads then issue no global loads and no address
// arithmetic: one thread = one 4-channel word of one output column (x = j * G + g < OW * G), it walks down its row strip
// with the 3x3 window unpacked in registers and reads 3 words per new input row from shared memory (neighbouring columns
// are neighbouring words; image borders are the zero-point).  Stores are coalesced 128-byte rows.
// ------------------------------------------------------------------------------------------------
constexpr int kDwSmemThreads = 192;

__device__ __forceinline__ void sm_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (done) break;
        if (spins > (1u << 22)) __trap();
    }
}

// Sample-resident depthwise 3x3.  Issue slots are what bound it (tools/ubench/pipes.cu: IMAD, IDP4A, PRMT, LOP3 and the
// packed FFMA2 all cost 2 clk per warp on their pipe, FFMA 1), so the MACs are IDP.4A: the three columns of an input row
// are transposed (6 PRMT, on the ALU pipe, overlapping the dot products) into one register per channel holding that
// channel's (left, centre, right, -) bytes, and one dp4a against (w[T][0], w[T][1], w[T][2], 0) is a whole kernel row.
// Input-stationary: a transposed row is used at once for every output row it feeds (kernel row 0 of one, 1 of the one
// before, 2 of the one before that), so no window of rows is kept and the output rows in flight are independent chains.
// Borders cost nothing in the row loop: each ring slot is [one row of in_zp][the sample][one row of in_zp], so rows -1 and
// H are ordinary loads, a window column outside the image reads a zero-point word through a pointer whose per-row stride
// is 0, and the accumulators start at -in_zp * sum(w).
// Slot layout (kDwHead bytes of header, then nbuf slots of buf_stride bytes, then kDwTail bytes):
//   header: nbuf "sample landed" mbarriers at +0, nbuf "warps finished with this slot" counters at +64
//   slot:   [one row of in_zp][the sample, filled by one cp.async.bulk][one row of in_zp]
// A thread addresses its window with ONE pointer: the three columns are at p, p + 4G and p + 8G bytes (G = channel words per
// pixel, a uniform offset) and a new input row is one pointer add.  A window column outside the image therefore reads
// whatever lies one pixel left / right in memory (the neighbouring row, the header or the tail padding): its three weight
// bytes are zeroed in that thread's registers, so the garbage is multiplied by 0, and the thread's correction term only sums
// the weights that remain -- exactly the reference's masked filter sum (conv_2d.rs:83-89 / depthwise_conv_2d.rs:80-86).
constexpr uint32_t kDwHead = 512, kDwTail = 768;     // head >= one pixel (C <= 256 bytes) + the 128-byte barrier block; tail >= two pixels (pair kernel) + slack

template <int S, bool FULL, int MINB>
__global__ void __launch_bounds__(kDwSmemThreads, MINB) dwconv3x3_smem_kernel(ConvArgs a, uint32_t in_bytes, uint32_t buf_stride, int nbuf, int xw, int nstrip,
                                                                            int rows_per_strip) {
    extern __shared__ __align__(128) uint8_t dsm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(dsm);
    uint8_t *bufs = dsm + kDwHead;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    const uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(bufs);
    const int tid = threadIdx.x;
    const long long first = blockIdx.x, step = gridDim.x;
    const int G = a.Cout >> 2;
    const uint32_t row_bytes = (uint32_t)(a.W * G) * 4u;
    if (tid == 0) {
        for (int k = 0; k < nbuf; ++k) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * k), "r"(1u) : "memory");
            reinterpret_cast<uint32_t *>(dsm + 64)[k] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // the zero-point rows above and below every slot's sample (the bulk copies never touch them)
        const uint32_t izw = (uint32_t)(a.in_zp & 0xff) * 0x01010101u;
        const int rw = (int)(row_bytes >> 2);
        for (int k = 0; k < nbuf; ++k) {
            uint32_t *top = reinterpret_cast<uint32_t *>(bufs + (size_t)k * buf_stride);
            uint32_t *bot = reinterpret_cast<uint32_t *>(bufs + (size_t)k * buf_stride + row_bytes + in_bytes);
            for (int i = tid; i < rw; i += kDwSmemThreads) { top[i] = izw; bot[i] = izw; }
        }
    }
    __syncthreads();
    auto request = [&](long long b, int slot) {                         // one thread
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * slot), "r"(in_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf0 + (uint32_t)slot * buf_stride + row_bytes),
                     "l"(a.in + (size_t)b * in_bytes), "r"(in_bytes), "r"(bar0 + 8u * slot)
                     : "memory");
    };
    pdl_trigger();
    if (tid == 0) {                                                     // the input is the previous kernel's output: first access after pdl_wait
        pdl_wait();
        for (int k = 0; k < nbuf; ++k)
            if (first + (long long)k * step < a.batch) request(first + (long long)k * step, k);
    }

    const bool active = tid < xw * nstrip;
    const int strip = active ? tid / xw : 0;
    const int x = active ? tid - strip * xw : 0;
    const int j = x / G, g = x - j * G;
    const int c0 = S * j - a.off_c;
    // wq[T][c] = (w[T][0][c], w[T][1][c], w[T][2][c], 0): the three taps of kernel row T of channel 4g + c, with the taps of
    // window columns outside the image zeroed
    uint32_t wq[3][4];
    int fresh[4];                                                       // accumulator init: float-conversion bias - in_zp * (sum of the remaining weights)
    {
        const uint32_t *ww = reinterpret_cast<const uint32_t *>(a.w);
        uint32_t keep = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if ((unsigned)(c0 + k) < (unsigned)a.W) keep |= 0xffu << (8 * k);
        int wsum[4] = {0, 0, 0, 0};
#pragma unroll
        for (int T = 0; T < 3; ++T) {
            transpose_3x4(__ldg(ww + (size_t)(3 * T) * G + g), __ldg(ww + (size_t)(3 * T + 1) * G + g), __ldg(ww + (size_t)(3 * T + 2) * G + g), wq[T]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                wq[T][c] &= keep;
                wsum[c] = __dp4a((int)wq[T][c], 0x01010101, wsum[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) fresh[c] = kAccBias - a.in_zp * wsum[c];
    }
    const float4 z = __ldg(reinterpret_cast<const float4 *>(a.c0z) + g);
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(a.c1) + g);
    const int i0 = strip * rows_per_strip;
    const int i1 = active ? min(a.OH, i0 + rows_per_strip) : i0;
    const int out_row_words = a.OW * G;
    const float lo = a.lo, hi = a.hi;
    // word offset (inside a slot) of window column 0 at the first input row of this strip; may be negative by up to G words
    const int poff = (int)(row_bytes >> 2) * (1 + S * i0 - a.off_r) + c0 * G + g;
    const int row_words = (int)(row_bytes >> 2);

    pdl_wait();                                                         // every thread: its stores must not overtake the previous kernel's reads
    // per-sample state is carried incrementally (ring slot, mbarrier phase, slot base, output pointer): the small feature
    // maps spend only a few hundred instructions per sample, so divisions and 64-bit multiplies per sample would show
    int slot = 0;
    uint32_t phase = 0;
    const uint32_t *pslot = reinterpret_cast<const uint32_t *>(bufs) + poff;
    uint32_t *osample = reinterpret_cast<uint32_t *>(a.out) + ((size_t)first * a.OH + i0) * out_row_words + x;
    const size_t ostep = (size_t)step * a.OH * out_row_words;
    const uint32_t slot_words = buf_stride >> 2;
    for (long long b = first; b < a.batch; b += step) {
        sm_mbar_wait(bar0 + 8u * slot, phase);
        if (i1 > i0) {
            const uint32_t *p = pslot;
            uint32_t *o = osample;
            auto take = [&](uint32_t (&t)[4]) {                         // the next input row, transposed
                const uint32_t v0 = p[0], v1 = p[G], v2 = p[2 * G];
                p += row_words;
                transpose_3x4(v0, v1, v2, t);
            };
            struct Acc { int c[4]; };
            const Acc init = {{fresh[0], fresh[1], fresh[2], fresh[3]}};
            auto mac = [&](Acc &A, const uint32_t (&t)[4], int T) {
#pragma unroll
                for (int c = 0; c < 4; ++c) A.c[c] = __dp4a((int)t[c], (int)wq[T][c], A.c[c]);
            };
            auto store = [&](const Acc &A) {
                *o = requant4_biased<FULL>(A.c[0], A.c[1], A.c[2], A.c[3], z, sc, lo, hi);
                o += out_row_words;
            };
            uint32_t t[4];
            Acc A = init, B = init, C = init;
            int left = i1 - i0;
            if (S == 1) {
                take(t); mac(A, t, 0);
                take(t); mac(A, t, 1); mac(B, t, 0);
                while (true) {
                    take(t); mac(A, t, 2); mac(B, t, 1); C = init; mac(C, t, 0); store(A); if (--left == 0) break;
                    take(t); mac(B, t, 2); mac(C, t, 1); A = init; mac(A, t, 0); store(B); if (--left == 0) break;
                    take(t); mac(C, t, 2); mac(A, t, 1); B = init; mac(B, t, 0); store(C); if (--left == 0) break;
                }
            } else {                                                    // stride 2: every second input row closes one output row and opens the next
                take(t); mac(A, t, 0);
                while (true) {
                    take(t); mac(A, t, 1);
                    take(t); mac(A, t, 2); B = init; mac(B, t, 0); store(A); if (--left == 0) break;
                    take(t); mac(B, t, 1);
                    take(t); mac(B, t, 2); A = init; mac(A, t, 0); store(B); if (--left == 0) break;
                }
            }
        }
        // No CTA-wide barrier per sample: warps run ahead into the slots that have already landed.  The last warp to finish
        // with a slot (acq_rel counter: every warp's reads happen-before the refill) re-arms it for sample b + nbuf * step.
        __syncwarp();
        if ((tid & 31) == 0) {
            uint32_t old;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(bar0 + 64u + 4u * slot) : "memory");
            if (old == (uint32_t)(kDwSmemThreads / 32 - 1)) {           // last warp out: reset the counter, refill the slot
                reinterpret_cast<volatile uint32_t *>(dsm + 64)[slot] = 0;
                if (b + (long long)nbuf * step < a.batch) request(b + (long long)nbuf * step, slot);
            }
        }
        osample += ostep;
        pslot += slot_words;
        if (++slot == nbuf) { slot = 0; phase ^= 1u; pslot -= (size_t)nbuf * slot_words; }
    }
}

// ------------------------------------------------------------------------------------------------
// Stride-1 variant with TWO output columns per thread.  The window of the column pair (2jj, 2jj+1) is four input columns:
// a 4x4 byte transpose (8 PRMT) leaves, per channel, one register holding (col0, col1, col2, col3), and the two outputs are
// dp4a against (w0, w1, w2, 0) and (0, w0, w1, w2).  Per 8 outputs: 4 LDS + 8 PRMT instead of 6 + 12, half the address / loop /
// barrier instructions, and two independent accumulator chains per thread for the scheduler to interleave.  Same slot layout,
// border handling (zeroed weight bytes + per-thread correction) and epilogue as dwconv3x3_smem_kernel; CTAs are 96 or 128
// threads (one strip of a 192-word-wide row is 96 pairs), 5 per SM.
// ------------------------------------------------------------------------------------------------
constexpr int kDwPairMaxThreads = 128;
template <bool FULL, int MINB>
__global__ void __launch_bounds__(kDwPairMaxThreads, MINB) dwconv3x3_pair_kernel(ConvArgs a, uint32_t in_bytes, uint32_t buf_stride, int nbuf, int xwp, int nstrip,
                                                                                int rows_per_strip) {
    extern __shared__ __align__(128) uint8_t dsm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(dsm);
    uint8_t *bufs = dsm + kDwHead;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    const uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(bufs);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const uint32_t nwarps = (uint32_t)nthreads >> 5;
    const long long first = blockIdx.x, step = gridDim.x;
    const int G = a.Cout >> 2;
    const uint32_t row_bytes = (uint32_t)(a.W * G) * 4u;
    if (tid == 0) {
        for (int k = 0; k < nbuf; ++k) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * k), "r"(1u) : "memory");
            reinterpret_cast<uint32_t *>(dsm + 64)[k] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const uint32_t izw = (uint32_t)(a.in_zp & 0xff) * 0x01010101u;
        const int rw = (int)(row_bytes >> 2);
        for (int k = 0; k < nbuf; ++k) {
            uint32_t *top = reinterpret_cast<uint32_t *>(bufs + (size_t)k * buf_stride);
            uint32_t *bot = reinterpret_cast<uint32_t *>(bufs + (size_t)k * buf_stride + row_bytes + in_bytes);
            for (int i = tid; i < rw; i += nthreads) { top[i] = izw; bot[i] = izw; }
        }
    }
    __syncthreads();
    auto request = [&](long long b, int slot) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * slot), "r"(in_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf0 + (uint32_t)slot * buf_stride + row_bytes),
                     "l"(a.in + (size_t)b * in_bytes), "r"(in_bytes), "r"(bar0 + 8u * slot)
                     : "memory");
    };
    pdl_trigger();
    if (tid == 0) {
        pdl_wait();
        for (int k = 0; k < nbuf; ++k)
            if (first + (long long)k * step < a.batch) request(first + (long long)k * step, k);
    }

    const bool active = tid < xwp * nstrip;
    const int strip = active ? tid / xwp : 0;
    const int x = active ? tid - strip * xwp : 0;
    const int jj = x / G, g = x - jj * G;
    const int j0 = 2 * jj;
    const bool second = j0 + 1 < a.OW;                                  // an odd output width leaves the last pair half empty
    const int c0 = j0 - a.off_c;                                        // leftmost of the four window columns
    // wa[T][c] = (w[T][0], w[T][1], w[T][2], 0) for column j0, wb[T][c] = (0, w[T][0], w[T][1], w[T][2]) for column j0 + 1, each with
    // the taps that fall on window columns outside the image zeroed
    uint32_t wa[3][4], wb[3][4];
    int fresh_a[4], fresh_b[4];
    {
        const uint32_t *ww = reinterpret_cast<const uint32_t *>(a.w);
        uint32_t keep = 0;                                              // byte k <-> window column c0 + k
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((unsigned)(c0 + k) < (unsigned)a.W) keep |= 0xffu << (8 * k);
        int sa[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
#pragma unroll
        for (int T = 0; T < 3; ++T) {
            uint32_t wq[4];
            transpose_3x4(__ldg(ww + (size_t)(3 * T) * G + g), __ldg(ww + (size_t)(3 * T + 1) * G + g), __ldg(ww + (size_t)(3 * T + 2) * G + g), wq);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t w3 = wq[c] & 0x00ffffffu;
                wa[T][c] = w3 & keep;
                wb[T][c] = (w3 << 8) & keep;
                sa[c] = __dp4a((int)wa[T][c], 0x01010101, sa[c]);
                sb[c] = __dp4a((int)wb[T][c], 0x01010101, sb[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { fresh_a[c] = kAccBias - a.in_zp * sa[c]; fresh_b[c] = kAccBias - a.in_zp * sb[c]; }
    }
    const float4 z = __ldg(reinterpret_cast<const float4 *>(a.c0z) + g);
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(a.c1) + g);
    const int i0 = strip * rows_per_strip;
    const int i1 = active ? min(a.OH, i0 + rows_per_strip) : i0;
    const int out_row_words = a.OW * G;
    const float lo = a.lo, hi = a.hi;
    const int poff = (int)(row_bytes >> 2) * (1 + i0 - a.off_r) + c0 * G + g;
    const int row_words = (int)(row_bytes >> 2);

    pdl_wait();
    int slot = 0;
    uint32_t phase = 0;
    const uint32_t *pslot = reinterpret_cast<const uint32_t *>(bufs) + poff;
    uint32_t *osample = reinterpret_cast<uint32_t *>(a.out) + ((size_t)first * a.OH + i0) * out_row_words + (size_t)j0 * G + g;
    const size_t ostep = (size_t)step * a.OH * out_row_words;
    const uint32_t slot_words = buf_stride >> 2;
    for (long long b = first; b < a.batch; b += step) {
        sm_mbar_wait(bar0 + 8u * slot, phase);
        if (i1 > i0) {
            const uint32_t *p = pslot;
            uint32_t *o = osample;
            auto take = [&](uint32_t (&t)[4]) {
                const uint32_t v0 = p[0], v1 = p[G], v2 = p[2 * G], v3 = p[3 * G];
                p += row_words;
                transpose_4x4(v0, v1, v2, v3, t);
            };
            struct Acc { int a[4], b[4]; };
            Acc init;
#pragma unroll
            for (int c = 0; c < 4; ++c) { init.a[c] = fresh_a[c]; init.b[c] = fresh_b[c]; }
            auto mac = [&](Acc &A, const uint32_t (&t)[4], int T) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    A.a[c] = __dp4a((int)t[c], (int)wa[T][c], A.a[c]);
                    A.b[c] = __dp4a((int)t[c], (int)wb[T][c], A.b[c]);
                }
            };
            auto store = [&](const Acc &A) {
                o[0] = requant4_biased<FULL>(A.a[0], A.a[1], A.a[2], A.a[3], z, sc, lo, hi);
                const uint32_t y1 = requant4_biased<FULL>(A.b[0], A.b[1], A.b[2], A.b[3], z, sc, lo, hi);
                if (second) o[G] = y1;
                o += out_row_words;
            };
            uint32_t t[4];
            Acc A = init, B = init, C = init;
            int left = i1 - i0;
            take(t); mac(A, t, 0);
            take(t); mac(A, t, 1); mac(B, t, 0);
            while (true) {
                take(t); mac(A, t, 2); mac(B, t, 1); C = init; mac(C, t, 0); store(A); if (--left == 0) break;
                take(t); mac(B, t, 2); mac(C, t, 1); A = init; mac(A, t, 0); store(B); if (--left == 0) break;
                take(t); mac(C, t, 2); mac(A, t, 1); B = init; mac(B, t, 0); store(C); if (--left == 0) break;
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) {
            uint32_t old;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(bar0 + 64u + 4u * slot) : "memory");
            if (old == nwarps - 1u) {
                reinterpret_cast<volatile uint32_t *>(dsm + 64)[slot] = 0;
                if (b + (long long)nbuf * step < a.batch) request(b + (long long)nbuf * step, slot);
            }
        }
        osample += ostep;
        pslot += slot_words;
        if (++slot == nbuf) { slot = 0; phase ^= 1u; pslot -= (size_t)nbuf * slot_words; }
    }
}

// shapes the sample-resident kernel takes: 3x3, stride 1x1 / 2x2, one block spans the full output width, whole input fits a ring buffer
bool dwconv3x3_smem_eligible(const ConvArgs &a) {
    if (!(dwconv_c4_eligible(a) && a.KH == 3 && a.KW == 3 && a.sh == a.sw && (a.sh == 1 || a.sh == 2))) return false;
    const long long in_bytes = (long long)a.H * a.W * a.Cin;
    const int xw = a.OW * (a.Cout / 4);
    // rows read: -off_r .. sh*(OH-1) - off_r + 2; the slot holds rows -1 .. H
    const bool rows_ok = a.off_r <= 1 && a.off_r >= 0 && a.sh * (a.OH - 1) - a.off_r + 2 <= a.H;
    return in_bytes % 16 == 0 && ((long long)a.W * a.Cin) % 16 == 0 && in_bytes <= 48 * 1024 && xw <= kDwSmemThreads && rows_ok && a.batch >= 148 * 2 && a.Cout <= 256 &&
           a.off_c >= 0 && a.off_c <= 1 &&   // a window column outside the image is at most one pixel (<= kDwTail bytes) away
          
           ((uintptr_t)a.in % 16) == 0;
}
bool dwconv3x3_uses_pair(const ConvArgs &a) {
    static const int env_pair = [] { const char *e = std::getenv("MF_DW_PAIR"); return e ? std::atoi(e) : 1; }();
    // an odd output width leaves the last pair half empty: measured slower on the 3-wide map (22.8 vs 20.8 us), so pairs need an
    // even width or one wide enough for the idle half-pair not to matter (MF_DW_PAIR=2 forces pairs everywhere, for tests)
    const bool width_ok = (a.OW % 2) == 0 || a.OW >= 9 || env_pair == 2;
    return env_pair && a.sh == 1 && width_ok && ((a.OW + 1) / 2) * (a.Cout / 4) <= kDwPairMaxThreads;
}
cudaError_t launch_dwconv3x3_smem(const ConvArgs &a, int num_sms, cudaStream_t s) {
    const uint32_t in_bytes = (uint32_t)(a.H * a.W * a.Cin);
    const uint32_t buf_stride = (in_bytes + 2u * (uint32_t)(a.W * a.Cin) + 127u) & ~127u;   // sample + a zero-point row either side
    const int xw = a.OW * (a.Cout / 4);
    const int xwp = ((a.OW + 1) / 2) * (a.Cout / 4);
    if (dwconv3x3_uses_pair(a)) {   // stride 1: two output columns per thread
        // CTA size and CTAs/SM measured on person_detect's layers (profiles/r01k_dw_pair_sweep.txt): 96/128 threads x 4 per SM
        // beats x 3, x 5, x 6 and 192-thread CTAs
        int nstrip = kDwPairMaxThreads / xwp;
        if (nstrip > a.OH) nstrip = a.OH;
        const int rows = (a.OH + nstrip - 1) / nstrip;
        nstrip = (a.OH + rows - 1) / rows;
        const int threads = (xwp * nstrip + 31) & ~31;
        static const int env_pminb = [] { const char *e = std::getenv("MF_DW_PAIR_MINB"); return e ? std::atoi(e) : 4; }();
        int per_sm = env_pminb < 2 ? 2 : (env_pminb > 6 ? 6 : env_pminb), nbuf = 0;
        const int minb = per_sm;
        for (; per_sm >= 1; --per_sm) {
            const long long share = (227ll * 1024) / per_sm - 1024 - (long long)(kDwHead + kDwTail);
            nbuf = (int)(share / buf_stride);
            if (nbuf > 4) nbuf = 4;
            if (nbuf >= 2 || per_sm == 1) break;
        }
        if (nbuf < 1) return cudaErrorInvalidConfiguration;
        const size_t smem = kDwHead + (size_t)nbuf * buf_stride + kDwTail;
        const bool full = a.lo == -128.f && a.hi == 127.f;
        using Fn = void (*)(ConvArgs, uint32_t, uint32_t, int, int, int, int);
#define MF_DWP_PICK(M) (full ? dwconv3x3_pair_kernel<true, M> : dwconv3x3_pair_kernel<false, M>)
        Fn fn = minb >= 6 ? MF_DWP_PICK(6) : (minb == 5 ? MF_DWP_PICK(5) : (minb == 4 ? MF_DWP_PICK(4) : (minb == 3 ? MF_DWP_PICK(3) : MF_DWP_PICK(2))));
#undef MF_DWP_PICK
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (e != cudaSuccess) return e;
        long long ctas = (long long)num_sms * per_sm;
        if (ctas > a.batch) ctas = a.batch;
        return launch_pdl(fn, dim3((unsigned)ctas), dim3((unsigned)threads), smem, s, a.pdl, a, in_bytes, buf_stride, nbuf, xwp, nstrip, rows);
    }
    int nstrip = kDwSmemThreads / xw;
    if (nstrip > a.OH) nstrip = a.OH;
    const int rows = (a.OH + nstrip - 1) / nstrip;
    nstrip = (a.OH + rows - 1) / rows;
    static const int env_minb = [] { const char *e = std::getenv("MF_DW_MINB"); return e ? std::atoi(e) : 4; }();
    const int minb = env_minb < 2 ? 2 : (env_minb > 6 ? 6 : env_minb);
    // persistent CTAs, samples taken grid-stride: as many CTAs per SM as the launch bound allows while every CTA still has
    // a ring of >= 2 sample slots (<= 4) in its share of the 227 KB
    int per_sm = minb, nbuf = 0;
    for (; per_sm >= 1; --per_sm) {
        const long long share = (227ll * 1024) / per_sm - 1024 - (long long)(kDwHead + kDwTail);
        nbuf = (int)(share / buf_stride);
        if (nbuf > 4) nbuf = 4;
        if (nbuf >= 2 || per_sm == 1) break;
    }
    if (nbuf < 1) return cudaErrorInvalidConfiguration;
    const size_t smem = kDwHead + (size_t)nbuf * buf_stride + kDwTail;
    const bool full = a.lo == -128.f && a.hi == 127.f;          // F2I.S8 saturation doubles as the clamp
    using Fn = void (*)(ConvArgs, uint32_t, uint32_t, int, int, int, int);
    Fn fn = nullptr;
#define MF_DW_PICK(M) (a.sh == 1 ? (full ? dwconv3x3_smem_kernel<1, true, M> : dwconv3x3_smem_kernel<1, false, M>) \
                                 : (full ? dwconv3x3_smem_kernel<2, true, M> : dwconv3x3_smem_kernel<2, false, M>))
    fn = minb == 6 ? MF_DW_PICK(6) : (minb == 5 ? MF_DW_PICK(5) : (minb == 4 ? MF_DW_PICK(4) : (minb == 3 ? MF_DW_PICK(3) : MF_DW_PICK(2))));
#undef MF_DW_PICK
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) return e;
    long long ctas = (long long)num_sms * per_sm;
    if (ctas > a.batch) ctas = a.batch;
    return launch_pdl(fn, dim3((unsigned)ctas), dim3(kDwSmemThreads), smem, s, a.pdl, a, in_bytes, buf_stride, nbuf, xw, nstrip, rows);
}

bool dwconv3x3_rows_eligible(const ConvArgs &a) {
    return dwconv_c4_eligible(a) && a.KH == 3 && a.KW == 3 && a.sh == a.sw && (a.sh == 1 || a.sh == 2);
}
cudaError_t launch_dwconv3x3_rows(const ConvArgs &a, cudaStream_t s) {
    if (a.batch <= 0) return cudaSuccess;
    const int G = a.Cout / 4;
    const uint32_t xw = (uint32_t)(a.OW * G);
    // strips of output rows: each strip pays an exposed load latency and ~80 instructions of setup once, so they should be
    // as long as the machine still fills: whole columns when that leaves >= 8 CTAs of 128 threads per SM
    uint32_t rows = (uint32_t)a.OH;
    while (rows > 4 && (long long)xw * a.batch * ((a.OH + rows - 1) / rows) < 148ll * 8 * 128) rows = (rows + 1) / 2;
    const uint32_t strips = (uint32_t)((a.OH + rows - 1) / rows);
    const long long per = (long long)strips * xw;
    const FastDiv fxw(xw), fg((uint32_t)G);
    const dim3 grid = grid2(per, 128, a.batch);
    static const int env_xu = [] { const char *e = std::getenv("MF_DW_XU"); return e ? std::atoi(e) : -1; }();
    int xu = env_xu >= 0 ? env_xu : 4;
    if (!(a.lo == -128.f && a.hi == 127.f)) xu = 0;             // the XU epilogue relies on F2I.S8 saturation = full int8 clamp
    xu = xu >= 4 ? 4 : (xu >= 2 ? 2 : 0);
    if (a.sh == 1) {
        if (xu == 4) dwconv3x3_rows_kernel<1, 4><<<grid, 128, 0, s>>>(a, (uint32_t)per, rows, fxw, fg);
        else if (xu == 2) dwconv3x3_rows_kernel<1, 2><<<grid, 128, 0, s>>>(a, (uint32_t)per, rows, fxw, fg);
        else dwconv3x3_rows_kernel<1, 0><<<grid, 128, 0, s>>>(a, (uint32_t)per, rows, fxw, fg);
    } else {
        if (xu == 4) dwconv3x3_rows_kernel<2, 4><<<grid, 128, 0, s>>>(a, (uint32_t)per, rows, fxw, fg);
        else if (xu == 2) dwconv3x3_rows_kernel<2, 2><<<grid, 128, 0, s>>>(a, (uint32_t)per, rows, fxw, fg);
        else dwconv3x3_rows_kernel<2, 0><<<grid, 128, 0, s>>>(a, (uint32_t)per, rows, fxw, fg);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// depthwise conv with a single input channel and a depth multiplier (person_detect layer 0: 3x3 s2 -> 8 ch,
// speech layer 1: 10x8 s2 -> 8 ch).  Output channel c reads input channel 0 (depthwise_conv_2d.rs:67).
// One thread = one output pixel, all COUT channels.  The sign-extended weights sit in shared memory as int32
// [tap][COUT] (broadcast LDS.128); the activation byte arrives sign-extended from LDG.S8; MACs are IMAD.
// ------------------------------------------------------------------------------------------------
template <int COUT, int KH_T, int KW_T>
__global__ void __launch_bounds__(128) dwconv_cin1_kernel(ConvArgs a, uint32_t px_per_sample, FastDiv fd_ow) {
    extern __shared__ int4 w_s4[];                       // [taps][COUT / 4]
    constexpr int Q = COUT / 4;
    const int KH = KH_T ? KH_T : a.KH, KW = KW_T ? KW_T : a.KW;
    for (int e = threadIdx.x; e < KH * KW * COUT; e += blockDim.x) reinterpret_cast<int *>(w_s4)[e] = (int)(int8_t)a.w[e];
    __syncthreads();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= px_per_sample) return;
    uint32_t i, j;
    fd_ow.divmod(idx, i, j);
    const int r0 = a.sh * (int)i - a.off_r, c0 = a.sw * (int)j - a.off_c;
    const float lo = a.lo, hi = a.hi;
    for (long long b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const int8_t *in = reinterpret_cast<const int8_t *>(a.in) + (size_t)b * a.H * a.W;
        int acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = 0;
#pragma unroll
        for (int m = 0; m < KH; ++m) {
            const int r = r0 + m;
            const bool rok = (unsigned)r < (unsigned)a.H;
#pragma unroll
            for (int n = 0; n < KW; ++n) {
                const int c = c0 + n;
                const bool ok = rok && (unsigned)c < (unsigned)a.W;
                const int v = ok ? (int)__ldg(in + (size_t)r * a.W + c) : a.in_zp;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const int4 wv = w_s4[(m * KW + n) * Q + q];
                    acc[4 * q + 0] += v * wv.x; acc[4 * q + 1] += v * wv.y; acc[4 * q + 2] += v * wv.z; acc[4 * q + 3] += v * wv.w;
                }
            }
        }
        uint32_t *out = reinterpret_cast<uint32_t *>(a.out) + ((size_t)b * px_per_sample + idx) * Q;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int4 kc = __ldg(reinterpret_cast<const int4 *>(a.kcorr) + q);
            const float4 z = __ldg(reinterpret_cast<const float4 *>(a.c0z) + q);
            const float4 s = __ldg(reinterpret_cast<const float4 *>(a.c1) + q);
            out[q] = pack4(requant_nx<false>(acc[4 * q + 0] - kc.x, z.x, s.x, lo, hi), requant_nx<false>(acc[4 * q + 1] - kc.y, z.y, s.y, lo, hi),
                           requant_nx<false>(acc[4 * q + 2] - kc.z, z.z, s.z, lo, hi), requant_nx<false>(acc[4 * q + 3] - kc.w, z.w, s.w, lo, hi));
        }
    }
}

// Depth-multiplier first layer (Cin == 1 -> 8 channels, 3x3), sample-resident like dwconv3x3_smem_kernel and built from the
// same parts: zero-point rows around the sample, per-slot warp counters, IDP.4A.  A thread owns one output column.  The
// three input bytes of a kernel row are consecutive in memory: two aligned LDS words and one PRMT with a per-thread
// selector give (left, centre, right, -), and one dp4a per channel against (w[T][0][c], w[T][1][c], w[T][2][c], 0) is that
// channel's whole kernel row.  A word that lies entirely left / right of the image is read from a zero-point word instead
// (pointer with stride 0), which needs W % 4 == 0.
template <int S, bool FULL>
__global__ void __launch_bounds__(kDwSmemThreads, 3) dwconv_cin1_smem_kernel(ConvArgs a, uint32_t in_bytes, uint32_t buf_stride, int nbuf, int nstrip, int rows_per_strip) {
    constexpr int COUT = 8;
    extern __shared__ __align__(128) uint8_t dsm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(dsm);                 // nbuf "sample landed" mbarriers; counters at dsm + 64
    uint8_t *bufs = dsm + 128;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    const uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(bufs);
    const int tid = threadIdx.x;
    const long long first = blockIdx.x, step = gridDim.x;
    const uint32_t row_bytes = (uint32_t)a.W;
    if (tid == 0) {
        for (int k = 0; k < nbuf; ++k) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * k), "r"(1u) : "memory");
            reinterpret_cast<uint32_t *>(dsm + 64)[k] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const uint32_t izw = (uint32_t)(a.in_zp & 0xff) * 0x01010101u;
        const int rw = (int)(row_bytes >> 2);
        for (int k = 0; k < nbuf; ++k) {
            uint32_t *top = reinterpret_cast<uint32_t *>(bufs + (size_t)k * buf_stride);
            uint32_t *bot = reinterpret_cast<uint32_t *>(bufs + (size_t)k * buf_stride + row_bytes + in_bytes);
            for (int i = tid; i < rw; i += kDwSmemThreads) { top[i] = izw; bot[i] = izw; }
        }
    }
    __syncthreads();
    auto request = [&](long long b, int slot) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * slot), "r"(in_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf0 + (uint32_t)slot * buf_stride + row_bytes),
                     "l"(a.in + (size_t)b * in_bytes), "r"(in_bytes), "r"(bar0 + 8u * slot)
                     : "memory");
    };
    pdl_trigger();
    if (tid == 0) {
        pdl_wait();
        for (int k = 0; k < nbuf; ++k)
            if (first + (long long)k * step < a.batch) request(first + (long long)k * step, k);
    }

    const bool active = tid < a.OW * nstrip;
    const int strip = active ? tid / a.OW : 0;
    const int j = active ? tid - strip * a.OW : 0;
    // wq[T][c] = (w[T][0][c], w[T][1][c], w[T][2][c], 0)
    uint32_t wq[3][COUT];
#pragma unroll
    for (int T = 0; T < 3; ++T)
#pragma unroll
        for (int c = 0; c < COUT; ++c)
            wq[T][c] = (uint32_t)__ldg(a.w + (3 * T) * COUT + c) | ((uint32_t)__ldg(a.w + (3 * T + 1) * COUT + c) << 8) | ((uint32_t)__ldg(a.w + (3 * T + 2) * COUT + c) << 16);
    int kcr[COUT];
    float zr[COUT], sr[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) { kcr[c] = kAccBias - __ldg(a.kcorr + c); zr[c] = __ldg(a.c0z + c); sr[c] = __ldg(a.c1 + c); }   // pre-biased accumulators (mf_device.cuh)
    const int c0 = S * j - a.off_c;                                     // leftmost window column, >= -1
    const int w0 = (c0 >= 0 ? c0 : c0 - 3) / 4;                         // floor(c0 / 4): the aligned word holding it
    const uint32_t kk = (uint32_t)(c0 - 4 * w0);                        // its byte inside that word
    const uint32_t sel = kk | ((kk + 1) << 4) | ((kk + 2) << 8) | ((kk + 2) << 12);
    const bool lo_ok = w0 >= 0, hi_ok = 4 * (w0 + 1) < a.W;
    const int i0 = strip * rows_per_strip;
    const int i1 = active ? min(a.OH, i0 + rows_per_strip) : i0;
    const float lo = a.lo, hi = a.hi;
    const uint32_t first_row_off = (uint32_t)((int)row_bytes * (1 + S * i0 - a.off_r));
    const uint32_t off_lo = lo_ok ? first_row_off + 4u * (uint32_t)w0 : 0u, off_hi = hi_ok ? first_row_off + 4u * (uint32_t)(w0 + 1) : 0u;
    const uint32_t st_lo = lo_ok ? row_bytes : 0u, st_hi = hi_ok ? row_bytes : 0u;

    pdl_wait();
    int slot = 0;                                                       // per-sample state carried incrementally (see dwconv3x3_smem_kernel)
    uint32_t phase = 0, base = buf0;
    uint2 *osample = reinterpret_cast<uint2 *>(a.out) + ((size_t)first * a.OH + i0) * a.OW + j;
    const size_t ostep = (size_t)step * a.OH * a.OW;
    for (long long b = first; b < a.batch; b += step) {
        sm_mbar_wait(bar0 + 8u * slot, phase);
        if (i1 > i0) {
            uint32_t plo = base + off_lo, phi = base + off_hi;
            uint2 *o = osample;
            auto take = [&]() {                                         // (left, centre, right, -) of the next input row
                const uint32_t vlo = lds_u32(plo), vhi = lds_u32(phi);
                plo += st_lo; phi += st_hi;
                uint32_t t;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(vlo), "r"(vhi), "r"(sel));
                return t;
            };
            struct Acc { int c[COUT]; };
            Acc fresh;
#pragma unroll
            for (int c = 0; c < COUT; ++c) fresh.c[c] = kcr[c];
            auto mac = [&](Acc &A, uint32_t t, int T) {
#pragma unroll
                for (int c = 0; c < COUT; ++c) A.c[c] = __dp4a((int)t, (int)wq[T][c], A.c[c]);
            };
            auto store = [&](const Acc &A) {
                *o = make_uint2(requant4_biased<FULL>(A.c[0], A.c[1], A.c[2], A.c[3], make_float4(zr[0], zr[1], zr[2], zr[3]), make_float4(sr[0], sr[1], sr[2], sr[3]), lo, hi),
                                requant4_biased<FULL>(A.c[4], A.c[5], A.c[6], A.c[7], make_float4(zr[4], zr[5], zr[6], zr[7]), make_float4(sr[4], sr[5], sr[6], sr[7]), lo, hi));
                o += a.OW;
            };
            Acc A = fresh, B = fresh, C = fresh;
            int left = i1 - i0;
            uint32_t t;
            if (S == 1) {
                t = take(); mac(A, t, 0);
                t = take(); mac(A, t, 1); mac(B, t, 0);
                while (true) {
                    t = take(); mac(A, t, 2); mac(B, t, 1); C = fresh; mac(C, t, 0); store(A); if (--left == 0) break;
                    t = take(); mac(B, t, 2); mac(C, t, 1); A = fresh; mac(A, t, 0); store(B); if (--left == 0) break;
                    t = take(); mac(C, t, 2); mac(A, t, 1); B = fresh; mac(B, t, 0); store(C); if (--left == 0) break;
                }
            } else {
                t = take(); mac(A, t, 0);
                while (true) {
                    t = take(); mac(A, t, 1);
                    t = take(); mac(A, t, 2); B = fresh; mac(B, t, 0); store(A); if (--left == 0) break;
                    t = take(); mac(B, t, 1);
                    t = take(); mac(B, t, 2); A = fresh; mac(A, t, 0); store(B); if (--left == 0) break;
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) {
            uint32_t old;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(bar0 + 64u + 4u * slot) : "memory");
            if (old == (uint32_t)(kDwSmemThreads / 32 - 1)) {           // last warp out: reset the counter, refill the slot
                reinterpret_cast<volatile uint32_t *>(dsm + 64)[slot] = 0;
                if (b + (long long)nbuf * step < a.batch) request(b + (long long)nbuf * step, slot);
            }
        }
        osample += ostep;
        base += buf_stride;
        if (++slot == nbuf) { slot = 0; phase ^= 1u; base = buf0; }
    }
}

bool dwconv_cin1_smem_eligible(const ConvArgs &a) {
    const long long in_bytes = (long long)a.H * a.W;
    const bool rows_ok = a.off_r >= 0 && a.off_r <= 1 && a.sh * (a.OH - 1) - a.off_r + 2 <= a.H;
    const bool cols_ok = a.off_c >= 0 && a.off_c <= 1 && a.sw * (a.OW - 1) - a.off_c + 2 <= a.W;   // at most one column outside, on either side
    return a.depthwise && !a.is_u8 && a.Cin == 1 && a.Cout == 8 && a.KH == 3 && a.KW == 3 && a.sh == a.sw && (a.sh == 1 || a.sh == 2) && a.kcorr != nullptr &&
           !a.big_acc && a.W % 16 == 0 && in_bytes <= 32 * 1024 && a.OW <= kDwSmemThreads && rows_ok && cols_ok && a.batch >= 148 * 2 && ((uintptr_t)a.in % 16) == 0 &&
           ((uintptr_t)a.out % 8) == 0;
}
cudaError_t launch_dwconv_cin1_smem(const ConvArgs &a, int num_sms, cudaStream_t s) {
    const uint32_t in_bytes = (uint32_t)(a.H * a.W);
    const uint32_t buf_stride = (in_bytes + 2u * (uint32_t)a.W + 127u) & ~127u;
    int nstrip = kDwSmemThreads / a.OW;
    if (nstrip > a.OH) nstrip = a.OH;
    const int rows = (a.OH + nstrip - 1) / nstrip;
    nstrip = (a.OH + rows - 1) / rows;
    const int per_sm = 3;
    int nbuf = (int)(((227ll * 1024) / per_sm - 1024 - 128) / buf_stride);
    if (nbuf > 4) nbuf = 4;
    if (nbuf < 2) return cudaErrorInvalidConfiguration;       // cannot happen: in_bytes <= 32 KB
    const size_t smem = 128 + (size_t)nbuf * buf_stride;
    const bool full = a.lo == -128.f && a.hi == 127.f;
    using Fn = void (*)(ConvArgs, uint32_t, uint32_t, int, int, int);
    Fn fn = a.sh == 1 ? (full ? dwconv_cin1_smem_kernel<1, true> : dwconv_cin1_smem_kernel<1, false>)
                      : (full ? dwconv_cin1_smem_kernel<2, true> : dwconv_cin1_smem_kernel<2, false>);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) return e;
    long long ctas = (long long)num_sms * per_sm;
    if (ctas > a.batch) ctas = a.batch;
    return launch_pdl(fn, dim3((unsigned)ctas), dim3(kDwSmemThreads), smem, s, a.pdl, a, in_bytes, buf_stride, nbuf, nstrip, rows);
}

// ------------------------------------------------------------------------------------------------
// Depth-multiplier first layer with a LARGE kernel (speech layer 1: 10x8 taps, Cin == 1 -> 8 channels, stride 2): 80 MACs per
// output, i.e. compute-bound -- the generic-shape kernel above spends one IMAD per tap and channel.  Here the whole sample sits
// in shared memory inside a frame of zero-point bytes (rows and columns the window can reach outside the image), so the taps of
// a kernel row are consecutive bytes of one padded row: two PRMT assemble them from three aligned words and ONE dp4a does four
// taps of one channel.  A thread owns PX = 2 output pixels x 8 channels, so the weight words (shared memory, broadcast LDS.128)
// are fetched once for both.  The next sample's words are prefetched into registers while the current one is computed
// (two slots, one __syncthreads per sample).
// ------------------------------------------------------------------------------------------------
constexpr int kTapsThreads = 256;
template <int NW, bool FULL>   // NW = ceil(KW / 4) words of taps per kernel row
__global__ void __launch_bounds__(kTapsThreads, 3) dwconv_cin1_taps_kernel(ConvArgs a, int Hp, int Wp, int pad_l, uint32_t slot_bytes) {
    constexpr int COUT = 8, PX = 2, PF = 4;                              // PF: prefetched words per thread (H * W / 4 <= PF * threads)
    extern __shared__ __align__(16) uint8_t tsm[];
    uint32_t *wsm = reinterpret_cast<uint32_t *>(tsm);                   // [KH][NW][COUT] packed tap words
    uint8_t *slots = tsm + (((size_t)a.KH * NW * COUT * 4 + 15) & ~(size_t)15);
    const int tid = threadIdx.x;
    const long long first = blockIdx.x, step = gridDim.x;
    for (int e = tid; e < a.KH * NW * COUT; e += kTapsThreads) {
        const int c = e % COUT, q = (e / COUT) % NW, m = e / (COUT * NW);
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k) {
            const int n = 4 * q + k;
            if (n < a.KW) v |= (uint32_t)a.w[(m * a.KW + n) * COUT + c] << (8 * k);
        }
        wsm[e] = v;
    }
    {   // both slots start as a frame of zero-point bytes; the sample interior is overwritten for every sample
        const uint32_t izw = (uint32_t)(a.in_zp & 0xff) * 0x01010101u;
        uint32_t *sw = reinterpret_cast<uint32_t *>(slots);
        for (uint32_t i = tid; i < 2 * (slot_bytes >> 2); i += kTapsThreads) sw[i] = izw;
    }
    const int npx = a.OH * a.OW;
    const int wrow = a.W >> 2;                                           // words per image row (W % 4 == 0)
    const int nwords = a.H * wrow;
    // where this thread's prefetched words go inside a slot (word k of the sample -> padded row off_r + k / wrow, column pad_l + ...)
    uint32_t dstw[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) {
        const int k = tid + u * kTapsThreads;
        const int r = k / wrow, cw = k - r * wrow;
        dstw[u] = k < nwords ? (uint32_t)(((r + a.off_r) * Wp + pad_l) >> 2) + (uint32_t)cw : 0xffffffffu;
    }
    // this thread's two output pixels: byte offset of the window origin inside a slot, and the PRMT selector of its misalignment
    uint32_t worg[PX], sel[PX];
    bool pvalid[PX];
#pragma unroll
    for (int u = 0; u < PX; ++u) {
        const int px = tid + u * kTapsThreads;
        pvalid[u] = px < npx;
        const int i = pvalid[u] ? px / a.OW : 0, j = pvalid[u] ? px - i * a.OW : 0;
        const int pc0 = a.sw * j - a.off_c + pad_l;                     // >= 0
        worg[u] = (uint32_t)(a.sh * i * Wp + (pc0 & ~3));
        const uint32_t kk = (uint32_t)(pc0 & 3);
        sel[u] = kk | ((kk + 1) << 4) | ((kk + 2) << 8) | ((kk + 3) << 12);
    }
    int init[COUT];
    float zr[COUT], sr[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) { init[c] = kAccBias - __ldg(a.kcorr + c); zr[c] = __ldg(a.c0z + c); sr[c] = __ldg(a.c1 + c); }
    const float lo = a.lo, hi = a.hi;
    pdl_trigger();
    pdl_wait();
    uint32_t pre[PF];
    auto fetch = [&](long long b) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.in + (size_t)b * a.H * a.W);
#pragma unroll
        for (int u = 0; u < PF; ++u)
            if (dstw[u] != 0xffffffffu) pre[u] = __ldg(src + tid + u * kTapsThreads);
    };
    auto stash = [&](int slot) {
        uint32_t *dst = reinterpret_cast<uint32_t *>(slots + (size_t)slot * slot_bytes);
#pragma unroll
        for (int u = 0; u < PF; ++u)
            if (dstw[u] != 0xffffffffu) dst[dstw[u]] = pre[u];
    };
    if (first < a.batch) { fetch(first); }
    __syncthreads();                                                     // frames and weights written
    if (first < a.batch) stash(0);
    __syncthreads();
    int slot = 0;
    for (long long b = first; b < a.batch; b += step, slot ^= 1) {
        const bool more = b + step < a.batch;
        if (more) fetch(b + step);                                       // in flight while this sample is computed
        const uint8_t *img = slots + (size_t)slot * slot_bytes;
        int acc[PX][COUT];
#pragma unroll
        for (int u = 0; u < PX; ++u)
#pragma unroll
            for (int c = 0; c < COUT; ++c) acc[u][c] = init[c];
        const uint32_t *rowp0 = reinterpret_cast<const uint32_t *>(img + worg[0]);
        const uint32_t *rowp1 = reinterpret_cast<const uint32_t *>(img + worg[1]);
        const int wp4 = Wp >> 2;
        const uint4 *wq = reinterpret_cast<const uint4 *>(wsm);
        for (int m = 0; m < a.KH; ++m) {
            uint32_t t[PX][NW];
            {
                uint32_t v[NW + 1];
#pragma unroll
                for (int q = 0; q <= NW; ++q) v[q] = rowp0[q];
#pragma unroll
                for (int q = 0; q < NW; ++q) asm("prmt.b32 %0, %1, %2, %3;" : "=r"(t[0][q]) : "r"(v[q]), "r"(v[q + 1]), "r"(sel[0]));
#pragma unroll
                for (int q = 0; q <= NW; ++q) v[q] = rowp1[q];
#pragma unroll
                for (int q = 0; q < NW; ++q) asm("prmt.b32 %0, %1, %2, %3;" : "=r"(t[1][q]) : "r"(v[q]), "r"(v[q + 1]), "r"(sel[1]));
            }
            rowp0 += wp4; rowp1 += wp4;
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                const uint4 wa = wq[(m * NW + q) * 2], wb = wq[(m * NW + q) * 2 + 1];
                const uint32_t wv[COUT] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int c = 0; c < COUT; ++c) {
                    acc[0][c] = __dp4a((int)t[0][q], (int)wv[c], acc[0][c]);
                    acc[1][c] = __dp4a((int)t[1][q], (int)wv[c], acc[1][c]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PX; ++u)
            if (pvalid[u]) {
                uint2 *o = reinterpret_cast<uint2 *>(a.out) + (size_t)b * npx + tid + u * kTapsThreads;
                *o = make_uint2(requant4_biased<FULL>(acc[u][0], acc[u][1], acc[u][2], acc[u][3], make_float4(zr[0], zr[1], zr[2], zr[3]), make_float4(sr[0], sr[1], sr[2], sr[3]), lo, hi),
                                requant4_biased<FULL>(acc[u][4], acc[u][5], acc[u][6], acc[u][7], make_float4(zr[4], zr[5], zr[6], zr[7]), make_float4(sr[4], sr[5], sr[6], sr[7]), lo, hi));
            }
        if (more) stash(slot ^ 1);                                       // the other slot was last read one iteration ago (barrier below)
        __syncthreads();
    }
}

// Window geometry of the padded slot: rows 0 .. Hp-1 cover input rows -off_r .. , columns start pad_l bytes left of the image
static void taps_geometry(const ConvArgs &a, int &Hp, int &Wp, int &pad_l) {
    pad_l = (a.off_c + 3) & ~3;
    const int need_r = std::max(a.H + a.off_r, (a.OH - 1) * a.sh + a.KH);
    const int need_c = std::max(a.W + pad_l, (a.OW - 1) * a.sw - a.off_c + pad_l + ((a.KW + 3) & ~3) + 4);   // + one word: the PRMT reads NW + 1 words
    Hp = need_r;
    Wp = (need_c + 3) & ~3;
}
bool dwconv_cin1_taps_eligible(const ConvArgs &a) {
    if (!(a.depthwise && !a.is_u8 && a.Cin == 1 && a.Cout == 8 && a.kcorr != nullptr && !a.big_acc)) return false;
    if (a.KW < 1 || a.KW > 8 || a.KH < 1 || a.KH * a.KW > 128 || a.KH * a.KW <= 9) return false;            // 3x3 has its own kernel
    if (a.W % 4 != 0 || a.off_r < 0 || a.off_c < 0 || a.batch < 148 * 2) return false;
    if ((long long)a.H * a.W / 4 > 4ll * kTapsThreads || a.OH * a.OW > 2 * kTapsThreads) return false;
    if (((uintptr_t)a.in % 4) != 0 || ((uintptr_t)a.out % 8) != 0 || ((long long)a.H * a.W) % 4 != 0) return false;
    int Hp, Wp, pad_l;
    taps_geometry(a, Hp, Wp, pad_l);
    return 2ll * (((long long)Hp * Wp + 15) & ~15ll) + a.KH * 2 * 8 * 4 + 64 <= 64 * 1024;
}
cudaError_t launch_dwconv_cin1_taps(const ConvArgs &a, int num_sms, cudaStream_t s) {
    int Hp, Wp, pad_l;
    taps_geometry(a, Hp, Wp, pad_l);
    const int NW = (a.KW + 3) / 4;
    const uint32_t slot_bytes = (uint32_t)(((size_t)Hp * Wp + 15) & ~(size_t)15);
    const size_t smem = (((size_t)a.KH * NW * 8 * 4 + 15) & ~(size_t)15) + 2 * (size_t)slot_bytes;
    const bool full = a.lo == -128.f && a.hi == 127.f;
    using Fn = void (*)(ConvArgs, int, int, int, uint32_t);
    Fn fn = NW == 1 ? (full ? dwconv_cin1_taps_kernel<1, true> : dwconv_cin1_taps_kernel<1, false>)
                    : (full ? dwconv_cin1_taps_kernel<2, true> : dwconv_cin1_taps_kernel<2, false>);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(64 * 1024));
    if (e != cudaSuccess) return e;
    long long ctas = (long long)num_sms * 3;
    if (ctas > a.batch) ctas = a.batch;
    return launch_pdl(fn, dim3((unsigned)ctas), dim3(kTapsThreads), smem, s, a.pdl, a, Hp, Wp, pad_l, slot_bytes);
}

bool dwconv_cin1_eligible(const ConvArgs &a) {
    return a.depthwise && !a.is_u8 && a.Cin == 1 && (a.Cout % 4) == 0 && a.Cout >= 4 && a.Cout <= 16 && a.kcorr != nullptr && !a.big_acc &&
           (size_t)a.KH * a.KW * a.Cout * 4 <= 40 * 1024;
}
cudaError_t launch_dwconv_cin1(const ConvArgs &a, cudaStream_t s) {
    const long long per = (long long)a.OH * a.OW;
    if (per <= 0 || a.batch <= 0) return cudaSuccess;
    const dim3 grid = grid2(per, 128, a.batch);
    const FastDiv fow((uint32_t)a.OW);
    const uint32_t n = (uint32_t)per;
    const size_t sm = (size_t)a.KH * a.KW * a.Cout * 4;
    const bool k33 = a.KH == 3 && a.KW == 3;
    switch (a.Cout) {
        case 4: dwconv_cin1_kernel<4, 0, 0><<<grid, 128, sm, s>>>(a, n, fow); break;
        case 8:
            if (k33) dwconv_cin1_kernel<8, 3, 3><<<grid, 128, sm, s>>>(a, n, fow);
            else dwconv_cin1_kernel<8, 0, 0><<<grid, 128, sm, s>>>(a, n, fow);
            break;
        case 12: dwconv_cin1_kernel<12, 0, 0><<<grid, 128, sm, s>>>(a, n, fow); break;
        case 16: dwconv_cin1_kernel<16, 0, 0><<<grid, 128, sm, s>>>(a, n, fow); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// 1x1 conv on CUDA cores (shapes the tcgen05 GEMM does not take, e.g. person_detect's final 256 -> 2 layer):
// int8, w_zp == 0, Cin % 4 == 0.  One thread = up to 4 output channels of one pixel; dp4a over the input channels.
// ------------------------------------------------------------------------------------------------
template <bool U8, bool WZP>
__global__ void __launch_bounds__(256) pwconv_dp4a_kernel(ConvArgs a, uint32_t items_per_sample, FastDiv fd_g, FastDiv fd_ow) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= items_per_sample) return;
    const int K4 = a.Cin >> 2;
    uint32_t p, g, i, j;
    fd_g.divmod(idx, p, g);
    fd_ow.divmod(p, i, j);
    const int co = 4 * (int)g;
    const int nco = min(4, a.Cout - co);
    const uint32_t *w0 = reinterpret_cast<const uint32_t *>(a.w) + (size_t)co * K4;
    const uint32_t *w1 = w0 + (nco > 1 ? K4 : 0), *w2 = w0 + (nco > 2 ? 2 * K4 : 0), *w3 = w0 + (nco > 3 ? 3 * K4 : 0);
    for (long long b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const uint32_t *x = reinterpret_cast<const uint32_t *>(a.in) + (((size_t)b * a.H + (size_t)a.sh * i) * a.W + (size_t)a.sw * j) * K4;
        int acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, vsum = 0;
#pragma unroll 4
        for (int k = 0; k < K4; ++k) {
            const uint32_t v = __ldg(x + k);
            acc0 = dot4<U8>(v, __ldg(w0 + k), acc0);
            acc1 = dot4<U8>(v, __ldg(w1 + k), acc1);
            acc2 = dot4<U8>(v, __ldg(w2 + k), acc2);
            acc3 = dot4<U8>(v, __ldg(w3 + k), acc3);
            if (WZP) vsum = dot4<U8>(v, 0x01010101u, vsum);              // conv_2d.rs:74-76: the view sum, once per pixel
        }
        const int accs[4] = {acc0, acc1, acc2, acc3};
        uint8_t *o = a.out + ((size_t)b * a.OH * a.OW + p) * a.Cout + co;
        for (int u = 0; u < nco; ++u) {
            int t = accs[u] - a.kcorr[co + u];
            if (WZP) { const int fz = a.w_zp[co + u]; t += fz * (a.Cin * a.in_zp - vsum); }     // - fz * sum(v) + len * Cin * iz * fz  (len = 1)
            o[u] = (uint8_t)requant_nx<true>(t, a.c0z[co + u], a.c1[co + u], a.lo, a.hi);
        }
    }
}

bool pwconv_dp4a_eligible(const ConvArgs &a) {
    // the kernel reads in[(sh*i)*W + sw*j] unchecked: every output position must map inside the input (a model whose declared
    // output is larger than the strided input would need the reference's padding semantics -> generic kernel).  int8 or uint8, any
    // weight zero-points.
    return !a.depthwise && a.KH == 1 && a.KW == 1 && (a.Cin % 4) == 0 && a.kcorr != nullptr && a.w_zp != nullptr && (long long)a.sh * (a.OH - 1) < a.H &&
           (long long)a.sw * (a.OW - 1) < a.W;
}
cudaError_t launch_pwconv_dp4a(const ConvArgs &a, cudaStream_t s) {
    const int G = (a.Cout + 3) / 4;
    const long long per = (long long)a.OH * a.OW * G;
    if (per <= 0 || a.batch <= 0) return cudaSuccess;
    const dim3 grid = grid2(per, 256, a.batch);
    const FastDiv fg((uint32_t)G), fow((uint32_t)a.OW);
    if (a.is_u8) {
        if (a.wzp_nonzero) pwconv_dp4a_kernel<true, true><<<grid, 256, 0, s>>>(a, (uint32_t)per, fg, fow);
        else pwconv_dp4a_kernel<true, false><<<grid, 256, 0, s>>>(a, (uint32_t)per, fg, fow);
    } else {
        if (a.wzp_nonzero) pwconv_dp4a_kernel<false, true><<<grid, 256, 0, s>>>(a, (uint32_t)per, fg, fow);
        else pwconv_dp4a_kernel<false, false><<<grid, 256, 0, s>>>(a, (uint32_t)per, fg, fow);
    }
    return cudaGetLastError();
}

// ================================================================================================
// FAST: fully connected with few outputs (speech: 4000 -> 4): one warp per sample, 128-bit loads, dp4a,
// shuffle reduction.  Pure bandwidth: K bytes per sample.
// ================================================================================================
template <int N_T, bool U8>
__global__ void __launch_bounds__(256) fc_warp_kernel(FcArgs a) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (warp >= a.batch) return;
    const int4 *x = reinterpret_cast<const int4 *>(a.in + (size_t)warp * a.K);
    const int chunks = a.K >> 4;
    int acc[N_T];
#pragma unroll
    for (int j = 0; j < N_T; ++j) acc[j] = 0;
    int rowsum = 0;
    for (int t = lane; t < chunks; t += 32) {
        const int4 v = __ldg(x + t);
        rowsum = dot4<U8>(v.x, 0x01010101u, rowsum); rowsum = dot4<U8>(v.y, 0x01010101u, rowsum);
        rowsum = dot4<U8>(v.z, 0x01010101u, rowsum); rowsum = dot4<U8>(v.w, 0x01010101u, rowsum);
#pragma unroll
        for (int j = 0; j < N_T; ++j) {
            if (j < a.N) {
                const int4 w = __ldg(reinterpret_cast<const int4 *>(a.w + (size_t)j * a.K) + t);
                acc[j] = dot4<U8>(v.x, w.x, acc[j]); acc[j] = dot4<U8>(v.y, w.y, acc[j]);
                acc[j] = dot4<U8>(v.z, w.z, acc[j]); acc[j] = dot4<U8>(v.w, w.w, acc[j]);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        rowsum += __shfl_xor_sync(0xffffffffu, rowsum, off);
#pragma unroll
        for (int j = 0; j < N_T; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
    }
    if (a.sm_out == nullptr) {
#pragma unroll
        for (int j = 0; j < N_T; ++j) {
            if (lane == j && j < a.N) {
                const int t = acc[j] - rowsum * a.w_zp - a.c2[j] + a.c3;
                a.out[(size_t)warp * a.N + j] = (uint8_t)requant(t, a.c0z[j], a.c1, a.lo, a.hi);
            }
        }
        return;
    }
    // fused softmax tail: every lane holds all N sums after the butterfly; lane 0 finishes the sample exactly as
    // fc (above) + softmax_kernel would (same summation order, same divisions)
    if (lane != 0) return;
    int q[N_T];
#pragma unroll
    for (int j = 0; j < N_T; ++j)
        q[j] = j < a.N ? (requant(acc[j] - rowsum * a.w_zp - a.c2[j] + a.c3, a.c0z[j], a.c1, a.lo, a.hi) & 0xff) : 0;
    if (a.out)
        for (int j = 0; j < a.N; ++j) a.out[(size_t)warp * a.N + j] = (uint8_t)q[j];
    float sumexp = 0.0f;
    for (int j = 0; j < a.sm_cols; ++j)
        for (int i = 0; i < a.sm_rows; ++i) sumexp = __fadd_rn(sumexp, __ldg(a.exp_lut + q[i * a.sm_cols + j]));
    for (int k = 0; k < a.N; ++k) {
        const float t = __fadd_rn(__fdiv_rn(__fdiv_rn(__ldg(a.exp_lut + q[k]), sumexp), a.sm_out_scale), a.sm_out_zp);
        const int y = round_clamp(t, a.sm_lo, a.sm_hi);
        a.sm_out[(size_t)warp * a.N + k] = (uint8_t)y;
        if (a.out_f32) a.out_f32[(size_t)warp * a.N + k] = __fmul_rn(a.dq_scale, __fsub_rn(__int2float_rn(y), a.dq_zp));   // dequantize_kernel's arithmetic
    }
}

// ================================================================================================
// FAST: classifier tail.  One warp per sample runs average_pool_2d (whole image -> 1x1), the 1x1 conv_2d with a handful of
// outputs and the softmax, each with exactly the arithmetic of its stand-alone kernel (pool_generic_kernel,
// pwconv_dp4a_kernel, softmax_kernel): three launches that each move a few bytes per sample become one.
// Lane l owns channel words l, l + 32, ... of every pixel, so the loads are fully coalesced.
// ================================================================================================
template <int WPL>   // channel words (4 channels) per lane: C == 128 * WPL
__global__ void __launch_bounds__(256) tail_fused_kernel(TailArgs a) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (b >= a.batch) return;
    const int CW = a.C >> 2;
    const uint32_t *x = reinterpret_cast<const uint32_t *>(a.in + (size_t)b * a.HW * a.C);
    int sum[WPL][4];
#pragma unroll
    for (int k = 0; k < WPL; ++k) sum[k][0] = sum[k][1] = sum[k][2] = sum[k][3] = 0;
    for (int p = 0; p < a.HW; ++p) {
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
            const uint32_t v = __ldg(x + (size_t)p * CW + lane + 32 * k);
            sum[k][0] += sx8<0>(v); sum[k][1] += sx8<1>(v); sum[k][2] += sx8<2>(v); sum[k][3] += sx8<3>(v);
        }
    }
    int acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
        int q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float xm = __fmul_rn(a.inv_len, __int2float_rn(sum[k][u]));
            q[u] = round_clamp(__fadd_rn(__fmul_rn(a.pool_c0, xm), a.pool_c1), a.pool_lo, a.pool_hi);
        }
        const int pv = (int)pack4(q[0], q[1], q[2], q[3]);
#pragma unroll
        for (int o = 0; o < 8; ++o)
            if (o < a.N) acc[o] = __dp4a(pv, (int)__ldg(reinterpret_cast<const uint32_t *>(a.w) + (size_t)o * CW + lane + 32 * k), acc[o]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
    if (lane != 0) return;
    int q[8];
#pragma unroll
    for (int o = 0; o < 8; ++o)
        q[o] = o < a.N ? (requant_nx<true>(acc[o] - __ldg(a.kcorr + o), __ldg(a.c0z + o), __ldg(a.c1 + o), a.conv_lo, a.conv_hi) & 0xff) : 0;
    if (a.logits)
        for (int o = 0; o < a.N; ++o) a.logits[(size_t)b * a.N + o] = (uint8_t)q[o];
    float sumexp = 0.0f;                                            // softmax_kernel's summation order
    for (int j = 0; j < a.sm_cols; ++j)
        for (int i = 0; i < a.sm_rows; ++i) sumexp = __fadd_rn(sumexp, __ldg(a.exp_lut + q[i * a.sm_cols + j]));
    for (int k = 0; k < a.N; ++k) {
        const float t = __fadd_rn(__fdiv_rn(__fdiv_rn(__ldg(a.exp_lut + q[k]), sumexp), a.out_scale), a.out_zp);
        const int y = round_clamp(t, a.sm_lo, a.sm_hi);
        a.out[(size_t)b * a.N + k] = (uint8_t)y;
        if (a.out_f32) a.out_f32[(size_t)b * a.N + k] = __fmul_rn(a.dq_scale, __fsub_rn(__int2float_rn(y), a.dq_zp));     // dequantize_kernel's arithmetic
    }
}

cudaError_t launch_tail_fused(const TailArgs &a, cudaStream_t s) {
    if (a.batch <= 0) return cudaSuccess;
    if (a.C % 128 != 0 || a.C > 512 || a.N < 1 || a.N > 8 || a.sm_rows * a.sm_cols != a.N) return cudaErrorInvalidValue;
    const unsigned grid = grid_for(a.batch * 32, 256);
    switch (a.C / 128) {
        case 1: return launch_pdl(tail_fused_kernel<1>, dim3(grid), dim3(256), 0, s, a.pdl, a);
        case 2: return launch_pdl(tail_fused_kernel<2>, dim3(grid), dim3(256), 0, s, a.pdl, a);
        case 3: return launch_pdl(tail_fused_kernel<3>, dim3(grid), dim3(256), 0, s, a.pdl, a);
        default: return launch_pdl(tail_fused_kernel<4>, dim3(grid), dim3(256), 0, s, a.pdl, a);
    }
}

bool fc_warp_eligible(const FcArgs &a) { return (a.K % 16) == 0 && a.N >= 1 && a.N <= 8; }     // int8 or uint8, any weight zero-point
cudaError_t launch_fc_warp(const FcArgs &a, cudaStream_t s) {
    if (a.batch <= 0) return cudaSuccess;
    const unsigned grid = grid_for(a.batch * 32, 256);
    if (a.is_u8) {
        if (a.N <= 4) return launch_pdl(fc_warp_kernel<4, true>, dim3(grid), dim3(256), 0, s, a.pdl, a);
        return launch_pdl(fc_warp_kernel<8, true>, dim3(grid), dim3(256), 0, s, a.pdl, a);
    }
    if (a.N <= 4) return launch_pdl(fc_warp_kernel<4, false>, dim3(grid), dim3(256), 0, s, a.pdl, a);
    return launch_pdl(fc_warp_kernel<8, false>, dim3(grid), dim3(256), 0, s, a.pdl, a);
}

}  // namespace mf
