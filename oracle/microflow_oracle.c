/*
 * microflow_oracle.c -- CPU restatement of MicroFlow's quantized op-kernel path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA product in
 * microflow_rs_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may build, load or call it.  The product never links or calls it.
 *
 * It restates, in plain scalar C, the algorithm of the reference (matteocarnelos/microflow-rs,
 * /root/reference at survey time).  Every function cites the reference file:line it follows.
 * The reference itself is Rust and cannot be compiled in this image (no rustc/cargo), so parity
 * is pinned against the reference's own golden vectors instead (tests/golden/, see
 * tests/test_oracle_*.py): the per-op KATs in src/ops/<op>.rs `mod tests`, the macro `preprocess`
 * KATs, the three end-to-end goldens in tests/<model>.rs, and the 500 rows of
 * analysis/accuracy/data/sine-microflow.csv.
 *
 * Third-party arithmetic that is NOT in /root/reference (Cargo.toml:27 `libm = "0.2"`, no
 * Cargo.lock): `roundf` (C semantics, half away from zero -> we call the C library's roundf) and
 * `expf` (Rust libm 0.2 = the FreeBSD/musl e_expf.c algorithm; restated below from the published
 * algorithm).  expf is pinned only by the vectors listed above: beyond them, softmax parity is
 * "unpinned at the expf boundary" (the pre-softmax int8 logits are fully pinned).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fwrapv; no -ffast-math, no FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define MFO_API __attribute__((visibility("default")))

enum { MFO_OK = 0, MFO_ERR_PARSE = 1, MFO_ERR_UNSUPPORTED = 2, MFO_ERR_OOB = 3, MFO_ERR_ARG = 4 };
enum { MFO_ACT_NONE = 0, MFO_ACT_RELU = 1, MFO_ACT_RELU6 = 3 }; /* tflite.fbs:552 ActivationFunctionType */
enum { MFO_PAD_SAME = 0, MFO_PAD_VALID = 1 };                    /* tflite.fbs:548 Padding */
enum { MFO_OP_AVGPOOL = 1, MFO_OP_CONV = 3, MFO_OP_DWCONV = 4, MFO_OP_FC = 9, MFO_OP_RESHAPE = 22,
       MFO_OP_SOFTMAX = 25 };                                    /* BuiltinOperator codes */

/* ---- scalar element access: T = i8 or u8 (src/quantize.rs:6-7 `Quantized`) ---------------- */
static inline int32_t ld_elem(const uint8_t *p, int is_u8) {
    return is_u8 ? (int32_t)(*p) : (int32_t)(int8_t)(*p);
}
/* Rust `x as T` for f32 -> i8/u8: saturating, NaN -> 0 (simba from_superset_unchecked). */
static inline int32_t sat_cast(float x, int is_u8) {
    if (x != x) return 0;
    if (is_u8) { if (x <= 0.0f) return 0; if (x >= 255.0f) return 255; return (int32_t)x; }
    if (x <= -128.0f) return -128;
    if (x >= 127.0f) return 127;
    return (int32_t)x;
}

/* ---- src/quantize.rs:16-18 / :27-29 -------------------------------------------------------- */
MFO_API int32_t mfo_quantize(float input, float scale, int32_t zero_point, int is_u8) {
    return sat_cast(roundf(input / scale + (float)zero_point), is_u8);
}
MFO_API float mfo_dequantize(int32_t input, float scale, int32_t zero_point) {
    return scale * ((float)input - (float)zero_point);
}

/* ---- src/activation.rs:21-23, :32-34 ------------------------------------------------------- */
MFO_API int32_t mfo_relu(int32_t input, int32_t zero_point) { return input > zero_point ? input : zero_point; }
MFO_API int32_t mfo_relu6(int32_t input, float scale, int32_t zero_point, int is_u8) {
    int32_t r = mfo_relu(input, zero_point);
    int32_t six = mfo_quantize(6.0f, scale, zero_point, is_u8);
    return r < six ? r : six;
}
static inline int32_t apply_act(int32_t y, int act, float out_scale, int32_t out_zp, int is_u8) {
    switch (act) {
        case MFO_ACT_RELU: return mfo_relu(y, out_zp);
        case MFO_ACT_RELU6: return mfo_relu6(y, out_scale, out_zp, is_u8);
        default: return y;
    }
}

/* ---- libm 0.2 expf (FreeBSD/musl e_expf.c algorithm; see header) -------------------------- */
static float mfo_scalbnf(float x, int n) {
    union { float f; uint32_t i; } u;
    float y = x;
    if (n > 127) {
        y *= 0x1p127f; n -= 127;
        if (n > 127) { y *= 0x1p127f; n -= 127; if (n > 127) n = 127; }
    } else if (n < -126) {
        y *= 0x1p-126f * 0x1p24f; n += 126 - 24;
        if (n < -126) { y *= 0x1p-126f * 0x1p24f; n += 126 - 24; if (n < -126) n = -126; }
    }
    u.i = (uint32_t)(0x7f + n) << 23;
    return y * u.f;
}
MFO_API float mfo_expf(float x) {
    static const float half[2] = {0.5f, -0.5f};
    const float ln2hi = 6.9314575195e-1f, ln2lo = 1.4286067653e-6f, invln2 = 1.4426950216e+0f;
    const float P1 = 1.6666625440e-1f, P2 = -2.7667332906e-3f;
    union { float f; uint32_t i; } u;
    float hi, lo, c, xx, y;
    int k, sign;
    uint32_t hx;
    u.f = x; hx = u.i;
    sign = (int)(hx >> 31);
    hx &= 0x7fffffff;
    if (hx >= 0x42aeac50) {
        if (hx > 0x7f800000) return x;
        if (hx >= 0x42b17218 && !sign) { x *= 0x1p127f; return x; }
        if (sign) { if (hx >= 0x42cff1b5) return 0.0f; }
    }
    if (hx > 0x3eb17218) {
        if (hx > 0x3f851592) k = (int)(invln2 * x + half[sign]);
        else k = 1 - sign - sign;
        hi = x - (float)k * ln2hi;
        lo = (float)k * ln2lo;
        x = hi - lo;
    } else if (hx > 0x39000000) {
        k = 0; hi = x; lo = 0.0f;
    } else {
        return 1.0f + x;
    }
    xx = x * x;
    c = x - xx * (P1 + xx * P2);
    y = 1.0f + (x * c / (2.0f - c) - lo + hi);
    if (k == 0) return y;
    return mfo_scalbnf(y, k);
}
/* src/activation.rs:44-46 */
MFO_API int32_t mfo_softmax_scalar(float input, float sum, float scale, int32_t zero_point, int is_u8) {
    return mfo_quantize(mfo_expf(input) / sum, scale, zero_point, is_u8);
}

/* ---- src/tensor.rs:180-228  Tensor4D::view ------------------------------------------------- *
 * Copies the KH x KW x C window anchored at focus (i,j) into `vbuf` (zero where out of bounds), fills
 * `mask` and returns len (= number of in-bounds taps), or -1 if a VALID view indexes out of bounds (the
 * reference would panic at tensor.rs:222). */
static int view_extract(const uint8_t *in, int H, int W, int C, int i, int j, int KH, int KW, int pad,
                        int sh, int sw, uint8_t *vbuf, uint8_t *mask) {
    int len = KH * KW;
    for (int m = 0; m < KH; ++m)
        for (int n = 0; n < KW; ++n) {
            uint8_t *dst = vbuf + ((size_t)m * KW + n) * C;
            int r, c, ok = 1;
            if (pad == MFO_PAD_SAME) {
                int shr = (KH - 1) / 2, shc = (KW - 1) / 2;          /* tensor.rs:193 */
                r = sh * i + m - shr; c = sw * j + n - shc;          /* checked_sub -> None if negative */
                if (r < 0 || c < 0 || r >= H || c >= W) ok = 0;      /* tensor.rs:196-218 */
            } else {
                r = sh * i + m; c = sw * j + n;                      /* tensor.rs:222 */
                if (r >= H || c >= W) return -1;
            }
            mask[m * KW + n] = (uint8_t)ok;
            if (ok) memcpy(dst, in + ((size_t)r * W + c) * C, (size_t)C);
            else { memset(dst, 0, (size_t)C); len -= 1; }
        }
    return len;
}

/* ---- src/ops/conv_2d.rs:28-108 ------------------------------------------------------------- *
 * in: [H][W][Cin] ; filt: OHWI [Cout][KH][KW][Cin] (microflow-macros/src/tensor.rs:176-200);
 * filt_zp has n_fq entries (per-channel if n_fq>1 else per-tensor, conv_2d.rs:59-63); c1 has n_c1. */
MFO_API int mfo_conv_2d(int is_u8, const uint8_t *in, int H, int W, int Cin, int32_t in_zp,
                        const uint8_t *filt, int Cout, int KH, int KW, const int32_t *filt_zp, int n_fq,
                        float out_scale, int32_t out_zp, int act, int pad, int sh, int sw,
                        const float *c0, const float *c1, int n_c1, uint8_t *out, int OH, int OW) {
    size_t vsz = (size_t)KH * KW * Cin;
    uint8_t *vbuf = (uint8_t *)malloc(vsz ? vsz : 1), *mask = (uint8_t *)malloc((size_t)KH * KW + 1);
    int rc = MFO_OK;
    for (int i = 0; i < OH && rc == MFO_OK; ++i)
        for (int j = 0; j < OW; ++j) {
            int len = view_extract(in, H, W, Cin, i, j, KH, KW, pad, sh, sw, vbuf, mask); /* :52-53 */
            if (len < 0) { rc = MFO_ERR_OOB; break; }
            for (int b = 0; b < Cout; ++b) {                                            /* :55 */
                const uint8_t *f = filt + (size_t)b * vsz;
                int32_t fz = filt_zp[b < n_fq ? b : 0];
                int32_t dot = 0, vsum = 0, fsum = 0;
                for (size_t t = 0; t < vsz; ++t) dot += ld_elem(vbuf + t, is_u8) * ld_elem(f + t, is_u8); /* :66-72 */
                for (size_t t = 0; t < vsz; ++t) vsum += ld_elem(vbuf + t, is_u8);     /* :74-76 */
                int32_t x1 = vsum * fz;
                for (int t = 0; t < KH * KW; ++t)                                       /* :83-89 */
                    if (mask[t]) for (int ch = 0; ch < Cin; ++ch) fsum += ld_elem(f + (size_t)t * Cin + ch, is_u8);
                int32_t k2 = in_zp * fsum;
                int32_t k3 = len * Cin * in_zp * fz;                                    /* :90 */
                float v = (float)out_zp + c0[b] + c1[b < n_c1 ? b : 0] * (float)(dot - x1 - k2 + k3); /* :93-98 */
                int32_t y = sat_cast(roundf(v), is_u8);
                y = apply_act(y, act, out_scale, out_zp, is_u8);                        /* :100-104 */
                out[((size_t)i * OW + j) * Cout + b] = (uint8_t)y;
            }
        }
    free(vbuf); free(mask);
    return rc;
}

/* ---- src/ops/depthwise_conv_2d.rs:28-105 --------------------------------------------------- *
 * w: [1][KH][KW][Cout]; output channel c reads input channel c, or channel 0 if c >= Cin (:67,:72). */
MFO_API int mfo_depthwise_conv_2d(int is_u8, const uint8_t *in, int H, int W, int Cin, int32_t in_zp,
                                  const uint8_t *w, int Cout, int KH, int KW, const int32_t *w_zp, int n_wq,
                                  float out_scale, int32_t out_zp, int act, int pad, int sh, int sw,
                                  const float *c0, const float *c1, int n_c1, uint8_t *out, int OH, int OW) {
    size_t vsz = (size_t)KH * KW * Cin;
    uint8_t *vbuf = (uint8_t *)malloc(vsz ? vsz : 1), *mask = (uint8_t *)malloc((size_t)KH * KW + 1);
    int rc = MFO_OK;
    for (int i = 0; i < OH && rc == MFO_OK; ++i)
        for (int j = 0; j < OW; ++j) {
            int len = view_extract(in, H, W, Cin, i, j, KH, KW, pad, sh, sw, vbuf, mask);
            if (len < 0) { rc = MFO_ERR_OOB; break; }
            for (int c = 0; c < Cout; ++c) {
                int ci = c < Cin ? c : 0;
                int32_t wz = w_zp[c < n_wq ? c : 0];
                int32_t dot = 0, vsum = 0, wsum = 0;
                for (int t = 0; t < KH * KW; ++t)                                       /* :66-69 */
                    dot += ld_elem(vbuf + (size_t)t * Cin + ci, is_u8) * ld_elem(w + (size_t)t * Cout + c, is_u8);
                for (int t = 0; t < KH * KW; ++t) vsum += ld_elem(vbuf + (size_t)t * Cin + ci, is_u8); /* :71-73 */
                int32_t x1 = vsum * wz;
                for (int t = 0; t < KH * KW; ++t) if (mask[t]) wsum += ld_elem(w + (size_t)t * Cout + c, is_u8); /* :79-86 */
                int32_t k2 = in_zp * wsum;
                int32_t k3 = len * in_zp * wz;                                          /* :87 */
                float v = (float)out_zp + c0[c] + c1[c < n_c1 ? c : 0] * (float)(dot - x1 - k2 + k3); /* :90-95 */
                int32_t y = sat_cast(roundf(v), is_u8);
                y = apply_act(y, act, out_scale, out_zp, is_u8);
                out[((size_t)i * OW + j) * Cout + c] = (uint8_t)y;
            }
        }
    free(vbuf); free(mask);
    return rc;
}

/* ---- src/ops/fully_connected.rs:24-82 ------------------------------------------------------ *
 * in: [R][K] row-major; w: TFLite bytes [N][K] (the reference's W[k][j] = bytes[j*K+k],
 * microflow-macros/src/tensor.rs:98-114). */
MFO_API int mfo_fully_connected(int is_u8, const uint8_t *in, int R, int K, const uint8_t *w, int N,
                                int32_t w_zp, float out_scale, int32_t out_zp, int act, const float *c0,
                                float c1, const int32_t *c2, int32_t c3, uint8_t *out) {
    for (int i = 0; i < R; ++i) {
        int32_t rowsum = 0;
        for (int k = 0; k < K; ++k) rowsum += ld_elem(in + (size_t)i * K + k, is_u8);   /* :58-64 */
        int32_t x1 = rowsum * w_zp;
        for (int j = 0; j < N; ++j) {
            int32_t dot = 0;
            for (int k = 0; k < K; ++k)                                                 /* :47-56 */
                dot += ld_elem(in + (size_t)i * K + k, is_u8) * ld_elem(w + (size_t)j * K + k, is_u8);
            float v = (float)out_zp + c0[j] + c1 * (float)(dot - x1 - c2[j] + c3);      /* :68-73 */
            int32_t y = sat_cast(roundf(v), is_u8);
            y = apply_act(y, act, out_scale, out_zp, is_u8);
            out[(size_t)i * N + j] = (uint8_t)y;
        }
    }
    return MFO_OK;
}

/* ---- src/ops/average_pool_2d.rs:29-66 ------------------------------------------------------ */
MFO_API int mfo_average_pool_2d(int is_u8, const uint8_t *in, int H, int W, int C, int FH, int FW,
                                float out_scale, int32_t out_zp, int act, int pad, int sh, int sw, float c0,
                                float c1, uint8_t *out, int OH, int OW) {
    size_t vsz = (size_t)FH * FW * C;
    uint8_t *vbuf = (uint8_t *)malloc(vsz ? vsz : 1), *mask = (uint8_t *)malloc((size_t)FH * FW + 1);
    int rc = MFO_OK;
    for (int i = 0; i < OH && rc == MFO_OK; ++i)
        for (int j = 0; j < OW; ++j) {
            int len = view_extract(in, H, W, C, i, j, FH, FW, pad, sh, sw, vbuf, mask);
            if (len < 0) { rc = MFO_ERR_OOB; break; }
            for (int c = 0; c < C; ++c) {
                int32_t s = 0;
                for (int t = 0; t < FH * FW; ++t) s += ld_elem(vbuf + (size_t)t * C + c, is_u8);
                float x = 1.0f / (float)len * (float)s;                                  /* :52-55 */
                int32_t y = sat_cast(roundf(c0 * x + c1), is_u8);                        /* :56 */
                y = apply_act(y, act, out_scale, out_zp, is_u8);
                out[((size_t)i * OW + j) * C + c] = (uint8_t)y;
            }
        }
    free(vbuf); free(mask);
    return rc;
}

/* ---- src/ops/softmax.rs:15-27 -------------------------------------------------------------- *
 * Over the WHOLE rows x cols buffer, nalgebra column-major iteration order for the sum. */
MFO_API int mfo_softmax(int is_u8, const uint8_t *in, int rows, int cols, float in_scale, float out_scale,
                        int32_t out_zp, uint8_t *out) {
    float sum = 0.0f;
    for (int j = 0; j < cols; ++j)
        for (int i = 0; i < rows; ++i)
            sum = sum + mfo_expf((float)ld_elem(in + (size_t)i * cols + j, is_u8) * in_scale);
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) {
            float e = (float)ld_elem(in + (size_t)i * cols + j, is_u8) * in_scale;
            out[(size_t)i * cols + j] = (uint8_t)mfo_softmax_scalar(e, sum, out_scale, out_zp, is_u8);
        }
    return MFO_OK;
}

/* ---- pre-processing constants (the proc-macro's job) ---------------------------------------- */
/* microflow-macros/src/ops/conv_2d.rs:94-114 and depthwise_conv_2d.rs:100-120 (same form; `n_out` is
 * filters.shape[0] for conv, weights.shape[3] for depthwise). */
MFO_API void mfo_conv_preprocess(float in_scale, const float *w_scale, int n_wq, const float *b_scale, int n_bq,
                                 const int32_t *bias, const int32_t *b_zp, int n_bzp, float out_scale, int n_out,
                                 float *c0, float *c1) {
    for (int b = 0; b < n_out; ++b)
        c0[b] = b_scale[b < n_bq ? b : 0] / out_scale * (float)(bias[b] - b_zp[b < n_bzp ? b : 0]);
    for (int b = 0; b < n_wq; ++b) c1[b] = in_scale * w_scale[b] / out_scale;
}
/* microflow-macros/src/ops/fully_connected.rs:100-123.  `shape1` = input.shape[1] as the macro sees it. */
MFO_API void mfo_fc_preprocess(int is_u8, float in_scale, int32_t in_zp, int shape1, const uint8_t *w, int N, int K,
                               float w_scale, int32_t w_zp, float b_scale, const int32_t *bias, int32_t b_zp,
                               float out_scale, float *c0, float *c1, int32_t *c2, int32_t *c3) {
    for (int j = 0; j < N; ++j) c0[j] = b_scale / out_scale * (float)(bias[j] + (-b_zp));
    *c1 = in_scale * w_scale / out_scale;
    for (int j = 0; j < N; ++j) {
        int32_t s = 0;
        for (int k = 0; k < K; ++k) s += ld_elem(w + (size_t)j * K + k, is_u8);
        c2[j] = s * in_zp;
    }
    *c3 = shape1 * in_zp * w_zp;
}
/* microflow-macros/src/ops/average_pool_2d.rs:77-83 */
MFO_API void mfo_pool_preprocess(float in_scale, int32_t in_zp, float out_scale, int32_t out_zp, float *c0, float *c1) {
    *c0 = in_scale / out_scale;
    *c1 = (float)out_zp - (in_scale * (float)in_zp) / out_scale;
}

/* =============================================================================================
 * Model loader: what microflow-macros/src/lib.rs:46-208 does at Rust compile time.
 * Hand-rolled FlatBuffers reader for the ~12 table fields of tflite.fbs the macro touches.
 * ============================================================================================= */
typedef struct { const uint8_t *p; size_t n; } fb_t;
static int fb_ok(const fb_t *fb, size_t off, size_t len) { return off <= fb->n && len <= fb->n - off; }
static uint32_t rd_u32(const fb_t *fb, size_t o) { uint32_t v = 0; if (fb_ok(fb, o, 4)) memcpy(&v, fb->p + o, 4); return v; }
static int32_t rd_i32(const fb_t *fb, size_t o) { return (int32_t)rd_u32(fb, o); }
static uint16_t rd_u16(const fb_t *fb, size_t o) { uint16_t v = 0; if (fb_ok(fb, o, 2)) memcpy(&v, fb->p + o, 2); return v; }
/* absolute position of field `id` of the table at `t`, or 0 when absent */
static size_t fb_field(const fb_t *fb, size_t t, int id) {
    if (!t || !fb_ok(fb, t, 4)) return 0;
    int64_t vt = (int64_t)t - rd_i32(fb, t);
    if (vt < 0 || !fb_ok(fb, (size_t)vt, 4)) return 0;
    uint16_t vsz = rd_u16(fb, (size_t)vt);
    size_t slot = 4 + 2 * (size_t)id;
    if (slot + 2 > vsz) return 0;
    uint16_t off = rd_u16(fb, (size_t)vt + slot);
    return off ? t + off : 0;
}
static size_t fb_indirect(const fb_t *fb, size_t pos) { return pos ? pos + rd_u32(fb, pos) : 0; }
static size_t fb_table(const fb_t *fb, size_t t, int id) { return fb_indirect(fb, fb_field(fb, t, id)); }
/* vector field: returns position of first element, *len = count */
static size_t fb_vec(const fb_t *fb, size_t t, int id, uint32_t *len) {
    size_t v = fb_indirect(fb, fb_field(fb, t, id));
    *len = 0;
    if (!v || !fb_ok(fb, v, 4)) return 0;
    *len = rd_u32(fb, v);
    return v + 4;
}
static size_t fb_vec_table(const fb_t *fb, size_t elems, uint32_t i) { return fb_indirect(fb, elems + 4 * (size_t)i); }
static int32_t fb_i32(const fb_t *fb, size_t t, int id, int32_t def) { size_t f = fb_field(fb, t, id); return f ? rd_i32(fb, f) : def; }
static int32_t fb_i8(const fb_t *fb, size_t t, int id, int32_t def) { size_t f = fb_field(fb, t, id); return f && fb_ok(fb, f, 1) ? (int8_t)fb->p[f] : def; }

#define MFO_MAX_DIMS 4
typedef struct {
    int rank, dims[MFO_MAX_DIMS];
    int type;              /* tflite TensorType: 9 = INT8, 3 = UINT8, 2 = INT32 */
    int n_scale, n_zp;
    float *scale;          /* owned */
    int64_t *zp;           /* owned, raw i64 (cast to T / i32 by truncation where used, tensor.rs:81-88) */
    const uint8_t *data;   /* points into the flatbuffer */
    size_t data_len;
} mfo_tinfo;

typedef struct {
    int op;                /* MFO_OP_* */
    int is_u8;
    int in_rank, in_dims[4], out_rank, out_dims[4];
    float in_scale, out_scale;
    int32_t in_zp, out_zp;
    int act, pad, sh, sw, KH, KW, Cout;
    const uint8_t *w;      /* into flatbuffer copy */
    int n_wq; int32_t *w_zp;
    float *c0, *c1; int n_c1;
    int32_t *c2, c3; float fc_c1; float pool_c0, pool_c1;
    size_t out_elems, in_elems;
} mfo_layer;

typedef struct mfo_model {
    uint8_t *buf; size_t len;
    int n_layers; mfo_layer *layers;
    int is_u8_in, is_u8_out;
    int in_rank, in_dims[4], out_rank, out_dims[4];
    float in_scale, out_scale; int32_t in_zp, out_zp;
    size_t in_elems, out_elems, max_elems;
} mfo_model;

static void tinfo_free(mfo_tinfo *t) { free(t->scale); free(t->zp); t->scale = NULL; t->zp = NULL; }
static int load_tinfo(const fb_t *fb, size_t tensors, uint32_t n_tensors, size_t buffers, uint32_t n_buffers,
                      int32_t idx, mfo_tinfo *ti) {
    memset(ti, 0, sizeof *ti);
    if (idx < 0 || (uint32_t)idx >= n_tensors) return MFO_ERR_PARSE;
    size_t t = fb_vec_table(fb, tensors, (uint32_t)idx);
    uint32_t n; size_t sh = fb_vec(fb, t, 0, &n);
    if (n > MFO_MAX_DIMS) return MFO_ERR_UNSUPPORTED;
    ti->rank = (int)n;
    for (uint32_t i = 0; i < n; ++i) ti->dims[i] = rd_i32(fb, sh + 4 * i);
    ti->type = fb_i8(fb, t, 1, 0);
    size_t q = fb_table(fb, t, 4);
    uint32_t ns = 0, nz = 0;
    size_t sv = q ? fb_vec(fb, q, 2, &ns) : 0, zv = q ? fb_vec(fb, q, 3, &nz) : 0;
    ti->n_scale = (int)ns; ti->n_zp = (int)nz;
    ti->scale = (float *)calloc(ns ? ns : 1, sizeof(float));
    ti->zp = (int64_t *)calloc(nz ? nz : 1, sizeof(int64_t));
    for (uint32_t i = 0; i < ns; ++i) { uint32_t b = rd_u32(fb, sv + 4 * i); memcpy(&ti->scale[i], &b, 4); }
    for (uint32_t i = 0; i < nz; ++i) { if (fb_ok(fb, zv + 8 * (size_t)i, 8)) memcpy(&ti->zp[i], fb->p + zv + 8 * (size_t)i, 8); }
    uint32_t bidx = (uint32_t)fb_i32(fb, t, 2, 0);
    if (bidx < n_buffers) {
        size_t b = fb_vec_table(fb, buffers, bidx);
        uint32_t dl; size_t d = fb_vec(fb, b, 0, &dl);
        if (d && fb_ok(fb, d, dl)) { ti->data = fb->p + d; ti->data_len = dl; }
    }
    return MFO_OK;
}
static int32_t zp_as_T(int64_t z, int is_u8) { return is_u8 ? (int32_t)(uint8_t)z : (int32_t)(int8_t)z; } /* tensor.rs:81-88 */
static size_t prod_dims(const int *d, int r) { size_t p = 1; for (int i = 0; i < r; ++i) p *= (size_t)d[i]; return p; }

static void layer_free(mfo_layer *L) { free(L->w_zp); free(L->c0); free(L->c1); free(L->c2); }
MFO_API void mfo_model_free(mfo_model *m) {
    if (!m) return;
    for (int i = 0; i < m->n_layers; ++i) layer_free(&m->layers[i]);
    free(m->layers); free(m->buf); free(m);
}

/* microflow-macros/src/lib.rs:46-208 (graph walk) + ops/<op>.rs `new`/`preprocess` */
MFO_API int mfo_model_load(const uint8_t *data, size_t len, mfo_model **out) {
    if (!data || len < 8 || !out) return MFO_ERR_ARG;
    mfo_model *m = (mfo_model *)calloc(1, sizeof *m);
    m->buf = (uint8_t *)malloc(len); memcpy(m->buf, data, len); m->len = len;
    fb_t fbv = {m->buf, len}; const fb_t *fb = &fbv;
    int rc = MFO_OK;
    size_t model = rd_u32(fb, 0);
    uint32_t n_codes, n_sub, n_buffers, n_tensors, n_in, n_out, n_ops;
    size_t codes = fb_vec(fb, model, 1, &n_codes), subs = fb_vec(fb, model, 2, &n_sub), buffers = fb_vec(fb, model, 4, &n_buffers);
    if (!subs || !n_sub || !codes) { mfo_model_free(m); return MFO_ERR_PARSE; }
    size_t sg = fb_vec_table(fb, subs, 0);                                   /* lib.rs:62 subgraph 0 only */
    size_t tensors = fb_vec(fb, sg, 0, &n_tensors), ins = fb_vec(fb, sg, 1, &n_in), outs = fb_vec(fb, sg, 2, &n_out),
           ops = fb_vec(fb, sg, 3, &n_ops);
    if (!tensors || !n_in || !n_out) { mfo_model_free(m); return MFO_ERR_PARSE; }

    mfo_tinfo ti;
    /* model input (lib.rs:66-128): 1-D shape -> [1,n]; type INT8/UINT8; rank 2 or 4 */
    if ((rc = load_tinfo(fb, tensors, n_tensors, buffers, n_buffers, rd_i32(fb, ins), &ti)) != MFO_OK) { mfo_model_free(m); return rc; }
    if (ti.type != 9 && ti.type != 3) { tinfo_free(&ti); mfo_model_free(m); return MFO_ERR_UNSUPPORTED; }
    m->is_u8_in = ti.type == 3;
    m->in_rank = ti.rank; memcpy(m->in_dims, ti.dims, sizeof ti.dims);
    if (m->in_rank == 1) { m->in_rank = 2; m->in_dims[1] = m->in_dims[0]; m->in_dims[0] = 1; }
    if ((m->in_rank != 2 && m->in_rank != 4) || ti.n_scale < 1 || ti.n_zp < 1) { tinfo_free(&ti); mfo_model_free(m); return MFO_ERR_UNSUPPORTED; }
    m->in_scale = ti.scale[0]; m->in_zp = zp_as_T(ti.zp[0], m->is_u8_in);
    m->in_elems = prod_dims(m->in_dims, m->in_rank); m->max_elems = m->in_elems;
    tinfo_free(&ti);

    m->layers = (mfo_layer *)calloc(n_ops ? n_ops : 1, sizeof(mfo_layer));
    for (uint32_t oi = 0; oi < n_ops && rc == MFO_OK; ++oi) {                /* lib.rs:130-151 */
        size_t op = fb_vec_table(fb, ops, oi);
        uint32_t code_idx = (uint32_t)fb_i32(fb, op, 0, 0);
        if (code_idx >= n_codes) { rc = MFO_ERR_PARSE; break; }
        int code = fb_i8(fb, fb_vec_table(fb, codes, code_idx), 0, 0);      /* deprecated_builtin_code, lib.rs:131-137 */
        uint32_t nin, nout; size_t oin = fb_vec(fb, op, 1, &nin), oout = fb_vec(fb, op, 2, &nout);
        size_t opt = fb_table(fb, op, 4);
        mfo_layer *L = &m->layers[m->n_layers];
        memset(L, 0, sizeof *L);
        L->op = code;
        mfo_tinfo tin, tout; memset(&tin, 0, sizeof tin); memset(&tout, 0, sizeof tout);
        if (!nin || !nout) { rc = MFO_ERR_PARSE; break; }
        if ((rc = load_tinfo(fb, tensors, n_tensors, buffers, n_buffers, rd_i32(fb, oin), &tin)) != MFO_OK) break;
        if ((rc = load_tinfo(fb, tensors, n_tensors, buffers, n_buffers, rd_i32(fb, oout), &tout)) != MFO_OK) { tinfo_free(&tin); break; }
        L->is_u8 = tin.type == 3;
        if (code != MFO_OP_RESHAPE && tin.type != 9 && tin.type != 3) rc = MFO_ERR_UNSUPPORTED;
        L->in_rank = tin.rank; memcpy(L->in_dims, tin.dims, sizeof tin.dims);
        L->out_rank = tout.rank; memcpy(L->out_dims, tout.dims, sizeof tout.dims);
        if (L->out_rank == 1) { L->out_rank = 2; L->out_dims[1] = L->out_dims[0]; L->out_dims[0] = 1; }
        if (tin.n_scale) L->in_scale = tin.scale[0];
        if (tin.n_zp) L->in_zp = zp_as_T(tin.zp[0], L->is_u8);
        if (tout.n_scale) L->out_scale = tout.scale[0];
        if (tout.n_zp) L->out_zp = zp_as_T(tout.zp[0], L->is_u8);
        L->in_elems = prod_dims(tin.dims, tin.rank);
        L->out_elems = prod_dims(tout.dims, tout.rank);
        if (L->out_elems > m->max_elems) m->max_elems = L->out_elems;

        if (rc == MFO_OK && (code == MFO_OP_CONV || code == MFO_OP_DWCONV)) {
            mfo_tinfo tw, tb;
            if (nin < 3 || tin.rank != 4 || tout.rank != 4) rc = MFO_ERR_UNSUPPORTED;
            if (rc == MFO_OK && (rc = load_tinfo(fb, tensors, n_tensors, buffers, n_buffers, rd_i32(fb, oin + 4), &tw)) == MFO_OK) {
                if ((rc = load_tinfo(fb, tensors, n_tensors, buffers, n_buffers, rd_i32(fb, oin + 8), &tb)) == MFO_OK) {
                    if (tw.rank != 4 || !tw.data || !tb.data || tw.n_scale < 1 || tw.n_zp < 1 || tb.n_scale < 1 || tb.n_zp < 1 ||
                        tw.data_len < prod_dims(tw.dims, 4)) rc = MFO_ERR_UNSUPPORTED;
                    else {
                        L->KH = tw.dims[1]; L->KW = tw.dims[2];
                        L->Cout = code == MFO_OP_CONV ? tw.dims[0] : tw.dims[3];
                        if (tb.data_len < (size_t)L->Cout * 4 || (code == MFO_OP_CONV && tw.dims[3] != tin.dims[3])) rc = MFO_ERR_UNSUPPORTED;
                    }
                    if (rc == MFO_OK) {
                        L->w = tw.data;
                        L->n_wq = tw.n_zp; L->w_zp = (int32_t *)calloc((size_t)tw.n_zp, 4);
                        for (int i = 0; i < tw.n_zp; ++i) L->w_zp[i] = zp_as_T(tw.zp[i], L->is_u8);
                        int32_t *bias = (int32_t *)malloc((size_t)L->Cout * 4), *bzp = (int32_t *)malloc((size_t)tb.n_zp * 4);
                        memcpy(bias, tb.data, (size_t)L->Cout * 4);
                        for (int i = 0; i < tb.n_zp; ++i) bzp[i] = (int32_t)tb.zp[i];
                        L->c0 = (float *)calloc((size_t)L->Cout, 4); L->n_c1 = tw.n_scale; L->c1 = (float *)calloc((size_t)tw.n_scale, 4);
                        mfo_conv_preprocess(L->in_scale, tw.scale, tw.n_scale, tb.scale, tb.n_scale, bias, bzp, tb.n_zp, L->out_scale, L->Cout, L->c0, L->c1);
                        free(bias); free(bzp);
                        /* options: Conv2DOptions{0 pad,1 sw,2 sh,3 act}; DepthwiseConv2DOptions{0 pad,1 sw,2 sh,4 act} */
                        L->pad = fb_i8(fb, opt, 0, 0); L->sw = fb_i32(fb, opt, 1, 0); L->sh = fb_i32(fb, opt, 2, 0);
                        L->act = fb_i8(fb, opt, code == MFO_OP_CONV ? 3 : 4, 0);
                    }
                    tinfo_free(&tb);
                }
                tinfo_free(&tw);
            }
        } else if (rc == MFO_OK && code == MFO_OP_FC) {
            mfo_tinfo tw, tb;
            if (nin < 3) rc = MFO_ERR_UNSUPPORTED;
            if (rc == MFO_OK && (rc = load_tinfo(fb, tensors, n_tensors, buffers, n_buffers, rd_i32(fb, oin + 4), &tw)) == MFO_OK) {
                if ((rc = load_tinfo(fb, tensors, n_tensors, buffers, n_buffers, rd_i32(fb, oin + 8), &tb)) == MFO_OK) {
                    if (tw.rank != 2 || !tw.data || !tb.data || tw.n_scale < 1 || tw.n_zp < 1 || tb.n_scale < 1 || tb.n_zp < 1) rc = MFO_ERR_UNSUPPORTED;
                    else {
                        int N = tw.dims[0], K = tw.dims[1];
                        int shape1 = tin.rank == 1 ? tin.dims[0] : tin.dims[1];            /* tensor.rs:67-70 */
                        if ((size_t)K != L->in_elems || tw.data_len < (size_t)N * K || tb.data_len < (size_t)N * 4) rc = MFO_ERR_UNSUPPORTED;
                        else {
                            L->Cout = N; L->KH = K; L->w = tw.data;
                            L->n_wq = 1; L->w_zp = (int32_t *)calloc(1, 4); L->w_zp[0] = zp_as_T(tw.zp[0], L->is_u8);
                            int32_t *bias = (int32_t *)malloc((size_t)N * 4); memcpy(bias, tb.data, (size_t)N * 4);
                            L->c0 = (float *)calloc((size_t)N, 4); L->c2 = (int32_t *)calloc((size_t)N, 4);
                            mfo_fc_preprocess(L->is_u8, L->in_scale, L->in_zp, shape1, tw.data, N, K, tw.scale[0], L->w_zp[0], tb.scale[0], bias,
                                              (int32_t)tb.zp[0], L->out_scale, L->c0, &L->fc_c1, L->c2, &L->c3);
                            free(bias);
                            L->act = fb_i8(fb, opt, 0, 0);
                        }
                    }
                    tinfo_free(&tb);
                }
                tinfo_free(&tw);
            }
        } else if (rc == MFO_OK && code == MFO_OP_AVGPOOL) {
            if (tin.rank != 4 || tout.rank != 4) rc = MFO_ERR_UNSUPPORTED;
            else {
                mfo_pool_preprocess(L->in_scale, L->in_zp, L->out_scale, L->out_zp, &L->pool_c0, &L->pool_c1);
                L->pad = fb_i8(fb, opt, 0, 0); L->sw = fb_i32(fb, opt, 1, 0); L->sh = fb_i32(fb, opt, 2, 0);
                L->KW = fb_i32(fb, opt, 3, 0); L->KH = fb_i32(fb, opt, 4, 0); L->act = fb_i8(fb, opt, 5, 0);
            }
        } else if (rc == MFO_OK && code == MFO_OP_SOFTMAX) {
            /* softmax.rs: only the output tensor is read; beta ignored */
        } else if (rc == MFO_OK && code == MFO_OP_RESHAPE) {
            if (L->out_rank != 2 && L->out_rank != 4) rc = MFO_ERR_UNSUPPORTED;      /* ops/reshape.rs:47-55 */
        } else if (rc == MFO_OK) {
            rc = MFO_ERR_UNSUPPORTED;                                                /* lib.rs:148 */
        }
        if (rc == MFO_OK && L->act != MFO_ACT_NONE && L->act != MFO_ACT_RELU && L->act != MFO_ACT_RELU6) rc = MFO_ERR_UNSUPPORTED; /* activation.rs:25-37 */
        if (rc == MFO_OK && L->pad != MFO_PAD_SAME && L->pad != MFO_PAD_VALID) rc = MFO_ERR_UNSUPPORTED;
        tinfo_free(&tin); tinfo_free(&tout);
        m->n_layers += 1;
    }
    if (rc == MFO_OK) {                                                              /* lib.rs:153-183 */
        if ((rc = load_tinfo(fb, tensors, n_tensors, buffers, n_buffers, rd_i32(fb, outs), &ti)) == MFO_OK) {
            if ((ti.type != 9 && ti.type != 3) || ti.n_scale < 1 || ti.n_zp < 1) rc = MFO_ERR_UNSUPPORTED;
            else {
                m->is_u8_out = ti.type == 3;
                m->out_rank = ti.rank; memcpy(m->out_dims, ti.dims, sizeof ti.dims);
                if (m->out_rank == 1) { m->out_rank = 2; m->out_dims[1] = m->out_dims[0]; m->out_dims[0] = 1; }
                if (m->out_rank != 2 && m->out_rank != 4) rc = MFO_ERR_UNSUPPORTED;
                m->out_scale = ti.scale[0]; m->out_zp = zp_as_T(ti.zp[0], m->is_u8_out);
                m->out_elems = prod_dims(m->out_dims, m->out_rank);
            }
            tinfo_free(&ti);
        }
    }
    if (rc != MFO_OK) { mfo_model_free(m); return rc; }
    *out = m;
    return MFO_OK;
}

/* io info: dims[0..3] padded with 1 */
MFO_API int mfo_model_io(const mfo_model *m, int *in_rank, int *in_dims, float *in_scale, int32_t *in_zp, int *out_rank,
                         int *out_dims, float *out_scale, int32_t *out_zp, int *is_u8) {
    *in_rank = m->in_rank; *out_rank = m->out_rank;
    for (int i = 0; i < 4; ++i) { in_dims[i] = i < m->in_rank ? m->in_dims[i] : 1; out_dims[i] = i < m->out_rank ? m->out_dims[i] : 1; }
    *in_scale = m->in_scale; *in_zp = m->in_zp; *out_scale = m->out_scale; *out_zp = m->out_zp; *is_u8 = m->is_u8_in;
    return MFO_OK;
}
MFO_API int mfo_model_num_layers(const mfo_model *m) { return m->n_layers; }
/* info = {op, out_elems, in_elems, act, pad, sh, sw, KH, KW, Cout, in_zp, out_zp, n_c1, n_wq, out_rank, out_dims[4], in_rank, in_dims[4]} */
MFO_API int mfo_model_layer_info(const mfo_model *m, int i, int32_t *info, float *scales) {
    if (i < 0 || i >= m->n_layers) return MFO_ERR_ARG;
    const mfo_layer *L = &m->layers[i];
    int32_t v[24] = {L->op, (int32_t)L->out_elems, (int32_t)L->in_elems, L->act, L->pad, L->sh, L->sw, L->KH, L->KW, L->Cout,
                     L->in_zp, L->out_zp, L->n_c1, L->n_wq, L->out_rank, L->out_dims[0], L->out_dims[1], L->out_dims[2], L->out_dims[3],
                     L->in_rank, L->in_dims[0], L->in_dims[1], L->in_dims[2], L->in_dims[3]};
    memcpy(info, v, sizeof v);
    scales[0] = L->in_scale; scales[1] = L->out_scale;
    return MFO_OK;
}
/* constants of layer i (for the loader parity test): copies up to n floats of c0 / c1 and ints of c2 */
MFO_API int mfo_model_layer_consts(const mfo_model *m, int i, float *c0, float *c1, int32_t *c2, int32_t *c3, int n) {
    if (i < 0 || i >= m->n_layers) return MFO_ERR_ARG;
    const mfo_layer *L = &m->layers[i];
    if (L->op == MFO_OP_AVGPOOL) { c0[0] = L->pool_c0; c1[0] = L->pool_c1; return MFO_OK; }
    for (int k = 0; k < n && k < L->Cout && L->c0; ++k) c0[k] = L->c0[k];
    if (L->op == MFO_OP_FC) { c1[0] = L->fc_c1; for (int k = 0; k < n && k < L->Cout; ++k) c2[k] = L->c2[k]; *c3 = L->c3; }
    else for (int k = 0; k < n && k < L->n_c1 && L->c1; ++k) c1[k] = L->c1[k];
    return MFO_OK;
}

/* The straight-line predict_inner body (lib.rs:198-201): each op consumes the previous op's output.
 * If layer_outs != NULL, layer_outs[i] receives layer i's quantized output (out_elems bytes). */
static int run_layers(const mfo_model *m, const uint8_t *in_q, uint8_t *final_q, uint8_t **layer_outs) {
    uint8_t *a = (uint8_t *)malloc(m->max_elems ? m->max_elems : 1), *b = (uint8_t *)malloc(m->max_elems ? m->max_elems : 1);
    memcpy(a, in_q, m->in_elems);
    int rc = MFO_OK;
    size_t cur_elems = m->in_elems;
    for (int i = 0; i < m->n_layers && rc == MFO_OK; ++i) {
        const mfo_layer *L = &m->layers[i];
        switch (L->op) {
            case MFO_OP_CONV:
                rc = mfo_conv_2d(L->is_u8, a, L->in_dims[1], L->in_dims[2], L->in_dims[3], L->in_zp, L->w, L->Cout, L->KH, L->KW, L->w_zp, L->n_wq,
                                 L->out_scale, L->out_zp, L->act, L->pad, L->sh, L->sw, L->c0, L->c1, L->n_c1, b, L->out_dims[1], L->out_dims[2]);
                break;
            case MFO_OP_DWCONV:
                rc = mfo_depthwise_conv_2d(L->is_u8, a, L->in_dims[1], L->in_dims[2], L->in_dims[3], L->in_zp, L->w, L->Cout, L->KH, L->KW, L->w_zp,
                                           L->n_wq, L->out_scale, L->out_zp, L->act, L->pad, L->sh, L->sw, L->c0, L->c1, L->n_c1, b, L->out_dims[1],
                                           L->out_dims[2]);
                break;
            case MFO_OP_FC:   /* 4-D input is flattened NHWC (tensor.rs:106-114): a no-op on this layout */
                rc = mfo_fully_connected(L->is_u8, a, 1, L->KH, L->w, L->Cout, L->w_zp[0], L->out_scale, L->out_zp, L->act, L->c0, L->fc_c1, L->c2,
                                         L->c3, b);
                break;
            case MFO_OP_AVGPOOL:
                rc = mfo_average_pool_2d(L->is_u8, a, L->in_dims[1], L->in_dims[2], L->in_dims[3], L->KH, L->KW, L->out_scale, L->out_zp, L->act,
                                         L->pad, L->sh, L->sw, L->pool_c0, L->pool_c1, b, L->out_dims[1], L->out_dims[2]);
                break;
            case MFO_OP_SOFTMAX:
                rc = mfo_softmax(L->is_u8, a, L->out_dims[0], (int)(L->out_elems / (size_t)L->out_dims[0]), L->in_scale, L->out_scale, L->out_zp, b);
                break;
            case MFO_OP_RESHAPE:  /* ops/reshape.rs:3-8 + tensor.rs:95-141: NHWC element order preserved */
                if (L->out_elems != cur_elems) rc = MFO_ERR_UNSUPPORTED; else memcpy(b, a, cur_elems);
                break;
            default: rc = MFO_ERR_UNSUPPORTED;
        }
        if (rc != MFO_OK) break;
        cur_elems = L->out_elems;
        if (layer_outs && layer_outs[i]) memcpy(layer_outs[i], b, L->out_elems);
        uint8_t *t = a; a = b; b = t;
    }
    if (rc == MFO_OK) { if (cur_elems != m->out_elems) rc = MFO_ERR_UNSUPPORTED; else memcpy(final_q, a, m->out_elems); }
    free(a); free(b);
    return rc;
}

/* predict_quantized (lib.rs:193-196): out_f32 = dequantize(final); out_q (optional) = final quantized bytes */
MFO_API int mfo_predict_quantized(const mfo_model *m, const uint8_t *in_q, float *out_f32, uint8_t *out_q, uint8_t **layer_outs) {
    uint8_t *fq = (uint8_t *)malloc(m->out_elems ? m->out_elems : 1);
    int rc = run_layers(m, in_q, fq, layer_outs);
    if (rc == MFO_OK) {
        for (size_t k = 0; k < m->out_elems; ++k) {
            if (out_f32) out_f32[k] = mfo_dequantize(ld_elem(fq + k, m->is_u8_out), m->out_scale, m->out_zp); /* tensor.rs:89-92 */
            if (out_q) out_q[k] = fq[k];
        }
    }
    free(fq);
    return rc;
}
/* predict (lib.rs:188-191): quantize f32 input (tensor.rs:80-86 / :246-256), run, dequantize */
MFO_API int mfo_predict(const mfo_model *m, const float *in_f32, float *out_f32, uint8_t *out_q) {
    uint8_t *q = (uint8_t *)malloc(m->in_elems ? m->in_elems : 1);
    for (size_t k = 0; k < m->in_elems; ++k) q[k] = (uint8_t)mfo_quantize(in_f32[k], m->in_scale, m->in_zp, m->is_u8_in);
    int rc = mfo_predict_quantized(m, q, out_f32, out_q, NULL);
    free(q);
    return rc;
}
/* n independent samples (the reference handles exactly one per call; this is a loop of calls), optionally
 * across host threads for the `--impl reference` CPU baseline (each thread takes the next unprocessed sample). */
typedef struct {
    const mfo_model *m; const uint8_t *in_q; size_t n; float *out_f32; uint8_t *out_q;
    volatile long long next; volatile int rc; pthread_mutex_t mu;
} mfo_many_job;
static void *many_worker(void *arg) {
    mfo_many_job *J = (mfo_many_job *)arg;
    for (;;) {
        long long s = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (s >= (long long)J->n) break;
        int rc = mfo_predict_quantized(J->m, J->in_q + (size_t)s * J->m->in_elems, J->out_f32 ? J->out_f32 + (size_t)s * J->m->out_elems : NULL,
                                       J->out_q ? J->out_q + (size_t)s * J->m->out_elems : NULL, NULL);
        if (rc != MFO_OK) { pthread_mutex_lock(&J->mu); J->rc = rc; pthread_mutex_unlock(&J->mu); }
    }
    return NULL;
}
MFO_API int mfo_predict_many_quantized(const mfo_model *m, const uint8_t *in_q, size_t n, float *out_f32, uint8_t *out_q, int threads) {
    mfo_many_job J = {m, in_q, n, out_f32, out_q, 0, MFO_OK, PTHREAD_MUTEX_INITIALIZER};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    pthread_t th[256];
    int started = 0;
    for (int t = 1; t < threads; ++t) if (pthread_create(&th[started], NULL, many_worker, &J) == 0) started++;
    many_worker(&J);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    return J.rc;
}
MFO_API int mfo_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
