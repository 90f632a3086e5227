// mf_loader.cpp -- TFLite flatbuffer walk + per-op pre-processing (host side of the product).
//
// Mirrors, at run time, what microflow-macros does at Rust compile time:
//   graph walk            microflow-macros/src/lib.rs:62-183
//   tensor decoding       microflow-macros/src/tensor.rs:64-114 (2-D), :148-200 (4-D)
//   conv   preprocess     microflow-macros/src/ops/conv_2d.rs:94-114
//   dwconv preprocess     microflow-macros/src/ops/depthwise_conv_2d.rs:100-120
//   fc     preprocess     microflow-macros/src/ops/fully_connected.rs:100-123
//   pool   preprocess     microflow-macros/src/ops/average_pool_2d.rs:77-83
// Only the ~12 table fields of tflite.fbs that the macro reads are decoded.
#include "mf_loader.h"

#include <cmath>
#include <cstring>

namespace mf {
namespace {

// ---------------------------------------------------------------------------------------------
// Minimal FlatBuffers reader (little-endian host assumed, like the reference's target boards).
// ---------------------------------------------------------------------------------------------
class Fb {
  public:
    Fb(const uint8_t *p, size_t n) : p_(p), n_(n) {}
    bool in(size_t off, size_t len) const { return off <= n_ && len <= n_ - off; }
    template <class T> T rd(size_t off) const {
        T v{};
        if (in(off, sizeof(T))) std::memcpy(&v, p_ + off, sizeof(T));
        else bad_ = true;
        return v;
    }
    const uint8_t *ptr(size_t off) const { return p_ + off; }
    bool bad() const { return bad_; }

    struct Table {
        const Fb *fb = nullptr;
        size_t pos = 0;
        explicit operator bool() const { return fb && pos; }
        size_t field(int id) const {  // absolute offset of field `id`, 0 if absent
            if (!*this) return 0;
            int32_t so = fb->rd<int32_t>(pos);
            int64_t vt = (int64_t)pos - so;
            if (vt < 0 || !fb->in((size_t)vt, 4)) return 0;
            uint16_t vsz = fb->rd<uint16_t>((size_t)vt);
            size_t slot = 4 + 2 * (size_t)id;
            if (slot + 2 > vsz) return 0;
            uint16_t o = fb->rd<uint16_t>((size_t)vt + slot);
            return o ? pos + o : 0;
        }
        template <class T> T scalar(int id, T def) const {
            size_t f = field(id);
            return f ? fb->rd<T>(f) : def;
        }
        Table table(int id) const {
            size_t f = field(id);
            return f ? Table{fb, f + fb->rd<uint32_t>(f)} : Table{};
        }
    };
    struct Vec {
        const Fb *fb = nullptr;
        size_t pos = 0;  // first element
        uint32_t len = 0;
        template <class T> T at(uint32_t i) const { return fb->rd<T>(pos + sizeof(T) * (size_t)i); }
        Table table_at(uint32_t i) const {
            size_t e = pos + 4 * (size_t)i;
            return Table{fb, e + fb->rd<uint32_t>(e)};
        }
    };
    Vec vec(const Table &t, int id) const {
        size_t f = t.field(id);
        if (!f) return {};
        size_t v = f + rd<uint32_t>(f);
        if (!in(v, 4)) return {};
        return Vec{this, v + 4, rd<uint32_t>(v)};
    }
    Table root() const { return Table{this, (size_t)rd<uint32_t>(0)}; }

  private:
    const uint8_t *p_;
    size_t n_;
    mutable bool bad_ = false;
};

struct TensorMeta {
    std::vector<int> shape;
    int type = 0;
    std::vector<float> scale;
    std::vector<int64_t> zp;
    const uint8_t *data = nullptr;
    size_t data_len = 0;
};

bool read_tensor(const Fb &fb, const Fb::Vec &tensors, const Fb::Vec &buffers, int32_t idx, TensorMeta &t) {
    if (idx < 0 || (uint32_t)idx >= tensors.len) return false;
    Fb::Table tt = tensors.table_at((uint32_t)idx);
    Fb::Vec sh = fb.vec(tt, 0);
    t.shape.clear();
    for (uint32_t i = 0; i < sh.len; ++i) t.shape.push_back(sh.at<int32_t>(i));
    t.type = tt.scalar<int8_t>(1, 0);
    Fb::Table q = tt.table(4);
    t.scale.clear();
    t.zp.clear();
    if (q) {
        Fb::Vec s = fb.vec(q, 2), z = fb.vec(q, 3);
        for (uint32_t i = 0; i < s.len; ++i) t.scale.push_back(s.at<float>(i));
        for (uint32_t i = 0; i < z.len; ++i) t.zp.push_back(z.at<int64_t>(i));
    }
    uint32_t b = tt.scalar<uint32_t>(2, 0);
    t.data = nullptr;
    t.data_len = 0;
    if (b < buffers.len) {
        Fb::Vec d = fb.vec(buffers.table_at(b), 0);
        if (d.pos && fb.in(d.pos, d.len)) {
            t.data = fb.ptr(d.pos);
            t.data_len = d.len;
        }
    }
    return !fb.bad();
}

// i64 -> T by truncation (microflow-macros/src/tensor.rs:81-88 `to_subset_unchecked`)
inline int zp_cast(int64_t z, bool is_u8) { return is_u8 ? (int)(uint8_t)z : (int)(int8_t)z; }
inline int elem(uint8_t b, bool is_u8) { return is_u8 ? (int)b : (int)(int8_t)b; }
// element count of a shape; dimensions are untrusted: a negative one counts as 0 and anything beyond kMaxElems (2^31, far above
// every tensor the kernels can index with 32-bit offsets) saturates to kTooMany, which every caller rejects
constexpr size_t kMaxElems = (size_t)1 << 31, kTooMany = ~(size_t)0;
inline size_t prod(const std::vector<int> &s) {
    size_t p = 1;
    for (int d : s) {
        if (d <= 0) return 0;
        if (p > kMaxElems / (size_t)d) return kTooMany;
        p *= (size_t)d;
    }
    return p;
}
inline bool finite_all(const std::vector<float> &v) {
    for (float f : v)
        if (!std::isfinite(f)) return false;
    return true;
}

// Rust `x as T`: saturating float -> int, NaN -> 0
inline int sat_cast(float x, bool is_u8) {
    if (x != x) return 0;
    if (is_u8) return x <= 0.f ? 0 : (x >= 255.f ? 255 : (int)x);
    return x <= -128.f ? -128 : (x >= 127.f ? 127 : (int)x);
}

float scalbn_f(float x, int n) {  // Rust libm scalbnf
    float y = x;
    if (n > 127) {
        y *= 0x1p127f; n -= 127;
        if (n > 127) { y *= 0x1p127f; n -= 127; if (n > 127) n = 127; }
    } else if (n < -126) {
        y *= 0x1p-126f * 0x1p24f; n += 126 - 24;
        if (n < -126) { y *= 0x1p-126f * 0x1p24f; n += 126 - 24; if (n < -126) n = -126; }
    }
    uint32_t bits = (uint32_t)(0x7f + n) << 23;
    float s;
    std::memcpy(&s, &bits, 4);
    return y * s;
}

}  // namespace

// Rust libm 0.2 `expf` = FreeBSD/musl e_expf.c; the crate is not vendored in the reference
// (Cargo.toml:27), so this is written from the published algorithm.  Used only to build the 256-entry
// softmax table at load time (call sites: src/ops/softmax.rs:21, src/activation.rs:45).
float libm_expf(float x) {
    const float ln2hi = 6.9314575195e-1f, ln2lo = 1.4286067653e-6f, invln2 = 1.4426950216e+0f;
    const float P1 = 1.6666625440e-1f, P2 = -2.7667332906e-3f;
    uint32_t hx;
    std::memcpy(&hx, &x, 4);
    int sign = (int)(hx >> 31);
    hx &= 0x7fffffffu;
    if (hx >= 0x42aeac50u) {
        if (hx > 0x7f800000u) return x;
        if (hx >= 0x42b17218u && !sign) return x * 0x1p127f;
        if (sign && hx >= 0x42cff1b5u) return 0.f;
    }
    float hi, lo;
    int k;
    if (hx > 0x3eb17218u) {
        if (hx > 0x3f851592u) k = (int)(invln2 * x + (sign ? -0.5f : 0.5f));
        else k = 1 - sign - sign;
        hi = x - (float)k * ln2hi;
        lo = (float)k * ln2lo;
        x = hi - lo;
    } else if (hx > 0x39000000u) {
        k = 0; hi = x; lo = 0.f;
    } else {
        return 1.f + x;
    }
    float xx = x * x;
    float c = x - xx * (P1 + xx * P2);
    float y = 1.f + (x * c / (2.f - c) - lo + hi);
    return k == 0 ? y : scalbn_f(y, k);
}

int features_from_bmp_gray8(const uint8_t *b, size_t len, uint8_t *out, size_t cap, int *height, int *width, std::string &err) {
    auto u16 = [&](size_t o) { return (uint32_t)b[o] | ((uint32_t)b[o + 1] << 8); };
    auto u32 = [&](size_t o) { return u16(o) | (u16(o + 2) << 16); };
    if (!b || len < 54 || b[0] != 'B' || b[1] != 'M') { err = "not a BMP file"; return MF_ERR_INVALID_ARG; }
    const uint32_t off = u32(10), hsz = u32(14);
    const int32_t w = (int32_t)u32(18), hraw = (int32_t)u32(22);
    const uint32_t planes = u16(26), bpp = u16(28), comp = u32(30);
    if (hsz < 40 || planes != 1 || bpp != 8 || comp != 0) { err = "only uncompressed 8-bit BMP images are supported"; return MF_ERR_UNSUPPORTED_TYPE; }
    const int32_t h = hraw < 0 ? -hraw : hraw;
    if (w <= 0 || h <= 0 || w > 16384 || h > 16384) { err = "bad BMP dimensions"; return MF_ERR_INVALID_ARG; }
    const size_t pitch = ((size_t)w + 3) & ~(size_t)3;                      // rows are padded to 4 bytes
    if ((size_t)off + pitch * (size_t)h > len) { err = "truncated BMP pixel array"; return MF_ERR_INVALID_ARG; }
    uint32_t ncol = u32(46);
    if (ncol == 0) ncol = 256;
    if (14 + (size_t)hsz + 4 * (size_t)ncol > off || ncol > 256) { err = "bad BMP palette"; return MF_ERR_INVALID_ARG; }
    for (uint32_t k = 0; k < ncol; ++k) {                                   // gray identity palette: index == intensity
        const uint8_t *e = b + 14 + hsz + 4 * (size_t)k;
        if (e[0] != k || e[1] != k || e[2] != k) { err = "BMP palette is not the identity gray ramp"; return MF_ERR_UNSUPPORTED_TYPE; }
    }
    if (height) *height = h;
    if (width) *width = w;
    if ((size_t)w * (size_t)h > cap) { err = "output buffer too small for the image"; return MF_ERR_INVALID_ARG; }
    for (int32_t r = 0; r < h; ++r) {
        const int32_t src = hraw < 0 ? r : h - 1 - r;                        // bottom-up storage -> top row first
        std::memcpy(out + (size_t)r * w, b + off + pitch * (size_t)src, (size_t)w);
    }
    return MF_OK;
}

int quantize_scalar(float x, float scale, int zp, bool is_u8) {  // src/quantize.rs:16-18
    return sat_cast(roundf(x / scale + (float)zp), is_u8);
}

// saturating cast followed by relu / relu6 on the quantized value == one clamp [lo, hi]
// (max first, then min: src/activation.rs:21-23, :32-34)
void activation_clamp(int act, float out_scale, int out_zp, bool is_u8, int &lo, int &hi) {
    lo = is_u8 ? 0 : -128;
    hi = is_u8 ? 255 : 127;
    if (act == MF_ACT_RELU || act == MF_ACT_RELU6) lo = out_zp > lo ? out_zp : lo;
    if (act == MF_ACT_RELU6) {
        int six = quantize_scalar(6.0f, out_scale, out_zp, is_u8);
        hi = six < hi ? six : hi;
    }
}

int parse_tflite(const uint8_t *buf, size_t len, ModelSpec &M, std::string &err) {
    if (!buf || len < 16) { err = "invalid model, please provide a valid TensorFlow Lite model"; return MF_ERR_INVALID_MODEL; }
    Fb fb(buf, len);
    Fb::Table model = fb.root();
    Fb::Vec codes = fb.vec(model, 1), subgraphs = fb.vec(model, 2), buffers = fb.vec(model, 4);
    if (!subgraphs.len || !codes.len || fb.bad()) { err = "invalid model, please provide a valid TensorFlow Lite model"; return MF_ERR_INVALID_MODEL; }
    Fb::Table sg = subgraphs.table_at(0);  // lib.rs:62: subgraph 0 only
    Fb::Vec tensors = fb.vec(sg, 0), g_in = fb.vec(sg, 1), g_out = fb.vec(sg, 2), ops = fb.vec(sg, 3);
    if (!tensors.len || !g_in.len || !g_out.len) { err = "invalid model: empty subgraph"; return MF_ERR_INVALID_MODEL; }

    auto fail = [&](int code, const std::string &msg) { err = msg; return code; };
    auto type_name = [](int t) { return std::string("TensorType(") + std::to_string(t) + ")"; };

    // ---- model input: first input of subgraph 0 (lib.rs:66-128) ---------------------------------
    TensorMeta ti;
    if (!read_tensor(fb, tensors, buffers, g_in.at<int32_t>(0), ti)) return fail(MF_ERR_INVALID_MODEL, "invalid model: bad input tensor");
    if (ti.type != MF_DTYPE_I8 && ti.type != MF_DTYPE_U8)
        return fail(MF_ERR_UNSUPPORTED_TYPE, "unsupported input tensor type: " + type_name(ti.type) + ". Supported input types are INT8 and UINT8");
    std::vector<int> ishape = ti.shape;
    if (ishape.size() == 1) ishape.insert(ishape.begin(), 1);  // lib.rs:68-70
    if (ishape.size() != 2 && ishape.size() != 4)
        return fail(MF_ERR_UNSUPPORTED_RANK, "unsupported input tensor rank: " + std::to_string(ishape.size()) + ". Supported ranks are 2 and 4");
    if (ti.scale.empty() || ti.zp.empty()) return fail(MF_ERR_INVALID_MODEL, "invalid model: input tensor has no quantization");
    M = ModelSpec{};
    M.is_u8_in = ti.type == MF_DTYPE_U8;
    M.in_rank = (int)ishape.size();
    for (int i = 0; i < M.in_rank; ++i) M.in_dims[i] = ishape[i];
    M.in_scale = ti.scale[0];
    M.in_zp = zp_cast(ti.zp[0], M.is_u8_in);
    M.in_elems = prod(ishape);
    if (M.in_elems == kTooMany) return fail(MF_ERR_UNSUPPORTED_SHAPE, "input tensor has more than 2^31 elements");
    M.max_elems = M.in_elems;

    // ---- operators, in execution order; each consumes the previous op's output (lib.rs:130-151, :198-201)
    size_t cur_elems = M.in_elems;
    for (uint32_t oi = 0; oi < ops.len; ++oi) {
        Fb::Table op = ops.table_at(oi);
        uint32_t ci = op.scalar<uint32_t>(0, 0);
        if (ci >= codes.len) return fail(MF_ERR_INVALID_MODEL, "invalid model: opcode index out of range");
        int code = codes.table_at(ci).scalar<int8_t>(0, 0);  // deprecated_builtin_code (lib.rs:131-137)
        Fb::Vec oin = fb.vec(op, 1), oout = fb.vec(op, 2);
        Fb::Table opt = op.table(4);
        if (!oin.len || !oout.len) return fail(MF_ERR_INVALID_MODEL, "invalid model: operator without inputs/outputs");
        TensorMeta tin, tout;
        if (!read_tensor(fb, tensors, buffers, oin.at<int32_t>(0), tin) || !read_tensor(fb, tensors, buffers, oout.at<int32_t>(0), tout))
            return fail(MF_ERR_INVALID_MODEL, "invalid model: bad operator tensor index");

        LayerSpec L;
        L.op = code;
        if (code != MF_OP_CONV_2D && code != MF_OP_DEPTHWISE_CONV_2D && code != MF_OP_FULLY_CONNECTED && code != MF_OP_AVERAGE_POOL_2D &&
            code != MF_OP_SOFTMAX && code != MF_OP_RESHAPE)
            return fail(MF_ERR_UNSUPPORTED_OP, "unsupported operator: BuiltinOperator(" + std::to_string(code) + ")");
        if (code != MF_OP_RESHAPE && tin.type != MF_DTYPE_I8 && tin.type != MF_DTYPE_U8)
            return fail(MF_ERR_UNSUPPORTED_TYPE, "operator " + std::to_string(oi) + " supports only INT8/UINT8 input tensors, got " + type_name(tin.type));
        L.is_u8 = tin.type == MF_DTYPE_U8;
        std::vector<int> oshape = tout.shape;
        if (oshape.size() == 1) oshape.insert(oshape.begin(), 1);
        if (tin.shape.size() > 4 || oshape.size() > 4) return fail(MF_ERR_UNSUPPORTED_RANK, "unsupported tensor rank > 4");
        L.in_rank = (int)tin.shape.size();
        L.out_rank = (int)oshape.size();
        for (int i = 0; i < L.in_rank; ++i) L.in_dims[i] = tin.shape[i];
        for (int i = 0; i < L.out_rank; ++i) L.out_dims[i] = oshape[i];
        L.in_elems = prod(tin.shape);
        L.out_elems = prod(oshape);
        if (L.in_elems == kTooMany || L.out_elems == kTooMany)
            return fail(MF_ERR_UNSUPPORTED_SHAPE, "operator " + std::to_string(oi) + ": tensor with more than 2^31 elements");
        if (!tin.scale.empty()) L.in_scale = tin.scale[0];
        if (!tin.zp.empty()) L.in_zp = zp_cast(tin.zp[0], L.is_u8);
        if (!tout.scale.empty()) L.out_scale = tout.scale[0];
        if (!tout.zp.empty()) L.out_zp = zp_cast(tout.zp[0], L.is_u8);
        if (L.in_elems != cur_elems)
            return fail(MF_ERR_UNSUPPORTED_SHAPE, "operator " + std::to_string(oi) + ": input has " + std::to_string(L.in_elems) +
                                                      " elements but the previous operator produced " + std::to_string(cur_elems));
        const bool needs_q = code != MF_OP_RESHAPE;
        if (needs_q && (tout.scale.empty() || tout.zp.empty() || (code != MF_OP_SOFTMAX && (tin.scale.empty() || tin.zp.empty()))))
            return fail(MF_ERR_INVALID_MODEL, "invalid model: operator tensor without quantization parameters");

        if (code == MF_OP_CONV_2D || code == MF_OP_DEPTHWISE_CONV_2D) {
            const bool dw = code == MF_OP_DEPTHWISE_CONV_2D;
            if (oin.len < 3) return fail(MF_ERR_INVALID_MODEL, "invalid model: convolution without filter/bias inputs");
            if (tin.shape.size() != 4 || tout.shape.size() != 4) return fail(MF_ERR_UNSUPPORTED_RANK, "convolution needs 4-D input and output tensors");
            TensorMeta tw, tb;
            if (!read_tensor(fb, tensors, buffers, oin.at<int32_t>(1), tw) || !read_tensor(fb, tensors, buffers, oin.at<int32_t>(2), tb))
                return fail(MF_ERR_INVALID_MODEL, "invalid model: bad filter/bias tensor");
            if (tw.shape.size() != 4 || !tw.data || !tb.data || tw.scale.empty() || tw.zp.empty() || tb.scale.empty() || tb.zp.empty())
                return fail(MF_ERR_INVALID_MODEL, "invalid model: convolution filter/bias without data or quantization");
            // the reference's ops take Tensor4D<T, 1, ...> (src/ops/conv_2d.rs:40: BATCHES fixed to 1) and depthwise weights of shape
            // [1, KH, KW, C]: anything else would not type-check there, and every kernel here strides samples by H*W*C
            if (tin.shape[0] != 1 || tout.shape[0] != 1 || (dw && tw.shape[0] != 1))
                return fail(MF_ERR_UNSUPPORTED_SHAPE, "convolution " + std::to_string(oi) + ": tensor batch dimension must be 1");
            L.H = tin.shape[1]; L.W = tin.shape[2]; L.Cin = tin.shape[3];
            L.OH = tout.shape[1]; L.OW = tout.shape[2];
            L.KH = tw.shape[1]; L.KW = tw.shape[2];
            L.Cout = dw ? tw.shape[3] : tw.shape[0];
            if (L.Cout != tout.shape[3] || (!dw && tw.shape[3] != L.Cin) || prod(tw.shape) == kTooMany || tw.data_len < prod(tw.shape) ||
                tb.data_len < (size_t)L.Cout * 4 || L.Cout <= 0 || L.KH <= 0 || L.KW <= 0)
                return fail(MF_ERR_UNSUPPORTED_SHAPE, "convolution " + std::to_string(oi) + ": inconsistent filter/bias/output shapes");
            L.w.assign(tw.data, tw.data + prod(tw.shape));
            for (int64_t z : tw.zp) L.w_zp.push_back(zp_cast(z, L.is_u8));
            // options (tflite.fbs:562 / :592)
            L.pad = opt.scalar<int8_t>(0, 0);
            L.sw = opt.scalar<int32_t>(1, 0);
            L.sh = opt.scalar<int32_t>(2, 0);     // strides = (stride_h, stride_w)  (ops/conv_2d.rs:80)
            L.act = opt.scalar<int8_t>(dw ? 4 : 3, 0);
            // preprocess (ops/conv_2d.rs:101-112 / depthwise_conv_2d.rs:107-118); f32, written operation order
            std::vector<int32_t> bias((size_t)L.Cout);
            std::memcpy(bias.data(), tb.data, (size_t)L.Cout * 4);
            L.c0.resize((size_t)L.Cout);
            for (int b = 0; b < L.Cout; ++b) {
                float bs = tb.scale[(size_t)b < tb.scale.size() ? (size_t)b : 0];
                int32_t bz = (int32_t)tb.zp[(size_t)b < tb.zp.size() ? (size_t)b : 0];
                L.c0[(size_t)b] = bs / L.out_scale * (float)(bias[(size_t)b] - bz);
            }
            L.c1.resize(tw.scale.size());
            for (size_t b = 0; b < tw.scale.size(); ++b) L.c1[b] = L.in_scale * tw.scale[b] / L.out_scale;
            L.macs = (uint64_t)L.OH * L.OW * L.Cout * L.KH * L.KW * (dw ? 1 : L.Cin);
        } else if (code == MF_OP_FULLY_CONNECTED) {
            if (oin.len < 3) return fail(MF_ERR_INVALID_MODEL, "invalid model: fully_connected without weights/bias inputs");
            TensorMeta tw, tb;
            if (!read_tensor(fb, tensors, buffers, oin.at<int32_t>(1), tw) || !read_tensor(fb, tensors, buffers, oin.at<int32_t>(2), tb))
                return fail(MF_ERR_INVALID_MODEL, "invalid model: bad weights/bias tensor");
            if (tw.shape.size() != 2 || !tw.data || !tb.data || tw.scale.empty() || tw.zp.empty() || tb.scale.empty() || tb.zp.empty())
                return fail(MF_ERR_INVALID_MODEL, "invalid model: fully_connected weights/bias without data or quantization");
            const int N = tw.shape[0], K = tw.shape[1];
            if ((size_t)K != L.in_elems || (size_t)N != L.out_elems || tw.data_len < (size_t)N * K || tb.data_len < (size_t)N * 4)
                return fail(MF_ERR_UNSUPPORTED_SHAPE, "fully_connected " + std::to_string(oi) + ": inconsistent weights/bias/output shapes");
            L.Cin = K; L.Cout = N;
            L.w.assign(tw.data, tw.data + (size_t)N * K);
            L.w_zp.push_back(zp_cast(tw.zp[0], L.is_u8));
            L.act = opt.scalar<int8_t>(0, 0);
            // preprocess (ops/fully_connected.rs:107-121).  `input.shape[1]` is the macro's view of the *unflattened*
            // input tensor shape (1-D -> [1,n], tensor.rs:67-70) -- replicated as is.
            std::vector<int32_t> bias((size_t)N);
            std::memcpy(bias.data(), tb.data, (size_t)N * 4);
            const int32_t bz = (int32_t)tb.zp[0];
            const float ratio = tb.scale[0] / L.out_scale;
            L.c0.resize((size_t)N);
            for (int j = 0; j < N; ++j) L.c0[(size_t)j] = ratio * (float)(bias[(size_t)j] + (-bz));
            L.c1.assign(1, L.in_scale * tw.scale[0] / L.out_scale);
            L.c2.resize((size_t)N);
            for (int j = 0; j < N; ++j) {
                int32_t s = 0;
                for (int k = 0; k < K; ++k) s += elem(L.w[(size_t)j * K + k], L.is_u8);
                L.c2[(size_t)j] = s * L.in_zp;
            }
            std::vector<int> s2 = tin.shape;
            if (s2.size() == 1) s2.insert(s2.begin(), 1);
            L.c3 = (s2.size() > 1 ? s2[1] : 1) * L.in_zp * L.w_zp[0];
            L.macs = (uint64_t)N * K;
        } else if (code == MF_OP_AVERAGE_POOL_2D) {
            if (tin.shape.size() != 4 || tout.shape.size() != 4) return fail(MF_ERR_UNSUPPORTED_RANK, "average_pool_2d needs 4-D input and output tensors");
            if (tin.shape[0] != 1 || tout.shape[0] != 1) return fail(MF_ERR_UNSUPPORTED_SHAPE, "average_pool_2d: tensor batch dimension must be 1");
            L.H = tin.shape[1]; L.W = tin.shape[2]; L.Cin = L.Cout = tin.shape[3];
            L.OH = tout.shape[1]; L.OW = tout.shape[2];
            if (tout.shape[3] != L.Cin) return fail(MF_ERR_UNSUPPORTED_SHAPE, "average_pool_2d: channel mismatch");
            L.pad = opt.scalar<int8_t>(0, 0);
            L.sw = opt.scalar<int32_t>(1, 0);
            L.sh = opt.scalar<int32_t>(2, 0);
            L.KW = opt.scalar<int32_t>(3, 0);
            L.KH = opt.scalar<int32_t>(4, 0);
            L.act = opt.scalar<int8_t>(5, 0);
            if (L.KH <= 0 || L.KW <= 0) return fail(MF_ERR_UNSUPPORTED_SHAPE, "average_pool_2d: empty filter");
            // preprocess (ops/average_pool_2d.rs:78-82)
            L.c0.assign(1, L.in_scale / L.out_scale);
            L.c1.assign(1, (float)L.out_zp - (L.in_scale * (float)L.in_zp) / L.out_scale);
        } else if (code == MF_OP_SOFTMAX) {
            // ops/softmax.rs: only the output tensor is read; `beta` is ignored; the input scale travels with the tensor.
            if (tin.scale.empty()) return fail(MF_ERR_INVALID_MODEL, "invalid model: softmax input without scale");
            if (L.out_elems != L.in_elems) return fail(MF_ERR_UNSUPPORTED_SHAPE, "softmax: shape mismatch");
            L.Cin = L.Cout = (int)L.in_elems;
            L.exp_lut.resize(256);
            for (int b = 0; b < 256; ++b) L.exp_lut[(size_t)b] = libm_expf((float)elem((uint8_t)b, L.is_u8) * L.in_scale);  // softmax.rs:20-21
        } else {  // RESHAPE (ops/reshape.rs:33-55): only the output tensor shape is used
            if (tout.shape.size() != 2 && tout.shape.size() != 4)
                return fail(MF_ERR_UNSUPPORTED_SHAPE, "Reshape supports only output tensor ranks 2 and 4, got rank " + std::to_string(tout.shape.size()));
            if (L.out_elems != L.in_elems) return fail(MF_ERR_UNSUPPORTED_SHAPE, "reshape: element count mismatch");
            L.is_u8 = M.layers.empty() ? M.is_u8_in : M.layers.back().is_u8;
        }
        if (L.act != MF_ACT_NONE && L.act != MF_ACT_RELU && L.act != MF_ACT_RELU6)
            return fail(MF_ERR_UNSUPPORTED_ACTIVATION, "unsupported fused activation: " + std::to_string(L.act) + ". Supported activations are NONE, RELU, and RELU6");
        if (L.pad != MF_PAD_SAME && L.pad != MF_PAD_VALID) return fail(MF_ERR_INVALID_MODEL, "invalid model: unknown padding");
        if ((code == MF_OP_CONV_2D || code == MF_OP_DEPTHWISE_CONV_2D || code == MF_OP_AVERAGE_POOL_2D)) {
            if (L.sh <= 0 || L.sw <= 0) return fail(MF_ERR_UNSUPPORTED_SHAPE, "operator " + std::to_string(oi) + ": non-positive stride");
            if (L.pad == MF_PAD_VALID && (L.sh * (L.OH - 1) + L.KH > L.H || L.sw * (L.OW - 1) + L.KW > L.W))
                return fail(MF_ERR_VIEW_OUT_OF_BOUNDS, "operator " + std::to_string(oi) + ": VALID view indexes outside the input (src/tensor.rs:222 would panic)");
        }
        if (!finite_all(L.c0) || !finite_all(L.c1))
            return fail(MF_ERR_NONFINITE_CONSTANT, "operator " + std::to_string(oi) + ": non-finite requantization constant (zero scale?)");
        activation_clamp(L.act, L.out_scale, L.out_zp, L.is_u8, L.act_lo, L.act_hi);
        cur_elems = L.out_elems;
        if (L.out_elems > M.max_elems) M.max_elems = L.out_elems;
        M.layers.push_back(std::move(L));
        if (fb.bad()) return fail(MF_ERR_INVALID_MODEL, "invalid model: truncated flatbuffer");
    }

    // ---- model output: first output of subgraph 0 (lib.rs:153-183) --------------------------------
    TensorMeta to;
    if (!read_tensor(fb, tensors, buffers, g_out.at<int32_t>(0), to)) return fail(MF_ERR_INVALID_MODEL, "invalid model: bad output tensor");
    if (to.type != MF_DTYPE_I8 && to.type != MF_DTYPE_U8)
        return fail(MF_ERR_UNSUPPORTED_TYPE, "unsupported output tensor type: " + type_name(to.type) + ". Supported output types are INT8 and UINT8");
    std::vector<int> oshape = to.shape;
    if (oshape.size() == 1) oshape.insert(oshape.begin(), 1);
    if (oshape.size() != 2 && oshape.size() != 4)
        return fail(MF_ERR_UNSUPPORTED_RANK, "unsupported output tensor rank: " + std::to_string(oshape.size()) + ". Supported ranks are 2 and 4");
    if (to.scale.empty() || to.zp.empty()) return fail(MF_ERR_INVALID_MODEL, "invalid model: output tensor has no quantization");
    M.is_u8_out = to.type == MF_DTYPE_U8;
    M.out_rank = (int)oshape.size();
    for (int i = 0; i < M.out_rank; ++i) M.out_dims[i] = oshape[i];
    M.out_scale = to.scale[0];
    M.out_zp = zp_cast(to.zp[0], M.is_u8_out);
    M.out_elems = prod(oshape);
    if (M.out_elems == kTooMany) return fail(MF_ERR_UNSUPPORTED_SHAPE, "output tensor has more than 2^31 elements");
    if (M.out_elems != cur_elems) return fail(MF_ERR_UNSUPPORTED_SHAPE, "model output shape does not match the last operator");
    if (M.in_elems == 0 || M.out_elems == 0) return fail(MF_ERR_UNSUPPORTED_SHAPE, "empty input or output tensor");
    return MF_OK;
}

}  // namespace mf
