// mf_engine.cu -- kernel selection and launch for one layer.
#include "mf_engine.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace mf {

const char *kernel_name(Kernel k) {
    switch (k) {
        case Kernel::None: return "none(reshape)";
        case Kernel::ConvGeneric: return "conv_generic_kernel";
        case Kernel::ConvTcPointwise: return "conv_tc_kernel(pointwise)";
        case Kernel::ConvTc3x3: return "conv_tc_kernel(3x3)";
        case Kernel::PwConvDp4a: return "pwconv_dp4a_kernel";
        case Kernel::DwConvC4: return "dwconv_c4_kernel";
        case Kernel::DwConv3x3Rows: return "dwconv3x3_rows_kernel";
        case Kernel::DwConvCin1: return "dwconv_cin1_kernel";
        case Kernel::FcGeneric: return "fc_generic_kernel";
        case Kernel::FcWarp: return "fc_warp_kernel";
        case Kernel::FcTc: return "conv_tc_kernel(fc)";
        case Kernel::PoolGeneric: return "pool_generic_kernel";
        case Kernel::Softmax: return "softmax_kernel";
    }
    return "?";
}

size_t BlobBuilder::add(const void *p, size_t bytes) {
    size_t off = (host_.size() + 255) & ~(size_t)255;
    host_.resize(off + bytes);
    if (bytes) std::memcpy(host_.data() + off, p, bytes);
    return off;
}

namespace {
inline int elem_i(uint8_t b, bool is_u8) { return is_u8 ? (int)b : (int)(int8_t)b; }
template <class T> std::vector<T> expand(const std::vector<T> &v, int n) {  // per-tensor -> per-channel (.get(b).unwrap_or([0]))
    std::vector<T> o((size_t)n);
    for (int i = 0; i < n; ++i) o[(size_t)i] = v[(size_t)i < v.size() ? (size_t)i : 0];
    return o;
}
}  // namespace

void LayerExec::plan(BlobBuilder &bb, int impl, bool have_device) {
    const LayerSpec &L = spec;
    kernel = Kernel::None;
    const size_t io_bytes = L.in_elems + L.out_elems;
    alg_bytes = io_bytes;
    weight_bytes = 0;
    if (L.op == MF_OP_RESHAPE) { alg_bytes = 0; return; }

    if (L.op == MF_OP_CONV_2D || L.op == MF_OP_DEPTHWISE_CONV_2D) {
        const bool dw = L.op == MF_OP_DEPTHWISE_CONV_2D;
        const int Cout = L.Cout;
        std::vector<int32_t> wzp = expand(L.w_zp, Cout);
        std::vector<float> c1 = expand(L.c1, Cout);
        std::vector<float> c0z((size_t)Cout);
        for (int b = 0; b < Cout; ++b) c0z[(size_t)b] = (float)L.out_zp + L.c0[(size_t)b];  // conv_2d.rs:94-95, same f32 add
        // kcorr[co] = in_zp * (sum of ALL weights of filter co): used by the pad-with-zero-point fast kernels
        std::vector<int32_t> kcorr((size_t)Cout, 0);
        const int taps = L.KH * L.KW;
        for (int co = 0; co < Cout; ++co) {
            int32_t s = 0;
            if (dw) for (int t = 0; t < taps; ++t) s += elem_i(L.w[(size_t)t * Cout + co], L.is_u8);
            else for (size_t k = 0; k < (size_t)taps * L.Cin; ++k) s += elem_i(L.w[(size_t)co * taps * L.Cin + k], L.is_u8);
            kcorr[(size_t)co] = L.in_zp * s;
        }
        // can the corrected accumulator exceed 2^22 ?  It equals sum_valid (v - in_zp) * (w - w_zp) whatever the kernel does with
        // borders (zero-point padding, zero fill + border-class table, masked weights), and |v - in_zp| <= 255, so 255 * sum |w - w_zp|
        // bounds it rigorously.  Beyond 2^22 the kernels take the general exact int -> float conversion instead of the biased one.
        big_acc = false;
        for (int co = 0; co < Cout; ++co) {
            long long sa = 0;
            const int fz = wzp[(size_t)co];
            if (dw) for (int t = 0; t < taps; ++t) sa += std::abs(elem_i(L.w[(size_t)t * Cout + co], L.is_u8) - fz);
            else for (size_t k = 0; k < (size_t)taps * L.Cin; ++k) sa += std::abs(elem_i(L.w[(size_t)co * taps * L.Cin + k], L.is_u8) - fz);
            if (sa * 255 > (1ll << 22)) big_acc = true;
        }
        o_w = bb.add(L.w.data(), L.w.size());
        o_wzp = bb.add(wzp.data(), wzp.size() * 4);
        o_c0z = bb.add(c0z.data(), c0z.size() * 4);
        o_c1 = bb.add(c1.data(), c1.size() * 4);
        o_kcorr = bb.add(kcorr.data(), kcorr.size() * 4);
        weight_bytes = L.w.size() + (size_t)Cout * 8;
        alg_bytes += weight_bytes;

        const bool wz0 = std::all_of(L.w_zp.begin(), L.w_zp.end(), [](int32_t z) { return z == 0; });
        kernel = Kernel::ConvGeneric;
        if (impl == 1) { why_not_fast = "generic kernels forced"; return; }
        if (L.is_u8 || !wz0) {
            // `T = u8` (microflow-macros/src/ops/conv_2d.rs:39-46) and general weight zero-points (the view-sum term, src/ops/conv_2d.rs:74-76,
            // depthwise_conv_2d.rs:71-73): the SIMT fast kernels that implement the full formula, or -- uint8 with weight zero-points 0 --
            // the tcgen05 path below with unsigned operand formats
            if (dw) {
                if (!big_acc && L.Cin == L.Cout && L.Cout % 4 == 0) kernel = Kernel::DwConvC4;
                else why_not_fast = "uint8 / weight zero-point depthwise shape not covered by the fast kernel";
                return;
            }
            if (!wz0) {
                if (L.KH == 1 && L.KW == 1 && L.Cin % 4 == 0) kernel = Kernel::PwConvDp4a;
                else why_not_fast = "non-zero weight zero-point on a KxK convolution: generic kernel";
                return;
            }
        }
        if (dw) {
            if (big_acc) why_not_fast = "accumulator range beyond 2^22: generic kernel";
            else if (L.Cin == L.Cout && L.Cout % 4 == 0)
                kernel = (L.KH == 3 && L.KW == 3 && L.sh == L.sw && (L.sh == 1 || L.sh == 2)) ? Kernel::DwConv3x3Rows : Kernel::DwConvC4;
            else if (L.Cin == 1 && L.Cout % 4 == 0 && L.Cout >= 4 && L.Cout <= 16) kernel = Kernel::DwConvCin1;
            else why_not_fast = "depthwise shape not covered by a fast kernel";
            return;
        }
        // dense conv: try the tensor core first
        if (impl == 0 && have_device) {
            std::string why;
            if (L.KH == 1 && L.KW == 1 && L.sh == 1 && L.sw == 1 && L.OH == L.H && L.OW == L.W) {
                const int P = conv_tc_pick_pack(L.Cin, L.Cout);
                if (P > 0 && ((long long)L.OH * L.OW) % P == 0) {
                    tc = ConvTcPlan{};
                    tc_P = P;
                    tc.P = P; tc.N = P * L.Cout; tc.Cout = L.Cout; tc.C = P * L.Cin; tc.CB = tc.C / 128;
                    tc.KH = tc.KW = 1; tc.TW = 128; tc.TH = 1; tc.off_r = tc.off_c = 0; tc.ncls = 1;
                    tc.lo = (float)L.act_lo; tc.hi = (float)L.act_hi; tc.big_acc = big_acc; tc.is_u8 = L.is_u8;
                    std::vector<uint8_t> wm = conv_tc_pack_pointwise(L.w.data(), L.Cout, L.Cin, P);
                    std::vector<float> ez((size_t)tc.N), es((size_t)tc.N);
                    std::vector<int32_t> ec((size_t)tc.N);
                    for (int n = 0; n < tc.N; ++n) { ez[(size_t)n] = c0z[(size_t)(n % Cout)]; es[(size_t)n] = c1[(size_t)(n % Cout)]; ec[(size_t)n] = kcorr[(size_t)(n % Cout)]; }
                    o_tc_w = bb.add(wm.data(), wm.size());
                    tc.h_c0z = ez; tc.h_c1 = es; tc.h_corr = ec;
                    kernel = Kernel::ConvTcPointwise;
                    return;
                }
                why = "pointwise shape cannot be packed into 128-byte rows";
            } else if (L.KH == 3 && L.KW == 3 && L.sh == 1 && L.sw == 1 && L.pad == MF_PAD_SAME && L.OH == L.H && L.OW == L.W && L.Cin % 128 == 0 &&
                       L.Cout % 32 == 0 && L.Cout <= 256 && L.H >= 2 && L.W >= 2) {
                tc = ConvTcPlan{};
                tc.P = 1; tc.N = L.Cout; tc.Cout = L.Cout; tc.C = L.Cin; tc.CB = L.Cin / 128;
                tc.KH = tc.KW = 3; tc.off_r = tc.off_c = 1; tc.ncls = 9;
                // tile shape: minimise padded work, prefer 16 x 8
                long long best = -1;
                for (int tw : {16, 32, 8, 64, 128}) {
                    const int th = 128 / tw;
                    const long long work = (long long)((L.OW + tw - 1) / tw) * tw * ((L.OH + th - 1) / th) * th;
                    if (best < 0 || work < best) { best = work; tc.TW = tw; tc.TH = th; }
                }
                {   // single-patch staging (mf_conv_tc.h): 16 x 8 tiles, the input patch is fetched once per tile
                    static const bool env_patch = [] { const char *e = std::getenv("MF_TC_PATCH"); return !e || std::atoi(e) != 0; }();
                    if (env_patch) { tc.patch = true; tc.TW = 8; tc.TH = 16; }
                }
                tc.lo = (float)L.act_lo; tc.hi = (float)L.act_hi; tc.big_acc = big_acc; tc.is_u8 = L.is_u8;
                std::vector<int32_t> corr = conv_tc_border_corr_3x3(L.w.data(), L.Cout, L.Cin, L.in_zp, L.H, L.W, L.is_u8);
                o_tc_w = o_w;  // OHWI is already the [N][K_total] matrix
                tc.h_c0z = c0z; tc.h_c1 = c1; tc.h_corr = corr;
                kernel = Kernel::ConvTc3x3;
                return;
            } else {
                why = "not a 1x1/s1 or 3x3/s1/SAME/Cin%128 convolution";
            }
            why_not_fast = "tensor core: " + why;
        }
        if (L.KH == 1 && L.KW == 1 && L.Cin % 4 == 0) kernel = Kernel::PwConvDp4a;
        return;
    }
    if (L.op == MF_OP_FULLY_CONNECTED) {
        std::vector<float> c0z((size_t)L.Cout);
        for (int j = 0; j < L.Cout; ++j) c0z[(size_t)j] = (float)L.out_zp + L.c0[(size_t)j];
        o_w = bb.add(L.w.data(), L.w.size());
        o_c0z = bb.add(c0z.data(), c0z.size() * 4);
        o_c2 = bb.add(L.c2.data(), L.c2.size() * 4);
        weight_bytes = L.w.size() + (size_t)L.Cout * 8;
        alg_bytes += weight_bytes;
        kernel = Kernel::FcGeneric;
        if (impl == 1) return;
        if (L.Cin % 16 == 0 && L.Cout <= 8) { kernel = Kernel::FcWarp; return; }     // int8 or uint8, any weight zero-point (row-sum term)
        // FullyConnected as the same tcgen05 GEMM as the pointwise convs: row = one sample (K bytes, K % 128 == 0), B = W [N][K].
        // With w_zp == 0 the reference's accumulator x.W - w_zp*rowsum - c2[j] + c3 (fully_connected.rs:71) is acc - c2[j]
        // (c3 = K*in_zp*w_zp = 0), i.e. exactly the conv epilogue with kcorr = c2; c1 is per-tensor and replicated.
        if (impl == 0 && have_device && L.w_zp[0] == 0 && L.c3 == 0 && L.Cin % 128 == 0 && L.Cout % 32 == 0 && L.Cout <= 256) {
            long long big = 0;
            for (int j = 0; j < L.Cout; ++j) {
                long long sa = 0;
                for (int k = 0; k < L.Cin; ++k) sa += std::abs(elem_i(L.w[(size_t)j * L.Cin + k], L.is_u8));
                big = std::max(big, sa * 255);     // acc - c2[j] = sum (x - in_zp) * w  (w_zp == 0)
            }
            big_acc = big > (1ll << 22);
            tc = ConvTcPlan{};
            tc_P = 1;
            tc.P = 1; tc.N = L.Cout; tc.Cout = L.Cout; tc.C = L.Cin; tc.CB = L.Cin / 128;
            tc.KH = tc.KW = 1; tc.TW = 128; tc.TH = 1; tc.off_r = tc.off_c = 0; tc.ncls = 1;
            tc.lo = (float)L.act_lo; tc.hi = (float)L.act_hi; tc.big_acc = big_acc; tc.is_u8 = L.is_u8;
            tc.h_c0z = c0z;
            tc.h_c1.assign((size_t)L.Cout, L.c1[0]);
            tc.h_corr = L.c2;
            o_tc_w = o_w;   // TFLite [N][K] bytes are already the K-major B matrix
            kernel = Kernel::FcTc;
        }
        return;
    }
    if (L.op == MF_OP_AVERAGE_POOL_2D) { kernel = Kernel::PoolGeneric; return; }
    if (L.op == MF_OP_SOFTMAX) {
        o_lut = bb.add(L.exp_lut.data(), L.exp_lut.size() * 4);
        kernel = Kernel::Softmax;
        return;
    }
}

bool LayerExec::resolve(const uint8_t *d, std::string *err) {
    const LayerSpec &L = spec;
    auto at = [&](size_t off) { return off == SIZE_MAX ? nullptr : d + off; };
    if (L.op == MF_OP_CONV_2D || L.op == MF_OP_DEPTHWISE_CONV_2D) {
        ConvArgs &a = conv;
        a = ConvArgs{};
        a.w = at(o_w);
        a.w_zp = reinterpret_cast<const int32_t *>(at(o_wzp));
        a.c0z = reinterpret_cast<const float *>(at(o_c0z));
        a.c1 = reinterpret_cast<const float *>(at(o_c1));
        a.kcorr = reinterpret_cast<const int32_t *>(at(o_kcorr));
        a.H = L.H; a.W = L.W; a.Cin = L.Cin; a.OH = L.OH; a.OW = L.OW; a.Cout = L.Cout; a.KH = L.KH; a.KW = L.KW; a.sh = L.sh; a.sw = L.sw;
        a.off_r = L.pad == MF_PAD_SAME ? (L.KH - 1) / 2 : 0;   // src/tensor.rs:193
        a.off_c = L.pad == MF_PAD_SAME ? (L.KW - 1) / 2 : 0;
        a.in_zp = L.in_zp; a.lo = (float)L.act_lo; a.hi = (float)L.act_hi; a.is_u8 = L.is_u8; a.depthwise = L.op == MF_OP_DEPTHWISE_CONV_2D;
        a.big_acc = big_acc;
        a.wzp_nonzero = !std::all_of(L.w_zp.begin(), L.w_zp.end(), [](int32_t z) { return z == 0; });
        // plan() chose from the LayerSpec alone; the launch-side predicates (alignment, shared-memory budget, window geometry)
        // have the last word, so a shape they refuse runs on the generic kernel instead of failing at launch
        if (kernel == Kernel::PwConvDp4a && !pwconv_dp4a_eligible(a)) { kernel = Kernel::ConvGeneric; why_not_fast = "1x1 conv reads outside the input: generic kernel"; }
        if (kernel == Kernel::DwConvCin1 && !dwconv_cin1_eligible(a)) { kernel = Kernel::ConvGeneric; why_not_fast = "Cin=1 depthwise kernel too large for the fast kernel"; }
        if (kernel == Kernel::DwConvC4 && !dwconv_c4_general_eligible(a)) { kernel = Kernel::ConvGeneric; why_not_fast = "depthwise shape refused by the fast kernel"; }
        if (kernel == Kernel::DwConv3x3Rows && !dwconv_c4_eligible(a)) { kernel = Kernel::ConvGeneric; why_not_fast = "depthwise shape refused by the fast kernel"; }
        if (kernel == Kernel::ConvTcPointwise || kernel == Kernel::ConvTc3x3) {
            tc.d_wmat = at(o_tc_w);
            std::string why;
            if (!conv_tc_finalize_plan(tc, &why)) {  // fall back to the SIMT path, never to the CPU
                why_not_fast = "tensor core plan rejected: " + why;
                kernel = (L.KH == 1 && L.KW == 1 && L.Cin % 4 == 0 && pwconv_dp4a_eligible(a)) ? Kernel::PwConvDp4a : Kernel::ConvGeneric;
            }
        }
    } else if (L.op == MF_OP_FULLY_CONNECTED) {
        FcArgs &a = fc;
        a = FcArgs{};
        a.w = at(o_w);
        a.c0z = reinterpret_cast<const float *>(at(o_c0z));
        a.c2 = reinterpret_cast<const int32_t *>(at(o_c2));
        a.c1 = L.c1[0]; a.c3 = L.c3; a.w_zp = L.w_zp[0]; a.K = L.Cin; a.N = L.Cout;
        a.lo = (float)L.act_lo; a.hi = (float)L.act_hi; a.is_u8 = L.is_u8;
        if (kernel == Kernel::FcTc) {
            tc.d_wmat = at(o_tc_w);
            std::string why;
            if (!conv_tc_finalize_plan(tc, &why)) { why_not_fast = "tensor core plan rejected: " + why; kernel = Kernel::FcGeneric; }
        }
    } else if (L.op == MF_OP_AVERAGE_POOL_2D) {
        PoolArgs &a = pool;
        a = PoolArgs{};
        a.H = L.H; a.W = L.W; a.C = L.Cin; a.OH = L.OH; a.OW = L.OW; a.KH = L.KH; a.KW = L.KW; a.sh = L.sh; a.sw = L.sw;
        a.off_r = L.pad == MF_PAD_SAME ? (L.KH - 1) / 2 : 0;
        a.off_c = L.pad == MF_PAD_SAME ? (L.KW - 1) / 2 : 0;
        a.c0 = L.c0[0]; a.c1 = L.c1[0]; a.lo = (float)L.act_lo; a.hi = (float)L.act_hi; a.is_u8 = L.is_u8;
    } else if (L.op == MF_OP_SOFTMAX) {
        SoftmaxArgs &a = sm;
        a = SoftmaxArgs{};
        a.exp_lut = reinterpret_cast<const float *>(at(o_lut));
        a.rows = L.out_rank >= 1 ? L.out_dims[0] : 1;
        a.cols = (int)(L.out_elems / (size_t)(a.rows > 0 ? a.rows : 1));
        a.out_scale = L.out_scale; a.out_zp = (float)L.out_zp;
        a.lo = L.is_u8 ? 0.f : -128.f; a.hi = L.is_u8 ? 255.f : 127.f;
    }
    (void)err;
    return true;
}

const char *LayerExec::launched_name(const uint8_t *in, uint8_t *out, long long batch) const {
    if (kernel == Kernel::DwConv3x3Rows || kernel == Kernel::DwConvCin1) {
        static const bool no_smem = std::getenv("MF_DW_NO_SMEM") != nullptr;
        ConvArgs a = conv;
        a.in = in; a.out = out; a.batch = batch;
        if (!no_smem && kernel == Kernel::DwConv3x3Rows && dwconv3x3_smem_eligible(a)) return dwconv3x3_uses_pair(a) ? "dwconv3x3_pair_kernel" : "dwconv3x3_smem_kernel";
        if (!no_smem && kernel == Kernel::DwConvCin1 && dwconv_cin1_smem_eligible(a)) return "dwconv_cin1_smem_kernel";
        if (!no_smem && kernel == Kernel::DwConvCin1 && dwconv_cin1_taps_eligible(a)) return "dwconv_cin1_taps_kernel";
    }
    if (kernel == Kernel::ConvTc3x3) {
        static const bool pair = [] { const char *e = std::getenv("MF_TC_PAIR"); return !e || std::atoi(e) != 0; }();
        if (pair && tc.pair_ok && tc.patch && (long long)batch * ((spec.OH + tc.TH - 1) / tc.TH) * ((spec.OW + tc.TW - 1) / tc.TW) >= 2) return "conv3x3_pair_kernel";
    }
    return kernel_name(kernel);
}

cudaError_t LayerExec::run(const uint8_t *in, uint8_t *out, long long batch, int num_sms, cudaStream_t s, std::string *err, int pdl) const {
    switch (kernel) {
        case Kernel::None: return cudaSuccess;
        case Kernel::ConvGeneric: case Kernel::PwConvDp4a: case Kernel::DwConvC4: case Kernel::DwConv3x3Rows: case Kernel::DwConvCin1: {
            ConvArgs a = conv;
            a.in = in; a.out = out; a.batch = batch; a.pdl = pdl;
            if (kernel == Kernel::ConvGeneric) return launch_conv_generic(a, s);
            if (kernel == Kernel::PwConvDp4a) return launch_pwconv_dp4a(a, s);
            if (kernel == Kernel::DwConvC4) return launch_dwconv_c4(a, s);
            if (kernel == Kernel::DwConv3x3Rows) {
                static const bool no_smem = std::getenv("MF_DW_NO_SMEM") != nullptr;
                if (!no_smem && dwconv3x3_smem_eligible(a)) return launch_dwconv3x3_smem(a, num_sms, s);
                return launch_dwconv3x3_rows(a, s);
            }
            {
                static const bool no_smem = std::getenv("MF_DW_NO_SMEM") != nullptr;
                if (!no_smem && dwconv_cin1_smem_eligible(a)) return launch_dwconv_cin1_smem(a, num_sms, s);
                if (!no_smem && dwconv_cin1_taps_eligible(a)) return launch_dwconv_cin1_taps(a, num_sms, s);
            }
            return launch_dwconv_cin1(a, s);
        }
        case Kernel::ConvTcPointwise: {
            ConvTcLaunch l;
            l.in = in; l.out = out; l.pdl = pdl;
            l.W = l.OW = batch * spec.OH * spec.OW / tc_P;
            l.H = l.OH = 1; l.B = 1;
            return conv_tc_launch(tc, l, num_sms, s, err);
        }
        case Kernel::ConvTc3x3: {
            ConvTcLaunch l;
            l.in = in; l.out = out; l.pdl = pdl;
            l.W = l.OW = spec.W; l.H = l.OH = spec.H; l.B = batch;
            return conv_tc_launch(tc, l, num_sms, s, err);
        }
        case Kernel::FcTc: {
            ConvTcLaunch l;
            l.in = in; l.out = out; l.pdl = pdl;
            l.W = l.OW = batch; l.H = l.OH = 1; l.B = 1;
            return conv_tc_launch(tc, l, num_sms, s, err);
        }
        case Kernel::FcGeneric: case Kernel::FcWarp: {
            FcArgs a = fc;
            a.in = in; a.out = out; a.batch = batch; a.pdl = pdl;
            return kernel == Kernel::FcWarp ? launch_fc_warp(a, s) : launch_fc_generic(a, s);
        }
        case Kernel::PoolGeneric: {
            PoolArgs a = pool;
            a.in = in; a.out = out; a.batch = batch;
            return launch_pool_generic(a, s);
        }
        case Kernel::Softmax: {
            SoftmaxArgs a = sm;
            a.in = in; a.out = out; a.batch = batch;
            return launch_softmax(a, s);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace mf
