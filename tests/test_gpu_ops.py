"""GPU parity, op level: every reference op KAT driven through the C-ABI per-op hooks, then seeded differential tests
CUDA-vs-oracle (bit-exact) over shapes of the three models, odd shapes, strides, paddings, non-zero weight
zero-points and uint8.  `impl=1` = generic kernels, `impl=0` = whatever the engine would pick (fast / tensor core)."""
import zlib

import numpy as np
import pytest

import microflow_rs_b200 as mf
import oracle
from conftest import f32

pytestmark = pytest.mark.gpu


def _arr(d, dtype=np.int8):
    return np.array(d["data"], dtype=dtype).reshape(d["shape"])


def rng(seed):
    return np.random.default_rng(seed)


# ---------------------------------------------------------------- reference KATs on the GPU ------------------
@pytest.mark.parametrize("impl", [0, 1])
def test_conv_2d_kat(kats, impl):
    k = kats["conv_2d"]
    x, f = _arr(k["input"]), _arr(k["filters"])
    out = mf.ops.conv_2d(x[None], k["input"]["zero_point"], f, k["filters"]["zero_point"], k["output_scale"], k["output_zero_point"], k["act"],
                         k["pad"], k["strides"], f32(k["constants"][0]), f32(k["constants"][1]), k["output"]["shape"][:2], impl=impl)
    np.testing.assert_array_equal(out[0], _arr(k["output"]))


@pytest.mark.parametrize("impl", [0, 1])
def test_depthwise_conv_2d_kat(kats, impl):
    k = kats["depthwise_conv_2d"]
    x, w = _arr(k["input"]), _arr(k["weights"])
    out = mf.ops.depthwise_conv_2d(x[None], k["input"]["zero_point"], w, k["weights"]["zero_point"], k["output_scale"], k["output_zero_point"],
                                   k["act"], k["pad"], k["strides"], f32(k["constants"][0]), f32(k["constants"][1]), k["output"]["shape"][:2],
                                   impl=impl)
    np.testing.assert_array_equal(out[0], _arr(k["output"]))


@pytest.mark.parametrize("impl", [0, 1])
def test_fully_connected_kat(kats, impl):
    k = kats["fully_connected"]
    x = _arr(k["input"])
    w_nk = np.ascontiguousarray(_arr(k["weights_kn"]).T)
    c0, c1, c2, c3 = k["constants"]
    out = mf.ops.fully_connected(x, w_nk, k["weights_kn"]["zero_point"], k["output_scale"], k["output_zero_point"], k["act"], f32(c0), c1, c2, c3,
                                 impl=impl)
    np.testing.assert_array_equal(out, _arr(k["output"]))


def test_average_pool_2d_kat(kats):
    k = kats["average_pool_2d"]
    out = mf.ops.average_pool_2d(_arr(k["input"])[None], k["filter_shape"], k["output_scale"], k["output_zero_point"], k["act"], k["pad"],
                                 k["strides"], k["constants"][0], k["constants"][1], k["output"]["shape"][:2])
    np.testing.assert_array_equal(out[0], _arr(k["output"]))


def test_softmax_kat(kats):
    k = kats["softmax"]
    out = mf.ops.softmax(_arr(k["input"])[None], k["input"]["scale"], k["output_scale"], k["output_zero_point"])
    np.testing.assert_array_equal(out[0], _arr(k["output"]))


def test_quantize_dequantize_kats(kats):
    t = kats["tensor"]["t2d"]
    np.testing.assert_array_equal(mf.ops.quantize(f32(t["buffer"]), t["scale"], t["zero_point"]), np.array(t["quantized"], np.int8))
    np.testing.assert_array_equal(mf.ops.dequantize(np.array(t["quantized"], np.int8), t["scale"], t["zero_point"]), f32(t["dequantized"]))
    t = kats["tensor"]["t4d"]
    np.testing.assert_array_equal(mf.ops.quantize(f32(t["buffer"]), t["scale"], t["zero_point"]), np.array(t["quantized"], np.int8))
    np.testing.assert_array_equal(mf.ops.dequantize(np.array(t["quantized"], np.int8), t["scale"], t["zero_point"]), f32(t["buffer"]))
    k = kats["quantize"]
    assert mf.ops.quantize(f32([k["value"]]), k["scale"], k["zero_point"])[0] == k["quantized"]


def test_quantize_matches_oracle_on_ties_and_extremes():
    xs = f32(np.concatenate([np.arange(-70, 70) * 0.05, [0.5, -0.5, 1.5, 2.5, -2.5, 1e9, -1e9, 0.49999997, 12.7, 12.75, -12.85, np.nan]]))
    for scale, zp in [(0.1, 2), (0.05, -3), (1.0, 0), (0.0078431, -1)]:
        got = mf.ops.quantize(xs, scale, zp)
        want = np.array([oracle.quantize(v, scale, zp) for v in xs], np.int8)
        np.testing.assert_array_equal(got, want)


# ---------------------------------------------------------------- differential helpers -----------------------
def _conv_case(r, B, H, W, Cin, Cout, KH, KW, sh, sw, pad, act, dw, dtype, wzp0, per_channel=True):
    lo, hi = (0, 256) if dtype == np.uint8 else (-128, 128)
    x = r.integers(lo, hi, (B, H, W, Cin)).astype(dtype)
    wshape = (1, KH, KW, Cout) if dw else (Cout, KH, KW, Cin)
    w = r.integers(lo, hi, wshape).astype(dtype)
    nq = Cout if per_channel else 1
    wzp = np.zeros(nq, np.int32) if wzp0 else r.integers(lo, hi, nq).astype(np.int32)
    in_zp = int(r.integers(lo, hi))
    out_zp = int(r.integers(lo, hi))
    kdim = KH * KW * (1 if dw else Cin)
    c1 = (r.uniform(0.2, 2.0, nq) / (kdim * 40.0)).astype(np.float32)
    c0 = r.uniform(-20, 20, Cout).astype(np.float32)
    if pad == "same":
        OH, OW = -(-H // sh), -(-W // sw)
    else:
        OH, OW = (H - KH) // sh + 1, (W - KW) // sw + 1
    out_scale = np.float32(r.uniform(0.02, 0.2))
    return dict(x=x, in_zp=in_zp, w=w, wzp=wzp, out_scale=out_scale, out_zp=out_zp, act=act, pad=pad, strides=(sh, sw), c0=c0, c1=c1,
                out_hw=(OH, OW), dw=dw)


def _run_conv(c, impl):
    got = mf.ops.conv_2d(c["x"], c["in_zp"], c["w"], c["wzp"], c["out_scale"], c["out_zp"], c["act"], c["pad"], c["strides"], c["c0"], c["c1"],
                         c["out_hw"], depthwise=c["dw"], impl=impl)
    kern = mf.ops.last_kernel
    want = np.stack([oracle.conv_2d(c["x"][b], c["in_zp"], c["w"], c["wzp"], c["out_scale"], c["out_zp"], c["act"], c["pad"], c["strides"], c["c0"],
                                    c["c1"], c["out_hw"], depthwise=c["dw"]) for b in range(c["x"].shape[0])])
    return got, want, kern


GENERIC_CASES = [
    # B, H, W, Cin, Cout, KH, KW, sh, sw, pad, act, dw, dtype, wzp0
    (2, 5, 7, 3, 4, 3, 3, 1, 1, "same", "none", False, np.int8, False),
    (2, 6, 6, 2, 5, 3, 3, 2, 2, "same", "relu", False, np.int8, False),     # stride 2, even input: MicroFlow pads top/left
    (1, 9, 8, 4, 3, 2, 3, 1, 2, "same", "relu6", False, np.int8, False),    # even kernel height
    (2, 8, 9, 3, 2, 3, 2, 2, 1, "valid", "none", False, np.int8, False),
    (2, 7, 5, 3, 3, 3, 3, 1, 1, "same", "relu6", True, np.int8, False),     # depthwise, general zero points
    (2, 6, 6, 1, 4, 3, 3, 2, 2, "same", "none", True, np.int8, False),      # depth multiplier (channel fallback to 0)
    (1, 5, 5, 2, 5, 2, 2, 1, 1, "valid", "relu", True, np.int8, False),     # Cout > Cin > 1: channels >= Cin read channel 0
    (2, 5, 6, 3, 4, 3, 3, 1, 1, "same", "relu", False, np.uint8, False),    # uint8
    (2, 5, 6, 4, 4, 3, 3, 2, 1, "same", "none", True, np.uint8, False),
    (1, 1, 1, 8, 4, 1, 1, 1, 1, "same", "none", False, np.int8, True),
    (1, 3, 3, 4, 4, 5, 5, 1, 1, "same", "none", False, np.int8, True),      # kernel larger than the image
]


@pytest.mark.parametrize("case", GENERIC_CASES)
def test_conv_generic_vs_oracle(case):
    B, H, W, Cin, Cout, KH, KW, sh, sw, pad, act, dw, dtype, wzp0 = case
    c = _conv_case(rng(zlib.crc32(repr(case).encode())), B, H, W, Cin, Cout, KH, KW, sh, sw, pad, act, dw, dtype, wzp0)
    got, want, kern = _run_conv(c, impl=1)
    assert "generic" in kern
    np.testing.assert_array_equal(got, want)


FAST_SIMT_CASES = [
    # shapes of person_detect / speech depthwise layers (scaled-down spatial extents keep the oracle fast) + odd ones
    (3, 12, 12, 8, 8, 3, 3, 1, 1, "same", "relu6", True, "dwconv3x3_rows"),
    (3, 12, 12, 16, 16, 3, 3, 2, 2, "same", "relu6", True, "dwconv3x3_rows"),
    (2, 6, 6, 128, 128, 3, 3, 1, 1, "same", "relu6", True, "dwconv3x3_rows"),
    (2, 3, 3, 256, 256, 3, 3, 1, 1, "same", "relu6", True, "dwconv3x3_rows"),
    (2, 7, 9, 12, 12, 5, 3, 2, 1, "same", "relu", True, "dwconv_c4"),
    (2, 9, 7, 4, 4, 3, 3, 1, 1, "valid", "none", True, "dwconv3x3_rows"),
    (3, 96, 96, 1, 8, 3, 3, 2, 2, "same", "relu6", True, "dwconv_cin1"),     # person_detect layer 0
    (3, 49, 40, 1, 8, 10, 8, 2, 2, "same", "relu", True, "dwconv_cin1"),     # speech layer 1
    (2, 11, 13, 1, 16, 3, 5, 1, 2, "valid", "none", True, "dwconv_cin1"),
    (2, 48, 48, 8, 8, 3, 3, 1, 1, "same", "relu", True, "dwconv3x3_rows"),          # several row strips, clamp != full int8 range
    (2, 24, 24, 32, 32, 3, 3, 2, 2, "same", "relu6", True, "dwconv3x3_rows"),       # stride 2 over several strips
    (2, 13, 11, 8, 8, 3, 3, 2, 2, "valid", "relu6", True, "dwconv3x3_rows"),
    (2, 17, 5, 4, 4, 3, 3, 1, 1, "same", "none", True, "dwconv3x3_rows"),
    (2, 9, 9, 8, 8, 3, 3, 2, 1, "same", "none", True, "dwconv_c4"),
    # batch >= 296 switches the 3x3 depthwise to the sample-resident (cp.async.bulk) kernel
    (300, 12, 12, 8, 8, 3, 3, 1, 1, "same", "relu6", True, "dwconv3x3_(smem|pair)"),
    (300, 8, 8, 32, 32, 3, 3, 2, 2, "same", "relu6", True, "dwconv3x3_(smem|pair)"),
    (333, 6, 6, 128, 128, 3, 3, 1, 1, "same", "relu", True, "dwconv3x3_(smem|pair)"),
    (300, 9, 7, 16, 16, 3, 3, 1, 1, "valid", "none", True, "dwconv3x3_(smem|rows|pair)"),
    (300, 11, 13, 8, 8, 3, 3, 2, 2, "valid", "relu6", True, "dwconv3x3_(smem|rows|pair)"),
    (300, 3, 3, 256, 256, 3, 3, 1, 1, "same", "relu6", True, "dwconv3x3_(smem|pair)"),
    # window columns outside the image are handled by zeroed per-thread weights + a per-thread correction term:
    (300, 7, 5, 16, 16, 3, 3, 2, 2, "same", "relu6", True, "dwconv3x3_(smem|pair)"),              # odd width, stride 2: the right column falls outside
    (300, 4, 1, 32, 32, 3, 3, 1, 1, "same", "none", True, "dwconv3x3_(smem|pair)"),               # one-pixel-wide image: left AND right outside for the same thread
    (300, 6, 6, 64, 64, 3, 3, 1, 1, "same", "relu", True, "dwconv3x3_(smem|pair)"),               # clamp narrower than int8 (FULL = false epilogue)
    (300, 2, 2, 8, 8, 3, 3, 1, 1, "same", "relu6", True, "dwconv3x3_(smem|pair)"),                # every output pixel is a corner
    (300, 32, 32, 1, 8, 3, 3, 2, 2, "same", "relu6", True, "dwconv_cin1_smem"),        # sample-resident Cin=1 kernel (layer-0 shape, scaled down)
    (300, 20, 16, 1, 8, 3, 3, 1, 1, "valid", "relu", True, "dwconv_cin1"),                 # same kernel, stride 1, no padding, partial clamp
    (300, 10, 32, 1, 8, 3, 3, 1, 1, "same", "none", True, "dwconv_cin1"),                  # stride 1: a column outside on both sides
    (300, 96, 96, 1, 8, 3, 3, 2, 2, "same", "relu6", True, "dwconv_cin1_smem"),            # person_detect layer 0 at full extent
    (300, 15, 48, 1, 8, 3, 3, 2, 2, "valid", "relu6", True, "dwconv_cin1"),
    (300, 20, 16, 1, 8, 3, 3, 2, 1, "same", "relu", True, "dwconv_cin1"),                  # mixed strides stay on the generic-shape fast kernel
    # large kernels with Cin == 1: dp4a along the taps of a kernel row inside a zero-point frame (speech layer 1 and relatives)
    (300, 49, 40, 1, 8, 10, 8, 2, 2, "same", "relu", True, "dwconv_cin1_taps"),            # speech layer 1 at full extent
    (300, 20, 16, 1, 8, 4, 4, 1, 1, "valid", "none", True, "dwconv_cin1_taps"),            # one word of taps per row, no frame, full clamp
    (300, 21, 24, 1, 8, 7, 5, 2, 1, "same", "relu6", True, "dwconv_cin1_taps"),            # 5 taps per row (zero-padded to 8), mixed strides
    (300, 9, 8, 1, 8, 12, 8, 1, 1, "same", "relu", True, "dwconv_cin1_taps"),              # kernel taller than the image
    (3, 1, 1, 256, 2, 1, 1, 1, 1, "same", "none", False, "pwconv_dp4a"),            # person_detect's last conv (Cout = 2)
    (2, 4, 4, 16, 7, 1, 1, 1, 1, "same", "relu", False, "pwconv_dp4a"),
    (3, 12, 12, 8, 16, 1, 1, 1, 1, "same", "relu6", False, "pwconv_dp4a|conv_tc"),   # person_detect layer 2 shape
    (2, 5, 5, 12, 20, 1, 1, 1, 1, "same", "none", False, "pwconv_dp4a"),
    (2, 6, 6, 8, 4, 1, 1, 2, 2, "same", "relu", False, "pwconv_dp4a"),
    (5, 1, 1, 256, 4, 1, 1, 1, 1, "same", "none", False, "pwconv_dp4a"),
]


@pytest.mark.parametrize("case", FAST_SIMT_CASES)
def test_conv_fast_simt_vs_oracle(case):
    import re
    B, H, W, Cin, Cout, KH, KW, sh, sw, pad, act, dw, kname = case
    c = _conv_case(rng(zlib.crc32(repr(case).encode())), B, H, W, Cin, Cout, KH, KW, sh, sw, pad, act, dw, np.int8, True)
    got, want, kern = _run_conv(c, impl=0)
    assert re.search(kname, kern), kern
    np.testing.assert_array_equal(got, want)
    got_g, _, _ = _run_conv(c, impl=1)
    np.testing.assert_array_equal(got_g, want)


FAST_GENERAL_CASES = [
    # uint8 tensors (`T = u8`, microflow-macros/src/ops/conv_2d.rs:39-46) and non-zero weight zero-points (the view-sum term,
    # src/ops/conv_2d.rs:74-76, depthwise_conv_2d.rs:71-73) on the FAST kernels: impl=2 refuses the generic kernel.
    # B, H, W, Cin, Cout, KH, KW, sh, sw, pad, act, dw, dtype, wzp0, kernel
    (3, 12, 12, 8, 8, 3, 3, 1, 1, "same", "relu6", True, np.uint8, False, "dwconv_c4"),
    (2, 7, 9, 12, 12, 5, 3, 2, 1, "same", "relu", True, np.int8, False, "dwconv_c4"),
    (300, 6, 6, 128, 128, 3, 3, 1, 1, "same", "relu6", True, np.uint8, True, "dwconv_c4"),
    (2, 9, 7, 4, 4, 3, 3, 1, 1, "valid", "none", True, np.int8, False, "dwconv_c4"),
    (2, 4, 4, 16, 7, 1, 1, 1, 1, "same", "relu", False, np.int8, False, "pwconv_dp4a"),
    (3, 12, 12, 8, 16, 1, 1, 1, 1, "same", "relu6", False, np.uint8, False, "pwconv_dp4a"),
    (2, 6, 6, 8, 4, 1, 1, 2, 2, "same", "relu", False, np.uint8, False, "pwconv_dp4a"),
    (2, 5, 5, 12, 20, 1, 1, 1, 1, "same", "none", False, np.uint8, True, "pwconv_dp4a"),
    (3, 24, 24, 32, 32, 1, 1, 1, 1, "same", "relu6", False, np.uint8, True, "conv_tc"),     # tcgen05 with unsigned operand formats, packed pixels
    (9, 6, 6, 128, 128, 1, 1, 1, 1, "same", "relu", False, np.uint8, True, "conv_tc"),
    (5, 3, 3, 256, 256, 1, 1, 1, 1, "same", "none", False, np.uint8, True, "conv_tc"),       # two 128-byte channel blocks
    (2, 9, 21, 128, 64, 3, 3, 1, 1, "same", "relu6", False, np.uint8, True, "conv_tc"),      # 3x3 implicit GEMM, zero-filled borders + class table
]


@pytest.mark.parametrize("case", FAST_GENERAL_CASES)
def test_uint8_and_weight_zero_points_on_fast_kernels(case):
    import re
    B, H, W, Cin, Cout, KH, KW, sh, sw, pad, act, dw, dtype, wzp0, kname = case
    c = _conv_case(rng(zlib.crc32(repr(case).encode())), B, H, W, Cin, Cout, KH, KW, sh, sw, pad, act, dw, dtype, wzp0)
    if kname == "conv_tc":                      # keep the requantized values inside the output range for a meaningful comparison
        c["c1"] = (c["c1"] / 4).astype(np.float32)
    if dtype == np.uint8 and wzp0:              # unsigned weights with zero-point 0 are all positive: large accumulators
        c["c1"] = (c["c1"] / 24).astype(np.float32)
        c["in_zp"] = 128
    nchk = min(B, 6)
    got = mf.ops.conv_2d(c["x"], c["in_zp"], c["w"], c["wzp"], c["out_scale"], c["out_zp"], c["act"], c["pad"], c["strides"], c["c0"], c["c1"], c["out_hw"],
                         depthwise=c["dw"], impl=2)
    # the tcgen05 path: 3x3 layers with splittable output channels run on CTA pairs (conv3x3_pair_kernel), the rest on conv_tc_kernel
    assert re.search("conv_tc|conv3x3_pair" if kname == "conv_tc" else kname, mf.ops.last_kernel), mf.ops.last_kernel
    want = np.stack([oracle.conv_2d(c["x"][b], c["in_zp"], c["w"], c["wzp"], c["out_scale"], c["out_zp"], c["act"], c["pad"], c["strides"], c["c0"],
                                    c["c1"], c["out_hw"], depthwise=c["dw"]) for b in range(nchk)])
    np.testing.assert_array_equal(got[:nchk], want)
    ref = mf.ops.conv_2d(c["x"], c["in_zp"], c["w"], c["wzp"], c["out_scale"], c["out_zp"], c["act"], c["pad"], c["strides"], c["c0"], c["c1"], c["out_hw"],
                         depthwise=c["dw"], impl=1)
    np.testing.assert_array_equal(got, ref)
    assert 0.01 < np.mean((got != got.flat[0])), "degenerate output"


@pytest.mark.parametrize("K,N,B,wzp,dtype,kname", [(48, 3, 40, 9, np.uint8, "fc_warp"), (64, 8, 6, -7, np.int8, "fc_warp"), (4000, 4, 9, 200, np.uint8, "fc_warp"),
                                                    (256, 64, 129, 0, np.uint8, "conv_tc"), (128, 32, 300, 0, np.uint8, "conv_tc")])
def test_fully_connected_uint8_and_weight_zero_point_on_fast_kernels(K, N, B, wzp, dtype, kname):
    r = rng(K * 7 + N)
    lo, hi = (0, 256) if dtype == np.uint8 else (-128, 128)
    x = r.integers(lo, hi, (B, K)).astype(dtype)
    w = r.integers(lo, hi, (N, K)).astype(dtype)
    in_zp = int(r.integers(lo, hi))
    c0 = r.uniform(-10, 10, N).astype(np.float32)
    c1 = np.float32(1.0 / (K * 60.0))
    c2 = (w.astype(np.int32).sum(1) * in_zp).astype(np.int32)
    c3 = K * in_zp * wzp
    out_zp = 100 if dtype == np.uint8 else 3
    got = mf.ops.fully_connected(x, w, wzp, 0.1, out_zp, "relu", c0, c1, c2, c3, impl=2)
    assert kname in mf.ops.last_kernel, mf.ops.last_kernel
    want = oracle.fully_connected(x, w, wzp, 0.1, out_zp, "relu", c0, c1, c2, c3)
    np.testing.assert_array_equal(got, want)


def test_requant_saturation_and_large_accumulators():
    """Accumulators beyond 2^24 (i32->f32 rounding), saturation at both ends, ties."""
    r = rng(7)
    B, H, W, Cin, Cout = 2, 4, 4, 1152, 8
    x = r.integers(100, 128, (B, H, W, Cin)).astype(np.int8)
    w = r.integers(100, 128, (Cout, 1, 1, Cin)).astype(np.int8)
    w[1] = -w[1]
    c1 = f32([1e-5, 1e-5, 3e-6, 1.0, 0.5, 0.25, 7.62939453125e-06, 2.0])
    c0 = f32([0, 0, 0.5, 0, -0.5, 1e9, -30.5, -1e9])
    for impl in (0, 1):
        got = mf.ops.conv_2d(x, -128, w, [0], 0.05, -3, "none", "same", (1, 1), c0, c1, (H, W), impl=impl)
        want = np.stack([oracle.conv_2d(x[b], -128, w, [0], 0.05, -3, "none", "same", (1, 1), c0, c1, (H, W)) for b in range(B)])
        np.testing.assert_array_equal(got, want)


# kernels whose requantize epilogue converts with the packed F2IP (mf_device.cuh f2i_pack4): accumulators are made EXACTLY 2 * x by a
# filter that is 2 at one tap and 0 elsewhere, and the per-channel constants walk through exact .5 ties of both signs (c1 = 0.25 ->
# t = c0 + x / 2), integers, quarter steps, saturation at both ends, huge offsets and vanishing scales.
TIE_CASES = [
    # B, H, W, Cin, Cout, KH, KW, stride, depthwise, act, kernel
    (3, 24, 24, 32, 32, 1, 1, 1, False, "none", "conv_tc"),                      # tcgen05 pointwise, packed pixels, pre-biased accumulators
    (3, 24, 24, 32, 32, 1, 1, 1, False, "relu6", "conv_tc"),                     # clamp narrower than int8: requant4_clamp
    (9, 6, 6, 128, 128, 1, 1, 1, False, "none", "conv_tc"),
    (2, 16, 16, 128, 128, 3, 3, 1, False, "none", "conv3x3_pair|conv_tc"),       # tcgen05 3x3 on CTA pairs (the I2F variant of the epilogue is
                                                                                 # covered by test_config5_full_size_image_vs_oracle)
    (300, 12, 12, 16, 16, 3, 3, 1, True, "none", "dwconv3x3_pair"),
    (300, 12, 12, 16, 16, 3, 3, 2, True, "none", "dwconv3x3_smem"),
    (300, 12, 12, 16, 16, 3, 3, 1, True, "relu", "dwconv3x3_pair"),              # FULL = false epilogue
    (300, 32, 32, 1, 8, 3, 3, 2, True, "none", "dwconv_cin1_smem"),
    (3, 12, 12, 8, 8, 3, 3, 1, True, "none", "dwconv3x3_rows"),
]


@pytest.mark.parametrize("case", TIE_CASES)
def test_epilogue_ties_and_saturation_on_fast_kernels(case):
    import re
    B, H, W, Cin, Cout, KH, KW, st, dw, act, kname = case
    r = rng(zlib.crc32(repr(case).encode()))
    x = r.integers(-128, 128, (B, H, W, Cin)).astype(np.int8)
    x[0, 0, 0, :] = 127
    x[0, 0, 1 % W, :] = -128
    if dw:
        w = np.zeros((1, KH, KW, Cout), np.int8)
        w[0, KH // 2, KW // 2, :] = 2
    else:
        w = np.zeros((Cout, KH, KW, Cin), np.int8)
        for n in range(Cout):
            w[n, KH // 2, KW // 2, n % Cin] = 2
    c1v = [0.25, 0.25, 0.25, 0.125, 0.75, 1.0, 64.0, 1e-3, 0.5, 0.375, 0.25, 3.0e38 / 256, 0.0, 0.24999999, 0.25000003, 0.5]
    c0v = [0.0, 0.5, -0.5, 0.25, 0.125, 0.0, 0.0, 0.499, -63.5, 64.5, 1e9, 0.0, -0.5, 0.0, 0.0, -1e9]
    c1 = np.array([c1v[n % 16] for n in range(Cout)], np.float32)
    c0 = np.array([c0v[n % 16] for n in range(Cout)], np.float32)
    OH, OW = -(-H // st), -(-W // st)
    args = (x, 0, w, [0], np.float32(0.0235294), -3 if act == "none" else -128, act, "same", (st, st), c0, c1, (OH, OW))
    got = mf.ops.conv_2d(*args, depthwise=dw, impl=0)
    assert re.search(kname, mf.ops.last_kernel), mf.ops.last_kernel
    want = np.stack([oracle.conv_2d(x[b], *args[1:], depthwise=dw) for b in range(min(B, 4))])
    np.testing.assert_array_equal(got[:min(B, 4)], want)
    np.testing.assert_array_equal(mf.ops.conv_2d(*args, depthwise=dw, impl=1), got)       # every sample against the generic kernel


@pytest.mark.parametrize("K,N,B,wzp,dtype", [(1, 16, 5, 0, np.int8), (16, 16, 7, 0, np.int8), (16, 1, 3, 0, np.int8), (4000, 4, 9, 0, np.int8),
                                              (37, 5, 4, 22, np.int8), (64, 8, 6, -7, np.int8), (48, 3, 4, 9, np.uint8)])
def test_fully_connected_vs_oracle(K, N, B, wzp, dtype):
    r = rng(K * 131 + N)
    lo, hi = (0, 256) if dtype == np.uint8 else (-128, 128)
    x = r.integers(lo, hi, (B, K)).astype(dtype)
    w = r.integers(lo, hi, (N, K)).astype(dtype)
    in_zp = int(r.integers(lo, hi))
    c0 = r.uniform(-10, 10, N).astype(np.float32)
    c1 = np.float32(1.0 / (K * 30.0))
    c2 = (w.astype(np.int32).sum(1) * in_zp).astype(np.int32)
    c3 = K * in_zp * wzp
    for impl in (0, 1):
        got = mf.ops.fully_connected(x, w, wzp, 0.1, 3, "relu", c0, c1, c2, c3, impl=impl)
        want = oracle.fully_connected(x, w, wzp, 0.1, 3, "relu", c0, c1, c2, c3)
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("K,N,B", [(128, 32, 300), (256, 64, 129), (384, 256, 40), (1152, 96, 33)])
def test_fully_connected_on_tensor_core_vs_oracle(K, N, B):
    """FullyConnected through the tcgen05 GEMM (w_zp == 0, K % 128 == 0, N % 32 == 0); K = 1152 exceeds the 2^22 accumulator bound."""
    r = rng(K + N)
    x = r.integers(-128, 128, (B, K)).astype(np.int8)
    w = r.integers(-128, 128, (N, K)).astype(np.int8)
    in_zp = int(r.integers(-128, 128))
    c0 = r.uniform(-10, 10, N).astype(np.float32)
    c1 = np.float32(1.0 / (K * 30.0))
    c2 = (w.astype(np.int32).sum(1) * in_zp).astype(np.int32)
    got = mf.ops.fully_connected(x, w, 0, 0.1, 3, "relu", c0, c1, c2, 0, impl=0)
    assert "conv_tc" in mf.ops.last_kernel, mf.ops.last_kernel
    want = oracle.fully_connected(x, w, 0, 0.1, 3, "relu", c0, c1, c2, 0)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(mf.ops.fully_connected(x, w, 0, 0.1, 3, "relu", c0, c1, c2, 0, impl=1), want)


@pytest.mark.parametrize("H,W,C,FH,FW,sh,sw,pad,dtype", [(3, 3, 256, 3, 3, 2, 2, "valid", np.int8), (7, 9, 5, 2, 3, 1, 2, "same", np.int8),
                                                           (8, 8, 4, 3, 3, 2, 2, "same", np.uint8)])
def test_average_pool_vs_oracle(H, W, C, FH, FW, sh, sw, pad, dtype):
    r = rng(H * 100 + C)
    lo, hi = (0, 256) if dtype == np.uint8 else (-128, 128)
    x = r.integers(lo, hi, (3, H, W, C)).astype(dtype)
    OH, OW = (-(-H // sh), -(-W // sw)) if pad == "same" else ((H - FH) // sh + 1, (W - FW) // sw + 1)
    c0, c1 = oracle.pool_preprocess(0.0235294, -128 if dtype == np.int8 else 3, 0.0186093, -128 if dtype == np.int8 else 7)
    got = mf.ops.average_pool_2d(x, (FH, FW), 0.0186093, -128 if dtype == np.int8 else 7, "none", pad, (sh, sw), c0, c1, (OH, OW))
    want = np.stack([oracle.average_pool_2d(x[b], (FH, FW), 0.0186093, -128 if dtype == np.int8 else 7, "none", pad, (sh, sw), c0, c1, (OH, OW))
                     for b in range(3)])
    np.testing.assert_array_equal(got, want)


def test_softmax_all_256_inputs_vs_oracle():
    """Every int8 value through the exp table, for the two scales the shipped models use."""
    for in_scale in (0.0917319, 0.0125188):
        x = np.arange(-128, 128, dtype=np.int8).reshape(64, 1, 4)
        got = mf.ops.softmax(x, in_scale, 1.0 / 256.0, -128)
        want = np.stack([oracle.softmax(x[b], in_scale, 1.0 / 256.0, -128) for b in range(64)])
        np.testing.assert_array_equal(got, want)
