"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(tests/golden/kats.json, transcribed from the reference's `mod tests`)."""
import numpy as np

import oracle
from conftest import f32


def _arr(d, dtype=np.int8):
    return np.array(d["data"], dtype=dtype).reshape(d["shape"])


def test_conv_2d_kat(kats):
    k = kats["conv_2d"]
    x, f = _arr(k["input"]), _arr(k["filters"])
    out = oracle.conv_2d(x, k["input"]["zero_point"], f, k["filters"]["zero_point"], k["output_scale"], k["output_zero_point"], k["act"],
                         k["pad"], k["strides"], f32(k["constants"][0]), f32(k["constants"][1]), k["output"]["shape"][:2])
    np.testing.assert_array_equal(out, _arr(k["output"]))


def test_depthwise_conv_2d_kat(kats):
    k = kats["depthwise_conv_2d"]
    x, w = _arr(k["input"]), _arr(k["weights"])
    out = oracle.depthwise_conv_2d(x, k["input"]["zero_point"], w, k["weights"]["zero_point"], k["output_scale"], k["output_zero_point"],
                                   k["act"], k["pad"], k["strides"], f32(k["constants"][0]), f32(k["constants"][1]), k["output"]["shape"][:2])
    np.testing.assert_array_equal(out, _arr(k["output"]))


def test_fully_connected_kat(kats):
    k = kats["fully_connected"]
    x = _arr(k["input"])
    w_nk = np.ascontiguousarray(_arr(k["weights_kn"]).T)   # reference W is [K,N]; TFLite bytes are [N,K]
    c0, c1, c2, c3 = k["constants"]
    out = oracle.fully_connected(x, w_nk, k["weights_kn"]["zero_point"], k["output_scale"], k["output_zero_point"], k["act"], f32(c0), c1, c2, c3)
    np.testing.assert_array_equal(out, _arr(k["output"]))


def test_average_pool_2d_kat(kats):
    k = kats["average_pool_2d"]
    out = oracle.average_pool_2d(_arr(k["input"]), k["filter_shape"], k["output_scale"], k["output_zero_point"], k["act"], k["pad"], k["strides"],
                                 k["constants"][0], k["constants"][1], k["output"]["shape"][:2])
    np.testing.assert_array_equal(out, _arr(k["output"]))


def test_softmax_kat(kats):
    k = kats["softmax"]
    out = oracle.softmax(_arr(k["input"]), k["input"]["scale"], k["output_scale"], k["output_zero_point"])
    np.testing.assert_array_equal(out, _arr(k["output"]))


def test_quantize_kat(kats):
    k = kats["quantize"]
    assert oracle.quantize(k["value"], k["scale"], k["zero_point"]) == k["quantized"]
    assert oracle.dequantize(k["quantized"], k["scale"], k["zero_point"]) == np.float32(k["dequantized"])


def test_activation_kats(kats):
    k = kats["activation"]
    s, zp = k["scale"], k["zero_point"]
    assert oracle.relu(k["relu_inactive"][0], zp) == k["relu_inactive"][1]
    assert oracle.relu(k["relu_active"][0], zp) == k["relu_active"][1]
    assert oracle.relu6(k["relu6_saturated"][0], s, zp) == k["relu6_saturated"][1]
    outs = [oracle.softmax_scalar(v, k["softmax_sum"], s, zp) for v in k["softmax_inputs"]]
    assert outs[0] == k["softmax_output_1"]
    assert sum(outs) == k["softmax_total"]


def test_tensor_quantize_dequantize_kats(kats):
    t = kats["tensor"]["t2d"]
    q = [oracle.quantize(v, t["scale"], t["zero_point"]) for v in t["buffer"]]
    assert q == t["quantized"]
    dq = [oracle.dequantize(v, t["scale"], t["zero_point"]) for v in t["quantized"]]
    np.testing.assert_array_equal(f32(dq), f32(t["dequantized"]))
    t = kats["tensor"]["t4d"]
    q = [oracle.quantize(v, t["scale"], t["zero_point"]) for v in t["buffer"]]
    assert q == t["quantized"]
    dq = [oracle.dequantize(v, t["scale"], t["zero_point"]) for v in t["quantized"]]
    np.testing.assert_array_equal(f32(dq), f32(t["buffer"]))


def test_tensor_view_kat(kats):
    """src/tensor.rs:391-401: the SAME-padding view (zero fill, mask, len) -- observed through a 1-filter conv whose
    filter is all ones: acc = sum(view) - in_zp * (#valid * C) with c1 = 1, c0 = 0, out_zp = 0."""
    t4, v = kats["tensor"]["t4d"], kats["tensor"]["view"]
    x = np.array(t4["quantized"], np.int8).reshape(t4["shape"])[v["batch"]]
    ones = np.ones((1, v["view_shape"][0], v["view_shape"][1], x.shape[2]), np.int8)
    # in_zp = 0 -> output = sum of the view buffer
    out = oracle.conv_2d(x, 0, ones, [0], 1.0, 0, "none", v["pad"], v["strides"], f32([0.0]), f32([1.0]), x.shape[:2])
    assert int(out[v["focus"][0], v["focus"][1], 0]) == min(127, sum(v["buffer"]))
    # in_zp = 1 -> output = sum(view) - len * C : pins `len`/mask
    x2 = np.ones_like(x)
    out = oracle.conv_2d(x2, 1, ones, [0], 1.0, 0, "none", v["pad"], v["strides"], f32([0.0]), f32([1.0]), x.shape[:2])
    assert int(out[v["focus"][0], v["focus"][1], 0]) == v["len"] * x.shape[2] - v["len"] * x.shape[2]
    x3 = np.full_like(x, 3)
    out = oracle.conv_2d(x3, 1, ones, [0], 1.0, 0, "none", v["pad"], v["strides"], f32([0.0]), f32([1.0]), x.shape[:2])
    assert int(out[v["focus"][0], v["focus"][1], 0]) == 2 * v["len"] * x.shape[2]


def test_preprocess_kats(kats):
    p = kats["preprocess"]
    k = p["conv_2d"]
    c0, c1 = oracle.conv_preprocess(k["input_scale"], k["filters_scale"], k["biases_scale"], k["biases"], k["biases_zero_point"],
                                    k["output_scale"], k["filters_shape"][0])
    np.testing.assert_array_equal(c0, f32(k["c0"]))
    np.testing.assert_array_equal(c1, f32(k["c1"]))
    k = p["depthwise_conv_2d"]
    c0, c1 = oracle.conv_preprocess(k["input_scale"], k["weights_scale"], k["biases_scale"], k["biases"], k["biases_zero_point"],
                                    k["output_scale"], k["weights_shape"][3])
    np.testing.assert_array_equal(c0, f32(k["c0"]))
    np.testing.assert_array_equal(c1, f32(k["c1"]))
    k = p["fully_connected"]
    w_nk = np.ascontiguousarray(np.array(k["weights_kn"]["data"], np.int8).reshape(k["weights_kn"]["shape"]).T)
    c0, c1, c2, c3 = oracle.fc_preprocess(k["input_scale"], k["input_zero_point"], k["input_shape"][1], w_nk, k["weights_kn"]["scale"],
                                          k["weights_kn"]["zero_point"], k["biases_scale"], k["biases"], k["biases_zero_point"], k["output_scale"])
    np.testing.assert_array_equal(c0, f32(k["c0"]))
    assert c1 == np.float32(k["c1"])
    assert list(c2) == k["c2"] and c3 == k["c3"]
    k = p["average_pool_2d"]
    c0, c1 = oracle.pool_preprocess(k["input_scale"], k["input_zero_point"], k["output_scale"], k["output_zero_point"])
    assert c0 == np.float32(k["c0"]) and c1 == np.float32(k["c1"])
