"""Pins the oracle's loader + graph runner against the reference's end-to-end goldens (tests/*.rs), the 500-row
sine accuracy CSV (analysis/accuracy/data/sine-microflow.csv) and the survey's derived pins for the sample inputs."""
import numpy as np
import pytest

import oracle
from conftest import GOLDEN, MODELS, f32


@pytest.fixture(scope="module")
def models():
    return {n: oracle.Model(MODELS / f"{n}.tflite") for n in ("sine", "speech", "person_detect")}


def test_model_io(models):
    assert models["sine"].in_shape == (1, 1) and models["sine"].out_shape == (1, 1)
    assert models["speech"].in_shape == (1, 1960) and models["speech"].out_shape == (1, 4)
    assert models["person_detect"].in_shape == (1, 96, 96, 1) and models["person_detect"].out_shape == (1, 2)
    assert [L["op"] for L in models["sine"].layers] == ["fully_connected"] * 3
    assert [L["op"] for L in models["speech"].layers] == ["reshape", "depthwise_conv_2d", "fully_connected", "softmax"]
    assert len(models["person_detect"].layers) == 31


@pytest.mark.parametrize("name", ["sine", "speech", "person_detect"])
def test_e2e_goldens(models, kats, name):
    k = kats["e2e"][name]
    m = models[name]
    x = np.full(m.in_shape, k["input_fill"], np.float32)
    out = m.predict(x)
    np.testing.assert_array_equal(out.reshape(-1), f32(k["output"]))


def test_e2e_derived_int8_pins(models):
    """quantized outputs / pre-softmax logits for the all-0.5 inputs (derived, not reference-asserted)."""
    m = models["speech"]
    x = np.full(m.in_shape, 0.5, np.float32)
    q_in = np.array([oracle.quantize(0.5, m.in_scale, m.in_zp)] * m.in_elems, np.int8)
    _, q, trace = m.predict_quantized(q_in, return_q=True, trace=True)
    assert q.reshape(-1).tolist() == [-88, -58, -58, -52]
    assert trace[-2].reshape(-1).tolist() == [9, 15, 15, 16]
    m = models["person_detect"]
    q_in = np.array([oracle.quantize(0.5, m.in_scale, m.in_zp)] * m.in_elems, np.int8)
    _, q, trace = m.predict_quantized(q_in, return_q=True, trace=True)
    assert q.reshape(-1).tolist() == [78, -78]
    assert trace[-3].reshape(-1).tolist() == [57, -56]
    _, q = models["sine"].predict(f32([[0.5]]), return_q=True)
    assert q.reshape(-1).tolist() == [57]


def test_sine_accuracy_csv_500_rows(models):
    rows = np.loadtxt(GOLDEN / "sine_microflow.csv", delimiter=",", skiprows=1, dtype=np.float32)
    assert rows.shape == (500, 2)
    m = models["sine"]
    got = np.array([m.predict(f32([[x]]))[0, 0] for x in rows[:, 0]], np.float32)
    np.testing.assert_array_equal(got, rows[:, 1])


def test_samples_classify(models, samples):
    sp, pd_ = models["speech"], models["person_detect"]
    np.testing.assert_array_equal(sp.predict_quantized(samples["YES"]).reshape(-1), f32([0, 0, 0.99609375, 0]))
    np.testing.assert_array_equal(sp.predict_quantized(samples["NO"]).reshape(-1), f32([0, 0.0546875, 0, 0.9453125]))
    np.testing.assert_array_equal(pd_.predict_quantized(samples["PERSON"]).reshape(-1), f32([0.26953125, 0.73046875]))
    np.testing.assert_array_equal(pd_.predict_quantized(samples["NO_PERSON"]).reshape(-1), f32([0.6171875, 0.3828125]))


def test_predict_many_threads_equal(models):
    from conftest import splitmix_bytes
    m = models["speech"]
    xs = splitmix_bytes(0x5EED0002, 16 * m.in_elems).reshape(16, -1)
    a, aq = m.predict_many_quantized(xs, threads=1)
    b, bq = m.predict_many_quantized(xs, threads=4)
    np.testing.assert_array_equal(aq, bq)
    np.testing.assert_array_equal(a, b)
    one = m.predict_quantized(xs[3])
    np.testing.assert_array_equal(one.reshape(-1), a[3])
