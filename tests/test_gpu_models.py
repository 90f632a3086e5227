"""GPU parity, model level, through the reference-shaped API (predict / predict_quantized / predict_many) of the C-ABI:
the reference's end-to-end goldens, the 500-row sine CSV, the sample inputs, and seeded random batches checked against
the oracle layer by layer (trace), at the final quantized output and at the pre-softmax logits."""
import numpy as np
import pytest

import microflow_rs_b200 as mf
import oracle
from conftest import GOLDEN, MODELS, f32, splitmix_bytes

pytestmark = pytest.mark.gpu
NAMES = ["sine", "speech", "person_detect"]
SEEDS = {"sine": 0x5EED0001, "speech": 0x5EED0002, "person_detect": 0x5EED0003}


@pytest.fixture(scope="module")
def gpu_models():
    ms = {n: mf.Model(MODELS / f"{n}.tflite") for n in NAMES}
    yield ms
    for m in ms.values():
        m.close()


@pytest.fixture(scope="module")
def ora():
    return {n: oracle.Model(MODELS / f"{n}.tflite", fast=True) for n in NAMES}


@pytest.mark.parametrize("name", NAMES)
def test_e2e_goldens(gpu_models, kats, name):
    k = kats["e2e"][name]
    m = gpu_models[name]
    out = m.predict(np.full(m.in_shape, k["input_fill"], np.float32))
    np.testing.assert_array_equal(out.reshape(-1), f32(k["output"]))


def test_sine_accuracy_csv_500_rows(gpu_models):
    rows = np.loadtxt(GOLDEN / "sine_microflow.csv", delimiter=",", skiprows=1, dtype=np.float32)
    m = gpu_models["sine"]
    one_by_one = np.array([m.predict(f32([[x]]))[0, 0] for x in rows[:50, 0]], np.float32)
    np.testing.assert_array_equal(one_by_one, rows[:50, 1])
    batched = m.predict_many(np.ascontiguousarray(rows[:, 0]))
    np.testing.assert_array_equal(batched.reshape(-1), rows[:, 1])


def test_samples(gpu_models, samples):
    sp, pd_ = gpu_models["speech"], gpu_models["person_detect"]
    np.testing.assert_array_equal(sp.predict_quantized(samples["YES"]).reshape(-1), f32([0, 0, 0.99609375, 0]))
    np.testing.assert_array_equal(sp.predict_quantized(samples["NO"]).reshape(-1), f32([0, 0.0546875, 0, 0.9453125]))
    np.testing.assert_array_equal(pd_.predict_quantized(samples["PERSON"]).reshape(-1), f32([0.26953125, 0.73046875]))
    np.testing.assert_array_equal(pd_.predict_quantized(samples["NO_PERSON"]).reshape(-1), f32([0.6171875, 0.3828125]))


@pytest.mark.parametrize("name,n", [("sine", 64), ("speech", 24), ("person_detect", 6)])
@pytest.mark.parametrize("flags", [0, mf.FLAG_NO_TENSOR_CORE, mf.FLAG_FORCE_GENERIC])
def test_trace_every_layer_vs_oracle(ora, name, n, flags):
    o = ora[name]
    m = mf.Model(MODELS / f"{name}.tflite", flags=flags)
    try:
        xs = splitmix_bytes(SEEDS[name], n * o.in_elems).reshape(n, -1)
        tr = m.predict_trace(xs)
        for s in range(n):
            _, q, ot = o.predict_quantized(xs[s], return_q=True, trace=True)
            for i, (a, b) in enumerate(zip(tr, ot)):
                assert np.array_equal(a[s], b.reshape(-1)), f"{name} sample {s}: layer {i} ({m.layers[i]['op']}, {m.layers[i]['kernel']}) differs"
    finally:
        m.close()


@pytest.mark.parametrize("name,n", [("sine", 3000), ("speech", 600), ("person_detect", 300)])
def test_predict_many_vs_oracle(gpu_models, ora, name, n):
    m, o = gpu_models[name], ora[name]
    xs = splitmix_bytes(SEEDS[name], n * o.in_elems).reshape(n, -1)
    want_f, want_q = o.predict_many_quantized(xs, threads=oracle.max_threads())
    got_f = m.predict_many_quantized(xs)
    np.testing.assert_array_equal(got_f, want_f)
    got_q, logits = m.predict_many_logits(xs)
    np.testing.assert_array_equal(got_q, want_q)
    if logits is not None:   # pre-softmax logits: the part of the output that is pinned independently of libm expf
        idx = max(i for i, L in enumerate(o.layers) if L["op"] == "softmax")
        for s in range(0, n, max(1, n // 16)):
            _, _, tr = o.predict_quantized(xs[s], return_q=True, trace=True)
            np.testing.assert_array_equal(logits[s], tr[idx - 1].reshape(-1))


def test_f32_predict_many_equals_quantize_then_predict(gpu_models, ora):
    m, o = gpu_models["speech"], ora["speech"]
    r = np.random.default_rng(5)
    xf = r.uniform(-14, 14, (40, o.in_elems)).astype(np.float32)
    got = m.predict_many(xf)
    want = np.stack([o.predict(xf[s]).reshape(-1) for s in range(40)])
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("name,n", [("speech", 4096), ("person_detect", 4099)])
def test_full_size_fast_equals_generic_and_chunking_is_invisible(name, n):
    """At (near) BASELINE batch sizes the oracle is too slow; use size-independent properties instead: the fast path equals the
    generic cross-check path bit for bit, results do not depend on the chunk size, and duplicated rows give duplicated outputs."""
    base = splitmix_bytes(SEEDS[name] + 7, 1024 * oracle.Model(MODELS / f"{name}.tflite").in_elems).reshape(1024, -1)
    xs = np.concatenate([base] * (n // 1024) + [base[: n % 1024]])
    a = mf.Model(MODELS / f"{name}.tflite", chunk=2048)
    b = mf.Model(MODELS / f"{name}.tflite", chunk=333, flags=mf.FLAG_FORCE_GENERIC)
    try:
        qa, la = a.predict_many_logits(xs)
        qb, lb = b.predict_many_logits(xs)
        np.testing.assert_array_equal(qa, qb)
        np.testing.assert_array_equal(la, lb)
        np.testing.assert_array_equal(qa[:1024], qa[1024:2048])
        assert a.launch_count() > 0 and any("conv_tc" in L["kernel"] for L in a.layers) == (name == "person_detect")
    finally:
        a.close(); b.close()


def test_device_resident_api(gpu_models, ora):
    torch = pytest.importorskip("torch")
    m, o = gpu_models["person_detect"], ora["person_detect"]
    n = 64
    xs = splitmix_bytes(0x5EED0004, n * o.in_elems).reshape(n, -1)
    d_in = torch.from_numpy(xs.copy()).cuda()
    d_out = torch.empty((n, o.out_elems), dtype=torch.float32, device="cuda")
    d_q = torch.empty((n, o.out_elems), dtype=torch.int8, device="cuda")
    m.predict_many_device(d_in.data_ptr(), n, d_out.data_ptr(), d_q.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want_f, want_q = o.predict_many_quantized(xs, threads=oracle.max_threads())
    np.testing.assert_array_equal(d_out.cpu().numpy(), want_f)
    np.testing.assert_array_equal(d_q.cpu().numpy(), want_q)


def test_async_predict_many_pipelines_and_matches_blocking(gpu_models, ora):
    """mf_predict_many_quantized_async: several enqueued calls on pinned buffers, one synchronize, same bits as the blocking call."""
    m, o = gpu_models["person_detect"], ora["person_detect"]
    n = 700
    ins = [mf.PinnedBuffer((n, o.in_elems), np.int8) for _ in range(3)]
    outs = [mf.PinnedBuffer((n, o.out_elems), np.float32) for _ in range(3)]
    for k, b in enumerate(ins):
        b.array[:] = splitmix_bytes(0x5EED0100 + k, n * o.in_elems).reshape(n, -1)
    for k in range(3):
        m.predict_many_quantized_async(ins[k].array, outs[k].array)
    m.synchronize()
    for k in range(3):
        np.testing.assert_array_equal(outs[k].array, m.predict_many_quantized(ins[k].array))
    want, _ = o.predict_many_quantized(ins[1].array[:64], threads=oracle.max_threads())
    np.testing.assert_array_equal(outs[1].array[:64], want)


@pytest.mark.parametrize("name", NAMES)
def test_small_calls_replay_a_cuda_graph_and_match_the_oracle(gpu_models, ora, name):
    """Host-path calls of <= 64 samples (the reference's one-sample predict() above all) replay a captured CUDA graph: the same
    graph must serve changing inputs, sizes either side of the limit must agree, and every variant (f32 in, logits out) is checked."""
    m, o = gpu_models[name], ora[name]
    xs = splitmix_bytes(SEEDS[name] + 11, 70 * o.in_elems).reshape(70, -1)
    want_f, want_q = o.predict_many_quantized(xs, threads=oracle.max_threads())
    for rep in range(3):                                   # capture on the first pass, replays afterwards
        for s in (0, 1, 69):
            np.testing.assert_array_equal(m.predict_quantized(xs[s]).reshape(-1), want_f[s])
    for n in (1, 2, 17, 64, 65, 70):                       # 65 and 70 take the stream path
        np.testing.assert_array_equal(m.predict_many_quantized(xs[:n]), want_f[:n])
        q, _ = m.predict_many_logits(xs[:n])
        np.testing.assert_array_equal(q, want_q[:n])
    xf = np.random.default_rng(3).uniform(-2, 2, (5, o.in_elems)).astype(np.float32)
    want = np.stack([o.predict(xf[s]).reshape(-1) for s in range(5)])
    for rep in range(2):
        np.testing.assert_array_equal(m.predict_many(xf), want)
        np.testing.assert_array_equal(m.predict(xf[3]).reshape(-1), want[3])


def test_nalgebra_layout_matches_nhwc(ora, samples):
    """mf_options.layout = MF_LAYOUT_NALGEBRA: host buffers in the reference's own column-major Buffer4D order ([col][row][chan],
    src/buffer.rs:10-16) give the same results as NHWC -- through the graph path (1 sample), the stream path, f32 and int8."""
    o = ora["person_detect"]
    a = mf.Model(MODELS / "person_detect.tflite")
    b = mf.Model(MODELS / "person_detect.tflite", layout=mf.LAYOUT_NALGEBRA)
    try:
        n = 130
        xs = splitmix_bytes(0x5EED0033, n * o.in_elems).reshape(n, 96, 96, 1)
        xs[0] = np.asarray(samples["PERSON"]).reshape(96, 96, 1)
        xt = np.ascontiguousarray(xs.transpose(0, 2, 1, 3))                      # [sample][col][row][chan]
        want = a.predict_many_quantized(xs.reshape(n, -1))
        np.testing.assert_array_equal(b.predict_many_quantized(xt.reshape(n, -1)), want)
        np.testing.assert_array_equal(b.predict_quantized(xt[0].reshape(-1)).reshape(-1), f32([0.26953125, 0.73046875]))   # tests/person_detect.rs sample
        qa, la = a.predict_many_logits(xs.reshape(n, -1)[:7])
        qb, lb = b.predict_many_logits(xt.reshape(n, -1)[:7])
        np.testing.assert_array_equal(qa, qb)
        np.testing.assert_array_equal(la, lb)
        xf = np.random.default_rng(9).uniform(-1, 1, (3, 96, 96, 1)).astype(np.float32)
        np.testing.assert_array_equal(b.predict_many(np.ascontiguousarray(xf.transpose(0, 2, 1, 3)).reshape(3, -1)), a.predict_many(xf.reshape(3, -1)))
    finally:
        a.close(); b.close()


@pytest.mark.parametrize("shape", [(3, 5, 7, 1), (2, 96, 96, 1), (4, 3, 3, 256), (1, 1, 9, 4), (2, 6, 1, 8)])
@pytest.mark.parametrize("dtype", [np.int8, np.float32])
def test_layout_transpose_kernel_both_directions(shape, dtype):
    r = np.random.default_rng(sum(shape))
    x = (r.integers(-128, 128, shape).astype(dtype) if dtype == np.int8 else r.standard_normal(shape).astype(dtype))
    t = mf.ops.layout_transpose(x, to_nalgebra=True)
    np.testing.assert_array_equal(t, x.transpose(0, 2, 1, 3))
    np.testing.assert_array_equal(mf.ops.layout_transpose(t, to_nalgebra=False), x)


def test_device_resident_full_batch_splits_into_two_streams(gpu_models, ora):
    """A device-resident call of >= 4096 samples per chunk runs as two half-chunks on the model's two internal streams, joined on
    the caller's stream: the result must equal the host path (oracle-checked above) row for row, and be ordered on that stream."""
    torch = pytest.importorskip("torch")
    m, o = gpu_models["person_detect"], ora["person_detect"]
    n = 8192 + 9
    base = splitmix_bytes(0x5EED0044, 1031 * o.in_elems).reshape(1031, -1)
    xs = np.concatenate([base] * (n // 1031) + [base[: n % 1031]])
    want = m.predict_many_quantized(xs)
    ref, _ = o.predict_many_quantized(xs[:32], threads=oracle.max_threads())
    np.testing.assert_array_equal(want[:32], ref)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d_in = torch.from_numpy(xs.copy()).cuda()
        d_out = torch.zeros((n, o.out_elems), dtype=torch.float32, device="cuda")
        for _ in range(3):                                   # back-to-back calls reuse both streams' buffers
            m.predict_many_device(d_in.data_ptr(), n, d_out.data_ptr(), None, st.cuda_stream)
        got = d_out.cpu().numpy()                            # ordered behind the calls on the same stream
    np.testing.assert_array_equal(got, want)


def test_predict_from_bmp_samples(gpu_models):
    """The step before the path: the reference's image samples as files -> staged -> person_detect, same outputs as its
    precomputed feature tensors give (examples/person_detect.rs:27-28 classes)."""
    from conftest import GOLDEN
    out = gpu_models["person_detect"].predict_many_bmp([(GOLDEN / "person.bmp").read_bytes(), (GOLDEN / "no_person.bmp").read_bytes()])
    np.testing.assert_array_equal(out, f32([[0.26953125, 0.73046875], [0.6171875, 0.3828125]]))
