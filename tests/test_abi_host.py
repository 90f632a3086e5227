"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/microflow_cuda.h declares,
fails loudly without a GPU (no CPU fallback), and its run-time loader (the proc-macro's job) reproduces the oracle's --
i.e. the reference's -- pre-processing constants bit for bit.  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import microflow_rs_b200 as mf
import oracle
from conftest import MODELS, ROOT


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "microflow_cuda.h").read_text()
    declared = sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(mf.ABI_SYMBOLS)
    L = mf.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.mf_abi_version() == 3


def _has_gpu():
    try:
        return mf.device_count() > 0
    except mf.MicroflowError:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(mf.MicroflowError) as e:
        mf.device_count()
    assert e.value.status == 10  # MF_ERR_NO_DEVICE
    with pytest.raises(mf.MicroflowError) as e:
        mf.Model(MODELS / "sine.tflite")
    assert e.value.status == 10
    m = mf.Model(MODELS / "sine.tflite", flags=mf.FLAG_HOST_ONLY)
    with pytest.raises(mf.MicroflowError) as e:
        m.predict(np.zeros((1, 1), np.float32))
    assert e.value.status == 10
    with pytest.raises(mf.MicroflowError):
        mf.ops.quantize(np.zeros(4, np.float32), 0.1, 0)


@pytest.mark.parametrize("name", ["sine", "speech", "person_detect"])
def test_loader_matches_oracle(name):
    m = mf.Model(MODELS / f"{name}.tflite", flags=mf.FLAG_HOST_ONLY)
    o = oracle.Model(MODELS / f"{name}.tflite")
    assert m.in_shape == o.in_shape and m.out_shape == o.out_shape
    assert m.in_scale == o.in_scale and m.in_zp == o.in_zp and m.out_scale == o.out_scale and m.out_zp == o.out_zp
    assert len(m.layers) == len(o.layers)
    for i, (a, b) in enumerate(zip(m.layers, o.layers)):
        assert a["op"] == b["op"], i
        assert tuple(a["out_shape"]) == tuple(b["out_shape"]), i
        if a["op"] in ("reshape", "softmax"):
            continue
        c0, c1, c2, c3 = m.layer_constants(i)
        oc0, oc1, oc2, oc3 = o.layer_consts(i)
        if a["op"] == "average_pool_2d":
            assert c0[0] == oc0[0] and c1[0] == oc1[0]
            continue
        np.testing.assert_array_equal(c0, oc0[: len(c0)])
        np.testing.assert_array_equal(c1, oc1[: len(c1)])
        if a["op"] == "fully_connected":
            np.testing.assert_array_equal(c2[: len(c0)], oc2[: len(c0)])
            assert c3 == oc3


def test_person_detect_graph_facts():
    """MAC counts and the ReLU6 clamp evaluating to the full int8 range."""
    m = mf.Model(MODELS / "person_detect.tflite", flags=mf.FLAG_HOST_ONLY)
    assert sum(L["macs"] for L in m.layers) == 7157888
    assert sum(L["macs"] for L in m.layers if L["op"] == "conv_2d") == 6193664
    assert all(L["clamp"] == (-128, 127) for L in m.layers if L["op"] in ("conv_2d", "depthwise_conv_2d"))
    s = mf.Model(MODELS / "speech.tflite", flags=mf.FLAG_HOST_ONLY)
    assert sum(L["macs"] for L in s.layers) == 336000


def test_loader_errors_map_to_reference_diagnostics(tmp_path):
    with pytest.raises(mf.MicroflowError) as e:
        mf.Model(tmp_path / "missing.tflite", flags=mf.FLAG_HOST_ONLY)
    assert e.value.status == 1 and "couldn't find" in e.value.text          # lib.rs:50-55
    with pytest.raises(mf.MicroflowError) as e:
        mf.Model(b"\x00" * 64, flags=mf.FLAG_HOST_ONLY)
    assert e.value.status == 2                                               # lib.rs:56-58
    data = bytearray((MODELS / "sine.tflite").read_bytes())
    with pytest.raises(mf.MicroflowError):
        mf.Model(bytes(data[:200]), flags=mf.FLAG_HOST_ONLY)                 # truncated flatbuffer


def test_model_dump(tmp_path):
    m = mf.Model(MODELS / "speech.tflite", flags=mf.FLAG_HOST_ONLY)
    p = tmp_path / "dump.txt"
    m.dump(p)
    text = p.read_text()
    assert "layer 1: op 4" in text and "c0:" in text


def test_options_struct_is_versioned_by_struct_size():
    """mf_options grew a `layout` field in ABI 2: a 16-byte ABI-1 struct is still accepted, an unknown layout is rejected."""
    import ctypes as C
    L = mf.lib()
    assert L.mf_abi_version() == 3
    data = (MODELS / "sine.tflite").read_bytes()

    class OldOptions(C.Structure):
        _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("chunk", C.c_uint32), ("flags", C.c_uint32)]

    h = C.c_void_p()
    old = OldOptions(C.sizeof(OldOptions), -1, 0, mf.FLAG_HOST_ONLY)
    assert L.mf_model_create_from_tflite(data, len(data), C.cast(C.byref(old), C.POINTER(mf._Options)), C.byref(h)) == 0
    L.mf_model_destroy(h)
    with pytest.raises(mf.MicroflowError) as e:
        mf.Model(MODELS / "sine.tflite", flags=mf.FLAG_HOST_ONLY, layout=7)
    assert e.value.status == 9
    mf.Model(MODELS / "person_detect.tflite", flags=mf.FLAG_HOST_ONLY, layout=mf.LAYOUT_NALGEBRA).close()


def _sass_by_function():
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists():
        pytest.skip("cuobjdump not available")
    mf.lib()                                                           # builds the library if needed
    out = subprocess.run([exe, "-sass", str(mf.LIB_PATH)], capture_output=True, text=True, timeout=600).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
        elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            cur.append(line)
    return funcs


def test_sass_epilogues_are_never_contracted_into_fma():
    """Bit-exactness rests on separate f32 multiply and add (src/ops/conv_2d.rs:93-98).  ptxas contracts mul.rn.f32x2 + add.rn.f32x2
    into FFMA2 even under -fmad=false, so the machine code of every conv / depthwise / fully-connected kernel is
    checked: no FFMA, no FFMA2.  (The pool, softmax, quantize and classifier-tail kernels contain FFMAs inside __fdiv_rn.)"""
    funcs = _sass_by_function()
    hot = {n: body for n, body in funcs.items() if re.search(r"conv_tc_kernel|conv3x3_pair_kernel|fused_chain_kernel|dwconv|pwconv|conv_generic|fc_generic|layout_transpose", n)}
    assert len(hot) >= 20, sorted(funcs)
    for name, body in hot.items():
        bad = [ln for ln in body if re.search(r"\bFFMA2?\b", ln)]
        assert not bad, f"{name}: {bad[:3]}"


def test_sass_shows_the_blackwell_paths():
    """The tcgen05 kernel issues UTCIMMA (tcgen05.mma kind::i8) fed by UTMALDG (TMA tensor loads), reads TMEM with LDTM and stores
    256-bit sectors; the sample-resident depthwise kernels use bulk TMA copies (UBLKCP), IDP.4A and packed FADD2."""
    funcs = _sass_by_function()
    text = {n: "\n".join(b) for n, b in funcs.items()}
    tc = [t for n, t in text.items() if "conv_tc_kernel" in n]
    assert tc and all("UTCIMMA" in t and "UTMALDG" in t and "LDTM" in t for t in tc)
    assert any("STG.E.ENL2.256" in t for t in tc)
    pair = [t for n, t in text.items() if re.search(r"(?<!dw)conv3x3_pair_kernel", n)]
    assert pair and all("UTCIMMA.2CTA" in t and "UTMALDG" in t and "LDTM" in t for t in pair)          # tcgen05.mma.cta_group::2 on CTA pairs
    chain = [t for n, t in text.items() if "fused_chain_kernel" in n]
    assert chain and all("UTCIMMA" in t and "UBLKCP" in t and "LDTM" in t and "IDP.4A" in t and "FADD2" in t for t in chain)
    dw = [t for n, t in text.items() if re.search(r"dwconv3x3_(smem|pair)_kernel|dwconv_cin1_smem_kernel", n)]
    assert dw and all("UBLKCP" in t and "IDP.4A" in t and "FADD2" in t for t in dw)


def test_sass_tensor_core_issue_is_not_serialised():
    """The MMA / TMA issuing thread is picked with elect.sync (tcptx::elect_one).  With `lane == 0` the compiler wraps every UTCIMMA and
    UTMALDG in an ELECT / BRA.U.ANY loop (~10 extra instructions per MMA, profiles/r02g_conv3x3_experiments.txt): pin their absence."""
    funcs = _sass_by_function()
    tc = {n: b for n, b in funcs.items() if re.search(r"conv_tc_kernel|(?<!dw)conv3x3_pair_kernel", n)}
    assert len(tc) >= 20
    for n, body in tc.items():
        assert not [ln for ln in body if "BRA.U.ANY" in ln], n
    for n, body in funcs.items():
        if "fused_chain_kernel" in n:          # its two remaining loops belong to the once-per-unit bulk copies, none to the MMAs
            assert sum("BRA.U.ANY" in ln for ln in body) <= 2, n
            mma = [i for i, ln in enumerate(body) if "UTCIMMA" in ln]
            assert mma and not any("BRA.U.ANY" in ln for ln in body[mma[0]:mma[-1]]), n


def test_sass_packed_float_to_int8_conversion():
    """f2i_pack4 (mf_device.cuh) relies on ptxas fusing {cvt.rzi.s32.f32 x2, cvt.pack.sat.s8.s32} into ONE F2IP.S8.F32.TRUNC.NTZ (two
    values per instruction, not on the quarter-rate XU pipe).  If a toolchain stops doing that the epilogues silently fall back to
    F2I.S32 + I2IP (five times the cost): pin it.  No XU-epilogue kernel may contain the old F2I.S8 either."""
    funcs = _sass_by_function()
    text = {n: "\n".join(b) for n, b in funcs.items()}
    hot = {n: t for n, t in text.items()
           if re.search(r"fused_chain_kernel|dwconv3x3_(smem|pair)_kernel|dwconv_cin1_(smem|taps)_kernel", n)
           or re.search(r"(conv_tc_kernel|(?<!dw)conv3x3_pair_kernel)ILb[01]ELb1E", n)}      # XU = true instantiations
    assert len(hot) >= 30, sorted(text)
    for n, t in hot.items():
        assert "F2IP.S8.F32.TRUNC.NTZ" in t, n
        assert "F2I.S8" not in t and "I2IP" not in t, n


def test_device_list_options_are_validated_without_a_gpu():
    """ABI 3: mf_options.n_devices / devices[].  A 20-byte ABI-2 struct (no device list) is still accepted; a list longer than
    MF_MAX_DEVICES is refused; with MF_FLAG_HOST_ONLY (parse + preprocess only) the device list is not consulted."""
    L = mf.lib()
    data = (MODELS / "sine.tflite").read_bytes()

    class Abi2Options(C.Structure):
        _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("chunk", C.c_uint32), ("flags", C.c_uint32), ("layout", C.c_uint32)]

    h = C.c_void_p()
    old = Abi2Options(C.sizeof(Abi2Options), -1, 0, mf.FLAG_HOST_ONLY, 0)
    assert L.mf_model_create_from_tflite(data, len(data), C.cast(C.byref(old), C.POINTER(mf._Options)), C.byref(h)) == 0
    assert L.mf_model_devices(h, None, 0) == 0            # host-only: runs nowhere
    assert L.mf_model_weight_broadcast(h) == b"none"
    L.mf_model_destroy(h)
    opt = mf._Options(C.sizeof(mf._Options), -1, 0, mf.FLAG_HOST_ONLY, 0)
    opt.n_devices = mf.MAX_DEVICES + 1
    assert L.mf_model_create_from_tflite(data, len(data), C.byref(opt), C.byref(h)) == 9      # MF_ERR_INVALID_ARG
    m = mf.Model(MODELS / "sine.tflite", flags=mf.FLAG_HOST_ONLY, devices=[0, 1, 2])
    assert m.devices == [] and m.launched_kernels() == [""] * len(m.layers)
    m.close()


def test_bmp_staging_reproduces_the_reference_features(samples):
    """samples/person.bmp and no_person.bmp (byte-identical copies under tests/golden/) -> exactly the int8 tensors the reference ships as
    features::PERSON / NO_PERSON (samples/features/person_detect.rs:5,104).  Host-side staging: needs no GPU."""
    from conftest import GOLDEN
    for f, key in (("person.bmp", "PERSON"), ("no_person.bmp", "NO_PERSON")):
        got = mf.features_from_bmp((GOLDEN / f).read_bytes())
        assert got.shape == (96, 96, 1) and got.dtype == np.int8
        np.testing.assert_array_equal(got.reshape(-1), np.asarray(samples[key]).reshape(-1))
    with pytest.raises(mf.MicroflowError):
        mf.features_from_bmp(b"BM" + b"\0" * 20)
    with pytest.raises(mf.MicroflowError):
        mf.features_from_bmp((GOLDEN / "person.bmp").read_bytes()[:5000])       # truncated pixel array


def _patched_model(name, old_shape, new_shape, which=0):
    """Returns the .tflite bytes with one int32 shape vector replaced (flatbuffer vector = length word + elements)."""
    data = bytearray((MODELS / f"{name}.tflite").read_bytes())
    pat = np.array([len(old_shape)] + list(old_shape), np.int32).tobytes()
    hits = [i for i in range(len(data) - len(pat)) if data[i:i + len(pat)] == pat]
    assert len(hits) > which, (name, old_shape, hits)
    data[hits[which] + 4: hits[which] + 4 + 4 * len(new_shape)] = np.array(new_shape, np.int32).tobytes()
    return bytes(data)


def test_loader_rejects_batched_and_overflowing_shapes():
    """The reference's ops take Tensor4D<T, 1, ...> (BATCHES = 1, src/ops/conv_2d.rs:40): a model whose op tensors carry a batch dimension
    other than 1 would not type-check there and must not be silently mis-strided here; dimensions are untrusted, so products that
    overflow are refused instead of wrapping to a small element count."""
    # layer 0's output / layer 1's input [1,48,48,8] -> batch 2: element counts no longer chain, or the batch check fires
    for which in (0, 1):
        with pytest.raises(mf.MicroflowError) as e:
            mf.Model(_patched_model("person_detect", [1, 48, 48, 8], [2, 48, 48, 8], which), flags=mf.FLAG_HOST_ONLY)
        assert e.value.status in (7, 2), e.value
    # the model input [1,96,96,1] -> batch 2 with layer 0 declaring the same: must be refused as an unsupported shape
    with pytest.raises(mf.MicroflowError) as e:
        mf.Model(_patched_model("person_detect", [1, 96, 96, 1], [2, 96, 96, 1]), flags=mf.FLAG_HOST_ONLY)
    assert e.value.status == 7, e.value
    # absurd dimensions: 2^31-1 squared overflows any size_t product check that is not done step by step
    with pytest.raises(mf.MicroflowError) as e:
        mf.Model(_patched_model("person_detect", [1, 96, 96, 1], [1, 2147483647, 2147483647, 1]), flags=mf.FLAG_HOST_ONLY)
    assert e.value.status == 7, e.value
    with pytest.raises(mf.MicroflowError):
        mf.Model(_patched_model("person_detect", [1, 96, 96, 1], [1, -96, 96, 1]), flags=mf.FLAG_HOST_ONLY)


def test_python_mirror_validates_caller_supplied_output_arrays():
    """A wrong-sized, wrong-typed or non-contiguous `out=` must be refused before its raw pointer reaches the C side."""
    m = mf.Model(MODELS / "sine.tflite", flags=mf.FLAG_HOST_ONLY)
    xs = np.zeros((4, 1), np.int8)
    for bad in (np.zeros((3, 1), np.float32), np.zeros((4, 1), np.float64), np.zeros((4, 2), np.float32)[:, ::2]):
        with pytest.raises(ValueError):
            m.predict_many_quantized(xs, out=bad)
        with pytest.raises(ValueError):
            m.predict_many_quantized_async(xs, bad)
    with pytest.raises(ValueError):
        m.predict_many_quantized_async(np.zeros((4, 1), np.float32), np.zeros((4, 1), np.float32))
    m.close()
