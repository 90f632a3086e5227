"""Fused low-resolution stage (mf_fused.cu): bit-exact against the oracle and against the layer-by-layer kernels, at op level on
random chains and at model level on person_detect (layers 13-22 run as one launch) with the fusion on and off."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import microflow_rs_b200 as mf
import oracle
from conftest import MODELS, splitmix_bytes

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent


def _run(which, timeout=600):
    p = subprocess.run([sys.executable, str(HERE / "fused_check.py"), which], capture_output=True, text=True, timeout=timeout)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, f"no result (rc={p.returncode})\nstdout:\n{p.stdout[-2000:]}\nstderr:\n{p.stderr[-3000:]}"
    res = json.loads(lines[-1])
    assert res["ok"], res.get("error")
    return res


def test_fused_chain_shapes_vs_oracle():
    assert _run("shapes")["cases"] == 12


def test_fused_chain_full_grid_vs_layerwise():
    _run("large")


def test_person_detect_runs_the_fused_chain_and_matches_the_oracle():
    o = oracle.Model(MODELS / "person_detect.tflite", fast=True)
    m = mf.Model(MODELS / "person_detect.tflite")
    try:
        names = [L["kernel"] for L in m.layers]
        assert names[13] == "fused_chain_kernel" and all(n == "(in fused_chain_kernel)" for n in names[14:23]), names
        for n in (1, 6, 7, 8, 65, 300):
            xs = splitmix_bytes(0x5EED0003 + n, n * o.in_elems).reshape(n, -1)
            want_f, want_q = o.predict_many_quantized(xs, threads=oracle.max_threads())
            q, _ = m.predict_many_logits(xs)
            np.testing.assert_array_equal(q, want_q)
            np.testing.assert_array_equal(m.predict_many_quantized(xs), want_f)
            launched = m.launched_kernels()
            assert launched[13] == "fused_chain_kernel" and all(k == "" for k in launched[14:23]), launched
    finally:
        m.close()


CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import microflow_rs_b200 as mf
from conftest import MODELS, splitmix_bytes
m = mf.Model(MODELS / "person_detect.tflite")
n = 4099
base = splitmix_bytes(0x5EED0077, 1031 * m.in_elems).reshape(1031, -1)
xs = np.concatenate([base] * (n // 1031) + [base[: n %% 1031]])
q, l = m.predict_many_logits(xs)
np.save(sys.argv[1], np.concatenate([q.reshape(n, -1), l.reshape(n, -1)], axis=1))
print(json.dumps({"kernels": [L["kernel"] for L in m.layers]}))
"""


def test_fusion_on_equals_fusion_off_at_full_size(tmp_path):
    """Size-independent property at (half) BASELINE batch: the model gives the same int8 outputs and logits with the chain fused
    (default) and with MF_NO_CHAIN_FUSE=1 (every layer on its own kernel, oracle-checked by the other tests)."""
    code = CHILD % {"root": str(HERE.parent), "tests": str(HERE)}
    outs = []
    for tag, env in (("on", {}), ("off", {"MF_NO_CHAIN_FUSE": "1"})):
        f = tmp_path / f"{tag}.npy"
        p = subprocess.run([sys.executable, "-c", code, str(f)], capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
        assert p.returncode == 0, p.stderr[-3000:]
        kern = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])["kernels"]
        assert ("fused_chain_kernel" in kern) == (tag == "on"), kern
        outs.append(np.load(f))
    np.testing.assert_array_equal(outs[0], outs[1])
