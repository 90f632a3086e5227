#!/bin/bash
# compute-sanitizer pass over the hot kernels (run under gpurun): memcheck, racecheck and synccheck on a workload that launches every
# fast kernel once at a batch that selects the sample-resident / tcgen05 variants.  Summaries land in gpurun_out/sanitize_*.txt.
set -u
mkdir -p gpurun_out
cat > /tmp/mf_sanitize_workload.py <<'PY'
import sys
from pathlib import Path
import numpy as np
ROOT = Path.cwd()
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import microflow_rs_b200 as mf
from conftest import MODELS, splitmix_bytes
for name, n in (("person_detect", 300), ("speech", 40), ("sine", 70)):
    m = mf.Model(MODELS / f"{name}.tflite", chunk=512)
    xs = splitmix_bytes(7, n * m.in_elems).reshape(n, -1)
    a = m.predict_many_quantized(xs)                  # stream path (PDL on)
    b = m.predict_many_quantized(xs[:8])              # CUDA-graph path
    assert np.array_equal(a[:8], b), name
    g = mf.Model(MODELS / f"{name}.tflite", flags=mf.FLAG_FORCE_GENERIC)
    assert np.array_equal(g.predict_many_quantized(xs[:16]), a[:16]), name
    m.close(); g.close()
# fused low-resolution chain at a batch with partial units, the CTA-pair 3x3 kernel, a two-replica multi-device model
import os
r = np.random.default_rng(1)
sys.path.insert(0, str(ROOT / "tests"))
import fused_check
layers = fused_check.make_chain(r, 6, 6, 5)
x = r.integers(-128, 128, (23, 6, 6, 128)).astype(np.int8)
assert np.array_equal(mf.ops.conv_chain(x, layers, fuse=True), mf.ops.conv_chain(x, layers, fuse=False))
# (the CTA-pair kernel is the default for this shape; MF_TC_* switches are latched at the first tcgen05 launch, so they must be set before the process starts)
x3 = r.integers(-128, 128, (2, 24, 16, 128)).astype(np.int8)
w3 = r.integers(-128, 128, (128, 3, 3, 128)).astype(np.int8)
c1 = (r.uniform(0.2, 2.0, 128) / 40000.0).astype(np.float32); c0 = r.uniform(-20, 20, 128).astype(np.float32)
a3 = mf.ops.conv_2d(x3, -128, w3, [0], 0.0235294, -128, "relu6", "same", (1, 1), c0, c1, (24, 16), impl=0)
assert "conv3x3_pair_kernel" in mf.ops.last_kernel, mf.ops.last_kernel
assert np.array_equal(a3, mf.ops.conv_2d(x3, -128, w3, [0], 0.0235294, -128, "relu6", "same", (1, 1), c0, c1, (24, 16), impl=1))
# the one-CTA 3x3 kernel (Cout = 32 is not splittable over a CTA pair)
w4 = r.integers(-128, 128, (32, 3, 3, 128)).astype(np.int8)
a4 = mf.ops.conv_2d(x3, -128, w4, [0], 0.0235294, -128, "relu6", "same", (1, 1), c0[:32], c1[:32], (24, 16), impl=0)
assert "conv_tc_kernel" in mf.ops.last_kernel, mf.ops.last_kernel
assert np.array_equal(a4, mf.ops.conv_2d(x3, -128, w4, [0], 0.0235294, -128, "relu6", "same", (1, 1), c0[:32], c1[:32], (24, 16), impl=1))
os.environ["MF_ALLOW_DUPLICATE_DEVICES"] = "1"
g = mf.Model(MODELS / "speech.tflite", devices=[0, 0])
one = mf.Model(MODELS / "speech.tflite")
xs = splitmix_bytes(9, 300 * g.in_elems).reshape(300, -1)
assert np.array_equal(g.predict_many_quantized(xs), one.predict_many_quantized(xs))
g.close(); one.close()
print("workload ok")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/mf_sanitize_workload.py > gpurun_out/sanitize_$tool.txt 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|workload ok|Error|hazard" gpurun_out/sanitize_$tool.txt | head -12
done
