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
print("workload ok")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/mf_sanitize_workload.py > gpurun_out/sanitize_$tool.txt 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|workload ok|Error|hazard" gpurun_out/sanitize_$tool.txt | head -12
done
