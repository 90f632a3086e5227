#!/bin/bash
# BASELINE config 5 on the 3x3 tcgen05 kernel: A-staging mode x batch.  Output: gpurun_out/conv2d_sweep.txt
mkdir -p gpurun_out
: > gpurun_out/conv2d_sweep.txt
for pm in 1 0; do
  for b in 8 16 32; do
    MF_TC_PATCH=$pm timeout 120 python -m microflow_rs_b200._convbench $b 20 2>&1 | tail -1 > /tmp/cb.json
    python - "$pm" "$b" >> gpurun_out/conv2d_sweep.txt <<'PY'
import json, sys
d = json.loads(open("/tmp/cb.json").read())
print("MF_TC_PATCH=%s batch %s  %.4f ms  frac %.4f  verified %s" % (sys.argv[1], sys.argv[2], d["ms_per_launch"], d["roofline"]["frac"], d["verified_vs_generic_kernel"]))
PY
  done
done
cat gpurun_out/conv2d_sweep.txt
