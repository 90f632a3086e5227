#!/bin/bash
# BASELINE config 5 on the 3x3 tcgen05 kernel: batch sweep {1,2,4,8,16,32} in the default A-staging mode, plus the
# three-patch mode at the large batches, next to the tcgen05 kind::i8 MMA-only ceiling (tools/ubench/mma_i8).
# Output: gpurun_out/conv2d_sweep.txt
mkdir -p gpurun_out
: > gpurun_out/conv2d_sweep.txt
[ -x tools/ubench/mma_i8 ] && tools/ubench/mma_i8 2000 >> gpurun_out/conv2d_sweep.txt 2>&1
for pm in 1 0; do
  for b in 1 2 4 8 16 32; do
    [ "$pm" = 0 ] && [ "$b" -lt 8 ] && continue
    MF_TC_PATCH=$pm timeout 120 python -m microflow_rs_b200._convbench $b 20 2>&1 | tail -1 > /tmp/cb.json
    python - "$pm" "$b" >> gpurun_out/conv2d_sweep.txt <<'PY'
import json, sys
d = json.loads(open("/tmp/cb.json").read())
r = d["roofline"]
print("MF_TC_PATCH=%s batch %2s  %.4f ms  %7.1f TOP/s  of 2 x bf16 burst %.4f  of nominal 4500 %.4f  verified %s" % (
    sys.argv[1], sys.argv[2], d["ms_per_launch"], r["achieved"], r["frac"], r["frac_of_nominal_4500"], d["verified_vs_generic_kernel"]))
PY
  done
done
cat gpurun_out/conv2d_sweep.txt
