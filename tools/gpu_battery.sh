#!/bin/bash
# Standard GPU session (run under gpurun): parity tests, smoke, benches, ncu launch list + full captures.
# Usage: tools/gpu_battery.sh [tests|bench|prof|all]...   (default: all).  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
what="${*:-all}"
has() { [[ " $what " == *" $1 "* || " $what " == *" all "* ]]; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
if has tests; then
  timeout 500 python -m pytest tests/test_gpu_tc.py -q -m gpu 2>&1 | tail -25 > gpurun_out/t_tc.log
  timeout 700 python -m pytest tests/test_gpu_ops.py -q -m gpu 2>&1 | tail -40 > gpurun_out/t_ops.log
  timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu 2>&1 | tail -60 > gpurun_out/t_models.log
  timeout 300 python -m pytest tests/test_gpu_multidevice.py -q -m gpu 2>&1 | tail -20 > gpurun_out/t_multi.log
  timeout 900 python -m pytest tests/test_gpu_fused.py -q -m gpu 2>&1 | tail -20 > gpurun_out/t_fused.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
fi
if has bench; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_pd.json 2> gpurun_out/bench_pd.err
  timeout 300 python bench.py --workload speech --steps 20 --warmup 5 --no-conv2d > gpurun_out/bench_speech.json 2> gpurun_out/bench_speech.err
  for c in 1024 2048 8192; do
    timeout 200 python bench.py --steps 10 --warmup 3 --chunk $c --no-cpu-baseline --no-conv2d > gpurun_out/bench_pd_chunk$c.json 2>> gpurun_out/bench_pd.err
  done
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
fi
if has prof; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-conv2d > gpurun_out/ncu_bench.log 2>&1
  # full captures of single layers at batch 8192 (tools/layer_bench.py launches one kernel type per process: deterministic)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3 -s 10 -c 1 -f -o gpurun_out/prof_dw \
      python tools/layer_bench.py 8192 L1_dw8 > gpurun_out/ncu_dw.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3 -s 10 -c 1 -f -o gpurun_out/prof_dw_s2 \
      python tools/layer_bench.py 8192 L3_dw16_s2 > gpurun_out/ncu_dw_s2.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 10 -c 1 -f -o gpurun_out/prof_tc_pw \
      python tools/layer_bench.py 8192 L6_pw32_32 > gpurun_out/ncu_tc.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 10 -c 1 -f -o gpurun_out/prof_tc_pw128 \
      python tools/layer_bench.py 8192 L14_pw128 > gpurun_out/ncu_tc128.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_chain -s 3 -c 1 -f -o gpurun_out/prof_fused \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-conv2d > gpurun_out/ncu_fused.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair -s 5 -c 1 -f -o gpurun_out/prof_conv3x3_pair \
      python -m microflow_rs_b200._convbench 16 4 > gpurun_out/ncu_conv3x3_pair.log 2>&1
  MF_TC_PAIR=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 5 -c 1 -f -o gpurun_out/prof_conv3x3 \
      python -m microflow_rs_b200._convbench 16 4 > gpurun_out/ncu_conv3x3.log 2>&1
fi
tail -n 4 gpurun_out/t_tc.log gpurun_out/t_ops.log gpurun_out/t_models.log gpurun_out/t_multi.log gpurun_out/t_fused.log gpurun_out/smoke.log 2>/dev/null
for f in gpurun_out/bench_pd.json gpurun_out/bench_speech.json gpurun_out/bench_pd_chunk*.json gpurun_out/bench_ref.json; do
  [ -f "$f" ] && python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    c = (d.get("conv2d") or {}).get("roofline") or {}
    print(sys.argv[1], "value=%.4g e2e=%.4g ms/step=%.3f dom=%s frac=%.3f conv2d=%s clocks=%s" % (
        d["value"], (d.get("e2e") or {}).get("value", 0), d["ms_per_step"], r.get("kernel"), r.get("frac", 0), c.get("frac"), d.get("clocks")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -n 3 gpurun_out/*.err 2>/dev/null | tail -20
if has cpuscale; then
  python - > gpurun_out/cpu_scaling.txt 2>&1 <<'PY'
import sys, time, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, oracle
from conftest import MODELS, splitmix_bytes
o = oracle.Model(MODELS / "person_detect.tflite", fast=True)
print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for th in (1, 4, 16, 32, 64, 128):
    n = 16 * th
    xs = splitmix_bytes(3, n * o.in_elems).reshape(n, -1)
    t = time.perf_counter(); o.predict_many_quantized(xs, threads=th); dt = time.perf_counter() - t
    print(th, "threads:", n / dt, "inf/s", (n / dt) / th, "per thread")
PY
  cat gpurun_out/cpu_scaling.txt
fi
if has exp; then
  : > gpurun_out/exp.txt
  for x in 0 5 8; do
    echo "== conv3x3 MF_TC_XUG=$x" >> gpurun_out/exp.txt
    MF_TC_XUG=$x timeout 200 python -m microflow_rs_b200._convbench 16 10 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_launch'], d['roofline']['frac'], d['verified_vs_generic_kernel'])" >> gpurun_out/exp.txt 2>&1
  done
  for dw in 0 2 4; do for tc in 0 5 8; do
    echo "== person_detect MF_DW_XU=$dw MF_TC_XUG=$tc" >> gpurun_out/exp.txt
    MF_DW_XU=$dw MF_TC_XUG=$tc timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-conv2d 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g' % d['value'], {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items() if v['ms_per_step']>0.1})" >> gpurun_out/exp.txt 2>&1
  done; done
  cat gpurun_out/exp.txt
fi
if has layers; then
  timeout 300 python tools/layer_bench.py 8192 > gpurun_out/layer_bench.txt 2>&1
  MF_DW_NO_SMEM=1 timeout 200 python tools/layer_bench.py 8192 L1_dw8,L5_dw32,L13_dw128 >> gpurun_out/layer_bench.txt 2>&1
  MF_TC_XUG=0 timeout 200 python tools/layer_bench.py 8192 L2_pw8_16,L6_pw32_32,L14_pw128 >> gpurun_out/layer_bench.txt 2>&1
  MF_TC_XUG=5 timeout 200 python tools/layer_bench.py 8192 L2_pw8_16,L6_pw32_32,L14_pw128 >> gpurun_out/layer_bench.txt 2>&1
  cat gpurun_out/layer_bench.txt
fi
if has prof2; then
  timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:conv_tc -s 10 -c 1 -f -o gpurun_out/prof_pw_nc \
      python tools/layer_bench.py 8192 L2_pw8_16 > gpurun_out/ncu_pw_nc.log 2>&1
  timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:dwconv3x3_smem -s 10 -c 1 -f -o gpurun_out/prof_dw_nc \
      python tools/layer_bench.py 8192 L1_dw8 > gpurun_out/ncu_dw_nc.log 2>&1
  timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:cin1 -s 10 -c 1 -f -o gpurun_out/prof_cin1_nc \
      python tools/layer_bench.py 8192 L0_dw_cin1 > gpurun_out/ncu_cin1_nc.log 2>&1
  tail -n 2 gpurun_out/ncu_pw_nc.log gpurun_out/ncu_dw_nc.log gpurun_out/ncu_cin1_nc.log
fi
