#!/bin/bash
# e2e (async) and blocking e2e of person_detect batch 8192 against the host-path piece size.  Output: gpurun_out/host_piece_sweep.txt
mkdir -p gpurun_out
: > gpurun_out/host_piece_sweep.txt
for p in 0 4096 2048 1024; do
  MF_HOST_PIECE=$p timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-conv2d 2>/dev/null | tail -1 > /tmp/hp.json
  python - "$p" >> gpurun_out/host_piece_sweep.txt <<'PY'
import json, sys
d = json.loads(open("/tmp/hp.json").read())
print("MF_HOST_PIECE=%s  e2e async %.4g  blocking %.4g  (device-resident %.4g)" % (sys.argv[1], d["e2e"]["value"], d["e2e"]["blocking"], d["value"]))
PY
done
cat gpurun_out/host_piece_sweep.txt
