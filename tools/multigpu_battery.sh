#!/bin/bash
# Multi-GPU session (run under gpurun --gpus N): PCIe / NUMA topology evidence, concurrent H2D sweep, multi-device tests, and bench.py in
# both deployment forms (one process per GPU under torchrun; one process driving all GPUs through mf_options.devices).
set -u
N="${1:-8}"
mkdir -p gpurun_out
{
  echo "== nvidia-smi topo -m"; nvidia-smi topo -m
  echo "== nvidia-smi topo -p2p r (first lines)"; nvidia-smi topo -p2p r 2>&1 | head -12
  echo "== lscpu (NUMA)"; lscpu | grep -Ei "model name|socket|numa|^cpu\(s\)|thread"
  echo "== lspci -tv (bridges + NVIDIA)"; (lspci -tv 2>/dev/null || echo "lspci not available") | grep -Ei "nvidia|\[[0-9a-f]{2}(-[0-9a-f]{2})?\]-|root" | head -80
  echo "== numactl -H"; (numactl -H 2>/dev/null || echo "numactl not available")
  echo "== /sys NUMA node of each GPU"; for d in /sys/bus/pci/devices/*; do v=$(cat $d/vendor 2>/dev/null); c=$(cat $d/class 2>/dev/null); if [ "$v" = "0x10de" ] && [[ "$c" == 0x0302* ]]; then echo "$(basename $d) numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi; done
} > gpurun_out/topology_${N}gpu.txt 2>&1
timeout 300 python tools/h2d_sweep.py $N > gpurun_out/h2d_sweep_${N}gpu.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_multidevice.py -q -m gpu 2>&1 | tail -8 > gpurun_out/t_multi_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 \
    > gpurun_out/bench_pd_${N}gpu.json 2> gpurun_out/bench_pd_${N}gpu.err
timeout 600 python bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_pd_${N}gpu_single_process.json 2> gpurun_out/bench_pd_${N}gpu_single_process.err
tail -n 30 gpurun_out/topology_${N}gpu.txt; cat gpurun_out/h2d_sweep_${N}gpu.txt; cat gpurun_out/t_multi_${N}gpu.log
for f in gpurun_out/bench_pd_${N}gpu.json gpurun_out/bench_pd_${N}gpu_single_process.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    e = d["e2e"]
    print(sys.argv[1], "value %.4g e2e %.4g blocking %.4g ms/step %.4f verified %s link/gpu %s" % (d["value"], e["value"], e.get("blocking", 0), d["ms_per_step"], d.get("verified"), e.get("link_bound_per_gpu")))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
tail -n 5 gpurun_out/bench_pd_${N}gpu.err gpurun_out/bench_pd_${N}gpu_single_process.err
