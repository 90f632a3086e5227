#!/bin/bash
# A/B runs of the fused chain kernel under different environment switches (run under gpurun): prints the chain's time per step
# (per-layer CUDA events) and the whole step.  Usage: tools/fused_ab.sh "VAR=val VAR2=val" "VAR=val" ...
for cfg in "$@"; do
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-conv2d 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernels']
print('$cfg', '| step %.4f ms (events %.4f) |' % (d['ms_per_step'], d['ms_per_step_with_layer_events']), {n: round(v['ms_per_step'] * 1e3, 1) for n, v in k.items() if 'fused' in n or 'tail' in n})"
done
