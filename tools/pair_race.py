"""One launch of conv3x3_pair_kernel checked against the generic kernel: the smallest workload for compute-sanitizer runs on the
CTA-pair path (`compute-sanitizer --tool racecheck python tools/pair_race.py`; see profiles/r02g_conv3x3_experiments.txt, item 4)."""
import sys, os
from pathlib import Path
import numpy as np
ROOT = Path.cwd()
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import microflow_rs_b200 as mf
r = np.random.default_rng(1)
x3 = r.integers(-128, 128, (2, 24, 16, 128)).astype(np.int8)
w3 = r.integers(-128, 128, (128, 3, 3, 128)).astype(np.int8)
c1 = (r.uniform(0.2, 2.0, 128) / 40000.0).astype(np.float32); c0 = r.uniform(-20, 20, 128).astype(np.float32)
a3 = mf.ops.conv_2d(x3, -128, w3, [0], 0.0235294, -128, "relu6", "same", (1, 1), c0, c1, (24, 16), impl=0)
print(mf.ops.last_kernel)
assert np.array_equal(a3, mf.ops.conv_2d(x3, -128, w3, [0], 0.0235294, -128, "relu6", "same", (1, 1), c0, c1, (24, 16), impl=1))
print("ok")
