"""Stand-alone tcgen05 kernel checks, run in a SUBPROCESS by test_gpu_tc.py (a device trap in an experimental
kernel must not poison the CUDA context of the other tests).  Prints one JSON line."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import microflow_rs_b200 as mf  # noqa: E402
import oracle  # noqa: E402


def conv_case(r, B, H, W, Cin, Cout, K, act, in_zp=None):
    x = r.integers(-128, 128, (B, H, W, Cin)).astype(np.int8)
    w = r.integers(-128, 128, (Cout, K, K, Cin)).astype(np.int8)
    in_zp = int(r.integers(-128, 128)) if in_zp is None else in_zp
    c1 = (r.uniform(0.2, 2.0, Cout) / (K * K * Cin * 40.0)).astype(np.float32)
    c0 = r.uniform(-20, 20, Cout).astype(np.float32)
    return x, w, in_zp, c0, c1, act


def main():
    which = sys.argv[1]
    r = np.random.default_rng(1234)
    res = {"which": which, "ok": False}
    if which == "pointwise":
        # (B, H, W, Cin, Cout): all person_detect 1x1 shapes + a ragged row count
        shapes = [(2, 48, 48, 8, 16), (2, 24, 24, 16, 32), (3, 24, 24, 32, 32), (3, 12, 12, 32, 64), (5, 12, 12, 64, 64), (7, 6, 6, 64, 128),
                  (9, 6, 6, 128, 128), (11, 3, 3, 128, 256), (13, 3, 3, 256, 256), (1, 6, 6, 128, 128), (300, 6, 6, 128, 128)]
        for (B, H, W, Cin, Cout) in shapes:
            x, w, in_zp, c0, c1, act = conv_case(r, B, H, W, Cin, Cout, 1, "relu6")
            got = mf.ops.conv_2d(x, in_zp, w, [0], 0.0235294, -128, act, "same", (1, 1), c0, c1, (H, W), impl=0)
            kern = mf.ops.last_kernel
            if "conv_tc" not in kern:
                res["error"] = f"shape {(B, H, W, Cin, Cout)} ran on {kern}"
                print(json.dumps(res)); return
            nchk = min(B, 4)
            want = np.stack([oracle.conv_2d(x[b], in_zp, w, [0], 0.0235294, -128, act, "same", (1, 1), c0, c1, (H, W)) for b in range(nchk)])
            ref_gpu = mf.ops.conv_2d(x, in_zp, w, [0], 0.0235294, -128, act, "same", (1, 1), c0, c1, (H, W), impl=1)
            if not np.array_equal(got[:nchk], want) or not np.array_equal(got, ref_gpu):
                bad = int((got != ref_gpu).sum())
                res["error"] = f"mismatch shape {(B, H, W, Cin, Cout)}: {bad} of {got.size} bytes differ from the generic kernel"
                print(json.dumps(res)); return
        res["ok"] = True
    elif which == "conv3x3":
        shapes = [(2, 16, 16, 128, 128), (1, 8, 24, 128, 64), (2, 9, 21, 128, 128), (1, 5, 3, 256, 32), (1, 2, 2, 128, 64), (3, 40, 40, 128, 128)]
        for (B, H, W, Cin, Cout) in shapes:
            x, w, in_zp, c0, c1, act = conv_case(r, B, H, W, Cin, Cout, 3, "relu6", in_zp=-128 if Cout == 128 else None)
            got = mf.ops.conv_2d(x, in_zp, w, [0], 0.0235294, -128, act, "same", (1, 1), c0, c1, (H, W), impl=0)
            kern = mf.ops.last_kernel
            import os
            pair_expected = os.environ.get("MF_TC_PAIR", "1") != "0" and Cin == 128 and Cout % 64 == 0 and B * (-(-H // 16)) * (-(-W // 8)) >= 2
            if ("conv3x3_pair_kernel" in kern) != pair_expected or ("conv_tc" not in kern and "conv3x3_pair" not in kern):
                res["error"] = f"shape {(B, H, W, Cin, Cout)} ran on {kern}"
                print(json.dumps(res)); return
            want = np.stack([oracle.conv_2d(x[b], in_zp, w, [0], 0.0235294, -128, act, "same", (1, 1), c0, c1, (H, W)) for b in range(B)])
            if not np.array_equal(got, want):
                bad = int((got != want).sum())
                res["error"] = f"mismatch shape {(B, H, W, Cin, Cout)}: {bad} of {got.size} bytes differ from the oracle"
                print(json.dumps(res)); return
        res["ok"] = True
    print(json.dumps(res))


if __name__ == "__main__":
    main()
