#!/usr/bin/env python
"""Times single person_detect layers in isolation (persistent ConvOp, device-resident, inputs rotate over >L2 buffers)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import microflow_rs_b200 as mf  # noqa: E402

LAYERS = {  # name: (H, W, Cin, Cout, K, stride, depthwise)
    "L0_dw_cin1": (96, 96, 1, 8, 3, 2, True), "L1_dw8": (48, 48, 8, 8, 3, 1, True), "L2_pw8_16": (48, 48, 8, 16, 1, 1, False),
    "L3_dw16_s2": (48, 48, 16, 16, 3, 2, True), "L4_pw16_32": (24, 24, 16, 32, 1, 1, False), "L5_dw32": (24, 24, 32, 32, 3, 1, True),
    "L6_pw32_32": (24, 24, 32, 32, 1, 1, False), "L13_dw128": (6, 6, 128, 128, 3, 1, True), "L14_pw128": (6, 6, 128, 128, 1, 1, False),
    "L26_pw256": (3, 3, 256, 256, 1, 1, False),
    "L7_dw32_s2": (24, 24, 32, 32, 3, 2, True), "L9_dw64": (12, 12, 64, 64, 3, 1, True), "L11_dw64_s2": (12, 12, 64, 64, 3, 2, True),
    "L23_dw128_s2": (6, 6, 128, 128, 3, 2, True), "L25_dw256": (3, 3, 256, 256, 3, 1, True), "L8_pw32_64": (12, 12, 32, 64, 1, 1, False),
    "L10_pw64_64": (12, 12, 64, 64, 1, 1, False), "L12_pw64_128": (6, 6, 64, 128, 1, 1, False), "L24_pw128_256": (3, 3, 128, 256, 1, 1, False),
}


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else list(LAYERS)
    torch.cuda.set_stream(torch.cuda.Stream())
    st = torch.cuda.current_stream().cuda_stream
    r = np.random.default_rng(0)
    for name in names:
        H, W, Cin, Cout, K, s, dw = LAYERS[name]
        OH, OW = -(-H // s), -(-W // s)
        w = r.integers(-128, 128, (1, K, K, Cout) if dw else (Cout, K, K, Cin)).astype(np.int8)
        c1 = r.uniform(1e-4, 1e-3, Cout).astype(np.float32)
        c0 = r.uniform(-4, 4, Cout).astype(np.float32)
        op = mf.ConvOp((H, W, Cin), -128, w, [0], 0.0235294, -128, "relu6", "same", (s, s), c0, c1, (OH, OW), depthwise=dw)
        nbuf = max(2, int(200e6 // (batch * H * W * Cin)) + 1)
        xs = [torch.randint(-128, 128, (batch, H, W, Cin), dtype=torch.int8, device="cuda") for _ in range(min(nbuf, 6))]
        y = torch.empty((batch, OH, OW, Cout), dtype=torch.int8, device="cuda")
        for i in range(3):
            op.run_device(xs[i % len(xs)].data_ptr(), y.data_ptr(), batch, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for i in range(n):
            op.run_device(xs[i % len(xs)].data_ptr(), y.data_ptr(), batch, st)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / n
        byts = batch * (H * W * Cin + OH * OW * Cout)
        print(json.dumps({"layer": name, "kernel": op.kernel, "batch": batch, "us": round(us, 1), "GBps": round(byts / us / 1e3, 1),
                          "Mout_per_s": round(batch * OH * OW * Cout / us, 1)}), flush=True)
        op.close()


if __name__ == "__main__":
    main()
