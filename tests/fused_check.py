"""Fused low-resolution chain kernel (mf_fused.cu) vs the oracle, bit for bit; run in a SUBPROCESS by test_gpu_fused.py (a device
trap in a kernel under development must not poison the CUDA context of the other tests).  Prints one JSON line.

The chain is driven through the C-ABI hook mf_op_conv_chain: n x [depthwise_conv_2d 3x3 s1 SAME, 128 ch] -> [conv_2d 1x1 128 -> 128]
with random weights, per-channel constants and zero-points, once as ONE fused launch and once layer by layer; the oracle applies
the reference's operators one after the other (src/ops/depthwise_conv_2d.rs:56-101, conv_2d.rs:56-104)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import microflow_rs_b200 as mf  # noqa: E402
import oracle  # noqa: E402

C = 128


def make_chain(r, H, W, pairs, act="relu6", narrow=False):
    layers = []
    zp = int(r.integers(-128, 128))
    for l in range(pairs):
        out_zp = -128 if not narrow else int(r.integers(-120, -60))
        dw = r.integers(-128, 128, (1, 3, 3, C)).astype(np.int8)
        layers.append(dict(in_zp=zp, filters=dw, filter_zp=[0], out_scale=0.0235294, out_zp=out_zp, act=act, pad="same", strides=(1, 1),
                           c0=r.uniform(-30, 30, C).astype(np.float32), c1=(r.uniform(0.3, 2.0, C) / 300.0).astype(np.float32), out_hw=(H, W),
                           depthwise=True))
        zp = out_zp
        out_zp = -128 if not narrow else int(r.integers(-120, -60))
        pw = r.integers(-128, 128, (C, 1, 1, C)).astype(np.int8)
        layers.append(dict(in_zp=zp, filters=pw, filter_zp=[0], out_scale=0.0235294, out_zp=out_zp, act=act, pad="same", strides=(1, 1),
                           c0=r.uniform(-30, 30, C).astype(np.float32), c1=(r.uniform(0.3, 2.0, C) / 2500.0).astype(np.float32), out_hw=(H, W),
                           depthwise=False))
        zp = out_zp
    return layers


def oracle_chain(x, layers):
    out = []
    for b in range(x.shape[0]):
        t = x[b]
        for L in layers:
            fn = oracle.depthwise_conv_2d if L["depthwise"] else oracle.conv_2d
            t = fn(t, L["in_zp"], L["filters"], L["filter_zp"], L["out_scale"], L["out_zp"], L["act"], L["pad"], L["strides"], L["c0"], L["c1"], L["out_hw"])
        out.append(t)
    return np.stack(out)


def main():
    which = sys.argv[1]
    r = np.random.default_rng(20261017)
    res = {"which": which, "ok": False, "cases": 0}
    if which == "shapes":
        # (H, W, pairs, batch, narrow clamp): unit sizes 7 / 16 / 17 / 4 / 2 / 256 samples, odd widths, partial units and tiles
        cases = [(6, 6, 5, 1, False), (6, 6, 5, 7, False), (6, 6, 5, 8, False), (6, 6, 2, 15, False), (6, 6, 5, 45, True), (4, 4, 3, 37, False),
                 (3, 5, 2, 40, False), (8, 8, 2, 9, False), (11, 11, 2, 5, True), (1, 1, 2, 300, False), (2, 7, 4, 33, False), (6, 6, 5, 300, False)]
    elif which == "large":
        cases = [(6, 6, 5, 2 * 148 * 7 + 3, False)]            # every CTA gets units for both teams, the last unit is partial
    else:
        raise SystemExit("unknown group")
    for (H, W, pairs, B, narrow) in cases:
        layers = make_chain(r, H, W, pairs, act="relu" if narrow else "relu6", narrow=narrow)
        x = r.integers(-128, 128, (B, H, W, C)).astype(np.int8)
        got = mf.ops.conv_chain(x, layers, fuse=True)
        if "fused_chain_kernel" not in mf.ops.last_kernel:
            res["error"] = f"case {(H, W, pairs, B)} ran on {mf.ops.last_kernel}"
            print(json.dumps(res)); return
        ref_gpu = mf.ops.conv_chain(x, layers, fuse=False)
        nchk = min(B, 24)
        sel = np.unique(np.concatenate([np.arange(min(B, 12)), np.arange(B - min(B, 12), B)]))[:nchk]
        want = oracle_chain(x[sel], layers)
        if not np.array_equal(got[sel], want) or not np.array_equal(got, ref_gpu):
            bad = int((got != ref_gpu).sum())
            badrows = np.unique(np.nonzero((got != ref_gpu).reshape(B, -1).any(axis=1))[0])[:8].tolist()
            res["error"] = (f"case {(H, W, pairs, B, narrow)}: {bad} of {got.size} bytes differ from the layer-by-layer kernels "
                            f"(samples {badrows}); oracle rows equal: {bool(np.array_equal(got[sel], want))}")
            print(json.dumps(res)); return
        res["cases"] += 1
    if which == "shapes":
        # a chain the kernel cannot hold (six 16 KB weight images + two units exceed 227 KB) must be refused, not mangled
        layers = make_chain(r, 6, 6, 6)
        x = r.integers(-128, 128, (3, 6, 6, C)).astype(np.int8)
        try:
            mf.ops.conv_chain(x, layers, fuse=True)
            res["error"] = "a 6-pair chain was accepted"
            print(json.dumps(res)); return
        except mf.MicroflowError as e:
            if e.status != 7:
                raise
        if not np.array_equal(mf.ops.conv_chain(x, layers, fuse=False), oracle_chain(x, layers)):
            res["error"] = "layer-by-layer 6-pair chain differs from the oracle"
            print(json.dumps(res)); return
    res["ok"] = True
    print(json.dumps(res))


if __name__ == "__main__":
    main()
