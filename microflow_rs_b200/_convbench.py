"""BASELINE config 5: synthetic Conv2D 224x224x128 -> 128, k3 s1 SAME, int8, ReLU6, on device-resident buffers.
Returns the `conv2d` roofline object of bench.py (tensor-pipe bound)."""
import numpy as np

from . import ConvOp


def run(torch, w, c0, c1, in_zp, out_zp, out_scale, H, W, batch, steps, warmup, peaks, seed, clock_sampler=None):
    Cout, _, _, Cin = w.shape
    fast = ConvOp((H, W, Cin), in_zp, w, [0], out_scale, out_zp, "relu6", "same", (1, 1), c0, c1, (H, W), impl=0)
    res = {"workload": f"synthetic Conv2D {H}x{W}x{Cin}->{Cout} k3 s1 SAME int8 ReLU6, batch {batch} (BASELINE configs[4])", "kernel": fast.kernel}
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    x = [torch.randint(-128, 128, (batch, H, W, Cin), dtype=torch.int8, device="cuda", generator=g) for _ in range(2)]
    y = torch.empty((batch, H, W, Cout), dtype=torch.int8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert st != 0, "run under a non-default torch stream so that the CUDA events bracket the kernel launches"
    # correctness guard inside the bench: one image against the generic direct kernel (bit-exact)
    gen = ConvOp((H, W, Cin), in_zp, w, [0], out_scale, out_zp, "relu6", "same", (1, 1), c0, c1, (H, W), impl=1)
    y1 = torch.empty((1, H, W, Cout), dtype=torch.int8, device="cuda")
    y2 = torch.empty((1, H, W, Cout), dtype=torch.int8, device="cuda")
    fast.run_device(x[0].data_ptr(), y1.data_ptr(), 1, st)
    gen.run_device(x[0].data_ptr(), y2.data_ptr(), 1, st)
    torch.cuda.synchronize()
    res["verified_vs_generic_kernel"] = bool(torch.equal(y1, y2))
    gen.close()
    for i in range(warmup):
        fast.run_device(x[i & 1].data_ptr(), y.data_ptr(), batch, st)
    torch.cuda.synchronize()
    from . import lib
    res["kernel"] = lib().mf_op_kernel_name(fast._h).decode()      # what the launches above really ran (CTA-pair or one-CTA kernel)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fast.run_device(x[i & 1].data_ptr(), y.data_ptr(), batch, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # SM clock under THIS kernel: the timed region is a millisecond, so sample NVML over ~0.3 s of the same launches right after it.
    # (The 3x3 kernel keeps tensor pipe, TMA and HBM busy at once and runs power-capped well below the 1965 MHz the network step holds.)
    clocks = None
    if clock_sampler is not None:
        import time
        cs = clock_sampler
        cs.start()
        cs.window(True)
        t_end = time.perf_counter() + 0.3
        i = 0
        while time.perf_counter() < t_end:
            for _ in range(50):
                fast.run_device(x[i & 1].data_ptr(), y.data_ptr(), batch, st)
                i += 1
            torch.cuda.synchronize()
        cs.window(False)
        c = cs.stop()
        clocks = {"sm_mhz": c.get("sm_mhz"), "sm_max_mhz": c.get("sm_max_mhz"), "reasons": c.get("reasons"), "samples": c.get("samples"),
                  "what": "NVML samples over 0.3 s of the same back-to-back launches, untimed, right after the timed region"}
    ops = 2.0 * fast.macs * batch
    tops = ops / (ms * 1e-3) / 1e12
    peak = 2.0 * peaks["bf16_tflops"]
    traffic = None
    try:   # dram__bytes_read + write of one launch at batch 16 from the committed ncu capture (profiles/r02e_conv3x3.txt)
        import json
        from pathlib import Path
        traffic = json.loads((Path(__file__).resolve().parent.parent / "profiles" / "traffic_latest.json").read_text()).get(res["kernel"]) if batch == 16 else None
    except Exception:
        traffic = None
    mma_only = 4423.0   # tools/ubench/mma_i8.cu on this pool's B200: tcgen05.mma kind::i8 M128 N128 K32 back to back, no TMA, no epilogue (profiles/r02d)
    res.update({"ms_per_launch": ms, "images_per_s": batch / (ms * 1e-3),
                "roofline": {"bound": "tensor", "achieved": tops, "peak": peak, "unit": "TOP/s", "frac": tops / peak, "traffic": traffic,
                             "peak_source": "2 x measured cuBLAS bf16 burst TFLOP/s (tcgen05 kind::i8 issues 2x the bf16 MAC rate); nominal dense int8 is 4500; "
                                            "the MMA-only ceiling of this instruction shape measured by tools/ubench/mma_i8.cu is 4423 TOP/s",
                             "frac_of_nominal_4500": tops / 4500.0, "frac_of_mma_only_ceiling_4423": tops / mma_only,
                             # the same ceiling in CYCLES (66.9 clk per M128 N128 K32 instruction, profiles/r02g_mma_i8_views.txt) at the SM clock
                             # this kernel sustains: separates what the kernel loses from what the power cap takes
                             "mma_only_ceiling_at_sampled_clock": (148 * 2.0 * 128 * 128 * 32 / 66.9 * clocks["sm_mhz"] * 1e6 / 1e12) if clocks and clocks.get("sm_mhz") else None,
                             "frac_of_mma_only_ceiling_at_sampled_clock": (tops / (148 * 2.0 * 128 * 128 * 32 / 66.9 * clocks["sm_mhz"] * 1e6 / 1e12)) if clocks and clocks.get("sm_mhz") else None},
                "clocks": clocks,
                "l2": f"input+output per launch {(x[0].numel() + y.numel()) / 1e6:.0f} MB > 126 MB L2; inputs alternate between 2 buffers",
                "algorithmic_bytes_per_launch": int(x[0].numel() + y.numel() + fast.weight_bytes)})
    fast.close()
    return res


def main():
    """python -m microflow_rs_b200._convbench [batch] [steps] -- stand-alone BASELINE config 5 run (used under ncu)."""
    import json
    import sys
    from pathlib import Path

    import torch
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    import bench
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    peaks, _ = bench.load_peaks()
    torch.cuda.set_stream(torch.cuda.Stream())
    H = W = 224
    Cin = Cout = 128
    seed = 0x5EED0005
    w = bench.splitmix_bytes(seed, Cout * 9 * Cin).reshape(Cout, 3, 3, Cin)
    r = np.random.default_rng(seed)
    c1 = r.uniform(1e-3, 1e-2, Cout).astype(np.float32)
    c0 = r.uniform(-4, 4, Cout).astype(np.float32)
    print(json.dumps(run(torch, w, c0, c1, in_zp=-128, out_zp=-128, out_scale=0.0235294, H=H, W=W, batch=batch, steps=steps, warmup=3, peaks=peaks,
                         seed=seed, clock_sampler=bench.ClockSampler(0))))


if __name__ == "__main__":
    main()
