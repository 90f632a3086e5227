#!/usr/bin/env python
"""bench.py -- throughput of MicroFlow's quantized hot path on B200 (driver contract: see the task brief).

    python bench.py --gpus N --steps K --warmup W                 our CUDA arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W  the reference algorithm on the host CPU cores

A "step" = one pass of the whole quantized graph over one batch of synthetic int8 inputs of the model's shape
(default workload: person_detect.tflite, batch 8192 per GPU = BASELINE.json configs[2]; weak scaling: batch 8192 on
every GPU, samples are independent, no collective on the inference path, one NCCL broadcast of the weight blob at init).

Printed JSON (one line, rank 0):
  value      inferences/s, inputs resident in HBM, CUDA-event timed on the launching stream, max over ranks
  e2e        same metric through the public C-ABI call mf_predict_many_quantized with pinned HOST buffers (H2D + D2H inside)
  roofline   dominant kernel of the step: algorithmic bytes / its summed launch time vs the measured HBM copy peak
  conv2d     BASELINE config 5 (synthetic 224x224x128->128 3x3 Conv2D) on the tcgen05 kernel vs the int8 tensor peak (N=1 only)
  cpu_baseline  the oracle (faithful CPU port of the reference algorithm; the Rust reference cannot be built here) timed on one
                host core over a bounded sample
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
MODELS = ROOT / "tests" / "golden" / "models"
SEEDS = {"person_detect": 0x5EED0003, "speech": 0x5EED0002, "sine": 0x5EED0001}
DEFAULT_BATCH = {"person_detect": 8192, "speech": 4096, "sine": 65536}
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def splitmix_bytes(seed, n, offset=0):
    k = (np.arange(offset, offset + n, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = k
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z & np.uint64(0xFF)).astype(np.uint8).view(np.int8)


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"]))}, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS, bf16_tflops_sustained=1400.0), "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed regions, polled through NVML every ~2 ms (the nvidia-smi -lms loop of
    B200_PROFILING.md starts too slowly for a timed region of a few tens of milliseconds)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self.active = False
        self.stop_flag = False
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            if self.active:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    try:
                        rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.samples.append((sm, rs))
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def window(self, on):
        self.active = on

    @staticmethod
    def _reasons(samples):
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
        return sorted({k for _, rs in samples for k, bit in names.items() if rs & bit})

    def stop(self, n_timed=None):
        """n_timed: how many of the samples fell inside the timed regions (the rest: the untimed extension under the same load).
        `sm_mhz` / `reasons` describe the timed regions when they hold samples, else the extension; both are reported."""
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1)
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        n_timed = len(self.samples) if n_timed is None else n_timed
        timed, ext = self.samples[:n_timed], self.samples[n_timed:]
        main = timed if timed else ext
        out = {"sm_mhz": float(np.median([s for s, _ in main])), "sm_max_mhz": mx, "reasons": self._reasons(main), "samples": len(self.samples),
               "samples_in_timed_regions": len(timed)}
        if ext:
            out["extension"] = {"sm_mhz": float(np.median([s for s, _ in ext])), "reasons": self._reasons(ext), "samples": len(ext),
                                "what": "0.5 s of the same device-resident steps, untimed, right after the timed regions"}
        return out


def cpu_baseline(workload, seconds=12.0, threads=1, max_samples=4096):
    """Times the oracle (-O3 -march=native build) on a bounded sample of the same synthetic workload."""
    import oracle
    o = oracle.Model(MODELS / f"{workload}.tflite", fast=True)
    block = {"person_detect": 64, "speech": 512, "sine": 8192}[workload] * max(1, threads)
    done, t_used, off = 0, 0.0, 0
    o.predict_many_quantized(splitmix_bytes(SEEDS[workload], 2 * o.in_elems).reshape(2, -1), threads=1)  # warm-up
    while t_used < seconds and done < max_samples * max(1, threads):
        xs = splitmix_bytes(SEEDS[workload], block * o.in_elems, offset=off).reshape(block, -1)
        t0 = time.perf_counter()
        o.predict_many_quantized(xs, threads=threads)
        t_used += time.perf_counter() - t0
        done += block
        off += block * o.in_elems
    return {"value": done / t_used, "unit": "inferences/s", "cores": threads, "kind": "port",
            "sample": f"{done} synthetic {workload} samples (splitmix64 seed {SEEDS[workload]:#x}), {t_used:.1f} s, "
                      f"oracle/microflow_oracle.c -O3 -march=native -ffp-contract=off, {threads} thread(s)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own algorithm on the box's host cores (oracle port; Rust cannot be built here).
    Every step processes the SAME batch as the CUDA arm (8192 person_detect samples) when the whole run then stays under
    ~6 minutes on this box; otherwise a bounded sample, and `config.same_config` says so.  MF_REF_THREADS fixes the thread count
    (default: every hardware thread), so that numbers from boxes with different core counts can be compared."""
    if rank != 0:
        return
    import oracle
    wl = args.workload
    threads = oracle.max_threads()
    if os.environ.get("MF_REF_THREADS"):
        threads = max(1, min(threads, int(os.environ["MF_REF_THREADS"])))
    o = oracle.Model(MODELS / f"{wl}.tflite", fast=True)
    batch = args.batch or DEFAULT_BATCH[wl]
    probe_n = {"person_detect": 16, "speech": 128, "sine": 4096}[wl] * threads
    xs = splitmix_bytes(SEEDS[wl], probe_n * o.in_elems).reshape(probe_n, -1)
    o.predict_many_quantized(xs, threads=threads)
    t0 = time.perf_counter()
    o.predict_many_quantized(xs, threads=threads)
    rate = probe_n / (time.perf_counter() - t0)
    budget_s = float(os.environ.get("MF_REF_BUDGET_S", "360"))
    per_step = batch
    if batch * (args.steps + args.warmup) / rate > budget_s:
        per_step = max(threads, int(rate * budget_s / (args.steps + args.warmup)) // threads * threads)
    same = per_step == batch
    xs = splitmix_bytes(SEEDS[wl], per_step * o.in_elems).reshape(per_step, -1)
    for _ in range(args.warmup):
        o.predict_many_quantized(xs, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.predict_many_quantized(xs, threads=threads)
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    sample = f"{per_step} synthetic {wl} samples per step on {threads} host threads (oracle port of the reference algorithm)"
    cfg = {"workload": f"{wl}.tflite int8, batch {per_step} per step" + (" (BASELINE configs[2])" if same and wl == "person_detect" else " (bounded CPU sample)"),
           "same_config": same, "threads": threads}
    line = {"impl": "reference", "metric": f"inferences/s {wl} int8", "value": val, "unit": "inferences/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": "inferences/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "inferences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def conv2d_roofline(torch, mf, peaks, steps, warmup, batch=16):
    """BASELINE config 5: synthetic Conv2D 224x224x128 -> 128, k3 s1 SAME, int8, ReLU6 on the tcgen05 kernel."""
    H = W = 224
    Cin = Cout = 128
    seed = 0x5EED0005
    w = splitmix_bytes(seed, Cout * 9 * Cin).reshape(Cout, 3, 3, Cin)
    r = np.random.default_rng(seed)
    c1 = r.uniform(1e-3, 1e-2, Cout).astype(np.float32)   # = in_scale * filter_scale[b] / out_scale with in_scale == out_scale
    c0 = r.uniform(-4, 4, Cout).astype(np.float32)
    # drive the op through a one-layer engine object (private helper of the package keeps device buffers resident)
    from microflow_rs_b200 import _convbench
    return _convbench.run(torch, w, c0, c1, in_zp=-128, out_zp=-128, out_scale=0.0235294, H=H, W=W, batch=batch, steps=steps, warmup=warmup,
                          peaks=peaks, seed=seed, clock_sampler=ClockSampler(torch.cuda.current_device()))


def verify_ranks(dist, torch, m, wl, rank, world, batch):
    """BASELINE configs[3]: "result must equal the 1-GPU result row-for-row".  Outside every timed region each rank runs the
    first 256 samples of RANK 0's shard plus 32 samples of its own shard; the int8 outputs and pre-softmax logits are
    all-gathered; rank 0 checks that every rank produced rank 0's bytes on the common rows (a broken weight broadcast or a
    mis-addressed shard cannot pass) and checks its own 32 rows against the oracle."""
    ie = m.in_elems
    common_n, own_n = 256, 32
    common = splitmix_bytes(SEEDS[wl], common_n * ie, offset=0).reshape(common_n, ie)
    own = splitmix_bytes(SEEDS[wl], own_n * ie, offset=rank * batch * ie).reshape(own_n, ie)
    xs = np.concatenate([common, own])
    out_q, logits = m.predict_many_logits(xs)
    mine = np.concatenate([out_q.reshape(len(xs), -1).view(np.uint8), logits.reshape(len(xs), -1).view(np.uint8)], axis=1) if logits is not None else \
        out_q.reshape(len(xs), -1).view(np.uint8)
    t = torch.from_numpy(np.ascontiguousarray(mine)).cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    res = None
    if rank == 0:
        import oracle
        ok_common = all(bool(torch.equal(parts[r][:common_n], parts[0][:common_n])) for r in range(world))
        o = oracle.Model(MODELS / f"{wl}.tflite", fast=True)
        _, want_q = o.predict_many_quantized(own)
        ok_oracle = bool(np.array_equal(out_q[common_n:].reshape(own_n, -1).view(np.uint8), np.asarray(want_q).reshape(own_n, -1).view(np.uint8)))
        res = {"ranks": world, "rows": common_n * world + own_n, "common_rows_equal_on_all_ranks": ok_common, "own_rows_equal_oracle": ok_oracle,
               "ok": ok_common and ok_oracle, "what": "first 256 samples of rank 0's shard on every rank (all_gather, byte-equal) + 32 own rows vs the oracle"}
    return res


def main_single_process(args):
    """`python bench.py --gpus N` WITHOUT torchrun: the N GPUs are driven by ONE process through the library's own multi-device
    model (mf_options.devices): weights broadcast once at create, predict_many shards the samples inside the C-ABI call."""
    import torch
    import microflow_rs_b200 as mf
    wl = args.workload
    G = args.gpus
    batch = args.batch or DEFAULT_BATCH[wl]
    m = mf.Model(MODELS / f"{wl}.tflite", devices=list(range(G)), chunk=args.chunk, flags=args.flags)
    ie, oe = m.in_elems, m.out_elems
    host_in = [mf.PinnedBuffer((G * batch, ie), np.int8) for _ in range(2)]
    for r_i, hb in enumerate(host_in):
        hb.array[:] = splitmix_bytes(SEEDS[wl] + r_i, G * batch * ie).reshape(G * batch, ie)
    host_out = [mf.PinnedBuffer((G * batch, oe), np.float32) for _ in range(2)]
    # device-resident leg: per device its contiguous shard, enqueued by one Python thread per device (ctypes drops the GIL)
    R = max(2, int(np.ceil(160e6 / (batch * ie))) + 1)
    d_in, d_out, streams, evs = [], [], [], []
    for g in range(G):
        with torch.cuda.device(g):
            d_in.append([torch.from_numpy(host_in[i % 2].array[g * batch:(g + 1) * batch]).cuda() for i in range(R)])
            d_out.append(torch.empty((batch, oe), dtype=torch.float32, device=f"cuda:{g}"))
            streams.append(torch.cuda.Stream(device=g))
            evs.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))

    def run_dev(g, first, count, timed):
        with torch.cuda.device(g):
            if timed:
                evs[g][0].record(streams[g])
            for i in range(first, first + count):
                m.predict_many_device_on(g, d_in[g][i % R].data_ptr(), batch, d_out[g].data_ptr(), None, streams[g].cuda_stream)
            if timed:
                evs[g][1].record(streams[g])

    def all_devs(first, count, timed):
        ths = [threading.Thread(target=run_dev, args=(g, first, count, timed)) for g in range(G)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for g in range(G):
            torch.cuda.synchronize(g)

    all_devs(0, args.warmup, False)
    launches0 = m.launch_count()
    sampler = ClockSampler(0)
    sampler.start()
    sampler.window(True)
    all_devs(args.warmup, args.steps, True)
    sampler.window(False)
    ms = max(evs[g][0].elapsed_time(evs[g][1]) for g in range(G))
    launches = m.launch_count() - launches0
    value = G * batch * args.steps / (ms * 1e-3)
    # e2e: ONE C-ABI call per step shards G * batch host samples over the devices (H2D + D2H inside)
    for i in range(3):
        m.predict_many_quantized(host_in[i & 1].array, out=host_out[i & 1].array)
    sampler.window(True)
    t0 = time.perf_counter()
    for i in range(args.steps):
        m.predict_many_quantized(host_in[i & 1].array, out=host_out[i & 1].array)
    e2e_block_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(args.steps):
        m.predict_many_quantized_async(host_in[i & 1].array, host_out[i & 1].array)
    m.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.window(False)
    clocks = sampler.stop()
    # row-for-row equality with the 1-GPU result (BASELINE configs[3]): the same rows through a one-device model
    m1 = mf.Model(MODELS / f"{wl}.tflite", device=0)
    nchk = min(G * batch, 4096)
    sel = np.linspace(0, G * batch - 1, nchk).astype(np.int64)       # rows from every shard
    xs = np.ascontiguousarray(host_in[0].array[sel])
    q1, l1 = m1.predict_many_logits(xs)
    qG, lG = m.predict_many_logits(host_in[0].array)
    ok = bool(np.array_equal(qG[sel], q1)) and (l1 is None or bool(np.array_equal(lG[sel], l1)))
    m1.close()
    line = {"metric": f"inferences/s {wl} int8", "value": value, "unit": "inferences/s", "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": f"{wl}.tflite int8, batch {batch} per GPU" + (" (BASELINE configs[3] shape)" if wl == "person_detect" else ""),
                       "global_batch": batch * G, "parallelism": f"dp{G} in ONE process: mf_options.devices = {m.devices}, weights by one '{m.weight_broadcast}' "
                                                                 "broadcast at create, contiguous shards inside mf_predict_many*",
                       "l2": f"inputs rotate over {R} device batches per GPU"},
            "e2e": {"value": G * batch * args.steps / e2e_s, "unit": "inferences/s", "h2d_bytes_per_step": G * batch * ie, "d2h_bytes_per_step": G * batch * oe * 4,
                    "api": "mf_predict_many_quantized_async (one call shards over all devices) x K + mf_model_synchronize", "blocking": G * batch * args.steps / e2e_block_s},
            "gpu_launches": int(launches), "clocks": clocks,
            "verified": {"devices": G, "rows": int(nchk), "ok": ok, "what": "multi-device predict_many rows == the same rows through a one-device model (int8 outputs and logits)"},
            "roofline": None, "cpu_baseline": None}
    print(json.dumps(line), flush=True)
    m.close()
    if not ok:
        raise SystemExit("bench.py: multi-device result differs from the 1-GPU result")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="person_detect", choices=["person_detect", "speech", "sine"])
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU per step (default: BASELINE config)")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-conv2d", action="store_true")
    ap.add_argument("--flags", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        main_single_process(args)
        return

    import torch
    import torch.distributed as dist
    import microflow_rs_b200 as mf

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    # NCCL prints its version banner on stdout when the communicator is created (NCCL_DEBUG=VERSION on these boxes); stdout must
    # carry exactly one JSON line, so file descriptor 1 points at stderr until the first collectives are through.
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = args.workload
    batch = args.batch or DEFAULT_BATCH[wl]
    peaks, peak_src = load_peaks()

    m = mf.Model(MODELS / f"{wl}.tflite", device=local_rank, chunk=args.chunk, flags=args.flags)
    # ---- init: ONE NCCL broadcast of the static weights/constants blob from rank 0 (ranks != 0 clear theirs first)
    ptr, nbytes = m.blob()
    if world > 1 and nbytes:
        from microflow_rs_b200 import sharding
        blob = torch.as_tensor(sharding.DeviceBytes(ptr, nbytes), device=f"cuda:{local_rank}")
        sharding.broadcast_weights(dist, blob, rank)
        torch.cuda.synchronize()
    if saved_stdout is not None:
        dist.barrier()                      # communicator exists on every rank now
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    ie, oe = m.in_elems, m.out_elems
    # inputs: contiguous shard of the global sample range [rank*batch, (rank+1)*batch), R rotating batches so that
    # every step reads inputs that have left the 126 MB L2 (R * batch * in_elems > L2)
    R = max(2, int(np.ceil(160e6 / (batch * ie))) + 1)
    R = min(R, 64)
    host_in = [mf.PinnedBuffer((batch, ie), np.int8) for _ in range(min(R, 4))]
    for r_i, hb in enumerate(host_in):
        hb.array[:] = splitmix_bytes(SEEDS[wl] + r_i, batch * ie, offset=rank * batch * ie).reshape(batch, ie)
    d_in = [torch.from_numpy(host_in[i % len(host_in)].array).cuda() for i in range(R)]
    d_out = torch.empty((batch, oe), dtype=torch.float32, device="cuda")
    tstream = torch.cuda.Stream()          # a real (non-NULL) stream: kernels, timing events and per-layer events all live on it
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step(i):
        m.predict_many_device(d_in[i % R].data_ptr(), batch, d_out.data_ptr(), None, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    verified = verify_ranks(dist, torch, m, wl, rank, world, batch) if world > 1 else None
    barrier()
    # ---- timed region: exactly K steps, CUDA events on the launching stream (consecutive layers overlap their launch
    # ramps through programmatic dependent launch, so no per-layer events here)
    launches0 = m.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.window(True)
    ev0.record(tstream)
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record(tstream)
    barrier()
    sampler.window(False)
    ms = ev0.elapsed_time(ev1)
    launches = m.launch_count() - launches0
    # ---- second timed pass of the same K steps with the engine's per-layer CUDA events (on the same stream) between the
    # kernels: this is what isolates each kernel's launch duration for the roofline and the per-kernel shares.  The events
    # serialise the launches (no programmatic overlap), so this pass is a little slower than the headline one; both are reported.
    m.set_profiling(True)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.window(True)
    ev2.record(tstream)
    for i in range(args.steps):
        step(args.warmup + args.steps + i)
    ev3.record(tstream)
    barrier()
    sampler.window(False)
    ms_profiled = ev2.elapsed_time(ev3)
    layer_ms = m.layer_times_ms()
    m.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * batch * args.steps / (ms * 1e-3)

    # ---- e2e through the public C-ABI with HOST buffers (H2D of the inputs and D2H of the result inside the timed region).
    # `e2e.value`: K back-to-back mf_predict_many_quantized_async calls + one mf_model_synchronize (the H2D of call k+1 overlaps
    # the kernels of call k; every step still copies its own inputs in and its own results out).  `e2e.blocking`: the same K
    # steps through the blocking mf_predict_many_quantized (each call returns only after its results are on the host).
    host_out = [mf.PinnedBuffer((batch, oe), np.float32) for _ in range(2)]
    for i in range(3):
        m.predict_many_quantized(host_in[i % len(host_in)].array, out=host_out[i & 1].array)
    barrier()
    sampler.window(True)
    t0 = time.perf_counter()
    for i in range(args.steps):
        m.predict_many_quantized(host_in[i % len(host_in)].array, out=host_out[i & 1].array)
    torch.cuda.synchronize()
    e2e_block_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        m.predict_many_quantized_async(host_in[i % len(host_in)].array, host_out[i & 1].array)
    m.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.window(False)
    n_timed = len(sampler.samples)
    # The timed regions above last a few tens of milliseconds and an NVML query can take longer than that when several ranks
    # poll at once, so the same device-resident steps keep running (untimed) for another half second while the sampler
    # goes on: the clocks line then always rests on samples taken under this workload, and says how many fell where.
    sampler.window(True)
    t_end = time.perf_counter() + 0.5
    i = 0
    while time.perf_counter() < t_end:
        step(i)
        i += 1
        if i % 8 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    sampler.window(False)
    clocks = sampler.stop(n_timed)
    t = torch.tensor([e2e_s, e2e_block_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * batch * args.steps / float(t[0].item())
    e2e_block_val = world * batch * args.steps / float(t[1].item())

    # ---- ceiling of the e2e number: plain pinned-host -> device copies of one step's input over this box's PCIe link, issued the
    # way the engine issues them (two streams, half a step each); best of three rounds of ten steps
    link_dst = torch.empty(batch * ie, dtype=torch.int8, device="cuda")
    link_src = torch.from_numpy(host_in[0].array.reshape(-1))      # pinned (PinnedBuffer)
    half_b = (batch // 2) * ie
    cs = [torch.cuda.Stream(), torch.cuda.Stream()]
    link_gbs = 0.0
    for rnd in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            with torch.cuda.stream(cs[0]):
                link_dst[:half_b].copy_(link_src[:half_b], non_blocking=True)
            with torch.cuda.stream(cs[1]):
                link_dst[half_b:].copy_(link_src[half_b:], non_blocking=True)
        torch.cuda.synchronize()
        if rnd:                                                     # round 0 warms up
            link_gbs = max(link_gbs, 10 * batch * ie / (time.perf_counter() - t0) / 1e9)
    del link_dst

    # ---- latency of the reference's own call shape: ONE sample through mf_predict_quantized (host buffers, blocking)
    one = np.ascontiguousarray(host_in[0].array[0])
    for _ in range(20):
        m.predict_quantized(one)
    lat = []
    for _ in range(200):
        t0 = time.perf_counter()
        m.predict_quantized(one)
        lat.append(time.perf_counter() - t0)
    lat_us = 1e6 * float(np.median(lat))

    if rank == 0:
        # ---- roofline of the dominant kernel: group layers by kernel, take the largest share of the step
        groups = {}
        for L, tms in zip(m.layers, layer_ms):
            if not L["kernel"] or L["kernel"].startswith("none"):
                continue
            # layers that run inside another layer's launch ("(in fused_chain_kernel)") count towards that kernel: its time is recorded on
            # the first layer of the launch, its algorithmic bytes / MACs are the sum over the layers it executes
            inside = L["kernel"].startswith("(in ")
            kname = L["kernel"][4:-1] if inside else L["kernel"]
            g = groups.setdefault(kname, {"ms": 0.0, "bytes": 0, "macs": 0, "launches": 0})
            g["ms"] += float(tms)
            g["bytes"] += (L["bytes"] - L["weight_bytes"]) * batch * args.steps + L["weight_bytes"] * args.steps
            g["macs"] += L["macs"] * batch * args.steps
            g["launches"] += 0 if inside else 1
            # compulsory traffic of a launch: what enters it and what leaves it (for a fused launch the tensors between its layers never
            # reach HBM; for a one-layer launch this equals the algorithmic figure minus nothing)
            in_b = int(np.prod(L["in_shape"])) if L["in_shape"] else 0
            if not inside:
                g.setdefault("comp", 0)
                g["comp"] += (in_b + L["out_elems"]) * batch * args.steps
                g["_last_out"] = L["out_elems"]
            else:
                g["comp"] += (L["out_elems"] - g["_last_out"]) * batch * args.steps     # the launch's output is its LAST layer's output
                g["_last_out"] = L["out_elems"]
        total_layer_ms = sum(g["ms"] for g in groups.values()) or 1.0
        dom_name, dom = max(groups.items(), key=lambda kv: kv[1]["ms"])
        achieved = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic_latest.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get(dom_name)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                    "traffic": traffic, "peak_source": peak_src, "share_of_step": dom["ms"] / total_layer_ms,
                    "launches_per_step": dom["launches"], "avg_launch_us": 1e3 * dom["ms"] / (dom["launches"] * args.steps),
                    "measured_in": f"second timed pass of the same {args.steps} steps with per-layer CUDA events on the launching stream "
                                   f"({ms_profiled / args.steps:.4f} ms/step; headline pass without the events: {ms / args.steps:.4f} ms/step)",
                    "traffic_note": "traffic = dram__bytes_read+write of ONE captured launch of this kernel (profiles/traffic_latest.json says which layer and its algorithmic bytes); achieved aggregates all layers that run on this kernel",
                    "note": "achieved = algorithmic bytes (layer input+output per sample x batch + weights once per launch) of this kernel's layers / its summed "
                            "CUDA-event time inside the timed region"}
        kernels = {k: {"ms_per_step": g["ms"] / args.steps, "share": g["ms"] / total_layer_ms, "GBps": g["bytes"] / (g["ms"] * 1e-3) / 1e9 if g["ms"] > 0 else None,
                       "compulsory_GBps": g.get("comp", 0) / (g["ms"] * 1e-3) / 1e9 if g["ms"] > 0 else None,
                       "TOPS": 2 * g["macs"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else None} for k, g in groups.items()}
        line = {
            "metric": f"inferences/s {wl} int8", "value": value, "unit": "inferences/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "ms_per_step_with_layer_events": ms_profiled / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": f"{wl}.tflite int8, batch {batch} per GPU (BASELINE configs[2])" if wl == "person_detect" else f"{wl}.tflite int8, batch {batch} per GPU",
                       "global_batch": batch * world, "parallelism": f"dp{world} (independent samples, contiguous shards, one NCCL weight broadcast at init)",
                       "l2": f"inputs rotate over {R} device batches ({R * batch * ie / 1e6:.0f} MB > 126 MB L2)", "chunk": args.chunk or 8192},
            "e2e": {"value": e2e_val, "unit": "inferences/s", "h2d_bytes_per_step": world * batch * ie, "d2h_bytes_per_step": world * batch * oe * 4,
                    "api": "mf_predict_many_quantized_async x K + mf_model_synchronize (pinned host buffers)", "blocking": e2e_block_val,
                    "h2d_link_GBps": link_gbs, "link_bound_per_gpu": link_gbs * 1e9 / ie,
                    "note": "h2d_link_GBps = plain pinned-host -> device copies of one step's input on this rank (two streams, ten steps, best of 3); "
                            "link_bound_per_gpu = that bandwidth / input bytes per sample: the ceiling of e2e per GPU on this host"},
            "single_sample_latency_us": {"value": lat_us, "api": "mf_predict_quantized (one sample, host buffers, blocking; CUDA-graph replay)",
                                         "note": "median of 200 calls incl. the Python/ctypes call overhead"},
            "gpu_launches": int(launches), "clocks": clocks, "verified": verified, "roofline": roofline, "kernels": kernels,
            "layers": [{"i": i, "op": L["op"], "kernel": L["kernel"], "us_per_step": round(1e3 * float(t) / args.steps, 2),
                        "GBps": round((L["bytes"] - L["weight_bytes"]) * batch * args.steps / (float(t) * 1e-3) / 1e9, 1) if t > 0 else None}
                       for i, (L, t) in enumerate(zip(m.layers, layer_ms)) if not L["kernel"].startswith("none")],
            "model_level": {"int8_TOPS": value * 2 * sum(L["macs"] for L in m.layers) / 1e12,
                            "layerwise_GBps": value * sum(L["bytes"] - L["weight_bytes"] for L in m.layers) / 1e9,
                            "compulsory_GBps": value * (ie + oe * 4) / 1e9},
        }
        if world == 1 and not args.no_conv2d:
            try:
                line["conv2d"] = conv2d_roofline(torch, mf, peaks, steps=max(5, args.steps // 2), warmup=3)
            except Exception as e:  # the headline line must still print
                line["conv2d"] = {"error": repr(e)[:300]}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and verified is not None and not verified["ok"]:
        raise SystemExit("bench.py: ranks disagree on the common rows or rank 0 differs from the oracle")


if __name__ == "__main__":
    main()
