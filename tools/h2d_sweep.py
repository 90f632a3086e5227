#!/usr/bin/env python
"""Host -> device copy bandwidth with 1 / 2 / 4 / 8 GPUs copying AT THE SAME TIME (the limiter of bench.py's e2e number at N >= 4).

One process per GPU (as bench.py under torchrun); every process copies a person_detect step's input (75.5 MB, two streams, half
each) from pinned host memory in a loop for ~1 s between two barriers.  Variants of where / how the pinned buffer is allocated:
    default     pinned buffer allocated and first touched by the process as launched
    bind        the process first binds itself (CPU affinity + first touch) to the cores `nvidia-smi topo -m` lists for its GPU
Prints one line per (variant, concurrency): per-GPU GB/s min / median / max and the sum.  Usage: tools/h2d_sweep.py [gpus]"""
import multiprocessing as mp
import os
import subprocess
import sys
import time


def gpu_cpu_affinity():
    """GPU index -> CPU list from `nvidia-smi topo -m` (column 'CPU Affinity')."""
    try:
        txt = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    except Exception:
        return {}
    aff = {}
    hdr = None
    for line in txt.splitlines():
        cols = line.split("\t")
        cols = [c.strip() for c in cols if c.strip() != ""]
        if not cols:
            continue
        if hdr is None and any("CPU Affinity" in c for c in cols):
            hdr = cols
            continue
        if hdr and cols[0].startswith("GPU"):
            try:
                k = int(cols[0][3:])
                i = [j for j, c in enumerate(hdr) if "CPU Affinity" in c][0] + 1      # the header has no label for the first column
                aff[k] = cols[i]
            except Exception:
                pass
    return aff


def parse_cpus(spec):
    cpus = []
    for part in spec.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus += list(range(int(a), int(b) + 1))
        elif part.strip().isdigit():
            cpus.append(int(part))
    return cpus


def worker(rank, world, variant, aff, barrier, q):
    os.environ["CUDA_VISIBLE_DEVICES"] = str(rank)
    if variant == "bind" and rank in aff:
        cpus = parse_cpus(aff[rank])
        if cpus:
            try:
                os.sched_setaffinity(0, cpus)
            except Exception:
                pass
    import torch
    torch.cuda.init()
    n = 75497472
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)                                       # first touch from this (possibly core-bound) process
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    half = n // 2

    def step():
        with torch.cuda.stream(s[0]):
            dev[:half].copy_(host[:half], non_blocking=True)
        with torch.cuda.stream(s[1]):
            dev[half:].copy_(host[half:], non_blocking=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    barrier.wait()
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < 1.0:
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        k += 4
    dt = time.perf_counter() - t0
    barrier.wait()
    q.put((rank, variant, world, k * n / dt / 1e9))


def main():
    gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    aff = gpu_cpu_affinity()
    print("nvidia-smi topo CPU affinity per GPU:", aff)
    print("host:", os.cpu_count(), "logical CPUs; this process may run on", len(os.sched_getaffinity(0)))
    ctx = mp.get_context("spawn")
    for variant in ("default", "bind"):
        for world in (1, 2, 4, 8):
            if world > gpus:
                continue
            barrier = ctx.Barrier(world)
            q = ctx.Queue()
            ps = [ctx.Process(target=worker, args=(r, world, variant, aff, barrier, q)) for r in range(world)]
            for p in ps:
                p.start()
            res = [q.get(timeout=120) for _ in ps]
            for p in ps:
                p.join(timeout=60)
            bw = sorted(r[3] for r in res if r[3] is not None)
            if bw:
                print(f"{variant:8s} {world} GPUs copying: per GPU min {bw[0]:6.1f} / median {bw[len(bw) // 2]:6.1f} / max {bw[-1]:6.1f} GB/s, sum {sum(bw):7.1f} GB/s"
                      f" -> {sum(bw) * 1e9 / 9216 / 1e6:6.1f} M person_detect samples/s", flush=True)


if __name__ == "__main__":
    main()
