#!/usr/bin/env python
"""Pinned-host -> device copy bandwidth of this box (the ceiling of bench.py's e2e number: 9216 B per person_detect sample)."""
import torch
n = 75497472
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for size in (n, n // 2, n // 8):
    for _ in range(3):
        d[:size].copy_(h[:size], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        d[:size].copy_(h[:size], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"H2D {size / 1e6:8.1f} MB  {ms:7.3f} ms  {size / ms / 1e6:6.1f} GB/s  -> {size / 9216 / ms / 1e3:6.2f} M person_detect samples/s")
