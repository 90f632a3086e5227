"""World-size-2 gloo test (CPU) of the host-side multi-GPU logic: contiguous shards partition the sample range, the one-time
weight broadcast fills every rank's blob from rank 0, and the sharded result equals the single-process result row for row.
The per-shard compute here is the oracle (tests may use it); on the GPU box bench.py runs the same logic with the CUDA model."""
import os
import socket

import numpy as np
import pytest

from microflow_rs_b200.sharding import shard_range


def test_shard_ranges_partition():
    for n in (0, 1, 7, 8192, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= lo <= hi <= n for lo, hi in spans)
            assert max(hi - lo for lo, hi in spans) <= -(-n // world) if n else True


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    sys.path.insert(0, str(root / "tests"))
    import torch
    import torch.distributed as dist
    import oracle
    from conftest import MODELS, splitmix_bytes
    from microflow_rs_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. one-time weight broadcast: rank 0 holds the "blob", the other rank starts from garbage
        ref_blob = torch.arange(1000, dtype=torch.int64).to(torch.uint8)
        blob = ref_blob.clone() if rank == 0 else torch.full((1000,), 7, dtype=torch.uint8)
        sharding.broadcast_weights(dist, blob, rank)
        ok_blob = bool(torch.equal(blob, ref_blob))
        # 2. sharded predict_many == single-process predict_many, row for row
        o = oracle.Model(MODELS / "speech.tflite", fast=True)
        n = 37
        xs = splitmix_bytes(0x5EED0004, n * o.in_elems).reshape(n, -1)
        full, (lo, hi) = sharding.predict_many_sharded(dist, lambda x: o.predict_many_quantized(x)[0], xs, rank, world, o.out_elems)
        want = o.predict_many_quantized(xs)[0]
        q.put((rank, ok_blob, bool(np.array_equal(full, want)), (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_predict_many_and_weight_broadcast():
    mp = pytest.importorskip("torch.multiprocessing")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert [r[1] for r in res] == [True, True]          # both ranks hold rank 0's weights
    assert [r[2] for r in res] == [True, True]          # gathered result identical to the single-process result
    assert res[0][3] == (0, 19) and res[1][3] == (19, 37)
