"""Host-side logic of the multi-GPU deployment: independent samples, contiguous shards, one
broadcast of the static weight blob at init, no collective on the inference path.  torch.distributed is plumbing only."""
import numpy as np


def shard_range(n, rank, world):
    """Rank r of G owns samples [r*ceil(n/G), min(n, (r+1)*ceil(n/G)))."""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class DeviceBytes:
    """Zero-copy view of a raw device allocation for torch.as_tensor (used to broadcast mf_model_blob in place)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def broadcast_weights(dist, blob, rank, src=0):
    """Init-time weight distribution: ranks != src clear their copy, then ONE broadcast fills it from `src`.
    `blob` is a torch uint8 tensor (device memory under NCCL, host memory under gloo in the CPU tests)."""
    if rank != src:
        blob.zero_()
    dist.broadcast(blob, src=src)
    return blob


def predict_many_sharded(dist, predict_fn, xs, rank, world, out_elems, gather=True):
    """Runs `predict_fn` on this rank's contiguous shard of `xs` [n, in_elems].  Nothing is exchanged on the inference path;
    `gather=True` additionally assembles the global result on every rank (verification / single-consumer deployments)."""
    import torch
    n = xs.shape[0]
    lo, hi = shard_range(n, rank, world)
    local = predict_fn(xs[lo:hi]) if hi > lo else np.zeros((0, out_elems), np.float32)
    if not gather:
        return local, (lo, hi)
    per = -(-n // world)
    pad = np.zeros((per, out_elems), np.float32)
    pad[: hi - lo] = local
    parts = [torch.zeros((per, out_elems), dtype=torch.float32) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(pad))
    full = np.concatenate([p.numpy() for p in parts])[:n]
    return full, (lo, hi)
