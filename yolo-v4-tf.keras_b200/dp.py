"""Data-parallel host plumbing (one process per GPU).  Images are independent units: rank r owns a contiguous
range of global image indices; the only exchange is an all-gather of the fixed-size result arrays.  On GPUs the
gather is the engine's ncclAllGather (y4_allgather_results); these helpers are the host-side index arithmetic
and a backend-agnostic gather used by callers that already hold results on the host (and by the gloo tests)."""
import numpy as np


def shard_range(rank, world, total):
    """Contiguous [lo, hi) of `total` items for `rank`; the first total % world ranks get one extra."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allgather_results(dist, local, counts):
    """local: tuple of arrays whose axis 0 is this rank's images; counts[r] = images of rank r.
    Returns the arrays of all ranks concatenated in rank order (ragged shards are padded for the collective)."""
    import torch
    world = dist.get_world_size()
    mx = max(counts)
    out = []
    for a in local:
        a = np.ascontiguousarray(a)
        pad = np.zeros((mx,) + a.shape[1:], a.dtype)
        pad[:a.shape[0]] = a
        t = torch.from_numpy(pad)
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t)
        out.append(np.concatenate([b.numpy()[:counts[r]] for r, b in enumerate(bufs)], axis=0))
    return tuple(out)
