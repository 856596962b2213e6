"""N>1 host path on CPU: world_size-2 gloo.  Each rank evaluates its shard of a global batch (with the oracle
standing in for the GPU), results are gathered with the package's helper, and must equal the single-process
evaluation of the whole batch — the same invariant bench.py --gpus N relies on (image i depends only on its
global index)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, size, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    import y4_oracle as O
    from y4b200 import dp
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = dp.shard_range(rank, world, total)
    heads = O.synth_heads(seed=9, batch=total, img_size=size, n_clusters=10)      # same seed on every rank
    local = O.decode_nms([h[lo:hi] for h in heads], size)
    counts = [dp.shard_range(r, world, total)[1] - dp.shard_range(r, world, total)[0] for r in range(world)]
    full = dp.allgather_results(dist, local, counts)
    if rank == 0:
        q.put([np.asarray(a) for a in full])
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range():
    from y4b200 import dp
    for total in (1, 5, 32, 255):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(r, world, total) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_equals_single_process():
    import torch.multiprocessing as mp
    import y4_oracle as O
    total, size, world = 5, 64, 2                     # ragged: 3 + 2 images
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, size, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = O.decode_nms(O.synth_heads(seed=9, batch=total, img_size=size, n_clusters=10), size)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
