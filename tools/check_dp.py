"""Run under torchrun (one rank per GPU): every rank processes its shard of a global batch on its own GPU, results
are gathered with the engine's ncclAllGather, and rank 0 checks them BITWISE against a single-engine evaluation of
the whole batch (image i depends only on its global index; tile shapes never change per-element arithmetic).
usage: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_dp.py [size] [per_gpu_batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import numpy as np  # noqa: E402
import torch  # noqa: E402  (first: its NCCL must be the one that gets loaded)
import torch.distributed as dist  # noqa: E402
import y4b200  # noqa: E402
import y4_oracle as O  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 416
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
dist.init_process_group('gloo', rank=rank, world_size=world)
blob = O.synth_weights(seed=1).to_darknet_bytes()
eng = y4b200.Engine(img_size=size, max_batch=B * world if rank == 0 else B, precision=y4b200.PREC_FP16, device=local)
eng.load_darknet_bytes(blob)
uid = torch.from_numpy(eng.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8))
dist.broadcast(uid, 0)
eng.comm_init(rank, world, uid.numpy())
eng.synth_fill(0, rank * B, B)
eng.run_resident(B)
got = eng.allgather_results(B)
ok = True
if rank == 0:
    eng.synth_fill(0, 0, B * world)
    eng.run_resident(B * world)
    ref = eng.fetch_results(B * world)
    ok = all(np.array_equal(a, b) for a, b in zip(got, ref))
    print(f'DP CHECK world={world} size={size} per_gpu_batch={B}: gathered == single-GPU bitwise: {ok}; valid={got[3].tolist()}', flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
