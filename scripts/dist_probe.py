"""Developer probe of the distributed path: the bench loop over N GPUs (torchrun), QB_DIST_TRACE timings of rank 0 on stderr.

    QB_DIST_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_probe.py [--parents P] [--passes K]
"""
import argparse
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quids_b200 as qb  # noqa: E402
from quids_b200 import qcgd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--parents", type=int, default=10**7)
ap.add_argument("--passes", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
qb.config.tolerance, qb.config.align_byte_length, qb.config.profile = 1e-18, 8, True
ctx = qb.default_context()
comm = qb.Communicator.from_torch(ctx, dist)
n = args.parents
sizes, data = qcgd.random_graphs(12, n, seed=rank)
mags = np.zeros((n, 2))
mags[:, 0] = qcgd.read_state_magnitude(n)[0]
a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
a.upload_packed(sizes, mags, data)
t = math.pi / 4
sm, ec, step = qb.Rule("split_merge", t, t, t), qb.Rule("erase_create", t, 0.0, 0.0), qb.Modifier("step")
for p in range(args.passes):
    for rule, x, y in ((sm, a, b), (ec, b, a)):
        qb.simulate(x, step)
        if rank == 0:
            print(f"---- pass {p} {rule.name}", file=sys.stderr, flush=True)
        qb.mpi_simulate(x, rule, y, sym, comm, n * world)
        if rank == 0:
            print(f"     phases {{k: round(v, 2) for k, v in sym.phase_ms.items() if v > 0.01}}".replace("{{", "{").replace("}}", "}"), file=sys.stderr, flush=True)
            print("     " + str({k: round(v, 2) for k, v in sym.phase_ms.items() if v > 0.01}), file=sys.stderr, flush=True)
comm.close()
dist.barrier()
dist.destroy_process_group()
