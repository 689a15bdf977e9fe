"""Developer probe: the bench loop (step; split_merge; step; erase_create) on one GPU, one JSON line per rule call with the
phase times -- the quick A/B tool behind the numbers in DESIGN.md (bench.py stays the contract line).

    python scripts/loop_probe.py [--parents N] [--passes P] [--skip S] [--binned 0|1|2] [--sort 0|1|2] [--c3]
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quids_b200 as qb  # noqa: E402
from quids_b200 import qcgd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--parents", type=int, default=10**7)
ap.add_argument("--passes", type=int, default=4)
ap.add_argument("--skip", type=int, default=2, help="passes not reported (warm-up)")
ap.add_argument("--binned", type=int, default=1)
ap.add_argument("--sort", type=int, default=1)
ap.add_argument("--c3", action="store_true", help="also the 24-qubit hadamard steps (configs[2])")
ap.add_argument("--tag", default="")
args = ap.parse_args()

qb.config.tolerance, qb.config.align_byte_length, qb.config.profile = 1e-18, 8, True
qb.config.binned_inserts, qb.config.locality_sort = args.binned, args.sort
ctx = qb.default_context()
stream = torch.cuda.ExternalStream(ctx.stream)
t = math.pi / 4
sm, ec, step = qb.Rule("split_merge", t, t, t), qb.Rule("erase_create", t, 0.0, 0.0), qb.Modifier("step")


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.synchronize()
    e0.record(stream)
    fn()
    e1.record(stream)
    ctx.synchronize()
    return e0.elapsed_time(e1)


n = args.parents
sizes, data = qcgd.random_graphs(12, n, seed=0)
mags = np.zeros((n, 2))
mags[:, 0] = qcgd.read_state_magnitude(n)[0]
a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
a.upload_packed(sizes, mags, data)
del sizes, mags, data
for p in range(args.passes):
    for rule, x, y in ((sm, a, b), (ec, b, a)):
        qb.simulate(x, step)
        ms = timed(lambda: qb.simulate(x, rule, y, sym, n))
        if p >= args.skip:
            print(json.dumps({"tag": args.tag, "pass": p, "rule": rule.name, "ms": round(ms, 3), "N_c": sym.num_object, "N_u": sym.num_object_after_interferences, "N_s": y.num_object,
                              "children_per_s": sym.num_object / ms * 1e3, "phase_ms": {k: round(v, 3) for k, v in sym.phase_ms.items() if v > 0}}), flush=True)

if args.c3:
    del a, b
    qb.config.align_byte_length, qb.config.tolerance = 0, 1e-30
    nq = 24
    a, b = qb.Iteration(), qb.Iteration()
    a.append(bytes(nq), 1.0)
    for bit in range(nq):
        qb.simulate(a, qb.Rule("hadamard", bit), b, sym)
        a, b = b, a
    for rep in range(3):
        for what in ("interfering 2^24 -> 2^23", "doubling 2^23 -> 2^24"):
            ms = timed(lambda: qb.simulate(a, qb.Rule("hadamard", 0), b, sym))
            print(json.dumps({"tag": args.tag, "c3": what, "rep": rep, "ms": round(ms, 3), "N_c": sym.num_object, "N_u": sym.num_object_after_interferences,
                              "children_per_s": sym.num_object / ms * 1e3, "phase_ms": {k: round(v, 3) for k, v in sym.phase_ms.items() if v > 0}}), flush=True)
            a, b = b, a
