"""First look at the CUDA path on a B200: per-phase device times of a few workloads (not the bench)."""
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import quids_b200 as qb
from quids_b200 import qcgd

qb.config.profile = True


def report(tag, sym, t_ms):
    ph = sym.phase_ms
    nc = sym.num_object
    print(f"{tag}: N_c={nc:.3e} N_u={sym.num_object_after_interferences:.3e} wall={t_ms:.2f} ms -> {nc / t_ms * 1e3:.3e} children/s | "
          + " ".join(f"{k}={v:.2f}" for k, v in ph.items() if v > 0) + f" | sym HBM {sym.device_bytes / 1e9:.2f} GB", flush=True)


def timed(fn):
    qb.default_context().synchronize()
    t0 = time.perf_counter()
    fn()
    qb.default_context().synchronize()
    return (time.perf_counter() - t0) * 1e3


def modifier_probe(n):
    qb.config.align_byte_length = 8
    it = qb.Iteration()
    rng = np.random.default_rng(1)
    data = rng.integers(0, 256, size=8 * n, dtype=np.uint8)
    mags = np.zeros((n, 2))
    mags[:, 0] = 1 / math.sqrt(n)
    it.upload_packed(np.full(n, 8, np.uint32), mags, data)
    for name, params in (("phase", (0.3,)), ("ygate", (3,))):
        m = qb.Modifier(name, *params)
        for _ in range(2):
            qb.simulate(it, m)
        t = min(timed(lambda: qb.simulate(it, m)) for _ in range(5))
        print(f"modifier {name}: {n:.2e} objects of 8 B: {t:.3f} ms -> {n / t * 1e3:.3e} obj/s, {n * 40 / t / 1e6:.0f} GB/s algorithmic (40 B/obj)", flush=True)


def hadamard_probe(nq):
    qb.config.align_byte_length = 0
    qb.config.tolerance = 1e-30
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.append(bytes(nq), 1.0)
    for bit in range(nq):
        t = timed(lambda: qb.simulate(a, qb.Rule("hadamard", bit), b, sym))
        a, b = b, a
        if bit >= nq - 2:
            report(f"hadamard {nq}q doubling bit {bit}", sym, t)
    for rep in range(2):
        t = timed(lambda: qb.simulate(a, qb.Rule("hadamard", 0), b, sym))
        report(f"hadamard {nq}q interfering", sym, t)
        t = timed(lambda: qb.simulate(b, qb.Rule("hadamard", 0), a, sym))
        report(f"hadamard {nq}q doubling again", sym, t)


def qcgd_probe(n_graphs, rule, k, n_node=12):
    qb.config.align_byte_length = 8
    qb.config.tolerance = 1e-18
    sizes, data = qcgd.random_graphs(n_node, n_graphs, seed=0)
    mags = np.zeros((n_graphs, 2))
    mags[:, 0] = qcgd.read_state_magnitude(n_graphs)[0]
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.upload_packed(sizes, mags, data)
    for rep in range(3):
        t = timed(lambda: qb.simulate(a, rule, b, sym, k))
        report(f"qcgd {rule.name} {n_graphs:.0e} parents k={k:.0e} rep {rep}", sym, t)
    print(f"   next: {b.num_object} objects, {b.num_bytes / 1e6:.1f} MB, P={b.total_proba:.6f}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["mod", "had", "qcgd"]
    if "mod" in what:
        modifier_probe(10**8)
    if "had" in what:
        hadamard_probe(22)
    if "qcgd" in what:
        for n in (10**5, 10**6):
            qcgd_probe(n, qb.Rule("erase_create", math.pi / 4), n)
            qcgd_probe(n, qb.Rule("split_merge", math.pi / 4, math.pi / 4, math.pi / 4), n)
    if "big" in what:
        qcgd_probe(10**7, qb.Rule("erase_create", math.pi / 4), 10**7)
