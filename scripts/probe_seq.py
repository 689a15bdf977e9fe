"""split_merge / erase_create sequence on random graphs: per-phase times and table trace (QB_TABLE_TRACE=1)"""
import math, sys, time
import numpy as np
sys.path.insert(0, ".")
import quids_b200 as qb
from quids_b200 import qcgd
qb.config.profile = True
qb.config.align_byte_length, qb.config.tolerance = 8, 1e-18
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**6
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sizes, data = qcgd.random_graphs(12, n, seed=0)
mags = np.zeros((n, 2)); mags[:, 0] = qcgd.read_state_magnitude(n)[0]
a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
a.upload_packed(sizes, mags, data)
t = math.pi / 4
step, sm, ec = qb.Modifier("step"), qb.Rule("split_merge", t, t, t), qb.Rule("erase_create", t, 0.0, 0.0)
for it in range(iters):
    for rule in (sm, ec):
        qb.simulate(a, step)
        t0 = time.perf_counter()
        qb.simulate(a, rule, b, sym, n)
        ms = (time.perf_counter() - t0) * 1e3
        print(f"iter {it} {rule.name}: {ms:.1f} ms N_c={sym.num_object:.3e} N_u={sym.num_object_after_interferences:.3e} N_s={b.num_object} "
              + " ".join(f"{k}={v:.1f}" for k, v in sym.phase_ms.items() if v > 0.05), flush=True)
        a, b = b, a
