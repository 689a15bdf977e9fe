"""The reference's own loop (qcgd_test.cpp: step; split_merge; step; erase_create) on a saturated state, for ncu
launch lists (never a bench number).  argv: parents, iterations."""
import math
import sys

import numpy as np

sys.path.insert(0, ".")
import quids_b200 as qb
from quids_b200 import qcgd

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**6
iterations = int(sys.argv[2]) if len(sys.argv) > 2 else 2
qb.config.tolerance = 1e-18
qb.config.profile = True
sizes, data = qcgd.random_graphs(12, n, seed=0)
mags = np.zeros((n, 2))
mags[:, 0] = qcgd.read_state_magnitude(n)[0]
a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
a.upload_packed(sizes, mags, data)
t = math.pi / 4
step, sm, ec = qb.Modifier("step"), qb.Rule("split_merge", t, t, t), qb.Rule("erase_create", t, 0.0, 0.0)
for it in range(iterations):
    for rule in (sm, ec):
        qb.simulate(a, step)
        qb.simulate(a, rule, b, sym, n)
        print(it, rule.name, sym.num_object, sym.num_object_after_interferences, b.num_object, {k: round(v, 3) for k, v in sym.phase_ms.items() if v > 0}, flush=True)
        a, b = b, a
