"""One rule iteration of a chosen workload, for ncu captures (never a bench number)."""
import math
import sys

import numpy as np

sys.path.insert(0, ".")
import quids_b200 as qb
from quids_b200 import qcgd

rule_name = sys.argv[1] if len(sys.argv) > 1 else "erase_create"
n_graphs = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**6
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
qb.config.tolerance = 1e-18
sizes, data = qcgd.random_graphs(12, n_graphs, seed=0)
mags = np.zeros((n_graphs, 2))
mags[:, 0] = qcgd.read_state_magnitude(n_graphs)[0]
a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
a.upload_packed(sizes, mags, data)
rule = qb.Rule(rule_name, math.pi / 4, math.pi / 4 if rule_name == "split_merge" else 0.0, math.pi / 4 if rule_name == "split_merge" else 0.0)
for _ in range(reps):
    qb.simulate(a, rule, b, sym, n_graphs)
print(sym.num_object, sym.num_object_after_interferences, b.num_object)
