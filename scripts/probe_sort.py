import sys, math, numpy as np
sys.path.insert(0,".")
import quids_b200 as qb
from quids_b200 import qcgd
qb.config.profile=True; qb.config.tolerance=1e-18
for n in (10**6, 10**7):
    sizes,data=qcgd.random_graphs(12,n,seed=0); mags=np.zeros((n,2)); mags[:,0]=qcgd.read_state_magnitude(n)[0]
    a,b,sym=qb.Iteration(),qb.Iteration(),qb.SymbolicIteration()
    a.upload_packed(sizes,mags,data)
    for mode in (0,1):
        qb.config.locality_sort=mode
        for i in range(3):
            qb.simulate(a, qb.Rule("erase_create", math.pi/4), b, sym, n)
        print(n, "sort", mode, {k:round(v,2) for k,v in sym.phase_ms.items() if v}, flush=True)
    del a,b,sym
