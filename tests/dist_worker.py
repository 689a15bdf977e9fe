"""One rank of the 2+ GPU parity check of the distributed path (launched by tests/test_gpu_dist.py
through torch.distributed.run).  The full state is built identically on every rank; rank r keeps the
parents i with i % world == r; the union of the results must equal the single-process CPU checker
on the full state (north_star: match quids::simulate on the gathered state)."""
import math
import os
import sys

import numpy as np
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402
import quids_b200 as qb  # noqa: E402
from gpu_engine import RULE_NAMES, NPARAMS  # noqa: E402


def share(p: orc.Packed, rank, world) -> orc.Packed:
    objs = p.objects()
    idx = [i for i in range(p.n) if i % world == rank]
    return orc.Packed.from_objects([objs[i] for i in idx], p.cmags[idx])


def run_case(comm, port, state, rid, params, k, tol, qcgd, what):
    rank, world = dist.get_rank(), dist.get_world_size()
    qb.config.tolerance = tol
    qb.config.align_byte_length = 8
    mine = share(state, rank, world)
    it, nxt, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    it.upload_packed(mine.sizes, mine.mags, mine.data)
    node = qb.mpi_simulate(it, qb.Rule(RULE_NAMES[rid], *params[:NPARAMS[rid]]), nxt, sym, comm, k)
    sizes, mags, data = nxt.download_packed()
    counts = comm.allreduce_u64([sym.num_object, sym.num_object_after_interferences, nxt.num_object])
    nodes = comm.allreduce_f64([node])
    gathered = [None] * world
    dist.all_gather_object(gathered, (sizes, mags, data, nxt.total_proba))
    if rank != 0:
        return
    assert abs(nodes[0] - 1) < 1e-12 or counts[2] == 0, nodes
    totals = {g[3] for g in gathered}
    assert len(totals) == 1, f"{what}: total_proba differs between ranks {totals}"
    got = orc.Packed(np.concatenate([g[0] for g in gathered]), np.concatenate([g[1] for g in gathered]), np.concatenate([g[2] for g in gathered]),
                     gathered[0][3])
    want, nc, nu = port.simulate(state, rid, params, k, tol)
    assert (int(counts[0]), int(counts[1])) == (nc, nu), f"{what}: counters {counts[:2]} vs {(nc, nu)}"
    hg, hw = port.hash_objects(got, rid), port.hash_objects(want, rid)
    if k == orc.NO_TRUNCATION or k >= nu:
        orc.assert_same_state(got, hg, want, hw, qcgd, what=what)
    else:
        full_src = state
        if k < state.n:
            order = np.sort(np.argsort(-(np.abs(state.cmags) ** 2), kind="stable")[:k])
            objs = state.objects()
            full_src = orc.Packed.from_objects([objs[j] for j in order], state.cmags[order])
        full, _, _ = port.simulate(full_src, rid, params, orc.NO_TRUNCATION, tol)
        orc.assert_same_truncated(got, hg, want, hw, full, port.hash_objects(full, rid), min(k, full.n), qcgd, what=what)
    print(f"ok {what}: N_c={nc} N_u={nu} N_s={got.n}", flush=True)


def main():
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    port = orc.Oracle(orc.PORT_SO)
    ctx = qb.default_context()
    comm = qb.Communicator.from_torch(ctx, dist)
    rng = np.random.default_rng(7)
    base = port.qcgd_random_state(8, 600, 4)
    mags = rng.normal(size=(600, 2))
    state = orc.Packed(base.sizes, mags / np.sqrt((mags ** 2).sum()), base.data)
    p = [0.37, 0.21, -0.4]
    for rid in orc.QCGD_RULES:
        run_case(comm, port, state, rid, p, orc.NO_TRUNCATION, 1e-18, True, f"rule {rid} no truncation")
        run_case(comm, port, state, rid, p, 900, 1e-18, True, f"rule {rid} children truncated")
        run_case(comm, port, state, rid, p, 250, 1e-18, True, f"rule {rid} parents and children truncated")
    # equal magnitudes: the ties at the threshold must be shared out between the ranks, exactly k kept
    tied = port.qcgd_random_state(7, 300, 9)
    run_case(comm, port, tied, orc.RULE_ERASE_CREATE, [math.pi / 4, 0, 0], 777, 1e-18, True, "ties across ranks")
    # one rank empty: fewer objects than ranks
    tiny = orc.Packed.from_objects([bytes([0, 1, 0, 1])], [1.0])
    run_case(comm, port, tiny, orc.RULE_HADAMARD, [2], orc.NO_TRUNCATION, 1e-30, False, "single object, other ranks empty")
    # interference across ranks: 12-qubit register to full superposition and back
    reg = orc.Packed.from_objects([bytes(10)], [1.0])
    cur = reg
    for bit in list(range(10)) + list(reversed(range(10))):
        want, _, _ = port.simulate(cur, orc.RULE_HADAMARD, [bit])
        run_case(comm, port, cur, orc.RULE_HADAMARD, [bit], orc.NO_TRUNCATION, 1e-30, False, f"register H({bit}) from {cur.n} objects")
        cur = want
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
