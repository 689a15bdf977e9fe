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


def skewed_share(p: orc.Packed, rank, world) -> orc.Packed:
    """rank 0 holds 70 % of the objects, the others split the rest: work for the load balancer"""
    objs = p.objects()
    cut = (p.n * 7) // 10
    idx = list(range(cut)) if rank == 0 else [i for i in range(cut, p.n) if (i - cut) % (world - 1) == rank - 1]
    return orc.Packed.from_objects([objs[i] for i in idx], p.cmags[idx]) if idx else orc.Packed.from_objects([], [])


def multiset(sizes, mags, data):
    begin = np.concatenate([[0], np.cumsum(sizes.astype(np.int64))])
    return sorted((data[begin[i]:begin[i + 1]].tobytes(), float(mags[i, 0]), float(mags[i, 1])) for i in range(len(sizes)))


def gather_states(it):
    world = dist.get_world_size()
    got = [None] * world
    dist.all_gather_object(got, it.download_packed())
    return got


def migration_cases(comm, port):
    """send/receive, distribute, gather, equalize (quids_mpi.hpp:124-231, 903-1077): objects are moved, never changed"""
    rank, world = dist.get_rank(), dist.get_world_size()
    qb.config.align_byte_length = 8
    base = port.qcgd_random_state(6, 240, 11)
    ragged, _, _ = port.simulate(base, orc.RULE_SPLIT_MERGE, [0.4, 0.1, 0.2], orc.NO_TRUNCATION, 1e-18)  # objects of many sizes
    assert len(set(ragged.sizes.tolist())) > 3
    want = multiset(ragged.sizes, ragged.mags, ragged.data)
    root = world - 1

    # distribute_objects from `root`: shares of quids_mpi.hpp:1042, taken from the tail in rank order
    it = qb.Iteration()
    if rank == root:
        it.upload_packed(ragged.sizes, ragged.mags, ragged.data, total_proba=0.75)
    it.distribute_objects(comm, root)
    states = gather_states(it)
    if rank == 0:
        n = ragged.n
        expect = {root: n - sum((n * (node + 1)) // world - (n * node) // world for node in range(1, world))}
        for node in range(1, world):
            expect[node - 1 if node <= root else node] = (n * (node + 1)) // world - (n * node) // world
        assert [len(s[0]) for s in states] == [expect[r] for r in range(world)], ([len(s[0]) for s in states], expect)
        merged = sorted(sum((multiset(*s) for s in states), []))
        assert merged == want, "distribute_objects changed the objects"
        print(f"ok distribute_objects: {[len(s[0]) for s in states]}", flush=True)
    # the layout that arrived must be usable: one rule iteration on the distributed state equals the oracle's
    nxt, sym = qb.Iteration(), qb.SymbolicIteration()
    qb.config.tolerance = 1e-18
    qb.mpi_simulate(it, qb.Rule("erase_create", 0.3, 0.1, 0.2), nxt, sym, comm, orc.NO_TRUNCATION)
    outs = gather_states(nxt)
    if rank == 0:
        ref, _, _ = port.simulate(ragged, orc.RULE_ERASE_CREATE, [0.3, 0.1, 0.2], orc.NO_TRUNCATION, 1e-18)
        got = orc.Packed(np.concatenate([g[0] for g in outs]), np.concatenate([g[1] for g in outs]), np.concatenate([g[2] for g in outs]), ref.total_proba)
        orc.assert_same_state(got, port.hash_objects(got, orc.RULE_ERASE_CREATE), ref, port.hash_objects(ref, orc.RULE_ERASE_CREATE), True,
                              what="rule iteration after distribute_objects")
        print("ok rule iteration after distribute_objects", flush=True)

    # gather_objects back to rank 0
    it.gather_objects(comm, 0)
    states = gather_states(it)
    if rank == 0:
        assert [len(s[0]) for s in states] == [ragged.n] + [0] * (world - 1)
        assert multiset(*states[0]) == want, "gather_objects changed the objects"
        print("ok gather_objects", flush=True)
    at_rank0 = orc.Packed(*states[0])

    # send_objects / receive_objects between rank 0 and rank 1; asking for more room than allowed moves nothing
    if rank == 0:
        assert it.send_objects(17, 1, comm) == 17
        assert it.send_objects(5, 1, comm) == 0 and it.num_object == ragged.n - 17  # refused by the receiver
        assert it.send_objects(0, 1, comm) == 0
    elif rank == 1:
        assert it.receive_objects(0, comm) == 17
        assert it.receive_objects(0, comm, max_mem=16) == 0 and it.num_object == 17
        assert it.receive_objects(0, comm) == 0
    states = gather_states(it)
    if rank == 0:
        assert [len(s[0]) for s in states][:2] == [ragged.n - 17, 17]
        assert multiset(*states[1]) == want_tail(at_rank0, 17), "send_objects must take the tail"
        assert sorted(sum((multiset(*s) for s in states), [])) == want
        print("ok send_objects / receive_objects", flush=True)

    # equalize by objects: every round pairs the fullest rank with the emptiest
    before = comm.allreduce_u64([it.num_object], op_max=True)[0]
    rounds = it.equalize(comm, max_rounds=8, min_equalize_size=10, equalize_inbalance=0.05, min_equalize_step=0.0)
    states = gather_states(it)
    after = max(len(s[0]) for s in states)
    if rank == 0:
        assert rounds >= 1 and after < before, (rounds, before, after)
        if world == 2:
            assert after <= int(1.06 * ragged.n / world) + 1, (before, after)
        assert sorted(sum((multiset(*s) for s in states), [])) == want, "equalize changed the objects"
        print(f"ok equalize by objects: {rounds} rounds, max per rank {before} -> {after}", flush=True)

    # equalize by children of a rule (equalize_symbolic): the children counts of a pair meet in the middle
    it.gather_objects(comm, 0)
    rule = qb.Rule("erase_create", 0.3, 0.1, 0.2)
    c_before = comm.allreduce_u64([it.count_children(rule)], op_max=True)[0]
    c_total = comm.allreduce_u64([it.count_children(rule)])[0]
    rounds = it.equalize(comm, rule, max_rounds=8, min_equalize_size=10, equalize_inbalance=0.05, min_equalize_step=0.0)
    c_after = comm.allreduce_u64([it.count_children(rule)], op_max=True)[0]
    states = gather_states(it)
    if rank == 0:
        assert c_before == c_total and rounds >= 1 and c_after < c_before
        if world == 2:
            assert c_after <= 0.56 * c_total, (c_after, c_total)
        assert sorted(sum((multiset(*s) for s in states), [])) == want
        print(f"ok equalize by children: {rounds} rounds, max children per rank {c_before} -> {c_after} of {c_total}", flush=True)


def want_tail(p: orc.Packed, n):
    sizes, mags = p.sizes[-n:], p.mags[-n:]
    begin = int(p.sizes[:-n].astype(np.int64).sum())
    return multiset(sizes, mags, p.data[begin:])


def run_case(comm, port, state, rid, params, k, tol, qcgd, what, share=share, equalize=0, family_routing=1, big=False, binned=1):
    rank, world = dist.get_rank(), dist.get_world_size()
    qb.config.tolerance = tol
    qb.config.align_byte_length = 8
    qb.config.equalize = equalize
    qb.config.family_routing = family_routing
    qb.config.binned_inserts = binned
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
    if big:  # thousands of contributions per object, some nearly cancelling: the large-state rule (tests/bigcmp.py)
        import bigcmp
        bigcmp.compare(got, hg, want, hw, qcgd, truncated_k=None if (k == orc.NO_TRUNCATION or k >= nu) else k, bytes_sample=None, what=what)
    elif k == orc.NO_TRUNCATION or k >= nu:
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


def failure_cases(comm, port, state):
    """a failure that only ONE rank sees must surface as an error on EVERY rank from the same call (no rank left waiting in
    a collective), and the communicator must stay usable (capacity checks are data dependent per rank)"""
    rank = dist.get_rank()
    qb.config.tolerance, qb.config.align_byte_length, qb.config.equalize = 1e-18, 8, 0
    mine = share(state, rank, dist.get_world_size())
    # the record-exchange path (a rule without families) and the family-routed path (erase_create), every phase of each
    plans = [(qb.Rule("split_merge", 0.37, 0.21, -0.4), ("local", "partition", "owner", "return", "finalize")),
             (qb.Rule("erase_create", 0.37, 0.21, -0.4), ("route", "local", "finalize"))]
    for rule, phases in plans:
        for phase in phases:
            it, nxt, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
            it.upload_packed(mine.sizes, mine.mags, mine.data)
            os.environ["QB_DIST_INJECT_FAILURE"] = f"1:{phase}"
            try:
                qb.mpi_simulate(it, rule, nxt, sym, comm, 900)
                raised = None
            except qb.QuidsError as e:
                raised = str(e)
            finally:
                del os.environ["QB_DIST_INJECT_FAILURE"]
            assert raised is not None, f"rank {rank}: no error although rank 1 failed in phase {phase} of {rule.name}"
            assert ("injected failure" in raised) == (rank == 1), raised
            assert rank == 1 or "rank 1 failed" in raised, raised
    if rank == 0:
        print("ok failure on one rank stops every rank", flush=True)
    run_case(comm, port, state, orc.RULE_ERASE_CREATE, [0.37, 0.21, -0.4], 900, 1e-18, True, "communicator usable after agreed failures")


def automatic_budget_cases(comm, port, state):
    """max_num_object = 0 on the distributed path (the reference's default, quids_mpi.hpp:423): per-rank parent budget, children
    budget agreed between the ranks.  With all parents kept, the result must be the oracle's top-k for the k that was agreed."""
    rank, world = dist.get_rank(), dist.get_world_size()
    qb.config.tolerance, qb.config.align_byte_length, qb.config.equalize = 1e-18, 8, 0
    mine = share(state, rank, world)
    rid, params = orc.RULE_ERASE_CREATE, [0.37, 0.21, -0.4]
    full, nc_full, nu_full = port.simulate(state, rid, params, orc.NO_TRUNCATION, 1e-18)
    truncated = 0
    try:
        for budget in (1 << 31, 40 << 20, 20 << 20, 10 << 20, 5 << 20, 3 << 20):
            qb.config.memory_budget = budget
            it, nxt, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
            it.upload_packed(mine.sizes, mine.mags, mine.data)
            try:
                qb.mpi_simulate(it, qb.Rule(RULE_NAMES[rid], *params), nxt, sym, comm, 0)
            except qb.QuidsError as e:
                assert "automatic budget" in str(e) or "failed during" in str(e), str(e)
                break
            counts = comm.allreduce_u64([sym.num_object, sym.num_object_after_interferences, nxt.num_object])
            gathered = [None] * world
            dist.all_gather_object(gathered, (*nxt.download_packed(), nxt.total_proba))
            if rank != 0 or int(counts[0]) != nc_full:
                continue  # parents were cut per rank: no single-process counterpart
            k = int(counts[2])
            assert int(counts[1]) == nu_full and 1 <= k <= nu_full
            got = orc.Packed(np.concatenate([g[0] for g in gathered]), np.concatenate([g[1] for g in gathered]), np.concatenate([g[2] for g in gathered]), gathered[0][3])
            want, _, _ = port.simulate(state, rid, params, k, 1e-18)
            hg, hw = port.hash_objects(got, rid), port.hash_objects(want, rid)
            if k == nu_full:
                orc.assert_same_state(got, hg, want, hw, True, what=f"automatic budget {budget}")
            else:
                truncated += 1
                orc.assert_same_truncated(got, hg, want, hw, full, port.hash_objects(full, rid), k, True, what=f"automatic budget {budget}")
    finally:
        qb.config.memory_budget = 0
    if rank == 0:
        assert truncated >= 1, "no budget of the sweep truncated the children"
        print(f"ok automatic budget on the distributed path ({truncated} truncating budgets)", flush=True)


def main():
    # the compaction of the family-routed path lists only the entries above a rank-local floor (default: tables of 2^24 slots and
    # more; here every table), and simulate() checks the floors against the global threshold
    os.environ["QB_COMPACT_FILTER_MIN"] = "0"
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
        if rid != orc.RULE_SPLIT_MERGE:  # erase_create / coin went through the family routing above: the record exchange must agree
            run_case(comm, port, state, rid, p, 900, 1e-18, True, f"rule {rid} children truncated, record exchange", family_routing=0)
            run_case(comm, port, state, rid, p, 250, 1e-18, True, f"rule {rid} parents and children truncated, record exchange", family_routing=0)
    # binned interference forced on both sides of the record exchange (local bins, owner-side bins with the fair representative)
    for rid in orc.QCGD_RULES:
        run_case(comm, port, state, rid, p, orc.NO_TRUNCATION, 1e-18, True, f"rule {rid} no truncation, bins on both sides", family_routing=0, binned=2)
        run_case(comm, port, state, rid, p, 250, 1e-18, True, f"rule {rid} parents and children truncated, bins on both sides", family_routing=0, binned=2)
    # a ragged grown state (names of many shapes, nodes merged and split): families must still be closed under erase_create / coin
    grown, _, _ = port.simulate(port.qcgd_random_state(9, 2000, 8), orc.RULE_SPLIT_MERGE, [0.4, 0.3, 0.2], orc.NO_TRUNCATION, 1e-18)
    gm = np.random.default_rng(3).normal(size=(grown.n, 2))
    grown = orc.Packed(grown.sizes, gm / np.sqrt((gm ** 2).sum()), grown.data)
    for rid in (orc.RULE_ERASE_CREATE, orc.RULE_COIN):
        run_case(comm, port, grown, rid, p, orc.NO_TRUNCATION, 1e-18, True, f"rule {rid} on a grown state, routed by family", big=True)
        run_case(comm, port, grown, rid, p, 20000, 1e-18, True, f"rule {rid} on a grown state, routed by family, truncated", big=True)
    # equal magnitudes: the ties at the threshold must be shared out between the ranks, exactly k kept
    tied = port.qcgd_random_state(7, 300, 9)
    run_case(comm, port, tied, orc.RULE_ERASE_CREATE, [math.pi / 4, 0, 0], 777, 1e-18, True, "ties across ranks")
    # at scale: 7e5 unique children, so that every rank's share takes the large-input path of the global select (candidates
    # copied out after two all-reduced digits), once with distinct probabilities and once saturated with ties
    big = port.qcgd_random_state(12, 6000, 13, 1.0)
    big_mags = np.random.default_rng(6).normal(size=(big.n, 2))
    big = orc.Packed(big.sizes, big_mags / np.sqrt((big_mags ** 2).sum()), big.data)
    run_case(comm, port, big, orc.RULE_ERASE_CREATE, [math.pi / 4, 0.1, 0.2], 100000, 1e-18, True, "global select at scale")
    run_case(comm, port, port.qcgd_random_state(12, 6000, 3), orc.RULE_ERASE_CREATE, [math.pi / 4, 0, 0], 100000, 1e-18, True, "global select at scale, ties")
    # a floor taken far too high on every rank (5 % of a rank's share of k): the check against the global threshold must catch it and redo the lists
    os.environ["QB_DIST_FLOOR_FACTOR"] = "0.05"
    run_case(comm, port, big, orc.RULE_ERASE_CREATE, [math.pi / 4, 0.1, 0.2], 100000, 1e-18, True, "global select at scale, floors too high")
    run_case(comm, port, big, orc.RULE_COIN, [math.pi / 4, 0.1, 0.2], 40000, 1e-18, True, "global select at scale, floors too high, coin")
    del os.environ["QB_DIST_FLOOR_FACTOR"]
    # load balancing at the head of mpi::simulate (quids_mpi.hpp:442-500): skewed shares, same result
    for rid in orc.QCGD_RULES:
        run_case(comm, port, state, rid, p, orc.NO_TRUNCATION, 1e-18, True, f"rule {rid} skewed shares, equalize by children", share=skewed_share, equalize=2)
    run_case(comm, port, state, orc.RULE_ERASE_CREATE, p, 900, 1e-18, True, "skewed shares, equalize by objects, children truncated", share=skewed_share,
             equalize=1)
    migration_cases(comm, port)
    run_case(comm, port, grown, orc.RULE_SPLIT_MERGE, p, 30000, 1e-18, True, "split_merge on a grown state, bins on both sides, truncated", binned=2, big=True)
    failure_cases(comm, port, state)
    automatic_budget_cases(comm, port, state)
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
