"""world_size-2 gloo test of the distributed protocol (CPU, no GPU): tests/dist_model.py mirrors the
host-side logic of qb_simulate_dist; its union over the ranks must equal the single-process checker
on the gathered state, truncation ties included."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dist_model  # noqa: E402
import orc  # noqa: E402


def test_owner_function_is_balanced_and_total():
    rng = np.random.default_rng(0)
    hashes = rng.integers(0, 2**63, size=20000, dtype=np.int64).astype(np.uint64)
    for world in (2, 3, 8):
        owners = np.array([dist_model.owner_of(int(h), world) for h in hashes])
        assert owners.min() == 0 and owners.max() == world - 1
        counts = np.bincount(owners, minlength=world)
        assert counts.max() < 1.1 * len(hashes) / world


def test_exchange_bins_group_the_records_by_owner_then_by_table_region():
    """the records a rank sends are ordered by bin = owner * regions + region (dist.inc.cuh): every owner's records stay
    contiguous, and inside an owner's segment the home slot in the owner's table never decreases from one region to
    the next, whatever the capacity -- which is what lets the owner's inserts sweep its table"""
    rng = np.random.default_rng(3)
    hashes = [int(h) for h in rng.integers(0, 2**63, size=5000, dtype=np.int64).astype(np.uint64)] + [0, 1, 2**64 - 1]
    assert [dist_model.owner_sub_buckets(w) for w in (1, 2, 8, 9, 16, 64, 4096)] == [256, 256, 256, 128, 128, 32, 1]
    for world in (2, 8, 12):
        sub = dist_model.owner_sub_buckets(world)
        assert world * sub <= 2048
        order = sorted(hashes, key=lambda h: dist_model.owner_bin(h, world, sub))
        owners = [dist_model.owner_of(h, world) for h in order]
        assert owners == sorted(owners), "an owner's records must be contiguous"
        for capacity in (1024, 999983, 31000000):
            for o in range(world):
                segment = [h for h in order if dist_model.owner_of(h, world) == o]
                regions = [dist_model.owner_bin(h, world, sub) for h in segment]
                homes = [dist_model.table_home(h, capacity) for h in segment]
                for a in range(1, len(segment)):
                    if regions[a] != regions[a - 1]:  # crossing into the next region: every home there is at or above the last one's
                        assert min(homes[a:]) >= max(x for x, r in zip(homes[:a], regions[:a]) if r == regions[a - 1])


def test_tie_sharing_serves_lower_ranks_first():
    assert [dist_model.share_ties(5, [2, 2, 4], r) for r in range(3)] == [2, 2, 1]
    assert [dist_model.share_ties(0, [2, 2, 4], r) for r in range(3)] == [0, 0, 0]
    assert sum(dist_model.share_ties(7, [3, 0, 9, 1], r) for r in range(4)) == 7


def _worker(rank, world, port_number, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_number))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    port = orc.Oracle(orc.PORT_SO)
    rng = np.random.default_rng(7)
    base = port.qcgd_random_state(7, 120, 4)
    mags = rng.normal(size=(120, 2))
    distinct = orc.Packed(base.sizes, mags / np.sqrt((mags ** 2).sum()), base.data)
    tied = port.qcgd_random_state(6, 90, 9)
    failures = []
    cases = [(distinct, orc.RULE_ERASE_CREATE, orc.NO_TRUNCATION), (distinct, orc.RULE_SPLIT_MERGE, orc.NO_TRUNCATION), (distinct, orc.RULE_COIN, 300),
             (distinct, orc.RULE_ERASE_CREATE, 50), (tied, orc.RULE_ERASE_CREATE, 211)]
    for state, rid, k in cases:
        params, tol = [0.37, 0.21, -0.4], 1e-18
        objs = state.objects()
        idx = [i for i in range(state.n) if i % world == rank]
        mine = orc.Packed.from_objects([objs[i] for i in idx], state.cmags[idx])
        nxt, nc, nu = dist_model.model_simulate(port, mine, None, rid, params, k, tol)
        gathered = [None] * world
        dist.all_gather_object(gathered, (nxt.sizes, nxt.mags, nxt.data, nxt.total_proba))
        if rank == 0:
            got = orc.Packed(np.concatenate([g[0] for g in gathered]), np.concatenate([g[1] for g in gathered]), np.concatenate([g[2] for g in gathered]),
                             gathered[0][3])
            want, wc, wu = port.simulate(state, rid, params, k, tol)
            try:
                assert (nc, nu) == (wc, wu), ((nc, nu), (wc, wu))
                hg, hw = port.hash_objects(got, rid), port.hash_objects(want, rid)
                if k >= wu:
                    orc.assert_same_state(got, hg, want, hw, True, what=f"rule {rid} k {k}")
                else:
                    src = state
                    if k < state.n:
                        order = np.sort(np.argsort(-(np.abs(state.cmags) ** 2), kind="stable")[:k])
                        src = orc.Packed.from_objects([objs[j] for j in order], state.cmags[order])
                    full, _, _ = port.simulate(src, rid, params, orc.NO_TRUNCATION, tol)
                    if state is tied:  # everything is tied: only the count and membership are defined
                        assert got.n == k and set(hg.tolist()) <= set(port.hash_objects(full, rid).tolist())
                    else:
                        orc.assert_same_truncated(got, hg, want, hw, full, port.hash_objects(full, rid), k, True, what=f"rule {rid} k {k}")
            except AssertionError as e:  # report through the queue: an exception here would hang the other rank
                failures.append(f"rule {rid} k {k}: {e}")
    if rank == 0:
        results.put(failures)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_protocol_matches_single_process(port):
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, results)) for r in range(2)]
    for p in procs:
        p.start()
    failures = results.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert not failures, failures
    assert all(p.exitcode == 0 for p in procs)


def _failure_worker(rank, world, port_number, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_number))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    port = orc.Oracle(orc.PORT_SO)
    state = port.qcgd_random_state(6, 60, 2)
    objs = state.objects()
    idx = [i for i in range(state.n) if i % world == rank]
    mine = orc.Packed.from_objects([objs[i] for i in idx], state.cmags[idx])
    seen = []
    for phase in ("local", "owner", "return"):
        try:
            dist_model.model_simulate(port, mine, None, orc.RULE_ERASE_CREATE, [0.3, 0.2, 0.1], 100, 1e-18, fail=(1, phase))
            seen.append((phase, None))
        except dist_model.RankFailure as e:
            seen.append((phase, str(e)))
    # the protocol is still in step on every rank: a clean iteration goes through
    nxt, nc, nu = dist_model.model_simulate(port, mine, None, orc.RULE_ERASE_CREATE, [0.3, 0.2, 0.1], 100, 1e-18)
    gathered = [None] * world
    dist.all_gather_object(gathered, (seen, nxt.n))
    if rank == 0:
        results.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_failure_on_one_rank_is_agreed_on_by_all():
    """an overflow / allocation failure that only rank 1 sees: every rank raises from the same call, none hangs
    (dist.inc.cuh allgather_agreed; the GPU counterpart is tests/dist_worker.py::failure_cases)"""
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    procs = [ctx.Process(target=_failure_worker, args=(r, 2, 29535, results)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = results.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, (seen, n_after) in enumerate(gathered):
        assert [ph for ph, _ in seen] == ["local", "owner", "return"]
        for phase, message in seen:
            assert message is not None, f"rank {rank} went on although rank 1 failed in {phase}"
            assert ("injected failure" in message) == (rank == 1), message
    assert sum(n for _, n in gathered) == 100


def test_pairing_and_shares_of_the_migration_logic():
    assert dist_model.make_equal_pairs([5, 9, 1, 7]) == [3, 2, 1, 0]
    assert dist_model.make_equal_pairs([4, 4, 4]) == [2, 1, 0]  # the middle rank is alone
    for n, world, root in ((10, 4, 0), (10, 4, 2), (7, 3, 2), (2, 8, 5), (1000003, 8, 1)):
        shares = dist_model.distribute_shares(n, world, root)
        assert sum(shares) == n and max(shares) - min(shares) <= 1, shares


def _equalize_worker(rank, world, port_number, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_number))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    everything = [(i, int(w)) for i, w in enumerate(rng.integers(1, 200, size=400))]
    objs = everything if rank == 0 else []
    # by children (weights), then by objects (weight 1 each) from the same skewed start
    by_children, r1 = dist_model.model_equalize(list(objs), lambda o: o[1], 4, 10, 0.05, 0.0)
    by_objects, r2 = dist_model.model_equalize(list(objs), None, 4, 10, 0.05, 0.0)
    gathered = [None] * world
    dist.all_gather_object(gathered, (by_children, by_objects, r1, r2))
    if rank == 0:
        results.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_equalize_model():
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    procs = [ctx.Process(target=_equalize_worker, args=(r, 2, 29534, results)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = results.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    children = [sum(w for _, w in g[0]) for g in gathered]
    objects = [len(g[1]) for g in gathered]
    assert sorted(o for g in gathered for o in g[0]) == sorted(o for g in gathered for o in g[1])  # nothing lost, nothing duplicated
    assert len({o for g in gathered for o in g[0]}) == 400
    assert gathered[0][2] >= 1 and abs(children[0] - children[1]) <= 200, children  # within one object's weight
    assert gathered[0][3] >= 1 and abs(objects[0] - objects[1]) <= 1, objects


def _floor_worker(rank, world, port_number, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_number))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    out = []
    k = 400
    balanced = [rng.integers(1, 2 ** 40, size=5000, dtype=np.uint64) for _ in range(world)]
    skewed = [np.concatenate([balanced[0] + np.uint64(2 ** 50), np.array([], dtype=np.uint64)]), balanced[1]]  # every survivor sits on rank 0
    tied = [np.full(3000, 7, dtype=np.uint64) for _ in range(world)]
    for name, shares, sample in (("balanced", balanced, None), ("skewed", skewed, None), ("tied", tied, None),
                                 ("sampled floor", balanced, lambda keys: rng.choice(keys, size=300))):
        mine = shares[rank]
        threshold, kept, rounds = dist_model.model_floored_topk(mine, k, sample=sample(mine) if sample else None)
        gathered = [None] * world
        dist.all_gather_object(gathered, (kept, rounds))
        if rank == 0:
            everything = np.sort(np.concatenate(shares))[::-1]
            out.append((name, threshold, int(everything[k - 1]), sum(len(g[0]) for g in gathered), int((everything >= everything[k - 1]).sum()), [g[1] for g in gathered]))
    if rank == 0:
        results.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_rank_local_floors_give_the_global_threshold():
    """the filtered lists of the family-routed path: whatever the floors, the threshold is the global k-th largest key and every key
    at or above it is listed; a rank that holds more than its share of the survivors relists (second round)"""
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    procs = [ctx.Process(target=_floor_worker, args=(r, 2, 29536, results)) for r in range(2)]
    for p in procs:
        p.start()
    out = results.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for name, threshold, want_threshold, kept, want_kept, rounds in out:
        assert threshold == want_threshold, (name, threshold, want_threshold)
        assert kept == want_kept, (name, kept, want_kept)
        assert len(set(rounds)) == 1, (name, rounds)  # the ranks stay in step
    by_name = {o[0]: o for o in out}
    assert by_name["balanced"][5][0] == 1 and by_name["skewed"][5][0] == 2, out
