"""CPU model of the distributed protocol of quids_b200/csrc/dist.inc.cuh, run over torch.distributed
(gloo) with the CPU checker doing each rank's local work.  It mirrors, step for step, what
qb_simulate_dist does on the GPUs -- local merge, owner(hash) partition, all-to-allv of the locally
unique records, owner-side merge + tolerance, GLOBAL top-k with the ties at the threshold shared out
in rank order, survivors back to the rank of their representative, global normalisation -- so that
the host-side logic of the N > 1 path is exercised without a GPU (world_size 2, gloo)."""
import numpy as np
import torch.distributed as dist

import orc

MASK = (1 << 64) - 1


def mix64(x: int) -> int:  # common.cuh mix64
    x ^= x >> 32
    x = (x * 0xd6e8feb86659fd93) & MASK
    x ^= x >> 32
    x = (x * 0xd6e8feb86659fd93) & MASK
    x ^= x >> 32
    return x


def owner_of(h: int, world: int) -> int:  # dist.inc.cuh owner_of: mulhi(mix64(hash ^ golden), world)
    return (mix64(h ^ 0x9e3779b97f4a7c15) * world) >> 64


def owner_sub_buckets(world: int) -> int:  # dist.inc.cuh owner_sub_buckets: world * regions <= 2048 bins
    sub = 256
    while sub > 1 and world * sub > 2048:
        sub >>= 1
    return sub


def owner_bin(h: int, world: int, sub: int) -> int:  # dist.inc.cuh owner_bin: owner-major, region of the owner's table inside
    region = (mix64(h) * sub) >> 64 if sub > 1 else 0
    return owner_of(h, world) * sub + region


def table_home(h: int, capacity: int) -> int:  # table.cuh table_home
    return (mix64(h) * capacity) >> 64


def share_ties(need: int, eq_counts, rank: int) -> int:
    """ties at the threshold are served to the lower ranks first (capi.cu select_keep)"""
    before = sum(eq_counts[:rank])
    return min(need - before, eq_counts[rank]) if need > before else 0


def all_to_all(lists):
    """lists[r] = python list for rank r -> list of what every rank sent to me, in rank order"""
    out = [None] * dist.get_world_size()
    dist.all_to_all_object_list(out, lists) if hasattr(dist, "all_to_all_object_list") else None
    if out[0] is None:  # gloo has no all_to_all: emulate with all_gather
        gathered = [None] * dist.get_world_size()
        dist.all_gather_object(gathered, lists)
        out = [gathered[src][dist.get_rank()] for src in range(dist.get_world_size())]
    return out


class RankFailure(RuntimeError):
    pass


class Pending:
    """dist.inc.cuh pending_error / allgather_agreed: a phase that fails on one rank is remembered, the status word travels
    with the next exchange of the protocol, and EVERY rank raises there -- nobody is left waiting in a collective"""

    def __init__(self):
        self.error = None

    def run(self, fn):
        if self.error is None:
            try:
                return fn()
            except Exception as e:  # noqa: BLE001 -- the model treats any failure of a phase alike
                self.error = str(e) or type(e).__name__
        return None

    def agree(self, phase):
        world = dist.get_world_size()
        status = [None] * world
        dist.all_gather_object(status, self.error)
        for r, e in enumerate(status):
            if e is not None:
                raise RankFailure(self.error + f" [rank {dist.get_rank()}, {phase}]" if self.error is not None else f"rank {r} failed during {phase}; this rank stops too")


def model_simulate(port: orc.Oracle, mine: orc.Packed, all_parent_norms_fn, rule_id, params, k, tol, fail=None):
    """returns (next state of this rank as Packed, N_c total, N_u total); fail = (rank, phase) injects a failure of that rank
    in "local", "owner" or "return" (the test knob QB_DIST_INJECT_FAILURE of capi.cu)"""
    rank, world = dist.get_rank(), dist.get_world_size()
    pending = Pending()

    def inject(phase):
        if fail is not None and fail == (rank, phase):
            raise MemoryError(f"injected failure in phase {phase}")

    # parent pre-truncation over ALL ranks (quids.hpp:613-642 applied to the gathered state)
    norms = np.abs(mine.cmags) ** 2
    gathered = [None] * world
    dist.all_gather_object(gathered, norms)
    n_global = sum(len(g) for g in gathered)
    keep = np.arange(mine.n)
    if k < n_global:
        flat = np.sort(np.concatenate(gathered))[::-1]
        threshold = flat[k - 1]
        need = k - int((flat > threshold).sum())
        eq_counts = [int((g == threshold).sum()) for g in gathered]
        take = share_ties(need, eq_counts, rank)
        ties = np.flatnonzero(norms == threshold)[:take]
        keep = np.sort(np.concatenate([np.flatnonzero(norms > threshold), ties]))
    objs = mine.objects()
    kept = orc.Packed.from_objects([objs[i] for i in keep], mine.cmags[keep])

    # 1. local children merged locally; tolerance -1 keeps every locally unique child
    def local_phase():
        inject("local")
        if not kept.n:
            return 0, []
        local, nc, _ = port.simulate(kept, rule_id, params, orc.NO_TRUNCATION, -1.0)
        scale = np.sqrt(local.total_proba)
        lh = port.hash_objects(local, rule_id, params)
        return nc, [(int(h), complex(m) * scale, o) for h, m, o in zip(lh.tolist(), local.cmags.tolist(), local.objects())]

    nc, records = pending.run(local_phase) or (0, [])

    # 2-3. partition by owner, all-to-allv (the object bytes ride along here only so that the model can
    #      hand them back; on the GPUs the representative's index travels and the origin rebuilds the bytes)
    outgoing = [[] for _ in range(world)]
    for h, m, o in records:
        outgoing[owner_of(h, world)].append((h, m, o))
    pending.agree("the local interference step")  # rides on the count exchange of the all-to-allv
    incoming = all_to_all(outgoing)

    # 4. owner: merge, first record seen is the representative; tolerance on the global sum
    merged = {}
    for src, recs in enumerate(incoming):
        for h, m, o in recs:
            if h in merged:
                merged[h][0] += m
            else:
                merged[h] = [m, src, o]
    alive = {h: v for h, v in merged.items() if abs(v[0]) ** 2 > tol}
    pending.run(lambda: inject("owner"))
    pending.agree("the owner-side interference step")  # rides on the all-gather of the unique counts

    # 5. global top-k with ties shared out in rank order
    my_norms = np.array([v[0].real ** 2 + v[0].imag ** 2 for v in alive.values()])
    gathered = [None] * world
    dist.all_gather_object(gathered, my_norms)
    nu_global = sum(len(g) for g in gathered)
    survivors = list(alive.items())
    if k < nu_global:
        flat = np.sort(np.concatenate(gathered))[::-1]
        threshold = flat[k - 1]
        need = k - int((flat > threshold).sum())
        eq_counts = [int((g == threshold).sum()) for g in gathered]
        take = share_ties(need, eq_counts, rank)
        above = [(h, v) for (h, v), p in zip(alive.items(), my_norms) if p > threshold]
        tied = [(h, v) for (h, v), p in zip(alive.items(), my_norms) if p == threshold][:take]
        survivors = above + tied

    # 6. survivors go back to the rank of their representative
    back = [[] for _ in range(world)]
    for h, (m, src, o) in survivors:
        back[src].append((m, o))
    pending.run(lambda: inject("return"))
    pending.agree("the return of the survivors")
    mine_back = [x for recs in all_to_all(back) for x in recs]

    # 7. global normalisation
    local_total = sum(abs(m) ** 2 for m, _ in mine_back)
    totals = [None] * world
    dist.all_gather_object(totals, local_total)
    total = sum(totals)
    nxt = orc.Packed.from_objects([o for _, o in mine_back], [m / np.sqrt(total) for m, _ in mine_back])
    nxt.total_proba = total
    counts = [None] * world
    dist.all_gather_object(counts, nc)
    return nxt, sum(counts), nu_global


# ---- object migration (quids_b200/csrc/migrate.inc.cuh; quids_mpi.hpp:903-1077) ------------------------------------
def make_equal_pairs(weights):
    """utils/mpi_utils.hpp:9-24: the i-th heaviest rank is paired with the i-th lightest (ties: lower rank first)"""
    ids = sorted(range(len(weights)), key=lambda r: (-weights[r], r))
    pair = [0] * len(weights)
    for i, r in enumerate(ids):
        pair[r] = ids[len(ids) - 1 - i]
    return pair


def distribute_shares(n, world, root):
    """objects every rank holds after distribute_objects(root) (quids_mpi.hpp:1031-1051)"""
    out = [0] * world
    for node in range(1, world):
        out[node - 1 if node <= root else node] = (n * (node + 1)) // world - (n * node) // world
    out[root] = n - sum(out)
    return out


def model_send_recv(objs, n_send, peer, sender):
    """the tail of a python list travels to the peer (send_objects pops, receive_objects appends)"""
    if sender:
        tail = objs[len(objs) - n_send:]
        dist.send_object_list([tail], dst=peer)
        return objs[:len(objs) - n_send]
    box = [None]
    dist.recv_object_list(box, src=peer)
    return objs + box[0]


def model_equalize(objs, weight_of, max_rounds, min_size, inbalance_limit, min_step):
    """equalize_loop of migrate.inc.cuh over python lists: returns (objects of this rank, rounds)"""
    rank, world = dist.get_rank(), dist.get_world_size()
    rounds, previous_diff, avg0 = 0, 0.0, None
    for i in range(max_rounds):
        begins = np.concatenate([[0], np.cumsum([weight_of(o) if weight_of else 1 for o in objs])]).astype(np.int64)
        mine = int(begins[-1])
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, len(objs)))
        weights = [g[0] for g in gathered]
        if avg0 is None:
            avg0 = sum(weights) / world
        diff = max(weights) - avg0
        inbalance = diff / max(weights) if max(weights) else 0.0
        if max(g[1] for g in gathered) < min_size or inbalance < inbalance_limit or (i > 0 and diff > previous_diff * (1 - min_step)):
            break
        other = make_equal_pairs(weights)[rank]
        if other != rank and mine != weights[other]:
            if mine > weights[other]:
                if weight_of is None:  # equalize: half of the difference (quids_mpi.hpp:954-955)
                    n_send = (mine - weights[other]) // 2
                else:  # equalize_symbolic: the objects holding the last half of the difference (quids_mpi.hpp:1013-1019)
                    target = mine - (mine - weights[other]) // 2
                    limit = max(int(np.searchsorted(begins[:len(objs)], target, side="left")) - 1, 0)  # std::lower_bound - 1
                    n_send = len(objs) - limit
                objs = model_send_recv(objs, n_send, other, True)
            else:
                objs = model_send_recv(objs, 0, other, False)
        previous_diff = diff
        rounds += 1
    return objs, rounds


def model_floored_topk(keys, k, factor=1.5, sample=None):
    """family-routed path (capi.cu simulate(), pipeline.cuh filtered compaction): every rank lists only its keys at or above a
    RANK-LOCAL floor (about the factor * k / world-th largest of its own keys, here exact or taken from `sample` keys), the
    global k-th largest key is selected over the lists, and a rank whose floor turns out to lie above that threshold relists
    everything (all ranks then select again).  Returns (threshold, this rank's keys at or above it, rounds)."""
    world = dist.get_world_size()
    keys = np.asarray(keys, dtype=np.uint64)
    local_k = int(factor * k / world) + 1
    basis = np.sort(np.asarray(sample if sample is not None else keys, dtype=np.uint64))[::-1]
    rank_in_basis = min(len(basis), max(1, int(np.ceil(local_k * len(basis) / max(1, len(keys)))))) if len(basis) else 0
    floor = int(basis[rank_in_basis - 1]) if rank_in_basis and local_k < len(keys) else 0
    listed = keys[keys >= np.uint64(floor)]
    rounds = 0
    while True:
        rounds += 1
        everyone = [None] * world
        dist.all_gather_object(everyone, listed)  # (the GPUs all-reduce digit histograms instead: same threshold)
        union = np.sort(np.concatenate(everyone))[::-1]
        threshold = int(union[k - 1]) if k <= len(union) else 0
        too_high = floor > threshold
        flags = [None] * world
        dist.all_gather_object(flags, bool(too_high))
        if not any(flags):
            return threshold, np.sort(listed[listed >= np.uint64(threshold)]), rounds
        if too_high:
            floor, listed = 0, keys
