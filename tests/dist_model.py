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


def model_simulate(port: orc.Oracle, mine: orc.Packed, all_parent_norms_fn, rule_id, params, k, tol):
    """returns (next state of this rank as Packed, N_c total, N_u total)"""
    rank, world = dist.get_rank(), dist.get_world_size()

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
    if kept.n:
        local, nc, _ = port.simulate(kept, rule_id, params, orc.NO_TRUNCATION, -1.0)
        scale = np.sqrt(local.total_proba)
        lh = port.hash_objects(local, rule_id, params)
        records = [(int(h), complex(m) * scale, o) for h, m, o in zip(lh.tolist(), local.cmags.tolist(), local.objects())]
    else:
        nc, records = 0, []

    # 2-3. partition by owner, all-to-allv (the object bytes ride along here only so that the model can
    #      hand them back; on the GPUs the representative's index travels and the origin rebuilds the bytes)
    outgoing = [[] for _ in range(world)]
    for h, m, o in records:
        outgoing[owner_of(h, world)].append((h, m, o))
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
