"""Generates the golden vectors of tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference, compiled by oracle/Makefile into
oracle/_ref/libquids_ref.so):

    python tests/golden/gen_golden.py [fixture names; default: all]

Each fixture is a "script": an initial packed state and a list of operations (rule iterations and
modifiers); the state after every operation, its hashes (rule->hasher), N_c, N_u and total_proba
are stored as produced by the reference (8 OpenMP threads, simple truncation).  QCGD object bytes
are stored canonical (sub_node padding masked, see tests/orc.py).  For truncating rule steps the
un-truncated result of the same step is stored too (key s{i}_full_*), because truncated parity is
defined relative to the tie band at the k-th probability.
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402

PI = math.pi


def rule(rid, params, k=orc.NO_TRUNCATION, tol=1e-30):
    return {"type": "rule", "id": rid, "params": list(params), "k": k, "tol": tol}


def mod(mid, params=()):
    return {"type": "mod", "id": mid, "params": list(params)}


def canon(p: orc.Packed, qcgd: bool) -> orc.Packed:
    if not qcgd:
        return p
    data = np.frombuffer(b"".join(orc.canonical_qcgd(o) for o in p.objects()), dtype=np.uint8) if p.n else np.zeros(0, np.uint8)
    return orc.Packed(p.sizes, p.mags, data, p.total_proba)


ONLY = set(sys.argv[1:])  # fixture names to (re)generate; none given = all


def run_script(R: orc.Oracle, name, init: orc.Packed, ops, hash_rule, qcgd):
    if ONLY and name not in ONLY:
        return
    out = {"ops": json.dumps({"ops": ops, "hash_rule": hash_rule, "qcgd": qcgd})}
    init = canon(init, qcgd)
    out["init_sizes"], out["init_mags"], out["init_data"] = init.sizes, init.mags, init.data
    state = init
    for i, op in enumerate(ops):
        meta = [0, 0, 0.0]
        if op["type"] == "mod":
            state = R.apply_modifier(state, op["id"], op["params"])
        else:
            nxt, nc, nu = R.simulate(state, op["id"], op["params"], op["k"], op["tol"])
            meta = [nc, nu, nxt.total_proba]
            if op["k"] != orc.NO_TRUNCATION:
                # the reference first keeps the k most probable PARENTS (quids.hpp:613-642); the tie-band
                # reference set is therefore the un-truncated step applied to those parents only
                src = state
                if op["k"] < state.n:
                    pr = np.abs(state.cmags) ** 2
                    order = np.argsort(-pr, kind="stable")
                    assert pr[order[op["k"] - 1]] > pr[order[op["k"]]] * (1 + 1e-9), "fixture needs untied parents"
                    keep = np.sort(order[:op["k"]])
                    objs = state.objects()
                    src = orc.Packed.from_objects([objs[j] for j in keep], state.cmags[keep])
                full, _, _ = R.simulate(src, op["id"], op["params"], orc.NO_TRUNCATION, op["tol"])
                full = canon(full, qcgd)
                out[f"s{i}_full_sizes"], out[f"s{i}_full_mags"], out[f"s{i}_full_data"] = full.sizes, full.mags, full.data
                out[f"s{i}_full_hashes"] = R.hash_objects(full, hash_rule, [0, 0, 0])
            state = nxt
        state = canon(state, qcgd)
        out[f"s{i}_sizes"], out[f"s{i}_mags"], out[f"s{i}_data"] = state.sizes, state.mags, state.data
        out[f"s{i}_hashes"] = R.hash_objects(state, hash_rule, [0, 0, 0])
        out[f"s{i}_meta"] = np.array(meta, dtype=np.float64)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(ops)} ops, final {state.n} objects, {os.path.getsize(path)} bytes")


def main():
    if not orc.have_reference():
        orc.build()
    R = orc.Oracle(orc.REF_SO)
    assert R.kind == "reference"
    H, EC, COIN, SM = orc.RULE_HADAMARD, orc.RULE_ERASE_CREATE, orc.RULE_COIN, orc.RULE_SPLIT_MERGE

    # 1. examples/quantum_computer_test.cpp:19-50 -- 4- and 5-qubit strings, every gate, then the inverse circuit
    s = 1 / math.sqrt(2)
    init = orc.Packed.from_objects([bytes([1, 1, 0, 0]), bytes([0, 1, 1, 0, 1])], [s, 1j * s])
    fwd = [rule(H, [1]), rule(H, [2]), mod(orc.MOD_CNOT, [1, 3]), mod(orc.MOD_XGATE, [2]), mod(orc.MOD_YGATE, [0]), mod(orc.MOD_ZGATE, [3])]
    bwd = [mod(orc.MOD_ZGATE, [3]), mod(orc.MOD_YGATE, [0]), mod(orc.MOD_XGATE, [2]), mod(orc.MOD_CNOT, [1, 3]), rule(H, [2]), rule(H, [1])]
    run_script(R, "qc_example", init, fwd + bwd, H, False)

    # 2. 10-qubit |0..0> register to full superposition and back (interference 1 -> 1024 -> 1)
    nq = 10
    init = orc.Packed.from_objects([bytes(nq)], [1])
    run_script(R, "qc_register10", init, [rule(H, [b]) for b in range(nq)] + [rule(H, [b]) for b in reversed(range(nq))], H, False)

    # 3. SURVEY appendix A.4 known answers: the 3-node graph left=110 right=101
    g3 = bytearray(R.qcgd_random_state(3, 1, 0, 1.0).objects()[0])
    g3[2:5] = bytes([1, 1, 0])
    g3[5:8] = bytes([1, 0, 1])
    g3 = orc.Packed.from_objects([bytes(g3)], [1])
    run_script(R, "qcgd_kat_erase_create", g3, [rule(EC, [0.3333, 0, 0])], EC, True)
    run_script(R, "qcgd_kat_coin", g3, [rule(COIN, [0.25, 0.25, 0])], EC, True)
    run_script(R, "qcgd_kat_split_merge", g3, [rule(SM, [0.25, 0.25, 0.25])], EC, True)

    # 4. examples/qcgd_test.cpp:26-49 -- forward then reversed sequence on one random 6-node graph
    #    (their variable `coin` is an erase_create(0.25, 0.25)); final state = 1 object, P = 1
    for seed in (1, 7):
        init = R.qcgd_random_state(6, 1, seed, 1.0)
        ec, co, sm, rsm = [0.3333, 0, 0], [0.25, 0.25, 0], [0.25, 0.25, 0.25], [0.25, 0.25, -0.25]
        tol = 1e-15
        ops = [mod(orc.MOD_STEP), rule(EC, co, tol=tol), rule(EC, co, tol=tol), rule(EC, ec, tol=tol), rule(SM, sm, tol=tol),
               mod(orc.MOD_STEP), rule(SM, sm, tol=tol),
               rule(SM, rsm, tol=tol), mod(orc.MOD_REVERSED_STEP), rule(SM, rsm, tol=tol), rule(EC, ec, tol=tol), mod(orc.MOD_REVERSED_STEP)]
        run_script(R, f"qcgd_example_seed{seed}", init, ops, EC, True)

    # 5. the production sequence step; split_merge; step; erase_create (+ coin) from a single graph, two rounds
    init = R.qcgd_random_state(5, 1, 3, 1.0)
    th = [PI / 4, PI / 4, PI / 4]
    ops = []
    for r in range(2):
        ops += [mod(orc.MOD_STEP), rule(SM, th, tol=1e-18), mod(orc.MOD_STEP), rule(EC, [PI / 4, 0, 0], tol=1e-18)]
        if r == 0:
            ops += [rule(COIN, [0.3, 0.2, 0.1], tol=1e-18)]
    run_script(R, "qcgd_grown", init, ops, EC, True)

    # 6. truncation, one step at a time from an identical input (SURVEY section 4 consequence 2):
    #    40 random 6-node graphs with distinct magnitudes, parents pre-truncated too (k < N_p on the 2nd op)
    base = R.qcgd_random_state(6, 40, 11, 1.0)
    rng = np.random.default_rng(5)
    mags = rng.normal(size=(40, 2))
    mags /= np.sqrt((mags ** 2).sum())
    base = orc.Packed(base.sizes, mags, base.data)
    run_script(R, "qcgd_truncate_children", base, [rule(EC, [PI / 4, 0.1, 0.2], k=200, tol=1e-18)], EC, True)
    run_script(R, "qcgd_truncate_parents", base, [rule(SM, th, k=25, tol=1e-18)], EC, True)

    # 7. random-density graphs through every QCGD rule (no truncation), wider graphs
    base = R.qcgd_random_state(9, 12, 21)
    for nm, rid, pr in (("erase_create", EC, [0.4, 0.3, 0.2]), ("coin", COIN, [0.4, 0.3, 0.2]), ("split_merge", SM, [0.4, 0.3, 0.2])):
        run_script(R, f"qcgd_random9_{nm}", base, [rule(rid, pr, tol=1e-18), mod(orc.MOD_STEP), rule(rid, pr, tol=1e-18)], EC, True)

    # 8. split_merge outside the GPU path's fast paths: graphs of 33-700 nodes with hand-placed split, merge and wrap-around
    #    sites (objects up to 14 KB) among ordinary 12-node graphs
    from quids_b200 import qcgd

    def graph(n, splits, merges, wrap):
        g = bytearray(qcgd.fresh_graph(n).tobytes())
        for i in splits:
            g[2 + i] = g[2 + n + i] = 1
        for i in merges:
            g[2 + i], g[2 + n + i + 1] = 1, 1
        if wrap:
            g[2 + n + 0], g[2 + n - 1] = 1, 1
        return bytes(g)

    rng = np.random.default_rng(21)
    objs = [graph(40, [3, 20], [10, 30], True), graph(40, [0, 17], [5], False), graph(33, [32], [1, 8], False),
            graph(700, [3, 400], [10, 650], True), graph(600, [0], [100], False)]
    for _ in range(20):
        g = bytearray(qcgd.fresh_graph(12).tobytes())
        g[2:2 + 24] = bytes(rng.integers(0, 2, size=24, dtype=np.uint8))
        objs.insert(int(rng.integers(0, len(objs) + 1)), bytes(g))
    mags = rng.normal(size=len(objs)) + 1j * rng.normal(size=len(objs))
    init = orc.Packed.from_objects(objs, mags / np.linalg.norm(mags))
    run_script(R, "qcgd_wide_split_merge", init, [rule(SM, [0.3, 0.2, 0.1], tol=1e-18)], EC, True)


if __name__ == "__main__":
    main()
