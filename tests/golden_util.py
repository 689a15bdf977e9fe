"""Replays the golden scripts of tests/golden/*.npz (made by gen_golden.py from the unmodified
reference) through any engine exposing

    simulate(packed, rule_id, params, max_num_object, tolerance) -> (packed, N_c, N_u)
    apply_modifier(packed, modifier_id, params) -> packed

and checks every intermediate state: hash set and canonical bytes exact, magnitudes within 1e-12
relative, total_proba within 1e-12, N_c and N_u equal; truncating steps through the tie-band rule.
Each step starts from the GOLDEN input state, so a (legal) tie choice never compounds.
"""
import glob
import json
import os

import numpy as np

import orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixtures():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def _packed(z, prefix, total=1.0):
    return orc.Packed(z[prefix + "_sizes"], z[prefix + "_mags"], z[prefix + "_data"], total)


def replay(name, engine, hasher: orc.Oracle, rtol=1e-12):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    script = json.loads(str(z["ops"]))
    ops, hash_rule, qcgd = script["ops"], script["hash_rule"], script["qcgd"]
    state = _packed(z, "init")
    for i, op in enumerate(ops):
        meta = z[f"s{i}_meta"]
        want = _packed(z, f"s{i}", meta[2] if op["type"] == "rule" else state.total_proba)
        want_h = z[f"s{i}_hashes"]
        what = f"{name} op {i} {op}"
        if op["type"] == "mod":
            got = engine.apply_modifier(state, op["id"], op["params"])
            got.total_proba = want.total_proba
        else:
            got, nc, nu = engine.simulate(state, op["id"], op["params"], op["k"], op["tol"])
            assert (nc, nu) == (int(meta[0]), int(meta[1])), f"{what}: counters {(nc, nu)} vs {(int(meta[0]), int(meta[1]))}"
        got_h = hasher.hash_objects(got, hash_rule, [0, 0, 0])
        if op["type"] == "rule" and op["k"] != orc.NO_TRUNCATION:
            full = _packed(z, f"s{i}_full")
            orc.assert_same_truncated(got, got_h, want, want_h, full, z[f"s{i}_full_hashes"], min(op["k"], full.n), qcgd, what=what)
        else:
            orc.assert_same_state(got, got_h, want, want_h, qcgd, rtol=rtol, what=what)
        state = want
    return len(ops)
