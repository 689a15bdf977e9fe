"""Vectorised comparison of two LARGE states (1e6 .. 1e8 objects) as hash-keyed sets -- the parity rule of
tests/orc.py (assert_same_state / assert_same_truncated) without a Python dictionary per object.

TEST INFRASTRUCTURE: used by tests/ and by bench.py's parity gates (bench.py may execute oracle/ only as the
checker, see oracle/oracle_api.h).

Rule (BASELINE.json north_star):
  * hashes (computed by the CHECKER's hasher over both states' bytes) are unique inside a state;
  * not truncated: the two hash sets are equal; truncated to k: both keep exactly k objects, and an object
    kept by one side only lies inside the tie band around the k-th probability
    (p <= threshold_of_the_other_side * (1 + band)): the reference itself picks arbitrarily among ties;
  * common objects: equal size; canonical bytes equal (all of them when `bytes_sample` is None, else a
    random sample -- the hash set already pins the bytes through an independent hasher); magnitudes, with
    the normalisation undone (mag * sqrt(total_proba)), within rtol relative to the larger modulus -- or, for an
    object whose contributions nearly cancel, to the state's RMS modulus (`rms_floor`): a sum carries the
    rounding of its TERMS, so an object left far below the typical modulus by interference cannot agree to
    1e-12 of its own modulus between two summation orders (the reference against itself with another thread
    count does not either).  The strict per-object figure is reported as max_rel_magnitude_error;
  * total_proba within tp_rtol relative.  north_star asks for 1e-12; a sum over >= 1e6 rounded terms in another order
    already moves the reference AGAINST ITSELF by more (port vs reference, both on the CPU, 2.5e6 unique children:
    1.15e-12; the reference with 1 vs 8 threads: 3e-15 per magnitude), hence 1e-11 for large states;
    the small-state tests keep 1e-12.
compare() returns a dict of what was measured and raises AssertionError on the first violated rule.
"""
import numpy as np

import orc


def _sorted(h):
    order = np.argsort(h, kind="stable")
    s = h[order]
    if s.shape[0] > 1:
        assert (s[1:] != s[:-1]).all(), "duplicate hash inside one state"
    return order, s


def _object_bytes(p: orc.Packed, begin, i, qcgd):
    o = bytes(p.data[int(begin[i]):int(begin[i]) + int(p.sizes[i])])
    return orc.canonical_qcgd(o) if qcgd else o


def compare(a: orc.Packed, ha, b: orc.Packed, hb, qcgd, truncated_k=None, rtol=1e-12, band=1e-12, tp_rtol=1e-11, bytes_sample=100000, seed=0, what="", rms_floor=1.0):
    ha, hb = np.asarray(ha, np.uint64), np.asarray(hb, np.uint64)
    assert ha.shape[0] == a.n and hb.shape[0] == b.n
    oa, sa = _sorted(ha)
    ob, sb = _sorted(hb)
    fa, fb = np.sqrt(a.total_proba), np.sqrt(b.total_proba)
    out = {"objects": [a.n, b.n]}
    if truncated_k is None:
        assert a.n == b.n, f"{what}: {a.n} vs {b.n} objects"
        assert np.array_equal(sa, sb), f"{what}: hash sets differ ({np.setxor1d(sa, sb, assume_unique=True).shape[0]} hashes on one side only)"
        ia, ib = oa, ob
        out["only_one_side"] = 0
    else:
        assert a.n == truncated_k and b.n == truncated_k, f"{what}: kept {a.n} / {b.n}, expected {truncated_k}"
        in_b = np.isin(sa, sb, assume_unique=True)
        in_a = np.isin(sb, sa, assume_unique=True)
        ia, ib = oa[in_b], ob[in_a]  # the common objects of both sides, in hash order
        # unnormalised probabilities; the threshold of a side = the smallest probability it kept
        pa = (a.mags ** 2).sum(axis=1) * a.total_proba
        pb = (b.mags ** 2).sum(axis=1) * b.total_proba
        ta, tb = pa.min(), pb.min()
        only_a, only_b = oa[~in_b], ob[~in_a]
        assert only_a.shape[0] == only_b.shape[0]
        if only_a.shape[0]:
            assert pa[only_a].max() <= tb * (1 + band) * (1 + 1e-15), f"{what}: an object above the threshold band is missing on the other side ({pa[only_a].max()} vs threshold {tb})"
            assert pb[only_b].max() <= ta * (1 + band) * (1 + 1e-15), f"{what}: an object above the threshold band is missing on one side ({pb[only_b].max()} vs threshold {ta})"
        assert abs(ta - tb) <= band * max(ta, tb) * 4, f"{what}: thresholds {ta} vs {tb}"
        out["only_one_side"] = int(only_a.shape[0])
        out["threshold"] = float(ta)
    # common objects, aligned by hash
    assert np.array_equal(a.sizes[ia], b.sizes[ib]), f"{what}: object sizes differ"
    ma, mb = a.mags[ia] * fa, b.mags[ib] * fb
    diff = np.sqrt(((ma - mb) ** 2).sum(axis=1))
    scale = np.maximum(np.sqrt((ma ** 2).sum(axis=1)), np.sqrt((mb ** 2).sum(axis=1)))
    rel = float((diff / np.maximum(scale, 1e-300)).max()) if diff.shape[0] else 0.0
    rms = np.sqrt(max(a.total_proba, b.total_proba) / max(1, max(a.n, b.n))) * rms_floor
    worst = float((diff / np.maximum(scale, max(rms, 1e-300))).max()) if diff.shape[0] else 0.0
    assert (diff <= rtol * np.maximum(scale, rms)).all(), f"{what}: magnitudes differ by up to {rel} relative ({worst} relative to max(modulus, RMS modulus))"
    out["max_rel_magnitude_error"] = rel
    out["max_magnitude_error_vs_rms"] = worst
    tp_err = abs(a.total_proba - b.total_proba) / max(abs(a.total_proba), abs(b.total_proba), 1e-300)
    assert tp_err <= tp_rtol, f"{what}: total_proba {a.total_proba} vs {b.total_proba}"
    out["total_proba"] = [a.total_proba, b.total_proba]
    # canonical bytes
    n = ia.shape[0]
    ba, bb = a.begin, b.begin
    same_layout = n and (a.sizes == a.sizes[0]).all() and (b.sizes == a.sizes[0]).all()
    if same_layout:
        # fixed-size objects: matrix comparison, a million objects at a time
        s = int(a.sizes[0])
        A, B = a.data.reshape(a.n, s), b.data.reshape(b.n, s)
        keep = np.ones(s, bool)
        if qcgd:  # mask the 4 indeterminate padding bytes of every sub_node
            nodes = int(np.frombuffer(A[0, :2].tobytes(), "<u2")[0])
            assert (A[:, :2] == A[0, :2]).all() and (B[:, :2] == A[0, :2]).all(), f"{what}: node counts differ inside a fixed-size state"
            col = np.arange(s) - (4 + 4 * nodes)
            keep = ~((col >= 0) & ((col & 15) >= 4) & ((col & 15) < 8))
        for c0 in range(0, n, 1 << 20):
            sl = slice(c0, min(n, c0 + (1 << 20)))
            assert np.array_equal(A[ia[sl]][:, keep], B[ib[sl]][:, keep]), f"{what}: object bytes differ"
        out["bytes_compared"] = int(n)
    else:
        pick = np.arange(n) if bytes_sample is None or n <= bytes_sample else np.random.default_rng(seed).choice(n, bytes_sample, replace=False)
        for j in pick.tolist():
            assert _object_bytes(a, ba, int(ia[j]), qcgd) == _object_bytes(b, bb, int(ib[j]), qcgd), f"{what}: bytes differ for hash {int(ha[ia[j]]):016x}"
        out["bytes_compared"] = int(pick.shape[0])
    out["common"] = int(n)
    return out
