"""Parity of the CUDA path (through the C ABI) against the CPU checker and the golden vectors.

Bar (BASELINE.json north_star): surviving set of object byte strings and hashes bit-exact
(sub_node padding masked), magnitudes within 1e-12 relative, total_proba within 1e-12; truncating
steps through the documented tie band."""
import math

import numpy as np
import pytest

import golden_util
import orc

pytestmark = pytest.mark.gpu

PI = math.pi


@pytest.fixture(scope="module")
def gpu():
    import quids_b200 as qb
    if qb.lib().qb_device_count() < 1:
        pytest.fail("no CUDA device: the CUDA path has no fallback")
    from gpu_engine import GpuEngine
    return GpuEngine


@pytest.mark.parametrize("suffix", ["", "_generic"])
@pytest.mark.parametrize("name", golden_util.fixtures())
def test_golden_vectors(gpu, port, name, suffix):
    assert golden_util.replay(name, gpu(suffix), port) > 0


@pytest.mark.parametrize("name", [n for n in golden_util.fixtures() if n.startswith("qcgd")])
def test_golden_vectors_with_parent_ordering(gpu, port, name):
    """the locality ordering of the parents (an engine knob) must not change any result"""
    import quids_b200 as qb
    qb.config.locality_sort = 2
    try:
        assert golden_util.replay(name, gpu(""), port) > 0
    finally:
        qb.config.locality_sort = 1


@pytest.mark.parametrize("name", golden_util.fixtures())
def test_golden_vectors_with_binned_inserts(gpu, port, name):
    """children deduplicated bin by bin in shared memory (table.cuh; automatic from 2^22 children on, forced here): the same
    results, for the rules written with the four reference methods and for the fused ones"""
    import quids_b200 as qb
    qb.config.binned_inserts = 2
    try:
        for suffix in ("", "_generic"):
            assert golden_util.replay(name, gpu(suffix), port) > 0
    finally:
        qb.config.binned_inserts = 1


def test_binned_interference_spill_and_fallback(gpu, port):
    """equal hashes pile up in one bin: 60000 identical parents send each of their children 60000 times to the same bin (far
    beyond the bin's fixed space: the spill list, sorted by bin, carries the rest); and a state with more unique children per
    bin than the shared-memory table holds cannot happen with mixed hashes, so the fallback is forced through a tiny limit"""
    import quids_b200 as qb
    one = port.qcgd_random_state(8, 1, 5)
    same = orc.Packed.from_objects(one.objects() * 60000, [1 / math.sqrt(60000)] * 60000)
    mixed = port.qcgd_random_state(9, 3000, 8)
    both = orc.Packed(np.concatenate([same.sizes, mixed.sizes]), np.concatenate([same.mags, mixed.mags]), np.concatenate([same.data, mixed.data]))
    qb.config.binned_inserts = 2
    try:
        for rid, params in ((orc.RULE_SPLIT_MERGE, [0.4, 0.3, 0.2]), (orc.RULE_ERASE_CREATE, [0.7, 0.2, 0.1])):
            eng = gpu("_generic")
            want, nc, nu = port.simulate(both, rid, params, orc.NO_TRUNCATION, 1e-18)
            got, gc, gu = eng.simulate(both, rid, params, tol=1e-18)
            assert (gc, gu) == (nc, nu)
            orc.assert_same_state(got, port.hash_objects(got, rid), want, port.hash_objects(want, rid), True, rtol=1e-11, what=f"binned with spill, rule {rid}")
    finally:
        qb.config.binned_inserts = 1


@pytest.mark.parametrize("align", [0, 4, 8, 16])
def test_golden_vectors_other_alignments(gpu, port, align):
    for name in ("qc_example", "qcgd_example_seed1", "qcgd_truncate_children"):
        if align == 0 and name.startswith("qcgd"):
            pass  # QCGD object sizes are multiples of 4, so unpadded storage stays 4-byte aligned
        golden_util.replay(name, gpu("", align=align), port)


def test_hashers_match(gpu, port):
    rng = np.random.default_rng(0)
    objs = [bytes(rng.integers(0, 256, size=l, dtype=np.uint8)) for l in range(0, 80)]
    st = orc.Packed.from_objects(objs, [1] * len(objs))
    for align in (0, 8):
        assert np.array_equal(gpu("", align=align).hash_objects(st, orc.RULE_HADAMARD, [0]), port.hash_objects(st, orc.RULE_HADAMARD, [0]))
    g = port.qcgd_random_state(9, 300, 4)
    grown, _, _ = port.simulate(g, orc.RULE_SPLIT_MERGE, [0.3, 0.2, 0.1], tolerance=1e-18)
    for rid in orc.QCGD_RULES:
        assert np.array_equal(gpu().hash_objects(grown, rid), port.hash_objects(grown, rid, [0, 0, 0]))


@pytest.mark.parametrize("suffix", ["", "_generic"])
@pytest.mark.parametrize("rule_id", orc.QCGD_RULES)
def test_qcgd_random_graphs_vs_oracle(gpu, port, rule_id, suffix):
    params = [0.37, 0.21, -0.4]
    eng = gpu(suffix)
    state = port.qcgd_random_state(8, 400, 9)
    for it in range(2):
        want, nc, nu = port.simulate(state, rule_id, params, tolerance=1e-18)
        got, gc, gu = eng.simulate(state, rule_id, params, tol=1e-18)
        assert (gc, gu) == (nc, nu)
        orc.assert_same_state(got, port.hash_objects(got, rule_id, params), want, port.hash_objects(want, rule_id, params), True,
                              what=f"rule {rule_id}{suffix} iteration {it}")
        state = port.apply_modifier(want, orc.MOD_STEP)
        assert eng.apply_modifier(want, orc.MOD_STEP).objects() == state.objects()
        if state.n > 60000:
            break


@pytest.mark.parametrize("rule_id", [orc.RULE_ERASE_CREATE, orc.RULE_COIN])
def test_qcgd_parent_ordering_and_groups_vs_oracle(gpu, port, rule_id):
    """12-node graphs (parents of up to 4096 children = several warp groups), parents ordered by
    locality key, with truncation of parents and children"""
    import quids_b200 as qb
    params = [0.37, 0.21, -0.4]
    state = port.qcgd_random_state(12, 300, 21)
    rng = np.random.default_rng(2)
    mags = rng.normal(size=(300, 2))
    state = orc.Packed(state.sizes, mags / np.sqrt((mags ** 2).sum()), state.data)
    qb.config.locality_sort = 2
    try:
        want, nc, nu = port.simulate(state, rule_id, params, tolerance=1e-18)
        got, gc, gu = gpu().simulate(state, rule_id, params, tol=1e-18)
        assert (gc, gu) == (nc, nu)
        orc.assert_same_state(got, port.hash_objects(got, rule_id), want, port.hash_objects(want, rule_id), True, what="ordered parents")
        k = 250  # fewer than the parents: pre-truncation, then ordering of the kept parents
        want, wc, wu = port.simulate(state, rule_id, params, k, 1e-18)
        got, gc, gu = gpu().simulate(state, rule_id, params, k, 1e-18)
        assert (gc, gu) == (wc, wu)
        order = np.sort(np.argsort(-(np.abs(state.cmags) ** 2), kind="stable")[:k])
        objs = state.objects()
        sub = orc.Packed.from_objects([objs[j] for j in order], state.cmags[order])
        full, _, _ = port.simulate(sub, rule_id, params, tolerance=1e-18)
        orc.assert_same_truncated(got, port.hash_objects(got, rule_id), want, port.hash_objects(want, rule_id), full, port.hash_objects(full, rule_id), k, True)
    finally:
        qb.config.locality_sort = 1


_region_cases = {}  # (states, the oracle's results per rule): computed once, used by both kernels


@pytest.mark.parametrize("batch", ["0", "1"])
@pytest.mark.parametrize("rule_id", [orc.RULE_ERASE_CREATE, orc.RULE_COIN])
def test_region_kernels_vs_oracle(gpu, port, rule_id, batch, monkeypatch):
    """sorted order with table regions, both kernels forced in turn (QB_ITEMS_BATCH): the accumulating one (long runs) and the
    batch one (32 items at a time, short runs) -- on states with one group per region, with many parents per family (several
    parent patterns per run: product chains and, beyond 6 items, the butterflies), with identical parents (stretches of 32),
    with graphs wider than the fold table (17-30 nodes), with and without truncation"""
    import quids_b200 as qb
    monkeypatch.setenv("QB_ITEMS_BATCH", batch)
    params = [0.37, 0.21, -0.4]
    rng = np.random.default_rng(5)

    def with_random_magnitudes(state):
        mags = rng.normal(size=(state.n, 2))
        return orc.Packed(state.sizes, mags / np.sqrt((mags ** 2).sum()), state.data)

    if "states" not in _region_cases:
        one = port.qcgd_random_state(9, 1, 5)
        grown, _, _ = port.simulate(port.qcgd_random_state(9, 400, 8), orc.RULE_SPLIT_MERGE, [0.4, 0.3, 0.2], orc.NO_TRUNCATION, 1e-18)
        _region_cases["states"] = {
            "12 nodes, distinct families": with_random_magnitudes(port.qcgd_random_state(12, 300, 21)),
            "6 nodes, every family many times": with_random_magnitudes(port.qcgd_random_state(6, 1500, 3)),
            "identical parents": orc.Packed.from_objects(one.objects() * 500, [1 / math.sqrt(500)] * 500),
            "grown state (ragged sizes and node counts)": with_random_magnitudes(grown),
            "wider than the fold table": port.qcgd_random_state(21, 6, 4),  # (its own uniform magnitudes keep total_proba's summation error small)
        }
    states = _region_cases["states"]
    qb.config.locality_sort = 2
    try:
        for what, state in states.items():
            if (rule_id, what) not in _region_cases:
                _region_cases[(rule_id, what)] = port.simulate(state, rule_id, params, tolerance=1e-18)
            want, nc, nu = _region_cases[(rule_id, what)]
            got, gc, gu = gpu().simulate(state, rule_id, params, tol=1e-18)
            assert (gc, gu) == (nc, nu), what
            # random complex magnitudes of several parents add up with partial cancellation: the rounding error is 1e-16 of the
            # largest term, i.e. up to a few 1e-12 of the sum (the summation order differs from the reference's)
            orc.assert_same_state(got, port.hash_objects(got, rule_id), want, port.hash_objects(want, rule_id), True, rtol=1e-11,
                                  what=f"{what}, batch={batch}")
        state = states["12 nodes, distinct families"]
        k = 5000
        full, _, _ = port.simulate(state, rule_id, params, tolerance=1e-18)
        assert full.n > k
        want, wc, wu = port.simulate(state, rule_id, params, k, 1e-18)
        got, gc, gu = gpu().simulate(state, rule_id, params, k, 1e-18)
        assert (gc, gu) == (wc, wu)
        orc.assert_same_truncated(got, port.hash_objects(got, rule_id), want, port.hash_objects(want, rule_id), full, port.hash_objects(full, rule_id), k, True)
    finally:
        qb.config.locality_sort = 1


def test_qcgd_wide_graphs_vs_oracle(gpu, port):
    # more than 64 nodes: the bit-mask fast path of erase_create / coin does not apply
    base = port.qcgd_random_state(70, 2, 1, 1.0)
    objs = []
    for o in base.objects():
        a = bytearray(o)
        n = 70
        for i in range(n):  # keep the fan-out small: only 6 eligible nodes per rule
            a[2 + i] = 1 if i < 6 else (i & 1)
            a[2 + n + i] = 1 if i < 6 else 1 - (i & 1)
        a[2 + 3] ^= 1
        objs.append(bytes(a))
    objs[1] = objs[1][:2 + 8] + bytes([1 - objs[1][2 + 8]]) + objs[1][2 + 9:]
    st = orc.Packed.from_objects(objs, [0.6, 0.8j])
    for rid in (orc.RULE_ERASE_CREATE, orc.RULE_COIN):
        want, nc, nu = port.simulate(st, rid, [0.3, 0.1, 0.2], tolerance=1e-18)
        got, gc, gu = gpu().simulate(st, rid, [0.3, 0.1, 0.2], tol=1e-18)
        assert (gc, gu) == (nc, nu)
        orc.assert_same_state(got, port.hash_objects(got, rid), want, port.hash_objects(want, rid), True, what=f"wide rule {rid}")


def test_hadamard_ragged_vs_oracle(gpu, port):
    rng = np.random.default_rng(3)
    objs = [bytes(rng.integers(0, 2, size=int(l), dtype=np.uint8)) for l in rng.integers(3, 40, size=300)]
    objs = list(dict.fromkeys(objs))
    st = orc.Packed.from_objects(objs, rng.normal(size=len(objs)) + 1j * rng.normal(size=len(objs)))
    for align in (0, 8):
        eng = gpu("", align=align)
        cur = st
        for bit in (0, 2, 1, 2, 0):
            want, nc, nu = port.simulate(cur, orc.RULE_HADAMARD, [bit])
            got, gc, gu = eng.simulate(cur, orc.RULE_HADAMARD, [bit])
            assert (gc, gu) == (nc, nu)
            orc.assert_same_state(got, port.hash_objects(got, 1, [0]), want, port.hash_objects(want, 1, [0]), False, what=f"H({bit}) align {align}")
            cur = want


def test_modifiers_vs_oracle(gpu, port):
    rng = np.random.default_rng(4)
    objs = [bytes(rng.integers(0, 2, size=6, dtype=np.uint8)) for _ in range(1000)]
    st = orc.Packed.from_objects(objs, rng.normal(size=1000) + 1j * rng.normal(size=1000))
    eng = gpu("", align=0)
    for mid, params in ((orc.MOD_CNOT, [1, 3]), (orc.MOD_XGATE, [2]), (orc.MOD_YGATE, [0]), (orc.MOD_ZGATE, [3]), (orc.MOD_PHASE, [0.3])):
        want, got = port.apply_modifier(st, mid, params), eng.apply_modifier(st, mid, params)
        assert got.objects() == want.objects()
        assert np.array_equal(got.mags, want.mags), mid  # modifiers are exact: no reassociation anywhere
        st = want
    g = port.qcgd_random_state(7, 500, 2)
    for mid in (orc.MOD_STEP, orc.MOD_REVERSED_STEP, orc.MOD_STEP):
        want, got = port.apply_modifier(g, mid), gpu().apply_modifier(g, mid)
        assert got.objects() == want.objects()
        g = want


@pytest.mark.parametrize("filtered_list", [False, True])
def test_truncation_band_vs_oracle(gpu, port, filtered_list, monkeypatch):
    if filtered_list:  # the compaction lists only the entries above a sampled lower bound of the k-th key (default: tables of 2^24 slots and more)
        monkeypatch.setenv("QB_COMPACT_FILTER_MIN", "0")
    base = port.qcgd_random_state(7, 300, 13, 1.0)
    rng = np.random.default_rng(6)
    mags = rng.normal(size=(300, 2))
    mags /= np.sqrt((mags ** 2).sum())
    st = orc.Packed(base.sizes, mags, base.data)
    params = [PI / 4, 0.1, 0.2]
    for rid, k in ((orc.RULE_ERASE_CREATE, 1000), (orc.RULE_COIN, 777), (orc.RULE_SPLIT_MERGE, 400)):  # k >= number of parents: only the children are truncated here
        full, _, nu = port.simulate(st, rid, params, tolerance=1e-18)
        want, wc, wu = port.simulate(st, rid, params, k, 1e-18)
        got, gc, gu = gpu().simulate(st, rid, params, k, 1e-18)
        assert (gc, gu) == (wc, wu) and nu == wu
        orc.assert_same_truncated(got, port.hash_objects(got, rid), want, port.hash_objects(want, rid), full, port.hash_objects(full, rid),
                                  min(k, full.n), True, what=f"truncated rule {rid}")


@pytest.mark.parametrize("filtered_list", [False, True])
def test_truncation_with_exact_ties(gpu, port, filtered_list, monkeypatch):
    if filtered_list:  # the compaction lists only the entries above a sampled lower bound of the k-th key (default: tables of 2^24 slots and more)
        monkeypatch.setenv("QB_COMPACT_FILTER_MIN", "0")
    # equal magnitudes everywhere: any k objects of the tied set are a legal answer, the count is not negotiable
    st = port.qcgd_random_state(6, 64, 3)
    k = 500
    full, _, nu = port.simulate(st, orc.RULE_ERASE_CREATE, [PI / 4, 0, 0], tolerance=1e-18)
    got, gc, gu = gpu().simulate(st, orc.RULE_ERASE_CREATE, [PI / 4, 0, 0], k, 1e-18)
    assert gu == nu and got.n == min(k, nu)
    kf = orc.keyed(full, port.hash_objects(full, orc.RULE_ERASE_CREATE), True)
    f, ff = math.sqrt(got.total_proba), math.sqrt(full.total_proba)
    probs = np.sort(np.abs(full.cmags) ** 2)[::-1]
    for h, (o, m) in orc.keyed(got, port.hash_objects(got, orc.RULE_ERASE_CREATE), True).items():
        assert h in kf and kf[h][0] == o
        assert abs(m * f - kf[h][1] * ff) <= 1e-12 * abs(m * f)
        assert abs(kf[h][1]) ** 2 >= probs[k - 1] * (1 - 1e-12)


@pytest.mark.parametrize("filtered_list", [False, True])
def test_parent_pretruncation(gpu, port, filtered_list, monkeypatch):
    if filtered_list:  # the compaction lists only the entries above a sampled lower bound of the k-th key (default: tables of 2^24 slots and more)
        monkeypatch.setenv("QB_COMPACT_FILTER_MIN", "0")
    st0 = port.qcgd_random_state(6, 200, 17, 1.0)
    rng = np.random.default_rng(8)
    mags = rng.normal(size=(200, 2))
    mags /= np.sqrt((mags ** 2).sum())
    st = orc.Packed(st0.sizes, mags, st0.data)
    k = 50  # < number of parents: the 50 most probable parents are expanded, then 50 children kept
    want, wc, wu = port.simulate(st, orc.RULE_COIN, [0.3, 0.2, 0.1], k, 1e-18)
    got, gc, gu = gpu().simulate(st, orc.RULE_COIN, [0.3, 0.2, 0.1], k, 1e-18)
    assert (gc, gu) == (wc, wu)
    order = np.argsort(-(np.abs(st.cmags) ** 2), kind="stable")[:k]
    objs = st.objects()
    sub = orc.Packed.from_objects([objs[j] for j in np.sort(order)], st.cmags[np.sort(order)])
    full, _, _ = port.simulate(sub, orc.RULE_COIN, [0.3, 0.2, 0.1], tolerance=1e-18)
    orc.assert_same_truncated(got, port.hash_objects(got, 3), want, port.hash_objects(want, 3), full, port.hash_objects(full, 3), k, True)


def test_empty_and_degenerate_states(gpu, port):
    eng = gpu()
    empty = orc.Packed.from_objects([], [])
    got, nc, nu = eng.simulate(empty, orc.RULE_HADAMARD, [0])
    assert (got.n, nc, nu, got.total_proba) == (0, 0, 0, 0.0)
    assert eng.apply_modifier(empty, orc.MOD_XGATE, [0]).n == 0
    # everything cancels: H on (|0> - |1>)/sqrt2 ... then a tolerance that removes all children
    st = orc.Packed.from_objects([bytes([0, 0])], [1])
    got, nc, nu = eng.simulate(st, orc.RULE_HADAMARD, [0], tol=10.0)
    assert (got.n, nc, nu, got.total_proba) == (0, 2, 0, 0.0)
    # a single object
    got, nc, nu = eng.simulate(st, orc.RULE_HADAMARD, [1])
    assert (got.n, nc, nu) == (2, 2, 2) and abs(got.total_proba - 1) < 1e-15


def test_phase_labels_in_reference_order(gpu):
    import quids_b200 as qb
    it, nxt, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    qb.config.align_byte_length = 0
    it.append(bytes([0, 1, 0]), 1.0)
    labels = []
    qb.simulate(it, qb.Rule("hadamard", 1), nxt, sym, qb.NO_TRUNCATION, labels.append)
    assert labels == ["num_child", "truncate_symbolic - prepare", "truncate_symbolic", "prepare_index", "symbolic_iteration",
                      "compute_collisions - prepare", "compute_collisions - insert", "compute_collisions - finalize", "truncate - prepare",
                      "truncate", "prepare_final", "final", "normalize", "end"]
    assert nxt.num_object == 2
    qb.config.align_byte_length = 8


def test_unsupported_requests_fail_loudly(gpu):
    import quids_b200 as qb
    it, nxt, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    it.append(bytes([0, 1, 0]), 1.0)
    qb.simulate(it, qb.Rule("hadamard", 1), nxt, sym, 0)  # automatic budget: everything is kept when it fits ...
    assert nxt.num_object == 2
    qb.config.safety_margin = 1.0  # ... and the call fails when not even one parent's children do
    try:
        with pytest.raises(qb.QuidsError, match="automatic budget"):
            qb.simulate(it, qb.Rule("hadamard", 1), nxt, sym, 0)
    finally:
        qb.config.safety_margin = 0.2
    with pytest.raises(qb.QuidsError):
        qb.simulate(it, qb.Rule("hadamard", 1), it, sym)
    with pytest.raises(qb.QuidsError):
        qb.Rule("not_a_rule")


# ---- size-independent properties at sizes the CPU checker would not finish in seconds ---------------------
def test_register_to_full_superposition_and_back(gpu):
    """SURVEY 8(d) C3 shape at 20 qubits: 1 -> 2^20 objects -> 1, exact cancellation, P = 1."""
    import quids_b200 as qb
    nq = 20
    qb.config.align_byte_length = 0
    qb.config.tolerance = 1e-30
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.append(bytes(nq), 1.0)
    for bit in range(nq):
        qb.simulate(a, qb.Rule("hadamard", bit), b, sym)
        assert sym.num_object == 2 ** (bit + 1) and b.num_object == 2 ** (bit + 1)
        a, b = b, a
    sizes, mags, data = a.download_packed()
    assert len(set(bytes(r) for r in data.reshape(-1, nq))) == 2 ** nq
    s = 1 / math.sqrt(2.)
    amp = 1.0
    for _ in range(nq):
        amp *= s
    np.testing.assert_allclose(mags[:, 0], amp, rtol=1e-12)
    assert np.all(mags[:, 1] == 0)
    for bit in reversed(range(nq)):
        qb.simulate(a, qb.Rule("hadamard", bit), b, sym)
        assert sym.num_object == 2 ** (bit + 2) and sym.num_object_after_interferences == 2 ** bit
        a, b = b, a
    obj, mag = a.get_object(0)
    assert a.num_object == 1 and obj == bytes(nq)
    assert abs(mag - 1) < 1e-12 and abs(a.total_proba - 1) < 1e-12
    qb.config.align_byte_length = 8


def test_qcgd_reversibility_large(gpu, port):
    """examples/qcgd_test.cpp property on a wider graph: forward then reversed rules return the
    single initial graph with magnitude 1 (interference of ~1e5-1e6 children down to one object)."""
    import quids_b200 as qb
    init = port.qcgd_random_state(9, 1, 5, 1.0)
    qb.config.align_byte_length = 8
    qb.config.tolerance = 1e-15
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.upload_packed(init.sizes, init.mags, init.data)
    ec, sm, rsm = qb.Rule("erase_create", 0.3333), qb.Rule("split_merge", 0.25, 0.25, 0.25), qb.Rule("split_merge", 0.25, 0.25, -0.25)
    step, rstep = qb.Modifier("step"), qb.Modifier("reversed_step")
    qb.simulate(a, step)
    qb.simulate(a, ec, b, sym)
    qb.simulate(b, sm, a, sym)
    qb.simulate(a, step)
    qb.simulate(a, sm, b, sym)
    grown = b.num_object
    assert grown > 1000
    qb.simulate(b, rsm, a, sym)
    qb.simulate(a, rstep)
    qb.simulate(a, rsm, b, sym)
    qb.simulate(b, ec, a, sym)
    qb.simulate(a, rstep)
    assert a.num_object == 1
    obj, mag = a.get_object(0)
    assert orc.canonical_qcgd(obj) == orc.canonical_qcgd(init.objects()[0])
    assert abs(abs(mag) - 1) < 1e-9 and abs(a.total_proba - 1) < 1e-9
    qb.config.tolerance = 1e-30


def test_probability_is_conserved_without_truncation(gpu, port):
    import quids_b200 as qb
    init = port.qcgd_random_state(10, 2000, 5)
    init = orc.Packed.from_objects(list(dict.fromkeys(init.objects())), [1] * len(dict.fromkeys(init.objects())))
    init.mags /= math.sqrt(init.n)
    qb.config.tolerance = 1e-18
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.upload_packed(init.sizes, init.mags, init.data)
    for rule in (qb.Rule("erase_create", PI / 4), qb.Rule("split_merge", PI / 4, PI / 4, PI / 4)):
        qb.simulate(a, rule, b, sym)
        assert abs(b.total_proba - 1) < 1e-11
        h = b.hashes(rule)
        assert len(np.unique(h)) == b.num_object == sym.num_object_after_interferences
        a, b = b, a
    qb.config.tolerance = 1e-30


def test_sorted_and_unsorted_order_agree_at_scale(gpu):
    """2e6 random 12-node parents (2.6e8 children): the sorted order with on-chip family accumulation
    and the plain order (every child goes to the table) must produce the same objects, the same
    hashes and magnitudes within 1e-12; erase_create is unitary, so without truncation the
    probability of the (duplicate-free) input is conserved."""
    import quids_b200 as qb
    from quids_b200 import qcgd
    n = 2_000_000
    sizes, data = qcgd.random_graphs(12, n, seed=3)
    # drop duplicate parents: the state must be a proper superposition for unitarity to show
    rows = np.unique(data.reshape(n, -1), axis=0)
    n = rows.shape[0]
    rng = np.random.default_rng(0)
    mags = rng.normal(size=(n, 2))
    mags /= np.sqrt((mags ** 2).sum())
    qb.config.tolerance = 1e-30
    qb.config.align_byte_length = 8
    rule = qb.Rule("erase_create", 0.7, 0.3, -0.2)
    results = []
    for mode in (0, 2):
        qb.config.locality_sort = mode
        a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
        a.upload_packed(np.full(n, rows.shape[1], np.uint32), mags, rows.reshape(-1))
        qb.simulate(a, rule, b, sym)
        h = b.hashes(rule)
        _, m, _ = b.download_packed()
        order = np.argsort(h)
        results.append((h[order], m[order], b.total_proba, sym.num_object, sym.num_object_after_interferences))
    qb.config.locality_sort = 1
    (h0, m0, p0, c0, u0), (h1, m1, p1, c1, u1) = results
    assert (c0, u0) == (c1, u1) and c0 > 2e8
    assert len(np.unique(h0)) == len(h0) and np.array_equal(h0, h1)
    scale = np.maximum(np.abs(m0).max(axis=1), 1e-300)
    assert (np.abs(m0 - m1).max(axis=1) <= 1e-12 * np.maximum(scale, np.abs(m0).max() * 1e-6)).all()
    assert abs(p0 - 1) < 1e-10 and abs(p1 - 1) < 1e-10
    qb.config.tolerance = 1e-30


def test_probabilistic_truncation(gpu, port):
    """the reference's default truncation (quids.hpp:594-608, 829-845): keep the k smallest u / |mag|^2.
    Not reproducible in the reference (seeded from rand()), so the checks are structural and statistical:
    exactly k objects, all of them genuine children with the right magnitudes, reproducible for a seed,
    different for another seed, and objects are kept more often the more probable they are."""
    import quids_b200 as qb
    base = port.qcgd_random_state(7, 200, 31, 1.0)
    rng = np.random.default_rng(9)
    mags = rng.normal(size=(200, 2)) * np.exp(rng.normal(size=(200, 1)))
    st = orc.Packed(base.sizes, mags / np.sqrt((mags ** 2).sum()), base.data)
    rid, params, k = orc.RULE_ERASE_CREATE, [0.6, 0.1, 0.2], 400
    full, nc, nu = port.simulate(st, rid, params, tolerance=1e-18)
    kf = orc.keyed(full, port.hash_objects(full, rid), True)
    fscale = math.sqrt(full.total_proba)
    assert nu > 3 * k
    eng = gpu()
    qb.config.simple_truncation = False
    try:
        kept_count = {h: 0 for h in kf}
        first = None
        seeds = range(24)
        for seed in seeds:
            qb.config.seed = seed
            got, gc, gu = eng.simulate(st, rid, params, k, 1e-18)
            assert (gc, gu, got.n) == (nc, nu, k)
            kg = orc.keyed(got, port.hash_objects(got, rid), True)
            scale = math.sqrt(got.total_proba)
            for h, (o, m) in kg.items():
                assert h in kf and kf[h][0] == o
                assert abs(m * scale - kf[h][1] * fscale) <= 1e-12 * abs(m * scale)
                kept_count[h] += 1
            if seed == 0:
                first = set(kg)
                again, _, _ = eng.simulate(st, rid, params, k, 1e-18)
                assert set(orc.keyed(again, port.hash_objects(again, rid), True)) == first  # same seed, same choice
            elif seed == 1:
                assert set(kg) != first
        probs = np.array([abs(kf[h][1]) ** 2 for h in kf])
        freq = np.array([kept_count[h] / len(seeds) for h in kf])
        order = np.argsort(probs)
        low, high = freq[order[:len(order) // 4]].mean(), freq[order[-len(order) // 4:]].mean()
        assert high > low + 0.3, (low, high)
    finally:
        qb.config.simple_truncation = True
        qb.config.seed = 0


def test_table_sized_from_misleading_history_is_redone(gpu, port):
    """the interference table is sized from the previous call of the same rule; when the state changes nature (many
    duplicates -> all distinct) the prediction is far too small: the kernels must stop early, the step is redone at the
    safe size, and the result is the oracle's (both orders of processing the children)."""
    import time
    import quids_b200 as qb
    one = port.qcgd_random_state(9, 1, 5)
    same = orc.Packed.from_objects(one.objects() * 70000, [1 / math.sqrt(70000)] * 70000)  # >= 2^16 groups: sorted order
    grown, _, _ = port.simulate(port.qcgd_random_state(9, 3000, 8), orc.RULE_SPLIT_MERGE, [0.4, 0.3, 0.2], orc.NO_TRUNCATION, 1e-18)
    assert grown.n > 20000
    rid, params = orc.RULE_ERASE_CREATE, [0.7, 0.2, 0.1]
    for sort in (1, 2, 0):
        qb.config.locality_sort = sort
        try:
            eng = gpu()
            got, gc, gu = eng.simulate(same, rid, params, tol=1e-18)  # N_u / N_c tiny: the history says "small table"
            assert gu <= 512 < gc // 1000
            want, nc, nu = port.simulate(grown, rid, params, orc.NO_TRUNCATION, 1e-18)
            t0 = time.perf_counter()
            got, gc, gu = eng.simulate(grown, rid, params, tol=1e-18)
            assert time.perf_counter() - t0 < 20, "a table that is too small must cost milliseconds, not a probe sequence per child"
            assert (gc, gu) == (nc, nu)
            orc.assert_same_state(got, port.hash_objects(got, rid), want, port.hash_objects(want, rid), True, what=f"redo after overflow, sort={sort}")
        finally:
            qb.config.locality_sort = 1


def test_automatic_budget_keeps_what_fits(gpu, port):
    """max_num_object = 0 (quids.hpp:459-485, 510-536): the most probable parents whose symbolic workspace fits the
    budget, then the most probable children the next state has room for.  The budget is given (qb_options.memory_budget)
    so that a small state exercises both truncations; the result must be the oracle's for the same two counts."""
    import quids_b200 as qb
    base = port.qcgd_random_state(8, 500, 21)
    rng = np.random.default_rng(4)
    mags = rng.normal(size=(500, 2)) * np.exp(rng.normal(size=(500, 1)))
    st = orc.Packed(base.sizes, mags / np.sqrt((mags ** 2).sum()), base.data)
    rid, params, tol = orc.RULE_ERASE_CREATE, [0.6, 0.1, 0.2], 1e-18
    objs = st.objects()
    order = np.argsort(-(np.abs(st.cmags) ** 2), kind="stable")
    children = np.array([port.simulate(orc.Packed.from_objects([objs[i]], [1.0]), rid, params, orc.NO_TRUNCATION, tol)[1] for i in order])
    cum = np.cumsum(children)
    eng = gpu()
    full, nc, nu = port.simulate(st, rid, params, orc.NO_TRUNCATION, tol)
    got, gc, gu = eng.simulate(st, rid, params, 0, tol)  # measured budget: a B200 holds all of it
    assert (gc, gu) == (nc, nu)
    orc.assert_same_state(got, port.hash_objects(got, rid), full, port.hash_objects(full, rid), True, what="automatic budget, everything fits")
    seen, refused = set(), 0
    try:
        for budget in [int(6e6 * 0.8 ** i) for i in range(22)]:  # 6 MB (everything fits) down to 55 KB
            qb.config.memory_budget = budget
            try:
                got, gc, gu = eng.simulate(st, rid, params, 0, tol)
            except qb.QuidsError as e:  # too small even for one parent, or no room left for a next state
                assert "automatic budget" in str(e)
                refused += 1
                continue
            assert 0 < got.n
            k_p = int(np.searchsorted(cum, gc)) + 1  # parents kept: the children counted are those of the k_p most probable
            assert k_p <= st.n and cum[k_p - 1] == gc, (k_p, gc)
            kept = np.sort(order[:k_p])
            src = orc.Packed.from_objects([objs[j] for j in kept], st.cmags[kept])
            fullk, nck, nuk = port.simulate(src, rid, params, orc.NO_TRUNCATION, tol)
            assert (gc, gu) == (nck, nuk)
            hg = port.hash_objects(got, rid)
            if got.n == nuk:
                orc.assert_same_state(got, hg, fullk, port.hash_objects(fullk, rid), True, what=f"automatic budget {budget}")
            else:
                # children-only truncation of the oracle's untruncated result (an explicit max_num_object would cut the parents too)
                raw = fullk.cmags * math.sqrt(fullk.total_proba)
                top = np.sort(np.argsort(-(np.abs(raw) ** 2), kind="stable")[:got.n])
                total = float((np.abs(raw[top]) ** 2).sum())
                fobjs = fullk.objects()
                want = orc.Packed.from_objects([fobjs[j] for j in top], raw[top] / math.sqrt(total))
                want.total_proba = total
                orc.assert_same_truncated(got, hg, want, port.hash_objects(want, rid), fullk, port.hash_objects(fullk, rid), got.n, True,
                                          what=f"automatic budget {budget}")
            seen.add((k_p < st.n, got.n < nuk))
    finally:
        qb.config.memory_budget = 0
    assert (False, False) in seen and (False, True) in seen, seen  # everything fits; only the children are truncated
    assert any(s[0] for s in seen), seen                           # parents are truncated


def test_async_transfers_double_buffered(gpu, port):
    """qb_iter_upload_async / qb_iter_download_async / qb_iter_wait: the pipelined use of bench.py's e2e leg (upload of
    step i+1 and download of step i-1 around the rule iteration of step i) gives the results of the synchronous calls"""
    import quids_b200 as qb
    qb.config.tolerance, qb.config.align_byte_length = 1e-18, 8
    rule = qb.Rule("erase_create", 0.4, 0.1, 0.2)
    states = [port.qcgd_random_state(7, 300 + 50 * i, 40 + i) for i in range(5)]
    sym = qb.SymbolicIteration()
    want = []
    for st in states:
        it, nxt = qb.Iteration(), qb.Iteration()
        it.upload_packed(st.sizes, st.mags, st.data)
        qb.simulate(it, rule, nxt, sym)
        want.append(nxt.download())
    hosts = []
    for st in states:  # the reference's storage layout on the host, as a driver would hold it
        it = qb.Iteration()
        it.upload_packed(st.sizes, st.mags, st.data)
        o, b, s, m = it.download()
        hosts.append((o.copy(), b.copy(), s.copy(), m.reshape(-1).copy()))
    ins, outs = [qb.Iteration(), qb.Iteration()], [qb.Iteration(), qb.Iteration()]
    results = [None] * len(states)
    bufs = [None, None]
    ins[0].upload_async(*hosts[0])
    for i in range(len(states)):
        cur = i % 2
        if i + 1 < len(states):
            ins[1 - cur].upload_async(*hosts[i + 1])
        if bufs[cur] is not None:  # the result of step i-2 must have left before its state is overwritten on the host side
            outs[cur].wait()
            results[i - 2] = [x.copy() for x in bufs[cur]]
        qb.simulate(ins[cur], rule, outs[cur], sym)
        n, nb = outs[cur].num_object, outs[cur].num_bytes
        bufs[cur] = [np.zeros(nb, np.uint8), np.zeros(n + 1, np.uint64), np.zeros(n, np.uint32), np.zeros(2 * n, np.float64)]
        outs[cur].download_async(*bufs[cur])
    for i in (len(states) - 2, len(states) - 1):
        outs[i % 2].wait()
        results[i] = bufs[i % 2]
    for got, (o, b, s, m) in zip(results, want):
        assert np.array_equal(got[1], b) and np.array_equal(got[2], s)
        # same kernels, same inputs: the objects arrive in the same order only if the table fills the same way; compare as sets
        key = lambda oo, bb, ss, mm: sorted((oo[int(bb[j]):int(bb[j]) + int(ss[j])].tobytes(), round(float(mm[2 * j]), 12), round(float(mm[2 * j + 1]), 12))
                                            for j in range(len(ss)))
        assert key(got[0], got[1], got[2], got[3]) == key(o, b, s, m.reshape(-1))


def test_float_magnitudes_at_the_boundary(gpu, port):
    """PROBA_TYPE = float (quids.hpp:21-23): complex<float> magnitudes cross the C ABI (qb_iter_upload_f32 /
    qb_iter_download_f32); the result agrees with the double oracle within 1e-5 relative (north_star's float tolerance)"""
    import quids_b200 as qb
    qb.config.tolerance, qb.config.align_byte_length = 1e-18, 8
    st = port.qcgd_random_state(8, 400, 13)
    rng = np.random.default_rng(2)
    mags = rng.normal(size=(400, 2))
    mags = (mags / np.sqrt((mags ** 2).sum())).astype(np.float32)
    st = orc.Packed(st.sizes, mags.astype(np.float64), st.data)  # the oracle sees exactly the float values, widened
    rid, params = orc.RULE_ERASE_CREATE, [0.5, 0.2, 0.1]
    want, nc, nu = port.simulate(st, rid, params, orc.NO_TRUNCATION, 1e-18)
    it, nxt, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    begin = np.concatenate([[0], np.cumsum((st.sizes.astype(np.uint64) + 7) // 8 * 8)]).astype(np.uint64)
    objects = np.zeros(int(begin[-1]), np.uint8)
    for j, o in enumerate(st.objects()):
        objects[int(begin[j]):int(begin[j]) + len(o)] = np.frombuffer(o, np.uint8)
    it.upload(objects, begin, st.sizes, mags)  # float32 magnitudes select the f32 entry point
    o64, b64, s64, m64 = it.download()
    assert np.array_equal(m64, mags.astype(np.float64))  # widened exactly
    qb.simulate(it, qb.Rule("erase_create", *params), nxt, sym)
    assert (sym.num_object, sym.num_object_after_interferences) == (nc, nu)
    o, b, s, m32 = nxt.download(np.float32)
    _, _, _, m = nxt.download()
    assert m32.dtype == np.float32 and np.array_equal(m32, m.astype(np.float32))  # narrowed with round-to-nearest
    sizes, mags_out, data = nxt.download_packed()
    got = orc.Packed(sizes, m32.astype(np.float64), data, nxt.total_proba)
    orc.assert_same_state(got, port.hash_objects(got, rid), want, port.hash_objects(want, rid), True, rtol=1e-5, what="float boundary")


def test_device_observables_vs_oracle(gpu, port):
    """iteration::average_value (quids.hpp:208-234) as a device reduction (qb_iter_average_value) against the checker:
    the four averages of utils::serialize in one pass ("qcgd_stats"), a qubit probability, sizes; 1e-12 relative"""
    import quids_b200 as qb
    rng = np.random.default_rng(12)
    g = port.qcgd_random_state(9, 5000, 6)
    g.mags[:] = rng.normal(size=g.mags.shape) / 70
    grown, _, _ = port.simulate(g, orc.RULE_SPLIT_MERGE, [0.3, 0.2, 0.1], tolerance=1e-18)
    it = gpu()._load(grown)
    stats = it.average_value("qcgd_stats")
    for got, oid in zip(stats, (orc.OBS_QCGD_SIZE, orc.OBS_QCGD_SQUARED_SIZE, orc.OBS_QCGD_DENSITY, orc.OBS_QCGD_SQUARED_DENSITY)):
        want = port.average_value(grown, oid)
        assert abs(got - want) <= 1e-12 * abs(want), (oid, got, want)
    assert abs(it.average_value("qcgd_size") - port.average_value(grown, orc.OBS_QCGD_SIZE)) <= 1e-12 * stats[0]
    assert abs(it.average_value("object_bytes") - port.average_value(grown, orc.OBS_BYTES)) <= 1e-12 * port.average_value(grown, orc.OBS_BYTES)
    reg = orc.Packed.from_objects([bytes(rng.integers(0, 2, size=l, dtype=np.uint8)) for l in rng.integers(3, 9, size=3000)], rng.normal(size=3000) + 1j * rng.normal(size=3000))
    it = gpu("", align=0)._load(reg)
    for bit in (0, 2, 7):
        want = port.average_value(reg, orc.OBS_QUBIT, [bit])
        assert abs(it.average_value("qubit", bit) - want) <= 1e-12 * want
    # an empty state averages to 0; an unknown observable fails loudly
    assert qb.Iteration().average_value("qcgd_size") == 0.0
    with pytest.raises(qb.QuidsError):
        it.average_value("no_such_observable")


def test_split_merge_wide_and_huge_graphs_vs_oracle(gpu, port):
    """split_merge outside its fast paths: graphs of more than 32 nodes (the children walk the object instead of the
    per-parent context) and objects larger than the finalisation's shared-memory stages (built in place), mixed with
    ordinary graphs in the same warp batches; two iterations, so that the second one meets grown names"""
    from quids_b200 import qcgd
    rng = np.random.default_rng(21)

    def graph(n, splits, merges, wrap):
        g = bytearray(qcgd.fresh_graph(n).tobytes())
        for i in splits:  # both particles: a split site
            g[2 + i] = g[2 + n + i] = 1
        for i in merges:  # left at i, right (and no left) at i + 1: a merge site
            g[2 + i], g[2 + n + i + 1] = 1, 1
        if wrap:  # right at node 0, left without right at the last node: the wrap-around merge
            g[2 + n + 0], g[2 + n - 1] = 1, 1
        return bytes(g)

    objs = [graph(40, [3, 20], [10, 30], True), graph(40, [0, 17], [5], False), graph(33, [32], [1, 8], False),
            graph(700, [3, 400], [10, 650], True), graph(600, [0], [100], False)]
    for _ in range(40):  # ordinary 12-node graphs around them
        n = 12
        g = bytearray(qcgd.fresh_graph(n).tobytes())
        g[2:2 + 2 * n] = bytes(rng.integers(0, 2, size=2 * n, dtype=np.uint8))
        objs.insert(int(rng.integers(0, len(objs) + 1)), bytes(g))
    mags = rng.normal(size=len(objs)) + 1j * rng.normal(size=len(objs))
    st = orc.Packed.from_objects(objs, mags / np.linalg.norm(mags))
    params = [0.3, 0.2, 0.1]
    for suffix in ("", "_generic"):
        state = st
        for iteration in range(2):
            want, nc, nu = port.simulate(state, orc.RULE_SPLIT_MERGE, params, tolerance=1e-18)
            got, gc, gu = gpu(suffix).simulate(state, orc.RULE_SPLIT_MERGE, params, tol=1e-18)
            assert (gc, gu) == (nc, nu), (suffix, iteration)
            orc.assert_same_state(got, port.hash_objects(got, orc.RULE_SPLIT_MERGE), want, port.hash_objects(want, orc.RULE_SPLIT_MERGE), True,
                                  what=f"split_merge{suffix} wide/huge iteration {iteration}")
            # second iteration: the grown graphs with fresh random magnitudes (iterating on the coherent result would un-split
            # everything and leave residues of cancelled sums, for which a relative tolerance means nothing)
            state = orc.Packed(want.sizes, np.random.default_rng(5).normal(size=want.mags.shape) / 10, want.data)


@pytest.mark.parametrize("two_pass", [False, True])
def test_truncation_at_scale_vs_oracle(gpu, port, two_pass):
    """7e5 unique children, k = 1e5: the radix select on its large-input path (candidates of the first two digits copied
    out, one-pass compaction; with QB_SELECT_TWO_PASS the compaction of inputs beyond 2^31 keys), against the checker;
    then the same with equal magnitudes everywhere, where the candidates do not fit and every pass reads all the keys"""
    import os
    params = [PI / 4, 0.1, 0.2]
    base = port.qcgd_random_state(12, 6000, 13, 1.0)
    rng = np.random.default_rng(6)
    mags = rng.normal(size=(base.n, 2))
    mags /= np.sqrt((mags ** 2).sum())
    st = orc.Packed(base.sizes, mags, base.data)
    k = 100000
    if two_pass:
        os.environ["QB_SELECT_TWO_PASS"] = "1"
    try:
        full, _, nu = port.simulate(st, orc.RULE_ERASE_CREATE, params, tolerance=1e-18)
        assert nu >= (1 << 18)
        want, wc, wu = port.simulate(st, orc.RULE_ERASE_CREATE, params, k, 1e-18)
        got, gc, gu = gpu().simulate(st, orc.RULE_ERASE_CREATE, params, k, 1e-18)
        assert (gc, gu) == (wc, wu)
        orc.assert_same_truncated(got, port.hash_objects(got, orc.RULE_ERASE_CREATE), want, port.hash_objects(want, orc.RULE_ERASE_CREATE), full,
                                  port.hash_objects(full, orc.RULE_ERASE_CREATE), k, True, what="truncation at scale")
        # exact ties: every parent has the same magnitude, the children a few hundred distinct ones
        tied = port.qcgd_random_state(12, 6000, 3)
        full, _, nu = port.simulate(tied, orc.RULE_ERASE_CREATE, [PI / 4, 0, 0], tolerance=1e-18)
        got, gc, gu = gpu().simulate(tied, orc.RULE_ERASE_CREATE, [PI / 4, 0, 0], k, 1e-18)
        assert gu == nu and got.n == k
        kf = orc.keyed(full, port.hash_objects(full, orc.RULE_ERASE_CREATE), True)
        f, ff = math.sqrt(got.total_proba), math.sqrt(full.total_proba)
        probs = np.sort(np.abs(full.cmags) ** 2)[::-1]
        for h, (o, m) in orc.keyed(got, port.hash_objects(got, orc.RULE_ERASE_CREATE), True).items():
            assert h in kf and kf[h][0] == o
            assert abs(m * f - kf[h][1] * ff) <= 1e-12 * abs(m * f)
            assert abs(kf[h][1]) ** 2 >= probs[k - 1] * (1 - 1e-12)
    finally:
        os.environ.pop("QB_SELECT_TWO_PASS", None)
