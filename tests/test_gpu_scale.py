"""Parity AT THE SCALE the numbers are quoted on ("parity gaps first"): the bench generator with its duplicate
parents kept, 1e6 parents, default work-item ordering (sorted order + table regions for erase_create / coin, the binned
interference path for split_merge), GPU vs the reference compiled here (oracle/_ref) when it travelled to the box, else the
port.  Hash sets and bytes exact, magnitudes 1e-12, counters equal; truncating steps through the tie band (tests/bigcmp.py).
Plus configs[2] (24-byte qubit registers, exact cancellation) at 2^22 objects and the configs[1] modifiers at 1e7 objects.
"""
import math

import numpy as np
import pytest

import bigcmp
import orc

pytestmark = pytest.mark.gpu
PI = math.pi


@pytest.fixture(scope="module")
def checker(port):
    """the unmodified reference where it was built/shipped (every host thread), else the port"""
    if orc.have_reference():
        o = orc.Oracle(orc.REF_SO)
        o.set_num_threads(0)
        return o
    return port


@pytest.fixture(scope="module")
def qb():
    import quids_b200
    if quids_b200.lib().qb_device_count() < 1:
        pytest.fail("no CUDA device: the CUDA path has no fallback")
    quids_b200.config.locality_sort = 1
    quids_b200.config.align_byte_length = 8
    quids_b200.config.simple_truncation = True
    return quids_b200


def bench_state(n, seed=0):
    """what bench.py feeds the loop: random density-1/2 12-node graphs, duplicates kept, magnitude 1/sqrt(n) in float arithmetic"""
    from quids_b200 import qcgd
    sizes, data = qcgd.random_graphs(12, n, seed=seed)
    mags = np.zeros((n, 2))
    mags[:, 0] = qcgd.read_state_magnitude(n)[0]
    return orc.Packed(sizes, mags, data)


def gpu_simulate(qb, state, name, params, k, tol=1e-18):
    qb.config.tolerance = tol
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.upload_packed(state.sizes, state.mags, state.data, state.total_proba)
    qb.simulate(a, qb.Rule(name, *params), b, sym, k)
    sizes, mags, data = b.download_packed()
    return orc.Packed(sizes, mags, data, b.total_proba), sym.num_object, sym.num_object_after_interferences, sym.phase_ms


RULES = [("erase_create", orc.RULE_ERASE_CREATE, [PI / 4, 0.0, 0.0]), ("coin", orc.RULE_COIN, [PI / 4, 0.3, -0.2]), ("split_merge", orc.RULE_SPLIT_MERGE, [PI / 4, PI / 4, PI / 4])]


@pytest.mark.parametrize("name,rid,params", RULES)
def test_bench_generator_1e6_parents_truncated(qb, checker, name, rid, params):
    """1e6 parents with their duplicates (about 3 % of 1e6 draws from 2^24 graphs repeat), k = 1e6 as in bench.py: 1.3e8
    children for erase_create / coin, 2.3e7 for split_merge; N_c and N_u equal, the kept sets equal up to the tie band"""
    n = 1_000_000
    state = bench_state(n)
    want, wc, wu = checker.simulate(state, rid, params, n, 1e-18)
    got, gc, gu, _ = gpu_simulate(qb, state, name, params, n)
    assert (gc, gu) == (wc, wu), f"{name}: counters {(gc, gu)} vs {(wc, wu)}"
    assert wu > n
    r = bigcmp.compare(got, checker.hash_objects(got, rid), want, checker.hash_objects(want, rid), True, truncated_k=n, what=f"{name} 1e6 parents k=1e6 vs {checker.kind}")
    assert r["common"] + r["only_one_side"] == n


@pytest.mark.parametrize("name,rid,params", RULES)
def test_bench_generator_without_truncation(qb, checker, name, rid, params):
    """no truncation: the whole interference result (3e5 parents: 3.9e7 children -> about 1.1e7 unique for erase_create / coin,
    7e6 children -> 5e6 unique for split_merge), every object, every magnitude"""
    n = 300_000
    state = bench_state(n, seed=5)
    want, wc, wu = checker.simulate(state, rid, params, orc.NO_TRUNCATION, 1e-18)
    got, gc, gu, _ = gpu_simulate(qb, state, name, params, orc.NO_TRUNCATION)
    assert (gc, gu) == (wc, wu), f"{name}: counters {(gc, gu)} vs {(wc, wu)}"
    r = bigcmp.compare(got, checker.hash_objects(got, rid), want, checker.hash_objects(want, rid), True, what=f"{name} 3e5 parents untruncated vs {checker.kind}")
    assert r["common"] == wu and r["bytes_compared"] >= min(wu, 100000)


def test_loop_state_second_pass(qb, checker):
    """the state bench.py times is NOT the fresh one: one pass of the loop on 2e5 parents, every call checked FROM THE
    CHECKER'S INPUT (grown names, ragged sizes, N_u / N_c ~ 0.5: the regime where on-chip merging cannot help)"""
    n = 200_000
    state = bench_state(n, seed=9)
    sm, ec = [PI / 4, PI / 4, PI / 4], [PI / 4, 0.0, 0.0]
    for _ in range(2):
        for name, rid, params in (("split_merge", orc.RULE_SPLIT_MERGE, sm), ("erase_create", orc.RULE_ERASE_CREATE, ec)):
            state = checker.apply_modifier(state, orc.MOD_STEP)
            want, wc, wu = checker.simulate(state, rid, params, n, 1e-18)
            got, gc, gu, _ = gpu_simulate(qb, state, name, params, n)
            assert (gc, gu) == (wc, wu), f"{name}: counters {(gc, gu)} vs {(wc, wu)}"
            bigcmp.compare(got, checker.hash_objects(got, rid), want, checker.hash_objects(want, rid), True, truncated_k=n if wu > n else None,
                           what=f"loop state {name} vs {checker.kind}")
            state = want


def register_superposition(nq, free_bits):
    """all bit strings over `free_bits` qubits of an nq-qubit register (the others 0), magnitude = the product of
    len(free_bits) factors 1/sqrt(2.) rounded one multiplication at a time, as a chain of hadamard iterations leaves it"""
    m = 1.0
    for _ in free_bits:
        m *= 1 / math.sqrt(2.0)
    n = 1 << len(free_bits)
    idx = np.arange(n, dtype=np.uint64)
    data = np.zeros((n, nq), np.uint8)
    for j, bit in enumerate(free_bits):
        data[:, bit] = (idx >> np.uint64(j)) & np.uint64(1)
    mags = np.zeros((n, 2))
    mags[:, 0] = m
    return orc.Packed(np.full(n, nq, np.uint32), mags, data.reshape(-1))


def test_register_2_pow_22_interfering_and_doubling_steps(qb, checker):
    """configs[2] at 2^22 objects of 22 bytes (align 0): H on a qubit of the FULL superposition -- 2^23 children, 2^22 hashes,
    half of which cancel to exactly 0 and fail the tolerance test -- and the doubling step 2^21 -> 2^22; survivors' magnitudes
    must match to 1e-12 (in fact to the bit before normalisation: all parents carry the same magnitude)"""
    nq = 22
    qb.config.align_byte_length = 0
    try:
        for free, bit, expect_nu in ((list(range(nq)), 0, 1 << (nq - 1)), (list(range(nq - 1)), nq - 1, 1 << nq)):
            state = register_superposition(nq, free)
            want, wc, wu = checker.simulate(state, orc.RULE_HADAMARD, [bit], orc.NO_TRUNCATION, 1e-30)
            got, gc, gu, _ = gpu_simulate(qb, state, "hadamard", [bit], orc.NO_TRUNCATION, tol=1e-30)
            assert (gc, gu) == (wc, wu) == (2 * state.n, expect_nu)
            bigcmp.compare(got, checker.hash_objects(got, orc.RULE_HADAMARD, [0]), want, checker.hash_objects(want, orc.RULE_HADAMARD, [0]), False,
                           what=f"hadamard({bit}) on {state.n} objects vs {checker.kind}")
    finally:
        qb.config.align_byte_length = 8


def test_modifiers_1e7_objects(qb, port):
    """configs[1] at 1e7 objects of 8 bytes: the phase modifier (reads the object, writes the magnitude) and Ygate (writes
    both) -- bytes exact, magnitudes to the bit (one complex product per object, same rounding as the reference)"""
    n = 10_000_000
    rng = np.random.default_rng(1)
    data = rng.integers(0, 2, size=8 * n, dtype=np.uint8)
    phi = 2 * np.pi * (np.arange(n) % 1024) / 1024
    mags = np.stack([np.cos(phi), np.sin(phi)], axis=1) / math.sqrt(n)
    state = orc.Packed(np.full(n, 8, np.uint32), mags, data)
    for mod, mid, params in (("phase", orc.MOD_PHASE, [0.3]), ("ygate", orc.MOD_YGATE, [3]), ("cnot", orc.MOD_CNOT, [1, 6])):
        it = qb.Iteration()
        it.upload_packed(state.sizes, state.mags, state.data)
        qb.simulate(it, qb.Modifier(mod, *params))
        sizes, gm, gd = it.download_packed()
        want = port.apply_modifier(state, mid, params)
        assert np.array_equal(gd, want.data), f"{mod}: bytes differ"
        assert np.array_equal(gm, want.mags), f"{mod}: magnitudes differ"


def test_pop_matches_the_checker(qb, port):
    """iteration::pop (quids.hpp:194-203): tail removed, optional normalisation with total_proba = the sum before it;
    pop to the empty state; pop(0) is a no-op"""
    base = port.qcgd_random_state(5, 40, 3)
    rng = np.random.default_rng(2)
    st = orc.Packed(base.sizes, rng.normal(size=(40, 2)), base.data, total_proba=0.8)
    ragged, _, _ = port.simulate(st, orc.RULE_SPLIT_MERGE, [0.3, 0.2, 0.1], orc.NO_TRUNCATION, 1e-18)
    for state in (st, ragged):
        for n, normalize in ((1, True), (1, False), (7, True), (0, True), (state.n - 1, True), (state.n, True), (state.n, False)):
            it = qb.Iteration()
            it.upload_packed(state.sizes, state.mags, state.data, state.total_proba)
            it.pop(n, normalize)
            want = port.pop(state, n, normalize)
            sizes, mags, data = it.download_packed()
            assert it.num_object == want.n and np.array_equal(sizes, want.sizes) and np.array_equal(data, want.data), (n, normalize)
            assert np.allclose(mags, want.mags, rtol=1e-14, atol=0), (n, normalize)
            assert abs(it.total_proba - want.total_proba) <= 1e-14 * max(1.0, abs(want.total_proba)), (n, normalize, it.total_proba, want.total_proba)
    with pytest.raises(qb.QuidsError):
        it = qb.Iteration()
        it.upload_packed(st.sizes, st.mags, st.data)
        it.pop(st.n + 1)
