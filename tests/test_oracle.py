"""Pins the CPU restatement (oracle/oracle.cpp): known answers, golden vectors made by the
unmodified reference, and -- where oracle/_ref exists -- the live reference on random inputs."""
import math

import numpy as np
import pytest

import golden_util
import orc


class PortEngine:
    def __init__(self, o):
        self.o = o

    def simulate(self, p, rid, params, k, tol):
        return self.o.simulate(p, rid, params, k, tol)

    def apply_modifier(self, p, mid, params):
        return self.o.apply_modifier(p, mid, params)


def test_default_hasher_known_answers(port):
    # libstdc++ std::hash<string_view>, g++ 13.3 in this image
    z = orc.Packed.from_objects([b"\0" * 12, bytes([1, 0, 0, 0]), bytes([1, 1, 0, 0])], [1, 1, 1])
    h = port.hash_objects(z, orc.RULE_HADAMARD, [0])
    assert [hex(x) for x in h] == ["0xabfb1a4626dbb342", "0x1e194d667accaf15", "0x3046654aacfd7af4"]


def test_qcgd_fresh_graph_known_answers(port):
    # make_graph with all bits 0
    for n, size, want in ((1, 24, 0x59572c1820095bfc), (3, 64, 0xe4556617d0d779a7), (12, 244, 0xbed7c6ba66b86de0)):
        g = port.qcgd_random_state(n, 1, 0, 1.0)
        fresh = bytearray(g.objects()[0])
        fresh[2:2 + 2 * n] = bytes(2 * n)
        st = orc.Packed.from_objects([bytes(fresh)], [1])
        assert st.sizes[0] == size
        assert int(port.hash_objects(st, orc.RULE_ERASE_CREATE, [0, 0, 0])[0]) == want


def test_hadamard_known_answer(port):
    st = orc.Packed.from_objects([bytes([1, 1, 0, 0])], [0.5 + 0.25j])
    nxt, nc, nu = port.simulate(st, orc.RULE_HADAMARD, [1])
    assert (nc, nu, nxt.n) == (2, 2, 2)
    f = math.sqrt(nxt.total_proba)
    got = dict(zip(nxt.objects(), (nxt.cmags * f).tolist()))
    assert got[bytes([1, 0, 0, 0])] == complex(0.35355339059327373, 0.17677669529663687)
    assert got[bytes([1, 1, 0, 0])] == -complex(0.35355339059327373, 0.17677669529663687)


def test_split_merge_known_answer(port):
    g = bytearray(port.qcgd_random_state(3, 1, 0, 1.0).objects()[0])
    g[2:5], g[5:8] = bytes([1, 1, 0]), bytes([1, 0, 1])
    st = orc.Packed.from_objects([bytes(g)], [1])
    nxt, nc, nu = port.simulate(st, orc.RULE_SPLIT_MERGE, [0.25, 0.25, 0.25])
    assert sorted(nxt.sizes.tolist()) == [64, 76, 116, 128]
    want = {0x09ae092a706239a2, 0xc67ae82883b35ab6, 0xf08a12507a782496, 0x2ed3a638b2767827}
    assert set(port.hash_objects(nxt, orc.RULE_SPLIT_MERGE, [0, 0, 0]).tolist()) == want


@pytest.mark.parametrize("name", golden_util.fixtures())
def test_port_reproduces_golden(port, name):
    assert golden_util.replay(name, PortEngine(port), port) > 0


def test_zero_max_num_object_is_refused(port):
    st = orc.Packed.from_objects([bytes(4)], [1])
    with pytest.raises(AssertionError):
        port.simulate(st, orc.RULE_HADAMARD, [0], max_num_object=0)


def test_empty_state(port):
    st = orc.Packed.from_objects([], [])
    nxt, nc, nu = port.simulate(st, orc.RULE_HADAMARD, [0])
    assert (nxt.n, nc, nu) == (0, 0, 0)
    assert nxt.total_proba == 0.0  # quids.hpp:986-992


@pytest.mark.parametrize("rule_id", orc.QCGD_RULES)
@pytest.mark.parametrize("n_node,n_graphs,seed", [(1, 3, 5), (2, 6, 6), (8, 20, 7), (11, 4, 8)])
def test_port_matches_live_reference(port, reference, rule_id, n_node, n_graphs, seed):
    params = [0.37, 0.21, -0.4]
    a = reference.qcgd_random_state(n_node, n_graphs, seed)
    b = port.qcgd_random_state(n_node, n_graphs, seed)
    assert [orc.canonical_qcgd(o) for o in a.objects()] == [orc.canonical_qcgd(o) for o in b.objects()]
    state = b
    for it in range(2):
        ra, nca, nua = reference.simulate(state, rule_id, params, tolerance=1e-18)
        rb, ncb, nub = port.simulate(state, rule_id, params, tolerance=1e-18)
        assert (nca, nua) == (ncb, nub)
        orc.assert_same_state(rb, port.hash_objects(rb, rule_id, params), ra, reference.hash_objects(ra, rule_id, params), True,
                              what=f"rule {rule_id} it {it}")
        state = port.apply_modifier(rb, orc.MOD_STEP)
        ref_state = reference.apply_modifier(rb, orc.MOD_STEP)
        assert state.objects() == ref_state.objects()
        if state.n > 3000:
            break


def test_port_matches_live_reference_hadamard_ragged(port, reference):
    rng = np.random.default_rng(3)
    objs = [bytes(rng.integers(0, 2, size=int(l), dtype=np.uint8)) for l in rng.integers(3, 40, size=50)]
    objs = list(dict.fromkeys(objs))
    mags = rng.normal(size=len(objs)) + 1j * rng.normal(size=len(objs))
    st = orc.Packed.from_objects(objs, mags)
    for bit in (0, 2, 1, 2):
        ra, nca, nua = reference.simulate(st, orc.RULE_HADAMARD, [bit])
        rb, ncb, nub = port.simulate(st, orc.RULE_HADAMARD, [bit])
        assert (nca, nua) == (ncb, nub)
        orc.assert_same_state(rb, port.hash_objects(rb, 1, [0]), ra, reference.hash_objects(ra, 1, [0]), False)
        st = rb


def test_modifiers_match_live_reference(port, reference):
    rng = np.random.default_rng(4)
    objs = [bytes(rng.integers(0, 2, size=6, dtype=np.uint8)) for _ in range(20)]
    st = orc.Packed.from_objects(objs, rng.normal(size=20) + 1j * rng.normal(size=20))
    for mid, params in ((orc.MOD_CNOT, [1, 3]), (orc.MOD_XGATE, [2]), (orc.MOD_YGATE, [0]), (orc.MOD_ZGATE, [3]), (orc.MOD_PHASE, [0.3])):
        a, b = port.apply_modifier(st, mid, params), reference.apply_modifier(st, mid, params)
        assert a.objects() == b.objects()
        assert np.array_equal(a.mags, b.mags)
        st = a


def test_average_value_port_matches_reference(port):
    """iteration::average_value (quids.hpp:208-234) with the observables of utils::serialize (qcgd.hpp:319-345):
    restatement against the reference's own average_value where the reference checker was built"""
    rng = np.random.default_rng(11)
    g = port.qcgd_random_state(9, 400, 6)
    g.mags[:] = rng.normal(size=g.mags.shape) / 20
    grown, _, _ = port.simulate(g, orc.RULE_SPLIT_MERGE, [0.3, 0.2, 0.1], tolerance=1e-18)  # graphs of several sizes
    reg = orc.Packed.from_objects([bytes(rng.integers(0, 2, size=l, dtype=np.uint8)) for l in rng.integers(3, 9, size=300)], rng.normal(size=300) + 1j * rng.normal(size=300))
    # hand-checked: two 2-node graphs, |mag|^2 = 0.25 and 0.75
    two = port.qcgd_random_state(2, 2, 1)
    two.mags[:] = [[0.5, 0], [0, math.sqrt(0.75)]]
    assert abs(port.average_value(two, orc.OBS_QCGD_SIZE) - 2.0) < 1e-15
    assert abs(port.average_value(two, orc.OBS_BYTES) - 44.0) < 1e-13
    if not orc.have_reference():
        pytest.skip("oracle/_ref not built here")
    ref = orc.Oracle(orc.REF_SO)
    for st, ids in ((grown, (orc.OBS_QCGD_SIZE, orc.OBS_QCGD_SQUARED_SIZE, orc.OBS_QCGD_DENSITY, orc.OBS_QCGD_SQUARED_DENSITY, orc.OBS_BYTES)), (reg, (orc.OBS_QUBIT, orc.OBS_BYTES))):
        for oid in ids:
            for params in ([0], [2], [7]) if oid == orc.OBS_QUBIT else ([],):
                a, b = port.average_value(st, oid, params), ref.average_value(st, oid, params)
                assert abs(a - b) <= 1e-12 * max(1.0, abs(b)), (oid, params, a, b)


def test_split_merge_wide_graphs_port_matches_reference(port):
    """graphs of 33-700 nodes with hand-placed split / merge / wrap-around sites (the inputs of the GPU test of the same
    name): the restatement against the reference, two iterations"""
    if not orc.have_reference():
        pytest.skip("oracle/_ref not built here")
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

    objs = [graph(40, [3, 20], [10, 30], True), graph(40, [0, 17], [5], False), graph(33, [32], [1, 8], False),
            graph(700, [3, 400], [10, 650], True), graph(600, [0], [100], False)]
    ref = orc.Oracle(orc.REF_SO)
    state = orc.Packed.from_objects(objs, [0.2, 0.4j, 0.4, 0.6, 0.5])
    for _ in range(2):
        a, ac, au = port.simulate(state, orc.RULE_SPLIT_MERGE, [0.3, 0.2, 0.1], tolerance=1e-18)
        b, bc, bu = ref.simulate(state, orc.RULE_SPLIT_MERGE, [0.3, 0.2, 0.1], tolerance=1e-18)
        assert (ac, au) == (bc, bu)
        orc.assert_same_state(a, port.hash_objects(a, orc.RULE_SPLIT_MERGE), b, ref.hash_objects(b, orc.RULE_SPLIT_MERGE), True, what="wide graphs")
        # the grown graphs with fresh random magnitudes: iterating on the coherent result would un-split everything and leave
        # residues of cancelled sums, for which a relative tolerance means nothing
        rng = np.random.default_rng(5)
        state = orc.Packed(b.sizes, rng.normal(size=b.mags.shape) / 10, b.data)


def test_port_matches_live_reference_on_random_rule_sequences(port, reference):
    """random sequences of QCGD rules and steps on small random states (25 seeds): the restatement follows the reference
    through grown names, merges back and interference, one operation at a time from the reference's state"""
    rng = np.random.default_rng(2024)
    for seed in range(25):
        n_node, n_graphs = int(rng.integers(2, 8)), int(rng.integers(1, 6))
        state = reference.qcgd_random_state(n_node, n_graphs, 100 + seed)
        mags = rng.normal(size=(state.n, 2))
        state = orc.Packed(state.sizes, mags / np.sqrt((mags ** 2).sum()), state.data)
        for step in range(4):
            rule_id = int(rng.choice(orc.QCGD_RULES))
            params = [float(x) for x in rng.uniform(-1.5, 1.5, size=3)]
            ra, nca, nua = reference.simulate(state, rule_id, params, tolerance=1e-18)
            rb, ncb, nub = port.simulate(state, rule_id, params, tolerance=1e-18)
            assert (nca, nua) == (ncb, nub), (seed, step, rule_id)
            orc.assert_same_state(rb, port.hash_objects(rb, rule_id, params), ra, reference.hash_objects(ra, rule_id, params), True, rtol=1e-10,
                                  what=f"seed {seed} step {step} rule {rule_id}")
            mod_id = int(rng.choice([orc.MOD_STEP, orc.MOD_REVERSED_STEP]))
            state = reference.apply_modifier(ra, mod_id)
            assert port.apply_modifier(ra, mod_id).objects() == state.objects()
            if state.n > 1500 or state.n == 0:
                break


def test_pop_port_matches_reference(port, reference):
    """iteration::pop + normalize (quids.hpp:194-203, 985-1017): the restatement against the reference's own method"""
    base = port.qcgd_random_state(5, 30, 3)
    rng = np.random.default_rng(2)
    st = orc.Packed(base.sizes, rng.normal(size=(30, 2)), base.data)
    for n, normalize in ((1, True), (1, False), (7, True), (0, True), (29, True), (30, True), (30, False)):
        a, b = port.pop(st, n, normalize), reference.pop(st, n, normalize)
        assert a.n == b.n == 30 - n and a.objects() == b.objects()
        assert np.allclose(a.mags, b.mags, rtol=1e-14, atol=0) and abs(a.total_proba - b.total_proba) <= 1e-14 * max(1.0, b.total_proba)
