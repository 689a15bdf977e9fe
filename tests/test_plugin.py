"""The plug-in boundary (north_star: "rules supplied as __device__-callable implementations", "the in-place
modifier lambda path"): device rules, a registered modifier and modifier LAMBDAS written by a user in examples/custom_rule.cu
are built OUT OF TREE -- nvcc against include/quids/device/ and libquids_b200.so, nothing from quids_b200/csrc -- loaded as a
plug-in, driven by name through the C ABI, and compared with the CPU checker.

reference interface: class rule (src/quids.hpp:105-146), modifier_t (:86), simulate(it_t&, modifier_t) (:436-438)."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "examples")
PLUGIN = os.path.join(EXAMPLES, "libcustom_rule.so")
DRIVER = os.path.join(EXAMPLES, "custom_rule.out")


def build_plugin():
    import quids_b200 as qb
    if not os.path.exists(qb.LIB_PATH):
        qb.build()
    out = subprocess.run(["make", "-C", EXAMPLES, "plugin"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout[-3000:]


@pytest.fixture(scope="module")
def plugin():
    """the user's module, loaded AFTER the library: its static initialisers call qb::register_rule / register_modifier"""
    import quids_b200 as qb
    build_plugin()
    qb.lib()
    return ctypes.CDLL(PLUGIN, mode=ctypes.RTLD_GLOBAL)


def test_plugin_builds_out_of_tree_and_registers_by_name(plugin):
    """no GPU needed: nvcc cross-compiles, loading the module registers the names in the library's registry"""
    import quids_b200 as qb
    lib = qb.lib()
    for name in ("user_hadamard", "user_ry"):
        assert lib.qb_rule_id(name.encode()) >= 1, name
    assert lib.qb_modifier_id(b"user_swap") >= 1
    src = open(os.path.join(EXAMPLES, "custom_rule.cu")).read()
    assert "csrc" not in src.replace("quids_b200/csrc", ""), "the example must not include anything from quids_b200/csrc"
    cmd = subprocess.run(["make", "-C", EXAMPLES, "-n", "-B", "libcustom_rule.so"], stdout=subprocess.PIPE, text=True).stdout
    assert "csrc" not in cmd and "-lquids_b200" in cmd and "include" in cmd


def random_register(n_qubits, n_objects, seed):
    rng = np.random.default_rng(seed)
    objs = sorted({bytes(rng.integers(0, 2, size=n_qubits, dtype=np.uint8)) for _ in range(n_objects)})
    mags = rng.normal(size=(len(objs), 2))
    return orc.Packed.from_objects(objs, mags[:, 0] + 1j * mags[:, 1])


@pytest.mark.gpu
def test_user_rule_matches_the_checker(plugin, port):
    """user_hadamard (the four reference methods only, compiled outside the library) against the checker's hadamard,
    interference included: H twice on the same qubit gives the state back"""
    from gpu_engine import GpuEngine
    import quids_b200 as qb
    eng = GpuEngine(align=0)
    eng.rule = lambda rid, params: qb.Rule("user_hadamard", *list(params)[:1])
    state = random_register(9, 300, 5)
    for bit in (0, 4, 8, 4):
        want, nc, nu = port.simulate(state, orc.RULE_HADAMARD, [bit])
        got, gc, gu = eng.simulate(state, orc.RULE_HADAMARD, [bit])
        assert (gc, gu) == (nc, nu)
        orc.assert_same_state(got, port.hash_objects(got, orc.RULE_HADAMARD, [0]), want, port.hash_objects(want, orc.RULE_HADAMARD, [0]), False, what=f"user_hadamard({bit})")
        state = want


@pytest.mark.gpu
def test_user_rule_the_library_does_not_ship(plugin):
    """user_ry(bit, theta): amplitudes of a product state, then the inverse rotations interfere back to one object"""
    import quids_b200 as qb
    qb.config.align_byte_length, qb.config.tolerance = 0, 1e-20
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.append(bytes(5), 1.0)
    theta = 1.1
    for bit in range(5):
        qb.simulate(a, qb.Rule("user_ry", bit, theta), b, sym)
        a, b = b, a
    sizes, mags, data = a.download_packed()
    assert a.num_object == 32
    ones = data.reshape(32, 5).sum(axis=1)
    expect = np.cos(theta / 2) ** (5 - ones) * np.sin(theta / 2) ** ones
    assert np.allclose(mags[:, 0], expect, rtol=1e-13, atol=0) and np.all(mags[:, 1] == 0)
    for bit in range(5):
        qb.simulate(a, qb.Rule("user_ry", bit, -theta), b, sym)
        a, b = b, a
    obj, mag = a.get_object(0)
    assert a.num_object == 1 and obj == bytes(5) and abs(mag - 1) < 1e-13
    qb.config.align_byte_length, qb.config.tolerance = 8, 1e-30


@pytest.mark.gpu
def test_modifier_lambdas_and_registered_modifier_match_the_checker(plugin, port):
    """[=] __device__ lambdas applied to a state handle (qb::apply_device_modifier) against the checker's phase modifier and
    Xgate; the modifier registered by name against numpy"""
    import quids_b200 as qb
    plugin.user_phase_lambda.argtypes = [ctypes.c_void_p, ctypes.c_double]
    plugin.user_xgate_lambda.argtypes = [ctypes.c_void_p, ctypes.c_uint]
    state = random_register(8, 500, 11)
    for align in (0, 8):
        qb.config.align_byte_length = align
        it = qb.Iteration()
        it.upload_packed(state.sizes, state.mags, state.data, align=align)
        assert plugin.user_phase_lambda(it.handle, 0.7) == 0
        want = port.apply_modifier(state, orc.MOD_PHASE, [0.7])
        sizes, mags, data = it.download_packed()
        assert np.array_equal(data, want.data) and np.allclose(mags, want.mags, rtol=1e-15, atol=1e-18)
        assert plugin.user_xgate_lambda(it.handle, 3) == 0
        want = port.apply_modifier(want, orc.MOD_XGATE, [3])
        sizes, mags2, data = it.download_packed()
        assert np.array_equal(data, want.data) and np.array_equal(mags2, mags)
        qb.simulate(it, qb.Modifier("user_swap", 1, 6))
        sizes, mags3, data = it.download_packed()
        swapped = want.data.reshape(-1, 8).copy()
        swapped[:, [1, 6]] = swapped[:, [6, 1]]
        assert np.array_equal(data, swapped.reshape(-1)) and np.array_equal(mags3, mags)
    qb.config.align_byte_length = 8


@pytest.mark.gpu
def test_driver_with_user_rules_and_lambda_through_the_header_api(plugin):
    """examples/custom_rule.out: quids::it_t / quids::simulate with a user rule class, quids::simulate(state, lambda), a
    registered modifier; the program checks itself against closed-form amplitudes"""
    out = subprocess.run([DRIVER], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0, out.stdout
    lines = [l for l in out.stdout.splitlines() if l.startswith(("ok", "FAILED"))]
    assert len(lines) == 5 and all(l.startswith("ok") for l in lines), out.stdout
