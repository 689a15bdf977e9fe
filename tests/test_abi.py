"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol declared in
include/quids_b200.h, and refuses to compute without a GPU instead of falling back."""
import ctypes
import os
import subprocess

import pytest

import quids_b200 as qb


@pytest.fixture(scope="module")
def library():
    if not os.path.exists(qb.LIB_PATH):
        qb.build()
    return qb.lib()


def test_every_declared_symbol_is_exported(library):
    names = qb.abi_symbols()
    assert len(names) >= 30
    raw = ctypes.CDLL(qb.LIB_PATH)
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing


def test_no_torch_or_oracle_dependency(library):
    out = subprocess.run(["ldd", qb.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "quids_ref" not in out


def test_registry_names(library):
    for name in ("hadamard", "erase_create", "coin", "split_merge", "hadamard_generic", "erase_create_generic", "coin_generic", "split_merge_generic"):
        assert library.qb_rule_id(name.encode()) >= 1
    for name in ("cnot", "xgate", "ygate", "zgate", "step", "reversed_step", "phase"):
        assert library.qb_modifier_id(name.encode()) >= 1
    for name, values in (("qcgd_stats", 4), ("qcgd_size", 1), ("qubit", 1), ("object_bytes", 1)):
        oid = library.qb_observable_id(name.encode())
        assert oid >= 1 and library.qb_observable_values(oid) == values
    assert library.qb_observable_id(b"no_such_observable") == -3
    assert library.qb_rule_id(b"no_such_rule") == -3
    assert b"no_such_rule" in library.qb_last_error()


def test_options_default(library):
    o = qb.qb_options()
    library.qb_options_default(ctypes.byref(o))
    assert (o.tolerance, o.align_byte_length, o.simple_truncation) == (1e-30, 8, 1)


def test_no_cpu_fallback(library):
    if library.qb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(qb.QuidsError):
        qb.Context(0)


def test_alignment_offset():
    assert [qb.get_alignment_offset(s, 8) for s in (0, 1, 7, 8, 9, 244)] == [0, 7, 1, 0, 7, 4]
    assert qb.get_alignment_offset(5, 0) == 0 and qb.get_alignment_offset(5, 1) == 0
