"""ctypes access to the CPU checkers of oracle/ (TEST INFRASTRUCTURE, see oracle/oracle_api.h).

Two libraries implement the same C interface:
  * oracle/liboracle.so          -- the from-scratch restatement ("port")
  * oracle/_ref/libquids_ref.so  -- the unmodified reference headers ("reference"); only present
                                    where it was built (this container) or shipped (gpurun)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libquids_ref.so")

RULE_HADAMARD, RULE_ERASE_CREATE, RULE_COIN, RULE_SPLIT_MERGE = 1, 2, 3, 4
MOD_CNOT, MOD_XGATE, MOD_YGATE, MOD_ZGATE, MOD_STEP, MOD_REVERSED_STEP, MOD_PHASE = 1, 2, 3, 4, 5, 6, 7
OBS_QCGD_SIZE, OBS_QCGD_SQUARED_SIZE, OBS_QCGD_DENSITY, OBS_QCGD_SQUARED_DENSITY, OBS_QUBIT, OBS_BYTES = 1, 2, 3, 4, 5, 6
NO_TRUNCATION = 2**64 - 1

QCGD_RULES = (RULE_ERASE_CREATE, RULE_COIN, RULE_SPLIT_MERGE)


def build():
    """(re)build the checkers with oracle/Makefile; the reference one only where /root/reference exists"""
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, env=env, stdout=subprocess.DEVNULL)


class Packed:
    """a state in the packed interchange form: sizes u32[n], mags f64[n,2], bytes u8[sum sizes]"""

    def __init__(self, sizes, mags, data, total_proba=1.0):
        self.sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        self.mags = np.ascontiguousarray(mags, dtype=np.float64).reshape(-1, 2)
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        self.total_proba = float(total_proba)
        assert self.mags.shape[0] == self.sizes.shape[0]
        assert int(self.sizes.sum(dtype=np.uint64)) == self.data.shape[0]

    @property
    def n(self):
        return int(self.sizes.shape[0])

    @property
    def begin(self):
        b = np.zeros(self.n + 1, dtype=np.uint64)
        np.cumsum(self.sizes, dtype=np.uint64, out=b[1:])
        return b

    def objects(self):
        b = self.begin
        return [bytes(self.data[int(b[i]):int(b[i + 1])]) for i in range(self.n)]

    @property
    def cmags(self):
        return self.mags[:, 0] + 1j * self.mags[:, 1]

    @staticmethod
    def from_objects(objs, mags):
        sizes = np.array([len(o) for o in objs], dtype=np.uint32)
        data = np.frombuffer(b"".join(objs), dtype=np.uint8) if objs else np.zeros(0, np.uint8)
        m = np.array([[complex(z).real, complex(z).imag] for z in mags], dtype=np.float64).reshape(-1, 2)
        return Packed(sizes, m, data)


def canonical_qcgd(obj: bytes) -> bytes:
    """mask the 4 indeterminate padding bytes of every sub_node (SURVEY 8c byte-exactness caveat)"""
    a = bytearray(obj)
    n = int.from_bytes(a[0:2], "little")
    first = 4 + 4 * n
    for off in range(first, len(a), 16):
        a[off + 4:off + 8] = b"\0\0\0\0"
    return bytes(a)


class Oracle:
    def __init__(self, path=PORT_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        vp, u64, u32, dbl, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_double, C.c_int
        L.orc_kind.restype = C.c_char_p
        L.orc_num_threads.restype = i32
        L.orc_set_num_threads.restype = i32
        L.orc_set_num_threads.argtypes = [i32]
        L.orc_state_create.restype = vp
        L.orc_state_destroy.argtypes = [vp]
        L.orc_state_load.argtypes = [vp, u64, vp, vp, vp]
        L.orc_state_num_object.argtypes = [vp]
        L.orc_state_num_object.restype = u64
        L.orc_state_num_bytes.argtypes = [vp]
        L.orc_state_num_bytes.restype = u64
        L.orc_state_total_proba.argtypes = [vp]
        L.orc_state_total_proba.restype = dbl
        L.orc_state_store.argtypes = [vp, vp, vp, vp]
        L.orc_qcgd_random_state.argtypes = [vp, u32, u64, u32, dbl, dbl]
        L.orc_hash_objects.argtypes = [vp, i32, vp, vp]
        L.orc_apply_modifier.argtypes = [vp, i32, vp]
        L.orc_simulate.argtypes = [vp, i32, vp, vp, u64, dbl, vp]
        L.orc_last_simulate_seconds.restype = dbl
        self.kind = L.orc_kind().decode()
        self.num_threads = L.orc_num_threads()

    def set_num_threads(self, n=0):
        """host threads of the following simulate calls (0 = every core: len(os.sched_getaffinity(0)))"""
        if n <= 0:
            n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.num_threads = self.lib.orc_set_num_threads(int(n))
        return self.num_threads

    # -- state handles ------------------------------------------------------------------
    def _new(self):
        return C.c_void_p(self.lib.orc_state_create())

    def _free(self, h):
        self.lib.orc_state_destroy(h)

    def _load(self, h, p: Packed):
        self.lib.orc_state_load(h, p.n, p.sizes.ctypes.data, p.mags.ctypes.data, p.data.ctypes.data)

    def _store(self, h) -> Packed:
        n = self.lib.orc_state_num_object(h)
        nb = self.lib.orc_state_num_bytes(h)
        sizes = np.zeros(n, np.uint32)
        mags = np.zeros((n, 2), np.float64)
        data = np.zeros(nb, np.uint8)
        self.lib.orc_state_store(h, sizes.ctypes.data, mags.ctypes.data, data.ctypes.data)
        return Packed(sizes, mags, data, self.lib.orc_state_total_proba(h))

    @staticmethod
    def _params(params):
        p = np.zeros(4, np.float64)
        p[:len(params)] = params
        return p

    # -- operations on packed states ----------------------------------------------------
    def qcgd_random_state(self, n_node, n_graphs, seed, re=None, im=0.0) -> Packed:
        if re is None:
            re = float(np.float32(1) / np.sqrt(np.float32(n_graphs)))  # qcgd.hpp:1126
        h = self._new()
        try:
            self.lib.orc_qcgd_random_state(h, n_node, n_graphs, seed, re, im)
            return self._store(h)
        finally:
            self._free(h)

    def hash_objects(self, p: Packed, rule_id, params=()) -> np.ndarray:
        h = self._new()
        try:
            self._load(h, p)
            out = np.zeros(p.n, np.uint64)
            pr = self._params(params)
            rc = self.lib.orc_hash_objects(h, rule_id, pr.ctypes.data, out.ctypes.data)
            assert rc == 0
            return out
        finally:
            self._free(h)

    def average_value(self, p: Packed, observable_id, params=()) -> float:
        """iteration::average_value (quids.hpp:208-234) with one of the ORC_OBS_* observables"""
        h = self._new()
        try:
            self._load(h, p)
            pr = self._params(params)
            out = C.c_double()
            self.lib.orc_average_value.restype = C.c_int
            self.lib.orc_average_value.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
            rc = self.lib.orc_average_value(h, observable_id, pr.ctypes.data, C.byref(out))
            assert rc == 0
            return out.value
        finally:
            self._free(h)

    def pop(self, p: Packed, n=1, normalize=True) -> Packed:
        """iteration::pop (quids.hpp:194-203)"""
        h = self._new()
        try:
            self._load(h, p)
            self.lib.orc_state_total_proba.restype = C.c_double
            self.lib.orc_state_pop.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
            rc = self.lib.orc_state_pop(h, n, 1 if normalize else 0)
            assert rc == 0, rc
            out = self._store(h)
            if not normalize or n < 1:
                out.total_proba = p.total_proba
            return out
        finally:
            self._free(h)

    def apply_modifier(self, p: Packed, modifier_id, params=()) -> Packed:
        h = self._new()
        try:
            self._load(h, p)
            pr = self._params(params)
            rc = self.lib.orc_apply_modifier(h, modifier_id, pr.ctypes.data)
            assert rc == 0
            out = self._store(h)
            out.total_proba = p.total_proba
            return out
        finally:
            self._free(h)

    def simulate(self, p: Packed, rule_id, params=(), max_num_object=NO_TRUNCATION, tolerance=1e-30):
        """returns (next state, N_c, N_u)"""
        a, b = self._new(), self._new()
        try:
            self._load(a, p)
            pr = self._params(params)
            counters = np.zeros(2, np.uint64)
            rc = self.lib.orc_simulate(a, rule_id, pr.ctypes.data, b, max_num_object, tolerance, counters.ctypes.data)
            assert rc == 0, rc
            self.last_seconds = self.lib.orc_last_simulate_seconds()
            return self._store(b), int(counters[0]), int(counters[1])
        finally:
            self._free(a)
            self._free(b)


class LoadedState:
    """a state kept inside the checker between calls (the reference's append() is quadratic: load once)"""

    def __init__(self, oracle: Oracle, p: Packed = None):
        self.o = oracle
        self.h = oracle._new()
        if p is not None:
            oracle._load(self.h, p)

    def close(self):
        if self.h is not None:
            self.o._free(self.h)
            self.h = None

    def store(self) -> Packed:
        return self.o._store(self.h)

    def apply_modifier(self, modifier_id, params=()):
        """in place (quids.hpp:973-980)"""
        pr = self.o._params(params)
        rc = self.o.lib.orc_apply_modifier(self.h, modifier_id, pr.ctypes.data)
        assert rc == 0, rc

    def simulate_into(self, out: "LoadedState", rule_id, params=(), max_num_object=NO_TRUNCATION, tolerance=1e-30):
        """returns (N_c, N_u, seconds inside simulate)"""
        pr = self.o._params(params)
        counters = np.zeros(2, np.uint64)
        rc = self.o.lib.orc_simulate(self.h, rule_id, pr.ctypes.data, out.h, max_num_object, tolerance, counters.ctypes.data)
        assert rc == 0, rc
        return int(counters[0]), int(counters[1]), self.o.lib.orc_last_simulate_seconds()


def have_reference():
    return os.path.exists(REF_SO)


# ---------------------------------------------------------------------------------------
# comparison of two states as hash-keyed sets (positional order is unspecified)
# ---------------------------------------------------------------------------------------
def keyed(p: Packed, hashes: np.ndarray, qcgd: bool):
    objs = p.objects()
    if qcgd:
        objs = [canonical_qcgd(o) for o in objs]
    out = {}
    for h, o, m in zip(hashes.tolist(), objs, p.cmags.tolist()):
        assert h not in out, "duplicate hash inside one state"
        out[h] = (o, m)
    return out


def assert_same_state(a: Packed, ha, b: Packed, hb, qcgd: bool, rtol=1e-12, what=""):
    """exact hash set + canonical bytes, magnitudes within rtol relative (of the larger modulus)"""
    ka, kb = keyed(a, ha, qcgd), keyed(b, hb, qcgd)
    assert len(ka) == len(kb), f"{what}: {len(ka)} vs {len(kb)} objects"
    assert ka.keys() == kb.keys(), f"{what}: hash sets differ"
    for h, (oa, ma) in ka.items():
        ob, mb = kb[h]
        assert oa == ob, f"{what}: bytes differ for hash {h:016x}"
        scale = max(abs(ma), abs(mb))
        assert abs(ma - mb) <= rtol * scale, f"{what}: magnitude {ma} vs {mb} for hash {h:016x}"
    assert abs(a.total_proba - b.total_proba) <= rtol * max(abs(a.total_proba), abs(b.total_proba)), \
        f"{what}: total_proba {a.total_proba} vs {b.total_proba}"


def assert_same_truncated(a: Packed, ha, b: Packed, hb, full: Packed, hfull, k, qcgd: bool, band=1e-12, rtol=1e-10, what=""):
    """truncated parity (consequence 2).

    `full` is the un-truncated, normalised result of the same step.  Every object whose
    probability is outside the tie band around the k-th probability must be present in both a
    and b; objects inside the band may differ; both must keep exactly k objects.
    Magnitudes are compared after undoing the (slightly different) normalisation."""
    assert a.n == k and b.n == k, f"{what}: kept {a.n} / {b.n}, expected {k}"
    probs = np.sort((np.abs(full.cmags) ** 2))[::-1]
    pk = probs[k - 1]
    kf = keyed(full, hfull, qcgd)
    ka, kb = keyed(a, ha, qcgd), keyed(b, hb, qcgd)
    for h, (o, m) in kf.items():
        p = abs(m) ** 2
        if p > pk * (1 + band):
            assert h in ka and h in kb, f"{what}: object above the threshold band missing"
        elif p < pk * (1 - band):
            assert h not in ka and h not in kb, f"{what}: object below the threshold band kept"
    fa, fb = np.sqrt(a.total_proba), np.sqrt(b.total_proba)
    for h in ka.keys() & kb.keys():
        assert ka[h][0] == kb[h][0]
        ma, mb = ka[h][1] * fa, kb[h][1] * fb
        assert abs(ma - mb) <= rtol * max(abs(ma), abs(mb)), f"{what}: magnitude {ma} vs {mb}"
    assert abs(a.total_proba - b.total_proba) <= 1e-9
