"""quids_b200 -- Python host mirror of the QuIDS rule-application API over the C ABI of
include/quids_b200.h (libquids_b200.so, hand-written CUDA for sm_100a).

The reference is a header-only C++ library; its drop-in C++ surface lives in include/quids/.  This
module is the thin ctypes binding used by the tests, bench.py and Python drivers.  It keeps the
reference's names and argument meaning:

    quids::it_t      -> Iteration      (append / num_object / total_proba / get_object / pop / normalize)
    quids::sy_it_t   -> SymbolicIteration (num_object / num_object_after_interferences)
    quids::rule_t    -> Rule(name, *params)        e.g. Rule("hadamard", 1), Rule("erase_create", theta, phi, xi)
    quids::modifier_t-> Modifier(name, *params)    e.g. Modifier("cnot", 1, 3), Modifier("step")
    quids::simulate(it, modifier)                         -> simulate(it, modifier)
    quids::simulate(it, rule, next, sy_it, max_num_object) -> simulate(it, rule, next, sy_it, max_num_object)

There is NO CPU fallback: loading fails loudly if the shared library has not been built, and every
compute call fails with QuidsError if no CUDA device is usable.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libquids_b200.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "quids_b200.h")

NO_TRUNCATION = 2**64 - 1
PHASES = ("num_child", "pre_truncate", "table_clear", "symbolic", "compact", "truncate", "finalize", "normalize", "exchange", "owner", "insert")


class QuidsError(RuntimeError):
    pass


def build(verbose=False):
    """compile libquids_b200.so in-tree (nvcc, sm_100a); no GPU is needed to build"""
    env = dict(os.environ)
    out = subprocess.run(["make", "-C", CSRC, "-j4"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise QuidsError("building libquids_b200.so failed")
    return LIB_PATH


class qb_options(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("align_byte_length", C.c_uint32), ("simple_truncation", C.c_int32),
                ("table_load", C.c_double), ("profile", C.c_int32), ("safety_margin", C.c_float), ("seed", C.c_uint32), ("locality_sort", C.c_int32), ("binned_inserts", C.c_int32), ("family_routing", C.c_int32), ("memory_budget", C.c_uint64),
                ("equalize", C.c_int32), ("equalize_inbalance", C.c_float), ("min_equalize_step", C.c_float), ("min_equalize_size", C.c_uint64)]


STEP_CB = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p)

_lib = None


def lib():
    """the loaded C ABI; raises if the library is missing (the product path has no fallback)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QuidsError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "or `make -C quids_b200/csrc` first (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_double
    P = C.POINTER
    sig = {
        "qb_options_default": (None, [P(qb_options)]),
        "qb_last_error": (C.c_char_p, []),
        "qb_version": (i32, []),
        "qb_device_count": (i32, []),
        "qb_ctx_create": (i32, [i32, P(vp)]),
        "qb_ctx_destroy": (i32, [vp]),
        "qb_ctx_synchronize": (i32, [vp]),
        "qb_ctx_stream": (vp, [vp]),
        "qb_ctx_launch_count": (u64, [vp]),
        "qb_host_alloc": (i32, [C.c_size_t, P(vp)]),
        "qb_host_free": (i32, [vp]),
        "qb_iter_create": (i32, [vp, P(vp)]),
        "qb_iter_destroy": (i32, [vp]),
        "qb_iter_upload": (i32, [vp, u64, vp, u64, vp, vp, vp, dbl]),
        "qb_iter_upload_f32": (i32, [vp, u64, vp, u64, vp, vp, vp, dbl]),
        "qb_iter_download_f32": (i32, [vp, vp, vp, vp, vp]),
        "qb_iter_upload_async": (i32, [vp, u64, vp, u64, vp, vp, vp, dbl]),
        "qb_iter_download_async": (i32, [vp, vp, vp, vp, vp]),
        "qb_iter_wait": (i32, [vp]),
        "qb_iter_append_state": (i32, [vp, vp]),
        "qb_iter_counts": (i32, [vp, P(u64), P(u64), P(dbl)]),
        "qb_iter_download": (i32, [vp, vp, vp, vp, vp]),
        "qb_iter_device_ptrs": (i32, [vp, P(vp), P(vp), P(vp), P(vp)]),
        "qb_iter_ctx": (vp, [vp]),
        "qb_ctx_device": (i32, [vp]),
        "qb_iter_pop": (i32, [vp, u64, i32]),
        "qb_iter_normalize": (i32, [vp]),
        "qb_sym_create": (i32, [vp, P(vp)]),
        "qb_sym_destroy": (i32, [vp]),
        "qb_sym_counts": (i32, [vp, P(u64), P(u64)]),
        "qb_sym_phase_ms": (i32, [vp, P(C.c_float)]),
        "qb_sym_device_bytes": (u64, [vp]),
        "qb_rule_id": (i32, [C.c_char_p]),
        "qb_modifier_id": (i32, [C.c_char_p]),
        "qb_apply_modifier": (i32, [vp, i32, P(dbl), u32]),
        "qb_observable_id": (i32, [C.c_char_p]),
        "qb_observable_values": (i32, [i32]),
        "qb_iter_average_value": (i32, [vp, i32, P(dbl), u32, P(dbl), u32]),
        "qb_simulate": (i32, [vp, i32, P(dbl), u32, vp, vp, u64, P(qb_options), STEP_CB, vp]),
        "qb_hash_objects": (i32, [vp, i32, P(dbl), u32, vp]),
        "qb_comm_unique_id": (i32, [vp]),
        "qb_comm_create": (i32, [vp, i32, i32, vp, P(vp)]),
        "qb_comm_destroy": (i32, [vp]),
        "qb_simulate_dist": (i32, [vp, i32, P(dbl), u32, vp, vp, vp, u64, P(qb_options), STEP_CB, vp, P(dbl)]),
        "qb_iter_send_objects": (i32, [vp, vp, u64, i32, P(u64)]),
        "qb_iter_receive_objects": (i32, [vp, vp, i32, u64, P(u64)]),
        "qb_iter_distribute_objects": (i32, [vp, vp, i32]),
        "qb_iter_gather_objects": (i32, [vp, vp, i32]),
        "qb_iter_num_symbolic_object": (i32, [vp, P(u64)]),
        "qb_iter_count_children": (i32, [vp, i32, P(dbl), u32, P(u64)]),
        "qb_iter_equalize": (i32, [vp, vp, i32, P(dbl), u32, i32, u64, C.c_float, C.c_float, P(i32)]),
        "qb_comm_allreduce_u64": (i32, [vp, vp, u32, i32]),
        "qb_comm_allreduce_f64": (i32, [vp, vp, u32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


ABI_SYMBOLS = None  # filled lazily by abi_symbols()


def abi_symbols():
    """every function name declared in include/quids_b200.h"""
    import re
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qb_[a-z0-9_]+)\s*\(", text)) - {"qb_step_cb"})


def _check(rc):
    if rc != 0:
        raise QuidsError(f"[{rc}] {lib().qb_last_error().decode(errors='replace')}")


# -------------------------------------------------------------------------------------------------
# mutable "namespace globals" of the reference (quids.hpp:60-75), passed per call through qb_options
# -------------------------------------------------------------------------------------------------
class _Globals:
    tolerance = 1e-30          # quids::tolerance
    align_byte_length = 8      # quids::align_byte_length
    simple_truncation = True   # quids::simple_truncation; False = probabilistic truncation (the reference's default)
    seed = 0                   # seed of the probabilistic truncation
    safety_margin = 0.2        # quids::safety_margin
    table_load = 0.0           # engine knob (0 = default)
    profile = False
    memory_budget = 0          # engine knob: bytes the automatic budget may spend (0 = measured on the GPU)
    equalize = 0               # distributed path: 0 off, 1 by objects, 2 by children (quids::mpi::equalize_children)
    equalize_inbalance = 0.1   # quids::mpi::equalize_inbalance
    min_equalize_step = 0.2    # quids::mpi::min_equalize_step
    min_equalize_size = 100    # quids::mpi::min_equalize_size
    locality_sort = 1          # engine knob: 0 off, 1 auto, 2 always
    binned_inserts = 1         # engine knob: 0 off, 1 for >= 2^22 children without heavy duplication (default), 2 always
    family_routing = 1         # engine knob, distributed path: parents routed to the owner of their family (erase_create, coin)

    def options(self):
        o = qb_options()
        lib().qb_options_default(C.byref(o))
        o.tolerance = self.tolerance
        o.align_byte_length = self.align_byte_length
        o.simple_truncation = 1 if self.simple_truncation else 0
        o.table_load = self.table_load
        o.safety_margin = self.safety_margin
        o.seed = self.seed
        o.profile = 1 if self.profile else 0
        o.locality_sort = self.locality_sort
        o.binned_inserts = self.binned_inserts
        o.family_routing = self.family_routing
        o.memory_budget = self.memory_budget
        o.equalize = self.equalize
        o.equalize_inbalance = self.equalize_inbalance
        o.min_equalize_step = self.min_equalize_step
        o.min_equalize_size = self.min_equalize_size
        return o


config = _Globals()


def get_alignment_offset(size, align=None):
    """quids.hpp:93-102"""
    align = config.align_byte_length if align is None else align
    if align <= 1:
        return 0
    off = align - size % align
    return 0 if off == align else off


class Context:
    """one GPU, one stream, one host thread"""

    def __init__(self, device=0):
        self.handle = C.c_void_p()
        _check(lib().qb_ctx_create(device, C.byref(self.handle)))
        self.device = device

    def close(self):
        if self.handle:
            lib().qb_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def synchronize(self):
        _check(lib().qb_ctx_synchronize(self.handle))

    @property
    def stream(self):
        return lib().qb_ctx_stream(self.handle)

    @property
    def launch_count(self):
        return int(lib().qb_ctx_launch_count(self.handle))


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def _params(params):
    arr = (C.c_double * max(1, len(params)))(*[float(p) for p in params])
    return arr, len(params)


class Rule:
    """quids::rule_t: a registered device rule + the constructor arguments of the reference class"""

    def __init__(self, name, *params):
        self.name = name
        self.id = lib().qb_rule_id(name.encode())
        if self.id < 1:
            raise QuidsError(f"unknown rule {name!r}")
        self.params = tuple(float(p) for p in params)


class Modifier:
    """quids::modifier_t: a registered device modifier"""

    def __init__(self, name, *params):
        self.name = name
        self.id = lib().qb_modifier_id(name.encode())
        if self.id < 1:
            raise QuidsError(f"unknown modifier {name!r}")
        self.params = tuple(float(p) for p in params)


class Iteration:
    """quids::iteration (quids.hpp:149-335): the state lives in HBM; host access goes through
    explicit upload/download of the reference's four arrays"""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self.handle = C.c_void_p()
        _check(lib().qb_iter_create(self.ctx.handle, C.byref(self.handle)))
        self._pending = []  # objects appended on the host, not yet uploaded

    def __del__(self):
        try:
            if self.handle and _lib is not None:
                _lib.qb_iter_destroy(self.handle)
        except Exception:
            pass

    # -- counters (quids.hpp:152-154) ---------------------------------------------------------
    def _counts(self):
        self._flush()
        n, nb, tp = C.c_uint64(), C.c_uint64(), C.c_double()
        _check(lib().qb_iter_counts(self.handle, C.byref(n), C.byref(nb), C.byref(tp)))
        return n.value, nb.value, tp.value

    @property
    def num_object(self):
        return self._counts()[0]

    @property
    def num_bytes(self):
        return self._counts()[1]

    @property
    def total_proba(self):
        return self._counts()[2]

    # -- host construction (quids.hpp:174-188) --------------------------------------------------
    def append(self, obj: bytes, mag=1.0):
        self._pending.append((bytes(obj), complex(mag)))

    def _flush(self):
        if not self._pending:
            return
        pend, self._pending = self._pending, []
        objects, begin, size, mag = self.download()
        align = config.align_byte_length
        chunks, begins, sizes, mags = [objects.tobytes()], list(begin), list(size), [mag]
        off = int(begin[-1])
        for o, m in pend:
            pad = get_alignment_offset(len(o), align)
            chunks.append(o + b"\0" * pad)
            off += len(o) + pad
            begins.append(off)
            sizes.append(len(o))
        mags.append(np.array([[m.real, m.imag] for _, m in pend], dtype=np.float64).reshape(-1, 2))
        self.upload(np.frombuffer(b"".join(chunks), dtype=np.uint8), np.array(begins, np.uint64), np.array(sizes, np.uint32),
                    np.concatenate(mags), self._raw_total_proba())

    def _raw_total_proba(self):
        tp = C.c_double()
        _check(lib().qb_iter_counts(self.handle, None, None, C.byref(tp)))
        return tp.value

    # -- bulk transfer in the reference's storage layout (quids.hpp:266-276) ----------------------
    def upload(self, objects, object_begin, object_size, magnitude, total_proba=1.0):
        objects = np.ascontiguousarray(objects, np.uint8)
        object_begin = np.ascontiguousarray(object_begin, np.uint64)
        object_size = np.ascontiguousarray(object_size, np.uint32)
        f32 = np.asarray(magnitude).dtype == np.float32  # PROBA_TYPE = float at the boundary (qb_iter_upload_f32)
        magnitude = np.ascontiguousarray(magnitude, np.float32 if f32 else np.float64).reshape(-1, 2)
        n = object_size.shape[0]
        assert object_begin.shape[0] == n + 1 and magnitude.shape[0] == n
        nbytes = int(object_begin[n]) if n else 0
        assert objects.shape[0] >= nbytes
        self._pending = []
        fn = lib().qb_iter_upload_f32 if f32 else lib().qb_iter_upload
        _check(fn(self.handle, n, objects.ctypes.data, nbytes, object_begin.ctypes.data, object_size.ctypes.data, magnitude.ctypes.data, total_proba))

    def upload_async(self, objects, object_begin, object_size, magnitude, total_proba=1.0):
        """upload on the copy stream, overlapping rule iterations on other states; the arrays must be page-locked,
        C-contiguous, of the exact dtypes, and stay untouched until wait() (qb_iter_upload_async)"""
        n = object_size.shape[0]
        assert objects.dtype == np.uint8 and object_begin.dtype == np.uint64 and object_size.dtype == np.uint32 and magnitude.dtype == np.float64
        assert object_begin.shape[0] == n + 1 and magnitude.size == 2 * n
        self._pending = []
        self._async_refs = ((getattr(self, "_async_refs", None) or []) + [(objects, object_begin, object_size, magnitude)])[-4:]  # alive until wait() (or four transfers later)
        _check(lib().qb_iter_upload_async(self.handle, n, objects.ctypes.data, int(object_begin[n]) if n else 0, object_begin.ctypes.data,
                                          object_size.ctypes.data, magnitude.ctypes.data, total_proba))

    def download_async(self, objects, object_begin, object_size, magnitude):
        """download into page-locked arrays on the copy stream; valid after wait() (qb_iter_download_async)"""
        n, nb, _ = self._counts_noflush()
        assert objects.nbytes >= nb and object_begin.shape[0] >= n + 1 and object_size.shape[0] >= n and magnitude.size >= 2 * n
        self._async_refs = ((getattr(self, "_async_refs", None) or []) + [(objects, object_begin, object_size, magnitude)])[-4:]  # alive until wait() (or four transfers later)
        _check(lib().qb_iter_download_async(self.handle, objects.ctypes.data, object_begin.ctypes.data, object_size.ctypes.data, magnitude.ctypes.data))
        return n, nb

    def wait(self):
        """block the host until the asynchronous transfers of this state are complete"""
        _check(lib().qb_iter_wait(self.handle))
        self._async_refs = None

    def upload_packed(self, sizes, mags, data, total_proba=1.0, align=None):
        """objects given back to back without padding: lay them out with align_byte_length"""
        sizes = np.ascontiguousarray(sizes, np.uint32)
        data = np.ascontiguousarray(data, np.uint8)
        align = config.align_byte_length if align is None else align
        n = sizes.shape[0]
        s64 = sizes.astype(np.uint64)
        if align > 1:
            padded = (s64 + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
        else:
            padded = s64
        begin = np.zeros(n + 1, np.uint64)
        np.cumsum(padded, out=begin[1:])
        src_begin = np.zeros(n + 1, np.uint64)
        np.cumsum(s64, out=src_begin[1:])
        objects = np.zeros(int(begin[n]) if n else 0, np.uint8)
        if n:
            if (padded == s64).all():
                objects[:] = data[:int(src_begin[n])]
            elif (sizes == sizes[0]).all():
                # same size everywhere: pad the rows of a matrix
                objects.reshape(n, int(padded[0]))[:, :int(sizes[0])] = data[:int(src_begin[n])].reshape(n, int(sizes[0]))
            else:
                # scatter every byte to its padded position
                owner = np.repeat(np.arange(n), sizes)
                pos = np.arange(int(src_begin[n]), dtype=np.uint64) - src_begin[owner] + begin[owner]
                objects[pos.astype(np.int64)] = data[:int(src_begin[n])]
        self.upload(objects, begin, sizes, mags, total_proba)

    def download(self, dtype=np.float64):
        """the reference's four arrays; dtype=np.float32 reads the magnitudes as complex<float> (qb_iter_download_f32)"""
        if self._pending:
            self._flush()
        n, nb, _ = self._counts_noflush()
        objects = np.zeros(nb, np.uint8)
        begin = np.zeros(n + 1, np.uint64)
        size = np.zeros(n, np.uint32)
        mag = np.zeros((n, 2), dtype)
        fn = lib().qb_iter_download_f32 if np.dtype(dtype) == np.float32 else lib().qb_iter_download
        _check(fn(self.handle, objects.ctypes.data, begin.ctypes.data, size.ctypes.data, mag.ctypes.data))
        return objects, begin, size, mag

    def _counts_noflush(self):
        n, nb, tp = C.c_uint64(), C.c_uint64(), C.c_double()
        _check(lib().qb_iter_counts(self.handle, C.byref(n), C.byref(nb), C.byref(tp)))
        return n.value, nb.value, tp.value

    def download_packed(self):
        """(sizes, mags, data) with the alignment padding removed"""
        objects, begin, size, mag = self.download()
        n = size.shape[0]
        if n == 0:
            return size, mag, np.zeros(0, np.uint8)
        total = int(size.sum(dtype=np.uint64))
        if total == int(begin[n]):
            return size, mag, objects[:total].copy()
        if (size == size[0]).all() and int(begin[n]) % n == 0:
            return size, mag, objects.reshape(n, int(begin[n]) // n)[:, :int(size[0])].reshape(-1).copy()
        owner = np.repeat(np.arange(n), size)
        dst_begin = np.zeros(n + 1, np.uint64)
        np.cumsum(size.astype(np.uint64), out=dst_begin[1:])
        pos = np.arange(total, dtype=np.uint64) - dst_begin[owner] + begin[owner]
        return size, mag, objects[pos.astype(np.int64)]

    def get_object(self, object_id):
        """(bytes, magnitude) of one object (quids.hpp:254-258)"""
        objects, begin, size, mag = self.download()
        b = int(begin[object_id])
        return objects[b:b + int(size[object_id])].tobytes(), complex(mag[object_id, 0], mag[object_id, 1])

    def append_state(self, other: "Iteration"):
        """append every object of `other` (HBM to HBM, no normalisation)"""
        self._flush()
        other._flush()
        _check(lib().qb_iter_append_state(self.handle, other.handle))

    def pop(self, n=1, normalize=True):
        self._flush()
        _check(lib().qb_iter_pop(self.handle, n, 1 if normalize else 0))

    def average_value(self, observable, *params):
        """iteration::average_value (quids.hpp:208-234) for a registered DEVICE observable: the sum over the objects
        of observable(object) * |mag|^2, reduced in HBM.  Returns a float, or a list for observables that produce
        several values per object ("qcgd_stats": nodes, nodes^2, density, density^2)."""
        self._flush()
        oid = lib().qb_observable_id(observable.encode())
        if oid < 1:
            raise QuidsError(f"unknown observable {observable!r}")
        k = lib().qb_observable_values(oid)
        out = (C.c_double * k)()
        p = (C.c_double * max(1, len(params)))(*[float(x) for x in params])
        _check(lib().qb_iter_average_value(self.handle, oid, p, len(params), out, k))
        return out[0] if k == 1 else list(out)

    def normalize(self):
        self._flush()
        _check(lib().qb_iter_normalize(self.handle))

    # -- object migration over NCCL (quids_mpi.hpp:124-231, 903-1077) ----------------------------
    def send_objects(self, num_object_sent, node, communicator):
        """send the last objects to rank `node` (which calls receive_objects) and pop them; returns how many moved"""
        self._flush()
        moved = C.c_uint64()
        _check(lib().qb_iter_send_objects(self.handle, communicator.handle, num_object_sent, node, C.byref(moved)))
        return moved.value

    def receive_objects(self, node, communicator, max_mem=NO_TRUNCATION):
        self._flush()
        moved = C.c_uint64()
        _check(lib().qb_iter_receive_objects(self.handle, communicator.handle, node, max_mem, C.byref(moved)))
        return moved.value

    def distribute_objects(self, communicator, node_id=0):
        self._flush()
        _check(lib().qb_iter_distribute_objects(self.handle, communicator.handle, node_id))

    def gather_objects(self, communicator, node_id=0):
        self._flush()
        _check(lib().qb_iter_gather_objects(self.handle, communicator.handle, node_id))

    def equalize(self, communicator, rule: "Rule" = None, max_rounds=1, min_equalize_size=0, equalize_inbalance=-1.0, min_equalize_step=0.0):
        """pairing rounds of equalize (rule None: by objects) or equalize_symbolic (by the children of `rule`); returns the rounds run"""
        self._flush()
        rounds = C.c_int()
        p, k = _params(rule.params) if rule else (None, 0)
        _check(lib().qb_iter_equalize(self.handle, communicator.handle, rule.id if rule else 0, p, k, max_rounds, min_equalize_size,
                                      equalize_inbalance, min_equalize_step, C.byref(rounds)))
        return rounds.value

    def count_children(self, rule: "Rule"):
        """children the local objects have under `rule` (compute_num_child, quids.hpp:548-569)"""
        self._flush()
        n = C.c_uint64()
        p, k = _params(rule.params)
        _check(lib().qb_iter_count_children(self.handle, rule.id, p, k, C.byref(n)))
        return n.value

    def hashes(self, rule: Rule):
        """rule->hasher over every object"""
        self._flush()
        n = self._counts_noflush()[0]
        out = np.zeros(n, np.uint64)
        p, k = _params(rule.params)
        _check(lib().qb_hash_objects(self.handle, rule.id, p, k, out.ctypes.data))
        return out


class SymbolicIteration:
    """quids::symbolic_iteration (quids.hpp:338-429): interference table and scratch, reused across calls"""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self.handle = C.c_void_p()
        _check(lib().qb_sym_create(self.ctx.handle, C.byref(self.handle)))

    def __del__(self):
        try:
            if self.handle and _lib is not None:
                _lib.qb_sym_destroy(self.handle)
        except Exception:
            pass

    def _counts(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(lib().qb_sym_counts(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def num_object(self):
        return self._counts()[0]

    @property
    def num_object_after_interferences(self):
        return self._counts()[1]

    @property
    def phase_ms(self):
        arr = (C.c_float * len(PHASES))()
        _check(lib().qb_sym_phase_ms(self.handle, arr))
        return dict(zip(PHASES, [float(x) for x in arr]))

    @property
    def device_bytes(self):
        return int(lib().qb_sym_device_bytes(self.handle))


def simulate(iteration: Iteration, rule, next_iteration: Iteration = None, symbolic_iteration: SymbolicIteration = None,
             max_num_object=NO_TRUNCATION, mid_step_function=None):
    """quids::simulate: the modifier overload (quids.hpp:436) when `rule` is a Modifier, the rule
    overload (quids.hpp:448) otherwise.  max_num_object = NO_TRUNCATION is the reference's -1."""
    iteration._flush()
    if isinstance(rule, Modifier):
        p, k = _params(rule.params)
        _check(lib().qb_apply_modifier(iteration.handle, rule.id, p, k))
        return
    assert next_iteration is not None and symbolic_iteration is not None
    next_iteration._pending = []
    opt = config.options()
    cb = STEP_CB(lambda label, user: mid_step_function(label.decode())) if mid_step_function else C.cast(None, STEP_CB)
    p, k = _params(rule.params)
    _check(lib().qb_simulate(iteration.handle, rule.id, p, k, next_iteration.handle, symbolic_iteration.handle, max_num_object,
                             C.byref(opt), cb, None))


# -------------------------------------------------------------------------------------------------
# distributed path: quids::mpi (quids_mpi.hpp) -- one process per GPU, NCCL inside the library
# -------------------------------------------------------------------------------------------------
class Communicator:
    """stands where MPI_Comm stands in quids::mpi::simulate (quids_mpi.hpp:423)"""

    def __init__(self, ctx, world_size, rank, unique_id: bytes):
        assert len(unique_id) == 128
        self.ctx, self.world_size, self.rank = ctx, world_size, rank
        self.handle = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(lib().qb_comm_create(ctx.handle, world_size, rank, buf, C.byref(self.handle)))

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _check(lib().qb_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch(cls, ctx, dist):
        """bootstrap from an initialised torch.distributed process group (any backend): rank 0
        creates the NCCL id, everybody receives it"""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(ctx, world, rank, box[0])

    def close(self):
        if self.handle:
            lib().qb_comm_destroy(self.handle)
            self.handle = C.c_void_p()

    def allreduce_u64(self, values, op_max=False):
        arr = np.ascontiguousarray(values, np.uint64).copy()
        _check(lib().qb_comm_allreduce_u64(self.handle, arr.ctypes.data, arr.shape[0], 1 if op_max else 0))
        return arr

    def allreduce_f64(self, values):
        arr = np.ascontiguousarray(values, np.float64).copy()
        _check(lib().qb_comm_allreduce_f64(self.handle, arr.ctypes.data, arr.shape[0]))
        return arr


def mpi_simulate(iteration: Iteration, rule: Rule, next_iteration: Iteration, symbolic_iteration: SymbolicIteration, communicator: Communicator,
                 max_num_object=NO_TRUNCATION, mid_step_function=None):
    """quids::mpi::simulate (quids_mpi.hpp:423): every rank passes its own share of the state; the
    truncation keeps the max_num_object most probable objects over ALL ranks.  Returns this rank's
    node_total_proba (quids_mpi.hpp:67,892); next_iteration.total_proba is the global sum."""
    iteration._flush()
    next_iteration._pending = []
    opt = config.options()
    cb = STEP_CB(lambda label, user: mid_step_function(label.decode())) if mid_step_function else C.cast(None, STEP_CB)
    p, k = _params(rule.params)
    node = C.c_double()
    _check(lib().qb_simulate_dist(iteration.handle, rule.id, p, k, next_iteration.handle, symbolic_iteration.handle, communicator.handle,
                                  max_num_object, C.byref(opt), cb, None, C.byref(node)))
    return node.value
