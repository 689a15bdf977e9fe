"""Every configuration of BASELINE.json on one B200, one JSON line each (bench.py stays the contract line for
configs[3]; this script is where the other configs get their measured rate and roofline fraction).

    python scripts/bench_configs.py [c1] [c2] [c3] [c4] [c5single]  > gpurun_out/configs.jsonl

Rates are device time per call (CUDA events on the library's stream around the whole call); the roofline
fraction uses SURVEY 8(d)'s algorithmic bytes and MEASURED_PEAKS.json's copy bandwidth.  Inputs are larger
than L2 except for c1 (the reference's own CPU-runnable case), which is reported for completeness.
"""
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quids_b200 as qb  # noqa: E402
from quids_b200 import qcgd  # noqa: E402

PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
qb.config.profile = True
ctx = qb.default_context()
stream = torch.cuda.ExternalStream(ctx.stream)


def device_ms(fn, reps=1):
    """CUDA events on the library's stream around `reps` calls"""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    ctx.synchronize()
    return e0.elapsed_time(e1) / reps


def emit(config, what, units, unit_name, ms, alg_bytes, extra=None):
    gbs = alg_bytes / (ms / 1e3) / 1e9
    line = {"config": config, "what": what, "ms": ms, "rate": units / (ms / 1e3), "unit": unit_name + "/s", "algorithmic_bytes": alg_bytes,
            "achieved_gbs": gbs, "peak_gbs": PEAK, "frac": gbs / PEAK}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def rule_bytes(n_p, s_p, n_c, n_u, n_s, s_s):
    return n_p * (s_p + 16) + 48 * n_c + 40 * n_u + n_s * (s_p + s_s + 16)


def c1():
    """configs[0]: hadamard on a 12-qubit register to full superposition and back (1 -> 4096 -> 1 objects)"""
    qb.config.align_byte_length, qb.config.tolerance = 0, 1e-30
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.append(bytes(12), 1.0)
    total_children, t0 = 0, time.perf_counter()
    for bit in list(range(12)) + list(reversed(range(12))):
        qb.simulate(a, qb.Rule("hadamard", bit), b, sym)
        total_children += sym.num_object
        a, b = b, a
    ms = (time.perf_counter() - t0) * 1e3
    assert a.num_object == 1
    emit("c1", "24 hadamard iterations on a 12-qubit register (launch-latency bound, in cache)", total_children, "children", ms, rule_bytes(8190, 12, total_children, 8190, 8190, 12),
         {"final_objects": a.num_object, "wall_clock": True})


def c2(n=5 * 10**8):
    """configs[1]: in-place modifier over 5e8 fixed-size 8-byte objects"""
    qb.config.align_byte_length = 8
    it = qb.Iteration()
    rng = np.random.default_rng(1)
    objects = rng.integers(0, 256, size=8 * n, dtype=np.uint8)
    begin = np.arange(n + 1, dtype=np.uint64) * 8
    size = np.full(n, 8, np.uint32)
    mags = np.zeros((n, 2))
    phi = 2 * np.pi * (np.arange(n) % 1024) / 1024
    mags[:, 0], mags[:, 1] = np.cos(phi) / math.sqrt(n), np.sin(phi) / math.sqrt(n)
    it.upload(objects, begin, size, mags)
    del objects, begin, size, mags, phi
    for name, params, writes_object in (("phase", (0.3,), False), ("ygate", (3,), True)):
        m = qb.Modifier(name, *params)
        for _ in range(3):
            qb.simulate(it, m)
        ms = device_ms(lambda: qb.simulate(it, m), reps=5)
        emit("c2", f"modifier {name} over {n:.0e} objects of 8 B", n, "objects", ms, n * (8 + 32 + (8 if writes_object else 0)))


def c3(nq=24):
    """configs[2]: hadamard on a 24-qubit register driven to full superposition, no truncation"""
    qb.config.align_byte_length, qb.config.tolerance = 0, 1e-30
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.append(bytes(nq), 1.0)
    for bit in range(nq):
        ms = device_ms(lambda: qb.simulate(a, qb.Rule("hadamard", bit), b, sym))
        if bit == nq - 1:
            n_p, n_c, n_u, n_s = a.num_object, sym.num_object, sym.num_object_after_interferences, b.num_object
            emit("c3", f"hadamard last doubling 2^{nq - 1} -> 2^{nq} (first call at this size)", n_c, "children", ms, rule_bytes(n_p, nq, n_c, n_u, n_s, nq),
                 {"N_p": n_p, "N_c": n_c, "N_u": n_u, "N_s": n_s, "phase_ms": sym.phase_ms})
        a, b = b, a
    # interfering step on the full superposition and the doubling that follows, repeated (steady state)
    for rep in range(3):
        ms = device_ms(lambda: qb.simulate(a, qb.Rule("hadamard", 0), b, sym))
        n_p, n_c, n_u, n_s = a.num_object, sym.num_object, sym.num_object_after_interferences, b.num_object
        emit("c3", f"hadamard interfering step 2^{nq} -> 2^{nq - 1} rep {rep}", n_c, "children", ms, rule_bytes(n_p, nq, n_c, n_u, n_s, nq),
             {"N_p": n_p, "N_c": n_c, "N_u": n_u, "N_s": n_s, "phase_ms": sym.phase_ms})
        ms = device_ms(lambda: qb.simulate(b, qb.Rule("hadamard", 0), a, sym))
        n_p, n_c, n_u, n_s = b.num_object, sym.num_object, sym.num_object_after_interferences, a.num_object
        emit("c3", f"hadamard doubling 2^{nq - 1} -> 2^{nq} rep {rep}", n_c, "children", ms, rule_bytes(n_p, nq, n_c, n_u, n_s, nq),
             {"N_p": n_p, "N_c": n_c, "N_u": n_u, "N_s": n_s, "phase_ms": sym.phase_ms})


def qcgd_state(n, seed=0):
    sizes, data = qcgd.random_graphs(12, n, seed=seed)
    mags = np.zeros((n, 2))
    mags[:, 0] = qcgd.read_state_magnitude(n)[0]
    return sizes, mags, data


def c4(n=10**7):
    """configs[3]: QCGD rules on 1e7 random 12-node graphs, max_num_object = 1e7"""
    qb.config.align_byte_length, qb.config.tolerance = 8, 1e-18
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.upload_packed(*qcgd_state(n))
    t = math.pi / 4
    for rule in (qb.Rule("erase_create", t, 0.0, 0.0), qb.Rule("coin", t, 0.0, 0.0), qb.Rule("split_merge", t, t, t)):
        for _ in range(2):
            qb.simulate(a, rule, b, sym, n)
        ms = device_ms(lambda: qb.simulate(a, rule, b, sym, n), reps=3)
        n_c, n_u, n_s = sym.num_object, sym.num_object_after_interferences, b.num_object
        s_s = b.num_bytes / max(1, n_s)
        emit("c4", f"{rule.name} on {n:.0e} random 12-node parents, k = {n:.0e}", n_c, "children", ms, rule_bytes(n, 248, n_c, n_u, n_s, s_s),
             {"N_p": n, "N_c": n_c, "N_u": n_u, "N_s": n_s, "mean_child_bytes": s_s, "phase_ms": sym.phase_ms})
    # the reference's sequence (qcgd_test.cpp): step; split_merge; step; erase_create, state saturated at k
    step = qb.Modifier("step")
    sm, ec = qb.Rule("split_merge", t, t, t), qb.Rule("erase_create", t, 0.0, 0.0)
    for iteration in range(3):
        for rule in (sm, ec):
            qb.simulate(a, step)
            n_p, s_p = a.num_object, a.num_bytes / max(1, a.num_object)
            ms = device_ms(lambda: qb.simulate(a, rule, b, sym, n))
            n_c, n_u, n_s = sym.num_object, sym.num_object_after_interferences, b.num_object
            s_s = b.num_bytes / max(1, n_s)
            emit("c4", f"sequence iteration {iteration}: {rule.name}", n_c, "children", ms, rule_bytes(n_p, s_p, n_c, n_u, n_s, s_s),
                 {"N_p": n_p, "N_c": n_c, "N_u": n_u, "N_s": n_s, "mean_parent_bytes": s_p, "mean_child_bytes": s_s, "phase_ms": sym.phase_ms})
            a, b = b, a


def c5single(n=10**8):
    """configs[4] on ONE GPU: 1e8 random 12-node parents (the reference needs 8 ranks for this: 1.3e10 children x 87 B)"""
    qb.config.align_byte_length, qb.config.tolerance = 8, 1e-18
    chunk = 10**7
    objects = np.zeros((n, 248), np.uint8)
    for i in range(n // chunk):
        _, data = qcgd.random_graphs(12, chunk, seed=i)
        objects[i * chunk:(i + 1) * chunk, :244] = data.reshape(chunk, 244)
    mags = np.zeros((n, 2))
    mags[:, 0] = qcgd.read_state_magnitude(n)[0]
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.upload(objects.reshape(-1), np.arange(n + 1, dtype=np.uint64) * 248, np.full(n, 244, np.uint32), mags)
    del objects, mags
    rule = qb.Rule("erase_create", math.pi / 4, 0.0, 0.0)
    for _ in range(2):
        qb.simulate(a, rule, b, sym, n)
    ms = device_ms(lambda: qb.simulate(a, rule, b, sym, n), reps=3)
    n_c, n_u, n_s = sym.num_object, sym.num_object_after_interferences, b.num_object
    free, total = torch.cuda.mem_get_info()
    emit("c5single", f"erase_create on {n:.0e} random 12-node parents on ONE B200, k = {n:.0e}", n_c, "children", ms, rule_bytes(n, 248, n_c, n_u, n_s, 248),
         {"N_p": n, "N_c": n_c, "N_u": n_u, "N_s": n_s, "phase_ms": sym.phase_ms, "hbm_used_gb": (total - free) / 1e9})


if __name__ == "__main__":
    todo = sys.argv[1:] or ["c1", "c2", "c3", "c4"]
    for name in todo:
        globals()[name]()
