#!/usr/bin/env python
"""bench.py -- children/sec per rule iteration on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3]): the reference's QCGD loop

    step; split_merge(pi/4, pi/4, pi/4); step; erase_create(pi/4)          (examples/qcgd_test.cpp, qcgd.hpp:1168-1176)

on a state of 12-node graphs saturated at max_num_object = parents (1e7 per GPU), simple truncation,
tolerance 1e-18.  A "step" is ONE pass of that loop = two rule iterations (quids::simulate, quids.hpp:448-543)
and two modifier passes (quids.hpp:436) over the resident state; the state starts as random density-1/2 graphs
and the warm-up passes bring it to the saturated regime the survey asks to time.  `value` = children generated
by the rule iterations of the timed steps / their device time.  At N > 1 every rank holds the same number of
parents (weak scaling), interference is hash-sharded over NCCL (qb_simulate_dist) and max_num_object is global.

One JSON line is printed by rank 0; see the task contract for the keys.  `value` is measured with the state
resident in HBM (CUDA events on the library's stream); `e2e` goes through the same C-ABI calls with HOST
buffers: upload of the input state and download of the result are inside the timed region.
`cpu_baseline` / `--impl reference` run the reference's own CPU implementation (oracle/_ref, every host
thread) through the same loop on a bounded sample, and the `parity` gate compares the GPU with it on that
sample before any timing counts (exit code 3 on a mismatch).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_NODE = 12
THETA = math.pi / 4
TOLERANCE = 1e-18
METRIC = "children_per_sec_per_rule_iteration"
UNIT = "children/s"
LOOP = "step; split_merge(pi/4, pi/4, pi/4); step; erase_create(pi/4)"
SM_PARAMS = [THETA, THETA, THETA]
EC_PARAMS = [THETA, 0.0, 0.0]


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_parents(n_parents, seed):
    from quids_b200 import qcgd
    sizes, data = qcgd.random_graphs(N_NODE, n_parents, seed=seed)
    mags = np.zeros((n_parents, 2))
    mags[:, 0] = qcgd.read_state_magnitude(n_parents)[0]
    return sizes, mags, data


class ClockSampler:
    """samples SM clocks and throttle reasons DURING the timed region: NVML in-process every ~2 ms (a step is a few
    milliseconds, nvidia-smi -lms would not deliver one sample in time); nvidia-smi is the fallback"""
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()
        self.thread = None
        self.source = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            index = device
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            if visible:
                try:
                    index = int(visible.split(",")[device])
                except (ValueError, IndexError):
                    index = device
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.source = "nvml"
        except Exception:
            self.nvml = None

    def _loop_nvml(self):
        n = self.nvml
        names = {n.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 n.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", n.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002)

    def _loop_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout
            except (OSError, subprocess.TimeoutExpired):
                break
            f = [x.strip() for x in out.strip().split(",")]
            if len(f) >= 6:
                try:
                    self.sm.append(float(f[0]))
                    self.max_mhz = float(f[1])
                except ValueError:
                    pass
                for name, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)

    def start(self):
        if self.nvml is None:
            self.source = "nvidia-smi"
        self.thread = threading.Thread(target=self._loop_nvml if self.nvml is not None else self._loop_smi, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=15)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self.source}


# ---- byte counts ------------------------------------------------------------------------
def step_bytes_design(n_p, s_p, n_u, n_s, s_s):
    """compulsory traffic of one rule iteration IN THIS DESIGN: parents read once; every unique child's (hash, magnitude,
    representative) written once as a 32-byte table slot and read once by the compaction, its 8-byte norm key written and
    read once by the selection (the children themselves never touch HBM: they are merged on chip / in the table);
    every survivor re-reads its parent and writes its bytes, offset, size and magnitude"""
    return n_p * (s_p + 16) + 80 * n_u + n_s * (s_p + s_s + 28)


def step_bytes_reference_dataflow(n_p, s_p, n_c, n_u, n_s, s_s):
    """the same iteration if every child's (hash, magnitude) went through memory once (the reference's data flow)"""
    return n_p * (s_p + 16) + 48 * n_c + 40 * n_u + n_s * (s_p + s_s + 16)


def symbolic_kernel_bytes(n_p, s_p, n_u):
    """compulsory traffic of the child-generation kernel: parents (+ magnitude) read once, one 32-byte slot written per unique child"""
    return n_p * (s_p + 16) + 32 * n_u


# ---- the CPU reference beside it ---------------------------------------------------------------------------------
def cpu_checker():
    """the reference's own CPU implementation (oracle/_ref) if it was built, else the port; every host thread"""
    import orc
    if orc.have_reference():
        o = orc.Oracle(orc.REF_SO)
    else:
        if not os.path.exists(orc.PORT_SO):
            orc.build()
        o = orc.Oracle(orc.PORT_SO)
    o.set_num_threads(0)  # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers
    return orc, o


def cpu_loop(sample_parents, seed, steps, warmup, keep_first=False):
    """the loop on the CPU implementation, state kept inside the checker (the reference's append() re-initialises an index
    array on every call, quids.hpp:279-306, so building a state is quadratic: it is loaded once).  Only quids::simulate is
    timed (the modifier passes are three orders of magnitude cheaper).  Returns a dict; with keep_first the states around
    the two rule calls of the first pass are kept for the parity gate."""
    orc, o = cpu_checker()
    sizes, mags, data = make_parents(sample_parents, seed)
    a, b = orc.LoadedState(o, orc.Packed(sizes, mags, data)), orc.LoadedState(o)
    kept = {"init": orc.Packed(sizes, mags, data)} if keep_first else {}
    children, seconds, per_rule = 0, 0.0, {"split_merge": [0, 0.0], "erase_create": [0, 0.0]}
    for i in range(warmup + steps):
        a.apply_modifier(orc.MOD_STEP)
        if keep_first and i == 0:
            kept["sm_in"] = a.store()
        nc1, nu1, s1 = a.simulate_into(b, orc.RULE_SPLIT_MERGE, SM_PARAMS, sample_parents, TOLERANCE)
        if keep_first and i == 0:
            kept["sm_out"], kept["sm_counts"] = b.store(), (nc1, nu1)
        b.apply_modifier(orc.MOD_STEP)
        if keep_first and i == 0:
            kept["ec_in"] = b.store()
        nc2, nu2, s2 = b.simulate_into(a, orc.RULE_ERASE_CREATE, EC_PARAMS, sample_parents, TOLERANCE)
        if keep_first and i == 0:
            kept["ec_out"], kept["ec_counts"] = a.store(), (nc2, nu2)
        if i >= warmup:
            children += nc1 + nc2
            seconds += s1 + s2
            per_rule["split_merge"][0] += nc1
            per_rule["split_merge"][1] += s1
            per_rule["erase_create"][0] += nc2
            per_rule["erase_create"][1] += s2
    a.close()
    b.close()
    return {"rate": children / seconds, "children_per_step": children / steps, "seconds_per_step": seconds / steps, "oracle": o, "orc": orc, "kept": kept,
            "per_rule": {k: {"children_per_step": v[0] / steps, "seconds_per_step": v[1] / steps, "rate": v[0] / v[1]} for k, v in per_rule.items()}}


def reference_sample_parents(args):
    """parents of the bounded CPU sample: the whole --steps/--warmup run of the reference arm must end within a few minutes
    (about 10 s per pass and million parents on 16 cores)"""
    if args.cpu_sample_parents:
        return args.cpu_sample_parents
    passes = args.steps + args.warmup
    return 1000000 if passes <= 8 else 250000


def run_reference(args):
    """--impl reference: the reference's CPU path, all host threads, same metric and loop (rank 0 only)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    parents = reference_sample_parents(args)
    r = cpu_loop(parents, 0, steps=args.steps, warmup=args.warmup)
    o = r["oracle"]
    assert o.kind == "port" or o.num_threads > 1 or (os.cpu_count() or 1) == 1, "the reference arm must use every host core"
    sample = f"{parents} random 12-node parents, max_num_object = {parents}, the loop [{LOOP}]: {r['children_per_step']:.0f} children and {r['seconds_per_step']:.2f} s " \
             f"inside quids::simulate per pass, {args.steps} passes after {args.warmup} warm-up"
    line = {"impl": "reference", "metric": METRIC, "value": r["rate"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus, parents),
            "per_rule": r["per_rule"],
            "cpu_baseline": {"value": r["rate"], "unit": UNIT, "cores": o.num_threads, "kind": o.kind, "sample": sample},
            "e2e": {"value": r["rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, parents_per_gpu=None):
    return {"workload": f"QCGD loop [{LOOP}] on 12-node graphs (random density-1/2 start, 244 B objects growing to ~330 B), state saturated at "
                        "max_num_object = parents, simple truncation, tolerance 1e-18 (BASELINE.json configs[3], SURVEY 8(d) C4; configs[4] when "
                        "hash-sharded over several GPUs); one step = one pass of the loop = 2 rule iterations + 2 modifier passes",
            "parents_per_gpu": parents_per_gpu, "n_node": N_NODE, "rules": ["split_merge", "erase_create"], "theta": THETA,
            "l2": "inputs larger than L2 (3 GB state, interference tables of 5-45 GB rewritten every rule iteration)",
            "parallelism": "1 GPU" if n_gpus == 1 else f"hash-sharded interference over {n_gpus} GPUs (NCCL)"}


# ---- parity gates ------------------------------------------------------------------------------------------------
def gpu_rule_call(qb, packed, rule, k, comm=None):
    """one rule iteration on the GPU from a packed host state -> (packed result, N_c, N_u)"""
    import orc
    a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
    a.upload_packed(packed.sizes, packed.mags, packed.data, packed.total_proba)
    qb.simulate(a, rule, b, sym, k)
    sizes, mags, data = b.download_packed()
    return orc.Packed(sizes, mags, data, b.total_proba), sym.num_object, sym.num_object_after_interferences


def parity_gate(qb, kept, o, k):
    """GPU vs the CPU implementation on the states the CPU loop went through (first pass of the sample): the modifier,
    then each rule call FROM THE REFERENCE'S INPUT (a legal tie choice must not compound), truncated to k"""
    import bigcmp
    import orc
    out = {"against": o.kind, "sample_parents": k, "rules": {}}
    sm = qb.Rule("split_merge", *SM_PARAMS)
    ec = qb.Rule("erase_create", *EC_PARAMS)
    it = qb.Iteration()
    it.upload_packed(kept["init"].sizes, kept["init"].mags, kept["init"].data)
    qb.simulate(it, qb.Modifier("step"))
    s2, m2, d2 = it.download_packed()
    assert np.array_equal(s2, kept["sm_in"].sizes) and np.array_equal(d2, kept["sm_in"].data) and np.array_equal(m2, kept["sm_in"].mags), "parity: the step modifier differs"
    out["modifier_step"] = "bytes and magnitudes identical"
    del it
    for name, rid, rule in (("split_merge", orc.RULE_SPLIT_MERGE, sm), ("erase_create", orc.RULE_ERASE_CREATE, ec)):
        key = "sm" if name == "split_merge" else "ec"
        want, (wc, wu) = kept[key + "_out"], kept[key + "_counts"]
        got, gc, gu = gpu_rule_call(qb, kept[key + "_in"], rule, k)
        assert (gc, gu) == (wc, wu), f"parity {name}: counters {(gc, gu)} vs {(wc, wu)}"
        r = bigcmp.compare(got, o.hash_objects(got, rid), want, o.hash_objects(want, rid), True, truncated_k=k if wu > k else None, what=f"parity {name}")
        r["N_c"], r["N_u"] = wc, wu
        out["rules"][name] = r
    out["ok"] = True
    return out


def dist_parity_gate(qb, comm, dist, rank, world, n_parents=100000, k=60000):
    """one quids::mpi::simulate per rule on a small state with a truncating k (fewer than the parents: the global
    pre-truncation of the parents runs too), gathered on rank 0 with the library's own gather_objects and compared with
    the CPU implementation's quids::simulate on the whole state (north_star's rule for the distributed path)"""
    import bigcmp
    import orc
    orc_mod, o = cpu_checker()
    sizes, _, data = make_parents(n_parents, 4242)
    rng = np.random.default_rng(7)
    mags = rng.normal(size=(n_parents, 2)) * np.exp(rng.normal(size=(n_parents, 1)))  # no ties among the parents
    mags /= np.sqrt((mags ** 2).sum())
    state = orc.Packed(sizes, mags, data)
    out = {"against": o.kind, "parents": n_parents, "k": k, "rules": {}}
    for name, rid, rule in (("split_merge", orc.RULE_SPLIT_MERGE, qb.Rule("split_merge", *SM_PARAMS)), ("erase_create", orc.RULE_ERASE_CREATE, qb.Rule("erase_create", *EC_PARAMS))):
        objs_begin = state.begin
        mine = np.arange(rank, state.n, world)
        a, b, sym = qb.Iteration(), qb.Iteration(), qb.SymbolicIteration()
        a.upload_packed(state.sizes[mine], state.mags[mine],
                        np.concatenate([state.data[int(objs_begin[i]):int(objs_begin[i + 1])] for i in mine]) if mine.shape[0] else np.zeros(0, np.uint8))
        qb.mpi_simulate(a, rule, b, sym, comm, k)
        total_proba = b.total_proba
        counts = comm.allreduce_u64([sym.num_object, sym.num_object_after_interferences])
        b.gather_objects(comm, 0)
        if rank == 0:
            want, wc, wu = o.simulate(state, rid, SM_PARAMS if name == "split_merge" else EC_PARAMS, k, TOLERANCE)
            s2, m2, d2 = b.download_packed()
            got = orc.Packed(s2, m2, d2, total_proba)
            assert (int(counts[0]), int(counts[1])) == (wc, wu), f"distributed parity {name}: counters {counts} vs {(wc, wu)}"
            r = bigcmp.compare(got, o.hash_objects(got, rid), want, o.hash_objects(want, rid), True, truncated_k=k if wu > k else None, what=f"distributed parity {name}")
            r["N_c"], r["N_u"] = wc, wu
            out["rules"][name] = r
            nxt = o.apply_modifier(want, orc.MOD_STEP)
            box = [nxt]
        else:
            box = [None]
        dist.broadcast_object_list(box, src=0)
        state = box[0]
    out["ok"] = True
    return out


# ---- our arm ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import quids_b200 as qb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    parents = args.parents
    qb.config.tolerance = TOLERANCE
    qb.config.align_byte_length = 8
    qb.config.profile = True
    ctx = qb.default_context()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    comm = qb.Communicator.from_torch(ctx, dist) if world > 1 else None
    k_total = parents * world
    peak, peak_src = measured_peak_gbs()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gates: before any timing counts -------------------------------------------------------------------
    parity, cpu, cpu_run = None, None, None
    try:
        if world == 1 and not args.no_cpu:
            cpu_parents = args.cpu_sample_parents or 1000000
            cpu_run = cpu_loop(cpu_parents, seed=0, steps=2, warmup=1, keep_first=True)
            o = cpu_run["oracle"]
            cpu = {"value": cpu_run["rate"], "unit": UNIT, "cores": o.num_threads, "kind": o.kind, "per_rule": cpu_run["per_rule"],
                   "sample": f"{cpu_parents} parents of the same generator, max_num_object = {cpu_parents}, the same loop: {cpu_run['children_per_step']:.0f} children and "
                             f"{cpu_run['seconds_per_step']:.2f} s inside quids::simulate per pass, 2 passes after 1 warm-up"}
            parity = parity_gate(qb, cpu_run["kept"], o, cpu_parents)
            cpu_run = None
        elif world > 1:
            parity = dist_parity_gate(qb, comm, dist, rank, world)
    except AssertionError as e:
        print(f"bench.py: PARITY GATE FAILED: {e}", file=sys.stderr, flush=True)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "parity": {"ok": False, "error": str(e)[:500]}}), flush=True)
        sys.stdout.flush()
        os._exit(3)

    # ---- the loop, state resident in HBM -----------------------------------------------------------------------------
    sizes, mags, data = make_parents(parents, seed=rank)
    a, b, sym = qb.Iteration(ctx), qb.Iteration(ctx), qb.SymbolicIteration(ctx)
    a.upload_packed(sizes, mags, data)
    del sizes, mags, data
    sm, ec, step_mod = qb.Rule("split_merge", *SM_PARAMS), qb.Rule("erase_create", *EC_PARAMS), qb.Modifier("step")

    def rule_call(src, rule, dst):
        if comm is None:
            qb.simulate(src, rule, dst, sym, parents)
        else:
            qb.mpi_simulate(src, rule, dst, sym, comm, k_total)

    def loop_pass(src, mid, record=None):
        """src -> (step, split_merge) -> mid -> (step, erase_create) -> src"""
        for rule, x, y in ((sm, src, mid), (ec, mid, src)):
            qb.simulate(x, step_mod)
            if record is not None:
                n_p, nb_p, _ = x._counts_noflush()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            rule_call(x, rule, y)
            if record is not None:
                e1.record(stream)
                n_s, nb_s, _ = y._counts_noflush()
                record.append({"rule": rule.name, "events": (e0, e1), "N_p": n_p, "S_p": nb_p / max(1, n_p), "N_c": sym.num_object, "N_u": sym.num_object_after_interferences,
                               "N_s": n_s, "S_s": nb_s / max(1, n_s), "phase_ms": sym.phase_ms})

    for _ in range(args.warmup):
        loop_pass(a, b)
    launches0 = ctx.launch_count
    sampler = ClockSampler(local)
    calls = []
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        loop_pass(a, b, calls)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    elapsed_ms = e0.elapsed_time(e1)
    children = sum(c["N_c"] for c in calls)
    if dist is not None:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        c = torch.tensor([children], device="cuda", dtype=torch.int64)
        dist.all_reduce(c)
        children_total = int(c[0].item())
    else:
        children_total = children
    ms_per_step = elapsed_ms / args.steps
    value = children_total / (elapsed_ms / 1e3)

    # per rule: time, counts, bytes, the kernel that dominates it
    per_rule = {}
    for name in ("split_merge", "erase_create"):
        mine = [c for c in calls if c["rule"] == name]
        n = len(mine)
        ms = sum(c["events"][0].elapsed_time(c["events"][1]) for c in mine) / n
        avg = {key: sum(c[key] for c in mine) / n for key in ("N_p", "S_p", "N_c", "N_u", "N_s", "S_s")}
        phase = {p: sum(c["phase_ms"][p] for c in mine) / n for p in mine[0]["phase_ms"]}
        design = step_bytes_design(avg["N_p"], avg["S_p"], avg["N_u"], avg["N_s"], avg["S_s"])
        refflow = step_bytes_reference_dataflow(avg["N_p"], avg["S_p"], avg["N_c"], avg["N_u"], avg["N_s"], avg["S_s"])
        sym_bytes = symbolic_kernel_bytes(avg["N_p"], avg["S_p"], avg["N_u"])
        per_rule[name] = {"ms_per_call": ms, "children_per_s": avg["N_c"] / (ms / 1e3), "counts": avg, "phase_ms": phase, "dominant_phase": max(phase, key=phase.get),
                          "whole_iteration": {"algorithmic_bytes": design, "achieved": design / (ms / 1e3) / 1e9, "frac": design / (ms / 1e3) / 1e9 / peak,
                                              "reference_dataflow_equivalent": {"bytes": refflow, "achieved": refflow / (ms / 1e3) / 1e9}},
                          "symbolic_kernel": {"algorithmic_bytes_per_launch": sym_bytes, "kernel_ms": phase["symbolic"],
                                              "achieved": sym_bytes / (phase["symbolic"] / 1e3) / 1e9 if phase["symbolic"] > 0 else None,
                                              "frac": sym_bytes / (phase["symbolic"] / 1e3) / 1e9 / peak if phase["symbolic"] > 0 else None}}
    for c in calls:
        del c["events"]

    # ---- end to end through the C ABI with HOST buffers: every step uploads its input state from pinned host memory,
    # runs one pass of the loop and downloads the result state; the transfers run on the library's copy streams and are
    # double-buffered, so the upload of step i+1 and the download of step i-1 overlap the rule iterations of step i ------
    objects, begin, size, mag = a.download()  # the saturated state every e2e step starts from
    h2d = objects.nbytes + begin.nbytes + size.nbytes + mag.nbytes
    pinned = [torch.from_numpy(x.copy()).pin_memory() for x in (objects, begin, size, mag.reshape(-1))]
    host = [p.numpy() for p in pinned]
    host[1], host[2] = host[1].view(np.uint64), host[2].view(np.uint32)
    del objects, begin, size, mag
    # three states in rotation: while step i computes, the input of step i+1 is uploaded into one and the result of step i-1
    # is downloaded from another, so that the two directions of the host link run at the same time
    ROT = 3
    ins, mids = [qb.Iteration(ctx) for _ in range(ROT)], b

    def pinned_result(cap_n, cap_b):
        return [torch.empty(cap_b, dtype=torch.uint8).pin_memory().numpy(), torch.empty(cap_n + 1, dtype=torch.int64).pin_memory().numpy().view(np.uint64),
                torch.empty(cap_n, dtype=torch.int32).pin_memory().numpy().view(np.uint32), torch.empty(cap_n * 2, dtype=torch.float64).pin_memory().numpy()]

    n1, nb1, _ = a._counts_noflush()
    out_host = [pinned_result(int(n1 * 1.5) + 4096, int(nb1 * 1.5) + 4096) for _ in range(ROT)]
    d2h_total, e2e_children = 0, 0

    def e2e_run(steps):
        nonlocal d2h_total, e2e_children
        d2h_total = e2e_children = 0
        ins[0].upload_async(*host)
        for i in range(steps):
            cur = i % ROT
            if i + 1 < steps:
                ins[(i + 1) % ROT].upload_async(*host)  # overlaps this step's rule iterations and the download of the previous result
            rec = []
            loop_pass(ins[cur], mids, rec)  # the result of the pass is back in ins[cur]
            e2e_children += sum(c["N_c"] for c in rec)
            n2, nb2, _ = ins[cur]._counts_noflush()
            if nb2 > out_host[cur][0].nbytes or n2 + 1 > out_host[cur][1].shape[0]:
                ins[cur].wait()
                out_host[cur] = pinned_result(int(n2 * 1.5) + 4096, int(nb2 * 1.5) + 4096)
            ins[cur].download_async(*out_host[cur])  # overlaps the next step
            d2h_total += nb2 + 8 * (n2 + 1) + 4 * n2 + 16 * n2
        for it in ins:
            it.wait()

    e2e_steps = max(2, min(args.steps, 10))
    e2e_run(ROT)  # untimed: allocations
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    d2h = d2h_total // e2e_steps
    last = (e2e_steps - 1) % ROT
    check = ins[last].download()
    n_last = check[2].shape[0]
    assert np.array_equal(check[3].reshape(-1), out_host[last][3][:2 * n_last]) and np.array_equal(check[2], out_host[last][2][:n_last]) and \
        np.array_equal(check[0], out_host[last][0][:check[0].shape[0]]), "e2e: the host copy differs from the state in HBM"
    del check
    e2e_children_total = e2e_children
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        c = torch.tensor([e2e_children], device="cuda", dtype=torch.int64)
        dist.all_reduce(c)
        e2e_children_total = int(c[0].item())
    e2e_value = e2e_children_total / e2e_steps / (e2e_ms / 1e3)
    del ins, out_host, pinned, host

    # ---- best case (round 1's headline, kept for continuity): erase_create on FRESH random graphs -- only 4^12 distinct
    # 12-node fresh-name graphs exist, so everything interferes on chip (N_u / N_c ~ 0.01) -----------------------------------
    best = None
    if world == 1 and not args.no_best_case:
        sizes, mags, data = make_parents(parents, seed=0)
        a.upload_packed(sizes, mags, data)
        del sizes, mags, data
        for _ in range(3):
            qb.simulate(a, ec, b, sym, parents)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        for _ in range(5):
            qb.simulate(a, ec, b, sym, parents)
        b1.record(stream)
        barrier()
        bms = b0.elapsed_time(b1) / 5
        n_c, n_u, n_s = sym.num_object, sym.num_object_after_interferences, b.num_object
        design = step_bytes_design(parents, 248, n_u, n_s, 248)
        best = {"workload": f"erase_create(pi/4) on {parents} fresh random 12-node graphs (duplicates kept), max_num_object = parents", "value": n_c / (bms / 1e3), "unit": UNIT,
                "ms_per_step": bms, "steps": 5, "counts": {"N_p": parents, "N_c": n_c, "N_u": n_u, "N_s": n_s}, "phase_ms": sym.phase_ms,
                "whole_iteration": {"algorithmic_bytes": design, "achieved": design / (bms / 1e3) / 1e9, "frac": design / (bms / 1e3) / 1e9 / peak}}
        # the 1e8-object configuration on ONE GPU (BASELINE.json configs[4] without the sharding), same generator
        if args.large_parents > parents:
            big, big_next, chunk_state = qb.Iteration(ctx), qb.Iteration(ctx), qb.Iteration(ctx)
            done = 0
            from quids_b200 import qcgd
            while done < args.large_parents:
                n_chunk = min(parents, args.large_parents - done)
                cs, cm, cd = make_parents(n_chunk, seed=1000 + done // parents)
                cm[:, 0] = qcgd.read_state_magnitude(args.large_parents)[0]
                chunk_state.upload_packed(cs, cm, cd)
                big.append_state(chunk_state)
                done += n_chunk
            del chunk_state
            for _ in range(2):
                qb.simulate(big, ec, big_next, sym, args.large_parents)
            barrier()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record(stream)
            for _ in range(3):
                qb.simulate(big, ec, big_next, sym, args.large_parents)
            l1.record(stream)
            barrier()
            lms = l0.elapsed_time(l1) / 3
            lc, lu, ls = sym.num_object, sym.num_object_after_interferences, big_next.num_object
            free_b, total_b = torch.cuda.mem_get_info()
            ldesign = step_bytes_design(args.large_parents, 248, lu, ls, 248)
            best["large"] = {"workload": f"the same on {args.large_parents} parents on one GPU", "value": lc / (lms / 1e3), "unit": UNIT, "ms_per_step": lms, "steps": 3,
                             "counts": {"N_p": args.large_parents, "N_c": lc, "N_u": lu, "N_s": ls}, "phase_ms": sym.phase_ms, "hbm_in_use_gb": (total_b - free_b) / 1e9,
                             "whole_iteration": {"algorithmic_bytes": ldesign, "achieved": ldesign / (lms / 1e3) / 1e9, "frac": ldesign / (lms / 1e3) / 1e9 / peak}}
            del big, big_next

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the child-generation kernel of the rule that takes most of the step -------
    dom_rule = max(per_rule, key=lambda r: per_rule[r]["symbolic_kernel"]["kernel_ms"])
    dk = per_rule[dom_rule]["symbolic_kernel"]
    kernel_name = {"erase_create": "symbolic_items_kernel<qcgd::erase_create_fused> (children of the sorted work items, families accumulated on chip, one 4 KB table region per run)",
                   "split_merge": "symbolic_kernel<qcgd::split_merge_fused> (one lane per child: walk, hash, record / interference-table insert)"}[dom_rule]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(dom_rule)
        if tj and tj.get("parents") == parents:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": kernel_name, "rule": dom_rule, "achieved": dk["achieved"], "peak": peak, "unit": "GB/s", "frac": dk["frac"], "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dk["algorithmic_bytes_per_launch"], "kernel_ms": dk["kernel_ms"],
                "algorithmic_bytes": "parents (object + magnitude) read once + one 32-byte (hash, magnitude, representative) slot written per unique child; the children "
                                     "themselves are merged on chip and never reach HBM (DESIGN.md section 4)",
                "kernel_share_of_step": sum(c["phase_ms"]["symbolic"] for c in calls if c["rule"] == dom_rule) / args.steps / ms_per_step,
                "traffic_frac": (traffic / (dk["kernel_ms"] / 1e3) / 1e9 / peak) if traffic else None,
                "whole_step": {"algorithmic_bytes": sum(per_rule[r]["whole_iteration"]["algorithmic_bytes"] for r in per_rule),
                               "frac": sum(per_rule[r]["whole_iteration"]["algorithmic_bytes"] for r in per_rule) / (ms_per_step / 1e3) / 1e9 / peak},
                "per_rule": {r: per_rule[r]["symbolic_kernel"] for r in per_rule}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, parents),
            "counts": {"children_per_step": children_total / args.steps, "parents_per_gpu": parents},
            "per_rule": per_rule, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "how": "per step: qb_iter_upload_async of the saturated input state from pinned host memory, one pass of the loop (2 x qb_apply_modifier, 2 x qb_simulate), "
                           "qb_iter_download_async of the result state into pinned host memory; three states in rotation on the library's copy streams (upload of step i+1 and "
                           "download of step i-1 overlap the rule iterations of step i), wall clock over all steps incl. the first upload and the last download"},
            "gpu_launches": int(launches), "best_case": best}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--parents", type=int, default=10**7, help="parents per GPU (= max_num_object per GPU)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the parity gate against it")
    ap.add_argument("--no-best-case", action="store_true", help="skip the best-case legs (fresh random graphs, 1e8 parents)")
    ap.add_argument("--cpu-sample-parents", type=int, default=0, help="parents of the bounded CPU sample (0 = sized from --steps/--warmup)")
    ap.add_argument("--large-parents", type=int, default=10**8, help="N = 1 only: also time erase_create on this many fresh parents on the one GPU (0 = skip)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: whatever the libraries print on file descriptor 1 meanwhile (NCCL's version
    # banner, for one) goes to stderr; the line itself is written to the saved descriptor by print()
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
