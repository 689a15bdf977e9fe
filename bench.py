#!/usr/bin/env python
"""bench.py -- children/sec per rule iteration on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one rule iteration (quids::simulate, quids.hpp:448-543) over one resident batch of
synthetic input.  Workload at N = 1 (BASELINE.json configs[3]): QCGD erase_create
(theta = pi/4) on 1e7 random density-1/2 12-node graphs (244-byte objects), max_num_object = 1e7,
simple truncation, tolerance 1e-18.  At N > 1 every rank holds the same number of parents (weak
scaling) and the interference step is hash-sharded over NCCL (qb_simulate_dist).

One JSON line is printed by rank 0; see the task contract for the keys.  `value` is measured with
the state resident in HBM (CUDA events on the library's stream); `e2e` goes through the same C-ABI
calls with HOST buffers: upload of the input state and download of the result are inside the timed
region.  `cpu_baseline` / `--impl reference` time the reference's own CPU implementation
(oracle/_ref, all host threads) on a bounded sample of the same workload.
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

N_NODE = 12
THETA = math.pi / 4
TOLERANCE = 1e-18
METRIC = "children_per_sec_per_rule_iteration"
UNIT = "children/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def qcgd_magnitude(n_parents):
    from quids_b200 import qcgd
    return qcgd.read_state_magnitude(n_parents)[0]


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


def algorithmic_bytes(n_p, s_p, n_c, n_u, n_s, s_s):
    """compulsory payload traffic of one rule iteration"""
    return n_p * (s_p + 16) + 48 * n_c + 40 * n_u + n_s * (s_p + s_s + 16)


def cpu_checker():
    """the reference's own CPU implementation (oracle/_ref) if it was built, else the port"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    if orc.have_reference():
        return orc, orc.Oracle(orc.REF_SO)
    if not os.path.exists(orc.PORT_SO):
        orc.build()
    return orc, orc.Oracle(orc.PORT_SO)


def cpu_reference_rate(sample_parents, seed, steps=1, warmup=0):
    """children/s of the CPU implementation on a bounded sample: (rate, oracle, N_c, mean seconds per step).
    The state is loaded once (the reference's append() re-initialises an index array on every call,
    quids.hpp:279-306, so building a state is quadratic) and only quids::simulate is timed."""
    orc, o = cpu_checker()
    sizes, mags, data = make_parents(sample_parents, seed)
    a, b = orc.LoadedState(o, orc.Packed(sizes, mags, data)), orc.LoadedState(o)
    times, nc = [], 0
    for i in range(warmup + steps):
        nc, nu, secs = a.simulate_into(b, orc.RULE_ERASE_CREATE, [THETA, 0, 0], sample_parents, TOLERANCE)
        if i >= warmup:
            times.append(secs)
    a.close()
    b.close()
    mean = sum(times) / len(times)
    return nc / mean, o, nc, mean


CPU_SAMPLE_PARENTS = 1000000  # 1.3e8 children: a few seconds of CPU work per step, ~11 GB of host memory


def run_reference(args):
    """--impl reference: the reference's CPU path, all host threads, same metric and config"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    parents = args.cpu_sample_parents
    rate, o, nc, secs = cpu_reference_rate(parents, 0, steps=args.steps, warmup=args.warmup)
    sample = f"{parents} random 12-node parents ({nc} children) per step, erase_create(pi/4), max_num_object={parents}"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": o.num_threads, "kind": o.kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, parents_per_gpu=None):
    return {"workload": "QCGD erase_create(theta=pi/4) on random density-1/2 12-node graphs (244 B objects), max_num_object = parents, "
                        "simple truncation, tolerance 1e-18 (BASELINE.json configs[3]; configs[4] when hash-sharded over several GPUs)",
            "parents_per_gpu": parents_per_gpu, "n_node": N_NODE, "rule": "erase_create", "theta": THETA,
            "l2": "inputs larger than L2 (2.6 GB state, GB-scale interference table rewritten every step)",
            "parallelism": "1 GPU" if n_gpus == 1 else f"hash-sharded interference over {n_gpus} GPUs (NCCL all-to-allv)"}


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

    sizes, mags, data = make_parents(parents, seed=rank)
    a, b, sym = qb.Iteration(ctx), qb.Iteration(ctx), qb.SymbolicIteration(ctx)
    a.upload_packed(sizes, mags, data)
    rule = qb.Rule("erase_create", THETA, 0.0, 0.0)
    comm = None
    if world > 1:
        comm = qb.Communicator.from_torch(ctx, dist)
    k_total = parents * world

    def step():
        if comm is None:
            qb.simulate(a, rule, b, sym, parents)
        else:
            qb.mpi_simulate(a, rule, b, sym, comm, k_total)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    launches0 = ctx.launch_count
    sampler = ClockSampler(local)
    phase_sum = {}
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
        for name, ms in sym.phase_ms.items():
            phase_sum[name] = phase_sum.get(name, 0.0) + ms
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    elapsed_ms = e0.elapsed_time(e1)
    n_c, n_u, n_s = sym.num_object, sym.num_object_after_interferences, b.num_object
    if dist is not None:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        c = torch.tensor([n_c, n_u, n_s], device="cuda", dtype=torch.int64)
        dist.all_reduce(c)
        n_c_total = int(c[0].item())
    else:
        n_c_total = n_c
    ms_per_step = elapsed_ms / args.steps
    value = n_c_total / (ms_per_step / 1e3)

    # ---- end to end through the C ABI with HOST buffers: every step uploads its input state from pinned host memory
    # and downloads its result state; the transfers run on the library's copy streams and are double-buffered, so
    # the upload of step i+1 and the download of step i-1 overlap the rule iteration of step i (qb_iter_*_async) -----
    objects, begin, size, mag = a.download()
    h2d = objects.nbytes + begin.nbytes + size.nbytes + mag.nbytes
    pinned = [torch.from_numpy(x.copy()).pin_memory() for x in (objects, begin, size, mag.reshape(-1))]
    host = [p.numpy() for p in pinned]
    host[1], host[2] = host[1].view(np.uint64), host[2].view(np.uint32)
    ins, outs = [qb.Iteration(ctx), qb.Iteration(ctx)], [qb.Iteration(ctx), qb.Iteration(ctx)]

    def pinned_result(cap_n, cap_b):
        return [torch.empty(cap_b, dtype=torch.uint8).pin_memory().numpy(), torch.empty(cap_n + 1, dtype=torch.int64).pin_memory().numpy().view(np.uint64),
                torch.empty(cap_n, dtype=torch.int32).pin_memory().numpy().view(np.uint32), torch.empty(cap_n * 2, dtype=torch.float64).pin_memory().numpy()]

    # on several GPUs the share of the survivors a rank materialises varies from step to step (whichever rank's child
    # created the owner's slot), hence the head room of the pinned result buffers
    n1, nb1, _ = b._counts_noflush()
    out_host = [pinned_result(int(n1 * 1.5) + 4096, int(nb1 * 1.5) + 4096) for _ in range(2)]
    d2h_total = 0

    def e2e_run(steps):
        nonlocal d2h_total
        d2h_total = 0
        ins[0].upload_async(*host)
        for i in range(steps):
            cur = i % 2
            if i + 1 < steps:
                ins[1 - cur].upload_async(*host)  # overlaps this step's rule iteration
            if comm is None:
                qb.simulate(ins[cur], rule, outs[cur], sym, parents)
            else:
                qb.mpi_simulate(ins[cur], rule, outs[cur], sym, comm, k_total)
            n2, nb2, _ = outs[cur]._counts_noflush()
            if nb2 > out_host[cur][0].nbytes or n2 + 1 > out_host[cur][1].shape[0]:
                outs[cur].wait()
                out_host[cur] = pinned_result(int(n2 * 1.5) + 4096, int(nb2 * 1.5) + 4096)
            outs[cur].download_async(*out_host[cur])  # overlaps the next step
            d2h_total += nb2 + 8 * (n2 + 1) + 4 * n2 + 16 * n2
        for it in ins + outs:
            it.wait()

    e2e_steps = max(2, args.steps)
    e2e_run(2)  # untimed: allocations
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    d2h = d2h_total // e2e_steps
    # the last result that reached the host is the state the device holds (spot check of the transfer path)
    last = (e2e_steps - 1) % 2
    n2 = outs[last]._counts_noflush()[0]
    check = outs[last].download()
    assert np.array_equal(check[3].reshape(-1)[:64], out_host[last][3][:64]) and np.array_equal(check[2][:64], out_host[last][2][:64]), "e2e: host copy differs"
    del check
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = n_c_total / (e2e_ms / 1e3)

    # ---- the 1e8-object QCGD configuration on ONE GPU (north_star's target line; BASELINE.json configs[4] without the
    # sharding): same generator, ten chunks of 1e7 parents appended in HBM, max_num_object = parents.  N = 1 only. ----
    large = None
    if world == 1 and args.large_parents > parents:
        del ins, outs, out_host, pinned, host
        big, big_next, chunk_state = qb.Iteration(ctx), qb.Iteration(ctx), qb.Iteration(ctx)
        done = 0
        while done < args.large_parents:
            n_chunk = min(parents, args.large_parents - done)
            cs, cm, cd = make_parents(n_chunk, seed=1000 + done // parents)
            cm[:, 0] = qcgd_magnitude(args.large_parents)
            chunk_state.upload_packed(cs, cm, cd)
            big.append_state(chunk_state)
            done += n_chunk
        del chunk_state
        for _ in range(2):
            qb.simulate(big, rule, big_next, sym, args.large_parents)
        barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(stream)
        large_steps = 3
        large_phase = {}
        for _ in range(large_steps):
            qb.simulate(big, rule, big_next, sym, args.large_parents)
            for name, ms in sym.phase_ms.items():
                large_phase[name] = large_phase.get(name, 0.0) + ms / large_steps
        l1.record(stream)
        barrier()
        large_ms = l0.elapsed_time(l1) / large_steps
        lc, lu, ls = sym.num_object, sym.num_object_after_interferences, big_next.num_object
        free_b, total_b = torch.cuda.mem_get_info()
        lbytes = algorithmic_bytes(args.large_parents, 248, lc, lu, ls, 248)
        peak_l, _ = measured_peak_gbs()
        large = {"workload": f"the same rule and generator on {args.large_parents} parents on one GPU, max_num_object = parents", "value": lc / (large_ms / 1e3),
                 "unit": UNIT, "ms_per_step": large_ms, "steps": large_steps, "counts": {"N_p": args.large_parents, "N_c": lc, "N_u": lu, "N_s": ls},
                 "whole_iteration": {"algorithmic_bytes": lbytes, "achieved": lbytes / (large_ms / 1e3) / 1e9, "frac": lbytes / (large_ms / 1e3) / 1e9 / peak_l},
                 "phase_ms": large_phase, "hbm_in_use_gb": (total_b - free_b) / 1e9}
        del big, big_next

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (symbolic_kernel: children -> (hash, magnitude) -> table) --------
    peak, peak_src = measured_peak_gbs()
    s_p = 4 + 20 * N_NODE + 4  # 244 B padded to 248
    phase_ms = {k: v / args.steps for k, v in phase_sum.items()}
    dominant = max(phase_ms, key=phase_ms.get)
    sym_ms = phase_ms["symbolic"]
    sym_bytes = parents * (s_p + 16) + 48 * n_c  # parents read once + 48 B per child (hash 8 + mag 16, written once, read once)
    achieved = sym_bytes / (sym_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("parents") == parents and tj.get("kernel") == "symbolic_kernel":
            traffic = tj.get("dram_bytes_per_launch")
    total_bytes = algorithmic_bytes(parents, s_p, n_c, n_u, n_s, s_p)
    roofline = {"bound": "hbm", "kernel": "symbolic_items_kernel<erase_create> (child generation in sorted order, on-chip family accumulation, interference-table insert)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sym_bytes, "kernel_ms": sym_ms, "kernel_share_of_step": sym_ms / ms_per_step,
                # what the kernel really moves: the (hash, magnitude) pairs of the algorithmic count are merged on chip and never
                # written, so `frac` can pass 1; this is the DRAM traffic of the ncu capture over the live kernel time
                "traffic_gbs": (traffic / (sym_ms / 1e3) / 1e9) if traffic else None,
                "traffic_frac": (traffic / (sym_ms / 1e3) / 1e9 / peak) if traffic else None,
                "dominant_phase": dominant, "phase_ms": phase_ms,
                "whole_iteration": {"algorithmic_bytes": total_bytes, "achieved": total_bytes / (ms_per_step / 1e3) / 1e9,
                                    "frac": total_bytes / (ms_per_step / 1e3) / 1e9 / peak}}

    # ---- CPU baseline beside it (bounded sample, rank 0, N = 1 only) ----------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        rate, o, nc_cpu, secs = cpu_reference_rate(args.cpu_sample_parents, seed=0, steps=3, warmup=1)
        cpu = {"value": rate, "unit": UNIT, "cores": o.num_threads, "kind": o.kind,
               "sample": f"{args.cpu_sample_parents} parents of the same generator ({nc_cpu} children, {secs:.2f} s per step, 3 steps after 1 warm-up), "
                         f"erase_create(pi/4), max_num_object={args.cpu_sample_parents}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, parents),
            "counts": {"N_p": parents * world, "N_c": n_c_total, "N_u_rank0": n_u, "N_s_rank0": n_s},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "how": "per step: qb_iter_upload_async of the input state from pinned host memory, qb_simulate, qb_iter_download_async of the result "
                           "state into pinned host memory; double-buffered on the library's copy streams, wall clock over all steps incl. the first upload and the last download"},
            "gpu_launches": int(launches)}
    if large is not None:
        line["large"] = large
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--parents", type=int, default=10**7, help="parents per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-sample-parents", type=int, default=CPU_SAMPLE_PARENTS, help="parents of the bounded CPU sample (cpu_baseline and --impl reference)")
    ap.add_argument("--large-parents", type=int, default=10**8, help="N = 1 only: also time this many parents on the one GPU (0 = skip)")
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
