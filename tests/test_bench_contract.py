"""The bench.py contract that can be checked without a GPU: the reference arm prints exactly one JSON line on stdout
with the keys the driver reads, and our arm fails loudly (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-parents", "20000"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "children_per_sec_per_rule_iteration" and line["unit"] == "children/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and "20000" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "children/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and out.stdout.strip() == "" and "no CUDA device" in out.stderr


def test_pinned_buffers_are_allocated_near_the_gpu(monkeypatch):
    """bench.near_gpu: the calling thread is bound to the CPUs NVML reports as local to the device while the pinned buffers of
    the e2e leg are allocated, and gets its affinity back; without NVML nothing changes"""
    import types
    sys.path.insert(0, ROOT)
    import bench
    allowed = os.sched_getaffinity(0)
    if len(allowed) < 2:
        pytest.skip("one CPU: nothing to choose from")
    local = set(sorted(allowed)[: len(allowed) // 2])
    fake = types.ModuleType("pynvml")
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByIndex = lambda i: i
    seen = {}

    def affinity(handle, n_words):
        seen["words"] = n_words
        return [sum(1 << (c - 64 * w) for c in local if 64 * w <= c < 64 * w + 64) for w in range(n_words)]

    fake.nvmlDeviceGetCpuAffinity = affinity
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    with bench.near_gpu(0) as n:
        assert os.sched_getaffinity(0) == local
        assert "local to the GPU" in n.placement
    assert os.sched_getaffinity(0) == allowed
    assert seen["words"] * 64 > max(allowed)
    # a cpuset that excludes the GPU's CPUs, or a failing NVML: nothing changes
    fake.nvmlDeviceGetCpuAffinity = lambda handle, n_words: [0] * n_words
    with bench.near_gpu(0) as n:
        assert os.sched_getaffinity(0) == allowed and n.placement == "default placement"

    def broken(handle, n_words):
        raise RuntimeError("no NVML")

    fake.nvmlDeviceGetCpuAffinity = broken
    with bench.near_gpu(0) as n:
        assert os.sched_getaffinity(0) == allowed and n.placement == "default placement"
    assert os.sched_getaffinity(0) == allowed
