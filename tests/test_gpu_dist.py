"""Distributed path on real GPUs (needs >= 2; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def test_two_gpu_parity_with_single_process_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(HERE, "dist_worker.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-4000:]
    assert "ok ties across ranks" in out.stdout
    for what in ("distribute_objects", "gather_objects", "send_objects / receive_objects", "equalize by objects", "equalize by children",
                 "rule iteration after distribute_objects", "failure on one rank stops every rank", "communicator usable after agreed failures",
                 "automatic budget on the distributed path"):
        assert f"ok {what}" in out.stdout, out.stdout[-4000:]


def test_mpi_example_driver_round_trip():
    """examples/mpi_test.cpp (the scenario of the reference's MPI example through the drop-in headers): gates on one rank,
    distribute_objects, global statistics, the gates undone with quids::mpi::simulate, gather_objects; the gathered state
    must be the initial one (SURVEY 8c fixture 3)."""
    import torch
    from test_examples import EXAMPLES, build_examples, normalise
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    build_examples()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29519",
           "--no-python", os.path.join(EXAMPLES, "mpi_test.out")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-4000:]
    text = out.stdout

    def section(title):
        body = text.split(f"\n{title}:\n", 1)[1]
        lines = []
        for line in body.splitlines():
            if line.startswith("\t") or line.startswith("    node "):
                lines.append(line)
            elif lines:
                break
        per_node, node = {}, None
        for line in lines:
            if line.startswith("    node "):
                node = int(line.split()[1].split("/")[0])
                per_node[node] = []
            else:
                per_node[node].append(line)
        return {k: normalise("\n".join(v)) for k, v in per_node.items()}

    first, last = section("initial state"), section("gathered all objects")
    assert len([l for l in first[1] if l]) == 3 and not [l for l in first[0] if l]
    assert last == first, (first, last)
    spread = section("distributed all objects")
    assert [len([l for l in spread[r] if l]) for r in (0, 1)] == [12, 12]  # 3 kets x 2^3 after three Hadamards, halved
    assert "the total number of objects is 24" in text and "the average size is" in text
    assert "P=1" in text
