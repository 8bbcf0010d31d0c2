"""BASELINE config 5 (seed sweep sharded one process per GPU) as an EQUALITY test: the per-seed final
latents of a 2-rank torchrun launch of the product entry point (`run.execute`) must be bit-identical to the 1-rank
run -- nothing on the path depends on which process or in which order a seed is generated (no data atomics, CPU RNG
streams per seed, CUDA graphs keyed on content).  With two GPUs the ranks use one GPU each and NCCL; on a 1-GPU box
both ranks share cuda:0 and the final gather goes through gloo."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "workers", "sweep_worker.py")
SEEDS = [28, 29, 30, 31, 32]          # uneven split over 2 ranks: 3 + 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, out, seeds=SEEDS):
    if world == 1:
        cmd = [sys.executable, WORKER]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER]
    cmd += ["--out", out, "--seeds", *map(str, seeds)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return torch.load(out)


@pytest.mark.timeout(3000)
def test_two_rank_sweep_equals_single_rank_bit_for_bit(tmp_path):
    one = _run(1, str(tmp_path / "w1.pt"))
    two = _run(2, str(tmp_path / "w2.pt"))
    assert one["world"] == 1 and two["world"] == 2
    assert one["latents"].shape == (len(SEEDS), 4, 64, 64) and one["latents"].dtype == torch.float16
    assert torch.isfinite(one["latents"].float()).all()
    for n, seed in enumerate(SEEDS):
        assert torch.equal(one["latents"][n], two["latents"][n]), f"seed {seed} differs between 1 and 2 ranks"
    assert not torch.equal(one["latents"][0], one["latents"][1])       # different seeds, different images
