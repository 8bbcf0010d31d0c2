"""CPU, world_size 2, gloo: the N>1 host logic of the seed sweep (sharding, per-seed isolation, final gather).  The hot
path has no collective; the only communication is the gather of final latents, exercised here with a stand-in generator."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from guided_attention_b200 import sweep


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_generate(seed):
    if seed == 31:
        raise RuntimeError("boom")          # a failing seed must not take the others down
    g = torch.Generator("cpu").manual_seed(seed)
    return torch.randn(4, 8, 8, generator=g)


def _worker(rank, world, port, seeds, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = sweep.init_distributed("gloo")
    assert (r, w) == (rank, world)
    full, local, failures = sweep.run_seed_sweep(_fake_generate, seeds, rank, world)
    torch.save({"full": full, "local": local, "failures": failures}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_seeds_round_robin():
    seeds = list(range(28, 92))
    parts = [sweep.shard_seeds(seeds, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == seeds and all(len(p) == 8 for p in parts)
    assert parts[3][:3] == [31, 39, 47]
    assert sweep.shard_seeds(seeds, 0, 1) == seeds


def test_two_rank_sweep_matches_single_process(tmp_path):
    seeds = list(range(28, 35))          # 7 seeds: uneven split, one failing seed
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, seeds, str(tmp_path)), nprocs=world, join=True)
    single_full, _, single_fail = sweep.run_seed_sweep(_fake_generate, seeds, 0, 1)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    for o in outs:                       # every rank holds the full result, in seed order, equal to the 1-process run
        assert torch.equal(o["full"], single_full)
    assert [s for s, _ in single_fail] == [31]
    assert sorted(s for o in outs for s, _ in o["failures"]) == [31]
    assert torch.equal(single_full[3], torch.zeros(4, 8, 8))       # the failed seed's slot
    assert outs[0]["local"].shape[0] == 4 and outs[1]["local"].shape[0] == 3
