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


def _only_even_ok(seed):
    if seed % 2:
        raise RuntimeError("odd seeds fail")     # with world 2 every seed of rank 1 fails
    return _fake_generate(seed)


def _always_fail(seed):
    raise RuntimeError("nope")


def _worker_edge(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sweep.init_distributed("gloo")
    res = {}
    # (a) one rank loses all of its seeds: it must still join the gather (round 1 returned early -> the others hung)
    res["all_fail_rank"] = sweep.run_seed_sweep(_only_even_ok, [28, 29, 30, 33], rank, world)
    # (b) fewer seeds than ranks: rank 1's shard is empty
    res["empty_shard"] = sweep.run_seed_sweep(_fake_generate, [28], rank, world)
    # (c) nobody has a result: None everywhere, no collective left dangling
    res["nobody"] = sweep.run_seed_sweep(_always_fail, [28, 29], rank, world)
    torch.save(res, os.path.join(out_dir, f"e{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_ranks_without_results_still_join_the_gather(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker_edge, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"e{r}.pt")) for r in range(world)]
    for o in outs:
        full, _, _ = o["all_fail_rank"]
        assert full.shape == (4, 4, 8, 8)
        assert torch.equal(full[0], _fake_generate(28)) and torch.equal(full[2], _fake_generate(30))
        assert not full[1].any() and not full[3].any()
        full, _, _ = o["empty_shard"]
        assert full.shape == (1, 4, 8, 8) and torch.equal(full[0], _fake_generate(28))
        full, local, failures = o["nobody"]
        assert full is None and local == [] and len(failures) == 1
    assert outs[1]["all_fail_rank"][1].shape[0] == 2 and len(outs[1]["all_fail_rank"][2]) == 2
    assert outs[1]["empty_shard"][1].shape[0] == 0
