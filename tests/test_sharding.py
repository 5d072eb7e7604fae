"""Multi-process (gloo, world size 2, CPU) tests of the image-sharding host logic used by bench.py / N>1 runs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openpsg_b200 import sharding, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _image_record(i):
    """Stand-in for head(inputs): a deterministic, image-dependent record of the same shape as the real result
    (selected pair indices + a checksum of the inputs), computed on CPU."""
    inp = synth.make_image_inputs(synth.WORKLOADS["cfg1"], i)
    pan = inp["object_info"][0]["pan_results"]
    feat = inp["mask_features"]
    return {"image": i, "pairs": [int(x) for x in torch.topk(feat.flatten()[:64], 5).indices],
            "checksum": int(pan.long().sum()) ^ int(feat.double().sum().item() * 1000)}


def _worker(rank, world, port, num_items, use_costs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        costs = [(i % 3 + 1) ** 2 for i in range(num_items)] if use_costs else None
        res = sharding.run_sharded(_image_record, num_items, costs)
        # max-over-ranks timing reduction as bench.py does it
        t = torch.tensor([10.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, res, t.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("use_costs", [False, True])
def test_two_rank_sharding_equals_single_process(use_costs):
    num_items, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_items, use_costs, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [_image_record(i) for i in range(num_items)]
    for rank, res, tmax in got:
        assert res == expect, f"rank {rank} gathered a different result list"
        assert tmax == 10.0 + world - 1


def test_shard_indices_partition():
    for n in (0, 1, 7, 32):
        for world in (1, 2, 4, 8):
            parts = [sharding.shard_indices(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_indices(4, 2, 2)


def test_lpt_assign_balances_quadratic_costs():
    costs = [80 ** 2, 40 ** 2, 40 ** 2, 40 ** 2, 40 ** 2, 8 ** 2]
    parts = sharding.lpt_assign(costs, 2)
    assert sorted(sum(parts, [])) == list(range(6))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) == 6464 and min(loads) == 6400     # the 80-object image sits alone with the 8-object one
    assert sharding.lpt_assign(costs, 2) == parts        # deterministic


def test_gather_detects_missing_and_duplicate_items():
    with pytest.raises(RuntimeError):
        sharding.gather_by_index({0: "a"}, 2)
    assert sharding.gather_by_index({1: "b", 0: "a"}, 2) == ["a", "b"]
