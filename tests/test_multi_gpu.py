"""2-GPU test: the same 8 images through the head on 1 rank and sharded over 2 ranks (NCCL, one process per
GPU, openpsg_b200.sharding.run_sharded) give IDENTICAL selected-pair lists and existence masks — the partition over images has
no effect on any image's result (no collective on the data path; PatchEmbed's split-K reduction is deterministic)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
NUM_IMAGES = 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _image_result(head, dev, i):
    from openpsg_b200 import synth
    head(synth.inputs_to(synth.make_image_inputs(synth.WORKLOADS["cfg1"], i), dev), is_generation=False)
    o = head.last_output
    return {"image": i, "topk": o.topk.cpu().tolist(), "mask": o.exist_mask.cpu().tolist(), "logits": o.logits.cpu().tolist()}


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from openpsg_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        head = synth.build_synthetic_head(max_object_num=8, device=dev)
        res = sharding.run_sharded(lambda i: _image_result(head, dev, i), NUM_IMAGES)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_give_the_single_rank_results():
    import torch.multiprocessing as mp
    from openpsg_b200 import synth
    dev = torch.device("cuda", 0)
    head = synth.build_synthetic_head(max_object_num=8, device=dev)
    single = [_image_result(head, dev, i) for i in range(NUM_IMAGES)]
    del head
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, res in got:
        assert [r["image"] for r in res] == list(range(NUM_IMAGES))
        for a, b in zip(res, single):
            assert a["topk"] == b["topk"] and a["mask"] == b["mask"], f"rank {rank} image {a['image']}"
            assert a["logits"] == b["logits"], "bit-identical logits on every rank"
