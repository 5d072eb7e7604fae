#!/usr/bin/env python
"""Development aid: cfg5-shaped images (80 objects, 6400 pair queries) through the relation-query path (a2-a8), timed
with CUDA events over a stream of resident images: ms per image and ordered pairs / s."""
import json
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import synth
from tests.helpers import build_product_head

wl = synth.WORKLOADS["cfg5"]
head = build_product_head(max_object_num=wl.num_objects, topk_pairs=100, device="cuda:0")
head.repack("cuda:0")
imgs = [synth.inputs_to(synth.make_image_inputs(wl, i), "cuda:0") for i in range(2)]
head.forward_batch(imgs * 2)
torch.cuda.synchronize()
n = 8
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
head.forward_batch(imgs * (n // 2))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"workload": "cfg5 image: 80 objects, 6400 pair queries (6320 ordered pairs), 256 image tokens, a2-a8",
                  "ms_per_image": ms, "pairs_per_s": wl.ordered_pairs / (ms * 1e-3), "topk": head.last_output.topk.numel()}))
