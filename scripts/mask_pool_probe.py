#!/usr/bin/env python
"""Development aid: K11 (mask pooling) on the cfg2 / cfg5 images, timed with CUDA events (L2 flushed between runs) and as a
target for `ncu`.  Algorithmic bytes = 67 MB of features + the label map once per channel block."""
import json
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import ops, synth

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name in ("cfg2", "cfg5"):
    wl = synth.WORKLOADS[name]
    inp = synth.make_image_inputs(wl, 0)
    ids = torch.tensor([int(i) for i in inp["object_info"][0]["object_id_list"]], dtype=torch.int32, device=dev)
    pan = inp["object_info"][0]["pan_results"].to(torch.int32).to(dev)
    feat = inp["mask_features"][0].to(dev)
    label, rep = ops.mask_pool_labels(pan, (wl.height, wl.width), (wl.height, wl.width), feat.shape[-2:], ids)
    ops.mask_pool_pairs(feat, label, len(ids), rep=rep)
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.profile_begin()
        ops.mask_pool_pairs(feat, label, len(ids), rep=rep)
        prof = ops.profile_end()
        times.append(prof["mask_pool_pairs"]["ms"])
    times.sort()
    nbytes = 4.0 * feat.numel() + 4.0 * label.numel()
    print(json.dumps({"kernel": "mask_pool_pairs (4 launches)", "workload": name, "ms_median": times[len(times) // 2], "ms_best": times[0],
                      "GBps_median": nbytes / times[len(times) // 2] / 1e6, "algorithmic_MB": nbytes / 1e6}), flush=True)
