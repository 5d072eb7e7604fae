#!/usr/bin/env python
"""Host-side breakdown of one cfg3 forward_batch step (8 images, grouped LLM decode): where the end-to-end time beyond the
device work goes.  Wraps the head's stages with wall-clock timers + stream synchronisations (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from openpsg_b200 import synth

dev = torch.device("cuda:0")
head = synth.build_synthetic_head(llm=synth.OPT_2P7B, max_object_num=80, topk_pairs=100, max_new_tokens=32, device=dev, llm_on_device=True)
head.repack(dev)
wl = synth.WORKLOADS["cfg3"]
host = []
for i in range(8):
    inp = synth.make_image_inputs(wl, i)
    inp["mask_features"] = inp["mask_features"].pin_memory()
    inp["object_info"][0]["pan_results"] = inp["object_info"][0]["pan_results"].to(torch.int32).pin_memory()
    host.append(inp)
for _ in range(3):
    head.forward_batch(host)
torch.cuda.synchronize()
T = {}
def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3; return r
    return w
head._run_queries = timed("run_queries (enqueue, 8 images)", head._run_queries)
head._prepare_host = timed("prepare_host", head._prepare_host)
head._to_device = timed("to_device (H2D enqueue)", head._to_device)
head._decode_group = timed("decode_group (total)", head._decode_group)
head._llm_engine.generate_rows = timed("  generate_rows (enqueue)", head._llm_engine.generate_rows)
head._llm_cache.lookup = timed("  llm prompt lookup", head._llm_cache.lookup)
head.llm_tokenizer.batch_decode = timed("  batch_decode", head.llm_tokenizer.batch_decode)
head._parse_relations = timed("  parse_relations", head._parse_relations)
n = 5
t0 = time.perf_counter()
for _ in range(n):
    head.forward_batch(host)
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) * 1e3 / n
print(f"forward_batch(8 cfg3 images): {tot:.1f} ms wall per call")
for k, v in T.items():
    print(f"  {k:36s} {v / n:8.2f} ms")
