#!/usr/bin/env python
"""Development aid: cfg3 decode leg (OPT-2.7B, 100 pairs x 32 tokens) as one CUDA graph, A/B over engine switches."""
import json
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import synth
from openpsg_b200.llm import build_llm_engine

dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.device(dev):
    lm = synth.build_causal_lm(synth.OPT_2P7B).eval()
    proj = torch.nn.Linear(768, 2560)
k, T, t_new = 100, 17, 32
g = torch.Generator().manual_seed(5)
hidden = torch.randn((1600 * 33, 768), generator=g).to(torch.bfloat16).to(dev)
sel = torch.randperm(1600, generator=g)[:k].to(torch.int32).to(dev)
ids = torch.randint(4, 50272, (k, T), generator=g).to(torch.int32).to(dev)
mask = torch.ones((k, T), dtype=torch.int32, device=dev)
ref = None
for tag, kw in (("baseline", dict(prefetch=0)), ("prefetch_gaps", dict(prefetch=2))):
    eng = build_llm_engine(lm, proj, dev)
    for a, b in kw.items():
        setattr(eng, a, b)
    for _ in range(3):
        out = eng.generate(hidden, sel, ids, mask, max_new_tokens=t_new)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = eng.generate(hidden, sel, ids, mask, max_new_tokens=t_new)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    toks = out.tokens.clone()
    same = True if ref is None else bool(torch.equal(ref, toks))
    ref = toks if ref is None else ref
    print(json.dumps({"variant": tag, "ms_per_image": ms, "tokens_per_s": k * t_new / ms * 1e3, "same_tokens_as_baseline": same}), flush=True)
    del eng
