#!/usr/bin/env python
"""Decode attention (K10b) alone at the stacked-batch size: nseq sequences x 32 heads x head_dim 80, context `ctx`.
  python scripts/decode_attn_probe.py [nseq] [ctx]     prints us per launch and achieved GB/s (CUDA events, L2 flushed by size)"""
import sys
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from openpsg_b200 import ops

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 800
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 65
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 32      # heads = 1 with nseq x 32: the same items, each contiguous in the caches
hd, max_ctx = 80, 81
d = heads * hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn((nseq, 3 * d), device="cuda", generator=g).to(torch.bfloat16)
layers = 4                                  # 4 x (K + V) caches of 265 MB each: nothing survives in L2 between launches
kc = [torch.randn((nseq, max_ctx, d), device="cuda", generator=g).to(torch.bfloat16) for _ in range(layers)]
vc = [torch.randn((nseq, max_ctx, d), device="cuda", generator=g).to(torch.bfloat16) for _ in range(layers)]
kmask = torch.ones((nseq, max_ctx), dtype=torch.uint8, device="cuda")
out = torch.empty((nseq, d), dtype=torch.bfloat16, device="cuda")
for i in range(layers):
    ops.llm_attn_append(qkv, kc[i], vc[i], kmask, nseq, ctx - 1, heads, hd, hd ** -0.5, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    for i in range(layers):
        ops.llm_attn_append(qkv, kc[i], vc[i], kmask, nseq, ctx - 1, heads, hd, hd ** -0.5, out)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * layers)
nbytes = 2.0 * nseq * ctx * d * 2
print(f"decode attention nseq={nseq} ctx={ctx}: {us:.1f} us per launch, {nbytes / us / 1e3:.0f} GB/s of K/V")
