#!/usr/bin/env python
"""The LLM Linears at the row counts of stacked decode steps (head.forward_batch: M = images x selected pairs), OPT-2.7B shapes:
  python scripts/gemm_medium_m.py [M ...]      us per call and TFLOP/s; operands rotate over 6 weight copies (nothing stays in L2)"""
import sys
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from openpsg_b200 import ops

DEV = "cuda:0"
Ms = [int(a) for a in sys.argv[1:]] or [800]
shapes = [("qkv", 7680, 2560, 0, False), ("out+res", 2560, 2560, 0, True), ("fc1 relu", 10240, 2560, ops.ACT_RELU, False),
          ("fc2+res", 2560, 10240, 0, True), ("lm_head f32", 50272, 2560, 0, False)]
for M in Ms:
    tot = 0.0
    for name, N, K, act, res in shapes:
        copies = 6
        ws = [(torch.randn((N, K), device=DEV) / K ** 0.5).to(torch.bfloat16) for _ in range(copies)]
        a = torch.randn((M, K), device=DEV).to(torch.bfloat16)
        bias = torch.randn(N, device=DEV)
        f32 = name.startswith("lm_head")
        r = torch.randn((M, N), device=DEV).to(torch.bfloat16) if res else None
        out = torch.empty((M, N), device=DEV, dtype=torch.float32 if f32 else torch.bfloat16)
        kw = dict(out=out) if f32 else dict(residual=r, act=act, out=out)
        for i in range(copies):
            (ops.gemm_medium_m if (os.environ.get("OPSG_PROBE_MEDIUM") and not f32) else ops.gemm)(a, ws[i], None if f32 else bias, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            for i in range(copies):
                (ops.gemm_medium_m if (os.environ.get("OPSG_PROBE_MEDIUM") and not f32) else ops.gemm)(a, ws[i], None if f32 else bias, **kw)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * copies)
        if not f32:
            tot += us
        print(f"M={M:5d} {name:12s} N={N:6d} K={K:6d}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s  (weights alone at 6.5 TB/s: {N * K * 2 / 6.5e6:6.1f} us)")
        del ws
    print(f"M={M:5d} layer GEMMs: {tot:.1f} us x 32 layers x 31 steps = {tot * 32 * 31 / 1e3:.1f} ms")
