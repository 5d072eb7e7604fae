#!/usr/bin/env python
"""Development aid: time sequences of decode-shaped small-M GEMMs inside one CUDA graph.
  python scripts/skinny_seq.py qofF qqqq oooo ffff FFFF qoqo fFfF   (q = qkv, o = out, f = fc1, F = fc2)"""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import ops
DEV = torch.device("cuda:0")
d, ffn, M, L = 2560, 10240, 100, 6
shapes = {"q": (3 * d, d), "o": (d, d), "f": (ffn, d), "F": (d, ffn)}
W = {k: [(torch.randn(s, device=DEV) * 0.02).to(torch.bfloat16) for _ in range(4 * L)] for k, s in shapes.items()}
A = {d: torch.randn((M, d), device=DEV).to(torch.bfloat16), ffn: torch.randn((M, ffn), device=DEV).to(torch.bfloat16)}
B = {k: torch.zeros(s[0], device=DEV) for k, s in shapes.items()}
O = {k: torch.empty((M, s[0]), device=DEV, dtype=torch.bfloat16) for k, s in shapes.items()}
for pat in sys.argv[1:]:
    for name, gemm in (("skinny", ops.gemm_small_m), ("tiled", ops.gemm)):
        def run():
            cnt = {k: 0 for k in shapes}
            for _ in range(L):
                for c in pat:
                    w = W[c][cnt[c] % len(W[c])]; cnt[c] += 1
                    gemm(A[shapes[c][1]], w, B[c], out=O[c])
        run(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        g.replay(); torch.cuda.synchronize()
        ms = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / (L * len(pat)))
        ms.sort()
        print(json.dumps({"pattern": pat, "kernel": name, "us_per_call": round(ms[5] * 1e3, 2)}), flush=True)
