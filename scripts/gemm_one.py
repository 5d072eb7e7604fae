#!/usr/bin/env python
"""Development aid: launch one cfg2-shaped GEMM a few times (target of `ncu -k regex:gemm2 --launch-skip 3 -c 1`)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import ops

CASES = {"qkv": (78400, 2304, 768, False, 0), "res": (78400, 768, 768, True, 0), "gelu": (52800, 3072, 768, False, 1),
         "down": (52800, 768, 3072, True, 0)}
M, N, K, use_res, act = CASES[sys.argv[1]]
a = torch.randn((M, K), device="cuda").to(torch.bfloat16)
w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
res = torch.randn((M, N), device="cuda").to(torch.bfloat16) if use_res else None
out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
for _ in range(5):
    ops.gemm(a, w, bias, residual=res, act=act, out=out)
torch.cuda.synchronize()
