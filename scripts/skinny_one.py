#!/usr/bin/env python
"""Development aid: a few decode-shaped small-M GEMM launches over different weight matrices (ncu target)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import ops

CASES = {"qkv": (100, 7680, 2560), "out": (100, 2560, 2560), "fc1": (100, 10240, 2560), "fc2": (100, 2560, 10240)}
M, N, K = CASES[sys.argv[1]]
a = torch.randn((M, K), device="cuda").to(torch.bfloat16)
ws = [(torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16) for _ in range(6)]
bias = torch.randn(N, device="cuda")
out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
for w in ws:
    ops.gemm_small_m(a, w, bias, out=out)
torch.cuda.synchronize()
