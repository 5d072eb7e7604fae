#!/usr/bin/env python
"""Development aid: epilogue timeline of CTA 0 of the CTA-pair GEMM (clock64 stamps through the debug hook)."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import _lib, ops

which = sys.argv[1] if len(sys.argv) > 1 else "plain"
M, N, K = 78400, 768, 768
a = torch.randn((M, K), device="cuda").to(torch.bfloat16)
w = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
res = torch.randn((M, N), device="cuda").to(torch.bfloat16) if which in ("res",) else None
act = 1 if which == "gelu" else 0
out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(a, w, bias, residual=res, act=act, out=out)
trace = torch.zeros(24 * 5 * 8, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.opsg_debug_gemm2_trace.argtypes = [ctypes.c_void_p]
lib.opsg_debug_gemm2_trace.restype = None
lib.opsg_debug_gemm2_trace(trace.data_ptr())
ops.gemm(a, w, bias, residual=res, act=act, out=out)
torch.cuda.synchronize()
lib.opsg_debug_gemm2_trace(None)
t = trace.cpu().view(24, 5, 8)
t0 = int(t[0, 4, 0])
print(f"case {which}: per tile: wait tmem_full | per slab: [ldtm_done bias_done slab_ready math_done sts_done arrived] (cycles from slab start)")
for i in range(2, 10):
    print(f"tile {i}: start {int(t[i,4,0])-t0:8d} tmem_full {int(t[i,4,1])-t0:8d} (tile period {int(t[i,4,0])-int(t[i-1,4,0])})")
    for sl in range(4):
        r = [int(x) - int(t[i, sl, 0]) for x in t[i, sl]]
        print("      slab", sl, r[1:7])
