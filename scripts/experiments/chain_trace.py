#!/usr/bin/env python
"""Development aid (trace build: make -C openpsg_b200/csrc trace; OPSG_B200_LIB=.../libopsg_b200_trace.so): per-CTA
%globaltimer stamps of the chained small-M GEMM for one decode-shaped GEMM (argv: N K)."""
import ctypes
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import ops, _lib

N, K = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (7680, 2560)
M = 100
dev = torch.device("cuda:0")
a = torch.randn((M, K), device=dev).to(torch.bfloat16)
ws = [(torch.randn((N, K), device=dev) / K ** 0.5).to(torch.bfloat16) for _ in range(6)]
bias = torch.randn(N, device=dev)
out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
lib = _lib.load()
trace = torch.zeros((148, 32, 16), dtype=torch.int64, device=dev)
for w in ws[:5]:
    ops.gemm_small_m(a, w, bias, out=out)
torch.cuda.synchronize()
lib.opsg_debug_chain_trace.argtypes = [ctypes.c_void_p]
lib.opsg_debug_chain_trace(trace.data_ptr())
ops.gemm_small_m(a, ws[5], bias, out=out)
torch.cuda.synchronize()
lib.opsg_debug_chain_trace(None)
t = trace.cpu().numpy()
t0 = t[:, 31, 0][t[:, 31, 0] > 0].min()
names = {1: "acc ready", 2: "stage free", 3: "store: stage full", 4: "store: done (wait_group)", 5: "store: counted", 6: "store: queued",
         7: "fin: item", 8: "fin: all slices", 9: "fin: rows done", 10: "fin: counted", 13: "fin: data landed", 11: "mma: last issued", 12: "mma: first issued"}
def rel(v):
    return (v - t0) / 1e3 if v > 0 else float("nan")
ends = t[:, 31, 2]
print(f"N={N} K={K}: kernel span {rel(ends.max()):.1f} us (CTA start min/max {rel(t[:,31,0].min()):.1f}/{rel(t[:,31,0].max()):.1f}, "
      f"A in TMEM min/max {rel(t[:,31,1][t[:,31,1]>0].min()):.1f}/{rel(t[:,31,1].max()):.1f}, end min/max {rel(ends[ends>0].min()):.1f}/{rel(ends.max()):.1f})")
for cta in (0, 1, 73, 147):
    print(f"-- CTA {cta}")
    for tile in range(8):
        row = t[cta, tile]
        if not (row[1:13] > 0).any():
            continue
        print(f"  tile {tile}: " + "  ".join(f"{names[k]} {rel(row[k]):.1f}" for k in (12, 11, 1, 2, 3, 4, 5, 6, 7, 8, 13, 9, 10) if row[k] > 0))
