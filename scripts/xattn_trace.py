#!/usr/bin/env python
"""Development aid: per-unit timeline of CTA 0 of xattn_pairs_kernel (clock64 stamps through the debug hook)."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import _lib, ops

N, L = (int(sys.argv[1]) if len(sys.argv) > 1 else 80), 256
B = N * N
g = torch.Generator().manual_seed(1)
q = torch.randn((B * 33, 768), generator=g).to(torch.bfloat16).cuda()
k = torch.randn((L, 768), generator=g).to(torch.bfloat16).cuda()
vt = torch.randn((768, L), generator=g).to(torch.bfloat16).cuda()
bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (N, 8), dtype=torch.int64, generator=g).to(torch.int32).cuda()
if len(sys.argv) > 2 and sys.argv[2] == "masks":      # panoptic masks of the synthetic image, keys in object order
    from openpsg_b200 import synth
    wl = synth.WORKLOADS["cfg2" if N == 40 else "cfg5"]
    inp = synth.make_image_inputs(wl, 0)
    pan = inp["object_info"][0]["pan_results"].to(torch.int32).cuda()
    ids = torch.tensor([int(i) for i in inp["object_info"][0]["object_id_list"]], dtype=torch.int32, device="cuda")
    bits = ops.pair_mask_bits(pan, (wl.height, wl.width), (wl.height, wl.width), (16, 16), ids)
    _, bits = ops.token_order(bits, L)
tiles = ops.xattn_bias_tiles(bits, N, B, 33, L)
for _ in range(3):
    ops.xattn_pairs(q, k, vt, bits, N, B, 33, L, 12, 64, bias_tiles=tiles)
trace = torch.zeros(8 * 256 + 4 * 160, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.opsg_debug_xattn_trace.argtypes = [ctypes.c_void_p]
lib.opsg_debug_xattn_trace.restype = None
lib.opsg_debug_xattn_trace(trace.data_ptr())
ops.xattn_pairs(q, k, vt, bits, N, B, 33, L, 12, 64, bias_tiles=tiles)
torch.cuda.synchronize()
lib.opsg_debug_xattn_trace(None)
full = trace.cpu()
t = full[:2048].view(-1, 8)
n = int((t[:, 0] != 0).sum())
t0 = int(t[0, 0])
print("unit  qk_issue pv_issue | s_full pass1_end pass2_end o_full epi_end   (cycles since first QK issue; buffers alternate)")
for i in range(min(n, 64)):
    r = [int(x) - t0 if x else -1 for x in t[i, :7]]
    print(f"{i:4d}  {r[0]:8d} {r[1]:8d} | {r[2]:7d} {r[3]:8d} {r[4]:8d} {r[5]:7d} {r[6]:8d}   pass1={r[3]-r[2]} pass2={r[4]-r[3]} p_ready->pv={r[1]-r[4]} pv->o_full={r[5]-r[1]} epi={r[6]-r[5]}")
if n > 12:
    print("steady-state cycles/unit:", (int(t[n - 2, 0]) - int(t[8, 0])) / (n - 10))

w = full[2048:].view(-1, 4)
w = w[w[:, 0] != 0]
if len(w):
    t_min = int(w[:, 0].min())
    rel = (w - t_min).float() / 1e3
    print(f"per-CTA wall clock (us since the first CTA started), {len(w)} CTAs:")
    print(f"  entry      : min {rel[:,0].min():.2f} max {rel[:,0].max():.2f}")
    print(f"  setup done : min {rel[:,1].min():.2f} max {rel[:,1].max():.2f}")
    print(f"  first QK   : min {rel[:,2].min():.2f} median {rel[:,2].median():.2f} max {rel[:,2].max():.2f}")
    print(f"  exit       : min {rel[:,3].min():.2f} median {rel[:,3].median():.2f} max {rel[:,3].max():.2f}")
    d = rel[:, 3] - rel[:, 2]
    print(f"  first QK -> exit: min {d.min():.2f} median {d.median():.2f} max {d.max():.2f}")
