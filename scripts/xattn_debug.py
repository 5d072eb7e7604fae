#!/usr/bin/env python
"""Development aid: run xattn_pairs repeatedly on fixed inputs, report which (tile, head) units differ from the reference."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import ops
from oracle import restated
from tests.test_kernels_gpu import _xattn_ref, _rand_bf16

L, N, B = 256, 40, 1600
g = torch.Generator().manual_seed(L * 100 + N)
nq, d = 33, 768
q = _rand_bf16((B * nq, d), g); k = _rand_bf16((L, d), g); v = _rand_bf16((L, d), g)
masks = torch.rand(N, L, generator=g) < 0.15
masks[N - 1] = False
bits = torch.from_numpy(restated.pack_mask_bits(masks.numpy()).view(np.int32)).cuda()
vt = v.t().contiguous().cuda()
qc, kc = q.cuda(), k.cuda()
ref = _xattn_ref(q, k, v, masks, N, None, nq)
tiles = ops.xattn_bias_tiles(bits, N, B, nq, L)
m_tiles = (B * nq + 127) // 128
per = -(-m_tiles * 12 // 148)
for it in range(8):
    out = ops.xattn_pairs(qc, kc, vt, bits, N, B, nq, L, 12, 64, bias_tiles=tiles).float().cpu()
    err = (out - ref).abs()
    bad = err > 2e-2
    print(f"run {it}: max err {err.max():.4f} bad elements {int(bad.sum())}")
    if bad.any():
        rows, cols = torch.nonzero(bad, as_tuple=True)
        units = sorted(set((int(r) // 128, int(c) // 64) for r, c in zip(rows.tolist(), cols.tolist())))
        desc = []
        for mt, h in units[:12]:
            u = h * m_tiles + mt
            cta, i = u // per, u % per
            rr = rows[(rows // 128 == mt) & (cols // 64 == h)]
            cc = cols[(rows // 128 == mt) & (cols // 64 == h)]
            desc.append(f"(mt={mt},h={h}: cta {cta} unit# {i}/{per}, rows {int(rr.min())%128}-{int(rr.max())%128}, cols {int(cc.min())%64}-{int(cc.max())%64}, n={len(rr)})")
        print("   bad units:", len(units), *desc)
