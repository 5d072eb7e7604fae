#!/usr/bin/env python
"""Development aid: one eager (no CUDA graph) cfg3-shaped generate() with few new tokens, for `ncu` launch lists."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from transformers import OPTConfig, OPTForCausalLM
from openpsg_b200 import synth
from openpsg_b200.llm import build_llm_engine

dev = torch.device("cuda:0")
n_new = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # > 1: the stacked batch of head.forward_batch (n_img x 100 sequences)
with torch.device(dev):
    lm = OPTForCausalLM(OPTConfig(**synth.OPT_2P7B)).eval()
    proj = torch.nn.Linear(768, 2560)
eng = build_llm_engine(lm, proj, dev, use_cuda_graphs=False)
del lm
k, T = 100 * n_img, 17
g = torch.Generator().manual_seed(5)
hidden = torch.randn((1600 * 33, 768), generator=g).to(torch.bfloat16).to(dev)
sel = torch.cat([torch.randperm(1600, generator=g)[:100] for _ in range(n_img)]).to(torch.int32).to(dev)
ids = torch.randint(4, 50272, (k, T), generator=g).to(torch.int32).to(dev)
mask = torch.ones((k, T), dtype=torch.int32, device=dev)
eng.generate(hidden, sel, ids, mask, max_new_tokens=2)      # warm-up (first-use attribute calls)
torch.cuda.synchronize()
out = eng.generate(hidden, sel, ids, mask, max_new_tokens=n_new)
torch.cuda.synchronize()
print("tokens", out.tokens.shape, int(out.tokens.long().sum()))
