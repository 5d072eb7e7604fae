#!/usr/bin/env python
"""Development aid: cfg3-shaped LLM leg (random-init OPT-2.7B, 100 pairs, 49-token prompt) as CUDA graphs: time per image with
32 new tokens and with 1 (prefill only) -> time per decode step and its share of the HBM roofline (weights + KV per step)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from transformers import OPTConfig, OPTForCausalLM
from openpsg_b200 import synth
from openpsg_b200.llm import build_llm_engine

dev = torch.device("cuda:0")
n_new = int(sys.argv[1]) if len(sys.argv) > 1 else 32
with torch.device(dev):
    lm = OPTForCausalLM(OPTConfig(**synth.OPT_2P7B)).eval()
    proj = torch.nn.Linear(768, 2560)
eng = build_llm_engine(lm, proj, dev, use_cuda_graphs=True)
n_layers, d, ffn, vocab = len(eng.w.layers), eng.w.d, eng.w.ffn, eng.w.vocab
del lm
k, T = 100, 17
g = torch.Generator().manual_seed(5)
hidden = torch.randn((1600 * 33, 768), generator=g).to(torch.bfloat16).to(dev)
sel = torch.randperm(1600, generator=g)[:k].to(torch.int32).to(dev)
ids = torch.randint(4, 50272, (k, T), generator=g).to(torch.int32).to(dev)
mask = torch.ones((k, T), dtype=torch.int32, device=dev)


def timed(new_tokens, reps=5):
    for _ in range(3):      # first sighting eager, second captures, third replays
        out = eng.generate(hidden, sel, ids, mask, max_new_tokens=new_tokens)
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.generate(hidden, sel, ids, mask, max_new_tokens=new_tokens)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2], int(out.tokens.long().sum())


full, chk = timed(n_new)
pre, _ = timed(1)
steps = n_new - 1
step_ms = (full - pre) / steps
w_bytes = 2.0 * (n_layers * (4 * d * d + 2 * d * ffn) + vocab * d)
Tp = 32 + T
kv_bytes = sum(n_layers * k * (Tp + s) * d * 2 * 2.0 for s in range(1, n_new)) / steps
print(f"image {full:.2f} ms  prefill {pre:.2f} ms  decode step {step_ms * 1e3:.1f} us  ({step_ms * 1e3 / n_layers:.1f} us/layer incl. lm_head)"
      f"  decode HBM {(w_bytes + kv_bytes) / step_ms / 1e6:.0f} GB/s  tokens/s {k * n_new / full * 1e3:.0f}  checksum {chk}")
