#!/usr/bin/env python
"""Development aid: decode-step GEMM chain of OPT-2.7B layers at M = 100 (out_proj+res -> LN -> fc1+ReLU -> fc2+res -> LN -> qkv),
16 layers of distinct weights in one CUDA graph: one persistent launch per layer (ops.gemm_chain, norms folded) against one
GEMM / LayerNorm kernel at a time."""
import json
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import ops

DEV = torch.device("cuda:0")
d, ffn, M, L = 2560, 10240, 100, 16
layers = []
for _ in range(L):
    layers.append(dict(
        w_qkv=(torch.randn((3 * d, d), device=DEV) * 0.02).to(torch.bfloat16), b_qkv=torch.zeros(3 * d, device=DEV),
        w_o=(torch.randn((d, d), device=DEV) * 0.02).to(torch.bfloat16), b_o=torch.zeros(d, device=DEV),
        w_fc1=(torch.randn((ffn, d), device=DEV) * 0.02).to(torch.bfloat16), b_fc1=torch.zeros(ffn, device=DEV),
        w_fc2=(torch.randn((d, ffn), device=DEV) * 0.02).to(torch.bfloat16), b_fc2=torch.zeros(d, device=DEV),
        g1=torch.ones(d, device=DEV), be1=torch.zeros(d, device=DEV), g2=torch.ones(d, device=DEV), be2=torch.zeros(d, device=DEV)))
h0 = torch.randn((M, d), device=DEV).to(torch.bfloat16)
ctx0 = torch.randn((M, d), device=DEV).to(torch.bfloat16)
n_st = (d + 63) // 64


def run_chain(phases_per_launch):
    h = h0.clone()
    f = torch.empty((M, ffn), dtype=torch.bfloat16, device=DEV)
    qkv = torch.empty((M, 3 * d), dtype=torch.bfloat16, device=DEV)
    qkv[:, :d] = ctx0
    st1 = torch.empty((M, n_st, 2), dtype=torch.float32, device=DEV)
    st2 = torch.empty_like(st1)
    for lw in layers:
        ph = [ops.chain_phase(qkv[:, :d], lw["w_o"], h, lw["b_o"], residual=h, stats_out=st1),
              ops.chain_phase(h, lw["w_fc1"], f, lw["b_fc1"], act=ops.ACT_RELU, norm=ops.NORM_LAYER, norm_gamma=lw["g2"],
                              norm_beta=lw["be2"], stats_in=st1),
              ops.chain_phase(f, lw["w_fc2"], h, lw["b_fc2"], residual=h, stats_out=st2),
              ops.chain_phase(h, lw["w_qkv"], qkv, lw["b_qkv"], norm=ops.NORM_LAYER, norm_gamma=lw["g1"], norm_beta=lw["be1"],
                              stats_in=st2)]
        for i in range(0, 4, phases_per_launch):
            assert ops.gemm_chain(ph[i:i + phases_per_launch])
    return h


def run_single():
    h = h0.clone()
    qkv = torch.empty((M, 3 * d), dtype=torch.bfloat16, device=DEV)
    qkv[:, :d] = ctx0
    for lw in layers:
        ops.gemm_small_m(qkv[:, :d], lw["w_o"], lw["b_o"], residual=h, out=h)
        x = ops.layernorm(h, lw["g2"], lw["be2"], 1e-5)
        f = ops.gemm_small_m(x, lw["w_fc1"], lw["b_fc1"], act=ops.ACT_RELU)
        ops.gemm_small_m(f, lw["w_fc2"], lw["b_fc2"], residual=h, out=h)
        x = ops.layernorm(h, lw["g1"], lw["be1"], 1e-5)
        ops.gemm_small_m(x, lw["w_qkv"], lw["b_qkv"], out=qkv)
    return h


import os
for name, fn, env in (("chain x4", lambda: run_chain(4), "1"), ("chain x2", lambda: run_chain(2), "1"), ("chain x1", lambda: run_chain(1), "1"),
                      ("one kernel per op (chain kernel)", run_single, "1"), ("one kernel per op (workspace kernels)", run_single, "0")):
    os.environ["OPSG_GEMM_CHAIN"] = env
    fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    graph.replay()
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / L)
    ms.sort()
    print(json.dumps({"variant": name, "us_per_layer_median": round(ms[5] * 1e3, 1)}), flush=True)
