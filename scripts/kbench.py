#!/usr/bin/env python
"""Kernel micro-benchmarks on one B200 (development aid; bench.py is the contract benchmark).

Times the cfg2-shaped launches of the two tensor-core kernels with CUDA events on the launching stream,
L2 flushed between iterations, and prints one JSON line per case:
  python scripts/kbench.py [gemm] [xattn] [--iters 20]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openpsg_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()                       # > L2 (126 MB): evicts operands between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2], ms[0]


def bench_gemm(iters, flush):
    B = 1600
    shapes = [  # name, M, N, K, bias, residual, act
        ("qkv", B * 49, 2304, 768, True, False, 0),
        ("self_out", B * 49, 768, 768, True, True, 0),
        ("cross_q", B * 33, 768, 768, True, False, 0),
        ("ffn_up_q_gelu", B * 33, 3072, 768, True, False, 1),
        ("ffn_down_q", B * 33, 768, 3072, True, True, 0),
        ("ffn_up_t_gelu", B * 16, 3072, 768, True, False, 1),
        ("square_8k", 8192, 8192, 8192, False, False, 0),
        ("opt_qkv_prefill", 4900, 7680, 2560, True, False, 0),
        ("opt_fc1_decode", 100, 10240, 2560, True, False, 2),
    ]
    for name, M, N, K, use_bias, use_res, act in shapes:
        a = torch.randn((M, K), device=DEV).to(torch.bfloat16)
        w = (torch.randn((N, K), device=DEV) / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=DEV) if use_bias else None
        res = torch.randn((M, N), device=DEV).to(torch.bfloat16) if use_res else None
        out = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
        med, best = timeit(lambda: ops.gemm(a, w, bias, residual=res, act=act, out=out), iters, flush)
        fl = 2.0 * M * N * K
        print(json.dumps({"kernel": "gemm", "case": name, "M": M, "N": N, "K": K, "ms_median": med, "ms_best": best,
                          "tflops_median": fl / med / 1e9, "tflops_best": fl / best / 1e9}), flush=True)
        del a, w, res, out


def bench_streamk(iters, flush):
    """Decode-shaped GEMMs (M = 100).  Event timing of single launches measures the host's enqueue latency for kernels this
    short, so each case is a CUDA graph of `reps` calls cycling over `reps` different weight matrices (> L2 in total:
    every call streams its weights from HBM, as consecutive decoder layers do), replayed and timed as a whole."""
    for name, M, N, K in (("opt_qkv_decode", 100, 7680, 2560), ("opt_out_decode", 100, 2560, 2560),
                          ("opt_fc1_decode", 100, 10240, 2560), ("opt_fc2_decode", 100, 2560, 10240),
                          ("opt_lm_head", 100, 50272, 2560)):
        reps = max(4, min(32, int(600e6 // (2 * N * K))))
        a = torch.randn((M, K), device=DEV).to(torch.bfloat16)
        ws = [(torch.randn((N, K), device=DEV) / K ** 0.5).to(torch.bfloat16) for _ in range(reps)]
        bias = torch.randn(N, device=DEV)
        out = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
        for fn_name, fn in (("skinny", lambda w: ops.gemm_small_m(a, w, bias, out=out)),
                            ("tiled", lambda w: ops.gemm(a, w, bias, out=out))):
            for w in ws[:2]:
                fn(w)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for w in ws:
                    fn(w)
            graph.replay()
            torch.cuda.synchronize()
            ms = []
            for _ in range(iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                graph.replay()
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1) / reps)
            ms.sort()
            med = ms[len(ms) // 2]
            print(json.dumps({"kernel": "gemm_" + fn_name, "case": name, "M": M, "N": N, "K": K, "reps_in_graph": reps,
                              "us_per_call_median": med * 1e3, "weight_GBps_median": 2.0 * N * K / med / 1e6}), flush=True)
        del ws


def bench_layers(iters, flush):
    """Decode-step GEMM chain of OPT-2.7B layers (qkv -> out -> fc1 -> fc2 with the real data dependencies, M = 100),
    16 layers of distinct weights (2.5 GB) in one CUDA graph: us per layer for the small-M kernel and the tiled one."""
    import os
    d, ffn, M, L = 2560, 10240, 100, int(os.environ.get('KB_L', '16'))
    nodep = os.environ.get('KB_NODEP', '0') == '1'
    layers = []
    for _ in range(L):
        layers.append(dict(
            w_qkv=(torch.randn((3 * d, d), device=DEV) * 0.02).to(torch.bfloat16), b_qkv=torch.zeros(3 * d, device=DEV),
            w_o=(torch.randn((d, d), device=DEV) * 0.02).to(torch.bfloat16), b_o=torch.zeros(d, device=DEV),
            w_fc1=(torch.randn((ffn, d), device=DEV) * 0.02).to(torch.bfloat16), b_fc1=torch.zeros(ffn, device=DEV),
            w_fc2=(torch.randn((d, ffn), device=DEV) * 0.02).to(torch.bfloat16), b_fc2=torch.zeros(d, device=DEV)))
    h0 = torch.randn((M, d), device=DEV).to(torch.bfloat16)
    for fn_name, gemm in (("skinny", ops.gemm_small_m), ("tiled", ops.gemm)):
        def run():
            h = h0.clone()
            if nodep:
                f0 = torch.zeros((M, ffn), device=DEV, dtype=torch.bfloat16)
                o1 = torch.empty((M, 3 * d), device=DEV, dtype=torch.bfloat16)
                o2 = torch.empty((M, d), device=DEV, dtype=torch.bfloat16)
                o3 = torch.empty((M, ffn), device=DEV, dtype=torch.bfloat16)
                for lw in layers:
                    gemm(h0, lw["w_qkv"], lw["b_qkv"], out=o1)
                    gemm(h0, lw["w_o"], lw["b_o"], out=o2)
                    gemm(h0, lw["w_fc1"], lw["b_fc1"], act=2, out=o3)
                    gemm(f0, lw["w_fc2"], lw["b_fc2"], out=o2)
                return h
            for lw in layers:
                qkv = gemm(h, lw["w_qkv"], lw["b_qkv"])
                gemm(qkv[:, :d], lw["w_o"], lw["b_o"], residual=h, out=h)
                f = gemm(h, lw["w_fc1"], lw["b_fc1"], act=2)
                gemm(f, lw["w_fc2"], lw["b_fc2"], residual=h, out=h)
            return h
        run()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run()
        graph.replay()
        torch.cuda.synchronize()
        ms = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / L)
        ms.sort()
        print(json.dumps({"kernel": "layer_gemms_" + fn_name, "layers": L, "us_per_layer_median": ms[len(ms) // 2] * 1e3,
                          "weight_GBps": 2.0 * (3 * d * d + d * d + 2 * d * ffn) / ms[len(ms) // 2] / 1e6}), flush=True)


def bench_xattn(iters, flush):
    for N, L in ((40, 256), (80, 256), (40, 252)):
        B = N * N
        g = torch.Generator(device="cpu").manual_seed(1)
        q = (torch.randn((B * 33, 768), generator=g)).to(torch.bfloat16).to(DEV)
        k = (torch.randn((L, 768), generator=g)).to(torch.bfloat16).to(DEV)
        Lp = (L + 7) // 8 * 8
        vt = torch.zeros((768, Lp), dtype=torch.bfloat16, device=DEV)
        vt[:, :L] = torch.randn((768, L), generator=g).to(torch.bfloat16).to(DEV)
        words = (L + 31) // 32
        bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (N, words), dtype=torch.int64, generator=g).to(torch.int32).to(DEV)
        out = torch.empty_like(q)
        med, best = timeit(lambda: ops.xattn_pairs(q, k, vt, bits, N, B, 33, L, 12, 64, out=out), iters, flush)
        fl = 4.0 * B * 33 * L * 768
        print(json.dumps({"kernel": "xattn_pairs", "N": N, "L": L, "ms_median": med, "ms_best": best,
                          "tflops_median": fl / med / 1e9, "tflops_best": fl / best / 1e9}), flush=True)


def bench_xattn_cfg2(iters, flush):
    """K5 on the panoptic masks of the synthetic cfg2 / cfg5 images (compact objects: most 32-key chunks invisible to a
    32-row quarter), mask-bias tiles prebuilt as in the pipeline."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from openpsg_b200 import synth
    for name in ("cfg2", "cfg5"):
        wl = synth.WORKLOADS[name]
        inp = synth.make_image_inputs(wl, 0)
        N, L = wl.num_objects, wl.image_tokens
        B = N * N
        pan = inp["object_info"][0]["pan_results"].to(torch.int32).to(DEV)
        ids = torch.tensor([int(i) for i in inp["object_info"][0]["object_id_list"]], dtype=torch.int32, device=DEV)
        bits = ops.pair_mask_bits(pan, (wl.height, wl.width), (wl.height, wl.width), (16, 16), ids)
        g = torch.Generator(device="cpu").manual_seed(1)
        q = (torch.randn((B * 33, 768), generator=g)).to(torch.bfloat16).to(DEV)
        k = (torch.randn((L, 768), generator=g)).to(torch.bfloat16).to(DEV)
        vt = torch.randn((768, L), generator=g).to(torch.bfloat16).to(DEV)
        out = torch.empty_like(q)
        perm, bits_sorted = ops.token_order(bits, L)
        for tag, bb in (("token_order", bits), ("object_order", bits_sorted)):
            tiles = ops.xattn_bias_tiles(bb, N, B, 33, L)
            vis = tiles[-((B * 33 + 127) // 128) * 16:].view(torch.int16).view(-1, 8)[:, :4].to(torch.int32) & 0xFFFF
            frac = sum(bin(int(x)).count("1") for x in vis.flatten().tolist()) / (16.0 * vis.numel())
            med, best = timeit(lambda: ops.xattn_pairs(q, k, vt, bb, N, B, 33, L, 12, 64, out=out, bias_tiles=tiles), iters, flush)
            fl = 4.0 * B * 33 * L * 768
            print(json.dumps({"kernel": "xattn_pairs_" + name + "_masks_" + tag, "N": N, "L": L, "visible_chunk_frac": round(frac, 3),
                              "ms_median": med, "ms_best": best, "tflops_median": fl / med / 1e9, "tflops_best": fl / best / 1e9}),
                  flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["gemm", "xattn"])
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    if "gemm" in args.which:
        bench_gemm(args.iters, flush)
    if "streamk" in args.which:
        bench_streamk(args.iters, flush)
    if "layers" in args.which:
        bench_layers(args.iters, flush)
    if "xattn" in args.which:
        bench_xattn(args.iters, flush)
        bench_xattn_cfg2(args.iters, flush)


if __name__ == "__main__":
    main()
