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
    for name, M, N, K in (("opt_qkv_decode", 100, 7680, 2560), ("opt_out_decode", 100, 2560, 2560),
                          ("opt_fc1_decode", 100, 10240, 2560), ("opt_fc2_decode", 100, 2560, 10240),
                          ("opt_lm_head", 100, 50272, 2560)):
        a = torch.randn((M, K), device=DEV).to(torch.bfloat16)
        w = (torch.randn((N, K), device=DEV) / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=DEV)
        out = torch.empty((M, N), device=DEV, dtype=torch.bfloat16)
        for fn_name, fn in (("streamk", lambda: ops.gemm_small_m(a, w, bias, out=out)), ("tiled", lambda: ops.gemm(a, w, bias, out=out))):
            med, best = timeit(fn, iters, flush)
            print(json.dumps({"kernel": "gemm_" + fn_name, "case": name, "M": M, "N": N, "K": K, "ms_median": med,
                              "weight_GBps_median": 2.0 * N * K / med / 1e6}), flush=True)


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
    if "xattn" in args.which:
        bench_xattn(args.iters, flush)


if __name__ == "__main__":
    main()
