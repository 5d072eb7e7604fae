"""torch.Tensor -> raw pointer marshalling for the C ABI (include/opsg_b200.h).

PyTorch is used only for device memory and streams; every function here enqueues hand-written CUDA
kernels from libopsg_b200.so on the current torch CUDA stream and returns without synchronising.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, OUT_BF16, OUT_F32, OUT_F32_ATOMIC  # noqa: F401

launch_count = 0   # kernels enqueued through this module (bench.py reports it as gpu_launches)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda(t: torch.Tensor, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (libopsg_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if t.device.index != torch.cuda.current_device():
        # kernels, TMA descriptors and the stream all belong to the CURRENT device: a tensor elsewhere would be launched on
        # the wrong GPU.  One process per GPU sets the device once (torch.cuda.set_device), as bench.py / sharding.py do.
        raise ValueError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                         "call torch.cuda.set_device (or use `with torch.cuda.device(...)`) before calling the head")
    return t


def _count(n=1):
    global launch_count
    launch_count += n


# ---- optional per-kernel timing (CUDA events on the launching stream; used by bench.py) ---------------
_profile = None      # None = off; else list of (name, start_event, stop_event, flops, bytes)


def profile_begin():
    global _profile
    _profile = []


def profile_end() -> dict:
    """-> {kernel name: {"ms": total device ms, "n": launches, "flops": algorithmic flops, "bytes": algorithmic bytes}}"""
    global _profile
    recs, _profile = _profile or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, flops, nbytes in recs:
        r = out.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
        r["ms"] += e0.elapsed_time(e1)
        r["n"] += 1
        r["flops"] += flops
        r["bytes"] += nbytes
    return out


class _timed:
    def __init__(self, name, flops=0.0, nbytes=0.0):
        self.name, self.flops, self.nbytes = name, flops, nbytes

    def __enter__(self):
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _profile is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _profile.append((self.name, self.e0, e1, self.flops, self.nbytes))
        return False


def pair_mask_bits(pan: torch.Tensor, img_hw, pad_hw, tok_hw, obj_ids: torch.Tensor, words: Optional[int] = None):
    """K2: int32 pan map [h,w] + object ids [N] -> uint32-as-int32 bit masks [N, words]."""
    pan = _cuda(pan, torch.int32, "pan").contiguous()
    obj_ids = _cuda(obj_ids, torch.int32, "obj_ids").contiguous()
    L = tok_hw[0] * tok_hw[1]
    words = words or max(1, (L + 31) // 32)
    bits = torch.empty((obj_ids.numel(), words), dtype=torch.int32, device=pan.device)
    with _timed("pair_mask_bits", 0.0, 4.0 * tok_hw[0] * tok_hw[1] + 4.0 * bits.numel()):
        _lib.check(_lib.load().opsg_pair_mask_bits(_ptr(pan), pan.shape[0], pan.shape[1], int(img_hw[0]), int(img_hw[1]),
                                                  int(pad_hw[0]), int(pad_hw[1]), int(tok_hw[0]), int(tok_hw[1]),
                                                  _ptr(obj_ids), obj_ids.numel(), _ptr(bits), words, _stream()))
    _count()
    return bits


def patch_im2col(feat: torch.Tensor, patch: int) -> torch.Tensor:
    """K1 operand: fp32 [C,h,w] -> bf16 [L, C*patch*patch]."""
    feat = _cuda(feat, torch.float32, "feat").contiguous()
    C, h, w = feat.shape
    L = (h // patch) * (w // patch)
    out = torch.empty((L, C * patch * patch), dtype=torch.bfloat16, device=feat.device)
    with _timed("patch_im2col", 0.0, 6.0 * out.numel()):
        _lib.check(_lib.load().opsg_patch_im2col(_ptr(feat), C, h, w, patch, _ptr(out), _stream()))
    _count()
    return out


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, residual: Optional[torch.Tensor] = None,
         act: int = ACT_NONE, out: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16, bias_along_m: bool = False,
         k_splits: int = 1, atomic: bool = False) -> torch.Tensor:
    """D = act(a @ w.T + bias + residual).  a bf16 [M,K], w bf16 [N,K] (nn.Linear layout)."""
    _cuda(a, torch.bfloat16, "a"); _cuda(w, torch.bfloat16, "w")
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    mode = OUT_F32_ATOMIC if atomic else (OUT_BF16 if out.dtype == torch.bfloat16 else OUT_F32)
    if bias is not None:
        _cuda(bias, torch.float32, "bias")
    if residual is not None:
        _cuda(residual, torch.bfloat16, "residual")
        assert residual.shape == (M, N) and residual.stride(1) == 1
    # profiling key: the large GEMMs (CTA-pair kernel, the dominant kernel of the step) apart from the small per-image ones
    # (K3 projections of 256 image tokens, the 33 shared query rows, PatchEmbed's split-K)
    key = "gemm_bf16" if 2.0 * M * N * K >= 2e9 and k_splits == 1 else "gemm_bf16_small"
    with _timed(key, 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N)):
        _lib.check(_lib.load().opsg_gemm_bf16(_ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(out), out.stride(0), M, N, K,
                                             _ptr(bias), int(bias_along_m), _ptr(residual),
                                             residual.stride(0) if residual is not None else 0, act, mode, k_splits, _stream()))
    _count()
    return out


def gemm_ln(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, residual: Optional[torch.Tensor] = None,
            act: int = ACT_NONE, out: Optional[torch.Tensor] = None, a_stats: Optional[torch.Tensor] = None,
            a_colsum: Optional[torch.Tensor] = None, r_stats: Optional[torch.Tensor] = None,
            r_gamma: Optional[torch.Tensor] = None, r_beta: Optional[torch.Tensor] = None,
            stats_out: Optional[torch.Tensor] = None, eps: float = 1e-12) -> torch.Tensor:
    """LayerNorm-folding GEMM (opsg_gemm_bf16_ln): a / residual may be un-normalised tensors with fp32 [rows, 2]
    (sum, sumsq) statistics; stats_out (zeroed by the caller) receives the statistics of the output rows."""
    _cuda(a, torch.bfloat16, "a"); _cuda(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    for t, n in ((a_stats, M), (r_stats, M), (stats_out, M)):
        assert t is None or (t.dtype == torch.float32 and t.shape == (n, 2) and t.is_contiguous())
    with _timed("gemm_bf16", 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N)):
        _lib.check(_lib.load().opsg_gemm_bf16_ln(_ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(out), out.stride(0), M, N, K,
                                                _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0,
                                                act, _ptr(a_stats), _ptr(a_colsum), _ptr(r_stats), _ptr(r_gamma), _ptr(r_beta),
                                                _ptr(stats_out), float(eps), _stream()))
    _count()
    return out


_streamk_ws = {}     # device -> uint8 workspace, grown on demand
_streamk_ws_keep = []  # outgrown workspaces stay alive: captured CUDA graphs may still point at them


def gemm_small_m(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
                 residual: Optional[torch.Tensor] = None, act: int = ACT_NONE, out: Optional[torch.Tensor] = None,
                 out_dtype=torch.bfloat16) -> torch.Tensor:
    """Weight-streaming GEMM for M <= 128 rows (LLM decode): K-sliced, activations resident in TMEM, deterministic
    fp32 partial reduction in a second kernel (csrc/gemm_skinny.cu)."""
    _cuda(a, torch.bfloat16, "a"); _cuda(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert M <= 128 and w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    lib = _lib.load()
    need = lib.opsg_gemm_streamk_workspace_bytes(N, K)
    ws = _streamk_ws.get(a.device)
    if ws is None or ws.numel() < need:
        if ws is not None:
            _streamk_ws_keep.append(ws)
        ws = torch.empty(max(need, 64 << 20), dtype=torch.uint8, device=a.device)
        _streamk_ws[a.device] = ws
    if residual is not None:
        _cuda(residual, torch.bfloat16, "residual")
    mode = OUT_BF16 if out.dtype == torch.bfloat16 else OUT_F32
    with _timed("gemm_streamk", 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N)):
        _lib.check(lib.opsg_gemm_bf16_streamk(_ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(out), out.stride(0), M, N, K,
                                              _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0,
                                              act, mode, _ptr(ws), ws.numel(), _stream()))
    _count(2)
    return out


def cast_f32_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(x, torch.float32, "x")
    rows, cols = x.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device)
    with _timed("cast_f32_bf16", 0.0, 6.0 * rows * cols):
        _lib.check(_lib.load().opsg_cast_f32_bf16(_ptr(x), x.stride(0), _ptr(out), out.stride(0), rows, cols, _stream()))
    _count()
    return out


def init_rows_f32(out: torch.Tensor, row: Optional[torch.Tensor]) -> torch.Tensor:
    _cuda(out, torch.float32, "out")
    with _timed("init_rows_f32", 0.0, 4.0 * out.numel()):
        _lib.check(_lib.load().opsg_init_rows_f32(_ptr(out), out.stride(0), _ptr(row), out.shape[0], out.shape[1], _stream()))
    _count()
    return out


def qformer_embed_ln(query, input_ids, word_emb, pos_emb, gamma, beta, eps, out=None):
    """K7: rows [B*nq + B*T, d] bf16 in the split (query rows | text rows) layout."""
    nq, d = query.shape
    B, T = input_ids.shape
    _cuda(input_ids, torch.int32, "input_ids")
    if out is None:
        out = torch.empty((B * (nq + T), d), dtype=torch.bfloat16, device=query.device)
    with _timed("qformer_embed_ln", 0.0, 2.0 * out.numel() + 4.0 * B * T * d):
        _lib.check(_lib.load().opsg_qformer_embed_ln(_ptr(query), nq, _ptr(input_ids), B, T, _ptr(word_emb), word_emb.shape[0],
                                                    _ptr(pos_emb), _ptr(gamma), _ptr(beta), float(eps), d, _ptr(out), _stream()))
    _count(2 if B > 1 else 1)          # LN kernel + broadcast of the pair-independent query rows
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out=None) -> torch.Tensor:
    _cuda(x, torch.bfloat16, "x")
    assert x.is_contiguous()
    rows, cols = x.shape
    if out is None:
        out = torch.empty_like(x)
    with _timed("layernorm_bf16", 0.0, 4.0 * rows * cols):
        _lib.check(_lib.load().opsg_layernorm_bf16(_ptr(x), _ptr(gamma), _ptr(beta), float(eps), _ptr(out), rows, cols, _stream()))
    _count()
    return out


def self_attn_small(qkv, text_mask, B, n_query, T, num_heads, head_dim, text_queries: bool, out=None, shared_query_qkv=None):
    """shared_query_qkv: bf16 [n_query, 3*d] q/k/v of the query rows when they are the same for every pair (layer 0)."""
    _cuda(qkv, torch.bfloat16, "qkv")
    if shared_query_qkv is not None:
        _cuda(shared_query_qkv, torch.bfloat16, "shared_query_qkv")
        assert shared_query_qkv.is_contiguous() and tuple(shared_query_qkv.shape) == (n_query, 3 * num_heads * head_dim)
    d = num_heads * head_dim
    assert qkv.is_contiguous() and qkv.shape[1] == 3 * d
    rows = B * (n_query + T) if text_queries else B * n_query
    if out is None:
        out = torch.empty((rows, d), dtype=torch.bfloat16, device=qkv.device)
    with _timed("self_attn_small", 4.0 * B * num_heads * (n_query + (T if text_queries else 0)) * (n_query + T) * head_dim, 2.0 * (qkv.numel() + out.numel())):
        _lib.check(_lib.load().opsg_self_attn_small(_ptr(qkv), _ptr(shared_query_qkv), _ptr(text_mask), B, n_query, T, num_heads, head_dim,
                                                   int(text_queries), _ptr(out), _stream()))
    _count()
    return out


def token_order(bits: torch.Tensor, L: int):
    """K5 key order: (perm int32 [L] = image tokens sorted by owning object, the object masks in that order)."""
    _cuda(bits, torch.int32, "bits")
    assert bits.is_contiguous()
    perm = torch.empty((L,), dtype=torch.int32, device=bits.device)
    bits_sorted = torch.empty_like(bits)
    with _timed("token_order", 0.0, 8.0 * bits.numel() + 4.0 * L):
        _lib.check(_lib.load().opsg_token_order(_ptr(bits), bits.shape[1], bits.shape[0], L, _ptr(perm), _ptr(bits_sorted), _stream()))
    _count()
    return perm, bits_sorted


def xattn_bias_tiles(bits, num_objects, B, n_query, L, pair_index=None):
    """K5 operand prep (once per image): pair masks as tensor-core bias tiles -> uint8 buffer for xattn_pairs."""
    _cuda(bits, torch.int32, "bits")
    lib = _lib.load()
    nbytes = lib.opsg_xattn_bias_tiles_bytes(B, n_query)
    tiles = torch.empty(nbytes, dtype=torch.uint8, device=bits.device)
    with _timed("xattn_bias_tiles", 0.0, float(nbytes)):
        _lib.check(lib.opsg_xattn_bias_tiles(_ptr(bits), bits.shape[1], _ptr(pair_index), num_objects, B, n_query, L,
                                             _ptr(tiles), _stream()))
    _count()
    return tiles


def xattn_pairs(q, k, vt, bits, num_objects, B, n_query, L, num_heads, head_dim, pair_index=None, out=None, bias_tiles=None):
    """K5: q bf16 [B*nq, d]; k bf16 [L, d]; vt bf16 [d, ld>=L]; bits int32 [N, words]; bias_tiles from
    xattn_bias_tiles (None -> built here; False -> none: the library's self-contained kernel)."""
    _cuda(q, torch.bfloat16, "q"); _cuda(k, torch.bfloat16, "k"); _cuda(vt, torch.bfloat16, "vt")
    assert q.is_contiguous()
    if out is None:
        out = torch.empty_like(q)
    if bias_tiles is None and L > 256:                    # the online-softmax kernel reads the mask bits directly
        bias_tiles = False
    if bias_tiles is None:
        bias_tiles = xattn_bias_tiles(bits, num_objects, B, n_query, L, pair_index)
    elif bias_tiles is False:
        bias_tiles = None
    with _timed("xattn_pairs", 4.0 * B * n_query * L * num_heads * head_dim, 4.0 * q.numel() + 4.0 * L * num_heads * head_dim):
        _lib.check(_lib.load().opsg_xattn_pairs(_ptr(q), _ptr(k), k.stride(0), _ptr(vt), vt.stride(0), _ptr(bits), bits.shape[1],
                                               _ptr(pair_index), num_objects, B, n_query, L, num_heads, head_dim,
                                               _ptr(bias_tiles), _ptr(out), _stream()))
    _count()
    return out


def exist_filter_topk(x: torch.Tensor, ld_x: int, B: int, d: int, w: torch.Tensor, b: torch.Tensor, threshold: float, k: int):
    """K8: returns (logits fp32 [B], probs fp32 [B], mask uint8 [B], topk int32 [k])."""
    dev = x.device
    logits = torch.empty(B, dtype=torch.float32, device=dev)
    probs = torch.empty(B, dtype=torch.float32, device=dev)
    mask = torch.empty(B, dtype=torch.uint8, device=dev)
    topk = torch.empty(max(k, 1), dtype=torch.int32, device=dev)
    with _timed("exist_filter_topk", 2.0 * B * d, 2.0 * B * d + 12.0 * B):
        _lib.check(_lib.load().opsg_exist_filter_topk(_ptr(x), ld_x, B, d, _ptr(w), _ptr(b), float(threshold), k, _ptr(logits),
                                                     _ptr(probs), _ptr(mask), _ptr(topk), _stream()))
    _count(2)
    return logits, probs, mask, topk[:k]


def mask_pool_labels(pan: torch.Tensor, img_hw, pad_hw, feat_hw, obj_ids: torch.Tensor):
    """K11 stage 1: (label int32 [h, w] = owning object slot or N, rep int32 [N]) from the panoptic map."""
    pan = _cuda(pan, torch.int32, "pan").contiguous()
    obj_ids = _cuda(obj_ids, torch.int32, "obj_ids").contiguous()
    n = obj_ids.numel()
    label = torch.empty((int(feat_hw[0]), int(feat_hw[1])), dtype=torch.int32, device=pan.device)
    rep = torch.empty((n,), dtype=torch.int32, device=pan.device)
    with _timed("mask_pool_labels", 0.0, 8.0 * label.numel()):
        _lib.check(_lib.load().opsg_mask_pool_labels(_ptr(pan), pan.shape[0], pan.shape[1], int(img_hw[0]), int(img_hw[1]),
                                                    int(pad_hw[0]), int(pad_hw[1]), int(feat_hw[0]), int(feat_hw[1]),
                                                    _ptr(obj_ids), n, _ptr(label), _ptr(rep), _stream()))
    _count()
    return label, rep


def mask_pool_pairs(feat: torch.Tensor, label: torch.Tensor, num_objects: int, with_pairs: bool = True, rep=None,
                    cls_table: Optional[torch.Tensor] = None, cls_ids: Optional[torch.Tensor] = None, cls_mode: str = "none",
                    use_background: bool = False):
    """K11: feat fp32 [C,h,w], label int32 [h,w] (ops.mask_pool_labels) -> (obj [N,C'], pair [N*N,2C'] or None)."""
    _cuda(feat, torch.float32, "feat"); _cuda(label, torch.int32, "label")
    feat, label = feat.contiguous(), label.contiguous()
    C, h, w = feat.shape
    mode = {"none": 0, "add": 1, "cat": 2}[cls_mode]
    cls_dim = 0
    if mode:
        _cuda(cls_table, torch.float32, "cls_table"); _cuda(cls_ids, torch.int32, "cls_ids")
        cls_table, cls_ids = cls_table.contiguous(), cls_ids.contiguous()
        cls_dim = cls_table.shape[1]
    Co = C + (cls_dim if mode == 2 else 0)
    lib = _lib.load()
    ws = torch.empty(lib.opsg_mask_pool_workspace_bytes(C, h, w, num_objects), dtype=torch.uint8, device=feat.device)
    obj = torch.empty((num_objects, Co), dtype=torch.float32, device=feat.device)
    pair = torch.empty((num_objects * num_objects, 2 * Co), dtype=torch.float32, device=feat.device) if with_pairs else None
    with _timed("mask_pool_pairs", 0.0, 4.0 * feat.numel() + 4.0 * label.numel()):
        _lib.check(lib.opsg_mask_pool_pairs(_ptr(feat), C, h, w, _ptr(label), _ptr(rep), num_objects, _ptr(cls_table if mode else None),
                                            _ptr(cls_ids if mode else None), cls_dim, mode, int(use_background), _ptr(ws), ws.numel(),
                                            _ptr(obj), _ptr(pair), _stream()))
    _count(4 if with_pairs else 3)
    return obj, pair


def gather_rows(src: torch.Tensor, row_elems: int, idx: torch.Tensor, out=None):
    _cuda(src, torch.bfloat16, "src"); _cuda(idx, torch.int32, "idx")
    n = idx.numel()
    if out is None:
        out = torch.empty((n, row_elems), dtype=torch.bfloat16, device=src.device)
    with _timed("gather_rows", 0.0, 4.0 * n * row_elems):
        _lib.check(_lib.load().opsg_gather_rows_bf16(_ptr(src), row_elems, _ptr(idx), n, _ptr(out), _stream()))
    _count()
    return out


def embed_gather(table, ids, out, pos_table=None, pos=None):
    _cuda(table, torch.bfloat16, "table"); _cuda(ids, torch.int32, "ids")
    n, d = ids.numel(), table.shape[1]
    with _timed("embed_gather", 0.0, 6.0 * n * d):
        _lib.check(_lib.load().opsg_embed_gather(_ptr(table), d, _ptr(ids), _ptr(pos_table), _ptr(pos), n, _ptr(out),
                                                out.stride(0), _stream()))
    _count()
    return out


def llm_build_prefix(proj, rows_per_seq, row0, n_prefix, table, ids, pos_table, pos, out):
    """a9: out [nseq, n_prefix+T, d] = cat(proj rows, table[ids]) + pos_table[pos]."""
    _cuda(proj, torch.bfloat16, "proj"); _cuda(table, torch.bfloat16, "table"); _cuda(ids, torch.int32, "ids")
    nseq, T = ids.shape
    d = table.shape[1]
    assert out.is_contiguous() and out.shape == (nseq, n_prefix + T, d)
    with _timed("llm_build_prefix", 0.0, 6.0 * out.numel()):
        _lib.check(_lib.load().opsg_llm_build_prefix(_ptr(proj), rows_per_seq, row0, n_prefix, _ptr(table), _ptr(ids), T,
                                                    _ptr(pos_table), _ptr(pos), nseq, d, _ptr(out), _stream()))
    _count()
    return out


def llm_attn(q, k_cache, v_cache, key_mask, nseq, q_len, q_pos0, num_heads, head_dim, scale, out):
    max_ctx = k_cache.shape[1]
    ctx = q_pos0 + q_len
    with _timed("llm_attn", 4.0 * nseq * q_len * ctx * num_heads * head_dim,
                4.0 * nseq * ctx * num_heads * head_dim + 4.0 * nseq * q_len * num_heads * head_dim):
        _lib.check(_lib.load().opsg_llm_attn(_ptr(q), q.stride(0), _ptr(k_cache), _ptr(v_cache), max_ctx, _ptr(key_mask), nseq,
                                            q_len, q_pos0, num_heads, head_dim, float(scale), _ptr(out), out.stride(0), _stream()))
    _count()
    return out


def llm_attn_append(qkv, k_cache, v_cache, key_mask, nseq, q_pos0, num_heads, head_dim, scale, out):
    """Decode step: append the new tokens' k / v (from the fused qkv rows) to the caches at q_pos0 and attend, one launch."""
    max_ctx = k_cache.shape[1]
    ctx = q_pos0 + 1
    with _timed("llm_attn", 4.0 * nseq * ctx * num_heads * head_dim, 4.0 * nseq * ctx * num_heads * head_dim):
        _lib.check(_lib.load().opsg_llm_attn_append(_ptr(qkv), qkv.stride(0), _ptr(k_cache), _ptr(v_cache), max_ctx, _ptr(key_mask),
                                                   nseq, q_pos0, num_heads, head_dim, float(scale), _ptr(out), out.stride(0),
                                                   _stream()))
    _count()
    return out


def kv_append(qkv, nseq, q_len, pos0, d_model, k_cache, v_cache):
    with _timed("kv_append", 0.0, 8.0 * nseq * q_len * d_model):
        _lib.check(_lib.load().opsg_kv_append(_ptr(qkv), qkv.stride(0), nseq, q_len, pos0, d_model, _ptr(k_cache), _ptr(v_cache),
                                             k_cache.shape[1], _stream()))
    _count()


def argmax_rows(logits: torch.Tensor, out=None):
    _cuda(logits, torch.float32, "logits")
    rows, cols = logits.shape
    if out is None:
        out = torch.empty(rows, dtype=torch.int32, device=logits.device)
    with _timed("argmax_rows", 0.0, 4.0 * rows * cols):
        _lib.check(_lib.load().opsg_argmax_rows(_ptr(logits), logits.stride(0), rows, cols, _ptr(out), _stream()))
    _count()
    return out


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float, out=None) -> torch.Tensor:
    """Llama RMSNorm over the last dim; x bf16 [rows, cols] (row stride may exceed cols), weight fp32 [cols]."""
    _cuda(x, torch.bfloat16, "x"); _cuda(weight, torch.float32, "weight")
    assert x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device)
    with _timed("rmsnorm_bf16", 0.0, 4.0 * rows * cols):
        _lib.check(_lib.load().opsg_rmsnorm_bf16(_ptr(x), x.stride(0), _ptr(weight), float(eps), _ptr(out), out.stride(0),
                                                rows, cols, _stream()))
    _count()
    return out


def rope(x: torch.Tensor, n_parts: int, num_heads: int, head_dim: int, pos: torch.Tensor, cos_table: torch.Tensor,
         sin_table: torch.Tensor) -> torch.Tensor:
    """In-place rotary embedding of the first n_parts (q, k) column blocks of x bf16 [rows, ld]; pos int32 [rows]."""
    _cuda(x, torch.bfloat16, "x"); _cuda(pos, torch.int32, "pos")
    _cuda(cos_table, torch.float32, "cos_table"); _cuda(sin_table, torch.float32, "sin_table")
    assert x.dim() == 2 and x.stride(1) == 1 and pos.numel() == x.shape[0] and pos.is_contiguous()
    assert cos_table.is_contiguous() and sin_table.is_contiguous() and cos_table.shape == sin_table.shape == (cos_table.shape[0], head_dim // 2)
    rows = x.shape[0]
    with _timed("rope_bf16", 0.0, 4.0 * rows * n_parts * num_heads * head_dim):
        _lib.check(_lib.load().opsg_rope_bf16(_ptr(x), x.stride(0), rows, n_parts, num_heads, head_dim, _ptr(pos), _ptr(cos_table),
                                             _ptr(sin_table), cos_table.shape[0], _stream()))
    _count()
    return x


def swiglu(gate_up: torch.Tensor, ffn: int, out=None) -> torch.Tensor:
    """out = silu(gate_up[:, :ffn]) * gate_up[:, ffn:2*ffn]; bf16."""
    _cuda(gate_up, torch.bfloat16, "gate_up")
    assert gate_up.dim() == 2 and gate_up.stride(1) == 1 and gate_up.shape[1] >= 2 * ffn
    rows = gate_up.shape[0]
    if out is None:
        out = torch.empty((rows, ffn), dtype=torch.bfloat16, device=gate_up.device)
    with _timed("swiglu_bf16", 0.0, 6.0 * rows * ffn):
        _lib.check(_lib.load().opsg_swiglu_bf16(_ptr(gate_up), gate_up.stride(0), rows, ffn, _ptr(out), out.stride(0), _stream()))
    _count()
    return out


class PromptLayout:
    __slots__ = ("pos", "key_mask", "last_rows", "dec_pos")


def llm_prompt_layout(text_mask: torch.Tensor, n_prefix: int, max_new_tokens: int, pos_offset: int) -> PromptLayout:
    """Positions / key mask / last-row indices / decode-step positions of the batched prompt (one kernel; see
    include/opsg_b200.h).  text_mask int32 [k, T] (left-padded)."""
    _cuda(text_mask, torch.int32, "text_mask")
    assert text_mask.is_contiguous()
    k, T = text_mask.shape
    dev = text_mask.device
    Tp = n_prefix + T
    lay = PromptLayout()
    lay.pos = torch.empty((k, Tp), dtype=torch.int32, device=dev)
    lay.key_mask = torch.empty((k, Tp + max_new_tokens), dtype=torch.uint8, device=dev)
    lay.last_rows = torch.empty((k,), dtype=torch.int32, device=dev)
    lay.dec_pos = torch.empty((max(max_new_tokens - 1, 1), k), dtype=torch.int32, device=dev)
    with _timed("llm_prompt_layout", 0.0, 4.0 * k * T + 5.0 * k * (Tp + max_new_tokens)):
        _lib.check(_lib.load().opsg_llm_prompt_layout(_ptr(text_mask), k, T, n_prefix, max_new_tokens, pos_offset, _ptr(lay.pos),
                                                     _ptr(lay.key_mask), _ptr(lay.last_rows), _ptr(lay.dec_pos), _stream()))
    _count()
    return lay


def copy_into(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst <- src.  Device-resident sources of the same dtype / shape go through the library's copy KERNEL (a copy-engine
    D2D copy would queue behind an in-flight host->device prefetch); anything else (host source, dtype conversion) is a
    plain asynchronous ``copy_``."""
    if src.is_cuda and src.device == dst.device and src.dtype == dst.dtype and src.shape == dst.shape \
            and src.is_contiguous() and dst.is_contiguous():
        if src.data_ptr() != dst.data_ptr():
            nbytes = src.numel() * src.element_size()
            with _timed("copy_bytes", 0.0, 2.0 * nbytes):
                _lib.check(_lib.load().opsg_copy_bytes(_ptr(dst), _ptr(src), nbytes, _stream()))
            _count()
    else:
        dst.copy_(src, non_blocking=True)
    return dst


def transpose_i32(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    _cuda(src, torch.int32, "src"); _cuda(dst, torch.int32, "dst")
    rows, cols = src.shape
    assert src.is_contiguous() and dst.is_contiguous() and dst.shape == (cols, rows)
    with _timed("transpose_i32", 0.0, 8.0 * rows * cols):
        _lib.check(_lib.load().opsg_transpose_i32(_ptr(src), rows, cols, _ptr(dst), _stream()))
    _count()
    return dst


def gemm_splitk(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], k_splits: int, out: Optional[torch.Tensor] = None,
                residual: Optional[torch.Tensor] = None):
    """Deterministic split-K GEMM for problems with too few output tiles for the machine (PatchEmbed: K = 65536; the LLM's
    out_proj / fc2 at a few hundred rows): every split writes its fp32 partial to its own slice, a second kernel sums the
    slices in split order and applies bias (+ residual, which may alias ``out``) -> bf16 [M, N]."""
    M, K = a.shape
    N = w.shape[0]
    part = torch.empty((k_splits, M, N), dtype=torch.float32, device=a.device)
    gemm(a, w, out=part.view(k_splits * M, N)[:M], k_splits=k_splits)          # the kernel strides the slices by M rows
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    if residual is not None:
        _cuda(residual, torch.bfloat16, "residual")
        assert residual.shape == (M, N) and residual.stride(1) == 1
    with _timed("splitk_reduce", 0.0, 4.0 * part.numel() + 2.0 * M * N):
        _lib.check(_lib.load().opsg_splitk_reduce_bf16(_ptr(part), k_splits, M, N, _ptr(bias), _ptr(residual),
                                                      residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0),
                                                      _stream()))
    _count()
    return out


def gemm_medium_m(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
                  residual: Optional[torch.Tensor] = None, act: int = ACT_NONE, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Linear for a few hundred rows (stacked LLM decode steps).  Wide outputs have enough 256 x 256 tiles for the CTA-pair
    kernel; N = hidden-size outputs with a LONG K (fc2 / down_proj: 40 tiles at 800 x 2560, K = 10240) do not, so K is split
    until the 128 x 256 tiles of the single-CTA kernel fill ONE wave, and the fixed-order reduction applies bias and residual.
    Measured at M = 800 (OPT-2.7B, scripts/gemm_medium_m.py, profiles/r2_gemm_medium_m.md): fc2 79.9 -> 51.6 us.  The fp32
    partial epilogue costs ~20 us, so short-K Linears (out_proj, K = 2560: 24.6 us plain, 46.7 us split) stay on the tiled path."""
    M, K = a.shape
    N = w.shape[0]
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    splits = min(148 // max(1, tiles), K // 512)
    if act != ACT_NONE or K < 6144 or splits < 2 or N % 4 or (M + 255) // 256 * ((N + 255) // 256) >= 74:
        return gemm(a, w, bias, residual=residual, act=act, out=out)
    return gemm_splitk(a, w, bias, splits, out=out, residual=residual)
