"""LLM relation decode engine (rows a9-a10 of SURVEY.md §8): selected pair features -> language projection ->
embedded prompt -> batched OPT prefill + greedy decode over a static KV cache, all on libopsg_b200 kernels.

Reference behaviour being replaced (``relation_transformer_head_v4.py:293-312``): one ``language_model.generate``
per selected pair, batch 1, HF ``DynamicCache`` — k x steps full passes over the LLM weights.  Here all k
selected pairs form ONE batch: prefill is k*(32+T) rows through every Linear (tensor-core bound), each decode
step streams the weights once for the whole batch (HBM bound).  Greedy, EOS never stops generation early: the
caller truncates at EOS when parsing, token ids past EOS are simply unused (``min_new_tokens == max_new_tokens``
in the oracle runs).

Arithmetic follows HF ``models/opt/modeling_opt.py`` (pre-LN decoder :202-253, learned positions with offset 2
:56-70, q scaled by head_dim**-0.5 :151, ReLU FFN, final LayerNorm, tied lm_head) in bf16 with fp32 accumulation,
fp32 softmax / LayerNorm statistics and fp32 logits.  Only ``do_layer_norm_before=True`` models whose
``word_embed_proj_dim == hidden_size`` are supported (OPT-1.3B ... 66B; not 350M).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import ops

N_PREFIX = 32          # relation_query rows handed to the LLM (out[:, 1:33], v4:215)
N_QUERY = 33
MAX_CTX_KERNEL = 128   # opsg_llm_attn capacity (keys per sequence)


def _bf16(t, device):
    return t.detach().to(device=device, dtype=torch.bfloat16).contiguous()


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class PackedOPT:
    """Kernel-ready device copy of an ``OPTForCausalLM`` (bf16 matrices, fp32 biases / LayerNorm)."""

    def __init__(self, language_model, device):
        cfg = language_model.config
        if getattr(cfg, "model_type", "") != "opt":
            raise NotImplementedError(f"libopsg_b200 LLM engine supports OPT decoders; got {getattr(cfg, 'model_type', '?')}")
        if not cfg.do_layer_norm_before or cfg.word_embed_proj_dim != cfg.hidden_size:
            raise NotImplementedError("OPT variants with post-LN or projected embeddings (350M) are not supported")
        if getattr(cfg, "activation_function", "relu") != "relu":
            raise NotImplementedError("OPT activation must be relu")
        self.d = cfg.hidden_size
        self.heads = cfg.num_attention_heads
        self.head_dim = self.d // self.heads
        if self.head_dim not in (64, 80, 128):
            raise NotImplementedError(f"head_dim {self.head_dim} unsupported by opsg_llm_attn")
        self.vocab = cfg.vocab_size
        self.n_layers = cfg.num_hidden_layers
        sd = language_model.state_dict()
        dp = "model.decoder."
        self.embed = _bf16(sd[dp + "embed_tokens.weight"], device)            # [V, d] (also the tied lm_head)
        self.pos = _bf16(sd[dp + "embed_positions.weight"], device)           # [max_pos + 2, d]
        self.final_ln = (_f32(sd[dp + "final_layer_norm.weight"], device), _f32(sd[dp + "final_layer_norm.bias"], device))
        lm_w = sd.get("lm_head.weight")
        self.lm_head = self.embed if lm_w is None or lm_w.data_ptr() == sd[dp + "embed_tokens.weight"].data_ptr() \
            else _bf16(lm_w, device)
        self.layers = []
        for l in range(self.n_layers):
            lp = f"{dp}layers.{l}."
            a = lp + "self_attn."
            self.layers.append(dict(
                ln1=(_f32(sd[lp + "self_attn_layer_norm.weight"], device), _f32(sd[lp + "self_attn_layer_norm.bias"], device)),
                w_qkv=_bf16(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0), device),
                b_qkv=_f32(torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0), device),
                w_o=_bf16(sd[a + "out_proj.weight"], device), b_o=_f32(sd[a + "out_proj.bias"], device),
                ln2=(_f32(sd[lp + "final_layer_norm.weight"], device), _f32(sd[lp + "final_layer_norm.bias"], device)),
                w_fc1=_bf16(sd[lp + "fc1.weight"], device), b_fc1=_f32(sd[lp + "fc1.bias"], device),
                w_fc2=_bf16(sd[lp + "fc2.weight"], device), b_fc2=_f32(sd[lp + "fc2.bias"], device),
            ))


@dataclass
class GenerationOutput:
    tokens: torch.Tensor                   # int32 [k, T_new] greedy token ids
    scores: Optional[torch.Tensor] = None  # fp32 [k, T_new, V] next-token logits (only if return_scores)
    prefix: Optional[torch.Tensor] = None  # bf16 [k, 32+T, d] embedded prompt incl. positions (only if return_scores)


class OPTDecodeEngine:
    LN_EPS = 1e-5

    def __init__(self, weights: PackedOPT, w_proj: torch.Tensor, b_proj: torch.Tensor, use_cuda_graphs: bool = True):
        self.w = weights
        self.w_proj, self.b_proj = w_proj, b_proj      # language_projection (v4:97): bf16 [d_llm, 768], fp32 [d_llm]
        if w_proj.shape[0] != weights.d:
            raise ValueError(f"language_projection out_features {w_proj.shape[0]} != LLM hidden size {weights.d}")
        self.use_cuda_graphs = use_cuda_graphs
        # decode steps (rows <= 128) take the K-sliced small-M GEMM (csrc/gemm_skinny.cu); OPSG_LLM_SMALL_M=0 keeps the tiled one
        self.small_m = os.environ.get("OPSG_LLM_SMALL_M", "1") == "1"
        self._graphs = {}          # (hidden shape, k, T, max_new_tokens) -> captured generate()

    # one decoder layer over `rows` = nseq * q_len token rows; h is updated in place
    def _layer(self, lw, h, k_cache, v_cache, key_mask, nseq, q_len, pos0):
        w = self.w
        d = w.d
        # Decode steps (rows <= 128) stream the weights through the K-sliced kernel with the activations resident in TMEM:
        # 17 / 11 / 21 / 20 us for qkv / out / fc1 / fc2 at k = 100 inside a graph against 22 / 13 / 24 / 40 us for the tiled
        # kernel (scripts/kbench.py streamk, profiles/r1_llm_decode.md).
        gemm = ops.gemm_small_m if (self.small_m and h.shape[0] <= 128) else ops.gemm
        x = ops.layernorm(h, lw["ln1"][0], lw["ln1"][1], self.LN_EPS)
        qkv = gemm(x, lw["w_qkv"], lw["b_qkv"])                                        # [rows, 3d]
        ctx = torch.empty((nseq * q_len, d), dtype=torch.bfloat16, device=h.device)
        if q_len == 1 and pos0 + 1 <= 128:
            # decode step: the attention kernel takes the new k / v from qkv and writes them into the caches itself
            ops.llm_attn_append(qkv, k_cache, v_cache, key_mask, nseq, pos0, w.heads, w.head_dim, w.head_dim ** -0.5, ctx)
        else:
            ops.kv_append(qkv, nseq, q_len, pos0, d, k_cache, v_cache)
            ops.llm_attn(qkv, k_cache, v_cache, key_mask, nseq, q_len, pos0, w.heads, w.head_dim, w.head_dim ** -0.5, ctx)
        gemm(ctx, lw["w_o"], lw["b_o"], residual=h, out=h)                             # h += out_proj(ctx)
        x = ops.layernorm(h, lw["ln2"][0], lw["ln2"][1], self.LN_EPS)
        f = gemm(x, lw["w_fc1"], lw["b_fc1"], act=ops.ACT_RELU)
        gemm(f, lw["w_fc2"], lw["b_fc2"], residual=h, out=h)                           # h += fc2(relu(fc1(x)))
        return h

    def _logits(self, h_last):
        w = self.w
        x = ops.layernorm(h_last, w.final_ln[0], w.final_ln[1], self.LN_EPS)
        # lm_head (N = 50272): the tiled kernel's 256-wide tiles stream it at 4.8 TB/s; K slicing would add 80 MB of partials
        return ops.gemm(x, w.lm_head, out_dtype=torch.float32)                         # fp32 [k, V]

    @torch.no_grad()
    def generate(self, hidden: torch.Tensor, selected: torch.Tensor, llm_ids: torch.Tensor, llm_mask: torch.Tensor,
                 max_new_tokens: int = 16, return_scores: bool = False,
                 forced_tokens: Optional[torch.Tensor] = None) -> GenerationOutput:
        """Greedy relation decode for the selected pairs.  The plain call replays one CUDA graph holding the whole
        prefill + decode loop (~300 launches per step; the loop has no host synchronisation); ``return_scores`` /
        ``forced_tokens`` (parity tests) and profiling runs execute eagerly."""
        if not self.use_cuda_graphs or return_scores or forced_tokens is not None or ops._profile is not None:
            return self._generate(hidden, selected, llm_ids, llm_mask, max_new_tokens, return_scores, forced_tokens)
        dev = hidden.device
        key = (tuple(hidden.shape), tuple(llm_ids.shape), int(max_new_tokens))
        e = self._graphs.get(key)
        if e is None:
            if len(self._graphs) >= 2:
                self._graphs.pop(next(iter(self._graphs)))
            e = dict(hidden=torch.empty_like(hidden),
                     selected=torch.empty(tuple(selected.shape), dtype=torch.int32, device=dev),
                     ids=torch.empty(tuple(llm_ids.shape), dtype=torch.int32, device=dev),
                     mask=torch.empty(tuple(llm_mask.shape), dtype=torch.int32, device=dev))
            for name, src in (("hidden", hidden), ("selected", selected), ("ids", llm_ids), ("mask", llm_mask)):
                e[name].copy_(src)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._generate(e["hidden"], e["selected"], e["ids"], e["mask"], max_new_tokens, False, None)
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            l0 = ops.launch_count
            with torch.cuda.graph(graph):
                out = self._generate(e["hidden"], e["selected"], e["ids"], e["mask"], max_new_tokens, False, None)
            e.update(graph=graph, out=out, launches=ops.launch_count - l0)
            self._graphs[key] = e
        if e["hidden"].data_ptr() != hidden.data_ptr():
            e["hidden"].copy_(hidden, non_blocking=True)
        e["selected"].copy_(selected, non_blocking=True)
        e["ids"].copy_(llm_ids, non_blocking=True)
        e["mask"].copy_(llm_mask, non_blocking=True)
        e["graph"].replay()
        ops._count(e["launches"])
        return e["out"]

    def _generate(self, hidden: torch.Tensor, selected: torch.Tensor, llm_ids: torch.Tensor, llm_mask: torch.Tensor,
                  max_new_tokens: int = 16, return_scores: bool = False,
                  forced_tokens: Optional[torch.Tensor] = None) -> GenerationOutput:
        # hidden bf16 [B*33, 768] (Q-Former output rows, pair-major); selected int32 [k] pair indices;
        # llm_ids / llm_mask int32 [k, T] left-padded prompt tokens (v4:260-266).  forced_tokens int32 [k, T_new]
        # teacher-forces the fed-back ids (parity tests: keeps our run on the oracle's trajectory).
        w = self.w
        dev = hidden.device
        k, T = llm_ids.shape
        d = w.d
        Tp = N_PREFIX + T
        max_ctx = Tp + max_new_tokens
        if max_ctx > MAX_CTX_KERNEL:
            raise ops._lib.OpsgError(ops._lib.OPSG_E_UNSUPPORTED, f"context {max_ctx} > {MAX_CTX_KERNEL} keys unsupported")
        llm_ids = llm_ids.to(device=dev, dtype=torch.int32).contiguous()
        llm_mask = llm_mask.to(device=dev, dtype=torch.int32).contiguous()
        selected = selected.to(device=dev, dtype=torch.int32).contiguous()

        # ---- a9: gather the selected pairs' 33 rows, project, assemble the embedded prompt -------------
        feat = ops.gather_rows(hidden, N_QUERY * hidden.shape[1], selected)            # [k, 33*768]
        proj = ops.gemm(feat.view(k * N_QUERY, hidden.shape[1]), self.w_proj, self.b_proj)      # [k*33, d]
        full_mask = torch.cat([torch.ones((k, N_PREFIX), dtype=torch.int32, device=dev), llm_mask], dim=1)   # [k, Tp]
        csum = torch.cumsum(full_mask, dim=1, dtype=torch.int32)
        pos = (csum * full_mask - 1 + 2).to(torch.int32).contiguous()                  # HF OPT :64-70 (pads -> row 1)
        n_valid = csum[:, -1].contiguous()                                             # [k]
        h = torch.empty((k, Tp, d), dtype=torch.bfloat16, device=dev)
        ops.llm_build_prefix(proj, N_QUERY, 1, N_PREFIX, w.embed, llm_ids, w.pos, pos, h)
        prefix = h.clone() if return_scores else None
        h = h.view(k * Tp, d)

        key_mask = torch.ones((k, max_ctx), dtype=torch.uint8, device=dev)
        key_mask[:, :Tp] = full_mask.to(torch.uint8)
        k_cache = torch.empty((w.n_layers, k, max_ctx, d), dtype=torch.bfloat16, device=dev)
        v_cache = torch.empty_like(k_cache)
        last_rows = (torch.arange(k, device=dev, dtype=torch.int32) * Tp + (Tp - 1)).contiguous()
        steps = torch.arange(1, max_new_tokens, device=dev, dtype=torch.int32)
        dec_pos = (n_valid[None, :] + steps[:, None] + 1).to(torch.int32).contiguous()  # [T_new-1, k]: cumsum-1+2

        tokens = torch.empty((max_new_tokens, k), dtype=torch.int32, device=dev)
        scores = torch.empty((k, max_new_tokens, w.vocab), dtype=torch.float32, device=dev) if return_scores else None

        # ---- prefill ---------------------------------------------------------------------------------
        for li, lw in enumerate(w.layers):
            self._layer(lw, h, k_cache[li], v_cache[li], key_mask, k, Tp, 0)
        logits = self._logits(ops.gather_rows(h, d, last_rows))
        ops.argmax_rows(logits, out=tokens[0])
        if scores is not None:
            scores[:, 0] = logits

        # ---- greedy decode (no host synchronisation inside the loop) -----------------------------------
        hd = torch.empty((k, d), dtype=torch.bfloat16, device=dev)
        for s in range(1, max_new_tokens):
            feed = tokens[s - 1] if forced_tokens is None else forced_tokens[:, s - 1].to(device=dev, dtype=torch.int32).contiguous()
            ops.embed_gather(w.embed, feed, hd, pos_table=w.pos, pos=dec_pos[s - 1])
            for li, lw in enumerate(w.layers):
                self._layer(lw, hd, k_cache[li], v_cache[li], key_mask, k, 1, Tp + s - 1)
            logits = self._logits(hd)
            ops.argmax_rows(logits, out=tokens[s])
            if scores is not None:
                scores[:, s] = logits
        return GenerationOutput(tokens=tokens.t().contiguous(), scores=scores, prefix=prefix)


def build_llm_engine(language_model, language_projection, device, use_cuda_graphs: bool = True):
    """Pack ``language_model`` (HF OPTForCausalLM) and ``language_projection`` (nn.Linear) for the kernels."""
    weights = PackedOPT(language_model, device)
    return OPTDecodeEngine(weights, _bf16(language_projection.weight, device), _f32(language_projection.bias, device),
                           use_cuda_graphs=use_cuda_graphs)
