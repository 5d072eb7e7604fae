"""LLM relation decode engine (rows a9-a10 of SURVEY.md §8): selected pair features -> language projection ->
embedded prompt -> batched prefill + greedy decode over a static KV cache, all on libopsg_b200 kernels.

Reference behaviour being replaced (``relation_transformer_head_v4.py:293-312``): one ``language_model.generate``
per selected pair, batch 1, HF ``DynamicCache`` — k x steps full passes over the LLM weights.  Here all k
selected pairs form ONE batch: prefill is k*(32+T) rows through every Linear (tensor-core bound), each decode
step streams the weights once for the whole batch (HBM bound).  Greedy, EOS never stops generation early: the
caller truncates at EOS when parsing, token ids past EOS are simply unused (``min_new_tokens == max_new_tokens``
in the oracle runs).

Two decoder families, selected by ``config.model_type``:

* ``llama`` — the LLM the shipped config names (``configs/psg/baseline_v4_ov.py:60-61``, ``v4:99-105``): HF
  ``models/llama/modeling_llama.py`` — RMSNorm (:52-69), rotary positions with the rotate_half convention (:137-166,
  positions ``cumsum(attention_mask) - 1`` as HF ``generate`` derives them, generation/utils.py:707-729), SwiGLU MLP
  (:170-184), no biases unless the config asks for them, untied lm_head, grouped-query attention (the k / v projection
  rows are repeated per query-head group at pack time, so the kernels always see num_heads key / value heads).
* ``opt`` — BASELINE.json's cfg3 model: HF ``models/opt/modeling_opt.py`` (pre-LN decoder :202-253, learned positions
  with offset 2 :56-70, q scaled by head_dim**-0.5 :151, ReLU FFN, final LayerNorm, tied lm_head); only
  ``do_layer_norm_before=True`` models whose ``word_embed_proj_dim == hidden_size`` (not 350M).

bf16 operands with fp32 accumulation, fp32 softmax / normalisation statistics and fp32 logits.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops
from .graphs import GraphCache

N_PREFIX = 32          # relation_query rows handed to the LLM (out[:, 1:33], v4:215)
N_QUERY = 33
MAX_CTX_KERNEL = 256   # opsg_llm_attn capacity (keys per sequence)
SUPPORTED_LLM_FAMILIES = ("llama", "opt")


def _bf16(t, device):
    return t.detach().to(device=device, dtype=torch.bfloat16).contiguous()


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _opt(t, device):
    return None if t is None else _f32(t, device)


def check_llm_supported(cfg) -> None:
    """Raise NotImplementedError with a precise reason when the decode engine cannot run this ``PretrainedConfig``;
    the head calls it at construction time (before any forward)."""
    family = getattr(cfg, "model_type", "?")
    if family not in SUPPORTED_LLM_FAMILIES:
        raise NotImplementedError(f"libopsg_b200 LLM engine supports {SUPPORTED_LLM_FAMILIES} decoders; got model_type={family!r}")
    d, heads = cfg.hidden_size, cfg.num_attention_heads
    head_dim = getattr(cfg, "head_dim", None) or d // heads
    if head_dim not in (64, 80, 128):
        raise NotImplementedError(f"head_dim {head_dim} unsupported by opsg_llm_attn (64, 80, 128)")
    if d > 8192 or d % 8:
        raise NotImplementedError(f"hidden size {d} unsupported (multiple of 8, <= 8192)")
    if heads * head_dim != d:
        raise NotImplementedError(f"num_attention_heads * head_dim = {heads * head_dim} != hidden_size {d}")
    if family == "opt":
        if not cfg.do_layer_norm_before or cfg.word_embed_proj_dim != cfg.hidden_size:
            raise NotImplementedError("OPT variants with post-LN or projected embeddings (350M) are not supported")
        if getattr(cfg, "activation_function", "relu") != "relu":
            raise NotImplementedError("OPT activation must be relu")
    else:
        if getattr(cfg, "hidden_act", "silu") != "silu":
            raise NotImplementedError("Llama hidden_act must be silu")
        rp = getattr(cfg, "rope_parameters", None) or {}
        rope_type = rp.get("rope_type", "default") if isinstance(rp, dict) else "default"
        if rope_type != "default":
            raise NotImplementedError(f"rope_type {rope_type!r} unsupported (default rotary embedding only)")
        if heads % (getattr(cfg, "num_key_value_heads", None) or heads):
            raise NotImplementedError("num_attention_heads must be a multiple of num_key_value_heads")


class PackedLLM:
    """Kernel-ready device copy of a HF causal LM (``OPTForCausalLM`` or ``LlamaForCausalLM``): bf16 matrices in
    nn.Linear layout, fp32 biases / norm weights.  Per layer: ``norm1`` -> fused ``w_qkv`` -> attention -> ``w_o`` (+res)
    -> ``norm2`` -> ``w_up`` (OPT fc1 | Llama [gate; up]) -> act -> ``w_down`` (+res)."""

    ROPE_TABLE_ROWS = 1024

    def __init__(self, language_model, device):
        cfg = language_model.config
        check_llm_supported(cfg)
        self.family = cfg.model_type
        self.d = cfg.hidden_size
        self.heads = cfg.num_attention_heads
        self.head_dim = getattr(cfg, "head_dim", None) or self.d // self.heads
        self.vocab = cfg.vocab_size
        sd = language_model.state_dict()
        if self.family == "opt":
            self._pack_opt(cfg, sd, device)
        else:
            self._pack_llama(cfg, sd, device)
        self.n_layers = len(self.layers)

    # ---- OPT ------------------------------------------------------------------------------------------------------
    def _pack_opt(self, cfg, sd, device):
        dp = "model.decoder."
        self.norm_eps = 1e-5
        self.ffn = cfg.ffn_dim
        self.embed = _bf16(sd[dp + "embed_tokens.weight"], device)            # [V, d] (also the tied lm_head)
        self.pos = _bf16(sd[dp + "embed_positions.weight"], device)           # [max_pos + 2, d]
        self.rope = None
        self.final_norm = (_f32(sd[dp + "final_layer_norm.weight"], device), _f32(sd[dp + "final_layer_norm.bias"], device))
        lm_w = sd.get("lm_head.weight")
        self.lm_head = self.embed if lm_w is None or lm_w.data_ptr() == sd[dp + "embed_tokens.weight"].data_ptr() \
            else _bf16(lm_w, device)
        self.layers = []
        n_layers = 0
        while f"{dp}layers.{n_layers}.fc1.weight" in sd:        # the head may have truncated the layer list (v4:101-103)
            n_layers += 1
        for l in range(n_layers):
            lp = f"{dp}layers.{l}."
            a = lp + "self_attn."
            self.layers.append(dict(
                norm1=(_f32(sd[lp + "self_attn_layer_norm.weight"], device), _f32(sd[lp + "self_attn_layer_norm.bias"], device)),
                w_qkv=_bf16(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0), device),
                b_qkv=_f32(torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0), device),
                w_o=_bf16(sd[a + "out_proj.weight"], device), b_o=_f32(sd[a + "out_proj.bias"], device),
                norm2=(_f32(sd[lp + "final_layer_norm.weight"], device), _f32(sd[lp + "final_layer_norm.bias"], device)),
                w_up=_bf16(sd[lp + "fc1.weight"], device), b_up=_f32(sd[lp + "fc1.bias"], device),
                w_down=_bf16(sd[lp + "fc2.weight"], device), b_down=_f32(sd[lp + "fc2.bias"], device),
            ))

    # ---- Llama ----------------------------------------------------------------------------------------------------
    def _pack_llama(self, cfg, sd, device):
        mp = "model."
        self.norm_eps = float(cfg.rms_norm_eps)
        self.ffn = cfg.intermediate_size
        self.embed = _bf16(sd[mp + "embed_tokens.weight"], device)
        self.pos = None
        self.final_norm = (_f32(sd[mp + "norm.weight"], device), None)
        lm_w = sd.get("lm_head.weight")
        self.lm_head = self.embed if lm_w is None or lm_w.data_ptr() == sd[mp + "embed_tokens.weight"].data_ptr() \
            else _bf16(lm_w, device)
        hd, H = self.head_dim, self.heads
        kvh = getattr(cfg, "num_key_value_heads", None) or H
        rep = H // kvh

        def expand_kv(w):        # [kvh*hd, ...] -> [H*hd, ...]: query head h reads key/value head h // rep (repeat_kv, :187-196)
            if rep == 1:
                return w
            return w.reshape(kvh, hd, *w.shape[1:]).repeat_interleave(rep, dim=0).reshape(H * hd, *w.shape[1:])

        # cos / sin tables exactly as LlamaRotaryEmbedding builds them (:97-135): fp32, inv_freq = theta^(-2e/hd)
        rp = getattr(cfg, "rope_parameters", None) or {}
        theta = float(rp.get("rope_theta", getattr(cfg, "rope_theta", 10000.0)) if isinstance(rp, dict) else 10000.0)
        inv_freq = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.int64).to(torch.float32) / hd))
        freqs = torch.arange(self.ROPE_TABLE_ROWS, dtype=torch.float32)[:, None] * inv_freq[None, :]      # [rows, hd/2]
        self.rope = (_f32(freqs.cos(), device), _f32(freqs.sin(), device))
        self.layers = []
        n_layers = 0
        while f"{mp}layers.{n_layers}.mlp.down_proj.weight" in sd:   # honours llm_truncate_num (v4:101-103)
            n_layers += 1
        for l in range(n_layers):
            lp = f"{mp}layers.{l}."
            a = lp + "self_attn."
            bq, bk, bv = sd.get(a + "q_proj.bias"), sd.get(a + "k_proj.bias"), sd.get(a + "v_proj.bias")
            b_qkv = None
            if bq is not None:
                b_qkv = _f32(torch.cat([bq, expand_kv(bk), expand_kv(bv)], 0), device)
            bg, bu = sd.get(lp + "mlp.gate_proj.bias"), sd.get(lp + "mlp.up_proj.bias")
            self.layers.append(dict(
                norm1=(_f32(sd[lp + "input_layernorm.weight"], device), None),
                w_qkv=_bf16(torch.cat([sd[a + "q_proj.weight"], expand_kv(sd[a + "k_proj.weight"]),
                                       expand_kv(sd[a + "v_proj.weight"])], 0), device),
                b_qkv=b_qkv,
                w_o=_bf16(sd[a + "o_proj.weight"], device), b_o=_opt(sd.get(a + "o_proj.bias"), device),
                norm2=(_f32(sd[lp + "post_attention_layernorm.weight"], device), None),
                w_up=_bf16(torch.cat([sd[lp + "mlp.gate_proj.weight"], sd[lp + "mlp.up_proj.weight"]], 0), device),   # [2*ffn, d]
                b_up=None if bg is None else _f32(torch.cat([bg, bu], 0), device),
                w_down=_bf16(sd[lp + "mlp.down_proj.weight"], device), b_down=_opt(sd.get(lp + "mlp.down_proj.bias"), device),
            ))


PackedOPT = PackedLLM      # round-1 name


@dataclass
class GenerationOutput:
    tokens: torch.Tensor                   # int32 [k, T_new] greedy token ids
    scores: Optional[torch.Tensor] = None  # fp32 [k, T_new, V] next-token logits (only if return_scores)
    prefix: Optional[torch.Tensor] = None  # bf16 [k, 32+T, d] embedded prompt incl. positions (only if return_scores)


class LLMDecodeEngine:
    def __init__(self, weights: PackedLLM, w_proj: torch.Tensor, b_proj: torch.Tensor, use_cuda_graphs: bool = True):
        self.w = weights
        self.w_proj, self.b_proj = w_proj, b_proj      # language_projection (v4:97): bf16 [d_llm, 768], fp32 [d_llm]
        if w_proj.shape[0] != weights.d:
            raise ValueError(f"language_projection out_features {w_proj.shape[0]} != LLM hidden size {weights.d}")
        self.use_cuda_graphs = use_cuda_graphs
        # decode steps (rows <= 128) take the K-sliced small-M GEMM (csrc/gemm_skinny.cu); OPSG_LLM_SMALL_M=0 keeps the tiled one
        self._graphs = GraphCache(max_entries=3)        # (kind, input shapes, max_new_tokens) -> captured prefill + decode loop

    def _norm(self, x, nw, out=None):
        w = self.w
        if w.family == "opt":
            return ops.layernorm(x, nw[0], nw[1], w.norm_eps, out=out)
        return ops.rmsnorm(x, nw[0], w.norm_eps, out=out)

    # one decoder layer over `rows` = nseq * q_len token rows; h is updated in place.  rope_pos: int32 [rows] (Llama only)
    def _layer(self, lw, h, k_cache, v_cache, key_mask, nseq, q_len, pos0, rope_pos):
        w = self.w
        d = w.d
        # Decode steps (rows <= 128) stream the weights through the K-sliced kernel with the activations resident in TMEM:
        # 17 / 11 / 21 / 20 us for qkv / out / fc1 / fc2 at k = 100 inside a graph against 22 / 13 / 24 / 40 us for the tiled
        # kernel (scripts/kbench.py streamk).  Tried and dropped (profiles/r2_decode_timeline.md): L2 prefetch of the next
        # GEMM's weights from a side kernel (164 ms per cfg3 image against 121), constant-weight early streaming inside the GEMM.
        # more rows (prefill; decode steps of a batch stacked over several images): the tiled kernels, with K split for the
        # N = hidden-size Linears while they have too few output tiles for the machine (ops.gemm_medium_m)
        gemm = ops.gemm_small_m if h.shape[0] <= 128 else ops.gemm_medium_m
        x = self._norm(h, lw["norm1"])
        qkv = gemm(x, lw["w_qkv"], lw["b_qkv"])                                        # [rows, 3d]
        if w.rope is not None:
            ops.rope(qkv, 2, w.heads, w.head_dim, rope_pos, w.rope[0], w.rope[1])      # q and k rotated in place
        ctx = torch.empty((nseq * q_len, d), dtype=torch.bfloat16, device=h.device)
        if q_len == 1:
            # decode step: the attention kernel takes the new k / v from qkv and writes them into the caches itself
            ops.llm_attn_append(qkv, k_cache, v_cache, key_mask, nseq, pos0, w.heads, w.head_dim, w.head_dim ** -0.5, ctx)
        else:
            ops.kv_append(qkv, nseq, q_len, pos0, d, k_cache, v_cache)
            ops.llm_attn(qkv, k_cache, v_cache, key_mask, nseq, q_len, pos0, w.heads, w.head_dim, w.head_dim ** -0.5, ctx)
        gemm(ctx, lw["w_o"], lw["b_o"], residual=h, out=h)                             # h += out_proj(ctx)
        x = self._norm(h, lw["norm2"])
        if w.family == "opt":
            f = gemm(x, lw["w_up"], lw["b_up"], act=ops.ACT_RELU)                      # relu(fc1(x))
        else:
            f = ops.swiglu(gemm(x, lw["w_up"], lw["b_up"]), w.ffn)                     # silu(gate(x)) * up(x)
        gemm(f, lw["w_down"], lw["b_down"], residual=h, out=h)                         # h += fc2 / down_proj
        return h

    def _logits(self, h_last):
        w = self.w
        x = self._norm(h_last, w.final_norm)
        # lm_head (N = 50272): the tiled kernel's 256-wide tiles stream it at 4.8 TB/s; K slicing would add 80 MB of partials
        return ops.gemm(x, w.lm_head, out_dtype=torch.float32)                         # fp32 [k, V]

    @torch.no_grad()
    def generate(self, hidden: torch.Tensor, selected: torch.Tensor, llm_ids: torch.Tensor, llm_mask: torch.Tensor,
                 max_new_tokens: int = 16, return_scores: bool = False,
                 forced_tokens: Optional[torch.Tensor] = None) -> GenerationOutput:
        """Greedy relation decode for the selected pairs of ONE image.  A call signature seen for the second time is
        captured and from then on replayed as ONE CUDA graph holding the whole prefill + decode loop (~300 launches per
        step; the loop has no host synchronisation); first sightings, ``return_scores`` / ``forced_tokens`` (parity tests)
        and profiling runs execute eagerly."""
        if not self.use_cuda_graphs or return_scores or forced_tokens is not None or ops._profile is not None:
            return self._generate(hidden, selected, llm_ids, llm_mask, max_new_tokens, return_scores, forced_tokens)
        key = ("pairs", tuple(hidden.shape), tuple(llm_ids.shape), int(max_new_tokens))
        return self._graphed(key, dict(hidden=hidden, selected=selected, ids=llm_ids, mask=llm_mask),
                             lambda e: self._generate(e["hidden"], e["selected"], e["ids"], e["mask"], max_new_tokens, False, None))

    @torch.no_grad()
    def generate_rows(self, rows: torch.Tensor, llm_ids: torch.Tensor, llm_mask: torch.Tensor, max_new_tokens: int = 16,
                      return_scores: bool = False, forced_tokens: Optional[torch.Tensor] = None) -> GenerationOutput:
        """Greedy relation decode for K sequences whose Q-Former rows were already gathered: ``rows`` bf16 [K, 33*768] (row
        0 of every pair is dropped by the prompt assembly, v4:215).  The K sequences need not come from one image:
        ``head.forward_batch`` stacks the selected pairs of several images, so that every decode step streams the LLM
        weights once for all of them (sequences are independent: v4:293-312 runs them one by one)."""
        if not self.use_cuda_graphs or return_scores or forced_tokens is not None or ops._profile is not None:
            return self._generate_rows(rows, llm_ids, llm_mask, max_new_tokens, return_scores, forced_tokens)
        key = ("rows", tuple(rows.shape), tuple(llm_ids.shape), int(max_new_tokens))
        return self._graphed(key, dict(rows=rows, ids=llm_ids, mask=llm_mask),
                             lambda e: self._generate_rows(e["rows"], e["ids"], e["mask"], max_new_tokens, False, None))

    def _graphed(self, key, inputs: dict, run):
        """Replay (or, at the second sighting of ``key``, capture) ``run(static inputs)`` as one CUDA graph."""
        dev = next(iter(inputs.values())).device
        e = self._graphs.lookup(key)
        if e is None:
            if not self._graphs.should_capture(key):
                return run(inputs)
            e = {n: torch.empty(tuple(t.shape), dtype=t.dtype if t.dtype == torch.bfloat16 else torch.int32, device=dev)
                 for n, t in inputs.items()}
            for name, src in inputs.items():
                e[name].copy_(src)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                run(e)
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            l0 = ops.launch_count
            with torch.cuda.graph(graph):
                out = run(e)
            e.update(graph=graph, out=out, launches=ops.launch_count - l0)
            self._graphs.insert(key, e)
        for name, src in inputs.items():
            ops.copy_into(e[name], src)
        e["graph"].replay()
        ops._count(e["launches"])
        return e["out"]

    def _generate(self, hidden: torch.Tensor, selected: torch.Tensor, llm_ids: torch.Tensor, llm_mask: torch.Tensor,
                  max_new_tokens: int = 16, return_scores: bool = False,
                  forced_tokens: Optional[torch.Tensor] = None) -> GenerationOutput:
        # hidden bf16 [B*33, 768] (Q-Former output rows, pair-major); selected int32 [k] pair indices
        selected = selected.to(device=hidden.device, dtype=torch.int32).contiguous()
        rows = ops.gather_rows(hidden, N_QUERY * hidden.shape[1], selected)            # [k, 33*768]
        return self._generate_rows(rows, llm_ids, llm_mask, max_new_tokens, return_scores, forced_tokens)

    def _generate_rows(self, feat: torch.Tensor, llm_ids: torch.Tensor, llm_mask: torch.Tensor,
                       max_new_tokens: int = 16, return_scores: bool = False,
                       forced_tokens: Optional[torch.Tensor] = None) -> GenerationOutput:
        # feat bf16 [k, 33*768]: the selected pairs' Q-Former rows; llm_ids / llm_mask int32 [k, T] left-padded prompt tokens
        # (v4:260-266).  forced_tokens int32 [k, T_new] teacher-forces the fed-back ids (parity tests: keeps our run on the
        # oracle's trajectory).
        w = self.w
        dev = feat.device
        k, T = llm_ids.shape
        d = w.d
        d_q = feat.shape[1] // N_QUERY
        Tp = N_PREFIX + T
        max_ctx = Tp + max_new_tokens
        if max_ctx > MAX_CTX_KERNEL:
            raise ops._lib.OpsgError(ops._lib.OPSG_E_UNSUPPORTED, f"context {max_ctx} > {MAX_CTX_KERNEL} keys unsupported")
        llm_ids = llm_ids.to(device=dev, dtype=torch.int32).contiguous()
        llm_mask = llm_mask.to(device=dev, dtype=torch.int32).contiguous()

        # ---- a9: project the selected pairs' rows, assemble the embedded prompt -------------
        proj = ops.gemm(feat.view(k * N_QUERY, d_q), self.w_proj, self.b_proj)         # [k*33, d]
        # positions / key mask of the prompt [1 x 32 ; left-padded text mask] and of every decode step, one small kernel:
        # OPT learned positions cumsum(mask)*mask - 1 + 2 (HF opt :64-70), Llama rotary positions cumsum(mask) - 1
        # (pads -> 0; HF generation/utils.py:707-729); generated token s sits at position n_valid + s (+2 for OPT)
        lay = ops.llm_prompt_layout(llm_mask, N_PREFIX, max_new_tokens, 2 if w.family == "opt" else 0)
        h = torch.empty((k, Tp, d), dtype=torch.bfloat16, device=dev)
        ops.llm_build_prefix(proj, N_QUERY, 1, N_PREFIX, w.embed, llm_ids, w.pos, lay.pos if w.pos is not None else None, h)
        prefix = h.clone() if return_scores else None
        h = h.view(k * Tp, d)
        rope_prefill = lay.pos.view(-1) if w.rope is not None else None

        k_cache = torch.empty((w.n_layers, k, max_ctx, d), dtype=torch.bfloat16, device=dev)
        v_cache = torch.empty_like(k_cache)

        tokens = torch.empty((max_new_tokens, k), dtype=torch.int32, device=dev)
        scores = torch.empty((k, max_new_tokens, w.vocab), dtype=torch.float32, device=dev) if return_scores else None

        # ---- prefill ---------------------------------------------------------------------------------
        for li, lw in enumerate(w.layers):
            self._layer(lw, h, k_cache[li], v_cache[li], lay.key_mask, k, Tp, 0, rope_prefill)
        logits = self._logits(ops.gather_rows(h, d, lay.last_rows))
        ops.argmax_rows(logits, out=tokens[0])
        if scores is not None:
            scores[:, 0] = logits

        # ---- greedy decode (no host synchronisation inside the loop) -----------------------------------
        hd = torch.empty((k, d), dtype=torch.bfloat16, device=dev)
        for s in range(1, max_new_tokens):
            feed = tokens[s - 1] if forced_tokens is None else forced_tokens[:, s - 1].to(device=dev, dtype=torch.int32).contiguous()
            step_pos = lay.dec_pos[s - 1]
            if w.pos is not None:
                ops.embed_gather(w.embed, feed, hd, pos_table=w.pos, pos=step_pos)
            else:
                ops.embed_gather(w.embed, feed, hd)
            for li, lw in enumerate(w.layers):
                self._layer(lw, hd, k_cache[li], v_cache[li], lay.key_mask, k, 1, Tp + s - 1, step_pos)
            logits = self._logits(hd)
            ops.argmax_rows(logits, out=tokens[s])
            if scores is not None:
                scores[:, s] = logits
        out_tokens = torch.empty((k, max_new_tokens), dtype=torch.int32, device=dev)
        ops.transpose_i32(tokens, out_tokens)
        return GenerationOutput(tokens=out_tokens, scores=scores, prefix=prefix)


OPTDecodeEngine = LLMDecodeEngine     # round-1 name


def build_llm_engine(language_model, language_projection, device, use_cuda_graphs: bool = True):
    """Pack ``language_model`` (HF OPTForCausalLM / LlamaForCausalLM) and ``language_projection`` (nn.Linear)."""
    weights = PackedLLM(language_model, device)
    return LLMDecodeEngine(weights, _bf16(language_projection.weight, device), _f32(language_projection.bias, device),
                           use_cuda_graphs=use_cuda_graphs)
