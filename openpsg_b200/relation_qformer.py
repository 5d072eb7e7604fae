"""RelationQueryTransformer: the N^2 pair-query Q-Former + existence filter on libopsg_b200 kernels.

Host-side orchestration of rows a3-a8 of SURVEY.md §8 for ONE image: what the reference does at
``relation_transformer_head_v4.py:155-209,236-237,408-435`` by calling HF's
``InstructBlipQFormerModel`` (modeling_instructblip.py:634-938) on N^2 expanded copies.  Here:

  * pair masks stay as N x ceil(L/32) bit words (K2) and are OR-ed inside the cross-attention kernel;
  * image tokens are projected to K / V^T ONCE per image per layer (K3), not once per pair;
  * all pairs' rows are stacked along M: rows [0, B*33) are query rows (pair-major), rows
    [B*33, B*33 + B*T) are instruction-text rows, so every Linear is one tcgen05 GEMM (K6);
  * the last layer only computes what ``[:, :33]`` (v4:185) can observe.

Every arithmetic step is a C-ABI call (``openpsg_b200.ops``); torch only owns the buffers.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import ops
from .graphs import GraphCache

NUM_HEADS = 12
HEAD_DIM = 64
N_QUERY = 33
MAX_TILE_KEYS = 256      # image tokens the tcgen05 cross-attention kernels keep in one score tile (more: online-softmax kernel)
LN_EPS = 1e-12


def _bf16(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.bfloat16).contiguous()


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class PackedQFormer:
    """Device-resident, kernel-ready copy of the head's relation-query weights (bf16 matrices, fp32
    biases / LayerNorm / embedding tables).  Built from a state dict with the reference's parameter
    names (SURVEY.md §5 checkpoint contract)."""

    def __init__(self, sd: Dict[str, torch.Tensor], device, num_layers: int = 2, patch: int = 16,
                 prefix: str = "relation_qformer."):
        self.device = device
        self.num_layers = num_layers
        self.patch = patch
        d = sd["rel_cls_query"].shape[-1]
        self.d = d
        self.query = _f32(torch.cat([sd["rel_cls_query"], sd["relation_query"]], dim=1)[0], device)      # [33,d]
        pw = sd["patch_embed.proj.weight"]
        self.patch_w = _bf16(pw.reshape(pw.shape[0], -1), device)                                           # [256, 65536]
        self.patch_b = _f32(sd["patch_embed.proj.bias"], device)
        e = prefix + "embeddings."
        self.word_emb = _f32(sd[e + "word_embeddings.weight"], device)
        self.pos_emb = _f32(sd[e + "position_embeddings.weight"], device)
        self.emb_ln = (_f32(sd[e + "layernorm.weight"], device), _f32(sd[e + "layernorm.bias"], device))
        self.layers = []
        for l in range(num_layers):
            lp = f"{prefix}encoder.layer.{l}."

            def W(name):
                return _bf16(sd[lp + name + ".weight"], device)

            def Bv(name):
                return _f32(sd[lp + name + ".bias"], device)

            def LN(name):
                return (_f32(sd[lp + name + ".weight"], device), _f32(sd[lp + name + ".bias"], device))

            a = "attention.attention."
            c = "crossattention.attention."
            layer = dict(
                w_qkv=_bf16(torch.cat([sd[lp + a + "query.weight"], sd[lp + a + "key.weight"], sd[lp + a + "value.weight"]], 0), device),
                b_qkv=_f32(torch.cat([sd[lp + a + "query.bias"], sd[lp + a + "key.bias"], sd[lp + a + "value.bias"]], 0), device),
                w_o=W("attention.output.dense"), b_o=Bv("attention.output.dense"), ln_self=LN("attention.output.LayerNorm"),
                w_cq=W(c + "query"), b_cq=Bv(c + "query"),
                w_ck=W(c + "key"), b_ck=Bv(c + "key"), w_cv=W(c + "value"), b_cv=Bv(c + "value"),
                w_co=W("crossattention.output.dense"), b_co=Bv("crossattention.output.dense"),
                ln_cross=LN("crossattention.output.LayerNorm"),
                w_iq=W("intermediate_query.dense"), b_iq=Bv("intermediate_query.dense"),
                w_oq=W("output_query.dense"), b_oq=Bv("output_query.dense"), ln_q=LN("output_query.LayerNorm"),
                w_it=W("intermediate.dense"), b_it=Bv("intermediate.dense"),
                w_ot=W("output.dense"), b_ot=Bv("output.dense"), ln_t=LN("output.LayerNorm"),
            )
            self.layers.append(layer)
        # LayerNorm folding (opsg_gemm_bf16_ln): consumer weights with the pending LayerNorm's gamma folded in, their
        # row sums, and biases that already contain W . beta.  Layer l's qkv consumes the previous layer's output
        # LayerNorms (query rows: output_query.LayerNorm, text rows: output.LayerNorm).
        def fold(wname, bname, gamma, beta, lp):
            W = sd[lp + wname].detach().float()
            b = sd[lp + bname].detach().float()
            Wf = (W * gamma.detach().float()[None, :]).to(torch.bfloat16)
            return (Wf.to(device).contiguous(), _f32(Wf.float().sum(1), device), _f32(b + W @ beta.detach().float(), device))

        for l, layer in enumerate(self.layers):
            lp = f"{prefix}encoder.layer.{l}."
            g_self, b_self = sd[lp + "attention.output.LayerNorm.weight"], sd[lp + "attention.output.LayerNorm.bias"]
            g_cross, b_cross = sd[lp + "crossattention.output.LayerNorm.weight"], sd[lp + "crossattention.output.LayerNorm.bias"]
            layer["f_cq"] = fold("crossattention.attention.query.weight", "crossattention.attention.query.bias", g_self, b_self, lp)
            layer["f_iq"] = fold("intermediate_query.dense.weight", "intermediate_query.dense.bias", g_cross, b_cross, lp)
            layer["f_it"] = fold("intermediate.dense.weight", "intermediate.dense.bias", g_self, b_self, lp)
            if l > 0:
                pp = f"{prefix}encoder.layer.{l - 1}."
                a = "attention.attention."
                Wqkv = torch.cat([sd[lp + a + "query.weight"], sd[lp + a + "key.weight"], sd[lp + a + "value.weight"]], 0).detach().float()
                bqkv = torch.cat([sd[lp + a + "query.bias"], sd[lp + a + "key.bias"], sd[lp + a + "value.bias"]], 0).detach().float()
                for tag, ln in (("f_qkv_q", "output_query.LayerNorm"), ("f_qkv_t", "output.LayerNorm")):
                    gamma, beta = sd[pp + ln + ".weight"].detach().float(), sd[pp + ln + ".bias"].detach().float()
                    Wf = (Wqkv * gamma[None, :]).to(torch.bfloat16)
                    layer[tag] = (Wf.to(device).contiguous(), _f32(Wf.float().sum(1), device), _f32(bqkv + Wqkv @ beta, device))
        self.exist_w = _f32(sd["binary_rel_cls_pred.weight"].reshape(-1), device)
        self.exist_b = _f32(sd["binary_rel_cls_pred.bias"].reshape(-1), device)


@dataclass
class RelationQueryOutput:
    hidden: torch.Tensor            # bf16 [B*33, d]: last_hidden_state[:, :33] stacked pair-major; with selected_rows_only
                                    # bf16 [k*33, d]: the 33 rows of the k pairs in ``topk``, in that order (hidden_pairs = k)
    logits: torch.Tensor            # fp32 [B]
    probs: torch.Tensor             # fp32 [B]
    exist_mask: torch.Tensor        # uint8 [B]
    topk: torch.Tensor              # int32 [k]
    mask_bits: torch.Tensor         # int32 [N, words]
    image_tokens: torch.Tensor      # bf16 [L, 256]
    intermediates: Optional[dict] = None
    hidden_pairs: Optional[int] = None   # None: ``hidden`` holds every pair; k: only the pairs of ``topk`` (selected_rows_only)

    def clone(self) -> "RelationQueryOutput":
        """Private copy of a result that lives in a CUDA graph's static buffers (overwritten by the next replay)."""
        return RelationQueryOutput(**{f: (getattr(self, f).clone() if torch.is_tensor(getattr(self, f)) else getattr(self, f))
                                      for f in self.__dataclass_fields__})


class GraphedRelationQuery:
    """CUDA-graph replay of ``RelationQueryTransformer.forward`` per input-shape signature.

    The per-image pipeline is ~100 launches of 5-500 us kernels; replaying it as one graph removes the launch gaps
    (measured 11 % of the step on cfg2) without touching the arithmetic.  Inputs are copied into static buffers
    (``copy_`` also performs the host->device transfer / dtype conversion when the caller's tensors live on the
    host), outputs are the static tensors of the captured run: they are overwritten by the next call with the same
    signature, so callers that keep results across calls must clone them (``RelationQueryOutput.clone()``)."""

    def __init__(self, engine: "RelationQueryTransformer", max_entries: int = 4, capture_after: int = 2):
        self.engine = engine
        self.cache = GraphCache(max_entries=max_entries, capture_after=capture_after)

    @property
    def entries(self):
        return self.cache.entries

    def run(self, feat, pan, img_hw, pad_hw, obj_ids, input_ids, text_mask, *, topk: int, threshold: float,
            selected_rows_only: bool = False):
        """A signature is replayed from its graph once it has been seen ``capture_after`` times; before that (and for
        signatures that never repeat, the common case on real PSG images) it runs eagerly through the same kernels."""
        dev = self.engine.w.device
        key = (tuple(feat.shape), tuple(pan.shape), tuple(int(x) for x in img_hw), tuple(int(x) for x in pad_hw),
               int(obj_ids.numel()), tuple(input_ids.shape), int(topk), float(threshold), bool(selected_rows_only))
        e = self.cache.lookup(key)
        if e is None:
            if not self.cache.should_capture(key):
                return self.engine.forward(self._dev(feat, torch.float32, dev), self._dev(pan, torch.int32, dev), img_hw, pad_hw,
                                           self._dev(obj_ids, torch.int32, dev), self._dev(input_ids, torch.int32, dev),
                                           self._dev(text_mask, torch.int32, dev), topk=topk, threshold=threshold,
                                           selected_rows_only=selected_rows_only)
            e = self._capture(key, feat, pan, img_hw, pad_hw, obj_ids, input_ids, text_mask, topk, threshold, dev,
                              selected_rows_only)
            self.cache.insert(key, e)
        for name, src in (("feat", feat), ("pan", pan), ("obj_ids", obj_ids), ("input_ids", input_ids), ("text_mask", text_mask)):
            ops.copy_into(e[name], src)
        e["graph"].replay()
        ops._count(e["launches"])
        return e["out"]

    @staticmethod
    def _dev(t, dtype, dev):
        return t.to(device=dev, dtype=dtype, non_blocking=True).contiguous()

    def _capture(self, key, feat, pan, img_hw, pad_hw, obj_ids, input_ids, text_mask, topk, threshold, dev,
                 selected_rows_only=False):
        st = dict(
            feat=torch.empty(tuple(feat.shape), dtype=torch.float32, device=dev),
            pan=torch.empty(tuple(pan.shape), dtype=torch.int32, device=dev),
            obj_ids=torch.empty(tuple(obj_ids.shape), dtype=torch.int32, device=dev),
            input_ids=torch.empty(tuple(input_ids.shape), dtype=torch.int32, device=dev),
            text_mask=torch.empty(tuple(text_mask.shape), dtype=torch.int32, device=dev),
        )
        for name, src in (("feat", feat), ("pan", pan), ("obj_ids", obj_ids), ("input_ids", input_ids), ("text_mask", text_mask)):
            st[name].copy_(src)

        def run():
            return self.engine.forward(st["feat"], st["pan"], img_hw, pad_hw, st["obj_ids"], st["input_ids"], st["text_mask"],
                                       topk=topk, threshold=threshold, selected_rows_only=selected_rows_only)

        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up outside capture: first-use cudaFuncSetAttribute calls, allocator
            run()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        l0 = ops.launch_count
        with torch.cuda.graph(graph):
            out = run()
        st.update(graph=graph, out=out, launches=ops.launch_count - l0)
        return st


class RelationQueryTransformer:
    def __init__(self, weights: PackedQFormer, fold_ln: Optional[bool] = None):
        self.w = weights
        # LayerNorm folding (opt-in, OPSG_FOLD_LN=1): activations stay un-normalised with per-row statistics and consumers
        # apply the pending LayerNorm in their GEMM epilogue (7 of the 8 LayerNorm passes per image disappear).  Measured
        # +0.6 % pairs/s only (the epilogue pays for it) and the fp32 atomics behind the row statistics make runs differ
        # in the last bf16 bit, so the default keeps the separate LayerNorm kernel.
        self.fold_ln = (os.environ.get("OPSG_FOLD_LN", "0") == "1") if fold_ln is None else bool(fold_ln)
        # layer 0 projects the (pair-independent) query rows once instead of B times; OPSG_SHARE_QUERY_ROWS=0 turns it off
        self.share_query_rows = os.environ.get("OPSG_SHARE_QUERY_ROWS", "1") != "0"
        # PatchEmbed's split-K is deterministic by default (per-split partial slices + fixed-order reduction);
        # OPSG_PATCH_DETERMINISTIC=0 selects the round-1 fp32-atomics accumulation (run-to-run differences in the last bit)
        self.deterministic_patch_embed = os.environ.get("OPSG_PATCH_DETERMINISTIC", "1") != "0"
        self._streams = {}          # device index -> (mask-chain stream, embedding stream)

    def _side_streams(self, dev):
        """Two side streams per device for the independent chains at the head of an image (see ``forward``).  Blocks they
        allocate return to their own pools and are only reused after the next forward's ``wait_stream(current)``, i.e. after
        everything the current stream did with them."""
        key = torch.device(dev).index
        st = self._streams.get(key)
        if st is None:
            st = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
            self._streams[key] = st
        return st

    # -- K1 ------------------------------------------------------------------------------------------
    def image_tokens(self, feat: torch.Tensor) -> torch.Tensor:
        """timm PatchEmbed (v4:410): Conv2d(k=s=16) as im2col + split-K tcgen05 GEMM.  feat fp32 [C,h,w]."""
        w = self.w
        a = ops.patch_im2col(feat, w.patch)                                  # bf16 [L, C*p*p]
        L, K = a.shape
        kb = K // 64
        m_tiles = (L + 127) // 128
        splits = max(1, min(kb, 148 // max(1, m_tiles)))
        if self.deterministic_patch_embed:
            # default: every K split writes its own fp32 partial slice, summed in split order with the bias by a second
            # kernel (ops.gemm_splitk) -- two runs of the product return the same bits, hence the same top-k set
            return ops.gemm_splitk(a, w.patch_w, w.patch_b, splits)
        acc = torch.empty((L, w.patch_w.shape[0]), dtype=torch.float32, device=feat.device)
        ops.init_rows_f32(acc, w.patch_b)
        ops.gemm(a, w.patch_w, out=acc, atomic=True, k_splits=splits)
        return ops.cast_f32_bf16(acc)

    @staticmethod
    def _vt_buffer(d, L, Lp, dev):
        # V^T rows are padded to a multiple of 8 tokens for the TMA box; the pad columns must read as zero
        if Lp == L:
            return torch.empty((d, Lp), dtype=torch.bfloat16, device=dev)
        return torch.zeros((d, Lp), dtype=torch.bfloat16, device=dev)

    # -- a3..a8 --------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, feat: torch.Tensor, pan: torch.Tensor, img_hw, pad_hw, obj_ids: torch.Tensor,
                input_ids: torch.Tensor, text_mask: torch.Tensor, *, topk: int = 20, threshold: float = 0.5,
                pair_index: Optional[torch.Tensor] = None, keep_intermediates: bool = False,
                selected_rows_only: bool = False) -> RelationQueryOutput:
        """feat fp32 [C,h,w]; pan int32 [Hp,Wp]; obj_ids int32 [N]; input_ids/text_mask int32 [B,T]
        (B = N*N unless pair_index int32 [B] selects a subset of pairs).

        ``selected_rows_only``: what the reference consumes of the last layer is row 0 of every pair (existence logit,
        v4:206-209) and rows 1..32 of the k SELECTED pairs (v4:215 gathers them for the LLM); every step after the last
        layer's self-attention is row-wise, so the last layer then runs on B rows (row 0 of every pair) up to the top-k and
        on k x 33 rows afterwards instead of on B x 33 rows: -34 % of the image's FLOPs.  Logits, mask, top-k and the selected
        pairs' rows are the same values; ``hidden`` then holds only the selected pairs (see RelationQueryOutput)."""
        w = self.w
        dev = feat.device
        d = w.d
        N = obj_ids.numel()
        B, T = input_ids.shape
        th, tw = feat.shape[-2] // w.patch, feat.shape[-1] // w.patch
        L = th * tw
        inter = {} if keep_intermediates else None
        if self.fold_ln:
            return self._forward_folded(feat, pan, img_hw, pad_hw, obj_ids, input_ids, text_mask, topk, threshold, pair_index,
                                        inter)

        # The three chains at the head of an image are independent and each too small to fill the machine (40 / 1 / 413 CTAs of
        # mask work, the PatchEmbed GEMM, 3 k CTAs of embeddings): they run as three branches -- two side streams forked off
        # the current one and joined before their results are used (inside a CUDA graph: three parallel branches).  Serial they
        # cost ~130 us per cfg2 image, the longest branch ~55 us.
        cur = torch.cuda.current_stream(dev)
        s_mask, s_embed = self._side_streams(dev)
        s_mask.wait_stream(cur)
        s_embed.wait_stream(cur)
        with torch.cuda.stream(s_mask):
            bits = ops.pair_mask_bits(pan, img_hw, pad_hw, (th, tw), obj_ids)   # K2
            # K5 sees the image tokens sorted by owning object (attention does not depend on the key order; contiguous
            # objects make most 16-key chunks invisible to a group of pair queries, whose exponentials are then skipped)
            if L <= MAX_TILE_KEYS:
                perm, bits_k = ops.token_order(bits, L)
                bias_tiles = ops.xattn_bias_tiles(bits_k, N, B, N_QUERY, L, pair_index)  # K5 mask operand tiles (both layers)
            else:
                # more image tokens than one tensor-memory score tile holds: K5 runs its online-softmax kernel straight from
                # the mask bits (csrc/xattn_pairs_long.cu); no key reordering, no operand tiles
                perm, bits_k, bias_tiles = None, bits, False
        with torch.cuda.stream(s_embed):
            h = ops.qformer_embed_ln(w.query, input_ids, w.word_emb, w.pos_emb, w.emb_ln[0], w.emb_ln[1], LN_EPS)   # K7
        X = self.image_tokens(feat)                                              # K1  [L,256]
        cur.wait_stream(s_mask)
        Xk = X if perm is None else ops.gather_rows(X, X.shape[1], perm)         # key-order copy for the K / V projections
        cur.wait_stream(s_embed)
        RQ = B * N_QUERY
        if inter is not None:
            inter["embeddings"] = h
        Lp = (L + 7) // 8 * 8
        for li, lw in enumerate(w.layers):
            last = li == len(w.layers) - 1
            # K3: per-image K and V^T for this layer's cross-attention (shared by all pairs)
            kc = ops.gemm(Xk, lw["w_ck"], lw["b_ck"])                            # [L, d]
            vt = self._vt_buffer(d, L, Lp, dev)
            ops.gemm(lw["w_cv"], Xk, lw["b_cv"], bias_along_m=True, out=vt[:, :L])  # V^T [d, L]
            # self-attention over the 33 + T rows of every pair
            if li == 0 and self.share_query_rows and T > 0:
                # Layer 0: the 33 query rows of every pair are the same LN(query tokens) (v4:158-159 expands them B times
                # and HF projects every copy), so their q/k/v are projected once and only the text rows go through the
                # big GEMM; the attention kernel reads the query rows from the shared table.
                qkv_q = ops.gemm(h[:N_QUERY], lw["w_qkv"], lw["b_qkv"])          # [33, 3d]
                qkv = torch.empty((h.shape[0], 3 * d), dtype=torch.bfloat16, device=dev)
                ops.gemm(h[RQ:], lw["w_qkv"], lw["b_qkv"], out=qkv[RQ:])          # text rows only
                ctx = ops.self_attn_small(qkv, text_mask, B, N_QUERY, T, NUM_HEADS, HEAD_DIM, text_queries=not last,
                                          shared_query_qkv=qkv_q)
            else:
                qkv = ops.gemm(h, lw["w_qkv"], lw["b_qkv"])                      # [R, 3d]
                ctx = ops.self_attn_small(qkv, text_mask, B, N_QUERY, T, NUM_HEADS, HEAD_DIM, text_queries=not last)
            if last and selected_rows_only and inter is None:
                return self._last_layer_selected(lw, ctx, h, kc, vt, bits_k, bits, X, N, B, L, pair_index, topk, threshold)
            rows = ctx.shape[0]                                                  # R, or B*33 on the last layer
            pre = ops.gemm(ctx, lw["w_o"], lw["b_o"], residual=h[:rows])
            h1 = ops.layernorm(pre, lw["ln_self"][0], lw["ln_self"][1], LN_EPS)
            hq = h1[:RQ]
            # K5: masked pair x image cross-attention on the query rows
            qc = ops.gemm(hq, lw["w_cq"], lw["b_cq"])
            cx = ops.xattn_pairs(qc, kc, vt, bits_k, N, B, N_QUERY, L, NUM_HEADS, HEAD_DIM, pair_index=pair_index,
                                 bias_tiles=bias_tiles)
            pre = ops.gemm(cx, lw["w_co"], lw["b_co"], residual=hq)
            hq2 = ops.layernorm(pre, lw["ln_cross"][0], lw["ln_cross"][1], LN_EPS)
            # FFN (query rows; text rows only where a later layer can still see them)
            h_next = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
            f = ops.gemm(hq2, lw["w_iq"], lw["b_iq"], act=ops.ACT_GELU)
            pre = ops.gemm(f, lw["w_oq"], lw["b_oq"], residual=hq2)
            ops.layernorm(pre, lw["ln_q"][0], lw["ln_q"][1], LN_EPS, out=h_next[:RQ])
            if not last and T > 0:
                ht = h1[RQ:]
                f = ops.gemm(ht, lw["w_it"], lw["b_it"], act=ops.ACT_GELU)
                pre = ops.gemm(f, lw["w_ot"], lw["b_ot"], residual=ht)
                ops.layernorm(pre, lw["ln_t"][0], lw["ln_t"][1], LN_EPS, out=h_next[RQ:])
            if inter is not None:
                inter[f"l{li}.self"] = h1
                inter[f"l{li}.xattn_q"] = qc
                inter[f"l{li}.xattn_ctx"] = cx
                inter[f"l{li}.cross"] = hq2
                inter[f"l{li}.out"] = h_next
                inter[f"l{li}.k"] = kc
                inter[f"l{li}.vt"] = vt
            h = h_next
        out = h[:RQ]
        logits, probs, mask, top = ops.exist_filter_topk(out, N_QUERY * d, B, d, w.exist_w, w.exist_b, threshold,
                                                         min(topk, B))           # K8
        return RelationQueryOutput(hidden=out, logits=logits, probs=probs, exist_mask=mask, topk=top, mask_bits=bits,
                                   image_tokens=X, intermediates=inter)

    def _query_rows_tail(self, lw, ctx, res, kc, vt, bits_k, N, n_pairs, n_query, L, pair_index):
        """Everything of a layer after self-attention for a set of query rows (row-wise ops + K5): ctx / res bf16
        [n_pairs * n_query, d] (row stride free), returns the layer output rows, contiguous."""
        pre = ops.gemm(ctx, lw["w_o"], lw["b_o"], residual=res)
        h1 = ops.layernorm(pre, lw["ln_self"][0], lw["ln_self"][1], LN_EPS)
        qc = ops.gemm(h1, lw["w_cq"], lw["b_cq"])
        # (the mask-bias tiles were built for 33 rows of every pair: these two calls take the kernel with in-kernel masks)
        cx = ops.xattn_pairs(qc, kc, vt, bits_k, N, n_pairs, n_query, L, NUM_HEADS, HEAD_DIM, pair_index=pair_index,
                             bias_tiles=False)
        pre = ops.gemm(cx, lw["w_co"], lw["b_co"], residual=h1)
        hq2 = ops.layernorm(pre, lw["ln_cross"][0], lw["ln_cross"][1], LN_EPS)
        f = ops.gemm(hq2, lw["w_iq"], lw["b_iq"], act=ops.ACT_GELU)
        pre = ops.gemm(f, lw["w_oq"], lw["b_oq"], residual=hq2)
        return ops.layernorm(pre, lw["ln_q"][0], lw["ln_q"][1], LN_EPS)

    def _last_layer_selected(self, lw, ctx, h, kc, vt, bits_k, bits, X, N, B, L, pair_index, topk, threshold):
        """Last layer after its self-attention, only for the rows the reference consumes (``selected_rows_only``)."""
        w = self.w
        d = w.d
        RQ = B * N_QUERY
        # row 0 of every pair (strided views: row p of the operand = row 33 p of ctx / h) -> existence logits -> top-k
        out0 = self._query_rows_tail(lw, ctx.view(B, N_QUERY, d)[:, 0], h[:RQ].view(B, N_QUERY, d)[:, 0], kc, vt, bits_k, N, B, 1,
                                     L, pair_index)
        k = min(topk, B)
        logits, probs, mask, top = ops.exist_filter_topk(out0, d, B, d, w.exist_w, w.exist_b, threshold, k)      # K8
        # the 33 rows of the k selected pairs
        if k > 0:
            ctx_sel = ops.gather_rows(ctx, N_QUERY * d, top).view(k * N_QUERY, d)
            h_sel = ops.gather_rows(h, N_QUERY * d, top).view(k * N_QUERY, d)
            pairs = top if pair_index is None else pair_index[top.long()].contiguous()
            hidden = self._query_rows_tail(lw, ctx_sel, h_sel, kc, vt, bits_k, N, k, N_QUERY, L, pairs)
        else:
            hidden = torch.empty((0, d), dtype=torch.bfloat16, device=ctx.device)
        return RelationQueryOutput(hidden=hidden, logits=logits, probs=probs, exist_mask=mask, topk=top, mask_bits=bits,
                                   image_tokens=X, intermediates=None, hidden_pairs=k)

    # -- a3..a8 with LayerNorm folding ---------------------------------------------------------------------------------
    def _forward_folded(self, feat, pan, img_hw, pad_hw, obj_ids, input_ids, text_mask, topk, threshold, pair_index, inter):
        """Same arithmetic as ``forward`` with every LayerNorm except the last applied inside the consuming GEMM
        (``ops.gemm_ln``): ``pre*`` tensors are the un-normalised Linear + residual outputs, ``st*`` their per-row
        (sum, sum of squares)."""
        w = self.w
        dev = feat.device
        d = w.d
        N = obj_ids.numel()
        B, T = input_ids.shape
        th, tw = feat.shape[-2] // w.patch, feat.shape[-1] // w.patch
        L = th * tw
        bits = ops.pair_mask_bits(pan, img_hw, pad_hw, (th, tw), obj_ids)
        perm, bits_k = ops.token_order(bits, L)
        bias_tiles = ops.xattn_bias_tiles(bits_k, N, B, N_QUERY, L, pair_index)
        X = self.image_tokens(feat)
        Xk = ops.gather_rows(X, X.shape[1], perm)
        h = ops.qformer_embed_ln(w.query, input_ids, w.word_emb, w.pos_emb, w.emb_ln[0], w.emb_ln[1], LN_EPS)
        RQ = B * N_QUERY
        if inter is not None:
            inter["embeddings"] = h
        Lp = (L + 7) // 8 * 8
        h_st = None                  # statistics of h when h is un-normalised (layers >= 1)
        pend_q = pend_t = None       # pending LayerNorm (gamma, beta) of the query / text rows of h

        def zeros_stats(rows):
            return torch.zeros((rows, 2), dtype=torch.float32, device=dev)

        for li, lw in enumerate(w.layers):
            last = li == len(w.layers) - 1
            kc = ops.gemm(Xk, lw["w_ck"], lw["b_ck"])
            vt = self._vt_buffer(d, L, Lp, dev)
            ops.gemm(lw["w_cv"], Xk, lw["b_cv"], bias_along_m=True, out=vt[:, :L])
            # ---- self-attention block ----
            qkv_q = None
            if h_st is None and self.share_query_rows and T > 0:
                qkv_q = ops.gemm(h[:N_QUERY], lw["w_qkv"], lw["b_qkv"])          # layer 0: query rows projected once
                qkv = torch.empty((h.shape[0], 3 * d), dtype=torch.bfloat16, device=dev)
                ops.gemm(h[RQ:], lw["w_qkv"], lw["b_qkv"], out=qkv[RQ:])
            elif h_st is None:
                qkv = ops.gemm(h, lw["w_qkv"], lw["b_qkv"])
            else:
                qkv = torch.empty((h.shape[0], 3 * d), dtype=torch.bfloat16, device=dev)
                Wq, cq_, bq_ = lw["f_qkv_q"]
                ops.gemm_ln(h[:RQ], Wq, bq_, a_stats=h_st[:RQ], a_colsum=cq_, out=qkv[:RQ], eps=LN_EPS)
                if h.shape[0] > RQ:
                    Wt, ct_, bt_ = lw["f_qkv_t"]
                    ops.gemm_ln(h[RQ:], Wt, bt_, a_stats=h_st[RQ:], a_colsum=ct_, out=qkv[RQ:], eps=LN_EPS)
            ctx = ops.self_attn_small(qkv, text_mask, B, N_QUERY, T, NUM_HEADS, HEAD_DIM, text_queries=not last,
                                      shared_query_qkv=qkv_q)
            rows = ctx.shape[0]
            pre1 = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
            st1 = zeros_stats(rows)
            if h_st is None:
                ops.gemm_ln(ctx, lw["w_o"], lw["b_o"], residual=h[:rows], stats_out=st1, out=pre1, eps=LN_EPS)
            else:
                ops.gemm_ln(ctx[:RQ], lw["w_o"], lw["b_o"], residual=h[:RQ], r_stats=h_st[:RQ], r_gamma=pend_q[0],
                            r_beta=pend_q[1], stats_out=st1[:RQ], out=pre1[:RQ], eps=LN_EPS)
                if rows > RQ:
                    ops.gemm_ln(ctx[RQ:], lw["w_o"], lw["b_o"], residual=h[RQ:rows], r_stats=h_st[RQ:rows], r_gamma=pend_t[0],
                                r_beta=pend_t[1], stats_out=st1[RQ:], out=pre1[RQ:], eps=LN_EPS)
            g_self, b_self = lw["ln_self"]
            # ---- cross-attention block (query rows) ----
            Wc, cc_, bc_ = lw["f_cq"]
            qc = ops.gemm_ln(pre1[:RQ], Wc, bc_, a_stats=st1[:RQ], a_colsum=cc_, eps=LN_EPS)
            cx = ops.xattn_pairs(qc, kc, vt, bits_k, N, B, N_QUERY, L, NUM_HEADS, HEAD_DIM, pair_index=pair_index,
                                 bias_tiles=bias_tiles)
            st2 = zeros_stats(RQ)
            pre2 = ops.gemm_ln(cx, lw["w_co"], lw["b_co"], residual=pre1[:RQ], r_stats=st1[:RQ], r_gamma=g_self, r_beta=b_self,
                               stats_out=st2, eps=LN_EPS)
            g_cross, b_cross = lw["ln_cross"]
            # ---- FFN (query rows; text rows only where a later layer can still see them) ----
            Wi, ci_, bi_ = lw["f_iq"]
            f = ops.gemm_ln(pre2, Wi, bi_, act=ops.ACT_GELU, a_stats=st2, a_colsum=ci_, eps=LN_EPS)
            pre3 = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
            st3 = zeros_stats(rows)
            ops.gemm_ln(f, lw["w_oq"], lw["b_oq"], residual=pre2, r_stats=st2, r_gamma=g_cross, r_beta=b_cross,
                        stats_out=st3[:RQ], out=pre3[:RQ], eps=LN_EPS)
            if not last and T > 0:
                Wt, ct_, bt_ = lw["f_it"]
                ft = ops.gemm_ln(pre1[RQ:], Wt, bt_, act=ops.ACT_GELU, a_stats=st1[RQ:], a_colsum=ct_, eps=LN_EPS)
                ops.gemm_ln(ft, lw["w_ot"], lw["b_ot"], residual=pre1[RQ:], r_stats=st1[RQ:], r_gamma=g_self, r_beta=b_self,
                            stats_out=st3[RQ:], out=pre3[RQ:], eps=LN_EPS)
            if inter is not None:       # materialise the virtual LayerNorm outputs for stage-by-stage parity tests
                h1 = ops.layernorm(pre1, g_self, b_self, LN_EPS)
                h_next = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
                ops.layernorm(pre3[:RQ].contiguous(), lw["ln_q"][0], lw["ln_q"][1], LN_EPS, out=h_next[:RQ])
                if rows > RQ:
                    ops.layernorm(pre3[RQ:].contiguous(), lw["ln_t"][0], lw["ln_t"][1], LN_EPS, out=h_next[RQ:])
                inter[f"l{li}.self"] = h1
                inter[f"l{li}.xattn_q"] = qc
                inter[f"l{li}.xattn_ctx"] = cx
                inter[f"l{li}.cross"] = ops.layernorm(pre2, g_cross, b_cross, LN_EPS)
                inter[f"l{li}.out"] = h_next
                inter[f"l{li}.k"] = kc
                inter[f"l{li}.vt"] = vt
            h, h_st, pend_q, pend_t = pre3, st3, lw["ln_q"], lw["ln_t"]
        out = ops.layernorm(h[:RQ], pend_q[0], pend_q[1], LN_EPS)                 # the one explicit LayerNorm left
        logits, probs, mask, top = ops.exist_filter_topk(out, N_QUERY * d, B, d, w.exist_w, w.exist_b, threshold, min(topk, B))
        return RelationQueryOutput(hidden=out, logits=logits, probs=probs, exist_mask=mask, topk=top, mask_bits=bits,
                                   image_tokens=X, intermediates=inter)
