"""TEST INFRASTRUCTURE — CPU oracle for the relation-head hot path.  Never imported by the product.

An independent fp32 restatement (torch CPU tensors / numpy integers) of what the reference computes
on the inference path of ``RelationTransformerHeadV4`` — control flow from
``kings_sgg/models/relation_heads/relation_transformer_head_v4.py:134-326,408-435`` and arithmetic
from the un-vendored dependency the reference calls, HuggingFace ``transformers`` (no version pinned
by the reference; the installed 5.5.0 is what we pin): ``models/instructblip/modeling_instructblip.py``
(MHA :486-538, SelfOutput :542-553, Layer :634-695, Embeddings :753-782, masks :822-863) and
``modeling_utils.py:858-878`` (finfo.min encoder mask), ``models/opt/modeling_opt.py`` for the LLM.

Parity pin: the reference ships no tests or golden vectors, so this restatement is
pinned against (i) the reference head file executed unmodified under ``oracle/ref_shims.py`` — frozen
as ``tests/golden/*.pt`` by ``oracle/make_golden.py`` — and (ii) the live HF modules on the GPU box
(``oracle/ref_port.py``).  tests/test_oracle.py asserts both.

Unlike the reference it projects K/V once per image and stacks all pairs' query rows; the results are
mathematically identical (validated to ~1e-6).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NUM_HEADS = 12
LN_EPS = 1e-12
QUERY_LEN = 33

# ----------------------------------------------------------------------------------------------
# a3: masks (integer, bit-exact)
# ----------------------------------------------------------------------------------------------


def legacy_nearest_index(out_size: int, in_size: int) -> np.ndarray:
    """PyTorch ``mode='nearest'`` source index: min(floor(dst * float32(in)/float32(out)), in-1).
    (ATen nearest_neighbor_compute_source_index; verified against F.interpolate in tests.)"""
    scale = np.float32(in_size) / np.float32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = np.floor(dst * scale).astype(np.int64)
    return np.minimum(src, in_size - 1)


def object_token_masks(pan: np.ndarray, img_hw: Tuple[int, int], pad_hw: Tuple[int, int],
                       feat_hw: Tuple[int, int], patch: int, object_ids: Sequence[int]) -> np.ndarray:
    """pan id map -> bool [N, L] (v4:416-429): nearest to img_shape, zero-pad to pad_shape, nearest to
    (feat_h//patch, feat_w//patch), compare with each object id.  The reference round-trips through
    float32 (v4:417,422); ids < 2^24 are exact in fp32, which callers must respect."""
    pan = np.asarray(pan)
    ph, pw = pan.shape
    ih, iw = img_hw
    Hp, Wp = pad_hw
    th, tw = feat_hw[0] // patch, feat_hw[1] // patch
    r2 = legacy_nearest_index(th, Hp)       # token row -> padded-image row
    c2 = legacy_nearest_index(tw, Wp)
    r1 = legacy_nearest_index(ih, ph)       # image row -> pan row
    c1 = legacy_nearest_index(iw, pw)
    tok = np.zeros((th, tw), dtype=np.float32)   # F.pad value 0 (v4:420-421)
    rin, cin = r2 < ih, c2 < iw
    rows = r1[r2[rin]]
    cols = c1[c2[cin]]
    tok[np.ix_(rin, cin)] = pan[np.ix_(rows, cols)].astype(np.float32)
    ids = np.asarray([int(i) for i in object_ids], dtype=np.int64).astype(np.float32)
    return tok.reshape(1, -1) == ids[:, None]


def pack_mask_bits(masks: np.ndarray) -> np.ndarray:
    """bool [N, L] -> uint32 [N, ceil(L/32)], bit (l % 32) of word (l // 32) = masks[:, l]."""
    n, L = masks.shape
    W = (L + 31) // 32
    padded = np.zeros((n, W * 32), dtype=np.uint64)
    padded[:, :L] = masks
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    return (padded.reshape(n, W, 32) * weights).sum(-1).astype(np.uint32)


def pair_masks(masks: np.ndarray) -> np.ndarray:
    """bool [N, L] -> bool [N*N, L]: pair p = (p // N, p % N) gets mask_i OR mask_j (v4:430-433)."""
    n = masks.shape[0]
    return (masks[:, None, :] | masks[None, :, :]).reshape(n * n, -1)

# ----------------------------------------------------------------------------------------------
# a3: patch embedding (timm PatchEmbed = Conv2d k=s=patch, flatten(2).transpose(1,2); v4:75,410)
# ----------------------------------------------------------------------------------------------


def patch_embed(feat: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, patch: int) -> torch.Tensor:
    """[1,C,h,w] -> [L, C_out]; row-major token order; partial patches floored away."""
    _, C, h, w = feat.shape
    th, tw = h // patch, w // patch
    x = feat[0, :, :th * patch, :tw * patch].reshape(C, th, patch, tw, patch)
    x = x.permute(1, 3, 0, 2, 4).reshape(th * tw, C * patch * patch)      # [L, C*p*p] in (c, py, px) order
    return x @ weight.reshape(weight.shape[0], -1).t() + bias

# ----------------------------------------------------------------------------------------------
# a4-a7: Q-Former
# ----------------------------------------------------------------------------------------------


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def _lin(x, sd, prefix):
    return x @ sd[prefix + ".weight"].t() + sd[prefix + ".bias"]


def _heads(x):  # [..., S, 768] -> [..., 12, S, 64]
    *lead, S, D = x.shape
    return x.reshape(*lead, S, NUM_HEADS, D // NUM_HEADS).transpose(-2, -3)


def _merge(x):  # [..., 12, S, 64] -> [..., S, 768]
    x = x.transpose(-2, -3)
    return x.reshape(*x.shape[:-2], -1)


def qformer_embeddings(sd: Dict[str, torch.Tensor], p: str, query: torch.Tensor, input_ids: torch.Tensor):
    """LN(cat(query, word_emb[ids] + pos_emb[0:T]))  (HF instructblip :753-782). query [33,768]."""
    B, T = input_ids.shape
    txt = sd[p + "embeddings.word_embeddings.weight"][input_ids] + sd[p + "embeddings.position_embeddings.weight"][:T]
    x = torch.cat([query.unsqueeze(0).expand(B, -1, -1), txt], dim=1)
    return _ln(x, sd[p + "embeddings.layernorm.weight"], sd[p + "embeddings.layernorm.bias"])


def qformer_forward(sd: Dict[str, torch.Tensor], query: torch.Tensor, input_ids: torch.Tensor,
                    text_mask: torch.Tensor, image_tokens: torch.Tensor, obj_masks: torch.Tensor,
                    pair_index: torch.Tensor | None = None, prefix: str = "relation_qformer.",
                    num_layers: int = 2, return_intermediates: bool = False):
    """Two-layer InstructBLIP Q-Former over B pair queries.

    query [33,768]; input_ids/text_mask [B,T]; image_tokens [L,256]; obj_masks bool [N,L];
    pair_index [B] (default arange(N*N)): pair p -> (p // N, p % N).
    Returns last_hidden_state[:, :33]  (what v4:185 keeps).  K/V are projected once per image.
    """
    B, T = input_ids.shape
    N, L = obj_masks.shape
    if pair_index is None:
        pair_index = torch.arange(B)
    pi, pj = pair_index // N, pair_index % N
    M = (obj_masks[pi] | obj_masks[pj])                                         # [B, L]
    cross_bias = (1.0 - M.float()) * torch.finfo(torch.float32).min             # modeling_utils.py:875-876
    self_mask = torch.cat([torch.ones(B, QUERY_LEN), text_mask.float()], dim=1)  # v4:158-159
    self_bias = (1.0 - self_mask) * -10000.0                                    # instructblip :861-862
    inter = {}
    h = qformer_embeddings(sd, prefix, query, input_ids)                        # [B,S,768]
    inter["embeddings"] = h
    scale = 1.0 / math.sqrt(64)
    for l in range(num_layers):
        lp = f"{prefix}encoder.layer.{l}."
        # self attention over all S rows
        q = _heads(_lin(h, sd, lp + "attention.attention.query"))
        k = _heads(_lin(h, sd, lp + "attention.attention.key"))
        v = _heads(_lin(h, sd, lp + "attention.attention.value"))
        s = q @ k.transpose(-1, -2) * scale + self_bias[:, None, None, :]
        a = _merge(torch.softmax(s, dim=-1) @ v)
        h = _ln(_lin(a, sd, lp + "attention.output.dense") + h, sd[lp + "attention.output.LayerNorm.weight"],
                sd[lp + "attention.output.LayerNorm.bias"])
        inter[f"l{l}.self"] = h
        hq, ht = h[:, :QUERY_LEN], h[:, QUERY_LEN:]
        # cross attention: query rows only; K/V shared by all pairs of the image
        qc = _heads(_lin(hq, sd, lp + "crossattention.attention.query"))       # [B,12,33,64]
        kc = _heads(_lin(image_tokens, sd, lp + "crossattention.attention.key"))    # [12,L,64]
        vc = _heads(_lin(image_tokens, sd, lp + "crossattention.attention.value"))
        sc = qc @ kc.transpose(-1, -2) * scale + cross_bias[:, None, None, :]
        c = _merge(torch.softmax(sc, dim=-1) @ vc)
        inter[f"l{l}.xattn_ctx"] = c
        hq = _ln(_lin(c, sd, lp + "crossattention.output.dense") + hq,
                 sd[lp + "crossattention.output.LayerNorm.weight"], sd[lp + "crossattention.output.LayerNorm.bias"])
        inter[f"l{l}.cross"] = hq
        # FFNs: separate weights for query rows and text rows (:663-677,687-695); erf GELU
        fq = _lin(F.gelu(_lin(hq, sd, lp + "intermediate_query.dense")), sd, lp + "output_query.dense")
        hq = _ln(fq + hq, sd[lp + "output_query.LayerNorm.weight"], sd[lp + "output_query.LayerNorm.bias"])
        if T > 0:
            ft = _lin(F.gelu(_lin(ht, sd, lp + "intermediate.dense")), sd, lp + "output.dense")
            ht = _ln(ft + ht, sd[lp + "output.LayerNorm.weight"], sd[lp + "output.LayerNorm.bias"])
        h = torch.cat([hq, ht], dim=1)
        inter[f"l{l}.out"] = h
    out = h[:, :QUERY_LEN]
    return (out, inter) if return_intermediates else out

# ----------------------------------------------------------------------------------------------
# a8: existence filter
# ----------------------------------------------------------------------------------------------


def existence_logits(cls_feature: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """z_p = w . O_p[0] + b  (v4:206-208); probability = sigmoid(z) (v4:209)."""
    return (cls_feature @ weight.reshape(-1, 1)).reshape(-1) + bias.reshape(())


def topk_pairs(z: torch.Tensor, k: int) -> List[int]:
    """``pred.topk(B).indices[:k]`` (v4:236-237) with ties broken toward the lower pair index
    (torch.topk leaves tie order unspecified).  Ranking on logits is the same
    order as ranking on sigmoid(z) wherever fp32 sigmoid has not saturated."""
    zz = np.asarray(z, dtype=np.float32).reshape(-1)
    order = np.lexsort((np.arange(zz.size), -zz.astype(np.float64)))
    return [int(i) for i in order[:k]]


def existence_mask(z: torch.Tensor, threshold: float = 0.5) -> np.ndarray:
    """sigmoid(z) > thr, decided in logit space: z > logit(thr) (thr=0.5 -> z > 0)."""
    t = math.log(threshold / (1.0 - threshold))
    return np.asarray(z, dtype=np.float32).reshape(-1) > np.float32(t)

# ----------------------------------------------------------------------------------------------
# a11: mask mean-pool + pair concat (detectors/openseed_relation.py:454-468, 502-527)
# ----------------------------------------------------------------------------------------------


def mask_pool_pairs(feature: torch.Tensor, masks: torch.Tensor):
    """feature [C,h,w], masks float/bool [N,h,w] ->
    object emb [N,C] = sum(feat*mask)/(sum(mask)+1e-8); pair emb [N*N, 2C] = cat(obj[i], obj[j])."""
    m = masks.float()
    num = torch.einsum("chw,nhw->nc", feature.double(), m.double())
    den = m.double().sum(dim=(1, 2))[:, None] + 1e-8
    obj = (num / den).float()
    n = obj.shape[0]
    pair = torch.cat([obj[:, None, :].expand(n, n, -1), obj[None, :, :].expand(n, n, -1)], dim=-1)
    return obj, pair.reshape(n * n, -1)

# ----------------------------------------------------------------------------------------------
# a9-a10: OPT greedy decode over an embedded prefix (HF models/opt/modeling_opt.py)
# ----------------------------------------------------------------------------------------------


def opt_positions(mask: torch.Tensor) -> torch.Tensor:
    """OPTLearnedPositionalEmbedding (:56-70): cumsum(mask)*mask - 1 + 2; pads -> row 1."""
    m = mask.long()
    return torch.cumsum(m, dim=1) * m - 1 + 2


def opt_forward(sd: Dict[str, torch.Tensor], cfg: dict, embeds: torch.Tensor, mask: torch.Tensor,
                prefix: str = "language_model.") -> torch.Tensor:
    """Full (non-cached) pre-LN OPT decoder forward.  embeds [B,S,d], mask [B,S] (1 = valid; pads may sit
    mid-sequence, v4:298-299).  Returns logits [B,S,V] (tied lm_head)."""
    B, S, d = embeds.shape
    H = cfg["num_attention_heads"]
    hd = d // H
    dp = prefix + "model.decoder."
    pos = sd[dp + "embed_positions.weight"][opt_positions(mask)]
    h = embeds + pos
    causal = torch.tril(torch.ones(S, S, dtype=torch.bool, device=embeds.device))
    allow = causal[None, :, :] & mask.bool()[:, None, :]
    bias = torch.zeros(B, 1, S, S, device=embeds.device).masked_fill(~allow[:, None], torch.finfo(torch.float32).min)
    for l in range(cfg["num_hidden_layers"]):
        lp = f"{dp}layers.{l}."
        x = F.layer_norm(h, (d,), sd[lp + "self_attn_layer_norm.weight"], sd[lp + "self_attn_layer_norm.bias"], 1e-5)
        q = _lin(x, sd, lp + "self_attn.q_proj") * (hd ** -0.5)
        k = _lin(x, sd, lp + "self_attn.k_proj")
        v = _lin(x, sd, lp + "self_attn.v_proj")
        q, k, v = (t.reshape(B, S, H, hd).transpose(1, 2) for t in (q, k, v))
        a = torch.softmax(q @ k.transpose(-1, -2) + bias, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, S, d)
        h = h + _lin(a, sd, lp + "self_attn.out_proj")
        x = F.layer_norm(h, (d,), sd[lp + "final_layer_norm.weight"], sd[lp + "final_layer_norm.bias"], 1e-5)
        h = h + _lin(F.relu(_lin(x, sd, lp + "fc1")), sd, lp + "fc2")
    h = F.layer_norm(h, (d,), sd[dp + "final_layer_norm.weight"], sd[dp + "final_layer_norm.bias"], 1e-5)
    return h @ sd[dp + "embed_tokens.weight"].t()


def opt_greedy_decode(sd, cfg, prefix_embeds: torch.Tensor, prefix_mask: torch.Tensor, max_new_tokens: int,
                      prefix: str = "language_model.", forced_tokens: torch.Tensor | None = None):
    """Greedy decode (v4:305-312; EOS never stops it in the synthetic setting: min_new = max_new).
    Recomputes the whole sequence each step (oracle clarity over speed).  Returns
    (tokens [B,T_new], scores [B,T_new,V]).  ``forced_tokens`` teacher-forces the fed-back ids."""
    emb = sd[prefix + "model.decoder.embed_tokens.weight"]
    embeds, mask = prefix_embeds, prefix_mask.long()
    toks, scores = [], []
    for t in range(max_new_tokens):
        logits = opt_forward(sd, cfg, embeds, mask, prefix)[:, -1]
        nxt = logits.argmax(dim=-1)
        scores.append(logits)
        toks.append(nxt)
        feed = nxt if forced_tokens is None else forced_tokens[:, t]
        embeds = torch.cat([embeds, emb[feed][:, None, :]], dim=1)
        mask = torch.cat([mask, torch.ones(mask.shape[0], 1, dtype=mask.dtype, device=mask.device)], dim=1)
    return torch.stack(toks, dim=1), torch.stack(scores, dim=1)


def build_llm_prefix(sd, pair_feature: torch.Tensor, llm_ids: torch.Tensor, llm_mask: torch.Tensor,
                     prefix: str = "language_model.", embed_key: str | None = None):
    """a9 (v4:294-301): U = pair_feature @ W_L^T + b_L  [k,32,d_llm]; cat with embed_tokens(ids);
    mask = cat(ones[k,32], left-padded text mask)."""
    U = _lin(pair_feature, sd, "language_projection")
    E = sd[embed_key or (prefix + "model.decoder.embed_tokens.weight")][llm_ids]
    embeds = torch.cat([U, E], dim=1)
    mask = torch.cat([torch.ones(U.shape[0], U.shape[1], dtype=torch.long, device=U.device), llm_mask.long()], dim=1)
    return embeds, mask

# ----------------------------------------------------------------------------------------------
# a9-a10 for the LLM the shipped config names: Llama greedy decode (HF models/llama/modeling_llama.py)
# ----------------------------------------------------------------------------------------------


def _lin_opt_bias(x, sd, prefix):
    y = x @ sd[prefix + ".weight"].t()
    b = sd.get(prefix + ".bias")
    return y if b is None else y + b


def _rmsnorm(x, w, eps):
    """LlamaRMSNorm (:52-69): weight * (x * rsqrt(mean(x^2) + eps))."""
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def llama_positions(mask: torch.Tensor) -> torch.Tensor:
    """position_ids as HF ``generate`` derives them from the attention mask (generation/utils.py:707-729):
    cumsum(mask) - 1, pad slots set to a dummy in-range value (their rows are never attended to)."""
    m = mask.long()
    return (torch.cumsum(m, dim=1) - 1).clamp_min(0)


def llama_forward(sd: Dict[str, torch.Tensor], cfg: dict, embeds: torch.Tensor, mask: torch.Tensor,
                  prefix: str = "language_model.") -> torch.Tensor:
    """Full (non-cached) Llama decoder forward: RMSNorm -> q/k/v (no bias) -> rotate_half RoPE on q, k (:137-166) ->
    causal + key-padding softmax(qk^T / sqrt(hd)) v with repeat_kv (:187-196) -> o_proj + residual -> RMSNorm ->
    down(silu(gate) * up) + residual (:170-184) -> final RMSNorm -> untied lm_head.  embeds [B,S,d], mask [B,S]."""
    B, S, d = embeds.shape
    H = cfg["num_attention_heads"]
    KV = cfg.get("num_key_value_heads") or H
    hd = cfg.get("head_dim") or d // H
    eps = cfg.get("rms_norm_eps", 1e-6)
    theta = cfg.get("rope_theta", 10000.0)
    mp = prefix + "model."
    pos = llama_positions(mask).float()                                           # [B,S]
    inv_freq = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.int64, device=embeds.device).float() / hd))
    freqs = pos[:, :, None] * inv_freq[None, None, :]                             # [B,S,hd/2]
    emb = torch.cat([freqs, freqs], dim=-1)
    cos, sin = emb.cos()[:, None], emb.sin()[:, None]                             # [B,1,S,hd]

    def rot(x):
        x1, x2 = x[..., : hd // 2], x[..., hd // 2:]
        return torch.cat([-x2, x1], dim=-1)

    h = embeds
    causal = torch.tril(torch.ones(S, S, dtype=torch.bool, device=embeds.device))
    allow = causal[None, :, :] & mask.bool()[:, None, :]
    bias = torch.zeros(B, 1, S, S, device=embeds.device).masked_fill(~allow[:, None], torch.finfo(torch.float32).min)
    n_layers = cfg["num_hidden_layers"]
    for l in range(n_layers):
        lp = f"{mp}layers.{l}."
        x = _rmsnorm(h, sd[lp + "input_layernorm.weight"], eps)
        q = _lin_opt_bias(x, sd, lp + "self_attn.q_proj").reshape(B, S, H, hd).transpose(1, 2)
        k = _lin_opt_bias(x, sd, lp + "self_attn.k_proj").reshape(B, S, KV, hd).transpose(1, 2)
        v = _lin_opt_bias(x, sd, lp + "self_attn.v_proj").reshape(B, S, KV, hd).transpose(1, 2)
        q = q * cos + rot(q) * sin
        k = k * cos + rot(k) * sin
        if KV != H:
            k = k.repeat_interleave(H // KV, dim=1)
            v = v.repeat_interleave(H // KV, dim=1)
        a = torch.softmax(q @ k.transpose(-1, -2) * (hd ** -0.5) + bias, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, S, H * hd)
        h = h + _lin_opt_bias(a, sd, lp + "self_attn.o_proj")
        x = _rmsnorm(h, sd[lp + "post_attention_layernorm.weight"], eps)
        g = _lin_opt_bias(x, sd, lp + "mlp.gate_proj")
        u = _lin_opt_bias(x, sd, lp + "mlp.up_proj")
        h = h + _lin_opt_bias(F.silu(g) * u, sd, lp + "mlp.down_proj")
    h = _rmsnorm(h, sd[mp + "norm.weight"], eps)
    lm = sd.get(prefix + "lm_head.weight")
    if lm is None:
        lm = sd[mp + "embed_tokens.weight"]
    return h @ lm.t()


def llama_greedy_decode(sd, cfg, prefix_embeds: torch.Tensor, prefix_mask: torch.Tensor, max_new_tokens: int,
                        prefix: str = "language_model.", forced_tokens: torch.Tensor | None = None):
    """Greedy decode as ``opt_greedy_decode`` (v4:305-312), Llama arithmetic."""
    emb = sd[prefix + "model.embed_tokens.weight"]
    embeds, mask = prefix_embeds, prefix_mask.long()
    toks, scores = [], []
    for t in range(max_new_tokens):
        logits = llama_forward(sd, cfg, embeds, mask, prefix)[:, -1]
        nxt = logits.argmax(dim=-1)
        scores.append(logits)
        toks.append(nxt)
        feed = nxt if forced_tokens is None else forced_tokens[:, t]
        embeds = torch.cat([embeds, emb[feed][:, None, :]], dim=1)
        mask = torch.cat([mask, torch.ones(mask.shape[0], 1, dtype=mask.dtype, device=mask.device)], dim=1)
    return torch.stack(toks, dim=1), torch.stack(scores, dim=1)


def greedy_decode(sd, cfg, prefix_embeds, prefix_mask, max_new_tokens, prefix: str = "language_model.", forced_tokens=None):
    """Dispatch on the config's ``model_type`` (OPT when absent)."""
    fn = llama_greedy_decode if cfg.get("model_type", "opt") == "llama" else opt_greedy_decode
    return fn(sd, cfg, prefix_embeds, prefix_mask, max_new_tokens, prefix, forced_tokens)


def embed_tokens_key(cfg, prefix: str = "language_model.") -> str:
    return prefix + ("model.embed_tokens.weight" if cfg.get("model_type", "opt") == "llama" else "model.decoder.embed_tokens.weight")

# ----------------------------------------------------------------------------------------------
# a11, whole chain (detectors/openseed_relation.py:441-493): masks from the panoptic map, pooling, class embedding,
# background feature
# ----------------------------------------------------------------------------------------------


def object_masks_feature_res(pan: np.ndarray, img_hw, pad_hw, feat_hw, object_ids: Sequence[int]) -> np.ndarray:
    """mask_o = nearest(pad0(nearest(pan == id_o -> img_shape) -> pad_shape) -> feature size)  (:441-462), bool [N, h, w].
    Resizing the 0/1 mask with 'nearest' = reading the panoptic id at the composed source pixel; padding reads as no object."""
    pan = np.asarray(pan)
    fh, fw = feat_hw
    r2 = legacy_nearest_index(fh, pad_hw[0])
    c2 = legacy_nearest_index(fw, pad_hw[1])
    inside = (r2 < img_hw[0])[:, None] & (c2 < img_hw[1])[None, :]
    r1 = legacy_nearest_index(img_hw[0], pan.shape[0])[np.minimum(r2, img_hw[0] - 1)]
    c1 = legacy_nearest_index(img_hw[1], pan.shape[1])[np.minimum(c2, img_hw[1] - 1)]
    ids = pan[r1][:, c1]
    return np.stack([(ids == int(o)) & inside for o in object_ids], axis=0)


def mask_pool_chain(feature: torch.Tensor, pan, img_hw, pad_hw, object_ids: Sequence[int], cls_table: torch.Tensor | None = None,
                    cls_mode: str = "none", use_background: bool = False):
    """feature [C,h,w] -> (object embedding [N, C'], pair embedding [N*N, 2C']); float64 accumulation."""
    masks = torch.from_numpy(object_masks_feature_res(np.asarray(pan), img_hw, pad_hw, feature.shape[-2:], object_ids))
    m = masks.double()
    f = feature.double()
    cnt = m.sum(dim=(1, 2))[:, None]
    obj = torch.einsum("chw,nhw->nc", f, m) / (cnt + 1e-8)
    if cls_mode != "none":
        e = cls_table[torch.tensor([int(o) % 1000 for o in object_ids])].double()
        obj = obj + e if cls_mode == "add" else torch.cat([obj, e], dim=-1)
    if use_background:
        bg = torch.einsum("chw,nhw->nc", f, 1.0 - m) / ((1.0 - m).sum(dim=(1, 2))[:, None] + 1e-8)
        obj = obj + bg
    obj = obj.float()
    n = obj.shape[0]
    pair = torch.cat([obj[:, None, :].expand(n, n, -1), obj[None, :, :].expand(n, n, -1)], dim=-1)
    return obj, pair.reshape(n * n, -1)

# ----------------------------------------------------------------------------------------------
# f1 / f2: detector glue and result wire format (numpy restatements; integer work -> bit-exact)
# ----------------------------------------------------------------------------------------------


def relabel_panoptic(pan_seg: np.ndarray, segments_info) -> Tuple[np.ndarray, List[int]]:
    """detectors/openseed_relation_v2.py:112-124: per segment, in list order, pan[pan_seg == id] = category + 1000 * (number of
    earlier segments of that category); everything else 0.  -> (pan_results, object ids)."""
    pan_seg = np.asarray(pan_seg)
    out = np.zeros_like(pan_seg)
    record: Dict[int, int] = {}
    ids = []
    for seg in segments_info:
        cat = int(seg["category_id"])
        record[cat] = record.get(cat, -1) + 1
        new = cat + 1000 * record[cat]
        out[np.where(pan_seg == int(seg["id"]))] = new
        ids.append(new)
    return out, ids


def submission_image(pan_results: np.ndarray, object_id_list, rng):
    """tools/infer.py:149-168: (uint8 [H, W, 3] image in R, G, B order as the PNG stores it, segments_info)."""
    pan_results = np.asarray(pan_results)
    acc = np.zeros(pan_results.shape + (3,), dtype=np.int64)
    segments_info = []
    for object_id in object_id_list:
        object_id = int(object_id)
        mask = pan_results == object_id
        if object_id == 133:
            continue
        r, g, b = rng.choices(range(0, 255), k=3)
        acc = acc + mask[..., None].astype(np.int64) * np.array([r, g, b]).reshape(1, 1, 3)
        segments_info.append(dict(category_id=int(object_id % 1000 + 1), id=r + 256 * g + 256 * 256 * b))
    return acc.astype(np.uint8), segments_info
