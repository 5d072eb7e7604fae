"""Seeded synthetic inputs, tokenizers and weights shared by the oracle, the tests and bench.py.

Nothing here is on the product compute path: it only *describes* workloads (BASELINE.json configs,
SURVEY.md §8d) so that the reference (run under ``oracle/ref_shims.py``), the CPU oracle and the CUDA
path all see byte-identical inputs.

Input dict layout = what ``OpenSeeDRelationV2.simple_test`` hands to the head
(reference ``kings_sgg/models/detectors/openseed_relation_v2.py:177-181`` and ``:128-141``).
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from .categories import INSTANCE_OFFSET, object_categories, relation_categories

# ----------------------------------------------------------------------------------------------
# Workload descriptions (BASELINE.json "configs")
# ----------------------------------------------------------------------------------------------


@dataclass(frozen=True)
class Workload:
    name: str
    height: int          # padded image height (pixels)
    width: int
    num_objects: int
    num_images: int = 1
    llm: bool = False
    topk_pairs: int = 20       # reference hard-codes 20 (relation_transformer_head_v4.py:237)
    max_new_tokens: int = 16   # reference hard-codes 16 (relation_transformer_head_v4.py:308)

    @property
    def queries(self) -> int:      # N^2 pair queries the head evaluates (diagonal included)
        return self.num_objects ** 2

    @property
    def ordered_pairs(self) -> int:  # N(N-1): BASELINE.json's reporting unit
        return self.num_objects * (self.num_objects - 1)

    @property
    def image_tokens(self) -> int:
        return (self.height // 4 // 16) * (self.width // 4 // 16)


WORKLOADS: Dict[str, Workload] = {
    "cfg1": Workload("cfg1", 256, 256, 8),
    "cfg2": Workload("cfg2", 1024, 1024, 40),
    "cfg3": Workload("cfg3", 1024, 1024, 40, llm=True, topk_pairs=100, max_new_tokens=32),
    "cfg4": Workload("cfg4", 1024, 1024, 40, num_images=32),
    "cfg5": Workload("cfg5", 1024, 1024, 80, num_images=8, llm=True, topk_pairs=100, max_new_tokens=32),
}

# ----------------------------------------------------------------------------------------------
# Images
# ----------------------------------------------------------------------------------------------


def _legacy_nearest_src(dst: np.ndarray, in_size: int, out_size: int) -> np.ndarray:
    scale = np.float32(in_size) / np.float32(out_size)
    src = np.floor(dst.astype(np.float32) * scale).astype(np.int64)
    return np.minimum(src, in_size - 1)


def object_ids(num_objects: int) -> List[int]:
    """id = category + 1000 * instance, category = i % 133."""
    n_cat = len(object_categories)
    return [(i % n_cat) + INSTANCE_OFFSET * (i // n_cat) for i in range(num_objects)]


def make_panoptic_map(height: int, width: int, num_objects: int, gen: torch.Generator,
                      ensure_token_coverage: bool = True) -> torch.Tensor:
    """N-seed Voronoi partition of the H x W grid; every pixel belongs to exactly one object."""
    ids = torch.tensor(object_ids(num_objects), dtype=torch.int32)
    th, tw = height // 64, width // 64
    rows = torch.from_numpy(_legacy_nearest_src(np.arange(th), height, th))
    cols = torch.from_numpy(_legacy_nearest_src(np.arange(tw), width, tw))
    yy = torch.arange(height, dtype=torch.float32)[:, None]
    xx = torch.arange(width, dtype=torch.float32)[None, :]
    for _ in range(1000):
        sy = torch.rand(num_objects, generator=gen) * height
        sx = torch.rand(num_objects, generator=gen) * width
        best = torch.full((height, width), float("inf"))
        label = torch.zeros((height, width), dtype=torch.int64)
        for o in range(num_objects):
            d = (yy - sy[o]) ** 2 + (xx - sx[o]) ** 2
            closer = d < best
            best = torch.where(closer, d, best)
            label = torch.where(closer, torch.full_like(label, o), label)
        if not ensure_token_coverage:
            break
        seen = torch.unique(label[rows][:, cols])
        if seen.numel() == num_objects:
            break
    else:  # pragma: no cover
        raise RuntimeError("could not draw a Voronoi map covering every object at token scale")
    return ids[label]


def make_image_inputs(workload: Workload, image_index: int = 0, *, ensure_token_coverage: bool = True,
                      pan_scale: float = 1.0, img_shape: Optional[tuple] = None) -> dict:
    """One image worth of head inputs (test-mode dict).

    ``pan_scale`` != 1 makes ``pan_results`` a different resolution than ``img_shape`` (exercises the
    first nearest resize); ``img_shape`` smaller than the padded shape exercises the zero padding.
    """
    gen = torch.Generator().manual_seed(1234 + image_index)
    H, W, N = workload.height, workload.width, workload.num_objects
    feats = torch.randn(1, 256, H // 4, W // 4, generator=gen)
    ih, iw = (img_shape or (H, W))[:2]
    ph, pw = max(1, int(round(ih * pan_scale))), max(1, int(round(iw * pan_scale)))
    if (ph, pw) == (H, W):
        pan = make_panoptic_map(H, W, N, gen, ensure_token_coverage)
    else:
        pan = make_panoptic_map(ph, pw, N, gen, ensure_token_coverage=False)
    ids = object_ids(N)
    return {
        "mask_features": feats,
        "img_metas": [{"img_shape": (ih, iw, 3), "pad_shape": (H, W, 3)}],
        "object_info": [{
            "object_id_list": [torch.tensor(i, dtype=torch.int32) for i in ids],
            "pan_results": pan,
        }],
    }


class BitmapMasksStandIn:
    """The one method of mmdet's ``BitmapMasks`` the head uses at train time (``to_tensor(dtype, device)``, v4:371)."""

    def __init__(self, masks: np.ndarray):
        self.masks = np.asarray(masks, dtype=np.uint8)          # [n_thing, H, W]

    def to_tensor(self, dtype, device):
        return torch.tensor(self.masks, dtype=dtype, device=device)


def make_train_inputs(workload: Workload, image_index: int = 0, num_rels: int = 6) -> dict:
    """One image worth of TRAIN-mode head inputs (``detectors/openseed_relation_v2.py:159-165``):
    ``mask_features``, ``img_metas`` with ``masks_info`` / ``gt_rels`` (``datasets/pipelines/loading.py:23-28``),
    ``gt_masks`` (thing bitmaps), ``gt_labels``, ``gt_semantic_seg``.  Objects alternate thing / stuff."""
    base = make_image_inputs(workload, image_index)
    gen = torch.Generator().manual_seed(4321 + image_index)
    N = workload.num_objects
    pan = base["object_info"][0]["pan_results"]
    ids = object_ids(N)
    infos = [dict(category=i % INSTANCE_OFFSET, is_thing=(k % 2 == 0)) for k, i in enumerate(ids)]
    thing = np.stack([(pan == i).numpy() for k, i in enumerate(ids) if infos[k]["is_thing"]]).astype(np.uint8)
    sem = (pan % INSTANCE_OFFSET).to(torch.int64)[None]            # [1, H, W] category map
    rels, seen = [], set()
    while len(rels) < num_rels:
        i, j = (int(x) for x in torch.randint(0, N, (2,), generator=gen))
        r = int(torch.randint(0, len(relation_categories), (1,), generator=gen))
        if i != j and (i, j, r) not in seen:
            seen.add((i, j, r))
            rels.append([i, j, r])
    meta = dict(base["img_metas"][0], masks_info=infos, gt_rels=[rels])
    return {
        "mask_features": base["mask_features"],
        "img_metas": [meta],
        "gt_labels": [torch.tensor([x["category"] for x in infos if x["is_thing"]])],
        "gt_masks": [BitmapMasksStandIn(thing)],
        "gt_semantic_seg": [sem],
    }


def inputs_to(inputs: dict, device) -> dict:
    out = dict(inputs)
    out["mask_features"] = inputs["mask_features"].to(device)
    oi = dict(inputs["object_info"][0])
    oi["pan_results"] = oi["pan_results"].to(device)
    oi["object_id_list"] = [t.to(device) for t in oi["object_id_list"]]
    out["object_info"] = [oi]
    return out

# ----------------------------------------------------------------------------------------------
# Tokenizers (no vocab files exist offline; ids are a pure function of the string)
# ----------------------------------------------------------------------------------------------


class _Encoding(dict):
    def __getattr__(self, k):
        return self[k]


class SyntheticTokenizer:
    """Deterministic stand-in for ``AutoTokenizer`` (reference call sites v4:85-86,104-105,149,263,313).

    ids are drawn from a generator seeded with crc32(text), so any two implementations given the same
    strings see the same ids.  ``padding=True`` pads to the fixed ``max_len`` (right or left according
    to ``padding_side``).  ``batch_decode`` maps a token id t to ``relation_categories[t % 56]`` and
    joins with two spaces between ``<s>`` and ``</s>`` — the format the reference parser expects
    (v4:313-318).
    """

    def __init__(self, kind: str = "qformer"):
        assert kind in ("qformer", "llm")
        self.kind = kind
        if kind == "qformer":   # BERT-like: ids in [1000, 30522), 12..16 valid, pad id 0
            self.lo, self.hi, self.min_len, self.max_len, self.pad_token_id = 1000, 30522, 12, 16, 0
            self.padding_side = "right"
        else:                   # OPT-like: ids in [4, vocab), 14..17 valid, pad = unk = 3
            self.lo, self.hi, self.min_len, self.max_len, self.pad_token_id = 4, 50272, 14, 17, 3
            self.padding_side = "left"
        self.unk_token = "<unk>"
        self.pad_token = "<pad>"
        self._cache: Dict[str, np.ndarray] = {}

    def set_vocab_size(self, vocab: int):
        self.hi = int(vocab)
        self._cache.clear()

    def _encode(self, text: str) -> np.ndarray:
        ids = self._cache.get(text)
        if ids is None:
            rs = np.random.RandomState(zlib.crc32(text.encode("utf-8")) & 0x7FFFFFFF)
            n = int(rs.randint(self.min_len, self.max_len + 1))
            ids = rs.randint(self.lo, self.hi, size=n).astype(np.int64)
            self._cache[text] = ids
        return ids

    def __call__(self, texts, return_tensors="pt", padding=True, return_attention_mask=True, **_):
        if isinstance(texts, str):
            texts = [texts]
        T = self.max_len
        ids = np.full((len(texts), T), self.pad_token_id, dtype=np.int64)
        mask = np.zeros((len(texts), T), dtype=np.int64)
        for r, t in enumerate(texts):
            e = self._encode(t)
            if self.padding_side == "left":
                ids[r, T - len(e):] = e
                mask[r, T - len(e):] = 1
            else:
                ids[r, :len(e)] = e
                mask[r, :len(e)] = 1
        return _Encoding(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask))

    def batch_decode(self, sequences, **_):
        out = []
        for seq in sequences:
            toks = [int(t) for t in (seq.tolist() if hasattr(seq, "tolist") else seq)]
            out.append("<s> " + "  ".join(relation_categories[t % len(relation_categories)] for t in toks) + "</s>")
        return out

# ----------------------------------------------------------------------------------------------
# Weights
# ----------------------------------------------------------------------------------------------


def init_parameters(module: torch.nn.Module, seed: int = 0, *, skip_prefixes=()) -> None:
    """Fill every parameter from a CPU generator seeded with (seed, crc32(parameter name)).

    Being a pure function of (name, shape, seed) it gives the reference head (built under the oracle
    shims), the oracle port and the drop-in head identical weights — whichever subset of modules each
    one owns — without shipping 270 MB fixtures.
    Scales: Linear/Conv weights ~ N(0, 0.04^2) (patch_embed ~ 1/sqrt(fan_in)), biases ~ N(0, 0.02^2),
    LayerNorm gamma = 1 + 0.1 N(0,1), embeddings ~ N(0, 0.02^2), learned queries ~ N(0, 1)
    (reference init v4:87-90), language-model weights ~ N(0, 0.02^2).
    """
    gen = torch.Generator()
    with torch.no_grad():
        for name, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):
            if any(name.startswith(pre) for pre in skip_prefixes):
                continue
            gen.manual_seed((seed * 1000003 + zlib.crc32(name.encode("utf-8"))) & 0x7FFFFFFFFFFF)
            lname = name.lower()
            r = torch.randn(p.shape, generator=gen, dtype=torch.float32)
            is_lm = name.startswith("language_model")
            if name in ("relation_query", "rel_cls_query"):
                v = r
            elif "layernorm" in lname or "layer_norm" in lname or lname.endswith("norm.weight") and p.dim() == 1:
                v = (1.0 + 0.1 * r) if name.endswith("weight") else 0.02 * r
            elif p.dim() == 1:
                v = 0.02 * r
            elif name.startswith("patch_embed"):
                fan_in = p[0].numel()
                v = r / float(np.sqrt(fan_in))
            elif "embeddings" in lname or "embed_" in lname:
                v = 0.02 * r
            elif is_lm:
                v = 0.02 * r
            else:
                v = 0.04 * r
            p.copy_(v.to(p.dtype))


OPT_2P7B = dict(vocab_size=50272, hidden_size=2560, num_hidden_layers=32, ffn_dim=10240,
                num_attention_heads=32, max_position_embeddings=2048, word_embed_proj_dim=2560,
                do_layer_norm_before=True, activation_function="relu")
OPT_TINY = dict(vocab_size=1024, hidden_size=320, num_hidden_layers=2, ffn_dim=1280,
                num_attention_heads=4, max_position_embeddings=256, word_embed_proj_dim=320,
                do_layer_norm_before=True, activation_function="relu")
# Llama family (the LLM configs/psg/baseline_v4_ov.py:60-61 names).  "model_type" selects the HF class in the builders.
LLAMA2_7B = dict(model_type="llama", vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                 num_attention_heads=32, num_key_value_heads=32, max_position_embeddings=4096, rms_norm_eps=1e-5)
LLAMA_TINY = dict(model_type="llama", vocab_size=1024, hidden_size=256, intermediate_size=704, num_hidden_layers=2,
                  num_attention_heads=2, num_key_value_heads=2, max_position_embeddings=512, rms_norm_eps=1e-5)
LLAMA_TINY_GQA = dict(LLAMA_TINY, hidden_size=512, num_attention_heads=4, num_key_value_heads=2, intermediate_size=1408)


def build_causal_lm(llm_config: dict):
    """Random-init HF causal LM from one of the config dicts above (OPT unless ``model_type == 'llama'``)."""
    cfg = dict(llm_config)
    if cfg.pop("model_type", "opt") == "llama":
        from transformers import LlamaConfig, LlamaForCausalLM
        return LlamaForCausalLM(LlamaConfig(**cfg))
    from transformers import OPTConfig, OPTForCausalLM
    return OPTForCausalLM(OPTConfig(**cfg))


WEIGHT_SEED = 0


def build_synthetic_head(llm: Optional[dict] = None, max_object_num: int = 80, topk_pairs: int = 20, max_new_tokens: int = 16,
                         device=None, llm_on_device: bool = False, **head_kwargs):
    """The drop-in head with synthetic tokenizers and seeded random weights (no vocab files / checkpoints offline).
    llm: None -> no language model (relation queries + existence filter only); a config dict (OPT_*, LLAMA_*) -> random-init
    LLM of that shape.  ``llm_on_device`` builds the LLM directly on ``device`` with the framework's own default init
    (multi-GB models: seconds instead of minutes; weights then differ from the CPU-seeded ones, fine for benchmarks)."""
    from .head import RelationTransformerHeadV4
    lm, ltok, d_llm = False, None, 4096
    if llm is not None:
        if llm_on_device and device is not None:
            torch.manual_seed(WEIGHT_SEED)             # same on-device weights in every process (bench.py and the parity test)
            with torch.device(device):
                lm = build_causal_lm(llm).eval()
        else:
            lm = build_causal_lm(llm)
        ltok = SyntheticTokenizer("llm")
        ltok.set_vocab_size(llm["vocab_size"])
        d_llm = llm["hidden_size"]
    head = RelationTransformerHeadV4(llm_feature_size=d_llm, max_object_num=max_object_num, topk_pairs=topk_pairs,
                                     max_new_tokens=max_new_tokens, qformer_tokenizer=SyntheticTokenizer("qformer"),
                                     llm_tokenizer=ltok, language_model=lm, **head_kwargs)
    init_parameters(head, WEIGHT_SEED, skip_prefixes=("language_model",) if (llm_on_device and llm is not None) else ())
    head.eval()
    if device is not None:
        head.to(device)
    return head


def make_stress_inputs() -> dict:
    """Edge-case image (tests only): non-square 256x320 padded shape (L = 4x5 = 20 tokens), image smaller
    than its padded shape (zero padding aliases panoptic id 0 = 'person' instance 0, v4:420-421), a
    panoptic map at half the image resolution (first nearest resize is not the identity), 6 Voronoi
    objects that need not all survive at token scale, plus a 7th listed object that owns no pixel at
    all -> an empty object mask, so pair (6,6) has an all-masked cross-attention row (uniform softmax)."""
    wl = Workload("stress", 256, 320, 6)
    inp = make_image_inputs(wl, image_index=7, ensure_token_coverage=False, pan_scale=0.5, img_shape=(200, 300, 3))
    inp["object_info"][0]["object_id_list"].append(torch.tensor(77 + INSTANCE_OFFSET, dtype=torch.int32))
    return inp


def make_mask_pool_case():
    """Inputs of the a11 (mask-pooled object / pair embedding) golden: the stress image (image smaller than its padded shape,
    panoptic map at half resolution, an object that owns no pixel) plus an object that repeats an id and one whose id is
    not in the map.  -> (mask_features [1,256,h,w], pan [Hp,Wp], object ids, meta, class-embedding table [133,256])."""
    inp = make_stress_inputs()
    info, meta = inp["object_info"][0], dict(inp["img_metas"][0])
    meta["ori_shape"] = tuple(info["pan_results"].shape) + (3,)
    ids = [int(i) for i in info["object_id_list"]]
    ids = ids + [ids[2], 2000 + 5]
    g = torch.Generator().manual_seed(99)
    table = torch.randn(133, 256, generator=g) * 0.5
    return inp["mask_features"], info["pan_results"], ids, meta, table
