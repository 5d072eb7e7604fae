"""TEST INFRASTRUCTURE — never imported by the product path.

Loads the UNMODIFIED reference head ``/root/reference/kings_sgg/models/relation_heads/
relation_transformer_head_v4.py`` on CPU under ``sys.modules`` shims for the packages that are not in
this image (mmcv, mmdet, timm), so that golden vectors can be generated from the reference itself
(SURVEY.md §8c).  Only works where /root/reference exists (the authoring container), so nothing in
``-m gpu`` tests, ``smoke()`` or ``bench.py`` uses it; they use the committed fixtures in tests/golden.

Shimmed, and why:
  * ``mmcv.runner.BaseModule``  -> ``torch.nn.Module``                        (v4:11)
  * ``mmdet.core.INSTANCE_OFFSET`` = 1000                                     (v4:12)
  * ``mmdet.models.builder.HEADS/DETECTORS`` -> identity ``register_module``  (v4:13,20)
  * ``timm.layers.PatchEmbed`` -> Conv2d(k=s=patch) + flatten(2).transpose(1,2) (v4:15,75; restated from
    the public timm >= 0.9 source: norm_layer=None -> Identity, no size check when img_size=None)
  * ``AutoTokenizer.from_pretrained`` -> ``openpsg_b200.synth.SyntheticTokenizer`` (no vocab offline)
  * ``AutoModelForCausalLM.from_pretrained`` -> random-init ``OPTForCausalLM`` / ``LlamaForCausalLM`` of the requested dims
  * ``torch.Tensor.cuda`` -> identity (the reference hard-codes ``.cuda()`` on 9 lines)
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

REFERENCE_ROOT = Path("/root/reference")
V4_PATH = REFERENCE_ROOT / "kings_sgg/models/relation_heads/relation_transformer_head_v4.py"
CATS_PATH = REFERENCE_ROOT / "kings_sgg/models/detectors/mask2former_relation_v2.py"


def reference_available() -> bool:
    return V4_PATH.exists()


def _mod(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


class _Registry:
    def __init__(self):
        self.modules = {}

    def register_module(self, *a, **k):
        def deco(cls):
            self.modules[cls.__name__] = cls
            return cls
        return deco

    def build(self, cfg):
        cfg = dict(cfg)
        return self.modules[cfg.pop("type")](**cfg)


class _PatchEmbed(nn.Module):
    def __init__(self, img_size=None, patch_size=16, in_chans=3, embed_dim=768, **_):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=True)
        self.norm = nn.Identity()

    def forward(self, x):
        return self.norm(self.proj(x).flatten(2).transpose(1, 2))


def install_shims():
    """Register the stub modules.  Idempotent."""
    import transformers  # noqa: F401  (must be imported before the stubs so it never sees them)
    from transformers import InstructBlipQFormerModel, OPTForCausalLM  # noqa: F401
    if "mmdet.models.builder" in sys.modules and getattr(sys.modules["mmdet"], "_opsg_shim", False):
        return sys.modules["mmdet.models.builder"]
    mmcv = _mod("mmcv")
    runner = _mod("mmcv.runner")
    runner.BaseModule = nn.Module
    mmcv.runner = runner
    mmdet = _mod("mmdet")
    mmdet._opsg_shim = True
    core = _mod("mmdet.core")
    core.INSTANCE_OFFSET = 1000
    core.bbox2result = lambda *a, **k: None
    models = _mod("mmdet.models")
    builder = _mod("mmdet.models.builder")
    builder.HEADS = _Registry()
    builder.DETECTORS = _Registry()
    builder.build_head = builder.HEADS.build
    dets = _mod("mmdet.models.detectors")
    dets.Mask2Former = type("Mask2Former", (nn.Module,), {})
    ss = _mod("mmdet.models.detectors.single_stage")
    ss.SingleStageDetector = type("SingleStageDetector", (nn.Module,), {})
    mmdet.core, mmdet.models, models.builder, models.detectors = core, models, builder, dets
    timm = _mod("timm")
    layers = _mod("timm.layers")
    layers.PatchEmbed = _PatchEmbed
    timm.layers = layers
    return builder


def _load_by_path(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _load_categories_only():
    """Evaluate just the category lists of mask2former_relation_v2.py (:22-37); the rest of that file
    needs a real mmdet."""
    src = CATS_PATH.read_text()
    i, j = src.index("def replace_name"), src.index("@DETECTORS.register_module()")
    ns: dict = {}
    exec(compile(src[i:j], str(CATS_PATH), "exec"), ns)
    return ns["object_categories"], ns["relation_categories"]


def load_reference_head_class(llm_config: dict):
    """Return the reference ``RelationTransformerHeadV4`` class, file executed unmodified."""
    from transformers import AutoModelForCausalLM, AutoTokenizer, OPTConfig, OPTForCausalLM
    from openpsg_b200.synth import SyntheticTokenizer

    install_shims()
    for pkg in ("kings_sgg", "kings_sgg.models", "kings_sgg.models.detectors", "kings_sgg.models.relation_heads"):
        if pkg not in sys.modules:
            _mod(pkg)
    cats = _mod("kings_sgg.models.detectors.mask2former_relation_v2")
    cats.object_categories, cats.relation_categories = _load_categories_only()

    def _tok_from_pretrained(name, *a, subfolder=None, **k):
        tok = SyntheticTokenizer("qformer" if subfolder == "qformer_tokenizer" else "llm")
        if tok.kind == "llm":
            tok.set_vocab_size(llm_config["vocab_size"])
        return tok

    def _lm_from_pretrained(name, *a, **k):
        from openpsg_b200.synth import build_causal_lm
        return build_causal_lm(llm_config)                 # OPTForCausalLM or LlamaForCausalLM, random init

    AutoTokenizer.from_pretrained = staticmethod(_tok_from_pretrained)
    AutoModelForCausalLM.from_pretrained = staticmethod(_lm_from_pretrained)
    torch.Tensor.cuda = lambda self, *a, **k: self
    mod = _load_by_path("kings_sgg.models.relation_heads.relation_transformer_head_v4", V4_PATH)
    return mod.RelationTransformerHeadV4


class Capture:
    """Records the intermediates the golden fixtures pin (inputs/outputs of the reference's own
    sub-module calls), by wrapping bound forwards of one reference head instance."""

    def __init__(self, head):
        self.head = head
        self.rec: dict = {"generate": []}
        self._wrap(head.relation_qformer, "qformer")
        self._wrap(head.binary_rel_cls_pred, "exist_logits")
        self._wrap(head.language_projection, "lang_proj", many=True)
        self._wrap(head.patch_embed, "image_tokens")
        gen = head.language_model.generate

        def generate(*a, **k):
            out = gen(*a, **k)
            self.rec["generate"].append({
                "inputs_embeds": k["inputs_embeds"].detach().clone(),
                "attention_mask": k["attention_mask"].detach().clone(),
                "sequences": out.sequences.detach().clone(),
                "scores": torch.stack([s[0] for s in out.scores]).detach().clone(),
            })
            return out
        head.language_model.generate = generate

    def _wrap(self, module, key, many=False):
        fwd = module.forward

        def wrapped(*a, **k):
            out = fwd(*a, **k)
            val = out["last_hidden_state"] if key == "qformer" else out
            if key == "qformer":
                self.rec["qformer_kwargs"] = {kk: (vv.detach().clone() if torch.is_tensor(vv) else vv)
                                              for kk, vv in k.items()}
            if many:
                self.rec.setdefault(key, []).append(val.detach().clone())
            else:
                self.rec[key] = val.detach().clone()
            return out
        module.forward = wrapped


def run_reference(head, inputs: dict) -> dict:
    """``head.eval()(inputs)`` under no_grad; tolerates the terminal UnboundLocalError of the shipped
    default ``rel_cls_type='binary'`` (reference defect, v4:355 reads names only bound in the multiclass
    branch) — every computation has finished by then."""
    head.eval()
    cap = Capture(head)
    out = None
    with torch.no_grad():
        try:
            out = head(inputs)
        except UnboundLocalError as e:  # v4:355
            cap.rec["terminal_error"] = repr(e)
    cap.rec["output"] = out
    return cap.rec


OPENSEED_REL_PATH = REFERENCE_ROOT / "kings_sgg/models/detectors/openseed_relation.py"


def reference_object_embedding(pan_result, object_id_list, feature_map, meta_info, *, object_cls_embed=None,
                               embedding_add_cls=False, merge_cls_type="add", use_background_feature=False):
    """Row a11: run the reference's OWN statements of ``OpenSeeDRelation._get_input`` (detectors/openseed_relation.py, from the
    ``def _get_input`` line up to the text-database lookups, i.e. :430-493: mask chain, mask mean-pool, class embedding,
    background feature) on the given tensors.  The file itself cannot be imported (mmdet, detectron2, openseed), so the
    statements are sliced out of its source text and executed unmodified as a function body against a stand-in ``self``
    that only carries the attributes those lines read.  Returns ``object_embedding`` [1, n, C'] (None when no object)."""
    import textwrap
    import types as _types
    import torch.nn.functional as F
    src = OPENSEED_REL_PATH.read_text()
    i = src.index("    def _get_input(self, pan_result, object_id_list, object_score_list, feature_map, meta_info):")
    j = src.index("        use_text_db = self.relation_head.use_pair_text_vision_cross", i)
    body = textwrap.dedent(src[i:j]) + "    return object_embedding\n"
    ns = {"torch": torch, "F": F, "INSTANCE_OFFSET": 1000}
    exec(compile(body, str(OPENSEED_REL_PATH), "exec"), ns)
    self = _types.SimpleNamespace(object_cls_embed=object_cls_embed, embedding_add_cls=embedding_add_cls,
                                  merge_cls_type=merge_cls_type, add_postional_encoding=False,
                                  use_background_feature=use_background_feature)
    with torch.no_grad():
        return ns["_get_input"](self, pan_result, object_id_list, None, feature_map, meta_info)
