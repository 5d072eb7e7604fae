"""TEST INFRASTRUCTURE — generates tests/golden/*.pt from the UNMODIFIED reference head.

Run in the authoring container (needs /root/reference):  python -m oracle.make_golden
Inputs and weights are pure functions of seeds (openpsg_b200/synth.py), so only the reference's
intermediates are stored, sub-sampled to keep the fixtures small.
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import torch

from openpsg_b200 import synth
from oracle import ref_shims

GOLDEN_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"
WEIGHT_SEED = 0


def build_reference_head(max_object_num=80, llm_config=None, **kwargs):
    llm_config = llm_config or synth.OPT_TINY
    cls = ref_shims.load_reference_head_class(llm_config)
    head = cls(llm_feature_size=llm_config["hidden_size"], max_object_num=max_object_num, **kwargs)
    synth.init_parameters(head, WEIGHT_SEED)
    return head


def _summarise(rec: dict, keep_pairs) -> dict:
    kw = rec["qformer_kwargs"]
    out33 = rec["qformer"][:, :33]
    g = {
        "image_tokens": rec["image_tokens"][0].clone(),
        "input_ids": kw["input_ids"].clone(),
        "attention_mask": kw["attention_mask"].clone(),
        "pair_masks": kw["encoder_attention_mask"][:, 0, :].clone(),     # bool [B, L]
        "keep_pairs": torch.tensor(keep_pairs),
        "qformer_out_keep": out33[keep_pairs].clone(),                     # [K,33,768]
        "cls_feature": out33[:, 0].clone(),                                # [B,768]
        "exist_logits": rec["exist_logits"][:, 0].clone(),
        "terminal_error": rec.get("terminal_error"),
    }
    gens = rec["generate"]
    g["selected_embeds0"] = gens[0]["inputs_embeds"][0].clone() if gens else None
    g["llm_masks"] = torch.stack([x["attention_mask"][0] for x in gens]) if gens else None
    g["sequences"] = [x["sequences"][0].clone() for x in gens]
    g["scores_first2"] = [x["scores"].clone() for x in gens[:2]]
    g["lang_proj_first2"] = [x.clone() for x in rec.get("lang_proj", [])[:2]]
    # the reference's own selection: recompute exactly as v4:236-237 from its probabilities
    prob = torch.sigmoid(rec["exist_logits"])
    g["selected"] = prob.squeeze(1).topk(prob.shape[0]).indices.tolist()[:20]
    return g


def main(argv=None):
    if not ref_shims.reference_available():
        print("reference not present; golden fixtures can only be generated in the authoring container")
        return 1
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    head = build_reference_head()
    cases = {
        "cfg1": (synth.make_image_inputs(synth.WORKLOADS["cfg1"], 0), list(range(0, 64, 4))),
        "stress": (synth.make_stress_inputs(), [0, 6, 8, 13, 20, 27, 34, 41, 42, 47, 48]),
        "cfg2": (synth.make_image_inputs(synth.WORKLOADS["cfg2"], 0), [0, 41, 399, 777, 1234, 1599]),
    }
    for name, (inputs, keep) in cases.items():
        t = time.time()
        rec = ref_shims.run_reference(head, inputs)
        g = _summarise(rec, keep)
        if name == "cfg2":   # 1600x768 fp32 cls rows would be 4.9 MB; the logits pin them
            g["cls_feature"] = g["cls_feature"][keep].clone()
            g["pair_masks"] = g["pair_masks"][keep].clone()
            g["input_ids_keep"] = g.pop("input_ids")[keep].clone()
            g["attention_mask_keep"] = g.pop("attention_mask")[keep].clone()
            g["image_tokens"] = g["image_tokens"][::16].clone()
        torch.save(g, GOLDEN_DIR / f"{name}.pt")
        print(f"{name}: {time.time() - t:.1f}s  ->  {(GOLDEN_DIR / (name + '.pt')).stat().st_size / 1e6:.2f} MB")
    make_llama_golden()
    make_train_golden()
    make_mask_pool_golden()
    return 0


def make_mask_pool_golden():
    """Row a11: object embeddings produced by the reference's own statements (ref_shims.reference_object_embedding)."""
    feat, pan, ids, meta, table = synth.make_mask_pool_case()
    emb = torch.nn.Embedding(133, 256)
    with torch.no_grad():
        emb.weight.copy_(table)
    g = {}
    for mode, add_cls, merge, bg in (("plain", False, "add", False), ("add", True, "add", False), ("cat", True, "cat", False),
                                     ("bg", False, "add", True), ("add+bg", True, "add", True)):
        out = ref_shims.reference_object_embedding(pan, ids, feat, meta, object_cls_embed=emb, embedding_add_cls=add_cls,
                                                   merge_cls_type=merge, use_background_feature=bg)
        g[mode] = out[0].clone()
        print("mask_pool", mode, tuple(out.shape))
    torch.save(g, GOLDEN_DIR / "mask_pool.pt")


def run_reference_train(head, inputs, seed, dropout):
    """The reference's TRAIN branch (v4:114-133,187-204,267-285,327-341) with seeded samplers.  ``dropout=False`` puts the
    two HF sub-modules in eval mode (the head itself stays in training mode) so the losses do not depend on dropout draws."""
    import random
    head.train()
    if not dropout:
        head.relation_qformer.eval()
        head.language_model.eval()
    torch.manual_seed(seed)
    random.seed(seed)
    out = head(inputs)
    return {k: v.detach().clone() for k, v in out.items()}


def make_train_golden():
    """Loss values of the unmodified reference head in training mode on a synthetic train-mode image (cfg1 geometry)."""
    t = time.time()
    g = {}
    for llm_name, llm_cfg in (("opt", synth.OPT_TINY), ("llama", synth.LLAMA_TINY)):
        for rel_cls_type in ("binary", "binary+multiclass"):
            head = build_reference_head(llm_config=llm_cfg, rel_cls_type=rel_cls_type)
            for dropout in (False, True):
                for image in (0, 1):
                    inputs = synth.make_train_inputs(synth.WORKLOADS["cfg1"], image)
                    key = (llm_name, rel_cls_type, dropout, image)
                    g[key] = run_reference_train(head, inputs, 100 + image, dropout)
                    print(key, {k: round(float(v), 6) for k, v in g[key].items()})
    torch.save(g, GOLDEN_DIR / "train_losses.pt")
    print(f"train_losses: {time.time() - t:.1f}s")


def make_llama_golden():
    """The shipped config's LLM family (configs/psg/baseline_v4_ov.py:60-61): the unmodified reference head with a tiny
    random LlamaForCausalLM behind the ``AutoModelForCausalLM.from_pretrained`` shim, cfg1 image.  Only the LLM leg is
    stored (the relation-query leg is the one cfg1.pt already pins)."""
    t = time.time()
    head = build_reference_head(llm_config=synth.LLAMA_TINY)
    rec = ref_shims.run_reference(head, synth.make_image_inputs(synth.WORKLOADS["cfg1"], 0))
    g = _summarise(rec, list(range(0, 64, 4)))
    keep = {k: g[k] for k in ("exist_logits", "selected", "selected_embeds0", "llm_masks", "sequences", "scores_first2",
                              "lang_proj_first2", "terminal_error")}
    keep["qformer_out_selected"] = rec["qformer"][:, :33][g["selected"][:4]].clone()     # rows the LLM leg consumed
    torch.save(keep, GOLDEN_DIR / "cfg1_llama.pt")
    print(f"cfg1_llama: {time.time() - t:.1f}s  ->  {(GOLDEN_DIR / 'cfg1_llama.pt').stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    sys.exit(main())
