"""CPU tests of the head's TRAINING branch (plain PyTorch, row f3 of SURVEY.md §8) against loss values produced by the
UNMODIFIED reference head run in training mode under the shims (tests/golden/train_losses.pt, oracle/make_golden.py:
``make_train_golden``) with the same seeds for ``torch.randint`` (pair sampler), dropout and ``random.sample``."""
import random

import pytest
import torch

from openpsg_b200 import synth
from openpsg_b200.head import RelationTransformerHeadV4

TOL = 1e-4


def _head(llm_cfg, rel_cls_type):
    tok = synth.SyntheticTokenizer("llm")
    tok.set_vocab_size(llm_cfg["vocab_size"])
    head = RelationTransformerHeadV4(llm_feature_size=llm_cfg["hidden_size"], max_object_num=80, rel_cls_type=rel_cls_type,
                                     qformer_tokenizer=synth.SyntheticTokenizer("qformer"), llm_tokenizer=tok,
                                     language_model=synth.build_causal_lm(llm_cfg))
    synth.init_parameters(head, 0)
    return head


@pytest.mark.parametrize("llm_name,llm_cfg", [("opt", synth.OPT_TINY), ("llama", synth.LLAMA_TINY)])
@pytest.mark.parametrize("rel_cls_type", ["binary", "binary+multiclass"])
def test_training_losses_match_reference(golden, llm_name, llm_cfg, rel_cls_type):
    g = golden("train_losses")
    head = _head(llm_cfg, rel_cls_type)
    for dropout in (False, True):
        for image in (0, 1):
            head.train()
            if not dropout:
                head.relation_qformer.eval()
                head.language_model.eval()
            torch.manual_seed(100 + image)
            random.seed(100 + image)
            out = head(synth.make_train_inputs(synth.WORKLOADS["cfg1"], image))
            ref = g[(llm_name, rel_cls_type, dropout, image)]
            assert set(out) == set(ref) and all("loss" in k for k in out)
            for k in ref:
                assert abs(float(out[k]) - float(ref[k])) <= TOL * max(1.0, abs(float(ref[k]))), (k, dropout, image, float(out[k]), float(ref[k]))


def test_training_losses_backpropagate_to_the_trainable_parameters():
    head = _head(synth.OPT_TINY, "binary")
    for p in head.language_model.parameters():          # configs/psg/baseline_v4_ov.py:65 freezes relation_head.language_model
        p.requires_grad_(False)
    head.train()
    torch.manual_seed(1); random.seed(1)
    losses = head(synth.make_train_inputs(synth.WORKLOADS["cfg1"], 0))
    sum(losses.values()).backward()
    for name in ("patch_embed.proj.weight", "relation_query", "rel_cls_query", "binary_rel_cls_pred.weight",
                 "language_projection.weight", "relation_qformer.encoder.layer.1.crossattention.attention.query.weight"):
        grad = dict(head.named_parameters())[name].grad
        assert grad is not None and torch.isfinite(grad).all() and grad.abs().sum() > 0, name
    assert all(p.grad is None for p in head.language_model.parameters())


def test_eval_after_train_still_refuses_cpu_inference():
    head = _head(synth.OPT_TINY, "binary").eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        head(synth.make_image_inputs(synth.WORKLOADS["cfg1"], 0))
