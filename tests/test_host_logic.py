"""CPU tests of the host-side mirror of the reference interface (no kernels involved)."""
import numpy as np
import torch

from openpsg_b200 import synth
from openpsg_b200.categories import object_categories, relation_categories
from openpsg_b200.head import PairInstructionCache
from tests.helpers import build_port_head, build_product_head


def test_parameter_names_match_the_reference_contract():
    """Checkpoint contract: same names/shapes as the reference head (here: its port, which is
    pinned against the unmodified reference by tests/test_oracle.py through identical seeded weights)."""
    prod = build_product_head(llm=synth.OPT_TINY)
    port = build_port_head(llm=synth.OPT_TINY)
    a = {k: tuple(v.shape) for k, v in prod.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in port.state_dict().items()}
    assert a == b
    for k in ("patch_embed.proj.weight", "relation_query", "rel_cls_query", "binary_rel_cls_pred.weight",
              "language_projection.weight", "relation_qformer.encoder.layer.1.crossattention.attention.key.weight",
              "language_model.model.decoder.layers.0.fc1.weight"):
        assert k in a
    for k in a:
        assert torch.equal(prod.state_dict()[k], port.state_dict()[k]), k
    port.load_state_dict(prod.state_dict())       # reference-named checkpoints load both ways


def test_registry_builds_head_from_config_dict():
    from openpsg_b200.registry import HEADS, build_head
    import kings_sgg.models.relation_heads.relation_transformer_head_v4 as dropin   # custom_imports path
    assert dropin.RelationTransformerHeadV4 is HEADS.get("RelationTransformerHeadV4")
    head = build_head(dict(type="RelationTransformerHeadV4", qformer_model_name="x", llm_model_name="y",
                           relation_classes=relation_categories, language_model=False,
                           qformer_tokenizer=synth.SyntheticTokenizer("qformer"), some_future_kwarg=1))
    assert head.max_object_num == 30 and head.topk_pairs == 20 and head.max_new_tokens == 16
    assert head.num_relation_classes == 56


def test_pair_instruction_cache_equals_batch_tokenisation():
    rs = np.random.RandomState(0)
    for kind, side, tmpl in (("qformer", "right", 'Is there a relation between {} and {}?'),
                             ("llm", "left", 'What are the relations between {} and {}? Assistant: ')):
        tok = synth.SyntheticTokenizer(kind)
        cache = PairInstructionCache(synth.SyntheticTokenizer(kind), tmpl, object_categories, side)
        for _ in range(3):
            a, b = rs.randint(0, 133, 50), rs.randint(0, 133, 50)
            tok.padding_side = side
            enc = tok([tmpl.format(object_categories[i], object_categories[j]) for i, j in zip(a, b)])
            ids, mask = cache.lookup(a, b)
            T = int(enc["attention_mask"].sum(1).max())
            ref_ids = enc["input_ids"][:, :T] if side == "right" else enc["input_ids"][:, -T:]
            ref_mask = enc["attention_mask"][:, :T] if side == "right" else enc["attention_mask"][:, -T:]
            assert torch.equal(mask.long(), ref_mask)
            assert torch.equal(ids.long() * mask.long(), ref_ids * ref_mask)


def test_category_tables():
    assert object_categories[0] == "person" and object_categories[132] == "rug" and object_categories[86] == "door"
    assert relation_categories[0] == "over" and relation_categories[55] == "leaning on"


def test_workload_table():
    w = synth.WORKLOADS["cfg2"]
    assert (w.queries, w.ordered_pairs, w.image_tokens) == (1600, 1560, 256)
    assert synth.WORKLOADS["cfg5"].ordered_pairs == 6320


def test_parse_relations_equals_the_reference_loop():
    """head._parse_relations (dict / set) against the reference's statements v4:315-326 (list membership, list.index, list-of-lists
    dedupe) on random generations, repeated names, unknown names and texts without the <s> marker."""
    import numpy as np
    from openpsg_b200.categories import relation_categories
    from openpsg_b200.head import RelationTransformerHeadV4

    def reference_loop(selected, n, texts):
        rel_pred, rel_score = [], []
        for si, text in zip(selected, texts):
            parts = text.split('<s>')
            body = (parts[1] if len(parts) > 1 else parts[0]).split('</s>')[0].strip()
            for name in body.split('  '):
                if name in relation_categories:
                    trip = [si // n, si % n, relation_categories.index(name)]
                    if trip not in rel_pred:
                        rel_pred.append(trip)
                        rel_score.append(1)
        return rel_pred, rel_score

    class _Holder:
        pass
    rng = np.random.default_rng(0)
    texts = []
    for _ in range(60):
        names = [relation_categories[i] for i in rng.integers(0, len(relation_categories), 12)]
        names.insert(int(rng.integers(0, 12)), "not a relation")
        texts.append("<s> " + "  ".join(names) + "</s> trailing")
    texts += ["over  over  beside</s> junk", "<s> nothing here</s>", "", "<s> on  on  in front of  on</s>"]
    selected = [int(x) for x in rng.integers(0, 64, len(texts))]          # repeated pairs included
    assert RelationTransformerHeadV4._parse_relations(_Holder(), selected, 8, texts) == reference_loop(selected, 8, texts)
