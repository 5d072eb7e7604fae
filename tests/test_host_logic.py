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


def test_decode_group_pads_prompts_and_splits_results_per_image():
    """head._decode_group (the LLM leg of forward_batch over a group of images) on CPU with a recording stand-in for the
    engine: prompts of different images are LEFT-padded to the group's longest one with masked pad tokens (v4:260-266 pads the
    same way inside one image), the stacked rows / top-k reach the engine in image order, every image gets its own slice of
    the generated tokens (on_result sees it as last_generation), images without objects yield empty results."""
    from types import SimpleNamespace
    from openpsg_b200.llm import GenerationOutput
    head = build_product_head(llm=synth.OPT_TINY, topk_pairs=4, max_new_tokens=5)
    head._packed = SimpleNamespace(device=torch.device("cpu"))
    calls = []

    class Engine:
        def generate_rows(self, rows, ids, mask, max_new_tokens=16):
            calls.append((rows.clone(), ids.clone(), mask.clone(), max_new_tokens))
            k = ids.shape[0]
            # sequence s "generates" relation class s % 56 five times (SyntheticTokenizer decodes token t to class t % 56)
            return GenerationOutput(tokens=(torch.arange(k, dtype=torch.int32)[:, None] % 56).repeat(1, max_new_tokens))
    head._llm_engine = Engine()
    width = 33 * 768
    n_a, n_b = 5, 3                                               # objects per image: 25 and 9 pair queries
    cats_a = np.array([0, 1, 2, 3, 4]); cats_b = np.array([130, 131, 132])      # different names -> different prompt lengths
    bufs = (torch.zeros((8, width), dtype=torch.bfloat16), torch.zeros((8,), dtype=torch.int32))
    top_a = torch.tensor([7, 0, 24, 13], dtype=torch.int32); top_b = torch.tensor([8, 1, 4], dtype=torch.int32)   # image b: k = 3 < topk
    bufs[1][:4] = top_a; bufs[1][4:7] = top_b
    bufs[0][:7, 0] = torch.arange(7, dtype=torch.bfloat16)
    rec_a = dict(prep=dict(n=n_a, cats=cats_a), out=object(), rows=bufs[0][:4], topk=bufs[1][:4], slot=0)
    rec_empty = dict(prep=dict(n=0), out=None)
    rec_b = dict(prep=dict(n=n_b, cats=cats_b), out=object(), rows=bufs[0][4:7], topk=bufs[1][4:7], slot=4)
    seen = []
    results = head._decode_group([rec_a, rec_empty, rec_b], bufs, on_result=lambda h: seen.append(h.last_generation.tokens.clone()))
    assert len(calls) == 1 and len(results) == 3 and results[1] == {"rel_pred": [], "rel_score": []}
    rows, ids, mask, mnt = calls[0]
    assert mnt == 5 and rows.shape == (7, width) and torch.equal(rows[:, 0].float(), torch.arange(7.0))
    # prompts: each image's own lookup, left-padded to the longest of the group
    ia, ma = head._llm_cache.lookup(cats_a[top_a.numpy() // n_a], cats_a[top_a.numpy() % n_a])
    ib, mb = head._llm_cache.lookup(cats_b[top_b.numpy() // n_b], cats_b[top_b.numpy() % n_b])
    T = max(ia.shape[1], ib.shape[1])
    assert ids.shape == (7, T) and mask.shape == (7, T) and ia.shape[1] != ib.shape[1]
    for got_i, got_m, want_i, want_m in ((ids[:4], mask[:4], ia, ma), (ids[4:], mask[4:], ib, mb)):
        pad = T - want_i.shape[1]
        assert torch.equal(got_i[:, pad:], want_i) and torch.equal(got_m[:, pad:], want_m)
        assert (got_m[:, :pad] == 0).all() and (got_i[:, :pad] == head._llm_cache.pad_id).all()
    # results: sequence s of the stacked batch said "relation s % 56" -> one triple per selected pair, in top-k order
    assert results[0]["rel_pred"] == [[7 // 5, 7 % 5, 0], [0, 0, 1], [24 // 5, 24 % 5, 2], [13 // 5, 13 % 5, 3]]
    assert results[2]["rel_pred"] == [[8 // 3, 8 % 3, 4], [0, 1, 5], [4 // 3, 4 % 3, 6]]
    assert [t.shape for t in seen] == [(4, 5), (3, 5)] and int(seen[1][0, 0]) == 4
