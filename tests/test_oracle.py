"""CPU tests: the oracle (oracle/restated.py, oracle/ref_port.py) against the golden vectors produced by
the UNMODIFIED reference head (oracle/make_golden.py, run where /root/reference exists)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from openpsg_b200 import synth
from oracle import restated
from oracle.ref_port import ReferencePortHead

TOL = 2e-4   # fp32 restatement vs fp32 reference (different summation order / dedup of K,V)


@pytest.fixture(scope="module")
def port_head():
    torch.manual_seed(0)
    head = ReferencePortHead(synth.OPT_TINY, llm_feature_size=synth.OPT_TINY["hidden_size"], max_object_num=80)
    synth.init_parameters(head, 0)
    return head.eval()


@pytest.fixture(scope="module")
def sd(port_head):
    return {k: v.detach() for k, v in port_head.state_dict().items()}


def _inputs(name):
    if name == "stress":
        return synth.make_stress_inputs()
    return synth.make_image_inputs(synth.WORKLOADS[name], 0)


def _obj_masks(inputs):
    meta, info = inputs["img_metas"][0], inputs["object_info"][0]
    return restated.object_token_masks(
        info["pan_results"].numpy(), meta["img_shape"][:2], meta["pad_shape"][:2],
        inputs["mask_features"].shape[-2:], 16, [int(i) for i in info["object_id_list"]])


@pytest.mark.parametrize("sizes", [(480, 427), (427, 640), (640, 10), (1333, 21), (800, 12), (7, 50), (256, 4), (3, 3)])
def test_legacy_nearest_matches_interpolate(sizes):
    n_in, n_out = sizes
    src = torch.arange(n_in, dtype=torch.float32)[None, None, :, None].expand(1, 1, n_in, 2)
    ref = F.interpolate(src, size=(n_out, 2), mode="nearest")[0, 0, :, 0].long().numpy()
    assert np.array_equal(restated.legacy_nearest_index(n_out, n_in), ref)


@pytest.mark.parametrize("name", ["cfg1", "stress", "cfg2"])
def test_masks_bit_exact(golden, name):
    g, inputs = golden(name), _inputs(name)
    m = _obj_masks(inputs)
    pm = restated.pair_masks(m)
    keep = g["keep_pairs"].numpy() if name == "cfg2" else np.arange(pm.shape[0])
    assert np.array_equal(pm[keep], g["pair_masks"].numpy())
    if name == "stress":
        assert not m[6].any(), "7th object owns no token: empty mask"
        assert not pm[48].any()
    bits = restated.pack_mask_bits(m)
    L = m.shape[1]
    unpacked = ((bits[:, np.arange(L) // 32] >> (np.arange(L) % 32).astype(np.uint32)) & 1).astype(bool)
    assert np.array_equal(unpacked, m)


@pytest.mark.parametrize("name", ["cfg1", "stress", "cfg2"])
def test_patch_embed(golden, sd, name):
    g, inputs = golden(name), _inputs(name)
    x = restated.patch_embed(inputs["mask_features"], sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], 16)
    ref = g["image_tokens"]
    if name == "cfg2":
        x = x[::16]
    assert x.shape == ref.shape
    assert (x - ref).abs().max() < TOL


@pytest.mark.parametrize("name", ["cfg1", "stress", "cfg2"])
def test_qformer_restatement(golden, sd, name):
    g, inputs = golden(name), _inputs(name)
    keep = g["keep_pairs"]
    m = torch.from_numpy(_obj_masks(inputs))
    tokens = restated.patch_embed(inputs["mask_features"], sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], 16)
    query = torch.cat([sd["rel_cls_query"], sd["relation_query"]], dim=1)[0]
    if name == "cfg2":
        ids, tmask = g["input_ids_keep"], g["attention_mask_keep"][:, 33:]
    else:
        ids, tmask = g["input_ids"][keep], g["attention_mask"][keep][:, 33:]
    out = restated.qformer_forward(sd, query, ids, tmask, tokens, m, pair_index=keep)
    assert (out - g["qformer_out_keep"]).abs().max() < TOL
    z = restated.existence_logits(out[:, 0], sd["binary_rel_cls_pred.weight"], sd["binary_rel_cls_pred.bias"])
    assert (z - g["exist_logits"][keep]).abs().max() < TOL
    assert torch.isfinite(out).all()


def test_synthetic_tokenizer_matches_reference_ids(golden):
    g = golden("cfg1")
    from openpsg_b200.categories import object_categories
    ids = synth.object_ids(8)
    names = [object_categories[i % 1000] for i in ids]
    texts = ['Is there a relation between {} and {}?'.format(names[p // 8], names[p % 8]) for p in range(64)]
    enc = synth.SyntheticTokenizer("qformer")(texts)
    assert torch.equal(enc["input_ids"], g["input_ids"])
    assert torch.equal(enc["attention_mask"], g["attention_mask"][:, 33:])


@pytest.mark.parametrize("name", ["cfg1", "stress", "cfg2"])
def test_topk_given_reference_logits(golden, name):
    g = golden(name)
    z = g["exist_logits"]
    sel = restated.topk_pairs(z, 20)
    assert sel == g["selected"]
    assert np.array_equal(restated.existence_mask(z), (torch.sigmoid(z) > 0.5).numpy())


def test_ref_port_matches_reference_cfg1(golden, port_head):
    g = golden("cfg1")
    q = port_head.relation_queries(_inputs("cfg1"))
    assert (q["exist_logits"] - g["exist_logits"]).abs().max() < TOL
    assert (q["qformer_out"][g["keep_pairs"]] - g["qformer_out_keep"]).abs().max() < TOL
    assert q["selected"] == g["selected"]
    d = port_head.decode_relations(q, max_pairs=3)
    for got, ref in zip(d["sequences"], g["sequences"][:3]):
        n = min(len(got), len(ref))     # the reference may stop at EOS; the port disables EOS
        assert torch.equal(got[:n], ref[:n])
    got, ref = d["scores"][0][: g["scores_first2"][0].shape[0]].clone(), g["scores_first2"][0].clone()
    got[:, 2] = ref[:, 2] = 0     # min_new_tokens masks the EOS logit (id 2) to -inf in the port
    assert (got - ref).abs().max() < 1e-3


def test_llm_restatement_matches_reference(golden, sd):
    g = golden("cfg1")
    embeds0 = g["selected_embeds0"][None]             # what the reference fed to generate for pair 0
    mask0 = g["llm_masks"][:1]
    n_new = g["scores_first2"][0].shape[0]
    toks, scores = restated.opt_greedy_decode(sd, synth.OPT_TINY, embeds0, mask0, n_new)
    assert torch.equal(toks[0], g["sequences"][0][:n_new])
    assert (scores[0] - g["scores_first2"][0]).abs().max() < 1e-3
    # a9: projection of the gathered pair feature (v4:294)
    sel0 = g["selected"][0]
    keep = g["keep_pairs"].tolist()
    if sel0 in keep:
        feat = g["qformer_out_keep"][keep.index(sel0)][1:]
        u = feat @ sd["language_projection.weight"].t() + sd["language_projection.bias"]
        assert (u - g["lang_proj_first2"][0]).abs().max() < TOL


def test_mask_pool_matches_reference_formula():
    # a11: (feat*mask).sum/(mask.sum+1e-8) and cat(obj[i], obj[j])  (detectors/openseed_relation.py:454-468,502-527)
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(16, 12, 10, generator=g)
    masks = torch.rand(5, 12, 10, generator=g) > 0.6
    masks[4] = False
    obj, pair = restated.mask_pool_pairs(feat, masks)
    ref = (feat[None] * masks[:, None].float()).sum(dim=(2, 3)) / (masks.float().sum(dim=(1, 2))[:, None] + 1e-8)
    assert (obj - ref).abs().max() < 1e-5
    assert torch.equal(pair[7], torch.cat([obj[1], obj[2]]))
    assert obj[4].abs().max() == 0


# ---- Llama family (the LLM the shipped config names) -----------------------------------------------------------------

@pytest.fixture(scope="module")
def llama_port_head():
    head = ReferencePortHead(synth.LLAMA_TINY, llm_feature_size=synth.LLAMA_TINY["hidden_size"], max_object_num=80)
    synth.init_parameters(head, 0)
    return head.eval()


def test_llama_restatement_matches_reference(golden, llama_port_head):
    """restated.llama_greedy_decode against what the UNMODIFIED reference head produced with a tiny LlamaForCausalLM behind
    its from_pretrained call (tests/golden/cfg1_llama.pt): ids and per-step scores of the first two generate calls."""
    g = golden("cfg1_llama")
    sd = {k: v.detach() for k, v in llama_port_head.state_dict().items()}
    embeds0 = g["selected_embeds0"][None]
    n_new = g["scores_first2"][0].shape[0]
    toks, scores = restated.llama_greedy_decode(sd, synth.LLAMA_TINY, embeds0, g["llm_masks"][:1], n_new)
    assert torch.equal(toks[0], g["sequences"][0][:n_new])
    assert (scores[0] - g["scores_first2"][0]).abs().max() < 1e-3
    # a9 on the Llama width: projection of the first selected pair's 32 relation rows
    feat = g["qformer_out_selected"][0][1:]
    u = feat @ sd["language_projection.weight"].t() + sd["language_projection.bias"]
    assert (u - g["lang_proj_first2"][0]).abs().max() < TOL
    emb, mask = restated.build_llm_prefix(sd, g["qformer_out_selected"][:1, 1:], torch.zeros((1, 0), dtype=torch.long),
                                          torch.zeros((1, 0), dtype=torch.long), embed_key=restated.embed_tokens_key(synth.LLAMA_TINY))
    assert (emb[0] - g["selected_embeds0"][:32]).abs().max() < TOL


def test_llama_port_matches_reference(golden, llama_port_head):
    g = golden("cfg1_llama")
    q = llama_port_head.relation_queries(_inputs("cfg1"))
    assert (q["exist_logits"] - g["exist_logits"]).abs().max() < TOL
    assert q["selected"] == g["selected"]
    d = llama_port_head.decode_relations(q, max_pairs=3)
    for got, ref in zip(d["sequences"], g["sequences"][:3]):
        n = min(len(got), len(ref))
        assert torch.equal(got[:n], ref[:n])


def test_llama_gqa_restatement_matches_hf():
    """Grouped-query attention + left-padded batch: the restatement against the live HF LlamaForCausalLM forward."""
    cfg = synth.LLAMA_TINY_GQA
    lm = synth.build_causal_lm(cfg).eval()
    synth.init_parameters(lm, 3)
    sd = {"language_model." + k: v.detach() for k, v in lm.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    embeds = torch.randn((2, 9, cfg["hidden_size"]), generator=g) * 0.1
    mask = torch.ones((2, 9), dtype=torch.long)
    mask[1, 3:5] = 0                                      # pads mid-sequence like [prefix ; left-padded text] (v4:298-299)
    with torch.no_grad():
        ref = lm(inputs_embeds=embeds, attention_mask=mask, position_ids=restated.llama_positions(mask)).logits
    got = restated.llama_forward(sd, cfg, embeds, mask)
    valid = mask.bool()
    assert (got - ref)[valid].abs().max() < 1e-4


MASK_POOL_MODES = (("plain", "none", False), ("add", "add", False), ("cat", "cat", False), ("bg", "none", True), ("add+bg", "add", True))


def test_mask_pool_chain_matches_reference_statements(golden):
    """Row a11: restated.mask_pool_chain against tests/golden/mask_pool.pt, which holds what the reference's OWN statements
    (detectors/openseed_relation.py:430-493, sliced out of the file and executed by oracle/ref_shims.reference_object_embedding)
    return: mask chain pan==id -> nearest -> pad -> nearest, mean-pool, class embedding add / cat, background feature."""
    g = golden("mask_pool")
    feat, pan, ids, meta, table = synth.make_mask_pool_case()
    for name, cls_mode, bg in MASK_POOL_MODES:
        obj, pair = restated.mask_pool_chain(feat[0], pan.numpy(), meta["img_shape"][:2], meta["pad_shape"][:2], ids, table, cls_mode, bg)
        assert obj.shape == g[name].shape
        assert (obj - g[name]).abs().max() < 1e-5, name
        n = len(ids)
        assert torch.equal(pair[1 * n + 3], torch.cat([obj[1], obj[3]]))
    m = restated.object_masks_feature_res(pan.numpy(), meta["img_shape"][:2], meta["pad_shape"][:2], feat.shape[-2:], ids)
    assert not m[6].any() and not m[8].any(), "the pixel-less object and the unlisted id own nothing"
    assert np.array_equal(m[7], m[2]), "an object repeating an id shares the mask"
