"""GPU parity of the Llama-family decode path — the LLM ``configs/psg/baseline_v4_ov.py:60-61`` names (rows a9-a10 / f4 of
SURVEY.md §8) — against the fp32 oracle (oracle/restated.py, pinned to the unmodified reference in tests/test_oracle.py)
and against the reference's own golden vectors (tests/golden/cfg1_llama.pt).

Tolerance: per-step logits |d| <= 1e-1 + 5e-2*|ref|; greedy ids must match wherever the oracle's
top-1 / top-2 gap exceeds 2x that tolerance."""
import math

import pytest
import torch

from openpsg_b200 import ops, synth
from openpsg_b200.categories import object_categories, relation_categories
from oracle import restated
from tests.helpers import build_product_head, margin_set_equal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _logit_tol(ref):
    return 1e-1 + 5e-2 * ref.abs()


def _check_scores(got, ref_scores, ref_toks, got_toks, tag, min_match=0.9):
    err = (got - ref_scores).abs()
    assert (err <= _logit_tol(ref_scores)).all(), (tag, (err - _logit_tol(ref_scores)).max())
    top2 = ref_scores.topk(2, dim=-1).values
    decided = (top2[..., 0] - top2[..., 1]) > 2 * _logit_tol(top2[..., 0])
    match = (got_toks == ref_toks).float().mean().item()
    print(f"[{tag}] max|dlogit|={err.max():.4f} (ref absmax {ref_scores.abs().max():.2f}) decided={decided.float().mean():.2f} "
          f"id_match={match:.3f}")
    assert torch.equal(got_toks[decided], ref_toks[decided])
    assert match >= min_match
    return decided


# ---- kernels ----------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("rows,cols,ld", [(7, 256, 256), (100, 4096, 4096), (33, 2560, 2568), (5, 8192, 8192)])
def test_rmsnorm_kernel(rows, cols, ld):
    g = torch.Generator().manual_seed(rows + cols)
    x = (torch.randn((rows, ld), generator=g) * 3).to(torch.bfloat16)
    w = 1 + 0.1 * torch.randn(cols, generator=g)
    xs = x.to(DEV)[:, :cols]
    y = ops.rmsnorm(xs, w.to(DEV), 1e-5)
    xf = x[:, :cols].float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5))
    assert (y.float().cpu() - ref).abs().max() <= 2e-2 * ref.abs().max()
    assert torch.equal(y.cpu(), ref.to(torch.bfloat16)) or (y.float().cpu() - ref).abs().max() <= 4e-2
    y2 = ops.rmsnorm(xs, w.to(DEV), 1e-5)
    assert torch.equal(y, y2), "bit-reproducible"


@pytest.mark.parametrize("heads,hd", [(2, 128), (4, 64), (3, 80)])
def test_rope_kernel(heads, hd):
    g = torch.Generator().manual_seed(hd)
    rows, d = 37, heads * hd
    x = torch.randn((rows, 3 * d), generator=g).to(torch.bfloat16)
    pos = torch.randint(0, 300, (rows,), generator=g, dtype=torch.int32)
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    fr = torch.arange(512, dtype=torch.float32)[:, None] * inv[None]
    xd = x.to(DEV).clone()
    ops.rope(xd, 2, heads, hd, pos.to(DEV), fr.cos().to(DEV), fr.sin().to(DEV))
    xf = x.float().reshape(rows, 3, heads, hd)
    c = torch.cat([fr, fr], -1).cos()[pos.long()][:, None, None, :]
    s = torch.cat([fr, fr], -1).sin()[pos.long()][:, None, None, :]
    rot = torch.cat([-xf[..., hd // 2:], xf[..., : hd // 2]], -1)
    ref = xf.clone()
    ref[:, :2] = (xf * c + rot * s)[:, :2]
    got = xd.float().cpu().reshape(rows, 3, heads, hd)
    assert torch.equal(got[:, 2], xf[:, 2]), "v part untouched"
    assert (got - ref).abs().max() <= 2e-2


def test_swiglu_kernel():
    g = torch.Generator().manual_seed(1)
    gu = torch.randn((19, 2 * 704), generator=g).to(torch.bfloat16)
    out = ops.swiglu(gu.to(DEV), 704)
    ref = torch.nn.functional.silu(gu[:, :704].float()) * gu[:, 704:].float()
    assert (out.float().cpu() - ref).abs().max() <= 2e-2


@pytest.mark.parametrize("pos_offset", [0, 2])
def test_prompt_layout_kernel(pos_offset):
    g = torch.Generator().manual_seed(2)
    k, T, npre, new = 13, 17, 32, 6
    lens = torch.randint(1, T + 1, (k,), generator=g)
    mask = (torch.arange(T)[None, :] >= (T - lens)[:, None]).to(torch.int32)          # left padded
    lay = ops.llm_prompt_layout(mask.to(DEV), npre, new, pos_offset)
    full = torch.cat([torch.ones((k, npre), dtype=torch.long), mask.long()], 1)
    cs = torch.cumsum(full, 1)
    if pos_offset == 2:
        assert torch.equal(lay.pos.cpu().long(), restated.opt_positions(full))
    else:
        assert torch.equal(lay.pos.cpu().long()[full.bool()], restated.llama_positions(full)[full.bool()])
    km = lay.key_mask.cpu()
    assert torch.equal(km[:, :npre + T].long(), full) and bool((km[:, npre + T:] == 1).all())
    assert torch.equal(lay.last_rows.cpu().long(), torch.arange(k) * (npre + T) + npre + T - 1)
    want = cs[:, -1][None, :] + torch.arange(new - 1)[:, None] + pos_offset
    assert torch.equal(lay.dec_pos.cpu().long(), want)


def test_copy_and_transpose_kernels():
    for n in (1, 17, 4096, 1 << 20):
        src = torch.randint(0, 255, (n + 3,), dtype=torch.uint8, device=DEV)
        dst = torch.zeros_like(src)
        ops.copy_into(dst[3:], src[3:])            # unaligned -> byte path
        assert torch.equal(dst[3:], src[3:])
    a = torch.randn((256, 300), device=DEV)
    b = torch.empty_like(a)
    ops.copy_into(b, a)
    assert torch.equal(a, b)
    t = torch.arange(7 * 11, dtype=torch.int32, device=DEV).reshape(7, 11).contiguous()
    o = torch.empty((11, 7), dtype=torch.int32, device=DEV)
    ops.transpose_i32(t, o)
    assert torch.equal(o, t.t().contiguous())


# ---- engine vs oracle -------------------------------------------------------------------------------------------------

def _prompts(inputs, sel, n, vocab):
    ids = [int(i) for i in inputs["object_info"][0]["object_id_list"]]
    names = [object_categories[i % 1000] for i in ids]
    tok = synth.SyntheticTokenizer("llm")
    tok.set_vocab_size(vocab)
    enc = tok(['What are the relations between {} and {}? Assistant: '.format(names[s // n], names[s % n]) for s in sel])
    return enc["input_ids"].to(torch.int32), enc["attention_mask"].to(torch.int32)


def _oracle(head, cfg, hidden, selected, l_ids, l_mask, n_new):
    sd = {k: v.detach().float().cpu() for k, v in head.state_dict().items()}
    feat = hidden.float().cpu().reshape(-1, 33, 768)[selected.cpu().long()][:, 1:]
    embeds, mask = restated.build_llm_prefix(sd, feat, l_ids.long(), l_mask.long(), embed_key=restated.embed_tokens_key(cfg))
    toks, scores = restated.greedy_decode(sd, cfg, embeds, mask, n_new)
    return embeds, mask, toks, scores


@pytest.fixture(scope="module")
def llama_head():
    return build_product_head(llm=synth.LLAMA_TINY, device=DEV)


@pytest.mark.parametrize("name", ["cfg1", "stress"])
def test_llama_decode_matches_oracle(llama_head, name):
    head = llama_head
    inputs = synth.make_stress_inputs() if name == "stress" else synth.make_image_inputs(synth.WORKLOADS[name], 0)
    head(synth.inputs_to(inputs, DEV), is_generation=False)
    out = head.last_output
    n = int(round(out.logits.numel() ** 0.5))
    sel = out.topk[:12].clone()
    l_ids, l_mask = _prompts(inputs, sel.cpu().tolist(), n, synth.LLAMA_TINY["vocab_size"])
    n_new = 16
    embeds, mask, ref_toks, ref_scores = _oracle(head, synth.LLAMA_TINY, out.hidden, sel, l_ids, l_mask, n_new)
    eng = head._llm_engine
    assert eng.w.family == "llama"
    gen = eng.generate(out.hidden, sel, l_ids.cuda(), l_mask.cuda(), max_new_tokens=n_new, return_scores=True,
                       forced_tokens=ref_toks.to(torch.int32).cuda())
    torch.cuda.synchronize()
    dp = (gen.prefix.float().cpu() - embeds).abs()[mask.bool()]           # Llama adds no position table to the prompt
    assert dp.max() <= 4e-2, dp.max()
    decided = _check_scores(gen.scores.cpu(), ref_scores, ref_toks, gen.tokens.cpu().long(), f"llama {name}")
    free = eng.generate(out.hidden, sel, l_ids.cuda(), l_mask.cuda(), max_new_tokens=n_new).tokens.cpu().long()
    for r in range(free.shape[0]):
        for t in range(n_new):
            if not decided[r, t]:
                break
            assert free[r, t] == ref_toks[r, t]


def test_llama_decode_matches_reference_golden(golden, llama_head):
    """The UNMODIFIED reference head with a tiny LlamaForCausalLM (tests/golden/cfg1_llama.pt): its first two generate calls."""
    g = golden("cfg1_llama")
    head = llama_head
    inputs = synth.make_image_inputs(synth.WORKLOADS["cfg1"], 0)
    head(synth.inputs_to(inputs, DEV), is_generation=False)
    out = head.last_output
    ok, bad = margin_set_equal(out.topk.cpu().tolist(), g["exist_logits"], 20, 3e-2)
    assert ok, bad
    sel = torch.tensor(g["selected"][:2], dtype=torch.int32)
    l_ids, l_mask = _prompts(inputs, sel.tolist(), 8, synth.LLAMA_TINY["vocab_size"])
    assert torch.equal(l_mask, g["llm_masks"][:2, 32:].to(torch.int32))
    for r in range(2):
        ref_scores = g["scores_first2"][r][None]
        n_new = ref_scores.shape[1]
        ref_toks = g["sequences"][r][:n_new][None]
        gen = head._llm_engine.generate(out.hidden, sel[r:r + 1].cuda(), l_ids[r:r + 1].cuda(), l_mask[r:r + 1].cuda(),
                                        max_new_tokens=n_new, return_scores=True, forced_tokens=ref_toks.to(torch.int32).cuda())
        got = gen.scores.cpu()
        _check_scores(got, ref_scores, ref_toks, got.argmax(-1), f"llama golden pair {r}", min_match=0.85)


def test_llama_gqa_decode():
    cfg = synth.LLAMA_TINY_GQA
    head = build_product_head(llm=cfg, device=DEV)
    g = torch.Generator().manual_seed(4)
    hidden = torch.randn((6 * 33, 768), generator=g).to(torch.bfloat16)
    sel = torch.tensor([5, 1, 3], dtype=torch.int32)
    tok = synth.SyntheticTokenizer("llm"); tok.set_vocab_size(cfg["vocab_size"])
    enc = tok(["a", "bb", "ccc"])
    l_ids, l_mask = enc["input_ids"].to(torch.int32), enc["attention_mask"].to(torch.int32)
    embeds, mask, ref_toks, ref_scores = _oracle(head, cfg, hidden, sel, l_ids, l_mask, 8)
    eng = head.repack(DEV)._llm_engine
    gen = eng.generate(hidden.cuda(), sel.cuda(), l_ids.cuda(), l_mask.cuda(), max_new_tokens=8, return_scores=True,
                       forced_tokens=ref_toks.to(torch.int32).cuda())
    _check_scores(gen.scores.cpu(), ref_scores, ref_toks, gen.tokens.cpu().long(), "llama gqa")


def test_llama2_7b_width():
    """Llama-2-7B layer geometry (d 4096, 32 heads x 128, SwiGLU 11008, vocab 32000, untied lm_head) at 2 layers."""
    cfg = dict(synth.LLAMA2_7B, num_hidden_layers=2)
    head = build_product_head(llm=cfg, device=DEV)
    g = torch.Generator().manual_seed(12)
    hidden = torch.randn((4 * 33, 768), generator=g).to(torch.bfloat16)
    sel = torch.tensor([2, 0], dtype=torch.int32)
    tok = synth.SyntheticTokenizer("llm"); tok.set_vocab_size(cfg["vocab_size"])
    enc = tok(["first prompt", "second, different prompt"])
    l_ids, l_mask = enc["input_ids"].to(torch.int32), enc["attention_mask"].to(torch.int32)
    embeds, mask, ref_toks, ref_scores = _oracle(head, cfg, hidden, sel, l_ids, l_mask, 6)
    eng = head.repack(DEV)._llm_engine
    gen = eng.generate(hidden.cuda(), sel.cuda(), l_ids.cuda(), l_mask.cuda(), max_new_tokens=6, return_scores=True,
                       forced_tokens=ref_toks.to(torch.int32).cuda())
    _check_scores(gen.scores.cpu(), ref_scores, ref_toks, gen.tokens.cpu().long(), "llama2-7b width", min_match=0.8)


def test_long_context_up_to_256_keys(llama_head):
    """Real tokenizers give prompts longer than the synthetic 17 tokens: 32 + 150 + 40 = 222 keys (the kernel limit is 256)."""
    head = llama_head
    eng = head._llm_engine or head.repack(DEV)._llm_engine
    cfg = synth.LLAMA_TINY
    g = torch.Generator().manual_seed(9)
    hidden = torch.randn((3 * 33, 768), generator=g).to(torch.bfloat16)
    sel = torch.tensor([0, 2], dtype=torch.int32)
    T = 150
    l_ids = torch.randint(4, cfg["vocab_size"], (2, T), generator=g, dtype=torch.int32)
    l_mask = torch.ones((2, T), dtype=torch.int32)
    l_mask[1, :37] = 0
    embeds, mask, ref_toks, ref_scores = _oracle(head, cfg, hidden, sel, l_ids, l_mask, 40)
    gen = eng.generate(hidden.cuda(), sel.cuda(), l_ids.cuda(), l_mask.cuda(), max_new_tokens=40, return_scores=True,
                       forced_tokens=ref_toks.to(torch.int32).cuda())
    _check_scores(gen.scores.cpu(), ref_scores, ref_toks, gen.tokens.cpu().long(), "ctx 222", min_match=0.85)


# ---- the shipped config, as mmdet would build it -------------------------------------------------------------------------

def test_shipped_config_head_runs(monkeypatch):
    """``build_head(dict(type='RelationTransformerHeadV4', qformer_model_name=..., llm_model_name='meta-llama/Llama-2-7b-hf',
    relation_classes=...))`` — the relation_head dict of configs/psg/baseline_v4_ov.py:58-63 — with ``from_pretrained``
    answered offline (random-init Llama of 7B width, 4 layers; synthetic tokenizers), then ``forward`` on the B200."""
    from transformers import AutoModelForCausalLM, AutoTokenizer
    from openpsg_b200.registry import build_head
    seen = {}

    def lm_from_pretrained(name, *a, **k):
        seen["llm"] = name
        return synth.build_causal_lm(dict(synth.LLAMA2_7B, num_hidden_layers=4))

    def tok_from_pretrained(name, *a, subfolder=None, **k):
        tok = synth.SyntheticTokenizer("qformer" if subfolder == "qformer_tokenizer" else "llm")
        if tok.kind == "llm":
            tok.set_vocab_size(synth.LLAMA2_7B["vocab_size"])
        return tok

    monkeypatch.setattr(AutoModelForCausalLM, "from_pretrained", staticmethod(lm_from_pretrained))
    monkeypatch.setattr(AutoTokenizer, "from_pretrained", staticmethod(tok_from_pretrained))
    import kings_sgg.models.relation_heads.relation_transformer_head_v4  # noqa: F401  (custom_imports path, baseline_v4_ov.py:11)
    head = build_head(dict(type='RelationTransformerHeadV4', qformer_model_name='Salesforce/instructblip-vicuna-7b',
                           llm_model_name='meta-llama/Llama-2-7b-hf', relation_classes=relation_categories,
                           llm_truncate_num=2))
    assert seen["llm"] == 'meta-llama/Llama-2-7b-hf'
    assert len(head.language_model.model.layers) == 2                    # v4:101-103
    assert head.language_projection.out_features == 4096
    synth.init_parameters(head, 0)
    head.eval().to(DEV)
    res = head(synth.inputs_to(synth.make_image_inputs(synth.WORKLOADS["cfg1"], 0), DEV))
    assert set(res) == {"rel_pred", "rel_score"} and len(res["rel_pred"]) == len(res["rel_score"])
    assert head._llm_engine.w.family == "llama" and head._llm_engine.w.n_layers == 2
    assert head.last_generation.tokens.shape == (20, 16)
    for sub, obj, rel in res["rel_pred"]:
        assert 0 <= sub < 8 and 0 <= obj < 8 and 0 <= rel < 56
    # a second image through the same head (graph capture happens on the second sighting of a signature)
    res2 = head(synth.inputs_to(synth.make_image_inputs(synth.WORKLOADS["cfg1"], 1), DEV))
    res3 = head(synth.inputs_to(synth.make_image_inputs(synth.WORKLOADS["cfg1"], 0), DEV))
    assert set(res2) == set(res3) == {"rel_pred", "rel_score"}


def test_head_rejects_unsupported_llm_at_construction():
    from transformers import GPT2Config, GPT2LMHeadModel
    from openpsg_b200.head import RelationTransformerHeadV4
    with pytest.raises(NotImplementedError, match="model_type"):
        RelationTransformerHeadV4(llm_feature_size=64, qformer_tokenizer=synth.SyntheticTokenizer("qformer"),
                                  llm_tokenizer=synth.SyntheticTokenizer("llm"),
                                  language_model=GPT2LMHeadModel(GPT2Config(n_layer=1, n_embd=64, n_head=2, vocab_size=128)))
