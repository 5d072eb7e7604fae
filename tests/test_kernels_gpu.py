"""GPU parity tests, kernel by kernel, through the C ABI (openpsg_b200.ops -> libopsg_b200.so).

Integer / bit outputs must be bit-exact against the oracle; floating-point outputs are compared with a
torch fp32 evaluation of the same bf16-rounded inputs, tolerance stated per test."""
import math

import numpy as np
import pytest
import torch

from openpsg_b200 import synth
from oracle import restated

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from openpsg_b200 import ops as _ops
    return _ops


def _dev():
    return torch.device("cuda:0")


def _rand_bf16(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen) * scale).to(torch.bfloat16)


# ----------------------------------------------------------------------------------------------
# K6: tcgen05 GEMM.  tolerance: |err| <= 1e-2 * max|ref| + bf16 output rounding (2^-8 relative)
# ----------------------------------------------------------------------------------------------
GEMM_CASES = [
    # M, N, K, bias, residual, act, out_dtype
    (128, 256, 64, False, False, 0, torch.bfloat16),
    (128, 256, 768, True, False, 0, torch.bfloat16),
    (300, 768, 768, True, True, 0, torch.bfloat16),
    (1000, 3072, 768, True, False, 1, torch.bfloat16),
    (517, 768, 3072, True, True, 0, torch.bfloat16),
    (2112, 2304, 768, True, False, 0, torch.bfloat16),
    (100, 2560, 320, True, False, 2, torch.bfloat16),
    (49, 1024, 320, False, False, 0, torch.float32),
    (20000, 768, 768, True, True, 0, torch.bfloat16),
    (130, 72, 136, True, False, 0, torch.float32),
]


@pytest.mark.parametrize("case", GEMM_CASES, ids=lambda c: f"{c[0]}x{c[1]}x{c[2]}")
def test_gemm(ops, case):
    M, N, K, use_bias, use_res, act, odt = case
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = _rand_bf16((M, K), g)
    w = _rand_bf16((N, K), g, 1.0 / math.sqrt(K))
    bias = torch.randn(N, generator=g) if use_bias else None
    res = _rand_bf16((M, N), g) if use_res else None
    ref = a.float() @ w.float().t()
    if bias is not None:
        ref = ref + bias
    if res is not None:
        ref = ref + res.float()
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    out = ops.gemm(a.cuda(), w.cuda(), bias.cuda() if bias is not None else None,
                   residual=res.cuda() if res is not None else None, act=act, out_dtype=odt)
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref).abs().max().item()
    tol = 1.2e-2 * ref.abs().max().item()
    assert err <= tol, f"max err {err} > {tol}"


def test_gemm_gelu_tanh_option(ops, monkeypatch):
    """The default tanh.approx GELU of the CTA-pair GEMM epilogue and the exponential form (OPSG_GELU_TANH=0) both stay within
    the GEMM tolerance of the exact GELU, and within 2.5e-4 |x| + one bf16 ulp of each other."""
    g = torch.Generator().manual_seed(3)
    M, N, K = 1000, 3072, 768
    a, w, bias = _rand_bf16((M, K), g), _rand_bf16((N, K), g, 1.0 / math.sqrt(K)), torch.randn(N, generator=g)
    ref = torch.nn.functional.gelu(a.float() @ w.float().t() + bias)
    fast = ops.gemm(a.cuda(), w.cuda(), bias.cuda(), act=1)
    monkeypatch.setenv("OPSG_GELU_TANH", "0")
    exact = ops.gemm(a.cuda(), w.cuda(), bias.cuda(), act=1)
    torch.cuda.synchronize()
    tol = 1.2e-2 * ref.abs().max().item()
    assert (exact.float().cpu() - ref).abs().max().item() <= tol and (fast.float().cpu() - ref).abs().max().item() <= tol
    assert (fast.float() - exact.float()).abs().max().item() <= 2.5e-4 * 6 + 2 ** -7 * 4      # formula error + one bf16 ulp near |x| ~ 4


def test_gemm_bias_along_m_strided_out(ops):
    g = torch.Generator().manual_seed(5)
    for L in (16, 20, 252, 256):
        a = _rand_bf16((768, 256), g, 0.06)           # W_v
        x = _rand_bf16((L, 256), g)                   # image tokens
        bias = torch.randn(768, generator=g)
        Lp = (L + 7) // 8 * 8
        out = torch.zeros((768, Lp), dtype=torch.bfloat16, device="cuda")
        ops.gemm(a.cuda(), x.cuda(), bias.cuda(), bias_along_m=True, out=out[:, :L])
        torch.cuda.synchronize()
        ref = a.float() @ x.float().t() + bias[:, None]
        assert (out[:, :L].float().cpu() - ref).abs().max() <= 1.2e-2 * ref.abs().max()
        assert out[:, L:].abs().max().item() == 0 if Lp > L else True


def test_gemm_split_k_atomic(ops):
    g = torch.Generator().manual_seed(9)
    M, N, K = 256, 256, 8192
    a = _rand_bf16((M, K), g)
    w = _rand_bf16((N, K), g, 1.0 / math.sqrt(K))
    bias = torch.randn(N, generator=g)
    acc = torch.empty((M, N), dtype=torch.float32, device="cuda")
    ops.init_rows_f32(acc, bias.cuda())
    ops.gemm(a.cuda(), w.cuda(), out=acc, atomic=True, k_splits=37)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    assert (acc.cpu() - ref).abs().max() <= 2e-3 * ref.abs().max()


@pytest.mark.parametrize("M,N,K,res,inplace", [(800, 2560, 10240, True, True), (800, 2560, 2560, True, False), (400, 2560, 2560, True, True),
                                               (200, 2560, 10240, False, False), (777, 2560, 2560, True, True), (300, 768, 3072, True, False),
                                               # shapes that stay on the plain tiled path (wide output / too many tiles / activation)
                                               (800, 7680, 2560, False, False), (4900, 2560, 2560, True, True)])
def test_gemm_medium_m_split_k(ops, M, N, K, res, inplace):
    """ops.gemm_medium_m: N = hidden-size Linears at a few hundred rows (stacked LLM decode steps) through the deterministic
    split-K path (fp32 partial slices + fixed-order reduction with bias and residual, in place over the residual) against
    fp32; two runs give the same bits.  tol as test_gemm: 1.2e-2 of the output scale."""
    g = torch.Generator().manual_seed(M + N + K)
    a = _rand_bf16((M, K), g).cuda()
    w = _rand_bf16((N, K), g, 1.0 / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=g).cuda()
    r = _rand_bf16((M, N), g).cuda() if res else None
    ref = a.float() @ w.float().t() + bias + (r.float() if res else 0)
    runs = []
    for _ in range(2):
        if inplace:
            h = r.clone()
            out = ops.gemm_medium_m(a, w, bias, residual=h, out=h)
            assert out.data_ptr() == h.data_ptr()
        else:
            out = ops.gemm_medium_m(a, w, bias, residual=r)
        runs.append(out.float())
    assert torch.equal(runs[0], runs[1])
    assert (runs[0] - ref).abs().max() <= 1.2e-2 * ref.abs().max()


STREAMK_CASES = [
    # M, N, K, bias, residual, act, out_dtype      (LLM decode shapes: OPT-2.7B qkv / out / fc1 / fc2 / lm_head)
    (100, 7680, 2560, True, False, 0, torch.bfloat16),
    (100, 2560, 2560, True, True, 0, torch.bfloat16),
    (100, 10240, 2560, True, False, 2, torch.bfloat16),
    (100, 2560, 10240, True, True, 0, torch.bfloat16),
    (20, 50272, 2560, False, False, 0, torch.float32),
    (1, 1024, 320, True, False, 1, torch.bfloat16),
    (128, 72, 136, True, True, 0, torch.float32),
    (37, 300, 64, False, False, 0, torch.bfloat16),
    # Llama-2-7B decode shapes (K = 4096 / 11008: 6 and 15 slices, the last one short)
    (100, 12288, 4096, False, False, 0, torch.bfloat16),
    (100, 4096, 11008, False, True, 0, torch.bfloat16),
    (128, 1000, 4096, True, True, 0, torch.float32),
    (3, 2560, 10240, True, True, 0, torch.bfloat16),
]


@pytest.mark.parametrize("case", STREAMK_CASES, ids=lambda c: f"{c[0]}x{c[1]}x{c[2]}")
def test_gemm_small_m_streamk(ops, case):
    """Stream-K weight-streaming GEMM (M <= 128) + fix-up kernel against fp32 matmul; also deterministic."""
    M, N, K, use_bias, use_res, act, odt = case
    g = torch.Generator().manual_seed(M * 11 + N * 5 + K)
    a = _rand_bf16((M, K), g)
    w = _rand_bf16((N, K), g, 1.0 / math.sqrt(K))
    bias = torch.randn(N, generator=g) if use_bias else None
    res = _rand_bf16((M, N), g) if use_res else None
    ref = a.float() @ w.float().t()
    if bias is not None:
        ref = ref + bias
    if res is not None:
        ref = ref + res.float()
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    kw = dict(residual=res.cuda() if res is not None else None, act=act, out_dtype=odt)
    out = ops.gemm_small_m(a.cuda(), w.cuda(), bias.cuda() if bias is not None else None, **kw)
    out2 = ops.gemm_small_m(a.cuda(), w.cuda(), bias.cuda() if bias is not None else None, **kw)
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 1.2e-2 * ref.abs().max().item(), err
    assert torch.equal(out, out2)
    if res is not None and odt == torch.bfloat16:      # in-place residual update as the decoder uses it
        h = res.cuda().clone()
        ops.gemm_small_m(a.cuda(), w.cuda(), bias.cuda() if bias is not None else None, residual=h, act=act, out=h)
        assert torch.equal(h, out)


def test_gemm_rejects_bad_arguments(ops):
    from openpsg_b200._lib import OpsgError
    a = torch.zeros((8, 12), dtype=torch.bfloat16, device="cuda")     # K=12 -> lda not multiple of 8
    w = torch.zeros((8, 12), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(OpsgError):
        ops.gemm(a, w)


# ----------------------------------------------------------------------------------------------
# K2: bit-exact masks
# ----------------------------------------------------------------------------------------------
def _mask_case(inputs):
    meta, info = inputs["img_metas"][0], inputs["object_info"][0]
    ids = [int(i) for i in info["object_id_list"]]
    fh, fw = inputs["mask_features"].shape[-2:]
    ref = restated.object_token_masks(info["pan_results"].numpy(), meta["img_shape"][:2], meta["pad_shape"][:2],
                                      (fh, fw), 16, ids)
    return info["pan_results"], meta, ids, (fh // 16, fw // 16), ref


@pytest.mark.parametrize("name", ["cfg1", "stress", "cfg2"])
def test_pair_mask_bits_bit_exact(ops, name):
    inputs = synth.make_stress_inputs() if name == "stress" else synth.make_image_inputs(synth.WORKLOADS[name], 0)
    pan, meta, ids, tok, ref = _mask_case(inputs)
    bits = ops.pair_mask_bits(pan.to(torch.int32).cuda(), meta["img_shape"][:2], meta["pad_shape"][:2], tok,
                              torch.tensor(ids, dtype=torch.int32).cuda())
    got = bits.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, restated.pack_mask_bits(ref))


def test_pair_mask_bits_ragged_shapes(ops):
    rs = np.random.RandomState(0)
    for (ph, pw, ih, iw, Hp, Wp) in [(480, 640, 800, 1067, 800, 1088), (427, 640, 800, 1199, 800, 1216),
                                     (100, 37, 333, 500, 352, 512), (64, 64, 64, 64, 64, 64)]:
        pan = rs.randint(0, 7, size=(ph, pw)).astype(np.int32)
        ids = [0, 1, 2, 3, 4, 5, 6, 1005]
        fh, fw = Hp // 4, Wp // 4
        ref = restated.object_token_masks(pan, (ih, iw), (Hp, Wp), (fh, fw), 16, ids)
        bits = ops.pair_mask_bits(torch.from_numpy(pan).cuda(), (ih, iw), (Hp, Wp), (fh // 16, fw // 16),
                                  torch.tensor(ids, dtype=torch.int32).cuda())
        assert np.array_equal(bits.cpu().numpy().view(np.uint32), restated.pack_mask_bits(ref))


# ----------------------------------------------------------------------------------------------
# K1 operand / K7 / LayerNorm
# ----------------------------------------------------------------------------------------------
def test_patch_im2col_exact(ops):
    g = torch.Generator().manual_seed(1)
    feat = torch.randn(8, 40, 72, generator=g)       # h=40 -> 2 token rows (floor), w=72 -> 4 token cols
    out = ops.patch_im2col(feat.cuda(), 16).float().cpu()
    x = feat[:, :32, :64].reshape(8, 2, 16, 4, 16).permute(1, 3, 0, 2, 4).reshape(8, 8 * 256)
    assert torch.equal(out, x.to(torch.bfloat16).float())


def test_qformer_embed_ln(ops):
    g = torch.Generator().manual_seed(2)
    d, nq, B, T, V = 768, 33, 5, 16, 1000
    query = torch.randn(nq, d, generator=g)
    word = torch.randn(V, d, generator=g) * 0.02
    pos = torch.randn(512, d, generator=g) * 0.02
    gamma = 1 + 0.1 * torch.randn(d, generator=g)
    beta = 0.02 * torch.randn(d, generator=g)
    ids = torch.randint(0, V, (B, T), generator=g)
    out = ops.qformer_embed_ln(query.cuda(), ids.to(torch.int32).cuda(), word.cuda(), pos.cuda(), gamma.cuda(),
                               beta.cuda(), 1e-12).float().cpu()
    ln = lambda x: torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-12)
    ref_q = ln(query)[None].expand(B, -1, -1).reshape(B * nq, d)
    ref_t = ln(word[ids] + pos[:T]).reshape(B * T, d)
    ref = torch.cat([ref_q, ref_t])
    assert (out - ref).abs().max() < 2.5e-2       # bf16 output rounding of O(4) values


@pytest.mark.parametrize("rows,cols", [(333, 768), (333, 320), (333, 2560), (5000, 2560), (100, 3072), (1, 1024),
                                       # >= 4096 rows of <= 1024 columns: the shared-memory ring kernel (slabs of 8 rows, partial
                                       # last slab, fewer slabs than ring stages per CTA, the cfg2 row counts)
                                       (4096, 768), (4099, 768), (52800, 768), (78400, 768), (25600, 768), (5003, 1024),
                                       (6001, 256), (40000, 328)])
def test_layernorm(ops, rows, cols):
    """Warp-per-row kernel, CTA-per-row kernel (<= 4096 rows of >= 1024 columns: LLM decode) and the bulk-copy ring kernel."""
    g = torch.Generator().manual_seed(3)
    x = _rand_bf16((rows, cols), g, 2.0)
    gamma = 1 + 0.1 * torch.randn(cols, generator=g)
    beta = 0.1 * torch.randn(cols, generator=g)
    out = ops.layernorm(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5).float().cpu()
    ref = torch.nn.functional.layer_norm(x.float(), (cols,), gamma, beta, 1e-5)
    assert (out - ref).abs().max() < 3e-2
    if rows >= 4096 and cols <= 1024:
        # same lane <-> column assignment and reduction order as the warp-per-row kernel: a prefix of the rows through that
        # kernel (fewer than 4096 rows) differs at most by the compiler's FMA contraction of the final scale-and-shift, i.e. by
        # one bf16 rounding step on a handful of elements
        head = ops.layernorm(x[:1000].cuda(), gamma.cuda(), beta.cuda(), 1e-5).float().cpu()
        diff = (head - out[:1000]).abs()
        assert diff.max() <= 3.2e-2 and (diff > 0).float().mean() < 1e-3, (diff.max(), (diff > 0).float().mean())


# ----------------------------------------------------------------------------------------------
# K4: small self-attention (Q-Former split layout).  tol 2e-2 abs on O(1) outputs (bf16 P and output)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("text_queries", [True, False])
def test_self_attn_small(ops, text_queries):
    g = torch.Generator().manual_seed(4)
    B, nq, T, H, hd = 7, 33, 16, 12, 64
    d = H * hd
    R = B * (nq + T)
    qkv = _rand_bf16((R, 3 * d), g)
    tmask = (torch.arange(T)[None, :] < torch.randint(10, T + 1, (B, 1), generator=g)).to(torch.int32)
    out = ops.self_attn_small(qkv.cuda(), tmask.cuda(), B, nq, T, H, hd, text_queries).float().cpu()
    f = qkv.float()
    for p in range(B):
        rows = list(range(p * nq, (p + 1) * nq)) + list(range(B * nq + p * T, B * nq + (p + 1) * T))
        x = f[rows]
        q, k, v = (x[:, i * d:(i + 1) * d].reshape(-1, H, hd).transpose(0, 1) for i in range(3))
        valid = torch.cat([torch.ones(nq, dtype=torch.bool), tmask[p].bool()])
        s = q @ k.transpose(-1, -2) / 8.0
        s = s.masked_fill(~valid[None, None, :], float("-inf"))
        ref = (torch.softmax(s, -1) @ v).transpose(0, 1).reshape(-1, d)
        nrows = len(rows) if text_queries else nq
        got = out[rows[:nrows]]
        assert (got - ref[:nrows]).abs().max() < 2e-2


@pytest.mark.parametrize("hd,heads,q_len,pos0", [(80, 32, 1, 48), (80, 32, 1, 80), (80, 32, 1, 127), (64, 12, 1, 17),
                                                 (128, 8, 1, 63), (80, 32, 49, 0), (64, 12, 5, 3),
                                                 # prefill from position 0, <= 64 tokens: the tcgen05 + TMA kernel
                                                 (128, 8, 49, 0), (128, 32, 64, 0), (64, 12, 33, 0), (80, 32, 64, 0), (80, 5, 2, 0),
                                                 (128, 4, 17, 0), (64, 3, 65, 0)])
def test_llm_attn_static_cache(ops, hd, heads, q_len, pos0):
    """K10a/b: causal attention of q_len new tokens (positions pos0 ..) over a static KV cache with a key-validity mask
    (left padding); q_len == 1 takes the shared-memory decode kernel.  tol 2e-2 abs (bf16 P and output)."""
    g = torch.Generator().manual_seed(hd + q_len + pos0)
    nseq, max_ctx = 7, 128
    d = heads * hd
    qkv = _rand_bf16((nseq * q_len, 3 * d), g)
    kc = _rand_bf16((nseq, max_ctx, d), g)
    vc = _rand_bf16((nseq, max_ctx, d), g)
    pad = torch.randint(0, 6, (nseq,), generator=g)
    kmask = (torch.arange(max_ctx)[None, :] >= pad[:, None]).to(torch.uint8)
    kmask[:, pos0:] = 1
    out = torch.zeros((nseq * q_len, d), dtype=torch.bfloat16, device="cuda")
    ops.llm_attn(qkv.cuda(), kc.cuda(), vc.cuda(), kmask.cuda(), nseq, q_len, pos0, heads, hd, hd ** -0.5, out)
    q = qkv[:, :d].float().reshape(nseq, q_len, heads, hd).permute(0, 2, 1, 3)
    ctx = pos0 + q_len
    k = kc[:, :ctx].float().reshape(nseq, ctx, heads, hd).permute(0, 2, 1, 3)
    v = vc[:, :ctx].float().reshape(nseq, ctx, heads, hd).permute(0, 2, 1, 3)
    sc = q @ k.transpose(-1, -2) * hd ** -0.5
    causal = torch.arange(ctx)[None, :] <= (pos0 + torch.arange(q_len))[:, None]
    ok = causal[None, None] & kmask[:, None, None, :ctx].bool()
    ref = (torch.softmax(sc.masked_fill(~ok, float("-inf")), -1) @ v).permute(0, 2, 1, 3).reshape(nseq * q_len, d)
    assert (out.float().cpu() - ref).abs().max() < 2e-2


@pytest.mark.parametrize("hd,heads,pos0,nseq", [(80, 32, 128, 9), (80, 32, 200, 40), (128, 8, 255, 13), (64, 12, 143, 5),
                                                (80, 32, 64, 300), (128, 32, 15, 64), (64, 3, 0, 2)])
def test_llm_attn_decode_long_context_many_sequences(ops, hd, heads, pos0, nseq):
    """K10b decode kernel (mma.sync arithmetic over cp.async-staged head slices): contexts up to 256 keys (the 16-tile
    instantiation), many sequences per launch (stacked images), random left padding.  tol 2e-2 abs (bf16 P and output)."""
    g = torch.Generator().manual_seed(hd + pos0 + nseq)
    max_ctx = 256
    d = heads * hd
    qkv = _rand_bf16((nseq, 3 * d), g)
    kc = _rand_bf16((nseq, max_ctx, d), g)
    vc = _rand_bf16((nseq, max_ctx, d), g)
    pad = torch.randint(0, min(pos0, 20) + 1, (nseq,), generator=g)
    kmask = (torch.arange(max_ctx)[None, :] >= pad[:, None]).to(torch.uint8)
    kmask[:, pos0:] = 1
    out = torch.zeros((nseq, d), dtype=torch.bfloat16, device="cuda")
    kc_d, vc_d = kc.cuda(), vc.cuda()
    ops.llm_attn_append(qkv.cuda(), kc_d, vc_d, kmask.cuda(), nseq, pos0, heads, hd, hd ** -0.5, out)
    ctx = pos0 + 1
    kc[:, pos0] = qkv[:, d:2 * d]
    vc[:, pos0] = qkv[:, 2 * d:]
    assert torch.equal(kc_d.cpu()[:, :ctx], kc[:, :ctx]) and torch.equal(vc_d.cpu()[:, :ctx], vc[:, :ctx])
    q = qkv[:, :d].float().reshape(nseq, 1, heads, hd).permute(0, 2, 1, 3)
    k = kc[:, :ctx].float().reshape(nseq, ctx, heads, hd).permute(0, 2, 1, 3)
    v = vc[:, :ctx].float().reshape(nseq, ctx, heads, hd).permute(0, 2, 1, 3)
    sc = q @ k.transpose(-1, -2) * hd ** -0.5
    ok = kmask[:, None, None, :ctx].bool()
    ref = (torch.softmax(sc.masked_fill(~ok, float("-inf")), -1) @ v).permute(0, 2, 1, 3).reshape(nseq, d)
    assert (out.float().cpu() - ref).abs().max() < 2e-2


@pytest.mark.parametrize("hd,heads,pos0", [(80, 32, 60), (64, 12, 0), (128, 8, 127)])
def test_llm_attn_append_equals_append_then_attend(ops, hd, heads, pos0):
    """The fused decode entry (new k / v taken from the qkv row and written to the caches by the attention kernel) gives
    the same output and the same caches as opsg_kv_append followed by opsg_llm_attn."""
    g = torch.Generator().manual_seed(hd + pos0)
    nseq, max_ctx = 5, 128
    d = heads * hd
    qkv = _rand_bf16((nseq, 3 * d), g).cuda()
    kc = _rand_bf16((nseq, max_ctx, d), g).cuda()
    vc = _rand_bf16((nseq, max_ctx, d), g).cuda()
    kmask = torch.ones((nseq, max_ctx), dtype=torch.uint8).cuda()
    kmask[1, :min(3, pos0)] = 0
    kc2, vc2 = kc.clone(), vc.clone()
    ref = torch.zeros((nseq, d), dtype=torch.bfloat16, device="cuda")
    ops.kv_append(qkv, nseq, 1, pos0, d, kc, vc)
    ops.llm_attn(qkv, kc, vc, kmask, nseq, 1, pos0, heads, hd, hd ** -0.5, ref)
    got = torch.zeros_like(ref)
    ops.llm_attn_append(qkv, kc2, vc2, kmask, nseq, pos0, heads, hd, hd ** -0.5, got)
    assert torch.equal(got, ref) and torch.equal(kc2, kc) and torch.equal(vc2, vc)


def test_self_attn_small_shared_query_rows(ops):
    """Layer-0 form: q/k/v of the query rows come from one [n_query, 3d] table shared by all pairs; bit-identical to the
    full-layout call on a qkv whose query rows are that table repeated (the query rows of qkv itself are never read)."""
    g = torch.Generator().manual_seed(41)
    B, nq, T, H, hd = 9, 33, 16, 12, 64
    d = H * hd
    shared = _rand_bf16((nq, 3 * d), g)
    text = _rand_bf16((B * T, 3 * d), g)
    full = torch.cat([shared.repeat(B, 1), text])
    tmask = (torch.arange(T)[None, :] < torch.randint(8, T + 1, (B, 1), generator=g)).to(torch.int32)
    ref = ops.self_attn_small(full.cuda(), tmask.cuda(), B, nq, T, H, hd, True)
    poisoned = torch.cat([torch.full((B * nq, 3 * d), float("nan"), dtype=torch.bfloat16), text])
    got = ops.self_attn_small(poisoned.cuda(), tmask.cuda(), B, nq, T, H, hd, True, shared_query_qkv=shared.cuda())
    assert torch.equal(got, ref)


# ----------------------------------------------------------------------------------------------
# K5: pair x image masked cross-attention (tcgen05).  tol 2e-2 abs (bf16 P, bf16 output)
# ----------------------------------------------------------------------------------------------
def _xattn_ref(q, k, v, masks, N, pair_index, nq):
    B = q.shape[0] // nq
    H, hd = 12, 64
    qh = q.float().reshape(B, nq, H, hd).permute(0, 2, 1, 3)
    kh = k.float().reshape(-1, H, hd).permute(1, 0, 2)
    vh = v.float().reshape(-1, H, hd).permute(1, 0, 2)
    pi = pair_index if pair_index is not None else torch.arange(B)
    M = masks[pi // N] | masks[pi % N]
    bias = (1.0 - M.float()) * torch.finfo(torch.float32).min
    s = qh @ kh.transpose(-1, -2) / 8.0 + bias[:, None, None, :]
    o = torch.softmax(s, -1) @ vh
    return o.permute(0, 2, 1, 3).reshape(B * nq, H * hd)


@pytest.mark.parametrize("L,N,B", [(256, 40, 1600), (16, 8, 64), (20, 7, 49), (252, 5, 25), (256, 3, 9)])
def test_xattn_pairs(ops, L, N, B):
    g = torch.Generator().manual_seed(L * 100 + N)
    nq, d = 33, 768
    q = _rand_bf16((B * nq, d), g)
    k = _rand_bf16((L, d), g)
    v = _rand_bf16((L, d), g)
    masks = torch.rand(N, L, generator=g) < 0.15
    masks[N - 1] = False                                  # empty object -> pair (N-1, N-1) is all-masked
    bits = torch.from_numpy(restated.pack_mask_bits(masks.numpy()).view(np.int32))
    Lp = (L + 7) // 8 * 8
    vt = torch.zeros((d, Lp), dtype=torch.bfloat16)
    vt[:, :L] = v.t()
    out = ops.xattn_pairs(q.cuda(), k.cuda(), vt.cuda(), bits.cuda(), N, B, nq, L, 12, 64).float().cpu()
    ref = _xattn_ref(q, k, v, masks, N, None, nq)
    err = (out - ref).abs().max().item()
    assert err < 2e-2, err
    # all-masked pair attends uniformly: its context is the mean of V
    last = out[(B - 1) * nq:(B - 1) * nq + 1]
    assert (last - v.float().mean(0, keepdim=True)).abs().max() < 2e-2


@pytest.mark.parametrize("L,N,subset", [(257, 5, False), (320, 6, True), (512, 4, False), (1000, 3, True)])
def test_xattn_pairs_more_than_256_keys(ops, L, N, subset):
    """More image tokens than one tensor-memory score tile: the online-softmax kernel (csrc/xattn_pairs_long.cu) -- same
    semantics (masked keys weigh 0, an all-masked pair gets the mean of V), pair_index subsets, ragged last key block."""
    g = torch.Generator().manual_seed(L + N)
    nq, d = 33, 768
    pairs = torch.randperm(N * N, generator=g)[: N * N - 3].to(torch.int32) if subset else None
    B = pairs.numel() if subset else N * N
    q, k, v = _rand_bf16((B * nq, d), g), _rand_bf16((L, d), g), _rand_bf16((L, d), g)
    masks = torch.rand(N, L, generator=g) < 0.1
    masks[N - 1] = False                                  # pair (N-1, N-1) is all-masked
    masks[0, : L // 2] = False                            # object 0 only sees the second half of the keys
    bits = torch.from_numpy(restated.pack_mask_bits(masks.numpy()).view(np.int32))
    Lp = (L + 7) // 8 * 8
    vt = torch.zeros((d, Lp), dtype=torch.bfloat16)
    vt[:, :L] = v.t()
    out = ops.xattn_pairs(q.cuda(), k.cuda(), vt.cuda(), bits.cuda(), N, B, nq, L, 12, 64,
                          pair_index=pairs.cuda() if subset else None, bias_tiles=False).float().cpu()
    ref = _xattn_ref(q, k, v, masks, N, pairs.long() if subset else None, nq)
    err = (out - ref).abs().max().item()
    assert err < 2e-2, err
    if not subset:
        last = out[(B - 1) * nq:(B - 1) * nq + 1]
        assert (last - v.float().mean(0, keepdim=True)).abs().max() < 2e-2


def test_xattn_pairs_without_bias_tiles(ops):
    """bias_tiles == NULL selects the self-contained kernel (per-row OR of the bit rows): same results."""
    g = torch.Generator().manual_seed(3)
    L, N, nq, d = 252, 6, 33, 768
    B = N * N
    q, k, v = _rand_bf16((B * nq, d), g), _rand_bf16((L, d), g), _rand_bf16((L, d), g)
    masks = torch.rand(N, L, generator=g) < 0.2
    masks[0] = False
    bits = torch.from_numpy(restated.pack_mask_bits(masks.numpy()).view(np.int32))
    vt = torch.zeros((d, 256), dtype=torch.bfloat16)
    vt[:, :L] = v.t()
    a = ops.xattn_pairs(q.cuda(), k.cuda(), vt.cuda(), bits.cuda(), N, B, nq, L, 12, 64, bias_tiles=False).float().cpu()
    b = ops.xattn_pairs(q.cuda(), k.cuda(), vt.cuda(), bits.cuda(), N, B, nq, L, 12, 64).float().cpu()
    ref = _xattn_ref(q, k, v, masks, N, None, nq)
    assert (a - ref).abs().max() < 2e-2 and (b - ref).abs().max() < 2e-2


def test_xattn_bias_tiles_bit_exact(ops):
    """The mask-bias operand tiles are integer work: compare every byte with a numpy restatement."""
    g = torch.Generator().manual_seed(4)
    L, N, nq = 200, 9, 33
    B = N * N
    masks = torch.rand(N, L, generator=g) < 0.1
    masks[3] = False
    bits = torch.from_numpy(restated.pack_mask_bits(masks.numpy()).view(np.int32))
    got = ops.xattn_bias_tiles(bits.cuda(), N, B, nq, L).cpu().numpy()
    rows = B * nq
    m_tiles = (rows + 127) // 128
    tiles, flags = got[:m_tiles * 12288].reshape(m_tiles, 12288), got[m_tiles * 12288:]
    vis = got[m_tiles * (12288 + 128):].view(np.uint16)      # 4 x uint16 per tile, 16 bytes reserved per tile
    pm = restated.pair_masks(masks.numpy())                       # [B, L]
    NEG, ONE = 0xC680, 0x3F80
    for mt in range(m_tiles):
        first = (mt * 128) // nq
        last = min(B - 1, (mt * 128 + 127) // nq)
        a = tiles[mt, :4096].view(np.uint16).reshape(16, 2, 8, 8)     # [row group][k-core][row][slot]
        b = tiles[mt, 4096:].view(np.uint16).reshape(32, 2, 8, 8)
        assert not a[:, 1].any() and not b[:, 1].any()
        for r in range(128):
            row = mt * 128 + r
            exp = np.zeros(8, np.uint16)
            if row < rows:
                exp[row // nq - first] = ONE
                assert flags[row] == (0 if pm[row // nq].any() else 1)
            assert np.array_equal(a[r // 8, 0, r % 8], exp)
        for quarter in range(4):       # chunk visibility of each 32-row quarter
            exp_vis = 0
            for r in range(quarter * 32, quarter * 32 + 32):
                row = mt * 128 + r
                if row < rows:
                    m = np.zeros(256, bool)
                    m[:L] = pm[row // nq]
                    exp_vis |= sum(1 << w for w in range(16) if m[16 * w:16 * w + 16].any())
            assert vis[mt * 4 + quarter] == exp_vis, (mt, quarter)     # chunk_vis is indexed [mt * 4 + quarter]
        for key in range(256):
            exp = np.zeros(8, np.uint16)
            for s in range(last - first + 1):
                if not (key < L and pm[first + s, key]):
                    exp[s] = NEG
            assert np.array_equal(b[key // 8, 0, key % 8], exp), (mt, key)


def test_xattn_pairs_sparse_masks_skip_chunks(ops):
    """Compact masks (few visible key chunks per row group, as panoptic segments give) exercise the chunk-skipping
    path; an object with no tokens and rows with all keys in one half are included."""
    g = torch.Generator().manual_seed(21)
    L, N, nq, d = 256, 12, 33, 768
    B = N * N
    q, k, v = _rand_bf16((B * nq, d), g), _rand_bf16((L, d), g), _rand_bf16((L, d), g)
    masks = torch.zeros(N, L, dtype=torch.bool)
    for o in range(N - 1):                       # object o owns 5 consecutive tokens; the last object owns none
        start = int(torch.randint(0, L - 5, (1,), generator=g))
        masks[o, start:start + 5] = True
    bits = torch.from_numpy(restated.pack_mask_bits(masks.numpy()).view(np.int32))
    out = ops.xattn_pairs(q.cuda(), k.cuda(), v.t().contiguous().cuda(), bits.cuda(), N, B, nq, L, 12, 64).float().cpu()
    ref = _xattn_ref(q, k, v, masks, N, None, nq)
    assert (out - ref).abs().max() < 2e-2
    assert torch.isfinite(out).all()


def test_xattn_pairs_with_pair_index(ops):
    g = torch.Generator().manual_seed(77)
    L, N, nq, d = 256, 12, 33, 768
    idx = torch.tensor([5, 143, 0, 77, 12, 12, 100], dtype=torch.int32)
    B = idx.numel()
    q = _rand_bf16((B * nq, d), g)
    k = _rand_bf16((L, d), g)
    v = _rand_bf16((L, d), g)
    masks = torch.rand(N, L, generator=g) < 0.1
    bits = torch.from_numpy(restated.pack_mask_bits(masks.numpy()).view(np.int32))
    out = ops.xattn_pairs(q.cuda(), k.cuda(), v.t().contiguous().cuda(), bits.cuda(), N, B, nq, L, 12, 64,
                          pair_index=idx.cuda()).float().cpu()
    ref = _xattn_ref(q, k, v, masks, N, idx.long(), nq)
    assert (out - ref).abs().max() < 2e-2


def test_xattn_operand_tiles_reject_more_than_256_keys(ops):
    """The tensor-memory kernels' helpers (key order, mask operand tiles) cover <= 256 image tokens and say so; opsg_xattn_pairs
    itself takes longer inputs through the online-softmax kernel (test_xattn_pairs_more_than_256_keys)."""
    from openpsg_b200._lib import OPSG_E_UNSUPPORTED, OpsgError
    bits = torch.zeros((1, 10), dtype=torch.int32, device="cuda")
    with pytest.raises(OpsgError) as e:
        ops.xattn_bias_tiles(bits, 1, 1, 33, 300, None)
    assert e.value.code == OPSG_E_UNSUPPORTED
    with pytest.raises(OpsgError):
        ops.token_order(bits, 300)


# ----------------------------------------------------------------------------------------------
# K8: existence filter: logits within 1e-4 of fp32; index set / mask bit-exact given the logits
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,k", [(64, 20), (1600, 20), (6400, 100), (49, 49)])
def test_exist_filter_topk(ops, B, k):
    g = torch.Generator().manual_seed(B)
    d, nq = 768, 33
    x = _rand_bf16((B * nq, d), g)
    w = torch.randn(d, generator=g) * 0.04
    b = torch.randn(1, generator=g) * 0.02
    logits, probs, mask, top = ops.exist_filter_topk(x.cuda(), nq * d, B, d, w.cuda(), b.cuda(), 0.5, k)
    torch.cuda.synchronize()
    ref = restated.existence_logits(x.float()[::nq], w, b)
    z = logits.cpu()
    assert (z - ref).abs().max() < 1e-4
    assert top.cpu().tolist() == restated.topk_pairs(z, k)            # bit-exact on identical filter inputs
    assert np.array_equal(mask.cpu().numpy().astype(bool), restated.existence_mask(z))
    assert (probs.cpu() - torch.sigmoid(z)).abs().max() < 1e-6


def test_topk_ties_go_to_lower_index(ops):
    d, nq, B = 768, 33, 300
    x = torch.zeros((B * nq, d), dtype=torch.bfloat16)
    x[::nq, 0] = torch.tensor([float(i % 3) for i in range(B)], dtype=torch.bfloat16)
    w = torch.zeros(d)
    w[0] = 1.0
    _, _, _, top = ops.exist_filter_topk(x.cuda(), nq * d, B, d, w.cuda(), torch.zeros(1).cuda(), 0.5, 10)
    assert top.cpu().tolist() == [2 + 3 * i for i in range(10)]


def test_topk_nan_logits_keep_a_total_order(ops):
    """NaN logits (only possible with non-finite weights) rank as -inf: every slot of the top-k list is written with a
    distinct in-range index, finite candidates first."""
    d, nq, B, k = 768, 33, 64, 64
    x = torch.zeros((B * nq, d), dtype=torch.bfloat16)
    x[::nq, 0] = torch.arange(B, dtype=torch.float32).to(torch.bfloat16)
    x[5 * nq, 0] = float("nan")
    x[9 * nq, 0] = float("nan")
    w = torch.zeros(d)
    w[0] = 1.0
    _, _, _, top = ops.exist_filter_topk(x.cuda(), nq * d, B, d, w.cuda(), torch.zeros(1).cuda(), 0.5, k)
    got = top.cpu().tolist()
    assert sorted(got) == list(range(B))
    assert got[-2:] == [5, 9] and got[0] == B - 1


# ----------------------------------------------------------------------------------------------
# K11: mask mean-pool + pair gather (fp32 sums, different order: tol 1e-4 relative)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(256, 64, 64, 8), (256, 256, 256, 40), (16, 33, 47, 5)])
def test_mask_pool_pairs(ops, shape):
    C, h, w, N = shape
    g = torch.Generator().manual_seed(C + h)
    feat = torch.randn(C, h, w, generator=g)
    wl = synth.Workload("t", h * 4, w * 4, N)
    label = torch.randint(-1, N, (h, w), generator=g).to(torch.int32)
    label[label == N - 1] = -1                            # last object owns nothing -> zero embedding
    if h >= 64:                                           # coherent regions like a panoptic map
        pan = synth.make_panoptic_map(h, w, N, g, ensure_token_coverage=False)
        label = (pan % 1000).to(torch.int32)
        label[label == N - 1] = -1
    obj, pair = ops.mask_pool_pairs(feat.cuda(), label.cuda(), N)
    masks = torch.stack([label == i for i in range(N)])
    ref_obj, ref_pair = restated.mask_pool_pairs(feat, masks)
    assert (obj.cpu() - ref_obj).abs().max() < 1e-4
    assert (pair.cpu() - ref_pair).abs().max() < 1e-4
    assert obj[N - 1].abs().max().item() == 0


def test_argmax_and_gathers(ops):
    g = torch.Generator().manual_seed(8)
    x = torch.randn(37, 50272, generator=g)
    x[3, 100] = x[3, 40000] = 99.0
    got = ops.argmax_rows(x.cuda()).cpu()
    assert torch.equal(got.long(), x.argmax(1)) and got[3] == 100
    src = _rand_bf16((50, 32 * 768), g)
    idx = torch.tensor([4, 4, 49, 0], dtype=torch.int32)
    assert torch.equal(ops.gather_rows(src.cuda(), 32 * 768, idx.cuda()).cpu(), src[idx.long()])


def test_token_order_sorts_tokens_by_owning_object(ops):
    """opsg_token_order: perm lists the image tokens by (owner, token index), owner = first object whose mask holds the
    token, unowned tokens last; bits_sorted are the same masks in that order."""
    g = torch.Generator().manual_seed(21)
    for N, L in ((7, 20), (40, 256), (80, 256), (5, 255)):
        words = 8
        owner = torch.randint(0, N + 2, (L,), generator=g)           # N, N + 1 = nobody
        m = torch.zeros((N, L), dtype=torch.bool)
        for o in range(N):
            m[o] = owner == o
        m[0] |= torch.rand(L, generator=g) > 0.9                     # overlapping masks: object 0 also claims other tokens
        bits = torch.zeros((N, words), dtype=torch.int64)
        for l in range(L):
            bits[:, l // 32] |= m[:, l].long() << (l % 32)
        bits32 = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
        perm, bs = ops.token_order(bits32.cuda(), L)
        first = torch.where(m.any(0), m.float().argmax(0), torch.full((L,), N))
        want = sorted(range(L), key=lambda l: (int(first[l]), l))
        assert perm.cpu().tolist() == want
        got = bs.cpu().numpy().view(np.uint32)
        obj = ((got[:, np.arange(L) // 32] >> (np.arange(L) % 32).astype(np.uint32)) & 1).astype(bool)
        assert np.array_equal(obj, m[:, want].numpy())
        assert not ((got[:, np.arange(L, words * 32) // 32] >> (np.arange(L, words * 32) % 32).astype(np.uint32)) & 1).any()


MASK_POOL_MODES = (("plain", "none", False), ("add", "add", False), ("cat", "cat", False), ("bg", "none", True), ("add+bg", "add", True))


def test_mask_pool_chain_matches_reference_golden(ops, golden):
    """Row a11 end to end on the GPU — panoptic map -> label map (the reference's nearest / pad / nearest mask chain) ->
    pooled object embeddings (+ class embedding, + background feature) -> pair embeddings — against what the reference's own
    statements produced (tests/golden/mask_pool.pt) and against the restatement; two runs are bit-identical."""
    g = golden("mask_pool")
    feat, pan, ids, meta, table = synth.make_mask_pool_case()
    f = feat[0].cuda()
    ids_t = torch.tensor(ids, dtype=torch.int32).cuda()
    label, rep = ops.mask_pool_labels(pan.to(torch.int32).cuda(), meta["img_shape"][:2], meta["pad_shape"][:2], f.shape[-2:], ids_t)
    m = restated.object_masks_feature_res(pan.numpy(), meta["img_shape"][:2], meta["pad_shape"][:2], f.shape[-2:], ids)
    lab = label.cpu().numpy()
    first = np.where(m.any(0), m.argmax(0), len(ids))
    assert np.array_equal(lab, first), "label map = first object whose mask holds the pixel (bit-exact)"
    assert rep.cpu().tolist() == [0, 1, 2, 3, 4, 5, 6, 2, 8]
    cls_ids = torch.tensor([i % 1000 for i in ids], dtype=torch.int32).cuda()
    for name, cls_mode, bg in MASK_POOL_MODES:
        obj, pair = ops.mask_pool_pairs(f, label, len(ids), rep=rep, cls_table=table.cuda(), cls_ids=cls_ids, cls_mode=cls_mode,
                                        use_background=bg)
        obj2, pair2 = ops.mask_pool_pairs(f, label, len(ids), rep=rep, cls_table=table.cuda(), cls_ids=cls_ids, cls_mode=cls_mode,
                                          use_background=bg)
        assert torch.equal(obj, obj2) and torch.equal(pair, pair2), "no atomics: bit-reproducible"
        d = (obj.cpu() - g[name]).abs().max().item()
        assert d < 1e-4, (name, d)
        ref_obj, ref_pair = restated.mask_pool_chain(feat[0], pan.numpy(), meta["img_shape"][:2], meta["pad_shape"][:2], ids, table,
                                                     cls_mode, bg)
        assert (pair.cpu() - ref_pair).abs().max() < 1e-4


@pytest.mark.parametrize("name", ["cfg2", "cfg5"])
def test_mask_pool_benchmark_shapes(ops, name):
    """The 1024x1024 images (67 MB feature map, 40 / 80 objects) against the restatement."""
    wl = synth.WORKLOADS[name]
    inp = synth.make_image_inputs(wl, 0)
    ids = [int(i) for i in inp["object_info"][0]["object_id_list"]]
    pan = inp["object_info"][0]["pan_results"]
    feat = inp["mask_features"][0]
    label, rep = ops.mask_pool_labels(pan.to(torch.int32).cuda(), (wl.height, wl.width), (wl.height, wl.width), feat.shape[-2:],
                                      torch.tensor(ids, dtype=torch.int32).cuda())
    obj, pair = ops.mask_pool_pairs(feat.cuda(), label, len(ids), rep=rep, use_background=True)
    ref_obj, ref_pair = restated.mask_pool_chain(feat, pan.numpy(), (wl.height, wl.width), (wl.height, wl.width), ids, None, "none", True)
    assert (obj.cpu() - ref_obj).abs().max() < 1e-4
    assert (pair.cpu() - ref_pair).abs().max() < 1e-4


@pytest.mark.parametrize("hd,heads,q_len", [(128, 32, 49), (80, 32, 49), (64, 12, 64), (128, 8, 20), (80, 7, 33)])
def test_llm_prefill_attn_left_padding(ops, hd, heads, q_len):
    """K10a on the tcgen05 + TMA kernel with LEFT-PADDED prompts (v4:262, pads sit between the 32 relation rows and the text in
    the real prompt; here at the front): valid query rows against fp32; rows that sit on padding come out finite (zeros)."""
    g = torch.Generator().manual_seed(hd * 3 + q_len)
    nseq, max_ctx = 9, q_len + 32
    d = heads * hd
    qkv = _rand_bf16((nseq * q_len, 3 * d), g)
    kc = _rand_bf16((nseq, max_ctx, d), g)
    vc = _rand_bf16((nseq, max_ctx, d), g)
    pad = torch.randint(0, q_len // 2, (nseq,), generator=g)
    pad[0] = 0
    kmask = (torch.arange(max_ctx)[None, :] >= pad[:, None]).to(torch.uint8)
    kmask[3, pad[3] + 2] = 0                                            # a hole in the middle (mid-sequence pads, v4:298-299)
    out = torch.full((nseq * q_len, d), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.llm_attn(qkv.cuda(), kc.cuda(), vc.cuda(), kmask.cuda(), nseq, q_len, 0, heads, hd, hd ** -0.5, out)
    q = qkv[:, :d].float().reshape(nseq, q_len, heads, hd).permute(0, 2, 1, 3)
    k = kc[:, :q_len].float().reshape(nseq, q_len, heads, hd).permute(0, 2, 1, 3)
    v = vc[:, :q_len].float().reshape(nseq, q_len, heads, hd).permute(0, 2, 1, 3)
    sc = q @ k.transpose(-1, -2) * hd ** -0.5
    causal = torch.arange(q_len)[None, :] <= torch.arange(q_len)[:, None]
    ok = causal[None, None] & kmask[:, None, None, :q_len].bool()
    any_key = ok.any(-1)                                                # [nseq, 1, q_len]
    ref = torch.softmax(sc.masked_fill(~ok, float("-inf")), -1)
    ref = torch.nan_to_num(ref, nan=0.0) @ v
    ref = ref.permute(0, 2, 1, 3).reshape(nseq * q_len, d)
    got = out.float().cpu()
    assert torch.isfinite(got).all(), "every row is written and finite"
    rows = any_key[:, 0].reshape(-1)
    assert (got[rows] - ref[rows]).abs().max() < 2e-2
