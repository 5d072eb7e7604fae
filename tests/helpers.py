"""Shared builders for tests: product head and oracle port with identical seeded weights."""
import torch

from openpsg_b200 import synth

WEIGHT_SEED = 0


def build_product_head(llm=None, max_object_num=80, topk_pairs=20, max_new_tokens=16, device=None):
    """llm: None -> no language model (relation queries + filter only); dict -> random-init LLM of that config."""
    return synth.build_synthetic_head(llm, max_object_num, topk_pairs, max_new_tokens, device)


def build_port_head(llm=None, max_object_num=80, topk_pairs=20, max_new_tokens=16):
    from oracle.ref_port import ReferencePortHead
    d_llm = llm["hidden_size"] if llm is not None else 4096
    head = ReferencePortHead(llm, llm_feature_size=d_llm, max_object_num=max_object_num, topk_pairs=topk_pairs,
                             max_new_tokens=max_new_tokens)
    synth.init_parameters(head, WEIGHT_SEED)
    return head.eval()


def margin_set_equal(sel_got, z_ref: torch.Tensor, k: int, tol: float):
    """Margin rule: every pair whose reference logit is further than 2*tol from the k-th
    largest reference logit must be classified identically; pairs inside the band are exempt."""
    z = z_ref.reshape(-1)
    kth = torch.topk(z, k).values[-1].item()
    must_in = set(torch.nonzero(z > kth + 2 * tol).reshape(-1).tolist())
    must_out = set(torch.nonzero(z < kth - 2 * tol).reshape(-1).tolist())
    got = set(int(i) for i in sel_got)
    return must_in.issubset(got) and not (got & must_out), (must_in - got, got & must_out)
