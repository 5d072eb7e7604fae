"""Training branch of ``RelationTransformerHeadV4`` as plain PyTorch (row f3 of SURVEY.md §8: API completeness, no kernels).

What the reference computes when ``self.training`` (``relation_transformer_head_v4.py:114-133,187-204,267-285,327-341``,
``prepare_train`` ``:360-406``, ``qformer_sampler`` ``:437-461``, ``loss_for_rel_cls_pred`` ``:463-482``,
``multilabel_categorical_crossentropy`` ``:484-495``), evaluated on the head's own HuggingFace modules so that autograd
reaches every trainable parameter under its reference name.  It is device-agnostic (the reference hard-codes ``.cuda()``),
which is what lets the CPU suite pin it against a golden made by running the unmodified reference file.

Random draws follow the reference's stream exactly — ``torch.randint`` for the pair sampler, then the Q-Former forward
(dropout), then Python's ``random.sample`` for the pairs sent to the LLM, then one batch-1 LM forward per pair — so a
seeded run reproduces the reference's losses bit for bit, dropout included.
"""
from __future__ import annotations

import random
from typing import Dict, List

import torch
import torch.nn.functional as F

from .categories import object_categories

IGNORE = -100


def relation_targets(gt_rels, n: int, num_rel: int, like: torch.Tensor) -> torch.Tensor:
    """[n, n, num_rel] multi-hot relation labels from (subject, object, relation) triples (v4:123-126)."""
    t = like.new_zeros((n, n, num_rel))
    trip = torch.as_tensor([list(map(int, r)) for r in gt_rels], dtype=torch.long).reshape(-1, 3)
    if trip.numel():
        t[trip[:, 0], trip[:, 1], trip[:, 2]] = 1
    return t


def token_masks_train(head, feat: torch.Tensor, meta: dict, gt_thing_masks, gt_semantic_seg: torch.Tensor) -> torch.Tensor:
    """Object masks at token resolution from the ground truth (v4:371-398): thing masks bilinear-resized and
    thresholded at 0.5 (taken in annotation order), stuff masks = nearest-resized semantic map == category.
    -> bool [n, th*tw]."""
    th, tw = feat.shape[-2] // head.patch_size, feat.shape[-1] // head.patch_size
    things = gt_thing_masks.to_tensor(feat.dtype, feat.device)
    things = F.interpolate(things[None].float(), size=(th, tw), mode='bilinear', align_corners=False)[0] > 0.5
    sem = F.interpolate(gt_semantic_seg.to(feat.dtype).to(feat.device)[None].float(), size=(th, tw), mode='nearest')[0]
    rows, next_thing = [], 0
    for info in meta['masks_info']:
        if info['is_thing']:
            rows.append(things[next_thing:next_thing + 1])
            next_thing += 1
        else:
            rows.append(sem == info['category'])
    return torch.cat(rows, dim=0).flatten(1)


def sample_pairs(head, target: torch.Tensor) -> torch.Tensor:
    """Pair indices the Q-Former is trained on (v4:437-461): every positive pair plus up to ``neg_over_pos`` negatives per
    positive while positives are fewer than ``sampled_qformer_batch_size``; otherwise a 1 : neg_over_pos draw with
    replacement.  Consumes the global torch generator exactly like the reference (one ``randint`` per group drawn)."""
    per_pair = target.reshape(-1, head.num_relation_classes).sum(1)
    pos = torch.nonzero(per_pair, as_tuple=False)[:, 0]
    neg = torch.nonzero(per_pair == 0, as_tuple=False)[:, 0]
    budget, ratio = head.sampled_qformer_batch_size, head.qformer_neg_over_pos
    if pos.numel() < budget:
        take_neg = min(budget - pos.numel(), pos.numel() * ratio)
        picked_pos = pos
    else:
        picked_pos = pos[torch.randint(0, pos.numel(), (budget // (ratio + 1),))]
        take_neg = budget * ratio // (ratio + 1)
    picked_neg = neg[torch.randint(0, neg.numel(), (take_neg,))]
    return torch.cat([picked_pos, picked_neg], dim=0)


def multilabel_categorical_crossentropy(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
    """Su Jianlin's multi-label softmax loss (https://kexue.fm/archives/7359) as used at v4:484-495."""
    signed = (1 - 2 * y_true) * y_pred
    zero = torch.zeros_like(signed[..., :1])
    neg = torch.logsumexp(torch.cat([signed - y_true * 9999, zero], dim=-1), dim=-1)
    pos = torch.logsumexp(torch.cat([signed - (1 - y_true) * 9999, zero], dim=-1), dim=-1)
    return neg + pos


def rel_cls_loss(head, pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """v4:463-482: BCE-with-logits for the binary existence head ([pairs]); for the multiclass head ([pairs, classes]) the
    multi-label softmax loss re-weighted by loss / max(loss); both scaled by ``rel_cls_loss_weight``."""
    if pred.dim() == 1:
        loss = F.binary_cross_entropy_with_logits(pred, label)
    else:
        loss = multilabel_categorical_crossentropy(label, pred)
        loss = (loss * (loss / loss.max())).mean()
    return loss * head.rel_cls_loss_weight


def forward_train(head, inputs: dict, is_generation=None) -> Dict[str, torch.Tensor]:
    """-> {'binary_rel_cls_loss', ['multiclass_rel_cls_loss'], 'rel_llm_loss'} (v4:346-351)."""
    feat = inputs['mask_features']
    meta = inputs['img_metas'][0]
    dev = feat.device
    infos = meta['masks_info']
    n = len(infos)
    names = [object_categories[x['category']] for x in infos]
    R = head.num_relation_classes
    target = relation_targets(meta['gt_rels'][0], n, R, feat)                        # [n, n, R]
    positives = torch.nonzero(target, as_tuple=False)

    # ---- relation queries on the sampled pairs (v4:146-185) -----------------------------------------------------------
    B = n * n
    enc = head.relation_qformer_tokenizer([head.qformer_instruction.format(names[p // n], names[p % n]) for p in range(B)],
                                          return_tensors="pt", padding=True, return_attention_mask=True)
    q_ids, q_mask = enc['input_ids'].to(dev), enc['attention_mask'].to(dev)
    query = torch.cat([head.rel_cls_query, head.relation_query], dim=1)             # [1, 33, d]
    nq = query.shape[1]
    tokens = head.patch_embed.proj(feat).flatten(2).transpose(1, 2)                 # timm PatchEmbed (v4:362)
    obj = token_masks_train(head, feat, meta, inputs['gt_masks'][0], inputs['gt_semantic_seg'][0])
    idx = sample_pairs(head, target).to(dev)
    S = idx.numel()
    pair_mask = (obj[idx // n] | obj[idx % n])[:, None, :].expand(-1, nq, -1)       # [S, 33, L]
    attn = torch.cat([torch.ones((S, nq), device=dev), q_mask[idx]], dim=1)
    out = head.relation_qformer(input_ids=q_ids[idx], attention_mask=attn, query_embeds=query.expand(S, -1, -1),
                                encoder_hidden_states=tokens.expand(S, -1, -1),
                                encoder_attention_mask=pair_mask)['last_hidden_state'][:, :nq]
    losses: Dict[str, torch.Tensor] = {}
    cls = out[:, 0]
    if 'binary' in head.rel_cls_type:
        exists = (target.sum(2) > 0).to(target.dtype).view(-1)[idx]
        losses['binary_rel_cls_loss'] = rel_cls_loss(head, head.binary_rel_cls_pred(cls).view(-1), exists)
    if 'multiclass' in head.rel_cls_type:
        losses['multiclass_rel_cls_loss'] = rel_cls_loss(head, head.multiclass_rel_cls_pred(cls).view(-1, R),
                                                         target.view(-1, R)[idx])
    # rows of pairs that were not sampled stay zero, as in the reference's scatter into a zero tensor (v4:178,186)
    pair_feature = out.new_zeros((B, nq - 1, out.shape[-1])).index_copy(0, idx, out[:, 1:])

    # ---- teacher-forced LM loss on a few ground-truth pairs (v4:219-228,267-285,327-341) --------------------------------
    chosen: List[int] = [int(i) * n + int(j) for i, j, _ in positives.tolist()]
    chosen = random.sample(chosen, min(len(chosen), head.max_llm_forward_num))
    if not chosen:
        chosen = random.sample(list(range(B)), min(B, head.max_llm_forward_num))
    tok = head.llm_tokenizer
    tok.padding_side = 'left'
    prompt = tok([head.llm_instruction.format(names[s // n], names[s % n]) for s in chosen],
                 return_tensors="pt", padding=True, return_attention_mask=True)
    flat = target.view(-1, R).tolist()
    answers = [''.join(' {} </s>'.format(head.relation_classes[r]) for r, on in enumerate(flat[s]) if on) for s in chosen]
    tok.padding_side = 'right'
    ans = tok(answers, return_tensors="pt", padding=True, return_attention_mask=True)
    ans_ids, ans_mask = ans['input_ids'].to(dev), ans['attention_mask'].to(dev)
    ids = torch.cat([prompt['input_ids'].to(dev), ans_ids], dim=1)
    mask = torch.cat([prompt['attention_mask'].to(dev), ans_mask], dim=1)
    T_ans = ans_ids.shape[1]
    embed = head.language_model.get_input_embeddings()
    per_pair = []
    if is_generation is None:
        is_generation = False
    if is_generation:
        raise NotImplementedError("is_generation=True in training mode (generate() under autograd) is not supported; "
                                  "call head.eval() for generation")
    for row, s in enumerate(chosen):
        u = head.language_projection(pair_feature[s])                                # [32, d_llm]
        x = torch.cat([u, embed(ids[row])], dim=0)[None]
        m = torch.cat([torch.ones(u.shape[0], dtype=torch.long, device=dev), mask[row]], dim=0)[None]
        logits = head.language_model(inputs_embeds=x, attention_mask=m).logits[:, -T_ans:, :]
        labels = torch.where(ans_mask[row:row + 1].bool(), ans_ids[row:row + 1], torch.full_like(ans_ids[row:row + 1], IGNORE))
        # next-token loss INSIDE the answer segment only (the first answer token is not predicted from the prompt)
        per_pair.append(F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), labels[:, 1:].reshape(-1),
                                        ignore_index=IGNORE, reduction="mean"))
    losses['rel_llm_loss'] = torch.stack(per_pair).mean()
    return losses
