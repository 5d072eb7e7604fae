"""TEST INFRASTRUCTURE / CPU baseline — never imported by the product path.

A port of the reference head's *inference* control flow (``relation_transformer_head_v4.py:134-326,
408-435``) that, like the reference, delegates the arithmetic to the live HuggingFace modules
(``InstructBlipQFormerModel``, ``OPTForCausalLM.generate``) with the reference's own call pattern:
image tokens expanded to N^2 copies, one ``[N^2,33,L]`` bool mask, K/V re-projected per pair, one
batch-1 ``generate`` per selected pair.  /root/reference does not exist on the GPU box, the
transformers wheel does, so this is (a) the oracle the end-to-end GPU parity tests compare against
and (b) the ``cpu_baseline`` (kind "port") that ``bench.py`` times on the box's host cores.

tests/test_oracle.py pins it against the unmodified reference file (via tests/golden fixtures made by
oracle/make_golden.py).  Module / parameter names equal the reference's (v4:75-105) so
``openpsg_b200.synth.init_parameters`` gives it, the reference and the product identical weights.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from openpsg_b200.categories import INSTANCE_OFFSET, object_categories, relation_categories
from openpsg_b200.synth import SyntheticTokenizer


class _PatchEmbed(nn.Module):  # timm.layers.PatchEmbed as used at v4:75-76 (img_size=None, no norm)
    def __init__(self, patch, cin, cout):
        super().__init__()
        self.proj = nn.Conv2d(cin, cout, kernel_size=patch, stride=patch, bias=True)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class ReferencePortHead(nn.Module):
    def __init__(self, llm_config: Optional[dict], llm_feature_size: int = 4096, max_object_num: int = 30,
                 qformer_instruction='Is there a relation between {} and {}?',
                 llm_instruction='What are the relations between {} and {}? Assistant: ',
                 patch_size=16, topk_pairs=20, max_new_tokens=16):
        super().__init__()
        from transformers import InstructBlipQFormerConfig, InstructBlipQFormerModel
        self.qformer_instruction, self.llm_instruction = qformer_instruction, llm_instruction
        self.patch_size, self.max_object_num = patch_size, max_object_num
        self.topk_pairs, self.max_new_tokens = topk_pairs, max_new_tokens
        self.patch_embed = _PatchEmbed(patch_size, 256, 256)
        self.relation_qformer = InstructBlipQFormerModel(InstructBlipQFormerConfig(
            hidden_size=768, num_hidden_layers=2, cross_attention_frequency=1, encoder_hidden_size=256))
        self.relation_query = nn.Parameter(torch.randn(1, 32, 768))
        self.rel_cls_query = nn.Parameter(torch.randn(1, 1, 768))
        self.binary_rel_cls_pred = nn.Linear(768, 1)
        self.language_projection = nn.Linear(768, llm_feature_size)
        self.relation_qformer_tokenizer = SyntheticTokenizer("qformer")
        self.llm_tokenizer = SyntheticTokenizer("llm")
        self.language_model = None
        if llm_config is not None:
            from openpsg_b200.synth import build_causal_lm
            self.language_model = build_causal_lm(llm_config)      # OPT or Llama (config "model_type")
            self.llm_tokenizer.set_vocab_size(llm_config["vocab_size"])

    # -- v4:408-435 -----------------------------------------------------------------------------
    def prepare_inference(self, feat, meta, object_id_list, pan):
        tokens = self.patch_embed(feat)
        fh, fw = feat.shape[-2:]
        ih, iw = meta['img_shape'][:2]
        ph, pw = meta['pad_shape'][:2]
        pan = F.interpolate(pan[None, None].float(), size=(ih, iw), mode='nearest')
        pan = F.pad(pan, (0, pw - iw, 0, ph - ih), value=0)
        pan = F.interpolate(pan, size=(fh // self.patch_size, fw // self.patch_size), mode='nearest')[0, 0]
        masks = torch.stack([pan == oid for oid in object_id_list], dim=0)        # [N,th,tw]
        n = masks.shape[0]
        pair = torch.stack([masks[i] | masks[j] for i in range(n) for j in range(n)], dim=0)
        return tokens, pair.reshape(n * n, 1, -1), masks.reshape(n, -1)

    # -- v4:134-237 -----------------------------------------------------------------------------
    @torch.no_grad()
    def relation_queries(self, inputs: dict, pair_subset=None) -> dict:
        feat = inputs['mask_features']
        assert feat.shape[0] == 1
        meta = inputs['img_metas'][0]
        info = inputs['object_info'][0]
        ids = info['object_id_list'][:self.max_object_num]
        n = len(ids)
        names = [object_categories[int(x) % INSTANCE_OFFSET] for x in ids]
        B = n * n
        texts = [self.qformer_instruction.format(names[p // n], names[p % n]) for p in range(B)]
        enc = self.relation_qformer_tokenizer(texts, return_tensors="pt", padding=True, return_attention_mask=True)
        query = torch.cat([self.rel_cls_query, self.relation_query], dim=1).expand(B, -1, -1)
        attn = torch.cat([torch.ones(B, query.shape[1]), enc['attention_mask']], dim=1)
        tokens, pair_masks, obj_masks = self.prepare_inference(feat, meta, ids, info['pan_results'])
        tokens = tokens.expand(B, -1, -1)
        pair_masks = pair_masks.expand(-1, query.shape[1], -1)
        # test mode: every pair (v4:175); a subset = the reference's qformer_sampled_idxes mechanism (v4:173)
        idx = torch.arange(B) if pair_subset is None else torch.as_tensor(pair_subset, dtype=torch.long)
        out = self.relation_qformer(
            input_ids=enc['input_ids'][idx], attention_mask=attn[idx], query_embeds=query[idx],
            encoder_hidden_states=tokens[idx], encoder_attention_mask=pair_masks[idx],
        )['last_hidden_state'][:, :query.shape[1]]
        logits = self.binary_rel_cls_pred(out[:, 0])
        prob = torch.sigmoid(logits)
        selected = idx[prob.squeeze(1).topk(idx.numel()).indices].tolist()[:self.topk_pairs]
        return dict(object_num=n, names=names, qformer_out=out, exist_logits=logits.squeeze(1),
                    exist_prob=prob.squeeze(1), selected=selected, obj_masks=obj_masks,
                    image_tokens=tokens[0], input_ids=enc['input_ids'], text_mask=enc['attention_mask'])

    # -- v4:259-326 -----------------------------------------------------------------------------
    @torch.no_grad()
    def decode_relations(self, q: dict, max_pairs: Optional[int] = None) -> dict:
        n, names, sel = q['object_num'], q['names'], q['selected']
        if max_pairs is not None:
            sel = sel[:max_pairs]
        pair_feature = q['qformer_out'][:, 1:]
        texts = [self.llm_instruction.format(names[s // n], names[s % n]) for s in sel]
        self.llm_tokenizer.padding_side = 'left'
        enc = self.llm_tokenizer(texts, return_tensors="pt", padding=True, return_attention_mask=True)
        rel_pred, rel_score, seqs, scores = [], [], [], []
        for i, s in enumerate(sel):
            u = self.language_projection(pair_feature[s])
            e = self.language_model.get_input_embeddings()(enc['input_ids'][i])
            embeds = torch.cat([u, e], dim=0)[None]
            mask = torch.cat([torch.ones(u.shape[0], dtype=torch.long), enc['attention_mask'][i]], dim=0)[None]
            out = self.language_model.generate(
                inputs_embeds=embeds, attention_mask=mask, max_new_tokens=self.max_new_tokens,
                min_new_tokens=self.max_new_tokens, num_beams=1, do_sample=False,
                return_dict_in_generate=True, output_scores=True)
            seqs.append(out.sequences[0])
            scores.append(torch.stack([x[0] for x in out.scores]))
            text = self.llm_tokenizer.batch_decode(out.sequences)[0]
            pred = text.split('<s>')[1].split('</s>')[0].strip()
            for name in pred.split('  '):
                if name in relation_categories:
                    trip = [s // n, s % n, relation_categories.index(name)]
                    if trip not in rel_pred:
                        rel_pred.append(trip)
                        rel_score.append(1)
        return dict(rel_pred=rel_pred, rel_score=rel_score, sequences=seqs, scores=scores,
                    llm_input_ids=enc['input_ids'], llm_mask=enc['attention_mask'])

    def forward(self, inputs: dict) -> dict:
        q = self.relation_queries(inputs)
        if self.language_model is None:
            return {'rel_pred': [], 'rel_score': [], **q}
        d = self.decode_relations(q)
        return {'rel_pred': d['rel_pred'], 'rel_score': d['rel_score']}
