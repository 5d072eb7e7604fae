"""Detector-side glue around the relation head (rows f1 / f2 of SURVEY.md §8), on the device.

f1 — ``relabel_panoptic``: what ``OpenSeeDRelationV2.forward_openseed`` does after the segmenter
(``kings_sgg/models/detectors/openseed_relation_v2.py:112-128``): the segmenter's panoptic map carries arbitrary segment ids;
the head wants ``category + 1000 * instance`` ids (instance = running count per category in list order) plus the
``object_id_list``.  The reference copies the map to the host, runs one ``np.where`` pass per segment and copies it back;
here the id tables (a few dozen ints, host data already) go to the device and one kernel rewrites the map — no D2H, no sync.

f2 — ``encode_submission_image`` / ``submission_record`` / ``write_submission``: the result wire format of
``tools/infer.py:149-187`` (``relation.json`` + RGB-encoded panoptic PNGs) that the external HiLo / PSG grader reads.
"""
from __future__ import annotations

import json
import os
import random as _random
import struct
import zlib
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from .categories import INSTANCE_OFFSET


def segment_id_tables(segments_info: Sequence[dict]) -> Tuple[List[int], List[int]]:
    """(segment ids, category + 1000 * instance ids) in list order; instance = how many earlier segments share the category
    (openseed_relation_v2.py:116-124)."""
    seen: dict = {}
    old, new = [], []
    for seg in segments_info:
        cat = int(seg['category_id'])
        seen[cat] = seen.get(cat, -1) + 1
        old.append(int(seg['id']))
        new.append(cat + INSTANCE_OFFSET * seen[cat])
    return old, new


def relabel_panoptic(pan_seg: torch.Tensor, segments_info: Sequence[dict]):
    """pan_seg: int tensor [H, W] of segment ids ON THE DEVICE -> (pan_results int32 [H, W] on the device, object_id_list,
    object_score_list) exactly as openseed_relation_v2.py:112-128 builds them."""
    if not pan_seg.is_cuda:
        raise ValueError("relabel_panoptic runs on the device (libopsg_b200 has no CPU path)")
    old, new = segment_id_tables(segments_info)
    src = pan_seg.to(torch.int32).contiguous()
    out = torch.empty_like(src)
    if old:
        tab = torch.tensor([old, new], dtype=torch.int32).pin_memory().to(src.device, non_blocking=True)
        t_old, t_new = tab[0], tab[1]
    else:
        t_old = t_new = None
    with ops._timed("pan_relabel", 0.0, 8.0 * src.numel()):
        _lib.check(_lib.load().opsg_pan_relabel(ops._ptr(src), src.numel(), ops._ptr(t_old), ops._ptr(t_new), len(old), ops._ptr(out),
                                               ops._stream()))
    ops._count()
    object_id_list = [torch.tensor(x, dtype=torch.int32) for x in new]
    object_score_list = [torch.tensor(1.0) for _ in new]
    return out, object_id_list, object_score_list


# ---- f2 ------------------------------------------------------------------------------------------------------------

def rgb2id(color) -> int:
    """panopticapi.utils.rgb2id for one colour triple (tools/infer.py:12,165)."""
    r, g, b = (int(c) for c in color)
    return r + 256 * g + 256 * 256 * b


def encode_submission_image(pan_results: torch.Tensor, object_id_list: Sequence[int], rng=_random):
    """tools/infer.py:149-168: one random colour per listed object (``random.choices(range(0, 255), k=3)``, object id 133 =
    background is skipped), painted over the pixels it owns.  pan_results: int tensor [H, W] on the device.
    -> (rgb uint8 tensor [H, W, 3] on the device, segments_info list)."""
    if not pan_results.is_cuda:
        raise ValueError("encode_submission_image runs on the device (libopsg_b200 has no CPU path)")
    ids, colors, segments_info = [], [], []
    for object_id in object_id_list:
        object_id = int(object_id)
        if object_id == 133:
            continue
        r, g, b = rng.choices(range(0, 255), k=3)
        ids.append(object_id)
        colors.append((r, g, b))
        segments_info.append(dict(category_id=int(object_id % INSTANCE_OFFSET + 1), id=rgb2id((r, g, b))))
    pan = pan_results.to(torch.int32).contiguous()
    rgb = torch.empty(tuple(pan.shape) + (3,), dtype=torch.uint8, device=pan.device)
    t_ids = torch.tensor(ids, dtype=torch.int32).to(pan.device) if ids else None
    t_col = torch.tensor(colors, dtype=torch.uint8).to(pan.device) if ids else None
    with ops._timed("pan_colorize", 0.0, 7.0 * pan.numel()):
        _lib.check(_lib.load().opsg_pan_colorize(ops._ptr(pan), pan.numel(), ops._ptr(t_ids), ops._ptr(t_col), len(ids), ops._ptr(rgb),
                                                ops._stream()))
    ops._count()
    return rgb, segments_info


def submission_record(relation: Sequence[Sequence[int]], segments_info: List[dict], test_idx: int, rng=_random) -> dict:
    """tools/infer.py:171-185: relations with 1-based predicate ids, the placeholders for empty results."""
    relation = [list(map(int, r)) for r in relation]
    if len(relation) == 0:
        relation = [[0, 0, 0]]
    if len(segments_info) == 0:
        r, g, b = rng.choices(range(0, 255), k=3)
        segments_info = [dict(category_id=1, id=rgb2id((r, g, b)))]
    return dict(relations=[[s, o, r + 1] for s, o, r in relation], segments_info=segments_info,
                pan_seg_file_name='%d.png' % test_idx)


def png_bytes(rgb: np.ndarray) -> bytes:
    """Minimal 8-bit RGB PNG encoder (cv2 / PIL are not needed for the submission files)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()     # filter byte 0 per row

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def write_submission(output_dir: str, results: Sequence[dict], rng=_random) -> str:
    """results: per image ``{'pan_results': device int tensor [H, W], 'rel_results': {'object_id_list': [...], 'relation':
    [[s, o, r], ...]}}`` (what ``simple_test`` returns, openseed_relation_v2.py:183-190).  Writes ``submission/panseg/<i>.png``
    and ``submission/relation.json`` (tools/infer.py:64-68,169,188-190); returns the json path."""
    panseg_dir = os.path.join(output_dir, 'submission/panseg')
    os.makedirs(panseg_dir, exist_ok=True)
    records = []
    for test_idx, res in enumerate(results):
        rgb, segments_info = encode_submission_image(res['pan_results'], res['rel_results']['object_id_list'], rng)
        with open(os.path.join(panseg_dir, '%d.png' % test_idx), 'wb') as f:
            f.write(png_bytes(rgb.cpu().numpy()))
        records.append(submission_record(res['rel_results']['relation'], segments_info, test_idx, rng))
    path = os.path.join(output_dir, 'submission', 'relation.json')
    with open(path, 'w') as f:
        json.dump(records, f, default=str)
    return path
