"""Rows f1 / f2: on-device panoptic relabel (openseed_relation_v2.py:112-128) and the submission wire format
(tools/infer.py:149-187).  Integer / byte work: bit-exact against the numpy restatement; the restatement of f1 is pinned to
the reference's own statements where /root/reference exists (oracle/ref_shims.py is not needed: the loop is sliced here)."""
import json
import random
from pathlib import Path

import numpy as np
import pytest
import torch

from openpsg_b200 import detector_glue as glue
from oracle import restated

REF = Path("/root/reference/kings_sgg/models/detectors/openseed_relation_v2.py")


def _case(seed=0, h=97, w=131, n_seg=9):
    g = np.random.RandomState(seed)
    seg_ids = g.permutation(np.arange(1, 40))[:n_seg]
    pan = g.choice(np.concatenate([seg_ids, [0, 77]]), size=(h, w)).astype(np.int64)      # 0 / 77: pixels of unlisted segments
    cats = g.randint(0, 133, size=n_seg)
    cats[3] = cats[1]                                                                      # two instances of one category
    cats[6] = cats[1]
    info = [dict(id=int(i), category_id=int(c), isthing=bool(c < 80)) for i, c in zip(seg_ids, cats)]
    return pan, info


def test_segment_tables_and_restatement_follow_the_reference_loop():
    pan, info = _case()
    old, new = glue.segment_id_tables(info)
    ref_pan, ref_ids = restated.relabel_panoptic(pan, info)
    assert new == ref_ids and old == [s["id"] for s in info]
    assert new[3] == info[1]["category_id"] + 1000 and new[6] == info[1]["category_id"] + 2000
    assert (ref_pan[(pan == 0) | (pan == 77)] == 0).all()
    if REF.exists():      # pin the restatement to the reference's own statements (authoring container only)
        src = REF.read_text()
        i = src.index("        _pan_results = openseed_output['panoptic_seg'][0].cpu().numpy()")
        j = src.index("        object_score_list = [torch.tensor(1.0)")
        import textwrap
        body = "def run(openseed_output, img, np, torch, INSTANCE_OFFSET):\n" + src[i:j] + "        return pan_results, object_id_list\n"
        ns = {}
        exec(compile(textwrap.dedent(body.replace("\n        ", "\n    ")), str(REF), "exec"), ns)
        out, ids = ns["run"]({"panoptic_seg": (torch.from_numpy(pan), info)}, torch.zeros(1), np, torch, 1000)
        assert np.array_equal(out.numpy(), ref_pan) and [int(x) for x in ids] == ref_ids


def test_submission_record_and_png_roundtrip(tmp_path):
    rec = glue.submission_record([], [], 3, random.Random(1))
    assert rec["relations"] == [[0, 0, 1]] and rec["pan_seg_file_name"] == "3.png" and rec["segments_info"][0]["category_id"] == 1
    rec = glue.submission_record([[1, 2, 5], [0, 1, 55]], [dict(category_id=4, id=9)], 0)
    assert rec["relations"] == [[1, 2, 6], [0, 1, 56]]
    rgb = np.random.RandomState(0).randint(0, 255, size=(13, 7, 3)).astype(np.uint8)
    data = glue.png_bytes(rgb)
    PIL = pytest.importorskip("PIL.Image")
    p = tmp_path / "a.png"
    p.write_bytes(data)
    assert np.array_equal(np.asarray(PIL.open(p).convert("RGB")), rgb)
    assert glue.rgb2id((1, 2, 3)) == 1 + 2 * 256 + 3 * 65536


@pytest.mark.gpu
def test_relabel_and_submission_on_device(tmp_path):
    for seed, (h, w) in enumerate([(97, 131), (1024, 1024), (5, 3)]):
        pan, info = _case(seed, h, w)
        got, ids, scores = glue.relabel_panoptic(torch.from_numpy(pan).cuda(), info)
        ref_pan, ref_ids = restated.relabel_panoptic(pan, info)
        assert got.dtype == torch.int32 and np.array_equal(got.cpu().numpy(), ref_pan)
        assert [int(x) for x in ids] == ref_ids and all(x.dtype == torch.int32 and x.dim() == 0 for x in ids) and len(scores) == len(ids)
        obj_ids = ref_ids + [133, ref_ids[0]]                  # background id is skipped; a repeated id paints twice (uint8 sum)
        rgb, seg = glue.encode_submission_image(got, obj_ids, random.Random(5))
        ref_rgb, ref_seg = restated.submission_image(ref_pan, obj_ids, random.Random(5))
        assert seg == ref_seg and np.array_equal(rgb.cpu().numpy(), ref_rgb)
    got0, _, _ = glue.relabel_panoptic(torch.from_numpy(pan).cuda(), [])
    assert int(got0.abs().sum()) == 0
    results = [{"pan_results": got, "rel_results": {"object_id_list": ref_ids, "relation": [[0, 1, 3]]}},
               {"pan_results": got, "rel_results": {"object_id_list": [], "relation": []}}]
    path = glue.write_submission(str(tmp_path), results, random.Random(2))
    recs = json.loads(Path(path).read_text())
    assert len(recs) == 2 and recs[0]["relations"] == [[0, 1, 4]] and recs[1]["relations"] == [[0, 0, 1]]
    assert (tmp_path / "submission/panseg/0.png").exists() and (tmp_path / "submission/panseg/1.png").exists()
    PIL = pytest.importorskip("PIL.Image")
    img = np.asarray(PIL.open(tmp_path / "submission/panseg/0.png").convert("RGB")).astype(np.int64)
    ids_img = img[..., 0] + 256 * img[..., 1] + 65536 * img[..., 2]          # rgb2id of the decoded PNG (tools/parse_predict.py:47-52)
    for seg_rec, oid in zip(recs[0]["segments_info"], ref_ids):
        assert np.array_equal(ids_img == seg_rec["id"], ref_pan == oid) and seg_rec["category_id"] == oid % 1000 + 1
