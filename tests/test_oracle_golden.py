"""Pin the CPU oracle (oracle/) to the golden fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from object_detectors_b200 import synthetic as syn
from oracle import cref, tv_ref, yolo_ref

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _idf(name):
    return torch.from_numpy(np.load(os.path.join(G, f"idf_{name}_smooth.npy")))


def _check_post(gold, recs, num_classes):
    assert int(gold["num_images"]) == len(recs)
    for i, rec in enumerate(recs):
        # candidate lists: indices bit-exact, values bit-exact (same torch ops, same order)
        np.testing.assert_array_equal(gold[f"anchor_{i}"], rec["anchor"].numpy().astype(np.int32))
        np.testing.assert_array_equal(gold[f"det6_{i}"], rec["det6"].numpy())
        np.testing.assert_array_equal(gold[f"kept_{i}"], rec["kept"].numpy())
        # the C restatement of the NMS loop agrees with the torch-level one
        ki, kl = cref.nms_majority(rec["det6"].numpy(), 0.6, num_classes)
        np.testing.assert_array_equal(ki, rec["keep"].numpy().astype(np.int32))
        np.testing.assert_array_equal(kl.astype(np.float32), rec["kept"][:, 5].numpy())


@pytest.mark.parametrize("tag,softmax,use_idf,sig", [("softmax_idf", True, True, False),
                                                     ("sigmoid_plain", False, False, True)])
def test_decode_tiny_bit_exact(tag, softmax, use_idf, sig):
    gold = np.load(os.path.join(G, f"decode_tiny_{tag}.npz"))
    heads = syn.yolo_heads(11, 2, 64, 80, syn.COCO_ANCHORS, "clustered", sigmoid_cls=sig, max_objects=3)
    assert _sha(heads) == str(gold["sha"])
    th = [torch.from_numpy(h) for h in heads]
    idf = _idf("coco") if use_idf else None
    pred = yolo_ref.decode(th, syn.COCO_ANCHORS, 64, 80, idf, softmax)
    np.testing.assert_array_equal(gold["pred"], pred.numpy())
    recs = yolo_ref.score_filter(pred, 0.1)
    for r in recs:
        r["kept"], r["keep"] = yolo_ref.nms_majority(r["det6"], 0.6)
    _check_post(gold, recs, 80)


@pytest.mark.parametrize("tag,use_idf", [("idf", True), ("plain", False)])
@pytest.mark.parametrize("seed", [3, 110])
def test_c1_416(tag, use_idf, seed):
    gold = np.load(os.path.join(G, f"c1_416_{tag}_seed{seed}.npz"))
    heads = syn.yolo_heads(seed, 1, 416, 80, syn.COCO_ANCHORS, "clustered")
    assert _sha(heads) == str(gold["sha"])
    th = [torch.from_numpy(h) for h in heads]
    idf = _idf("coco") if use_idf else None
    pred = yolo_ref.decode(th, syn.COCO_ANCHORS, 416, 80, idf, True)
    assert pred.shape == (1, 10647, 85)
    np.testing.assert_array_equal(gold["sample_pred"], pred[0, torch.from_numpy(gold["sample_rows"])].numpy())
    np.testing.assert_array_equal(gold["colsum"], pred.double().sum(dim=1).numpy())
    recs = yolo_ref.postprocess(th, syn.COCO_ANCHORS, 416, 80, idf, True)
    _check_post(gold, recs, 80)
    assert sum(r["det6"].shape[0] for r in recs) > 50   # the generator really produces clusters


def test_c2_608_b4():
    gold = np.load(os.path.join(G, "c2_608_b4_seed203.npz"))
    heads = syn.yolo_heads(203, 4, 608, 80, syn.COCO_ANCHORS, "clustered")
    assert _sha(heads) == str(gold["sha"])
    th = [torch.from_numpy(h) for h in heads]
    recs = yolo_ref.postprocess(th, syn.COCO_ANCHORS, 608, 80, _idf("coco"), True)
    assert syn.num_anchors_total(608) == 22743
    _check_post(gold, recs, 80)


def test_c3_lvis():
    gold = np.load(os.path.join(G, "c3_lvis_96_b2_seed6.npz"))
    heads = syn.yolo_heads(6, 2, 96, 1203, syn.LVIS_ANCHORS, "clustered", max_objects=4)
    assert _sha(heads) == str(gold["sha"])
    th = [torch.from_numpy(h) for h in heads]
    pred = yolo_ref.decode(th, syn.LVIS_ANCHORS, 96, 1203, _idf("lvis"), True)
    np.testing.assert_array_equal(gold["colsum"], pred.double().sum(dim=1).numpy())
    recs = yolo_ref.score_filter(pred, 0.1)
    for r in recs:
        r["kept"], r["keep"] = yolo_ref.nms_majority(r["det6"], 0.6)
    _check_post(gold, recs, 1203)


def test_nms_majority_boxes_and_relabel():
    gold = np.load(os.path.join(G, "nms_majority_boxes.npz"))
    relabels = 0
    for i in range(4):
        seed, n, k, c = (int(v) for v in gold[f"args_{i}"])
        boxes, scores, labels = syn.random_boxes(seed, n, clusters=k, num_classes=c)
        det6 = np.concatenate([boxes, scores[:, None], labels[:, None].astype(np.float32)], 1)
        rows, keep = yolo_ref.nms_majority(torch.from_numpy(det6), 0.6)
        np.testing.assert_array_equal(gold[f"kept_{i}"], rows.numpy())
        ki, kl = cref.nms_majority(det6, 0.6, c)
        np.testing.assert_array_equal(ki, keep.numpy().astype(np.int32))
        np.testing.assert_array_equal(kl.astype(np.float32), rows[:, 5].numpy())
        relabels += int((rows[:, 5].numpy() != det6[keep.numpy(), 5]).sum())
    assert relabels > 0, "fixtures must exercise the majority relabel rule"


def test_match_416():
    gold = np.load(os.path.join(G, "match_416.npz"))
    cx, inw = yolo_ref.grid_table(syn.COCO_ANCHORS, 416, (13, 26, 52))
    assert cx.shape == (10647, 4)
    np.testing.assert_array_equal(gold["cxypwh_sample"], cx[::101].numpy())
    targets = syn.gt_targets(31, 3, 80, max_gt=20)
    tt = [{k: torch.from_numpy(v) for k, v in t.items()} for t in targets]
    for kind in (0, 1):
        tgt, tcls, obj, noobj = yolo_ref.get_target(tt, cx, inw, 80, 0.5, kind)
        np.testing.assert_array_equal(gold[f"tgt_{kind}"], tgt.numpy())
        np.testing.assert_array_equal(gold[f"obj_{kind}"], torch.cat(obj).numpy())
        np.testing.assert_array_equal(gold[f"noobj_{kind}"], np.packbits(noobj.numpy()))
        np.testing.assert_array_equal(gold[f"tcls_argmax_{kind}"], tcls.argmax(1).numpy())
        # scalar C restatement: same argmax / mask
        for t, o, nm in zip(targets, obj, noobj):
            best, free, _ = cref.iou_match(t["bbox"], cx.numpy(), kind, 0.5)
            np.testing.assert_array_equal(best, o.numpy())
            np.testing.assert_array_equal(free, nm.numpy())
    b1 = torch.from_numpy(targets[0]["bbox"])
    for kind in (0, 1, 2, 3):
        got = yolo_ref.bbox_iou(b1.unsqueeze(1), cx[::53].unsqueeze(0), kind)
        np.testing.assert_array_equal(gold[f"iou_kind{kind}"], got.numpy())
    _, _, iou_c = cref.iou_match(targets[0]["bbox"], cx[::53].numpy(), 1, 0.5, want_iou=True)
    np.testing.assert_array_equal(gold["iou_kind1"], iou_c)


def test_rpn_filter():
    gold = np.load(os.path.join(G, "rpn_filter.npz"))
    for tag in ("s", "m"):
        seed, bsz, ih, iw, pre, post = (int(v) for v in gold[f"{tag}_args"])
        obj, deltas, anchors, per_level = syn.rpn_inputs(seed, bsz, ih, iw)
        fb, fs, _ = tv_ref.filter_proposals(torch.from_numpy(obj), torch.from_numpy(deltas),
                                            torch.from_numpy(anchors), per_level, [(ih, iw)] * bsz,
                                            pre, post)
        for i in range(bsz):
            np.testing.assert_array_equal(gold[f"{tag}_boxes_{i}"], fb[i].numpy())
            np.testing.assert_array_equal(gold[f"{tag}_scores_{i}"], fs[i].numpy())


@pytest.mark.parametrize("tag", ["h13", "h38"])
def test_legacy_yolo_loss_decode(tag):
    """oracle.yolo_ref.legacy_decode == YOLOLoss.forward(input) of the reference (yolo_loss.py:34-105), bit for bit."""
    gold = np.load(os.path.join(G, "legacy_yolo_loss.npz"))
    seed, grid, head_idx, classes, img = [int(v) for v in gold[f"{tag}_args"]]
    x = torch.from_numpy(syn.legacy_head(seed, 2, 3, classes, grid))
    out = yolo_ref.legacy_decode(x, syn.COCO_ANCHORS[head_idx], classes, img)
    rows = torch.from_numpy(gold[f"{tag}_rows"])
    np.testing.assert_array_equal(out[:, rows].numpy(), gold[f"{tag}_sample"])
    np.testing.assert_array_equal(out.double().sum(dim=1).numpy(), gold[f"{tag}_colsum"])


@pytest.mark.parametrize("tag,activation", [("ce", "ce"), ("gombit", "gombit_x"), ("sigmoid", "bce")])
def test_roi_postprocess(tag, activation):
    """oracle.tv_ref.roi_postprocess == RoIHeads.postprocess_detections of the reference (roi_heads.py:715-781)."""
    gold = np.load(os.path.join(G, "roi_postprocess.npz"))
    seed, r0, r1, c, ih, iw = [int(v) for v in gold["args"]]
    logits, regs, props = syn.roi_inputs(seed, [r0, r1], c, ih, iw)
    res = tv_ref.roi_postprocess(torch.from_numpy(logits), torch.from_numpy(regs), [torch.from_numpy(p) for p in props],
                                 [(ih, iw)] * 2, torch.from_numpy(gold["idf"]), activation, strategy="torchvision")
    for i, (b, s, l, _) in enumerate(res):
        np.testing.assert_array_equal(l.numpy(), gold[f"{tag}_labels_{i}"])
        np.testing.assert_array_equal(s.numpy(), gold[f"{tag}_scores_{i}"])
        np.testing.assert_array_equal(b.numpy(), gold[f"{tag}_boxes_{i}"])


def test_retinanet_postprocess_oracle_matches_reference_golden():
    """oracle/tv_ref.retinanet_postprocess == the UNMODIFIED reference's RetinaNet.postprocess_detections
    (tests/golden/retinanet_postprocess.npz, written by make_golden_extra.py), bit for bit."""
    import os
    import numpy as np
    import torch
    from object_detectors_b200 import synthetic as syn
    from oracle import tv_ref
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "retinanet_postprocess.npz"))
    seed, bsz, ih, iw, c = [int(v) for v in g["args"]]
    logits, regs, anchors, per_level = syn.retina_inputs(seed, bsz, ih, iw, c)
    out = tv_ref.retinanet_postprocess(torch.from_numpy(logits), torch.from_numpy(regs), torch.from_numpy(anchors), per_level,
                                       [(ih, iw)] * bsz, torch.from_numpy(g["idf"]))
    for i, (b, s, l) in enumerate(out):
        np.testing.assert_array_equal(b.numpy(), g[f"boxes_{i}"])
        np.testing.assert_array_equal(s.numpy(), g[f"scores_{i}"])
        np.testing.assert_array_equal(l.numpy(), g[f"labels_{i}"])


def test_ssd_postprocess_oracle_matches_reference_golden():
    """oracle/tv_ref.ssd_postprocess == the UNMODIFIED reference's SSD.postprocess_detections
    (tests/golden/ssd_postprocess.npz, written by make_golden_extra.py), bit for bit."""
    import os
    import numpy as np
    import torch
    from object_detectors_b200 import synthetic as syn
    from oracle import tv_ref
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ssd_postprocess.npz"))
    seed, bsz, a, c, img = [int(v) for v in g["args"]]
    logits, regs, anchors = syn.ssd_inputs(seed, bsz, a, c, img)
    out = tv_ref.ssd_postprocess(torch.from_numpy(logits), torch.from_numpy(regs), [torch.from_numpy(anchors)] * bsz,
                                 [(img, img)] * bsz, torch.from_numpy(g["idf"]))
    for i, (b, s, l) in enumerate(out):
        np.testing.assert_array_equal(b.numpy(), g[f"boxes_{i}"])
        np.testing.assert_array_equal(s.numpy(), g[f"scores_{i}"])
        np.testing.assert_array_equal(l.numpy(), g[f"labels_{i}"])
