"""GPU tests of the reference-signature drop-ins (object_detectors_b200.yolo.*, .tvision.*): the
Python call sites of the reference keep their names, arguments and layouts, so these tests read like
calls into the reference; expected values come from the CPU oracle / golden fixtures."""
import os
import types

import numpy as np
import pytest
import torch

from object_detectors_b200 import synthetic as syn
from oracle import cref, tv_ref, yolo_ref

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def _cfg(img, classes, anchors, dset, class_loss, tfidf):
    yolo = AttrDict(classes=classes, img_size=img, ignore_threshold=0.5, iou_type=1, class_loss=class_loss,
                    inf_confidence=0.1, inf_iou_threshold=0.6, tfidf=tfidf, tfidf_variant="smooth", tfidf_norm=0)
    dataset = AttrDict(anchors=[[list(p) for p in s] for s in anchors], dset_name=dset, inp_dim=img,
                       num_classes=classes)
    return AttrDict(yolo=yolo, dataset=dataset)


def _idf(name):
    return torch.from_numpy(np.load(os.path.join(G, f"idf_{name}_smooth.npy")))


# ------------------------------------------------------------------------------------- helper.py
def test_helper_get_abs_coord_and_bbox_iou():
    from object_detectors_b200.yolo.utilities import helper
    g = np.random.Generator(np.random.PCG64(5))
    box = torch.from_numpy(g.random((7, 33, 4)).astype(np.float32))
    np.testing.assert_array_equal(helper.get_abs_coord(box.cuda()).cpu().numpy(), yolo_ref.abs_coord(box).numpy())
    np.testing.assert_array_equal(helper.get_abs_coord(box[0].cuda()).cpu().numpy(), yolo_ref.abs_coord(box[0]).numpy())
    cx, _ = yolo_ref.grid_table(syn.COCO_ANCHORS, 416, (13, 26, 52))
    gt = torch.from_numpy(syn.gt_targets(31, 1, 80, max_gt=20, min_gt=20)[0]["bbox"])
    for kind in (0, 1, 2):
        want = yolo_ref.bbox_iou(gt.unsqueeze(1), cx.unsqueeze(0), kind).numpy()
        got = helper.bbox_iou(gt.cuda().unsqueeze(1), cx.cuda().unsqueeze(0), kind).cpu().numpy()
        np.testing.assert_array_equal(got, want)
        pw = yolo_ref.bbox_iou(gt, cx[:20], kind).numpy()
        np.testing.assert_array_equal(helper.bbox_iou(gt.cuda(), cx[:20].cuda(), kind).cpu().numpy(), pw)
    # differentiable call (loss side): same values, gradient flows
    a = gt.cuda().clone().requires_grad_(True)
    out = helper.bbox_iou(a, cx[:20].cuda(), 1)
    out.sum().backward()
    assert a.grad is not None and torch.isfinite(a.grad).all()
    np.testing.assert_allclose(out.detach().cpu().numpy(), yolo_ref.bbox_iou(gt, cx[:20], 1).numpy(), rtol=1e-6, atol=1e-7)


def test_helper_nms_majority_matches_reference_golden(nms_path):
    from object_detectors_b200.yolo.utilities import helper
    gold = np.load(os.path.join(G, "nms_majority_boxes.npz"))
    for i in range(4):
        seed, n, k, c = (int(v) for v in gold[f"args_{i}"])
        boxes, scores, labels = syn.random_boxes(seed, n, clusters=k, num_classes=c)
        P = torch.from_numpy(np.concatenate([boxes, scores[:, None], labels[:, None].astype(np.float32)], 1)).cuda()
        before = P.clone()
        kept = helper.nms_majority(P, 0.6)
        np.testing.assert_array_equal(kept.cpu().numpy(), gold[f"kept_{i}"])          # what the reference returned
        # in-place relabel of the kept rows, nothing else touched
        changed = (P != before).any(dim=1).nonzero().flatten().cpu().numpy()
        ki, kl = cref.nms_majority(before.cpu().numpy(), 0.6, c)
        relabelled = ki[kl != before.cpu().numpy()[ki, 5].astype(np.int32)]
        np.testing.assert_array_equal(np.sort(changed), np.sort(relabelled))


# ------------------------------------------------------------------------------------ YOLOForw
def test_yoloforw_forward_and_get_target():
    from object_detectors_b200.yolo.nets.yolo_forw import YOLOForw
    heads = syn.yolo_heads(3, 1, 416, 80, syn.COCO_ANCHORS, "clustered")
    idf = _idf("coco")
    model = YOLOForw(_cfg(416, 80, syn.COCO_ANCHORS, "coco", 1, [0, 0]), idf_logits=idf)
    pred = model([torch.from_numpy(h).cuda() for h in heads]).cpu().numpy()
    ref = yolo_ref.decode([torch.from_numpy(h) for h in heads], syn.COCO_ANCHORS, 416, 80, idf, True).numpy()
    assert pred.shape == (1, 10647, 85)
    assert np.all(np.abs(pred[..., :4] - ref[..., :4]) <= 1e-5 * np.maximum(np.abs(ref[..., :4]), 416))
    assert np.all(np.abs(pred[..., 4:] - ref[..., 4:]) <= 1e-5 * np.abs(ref[..., 4:]) + 1e-12)
    with pytest.raises(NotImplementedError):
        model([torch.from_numpy(h).cuda() for h in heads], targets=[{}])

    gold = np.load(os.path.join(G, "match_416.npz"))
    targets = syn.gt_targets(31, 3, 80, max_gt=20)
    tt = [{k: torch.from_numpy(v).cuda() for k, v in t.items()} for t in targets]
    cx, inw = model.grid_table((13, 26, 52))
    rcx, rinw = yolo_ref.grid_table(syn.COCO_ANCHORS, 416, (13, 26, 52))
    np.testing.assert_array_equal(cx.cpu().numpy(), rcx.numpy())
    for kind in (0, 1):
        model.iou_type = kind
        tgt, tcls, obj, noobj = model.get_target(tt, cx, inw, ignore_threshold=0.5)
        np.testing.assert_array_equal(torch.cat(obj).cpu().numpy(), gold[f"obj_{kind}"])
        np.testing.assert_array_equal(np.packbits(noobj.cpu().numpy()), gold[f"noobj_{kind}"])
        np.testing.assert_array_equal(tcls.argmax(1).cpu().numpy(), gold[f"tcls_argmax_{kind}"])
        np.testing.assert_allclose(tgt.cpu().numpy(), gold[f"tgt_{kind}"], rtol=1e-5, atol=1e-6)


# -------------------------------------------------------------------------------- test_one_epoch
class _Loader(list):
    dset_name = "coco"


def test_test_one_epoch_reproduces_reference_results():
    from object_detectors_b200.yolo.nets.yolo_forw import YOLOForw
    from object_detectors_b200.yolo.procedures.test_one_epoch import test_one_epoch as run
    cfg = _cfg(608, 80, syn.COCO_ANCHORS, "coco", 1, [0, 0])
    idf = _idf("coco")
    heads = syn.yolo_heads(203, 4, 608, 80, syn.COCO_ANCHORS, "clustered")
    # blank out image 1 so that the reference's "drop empty images" path (and its index shift) is hit
    for h in heads:
        h[1, 4::85] = -20.0
    yolo = YOLOForw(cfg, idf_logits=idf)
    model = types.SimpleNamespace(eval=lambda: None, __call__=None)

    class _Model:
        def eval(self):
            return self

        def __call__(self, images):
            return [torch.from_numpy(h).cuda() for h in heads]

    sizes = [(480, 640), (375, 500), (427, 640), (600, 800)]
    targets = [{"img_size": torch.tensor(s), "image_id": torch.tensor(1000 + i)} for i, s in enumerate(sizes)]
    loader = _Loader([(torch.zeros(4, 3, 8, 8), targets)])
    got = run(loader, _Model(), yolo, cfg)

    recs = yolo_ref.postprocess([torch.from_numpy(h) for h in heads], syn.COCO_ANCHORS, 608, 80, idf, True)
    assert recs[1]["det6"].shape[0] == 0
    kept = [r["kept"] for r in recs if r["kept"].shape[0] > 0]           # reference :34,:37
    coco91 = [c for c in range(1, 91) if c not in (12, 26, 29, 30, 45, 66, 68, 69, 71, 83)]
    want = []
    for i, rows in enumerate(kept):                                        # reference :41-66, targets[i] by position
        h, w = sizes[i]
        for r in rows.numpy():
            x1, y1, x2, y2 = r[0] / 608 * w, r[1] / 608 * h, r[2] / 608 * w, r[3] / 608 * h
            want.append((1000 + i, coco91[int(r[5])], float(r[4]), [x1, y1, x2 - x1, y2 - y1]))
    assert len(got) == len(want) > 0
    for g, (iid, cat, score, bbox) in zip(got, want):
        assert g["image_id"] == iid and g["category_id"] == cat
        assert abs(g["score"] - score) <= 1e-5 * abs(score)
        np.testing.assert_allclose(g["bbox"], bbox, rtol=1e-4, atol=1e-2)
        assert abs(g["area"] - g["bbox"][2] * g["bbox"][3]) <= 1e-3 * max(g["area"], 1.0)
    fixed = run(loader, _Model(), yolo, cfg, strict_reference=False)
    assert sorted({d["image_id"] for d in fixed}) == [1000, 1002, 1003]


# ------------------------------------------------------------------------------------- tvision
def test_tvision_boxes_match_torchvision():
    from torchvision.ops import boxes as tvb
    from object_detectors_b200.tvision import boxes as b2
    b, s, l = syn.random_boxes(7, 900, clusters=15, num_classes=6)
    tb, ts, tl = torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(l)
    np.testing.assert_array_equal(b2.nms(tb.cuda(), ts.cuda(), 0.5).cpu().numpy(), tvb.nms(tb, ts, 0.5).numpy())
    np.testing.assert_array_equal(b2.batched_nms(tb.cuda(), ts.cuda(), tl.cuda(), 0.5, "vanilla").cpu().numpy(),
                                  tvb._batched_nms_vanilla(tb, ts, tl, 0.5).numpy())
    np.testing.assert_array_equal(b2.batched_nms(tb.cuda(), ts.cuda(), tl.cuda(), 0.5).cpu().numpy(),
                                  tvb._batched_nms_coordinate_trick(tb, ts, tl, 0.5).numpy())
    np.testing.assert_array_equal(b2.box_iou(tb[:50].cuda(), tb.cuda()).cpu().numpy(), tvb.box_iou(tb[:50], tb).numpy())
    assert b2.nms(tb[:0].cuda(), ts[:0].cuda(), 0.5).numel() == 0
    np.testing.assert_array_equal(b2.clip_boxes_to_image(tb.cuda(), (300, 400)).cpu().numpy(),
                                  tvb.clip_boxes_to_image(tb, (300, 400)).numpy())
    np.testing.assert_array_equal(b2.remove_small_boxes(tb.cuda(), 20.0).cpu().numpy(),
                                  tvb.remove_small_boxes(tb, 20.0).numpy())


def test_boxcoder_and_matcher():
    from object_detectors_b200.tvision._utils import BoxCoder, Matcher
    obj, deltas, anchors, per_level = syn.rpn_inputs(41, 1, 224, 320)
    d, a = torch.from_numpy(deltas[0]), torch.from_numpy(anchors)
    for w in ((1.0, 1.0, 1.0, 1.0), (10.0, 10.0, 5.0, 5.0)):
        got = BoxCoder(w).decode_single(d.cuda(), a.cuda()).cpu().numpy()
        want = tv_ref.decode_single(d, a, w).numpy()
        assert np.all(np.abs(got - want) <= 1e-5 * np.maximum(np.abs(want), 320))
    got = BoxCoder((1.0, 1.0, 1.0, 1.0)).decode(d.cuda(), [a.cuda()]).cpu().numpy()
    assert got.shape == (a.shape[0], 1, 4)
    # Matcher on a real IoU matrix with exact ties (duplicated predictions)
    gt, _, _ = syn.random_boxes(3, 12, clusters=3)
    pr, _, _ = syn.random_boxes(4, 700, clusters=3)
    pr[100:110] = pr[0:10]
    q = tv_ref.box_iou(torch.from_numpy(gt), torch.from_numpy(pr))
    for allow in (False, True):
        for hi, lo in ((0.7, 0.3), (0.5, 0.5)):
            want = tv_ref.matcher(q.clone(), hi, lo, allow).numpy()
            got = Matcher(hi, lo, allow)(q.cuda()).cpu().numpy()
            np.testing.assert_array_equal(got, want)


def test_rpn_filter_proposals_bound_like_the_reference():
    from object_detectors_b200.tvision import rpn as b200_rpn
    gold = np.load(os.path.join(G, "rpn_filter.npz"))
    for tag in ("s", "m"):
        seed, bsz, ih, iw, pre, post = (int(v) for v in gold[f"{tag}_args"])
        obj, deltas, anchors, per_level = syn.rpn_inputs(seed, bsz, ih, iw)
        a = torch.from_numpy(anchors)
        props = torch.stack([tv_ref.decode_single(torch.from_numpy(deltas[i]), a) for i in range(bsz)])
        fake_self = types.SimpleNamespace(pre_nms_top_n=lambda: pre, post_nms_top_n=lambda: post, nms_thresh=0.7,
                                          score_thresh=0.0, min_size=1e-3)
        # the reference ran on CPU, where torchvision picks the strategy by numel > 4000
        n_kept = sum(min(pre, n) for n in per_level)
        strategy = "vanilla" if n_kept * 4 > 4000 else "coordinate_trick"
        boxes, scores = b200_rpn.filter_proposals(fake_self, props.cuda(), torch.from_numpy(obj).reshape(-1, 1).cuda(),
                                                  [(ih, iw)] * bsz, per_level, strategy=strategy)
        for i in range(bsz):
            np.testing.assert_array_equal(boxes[i].cpu().numpy(), gold[f"{tag}_boxes_{i}"])
            np.testing.assert_allclose(scores[i].cpu().numpy(), gold[f"{tag}_scores_{i}"], rtol=1e-5)


# --------------------------------------------------------------------------- yolo_loss.py (legacy)
@pytest.mark.parametrize("grid,head_idx,classes,img", [(13, 0, 80, 416), (38, 1, 20, 608), (7, 2, 3, 224)])
def test_legacy_yololoss_forward(grid, head_idx, classes, img):
    """YOLOLoss(cfg, head).forward(input): rows ordered (a, h, w), tolerance of the decode contract
    (coordinates 1e-5 * max(|ref|, img_size), probabilities 1e-5 * |ref| + 1e-12)."""
    from object_detectors_b200.yolo.nets.yolo_loss import YOLOLoss
    x = torch.from_numpy(syn.legacy_head(77 + grid, 2, 3, classes, grid))
    cfg = dict(anchors=[[list(a) for a in s] for s in syn.COCO_ANCHORS], classes=classes, img_size=img,
               ignore_threshold=0.5)
    got = YOLOLoss(cfg, head_idx)(x.cuda()).cpu().numpy()
    ref = yolo_ref.legacy_decode(x, syn.COCO_ANCHORS[head_idx], classes, img).numpy()
    assert got.shape == ref.shape == (2, 3 * grid * grid, 5 + classes)
    assert np.all(np.abs(got[..., :4] - ref[..., :4]) <= 1e-5 * np.maximum(np.abs(ref[..., :4]), img))
    assert np.all(np.abs(got[..., 4:] - ref[..., 4:]) <= 1e-5 * np.abs(ref[..., 4:]) + 1e-12)
    if grid == 13:      # and against the reference's own output
        gold = np.load(os.path.join(G, "legacy_yolo_loss.npz"))
        s = gold["h13_sample"]
        g = got[:, gold["h13_rows"]]
        assert np.all(np.abs(g[..., :4] - s[..., :4]) <= 1e-5 * np.maximum(np.abs(s[..., :4]), img))
        assert np.all(np.abs(g[..., 4:] - s[..., 4:]) <= 1e-5 * np.abs(s[..., 4:]) + 1e-12)


def test_legacy_yololoss_get_target():
    from object_detectors_b200.yolo.nets.yolo_loss import YOLOLoss
    cfg = dict(anchors=[[list(a) for a in s] for s in syn.COCO_ANCHORS], classes=80, img_size=416, ignore_threshold=0.5)
    layer = YOLOLoss(cfg, 1)
    targets = [{k: torch.from_numpy(v) for k, v in t.items()} for t in syn.gt_targets(33, 2, 80, max_gt=12)]
    stride = 416 / 26
    scaled = [(a_w / stride, a_h / stride) for a_w, a_h in syn.COCO_ANCHORS[1]]
    mask, noobj, tx, ty, tw, th, tconf, tcls = layer.get_target(targets, scaled, 26, 26, 0.5)
    # restated with torchvision's CPU box_iou exactly as yolo_loss.py:107-161 does
    for b, t in enumerate(targets):
        bbox = t["bbox"]
        gx = torch.clamp(bbox[:, 0] * 26, 0, 26 - 1e-4); gy = torch.clamp(bbox[:, 1] * 26, 0, 26 - 1e-4)
        gt_box = torch.zeros(bbox.shape); gt_box[:, 2] = bbox[:, 2] * 26; gt_box[:, 3] = bbox[:, 3] * 26
        shapes = torch.cat([torch.zeros(3, 2), torch.tensor(scaled, dtype=torch.float32)], 1)
        ious = tv_ref.box_iou(gt_box, shapes)
        best = ious.max(1)[1]
        gi, gj = gx.long(), gy.long()
        assert bool((mask[b, best, gj, gi] == 1).all()) and bool((noobj[b, best, gj, gi] == 0).all())
        assert int(mask[b].sum()) == len(set(zip(best.tolist(), gj.tolist(), gi.tolist())))
        want_noobj = torch.ones(3, 26, 26)
        for i, row in enumerate(ious):
            want_noobj[row > 0.5, gj[i], gi[i]] = 0
        want_noobj[best, gj, gi] = 0
        np.testing.assert_array_equal(noobj[b].cpu().numpy(), want_noobj.numpy())


# --------------------------------------------------------------------------------- roi_heads.py
@pytest.mark.parametrize("tag,loss_name", [("ce", "ce"), ("gombit", "gombit_x"), ("sigmoid", "bce")])
@pytest.mark.parametrize("strategy", ["vanilla", "coordinate_trick", "torchvision"])
def test_roi_postprocess_detections_bound_like_the_reference(tag, loss_name, strategy, nms_path):
    """postprocess_detections bound onto a RoIHeads-like object; expected values: the oracle (pinned to the
    reference by tests/test_oracle_golden.py) with the same batched_nms strategy, and for the installed
    torchvision's own switch (4000 coordinates on CPU) the reference's golden output itself."""
    from object_detectors_b200 import _lib
    from object_detectors_b200.tvision import _utils as det_utils, roi_heads as b200_roi
    gold = np.load(os.path.join(G, "roi_postprocess.npz"))
    seed, r0, r1, c, ih, iw = [int(v) for v in gold["args"]]
    logits, regs, props = syn.roi_inputs(seed, [r0, r1], c, ih, iw)
    idf = torch.from_numpy(gold["idf"])
    head = types.SimpleNamespace(box_coder=det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0)), loss_function_name=loss_name,
                                 tfidf_post=idf.cuda(), score_thresh=0.05, nms_thresh=0.5, detections_per_img=100)
    lib = _lib.load()
    if strategy == "torchvision":
        lib.b200_set_batched_nms_auto_limit(4000)        # the CPU oracle's switch point
    try:
        boxes, scores, labels = b200_roi.postprocess_detections(
            head, torch.from_numpy(logits).cuda(), torch.from_numpy(regs).cuda(), [torch.from_numpy(p).cuda() for p in props],
            [(ih, iw)] * 2, strategy=strategy)
    finally:
        lib.b200_set_batched_nms_auto_limit(100000)
    ref = tv_ref.roi_postprocess(torch.from_numpy(logits), torch.from_numpy(regs), [torch.from_numpy(p) for p in props],
                                 [(ih, iw)] * 2, idf, loss_name, strategy=strategy)
    for i, (rb, rs, rl, _) in enumerate(ref):
        np.testing.assert_array_equal(labels[i].cpu().numpy(), rl.numpy())
        s, b = scores[i].cpu().numpy(), boxes[i].cpu().numpy()
        assert np.all(np.abs(s - rs.numpy()) <= 1e-5 * np.abs(rs.numpy()) + 1e-12)
        assert np.all(np.abs(b - rb.numpy()) <= 1e-5 * np.maximum(np.abs(rb.numpy()), max(ih, iw)))
        if strategy == "torchvision":
            np.testing.assert_array_equal(labels[i].cpu().numpy(), gold[f"{tag}_labels_{i}"])
            assert np.all(np.abs(s - gold[f"{tag}_scores_{i}"]) <= 1e-5 * np.abs(gold[f"{tag}_scores_{i}"]) + 1e-12)


# ------------------------------------------------------------------- differentiable paired IoU (f3)
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("xcycwh", [True, False])
def test_bbox_iou_paired_backward_matches_torch_autograd(kind, xcycwh):
    """helper.bbox_iou on [K,4] x [K,4] under autograd (yolo_forw.py:125): forward bit-equal to the forward-only
    kernel, gradients equal to torch autograd on the reference's expression (oracle.yolo_ref.bbox_iou) within
    1e-4 relative of the largest gradient component of the pair + 2e-5 absolute (fp32, different summation order;
    for identical boxes the O(1..10) partial terms cancel to exactly 0 in one order and to rounding residue in
    another)."""
    from object_detectors_b200.yolo.utilities import helper
    g = np.random.Generator(np.random.PCG64(90 + kind))
    k = 4096
    c1 = g.uniform(0.2, 0.8, (k, 2)); s1 = np.exp(g.uniform(-3.5, -0.7, (k, 2)))
    c2 = c1 + g.normal(0, 0.6, (k, 2)) * s1; s2 = s1 * np.exp(g.normal(0, 0.4, (k, 2)))
    a = np.concatenate([c1, s1], 1).astype(np.float32)
    b = np.concatenate([c2, s2], 1).astype(np.float32)
    if not xcycwh:
        a = np.concatenate([a[:, :2] - a[:, 2:] / 2, a[:, :2] + a[:, 2:] / 2], 1)
        b = np.concatenate([b[:, :2] - b[:, 2:] / 2, b[:, :2] + b[:, 2:] / 2], 1)
    a[:8] = b[:8]                                        # identical boxes: min/max ties
    wgt = g.normal(0, 1, k).astype(np.float32)
    ta, tb = torch.from_numpy(a).requires_grad_(True), torch.from_numpy(b).requires_grad_(True)
    ref = yolo_ref.bbox_iou(ta, tb, kind, xcycwh=xcycwh)
    (ref * torch.from_numpy(wgt)).sum().backward()
    ga, gb = torch.from_numpy(a).cuda().requires_grad_(True), torch.from_numpy(b).cuda().requires_grad_(True)
    out = helper.bbox_iou(ga, gb, kind, xcycwh=xcycwh)
    (out * torch.from_numpy(wgt).cuda()).sum().backward()
    if kind == 3:       # atan: ulp-level libm differences
        np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    else:
        np.testing.assert_array_equal(out.detach().cpu().numpy(), ref.detach().numpy())
    for got, want in ((ga.grad, ta.grad), (gb.grad, tb.grad)):
        got, want = got.cpu().numpy(), want.numpy()
        # CIoU of identical boxes is 0/0 in the reference's alpha (helper.py:273): NaN on both sides
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        scale = np.nanmax(np.abs(np.where(ok, want, 0.0)), axis=1, keepdims=True) + 1e-6
        err = np.where(ok, np.abs(got - want), 0.0)
        assert np.all(err <= 1e-4 * scale + 2e-5), float(err.max())


def test_roi_postprocess_edge_cases():
    """Images without proposals, images where nothing passes the score threshold, and slab overflow."""
    from object_detectors_b200 import ops
    c = 7
    logits, regs, props = syn.roi_inputs(61, [40, 0, 25], c, 480, 640)
    logits[40:] = -20.0                      # third image: every foreground probability ~ 0
    logits[40:, 0] = 20.0
    lg, rg = torch.from_numpy(logits).cuda(), torch.from_numpy(regs).cuda()
    pr = [torch.from_numpy(p).cuda() for p in props]
    shapes = [(480, 640)] * 3
    det, keep, dcnt, ccnt, status = ops.roi_postprocess(lg, rg, pr, shapes, None, ops.ROI_SOFTMAX, nms_mode=ops.NMS_TV_CLASS)
    ref = tv_ref.roi_postprocess(torch.from_numpy(logits), torch.from_numpy(regs), [torch.from_numpy(p) for p in props],
                                 shapes, torch.ones(c), "ce", strategy="vanilla")
    assert int(status.item()) & 1 == 0
    assert dcnt.tolist()[1:] == [0, 0] and ccnt.tolist()[1:] == [0, 0]
    assert int(dcnt[0]) == ref[0][0].shape[0] and ref[1][0].shape[0] == 0 and ref[2][0].shape[0] == 0
    np.testing.assert_array_equal(det[0, :int(dcnt[0]), 5].cpu().numpy(), ref[0][2].numpy().astype(np.float32))
    np.testing.assert_array_equal(keep[0, :int(dcnt[0])].cpu().numpy(), ref[0][3].numpy().astype(np.int32))
    # a slab that is too small must say so (status bit 0), never drop candidates silently
    _, _, _, ccnt2, status2 = ops.roi_postprocess(lg, rg, pr, shapes, None, ops.ROI_SOFTMAX, nms_mode=ops.NMS_TV_CLASS,
                                                  capacity=max(1, int(ccnt[0]) // 2))
    assert int(status2.item()) & 1 == 1 and int(ccnt2[0]) == int(ccnt[0])


def test_boxcoder_encode_matches_oracle():
    """BoxCoder.encode / encode_single (tvision/_utils.py:80-125): dx, dy bit-exact (IEEE mul/sub/div in the
    reference order), dw, dh within 1e-5 relative + 1e-6 (logf vs the CPU log)."""
    from object_detectors_b200.tvision import _utils as det_utils
    g = np.random.Generator(np.random.PCG64(8))
    n = 3000
    c = g.uniform(50, 500, (n, 2)); s = np.exp(g.uniform(2, 5, (n, 2)))
    prop = torch.from_numpy(np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32))
    c2 = c + g.normal(0, 5, (n, 2)); s2 = s * np.exp(g.normal(0, 0.2, (n, 2)))
    ref = torch.from_numpy(np.concatenate([c2 - s2 / 2, c2 + s2 / 2], 1).astype(np.float32))
    coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
    got = coder.encode_single(ref.cuda(), prop.cuda()).cpu().numpy()
    want = tv_ref.encode_boxes(ref, prop, (10.0, 10.0, 5.0, 5.0)).numpy()
    np.testing.assert_array_equal(got[:, :2], want[:, :2])
    np.testing.assert_allclose(got[:, 2:], want[:, 2:], rtol=1e-5, atol=1e-6)
    parts = coder.encode([ref[:1000].cuda(), ref[1000:].cuda()], [prop[:1000].cuda(), prop[1000:].cuda()])
    assert [p.shape[0] for p in parts] == [1000, 2000]
    np.testing.assert_array_equal(torch.cat(parts).cpu().numpy(), got)


def test_emit_results_kernel_bit_exact():
    """b200_emit_results against the reference's arithmetic (test_one_epoch.py:41-66) in numpy fp32, operation by
    operation: x / inp_dim * size, w = x2 - x1, area = w * h; COCO 80 -> 91 ids; empty images shift the owner."""
    from object_detectors_b200 import ops
    g = np.random.default_rng(9)
    b, max_det, inp = 7, 33, np.float32(608)
    det = (g.random((b, max_det, 6)) * 600).astype(np.float32)
    det[..., 5] = g.integers(0, 80, size=(b, max_det)).astype(np.float32)
    cnt = g.integers(0, max_det + 1, size=(b,)).astype(np.int32)
    cnt[2] = 0; cnt[5] = 0; cnt[0] = max_det
    hw = np.stack([g.integers(300, 900, size=b), g.integers(300, 900, size=b)], 1).astype(np.float32)
    ids = (np.arange(b) * 7 + 100000000000).astype(np.int64)
    coco91 = np.array([c for c in range(1, 91) if c not in (12, 26, 29, 30, 45, 66, 68, 69, 71, 83)], np.int32)
    for strict in (True, False):
        for cmap in (coco91, None):
            rec, cat, img, total = ops.emit_results(torch.from_numpy(det).cuda(), torch.from_numpy(cnt).cuda(),
                                                    torch.from_numpy(hw).cuda(), torch.from_numpy(ids).cuda(), float(inp),
                                                    None if cmap is None else torch.from_numpy(cmap).cuda(), strict)
            n = int(total.item())
            assert n == int(cnt.sum())
            want_rec, want_cat, want_img = [], [], []
            owner = 0
            for i in range(b):
                if cnt[i] == 0:
                    continue
                o = owner if strict else i
                owner += 1
                d = det[i, :cnt[i]]
                x1 = d[:, 0] / inp * hw[o, 1]; y1 = d[:, 1] / inp * hw[o, 0]
                x2 = d[:, 2] / inp * hw[o, 1]; y2 = d[:, 3] / inp * hw[o, 0]
                w, h = x2 - x1, y2 - y1
                want_rec.append(np.stack([x1, y1, w, h, w * h, d[:, 4]], 1))
                lab = d[:, 5].astype(np.int64)
                want_cat.append(cmap[lab] if cmap is not None else (lab + 1).astype(np.int32))
                want_img.append(np.full(cnt[i], ids[o]))
            np.testing.assert_array_equal(rec[:n].cpu().numpy(), np.concatenate(want_rec).astype(np.float32))
            np.testing.assert_array_equal(cat[:n].cpu().numpy(), np.concatenate(want_cat))
            np.testing.assert_array_equal(img[:n].cpu().numpy(), np.concatenate(want_img))


def test_yoloforw_postprocess_retries_with_the_true_candidate_count():
    """A slab that is too small is reported, never silent, and the drop-in repeats the call with what the batch needs
    (the kernels return the true per-image candidate counts) -- same detections as a roomy call; plans are cached."""
    from object_detectors_b200.yolo.nets.yolo_forw import YOLOForw
    cfg = _cfg(416, 80, syn.COCO_ANCHORS, "coco", 1, [0, 0])
    yolo = YOLOForw(cfg, idf_logits=_idf("coco"))
    heads = [torch.from_numpy(h).cuda() for h in syn.yolo_heads(3, 2, 416, 80, syn.COCO_ANCHORS, "clustered")]
    roomy = [t.clone() for t in yolo.postprocess(heads, capacity=4096)]
    tight = yolo.postprocess(heads, capacity=32)
    assert int(roomy[3].max()) > 32
    k = roomy[3]
    np.testing.assert_array_equal(tight[3].cpu().numpy(), k.cpu().numpy())
    for i in range(2):
        n = int(k[i])
        np.testing.assert_array_equal(tight[0][i, :n].cpu().numpy(), roomy[0][i, :n].cpu().numpy())
        np.testing.assert_array_equal(tight[1][i, :n].cpu().numpy(), roomy[1][i, :n].cpu().numpy())
    n_plans = len(yolo._plans)
    yolo.postprocess(heads, capacity=4096)
    assert len(yolo._plans) == n_plans                  # served from the cache


@pytest.mark.parametrize("shape,bsz,pre", [((224, 320), 3, 300), ((800, 1344), 2, 2000), ((800, 1344), 2, 1000)])
def test_rpn_get_top_n_idx(shape, bsz, pre):
    """RegionProposalNetwork._get_top_n_idx (rpn.py:215-228) bound like the reference: the index list of split /
    topk / cat per level, bit-exact (tie-free logits)."""
    from object_detectors_b200.tvision import rpn as b200_rpn
    ih, iw = shape
    obj, _, _, per_level = syn.rpn_inputs(43, bsz, ih, iw)
    fake_self = types.SimpleNamespace(pre_nms_top_n=lambda: pre)
    got = b200_rpn._get_top_n_idx(fake_self, torch.from_numpy(obj).cuda(), per_level).cpu()
    want, off = [], 0
    for ob in torch.from_numpy(obj).split(per_level, 1):                      # the reference's loop, on the CPU
        k = min(pre, ob.shape[1])
        want.append(ob.topk(k, dim=1)[1] + off)
        off += ob.shape[1]
    want = torch.cat(want, 1)
    assert got.dtype == torch.int64 and got.shape == want.shape
    np.testing.assert_array_equal(got.numpy(), want.numpy())


def test_rpn_top_n_idx_with_ties():
    """Equal logits around the k-th value: the selection is a valid top-k (multiset of values equal to topk's) and,
    among equal logits, the lower index comes first."""
    from object_detectors_b200 import ops
    g = np.random.default_rng(3)
    per_level = [5000, 700]
    obj = np.round(g.standard_normal((2, 5700)) * 4).astype(np.float32) / 4 + np.float32(0.0)     # heavy ties, no -0.0
    got = ops.rpn_top_n_idx(torch.from_numpy(obj).cuda(), per_level, 600).cpu().numpy()
    off = 0
    col = 0
    for n in per_level:
        k = min(600, n)
        for b in range(2):
            idx = got[b, col:col + k] - off
            vals = obj[b, off:off + n][idx]
            want_vals = np.sort(obj[b, off:off + n])[::-1][:k]
            np.testing.assert_array_equal(vals, want_vals)
            assert len(set(idx.tolist())) == k
            same = vals[1:] == vals[:-1]
            assert np.all(idx[1:][same] > idx[:-1][same])
        off += n
        col += k


@pytest.mark.parametrize("m,n,seed", [(12, 700, 3), (1, 257, 5), (50, 268569, 7), (300, 5000, 9)])
def test_fused_box_iou_matcher(m, n, seed):
    """b200_match_boxes == Matcher(...)(box_iou(gt, boxes)) of the reference (torchvision's CPU box_iou + the oracle's
    restatement of _utils.py:271-361), bit for bit, without the [M, N] matrix: both threshold settings, the
    low-quality restore incl. exact ties (duplicated boxes), a ground truth that meets no box, and SSDMatcher."""
    from object_detectors_b200.tvision._utils import Matcher, SSDMatcher
    if n > 100000:
        _, _, anchors, _ = syn.rpn_inputs(41, 1, 800, 1344)
        pr = anchors
        gt, _, _ = syn.random_boxes(seed, m, extent=1300.0, clusters=6)
    else:
        gt, _, _ = syn.random_boxes(seed, m, clusters=3)
        pr, _, _ = syn.random_boxes(seed + 1, n, clusters=3)
        pr[100:110] = pr[0:10]                                     # exact ties
    gt = gt.copy()
    gt[-1] = [5000.0, 5000.0, 5040.0, 5030.0] if m > 1 else gt[-1]  # intersects nothing: an all-zero row
    tg, tp = torch.from_numpy(gt), torch.from_numpy(pr)
    q = tv_ref.box_iou(tg, tp)
    for allow in (False, True):
        for hi, lo in ((0.7, 0.3), (0.5, 0.5)):
            want = tv_ref.matcher(q.clone(), hi, lo, allow).numpy()
            got = Matcher(hi, lo, allow).match_boxes(tg.cuda(), tp.cuda()).cpu().numpy()
            np.testing.assert_array_equal(got, want, err_msg=f"allow={allow} thr=({hi},{lo})")
    want = tv_ref.ssd_matcher(q.clone(), 0.5).numpy()
    np.testing.assert_array_equal(SSDMatcher(0.5).match_boxes(tg.cuda(), tp.cuda()).cpu().numpy(), want)
    if n <= 5000:
        np.testing.assert_array_equal(SSDMatcher(0.5)(q.cuda()).cpu().numpy(), want)
    from object_detectors_b200 import ops
    _, vals = ops.match_boxes(tg.cuda(), tp.cuda(), 0.7, 0.3, return_vals=True)
    np.testing.assert_array_equal(vals.cpu().numpy(), q.max(dim=0)[0].numpy())


@pytest.mark.parametrize("strategy", ["vanilla", "coordinate_trick", "torchvision"])
def test_retinanet_postprocess_detections_bound_like_the_reference(strategy, nms_path):
    """RetinaNet.postprocess_detections (retinanet.py:414-472) bound onto a RetinaNet-like object, against the oracle
    (pinned bit-exactly to the reference by tests/test_oracle_golden.py) and -- for the installed torchvision's own
    strategy switch -- against the reference's golden output itself: labels and counts exact, boxes
    |d| <= 1e-5 * max(|ref|, img), scores 1e-5 relative."""
    from object_detectors_b200 import _lib
    from object_detectors_b200.tvision import retinanet as b200_retina
    gold = np.load(os.path.join(G, "retinanet_postprocess.npz"))
    seed, bsz, ih, iw, c = [int(v) for v in gold["args"]]
    logits, regs, anchors, per_level = syn.retina_inputs(seed, bsz, ih, iw, c)
    idf = torch.from_numpy(gold["idf"])
    tl, tr, ta = torch.from_numpy(logits), torch.from_numpy(regs), torch.from_numpy(anchors)
    head = {"cls_logits": [t.cuda() for t in tl.split(per_level, 1)], "bbox_regression": [t.cuda() for t in tr.split(per_level, 1)]}
    fake_self = types.SimpleNamespace(tfidf_post=idf.cuda(), score_thresh=0.05, topk_candidates=1000, nms_thresh=0.5,
                                      detections_per_img=300)
    lib = _lib.load()
    if strategy == "torchvision":
        lib.b200_set_batched_nms_auto_limit(4000)        # the CPU oracle's switch point
    try:
        got = b200_retina.postprocess_detections(fake_self, head, [[a.cuda() for a in ta.split(per_level, 0)]] * bsz,
                                                 [(ih, iw)] * bsz, strategy=strategy)
    finally:
        lib.b200_set_batched_nms_auto_limit(100000)
    ref = tv_ref.retinanet_postprocess(tl, tr, ta, per_level, [(ih, iw)] * bsz, idf, strategy=strategy)
    for i, (rb, rs, rl) in enumerate(ref):
        assert got[i]["boxes"].shape[0] == rb.shape[0] > 0
        np.testing.assert_array_equal(got[i]["labels"].cpu().numpy(), rl.numpy())
        s, b = got[i]["scores"].cpu().numpy(), got[i]["boxes"].cpu().numpy()
        assert np.all(np.abs(s - rs.numpy()) <= 1e-5 * np.abs(rs.numpy()) + 1e-12)
        assert np.all(np.abs(b - rb.numpy()) <= 1e-5 * np.maximum(np.abs(rb.numpy()), max(ih, iw)))
        if strategy == "torchvision":
            np.testing.assert_array_equal(got[i]["labels"].cpu().numpy(), gold[f"labels_{i}"])
            assert np.all(np.abs(b - gold[f"boxes_{i}"]) <= 1e-5 * np.maximum(np.abs(gold[f"boxes_{i}"]), max(ih, iw)))


@pytest.mark.parametrize("strategy", ["vanilla", "torchvision"])
def test_ssd_postprocess_detections_bound_like_the_reference(strategy):
    """SSD.postprocess_detections (ssd.py:386-430) bound onto an SSD-like object: softmax of tfidf * logits, one decoded box
    per anchor, per-class threshold + top-400 (one class has 571 candidates), class-aware NMS, top 200 -- against the
    oracle (pinned bit-exactly to the reference) and the reference's golden output; the first image's 17 563 candidates
    overflow the default slab and exercise the exact-capacity retry.  Labels and counts exact, values in tolerance."""
    from object_detectors_b200 import _lib
    from object_detectors_b200.tvision import ssd as b200_ssd
    from object_detectors_b200.tvision._utils import BoxCoder
    gold = np.load(os.path.join(G, "ssd_postprocess.npz"))
    seed, bsz, a, c, img = [int(v) for v in gold["args"]]
    logits, regs, anchors = syn.ssd_inputs(seed, bsz, a, c, img)
    idf = torch.from_numpy(gold["idf"])
    fake_self = types.SimpleNamespace(tfidf_post=idf.cuda(), box_coder=BoxCoder((10.0, 10.0, 5.0, 5.0)), score_thresh=0.01,
                                      topk_candidates=400, nms_thresh=0.45, detections_per_img=200)
    head = {"cls_logits": torch.from_numpy(logits).cuda(), "bbox_regression": torch.from_numpy(regs).cuda()}
    lib = _lib.load()
    if strategy == "torchvision":
        lib.b200_set_batched_nms_auto_limit(4000)        # the CPU oracle's switch point
    try:
        got = b200_ssd.postprocess_detections(fake_self, head, [torch.from_numpy(anchors).cuda()] * bsz, [(img, img)] * bsz,
                                              strategy=strategy)
    finally:
        lib.b200_set_batched_nms_auto_limit(100000)
    ref = tv_ref.ssd_postprocess(torch.from_numpy(logits), torch.from_numpy(regs), [torch.from_numpy(anchors)] * bsz,
                                 [(img, img)] * bsz, idf, strategy=strategy)
    for i, (rb, rs, rl) in enumerate(ref):
        assert got[i]["boxes"].shape[0] == rb.shape[0] > 0
        np.testing.assert_array_equal(got[i]["labels"].cpu().numpy(), rl.numpy())
        s, b = got[i]["scores"].cpu().numpy(), got[i]["boxes"].cpu().numpy()
        assert np.all(np.abs(s - rs.numpy()) <= 1e-5 * np.abs(rs.numpy()) + 1e-12)
        assert np.all(np.abs(b - rb.numpy()) <= 1e-5 * np.maximum(np.abs(rb.numpy()), img))
        if strategy == "torchvision":
            np.testing.assert_array_equal(got[i]["labels"].cpu().numpy(), gold[f"labels_{i}"])


@pytest.mark.parametrize("case", ["overflow", "underflow"])
def test_rpn_top_n_idx_when_the_sample_misjudges_the_level(case):
    """Large levels estimate their threshold from 4096 evenly spaced logits.  Adversarial inputs: the largest logits all
    sit on unsampled positions and outnumber the candidate list (overflow), or ONLY the sampled positions are large
    (underflow: far too few candidates).  Both must fall back to the exhaustive select and still return topk's list."""
    from object_detectors_b200 import ops
    g = np.random.default_rng(17)
    n, k = 201600, 2000
    obj = (g.standard_normal((2, n)) * 2 - 3).astype(np.float32)
    sampled = np.unique((np.arange(4096, dtype=np.int64) * n) // 4096)
    if case == "overflow":
        free = np.setdiff1d(np.arange(n), sampled)
        hot = g.choice(free, size=20000, replace=False)
        for b in range(2):                                                    # tie-free: a shuffled ramp
            obj[b, hot] = (10.0 + g.permutation(20000) * 2e-4).astype(np.float32)
    else:
        for b in range(2):
            obj[b, sampled] = (20.0 + g.permutation(sampled.size) * 2e-3).astype(np.float32)
    got = ops.rpn_top_n_idx(torch.from_numpy(obj).cuda(), [n], k).cpu()
    want = torch.from_numpy(obj).topk(k, dim=1)[1]
    np.testing.assert_array_equal(got.numpy(), want.numpy())


@pytest.mark.gpu
def test_clip_boxes_and_remove_small_boxes_match_torch():
    """tvision/boxes.py clip_boxes_to_image / remove_small_boxes as kernels: same values / indices as the torch
    expressions of the reference (x clamped to [0, w], y to [0, h]; w >= min and h >= min), incl. NaN rows, leading
    batch dimensions, empty inputs and exact-threshold sizes."""
    from object_detectors_b200.tvision import boxes as bx
    g = np.random.default_rng(77)
    b = (g.standard_normal((3, 1777, 4)) * 400 + 300).astype(np.float32)
    b[1, 5] = np.nan
    t = torch.from_numpy(b).cuda()
    got = bx.clip_boxes_to_image(t, (480, 640)).cpu()
    x = torch.from_numpy(b)[..., 0::2].clamp(min=0, max=640)
    y = torch.from_numpy(b)[..., 1::2].clamp(min=0, max=480)
    want = torch.stack((x, y), dim=3).reshape(b.shape)
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(want, nan=-7.0))
    assert bx.clip_boxes_to_image(torch.empty((0, 4), device="cuda"), (4, 4)).shape == (0, 4)

    f = b[0].copy()
    f[:, 2:] = f[:, :2] + np.abs(g.standard_normal((f.shape[0], 2)).astype(np.float32)) * 2
    f[10, 2] = f[10, 0] + np.float32(1.0)             # exactly the threshold (when the subtraction is exact)
    f[11] = np.nan
    tf = torch.from_numpy(f).cuda()
    for ms in (1e-3, 1.0, 2.5):
        gotk = bx.remove_small_boxes(tf, ms).cpu()
        ws, hs = torch.from_numpy(f)[:, 2] - torch.from_numpy(f)[:, 0], torch.from_numpy(f)[:, 3] - torch.from_numpy(f)[:, 1]
        wantk = torch.where((ws >= ms) & (hs >= ms))[0]
        assert torch.equal(gotk, wantk)
    assert bx.remove_small_boxes(torch.empty((0, 4), device="cuda"), 1.0).numel() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("per_level,k", [
    ([50400], 7000),              # multi-slice level that does not qualify for sampling: local top-k per slice + merge
    ([30000, 9000], 8192),        # two slices with k close to the slice length; a single-slice level with k < n
    ([70000, 200, 33], 3000),     # sampled level beside levels smaller than k (taken whole) and smaller than a warp
    ([20480, 20481], 512),        # the single-slice limit and one logit more
    ([300000], 8192),             # sampled level with the largest k (candidate list near its capacity)
])
def test_rpn_top_n_idx_select_plans(per_level, k):
    """Every shared-memory plan of the per-level select (single slice, slices + merge, sampled threshold, bucket sort
    and its fallbacks) returns torch.topk's index list on tie-free logits."""
    from object_detectors_b200 import ops
    g = np.random.default_rng(k + len(per_level))
    total = sum(per_level)
    obj = np.empty((3, total), dtype=np.float32)
    for b in range(3):                     # tie-free by construction: a shuffled ramp with Gaussian-like spacing
        ramp = np.sort(g.standard_normal(total) * 2 - 3).astype(np.float64) + np.arange(total) * 1e-4
        obj[b] = g.permutation(ramp).astype(np.float32)
        assert np.unique(obj[b]).size == total
    got = ops.rpn_top_n_idx(torch.from_numpy(obj).cuda(), per_level, k).cpu().numpy()
    off = col = 0
    for n in per_level:
        kk = min(k, n)
        want = torch.from_numpy(obj[:, off:off + n]).topk(kk, dim=1)[1].numpy() + off
        np.testing.assert_array_equal(got[:, col:col + kk], want)
        off += n
        col += kk
