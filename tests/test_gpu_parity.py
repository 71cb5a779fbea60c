"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Contract, tolerances written here:
  * integer / index results (candidate anchor sets, labels, counts, NMS keep indices, argmax
    matches, masks, top-k index sets): BIT-EXACT.
  * decoded coordinates: |d| <= 1e-5 * max(|ref|, img_size);  scores / probabilities:
    |d| <= 1e-5 * |ref| + 1e-12  (transcendentals: libdevice vs Sleef differ by ulps).
  * NMS given identical inputs: bit-exact (the IoU arithmetic is reproduced operation by
    operation).  End-to-end keep sets are additionally compared with the golden fixtures on
    margin-screened seeds.
"""
import os

import numpy as np
import pytest
import torch

from object_detectors_b200 import synthetic as syn
from oracle import cref, tv_ref, yolo_ref

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5


def _ops():
    from object_detectors_b200 import ops
    return ops


def _idf(name):
    return torch.from_numpy(np.load(os.path.join(G, f"idf_{name}_smooth.npy")))


def _close_coord(got, ref, img):
    tol = RTOL * np.maximum(np.abs(ref), img)
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{bad.sum()} coordinates off, worst {np.abs(got - ref).max()}"


def _close_score(got, ref):
    bad = np.abs(got - ref) > RTOL * np.abs(ref) + 1e-12
    assert not bad.any(), f"{bad.sum()} scores off, worst rel {(np.abs(got - ref) / np.abs(ref)).max()}"


def _gpu_heads(heads):
    return [torch.from_numpy(h).cuda() for h in heads]


CASES = {
    # name: (seed, B, img, C, anchors, idf name, softmax, sigmoid_cls generator, max_objects)
    "tiny_softmax_idf": (11, 2, 64, 80, syn.COCO_ANCHORS, "coco", True, False, 3),
    "tiny_sigmoid": (11, 2, 64, 80, syn.COCO_ANCHORS, None, False, True, 3),
    "c1_416_idf": (3, 1, 416, 80, syn.COCO_ANCHORS, "coco", True, False, 15),
    "c1_416_plain": (110, 1, 416, 80, syn.COCO_ANCHORS, None, True, False, 15),
    "c2_608_b4": (203, 4, 608, 80, syn.COCO_ANCHORS, "coco", True, False, 15),
    "lvis_96_a6": (6, 2, 96, 1203, syn.LVIS_ANCHORS, "lvis", True, False, 4),
    "odd_grid_352": (17, 3, 352, 80, syn.COCO_ANCHORS, "coco", True, False, 8),   # grids 11,22,44
    "one_class": (19, 2, 128, 1, syn.COCO_ANCHORS, None, False, True, 4),
}


def _case(name):
    seed, b, img, c, anchors, idfn, softmax, sig, mo = CASES[name]
    heads = syn.yolo_heads(seed, b, img, c, anchors, "clustered", sigmoid_cls=sig, max_objects=mo)
    idf = _idf(idfn) if idfn else None
    return heads, b, img, c, anchors, idf, softmax


@pytest.mark.parametrize("name", list(CASES))
def test_decode_filter_candidates(name):
    """fused decode+filter == oracle decode -> get_abs_coord -> score -> mask -> gather."""
    ops = _ops()
    heads, b, img, c, anchors, idf, softmax = _case(name)
    ref = yolo_ref.score_filter(yolo_ref.decode([torch.from_numpy(h) for h in heads], anchors, img, c, idf, softmax), 0.1)
    out = ops.yolo_decode_filter(_gpu_heads(heads), anchors, img, c, None if idf is None else idf.cuda(), softmax, 0.1)
    assert int(out["status"].item()) == 0
    cnt = out["count"].cpu().numpy()
    total = 0
    for i in range(b):
        n = int(cnt[i])
        assert n == ref[i]["det6"].shape[0], f"image {i}: {n} candidates vs {ref[i]['det6'].shape[0]}"
        np.testing.assert_array_equal(out["anchor"][i, :n].cpu().numpy(), ref[i]["anchor"].numpy().astype(np.int32))
        d = ref[i]["det6"].numpy()
        np.testing.assert_array_equal(out["label"][i, :n].cpu().numpy(), d[:, 5].astype(np.int32))
        _close_coord(out["box"][i, :n].cpu().numpy(), d[:, :4], img)
        _close_score(out["score"][i, :n].cpu().numpy(), d[:, 4])
        total += n
    assert total > 0


@pytest.mark.parametrize("name", ["tiny_softmax_idf", "tiny_sigmoid", "c1_416_idf", "lvis_96_a6", "odd_grid_352", "one_class"])
def test_decode_dense(name):
    """YOLOForw.forward drop-in tensor [B,N,5+C] within tolerance of the oracle."""
    ops = _ops()
    heads, b, img, c, anchors, idf, softmax = _case(name)
    ref = yolo_ref.decode([torch.from_numpy(h) for h in heads], anchors, img, c, idf, softmax).numpy()
    got = ops.yolo_decode_dense(_gpu_heads(heads), anchors, img, c, None if idf is None else idf.cuda(), softmax).cpu().numpy()
    assert got.shape == ref.shape
    _close_coord(got[..., :4], ref[..., :4], img)
    _close_score(got[..., 4:], ref[..., 4:])


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("mode", ["majority", "tv", "tv_class"])
def test_postprocess_two_stage(name, mode, nms_path):
    """decode+filter+NMS in one call.  Stage 2 is checked bit-exactly by running the C oracle on
    the GPU's own candidate list (identical inputs -> identical keep / labels)."""
    ops = _ops()
    heads, b, img, c, anchors, idf, softmax = _case(name)
    gh = _gpu_heads(heads)
    gi = None if idf is None else idf.cuda()
    cand = ops.yolo_decode_filter(gh, anchors, img, c, gi, softmax, 0.1)
    m = {"majority": ops.NMS_MAJORITY, "tv": ops.NMS_TV, "tv_class": ops.NMS_TV_CLASS}[mode]
    det, keep, anchor, dcnt, ccnt = ops.yolo_postprocess(gh, anchors, img, c, gi, softmax, 0.1, 0.6, m)
    np.testing.assert_array_equal(ccnt.cpu().numpy(), cand["count"].cpu().numpy())
    for i in range(b):
        n = int(cand["count"][i])
        box = cand["box"][i, :n].cpu().numpy()
        sc = cand["score"][i, :n].cpu().numpy()
        lab = cand["label"][i, :n].cpu().numpy()
        if mode == "majority":
            det6 = np.concatenate([box, sc[:, None], lab[:, None].astype(np.float32)], 1)
            ki, kl = cref.nms_majority(det6, 0.6, c)
        else:
            ki = cref.nms_tv(box, sc, 0.6, lab.astype(np.int64) if mode == "tv_class" else None).astype(np.int32)
            kl = lab[ki]
        k = int(dcnt[i])
        assert k == len(ki)
        np.testing.assert_array_equal(keep[i, :k].cpu().numpy(), ki)
        d = det[i, :k].cpu().numpy()
        np.testing.assert_array_equal(d[:, :4], box[ki])
        np.testing.assert_array_equal(d[:, 4], sc[ki])
        np.testing.assert_array_equal(d[:, 5], kl.astype(np.float32))
        np.testing.assert_array_equal(anchor[i, :k].cpu().numpy(), cand["anchor"][i, :n].cpu().numpy()[ki])


@pytest.mark.parametrize("fixture,name", [("decode_tiny_softmax_idf", "tiny_softmax_idf"),
                                          ("decode_tiny_sigmoid_plain", "tiny_sigmoid"),
                                          ("c1_416_idf_seed3", "c1_416_idf"),
                                          ("c1_416_plain_seed110", "c1_416_plain"),
                                          ("c2_608_b4_seed203", "c2_608_b4"),
                                          ("c3_lvis_96_b2_seed6", "lvis_96_a6")])
def test_postprocess_end_to_end_vs_reference_golden(fixture, name):
    """Whole path against what the UNMODIFIED reference produced (tests/golden): candidate and
    kept counts, kept candidate indices and relabelled classes bit-exact, values in tolerance."""
    ops = _ops()
    gold = np.load(os.path.join(G, fixture + ".npz"))
    heads, b, img, c, anchors, idf, softmax = _case(name)
    det, keep, anchor, dcnt, ccnt = ops.yolo_postprocess(_gpu_heads(heads), anchors, img, c,
                                                         None if idf is None else idf.cuda(), softmax, 0.1, 0.6,
                                                         ops.NMS_MAJORITY)
    for i in range(b):
        ganchor, gkept = gold[f"anchor_{i}"], gold[f"kept_{i}"]
        assert int(ccnt[i]) == len(ganchor)
        k = int(dcnt[i])
        assert k == gkept.shape[0]
        if k == 0:
            continue
        d = det[i, :k].cpu().numpy()
        np.testing.assert_array_equal(anchor[i, :k].cpu().numpy(), ganchor[keep[i, :k].cpu().numpy()])
        # the reference's kept rows identify their candidate through the (tie-free) score
        gdet6 = gold[f"det6_{i}"]
        gkeep = np.array([int(np.nonzero(gdet6[:, 4] == s)[0][0]) for s in gkept[:, 4]])
        np.testing.assert_array_equal(keep[i, :k].cpu().numpy(), gkeep)
        np.testing.assert_array_equal(d[:, 5], gkept[:, 5])
        _close_coord(d[:, :4], gkept[:, :4], img)
        _close_score(d[:, 4], gkept[:, 4])


# ------------------------------------------------------------------------------------------ NMS
def _segments(sizes, seed, clusters, num_classes):
    bs, ss, ls, off = [], [], [], [0]
    for j, n in enumerate(sizes):
        b, s, l = syn.random_boxes(seed + j, n, clusters=clusters if n > 4 else 0, num_classes=num_classes)
        bs.append(b); ss.append(s); ls.append(l); off.append(off[-1] + n)
    return (np.concatenate(bs) if bs else np.zeros((0, 4), np.float32), np.concatenate(ss), np.concatenate(ls),
            np.array(off, np.int32), bs, ss, ls)


@pytest.mark.parametrize("sizes", [[300, 0, 1, 64, 65, 1000, 2, 129], [4096], [5000]])
def test_nms_modes_bit_exact(sizes, nms_path):
    ops = _ops()
    boxes, scores, labels, off, bs, ss, ls = _segments(sizes, 50, 12, 7)
    tb, ts = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    tl, to = torch.from_numpy(labels).cuda(), torch.from_numpy(off).cuda()
    for mode, thr in ((ops.NMS_MAJORITY, 0.6), (ops.NMS_TV, 0.5), (ops.NMS_TV_CLASS, 0.5), (ops.NMS_TV_TRICK, 0.5),
                      (ops.NMS_TV, 0.7)):
        keep, kc, lout = ops.nms_segments(tb, ts, tl, to, thr, mode)
        keep, kc = keep.cpu().numpy(), kc.cpu().numpy()
        for s, n in enumerate(sizes):
            got = keep[off[s]:off[s] + kc[s]]
            if mode == ops.NMS_MAJORITY:
                det6 = np.concatenate([bs[s], ss[s][:, None], ls[s][:, None].astype(np.float32)], 1)
                want, wl = cref.nms_majority(det6, thr, 7)
                np.testing.assert_array_equal(lout.cpu().numpy()[off[s]:off[s] + kc[s]], wl)
            elif mode == ops.NMS_TV:
                want = cref.nms_tv(bs[s], ss[s], thr)
            elif mode == ops.NMS_TV_CLASS:
                want = cref.nms_tv(bs[s], ss[s], thr, ls[s])
            else:
                want = tv_ref.batched_nms_coordinate_trick(torch.from_numpy(bs[s]), torch.from_numpy(ss[s]),
                                                           torch.from_numpy(ls[s]), thr).numpy()
            np.testing.assert_array_equal(got, want, err_msg=f"mode {mode} segment {s} (n={n})")


@pytest.mark.parametrize("quantum", [0.05, 1e-3])
def test_nms_with_tied_scores(quantum, nms_path):
    """Scores quantised to a few (0.05) or many (1e-3) levels: runs of equal scores longer than the bucket sort of the
    single-launch kernel ranks by counting (-> its sorting-network fallback) and short runs inside buckets; equal
    scores keep their index order (the oracle's canonical order)."""
    ops = _ops()
    sizes = [1500, 2000, 3000, 257]
    boxes, scores, labels, off, bs, ss, ls = _segments(sizes, 91, 20, 5)
    scores = (np.round(scores / quantum) * quantum).astype(np.float32)
    ss = [scores[off[i]:off[i + 1]] for i in range(len(sizes))]
    tb, ts = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    tl, to = torch.from_numpy(labels).cuda(), torch.from_numpy(off).cuda()
    for mode, thr in ((ops.NMS_TV, 0.5), (ops.NMS_MAJORITY, 0.6), (ops.NMS_TV_CLASS, 0.5)):
        keep, kc, lout = ops.nms_segments(tb, ts, tl, to, thr, mode)
        keep, kc = keep.cpu().numpy(), kc.cpu().numpy()
        for s_, n in enumerate(sizes):
            got = keep[off[s_]:off[s_] + kc[s_]]
            if mode == ops.NMS_MAJORITY:
                det6 = np.concatenate([bs[s_], ss[s_][:, None], ls[s_][:, None].astype(np.float32)], 1)
                want, wl = cref.nms_majority(det6, thr, 5)
                np.testing.assert_array_equal(lout.cpu().numpy()[off[s_]:off[s_] + kc[s_]], wl)
            else:
                want = cref.nms_tv(bs[s_], ss[s_], thr, ls[s_] if mode == ops.NMS_TV_CLASS else None)
            np.testing.assert_array_equal(got, want, err_msg=f"mode {mode} segment {s_} (n={n})")


def test_nms_threshold_semantics(nms_path):
    """IoU == fp32(0.6) exactly: torchvision suppresses ((double)0.6f > 0.6), nms_majority removes it
    without a vote; zero-area pairs: NaN IoU never suppresses in torchvision, is removed in majority."""
    ops = _ops()
    b = torch.tensor([[0.0, 0.0, 4.0, 4.0], [0.0, 0.0, 4.0, 2.4], [9.0, 9.0, 9.0, 9.0], [9.0, 9.0, 9.0, 9.0]]).cuda()
    s = torch.tensor([0.9, 0.8, 0.7, 0.6]).cuda()
    l = torch.tensor([0, 1, 2, 3], dtype=torch.int32).cuda()
    off = torch.tensor([0, 4], dtype=torch.int32).cuda()
    keep, kc, _ = ops.nms_segments(b, s, l, off, 0.6, ops.NMS_TV)
    assert keep[:int(kc[0])].tolist() == cref.nms_tv(b.cpu().numpy(), s.cpu().numpy(), 0.6).tolist() == [0, 2, 3]
    keep, kc, lout = ops.nms_segments(b, s, l, off, 0.6, ops.NMS_MAJORITY)
    det6 = torch.cat([b.cpu(), s.cpu()[:, None], l.cpu()[:, None].float()], 1).numpy()
    ki, kl = cref.nms_majority(det6, 0.6, 4)
    assert keep[:int(kc[0])].tolist() == ki.tolist() == [0, 2]
    assert lout[:int(kc[0])].tolist() == kl.tolist()


# ------------------------------------------------------------------------------------------ IoU
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_bbox_iou_matrix(kind):
    ops = _ops()
    gt = syn.gt_targets(31, 1, 80, max_gt=40, min_gt=40)[0]["bbox"]
    cx, _ = yolo_ref.grid_table(syn.COCO_ANCHORS, 416, (13, 26, 52))
    ref = yolo_ref.bbox_iou(torch.from_numpy(gt).unsqueeze(1), cx.unsqueeze(0), kind).numpy()
    got = ops.box_iou(torch.from_numpy(gt).cuda(), cx.cuda(), kind, xcycwh=True).cpu().numpy()
    if kind == 3:   # atan: ulp-level libm differences
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
    else:
        np.testing.assert_array_equal(got, ref)
    paired = ops.box_iou_paired(torch.from_numpy(gt).cuda(), cx[:40].cuda(), kind, xcycwh=True).cpu().numpy()
    if kind != 3:
        np.testing.assert_array_equal(paired, np.diag(ref[:, :40]))


def test_box_iou_torchvision_flavour():
    ops = _ops()
    b1, _, _ = syn.random_boxes(11, 37, clusters=4)
    b2, _, _ = syn.random_boxes(12, 501, clusters=4)
    got = ops.box_iou(torch.from_numpy(b1).cuda(), torch.from_numpy(b2).cuda(), ops.IOU_TV).cpu().numpy()
    np.testing.assert_array_equal(got, cref.box_iou_tv(b1, b2))


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("img,batch,max_gt", [(416, 3, 20), (608, 8, 100)])
def test_iou_match(kind, img, batch, max_gt):
    """get_target reductions: first argmax per GT and the no-object mask, bit-exact."""
    ops = _ops()
    targets = syn.gt_targets(31 + img, batch, 80, max_gt=max_gt)
    cx, _ = yolo_ref.grid_table(syn.COCO_ANCHORS, img, (img // 32, img // 16, img // 8))
    gt = np.zeros((batch, max_gt, 4), np.float32)
    cnt = np.zeros((batch,), np.int32)
    for i, t in enumerate(targets):
        cnt[i] = t["bbox"].shape[0]
        gt[i, :cnt[i]] = t["bbox"]
    best, noobj = ops.iou_match(torch.from_numpy(gt).cuda(), torch.from_numpy(cnt).cuda(), cx.cuda(), kind, 0.5)
    for i, t in enumerate(targets):
        wb, wf, _ = cref.iou_match(t["bbox"], cx.numpy(), kind, 0.5)
        np.testing.assert_array_equal(best[i, :cnt[i]].cpu().numpy(), wb)
        np.testing.assert_array_equal(noobj[i].cpu().numpy(), wf)


# ------------------------------------------------------------------------------------------ RPN
@pytest.mark.parametrize("shape,bsz,pre,post", [((224, 320), 2, 300, 300), ((416, 608), 2, 1000, 1000),
                                                ((800, 1344), 2, 2000, 2000)])
@pytest.mark.parametrize("strategy", ["vanilla", "trick"])
def test_rpn_filter(shape, bsz, pre, post, strategy, nms_path):
    ops = _ops()
    ih, iw = shape
    obj, deltas, anchors, per_level = syn.rpn_inputs(41, bsz, ih, iw)
    # decode+clip values carry transcendental ulps, so the oracle NMS runs on the GPU's boxes:
    # stage 1 (top-k index set, order, boxes, scores) against the oracle in tolerance / exactly,
    # stage 2 (NMS + top-n) bit-exact on identical inputs.
    mode = ops.NMS_TV_CLASS if strategy == "vanilla" else ops.NMS_TV_TRICK
    hw = torch.tensor([[ih, iw]] * bsz, dtype=torch.float32).cuda()
    # (a) no suppression at all (thr=1.0): output == filtered top-k list, checks select/decode/filter
    everything = sum(min(pre, n) for n in per_level)
    boxes, scores, index, count = ops.rpn_filter(torch.from_numpy(obj).cuda(), torch.from_numpy(deltas).cuda(),
                                                 torch.from_numpy(anchors).cuda(), per_level, hw, pre, everything,
                                                 1.0, 0.0, 1e-3, mode)
    tobj, tdel, tanc = torch.from_numpy(obj), torch.from_numpy(deltas), torch.from_numpy(anchors)
    fb, fs, fi = tv_ref.filter_proposals(tobj, tdel, tanc, per_level, [(ih, iw)] * bsz, pre, everything,
                                         nms_thresh=1.0)
    for i in range(bsz):
        k = int(count[i])
        assert k == fb[i].shape[0]
        # same index SET (top-k per level + filters); the global score order may swap neighbours
        # whose sigmoid differs by an ulp between libdevice and Sleef, so compare index-aligned
        gi_, ri_ = index[i, :k].cpu().numpy(), fi[i].numpy().astype(np.int32)
        go, ro = np.argsort(gi_, kind="stable"), np.argsort(ri_, kind="stable")
        np.testing.assert_array_equal(gi_[go], ri_[ro])
        _close_coord(boxes[i, :k].cpu().numpy()[go], fb[i].numpy()[ro], max(ih, iw))
        _close_score(scores[i, :k].cpu().numpy()[go], fs[i].numpy()[ro])
        assert (np.diff(scores[i, :k].cpu().numpy()) <= 0).all()     # emitted in descending score
    # (b) real threshold: compare with the oracle NMS fed with the GPU's stage-1 rows
    gb, gs, gi, gc = boxes, scores, index, count
    boxes, scores, index, count = ops.rpn_filter(torch.from_numpy(obj).cuda(), torch.from_numpy(deltas).cuda(),
                                                 torch.from_numpy(anchors).cuda(), per_level, hw, pre, post,
                                                 0.7, 0.0, 1e-3, mode)
    lvl_of = np.repeat(np.arange(len(per_level)), per_level)
    for i in range(bsz):
        k1 = int(gc[i])
        b1, s1, i1 = gb[i, :k1].cpu(), gs[i, :k1].cpu(), gi[i, :k1].cpu().numpy()
        # stage-1 rows come out sorted by score; restore the reference's level-major order
        order = np.lexsort((-s1.numpy(), lvl_of[i1]))
        b1, s1, i1 = b1[order], s1[order], i1[order]
        lv = torch.from_numpy(lvl_of[i1])
        fn = tv_ref.batched_nms_vanilla if strategy == "vanilla" else tv_ref.batched_nms_coordinate_trick
        want = fn(b1, s1, lv, 0.7)[:post].numpy()
        k = int(count[i])
        assert k == len(want)
        np.testing.assert_array_equal(index[i, :k].cpu().numpy(), i1[want])
        np.testing.assert_array_equal(boxes[i, :k].cpu().numpy(), b1.numpy()[want])


@pytest.mark.parametrize("name", ["c1_416_idf", "c2_608_b4", "odd_grid_352", "lvis_96_a6", "tiny_sigmoid"])
def test_decode_variants_agree(name):
    """The three fetch strategies of the fused decode kernel (gated / stream / TMA ring) must give the
    same candidates: identical anchor sets and labels, values equal up to the softmax summation order."""
    from object_detectors_b200 import _lib
    ops = _ops()
    lib = _lib.load()
    heads, b, img, c, anchors, idf, softmax = _case(name)
    gh = _gpu_heads(heads)
    gi = None if idf is None else idf.cuda()
    outs = []
    try:
        for variant in (_lib.DECODE_GATED, _lib.DECODE_STREAM, _lib.DECODE_RING):
            assert lib.b200_set_decode_variant(variant) == 0
            o = ops.yolo_decode_filter(gh, anchors, img, c, gi, softmax, 0.1)
            torch.cuda.synchronize()
            outs.append({k: v.clone() for k, v in o.items()})
    finally:
        lib.b200_set_decode_variant(_lib.DECODE_RING)
    ref = outs[0]
    for o in outs[1:]:
        np.testing.assert_array_equal(o["count"].cpu().numpy(), ref["count"].cpu().numpy())
        for i in range(b):
            n = int(ref["count"][i])
            np.testing.assert_array_equal(o["anchor"][i, :n].cpu().numpy(), ref["anchor"][i, :n].cpu().numpy())
            np.testing.assert_array_equal(o["label"][i, :n].cpu().numpy(), ref["label"][i, :n].cpu().numpy())
            np.testing.assert_array_equal(o["box"][i, :n].cpu().numpy(), ref["box"][i, :n].cpu().numpy())
            _close_score(o["score"][i, :n].cpu().numpy(), ref["score"][i, :n].cpu().numpy())


def test_c2_full_size_end_to_end(nms_path):
    """The bench.py workload itself (BASELINE configs[1]: 608 / COCO-80 / batch 64, seed 1000) through the fused
    call, against the oracle run end to end on the CPU: per-image candidate anchor sets, kept anchor lists (in
    score order) and labels after the majority vote bit-exact, 40 762 candidates and 8 365 detections in total.
    (The workload's closest score is 9.5e-5 relative away from the confidence threshold, far above fp32 rounding.)"""
    ops = _ops()
    b, img, c = 64, 608, 80
    heads = syn.yolo_heads(1000, b, img, c, syn.COCO_ANCHORS, "clustered")
    idf = _idf("coco")
    gh = _gpu_heads(heads)
    cand = ops.yolo_decode_filter(gh, syn.COCO_ANCHORS, img, c, idf.cuda(), True, 0.1)
    det, keep, anchor, dcnt, ccnt = ops.yolo_postprocess(gh, syn.COCO_ANCHORS, img, c, idf.cuda(), True, 0.1, 0.6,
                                                         ops.NMS_MAJORITY)
    torch.cuda.synchronize()
    total_c = total_k = 0
    for b0 in range(0, b, 16):
        recs = yolo_ref.score_filter(yolo_ref.decode([torch.from_numpy(h[b0:b0 + 16]) for h in heads], syn.COCO_ANCHORS,
                                                     img, c, idf, True), 0.1)
        for j, r in enumerate(recs):
            i = b0 + j
            d = r["det6"].numpy()
            n = d.shape[0]
            assert int(ccnt[i]) == n == int(cand["count"][i])
            ref_anchor = r["anchor"].numpy().astype(np.int32)
            np.testing.assert_array_equal(cand["anchor"][i, :n].cpu().numpy(), ref_anchor)
            ki, kl = cref.nms_majority(d, 0.6, c) if n else (np.zeros(0, np.int32), np.zeros(0, np.int32))
            k = int(dcnt[i])
            assert k == len(ki), f"image {i}: {k} kept vs {len(ki)}"
            np.testing.assert_array_equal(anchor[i, :k].cpu().numpy(), ref_anchor[ki])
            np.testing.assert_array_equal(keep[i, :k].cpu().numpy(), ki)
            np.testing.assert_array_equal(det[i, :k, 5].cpu().numpy(), kl.astype(np.float32))
            _close_coord(det[i, :k, :4].cpu().numpy(), d[ki, :4], img)
            _close_score(det[i, :k, 4].cpu().numpy(), d[ki, 4])
            total_c += n
            total_k += k
    assert (total_c, total_k) == (40762, 8365)


def test_sigmoid_label_is_first_maximum_of_the_probabilities():
    """in sigmoid mode the reference's label is the first maximum of the fp32 PROBABILITIES
    (test_one_epoch.py:35), not of the logits: where the sigmoid is saturated (1.0f above ~16.6) or flat, several classes
    tie and the lowest class index wins.  Cells are planted with such ties; every decode variant must label them like
    the oracle."""
    from object_detectors_b200 import _lib
    ops = _ops()
    lib = _lib.load()
    img, c, b = 128, 80, 2
    heads = syn.yolo_heads(23, b, img, c, syn.COCO_ANCHORS, "clustered", sigmoid_cls=True, max_objects=4)
    g = np.random.default_rng(5)
    planted = 0
    for h in heads:
        bsz, ch, gh, gw = h.shape
        t = h.reshape(bsz, 3, 5 + c, gh, gw)
        for _ in range(6):
            bi, a, y, x = int(g.integers(bsz)), int(g.integers(3)), int(g.integers(gh)), int(g.integers(gw))
            t[bi, a, 4, y, x] = 4.0                                  # live cell
            lo, hi = sorted(g.choice(c, size=2, replace=False).tolist())
            t[bi, a, 5 + lo, y, x] = 18.0 + g.random()               # saturated: sigmoid == 1.0f
            t[bi, a, 5 + hi, y, x] = 30.0 + g.random()               # larger logit, same probability, HIGHER index
            planted += 1
    ref = yolo_ref.score_filter(yolo_ref.decode([torch.from_numpy(h) for h in heads], syn.COCO_ANCHORS, img, c, None, False), 0.1)
    gh_ = _gpu_heads(heads)
    try:
        for variant in (_lib.DECODE_GATED, _lib.DECODE_STREAM, _lib.DECODE_RING):
            assert lib.b200_set_decode_variant(variant) == 0
            out = ops.yolo_decode_filter(gh_, syn.COCO_ANCHORS, img, c, None, False, 0.1)
            for i in range(b):
                n = int(out["count"][i])
                assert n == ref[i]["det6"].shape[0]
                np.testing.assert_array_equal(out["anchor"][i, :n].cpu().numpy(), ref[i]["anchor"].numpy().astype(np.int32))
                np.testing.assert_array_equal(out["label"][i, :n].cpu().numpy(), ref[i]["det6"].numpy()[:, 5].astype(np.int32),
                                              err_msg=f"variant {variant}")
    finally:
        lib.b200_set_decode_variant(_lib.DECODE_RING)
    assert planted == 18
