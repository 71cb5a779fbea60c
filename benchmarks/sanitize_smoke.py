#!/usr/bin/env python
"""One small invocation of every kernel family, for `compute-sanitizer --tool memcheck|racecheck|synccheck` (one tool per
gpurun call; the summaries go to profiles/).  Shapes are tiny but hit the code paths that matter: aligned and odd
grids of the ring decode (tensor-map and row-copy staging incl. the tensor's first / last row), chunked LVIS rows,
both NMS paths, RPN sliced select (multi-slice levels), ROI heads, matching, fused matcher, exchange, emission.

    compute-sanitizer --tool memcheck python benchmarks/sanitize_smoke.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detectors_b200 import _lib, ops, synthetic as syn  # noqa: E402
from object_detectors_b200.distributed import PeerExchange  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    idf = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_coco_smooth.npy"))).to(dev)
    done = []
    for img, c, anchors, cls_idf in ((128, 80, syn.COCO_ANCHORS, idf), (352, 80, syn.COCO_ANCHORS, idf), (96, 1203, syn.LVIS_ANCHORS, None)):
        heads = [torch.from_numpy(h).to(dev) for h in syn.yolo_heads(11, 2, img, c, anchors, "clustered", max_objects=4)]
        # exact-size buffers: a row copy that strays outside the tensor would be an out-of-bounds access here
        heads = [h.clone() for h in heads]
        for path in (1, 0):
            lib.b200_debug_set_nms_path(path)
            det, keep, anchor, dcnt, ccnt = ops.yolo_postprocess(heads, anchors, img, c, cls_idf, True, 0.1, 0.6, ops.NMS_MAJORITY)
            ops.yolo_postprocess(heads, anchors, img, c, cls_idf, True, 0.1, 0.6, ops.NMS_TV_CLASS)
        lib.b200_debug_set_nms_path(-1)
        ops.yolo_decode_filter(heads, anchors, img, c, cls_idf, True, 0.1)
        ops.yolo_decode_dense(heads, anchors, img, c, cls_idf, True)
        done.append(f"yolo{img}/{c}:{int(dcnt.sum())}")
    b, s, l = syn.random_boxes(5, 900, clusters=6, num_classes=5)
    tb, ts, tl = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(l).to(dev)
    off = torch.tensor([0, 300, 300, 900], dtype=torch.int32, device=dev)
    for path in (1, 0):
        lib.b200_debug_set_nms_path(path)
        for mode in (ops.NMS_MAJORITY, ops.NMS_TV, ops.NMS_TV_CLASS, ops.NMS_TV_TRICK):
            ops.nms_segments(tb, ts, tl, off, 0.5, mode)
    lib.b200_debug_set_nms_path(-1)
    done.append("nms")
    obj, deltas, anchors, per_level = syn.rpn_inputs(41, 2, 416, 608)
    to, td, ta = torch.from_numpy(obj).to(dev), torch.from_numpy(deltas).to(dev), torch.from_numpy(anchors).to(dev)
    hw = torch.tensor([[416, 608]] * 2, dtype=torch.float32, device=dev)
    ops.rpn_filter(to, td, ta, per_level, hw, 600, 300, 0.7, 0.0, 1e-3, ops.NMS_TV_CLASS)
    ops.rpn_filter(to, td, ta, per_level, hw, 600, 300, 0.7, 0.0, 1e-3, ops.NMS_TV_TRICK)
    ops.rpn_top_n_idx(to, per_level, 600)
    done.append("rpn")
    logits, regs, props = syn.roi_inputs(51, [200, 150], 91, 416, 608)
    ops.roi_postprocess(torch.from_numpy(logits).to(dev), torch.from_numpy(regs).to(dev), [torch.from_numpy(p).to(dev) for p in props],
                        [(416, 608)] * 2, None, ops.ROI_SOFTMAX)
    done.append("roi")
    from oracle import yolo_ref
    cx, _ = yolo_ref.grid_table(syn.COCO_ANCHORS, 128, (4, 8, 16))
    targets = syn.gt_targets(31, 2, 80, max_gt=9)
    gt = np.zeros((2, 9, 4), np.float32)
    cnt = np.zeros((2,), np.int32)
    for i, t in enumerate(targets):
        cnt[i] = t["bbox"].shape[0]
        gt[i, :cnt[i]] = t["bbox"]
    for kind in (0, 1, 2):
        ops.iou_match(torch.from_numpy(gt).to(dev), torch.from_numpy(cnt).to(dev), cx.to(dev), kind, 0.5)
    g4, _, _ = syn.random_boxes(3, 12, clusters=3)
    ops.match_boxes(torch.from_numpy(g4).to(dev), tb, 0.7, 0.3, True)
    ops.match_boxes(torch.from_numpy(g4).to(dev), tb, 0.5, 0.5, False, ssd=True)
    q = ops.box_iou(torch.from_numpy(g4).to(dev), tb, ops.IOU_TV)
    ops.matcher(q, 0.7, 0.3, True)
    ops.matcher_ssd(q, 0.5)
    ops.box_iou_paired(tb[:100], tb[100:200], ops.GIOU)
    done.append("match")
    x = PeerExchange(2, det.shape[1], dev, slots=2)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    for _ in range(5):
        x.push(det, dcnt, st)
        x.wait(st)
    x.read(4, st)
    st.synchronize()
    x.close()
    ops.pack_detections(det, dcnt)
    ops.emit_results(det, dcnt, torch.tensor([[480., 640.], [375., 500.]], device=dev), torch.tensor([7, 9], device=dev), 96.0, None, True)
    done.append("exchange+emit")
    torch.cuda.synchronize()
    print("SANITIZE_SMOKE_OK " + " ".join(done))


if __name__ == "__main__":
    main()
