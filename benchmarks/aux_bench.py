#!/usr/bin/env python
"""Device timings of the other BASELINE.json configurations (C3 LVIS decode+NMS, C4 target matching,
C5 RPN proposal filter).  These are parity-test shapes, not bench.py lines; the numbers go into
DESIGN.md / profiles/.  One JSON line per configuration.

    python benchmarks/aux_bench.py [--reps 20]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detectors_b200 import ops, synthetic as syn  # noqa: E402

HBM_GBS = 6534.1
try:
    HBM_GBS = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
LANE_OPS_PEAK = 148 * 4 * 32 * 1.965e9      # lane-instructions / s at the max SM clock


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


_c3_cache = {}


def c3_lvis(reps, batch=8):
    """LVIS-1203, 6 anchors / scale, 608: 219.8 MB per image (C3 is quoted at b32 = 7.0 GB).  Input: the survey's
    CLUSTERED generator (8 distinct images, repeated to fill larger batches); the first two images are checked against
    the CPU oracle (candidate counts, kept counts) before anything is timed."""
    from oracle import cref, yolo_ref
    if "heads" not in _c3_cache:
        _c3_cache["np"] = syn.yolo_heads(3001, 8, 608, 1203, syn.LVIS_ANCHORS, "clustered")
        _c3_cache["heads"] = [torch.from_numpy(h).cuda() for h in _c3_cache["np"]]
    base = _c3_cache["heads"]
    heads = [h if batch == 8 else torch.cat([h] * (batch // 8), 0).contiguous() for h in base] if batch >= 8 else [h[:batch].contiguous() for h in base]
    idf_cpu = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_lvis_smooth.npy")))
    idf = idf_cpu.cuda()
    plan = ops.YoloPostprocess([19, 38, 76], batch, syn.LVIS_ANCHORS, 608, 1203, True, 0.1, 0.6, ops.NMS_MAJORITY,
                               4096, 512, "cuda")
    plan(heads, idf)
    plan.check_status()
    cand = ops.yolo_decode_filter([h[:2].contiguous() for h in heads], syn.LVIS_ANCHORS, 608, 1203, idf, True, 0.1, capacity=4096)
    for i in range(2):       # oracle check of the timed input (same images whatever the batch)
        rec = yolo_ref.score_filter(yolo_ref.decode([torch.from_numpy(h[i:i + 1]) for h in _c3_cache["np"]], syn.LVIS_ANCHORS, 608,
                                                    1203, idf_cpu, True), 0.1)[0]
        n = rec["det6"].shape[0]
        assert int(plan.cand_count[i]) == n, (int(plan.cand_count[i]), n)
        g6 = torch.cat([cand["box"][i, :n], cand["score"][i, :n, None], cand["label"][i, :n, None].float()], 1).cpu().numpy()
        ki, _ = cref.nms_majority(g6, 0.6, 1203)
        assert int(plan.det_count[i]) == len(ki), (int(plan.det_count[i]), len(ki))
    ms = timeit(lambda: plan(heads, idf), reps)
    nbytes = sum(h.numel() * 4 for h in heads)
    return {"config": "C3 YOLOv3-608 LVIS-1203 A=6 decode+filter+nms_majority (clustered generator, oracle-checked)",
            "batch": batch, "ms": ms, "images_per_s": batch / ms * 1e3, "algorithmic_GBs": nbytes / ms / 1e6,
            "hbm_frac": nbytes / ms / 1e6 / HBM_GBS, "candidates": int(plan.cand_count.sum()), "kept": int(plan.det_count.sum())}


def c2_uniform(reps, batch=64):
    """C2 on the UNIFORM ("stress") generator: ~10 % of the cells pass instead of ~3 % -- the same bytes, three times
    the live cells and ~2300 candidates per image for the NMS.  Serial calls on one stream (no software pipeline)."""
    base = [torch.from_numpy(h).cuda() for h in syn.yolo_heads(2001, 8, 608, 80, syn.COCO_ANCHORS, "uniform")]
    heads = [torch.cat([h] * (batch // 8), 0).contiguous() for h in base]
    idf = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_coco_smooth.npy"))).cuda()
    plan = ops.YoloPostprocess([19, 38, 76], batch, syn.COCO_ANCHORS, 608, 80, True, 0.1, 0.6, ops.NMS_MAJORITY, 4096, 4096, "cuda")
    ms_dec = timeit(lambda: plan.decode(heads, idf), reps)
    ms = timeit(lambda: plan(heads, idf), reps)
    plan.check_status()
    nbytes = sum(h.numel() * 4 for h in heads)
    return {"config": "C2 YOLOv3-608 COCO b64, UNIFORM generator (stress: ~10 % live cells), serial decode+NMS", "batch": batch,
            "ms": ms, "decode_ms": ms_dec, "images_per_s": batch / ms * 1e3, "decode_GBs": nbytes / ms_dec / 1e6,
            "decode_hbm_frac": nbytes / ms_dec / 1e6 / HBM_GBS, "candidates": int(plan.cand_count.sum()),
            "kept": int(plan.det_count.sum())}


def c4_match(reps, batch=64, max_gt=100):
    from oracle import yolo_ref
    cx, _ = yolo_ref.grid_table(syn.COCO_ANCHORS, 608, (19, 38, 76))
    targets = syn.gt_targets(77, batch, 80, max_gt=max_gt)
    gt = np.zeros((batch, max_gt, 4), np.float32)
    cnt = np.zeros((batch,), np.int32)
    for i, t in enumerate(targets):
        cnt[i] = t["bbox"].shape[0]
        gt[i, :cnt[i]] = t["bbox"]
    g, c, a = torch.from_numpy(gt).cuda(), torch.from_numpy(cnt).cuda(), cx.cuda()
    out = []
    for kind, name, ops_per_pair in ((0, "IoU", 12), (1, "GIoU", 23)):
        ms = timeit(lambda: ops.iou_match(g, c, a, kind, 0.5), reps)
        pairs = int(cnt.sum()) * a.shape[0]
        out.append({"config": f"C4 target matching {name}, b{batch}, M<=100, N=22743", "ms": ms, "pairs": pairs,
                    "pairs_per_s": pairs / ms * 1e3, "images_per_s": batch / ms * 1e3,
                    "sm_issue_frac": pairs * ops_per_pair / (ms * 1e-3) / LANE_OPS_PEAK})
    return out


def c5_rpn(reps, batch=16):
    obj, deltas, anchors, per_level = syn.rpn_inputs(41, batch, 800, 1344)
    o, d, a = torch.from_numpy(obj).cuda(), torch.from_numpy(deltas).cuda(), torch.from_numpy(anchors).cuda()
    hw = torch.tensor([[800.0, 1333.0]] * batch).cuda()
    out = []
    for pre in (2000, 1000):
        for mode, name in ((ops.NMS_TV_CLASS, "vanilla"), (ops.NMS_TV_TRICK, "coordinate_trick")):
            ms = timeit(lambda: ops.rpn_filter(o, d, a, per_level, hw, pre, pre, 0.7, 0.0, 1e-3, mode), reps)
            b, s, i, c = ops.rpn_filter(o, d, a, per_level, hw, pre, pre, 0.7, 0.0, 1e-3, mode)
            out.append({"config": f"C5 RPN filter 800x1344 b{batch} pre/post={pre} {name}", "ms": ms,
                        "images_per_s": batch / ms * 1e3, "proposals_kept": int(c.sum()),
                        "objectness_GBs": o.numel() * 4 / ms / 1e6})
    return out


def dense_decode(reps, batch=64):
    heads = [torch.from_numpy(h).cuda() for h in syn.yolo_heads(1000, batch, 608, 80, syn.COCO_ANCHORS, "clustered")]
    idf = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_coco_smooth.npy"))).cuda()
    ms = timeit(lambda: ops.yolo_decode_dense(heads, syn.COCO_ANCHORS, 608, 80, idf, True), reps)
    nbytes = 2 * sum(h.numel() * 4 for h in heads)          # one read + one write of [B, N, 85]
    out = [{"config": f"a1 YOLOForw.forward dense decode 608 COCO b{batch} (softmax x IDF)", "ms": ms,
            "images_per_s": batch / ms * 1e3, "algorithmic_GBs": nbytes / ms / 1e6, "hbm_frac": nbytes / ms / 1e6 / HBM_GBS}]
    x = torch.from_numpy(syn.legacy_head(5, batch, 3, 80, 76)).cuda()
    ms = timeit(lambda: ops.yolo_legacy_decode(x, syn.COCO_ANCHORS[2], 80, 608), reps)
    nbytes = 2 * x.numel() * 4
    out.append({"config": f"a9 legacy YOLOLoss.forward decode, 76x76 head b{batch}", "ms": ms,
                "algorithmic_GBs": nbytes / ms / 1e6, "hbm_frac": nbytes / ms / 1e6 / HBM_GBS})
    return out


def roi_heads(reps):
    out = []
    for name, c, batch, rows in (("COCO-91", 91, 16, 1000), ("LVIS-1204", 1204, 4, 1000)):
        logits, regs, props = syn.roi_inputs(51, [rows] * batch, c, 800, 1216)
        lg, rg = torch.from_numpy(logits).cuda(), torch.from_numpy(regs).cuda()
        pr = [torch.from_numpy(p).cuda() for p in props]
        shapes = [(800, 1216)] * batch
        ms = timeit(lambda: ops.roi_postprocess(lg, rg, pr, shapes, None, ops.ROI_SOFTMAX, capacity=8192), reps)
        det, keep, dcnt, ccnt, status = ops.roi_postprocess(lg, rg, pr, shapes, None, ops.ROI_SOFTMAX, capacity=8192)
        nbytes = (lg.numel() + rg.numel()) * 4
        out.append({"config": f"a13 RoIHeads.postprocess_detections {name} b{batch} x {rows} proposals", "ms": ms,
                    "images_per_s": batch / ms * 1e3, "input_GBs": nbytes / ms / 1e6, "candidates": int(ccnt.sum()),
                    "kept": int(dcnt.sum()), "status": int(status.item())})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rows = [c3_lvis(max(args.reps // 4, 3)), c3_lvis(max(args.reps // 4, 3), batch=32), c2_uniform(args.reps)] + c4_match(args.reps) + c5_rpn(args.reps) + dense_decode(args.reps) + roi_heads(args.reps)
    for r in rows:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
