#!/usr/bin/env python
"""C5 (Faster R-CNN FPN RPN, batch 16, 800 x 1344, 268 569 anchors per image): device time of the proposal filter
(select + decode + NMS + merge) and of the per-level top-k select alone (RegionProposalNetwork._get_top_n_idx).

    python benchmarks/rpn_bench.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detectors_b200 import ops, synthetic as syn  # noqa: E402

INNER = 10


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(INNER):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / INNER)
    return float(np.median(ts))


def main():
    peak = 6534.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    bsz, ih, iw = 16, 800, 1344
    obj, deltas, anchors, per_level = syn.rpn_inputs(41, bsz, ih, iw)
    to, td, ta = torch.from_numpy(obj).cuda(), torch.from_numpy(deltas).cuda(), torch.from_numpy(anchors).cuda()
    hw = torch.tensor([[ih, iw]] * bsz, dtype=torch.float32).cuda()
    obj_bytes = obj.size * 4
    for k in (2000, 1000):
        ms = timed(lambda: ops.rpn_top_n_idx(to, per_level, k))
        print(json.dumps({"config": f"C5 _get_top_n_idx b16 800x1344 k={k}", "ms": ms, "objectness_GBs": obj_bytes / ms / 1e6,
                          "hbm_frac": obj_bytes / ms / 1e6 / peak}), flush=True)
        for name, mode in (("vanilla", ops.NMS_TV_CLASS), ("coordinate_trick", ops.NMS_TV_TRICK)):
            ms = timed(lambda: ops.rpn_filter(to, td, ta, per_level, hw, k, k, 0.7, 0.0, 1e-3, mode))
            cnt = ops.rpn_filter(to, td, ta, per_level, hw, k, k, 0.7, 0.0, 1e-3, mode)[3]
            print(json.dumps({"config": f"C5 RPN filter b16 800x1344 pre/post={k} {name}", "ms": ms,
                              "images_per_s": bsz / ms * 1e3, "proposals_kept": int(cnt.sum())}), flush=True)


if __name__ == "__main__":
    main()
