#!/usr/bin/env python
"""Latency of the NMS stage alone and of the serial post-process (one stream) on the C2 workload, for the
single-launch path (nms_fused.cu) and the general three-launch path (nms.cu).

    python benchmarks/nms_latency.py [--reps 200] [--gen clustered|uniform]
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
from object_detectors_b200 import _lib, ops, synthetic as syn  # noqa: E402

IMG, NC, BATCH = 608, 80, 64


INNER = 10


def timed(fn, reps, flush=None):
    """Device time per call: INNER calls are enqueued back to back between two events, so the host's enqueue cost
    (Python + ctypes, ~15 us per call) hides behind the previous call's kernels instead of inflating the number."""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(max(1, reps // INNER)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(INNER):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / INNER)
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--gen", default="clustered")
    ap.add_argument("--batch", type=int, default=BATCH)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    heads = [torch.from_numpy(h).to(dev) for h in syn.yolo_heads(1000, args.batch, IMG, NC, syn.COCO_ANCHORS, args.gen)]
    idf = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_coco_smooth.npy"))).to(dev)
    grids = [h.shape[2] for h in heads]
    plan = ops.YoloPostprocess(grids, args.batch, syn.COCO_ANCHORS, IMG, NC, True, 0.1, 0.6, ops.NMS_MAJORITY, 4096, 256, dev)
    out = {}
    ref = None
    for name, general in (("fused", 0), ("general", 1)):
        lib.b200_debug_set_nms_path(general)
        plan.decode(heads, idf)
        torch.cuda.synchronize()
        nms_med, nms_min = timed(lambda: plan.nms(), args.reps)
        full_med, full_min = timed(lambda: plan(heads, idf), args.reps)
        dec_med, dec_min = timed(lambda: plan.decode(heads, idf), args.reps)
        plan.check_status()
        res = (plan.det_count.clone(), plan.det.clone(), plan.det_keep.clone())
        if ref is None:
            ref = res
        else:
            assert torch.equal(ref[0], res[0]), "paths disagree on counts"
            for b in range(args.batch):
                k = int(ref[0][b])
                assert torch.equal(ref[1][b, :k], res[1][b, :k]) and torch.equal(ref[2][b, :k], res[2][b, :k]), f"image {b}"
        out[name] = {"nms_us": nms_med, "nms_us_min": nms_min, "postprocess_us": full_med, "postprocess_us_min": full_min,
                     "decode_us": dec_med, "decode_us_min": dec_min}
    lib.b200_debug_set_nms_path(0)
    # phase stamps of the fused kernel (globaltimer, ns): per CTA [start, seg|team|n, ranked, strips, arrived, blocks, votes, emit]
    import ctypes as C
    prof = torch.zeros((160, 8), dtype=torch.int64, device=dev)
    plan.decode(heads, idf)
    lib.b200_debug_set_resolve_prof(C.c_void_p(prof.data_ptr()))
    plan.nms()
    torch.cuda.synchronize()
    lib.b200_debug_set_resolve_prof(None)
    lib.b200_debug_set_nms_path(-1)         # back to the library default
    pr = prof.cpu().numpy()
    pr = pr[pr[:, 0] > 0]
    t0 = pr[:, 0].min()
    resolver = pr[:, 7] > 0
    us = lambda a: float(np.round(a / 1e3, 2))
    out["fused_phases_us"] = {
        "ctas": int(pr.shape[0]), "rank_mean": us((pr[:, 2] - pr[:, 0]).mean()), "rank_max": us((pr[:, 2] - pr[:, 0]).max()),
        "strips_mean": us((pr[:, 3] - pr[:, 2]).mean()), "strips_max": us((pr[:, 3] - pr[:, 2]).max()),
        "arrive_mean": us((pr[:, 4] - pr[:, 3]).mean()),
        "blocks_mean": us((pr[resolver, 5] - pr[resolver, 4]).mean()), "blocks_max": us((pr[resolver, 5] - pr[resolver, 4]).max()),
        "votes_mean": us((pr[resolver, 6] - pr[resolver, 5]).mean()), "relabel_emit_mean": us((pr[resolver, 7] - pr[resolver, 6]).mean()),
        "first_start_to_last_emit": us(pr[resolver, 7].max() - t0), "start_spread": us(pr[:, 0].max() - t0),
        "team_max": int(((pr[:, 1] >> 32) & 0xff).max()), "n_max": int((pr[:, 1] & 0xffffffff).max())}
    out["candidates"] = int(plan.cand_count.sum())
    out["kept"] = int(plan.det_count.sum())
    out["max_candidates_per_image"] = int(plan.cand_count.max())
    out["gen"] = args.gen
    print(json.dumps(out))


if __name__ == "__main__":
    main()
