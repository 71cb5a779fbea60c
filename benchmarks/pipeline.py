#!/usr/bin/env python
"""Explores the decode-stream / NMS-stream software pipeline on the C2 workload.

Decode launches of consecutive steps go round-robin over `--dstreams` streams (kernels on one stream are
serial, so exactly that many decode kernels are in flight), the NMS chain of a step runs on one of
`--nstreams` other streams after the step's decode (event), and `--plans` workspaces rotate.

    python benchmarks/pipeline.py --dstreams 2 --nstreams 2 --plans 4 --ring 4,1,101
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detectors_b200 import _lib, ops, synthetic as syn  # noqa: E402

IMG, NC, BATCH = 608, 80, 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="ring")
    ap.add_argument("--ring", default="4,1,101")
    ap.add_argument("--dstreams", type=int, default=2)
    ap.add_argument("--nstreams", type=int, default=2)
    ap.add_argument("--plans", type=int, default=4)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--hiprio", action="store_true", help="NMS streams get high priority")
    ap.add_argument("--no-nms", action="store_true", help="decode kernels only (how fast is the HBM stage alone?)")
    ap.add_argument("--conf", type=float, default=0.1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    lib.b200_set_decode_variant({"gated": 0, "stream": 1, "ring": 3}[args.variant])
    lib.b200_debug_set_ring(*[int(x) for x in args.ring.split(",")])
    heads = [torch.from_numpy(h).to(dev) for h in syn.yolo_heads(1000, BATCH, IMG, NC, syn.COCO_ANCHORS, "clustered")]
    idf = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_coco_smooth.npy"))).to(dev)
    grids = [h.shape[2] for h in heads]
    plans = [ops.YoloPostprocess(grids, BATCH, syn.COCO_ANCHORS, IMG, NC, True, args.conf, 0.6, ops.NMS_MAJORITY, 4096, 256, dev)
             for _ in range(args.plans)]
    # every workspace reads its own copy of the input (batch rolled): concurrently running decode kernels must not
    # find each other's lines in L2
    inputs = [[h.roll(7 * k, 0).contiguous() for h in heads] for k in range(args.plans)]
    ds = [torch.cuda.Stream(device=dev) for _ in range(args.dstreams)]
    nst = [torch.cuda.Stream(device=dev, priority=-1 if args.hiprio else 0) for _ in range(args.nstreams)]
    dec_done = [torch.cuda.Event() for _ in plans]
    nms_done = [torch.cuda.Event() for _ in plans]
    for e in nms_done:
        e.record()

    def step(i, ev=None):
        k = i % len(plans)
        d, n = ds[i % len(ds)], nst[i % len(nst)]
        with torch.cuda.stream(d):
            d.wait_event(nms_done[k])                 # the workspace is free again
            if ev is not None:
                lib.b200_debug_set_decode_events(C.c_void_p(ev[0].cuda_event), C.c_void_p(ev[1].cuda_event))
            plans[k].decode(inputs[k], idf)
            dec_done[k].record(d)
        if args.no_nms:
            nms_done[k].record(d)
            return
        with torch.cuda.stream(n):
            n.wait_event(dec_done[k])
            if ev is not None:
                lib.b200_debug_set_timeline(C.c_void_p(ev[2].cuda_event), C.c_void_p(ev[3].cuda_event), C.c_void_p(ev[4].cuda_event))
            plans[k].nms()
            nms_done[k].record(n)

    for i in range(16):
        step(i)
    torch.cuda.synchronize()
    ref = (int(plans[0].cand_count.sum()), int(plans[0].det_count.sum()))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    for row in evs:
        for e in row:
            e.record()
    t0 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    ev0 = torch.cuda.Event(); ev0.record()
    for st in ds + nst:
        st.wait_event(ev0)
    for i in range(args.steps):
        step(i, evs[i])
    lib.b200_debug_set_decode_events(None, None)
    lib.b200_debug_set_timeline(None, None, None)
    torch.cuda.synchronize()
    for pl in plans:
        pl.check_status()
        assert args.no_nms or (int(pl.cand_count.sum()), int(pl.det_count.sum())) == ref
    if args.no_nms:
        for row in evs:
            row[2] = row[3] = row[4] = row[1]
    ts = np.array([[t0.elapsed_time(e) * 1e3 for e in row] for row in evs])
    lo = args.steps // 4
    period = (ts[-1, 4] - ts[lo, 4]) / (args.steps - 1 - lo)
    iv = sorted((a, b) for a, b in ts[lo:, :2])
    busy, ca, cb = 0.0, iv[0][0], iv[0][1]
    for a, b in iv[1:]:
        if a > cb:
            busy += cb - ca; ca, cb = a, b
        else:
            cb = max(cb, b)
    busy = (busy + cb - ca) / len(iv)
    print(f"variant={args.variant} ring={args.ring} D={args.dstreams} N={args.nstreams} P={args.plans} hiprio={args.hiprio}: "
          f"period {period:.1f} us, decode launch-to-end {np.mean(ts[lo:, 1] - ts[lo:, 0]):.1f}, busy/launch {busy:.1f}, "
          f"nms chain {np.mean(ts[lo:, 4] - ts[lo:, 1]):.1f} (plan {np.mean(ts[lo:, 2] - ts[lo:, 1]):.1f} pairs {np.mean(ts[lo:, 3] - ts[lo:, 2]):.1f} "
          f"resolve {np.mean(ts[lo:, 4] - ts[lo:, 3]):.1f}) kept={ref}", flush=True)


if __name__ == "__main__":
    main()
