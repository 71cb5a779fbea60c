#!/usr/bin/env python
"""Per-kernel timeline of the software-pipelined post-process (C2 workload): start/end of the decode,
plan, pairs and resolve kernels of consecutive steps on their streams, from CUDA events.

    python benchmarks/timeline.py --variant ring --streams 3 --steps 40
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
    ap.add_argument("--variant", default="gated")
    ap.add_argument("--ring", default="4,1,1")
    ap.add_argument("--streams", type=int, default=3)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--show", type=int, default=12)
    ap.add_argument("--serial", action="store_true", help="decode kernels never overlap each other: decode i waits for decode i-1")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    lib.b200_set_decode_variant({"gated": 0, "stream": 1, "ring": 3}[args.variant])
    lib.b200_debug_set_ring(*[int(x) for x in args.ring.split(",")])
    heads = [torch.from_numpy(h).to(dev) for h in syn.yolo_heads(1000, BATCH, IMG, NC, syn.COCO_ANCHORS, "clustered")]
    idf = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_coco_smooth.npy"))).to(dev)
    grids = [h.shape[2] for h in heads]
    ns = args.streams
    plans = [ops.YoloPostprocess(grids, BATCH, syn.COCO_ANCHORS, IMG, NC, True, 0.1, 0.6, ops.NMS_MAJORITY, 4096, 256, dev)
             for _ in range(ns)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]

    def step(i):
        with torch.cuda.stream(streams[i % ns]):
            plans[i % ns](heads, idf)

    for i in range(12):
        step(i)
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    for row in evs:
        for e in row:
            e.record()
    t0 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    ev0 = torch.cuda.Event(); ev0.record()
    for st in streams:
        st.wait_event(ev0)
    for i in range(args.steps):
        e = evs[i]
        lib.b200_debug_set_decode_events(C.c_void_p(e[0].cuda_event), C.c_void_p(e[1].cuda_event))
        lib.b200_debug_set_timeline(C.c_void_p(e[2].cuda_event), C.c_void_p(e[3].cuda_event), C.c_void_p(e[4].cuda_event))
        if args.serial and i > 0:
            streams[i % ns].wait_event(evs[i - 1][1])
        step(i)
    lib.b200_debug_set_decode_events(None, None)
    lib.b200_debug_set_timeline(None, None, None)
    torch.cuda.synchronize()
    ts = np.array([[t0.elapsed_time(e) * 1e3 for e in row] for row in evs])     # microseconds
    lo = args.steps - args.show
    print(f"variant={args.variant} ring={args.ring} streams={ns} serial={args.serial}: decode {np.mean(ts[lo:, 1] - ts[lo:, 0]):.1f} us, step period "
          f"{(ts[-1, 4] - ts[lo, 4]) / (args.steps - 1 - lo):.1f} us")
    if args.show > 20:
        return
    print("step strm | decode start..end (dur) | plan end (dur) | pairs end (dur) | resolve end (dur)")
    for i in range(lo, args.steps):
        a, b, c, d, e = ts[i] - ts[lo, 0]
        print(f"{i:4d} {i % ns:4d} | {a:7.1f} .. {b:7.1f} ({b - a:5.1f}) | {c:7.1f} ({c - b:5.1f}) | {d:7.1f} ({d - c:5.1f}) | {e:7.1f} ({e - d:5.1f})")


if __name__ == "__main__":
    main()
