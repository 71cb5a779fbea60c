#!/usr/bin/env python
"""The driver's protocol (5 warm-up + 20 timed steps) on the C2 workload for different pipeline shapes.

`--chains C`: C streams, each with its own workspace; a step runs [memset, decode, NMS] serially on stream
step % C, so C steps are in flight and no cross-stream events are needed.
`--split D,N,P`: round 1's shape -- decode kernels round-robin over D streams, NMS chains over N other streams,
P workspaces.

    python benchmarks/pipe20.py --chains 3 --ring 4,1,101 --nms fused
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ring", default="4,1,101")
    ap.add_argument("--chains", type=int, default=3)
    ap.add_argument("--split", default="")
    ap.add_argument("--nms", default="auto", choices=["auto", "general", "fused", "fused256"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--long", type=int, default=1000)
    ap.add_argument("--gen", default="clustered")
    ap.add_argument("--resolve", default="", help="general NMS path: resolve CTA shape 'threads,smem_kb' (b200_debug_set_resolve)")
    ap.add_argument("--timeline", action="store_true", help="print decode start/end and NMS end of every timed step (split mode)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    lib.b200_set_decode_variant(3)
    lib.b200_debug_set_ring(*[int(x) for x in args.ring.split(",")])
    lib.b200_debug_set_nms_path({"auto": -1, "general": 1, "fused": 0, "fused256": 2}[args.nms])
    if args.resolve:
        lib.b200_debug_set_resolve(*[int(x) for x in args.resolve.split(",")])
    heads = [torch.from_numpy(h).to(dev) for h in syn.yolo_heads(1000, BATCH, IMG, NC, syn.COCO_ANCHORS, args.gen)]
    idf = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_coco_smooth.npy"))).to(dev)
    grids = [h.shape[2] for h in heads]
    if args.split:
        n_d, n_n, n_p = [int(x) for x in args.split.split(",")]
    else:
        n_d = n_n = 0
        n_p = args.chains
    plans = [ops.YoloPostprocess(grids, BATCH, syn.COCO_ANCHORS, IMG, NC, True, 0.1, 0.6, ops.NMS_MAJORITY, 4096, 256, dev)
             for _ in range(n_p)]
    inputs = [[h.roll(7 * k, 0).contiguous() for h in heads] for k in range(n_p)]
    if args.split:
        ds = [torch.cuda.Stream(device=dev) for _ in range(n_d)]
        ns = [torch.cuda.Stream(device=dev) for _ in range(n_n)]
        streams = ds + ns
        dec_done = [torch.cuda.Event() for _ in plans]
        nms_done = [torch.cuda.Event() for _ in plans]

        def step(i):
            k = i % n_p
            d, n = ds[i % n_d], ns[i % n_n]
            d.wait_event(nms_done[k])
            plans[k].decode(inputs[k], idf, d)
            dec_done[k].record(d)
            n.wait_event(dec_done[k])
            plans[k].nms(n)
            nms_done[k].record(n)
    else:
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_p)]
        nms_done = []

        def step(i):
            k = i % n_p
            plans[k].decode(inputs[k], idf, streams[k])
            plans[k].nms(streams[k])

    def run(steps, warm):
        def fence_in():
            ev = torch.cuda.Event()
            ev.record()
            for st in streams:
                st.wait_event(ev)
            for e in nms_done:
                e.record()

        def fence_out():
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)
        fence_in()
        for i in range(warm):
            step(i)
        fence_out()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fence_in()
        for i in range(steps):
            step(i)
        fence_out()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    if args.timeline and args.split:
        import ctypes as C
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        for row in evs:
            for e in row:
                e.record()
        torch.cuda.synchronize()
        base_step = step

        def step(i, _b=base_step):      # noqa: F811
            lib.b200_debug_set_decode_events(C.c_void_p(evs[i][0].cuda_event), C.c_void_p(evs[i][1].cuda_event))
            _b(i)
            evs[i][2].record(ns[i % n_n])
        run(args.steps, args.warmup)
        t0 = torch.cuda.Event(enable_timing=True)
        ms0 = run.__globals__ if False else None
    ms = [run(args.steps, args.warmup) for _ in range(args.reps)]
    if args.timeline and args.split:
        lib.b200_debug_set_decode_events(None, None)
        ref = evs[0][0]
        print("step | decode start .. end (dur) | nms end (after decode end)")
        for i, (a, b, c) in enumerate(evs):
            ta, tb, tc = ref.elapsed_time(a) * 1e3, ref.elapsed_time(b) * 1e3, ref.elapsed_time(c) * 1e3
            print(f"{i:4d} | {ta:8.1f} .. {tb:8.1f} ({tb - ta:6.1f}) | {tc:8.1f} (+{tc - tb:6.1f})")
    for pl in plans:
        pl.check_status()
    long_ms = run(args.long, 10) if args.long else 0.0
    print(json.dumps({"ring": args.ring, "chains": args.chains, "split": args.split, "nms": args.nms, "steps": args.steps, "resolve": args.resolve,
                      "ms": [round(m, 4) for m in ms], "img_per_s_best": BATCH * args.steps / (min(ms) * 1e-3),
                      "img_per_s_first": BATCH * args.steps / (ms[0] * 1e-3),
                      "us_per_step_20": 1e3 * float(np.median(ms)) / args.steps,
                      "us_per_step_long": 1e3 * long_ms / max(args.long, 1), "kept": int(plans[0].det_count.sum())}), flush=True)


if __name__ == "__main__":
    main()
