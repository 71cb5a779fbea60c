#!/usr/bin/env python
"""Sweeps the fused decode+filter kernel variants on the C2 workload (608 / COCO / batch 64).

For every configuration it reports (a) the kernel's duration alone on one stream (CUDA events around
the launch, L2-exceeding input) and (b) the pipelined whole-step rate with the NMS kernels of other
steps overlapping it on 3 streams -- the number bench.py reports as `value`.

    python benchmarks/decode_sweep.py [--steps 300] > gpurun_out/decode_sweep.jsonl
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detectors_b200 import _lib, ops, synthetic as syn  # noqa: E402

IMG, NC, BATCH = 608, 80, 64


CONF = 0.1


def run(lib, heads, idf, grids, dev, n_streams, steps, check=None):
    plans = [ops.YoloPostprocess(grids, BATCH, syn.COCO_ANCHORS, IMG, NC, True, CONF, 0.6, ops.NMS_MAJORITY,
                                 4096, 256, dev) for _ in range(n_streams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]

    def step(i):
        with torch.cuda.stream(streams[i % n_streams]):
            plans[i % n_streams](heads, idf)

    def fence_in():
        ev = torch.cuda.Event(); ev.record()
        for st in streams:
            st.wait_event(ev)

    def fence_out():
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)

    fence_in()
    for i in range(10):
        step(i)
    fence_out()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record(); b.record()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    fence_in()
    for i in range(steps):
        lib.b200_debug_set_decode_events(C.c_void_p(evs[i][0].cuda_event), C.c_void_p(evs[i][1].cuda_event))
        step(i)
    lib.b200_debug_set_decode_events(None, None)
    fence_out()
    t1.record()
    torch.cuda.synchronize()
    for pl in plans:
        pl.check_status()
    sig = (int(plans[0].cand_count.sum()), int(plans[0].det_count.sum()),
           float(plans[0].det[:, :, :5].double().sum()))
    if check is not None:
        assert sig[:2] == check[:2], (sig, check)
    return t0.elapsed_time(t1) / steps, float(np.mean([a.elapsed_time(b) for a, b in evs])), sig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--only", default=None, help="run only the configuration with this name")
    ap.add_argument("--streams", default="1,3")
    ap.add_argument("--conf", type=float, default=0.1, help="confidence threshold (2.0: nothing passes = pure streaming)")
    args = ap.parse_args()
    global CONF
    CONF = args.conf
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib = _lib.load()
    heads = [torch.from_numpy(h).to(dev) for h in syn.yolo_heads(1000, BATCH, IMG, NC, syn.COCO_ANCHORS, "clustered")]
    idf = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_coco_smooth.npy"))).to(dev)
    grids = [h.shape[2] for h in heads]
    algo = BATCH * sum(g * g * 3 for g in grids) * (5 + NC) * 4
    configs = [("gated", 0, None), ("stream", 1, None)]
    for warps, slots, ctas in ((4, 1, 1), (6, 1, 1), (8, 1, 1), (4, 2, 1), (3, 2, 1), (2, 3, 1), (3, 1, 2), (4, 1, 2)):
        configs.append((f"ring w{warps} s{slots} c{ctas}", 3, (warps, slots, ctas)))
    ref = None
    if args.only:
        configs = [c for c in configs if c[0] == args.only]
    for name, variant, tune in configs:
        lib.b200_set_decode_variant(variant)
        if tune:
            lib.b200_debug_set_ring(*tune)
        out = {"config": name}
        for ns in [int(x) for x in args.streams.split(',')]:
            ms_step, ms_k, sig = run(lib, heads, idf, grids, dev, ns, args.steps, ref)
            ref = ref or sig
            out[f"streams{ns}"] = {"ms_per_step": ms_step, "images_per_s": BATCH / (ms_step * 1e-3), "kernel_ms": ms_k,
                                   "kernel_algo_GBs": algo / (ms_k * 1e-3) / 1e9}
        out["sig"] = sig
        print(json.dumps(out), flush=True)
    lib.b200_set_decode_variant(0)


if __name__ == "__main__":
    main()
