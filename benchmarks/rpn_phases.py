#!/usr/bin/env python
"""Where the sliced RPN select spends its time: globaltimer stamps of every CTA (debug hook), summarised per level.

    python benchmarks/rpn_phases.py [--k 2000]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()
    lib = _lib.load()
    obj, deltas, anchors, per_level = syn.rpn_inputs(41, args.batch, 800, 1344)
    to, td, ta = torch.from_numpy(obj).cuda(), torch.from_numpy(deltas).cuda(), torch.from_numpy(anchors).cuda()
    hw = torch.tensor([[800, 1344]] * args.batch, dtype=torch.float32).cuda()
    prof = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
    for it in range(3):
        if it == 2:
            lib.b200_debug_set_rpn_prof(C.c_void_p(prof.data_ptr()))
        ops.rpn_filter(to, td, ta, per_level, hw, args.k, args.k, 0.7, 0.0, 1e-3, ops.NMS_TV_CLASS)
        torch.cuda.synchronize()
    lib.b200_debug_set_rpn_prof(None)
    p = prof.cpu().numpy().reshape(-1, 8)
    p = p[p[:, 0] > 0]
    t00 = p[:, 0].min()
    print(f"CTAs {len(p)}  kernel span {(p[:, :6].max() - t00) / 1e3:.1f} us   levels {per_level}")
    print("level | CTAs | start us (min..max) | tail CTAs: staged, selected, sorted, done (median / max us since kernel start)")
    for l in sorted(set(p[:, 6])):
        q = p[p[:, 6] == l]
        tails = q[q[:, 5] > 0]
        rel = lambda c, f: f((tails[:, c] - t00) / 1e3) if len(tails) else float("nan")   # noqa: E731
        print(f"{int(l):5d} | {len(q):4d} | {(q[:, 0].min() - t00) / 1e3:6.1f} .. {(q[:, 0].max() - t00) / 1e3:6.1f} | "
              + "  ".join(f"{rel(c, np.median):6.1f}/{rel(c, np.max):6.1f}" for c in (2, 3, 4, 5)))
        if len(tails):
            d = np.diff(tails[:, [0, 2, 3, 4, 5]].astype(np.float64), axis=1) / 1e3
            print("        tail phase durations (median us): stage %.1f  select %.1f  sort %.1f  decode/emit %.1f" % tuple(np.median(d, axis=0)))


if __name__ == "__main__":
    main()
