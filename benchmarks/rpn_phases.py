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
    nms_phases(lib, lambda: ops.rpn_filter(to, td, ta, per_level, hw, args.k, args.k, 0.7, 0.0, 1e-3, ops.NMS_TV_CLASS))


def nms_phases(lib, run):
    """phase stamps of the single-launch NMS kernel inside the filter (per CTA: start, seg|team|n, ranked, strips,
    arrived, blocks, votes, emit)"""
    prof = torch.zeros((160, 8), dtype=torch.int64, device="cuda")
    lib.b200_debug_set_resolve_prof(C.c_void_p(prof.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.b200_debug_set_resolve_prof(None)
    pr = prof.cpu().numpy()
    pr = pr[pr[:, 0] > 0]
    t0 = pr[:, 0].min()
    n = pr[:, 1] & 0xffffffff
    team = (pr[:, 1] >> 32) & 0xff
    res = pr[:, 7] > 0
    us = lambda a: np.round(np.asarray(a, dtype=np.float64) / 1e3, 1)   # noqa: E731
    print(f"NMS kernel: {len(pr)} CTAs, span {us(pr[:, 2:].max() - t0)} us, teams {sorted(set(team.tolist()))}")
    seg = pr[:, 1] >> 40
    for lv in range(5):
        m = (seg % 5) == lv
        if m.any():
            print(f"  level {lv}: {int(m.sum())} CTAs, n {int(n[m].min())}..{int(n[m].max())} | rank {us((pr[m, 2] - pr[m, 0]).mean())} | strips mean "
                  f"{us((pr[m, 3] - pr[m, 2]).mean())} max {us((pr[m, 3] - pr[m, 2]).max())}")
    for lo, hi in ((0, 1000), (1000, 4097)):
        m = (n > lo) & (n <= hi)
        if not m.any():
            continue
        r = m & res
        print(f"  segments of {lo + 1}..{hi} boxes: {int(m.sum())} CTAs | rank {us((pr[m, 2] - pr[m, 0]).mean())} | strips mean "
              f"{us((pr[m, 3] - pr[m, 2]).mean())} max {us((pr[m, 3] - pr[m, 2]).max())} | strips end (since start) max {us((pr[m, 3] - t0).max())}"
              f" | resolver: blocks {us((pr[r, 5] - pr[r, 4]).mean())} votes+emit {us((pr[r, 7] - pr[r, 5]).mean())} end max {us((pr[r, 7] - t0).max())}")


if __name__ == "__main__":
    main()
