#!/usr/bin/env python
"""Cycle stamps of the resolve kernel's phases per image (C2 workload), via b200_debug_set_resolve_prof."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detectors_b200 import _lib, ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
heads = [torch.from_numpy(h).to(dev) for h in syn.yolo_heads(1000, 64, 608, 80, syn.COCO_ANCHORS, "clustered")]
idf = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_coco_smooth.npy"))).to(dev)
plan = ops.YoloPostprocess([19, 38, 76], 64, syn.COCO_ANCHORS, 608, 80, True, 0.1, 0.6, ops.NMS_MAJORITY, 4096, 256, dev)
for _ in range(5):
    plan(heads, idf)
torch.cuda.synchronize()
buf = torch.zeros((64, 16), dtype=torch.int64, device=dev)
lib.b200_debug_set_resolve_prof(C.c_void_p(buf.data_ptr()))
plan(heads, idf)
torch.cuda.synchronize()
lib.b200_debug_set_resolve_prof(None)
b = buf.cpu().numpy()
d = np.diff(b[:, :6], axis=1)
order = np.argsort(-d.sum(1))
print("phase cycles: A stage | B fixed point | C vote | D order | E emit | total   (n, K)")
for i in list(order[:6]) + list(order[-3:]):
    print(f"img {i:2d}: " + " | ".join(f"{x:7d}" for x in d[i]) + f" | {d[i].sum():7d}   ({b[i, 6]}, {b[i, 7]})")
c = b[:, [2, 8, 9, 10, 3]]
print("phase C split (first suppressor+vote | scan | scatter | majority):", " | ".join(f"{x:7.0f}" for x in np.diff(c, axis=1).mean(0)))
c = b[:, [3, 11, 12, 4]]
print("phase D split (compaction | key fill | rank count):", " | ".join(f"{x:7.0f}" for x in np.diff(c, axis=1).mean(0)))
print("mean   : " + " | ".join(f"{x:7.0f}" for x in d.mean(0)) + f" | {d.sum(1).mean():7.0f}")
