"""Seeded synthetic inputs for the box-ops hot path.

Everything here is plain numpy driven by ``numpy.random.Generator(PCG64(seed))`` so the
same seed yields the same bytes in the build container and on the GPU box (same image,
same numpy).  Nothing in this module touches a GPU or the oracle; it only manufactures
head tensors / boxes of the shapes BASELINE.json names.

Generators
----------
* ``yolo_heads``        raw YOLO head tensors ``[B, A*(5+C), H, H]`` for the three strides
                        (32, 16, 8), either *clustered* (a handful of planted objects per
                        image, realistic duplicate clusters for NMS) or *uniform* (stress).
* ``gt_targets``        relative ``xc,yc,w,h`` ground-truth boxes for target matching (C4).
* ``rpn_inputs``        objectness / deltas / anchors for the RPN proposal filter (C5).
* ``random_boxes``      xyxy boxes + tie-free scores for stand-alone NMS / IoU tests.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

# anchor sets of the reference configs (yolo/hydra/dataset/coco.yaml:6-8, lvis.yaml:6-11),
# coarse -> fine, pixel units at the network input resolution.
COCO_ANCHORS: Tuple[Tuple[Tuple[float, float], ...], ...] = (
    ((116, 90), (156, 198), (373, 326)),
    ((30, 61), (62, 45), (59, 119)),
    ((10, 13), (16, 30), (33, 23)),
)
LVIS_ANCHORS: Tuple[Tuple[Tuple[float, float], ...], ...] = (
    ((155.78819651, 244.03609716), (320.272707, 116.94313185), (293.30877626, 232.00399174),
     (116, 90), (156, 198), (373, 326)),
    ((56.46791643, 96.62934705), (89.66263185, 59.3598243), (127.82328124, 40.61556824),
     (30, 61), (62, 45), (59, 119)),
    ((13.5255288, 23.31384949), (31.50078774, 9.86228439), (20.81998901, 13.66625921),
     (10, 13), (16, 30), (33, 23)),
)
STRIDES = (32, 16, 8)


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def num_anchors_total(img: int, anchors=COCO_ANCHORS) -> int:
    return sum((img // s) ** 2 * len(a) for s, a in zip(STRIDES, anchors))


def yolo_heads(seed: int, batch: int, img: int, num_classes: int,
               anchors: Sequence[Sequence[Tuple[float, float]]] = COCO_ANCHORS,
               mode: str = "clustered", sigmoid_cls: bool = False,
               max_objects: int = 15) -> List[np.ndarray]:
    """Raw head tensors, one per stride, NCHW float32 contiguous.

    clustered: background logits far below the objectness threshold plus, per planted
    object, the 3x3 cells around its centre on every scale and every anchor firing with
    a box regressed onto the object -> ~1-3 % of cells pass ``conf > 0.1`` and form
    tight duplicate clusters (what NMS exists for).
    uniform: i.i.d. logits, one boosted class per cell, ~10 % pass; little overlap.
    """
    g = _rng(seed)
    ch = 5 + num_classes
    heads5d = []
    for s, anc in zip(STRIDES, anchors):
        h = img // s
        a = len(anc)
        t = g.standard_normal((batch, a, ch, h, h), dtype=np.float32)
        if mode == "clustered":
            t[:, :, 2:4] *= np.float32(0.3)
            t[:, :, 4] = t[:, :, 4] * np.float32(1.5) - np.float32(6.0)
            if sigmoid_cls:
                t[:, :, 5:] -= np.float32(4.0)
        elif mode == "uniform":
            t[:, :, 2:4] *= np.float32(0.5)
            t[:, :, 4] = t[:, :, 4] * np.float32(2.5) - np.float32(5.0)
            hot = g.integers(0, num_classes, size=(batch, a, h, h))
            bi, ai, hi, wi = np.meshgrid(np.arange(batch), np.arange(a), np.arange(h),
                                         np.arange(h), indexing="ij")
            t[bi, ai, 5 + hot, hi, wi] += np.float32(6.0)
        else:
            raise ValueError(f"unknown mode {mode!r}")
        heads5d.append(t)

    if mode == "clustered":
        lo, hi_ = math.log(16.0), math.log(0.6 * img)
        for b in range(batch):
            k = int(g.integers(1, max_objects + 1))
            cxy = (g.random((k, 2)) * 0.8 + 0.1) * img
            wh = np.exp(g.random((k, 2)) * (hi_ - lo) + lo)
            cls = g.integers(0, num_classes, size=k)
            for o in range(k):
                for si, (s, anc) in enumerate(zip(STRIDES, anchors)):
                    h = img // s
                    t = heads5d[si]
                    gx, gy = cxy[o, 0] / s, cxy[o, 1] / s
                    for dj in (-1, 0, 1):
                        for di in (-1, 0, 1):
                            ci, cj = int(gx) + di, int(gy) + dj
                            if not (0 <= ci < h and 0 <= cj < h):
                                continue
                            fx = min(max(gx - ci, 0.02), 0.98)
                            fy = min(max(gy - cj, 0.02), 0.98)
                            for ai, (aw, ah) in enumerate(anc):
                                e = g.standard_normal(5)
                                t[b, ai, 0, cj, ci] = math.log(fx / (1 - fx)) + 0.3 * e[0]
                                t[b, ai, 1, cj, ci] = math.log(fy / (1 - fy)) + 0.3 * e[1]
                                t[b, ai, 2, cj, ci] = math.log(wh[o, 0] / aw) + 0.15 * e[2]
                                t[b, ai, 3, cj, ci] = math.log(wh[o, 1] / ah) + 0.15 * e[3]
                                t[b, ai, 4, cj, ci] = 2.0 + 1.5 * e[4]
                                t[b, ai, 5 + int(cls[o]), cj, ci] += np.float32(8.0)
    return [np.ascontiguousarray(t.reshape(batch, -1, t.shape[3], t.shape[4])) for t in heads5d]


def gt_targets(seed: int, batch: int, num_classes: int, max_gt: int = 100,
               min_gt: int = 1) -> List[dict]:
    """Ground truth per image as the reference's datasets emit it: relative xc,yc,w,h
    float32 plus int64 class ids (yolo/dsets/transformations.py:44-46)."""
    g = _rng(seed)
    out = []
    for _ in range(batch):
        m = int(g.integers(min_gt, max_gt + 1))
        c = g.random((m, 2)) * 0.9 + 0.05
        wh = np.exp(g.random((m, 2)) * (-0.4 + 4.0) - 4.0)
        out.append({
            "bbox": np.concatenate([c, wh], axis=1).astype(np.float32),
            "category_id": g.integers(0, num_classes, size=m).astype(np.int64),
        })
    return out


def random_boxes(seed: int, n: int, extent: float = 608.0, clusters: int = 0,
                 num_classes: int = 0):
    """xyxy float32 boxes, tie-free float32 scores (a shuffled strictly increasing ramp
    plus noise below half a step) and optional int64 labels.  ``clusters > 0`` plants
    that many centres and jitters boxes around them so IoUs cover the whole [0,1] range."""
    g = _rng(seed)
    if clusters > 0:
        cen = g.random((clusters, 2)) * extent * 0.8 + extent * 0.1
        size = np.exp(g.random((clusters, 2)) * 2.5 + 2.5)
        which = g.integers(0, clusters, size=n)
        c = cen[which] + g.standard_normal((n, 2)) * size[which] * 0.15
        wh = size[which] * np.exp(g.standard_normal((n, 2)) * 0.2)
    else:
        c = g.random((n, 2)) * extent
        wh = np.exp(g.random((n, 2)) * 3.0 + 2.0)
    boxes = np.concatenate([c - wh / 2, c + wh / 2], axis=1).astype(np.float32)
    ramp = (np.arange(n, dtype=np.float64) + 0.25 + 0.5 * g.random(n)) / max(n, 1)
    scores = g.permutation(ramp).astype(np.float32)
    if len(np.unique(scores)) != n:  # float32 collisions only for huge n; spread again
        scores = g.permutation(np.linspace(0.001, 0.999, n)).astype(np.float32)
    labels = g.integers(0, num_classes, size=n).astype(np.int64) if num_classes > 0 else None
    return boxes, scores, labels


# ----------------------------------------------------------------------------------------
# RPN (config C5): torchvision AnchorGenerator geometry restated (anchor_utils.py:60-134)
# ----------------------------------------------------------------------------------------
def rpn_level_shapes(img_h: int = 800, img_w: int = 1344) -> List[Tuple[int, int]]:
    """FPN feature-map sizes for strides 4,8,16,32 plus the stride-2 max-pool of the last."""
    shapes = []
    for s in (4, 8, 16, 32):
        shapes.append((math.ceil(img_h / s), math.ceil(img_w / s)))
    lh, lw = shapes[-1]
    shapes.append((math.ceil(lh / 2), math.ceil(lw / 2)))
    return shapes


def rpn_anchors(img_h: int = 800, img_w: int = 1344,
                sizes=((32,), (64,), (128,), (256,), (512,)),
                ratios=(0.5, 1.0, 2.0)) -> Tuple[np.ndarray, List[int]]:
    """Anchor table ``[sum A, 4]`` float32 and anchors per level, ordered (y, x, a) within
    a level exactly like ``AnchorGenerator.grid_anchors`` (anchor_utils.py:98-134): base
    anchors are ``round([-w,-h,w,h]/2)`` (:60-71), shifts are integer strides."""
    shapes = rpn_level_shapes(img_h, img_w)
    per_level, tables = [], []
    for (fh, fw), sz in zip(shapes, sizes):
        stride_h, stride_w = img_h // fh, img_w // fw
        r = np.asarray(ratios, dtype=np.float32)
        hr = np.sqrt(r)
        wr = np.float32(1.0) / hr
        scales = np.asarray(sz, dtype=np.float32)
        ws = (wr[:, None] * scales[None, :]).reshape(-1)
        hs = (hr[:, None] * scales[None, :]).reshape(-1)
        base = np.round(np.stack([-ws, -hs, ws, hs], axis=1) / np.float32(2)).astype(np.float32)
        sx = (np.arange(fw, dtype=np.int32) * stride_w).astype(np.float32)
        sy = (np.arange(fh, dtype=np.int32) * stride_h).astype(np.float32)
        yy, xx = np.meshgrid(sy, sx, indexing="ij")
        shifts = np.stack([xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)], axis=1)
        t = (shifts[:, None, :] + base[None, :, :]).reshape(-1, 4).astype(np.float32)
        tables.append(t)
        per_level.append(t.shape[0])
    return np.concatenate(tables, axis=0), per_level


def rpn_inputs(seed: int, batch: int, img_h: int = 800, img_w: int = 1344):
    """objectness ``[B, sumA]`` ~ N(-3, 2^2) (tie-free by construction check), deltas
    ``[B, sumA, 4]`` ~ N(0, 0.2^2), anchors ``[sumA, 4]``, per-level counts."""
    g = _rng(seed)
    anchors, per_level = rpn_anchors(img_h, img_w)
    n = anchors.shape[0]
    obj = g.standard_normal((batch, n), dtype=np.float32) * np.float32(2.0) - np.float32(3.0)
    deltas = g.standard_normal((batch, n, 4), dtype=np.float32) * np.float32(0.2)
    return obj, deltas, anchors, per_level


def retina_inputs(seed: int, batch: int, img_h: int, img_w: int, num_classes: int):
    """Head outputs of a RetinaNet-style many-class head on the FPN anchor layout of ``rpn_anchors``:
    ``cls_logits [B, sumA, C]`` ~ N(-5, 1.2^2) (a few percent pass sigmoid > 0.05: the large levels overflow the
    1000-candidate top-k, the small ones do not) with clustered boosts so that the class-aware NMS has duplicates to
    suppress, ``bbox_regression [B, sumA, 4]`` ~ N(0, 0.3^2), ``anchors [sumA, 4]``, anchors per level."""
    g = _rng(seed)
    anchors, per_level = rpn_anchors(img_h, img_w)
    n = anchors.shape[0]
    logits = g.standard_normal((batch, n, num_classes), dtype=np.float32) * np.float32(1.2) - np.float32(5.0)
    regs = g.standard_normal((batch, n, 4), dtype=np.float32) * np.float32(0.3)
    cx, cy = (anchors[:, 0] + anchors[:, 2]) * 0.5, (anchors[:, 1] + anchors[:, 3]) * 0.5
    for b in range(batch):
        for _ in range(int(g.integers(3, 8))):
            ox, oy = g.uniform(0.2, 0.8) * img_w, g.uniform(0.2, 0.8) * img_h
            near = np.nonzero((np.abs(cx - ox) < 24) & (np.abs(cy - oy) < 24))[0]
            logits[b, near, int(g.integers(0, num_classes))] += np.float32(6.0) + g.standard_normal(near.shape[0]).astype(np.float32)
    return logits, regs, anchors, per_level


def ssd_inputs(seed: int, batch: int, num_anchors: int, num_classes: int, img: int = 300):
    """Head outputs of an SSD300-style head: ``cls_logits [B, A, C]`` (background column 0 dominant; a few object
    clusters whose class passes 0.01 on many anchors -- one of them on more than 400, so the per-class top-k bites),
    ``bbox_regression [B, A, 4]`` ~ N(0, 0.5^2), default boxes ``[A, 4]`` clustered around the objects."""
    g = _rng(seed)
    logits = g.standard_normal((batch, num_anchors, num_classes), dtype=np.float32)
    logits[:, :, 0] += np.float32(10.0)
    regs = g.standard_normal((batch, num_anchors, 4), dtype=np.float32) * np.float32(0.5)
    k = 6
    centres = g.uniform(0.15, 0.85, size=(k, 2)) * img
    sizes = np.exp(g.uniform(np.log(20), np.log(0.5 * img), size=(k, 2)))
    which = g.integers(0, k, size=num_anchors)
    c = centres[which] + g.standard_normal((num_anchors, 2)) * sizes[which] * 0.1
    s = sizes[which] * np.exp(g.standard_normal((num_anchors, 2)) * 0.15)
    anchors = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
    for b in range(batch):
        cls = g.integers(1, num_classes, size=k)
        for j in range(k):
            rows = np.nonzero(which == j)[0]
            take = rows if j == 0 else rows[: max(8, len(rows) // 6)]           # cluster 0: every anchor fires (> 400)
            logits[b, take, cls[j]] += np.float32(6.5) + g.standard_normal(len(take)).astype(np.float32)
    return logits, regs, anchors


def roi_inputs(seed: int, rows_per_image: Sequence[int], num_classes: int, img_h: int = 800, img_w: int = 1216):
    """Box-head outputs of a Faster R-CNN ROI head: ``class_logits [R, C]`` (background column 0 dominant except for
    a few foreground rows per cluster), ``box_regression [R, 4C]`` ~ N(0, 0.5^2), ``proposals`` list of ``[r_i, 4]``
    clustered around a handful of objects so that the per-class NMS has something to suppress."""
    g = _rng(seed)
    props, logits, regs = [], [], []
    for r in rows_per_image:
        k = int(g.integers(2, 7))
        centres = g.uniform(0.15, 0.85, size=(k, 2)) * np.array([img_w, img_h])
        sizes = np.exp(g.uniform(np.log(32), np.log(0.5 * min(img_h, img_w)), size=(k, 2)))
        cls = g.integers(1, num_classes, size=k)
        which = g.integers(0, k, size=r)
        c = centres[which] + g.standard_normal((r, 2)) * sizes[which] * 0.08
        s = sizes[which] * np.exp(g.standard_normal((r, 2)) * 0.1)
        p = np.concatenate([c - s / 2, c + s / 2], 1)
        p[:, 0::2] = np.clip(p[:, 0::2], 0, img_w)
        p[:, 1::2] = np.clip(p[:, 1::2], 0, img_h)
        lg = g.standard_normal((r, num_classes)).astype(np.float32)
        lg[:, 0] += 2.0
        fg = g.random(r) < 0.6
        lg[np.arange(r)[fg], cls[which][fg]] += g.uniform(3.0, 7.0, size=int(fg.sum())).astype(np.float32)
        props.append(p.astype(np.float32))
        logits.append(lg)
        regs.append((g.standard_normal((r, 4 * num_classes)) * 0.5).astype(np.float32))
    return np.concatenate(logits), np.concatenate(regs), props


def legacy_head(seed: int, batch: int, num_anchors: int, num_classes: int, grid: int) -> np.ndarray:
    """One raw head tensor ``[B, A*(5+C), grid, grid]`` ~ N(0, 1.5^2) for the legacy YOLOLoss layer."""
    g = _rng(seed)
    return (g.standard_normal((batch, num_anchors * (5 + num_classes), grid, grid)) * 1.5).astype(np.float32)
