"""oracle/yolo_ref.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch CPU tensors, fp32) of the YOLO half of the detection box-ops hot
path of kostas1515/object_detectors.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` leg may import this module; the product
package ``object_detectors_b200`` never does.

Why torch CPU ops and not numpy: the reference *is* a sequence of torch CPU/CUDA tensor
ops, and its fp32 transcendentals (sigmoid / exp / softmax) come from ATen's vectorised
kernels.  Restating the path with the same primitive ops, applied in the same order,
makes this oracle bit-identical to the reference on CPU, which is what
``tests/golden/make_golden.py`` (run in the build container against /root/reference) and
``tests/test_oracle_golden.py`` pin.  Sequential loops additionally exist in scalar C
(oracle/boxops_ref.c) for the full-size parity runs.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# decode  (yolo/nets/yolo_forw.py:81-119, 163-176)
# --------------------------------------------------------------------------------------
def grid_table(anchors: Sequence[Sequence[Tuple[float, float]]], img_size: float,
               grid_sizes: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """Per-anchor constants ``cxypwh [N,4]`` = (cell centre x, y, anchor w, h, all divided by
    the grid width) and ``inw [N]`` (grid width as fp32), scales concatenated in input order,
    flat index ``n = (h*W + w)*A + a`` inside a scale.  yolo_forw.py:93-119:
    ``scaled = tensor(a_w / (img/in_w))`` (python double -> fp32, :99), ``grid = (idx+0.5)/in_w``
    (:104-107), ``anchor = scaled/in_w`` (:108-113)."""
    rows, widths = [], []
    for k, g in enumerate(grid_sizes):
        a = len(anchors[k])
        stride = img_size / g                                   # python float, :97-98
        scaled = torch.tensor([(aw / stride, ah / stride) for aw, ah in anchors[k]],
                              dtype=torch.float32)              # :99
        col = (torch.arange(g, dtype=torch.float32) + 0.5) / g  # (w+0.5)/in_w in fp32
        gx = col.view(1, g, 1).expand(g, g, a)                  # varies along w
        gy = col.view(g, 1, 1).expand(g, g, a)                  # varies along h
        aw = (scaled[:, 0] / g).view(1, 1, a).expand(g, g, a)
        ah = (scaled[:, 1] / g).view(1, 1, a).expand(g, g, a)
        rows.append(torch.stack((gx, gy, aw, ah), dim=-1).reshape(-1, 4))
        widths.append(torch.full((g * g * a,), float(g), dtype=torch.float32))   # :116
    return torch.cat(rows, 0).contiguous(), torch.cat(widths, 0)


def flatten_heads(heads: Sequence[Tensor], num_anchors: Sequence[int], num_classes: int) -> Tensor:
    """``[B, A*(5+C), H, W]`` x3 -> ``[B, N, 5+C]`` with n = (h*W+w)*A + a (yolo_forw.py:101-103,118)."""
    ch = 5 + num_classes
    flat = []
    for t, a in zip(heads, num_anchors):
        b, _, h, w = t.shape
        flat.append(t.reshape(b, a, ch, h, w).permute(0, 3, 4, 1, 2).reshape(b, h * w * a, ch))
    return torch.cat(flat, dim=1)


def decode(heads: Sequence[Tensor], anchors, img_size: float, num_classes: int,
           idf: Optional[Tensor] = None, softmax: bool = True) -> Tensor:
    """Inference branch of ``YOLOForw.forward`` (yolo_forw.py:163-176): centre-format pixel
    boxes, objectness, class probabilities ``[B, N, 5+C]``.
    ``idf`` None reproduces ``idf_logits = tensor(1)`` (:38); softmax=True is
    ``class_loss == CrossEntropyLoss`` (:168-169)."""
    grids = [int(t.shape[3]) for t in heads]
    cxypwh, inw = grid_table(anchors, img_size, grids)
    raw = flatten_heads(heads, [len(a) for a in anchors], num_classes)
    stride = (img_size / inw).unsqueeze(1)                      # :164
    inw1 = inw.unsqueeze(1)                                     # :165
    xy = (torch.sigmoid(raw[..., 0:2]) + cxypwh[:, :2] * inw1 - 0.5) * stride     # :166
    wh = torch.exp(raw[..., 2:4]) * cxypwh[:, 2:4] * inw1 * stride                # :167
    conf = torch.sigmoid(raw[:, :, 4:5])                                          # :168
    scale = torch.tensor(1) if idf is None else idf
    logits = scale.unsqueeze(0).unsqueeze(0) * raw[:, :, 5:]
    cls = torch.softmax(logits, dim=2) if softmax else torch.sigmoid(logits)      # :169-173
    return torch.cat((xy, wh, conf, cls), dim=2)


def abs_coord(box: Tensor) -> Tensor:
    """helper.get_abs_coord (yolo/utilities/helper.py:203-217): centre format -> corners."""
    half_w, half_h = box[..., 2] / 2, box[..., 3] / 2
    return torch.stack((box[..., 0] - half_w, box[..., 1] - half_h,
                        box[..., 0] + half_w, box[..., 1] + half_h), dim=-1)


# --------------------------------------------------------------------------------------
# filter / compaction  (yolo/procedures/test_one_epoch.py:24-35)
# --------------------------------------------------------------------------------------
def score_filter(pred: Tensor, conf_thr: float) -> List[Dict[str, Tensor]]:
    """``pred`` is the decode output.  Per image (empties are kept here; the reference drops
    them from its python list at :34): ``anchor`` int64 [n_i] ascending flat anchor index,
    ``det6`` fp32 [n_i, 6] = x1,y1,x2,y2, conf*max_c cls, float(argmax_c cls).
    The score is formed twice in the reference (:25 on the dense tensor, :35 on the gathered
    rows); both are the same fp32 product, restated once."""
    pred = pred.clone()
    pred[:, :, :4] = abs_coord(pred[:, :, :4])                       # :24
    best, label = pred[:, :, 5:].max(dim=2)
    score = pred[:, :, 4] * best                                     # :25
    mask = score > conf_thr                                          # :26
    out = []
    for b in range(pred.shape[0]):
        idx = torch.nonzero(mask[b]).flatten()                       # :27-28 order preserving
        det6 = torch.cat((pred[b, idx, :4], score[b, idx, None],
                          label[b, idx, None].to(torch.float32)), dim=1)   # :35
        out.append({"anchor": idx, "det6": det6})
    return out


# --------------------------------------------------------------------------------------
# class-agnostic greedy NMS with majority-vote relabel  (yolo/utilities/helper.py:280-382)
# --------------------------------------------------------------------------------------
def nms_majority(det6: Tensor, thresh_iou: float = 0.6) -> Tuple[Tensor, Tensor]:
    """Returns ``(kept_rows [K,6], keep_idx int64 [K])`` in descending score; kept rows carry
    the relabelled class.  ``det6`` is not modified (the reference mutates ``P[:,5]`` in
    place, the drop-in layer reproduces that side effect)."""
    n = det6.shape[0]
    if n == 0:
        return det6.new_zeros((0, 6)), torch.zeros(0, dtype=torch.int64)
    x1, y1, x2, y2, score = (det6[:, i] for i in range(5))
    votes_from = det6[:, 5].to(torch.int32)                          # snapshot, :301
    area = (x2 - x1) * (y2 - y1)                                     # :304
    # descending score, lower index first on ties (the reference's argsort at :308 leaves
    # tie order undefined; golden inputs are tie-free)
    remaining = torch.sort(score, descending=True, stable=True)[1]
    keep_idx, keep_lab = [], []
    while remaining.numel() > 0:
        s = int(remaining[0])
        remaining = remaining[1:]
        label = float(det6[s, 5])
        if remaining.numel() > 0:
            w = (torch.minimum(x2[remaining], x2[s]) - torch.maximum(x1[remaining], x1[s])).clamp(min=0.0)
            h = (torch.minimum(y2[remaining], y2[s]) - torch.maximum(y1[remaining], y1[s])).clamp(min=0.0)
            inter = w * h                                            # :358
            union = (area[remaining] - inter) + area[s]              # :363
            iou = inter / union                                      # :366
            voters = votes_from[remaining[iou > thresh_iou]]         # :369
            if voters.numel() > 0:
                cats, cnts = torch.unique(voters, return_counts=True)
                if cats.numel() > 1:                                 # :372
                    label = float(cats[int(torch.argmax(cnts))])     # first max -> smallest id
            remaining = remaining[iou < thresh_iou]                  # :368,:380
        keep_idx.append(s)
        keep_lab.append(label)
    keep = torch.tensor(keep_idx, dtype=torch.int64)
    rows = det6[keep].clone()
    rows[:, 5] = torch.tensor(keep_lab, dtype=torch.float32)
    return rows, keep


def postprocess(heads: Sequence[Tensor], anchors, img_size: float, num_classes: int,
                idf: Optional[Tensor], softmax: bool, conf_thr: float = 0.1,
                nms_thr: float = 0.6) -> List[Dict[str, Tensor]]:
    """decode -> xyxy -> filter -> nms_majority for a batch (test_one_epoch.py:22-36).
    Per image: ``anchor``/``det6`` (candidates, ascending anchor index) and ``keep`` (indices
    into the candidate list, descending score) / ``kept`` rows."""
    out = score_filter(decode(heads, anchors, img_size, num_classes, idf, softmax), conf_thr)
    for rec in out:
        rec["kept"], rec["keep"] = nms_majority(rec["det6"], nms_thr)
    return out


# --------------------------------------------------------------------------------------
# pairwise IoU family + target matching  (helper.py:221-277, yolo_forw.py:178-208)
# --------------------------------------------------------------------------------------
def bbox_iou(bb1: Tensor, bb2: Tensor, iou_type: int = 0, xcycwh: bool = True) -> Tensor:
    """Broadcasting IoU (0) / GIoU (1) / DIoU (2) / CIoU (3), helper.py:221-277."""
    b1 = abs_coord(bb1) if xcycwh else bb1
    b2 = abs_coord(bb2) if xcycwh else bb2
    ax1, ay1, ax2, ay2 = b1[..., 0], b1[..., 1], b1[..., 2], b1[..., 3]
    bx1, by1, bx2, by2 = b2[..., 0], b2[..., 1], b2[..., 2], b2[..., 3]
    inter = (torch.min(ax2, bx2) - torch.max(ax1, bx1)).clamp(0) * \
            (torch.min(ay2, by2) - torch.max(ay1, by1)).clamp(0)                  # :249-250
    w1, h1, w2, h2 = ax2 - ax1, ay2 - ay1, bx2 - bx1, by2 - by1
    union = (w1 * h1 + 1e-16) + w2 * h2 - inter                                   # :255
    iou = inter / union
    if iou_type == 0 or iou_type not in (1, 2, 3):
        return iou
    cw = torch.max(ax2, bx2) - torch.min(ax1, bx1)                                # :259
    ch = torch.max(ay2, by2) - torch.min(ay1, by1)
    if iou_type == 1:
        c_area = cw * ch + 1e-16
        return iou - (c_area - union) / c_area                                    # :263
    c2 = cw ** 2 + ch ** 2 + 1e-16                                                # :266
    rho2 = ((bx1 + bx2) - (ax1 + ax2)) ** 2 / 4 + ((by1 + by2) - (ay1 + ay2)) ** 2 / 4
    if iou_type == 2:
        return iou - rho2 / c2
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
    alpha = (v / (1 - iou + v)).detach()
    return iou - (rho2 / c2 + v * alpha)


def get_target(targets: Sequence[Dict[str, Tensor]], cxypwh: Tensor, inw: Tensor,
               num_classes: int, ignore_threshold: float = 0.5, iou_type: int = 1):
    """YOLOForw.get_target (yolo_forw.py:178-208): per image best anchor per GT (first argmax
    of the IoU row), regression targets, and the no-object mask = every GT IoU below the
    ignore threshold, cleared at matched anchors."""
    tgt, tcls, obj, noobj = [], [], [], []
    for t in targets:
        box = t["bbox"]
        tcls.append(torch.nn.functional.one_hot(t["category_id"], num_classes).float())   # :185
        iou = bbox_iou(box.unsqueeze(1), cxypwh.unsqueeze(0), iou_type)                   # :186
        best = iou.max(dim=1)[1]                                                          # :187
        anchor = cxypwh[best]
        width = inw[best]
        px, py = box[:, 0] * width, box[:, 1] * width
        gx = torch.clamp(px - px.long(), 0.0001, 0.9999)                                  # :191-194
        gy = torch.clamp(py - py.long(), 0.0001, 0.9999)
        gw = torch.log(box[:, 2] / anchor[:, 2] + 1e-16)                                  # :196-197
        gh = torch.log(box[:, 3] / anchor[:, 3] + 1e-16)
        tgt.append(torch.stack((gx, gy, gw, gh), dim=1))
        free = (iou < ignore_threshold).all(dim=0)                                        # :200
        free[best] = False                                                                # :201
        noobj.append(free)
        obj.append(best)
    return torch.cat(tgt, 0), torch.cat(tcls, 0), obj, torch.stack(noobj, 0)


# --------------------------------------------------------------------------------------
# margin screening: golden inputs must not sit on a threshold
# --------------------------------------------------------------------------------------
def screen_margins(heads, anchors, img_size, num_classes, idf, softmax, conf_thr=0.1,
                   nms_thr=0.6, rel=1e-4) -> Dict[str, float]:
    """Smallest distances of any decision from its threshold.  A vector is usable for
    bit-exact count/keep parity when every margin is > ``rel`` (scores: relative to the
    threshold; IoU: absolute) and no two candidate scores of an image coincide."""
    pred = decode(heads, anchors, img_size, num_classes, idf, softmax)
    cands = score_filter(pred, conf_thr)
    best = pred[:, :, 5:].max(dim=2)[0]
    score = pred[:, :, 4] * best
    m_score = float(((score - conf_thr).abs() / conf_thr).min())
    top2 = pred[:, :, 5:].topk(min(2, num_classes), dim=2)[0]
    live = score > conf_thr
    m_label = float("inf")
    if num_classes > 1 and bool(live.any()):
        gap = (top2[..., 0] - top2[..., 1])[live]
        m_label = float((gap / top2[..., 0][live]).min())
    m_iou, ties = float("inf"), 0
    for rec in cands:
        d = rec["det6"]
        if d.shape[0] < 2:
            continue
        ties += int(d.shape[0] - torch.unique(d[:, 4]).numel())
        x1, y1, x2, y2 = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
        area = (x2 - x1) * (y2 - y1)
        w = (torch.min(x2[:, None], x2[None]) - torch.max(x1[:, None], x1[None])).clamp(min=0)
        h = (torch.min(y2[:, None], y2[None]) - torch.max(y1[:, None], y1[None])).clamp(min=0)
        inter = w * h
        iou = inter / ((area[:, None] - inter) + area[None])
        iou.fill_diagonal_(0)
        m_iou = min(m_iou, float((iou - nms_thr).abs().min()))
    return {"score": m_score, "label": m_label, "iou": m_iou, "score_ties": ties,
            "candidates": [int(r["det6"].shape[0]) for r in cands]}


# --------------------------------------------------------------------------------------
# legacy per-head layer  (yolo/nets/yolo_loss.py:34-105, inference branch)
# --------------------------------------------------------------------------------------
def legacy_decode(head: Tensor, anchors_px, num_classes: int, img_size) -> Tensor:
    """YOLOLoss.forward(input, targets=None): [B, A*(5+C), H, W] -> [B, A*H*W, 5+C], rows ordered (a, h, w)."""
    bs, in_h, in_w = head.size(0), head.size(2), head.size(3)
    na = len(anchors_px)
    stride_h, stride_w = img_size / in_h, img_size / in_w                                   # :38-39
    scaled = [(a_w / stride_w, a_h / stride_h) for a_w, a_h in anchors_px]                  # :40
    pred = head.view(bs, na, 5 + num_classes, in_h, in_w).permute(0, 1, 3, 4, 2).contiguous()   # :42
    x, y = torch.sigmoid(pred[..., 0]), torch.sigmoid(pred[..., 1])                          # :76-77
    w, h = pred[..., 2], pred[..., 3]
    conf, cls = torch.sigmoid(pred[..., 4]), torch.sigmoid(pred[..., 5:])                    # :80-81
    grid_x = torch.linspace(0, in_w - 1, in_w).repeat(in_w, 1).repeat(bs * na, 1, 1).view(x.shape)      # :86-87
    grid_y = torch.linspace(0, in_h - 1, in_h).repeat(in_h, 1).t().repeat(bs * na, 1, 1).view(y.shape)  # :88-89
    sa = torch.tensor(scaled, dtype=torch.float32)
    anchor_w = sa[:, 0:1].repeat(bs, 1).repeat(1, 1, in_h * in_w).view(w.shape)              # :91-94
    anchor_h = sa[:, 1:2].repeat(bs, 1).repeat(1, 1, in_h * in_w).view(h.shape)
    boxes = torch.empty(pred[..., :4].shape)
    boxes[..., 0] = x + grid_x                                                               # :97-100
    boxes[..., 1] = y + grid_y
    boxes[..., 2] = torch.exp(w) * anchor_w
    boxes[..., 3] = torch.exp(h) * anchor_h
    scale = torch.tensor([stride_w, stride_h] * 2, dtype=torch.float32)                      # :102
    return torch.cat((boxes.view(bs, -1, 4) * scale, conf.view(bs, -1, 1), cls.view(bs, -1, num_classes)), -1)
