"""oracle/tv_ref.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the torchvision-side box ops the reference's hot path calls.

The arithmetic of ``nms`` / ``batched_nms`` / ``box_iou`` lives in a third-party dependency
that is NOT vendored under /root/reference: **torchvision** (version unpinned by the
reference; the vendored detection files come from pytorch/vision ~v0.10; 0.26.0+cu128 is what
the image ships).  Its published algorithm is restated here (stable descending sort, suppress
iff ``inter / (area_i + area_j - inter) > thr`` with a double threshold) and pinned against
the installed ``torchvision.ops`` CPU kernels by tests/test_oracle_pin.py, anchored on the
reference's call sites: rpn.py:272, roi_heads.py:771, retinanet.py:463, ssd.py:423,
yolo/benchmark.py:100, telemetry.py:207/248.

The RPN pieces (``BoxCoder.decode_single``, ``filter_proposals``, ``Matcher``) are vendored in
the reference and restated from there (torchvision_models/tvision/{_utils,rpn}.py).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


def _areas(boxes: Tensor) -> Tensor:
    return (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])


def nms(boxes: Tensor, scores: Tensor, iou_threshold: float) -> Tensor:
    """torchvision.ops.nms (csrc/ops/cpu/nms_kernel.cpp): int64 kept indices, descending score
    (stable: lower index first on ties)."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.int64)
    order = torch.sort(scores, descending=True, stable=True)[1]
    b = boxes[order]
    area = _areas(b)
    dead = torch.zeros(n, dtype=torch.bool)
    keep = []
    thr = float(iou_threshold)
    for i in range(n):
        if dead[i]:
            continue
        keep.append(int(order[i]))
        if i + 1 == n:
            break
        r = b[i + 1:]
        w = (torch.minimum(r[:, 2], b[i, 2]) - torch.maximum(r[:, 0], b[i, 0])).clamp(min=0)
        h = (torch.minimum(r[:, 3], b[i, 3]) - torch.maximum(r[:, 1], b[i, 1])).clamp(min=0)
        inter = w * h
        ovr = inter / (area[i] + area[i + 1:] - inter)
        dead[i + 1:] |= ovr.double() > thr
    return torch.tensor(keep, dtype=torch.int64)


def batched_nms_vanilla(boxes: Tensor, scores: Tensor, idxs: Tensor, iou_threshold: float) -> Tensor:
    """boxes.py ``_batched_nms_vanilla``: nms per distinct idx, union sorted by score."""
    mask = torch.zeros_like(scores, dtype=torch.bool)
    for c in torch.unique(idxs):
        sel = torch.where(idxs == c)[0]
        mask[sel[nms(boxes[sel], scores[sel], iou_threshold)]] = True
    keep = torch.where(mask)[0]
    return keep[torch.sort(scores[keep], descending=True, stable=True)[1]]


def batched_nms_coordinate_trick(boxes: Tensor, scores: Tensor, idxs: Tensor,
                                 iou_threshold: float) -> Tensor:
    """boxes.py ``_batched_nms_coordinate_trick``: shift every class by (max+1)*idx in fp32."""
    if boxes.numel() == 0:
        return torch.zeros(0, dtype=torch.int64)
    offsets = idxs.to(boxes) * (boxes.max() + torch.tensor(1).to(boxes))
    return nms(boxes + offsets[:, None], scores, iou_threshold)


def batched_nms(boxes: Tensor, scores: Tensor, idxs: Tensor, iou_threshold: float,
                device_type: str = "cpu") -> Tensor:
    """Dispatch rule of torchvision 0.26 ``batched_nms`` (numel > 4000 on CPU / 100000 on CUDA
    -> vanilla, else coordinate trick)."""
    if boxes.numel() > (4000 if device_type == "cpu" else 100_000):
        return batched_nms_vanilla(boxes, scores, idxs, iou_threshold)
    return batched_nms_coordinate_trick(boxes, scores, idxs, iou_threshold)


def box_iou(boxes1: Tensor, boxes2: Tensor) -> Tensor:
    """torchvision.ops.box_iou, xyxy: union = (area1 + area2) - inter."""
    a1, a2 = _areas(boxes1), _areas(boxes2)
    lt = torch.max(boxes1[:, None, :2], boxes2[None, :, :2])
    rb = torch.min(boxes1[:, None, 2:], boxes2[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (a1[:, None] + a2[None, :] - inter)


# --------------------------------------------------------------------------------------
# RPN proposal filter  (torchvision_models/tvision/rpn.py:215-280, _utils.py:186-223)
# --------------------------------------------------------------------------------------
BBOX_XFORM_CLIP = math.log(1000.0 / 16)       # _utils.py:134


def decode_single(rel_codes: Tensor, boxes: Tensor,
                  weights: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)) -> Tensor:
    """BoxCoder.decode_single for one box per row (_utils.py:186-223)."""
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    dx, dy = rel_codes[:, 0] / weights[0], rel_codes[:, 1] / weights[1]
    dw = torch.clamp(rel_codes[:, 2] / weights[2], max=BBOX_XFORM_CLIP)
    dh = torch.clamp(rel_codes[:, 3] / weights[3], max=BBOX_XFORM_CLIP)
    pcx, pcy = dx * w + cx, dy * h + cy
    pw, ph = torch.exp(dw) * w, torch.exp(dh) * h
    half = torch.tensor(0.5, dtype=pw.dtype)
    return torch.stack((pcx - half * pw, pcy - half * ph, pcx + half * pw, pcy + half * ph), dim=1)


def encode_boxes(reference_boxes: Tensor, proposals: Tensor,
                 weights: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)) -> Tensor:
    """encode_boxes (_utils.py:80-125): regression targets of ``proposals`` w.r.t. ``reference_boxes``."""
    wts = torch.as_tensor(weights, dtype=reference_boxes.dtype)
    ew = proposals[:, 2:3] - proposals[:, 0:1]
    eh = proposals[:, 3:4] - proposals[:, 1:2]
    ex = proposals[:, 0:1] + 0.5 * ew
    ey = proposals[:, 1:2] + 0.5 * eh
    gw = reference_boxes[:, 2:3] - reference_boxes[:, 0:1]
    gh = reference_boxes[:, 3:4] - reference_boxes[:, 1:2]
    gx = reference_boxes[:, 0:1] + 0.5 * gw
    gy = reference_boxes[:, 1:2] + 0.5 * gh
    return torch.cat((wts[0] * (gx - ex) / ew, wts[1] * (gy - ey) / eh,
                      wts[2] * torch.log(gw / ew), wts[3] * torch.log(gh / eh)), dim=1)


def decode_multi(rel_codes: Tensor, boxes: Tensor,
                 weights: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)) -> Tensor:
    """BoxCoder.decode_single with k boxes per row: rel_codes [n, 4k] -> [n, 4k] (_utils.py:186-223)."""
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    dx, dy = rel_codes[:, 0::4] / weights[0], rel_codes[:, 1::4] / weights[1]
    dw = torch.clamp(rel_codes[:, 2::4] / weights[2], max=BBOX_XFORM_CLIP)
    dh = torch.clamp(rel_codes[:, 3::4] / weights[3], max=BBOX_XFORM_CLIP)
    pcx, pcy = dx * w[:, None] + cx[:, None], dy * h[:, None] + cy[:, None]
    pw, ph = torch.exp(dw) * w[:, None], torch.exp(dh) * h[:, None]
    half = torch.tensor(0.5, dtype=pw.dtype)
    return torch.stack((pcx - half * pw, pcy - half * ph, pcx + half * pw, pcy + half * ph), dim=2).flatten(1)


def filter_proposals(objectness: Tensor, deltas: Tensor, anchors: Tensor,
                     per_level: Sequence[int], image_shapes: Sequence[Tuple[int, int]],
                     pre_nms_top_n: int, post_nms_top_n: int, nms_thresh: float = 0.7,
                     score_thresh: float = 0.0, min_size: float = 1e-3,
                     device_type: str = "cpu") -> Tuple[List[Tensor], List[Tensor], List[Tensor]]:
    """RegionProposalNetwork.filter_proposals (rpn.py:230-280) fed with raw deltas:
    the reference decodes every anchor first (rpn.py:355) and then gathers the per-level
    top-k; decoding only the gathered rows gives the same values.  Returns boxes, scores and
    (extra, for index parity) the flat anchor index of every surviving proposal."""
    bsz, total = objectness.shape
    picks, lvl = [], []
    off = 0
    for li, n_l in enumerate(per_level):                                   # rpn.py:215-228
        k = min(pre_nms_top_n, n_l)
        picks.append(objectness[:, off:off + n_l].topk(k, dim=1)[1] + off)
        lvl.append(torch.full((k,), li, dtype=torch.int64))
        off += n_l
    top = torch.cat(picks, dim=1)
    lvl = torch.cat(lvl)
    out_b, out_s, out_i = [], [], []
    for b in range(bsz):
        idx = top[b]
        prob = torch.sigmoid(objectness[b, idx])                           # :255
        boxes = decode_single(deltas[b, idx], anchors[idx])
        hh, ww = image_shapes[b]
        boxes = torch.stack((boxes[:, 0].clamp(0, ww), boxes[:, 1].clamp(0, hh),
                             boxes[:, 2].clamp(0, ww), boxes[:, 3].clamp(0, hh)), dim=1)  # :260
        ok = ((boxes[:, 2] - boxes[:, 0]) >= min_size) & ((boxes[:, 3] - boxes[:, 1]) >= min_size)
        ok &= prob >= score_thresh                                         # :263-269
        sel = torch.where(ok)[0]
        boxes, prob, lv, ids = boxes[sel], prob[sel], lvl[sel], idx[sel]
        keep = batched_nms(boxes, prob, lv, nms_thresh, device_type)[:post_nms_top_n]   # :272-275
        out_b.append(boxes[keep]); out_s.append(prob[keep]); out_i.append(ids[keep])
    return out_b, out_s, out_i


def matcher(quality: Tensor, high: float, low: float, allow_low_quality: bool = False) -> Tensor:
    """Matcher.__call__ (_utils.py:271-344) on an ``[M, N]`` quality matrix."""
    vals, matches = quality.max(dim=0)
    every = matches.clone()
    matches[vals < low] = -1
    matches[(vals >= low) & (vals < high)] = -2
    if allow_low_quality:
        row_best = quality.max(dim=1)[0]
        tied = torch.where(quality == row_best[:, None])[1]                # _utils.py:315-344
        matches[tied] = every[tied]
    return matches


# --------------------------------------------------------------------------------------
# ROI-head box post-process  (torchvision_models/tvision/roi_heads.py:715-781)
# --------------------------------------------------------------------------------------
def roi_postprocess(class_logits: Tensor, box_regression: Tensor, proposals: Sequence[Tensor], image_shapes,
                    tfidf, activation: str = "ce", weights=(10.0, 10.0, 5.0, 5.0), score_thresh: float = 0.05,
                    nms_thresh: float = 0.5, detections_per_img: int = 100, strategy: str = "vanilla"):
    """RoIHeads.postprocess_detections; ``strategy`` picks the batched_nms arithmetic explicitly
    ("vanilla" | "coordinate_trick" | "torchvision" = the installed CPU switch at 4000 coordinates).
    Returns per image (boxes, scores, labels, keep) with ``keep`` indexing the filtered candidate list."""
    num_classes = class_logits.shape[-1]
    rows = [len(p) for p in proposals]
    concat = torch.cat(list(proposals), 0)
    pred_boxes = decode_multi(box_regression.reshape(concat.shape[0], -1), concat, weights)       # :721
    pred_boxes = pred_boxes.reshape(concat.shape[0], -1, 4)
    if activation == "ce":
        scores = torch.softmax(tfidf * class_logits, -1)                                           # :725
    elif activation.startswith("gombit"):
        scores = 1 / (torch.exp(torch.exp(-tfidf * (class_logits - 1.96))))                        # :727
    else:
        scores = torch.sigmoid(tfidf * class_logits)                                               # :729
    out = []
    for boxes, sc, shape in zip(pred_boxes.split(rows, 0), scores.split(rows, 0), image_shapes):
        h, w = shape
        bx = boxes[..., 0::2].clamp(min=0, max=w)                                                  # clip_boxes_to_image
        by = boxes[..., 1::2].clamp(min=0, max=h)
        boxes = torch.stack((bx, by), dim=boxes.dim()).reshape(boxes.shape)
        labels = torch.arange(num_classes).view(1, -1).expand_as(sc)
        boxes, sc, labels = boxes[:, 1:].reshape(-1, 4), sc[:, 1:].reshape(-1), labels[:, 1:].reshape(-1)   # :752-759
        inds = torch.nonzero(sc > score_thresh).squeeze(1)                                         # :763
        boxes, sc, labels = boxes[inds], sc[inds], labels[inds]
        ws, hs = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        keep = torch.nonzero((ws >= 1e-2) & (hs >= 1e-2)).squeeze(1)                               # :767
        boxes, sc, labels = boxes[keep], sc[keep], labels[keep]
        if strategy == "vanilla":
            k = batched_nms_vanilla(boxes, sc, labels, nms_thresh)
        elif strategy == "coordinate_trick":
            k = batched_nms_coordinate_trick(boxes, sc, labels, nms_thresh)
        else:
            k = batched_nms(boxes, sc, labels, nms_thresh)
        k = k[:detections_per_img]                                                                 # :774
        out.append((boxes[k], sc[k], labels[k], k))
    return out


def ssd_matcher(quality: Tensor, threshold: float) -> Tensor:
    """SSDMatcher.__call__ (_utils.py:347-361): threshold matching, then every ground truth keeps its best prior."""
    matches = matcher(quality, threshold, threshold, False)
    best_pred = quality.max(dim=1)[1]
    matches[best_pred] = torch.arange(best_pred.size(0), dtype=torch.int64)
    return matches


def retinanet_postprocess(cls_logits: Tensor, bbox_regression: Tensor, anchors: Tensor, level_anchors: Sequence[int],
                          image_shapes, tfidf, score_thresh: float = 0.05, topk_candidates: int = 1000,
                          nms_thresh: float = 0.5, detections_per_img: int = 300, strategy: str = "torchvision"):
    """RetinaNet.postprocess_detections (retinanet.py:414-472), same torch CPU ops in the same order.
    cls_logits [B, sumA, C]; returns per image (boxes, scores, labels)."""
    out = []
    scale = tfidf if tfidf is not None else 1.0
    for b in range(cls_logits.shape[0]):
        ib, isc, il = [], [], []
        for lg, rg, an in zip((cls_logits[b] * scale).split(list(level_anchors), 0), bbox_regression[b].split(list(level_anchors), 0),
                              anchors.split(list(level_anchors), 0)):
            c = lg.shape[-1]
            s = torch.sigmoid(lg).flatten()
            keep = s > score_thresh
            s = s[keep]
            idx = torch.where(keep)[0]
            k = min(topk_candidates, idx.size(0))
            s, order = s.topk(k)
            idx = idx[order]
            a_idx, lab = idx // c, idx % c
            bx = decode_single(rg[a_idx], an[a_idx])
            h, w = image_shapes[b]
            bx = torch.stack((bx[:, 0].clamp(0, w), bx[:, 1].clamp(0, h), bx[:, 2].clamp(0, w), bx[:, 3].clamp(0, h)), 1)
            ib.append(bx); isc.append(s); il.append(lab)
        ib, isc, il = torch.cat(ib), torch.cat(isc), torch.cat(il)
        if strategy == "vanilla":
            keep = batched_nms_vanilla(ib, isc, il, nms_thresh)
        elif strategy == "coordinate_trick":
            keep = batched_nms_coordinate_trick(ib, isc, il, nms_thresh)
        else:
            keep = batched_nms_vanilla(ib, isc, il, nms_thresh) if ib.numel() > 4000 else batched_nms_coordinate_trick(ib, isc, il, nms_thresh)
        keep = keep[:detections_per_img]
        out.append((ib[keep], isc[keep], il[keep]))
    return out


def ssd_postprocess(cls_logits: Tensor, bbox_regression: Tensor, image_anchors: Sequence[Tensor], image_shapes, tfidf,
                    weights=(10.0, 10.0, 5.0, 5.0), score_thresh: float = 0.01, topk_candidates: int = 400,
                    nms_thresh: float = 0.45, detections_per_img: int = 200, strategy: str = "torchvision"):
    """SSD.postprocess_detections (ssd.py:386-430), same torch CPU ops in the same order.  cls_logits [B, A, C];
    returns per image (boxes, scores, labels)."""
    scale = tfidf.unsqueeze(0) if tfidf is not None else 1.0
    pred_scores = torch.softmax(scale * cls_logits, dim=-1)
    num_classes = pred_scores.size(-1)
    out = []
    for boxes, scores, anchors, (h, w) in zip(bbox_regression, pred_scores, image_anchors, image_shapes):
        boxes = decode_single(boxes, anchors, weights)
        boxes = torch.stack((boxes[:, 0].clamp(0, w), boxes[:, 1].clamp(0, h), boxes[:, 2].clamp(0, w), boxes[:, 3].clamp(0, h)), 1)
        ib, isc, il = [], [], []
        for label in range(1, num_classes):
            score = scores[:, label]
            keep = score > score_thresh
            score, box = score[keep], boxes[keep]
            k = min(topk_candidates, score.size(0))
            score, idx = score.topk(k)
            ib.append(box[idx]); isc.append(score); il.append(torch.full_like(score, fill_value=label, dtype=torch.int64))
        ib, isc, il = torch.cat(ib), torch.cat(isc), torch.cat(il)
        if strategy == "vanilla":
            keep = batched_nms_vanilla(ib, isc, il, nms_thresh)
        elif strategy == "coordinate_trick":
            keep = batched_nms_coordinate_trick(ib, isc, il, nms_thresh)
        else:
            keep = batched_nms_vanilla(ib, isc, il, nms_thresh) if ib.numel() > 4000 else batched_nms_coordinate_trick(ib, isc, il, nms_thresh)
        keep = keep[:detections_per_img]
        out.append((ib[keep], isc[keep], il[keep]))
    return out


def clip_boxes_to_image(boxes: Tensor, size: Tuple[int, int]) -> Tensor:
    """torchvision.ops.boxes.clip_boxes_to_image as the reference calls it (rpn.py:260, roi_heads.py:746,
    retinanet.py:452, ssd.py:397): x to [0, width], y to [0, height]."""
    height, width = size
    x = boxes[..., 0::2].clamp(min=0, max=width)
    y = boxes[..., 1::2].clamp(min=0, max=height)
    return torch.stack((x, y), dim=boxes.dim()).reshape(boxes.shape)


def remove_small_boxes(boxes: Tensor, min_size: float) -> Tensor:
    """torchvision.ops.boxes.remove_small_boxes (rpn.py:263, roi_heads.py:767): indices of the boxes whose width and
    height are both >= min_size."""
    ws, hs = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    return torch.where((ws >= min_size) & (hs >= min_size))[0]

