"""Drop-in for the ``torchvision.ops.boxes`` functions the reference calls on its hot path
(``from torchvision.ops import boxes`` at yolo_forw.py:9, test_one_epoch.py:3, telemetry.py:4;
``from torchvision.ops import boxes as box_ops`` at tvision/rpn.py:7, roi_heads.py:8):

    nms(boxes, scores, iou_threshold)                       -> int64 kept indices, descending score
    batched_nms(boxes, scores, idxs, iou_threshold)         -> same, no suppression across idxs
    box_iou(boxes1, boxes2)                                 -> [N, M]
    clip_boxes_to_image / remove_small_boxes                -> element-wise, kept in torch

``batched_nms`` reproduces torchvision 0.26's own strategy switch (coordinate trick up to 100 000
coordinates on CUDA, per-class "vanilla" above) unless ``strategy`` says otherwise -- the two
strategies round differently and disagree on ~1 % of inputs (hard part 2).
CUDA tensors only.
"""
from __future__ import annotations

from typing import Tuple

import torch

from .. import ops

Tensor = torch.Tensor


def _single_segment(n: int, device) -> Tensor:
    return torch.tensor([0, n], dtype=torch.int32, device=device)


def nms(boxes: Tensor, scores: Tensor, iou_threshold: float) -> Tensor:
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    keep, cnt, _ = ops.nms_segments(boxes.float(), scores.float(), None, _single_segment(n, boxes.device),
                                    iou_threshold, ops.NMS_TV, n)
    return keep[:int(cnt[0])]


def batched_nms(boxes: Tensor, scores: Tensor, idxs: Tensor, iou_threshold: float, strategy: str = "torchvision") -> Tensor:
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    if strategy == "torchvision":
        strategy = "vanilla" if boxes.numel() > 100_000 else "coordinate_trick"
    mode = ops.NMS_TV_CLASS if strategy == "vanilla" else ops.NMS_TV_TRICK
    keep, cnt, _ = ops.nms_segments(boxes.float(), scores.float(), idxs, _single_segment(n, boxes.device),
                                    iou_threshold, mode, n)
    return keep[:int(cnt[0])]


def box_iou(boxes1: Tensor, boxes2: Tensor) -> Tensor:
    return ops.box_iou(boxes1.float(), boxes2.float(), ops.IOU_TV, xcycwh=False)


def clip_boxes_to_image(boxes: Tensor, size: Tuple[int, int]) -> Tensor:
    if boxes.numel() == 0:
        return boxes.clone()
    return ops.clip_boxes_to_image(boxes.float(), size).to(boxes.dtype)


def remove_small_boxes(boxes: Tensor, min_size: float) -> Tensor:
    if boxes.shape[0] == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    return ops.remove_small_boxes(boxes.float(), min_size)
