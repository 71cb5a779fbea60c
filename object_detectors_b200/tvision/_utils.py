"""Drop-ins for ``torchvision_models/tvision/_utils.py`` (reference): ``BoxCoder.decode`` /
``decode_single`` (:168-223) and ``Matcher.__call__`` (:271-344) with the reference's constructor
arguments and return layouts, plus ``BoxCoder.encode`` / ``encode_single`` (:80-125,144-166), backed by
libb200det.so.  The samplers (``randperm``-based training glue) are not provided."""
from __future__ import annotations

import math
from typing import List, Tuple

import torch

from .. import ops

Tensor = torch.Tensor


class BoxCoder:
    def __init__(self, weights: Tuple[float, float, float, float], bbox_xform_clip: float = math.log(1000.0 / 16)):
        self.weights = weights
        self.bbox_xform_clip = bbox_xform_clip

    def encode(self, reference_boxes: List[Tensor], proposals: List[Tensor]):
        boxes_per_image = [len(b) for b in reference_boxes]
        targets = self.encode_single(torch.cat(reference_boxes, dim=0), torch.cat(proposals, dim=0))
        return targets.split(boxes_per_image, 0)

    def encode_single(self, reference_boxes: Tensor, proposals: Tensor) -> Tensor:
        return ops.boxcoder_encode(reference_boxes, proposals, self.weights)

    def decode(self, rel_codes: Tensor, boxes: List[Tensor]) -> Tensor:
        assert isinstance(boxes, (list, tuple))
        concat = torch.cat(boxes, dim=0)
        total = concat.shape[0]
        if total > 0:
            rel_codes = rel_codes.reshape(total, -1)
        pred = self.decode_single(rel_codes, concat)
        if total > 0:
            pred = pred.reshape(total, -1, 4)
        return pred

    def decode_single(self, rel_codes: Tensor, boxes: Tensor) -> Tensor:
        return ops.boxcoder_decode(rel_codes.float(), boxes.to(rel_codes.dtype), self.weights, self.bbox_xform_clip)


class Matcher:
    BELOW_LOW_THRESHOLD = -1
    BETWEEN_THRESHOLDS = -2

    def __init__(self, high_threshold: float, low_threshold: float, allow_low_quality_matches: bool = False):
        assert low_threshold <= high_threshold
        self.high_threshold = high_threshold
        self.low_threshold = low_threshold
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, match_quality_matrix: Tensor) -> Tensor:
        if match_quality_matrix.numel() == 0:
            if match_quality_matrix.shape[0] == 0:
                raise ValueError("No ground-truth boxes available for one of the images during training")
            raise ValueError("No proposal boxes available for one of the images during training")
        return ops.matcher(match_quality_matrix.float(), self.high_threshold, self.low_threshold,
                           self.allow_low_quality_matches)

    def match_boxes(self, gt_boxes: Tensor, boxes: Tensor) -> Tensor:
        """``self(box_ops.box_iou(gt_boxes, boxes))`` without the [M, N] matrix (rpn.py:192-193, roi_heads.py:633-634,
        retinanet.py:409-410): IoU, column max, thresholds and the low-quality restore in two fused kernels."""
        return ops.match_boxes(gt_boxes, boxes, self.high_threshold, self.low_threshold, self.allow_low_quality_matches)


class SSDMatcher(Matcher):
    """``_utils.py:347-361``: threshold matching plus "every ground truth keeps its best prior"."""

    def __init__(self, threshold: float):
        super().__init__(threshold, threshold, allow_low_quality_matches=False)

    def __call__(self, match_quality_matrix: Tensor) -> Tensor:
        if match_quality_matrix.numel() == 0:
            return super().__call__(match_quality_matrix)       # raises like the reference
        return ops.matcher_ssd(match_quality_matrix.float(), self.high_threshold)

    def match_boxes(self, gt_boxes: Tensor, boxes: Tensor) -> Tensor:
        """``self(box_ops.box_iou(gt_boxes, boxes))`` (ssd.py:371-372) without the matrix."""
        return ops.match_boxes(gt_boxes, boxes, self.high_threshold, self.low_threshold, False, ssd=True)
