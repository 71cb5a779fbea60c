"""Drop-in for the proposal filter of ``torchvision_models/tvision/rpn.py`` (reference).

``filter_proposals(self, proposals, objectness, image_shapes, num_anchors_per_level)`` has the
reference's signature (rpn.py:230) and is written to be bound onto the reference's
``RegionProposalNetwork`` (it reads ``self.pre_nms_top_n()``, ``self.post_nms_top_n()``,
``self.nms_thresh``, ``self.score_thresh``, ``self.min_size``):

    from object_detectors_b200.tvision import rpn as b200_rpn
    RegionProposalNetwork.filter_proposals = b200_rpn.filter_proposals

``_get_top_n_idx(self, objectness, num_anchors_per_level)`` (rpn.py:215) is the per-level top-k index list alone.
``filter_from_deltas`` is the fused variant that also replaces ``box_coder.decode`` (rpn.py:355) by
decoding only the per-level top-k anchors.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from .. import ops

Tensor = torch.Tensor


def _strategy_mode(num_boxes: int, strategy: str) -> int:
    if strategy == "torchvision":      # torchvision 0.26 batched_nms switch on CUDA
        strategy = "vanilla" if num_boxes * 4 > 100_000 else "coordinate_trick"
    return ops.NMS_TV_CLASS if strategy == "vanilla" else ops.NMS_TV_TRICK


def _split(boxes: Tensor, scores: Tensor, count: Tensor) -> Tuple[List[Tensor], List[Tensor]]:
    counts = count.tolist()
    return [boxes[i, :k] for i, k in enumerate(counts)], [scores[i, :k] for i, k in enumerate(counts)]


def filter_proposals(self, proposals: Tensor, objectness: Tensor, image_shapes: Sequence[Tuple[int, int]],
                     num_anchors_per_level: Sequence[int], strategy: str = "torchvision"):
    num_images = proposals.shape[0]
    objectness = objectness.detach().reshape(num_images, -1).float()
    pre, post = self.pre_nms_top_n(), self.post_nms_top_n()
    hw = torch.tensor([[float(h), float(w)] for h, w in image_shapes], dtype=torch.float32, device=proposals.device)
    kept_per_image = sum(min(pre, n) for n in num_anchors_per_level)
    boxes, scores, _, count = ops.rpn_filter_proposals(
        objectness, proposals.detach().float(), list(num_anchors_per_level), hw, pre, post, self.nms_thresh,
        self.score_thresh, self.min_size, _strategy_mode(kept_per_image, strategy))
    return _split(boxes, scores, count)


def _get_top_n_idx(self, objectness: Tensor, num_anchors_per_level: Sequence[int]) -> Tensor:
    """``RegionProposalNetwork._get_top_n_idx(objectness, num_anchors_per_level)`` (rpn.py:215-228): per level the
    indices of the ``min(pre_nms_top_n, n_l)`` largest logits in descending order, offset by the level start,
    concatenated over the levels -> int64 ``[B, sum_l k_l]``.  One cluster-select kernel instead of a ``split`` /
    ``topk`` / ``cat`` per level; bind it like ``filter_proposals``."""
    return ops.rpn_top_n_idx(objectness.detach().float(), list(num_anchors_per_level), self.pre_nms_top_n())


def filter_from_deltas(objectness: Tensor, pred_bbox_deltas: Tensor, anchors: Tensor,
                       image_shapes: Sequence[Tuple[int, int]], num_anchors_per_level: Sequence[int],
                       pre_nms_top_n: int, post_nms_top_n: int, nms_thresh: float = 0.7, score_thresh: float = 0.0,
                       min_size: float = 1e-3, strategy: str = "vanilla"):
    """objectness [B, sumA], pred_bbox_deltas [B, sumA, 4], anchors [sumA, 4] (identical for all images)."""
    hw = torch.tensor([[float(h), float(w)] for h, w in image_shapes], dtype=torch.float32, device=objectness.device)
    kept_per_image = sum(min(pre_nms_top_n, n) for n in num_anchors_per_level)
    boxes, scores, index, count = ops.rpn_filter(objectness.float(), pred_bbox_deltas.float(), anchors.float(),
                                                 list(num_anchors_per_level), hw, pre_nms_top_n, post_nms_top_n,
                                                 nms_thresh, score_thresh, min_size,
                                                 _strategy_mode(kept_per_image, strategy))
    b, s = _split(boxes, scores, count)
    return b, s, [index[i, :k] for i, k in enumerate(count.tolist())]
