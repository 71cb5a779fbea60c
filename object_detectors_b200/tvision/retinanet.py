"""Drop-in for ``RetinaNet.postprocess_detections`` of ``torchvision_models/tvision/retinanet.py:414-472``
(reference): same signature and return value, written to be bound onto the reference's ``RetinaNet`` (it reads
``self.tfidf_post``, ``self.score_thresh``, ``self.topk_candidates``, ``self.nms_thresh``, ``self.detections_per_img``):

    from object_detectors_b200.tvision import retinanet as b200_retina
    RetinaNet.postprocess_detections = b200_retina.postprocess_detections

The reference loops over images and levels (sigmoid of every logit, boolean mask, ``topk``, decode, ``batched_nms``);
here the whole batch is one sliced top-k select per level over the raw logits (only the <= topk survivors are scored,
decoded and clipped), one class-aware NMS per image and one merge -- ``b200_retinanet_postprocess``.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from .. import ops

Tensor = torch.Tensor


def _strategy_mode(strategy: str) -> int:
    return {"vanilla": ops.NMS_TV_CLASS, "coordinate_trick": ops.NMS_TV_TRICK, "torchvision": ops.NMS_TV_AUTO}[strategy]


def postprocess_detections(self, head_outputs: Dict[str, List[Tensor]], anchors: List[List[Tensor]],
                           image_shapes: List[Tuple[int, int]], strategy: str = "torchvision") -> List[Dict[str, Tensor]]:
    cls_levels, reg_levels = head_outputs["cls_logits"], head_outputs["bbox_regression"]
    level_anchors = [int(t.shape[1]) for t in cls_levels]
    logits = torch.cat([t.detach().float() for t in cls_levels], dim=1).contiguous()          # [B, sumA, C]
    regs = torch.cat([t.detach().float() for t in reg_levels], dim=1).contiguous()            # [B, sumA, 4]
    anc = torch.cat([a.float() for a in anchors[0]], dim=0).contiguous()                      # identical for every image
    tfidf = getattr(self, "tfidf_post", None)
    boxes, scores, labels, count = ops.retinanet_postprocess(
        logits, regs, anc, level_anchors, image_shapes, tfidf, self.score_thresh, self.topk_candidates, self.nms_thresh,
        self.detections_per_img, _strategy_mode(strategy))
    out = []
    for i, k in enumerate(count.tolist()):
        out.append({"boxes": boxes[i, :k], "scores": scores[i, :k], "labels": labels[i, :k].long()})
    return out
