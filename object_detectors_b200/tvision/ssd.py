"""Drop-in for ``SSD.postprocess_detections`` of ``torchvision_models/tvision/ssd.py:386-430`` (reference): same
signature and return value, written to be bound onto the reference's ``SSD`` (it reads ``self.tfidf_post``,
``self.box_coder.weights``, ``self.score_thresh``, ``self.topk_candidates``, ``self.nms_thresh``,
``self.detections_per_img``):

    from object_detectors_b200.tvision import ssd as b200_ssd
    SSD.postprocess_detections = b200_ssd.postprocess_detections

The reference materialises the [A, C] softmax, decodes all anchors and loops over the classes (mask, ``topk``,
``full_like``) per image; here one kernel per batch computes the softmax statistics per anchor, decodes an anchor only
if one of its classes passes the threshold and compacts the candidates, one kernel enforces the per-class top-k, and
the shared class-aware NMS emits the first ``detections_per_img`` -- ``b200_ssd_postprocess``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .. import ops

Tensor = torch.Tensor


def postprocess_detections(self, head_outputs: Dict[str, Tensor], image_anchors: List[Tensor],
                           image_shapes: List[Tuple[int, int]], strategy: str = "torchvision") -> List[Dict[str, Tensor]]:
    mode = {"vanilla": ops.NMS_TV_CLASS, "coordinate_trick": ops.NMS_TV_TRICK, "torchvision": ops.NMS_TV_AUTO}[strategy]
    logits, regs = head_outputs["cls_logits"].detach().float(), head_outputs["bbox_regression"].detach().float()
    coder = getattr(self, "box_coder", None)
    weights = getattr(coder, "weights", (10.0, 10.0, 5.0, 5.0))
    clip = getattr(coder, "bbox_xform_clip", math.log(1000.0 / 16))
    a, c = logits.shape[1], logits.shape[2]
    cap = None
    for attempt in range(3):
        det, dcnt, ccnt, status = ops.ssd_postprocess(logits, regs, image_anchors, image_shapes, getattr(self, "tfidf_post", None),
                                                      weights, clip, self.score_thresh, self.topk_candidates, self.nms_thresh,
                                                      self.detections_per_img, mode, cap)
        if not (int(status.item()) & 1):
            break
        # slab overflow: repeat once with what the batch needs (the kernels report the true candidate counts), then
        # with the worst case; never loop
        cap = a * (c - 1) if attempt else min(a * (c - 1), -(-int(ccnt.max().item()) // 256) * 256)
    else:
        raise RuntimeError("SSD candidate slab overflow persists at the worst-case capacity")
    out = []
    for i, k in enumerate(dcnt.tolist()):
        out.append({"boxes": det[i, :k, :4], "scores": det[i, :k, 4], "labels": det[i, :k, 5].long()})
    return out
