"""Drop-in for the box post-process of ``torchvision_models/tvision/roi_heads.py`` (reference).

``postprocess_detections(self, class_logits, box_regression, proposals, image_shapes)`` has the reference's
signature (roi_heads.py:715) and is written to be bound onto the reference's ``RoIHeads``: it reads
``self.box_coder.weights``, ``self.box_coder.bbox_xform_clip``, ``self.loss_function_name``, ``self.tfidf_post``,
``self.score_thresh``, ``self.nms_thresh`` and ``self.detections_per_img``:

    from object_detectors_b200.tvision import roi_heads as b200_roi
    RoIHeads.postprocess_detections = b200_roi.postprocess_detections

Everything from the class activation to the top-k (roi_heads.py:721-776) runs in libb200det.so
(b200_roi_postprocess: fused score / decode / clip / filter kernel + the shared NMS kernels); the only host
work is slicing the fixed-size outputs into the reference's per-image lists (one count transfer).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .. import ops

Tensor = torch.Tensor


def _activation(name: str) -> int:
    if name == "ce":
        return ops.ROI_SOFTMAX                 # F.softmax(tfidf_post * logits)            (:724-725)
    if name.startswith("gombit"):
        return ops.ROI_GOMBIT                  # 1 / exp(exp(-tfidf_post * (logits-1.96)))  (:726-727)
    return ops.ROI_SIGMOID                     # sigmoid(tfidf_post * logits)              (:728-729)


def postprocess_detections(self, class_logits: Tensor, box_regression: Tensor, proposals: List[Tensor],
                           image_shapes: List[Tuple[int, int]], strategy: str = "torchvision"):
    # "torchvision": batched_nms's own per-image switch between the coordinate trick and the per-class loop
    mode = {"torchvision": ops.NMS_TV_AUTO, "vanilla": ops.NMS_TV_CLASS, "coordinate_trick": ops.NMS_TV_TRICK}[strategy]
    tfidf = getattr(self, "tfidf_post", None)
    kwargs = dict(tfidf=tfidf, activation=_activation(getattr(self, "loss_function_name", "ce")),
                  weights=self.box_coder.weights, xform_clip=self.box_coder.bbox_xform_clip,
                  score_thresh=self.score_thresh, nms_thresh=self.nms_thresh,
                  detections_per_img=self.detections_per_img, nms_mode=mode)
    det, _, dcnt, ccnt, status = ops.roi_postprocess(class_logits, box_regression, proposals, image_shapes, **kwargs)
    counts = dcnt.tolist()                                          # the one sync of the batch
    if int(status.item()) & 1:
        # more candidates than the default slab holds: cand_count has the true numbers, run again with room for them
        need = int(ccnt.max().item())
        det, _, dcnt, ccnt, status = ops.roi_postprocess(class_logits, box_regression, proposals, image_shapes,
                                                         capacity=need, **kwargs)
        counts = dcnt.tolist()
        if int(status.item()) & 1:
            raise RuntimeError("candidate slab overflow in b200_roi_postprocess")
    all_boxes, all_scores, all_labels = [], [], []
    for i, k in enumerate(counts):
        all_boxes.append(det[i, :k, :4])
        all_scores.append(det[i, :k, 4])
        all_labels.append(det[i, :k, 5].to(torch.int64))
    return all_boxes, all_scores, all_labels
