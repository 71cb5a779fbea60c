"""Drop-in for the inference branch of the legacy per-head layer ``yolo/nets/yolo_loss.py`` (reference; still
used by the stale ``yolo/benchmark.py:63`` and copied into ``utilities/telemetry.py:46-92``).

``YOLOLoss(cfg, head)`` keeps the reference's constructor (``cfg['anchors'][head]``, ``cfg['classes']``,
``cfg['img_size']``, ...) and ``forward(input, targets=None)`` returns the reference's ``[B, A*H*W, 5+C]`` tensor
(rows ordered (a, h, w); xywh in pixels, sigmoid objectness and classes) from one transposing kernel
(b200_yolo_legacy_decode).  ``get_target`` is the reference's anchor-shape matching (:107-161): its only array
arithmetic is ``boxes.box_iou`` of [M, 4] x [A, 4], served by the drop-in; the scatter into the dense target
tensors is index glue kept in torch.  The loss branch is autograd code that stays in the reference class.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ...tvision import boxes


class YOLOLoss(nn.Module):
    def __init__(self, cfg, head):
        super().__init__()
        self.anchors = cfg["anchors"][head]
        self.num_anchors = len(self.anchors)
        self.num_classes = cfg["classes"]
        self.bbox_attrs = 5 + self.num_classes
        self.img_size = cfg["img_size"]
        self.ignore_threshold = cfg.get("ignore_threshold", 0.5) if hasattr(cfg, "get") else cfg["ignore_threshold"]
        self.device = torch.device("cuda")

    def forward(self, input, targets=None):
        if targets is not None:
            raise NotImplementedError("the loss branch (yolo_loss.py:52-73) is autograd code: keep the reference "
                                      "class for training and bind get_target onto it (INTEGRATION.md)")
        return ops.yolo_legacy_decode(input.to(self.device).float().contiguous(), self.anchors, self.num_classes,
                                      self.img_size)

    def get_target(self, targets, anchors, in_w, in_h, ignore_threshold=0.5):
        """yolo_loss.py:107-161 with the same return tuple; ``anchors`` are the scaled anchors of the head."""
        bs, na, dev = len(targets), len(anchors), self.device
        z = lambda *shape: torch.zeros(*shape, device=dev)      # noqa: E731
        mask, noobj_mask = z(bs, na, in_h, in_w), torch.ones(bs, na, in_h, in_w, device=dev)
        tx, ty, tw, th, tconf = (z(bs, na, in_h, in_w) for _ in range(5))
        tcls = z(bs, na, in_h, in_w, self.num_classes)
        anc = torch.tensor(anchors, dtype=torch.float32, device=dev)
        anchor_shapes = torch.cat([torch.zeros((na, 2), device=dev), anc], 1)
        for b, target in enumerate(targets):
            bbox = target["bbox"].to(dev)
            categories = target["category_id"].to(dev)
            if bbox.shape[0] == 0:
                continue
            gx = torch.clamp(bbox[:, 0] * in_w, 0, in_w - 1e-4)
            gy = torch.clamp(bbox[:, 1] * in_h, 0, in_h - 1e-4)
            gw, gh = bbox[:, 2] * in_w, bbox[:, 3] * in_h
            gi, gj = gx.long(), gy.long()
            gt_box = torch.zeros(bbox.shape, dtype=torch.float32, device=dev)
            gt_box[:, 2], gt_box[:, 3] = gw, gh
            anch_ious = boxes.box_iou(gt_box, anchor_shapes)              # b200_box_iou (torchvision flavour)
            over = anch_ious > ignore_threshold                           # [M, A]
            m_idx, a_idx = torch.nonzero(over, as_tuple=True)
            noobj_mask[b, a_idx, gj[m_idx], gi[m_idx]] = 0
            best_n = torch.max(anch_ious, axis=1)[1]
            mask[b, best_n, gj, gi] = 1
            noobj_mask[b, best_n, gj, gi] = 0
            tx[b, best_n, gj, gi] = gx - gi
            ty[b, best_n, gj, gi] = gy - gj
            tw[b, best_n, gj, gi] = torch.log(gw / anc[best_n][:, 0] + 1e-16)
            th[b, best_n, gj, gi] = torch.log(gh / anc[best_n][:, 1] + 1e-16)
            tconf[b, best_n, gj, gi] = 1
            tcls[b, best_n, gj, gi, categories] = 1
        return mask, noobj_mask, tx, ty, tw, th, tconf, tcls
