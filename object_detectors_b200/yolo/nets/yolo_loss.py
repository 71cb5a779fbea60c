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
        """Anchor-shape matching of yolo_loss.py:107-161, same return tuple (mask, noobj_mask, tx, ty, tw, th, tconf,
        tcls); ``anchors`` are the head's scaled anchors.  The reference loops over images and ground-truth boxes; here
        every ground-truth box of the batch goes through ONE [M, A] shape-IoU call and the dense target tensors are
        filled by batched scatters (cells shared by several boxes resolve like the reference's index assignments: the last
        box writes the regression targets)."""
        bs, na, dev = len(targets), len(anchors), self.device
        dense = lambda *tail: torch.zeros((bs, na, in_h, in_w) + tail, device=dev)      # noqa: E731
        mask, noobj_mask = dense(), torch.ones((bs, na, in_h, in_w), device=dev)
        tx, ty, tw, th, tconf, tcls = dense(), dense(), dense(), dense(), dense(), dense(self.num_classes)
        counts = [int(t["bbox"].shape[0]) for t in targets]
        if sum(counts) == 0:
            return mask, noobj_mask, tx, ty, tw, th, tconf, tcls
        gt = torch.cat([t["bbox"].reshape(-1, 4) for t in targets]).to(dev).float()
        cls = torch.cat([t["category_id"].reshape(-1) for t in targets]).to(dev).long()
        img = torch.repeat_interleave(torch.arange(bs, device=dev), torch.tensor(counts, device=dev))
        scale = torch.tensor([in_w, in_h, in_w, in_h], dtype=torch.float32, device=dev)
        g = gt * scale                                                     # centre and size in grid units (:121-127)
        gx = g[:, 0].clamp(0, in_w - 1e-4)
        gy = g[:, 1].clamp(0, in_h - 1e-4)
        col, row = gx.long(), gy.long()
        anc = torch.tensor(anchors, dtype=torch.float32, device=dev)
        as_shape = lambda wh: torch.cat([torch.zeros_like(wh), wh], 1)     # noqa: E731  (0, 0, w, h) boxes (:132-138)
        shape_iou = boxes.box_iou(as_shape(g[:, 2:4]), as_shape(anc))      # b200_box_iou, torchvision flavour
        m_ign, a_ign = torch.nonzero(shape_iou > ignore_threshold, as_tuple=True)
        noobj_mask[img[m_ign], a_ign, row[m_ign], col[m_ign]] = 0          # :143-144
        best = shape_iou.max(dim=1)[1]                                     # first maximum (:146)
        mask[img, best, row, col] = 1
        noobj_mask[img, best, row, col] = 0
        tconf[img, best, row, col] = 1
        tcls[img, best, row, col, cls] = 1                                 # one-hot bits accumulate over shared cells
        # Several GT boxes may share a cell; the reference's index assignment (:150-155) keeps the LAST one.  A scatter
        # with duplicate indices is unordered on the GPU, so only the last GT box of every cell writes its regression
        # targets.
        lin = ((img * na + best) * in_h + row) * in_w + col
        order = torch.arange(lin.numel(), device=dev)
        last = torch.full((bs * na * in_h * in_w,), -1, dtype=torch.long, device=dev).scatter_reduce_(0, lin, order, "amax")
        w = last[lin] == order
        at = (img[w], best[w], row[w], col[w])
        tx[at] = (gx - col)[w]
        ty[at] = (gy - row)[w]
        tw[at] = torch.log(g[w, 2] / anc[best[w], 0] + 1e-16)
        th[at] = torch.log(g[w, 3] / anc[best[w], 1] + 1e-16)
        return mask, noobj_mask, tx, ty, tw, th, tconf, tcls
