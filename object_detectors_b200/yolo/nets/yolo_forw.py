"""Drop-in for the inference / target-assignment half of ``yolo/nets/yolo_forw.py`` (reference).

``YOLOForw`` keeps the reference's constructor (a Hydra-style ``config`` with ``.yolo`` and
``.dataset``), ``forward(input, targets=None)``, ``get_target(targets, cxypwh, inw_inh,
ignore_threshold)`` and ``set_img_size``; the arithmetic runs in libb200det.so:

    forward(input)             -> b200_yolo_decode_dense            (yolo_forw.py:81-119,163-176)
    get_target(...)            -> b200_iou_match + tiny gathers      (yolo_forw.py:178-208)
    postprocess(input, ...)    -> b200_yolo_postprocess (fused path used by test_one_epoch)

The loss terms of the training branch (yolo_forw.py:122-160: MSE / focal BCE / CE on gathered rows)
are autograd code that stays in PyTorch in the reference's own class; to accelerate training bind
``get_target`` onto it.  Calling ``forward`` with targets here raises.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from ... import ops


def _cfg_get(cfg, key, default=None):
    try:
        return cfg[key]
    except Exception:
        return getattr(cfg, key, default)


def load_idf_column(dset_name: str, variant: str = "smooth", root: Optional[str] = None) -> torch.Tensor:
    """The reference reads ``$owd/<dset>_files/idf.csv`` (custom.py:166-176,249-254)."""
    root = root or os.getenv("owd")
    if not root:
        raise RuntimeError("tfidf weights requested but the 'owd' environment variable (reference yolo dir) is unset")
    path = os.path.join(root, f"{dset_name}_files", "idf.csv")
    import csv
    with open(path, newline="") as fh:
        rows = list(csv.DictReader(fh))
    return torch.tensor([float(r[variant]) for r in rows], dtype=torch.float32)


class YOLOForw(nn.Module):
    def __init__(self, config, idf_logits: Optional[torch.Tensor] = None):
        super().__init__()
        cfg = config.yolo if hasattr(config, "yolo") else config["yolo"]
        dataset = config.dataset if hasattr(config, "dataset") else config["dataset"]
        self.anchors = [[tuple(a) for a in scale] for scale in _cfg_get(dataset, "anchors")]
        self.num_anchors = len(self.anchors)
        self.num_classes = int(_cfg_get(cfg, "classes"))
        self.bbox_attrs = 5 + self.num_classes
        self.img_size = _cfg_get(cfg, "img_size")
        self.ignore_threshold = _cfg_get(cfg, "ignore_threshold", 0.5)
        self.iou_type = _cfg_get(cfg, "iou_type", 1)
        self.softmax = _cfg_get(cfg, "class_loss", 1) == 1        # CrossEntropyLoss -> softmax (yolo_forw.py:168)
        self.tfidf_norm = _cfg_get(cfg, "tfidf_norm", 0)
        self.device = torch.device("cuda")
        tfidf = _cfg_get(cfg, "tfidf", [0, 0])
        self.idf_logits = None                                     # == torch.tensor(1) in the reference (:38)
        if idf_logits is not None:
            self.idf_logits = torch.as_tensor(idf_logits, dtype=torch.float32)
        elif tfidf[1] == 1:
            self.idf_logits = load_idf_column(_cfg_get(dataset, "dset_name"), _cfg_get(cfg, "tfidf_variant", "smooth"))
        if self.idf_logits is not None:
            if self.tfidf_norm != 0:
                self.idf_logits = self.idf_logits / torch.norm(self.idf_logits, p=self.tfidf_norm)   # :63-67
            self.idf_logits = self.idf_logits.to(self.device)
        self._table_cache = {}
        self._plans = {}

    # ---------------------------------------------------------------------------------- inference
    def forward(self, input, targets=None):
        if targets is not None:
            raise NotImplementedError(
                "the training branch (loss terms) stays in the reference's YOLOForw; bind "
                "object_detectors_b200's get_target onto it instead (see INTEGRATION.md)")
        heads = [t.to(self.device, non_blocking=True).float() for t in input]
        return ops.yolo_decode_dense(heads, self.anchors, self.img_size, self.num_classes, self.idf_logits,
                                     self.softmax)

    def postprocess(self, input, conf_thr: float = 0.1, nms_thr: float = 0.6, nms_mode: int = ops.NMS_MAJORITY,
                    capacity: Optional[int] = None, max_det: Optional[int] = None):
        """Fused decode -> xyxy -> score filter -> NMS (what test_one_epoch.py:22-36 computes).

        The plan (outputs + workspace) is cached per (grids, batch, capacity, ...) on the module, so a steady stream of
        batches allocates nothing.  If the candidate slab or ``max_det`` overflows, the call is repeated once with
        exactly what the batch needs -- the kernels report the TRUE per-image candidate counts -- rounded up to a
        multiple of 256, not with the worst case N (whose NMS scratch grows with capacity^2 / 8 bytes per image)."""
        heads = [t.to(self.device, non_blocking=True).float().contiguous() for t in input]
        grids = tuple(int(h.shape[2]) for h in heads)
        batch = int(heads[0].shape[0])
        n = sum(g * g * len(self.anchors[0]) for g in grids)

        def plan_for(cap, md):
            key = (grids, batch, cap, md, float(conf_thr), float(nms_thr), int(nms_mode), float(self.img_size))
            plan = self._plans.get(key)
            if plan is None:
                if len(self._plans) >= 4:                 # a module sees few distinct shapes; do not hoard workspaces
                    self._plans.pop(next(iter(self._plans)))
                plan = ops.YoloPostprocess(list(grids), batch, self.anchors, self.img_size, self.num_classes, self.softmax,
                                           conf_thr, nms_thr, nms_mode, cap, md, self.device)
                self._plans[key] = plan
            return plan

        cap = int(capacity or min(n, ops.DEFAULT_CAPACITY))
        md = int(max_det or cap)
        plan = plan_for(cap, md)
        out = plan(heads, self.idf_logits)
        st = int(plan.status.item())
        if st:
            plan.status.zero_()
            need = int(plan.cand_count.max().item())
            cap = min(n, max(cap, -(-need // 256) * 256))
            md = cap if (st & 2) or max_det is None else md
            plan = plan_for(cap, md)
            out = plan(heads, self.idf_logits)
            plan.check_status()
        return out

    # ---------------------------------------------------------------------------- target assignment
    def grid_table(self, grid_sizes):
        """``cxypwh [N,4]`` and ``inw_inh [N]`` exactly as forward() builds them (yolo_forw.py:93-119)."""
        key = (tuple(grid_sizes), float(self.img_size))
        hit = self._table_cache.get(key)
        if hit is not None:
            return hit
        rows, widths = [], []
        for k, g in enumerate(grid_sizes):
            a = len(self.anchors[k])
            stride = self.img_size / g
            scaled = torch.tensor([(aw / stride, ah / stride) for aw, ah in self.anchors[k]], dtype=torch.float32)
            col = (torch.arange(g, dtype=torch.float32) + 0.5) / g
            gx = col.view(1, g, 1).expand(g, g, a)
            gy = col.view(g, 1, 1).expand(g, g, a)
            aw = (scaled[:, 0] / g).view(1, 1, a).expand(g, g, a)
            ah = (scaled[:, 1] / g).view(1, 1, a).expand(g, g, a)
            rows.append(torch.stack((gx, gy, aw, ah), dim=-1).reshape(-1, 4))
            widths.append(torch.full((g * g * a,), float(g), dtype=torch.float32))
        out = (torch.cat(rows, 0).contiguous().to(self.device), torch.cat(widths, 0).to(self.device))
        self._table_cache[key] = out
        return out

    def get_target(self, targets, cxypwh, inw_inh, ignore_threshold=0.5):
        """Same contract as the reference: ``(tgt [sumM,4], tcls [sumM,C], obj_mask list of int64 [M],
        noobj_mask bool [B,N])``.  The [M,N] IoU matrices are never materialised."""
        b = len(targets)
        counts = [int(t["bbox"].shape[0]) for t in targets]
        mmax = max(max(counts), 1)
        gt = torch.zeros((b, mmax, 4), dtype=torch.float32, device=self.device)
        for i, t in enumerate(targets):
            gt[i, :counts[i]] = t["bbox"].to(self.device, torch.float32)
        cnt = torch.tensor(counts, dtype=torch.int32, device=self.device)
        cxypwh = cxypwh.to(self.device, torch.float32).contiguous()
        inw_inh = inw_inh.to(self.device, torch.float32)
        best, noobj = ops.iou_match(gt, cnt, cxypwh, self.iou_type if self.iou_type in (1, 2, 3) else 0,
                                    ignore_threshold)
        tgt, tcls, obj = [], [], []
        for i, t in enumerate(targets):
            m = counts[i]
            idx = best[i, :m]
            box = gt[i, :m]
            tcls.append(torch.nn.functional.one_hot(t["category_id"].to(self.device), self.num_classes).float())
            anchor = cxypwh[idx]
            width = inw_inh[idx]
            px, py = box[:, 0] * width, box[:, 1] * width
            gx = torch.clamp(px - px.long(), 0.0001, 0.9999)                  # yolo_forw.py:191-194
            gy = torch.clamp(py - py.long(), 0.0001, 0.9999)
            gw = torch.log(box[:, 2] / anchor[:, 2] + 1e-16)                  # :196-197
            gh = torch.log(box[:, 3] / anchor[:, 3] + 1e-16)
            tgt.append(torch.stack((gx, gy, gw, gh), dim=1))
            obj.append(idx)
        return torch.cat(tgt, 0), torch.cat(tcls, 0), obj, noobj

    def set_img_size(self, img_size):
        self.img_size = img_size
