#!/usr/bin/env python
"""Golden fixtures for the rows added after the first batch (same rules as make_golden.py: the UNMODIFIED
reference under /root/reference is run here, in the build container, on seeded synthetic inputs; only the
outputs are committed).

  legacy_yolo_loss.npz : YOLOLoss.forward(input) of yolo/nets/yolo_loss.py (inference branch), two heads
  legacy_get_target.npz: YOLOLoss.get_target of yolo/nets/yolo_loss.py:107-161 (mask, noobj_mask, tx, ty, tw, th, tconf, tcls)
  retinanet_postprocess.npz : RetinaNet.postprocess_detections of torchvision_models/tvision/retinanet.py:414-472
  ssd_postprocess.npz  : SSD.postprocess_detections of torchvision_models/tvision/ssd.py:386-430
  roi_postprocess.npz  : RoIHeads.postprocess_detections of torchvision_models/tvision/roi_heads.py for the three
                         activations (ce / gombit / sigmoid), tfidf on, COCO-91 shaped head

    python tests/golden/make_golden_extra.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torchvision  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from object_detectors_b200 import synthetic as syn  # noqa: E402
import make_golden as mg  # noqa: E402  (shim: stubs + torch proxy)


def legacy():
    mg._install()
    from nets import yolo_loss
    yolo_loss.torch = mg._TorchProxy()
    pack = {}
    for tag, (grid, head_idx, classes, img) in {"h13": (13, 0, 80, 416), "h38": (38, 1, 20, 608)}.items():
        cfg = dict(anchors=[[list(a) for a in s] for s in syn.COCO_ANCHORS], classes=classes, img_size=img,
                   ignore_threshold=0.5, lambda_xy=1, lambda_wh=1, lambda_conf=1, lambda_no_conf=1, lambda_cls=1)
        layer = yolo_loss.YOLOLoss(cfg, head_idx)
        x = syn.legacy_head(77 + grid, 2, 3, classes, grid)
        out = layer(torch.from_numpy(x))
        rows = torch.arange(0, out.shape[1], 7)
        pack[f"{tag}_rows"] = rows.numpy()
        pack[f"{tag}_sample"] = out[:, rows].numpy()
        pack[f"{tag}_colsum"] = out.double().sum(dim=1).numpy()
        pack[f"{tag}_args"] = np.array([77 + grid, grid, head_idx, classes, img])
    np.savez_compressed(os.path.join(HERE, "legacy_yolo_loss.npz"), **pack)


def legacy_get_target():
    """The reference's anchor-shape matching on seeded ground truth (two heads, incl. duplicate cells)."""
    mg._install()
    from nets import yolo_loss
    yolo_loss.torch = mg._TorchProxy()
    pack = {}
    for tag, (grid, head_idx, img, seed, bsz, max_gt) in {"h26": (26, 1, 416, 33, 2, 12), "h19": (19, 0, 608, 34, 3, 40)}.items():
        cfg = dict(anchors=[[list(a) for a in s] for s in syn.COCO_ANCHORS], classes=80, img_size=img,
                   ignore_threshold=0.5, lambda_xy=1, lambda_wh=1, lambda_conf=1, lambda_no_conf=1, lambda_cls=1)
        layer = yolo_loss.YOLOLoss(cfg, head_idx)
        targets = [{k: torch.from_numpy(v) for k, v in t.items()} for t in syn.gt_targets(seed, bsz, 80, max_gt=max_gt)]
        stride = img / grid
        scaled = [(a_w / stride, a_h / stride) for a_w, a_h in syn.COCO_ANCHORS[head_idx]]
        out = layer.get_target(targets, scaled, grid, grid, 0.5)
        for name, t in zip(("mask", "noobj", "tx", "ty", "tw", "th", "tconf"), out[:7]):
            pack[f"{tag}_{name}"] = t.numpy()
        pack[f"{tag}_tcls_idx"] = torch.nonzero(out[7]).numpy().astype(np.int32)      # one-hot tensor, stored sparse
        pack[f"{tag}_args"] = np.array([grid, head_idx, img, seed, bsz, max_gt])
    np.savez_compressed(os.path.join(HERE, "legacy_get_target.npz"), **pack)


def roi():
    sys.path.insert(0, os.path.join(REF, "torchvision_models"))
    from tvision import roi_heads as ref_roi
    from torchvision.models.detection import _utils as det_utils
    pack = {}
    C = 91
    idf = torch.from_numpy(np.linspace(0.6, 1.8, C).astype(np.float32))
    for tag, loss_name in (("ce", "ce"), ("gombit", "gombit_x"), ("sigmoid", "bce")):
        rows = [300, 257]
        logits, regs, props = syn.roi_inputs(51, rows, C, 800, 1216)
        r = ref_roi.RoIHeads.__new__(ref_roi.RoIHeads)
        torch.nn.Module.__init__(r)
        r.box_coder = det_utils.BoxCoder((10.0, 10.0, 5.0, 5.0))
        r.loss_function_name = loss_name
        r.tfidf_post = idf.clone()
        r.score_thresh, r.nms_thresh, r.detections_per_img = 0.05, 0.5, 100
        b, s, l = r.postprocess_detections(torch.from_numpy(logits), torch.from_numpy(regs),
                                           [torch.from_numpy(p) for p in props], [(800, 1216), (800, 1216)])
        for i in range(len(rows)):
            pack[f"{tag}_boxes_{i}"] = b[i].numpy()
            pack[f"{tag}_scores_{i}"] = s[i].numpy()
            pack[f"{tag}_labels_{i}"] = l[i].numpy()
    pack["args"] = np.array([51, 300, 257, C, 800, 1216])
    pack["idf"] = idf.numpy()
    np.savez_compressed(os.path.join(HERE, "roi_postprocess.npz"), **pack)


def retinanet():
    """The reference's RetinaNet post-process on seeded head outputs (two images, five levels, tfidf on)."""
    sys.path.insert(0, os.path.join(REF, "torchvision_models"))
    from tvision import retinanet as ref_retina
    from torchvision.models.detection import _utils as det_utils
    C = 91
    r = ref_retina.RetinaNet.__new__(ref_retina.RetinaNet)
    torch.nn.Module.__init__(r)
    r.tfidf_post = torch.from_numpy(np.linspace(0.6, 1.8, C).astype(np.float32))
    r.score_thresh, r.topk_candidates, r.nms_thresh, r.detections_per_img = 0.05, 1000, 0.5, 300
    r.box_coder = det_utils.BoxCoder(weights=(1.0, 1.0, 1.0, 1.0))
    logits, regs, anchors, per_level = syn.retina_inputs(61, 2, 256, 320, C)
    tl, tr, ta = torch.from_numpy(logits), torch.from_numpy(regs), torch.from_numpy(anchors)
    head = {"cls_logits": list(tl.split(per_level, 1)), "bbox_regression": list(tr.split(per_level, 1))}
    out = r.postprocess_detections(head, [list(ta.split(per_level, 0)) for _ in range(2)], [(256, 320)] * 2)
    pack = {"args": np.array([61, 2, 256, 320, C]), "idf": r.tfidf_post.numpy()}
    for i, o in enumerate(out):
        pack[f"boxes_{i}"] = o["boxes"].numpy()
        pack[f"scores_{i}"] = o["scores"].numpy()
        pack[f"labels_{i}"] = o["labels"].numpy()
    np.savez_compressed(os.path.join(HERE, "retinanet_postprocess.npz"), **pack)


def ssd():
    """The reference's SSD post-process on seeded head outputs (two images, 91 classes, tfidf on, one class over the
    400-candidate limit)."""
    sys.path.insert(0, os.path.join(REF, "torchvision_models"))
    from tvision import ssd as ref_ssd
    from torchvision.models.detection import _utils as det_utils
    C, A = 91, 3000
    r = ref_ssd.SSD.__new__(ref_ssd.SSD)
    torch.nn.Module.__init__(r)
    r.tfidf_post = torch.from_numpy(np.linspace(0.6, 1.8, C).astype(np.float32))
    r.score_thresh, r.topk_candidates, r.nms_thresh, r.detections_per_img = 0.01, 400, 0.45, 200
    r.box_coder = det_utils.BoxCoder(weights=(10.0, 10.0, 5.0, 5.0))
    logits, regs, anchors = syn.ssd_inputs(71, 2, A, C)
    head = {"cls_logits": torch.from_numpy(logits), "bbox_regression": torch.from_numpy(regs)}
    out = r.postprocess_detections(head, [torch.from_numpy(anchors)] * 2, [(300, 300)] * 2)
    pack = {"args": np.array([71, 2, A, C, 300]), "idf": r.tfidf_post.numpy()}
    for i, o in enumerate(out):
        pack[f"boxes_{i}"] = o["boxes"].numpy()
        pack[f"scores_{i}"] = o["scores"].numpy()
        pack[f"labels_{i}"] = o["labels"].numpy()
    np.savez_compressed(os.path.join(HERE, "ssd_postprocess.npz"), **pack)


if __name__ == "__main__":
    legacy()
    ssd()
    retinanet()
    legacy_get_target()
    roi()
    for f in ("legacy_yolo_loss.npz", "legacy_get_target.npz", "retinanet_postprocess.npz", "ssd_postprocess.npz", "roi_postprocess.npz"):
        print(f"  {f:32s} {os.path.getsize(os.path.join(HERE, f)):>9d} B")
