"""Drop-in replacements for the box utilities of ``yolo/utilities/helper.py`` (reference), same
names, argument meaning and tensor layouts, backed by libb200det.so.  CUDA tensors only.

    get_abs_coord(box)                                   helper.py:203-217
    bbox_iou(bb1, bb2, iou_type, CUDA=True, xcycwh=True) helper.py:221-277
    nms_majority(P, thresh_iou=0.6)                      helper.py:280-382
    torch80_to_91(label) / coco80_to_coco91_class(label) helper.py:8-24 (COCO id table)
"""
from __future__ import annotations

import torch

from ... import ops

# COCO category ids of the 80 contiguous training classes (public COCO annotation ids)
_COCO91 = [c for c in range(1, 91) if c not in (12, 26, 29, 30, 45, 66, 68, 69, 71, 83)]
_coco91_cache = {}


def coco80_to_coco91_class(label):
    return _COCO91[int(label)]


def torch80_to_91(label: torch.Tensor) -> torch.Tensor:
    """``label``: integer tensor of contiguous class ids -> COCO category ids (same device)."""
    table = _coco91_cache.get(label.device)
    if table is None:
        table = torch.tensor(_COCO91, device=label.device)
        _coco91_cache[label.device] = table
    return table[label]


def _to_cuda(t: torch.Tensor) -> torch.Tensor:
    # the reference moves its inputs to the GPU when one is available (helper.py:204-205, 241-243)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("object_detectors_b200.helper needs a CUDA device (no CPU fallback)")
        t = t.cuda()
    return t


def get_abs_coord(box: torch.Tensor) -> torch.Tensor:
    """centre format ``[n,4]`` or ``[B,N,4]`` -> corner format, new tensor of the same shape."""
    box = _to_cuda(box)
    if box.dim() not in (2, 3) or box.shape[-1] != 4:
        raise RuntimeError(f"get_abs_coord expects [n,4] or [B,N,4], got {tuple(box.shape)}")
    return ops.abs_coord(box.float())


def bbox_iou(bb1: torch.Tensor, bb2: torch.Tensor, iou_type, CUDA: bool = True, xcycwh: bool = True) -> torch.Tensor:
    """IoU (0) / GIoU (1) / DIoU (2) / CIoU (3) with the reference's broadcasting patterns:
    ``[M,1,4] x [1,N,4] -> [M,N]`` (target matching, yolo_forw.py:186) and ``[K,4] x [K,4] -> [K]``
    (loss side, yolo_forw.py:125).  The paired form is differentiable in CUDA (forward + backward kernels,
    SURVEY.md 8f3); broadcast forms that require grad (not used by the reference) are evaluated with
    differentiable torch ops in the reference's operation order."""
    kind = iou_type if iou_type in (1, 2, 3) else 0
    bb1, bb2 = _to_cuda(bb1), _to_cuda(bb2)
    if bb1.requires_grad or bb2.requires_grad:
        if bb1.dim() == 2 and bb1.shape == bb2.shape and bb1.shape[1] == 4:
            return ops.box_iou_paired_autograd(bb1, bb2, kind, xcycwh)
        return _bbox_iou_autograd(bb1, bb2, kind, xcycwh)
    if bb1.dim() == 3 and bb2.dim() == 3 and bb1.shape[1] == 1 and bb2.shape[0] == 1:
        return ops.box_iou(bb1[:, 0].float(), bb2[0].float(), kind, xcycwh)
    if bb1.dim() == 2 and bb1.shape == bb2.shape:
        return ops.box_iou_paired(bb1.float(), bb2.float(), kind, xcycwh)
    if bb1.dim() == 1 and bb2.dim() == 2:          # "box1 is 4, box2 is nx4" (helper.py:222)
        return ops.box_iou(bb1[None].float(), bb2.float(), kind, xcycwh)[0]
    raise RuntimeError(f"bbox_iou: unsupported broadcast {tuple(bb1.shape)} x {tuple(bb2.shape)}")


def _bbox_iou_autograd(bb1, bb2, kind, xcycwh):
    import math

    def corners(b):
        hw, hh = b[..., 2] / 2, b[..., 3] / 2
        return b[..., 0] - hw, b[..., 1] - hh, b[..., 0] + hw, b[..., 1] + hh
    ax1, ay1, ax2, ay2 = corners(bb1) if xcycwh else (bb1[..., 0], bb1[..., 1], bb1[..., 2], bb1[..., 3])
    bx1, by1, bx2, by2 = corners(bb2) if xcycwh else (bb2[..., 0], bb2[..., 1], bb2[..., 2], bb2[..., 3])
    inter = (torch.min(ax2, bx2) - torch.max(ax1, bx1)).clamp(0) * (torch.min(ay2, by2) - torch.max(ay1, by1)).clamp(0)
    w1, h1, w2, h2 = ax2 - ax1, ay2 - ay1, bx2 - bx1, by2 - by1
    union = (w1 * h1 + 1e-16) + w2 * h2 - inter
    iou = inter / union
    if kind == 0:
        return iou
    cw = torch.max(ax2, bx2) - torch.min(ax1, bx1)
    ch = torch.max(ay2, by2) - torch.min(ay1, by1)
    if kind == 1:
        c_area = cw * ch + 1e-16
        return iou - (c_area - union) / c_area
    c2 = cw ** 2 + ch ** 2 + 1e-16
    rho2 = ((bx1 + bx2) - (ax1 + ax2)) ** 2 / 4 + ((by1 + by2) - (ay1 + ay2)) ** 2 / 4
    if kind == 2:
        return iou - rho2 / c2
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
    with torch.no_grad():
        alpha = v / (1 - iou + v)
    return iou - (rho2 / c2 + v * alpha)


def nms_majority(P: torch.Tensor, thresh_iou: float = 0.6) -> torch.Tensor:
    """``P``: ``[n,6]`` rows x1,y1,x2,y2,score,label.  Returns the kept rows ``[K,6]`` in descending
    score with the majority-vote label, and -- like the reference, which appends views of ``P`` and
    relabels in place (helper.py:326,375) -- writes the new labels into ``P[:, 5]`` as well."""
    P = _to_cuda(P)
    n = P.shape[0]
    if n == 0:
        return torch.stack([])          # the reference raises on an empty stack as well
    Pc = P.float().contiguous()
    seg = torch.tensor([0, n], dtype=torch.int32, device=P.device)
    keep, cnt, labels = ops.nms_segments(Pc[:, :4].contiguous(), Pc[:, 4].contiguous(), Pc[:, 5].to(torch.int32),
                                         seg, thresh_iou, ops.NMS_MAJORITY, n)
    k = int(cnt[0])
    keep = keep[:k]
    P[keep, 5] = labels[:k].to(P.dtype)
    return P[keep]
