"""Tensor-level front end of libb200det.so: validates CUDA tensors, owns scratch buffers and
passes raw device pointers + the current CUDA stream through the C ABI.  Everything here is
plumbing; the arithmetic lives in csrc/*.cu.  CUDA tensors only -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (CIOU, DIOU, GIOU, IOU, IOU_TV, NMS_MAJORITY, NMS_TV, NMS_TV_CLASS,  # noqa: F401
                   NMS_TV_AUTO, NMS_TV_TRICK)

Tensor = torch.Tensor

_scratch: Dict[Tuple[int, int, str], Tensor] = {}

# Candidate rows kept per image unless the caller says otherwise.  The reference has no cap; a slab
# overflow is never silent (device status word -> RuntimeError) and the drop-ins retry with the
# worst case N.  The NMS bitmask scratch grows with capacity^2/8 bytes per image.
DEFAULT_CAPACITY = 4096


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _first_cuda_tensor(obj):
    if isinstance(obj, Tensor):
        return obj if obj.is_cuda else None
    if isinstance(obj, (list, tuple)):
        for o in obj:
            t = _first_cuda_tensor(o)
            if t is not None:
                return t
    return None


def _device_guard(fn):
    """Runs ``fn`` with the device of its first CUDA tensor argument current: the library launches on the current
    device's stream and never calls cudaSetDevice itself, so a tensor that lives on another GPU than the current one
    must not be launched from here (one process per GPU is the supported layout, like the reference's mp.spawn)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = _first_cuda_tensor(list(args) + list(kwargs.values()))
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapper


def _ptr(t: Optional[Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(t: Tensor, name: str, dtype=None) -> Tensor:
    if not isinstance(t, Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (object_detectors_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def workspace(nbytes: int, device: torch.device, tag: str = "ws") -> Tensor:
    """Grow-only byte buffer per (device, current stream, tag); torch's allocator returns >= 512 B alignment.
    Keyed by stream because the kernels that use it are asynchronous: two streams must never share scratch, and a
    buffer that is replaced by a larger one is then only ever reused in stream order."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


# --------------------------------------------------------------------------------------- YOLO
def make_layout(grids: Sequence[int], batch: int, anchors, img_size: float, num_classes: int,
                softmax: bool) -> _lib.YoloLayout:
    """Host-side geometry with the reference's roundings (yolo_forw.py:97-99,108-113):
    ``scaled = fp32(a / (img/grid))`` (python double division, then cast), ``rel = scaled/grid``
    as an fp32 division."""
    if len(grids) != len(anchors) or not 1 <= len(grids) <= _lib.MAX_SCALES:
        raise RuntimeError("anchors must hold one list per head tensor (<= 4 scales)")
    na = len(anchors[0])
    if any(len(a) != na for a in anchors) or not 1 <= na <= _lib.MAX_ANCHORS:
        raise RuntimeError("every scale must use the same number of anchors (1..8)")
    lay = _lib.YoloLayout()
    lay.num_scales, lay.num_anchors, lay.num_classes = len(grids), na, int(num_classes)
    lay.batch, lay.softmax, lay.img_size = int(batch), int(bool(softmax)), float(img_size)
    for s, g in enumerate(grids):
        lay.grid[s] = int(g)
        stride = img_size / g
        for a, (aw, ah) in enumerate(anchors[s]):
            lay.anchor_rel[s][a][0] = float(np.float32(aw / stride) / np.float32(g))
            lay.anchor_rel[s][a][1] = float(np.float32(ah / stride) / np.float32(g))
    return lay


def _heads_args(heads: Sequence[Tensor], anchors, img_size, num_classes, softmax):
    hs = [_need_cuda(h, f"heads[{i}]", torch.float32) for i, h in enumerate(heads)]
    b = hs[0].shape[0]
    na = len(anchors[0])
    grids = []
    for h in hs:
        if h.dim() != 4 or h.shape[0] != b or h.shape[1] != na * (5 + num_classes) or h.shape[2] != h.shape[3]:
            raise RuntimeError(f"head tensor of shape {tuple(h.shape)} is not [B, A*(5+C), G, G] "
                               f"with A={na}, C={num_classes}")
        grids.append(h.shape[2])
    lay = make_layout(grids, b, anchors, img_size, num_classes, softmax)
    arr = (C.c_void_p * len(hs))(*[h.data_ptr() for h in hs])
    n_total = sum(g * g * na for g in grids)
    return hs, lay, arr, n_total


def _idf_arg(idf: Optional[Tensor], num_classes: int, device) -> Optional[Tensor]:
    if idf is None:
        return None
    idf = torch.as_tensor(idf, dtype=torch.float32, device=device).contiguous()
    if idf.numel() == 1:          # the reference's `idf_logits = tensor(1)` scalar (yolo_forw.py:38)
        idf = idf.reshape(1).expand(num_classes).contiguous()
    if idf.numel() != num_classes:
        raise RuntimeError("idf must hold one weight per class")
    return idf


@_device_guard
def yolo_decode_dense(heads, anchors, img_size, num_classes, idf=None, softmax=True) -> Tensor:
    """-> [B, N, 5+C]  (YOLOForw.forward inference branch)."""
    lib = _lib.load()
    hs, lay, arr, n = _heads_args(heads, anchors, img_size, num_classes, softmax)
    idf = _idf_arg(idf, num_classes, hs[0].device)
    out = torch.empty((lay.batch, n, 5 + num_classes), dtype=torch.float32, device=hs[0].device)
    _lib.check(lib.b200_yolo_decode_dense(C.byref(lay), arr, _ptr(idf), _ptr(out), _stream()),
               "b200_yolo_decode_dense")
    return out


@_device_guard
def yolo_decode_filter(heads, anchors, img_size, num_classes, idf=None, softmax=True,
                       conf_thr: float = 0.1, capacity: Optional[int] = None):
    """-> dict(box [B,cap,4], score [B,cap], label [B,cap] i32, anchor [B,cap] i32, count [B] i32)."""
    lib = _lib.load()
    hs, lay, arr, n = _heads_args(heads, anchors, img_size, num_classes, softmax)
    dev = hs[0].device
    idf = _idf_arg(idf, num_classes, dev)
    cap = int(capacity or min(n, DEFAULT_CAPACITY))
    b = lay.batch
    box = torch.empty((b, cap, 4), dtype=torch.float32, device=dev)
    score = torch.empty((b, cap), dtype=torch.float32, device=dev)
    label = torch.empty((b, cap), dtype=torch.int32, device=dev)
    anchor = torch.empty((b, cap), dtype=torch.int32, device=dev)
    count = torch.empty((b,), dtype=torch.int32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    nbytes = lib.b200_yolo_workspace_bytes(C.byref(lay), cap)
    ws = workspace(nbytes, dev, "yolo")
    _lib.check(lib.b200_yolo_decode_filter(C.byref(lay), arr, _ptr(idf), float(np.float32(conf_thr)), cap,
                                           _ptr(box), _ptr(score), _ptr(label), _ptr(anchor), _ptr(count),
                                           _ptr(status), _ptr(ws), ws.numel(), _stream()),
               "b200_yolo_decode_filter")
    return {"box": box, "score": score, "label": label, "anchor": anchor, "count": count, "status": status}


class YoloPostprocess:
    """Reusable plan for b200_yolo_postprocess: owns outputs + workspace so repeated calls
    (benchmark loop, CUDA-graph capture) allocate nothing."""

    def __init__(self, grids, batch, anchors, img_size, num_classes, softmax=True, conf_thr=0.1,
                 nms_thr=0.6, nms_mode=NMS_MAJORITY, capacity=None, max_det=None, device="cuda"):
        self.lib = _lib.load()
        self.dev = torch.device(device)
        self.lay = make_layout(grids, batch, anchors, img_size, num_classes, softmax)
        self.num_classes = num_classes
        na = len(anchors[0])
        self.head_shapes = [(batch, na * (5 + num_classes), g, g) for g in grids]
        self.n = sum(g * g * na for g in grids)
        self.cap = int(capacity or min(self.n, DEFAULT_CAPACITY))
        self.max_det = int(max_det or self.cap)
        self.conf_thr = float(np.float32(conf_thr))
        self.nms_thr = float(nms_thr)   # double: MAJORITY rounds to fp32 in C, TV modes compare in double
        self.nms_mode = int(nms_mode)
        b = batch
        self.det = torch.empty((b, self.max_det, 6), dtype=torch.float32, device=self.dev)
        self.det_keep = torch.empty((b, self.max_det), dtype=torch.int32, device=self.dev)
        self.det_anchor = torch.empty((b, self.max_det), dtype=torch.int32, device=self.dev)
        self.det_count = torch.zeros((b,), dtype=torch.int32, device=self.dev)
        self.cand_count = torch.zeros((b,), dtype=torch.int32, device=self.dev)
        self.status = torch.zeros((1,), dtype=torch.int32, device=self.dev)
        self.ws_bytes = self.lib.b200_yolo_workspace_bytes(C.byref(self.lay), self.cap)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.dev)

    def __call__(self, heads: Sequence[Tensor], idf: Optional[Tensor] = None):
        if self.dev.index is not None and self.dev.index != torch.cuda.current_device():
            with torch.cuda.device(self.dev):
                return self.__call__(heads, idf)
        for h, shp in zip(heads, self.head_shapes):
            if tuple(h.shape) != shp or not h.is_cuda or h.dtype != torch.float32 or not h.is_contiguous():
                raise RuntimeError(f"head tensor {tuple(h.shape)} does not match the plan {shp} "
                                   "(CUDA, fp32, contiguous)")
        arr = (C.c_void_p * len(heads))(*[h.data_ptr() for h in heads])
        idf = _idf_arg(idf, self.num_classes, self.dev)
        _lib.check(self.lib.b200_yolo_postprocess(
            C.byref(self.lay), arr, _ptr(idf), self.conf_thr, self.nms_thr, self.nms_mode, self.cap,
            self.max_det, _ptr(self.det), _ptr(self.det_keep), _ptr(self.det_anchor), _ptr(self.det_count),
            _ptr(self.cand_count), _ptr(self.status), _ptr(self.ws), self.ws_bytes, _stream()),
            "b200_yolo_postprocess")
        return self.det, self.det_keep, self.det_anchor, self.det_count, self.cand_count

    def decode(self, heads: Sequence[Tensor], idf: Optional[Tensor] = None, stream: Optional[torch.cuda.Stream] = None):
        """Phase 1 (b200_yolo_postprocess_decode) on `stream` (default: current): candidates into the plan's slab."""
        arr = (C.c_void_p * len(heads))(*[h.data_ptr() for h in heads])
        idf = _idf_arg(idf, self.num_classes, self.dev)
        st = _stream() if stream is None else C.c_void_p(stream.cuda_stream)
        _lib.check(self.lib.b200_yolo_postprocess_decode(
            C.byref(self.lay), arr, _ptr(idf), self.conf_thr, self.cap, _ptr(self.status), _ptr(self.ws),
            self.ws_bytes, st), "b200_yolo_postprocess_decode")

    def nms(self, stream: Optional[torch.cuda.Stream] = None):
        """Phase 2 (b200_yolo_postprocess_nms) on `stream` (default: current), ordered after phase 1."""
        st = _stream() if stream is None else C.c_void_p(stream.cuda_stream)
        _lib.check(self.lib.b200_yolo_postprocess_nms(
            C.byref(self.lay), self.nms_thr, self.nms_mode, self.cap, self.max_det, _ptr(self.det),
            _ptr(self.det_keep), _ptr(self.det_anchor), _ptr(self.det_count), _ptr(self.cand_count),
            _ptr(self.status), _ptr(self.ws), self.ws_bytes, st), "b200_yolo_postprocess_nms")
        return self.det, self.det_keep, self.det_anchor, self.det_count, self.cand_count

    def check_status(self):
        """Raises if a batch since the last check overflowed the slab or `max_det`; the status word is sticky on the
        device (kernels only OR into it), so it is cleared here once it has been reported."""
        st = int(self.status.item())
        if st:
            self.status.zero_()
        if st & 1:
            raise RuntimeError("candidate slab overflow: raise `capacity`")
        if st & 2:
            raise RuntimeError("more detections than `max_det`")


@_device_guard
def yolo_postprocess(heads, anchors, img_size, num_classes, idf=None, softmax=True, conf_thr=0.1,
                     nms_thr=0.6, nms_mode=NMS_MAJORITY, capacity=None, max_det=None):
    hs, lay, arr, n = _heads_args(heads, anchors, img_size, num_classes, softmax)
    plan = YoloPostprocess([h.shape[2] for h in hs], lay.batch, anchors, img_size, num_classes, softmax,
                           conf_thr, nms_thr, nms_mode, capacity, max_det, hs[0].device)
    out = plan(hs, idf)
    plan.check_status()
    return out


def yolo_postprocess_host(heads_host: Sequence[Tensor], anchors, img_size, num_classes, idf_host=None,
                          softmax=True, conf_thr=0.1, nms_thr=0.6, nms_mode=NMS_MAJORITY,
                          capacity=None, max_det=300, out=None):
    """End-to-end entry with HOST tensors (pinned recommended): H2D, decode+NMS, D2H."""
    lib = _lib.load()
    b = heads_host[0].shape[0]
    na = len(anchors[0])
    grids = [h.shape[2] for h in heads_host]
    for h in heads_host:
        if h.is_cuda or h.dtype != torch.float32 or not h.is_contiguous():
            raise RuntimeError("yolo_postprocess_host expects contiguous fp32 HOST tensors")
    lay = make_layout(grids, b, anchors, img_size, num_classes, softmax)
    n = sum(g * g * na for g in grids)
    cap = int(capacity or min(n, DEFAULT_CAPACITY))
    if out is None:
        out = (torch.empty((b, max_det, 6), dtype=torch.float32).pin_memory(),
               torch.empty((b, max_det), dtype=torch.int32).pin_memory(),
               torch.empty((b,), dtype=torch.int32).pin_memory(),
               torch.zeros((1,), dtype=torch.int32).pin_memory())
    det, keep, cnt, status = out
    arr = (C.c_void_p * len(heads_host))(*[h.data_ptr() for h in heads_host])
    idf_p = C.c_void_p(0)
    if idf_host is not None:
        idf_host = torch.as_tensor(idf_host, dtype=torch.float32).contiguous()
        idf_p = C.c_void_p(idf_host.data_ptr())
    _lib.check(lib.b200_yolo_postprocess_host(C.byref(lay), arr, idf_p, float(np.float32(conf_thr)),
                                              float(nms_thr), int(nms_mode), cap, int(max_det),
                                              _ptr(det), _ptr(keep), _ptr(cnt), _ptr(status)),
               "b200_yolo_postprocess_host")
    return det, keep, cnt, status


# ---------------------------------------------------------------------------------------- NMS
@_device_guard
def nms_segments(boxes: Tensor, scores: Tensor, labels: Optional[Tensor], seg_offsets: Tensor,
                 iou_thr: float, mode: int, max_segment: int = 0):
    """Batched NMS.  -> (keep int64 [T], keep_count int32 [S], labels_out int32 [T]).
    ``max_segment``: host-known bound of the largest segment (0 = T); it sizes the bitmask."""
    lib = _lib.load()
    boxes = _need_cuda(boxes, "boxes", torch.float32)
    scores = _need_cuda(scores, "scores", torch.float32)
    seg_offsets = _need_cuda(seg_offsets, "seg_offsets", torch.int32)
    t = boxes.shape[0]
    s = seg_offsets.numel() - 1
    if labels is not None:
        labels = _need_cuda(labels, "labels").to(torch.int32).contiguous()
    dev = boxes.device
    keep = torch.empty((max(t, 1),), dtype=torch.int64, device=dev)
    keep_count = torch.zeros((max(s, 1),), dtype=torch.int32, device=dev)
    labels_out = torch.empty((max(t, 1),), dtype=torch.int32, device=dev) if mode == NMS_MAJORITY else None
    nbytes = lib.b200_nms_workspace_bytes(t, s, int(max_segment))
    ws = workspace(nbytes, dev, "nms")
    if boxes.data_ptr() % 16:
        boxes = boxes.clone()
    _lib.check(lib.b200_nms(_ptr(boxes), _ptr(scores), _ptr(labels), _ptr(seg_offsets), s, t, int(max_segment),
                            float(iou_thr), int(mode), _ptr(keep), _ptr(keep_count), _ptr(labels_out), _ptr(ws),
                            ws.numel(), _stream()), "b200_nms")
    return keep, keep_count, labels_out


# ---------------------------------------------------------------------------------------- IoU
@_device_guard
def box_iou(b1: Tensor, b2: Tensor, kind: int = IOU, xcycwh: bool = False) -> Tensor:
    lib = _lib.load()
    b1 = _need_cuda(b1, "boxes1", torch.float32)
    b2 = _need_cuda(b2, "boxes2", torch.float32)
    m, n = b1.shape[0], b2.shape[0]
    out = torch.empty((m, n), dtype=torch.float32, device=b1.device)
    if b1.data_ptr() % 16:
        b1 = b1.clone()
    if b2.data_ptr() % 16:
        b2 = b2.clone()
    _lib.check(lib.b200_box_iou(_ptr(b1), m, _ptr(b2), n, int(kind), int(bool(xcycwh)), _ptr(out), _stream()),
               "b200_box_iou")
    return out


@_device_guard
def box_iou_paired(b1: Tensor, b2: Tensor, kind: int = IOU, xcycwh: bool = False) -> Tensor:
    lib = _lib.load()
    b1 = _need_cuda(b1, "boxes1", torch.float32)
    b2 = _need_cuda(b2, "boxes2", torch.float32)
    if b1.shape != b2.shape:
        raise RuntimeError("paired IoU needs equal shapes")
    k = b1.shape[0]
    out = torch.empty((k,), dtype=torch.float32, device=b1.device)
    if b1.data_ptr() % 16:
        b1 = b1.clone()
    if b2.data_ptr() % 16:
        b2 = b2.clone()
    _lib.check(lib.b200_box_iou_paired(_ptr(b1), _ptr(b2), k, int(kind), int(bool(xcycwh)), _ptr(out),
                                       _stream()), "b200_box_iou_paired")
    return out


class _PairedIoU(torch.autograd.Function):
    """Differentiable paired IoU family: forward = b200_box_iou_paired (bit-identical to the forward-only call),
    backward = b200_box_iou_paired_backward."""

    @staticmethod
    def forward(ctx, b1, b2, kind, xcycwh):
        b1c, b2c = b1.detach().float().contiguous(), b2.detach().float().contiguous()
        ctx.save_for_backward(b1c, b2c)
        ctx.kind, ctx.xcycwh = int(kind), bool(xcycwh)
        return box_iou_paired(b1c, b2c, kind, xcycwh)

    @staticmethod
    def backward(ctx, grad_out):
        b1, b2 = ctx.saved_tensors
        lib = _lib.load()
        go = grad_out.float().contiguous()
        k = b1.shape[0]
        g1 = torch.empty_like(b1) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(b2) if ctx.needs_input_grad[1] else None
        if k:
            _lib.check(lib.b200_box_iou_paired_backward(_ptr(b1), _ptr(b2), _ptr(go), k, ctx.kind, int(ctx.xcycwh),
                                                        _ptr(g1), _ptr(g2), _stream()), "b200_box_iou_paired_backward")
        return g1, g2, None, None


def box_iou_paired_autograd(b1: Tensor, b2: Tensor, kind: int = IOU, xcycwh: bool = False) -> Tensor:
    """``[K,4] x [K,4] -> [K]`` with gradients to either input (kinds IoU / GIoU / DIoU / CIoU)."""
    b1 = _need_cuda(b1, "boxes1")
    b2 = _need_cuda(b2, "boxes2")
    if b1.shape != b2.shape or b1.dim() != 2 or b1.shape[1] != 4:
        raise RuntimeError("paired IoU needs two [K,4] tensors")
    return _PairedIoU.apply(b1, b2, int(kind), bool(xcycwh))


@_device_guard
def iou_match(gt: Tensor, gt_count: Tensor, anchors: Tensor, kind: int = GIOU, ignore_thr: float = 0.5):
    """gt [B, Mmax, 4] rel xc,yc,w,h; gt_count [B] i32; anchors [N,4] (cxypwh).
    -> best_anchor int64 [B, Mmax], noobj bool [B, N]."""
    lib = _lib.load()
    gt = _need_cuda(gt, "gt", torch.float32)
    gt_count = _need_cuda(gt_count, "gt_count", torch.int32)
    anchors = _need_cuda(anchors, "anchors", torch.float32)
    b, mmax = gt.shape[0], gt.shape[1]
    n = anchors.shape[0]
    dev = gt.device
    best = torch.empty((b, mmax), dtype=torch.int64, device=dev)
    noobj = torch.empty((b, n), dtype=torch.uint8, device=dev)
    nbytes = lib.b200_iou_match_workspace_bytes(b, mmax)
    ws = workspace(nbytes, dev, "match")
    _lib.check(lib.b200_iou_match(_ptr(gt), _ptr(gt_count), b, mmax, _ptr(anchors), n, int(kind),
                                  float(np.float32(ignore_thr)), _ptr(best), _ptr(noobj), _ptr(ws), ws.numel(),
                                  _stream()), "b200_iou_match")
    return best, noobj.view(torch.bool)


# ---------------------------------------------------------------------------------------- RPN
@_device_guard
def rpn_filter(objectness: Tensor, deltas: Tensor, anchors: Tensor, level_sizes: Sequence[int],
               image_hw: Tensor, pre_nms_top_n: int, post_nms_top_n: int, nms_thr: float = 0.7,
               score_thr: float = 0.0, min_size: float = 1e-3, nms_mode: int = NMS_TV_CLASS):
    """-> boxes [B,post,4], scores [B,post], index [B,post] i32, count [B] i32."""
    lib = _lib.load()
    objectness = _need_cuda(objectness, "objectness", torch.float32)
    deltas = _need_cuda(deltas, "deltas", torch.float32)
    anchors = _need_cuda(anchors, "anchors", torch.float32)
    image_hw = _need_cuda(image_hw, "image_hw", torch.float32)
    b, total = objectness.shape
    dev = objectness.device
    lv = (C.c_int32 * len(level_sizes))(*[int(v) for v in level_sizes])
    boxes = torch.zeros((b, post_nms_top_n, 4), dtype=torch.float32, device=dev)
    scores = torch.zeros((b, post_nms_top_n), dtype=torch.float32, device=dev)
    index = torch.zeros((b, post_nms_top_n), dtype=torch.int32, device=dev)
    count = torch.zeros((b,), dtype=torch.int32, device=dev)
    nbytes = lib.b200_rpn_workspace_bytes(b, total, len(level_sizes), pre_nms_top_n)
    ws = workspace(nbytes, dev, "rpn")
    _lib.check(lib.b200_rpn_filter(_ptr(objectness), _ptr(deltas), _ptr(anchors), b, total, lv,
                                   len(level_sizes), _ptr(image_hw), int(pre_nms_top_n), int(post_nms_top_n),
                                   float(nms_thr), float(np.float32(score_thr)),
                                   float(np.float32(min_size)), int(nms_mode), _ptr(boxes), _ptr(scores),
                                   _ptr(index), _ptr(count), _ptr(ws), ws.numel(), _stream()),
               "b200_rpn_filter")
    return boxes, scores, index, count


@_device_guard
def rpn_top_n_idx(objectness: Tensor, level_sizes: Sequence[int], pre_nms_top_n: int) -> Tensor:
    """RegionProposalNetwork._get_top_n_idx: objectness [B, total] raw logits -> int64 [B, sum_l min(k, n_l)]."""
    lib = _lib.load()
    objectness = _need_cuda(objectness, "objectness", torch.float32)
    b, total = objectness.shape
    lv = (C.c_int32 * len(level_sizes))(*[int(v) for v in level_sizes])
    ktot = sum(min(int(pre_nms_top_n), int(v)) for v in level_sizes)
    out = torch.empty((b, ktot), dtype=torch.int64, device=objectness.device)
    ws = workspace(lib.b200_rpn_top_n_idx_workspace_bytes(b, total, len(level_sizes), int(pre_nms_top_n)), objectness.device,
                   "rpn_topk")
    _lib.check(lib.b200_rpn_top_n_idx(_ptr(objectness), b, total, lv, len(level_sizes), int(pre_nms_top_n), _ptr(out),
                                      _ptr(ws), ws.numel(), _stream()), "b200_rpn_top_n_idx")
    return out


@_device_guard
def rpn_filter_proposals(objectness: Tensor, proposals: Tensor, level_sizes: Sequence[int], image_hw: Tensor,
                         pre_nms_top_n: int, post_nms_top_n: int, nms_thr: float = 0.7, score_thr: float = 0.0,
                         min_size: float = 1e-3, nms_mode: int = NMS_TV_CLASS):
    """As rpn_filter, on boxes that are already decoded ([B, total, 4])."""
    lib = _lib.load()
    objectness = _need_cuda(objectness, "objectness", torch.float32)
    proposals = _need_cuda(proposals, "proposals", torch.float32)
    image_hw = _need_cuda(image_hw, "image_hw", torch.float32)
    b, total = objectness.shape
    dev = objectness.device
    lv = (C.c_int32 * len(level_sizes))(*[int(v) for v in level_sizes])
    boxes = torch.zeros((b, post_nms_top_n, 4), dtype=torch.float32, device=dev)
    scores = torch.zeros((b, post_nms_top_n), dtype=torch.float32, device=dev)
    index = torch.zeros((b, post_nms_top_n), dtype=torch.int32, device=dev)
    count = torch.zeros((b,), dtype=torch.int32, device=dev)
    nbytes = lib.b200_rpn_workspace_bytes(b, total, len(level_sizes), pre_nms_top_n)
    ws = workspace(nbytes, dev, "rpn")
    _lib.check(lib.b200_rpn_filter_proposals(_ptr(objectness), _ptr(proposals), b, total, lv, len(level_sizes),
                                             _ptr(image_hw), int(pre_nms_top_n), int(post_nms_top_n),
                                             float(nms_thr), float(np.float32(score_thr)),
                                             float(np.float32(min_size)), int(nms_mode), _ptr(boxes), _ptr(scores),
                                             _ptr(index), _ptr(count), _ptr(ws), ws.numel(), _stream()),
               "b200_rpn_filter_proposals")
    return boxes, scores, index, count


# ------------------------------------------------------------------------------- element-wise
@_device_guard
def abs_coord(box: Tensor) -> Tensor:
    """[..., 4] xc,yc,w,h -> x1,y1,x2,y2 (helper.get_abs_coord)."""
    lib = _lib.load()
    box = _need_cuda(box, "box", torch.float32)
    if box.data_ptr() % 16:
        box = box.clone()
    out = torch.empty_like(box)
    _lib.check(lib.b200_abs_coord(_ptr(box), box.numel() // 4, _ptr(out), _stream()), "b200_abs_coord")
    return out


@_device_guard
def boxcoder_decode(rel_codes: Tensor, boxes: Tensor, weights=(1.0, 1.0, 1.0, 1.0),
                    xform_clip: float = 4.135166556742356) -> Tensor:
    """rel_codes [n, 4k], boxes [n,4] -> [n, 4k] (BoxCoder.decode_single)."""
    lib = _lib.load()
    rel_codes = _need_cuda(rel_codes, "rel_codes", torch.float32)
    boxes = _need_cuda(boxes, "boxes").to(torch.float32).contiguous()
    n = boxes.shape[0]
    k = rel_codes.shape[1] // 4 if n else 1
    if rel_codes.data_ptr() % 16:
        rel_codes = rel_codes.clone()
    if boxes.data_ptr() % 16:
        boxes = boxes.clone()
    out = torch.empty_like(rel_codes)
    w = (C.c_float * 4)(*[float(v) for v in weights])
    _lib.check(lib.b200_boxcoder_decode(_ptr(rel_codes), _ptr(boxes), n, k, w, float(np.float32(xform_clip)),
                                        _ptr(out), _stream()), "b200_boxcoder_decode")
    return out


@_device_guard
def boxcoder_encode(reference_boxes: Tensor, proposals: Tensor, weights=(1.0, 1.0, 1.0, 1.0)) -> Tensor:
    """reference_boxes [n,4], proposals [n,4] -> regression targets [n,4] (BoxCoder.encode_single)."""
    lib = _lib.load()
    ref = _need_cuda(reference_boxes, "reference_boxes").to(torch.float32).contiguous()
    prop = _need_cuda(proposals, "proposals").to(torch.float32).contiguous()
    if ref.shape != prop.shape or ref.dim() != 2 or ref.shape[1] != 4:
        raise RuntimeError("encode needs two [n,4] tensors")
    if ref.data_ptr() % 16:
        ref = ref.clone()
    if prop.data_ptr() % 16:
        prop = prop.clone()
    out = torch.empty_like(ref)
    w = (C.c_float * 4)(*[float(v) for v in weights])
    _lib.check(lib.b200_boxcoder_encode(_ptr(ref), _ptr(prop), ref.shape[0], w, _ptr(out), _stream()),
               "b200_boxcoder_encode")
    return out


@_device_guard
def matcher(quality: Tensor, high: float, low: float, allow_low_quality: bool = False) -> Tensor:
    """[M,N] quality -> int64 [N] matches (Matcher.__call__)."""
    lib = _lib.load()
    quality = _need_cuda(quality, "match_quality_matrix", torch.float32)
    m, n = quality.shape
    matches = torch.empty((n,), dtype=torch.int64, device=quality.device)
    ws = workspace(8 * n, quality.device, "matcher")
    _lib.check(lib.b200_matcher(_ptr(quality), m, n, float(np.float32(high)), float(np.float32(low)),
                                int(bool(allow_low_quality)), _ptr(matches), _ptr(ws), ws.numel(), _stream()),
               "b200_matcher")
    return matches


@_device_guard
def match_boxes(gt_boxes: Tensor, boxes: Tensor, high: float, low: float, allow_low_quality: bool = False,
                ssd: bool = False, return_vals: bool = False):
    """``Matcher(high, low, allow_low_quality)(box_iou(gt_boxes, boxes))`` (or SSDMatcher with ``ssd=True``) in one
    fused pass over the pairs: the [M, N] matrix is never materialised.  -> int64 [N] (and matched_vals fp32 [N])."""
    lib = _lib.load()
    gt_boxes = _need_cuda(gt_boxes, "gt_boxes").to(torch.float32).contiguous()
    boxes = _need_cuda(boxes, "boxes").to(torch.float32).contiguous()
    m, n = gt_boxes.shape[0], boxes.shape[0]
    if m == 0:
        raise ValueError("No ground-truth boxes available for one of the images during training")
    if n == 0:
        raise ValueError("No proposal boxes available for one of the images during training")
    if gt_boxes.data_ptr() % 16:
        gt_boxes = gt_boxes.clone()
    if boxes.data_ptr() % 16:
        boxes = boxes.clone()
    matches = torch.empty((n,), dtype=torch.int64, device=boxes.device)
    vals = torch.empty((n,), dtype=torch.float32, device=boxes.device) if return_vals else None
    ws = workspace(lib.b200_match_boxes_workspace_bytes(m, n), boxes.device, "match_boxes")
    _lib.check(lib.b200_match_boxes(_ptr(gt_boxes), m, _ptr(boxes), n, float(np.float32(high)), float(np.float32(low)),
                                    int(bool(allow_low_quality)), int(bool(ssd)), _ptr(matches), _ptr(vals), _ptr(ws),
                                    ws.numel(), _stream()), "b200_match_boxes")
    return (matches, vals) if return_vals else matches


@_device_guard
def matcher_ssd(quality: Tensor, threshold: float) -> Tensor:
    """SSDMatcher.__call__ on a materialised [M,N] matrix (_utils.py:347-361)."""
    lib = _lib.load()
    quality = _need_cuda(quality, "match_quality_matrix", torch.float32)
    m, n = quality.shape
    matches = matcher(quality, threshold, threshold, False)
    ws = workspace(8 * m, quality.device, "matcher_ssd")
    _lib.check(lib.b200_matcher_ssd_override(_ptr(quality), m, n, _ptr(matches), _ptr(ws), ws.numel(), _stream()),
               "b200_matcher_ssd_override")
    return matches


@_device_guard
def yolo_legacy_decode(head: Tensor, anchors_px, num_classes: int, img_size) -> Tensor:
    """One head of the legacy YOLOLoss layer (yolo/nets/yolo_loss.py:34-105, inference branch):
    [B, A*(5+C), H, W] -> [B, A*H*W, 5+C], rows ordered (a, h, w)."""
    lib = _lib.load()
    head = _need_cuda(head, "input", torch.float32)
    b, _, in_h, in_w = head.shape
    na = len(anchors_px)
    if head.shape[1] != na * (5 + num_classes):
        raise RuntimeError("head channels do not match anchors x (5 + classes)")
    stride_h, stride_w = img_size / in_h, img_size / in_w           # python floats, as the reference (:38-39)
    scaled = torch.tensor([(a_w / stride_w, a_h / stride_h) for a_w, a_h in anchors_px], dtype=torch.float32,
                          device=head.device)                        # FloatTensor(scaled_anchors) (:91-92)
    out = torch.empty((b, na * in_h * in_w, 5 + num_classes), dtype=torch.float32, device=head.device)
    _lib.check(lib.b200_yolo_legacy_decode(_ptr(head), b, na, num_classes, in_h, in_w, float(np.float32(stride_w)),
                                           float(np.float32(stride_h)), _ptr(scaled), _ptr(out), _stream()),
               "b200_yolo_legacy_decode")
    return out


ROI_SOFTMAX, ROI_GOMBIT, ROI_SIGMOID = 0, 1, 2
# Default candidate slab rows per image for the ROI post-process.  The worst case is rows x (classes - 1) (90 000
# for COCO, 1.2 M for LVIS) but the NMS bitmask scratch grows with capacity^2 / 8 bytes per image, and after the
# score threshold an image holds a few thousand candidates; overflow sets status bit 0 and `cand_count` holds the
# true counts, so a caller can retry with exactly what is needed (tvision/roi_heads.py does).
ROI_DEFAULT_CAPACITY = 16384


@_device_guard
def roi_postprocess(class_logits: Tensor, box_regression: Tensor, proposals: Sequence[Tensor], image_shapes,
                    tfidf: Optional[Tensor] = None, activation: int = ROI_SOFTMAX,
                    weights=(10.0, 10.0, 5.0, 5.0), xform_clip: float = 4.135166556742356,
                    score_thresh: float = 0.05, nms_thresh: float = 0.5, detections_per_img: int = 100,
                    min_size: float = 1e-2, nms_mode: int = NMS_TV_AUTO, capacity: Optional[int] = None):
    """RoIHeads.postprocess_detections (roi_heads.py:715-781) for the whole batch in two launches + NMS.
    Returns (det [B, D, 6], det_keep [B, D], det_count [B], cand_count [B])."""
    lib = _lib.load()
    class_logits = _need_cuda(class_logits, "class_logits", torch.float32)
    box_regression = _need_cuda(box_regression, "box_regression", torch.float32)
    dev = class_logits.device
    rows = [int(p.shape[0]) for p in proposals]
    b, r, c = len(rows), int(class_logits.shape[0]), int(class_logits.shape[1])
    if sum(rows) != r or box_regression.shape[1] != 4 * c:
        raise RuntimeError("class_logits / box_regression / proposals disagree on the number of rows or classes")
    prop = torch.cat([p.to(torch.float32) for p in proposals], 0).contiguous() if r else torch.zeros((0, 4), device=dev)
    prop = _need_cuda(prop, "proposals", torch.float32)
    off = torch.tensor(np.concatenate([[0], np.cumsum(rows)]), dtype=torch.int32, device=dev)
    hw = torch.tensor([[float(h), float(w)] for h, w in image_shapes], dtype=torch.float32, device=dev)
    if tfidf is not None:
        tfidf = torch.as_tensor(tfidf, dtype=torch.float32, device=dev).expand(c).contiguous()
    cap = int(capacity or max(1, min(max(rows) * (c - 1), ROI_DEFAULT_CAPACITY)))
    d = int(detections_per_img)
    det = torch.empty((b, d, 6), dtype=torch.float32, device=dev)
    keep = torch.empty((b, d), dtype=torch.int32, device=dev)
    dcnt = torch.zeros((b,), dtype=torch.int32, device=dev)
    ccnt = torch.zeros((b,), dtype=torch.int32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws = workspace(lib.b200_roi_workspace_bytes(b, cap), dev, "roi")
    w4 = (C.c_float * 4)(*[float(v) for v in weights])
    _lib.check(lib.b200_roi_postprocess(
        _ptr(class_logits), _ptr(box_regression), _ptr(prop), _ptr(off), b, r, c, _ptr(hw), _ptr(tfidf), int(activation),
        w4, float(np.float32(xform_clip)), float(np.float32(score_thresh)), float(np.float32(min_size)),
        float(nms_thresh), int(nms_mode), cap, d, _ptr(det), _ptr(keep), _ptr(dcnt), _ptr(ccnt), _ptr(status), _ptr(ws),
        ws.numel(), _stream()), "b200_roi_postprocess")
    return det, keep, dcnt, ccnt, status


@_device_guard
def ssd_postprocess(cls_logits: Tensor, bbox_regression: Tensor, image_anchors: Sequence[Tensor], image_shapes,
                    tfidf: Optional[Tensor] = None, weights=(10.0, 10.0, 5.0, 5.0), xform_clip: float = 4.135166556742356,
                    score_thresh: float = 0.01, topk_candidates: int = 400, nms_thresh: float = 0.45,
                    detections_per_img: int = 200, nms_mode: int = NMS_TV_AUTO, capacity: Optional[int] = None):
    """SSD.postprocess_detections (ssd.py:386-430) for the whole batch.  cls_logits [B, A, C], bbox_regression
    [B, A, 4], image_anchors list of [A, 4].  -> det [B, D, 6], det_count [B], cand_count [B], status [1]."""
    lib = _lib.load()
    cls_logits = _need_cuda(cls_logits, "cls_logits", torch.float32)
    bbox_regression = _need_cuda(bbox_regression, "bbox_regression", torch.float32)
    b, a, c = cls_logits.shape
    dev = cls_logits.device
    anc = torch.cat([t.to(torch.float32) for t in image_anchors], 0).contiguous()
    if bbox_regression.shape != (b, a, 4) or anc.shape != (b * a, 4):
        raise RuntimeError("cls_logits / bbox_regression / anchors disagree")
    off = torch.arange(0, (b + 1) * a, a, dtype=torch.int32, device=dev)
    hw = torch.tensor([[float(h), float(w)] for h, w in image_shapes], dtype=torch.float32, device=dev)
    if tfidf is not None:
        tfidf = torch.as_tensor(tfidf, dtype=torch.float32, device=dev).expand(c).contiguous()
    cap = int(capacity or max(1, min(a * (c - 1), ROI_DEFAULT_CAPACITY)))
    d = int(detections_per_img)
    det = torch.empty((b, d, 6), dtype=torch.float32, device=dev)
    dcnt = torch.zeros((b,), dtype=torch.int32, device=dev)
    ccnt = torch.zeros((b,), dtype=torch.int32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws = workspace(lib.b200_roi_workspace_bytes(b, cap), dev, "roi")
    w4 = (C.c_float * 4)(*[float(v) for v in weights])
    _lib.check(lib.b200_ssd_postprocess(
        _ptr(cls_logits.reshape(b * a, c)), _ptr(bbox_regression.reshape(b * a, 4)), _ptr(anc), _ptr(off), b, b * a, c, _ptr(hw),
        _ptr(tfidf), w4, float(np.float32(xform_clip)), float(np.float32(score_thresh)), int(topk_candidates), float(nms_thresh),
        int(nms_mode), cap, d, _ptr(det), _ptr(dcnt), _ptr(ccnt), _ptr(status), _ptr(ws), ws.numel(), _stream()),
        "b200_ssd_postprocess")
    return det, dcnt, ccnt, status


@_device_guard
def retinanet_postprocess(cls_logits: Tensor, bbox_regression: Tensor, anchors: Tensor, level_anchors: Sequence[int],
                          image_shapes, tfidf: Optional[Tensor] = None, score_thresh: float = 0.05,
                          topk_candidates: int = 1000, nms_thresh: float = 0.5, detections_per_img: int = 300,
                          nms_mode: int = NMS_TV_AUTO):
    """RetinaNet.postprocess_detections (retinanet.py:414-472) for the whole batch.  cls_logits [B, sumA, C],
    bbox_regression [B, sumA, 4], anchors [sumA, 4].  -> boxes [B,D,4], scores [B,D], labels [B,D] i32, count [B] i32."""
    lib = _lib.load()
    cls_logits = _need_cuda(cls_logits, "cls_logits", torch.float32)
    bbox_regression = _need_cuda(bbox_regression, "bbox_regression", torch.float32)
    anchors = _need_cuda(anchors, "anchors", torch.float32)
    b, total, c = cls_logits.shape
    dev = cls_logits.device
    if bbox_regression.shape != (b, total, 4) or anchors.shape != (total, 4) or sum(level_anchors) != total:
        raise RuntimeError("cls_logits / bbox_regression / anchors / level sizes disagree")
    hw = torch.tensor([[float(h), float(w)] for h, w in image_shapes], dtype=torch.float32, device=dev)
    if tfidf is not None:
        tfidf = torch.as_tensor(tfidf, dtype=torch.float32, device=dev).expand(c).contiguous()
    lv = (C.c_int32 * len(level_anchors))(*[int(v) for v in level_anchors])
    d = int(detections_per_img)
    boxes = torch.zeros((b, d, 4), dtype=torch.float32, device=dev)
    scores = torch.zeros((b, d), dtype=torch.float32, device=dev)
    labels = torch.zeros((b, d), dtype=torch.int32, device=dev)
    count = torch.zeros((b,), dtype=torch.int32, device=dev)
    ws = workspace(lib.b200_retinanet_workspace_bytes(b, total, c, len(level_anchors), int(topk_candidates)), dev, "retina")
    _lib.check(lib.b200_retinanet_postprocess(_ptr(cls_logits), _ptr(bbox_regression), _ptr(anchors), b, total, c, lv,
                                              len(level_anchors), _ptr(tfidf), _ptr(hw), int(topk_candidates),
                                              float(np.float32(score_thresh)), float(nms_thresh), int(nms_mode), d,
                                              _ptr(boxes), _ptr(scores), _ptr(labels), _ptr(count), _ptr(ws), ws.numel(),
                                              _stream()), "b200_retinanet_postprocess")
    return boxes, scores, labels, count


@_device_guard
def clip_boxes_to_image(boxes: Tensor, size) -> Tensor:
    """b200_clip_boxes_to_image: ``boxes [..., 4]`` xyxy clamped to the image ``size = (height, width)``."""
    lib = _lib.load()
    b = _need_cuda(boxes, "boxes", torch.float32)
    out = torch.empty_like(b)
    h, w = size
    _lib.check(lib.b200_clip_boxes_to_image(_ptr(b), b.numel() // 4, float(h), float(w), _ptr(out), _stream()),
               "b200_clip_boxes_to_image")
    return out


@_device_guard
def remove_small_boxes(boxes: Tensor, min_size: float) -> Tensor:
    """b200_remove_small_boxes: int64 indices of the boxes whose width and height are both >= ``min_size``."""
    lib = _lib.load()
    b = _need_cuda(boxes, "boxes", torch.float32)
    n = b.shape[0]
    keep = torch.empty((n,), dtype=torch.int64, device=b.device)
    cnt = torch.zeros((1,), dtype=torch.int32, device=b.device)
    _lib.check(lib.b200_remove_small_boxes(_ptr(b), n, float(np.float32(min_size)), _ptr(keep), _ptr(cnt), _stream()),
               "b200_remove_small_boxes")
    return keep[:int(cnt.item())]          # the reference's torch.where() synchronises here as well


@_device_guard
def emit_results(det: Tensor, det_count: Tensor, img_hw: Tensor, image_id: Tensor, inp_dim: float,
                 class_map: Optional[Tensor] = None, strict_reference: bool = True):
    """b200_emit_results: packed evaluation records of a whole batch.  -> (records [B*max_det, 6] fp32,
    category [B*max_det] i32, image [B*max_det] i64, total [1] i32); only the first ``total`` rows are defined."""
    lib = _lib.load()
    det = _need_cuda(det, "det", torch.float32)
    det_count = _need_cuda(det_count, "det_count", torch.int32)
    img_hw = _need_cuda(img_hw, "img_hw", torch.float32)
    image_id = _need_cuda(image_id, "image_id", torch.int64)
    b, max_det = det.shape[0], det.shape[1]
    dev = det.device
    rec = torch.empty((b * max_det, 6), dtype=torch.float32, device=dev)
    cat = torch.empty((b * max_det,), dtype=torch.int32, device=dev)
    img = torch.empty((b * max_det,), dtype=torch.int64, device=dev)
    total = torch.zeros((1,), dtype=torch.int32, device=dev)
    if class_map is not None:
        class_map = _need_cuda(class_map, "class_map", torch.int32)
    _lib.check(lib.b200_emit_results(_ptr(det), _ptr(det_count), b, max_det, _ptr(img_hw), _ptr(image_id),
                                     float(np.float32(inp_dim)), _ptr(class_map), 0 if class_map is None else class_map.numel(),
                                     int(bool(strict_reference)), _ptr(rec), _ptr(cat), _ptr(img), _ptr(total), _stream()),
               "b200_emit_results")
    return rec, cat, img, total


@_device_guard
def pack_detections(det: Tensor, det_count: Tensor) -> Tensor:
    lib = _lib.load()
    b, max_det = det.shape[0], det.shape[1]
    msg = torch.empty((b * (1 + max_det * 6),), dtype=torch.float32, device=det.device)
    _lib.check(lib.b200_pack_detections(_ptr(det), _ptr(det_count), b, max_det, _ptr(msg), _stream()),
               "b200_pack_detections")
    return msg
