"""oracle/cref.py -- TEST INFRASTRUCTURE ONLY: ctypes front end of oracle/boxops_ref.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "boxops_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ref_nms_majority.restype = C.c_int
        _lib.ref_nms_tv.restype = C.c_int
        _lib.ref_pair_iou.restype = C.c_float
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def nms_majority(det6: np.ndarray, thr: float, num_classes: int):
    """-> (keep_idx int32 [K], keep_label int32 [K]) in descending score."""
    det6 = np.ascontiguousarray(det6, dtype=np.float32)
    n = det6.shape[0]
    ki = np.zeros(max(n, 1), np.int32)
    kl = np.zeros(max(n, 1), np.int32)
    k = lib().ref_nms_majority(_p(det6, C.c_float), C.c_int(n), C.c_float(np.float32(thr)),
                               C.c_int(num_classes), _p(ki, C.c_int32), _p(kl, C.c_int32))
    return ki[:k].copy(), kl[:k].copy()


def nms_tv(boxes: np.ndarray, scores: np.ndarray, thr: float, labels: np.ndarray | None = None):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    keep = np.zeros(max(n, 1), np.int64)
    lp = None
    if labels is not None:
        labels = np.ascontiguousarray(labels, dtype=np.int64)
        lp = _p(labels, C.c_int64)
    k = lib().ref_nms_tv(_p(boxes, C.c_float), _p(scores, C.c_float), lp, C.c_int(n),
                         C.c_double(thr), _p(keep, C.c_int64))
    return keep[:k].copy()


def iou_match(gt: np.ndarray, anc: np.ndarray, kind: int, ignore_thr: float, want_iou: bool = False):
    gt = np.ascontiguousarray(gt, dtype=np.float32)
    anc = np.ascontiguousarray(anc, dtype=np.float32)
    m, n = gt.shape[0], anc.shape[0]
    best = np.zeros(max(m, 1), np.int64)
    noobj = np.zeros(max(n, 1), np.uint8)
    iou = np.zeros((m, n), np.float32) if want_iou else None
    lib().ref_iou_match(_p(gt, C.c_float), C.c_int(m), _p(anc, C.c_float), C.c_int(n), C.c_int(kind),
                        C.c_float(np.float32(ignore_thr)), _p(best, C.c_int64), _p(noobj, C.c_uint8),
                        _p(iou, C.c_float) if want_iou else None)
    return best[:m].copy(), noobj[:n].astype(bool), iou


def box_iou_tv(b1: np.ndarray, b2: np.ndarray):
    b1 = np.ascontiguousarray(b1, dtype=np.float32)
    b2 = np.ascontiguousarray(b2, dtype=np.float32)
    out = np.zeros((b1.shape[0], b2.shape[0]), np.float32)
    lib().ref_box_iou_tv(_p(b1, C.c_float), C.c_int(b1.shape[0]), _p(b2, C.c_float),
                         C.c_int(b2.shape[0]), _p(out, C.c_float))
    return out
