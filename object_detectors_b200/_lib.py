"""ctypes binding of libb200det.so (include/b200det.h).  No fallback: a missing library or a
non-zero status raises."""
from __future__ import annotations

import ctypes as C
import os

MAX_SCALES = 4
MAX_ANCHORS = 8

NMS_MAJORITY, NMS_TV, NMS_TV_CLASS, NMS_TV_TRICK, NMS_TV_AUTO = 0, 1, 2, 3, 4
IOU, GIOU, DIOU, CIOU, IOU_TV = 0, 1, 2, 3, 4
DECODE_GATED, DECODE_STREAM, DECODE_RING = 0, 1, 3

LIB_PATH = os.environ.get("B200DET_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc",
                                                         "libb200det.so")


class YoloLayout(C.Structure):
    _fields_ = [
        ("num_scales", C.c_int32), ("num_anchors", C.c_int32), ("num_classes", C.c_int32),
        ("batch", C.c_int32), ("softmax", C.c_int32), ("img_size", C.c_float),
        ("grid", C.c_int32 * MAX_SCALES),
        ("anchor_rel", ((C.c_float * 2) * MAX_ANCHORS) * MAX_SCALES),
    ]


_p = C.c_void_p
_i32, _i64, _f32, _f64, _sz = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_size_t
_LP = C.POINTER(YoloLayout)
_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/b200det.h declares
SIGNATURES = {
    "b200_abi_version": (C.c_int, []),
    "b200_error_string": (C.c_char_p, [C.c_int]),
    "b200_device_info": (C.c_int, [_p, _p, _p]),
    "b200_yolo_decode_dense": (C.c_int, [_LP, _PP, _p, _p, _p]),
    "b200_yolo_workspace_bytes": (_sz, [_LP, _i32]),
    "b200_yolo_decode_filter": (C.c_int, [_LP, _PP, _p, _f32, _i32, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "b200_yolo_postprocess": (C.c_int, [_LP, _PP, _p, _f32, _f64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p,
                                        _p, _sz, _p]),
    "b200_yolo_postprocess_decode": (C.c_int, [_LP, _PP, _p, _f32, _i32, _p, _p, _sz, _p]),
    "b200_yolo_postprocess_nms": (C.c_int, [_LP, _f64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "b200_yolo_postprocess_host": (C.c_int, [_LP, _PP, _p, _f32, _f64, _i32, _i32, _i32, _p, _p, _p, _p]),
    "b200_debug_set_decode_events": (C.c_int, [_p, _p]),
    "b200_debug_set_timeline": (C.c_int, [_p, _p, _p]),
    "b200_debug_set_resolve_prof": (C.c_int, [_p]),
    "b200_debug_set_rpn_prof": (C.c_int, [_p]),
    "b200_debug_set_serial_split": (C.c_int, [C.c_int]),
    "b200_debug_set_nms_path": (C.c_int, [C.c_int]),
    "b200_debug_set_resolve": (C.c_int, [C.c_int, C.c_int]),
    "b200_set_decode_variant": (C.c_int, [C.c_int]),
    "b200_debug_set_ring": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "b200_yolo_legacy_decode": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _p, _p, _p]),
    "b200_roi_workspace_bytes": (_sz, [_i32, _i32]),
    "b200_roi_postprocess": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _i32, C.POINTER(C.c_float), _f32, _f32,
                                       _f32, _f64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "b200_set_batched_nms_auto_limit": (C.c_int, [_i64]),
    "b200_nms_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "b200_nms": (C.c_int, [_p, _p, _p, _p, _i32, _i64, _i32, _f64, _i32, _p, _p, _p, _p, _sz, _p]),
    "b200_box_iou": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _p, _p]),
    "b200_box_iou_paired": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "b200_box_iou_paired_backward": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "b200_iou_match_workspace_bytes": (_sz, [_i32, _i32]),
    "b200_iou_match": (C.c_int, [_p, _p, _i32, _i32, _p, _i32, _i32, _f32, _p, _p, _p, _sz, _p]),
    "b200_rpn_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "b200_rpn_filter": (C.c_int, [_p, _p, _p, _i32, _i32, C.POINTER(C.c_int32), _i32, _p, _i32, _i32, _f64,
                                  _f32, _f32, _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "b200_rpn_top_n_idx_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "b200_rpn_top_n_idx": (C.c_int, [_p, _i32, _i32, C.POINTER(C.c_int32), _i32, _i32, _p, _p, _sz, _p]),
    "b200_rpn_filter_proposals": (C.c_int, [_p, _p, _i32, _i32, C.POINTER(C.c_int32), _i32, _p, _i32, _i32, _f64,
                                            _f32, _f32, _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "b200_abs_coord": (C.c_int, [_p, _i64, _p, _p]),
    "b200_boxcoder_decode": (C.c_int, [_p, _p, _i64, _i32, C.POINTER(C.c_float), _f32, _p, _p]),
    "b200_boxcoder_encode": (C.c_int, [_p, _p, _i64, C.POINTER(C.c_float), _p, _p]),
    "b200_matcher": (C.c_int, [_p, _i32, _i32, _f32, _f32, _i32, _p, _p, _sz, _p]),
    "b200_retinanet_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "b200_retinanet_postprocess": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, C.POINTER(C.c_int32), _i32, _p, _p, _i32, _f32, _f64, _i32,
                                             _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "b200_ssd_postprocess": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p, _p, C.POINTER(C.c_float), _f32, _f32, _i32, _f64, _i32,
                                       _i32, _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "b200_match_boxes_workspace_bytes": (_sz, [_i32, _i32]),
    "b200_match_boxes": (C.c_int, [_p, _i32, _p, _i32, _f32, _f32, _i32, _i32, _p, _p, _p, _sz, _p]),
    "b200_matcher_ssd_override": (C.c_int, [_p, _i32, _i32, _p, _p, _sz, _p]),
    "b200_clip_boxes_to_image": (C.c_int, [_p, _i64, _f32, _f32, _p, _p]),
    "b200_remove_small_boxes": (C.c_int, [_p, _i32, _f32, _p, _p, _p]),
    "b200_emit_results": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _f32, _p, _i32, _i32, _p, _p, _p, _p, _p]),
    "b200_pack_detections": (C.c_int, [_p, _p, _i32, _i32, _p, _p]),
    "b200_allgather_dets": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _p]),
    "b200_exchange_create": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _PP]),
    "b200_exchange_handle": (C.c_int, [_p, _p]),
    "b200_exchange_connect": (C.c_int, [_p, _i32, _p]),
    "b200_exchange_push": (C.c_int, [_p, _p, _p, _p]),
    "b200_exchange_wait": (C.c_int, [_p, _p]),
    "b200_exchange_message": (_p, [_p, _i64, _i32]),
    "b200_exchange_read": (C.c_int, [_p, _i64, _p, _p]),
    "b200_exchange_steps": (C.c_int, [_p, _p, _p]),
    "b200_exchange_destroy": (C.c_int, [_p]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the in-tree CUDA library.  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m object_detectors_b200.build` "
            "(nvcc, sm_100a).  object_detectors_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    if lib.b200_abi_version() != 1:
        raise RuntimeError("libb200det.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (status {rc})")
