"""CPU-only checks of the C-ABI boundary: the library builds/loads, exports every symbol that
include/b200det.h declares, and the product refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from object_detectors_b200 import _lib, build
    build.build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200det.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from object_detectors_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200det.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared


def test_abi_version_and_error_strings(lib):
    assert lib.b200_abi_version() == 1
    assert lib.b200_error_string(0) == b"ok"
    assert b"workspace" in lib.b200_error_string(-3)


def test_layout_struct_matches_header(lib):
    from object_detectors_b200 import _lib
    # 5 int32 + float + 4 int32 + 4*8*2 float
    assert C.sizeof(_lib.YoloLayout) == 4 * (5 + 1 + 4 + 64)
    from object_detectors_b200 import ops, synthetic as syn
    lay = ops.make_layout([19, 38, 76], 64, syn.COCO_ANCHORS, 608, 80, True)
    assert lay.num_anchors == 3 and lay.grid[2] == 76
    # reference value of cxypwh[0, 2:4] at 608: 0.1908, 0.1480
    assert abs(lay.anchor_rel[0][0][0] - 0.1908) < 1e-4 and abs(lay.anchor_rel[0][0][1] - 0.1480) < 1e-4
    nbytes = lib.b200_yolo_workspace_bytes(C.byref(lay), 4096)
    assert nbytes > 64 * 4096 * 32


def test_invalid_arguments_are_rejected_without_touching_the_gpu(lib):
    assert lib.b200_nms(None, None, None, None, -1, 0, 0, 0.5, 1, None, None, None, None, 0, None) == -1
    assert lib.b200_box_iou(None, 3, None, 3, 9, 0, None, None) == -1
    assert lib.b200_yolo_workspace_bytes(None, 16) == 0
    assert lib.b200_clip_boxes_to_image(None, 5, 4.0, 4.0, None, None) == -1
    assert lib.b200_clip_boxes_to_image(None, 0, 4.0, 4.0, None, None) == 0          # empty input: nothing to do
    assert lib.b200_remove_small_boxes(None, 5, 1.0, None, None, None) == -1
    assert lib.b200_match_boxes_workspace_bytes(0, 10) == 0
    assert lib.b200_debug_set_serial_split(1500) == 0 and lib.b200_debug_set_nms_path(-1) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    from object_detectors_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.box_iou(torch.zeros(2, 4), torch.zeros(3, 4))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.nms_segments(torch.zeros(2, 4), torch.zeros(2), None, torch.zeros(2, dtype=torch.int32), 0.5, 1)
    assert lib.b200_device_info(None, None, None) == -2     # B200_ERR_CUDA, no device


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "object_detectors_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), \
                    f"{f} imports the oracle: product code must never route through it"


def test_header_is_plain_c():
    """include/b200det.h is the drop-in boundary: it must compile as C99 on its own (no C++ or CUDA types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "b200det.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_python_file_compiles():
    """bench.py, __graft_entry__.py, benchmarks/ and the package parse (guards the scripts that only run on a GPU box)."""
    import py_compile
    for top in ("bench.py", "__graft_entry__.py"):
        py_compile.compile(os.path.join(ROOT, top), doraise=True)
    for sub in ("benchmarks", "object_detectors_b200", "oracle", "profiles"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith(".py"):
                    py_compile.compile(os.path.join(dirpath, f), doraise=True)
