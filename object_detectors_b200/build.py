"""Builds object_detectors_b200/csrc/libb200det.so with nvcc for sm_100a (in-tree, so the binary
travels with the repository snapshot to the GPU box).

    python -m object_detectors_b200.build [--force] [--verbose]

Flags that matter for parity: ``--fmad=false`` (no FMA contraction: threshold decisions must see
the same fp32 roundings as the reference's unfused tensor ops; the kernels additionally use
explicit ``__f*_rn`` intrinsics) and no ``--use_fast_math`` (IEEE division, accurate expf,
denormals kept).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libb200det.so")
SOURCES = ("api.cu", "decode.cu", "decode_ring.cu", "nms.cu", "nms_fused.cu", "iou.cu", "rpn.cu", "misc.cu", "roi.cu", "exchange.cu", "emit.cu", "match.cu")
HEADERS = ("common.cuh", "decode.cuh", "decode_ring.cuh", "nms.cuh", "nms_dev.cuh", os.path.join("..", "..", "include", "b200det.h"))

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--fmad=false",
    "-Xcompiler", "-fPIC,-O2",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200det.so cannot be built (no CPU fallback exists)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for s in SOURCES:
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- nvcc {s} (exit {p.returncode})\n{out}", flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed, see output above")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-Xcompiler", "-fPIC", "-o", LIB, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
