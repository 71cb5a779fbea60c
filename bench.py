#!/usr/bin/env python
"""bench.py -- images/sec of decode + conf filter + NMS on synthetic YOLOv3-608 COCO head tensors.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (b200_yolo_postprocess_decode: fused decode+filter kernel, then
b200_yolo_postprocess_nms: the NMS kernels) over one batch of 64 images per GPU (weak scaling: every rank owns its own
64-image batch, as the reference's DistributedSampler does; --scaling strong splits the same 64 images over the ranks);
at N > 1 each step ends with the one exchange the path has: b200_exchange_push stores the kept-detection lists straight
into every peer's receive buffer over NVLink, b200_exchange_wait completes when all ranks' lists of the step are there.
Steps are software pipelined: decode kernels rotate over 3 decode streams (so the HBM stream never drains), NMS chains
over 3 other streams, pushes and waits have a stream each; 6 workspaces rotate.  The timed loop is captured into ONE
CUDA graph after the warm-up and timed as one replay.

Printed JSON (one line, rank 0):
  value     whole-job images/s with the head tensors resident in HBM (device-timed, max over ranks)
  roofline  algorithmic bytes per step (one read of the head tensors, B*N*(5+C)*4, per GPU) / the step period, against the
            measured HBM copy bandwidth (MEASURED_PEAKS.json); decode-kernel intervals and the isolated launch as extras
  e2e       same metric through the host-buffer C-ABI entry (pinned host -> device copies of all
            head tensors and device -> host copy of the detections inside the timed region)
  cpu_baseline  the CPU oracle port of the reference path on a bounded sample, all host threads (N = 1 only)
--impl reference: the reference's own CPU implementation is Python and cannot travel to the GPU
box, so the oracle port (oracle/yolo_ref.py: same torch CPU ops in the same order, pinned
bit-exactly to the reference by tests/golden) is timed on the host cores instead.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# one hardware work queue per stream (default 8): the pipeline uses 8 streams + the exchange's spinning wait
# kernels, which must never sit in front of unrelated work in a shared queue
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from object_detectors_b200 import synthetic as syn  # noqa: E402

METRIC = "images/sec decode+NMS, YOLOv3-608 COCO b64"
NO_EXCHANGE = bool(os.environ.get("B200_BENCH_NO_EXCHANGE"))     # diagnostic only: N ranks without the all-gather
IMG, NUM_CLASSES, BATCH = 608, 80, 64
CONF_THR, NMS_THR = 0.1, 0.6
MAX_DET = 256          # kept-detection capacity per image in the exchanged message
CAPACITY = 4096        # candidate slab rows per image (overflow is reported, never silent)

def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # nvidia-smi needs ~100 ms to deliver its first sample, so it is started before the warm-up; every line is
    # stamped on arrival and only those that fall inside [t0, t1] of the timed region are reported

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        inside = [ln for ts, ln in self.lines if t0 is None or (t0 <= ts <= t1 + 0.03)]
        window = "timed region"
        if not inside:                      # region shorter than the sampling period: fall back to the whole run under load
            inside, window = [ln for _, ln in self.lines], "whole run (timed region shorter than one sample)"
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def make_heads(seed: int, batch: int):
    return syn.yolo_heads(seed, batch, IMG, NUM_CLASSES, syn.COCO_ANCHORS, "clustered")


def load_idf():
    return torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "idf_coco_smooth.npy")))


# ----------------------------------------------------------------------------------------- CPU
def workload_config():
    """The workload both arms run (BASELINE.json configs[1]); identical dict in both JSON lines."""
    return {"workload": "C2: YOLOv3-608 COCO-80 decode + conf filter + nms_majority, batch 64, 22743 anchors/image, "
                        "softmax classes x IDF, conf 0.1, NMS 0.6",
            "img_size": IMG, "num_classes": NUM_CLASSES, "batch": BATCH, "conf_thr": CONF_THR, "nms_thr": NMS_THR,
            "nms": "majority", "generator": "clustered", "seed": 1000}


def cpu_reference_pass(heads_cpu, idf):
    """The reference's eval post-process (oracle port, torch CPU ops): decode -> xyxy -> filter ->
    nms_majority.  Returns the number of kept detections."""
    from oracle import yolo_ref
    recs = yolo_ref.postprocess(heads_cpu, syn.COCO_ANCHORS, IMG, NUM_CLASSES, idf, True, CONF_THR, NMS_THR)
    return sum(int(r["keep"].numel()) for r in recs)


def time_cpu(sample_batch: int, reps: int, warm: int = 1, heads_np=None):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    heads = [torch.from_numpy(h) for h in (heads_np if heads_np is not None else make_heads(1000, sample_batch))]
    idf = load_idf()
    for _ in range(warm):
        cpu_reference_pass(heads, idf)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_reference_pass(heads, idf)
        ts.append(time.perf_counter() - t0)
    return sample_batch / float(np.median(ts)), cores, ts


def run_reference(args):
    """The reference arm: the reference's own CPU code path (oracle port, all host threads) on the same workload.
    Each step processes a bounded sample of the 64-image batch, sized so that the whole run (warm-up + steps) is
    about a minute of CPU work at the ~80 images/s this path reaches on the GPU boxes' hosts."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    warm = max(args.warmup, 1)
    sample = max(1, min(BATCH, 4800 // max(args.steps + warm, 1)))
    heads = [torch.from_numpy(h) for h in make_heads(1000, sample)]
    idf = load_idf()
    for _ in range(warm):
        cpu_reference_pass(heads, idf)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(heads, idf)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = (f"{sample} images of the {BATCH}-image 608/COCO batch per step (clustered synthetic heads, seed 1000), "
            f"{args.steps} steps")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(),
        "detail": {"sample_batch": sample, "threads": cores, "engine": "oracle port of the reference (torch CPU ops)"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one fused decode+filter launch on this workload, as written
    by profiles/summarize.py from this round's `ncu --set full` capture (null when no capture is on file)."""
    p = os.path.join(ROOT, "profiles", "r02_decode_traffic.json")
    try:
        d = json.load(open(p))
        return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])
    except Exception:
        return None


# ----------------------------------------------------------------------------------------- GPU
def run_b200(args):
    # NCCL / torch may print banners on stdout; the contract is ONE JSON line there, so fd 1 is pointed at
    # stderr for the duration of the run and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    from object_detectors_b200 import _lib, ops
    from object_detectors_b200.distributed import DetectionExchange, PeerExchange, message_len, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: object_detectors_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    VARIANTS = {"gated": 0, "stream": 1, "ring": 3}
    lib.b200_set_decode_variant(VARIANTS[args.variant])
    lib.b200_debug_set_nms_path({"auto": -1, "general": 1, "fused": 0}[args.nms_path])

    # weak scaling (default, the reference's DistributedSampler data parallelism): every rank owns its own batch of
    # 64; strong scaling: the SAME 64 images split into contiguous blocks of 64 / world
    strong = args.scaling == "strong"
    if strong:
        lo, hi = shard_range(BATCH, rank, world)
        heads_np = [h[lo:hi] for h in make_heads(1000, BATCH)]
    else:
        heads_np = make_heads(1000 + rank, BATCH)
    batch = heads_np[0].shape[0]
    heads = [torch.from_numpy(np.ascontiguousarray(h)).to(dev) for h in heads_np]
    idf = load_idf().to(dev)
    grids = [h.shape[2] for h in heads]
    n_anchor = sum(g * g * 3 for g in grids)
    algo_bytes = batch * n_anchor * (5 + NUM_CLASSES) * 4
    # Software pipeline: decode launches of consecutive steps go round-robin over `dstreams` streams (kernels on
    # one stream are serial, so exactly that many decode kernels are in flight and the HBM stream never drains),
    # the latency-bound NMS chain of a step runs on one of `nstreams` other streams after the step's decode (event),
    # and `plans` workspaces rotate (a workspace is reused only after its detections have been pushed).  At N > 1
    # every step ends with the exchange: push (one-sided stores into every peer's receive buffer) on its own stream,
    # wait (all ranks' messages of the step have arrived) on another.  Every step's work completes inside the timed
    # region (the end event waits for all streams).
    n_d, n_n = max(1, args.dstreams), max(1, args.nstreams)
    n_p = max(args.plans, n_d + 1)
    lib.b200_debug_set_ring(*[int(x) for x in args.ring.split(",")])
    plans = [ops.YoloPostprocess(grids, batch, syn.COCO_ANCHORS, IMG, NUM_CLASSES, True, CONF_THR, NMS_THR,
                                 ops.NMS_MAJORITY, CAPACITY, MAX_DET, dev) for _ in range(n_p)]
    # Each workspace reads its OWN copy of the synthetic batch (images rolled by 7k): decode kernels of neighbouring
    # steps run concurrently and must not find each other's lines in L2; a step's input (495 MB) exceeds L2 (126 MB)
    # and the copy it reads was last touched n_p steps (~3 GB of traffic) ago.
    inputs = [[h.roll(7 * k, 0).contiguous() for h in heads] for k in range(n_p)]
    d_streams = [torch.cuda.Stream(device=dev) for _ in range(n_d)]
    n_streams_ = [torch.cuda.Stream(device=dev, priority=args.nms_priority) for _ in range(n_n)]
    p_stream = torch.cuda.Stream(device=dev)      # pushes, in step order
    w_stream = torch.cuda.Stream(device=dev)      # waits, in step order
    streams = d_streams + n_streams_ + [p_stream, w_stream]
    dec_done = [torch.cuda.Event() for _ in plans]
    free = [torch.cuda.Event() for _ in plans]          # workspace + outputs of plan k may be overwritten
    det_ready = [torch.cuda.Event() for _ in plans]
    mode = "none" if (world == 1 or NO_EXCHANGE) else args.exchange
    exchange = None
    if mode == "p2p":
        try:
            exchange = PeerExchange(batch, MAX_DET, dev, slots=args.slots)
        except Exception as e:          # peer mapping unavailable: same bytes through NCCL
            print(f"[bench] one-sided exchange unavailable ({e}); falling back to ncclAllGather", file=sys.stderr, flush=True)
            mode = "nccl"
    if mode == "nccl":
        exchange = DetectionExchange(batch, MAX_DET, dev, bucket=args.exchange_every)
    serial = [False]          # True: one stream does everything, no exchange (the isolated-kernel measurement)
    steps_done = [0]

    def step(i):
        k = i % n_p
        pl = plans[k]
        if serial[0]:
            pl.decode(inputs[k], idf, d_streams[0])
            pl.nms(d_streams[0])
            return
        d, n = d_streams[i % n_d], n_streams_[i % n_n]
        d.wait_event(free[k])                    # workspace k is free again
        pl.decode(inputs[k], idf, d)
        dec_done[k].record(d)
        n.wait_event(dec_done[k])
        pl.nms(n)
        if mode == "none":
            free[k].record(n)
            return
        det_ready[k].record(n)
        p_stream.wait_event(det_ready[k])
        if mode == "p2p":
            exchange.push(pl.det, pl.det_count, p_stream)
            free[k].record(p_stream)
            exchange.wait(w_stream)              # completes when every rank's message of this step is here
        else:
            exchange(pl.det, pl.det_count, p_stream)        # pack (+ all-gather when the bucket is full)
            free[k].record(p_stream)
        steps_done[0] += 1

    def fence_in():
        ev = torch.cuda.Event()
        ev.record()
        for st in streams:
            st.wait_event(ev)
        for e in free:
            e.record()

    def fence_out():
        if mode == "nccl" and not serial[0]:
            exchange.flush(p_stream)             # partial bucket: nothing stays behind
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)

    def timed_loop(steps, warm, sample_clocks=False):
        """`warm` untimed steps, then exactly `steps` timed ones bracketed by barrier + synchronize; returns
        (total ms, mean decode-kernel ms from per-step CUDA events on the launching stream, union-busy ms per launch),
        max over ranks."""
        fence_in()
        for i in range(warm):
            step(i)
        fence_out()
        torch.cuda.synchronize()
        for pl in plans:
            pl.check_status()
        k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in k_ev:
            a.record(); b.record()          # materialise the cudaEvent_t handles
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sample_clocks:
            window[0] = time.perf_counter()
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin.record()
        fence_in()
        h0 = time.perf_counter()
        for i in range(steps):
            lib.b200_debug_set_decode_events(C.c_void_p(k_ev[i][0].cuda_event), C.c_void_p(k_ev[i][1].cuda_event))
            step(i)
        host_us[0] = (time.perf_counter() - h0) / steps * 1e6      # host enqueue time per step (diagnostic)
        lib.b200_debug_set_decode_events(None, None)
        fence_out()
        t_end.record()
        torch.cuda.synchronize()
        if sample_clocks:
            window[1] = time.perf_counter()
        if world > 1:
            dist.barrier()
        ms = t_begin.elapsed_time(t_end)
        k = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
        # Decode kernels of neighbouring steps run concurrently, so launch-to-end times double count the shared
        # interval.  busy = length of the UNION of the decode kernels' [start, end] intervals / launches.
        iv = sorted((t_begin.elapsed_time(a), t_begin.elapsed_time(b)) for a, b in k_ev)
        busy, cur_a, cur_b = 0.0, iv[0][0], iv[0][1]
        for a_, b_ in iv[1:]:
            if a_ > cur_b:
                busy += cur_b - cur_a
                cur_a, cur_b = a_, b_
            else:
                cur_b = max(cur_b, b_)
        busy = (busy + cur_b - cur_a) / steps
        t = torch.tensor([ms, k, busy], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for pl in plans:
            pl.check_status()
        return float(t[0]), float(t[1]), float(t[2])

    def graph_loop(steps, warm, sample_clocks=False):
        """The same `steps` steps captured ONCE into a CUDA graph (all streams fork from / join the capturing stream
        exactly like fence_in / fence_out) and timed as one replay: the host enqueues one launch for the whole loop.
        Untimed before the measurement: `warm` eager steps, the capture, and one replay (graph upload)."""
        fence_in()
        for i in range(warm):
            step(i)
        fence_out()
        torch.cuda.synchronize()
        for pl in plans:
            pl.check_status()
        cap = torch.cuda.Stream(device=dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cap):
            graph.capture_begin(capture_error_mode="thread_local")
            fence_in()
            for i in range(steps):
                step(i)
            fence_out()
            graph.capture_end()
            graph.replay()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sample_clocks:
            window[0] = time.perf_counter()
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(cap):
            t_begin.record()
            h0 = time.perf_counter()
            graph.replay()
            host_us[0] = (time.perf_counter() - h0) / steps * 1e6
            t_end.record()
        torch.cuda.synchronize()
        if sample_clocks:
            window[1] = time.perf_counter()
        if world > 1:
            dist.barrier()
        t = torch.tensor([t_begin.elapsed_time(t_end)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for pl in plans:
            pl.check_status()
        return float(t[0]), graph

    host_us = [0.0]
    window = [None, None]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)          # let nvidia-smi deliver its first sample before the (sub-second) timed region
    warm = max(args.warmup, 3)
    used_graph = False
    graph = None
    if args.graph and mode != "nccl":
        try:
            ms_total, graph = graph_loop(args.steps, warm, sample_clocks=True)
            used_graph = True
        except Exception as e:          # capture is an optimisation of the host side only: fall back, loudly
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing the eager loop", file=sys.stderr, flush=True)
            torch.cuda.synchronize()
    if not used_graph:
        ms_total, k_ms, k_busy = timed_loop(args.steps, warm, sample_clocks=True)
    clocks = sampler.stop(window[0], window[1]) if rank == 0 else None
    host_enqueue_us = host_us[0]
    last_plan = plans[(args.steps - 1) % n_p]
    kept = int(last_plan.det_count.sum())
    cands = int(last_plan.cand_count.sum())

    # ---- the gathered bytes of the LAST timed step, checked on the GPUs: what arrived through the exchange on this
    #      rank must equal, bit for bit, every rank's own detections (an NCCL all-gather of the packed lists is the
    #      independent witness) -------------------------------------------------------------------------------------
    exchange_verified = None
    if mode == "p2p":
        got = exchange.read(exchange.steps()[0] - 1).reshape(world, batch, -1)     # the device's own step counter
        msg = ops.pack_detections(last_plan.det, last_plan.det_count)
        truth = torch.empty((world, message_len(batch, MAX_DET)), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(truth.view(-1), msg)
        truth = truth.reshape(world, batch, -1)
        cnt_t = truth[:, :, 0].contiguous().view(torch.int32)
        cnt_g = got[:, :, 0].contiguous().view(torch.int32)
        valid = (torch.arange(6 * MAX_DET, device=dev)[None, None, :] < 6 * cnt_t[:, :, None])
        same = torch.equal(cnt_t, cnt_g) and bool(((got[:, :, 1:].view(torch.int32) == truth[:, :, 1:].view(torch.int32)) | ~valid).all())
        flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        exchange_verified = bool(int(flag[0]))
        assert exchange_verified, "one-sided exchange delivered bytes that differ from the ranks' own detections"

    # ---- diagnostic: the same loop enqueued eagerly, with CUDA events around every decode kernel ----------------------
    eager_ms = None
    if used_graph:
        eager_ms, k_ms, k_busy = timed_loop(args.steps, 3)
        eager_host_us = host_us[0]
    else:
        eager_host_us = host_enqueue_us

    # ---- for the record: the decode kernel alone (one stream, nothing overlapping it, no exchange) -----------------
    serial[0] = True
    iso_steps = min(args.steps, 200)
    lib.b200_debug_set_ring(8, 1, 1)       # the stand-alone launch shape: 8 warps per SM, 64-cell tiles
    iso_ms, iso_k, _ = timed_loop(iso_steps, 5)
    lib.b200_debug_set_ring(*[int(x) for x in args.ring.split(",")])
    serial[0] = False
    isolated = {"kernel_ms": iso_k, "achieved_GBs": algo_bytes / (iso_k * 1e-3) / 1e9, "serial_step_ms": iso_ms / iso_steps,
                "note": "ONE launch alone on an idle GPU in its stand-alone shape (8 warps per SM, 64-cell tiles); "
                        "one stream, the NMS kernels follow it serially (serial_step_ms = decode + NMS, one stream)"}

    # ---- e2e through the host-buffer entry point: H2D of every head tensor + D2H of detections --
    heads_pin = [torch.from_numpy(np.ascontiguousarray(h)).pin_memory() for h in heads_np]
    idf_host = load_idf()
    out = (torch.empty((batch, MAX_DET, 6), dtype=torch.float32).pin_memory(),
           torch.empty((batch, MAX_DET), dtype=torch.int32).pin_memory(),
           torch.empty((batch,), dtype=torch.int32).pin_memory(),
           torch.zeros((1,), dtype=torch.int32).pin_memory())
    e2e_steps = max(2, min(args.steps, 10))
    for _ in range(2):
        ops.yolo_postprocess_host(heads_pin, syn.COCO_ANCHORS, IMG, NUM_CLASSES, idf_host, True, CONF_THR, NMS_THR,
                                  ops.NMS_MAJORITY, CAPACITY, MAX_DET, out)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ops.yolo_postprocess_host(heads_pin, syn.COCO_ANCHORS, IMG, NUM_CLASSES, idf_host, True, CONF_THR, NMS_THR,
                                  ops.NMS_MAJORITY, CAPACITY, MAX_DET, out)
    e2e_s = time.perf_counter() - t0
    assert int(out[3][0]) == 0, "e2e slab overflow"
    e2e_kept = int(out[2].sum())
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])
    h2d = sum(h.numel() * 4 for h in heads_pin) + idf_host.numel() * 4
    d2h = sum(o.numel() * o.element_size() for o in out)

    if rank == 0:
        peak, peak_src = _peaks()
        step_ms = ms_total / args.steps
        achieved = algo_bytes / (step_ms * 1e-3) / 1e9
        cpu_v, cores, cpu_ts = time_cpu(sample_batch=BATCH, reps=12) if world == 1 else (None, os.cpu_count() or 1, [])
        nms_kernels = {"auto": 4, "general": 3, "fused": 1}[args.nms_path]
        launches_per_step = 1 + nms_kernels + (2 if mode == "p2p" else 1 if mode == "nccl" else 0)
        exch = {"none": "none (1 GPU)" if world == 1 else "disabled (diagnostic)",
                "p2p": f"one-sided: push kernel stores the kept lists into every peer's receive buffer over NVLink (CUDA IPC "
                       f"mapping), per-(slot, rank) flags, {args.slots} slots; wait kernel per step; gathered bytes of the "
                       "last step verified against an NCCL all-gather",
                "nccl": f"ncclAllGather of fixed-capacity kept lists, {args.exchange_every} steps per bucket"}[mode]
        line = {
            "metric": METRIC, "value": world * batch * args.steps / (ms_total * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(),
            "detail": {"batch_per_gpu": batch, "global_batch": batch * world,
                       "l2_policy": f"inputs ({algo_bytes / 1e6:.1f} MB/step) larger than L2 (126 MB); {n_p} distinct input copies "
                                    "rotate, so concurrent decode kernels never read the same lines; no flush needed",
                       "candidates_per_step": cands, "kept_per_step": kept,
                       "pipeline": {"decode_streams": n_d, "nms_streams": n_n, "nms_stream_priority": args.nms_priority, "workspaces": n_p, "ring": args.ring,
                                    "nms_path": args.nms_path},
                       "decode_variant": args.variant, "exchange": exch, "exchange_verified": exchange_verified},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(), "kernel": "k_decode_filter_ring",
                         "how": "algorithmic bytes per step (one read of the head tensors, per GPU) / the step period "
                                "(timed region / steps, fill and drain included): the decode launches of consecutive steps "
                                "overlap, so the step period IS the decode stage's launch-to-launch time",
                         "step_ms": step_ms, "decode_busy_ms": k_busy, "decode_busy_GBs": algo_bytes / (k_busy * 1e-3) / 1e9,
                         "decode_launch_to_end_ms": k_ms, "algorithmic_bytes": algo_bytes, "peak_source": peak_src,
                         "isolated": isolated},
            "cpu_baseline": ({"value": cpu_v, "unit": "images/s", "cores": cores, "kind": "port",
                              "sample": f"the {BATCH}-image 608/COCO batch, 12 timed passes (median; ~10 s of CPU work), "
                                        "oracle port (torch CPU ops, all host threads)"} if cpu_v is not None else None),
            "e2e": {"value": world * batch * e2e_steps / e2e_s, "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "kept": e2e_kept},
            "gpu_launches": args.steps * launches_per_step,
            "host_enqueue_us_per_step": host_enqueue_us,
            "launch": ("one CUDA graph replay for the whole timed loop (captured after the warm-up; every stream forks from and "
                       "joins the capturing stream)" if used_graph else "eager: every launch enqueued from Python"),
            "eager_loop": ({"ms_per_step": eager_ms / args.steps, "host_enqueue_us_per_step": eager_host_us,
                            "value": world * batch * args.steps / (eager_ms * 1e-3)} if eager_ms is not None else None),
            "clocks": clocks,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if exchange is not None:
        exchange.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 64 images per GPU; strong: the same 64 images split over the GPUs (SURVEY 8e)")
    ap.add_argument("--variant", default="ring", choices=["ring", "gated", "stream"],
                    help="fused decode kernel variant (include/b200det.h: B200_DECODE_*)")
    ap.add_argument("--nms-path", default="auto", choices=["auto", "general", "fused"],
                    help="NMS kernels: auto = the library default (general path for segments of <= 1500 boxes + the "
                         "single-launch path for larger ones), general / fused = one path only")
    ap.add_argument("--dstreams", type=int, default=3, help="decode streams = decode kernels in flight")
    ap.add_argument("--nstreams", type=int, default=3, help="streams for the NMS chains")
    ap.add_argument("--nms-priority", type=int, default=-5,
                    help="CUDA stream priority of the NMS streams (lower = more urgent; clamped to the device's range): their CTAs are "
                         "placed before those of the next decode launch, which shortens the NMS chains (+4 %% at 20 steps)")
    ap.add_argument("--plans", type=int, default=6, help="rotating workspaces")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: one-sided peer-memory exchange (default) or bucketed ncclAllGather")
    ap.add_argument("--slots", type=int, default=32, help="p2p exchange: receive slots per rank (steps a rank may run ahead)")
    ap.add_argument("--exchange-every", type=int, default=4,
                    help="nccl exchange: steps per all-gather bucket (1 = gather after every step)")
    ap.add_argument("--graph", type=int, default=1, help="1: time one CUDA-graph replay of the loop (default); 0: eager launches")
    ap.add_argument("--ring", default="4,1,101", help="RING decode: warps per CTA, stages per warp, CTAs per SM (+100: 32-cell tiles)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
