"""CPU test of the N>1 host logic: image sharding and the fixed-capacity all-gather of kept lists,
on the gloo backend with world_size 2 (the GPU run uses NCCL with the same code path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from object_detectors_b200 import distributed as b200_dist

BATCH, MAX_DET = 3, 5


def _pack_host(dets, max_det):
    """Host restatement of the b200_pack_detections layout: [count-as-int-bits, max_det*6 floats] per image."""
    msg = np.zeros((len(dets), 1 + max_det * 6), np.float32)
    for i, d in enumerate(dets):
        k = min(len(d), max_det)
        msg[i, 0] = np.array([k], np.int32).view(np.float32)[0]
        msg[i, 1:1 + 6 * k] = d[:k].reshape(-1)
    return torch.from_numpy(msg.reshape(-1))


def _fake(rank):
    g = np.random.Generator(np.random.PCG64(100 + rank))
    return [g.random((int(g.integers(0, MAX_DET + 1)), 6)).astype(np.float32) for _ in range(BATCH)]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        msg = _pack_host(_fake(rank), MAX_DET)
        assert msg.numel() == b200_dist.message_len(BATCH, MAX_DET)
        gathered = b200_dist.all_gather_detections(msg)
        lists = b200_dist.unpack_detections(gathered, world, BATCH, MAX_DET)
        want = [d for r in range(world) for d in _fake(r)]
        assert len(lists) == world * BATCH
        for got, w in zip(lists, want):
            np.testing.assert_array_equal(got.numpy(), w)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_all_gather_kept_lists_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def test_shard_ranges_cover_the_batch():
    for n in (64, 63, 7, 1, 0):
        for world in (1, 2, 4, 8):
            spans = [b200_dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    msg = _pack_host(_fake(0), MAX_DET)
    out = b200_dist.all_gather_detections(msg)
    lists = b200_dist.unpack_detections(out, 1, BATCH, MAX_DET)
    for got, w in zip(lists, _fake(0)):
        np.testing.assert_array_equal(got.numpy(), w)


# ------------------------------------------------------------------ bucketed exchange (DetectionExchange)
def _fake_step(rank, step):
    g = np.random.Generator(np.random.PCG64(1000 + 17 * rank + step))
    return [g.random((int(g.integers(0, MAX_DET + 1)), 6)).astype(np.float32) for _ in range(BATCH)]


def _bucket_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def pack(det, det_count, dst, stream):        # host stand-in for the CUDA pack kernel, same layout
            dst.copy_(_pack_host(det, MAX_DET))
        ex = b200_dist.DetectionExchange(BATCH, MAX_DET, "cpu", bucket=2, pack_fn=pack)
        outs = []
        for step in range(3):                          # one full bucket (2 steps) + a partial one
            r = ex(_fake_step(rank, step), None, None)
            if r is not None:
                outs.append((r.clone(), 2, step - 1))
        r = ex.flush(None)
        assert r is not None and ex.flush(None) is None
        outs.append((r.clone(), 1, 2))
        assert ex.gathers == 2
        for gathered, steps, first in outs:
            lists = b200_dist.unpack_bucket(gathered, world, steps, BATCH, MAX_DET)
            for rr in range(world):
                for s in range(steps):
                    for got, want in zip(lists[rr][s], _fake_step(rr, first + s)):
                        np.testing.assert_array_equal(got.numpy(), want)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_bucketed_exchange_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_bucket_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)
