"""GPU tests of the exchange step (csrc/exchange.cu): the pack kernel, the one-sided push / wait on one rank
(slot reuse, flow control) and -- when the box has at least two GPUs -- two ranks over CUDA IPC / NVLink, whose
gathered bytes are compared bit for bit with every rank's own detections (an NCCL all-gather of the same
tensors is the independent witness)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _random_dets(seed, batch, max_det):
    g = np.random.default_rng(seed)
    det = g.standard_normal((batch, max_det, 6)).astype(np.float32)
    cnt = g.integers(0, max_det + 1, size=(batch,)).astype(np.int32)
    cnt[0] = 0
    cnt[-1] = max_det
    return det, cnt


def _expect_message(det, cnt, max_det, zero_tail):
    b = det.shape[0]
    msg = np.zeros((b, 1 + 6 * max_det), np.float32)
    for i in range(b):
        msg[i, 0] = np.array([cnt[i]], np.int32).view(np.float32)[0]
        msg[i, 1:1 + 6 * cnt[i]] = det[i, :cnt[i]].reshape(-1)
    return msg


def test_pack_detections():
    from object_detectors_b200 import ops
    det, cnt = _random_dets(5, 9, 17)
    msg = ops.pack_detections(torch.from_numpy(det).cuda(), torch.from_numpy(cnt).cuda()).cpu().numpy().reshape(9, -1)
    want = _expect_message(det, cnt, 17, True)
    np.testing.assert_array_equal(msg.view(np.int32), want.view(np.int32))
    # counts above the capacity are clipped
    cnt2 = cnt.copy(); cnt2[3] = 40
    msg = ops.pack_detections(torch.from_numpy(det).cuda(), torch.from_numpy(cnt2).cuda()).cpu().numpy().reshape(9, -1)
    assert msg[3, :1].view(np.int32)[0] == 17


def test_exchange_single_rank_slot_reuse():
    """world = 1: push / wait / read through more steps than there are slots (slot reuse + flow control),
    every gathered message equal to the rows that exist."""
    from object_detectors_b200.distributed import PeerExchange, unpack_detections
    batch, max_det, slots = 6, 11, 4
    x = PeerExchange(batch, max_det, "cuda", slots=slots)
    st = torch.cuda.Stream()
    try:
        for step in range(3 * slots + 1):
            det, cnt = _random_dets(100 + step, batch, max_det)
            d, c = torch.from_numpy(det).cuda(), torch.from_numpy(cnt).cuda()
            torch.cuda.synchronize()
            x.push(d, c, st)
            x.wait(st)
            got = x.read(step, st)
            st.synchronize()
            lists = unpack_detections(got.reshape(-1), 1, batch, max_det)
            for i in range(batch):
                assert lists[i].shape[0] == cnt[i]
                np.testing.assert_array_equal(lists[i].numpy(), det[i, :cnt[i]])
    finally:
        x.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_exchange_two_ranks_bit_exact():
    """Two processes, one GPU each, launched with torch.distributed.run (tests/helpers/exchange_two_rank.py)."""
    import subprocess
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "helpers", "exchange_two_rank.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("EXCHANGE_OK") == 2, r.stdout[-3000:]
