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


def _two_rank_worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        from object_detectors_b200 import ops, synthetic as syn
        from object_detectors_b200.distributed import PeerExchange, message_len
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        batch, img, c, max_det = 4, 416, 80, 128
        heads = [torch.from_numpy(h).to(dev) for h in syn.yolo_heads(700 + rank, batch, img, c, syn.COCO_ANCHORS, "clustered")]
        plan = ops.YoloPostprocess([h.shape[2] for h in heads], batch, syn.COCO_ANCHORS, img, c, True, 0.1, 0.6,
                                   ops.NMS_MAJORITY, 2048, max_det, dev)
        x = PeerExchange(batch, max_det, dev, slots=3)
        st = torch.cuda.Stream(device=dev)
        ok = True
        for step in range(7):                       # > slots: receive slots are reused
            hs = [h.roll(step, 0).contiguous() for h in heads]
            plan(hs, None)
            plan.check_status()
            st.wait_stream(torch.cuda.current_stream(dev))
            x.push(plan.det, plan.det_count, st)
            x.wait(st)
            got = x.read(step, st)
            st.synchronize()
            # independent witness: NCCL all-gather of the packed message
            msg = ops.pack_detections(plan.det, plan.det_count)
            truth = torch.empty((world, message_len(batch, max_det)), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(truth.view(-1), msg)
            g, t = got.cpu().numpy().reshape(world, batch, -1), truth.cpu().numpy().reshape(world, batch, -1)
            for r in range(world):
                for i in range(batch):
                    k = int(t[r, i, :1].view(np.int32)[0])
                    ok &= int(g[r, i, :1].view(np.int32)[0]) == k and k > 0
                    ok &= np.array_equal(g[r, i, 1:1 + 6 * k].view(np.int32), t[r, i, 1:1 + 6 * k].view(np.int32))
        x.close()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_exchange_two_ranks_bit_exact():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err in results:
        assert ok, f"rank {rank}: {err or 'gathered bytes differ from the NCCL all-gather'}"
