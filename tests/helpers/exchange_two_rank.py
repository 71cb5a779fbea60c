"""Worker of tests/test_gpu_exchange.py::test_exchange_two_ranks_bit_exact, launched with torch.distributed.run
(one process per GPU): every rank post-processes its own batches, pushes the kept lists through the one-sided
exchange and compares what it received with an NCCL all-gather of the same packed lists, bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from object_detectors_b200 import ops, synthetic as syn  # noqa: E402
from object_detectors_b200.distributed import PeerExchange, message_len  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    batch, img, c, max_det = 4, 416, 80, 512
    heads = [torch.from_numpy(h).to(dev) for h in syn.yolo_heads(700 + rank, batch, img, c, syn.COCO_ANCHORS, "clustered")]
    plan = ops.YoloPostprocess([h.shape[2] for h in heads], batch, syn.COCO_ANCHORS, img, c, True, 0.1, 0.6,
                               ops.NMS_MAJORITY, 2048, max_det, dev)
    x = PeerExchange(batch, max_det, dev, slots=3)
    st = torch.cuda.Stream(device=dev)
    total = 0
    for step in range(7):                       # > slots: receive slots are reused
        hs = [h.roll(step, 0).contiguous() for h in heads]
        plan(hs, None)
        plan.check_status()
        st.wait_stream(torch.cuda.current_stream(dev))
        x.push(plan.det, plan.det_count, st)
        x.wait(st)
        got = x.read(step, st)
        st.synchronize()
        msg = ops.pack_detections(plan.det, plan.det_count)          # independent witness: NCCL all-gather
        truth = torch.empty((world, message_len(batch, max_det)), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(truth.view(-1), msg)
        g, t = got.cpu().numpy().reshape(world, batch, -1), truth.cpu().numpy().reshape(world, batch, -1)
        for r in range(world):
            for i in range(batch):
                k = int(t[r, i, :1].view(np.int32)[0])
                assert int(g[r, i, :1].view(np.int32)[0]) == k, f"step {step} rank {r} image {i}: count differs"
                assert np.array_equal(g[r, i, 1:1 + 6 * k].view(np.int32), t[r, i, 1:1 + 6 * k].view(np.int32)), \
                    f"step {step} rank {r} image {i}: rows differ"
                total += k
    assert total > 0
    x.close()
    dist.destroy_process_group()
    print(f"EXCHANGE_OK rank {rank}: {total} gathered rows identical", flush=True)


if __name__ == "__main__":
    main()
