"""Multi-GPU plumbing: images are sharded across ranks (one process per GPU, like the reference's
DistributedSampler data parallelism, yolo/procedures/init_dataset.py:82-83) and the path has exactly
one exchange step at its end -- an all-gather of the variable-length kept-detection lists, sent as
one fixed-capacity message per rank (counts travel inside the payload) so a single NCCL collective
replaces the reference's per-rank pickle files + barrier (yolo/procedures/eval_results.py:12-31,
yolo/main.py:102-105) and its size-exchange/pad/gather ``utils.all_gather``
(torchvision_models/detection/utils.py:75-115).

The collective itself is ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests); the packing of
the message is a CUDA kernel (b200_pack_detections).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(num_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of images owned by ``rank`` (first ``num_images % world`` ranks get one more)."""
    base, extra = divmod(num_images, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def message_len(batch: int, max_det: int) -> int:
    return batch * (1 + max_det * 6)


def all_gather_detections(message: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """``message``: this rank's packed detections (``ops.pack_detections``).  Returns
    ``[world * len(message)]``; a no-op copy when no process group is initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return message if out is None else out.copy_(message)
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * message.numel(),), dtype=message.dtype, device=message.device)
    dist.all_gather_into_tensor(out, message)
    return out


def unpack_detections(gathered: torch.Tensor, world: int, batch: int, max_det: int) -> List[torch.Tensor]:
    """Inverse of the pack kernel on the host: list of ``[k_i, 6]`` tensors, images ordered rank-major
    (rank 0's images first), each in descending score."""
    stride = 1 + max_det * 6
    rows = gathered.reshape(world * batch, stride).cpu()
    counts = rows[:, 0].contiguous().view(torch.int32).tolist()        # count is stored as an int bit pattern
    return [rows[i, 1:1 + 6 * k].reshape(k, 6) for i, k in enumerate(counts)]
