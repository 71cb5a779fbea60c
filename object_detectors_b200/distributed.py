"""Multi-GPU plumbing: images are sharded across ranks (one process per GPU, like the reference's
DistributedSampler data parallelism, yolo/procedures/init_dataset.py:82-83) and the path has exactly
one exchange step at its end -- an all-gather of the variable-length kept-detection lists -- which replaces the
reference's per-rank pickle files + barrier (yolo/procedures/eval_results.py:12-31, yolo/main.py:102-105) and its
size-exchange/pad/gather ``utils.all_gather`` (torchvision_models/detection/utils.py:75-115).

``PeerExchange`` is the default: one-sided stores into the peers' receive buffers over NVLink (csrc/exchange.cu),
no collective kernel.  ``DetectionExchange`` is the NCCL form (pack kernel + bucketed ``ncclAllGather``), kept for
setups without peer mapping and for the gloo CPU tests of the host-side bookkeeping.
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


def unpack_bucket(gathered: torch.Tensor, world: int, steps: int, batch: int, max_det: int) -> List[List[List[torch.Tensor]]]:
    """A gathered bucket ``[world, steps, message_len]`` -> ``lists[rank][step][image]`` of ``[k, 6]`` tensors."""
    n = message_len(batch, max_det)
    g = gathered.reshape(world, steps, n)
    return [[unpack_detections(g[r, s], 1, batch, max_det) for s in range(steps)] for r in range(world)]


class DetectionExchange:
    """The path's one exchange step for a fixed (batch, max_det): pack kernel + all-gather, enqueued on the
    caller's stream with plain ctypes calls (~5 us of host time; ``all_gather_into_tensor`` costs ~10x that in
    Python/c10d dispatch, which at ~90 us per step is what decides multi-GPU scaling).

    Messages are bucketed: every step's kept lists are packed into the next slot of a bucket on the device and
    one ``ncclAllGather`` moves ``bucket`` steps at a time (``flush`` sends a partial bucket), so the ranks
    synchronise -- and pay the collective's launch + 8-hop latency -- once per bucket instead of once per
    step.  The payload is tiny (393 KB per rank and step for 64 images x 256 detections), NVLink bandwidth is
    never the issue; latency and rank skew are.  ``bucket=1`` gathers after every step.

    The collective is NCCL's ``ncclAllGather`` on a communicator of our own, created through the NCCL library
    the process has already loaded (torch's), bootstrapped over the existing process group.  If that library
    cannot be bound the exchange falls back to ``torch.distributed`` (same bytes, more host time).
    """

    NCCL_FLOAT = 7

    def __init__(self, batch: int, max_det: int, device, bucket: int = 1, pack_fn=None):
        """``pack_fn(det, det_count, dst_message_view, stream)``: replaces the CUDA pack kernel; it exists so that
        the bucket / flush / gather bookkeeping can be tested on the CPU with gloo (tests/test_dist_gloo.py)."""
        import ctypes as C
        self.batch, self.max_det, self.bucket = batch, max_det, max(1, int(bucket))
        self.dev = torch.device(device)
        self.pack_fn = pack_fn
        if pack_fn is None:
            from . import _lib
            self.lib = _lib.load()
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.n = message_len(batch, max_det)
        # two buckets alternate: one is being filled while the previous one may still be read by a consumer
        self.msg = [torch.zeros((self.bucket * self.n,), dtype=torch.float32, device=self.dev) for _ in range(2)]
        self.out = [torch.zeros((self.world * self.bucket * self.n,), dtype=torch.float32, device=self.dev) for _ in range(2)]
        self.fill, self.cur, self.gathers = 0, 0, 0
        self.nccl, self.comm = None, None
        if self.world > 1 and self.dev.type == "cuda":
            try:
                self._init_nccl(C)
            except Exception as e:      # keep working through c10d
                print(f"[b200det] direct NCCL binding unavailable ({e}); using torch.distributed", flush=True)
                self.nccl, self.comm = None, None

    def _init_nccl(self, C):
        nccl = C.CDLL("libnccl.so.2")

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_byte * 128)]

        nccl.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
        nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        nccl.ncclAllGather.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        nccl.ncclCommDestroy.argtypes = [C.c_void_p]
        uid = UniqueId()
        if self.rank == 0 and nccl.ncclGetUniqueId(C.byref(uid)) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).to(self.dev)
        dist.broadcast(t, src=0)
        C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        comm = C.c_void_p()
        torch.cuda.synchronize(self.dev)
        if nccl.ncclCommInitRank(C.byref(comm), self.world, uid, self.rank) != 0:
            raise RuntimeError("ncclCommInitRank failed")
        self.nccl, self.comm = nccl, comm

    def _gather(self, stream: torch.cuda.Stream, steps: int) -> torch.Tensor:
        import ctypes as C
        msg, out = self.msg[self.cur], self.out[self.cur]
        count = steps * self.n
        if self.world > 1:
            if self.nccl is not None:
                rc = self.nccl.ncclAllGather(C.c_void_p(msg.data_ptr()), C.c_void_p(out.data_ptr()), count, self.NCCL_FLOAT,
                                             self.comm, C.c_void_p(stream.cuda_stream))
                if rc != 0:
                    raise RuntimeError(f"ncclAllGather failed ({rc})")
            elif self.dev.type == "cuda":
                with torch.cuda.stream(stream):
                    dist.all_gather_into_tensor(out[:self.world * count], msg[:count])
            else:
                dist.all_gather_into_tensor(out[:self.world * count], msg[:count])
        self.gathers += 1
        self.fill, self.cur = 0, self.cur ^ 1
        return out[:self.world * count] if self.world > 1 else msg[:count]

    def __call__(self, det: torch.Tensor, det_count: torch.Tensor, stream: torch.cuda.Stream):
        """Packs one step's detections into the current bucket on ``stream``; when the bucket is full the whole
        bucket is all-gathered on the same stream and the gathered buffer is returned
        (``[world, steps, message_len]``, rank-major), else ``None``."""
        if self.pack_fn is not None:
            self.pack_fn(det, det_count, self.msg[self.cur][self.fill * self.n:(self.fill + 1) * self.n], stream)
        else:
            import ctypes as C
            from . import _lib
            dst = self.msg[self.cur].data_ptr() + 4 * self.fill * self.n
            _lib.check(self.lib.b200_pack_detections(C.c_void_p(det.data_ptr()), C.c_void_p(det_count.data_ptr()),
                                                     self.batch, self.max_det, C.c_void_p(dst),
                                                     C.c_void_p(stream.cuda_stream)), "b200_pack_detections")
        self.fill += 1
        if self.fill == self.bucket:
            return self._gather(stream, self.bucket)
        return None

    def flush(self, stream: torch.cuda.Stream):
        """All-gathers a partially filled bucket (end of an epoch / of the timed region)."""
        if self.fill:
            return self._gather(stream, self.fill)
        return None

    def close(self):
        if self.nccl is not None and self.comm is not None:
            self.nccl.ncclCommDestroy(self.comm)
            self.comm = None


class PeerExchange:
    """One-sided form of the exchange (``b200_exchange_*``, csrc/exchange.cu): every rank maps its peers' receive
    buffers through CUDA IPC; ``push`` stores this rank's kept lists straight into all of them over NVLink and
    releases a flag, ``wait`` completes when every rank's message of the next step has arrived.  No collective
    kernel runs and no rank waits for another one between pushes (flow control aside), so a late rank delays
    nobody's compute.  One process per GPU on one node; ``torch.distributed`` is used once, to hand the 64-byte
    IPC handles around.

    Pushes must be enqueued in step order on streams that order them (use one stream for all pushes), and so
    must waits."""

    def __init__(self, batch: int, max_det: int, device, slots: int = 32):
        import ctypes as C
        from . import _lib
        self.C, self._lib = C, _lib
        self.lib = _lib.load()
        self.batch, self.max_det, self.slots = int(batch), int(max_det), int(slots)
        self.dev = torch.device(device)
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.n = message_len(batch, max_det)
        self.ctx = C.c_void_p()
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.b200_exchange_create(self.rank, self.world, self.batch, self.max_det, self.slots,
                                                     C.byref(self.ctx)), "b200_exchange_create")
            if self.world > 1:
                buf = (C.c_ubyte * 64)()
                _lib.check(self.lib.b200_exchange_handle(self.ctx, buf), "b200_exchange_handle")
                mine = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).to(self.dev)
                every = torch.empty((self.world, 64), dtype=torch.uint8, device=self.dev)
                dist.all_gather_into_tensor(every.view(-1), mine)
                every = every.cpu().numpy()
                for p in range(self.world):
                    if p == self.rank:
                        continue
                    h = (C.c_ubyte * 64)(*every[p].tolist())
                    _lib.check(self.lib.b200_exchange_connect(self.ctx, p, h), f"b200_exchange_connect({p})")
                dist.barrier()          # every rank has mapped every buffer before anyone pushes
        self.pushed = self.waited = 0

    def push(self, det: torch.Tensor, det_count: torch.Tensor, stream: torch.cuda.Stream):
        C = self.C
        self._lib.check(self.lib.b200_exchange_push(self.ctx, C.c_void_p(det.data_ptr()), C.c_void_p(det_count.data_ptr()),
                                                    C.c_void_p(stream.cuda_stream)), "b200_exchange_push")
        self.pushed += 1

    def wait(self, stream: torch.cuda.Stream):
        self._lib.check(self.lib.b200_exchange_wait(self.ctx, self.C.c_void_p(stream.cuda_stream)), "b200_exchange_wait")
        self.waited += 1

    def read(self, step: int, stream: torch.cuda.Stream = None) -> torch.Tensor:
        """Gathered messages of ``step`` as ``[world, message_len]`` (rank-major), copied out of the receive slot."""
        out = torch.empty((self.world, self.n), dtype=torch.float32, device=self.dev)
        st = stream if stream is not None else torch.cuda.current_stream(self.dev)
        self._lib.check(self.lib.b200_exchange_read(self.ctx, int(step), self.C.c_void_p(out.data_ptr()),
                                                    self.C.c_void_p(st.cuda_stream)), "b200_exchange_read")
        return out

    def steps(self):
        """(steps pushed, steps waited for) as the device counts them (host-synchronous)."""
        a, b = self.C.c_int64(0), self.C.c_int64(0)
        self._lib.check(self.lib.b200_exchange_steps(self.ctx, self.C.byref(a), self.C.byref(b)), "b200_exchange_steps")
        return int(a.value), int(b.value)

    def close(self):
        if self.ctx:
            self.lib.b200_exchange_destroy(self.ctx)
            self.ctx = self.C.c_void_p()
