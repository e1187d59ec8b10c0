"""Data-parallel descriptor extraction over point clouds.

The reference is single-process / single-device (eval/evaluate.py:454-466); each cloud's forward is independent
of every other (eval-mode BatchNorm is affine; ECA and GeM reduce per cloud: layers/eca_block.py:23,
layers/pooling.py:84-86), so the path shards over clouds with NO collective inside the network.  One process
per GPU; the only exchange is one all-gather of the (B_local, 256) global descriptors (NCCL over NVLink on the
B200 box, gloo in the CPU tests).  Local descriptors / keypoints stay on the rank that produced them.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class Communicator:
    """The engine's own NCCL communicator (``egn_comm_*`` / ``egn_allgather_global`` of the C ABI): one per process,
    on the process's GPU.  ``torch.distributed`` is used once, as the out-of-band channel that carries rank 0's
    128-byte NCCL id to the other ranks; the all-gather itself is issued by the library on the caller's CUDA stream
    (no PyTorch collective, no cross-stream coupling through PyTorch's internal NCCL stream)."""

    def __init__(self, device: torch.device, group=None):
        from . import lib as L
        assert dist.is_initialized(), "init torch.distributed first (it only carries the NCCL id)"
        self.lib = L.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        ident = (C.c_ubyte * 128)()
        if self.rank == 0:
            L.check(self.lib.egn_comm_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0, group=group)
        ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
        self._comm = C.c_void_p()
        L.check(self.lib.egn_comm_create(C.byref(self._comm), self.device.index or 0, self.rank, self.world, ident))
        # NCCL connects its peers lazily, inside the first collective of a communicator (host-side rendezvous of all ranks
        # plus device allocations).  Do that here, where every rank is at the same program point and nothing else runs,
        # not inside a pipeline of worker threads.
        with torch.cuda.device(self.device):
            warm = torch.zeros((1, 32), dtype=torch.float32, device=self.device)
            self.all_gather(warm)
            torch.cuda.current_stream(self.device).synchronize()

    def all_gather(self, send: torch.Tensor, recv: Optional[torch.Tensor] = None) -> torch.Tensor:
        """recv (world * n, D) <- send (n, D) of every rank, rank order, on the current CUDA stream."""
        from . import lib as L
        assert send.is_cuda and send.dtype == torch.float32 and send.is_contiguous()
        if recv is None:
            recv = torch.empty((self.world * send.shape[0],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        L.check(self.lib.egn_allgather_global(self._comm, C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()), send.numel(),
                                              C.c_void_p(torch.cuda.current_stream(send.device).cuda_stream)))
        return recv

    def close(self):
        if getattr(self, "_comm", None) is not None and self._comm.value:
            self.lib.egn_comm_destroy(self._comm)
            self._comm = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_clouds(sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy balance: clouds sorted by voxel (or point) count, each to the currently lightest rank.
    Returns, per rank, the ORIGINAL indices of its clouds in ascending order; deterministic on every rank."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += int(sizes[i])
    return [sorted(p) for p in parts]


def gather_global(local: torch.Tensor, parts: List[List[int]], group=None, comm: Optional[Communicator] = None) -> torch.Tensor:
    """All-gather the per-rank (B_r, D) global descriptors and restore the original cloud order: (B, D).
    Ranks may own different numbers of clouds: rows are padded to the largest share for the collective.
    ``comm``: the engine's NCCL communicator (egn_allgather_global, CUDA tensors); without it the collective goes
    through ``torch.distributed`` (the gloo CPU tests)."""
    world = comm.world if comm is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
    total = sum(len(p) for p in parts)
    if world == 1:
        out = torch.empty((total, local.shape[1]), dtype=local.dtype, device=local.device)
        out[torch.tensor(parts[0], device=local.device, dtype=torch.long)] = local
        return out
    bmax = max(len(p) for p in parts)
    pad = torch.zeros((bmax, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * bmax, local.shape[1]), dtype=local.dtype, device=local.device)
    if comm is not None:
        comm.all_gather(pad, buf)
    else:
        dist.all_gather_into_tensor(buf, pad, group=group)
    out = torch.empty((total, local.shape[1]), dtype=local.dtype, device=local.device)
    for r, p in enumerate(parts):
        if p:
            out[torch.tensor(p, device=local.device, dtype=torch.long)] = buf[r * bmax: r * bmax + len(p)]
    return out


class OrderedGatherer:
    """The collectives of one process, issued by ONE host thread on ONE CUDA stream through ONE communicator, in ticket
    order - whatever the order in which the worker threads finish their steps.

    NCCL kernels wait for their peers on the device, so collectives of different communicators (or of one communicator in
    different orders on different ranks) that run concurrently on a GPU can wait for each other across GPUs for ever; an
    8-GPU run with one communicator per compute stream did exactly that.  The safe shape is the classic one: a single
    communicator, a single stream, the same order on every rank.  Worker threads ``submit(ticket, fn, *tensors)`` after
    enqueueing the producer of ``tensors`` on their own stream; the gather thread runs ``fn()`` (which calls
    ``comm.all_gather`` / ``gather_global``) on the communication stream behind an event, strictly by ascending ticket.
    Every rank must use the same tickets (consecutive integers) - e.g. the global step number."""

    def __init__(self, device: torch.device, first_ticket: int = 0):
        import threading
        self.device = torch.device(device)
        self.stream = self._make_stream()
        self._cv = threading.Condition()
        self._pending: Dict[int, tuple] = {}
        self._results: Dict[int, object] = {}
        self._next = first_ticket
        self._closed = False
        self._error: Optional[BaseException] = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    # CUDA plumbing, overridable (the CPU unit test replaces these three)
    def _make_stream(self):
        return torch.cuda.Stream(device=self.device)

    def _record_event(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return ev

    def _run_on_stream(self, ev, fn, tensors):
        torch.cuda.set_device(self.device)
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            out = fn()
        for t in tensors:
            t.record_stream(self.stream)
        return out

    def submit(self, ticket: int, fn, *tensors):
        """Called by a worker thread right after it has enqueued the producer of ``tensors`` on its current stream.
        ``fn is None`` consumes the ticket without a collective (a step that failed must still release its successors)."""
        ev = self._record_event() if fn is not None else None
        with self._cv:
            assert ticket >= self._next and ticket not in self._pending, f"ticket {ticket} submitted twice or too late"
            self._pending[ticket] = (ev, fn, tensors)
            self._cv.notify_all()

    def result(self, ticket: int):
        """Blocks until the collective of ``ticket`` has been ENQUEUED on the communication stream; returns what ``fn``
        returned (tensors that are ready on ``self.stream``)."""
        with self._cv:
            while ticket not in self._results and self._error is None:
                self._cv.wait()
            if self._error is not None:
                raise self._error
            return self._results.pop(ticket)

    def drain(self, up_to: int):
        """Blocks until every ticket below ``up_to`` has been enqueued (host side; synchronise ``self.stream`` for the device)."""
        with self._cv:
            while self._next < up_to and self._error is None:
                self._cv.wait()
            if self._error is not None:
                raise self._error

    def _run(self):
        while True:
            with self._cv:
                while self._next not in self._pending and not self._closed:
                    self._cv.wait()
                if self._closed and self._next not in self._pending:
                    return
                ticket = self._next
                ev, fn, tensors = self._pending.pop(ticket)
            try:
                out = self._run_on_stream(ev, fn, tensors) if fn is not None else None
            except BaseException as exc:                      # surfaced to whoever waits
                with self._cv:
                    self._error = exc
                    self._cv.notify_all()
                return
            with self._cv:
                self._results[ticket] = out
                self._next = ticket + 1
                self._cv.notify_all()

    def forget_results(self):
        with self._cv:
            self._results.clear()

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._thread.join()


class ShardedBatch:
    """This rank's share of a batch of voxelised clouds: the greedy partition, the batched coordinates of its clouds
    (``ME.utils.batched_coordinates`` of the share) and the all-ones features - everything that does not change when the
    same batch is extracted again."""

    def __init__(self, clouds_coords: List[torch.Tensor], batched_coordinates, rank: int, world: int):
        self.sizes = [int(c.shape[0]) for c in clouds_coords]
        self.parts = shard_clouds(self.sizes, world)
        self.mine = self.parts[rank]
        self.device = clouds_coords[0].device
        self.loads = [sum(self.sizes[i] for i in p) for p in self.parts]          # voxels per rank
        if self.mine:
            self.coords = batched_coordinates([clouds_coords[i] for i in self.mine]).contiguous()
            self.features = torch.ones((self.coords.shape[0], 1), device=self.device)
        else:
            self.coords = self.features = None

    @property
    def imbalance(self) -> float:
        """max / mean voxels per rank (1.0 = perfectly balanced)."""
        return max(self.loads) / (sum(self.loads) / len(self.loads)) if sum(self.loads) else 1.0


def on_side_stream(side: torch.cuda.Stream, fn, *tensors):
    """Run ``fn()`` on ``side`` after everything enqueued so far on the current stream: the caller's stream goes on with
    its next batch instead of waiting for the other ranks inside the collective.  ``tensors`` are inputs of ``fn`` that
    the caching allocator must keep alive until ``side`` has used them."""
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(cur)
    side.wait_event(ev)
    with torch.cuda.stream(side):
        out = fn()
    for t in tensors:
        t.record_stream(side)
    return out


def run_sharded(model, sb: ShardedBatch, group=None, comm: Optional[Communicator] = None,
                comm_stream: Optional[torch.cuda.Stream] = None) -> Tuple[torch.Tensor, Dict]:
    """Forward of this rank's share + the ONE all-gather of the path: (all global descriptors (B,256) in original cloud
    order, this rank's packed local outputs with ``cloud_ids``).  With ``comm_stream`` the collective is issued there
    (the result is then ready on THAT stream): a rank's compute stream never waits for the other ranks."""
    if sb.mine:
        local = model.forward_packed({"coords": sb.coords, "features": sb.features})
        g = local["global"]
    else:
        local, g = {}, torch.zeros((0, model.global_descriptor_size), device=sb.device)
    local["cloud_ids"] = sb.mine
    local["parts"] = sb.parts
    if comm_stream is not None:
        return on_side_stream(comm_stream, lambda: gather_global(g, sb.parts, group, comm), g), local
    return gather_global(g, sb.parts, group, comm), local


def extract_sharded(model, clouds_coords: List[torch.Tensor], batched_coordinates, group=None,
                    comm: Optional[Communicator] = None) -> Tuple[torch.Tensor, Dict]:
    """Run ``model.forward_packed`` on this rank's share of ``clouds_coords`` (list of (Mi,3) int32 voxel coords on
    the rank's device; the single-device original is the loop of eval/evaluate.py:454-466) and return (all global
    descriptors (B,256) in original order, this rank's packed local outputs with ``cloud_ids`` = original indices of
    its clouds)."""
    world = comm.world if comm is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
    rank = comm.rank if comm is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
    return run_sharded(model, ShardedBatch(clouds_coords, batched_coordinates, rank, world), group, comm)
