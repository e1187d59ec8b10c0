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
