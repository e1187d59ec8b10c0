"""Pipelined descriptor extraction from raw host point clouds - the caller loop of the reference
(``MinkLocGLEvaluator.compute_embedding`` eval/evaluate.py:327-350 inside the per-scan loop :454-466) for a STREAM of
batches: quantise, batch, forward, select the top-k keypoints by sigma, bring global descriptor + keypoints + their
descriptors back to the host.

One step of the engine blocks its host thread once (the row counts of ``egn_coords_build`` must reach the host), so a
single-threaded loop leaves the GPU waiting for the host.  ``Extractor`` feeds ``streams`` CUDA streams from as many host
threads (the C ABI releases the GIL), each with its own engine context, copy stream and two staging slots: per batch ONE
host-to-device copy (points + first-point offsets), the fused ingest ``egn_coords_build_points``, ``egn_forward``,
``egn_topk_smallest``, ``egn_pack_topk`` and ONE device-to-host copy.  Results come back in submission order.
"""
from __future__ import annotations

import queue
import threading
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Union

import numpy as np
import torch

from .engine import pack_topk, topk_smallest

Batch = Union[Sequence[np.ndarray], Sequence[torch.Tensor], "StagedBatch"]


class StagedBatch:
    """A batch of raw clouds already laid out for ONE host-to-device copy: a pinned f32 buffer [points (n,3) | first-point
    offsets (B+1, int32 bits)].  Build it where the clouds are read (e.g. in data-loader workers) with ``stage_batch``."""

    def __init__(self, buf: torch.Tensor, n_points: int, n_clouds: int):
        self.buf, self.n_points, self.n_clouds = buf, n_points, n_clouds
        self.words = n_points * 3 + n_clouds + 1


def stage_batch(clouds: Sequence, pin: bool = True) -> StagedBatch:
    n = [int(c.shape[0]) for c in clouds]
    total, b = int(sum(n)), len(n)
    buf = torch.empty((total * 3 + b + 1,), dtype=torch.float32)
    if pin:
        buf = buf.pin_memory()
    pts = buf[: total * 3].view(total, 3)
    o = 0
    for c, k in zip(clouds, n):
        pts[o:o + k] = torch.as_tensor(c, dtype=torch.float32)
        o += k
    starts = np.zeros(b + 1, dtype=np.int32)
    starts[1:] = np.cumsum(n)
    buf[total * 3:] = torch.from_numpy(starts).view(torch.float32)
    return StagedBatch(buf, total, b)


class _Slot:
    def __init__(self):
        self.host_in: Optional[torch.Tensor] = None      # pinned [points | offsets] staging
        self.dev_in: Optional[torch.Tensor] = None
        self.dev_out: Optional[torch.Tensor] = None
        self.host_out: Optional[torch.Tensor] = None     # pinned packed result
        self.host_off: Optional[torch.Tensor] = None     # pinned per-cloud row offsets of the local level
        self.host_all: Optional[torch.Tensor] = None     # pinned all-gathered global descriptors (multi-GPU)
        self.gather_ticket = None
        self.done = None                                 # event: the slot's last device work has finished


class Extractor:
    """``Extractor(model, streams=4, topk=256).extract(batches)`` yields one dict per batch, in order:
    ``global (B,G)``, ``keypoints (B,k,3)``, ``descriptors (B,k,D)`` (CPU tensors; zeros where a cloud has fewer than k
    keypoints), ``n_keypoints (B,)``.  ``batches``: iterable of lists of per-cloud (n_i, 3) float32 arrays / CPU tensors,
    or of ``StagedBatch`` objects (``stage_batch(clouds)``: already pinned and laid out, no host copy in the pipeline).
    The model must live on a CUDA device (there is no CPU path)."""

    def __init__(self, model, streams: int = 4, topk: int = 256, device: Optional[torch.device] = None, comm=None):
        """``comm``: an ``egonn_b200.parallel.Communicator`` (multi-GPU, one process per GPU): every batch's global
        descriptors are all-gathered over the ranks (``egn_allgather_global``) and returned as ``global_all (world * B, G)``.
        The collectives are issued by one thread on one stream in batch order (``parallel.OrderedGatherer``); every rank
        must submit batches of the same cloud count in the same order, and nothing else may use ``comm`` meanwhile."""
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("egonn_b200 has no CPU path: move the model to a CUDA device first")
        self.model, self.device = model, (torch.device(device) if device is not None else p.device)
        self.S, self.topk = max(1, int(streams)), int(topk)
        self.gdim = model.global_descriptor_size
        self.ddim = model.local_descriptor_size
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.S)]
        self.copy_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.S)]
        self.slots = [[_Slot(), _Slot()] for _ in range(self.S)]
        self.comm = comm
        self.gatherer = None                               # parallel.OrderedGatherer, created with the worker threads
        model._pack(self.device)                           # weights packed before the worker threads start

    # -- staging ------------------------------------------------------------------------------------------
    def _stage(self, slot: _Slot, clouds: Batch):
        """Returns (host staging buffer, points, clouds, words, device buffer freshly allocated); a StagedBatch is used as it
        is (no host copy)."""
        if isinstance(clouds, StagedBatch):
            sb = clouds
        else:
            sb = stage_batch(clouds, pin=False)                       # laid out once ...
            if slot.host_in is None or slot.host_in.numel() < sb.words:
                slot.host_in = torch.empty((int(sb.words * 1.25) + 64,), dtype=torch.float32).pin_memory()
            slot.host_in[: sb.words].copy_(sb.buf)                    # ... and moved into the slot's pinned buffer
            sb = StagedBatch(slot.host_in, sb.n_points, sb.n_clouds)
        fresh = slot.dev_in is None or slot.dev_in.numel() < sb.words
        if fresh:
            slot.dev_in = torch.empty((int(sb.words * 1.25) + 64,), dtype=torch.float32, device=self.device)
        per = self.gdim + self.topk * (3 + self.ddim)
        if slot.host_out is None or slot.host_out.shape[0] < sb.n_clouds:
            slot.host_out = torch.empty((sb.n_clouds, per), dtype=torch.float32).pin_memory()
            slot.dev_out = torch.empty((sb.n_clouds, per), dtype=torch.float32, device=self.device)
        return sb.buf, sb.n_points, sb.n_clouds, sb.words, fresh

    def _launch(self, t: int, slot: _Slot, clouds: Batch, seq: int = 0):
        """Enqueue one batch on stream t (no host wait): H2D on the thread's copy stream, ingest + forward + top-k + pack on
        its compute stream, D2H of the packed rows and of the per-cloud row offsets."""
        cur, cs = self.streams[t], self.copy_streams[t]
        if slot.done is not None:
            slot.done.synchronize()                        # the slot's previous batch has left the device
        host_in, total, b, words, fresh = self._stage(slot, clouds)
        if fresh:
            # the caching allocator may have recycled a block that kernels still queued on the COMPUTE stream write (outputs of
            # the previous batch, freed on the host already): the copy stream must not write it before they have run
            cs.wait_stream(cur)
        with torch.cuda.stream(cs):
            slot.dev_in[:words].copy_(host_in[:words], non_blocking=True)          # ONE host-to-device copy
            up = torch.cuda.Event()
            up.record(cs)
        cur.wait_event(up)
        pts = slot.dev_in[: total * 3].view(total, 3)
        off = slot.dev_in[total * 3: words].view(torch.int32)
        p = self.model.forward_points(pts, off)            # blocks once: the row counts of the pyramid reach the host
        idx = topk_smallest(p["sigma"], p["local_offsets"], self.topk)
        pack_topk(idx, p["local_offsets"], p["keypoints"], p["descriptors"], p["global"], out=slot.dev_out[:b])
        slot.host_out[:b].copy_(slot.dev_out[:b], non_blocking=True)               # ONE device-to-host copy (+ B+1 offsets)
        if slot.host_off is None or slot.host_off.numel() < b + 1:
            slot.host_off = torch.empty((b + 1,), dtype=torch.int32).pin_memory()
        slot.host_off[: b + 1].copy_(p["local_offsets"], non_blocking=True)
        slot.done = torch.cuda.Event()
        slot.done.record(cur)
        slot.gather_ticket = None
        if self.comm is not None:                          # the ONE collective of the path: by the gather thread, in batch order
            comm, g = self.comm, p["global"]
            if slot.host_all is None or slot.host_all.shape[0] < comm.world * b:
                slot.host_all = torch.empty((comm.world * b, self.gdim), dtype=torch.float32).pin_memory()

            def gather():
                allg = comm.all_gather(g)
                slot.host_all[: comm.world * b].copy_(allg, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                return ev

            self.gatherer.submit(seq, gather, g)
            slot.gather_ticket = seq
        return slot, b

    def _finish(self, handle) -> Dict[str, torch.Tensor]:
        slot, b = handle
        slot.done.synchronize()
        out = slot.host_out[:b]
        k, g, d = self.topk, self.gdim, self.ddim
        off = slot.host_off[: b + 1]
        extra = {}
        if slot.gather_ticket is not None:
            self.gatherer.result(slot.gather_ticket).synchronize()
            extra["global_all"] = slot.host_all[: self.comm.world * b].clone()
        return {**extra, "global": out[:, :g].clone(), "keypoints": out[:, g:g + 3 * k].reshape(b, k, 3).clone(),
                "descriptors": out[:, g + 3 * k:].reshape(b, k, d).clone(), "n_keypoints": (off[1:] - off[:-1]).clamp(max=k).clone()}

    # -- the pipeline ---------------------------------------------------------------------------------------
    def _ensure_workers(self):
        """Worker threads live as long as the extractor (started at the first extract(), stopped by close())."""
        if getattr(self, "_threads", None):
            return
        self._works = [queue.Queue() for _ in range(self.S)]   # batch i goes to stream i % S
        if self.comm is not None:
            from .parallel import OrderedGatherer
            self.gatherer = OrderedGatherer(self.device, first_ticket=getattr(self, "_seq_base", 0))
        self._done: Dict[int, object] = {}
        self._cond = threading.Condition()

        def deliver(seq, res):
            with self._cond:
                self._done[seq] = res
                self._cond.notify_all()

        def worker(t: int):
            """Two batches in flight per thread: batch j+1 is enqueued (its points travel, its kernels queue) before the
            thread waits for batch j's results."""
            torch.cuda.set_device(self.device)
            work = self._works[t]
            j, pending = 0, None                           # pending = (seq, handle) enqueued but not yet delivered
            with torch.cuda.stream(self.streams[t]):
                while True:
                    if pending is None:
                        item = work.get()
                    else:
                        try:
                            item = work.get_nowait()
                        except queue.Empty:                # nothing new to enqueue: deliver what is in flight first
                            seq, h = pending
                            pending = None
                            try:
                                deliver(seq, self._finish(h))
                            except BaseException as exc:
                                deliver(seq, exc)
                            continue
                    new = None
                    if item is not None:
                        seq, clouds = item
                        try:
                            new = (seq, self._launch(t, self.slots[t][j & 1], clouds, seq))
                        except BaseException as exc:       # delivered to the consumer in order
                            if self.gatherer is not None and self.slots[t][j & 1].gather_ticket != seq:
                                self.gatherer.submit(seq, None)   # a failed batch still releases the collectives after it
                            deliver(seq, exc)
                        j += 1
                    if pending is not None:
                        pseq, h = pending
                        try:
                            deliver(pseq, self._finish(h))
                        except BaseException as exc:
                            deliver(pseq, exc)
                    if item is None:
                        return
                    pending = new

        self._threads = [threading.Thread(target=worker, args=(t,), daemon=True) for t in range(self.S)]
        for th in self._threads:
            th.start()

    def close(self):
        if getattr(self, "_threads", None):
            for w in self._works:
                w.put(None)
            for th in self._threads:
                th.join()
            self._threads = None
            if self.gatherer is not None:
                self.gatherer.close()
                self.gatherer = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def extract(self, batches: Iterable[Batch]) -> Iterator[Dict[str, torch.Tensor]]:
        """Yields results in submission order; at most 2 * streams batches are in flight.  One extract() at a time."""
        self._ensure_workers()
        with self._cond:
            self._done.clear()
        base = getattr(self, "_seq_base", 0)               # sequence numbers keep growing: batch i -> stream i % S across calls
        submitted = delivered = base
        try:
            it = iter(batches)
            exhausted = False
            while not exhausted or delivered < submitted:
                while not exhausted and submitted - delivered < 2 * self.S:
                    try:
                        clouds = next(it)
                    except StopIteration:
                        exhausted = True
                        break
                    self._works[submitted % self.S].put((submitted, clouds))
                    submitted += 1
                if delivered < submitted:
                    with self._cond:
                        while delivered not in self._done:
                            self._cond.wait()
                        res = self._done.pop(delivered)
                    delivered += 1
                    if isinstance(res, BaseException):
                        raise res
                    yield res
        finally:
            # batches still in flight (the consumer stopped early, or an error was raised) are drained and dropped
            while delivered < submitted:
                with self._cond:
                    while delivered not in self._done:
                        self._cond.wait()
                    self._done.pop(delivered)
                delivered += 1
            self._seq_base = submitted
