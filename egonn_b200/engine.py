"""Python handle on one ``egn_ctx`` (coordinate manager + scratch) of the CUDA engine.

Host-side plumbing only: tensors are torch CUDA tensors, every operator is a call through the C ABI in
``include/egonn_b200.h`` on the current torch stream.  No operator here has a CPU path."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import lib as L


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise L.EgnError(f"{name} must be a CUDA tensor: egonn_b200 has no CPU path (got device {t.device})")


@dataclass
class CoordsInfo:
    n_batches: int
    n_input: int
    n_rows: List[int]


class Engine:
    def __init__(self, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise L.EgnError("no CUDA device: egonn_b200 has no CPU path")
        self.lib = L.load()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self._ctx = C.c_void_p()
        L.check(self.lib.egn_ctx_create(C.byref(self._ctx), dev.index))
        self.info: Optional[CoordsInfo] = None

    def __del__(self):
        try:
            if getattr(self, "_ctx", None) is not None and self._ctx.value:
                self.lib.egn_ctx_destroy(self._ctx)
                self._ctx = C.c_void_p()
        except Exception:
            pass

    # -- quantisation (datasets/quantization.py) -------------------------------------------------------
    def quantize(self, points: torch.Tensor, step, polar: bool):
        """ME.utils.sparse_quantize semantics on the GPU: returns (coords (m,3) int32, index (m,) int64)."""
        _need_cuda(points, "points")
        pts = points.detach().to(torch.float32).contiguous()
        assert pts.dim() == 2 and pts.shape[1] == 3
        n = pts.shape[0]
        coords = torch.empty((n, 3), dtype=torch.int32, device=pts.device)
        index = torch.empty((n,), dtype=torch.int64, device=pts.device)
        if n == 0:
            return coords, index
        st = step if isinstance(step, (list, tuple)) else [step, step, step]
        cstep = (C.c_float * 3)(*[float(v) for v in st])
        n_out = C.c_int64(0)
        with torch.cuda.device(pts.device):
            L.check(self.lib.egn_quantize(self._ctx, _ptr(pts), n, cstep, int(polar), _ptr(coords), _ptr(index),
                                          C.byref(n_out), _stream()))
        return coords[: n_out.value], index[: n_out.value]

    # -- coordinate manager --------------------------------------------------------------------------------
    def build(self, coords: torch.Tensor) -> CoordsInfo:
        _need_cuda(coords, "coords")
        assert coords.dim() == 2 and coords.shape[1] == 4, "coords must be (N,4) [batch,x,y,z]"
        c = coords.detach().to(torch.int32).contiguous()
        if c.shape[0] == 0:
            raise L.EgnError("empty coordinate set")
        info = L.CoordsInfo()
        with torch.cuda.device(c.device):
            L.check(self.lib.egn_coords_build(self._ctx, _ptr(c), c.shape[0], C.byref(info), _stream()))
        self._coords_keepalive = c
        self._derived = {}                                   # index tensors derived from the pyramid (egonn_b200.autograd)
        self.info = CoordsInfo(info.n_batches, info.n_input, list(info.n_rows))
        return self.info

    def build_points(self, points: torch.Tensor, cloud_offsets: torch.Tensor, step, polar: bool) -> CoordsInfo:
        """Fused ingest (egn_coords_build_points): concatenated raw points (n,3) f32 + (B+1) int32 first-point offsets,
        both on the device -> quantised, de-duplicated, batched pyramid in one sort."""
        _need_cuda(points, "points")
        _need_cuda(cloud_offsets, "cloud_offsets")
        pts = points.detach().to(torch.float32).contiguous()
        off = cloud_offsets.detach().to(torch.int32).contiguous()
        assert pts.dim() == 2 and pts.shape[1] == 3 and off.dim() == 1 and off.shape[0] >= 2
        st = step if isinstance(step, (list, tuple)) else [step, step, step]
        cstep = (C.c_float * 3)(*[float(v) for v in st])
        info = L.CoordsInfo()
        with torch.cuda.device(pts.device):
            L.check(self.lib.egn_coords_build_points(self._ctx, _ptr(pts), pts.shape[0], _ptr(off), off.shape[0] - 1, cstep,
                                                     int(polar), C.byref(info), _stream()))
        self._coords_keepalive = (pts, off)
        self._derived = {}
        self.info = CoordsInfo(info.n_batches, info.n_input, list(info.n_rows))
        return self.info

    def _new(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def level_coords(self, level: int) -> torch.Tensor:
        out = self._new((self.info.n_rows[level], 4), torch.int32)
        L.check(self.lib.egn_coords_get(self._ctx, level, _ptr(out), _stream()))
        return out

    def input_rows(self) -> torch.Tensor:
        out = self._new((self.info.n_rows[0],), torch.int32)
        L.check(self.lib.egn_coords_input_rows(self._ctx, _ptr(out), _stream()))
        return out

    def batch_offsets(self, level: int) -> torch.Tensor:
        out = self._new((self.info.n_batches + 1,), torch.int32)
        L.check(self.lib.egn_coords_batch_offsets(self._ctx, level, _ptr(out), _stream()))
        return out

    def neighbors(self, level: int) -> torch.Tensor:
        out = self._new((self.info.n_rows[level], 27), torch.int32)
        L.check(self.lib.egn_coords_neighbors(self._ctx, level, _ptr(out), _stream()))
        return out

    def weights_resident(self, blob: Optional[torch.Tensor]):
        """Pin the weight blob in L2 (persisting access-policy window) for this context's forwards."""
        if blob is None:
            L.check(self.lib.egn_weights_resident(self._ctx, None, 0))
        else:
            L.check(self.lib.egn_weights_resident(self._ctx, _ptr(blob), blob.numel() * blob.element_size()))

    # -- whole forward -----------------------------------------------------------------------------------------
    def forward(self, net: L.Net, blob: torch.Tensor, features: torch.Tensor, want_global=True, want_local=True) -> Dict:
        assert self.info is not None, "build() first"
        f = None
        if features is not None:
            _need_cuda(features, "features")
            f = features.detach().to(torch.float32).contiguous().reshape(-1)
            assert f.shape[0] == self.info.n_input, "features must have one row per input coordinate"
        out: Dict[str, torch.Tensor] = {}
        g = d = k = s = None
        if want_global:
            gdim = net.global_mlp[1].cout if net.global_mlp[0].cin else net.global_head.out_channels
            g = self._new((self.info.n_batches, gdim), torch.float32)
            out["global"] = g
        if want_local:
            lvl = net.local_head.levels[0]
            n = self.info.n_rows[lvl]
            d = self._new((n, net.desc_mlp[1].cout), torch.float32)
            k = self._new((n, 3), torch.float32)
            s = self._new((n, 1), torch.float32)
            out.update(descriptors=d, keypoints=k, sigma=s, local_level=lvl)
        L.check(self.lib.egn_forward(self._ctx, C.byref(net), _ptr(blob), _ptr(f), _ptr(g), _ptr(d), _ptr(k), _ptr(s), _stream()))
        return out

    def tap(self, which: int, level: int, channels: int) -> torch.Tensor:
        lvl = 0 if which == 0 else level
        out = self._new((self.info.n_rows[lvl], channels), torch.float32)
        L.check(self.lib.egn_forward_tap(self._ctx, which, level, _ptr(out), _stream()))
        return out

    # -- measurement hooks ------------------------------------------------------------------------------------
    def profile(self, enable: bool):
        L.check(self.lib.egn_profile_enable(self._ctx, int(enable)))

    def profile_read(self, reset=True) -> List[dict]:
        buf = (L.ProfileEntry * 256)()
        n = C.c_int(0)
        L.check(self.lib.egn_profile_read(self._ctx, buf, 256, C.byref(n), int(reset)))
        return [dict(name=buf[i].name.decode(), launches=buf[i].launches, ms=buf[i].ms, alg_bytes=buf[i].alg_bytes,
                     flops=buf[i].flops) for i in range(n.value)]

    def launch_count(self) -> int:
        return int(self.lib.egn_launch_count(self._ctx))

    # -- single operators (used by egonn_b200.minkowski) -----------------------------------------------------
    def conv(self, level_in: int, ksize: int, transposed: bool, x: torch.Tensor, kernel: torch.Tensor,
             scale=None, shift=None, relu=False, out: Optional[torch.Tensor] = None, accumulate=False) -> torch.Tensor:
        _need_cuda(x, "features")
        k = kernel if kernel.dim() == 3 else kernel.unsqueeze(0)
        cin, cout = k.shape[1], k.shape[2]
        x = x.detach().to(torch.float32).contiguous()
        k = k.detach().to(torch.float32).contiguous()
        if ksize == 2:
            lvl_out = level_in - 1 if transposed else level_in + 1
        else:
            lvl_out = level_in
        if out is None:
            out = self._new((self.info.n_rows[lvl_out], cout), torch.float32)
        L.check(self.lib.egn_conv(self._ctx, level_in, ksize, int(transposed), cin, cout, _ptr(x), _ptr(k), _ptr(scale),
                                  _ptr(shift), int(relu), int(accumulate), _ptr(out), _stream()))
        return out

    def conv_tc(self, level_in: int, ksize: int, x: torch.Tensor, kernel: torch.Tensor, scale=None, shift=None,
                relu=False, transposed=False) -> torch.Tensor:
        """tcgen05 path of one convolution (ksize 3, or 2 with stride 2); packs the kernel on the fly."""
        from .weights import pack_tc
        _need_cuda(x, "features")
        cin, cout = kernel.shape[1], kernel.shape[2]
        x = x.detach().to(torch.float32).contiguous()
        wpack = pack_tc(kernel).to(x.device)
        lvl_out = (level_in - 1 if transposed else level_in + 1) if ksize == 2 else level_in
        out = self._new((self.info.n_rows[lvl_out], cout), torch.float32)
        L.check(self.lib.egn_conv_tc(self._ctx, level_in, ksize, int(transposed), cin, cout, _ptr(x), _ptr(wpack), _ptr(scale), _ptr(shift),
                                     int(relu), _ptr(out), _stream()))
        return out

    def set_tensor_cores(self, enable: bool):
        L.check(self.lib.egn_set_tensor_cores(self._ctx, int(enable)))

    def global_pool(self, level: int, x: torch.Tensor, is_max=False) -> torch.Tensor:
        _need_cuda(x, "features")
        x = x.detach().to(torch.float32).contiguous()
        out = self._new((self.info.n_batches, x.shape[1]), torch.float32)
        L.check(self.lib.egn_global_pool(self._ctx, level, x.shape[1], _ptr(x), int(is_max), _ptr(out), _stream()))
        return out

    def broadcast_mul(self, level: int, x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
        _need_cuda(x, "features")
        x = x.detach().to(torch.float32).contiguous()
        g = g.detach().to(torch.float32).contiguous()
        out = torch.empty_like(x)
        L.check(self.lib.egn_broadcast_mul(self._ctx, level, x.shape[1], _ptr(x), _ptr(g), _ptr(out), _stream()))
        return out


_retrieval_engines = {}


def _stream_key(device):
    """(device index, current CUDA stream, host thread): an ``egn_ctx`` holds single-stream scratch, see quantization._engine."""
    import threading
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return (idx, torch.cuda.current_stream(idx).cuda_stream, threading.get_ident())


def knn_global(query: torch.Tensor, map_embeddings: torch.Tensor, k: int):
    """Nearest neighbours of every query descriptor in the map set by Euclidean distance (eval/evaluate.py:173-176).
    Returns (idx (Q,k) int64 map rows, ascending distance; dist (Q,k) f32).  CUDA tensors only."""
    _need_cuda(query, "query")
    _need_cuda(map_embeddings, "map_embeddings")
    q = query.detach().to(torch.float32).contiguous()
    m = map_embeddings.detach().to(torch.float32).contiguous()
    assert q.dim() == 2 and m.dim() == 2 and q.shape[1] == m.shape[1]
    eng = _shared_engine(q.device)
    idx = torch.empty((q.shape[0], k), dtype=torch.int32, device=q.device)
    dist = torch.empty((q.shape[0], m.shape[0]), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        L.check(eng.lib.egn_knn_l2(eng._ctx, _ptr(q), _ptr(m), q.shape[0], m.shape[0], q.shape[1], k, _ptr(idx), _ptr(dist), _stream()))
    idx = idx.long()
    return idx, torch.gather(dist, 1, idx.clamp_min(0))


def _shared_engine(device) -> "Engine":
    key = _stream_key(device)
    if key not in _retrieval_engines:
        _retrieval_engines[key] = Engine(torch.device("cuda", key[0]))
    return _retrieval_engines[key]


def match_descriptors(desc_a: torch.Tensor, desc_b: torch.Tensor, mutual: bool = True):
    """Correspondences between the local descriptors of two clouds - the feature-matching step of
    eval/evaluate.py:381-399 (Open3D ransac_based_on_feature_matching, mutual_filter=True).  Returns (idx (n_a,) int64 row of
    ``desc_b`` per row of ``desc_a``, -1 where the nearest neighbour is not mutual; dist (n_a,) f32).  CUDA tensors only."""
    _need_cuda(desc_a, "desc_a")
    _need_cuda(desc_b, "desc_b")
    a = desc_a.detach().to(torch.float32).contiguous()
    b = desc_b.detach().to(torch.float32).contiguous()
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1]
    eng = _shared_engine(a.device)
    idx = torch.empty((a.shape[0],), dtype=torch.int32, device=a.device)
    dist = torch.empty((a.shape[0],), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        L.check(eng.lib.egn_match_mutual(eng._ctx, _ptr(a), _ptr(b), a.shape[0], b.shape[0], a.shape[1], int(mutual), _ptr(idx), _ptr(dist),
                                         _stream()))
    return idx.long(), dist


def filter_points(records: torch.Tensor, remove_zero_points: bool = True, remove_ground_plane: bool = True,
                  ground_plane_level: float = -1.5) -> torch.Tensor:
    """PointCloudLoader.__call__ (misc/point_clouds.py:95-111) on the device: (n, >=3) f32 records (x, y, z[, reflectance]) ->
    (m, 3) points without the all-zero points and the points at or below the ground plane, input order kept."""
    _need_cuda(records, "records")
    r = records.detach().to(torch.float32).contiguous()
    assert r.dim() == 2 and r.shape[1] >= 3
    out = torch.empty((r.shape[0], 3), dtype=torch.float32, device=r.device)
    if r.shape[0] == 0:
        return out
    eng = _shared_engine(r.device)
    n_out = C.c_int64(0)
    with torch.cuda.device(r.device):
        L.check(eng.lib.egn_filter_points(eng._ctx, _ptr(r), r.shape[0], r.shape[1], int(remove_zero_points), int(remove_ground_plane),
                                          C.c_float(ground_plane_level), _ptr(out), C.byref(n_out), _stream()))
    return out[: n_out.value]


def topk_smallest(sigma: torch.Tensor, offsets: torch.Tensor, k: int) -> torch.Tensor:
    """Per-cloud indices of the k smallest sigma, ascending (eval/evaluate.py:352-361); -1 padded."""
    _need_cuda(sigma, "sigma")
    s = sigma.detach().to(torch.float32).contiguous().reshape(-1)
    off = offsets.to(torch.int32).contiguous()
    nb = off.shape[0] - 1
    out = torch.empty((nb, k), dtype=torch.int32, device=s.device)
    L.check(L.load().egn_topk_smallest(_ptr(s), _ptr(off), nb, k, _ptr(out), _stream()))
    return out


def pack_topk(idx: torch.Tensor, offsets: torch.Tensor, keypoints: torch.Tensor, descriptors: torch.Tensor,
              global_desc: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The selected keypoints / descriptors of every cloud next to its global descriptor (eval/evaluate.py:339-350), packed
    as (B, G + k*3 + k*D) f32 rows [global | keypoints | descriptors] for one device-to-host copy; zeros for -1 padding."""
    _need_cuda(idx, "idx")
    idx = idx.to(torch.int32).contiguous()
    off = offsets.to(torch.int32).contiguous()
    kp = keypoints.detach().to(torch.float32).contiguous()
    ds = descriptors.detach().to(torch.float32).contiguous()
    nb, k = idx.shape
    D = ds.shape[1]
    G = 0 if global_desc is None else global_desc.shape[1]
    g = None if global_desc is None else global_desc.detach().to(torch.float32).contiguous()
    if out is None:
        out = torch.empty((nb, G + k * 3 + k * D), dtype=torch.float32, device=idx.device)
    L.check(L.load().egn_pack_topk(_ptr(idx), _ptr(off), nb, k, _ptr(kp), _ptr(ds), D, _ptr(g), G, _ptr(out), _stream()))
    return out
