"""Seeded synthetic LiDAR-shaped point clouds for the BASELINE.json configs.

There is no network and no dataset in the image, so every workload is generated on the host from a
numpy ``default_rng(seed)`` stream (bit-reproducible across machines).  The shapes follow the three
sensors the reference reads (datasets/kitti/kitti_raw.py, datasets/mulran/mulran_raw.py,
datasets/southbay/southbay_raw.py): a spinning multi-beam scanner ray-cast against a crude street
scene (ground plane, vertical cylinders, two street walls).
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (generator kwargs, batch, voxel size, description)
    "cfg1": dict(kind="uniform", n=4096, batch=1, voxel=0.3,
                 desc="single synthetic 4096-pt cloud, 0.3 m voxel"),
    "cfg2": dict(kind="spin", beams=64, azimuths=2083, elev=(-24.8, 2.0), height=1.73, max_range=80.0,
                 batch=16, voxel=0.10, desc="batch=16 synthetic KITTI-64-beam clouds (~120k pts), 0.10 m voxel"),
    "cfg3": dict(kind="spin", beams=64, azimuths=1024, elev=(-22.5, 22.5), height=1.9, max_range=100.0,
                 batch=64, voxel=0.15, desc="batch=64 MulRan-Ouster-shape clouds (~65k pts), 0.15 m voxel"),
    "cfg4": dict(kind="spin", beams=64, azimuths=1800, elev=(-24.8, 2.0), height=1.8, max_range=100.0,
                 batch=256, voxel=0.30, desc="batch=256 Apollo-SouthBay-shape clouds, 0.30 m voxel"),
    "cfg5": dict(kind="tile", n=1_000_000, batch=1, voxel=0.05,
                 desc="single dense map-tile cloud 1M pts, 0.05 m voxel"),
}


def uniform_cloud(n: int, seed: int) -> np.ndarray:
    """cfg1: n points uniform in [-40,40]^2 x [-2,6] m."""
    rng = np.random.default_rng(seed)
    pc = rng.random((n, 3), dtype=np.float32)
    pc[:, 0] = pc[:, 0] * 80.0 - 40.0
    pc[:, 1] = pc[:, 1] * 80.0 - 40.0
    pc[:, 2] = pc[:, 2] * 8.0 - 2.0
    return pc


def spinning_lidar_cloud(seed: int, beams: int = 64, azimuths: int = 2083, elev=(-24.8, 2.0),
                         height: float = 1.73, max_range: float = 80.0, n_cylinders: int = 120,
                         noise: float = 0.02, pose=None, noise_seed=None) -> np.ndarray:
    """One revolution of a ``beams`` x ``azimuths`` scanner at ``height`` m above a ground plane, ray-cast
    against ~n_cylinders random vertical cylinders (trunks, poles, cars as fat cylinders) and two
    street walls; Gaussian range noise; misses dropped.  Returns (N,3) float32 in the sensor frame.

    ``pose = (px, py, yaw)`` re-scans the SAME scene (the one ``seed`` draws) from a sensor displaced by (px, py) m
    and rotated by yaw rad - a "revisit"; ``noise_seed`` then draws independent range noise.  The defaults leave the
    generated clouds bit-identical to the pose-free generator the golden vectors and bench workloads were made with."""
    rng = np.random.default_rng(seed)
    el = np.deg2rad(np.linspace(elev[0], elev[1], beams))
    az = np.linspace(-np.pi, np.pi, azimuths, endpoint=False) + rng.uniform(0, 2 * np.pi / azimuths)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    sdx = (ce * np.cos(az)[None, :]).ravel()          # ray directions in the SENSOR frame (what the output uses)
    sdy = (ce * np.sin(az)[None, :]).ravel()
    dz = np.broadcast_to(se, (beams, azimuths)).ravel()
    px, py, yaw = (0.0, 0.0, 0.0) if pose is None else pose
    if pose is None:
        dx, dy = sdx, sdy
    else:                                             # ray directions in the scene frame
        dx = (ce * np.cos(az + yaw)[None, :]).ravel()
        dy = (ce * np.sin(az + yaw)[None, :]).ravel()
    t = np.full(dx.shape, np.inf)
    # ground plane z = -height
    down = dz < -1e-6
    t[down] = np.minimum(t[down], -height / dz[down])
    # street walls: y = +-w, from the ground to wall_h
    w = rng.uniform(12.0, 30.0, size=2)
    wall_h = rng.uniform(4.0, 12.0, size=2)
    for sign, wi, hi in ((1.0, w[0], wall_h[0]), (-1.0, w[1], wall_h[1])):
        with np.errstate(divide="ignore", invalid="ignore"):
            tw = (sign * wi - py) / dy
        z = tw * dz
        ok = (tw > 0) & (z > -height) & (z < hi - height)
        t = np.where(ok & (tw < t), tw, t)
    # vertical cylinders (cx, cy, r, top)
    r_c = rng.uniform(3.0, max_range * 0.9, size=n_cylinders)
    a_c = rng.uniform(-np.pi, np.pi, size=n_cylinders)
    cx, cy = r_c * np.cos(a_c), np.clip(r_c * np.sin(a_c), -min(w) + 0.5, min(w) - 0.5)
    rad = np.where(rng.random(n_cylinders) < 0.3, rng.uniform(0.8, 1.6, n_cylinders), rng.uniform(0.1, 0.4, n_cylinders))
    top = rng.uniform(1.2, 8.0, size=n_cylinders)
    a2 = dx * dx + dy * dy
    cx, cy = cx - px, cy - py
    for j in range(n_cylinders):
        b = dx * cx[j] + dy * cy[j]
        cc = cx[j] ** 2 + cy[j] ** 2 - rad[j] ** 2
        disc = b * b - a2 * cc
        ok = disc > 0
        tc = np.where(ok, (b - np.sqrt(np.where(ok, disc, 0.0))) / np.maximum(a2, 1e-12), np.inf)
        z = tc * dz
        ok &= (tc > 0) & (z < top[j] - height) & (z > -height)
        t = np.where(ok & (tc < t), tc, t)
    hit = np.isfinite(t) & (t < max_range) & (t > 0.5)
    if noise_seed is not None:
        rng = np.random.default_rng(noise_seed)
    t = t[hit] + rng.normal(0.0, noise, size=int(hit.sum()))
    pc = np.stack([sdx[hit] * t, sdy[hit] * t, dz[hit] * t], axis=1)
    return pc.astype(np.float32)


def map_tile_cloud(n: int, seed: int, extent: float = 100.0, n_walls: int = 40) -> np.ndarray:
    """cfg5: half of the points on an extent x extent ground sheet (sigma_z 3 cm), half on random
    vertical wall patches."""
    rng = np.random.default_rng(seed)
    n_ground = n // 2
    g = np.empty((n_ground, 3))
    g[:, 0] = rng.uniform(-extent / 2, extent / 2, n_ground)
    g[:, 1] = rng.uniform(-extent / 2, extent / 2, n_ground)
    g[:, 2] = rng.normal(0.0, 0.03, n_ground)
    per = (n - n_ground) // n_walls
    walls = []
    for j in range(n_walls):
        m = per if j < n_walls - 1 else n - n_ground - per * (n_walls - 1)
        c = rng.uniform(-extent / 2 + 5, extent / 2 - 5, 2)
        ang = rng.uniform(0, np.pi)
        length, hgt = rng.uniform(5.0, 20.0), rng.uniform(2.5, 10.0)
        u = rng.uniform(-length / 2, length / 2, m)
        p = np.empty((m, 3))
        off = rng.normal(0.0, 0.02, m)
        p[:, 0] = c[0] + u * np.cos(ang) - off * np.sin(ang)
        p[:, 1] = c[1] + u * np.sin(ang) + off * np.cos(ang)
        p[:, 2] = rng.uniform(0.0, hgt, m)
        walls.append(p)
    pc = np.concatenate([g] + walls, axis=0)
    return pc[rng.permutation(pc.shape[0])].astype(np.float32)


def make_cloud(cfg: str, seed: int) -> np.ndarray:
    c = CONFIGS[cfg]
    if c["kind"] == "uniform":
        return uniform_cloud(c["n"], seed)
    if c["kind"] == "tile":
        return map_tile_cloud(c["n"], seed)
    return spinning_lidar_cloud(seed, beams=c["beams"], azimuths=c["azimuths"], elev=c["elev"],
                                height=c["height"], max_range=c["max_range"])


def make_batch(cfg: str, batch: int | None = None, first_seed: int | None = None):
    """List of ``batch`` clouds for a BASELINE config (seeds: cfg1 -> 0, others 1..B as SURVEY §8d)."""
    c = CONFIGS[cfg]
    b = c["batch"] if batch is None else batch
    s0 = (0 if cfg == "cfg1" else 1) if first_seed is None else first_seed
    return [make_cloud(cfg, s0 + i) for i in range(b)]
