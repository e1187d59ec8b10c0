"""TEST INFRASTRUCTURE (see oracle/__init__.py): CPU restatements of the two reference steps either side of the forward that
the engine also runs on the device ("next" rows f2 / f3 of SURVEY.md 8f).  numpy only.

  filter_points   - PointCloudLoader.__call__ after read_pc, /root/reference/misc/point_clouds.py:95-111 with the
                    loaders' records (datasets/kitti/kitti_raw.py:16-22, datasets/mulran/mulran_raw.py:19-25)
  match_mutual    - the correspondence step inside get_ransac_result, /root/reference/eval/evaluate.py:381-399: Open3D
                    registration_ransac_based_on_feature_matching(..., mutual_filter=True) matches every source
                    feature to its nearest target feature (KD-tree, Euclidean) and keeps the mutual pairs.  Open3D is
                    not installed here; this is its published behaviour restated by brute force ("parity unpinned").
"""
import numpy as np


def filter_points(records: np.ndarray, remove_zero_points=True, remove_ground_plane=True, ground_plane_level=-1.5) -> np.ndarray:
    pc = np.asarray(records, dtype=np.float32)[:, :3]                  # read_pc: reshape(-1, 4)[:, :3]
    if remove_zero_points:                                             # misc/point_clouds.py:103-105
        mask = np.all(np.isclose(pc, 0), axis=1)
        pc = pc[~mask]
    if remove_ground_plane:                                            # :107-109
        mask = pc[:, 2] > ground_plane_level
        pc = pc[mask]
    return pc


def match_mutual(feat_a: np.ndarray, feat_b: np.ndarray, mutual=True):
    """(idx (n_a,) int64, dist (n_a,) f32): nearest row of feat_b per row of feat_a (ties: lower row), -1 where not mutual."""
    a = np.asarray(feat_a, dtype=np.float32)
    b = np.asarray(feat_b, dtype=np.float32)
    d2 = ((a[:, None, :].astype(np.float64) - b[None, :, :].astype(np.float64)) ** 2).sum(-1)
    ab = d2.argmin(axis=1)
    dist = np.sqrt(d2[np.arange(a.shape[0]), ab]).astype(np.float32)
    if mutual:
        ba = d2.argmin(axis=0)
        ab = np.where(ba[ab] == np.arange(a.shape[0]), ab, -1)
    return ab.astype(np.int64), dist
