"""Synthetic "map snapshot" for the projection-guided searches: map points are created from
the keypoints of one frame of the synthetic sequence (back-projected with the rendered depth), exactly the data
ORBmatcher::SearchByProjection reads from MapPoint / Frame objects, as plain arrays."""
from __future__ import annotations

import numpy as np

from . import synth

SCALE_FACTORS = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2))])).astype(np.float32)


def frame_arrays(kps, desc, depth_u16, K=synth.TUM3_K, bf=40.0, depth_factor=5000.0):
    """mvKeysUn (no distortion), mvuRight, mvDepth of a Frame (src/Frame.cc:603-621)."""
    fx, fy, cx, cy = K
    n = len(kps)
    d = (depth_u16.astype(np.float32) * np.float32(1.0 / depth_factor))[kps["y"].astype(np.int64), kps["x"].astype(np.int64)]
    ur = np.where(d > 0, kps["x"] - np.float32(bf) / np.where(d > 0, d, 1), np.float32(-1)).astype(np.float32)
    return dict(n=n, keys_un=np.ascontiguousarray(kps), u_right=np.ascontiguousarray(ur), depth=d, desc=np.ascontiguousarray(desc))


def map_from_frame(fa, Tcw, K=synth.TUM3_K, n_levels=8, rng=None, noise=0.0):
    """Map points from every keypoint with depth: world position, mean viewing normal, scale-invariance distances,
    descriptor (MapPoint::UpdateNormalAndDepth, src/MapPoint.cc:340-386)."""
    fx, fy, cx, cy = K
    ok = fa["depth"] > 0
    idx = np.nonzero(ok)[0]
    k = fa["keys_un"][idx]
    z = fa["depth"][idx].astype(np.float64)
    Xc = np.stack([(k["x"] - cx) * z / fx, (k["y"] - cy) * z / fy, z], 1)
    R, t = np.asarray(Tcw, np.float64)[:3, :3], np.asarray(Tcw, np.float64)[:3, 3]
    Xw = (Xc - t) @ R                                  # R^T (Xc - t)
    if rng is not None and noise > 0:
        Xw = Xw + rng.normal(0, noise, Xw.shape)
    Ow = -R.T @ t
    PC = Xw - Ow
    dist = np.linalg.norm(PC, axis=1)
    normal = (PC / dist[:, None]).astype(np.float32)
    lvl = k["octave"]
    maxd = (dist.astype(np.float32) * SCALE_FACTORS[lvl]).astype(np.float32)
    mind = (maxd / SCALE_FACTORS[n_levels - 1]).astype(np.float32)
    n = len(idx)
    return dict(n=n, pos=np.ascontiguousarray(Xw, np.float32), normal=np.ascontiguousarray(normal), max_distance=maxd, min_distance=mind,
                desc=np.ascontiguousarray(fa["desc"][idx]), skip=np.zeros(n, np.uint8), has_obs=np.ones(n, np.uint8), src_index=idx)


def frame_view(fa, Tcw, K=synth.TUM3_K, bf=40.0, width=640, height=480):
    return dict(n=fa["n"], keys_un=fa["keys_un"], u_right=fa["u_right"], desc=fa["desc"], Tcw=np.ascontiguousarray(Tcw, np.float32),
                fx=K[0], fy=K[1], cx=K[2], cy=K[3], bf=bf, min_x=0.0, max_x=float(width), min_y=0.0, max_y=float(height), n_levels=8,
                scale_factors=SCALE_FACTORS, log_scale_factor=float(np.log(np.float32(1.2), dtype=np.float32)))


def true_pose(frame, n_frames=64):
    R_wc, t_wc = synth.camera_pose(frame, n_frames)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R_wc.T, -R_wc.T @ t_wc
    return T
