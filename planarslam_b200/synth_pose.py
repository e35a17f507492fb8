"""Seeded synthetic pose-optimisation problems (BASELINE.json configs 3/4): map points, map lines and map planes
of the room-corner scene observed from a known camera pose, with pixel noise and gross outliers, plus a
perturbed initial pose — everything Optimizer::PoseOptimization reads from a Frame (src/Optimizer.cc:593-981)."""
from __future__ import annotations

import numpy as np

from . import synth

# Examples/RGB-D/TUM3.yaml:73-110
PLANE_SETTINGS = dict(angle_info=0.5, dist_info=50.0, par_info=0.1, ver_info=0.1, plane_chi=100.0, vp_chi=50.0)
WORLD_PLANES = [(np.array([0.0, 1.0, 0.0]), 1.2), (np.array([1.0, 0.0, 0.0]), -1.6), (np.array([0.0, 0.0, 1.0]), 3.2)]  # n.x = d


def _rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def make_pose_problem(seed: int, frame: int = 0, n_points: int = 1000, n_lines: int = 40, n_planes: int = 3, n_par: int = 1,
                      n_ver: int = 2, outlier_frac: float = 0.05, rot_pert: float = 0.02, trans_pert: float = 0.03,
                      K=synth.TUM3_K, bf: float = 40.0, width: int = 640, height: int = 480):
    """Returns a dict of numpy arrays laid out like the C ABI's pslam_pose_problem plus 'Tcw0' (float 4x4 initial
    pose) and 'Tcw_true'."""
    rng = np.random.Generator(np.random.Philox(key=int(seed) * 7919 + int(frame)))
    fx, fy, cx, cy = K
    R_wc, t_wc = synth.camera_pose(frame, 64)
    R_cw, t_cw = R_wc.T, -R_wc.T @ t_wc
    Tcw_true = np.eye(4)
    Tcw_true[:3, :3], Tcw_true[:3, 3] = R_cw, t_cw

    def sample_on_planes(n):
        pts = []
        while len(pts) < n:
            u, v = rng.uniform(8, width - 8), rng.uniform(8, height - 8)
            d = R_wc @ np.array([(u - cx) / fx, (v - cy) / fy, 1.0])
            best = np.inf
            for nrm, dd in WORLD_PLANES:
                den = d @ nrm
                if abs(den) > 1e-9:
                    t = (dd - t_wc @ nrm) / den
                    if 0.3 < t < best:
                        best = t
            if np.isfinite(best):
                pts.append(t_wc + best * d)
        return np.array(pts, np.float64).reshape(-1, 3)

    def project(Xw):
        Xc = Xw @ R_cw.T + t_cw
        return np.stack([fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy], 1), Xc[:, 2]

    # ---- points ----
    Xw = sample_on_planes(n_points).astype(np.float32)
    uv, z = project(Xw.astype(np.float64))
    octave = rng.integers(0, 8, n_points)
    sig = 1.2 ** octave
    uv = uv + rng.normal(0, 1, uv.shape) * sig[:, None] * 0.7
    n_out = int(outlier_frac * n_points)
    out_idx = rng.choice(n_points, n_out, replace=False)
    uv[out_idx] = np.stack([rng.uniform(0, width, n_out), rng.uniform(0, height, n_out)], 1)
    ur = uv[:, 0] - bf / z + rng.normal(0, 0.5, n_points)
    ur[rng.random(n_points) < 0.2] = -1.0                       # monocular observations (no depth)
    obs = np.concatenate([uv, ur[:, None]], 1).astype(np.float32)
    inv_sigma2 = (1.0 / (np.float32(1.2) ** octave.astype(np.float32)) ** 2).astype(np.float32)

    # ---- lines: two map endpoints, observed line through the noisy projections ----
    A, B = sample_on_planes(n_lines), sample_on_planes(n_lines)
    line_Xw = np.concatenate([A, B], 1).astype(np.float64)
    pa, _ = project(A)
    pb, _ = project(B)
    pa += rng.normal(0, 0.8, pa.shape)
    pb += rng.normal(0, 0.8, pb.shape)
    n_lout = max(1, int(outlier_frac * n_lines)) if n_lines else 0
    if n_lines:
        pb[rng.choice(n_lines, n_lout, replace=False)] += rng.uniform(40, 120, (n_lout, 2))
    l = np.cross(np.concatenate([pa, np.ones((n_lines, 1))], 1), np.concatenate([pb, np.ones((n_lines, 1))], 1))
    line_obs = (l / np.linalg.norm(l, axis=1, keepdims=True)).astype(np.float64) if n_lines else np.zeros((0, 3))

    # ---- planes: map plane (world, float4 "n.x + d = 0") and the frame's measured plane (camera, float4) ----
    def world_coeff(i):
        nrm, dd = WORLD_PLANES[i % 3]
        return np.array([nrm[0], nrm[1], nrm[2], -dd])

    def to_camera(pw, noise):
        n_c = R_cw @ pw[:3]
        d_c = pw[3] - t_cw @ n_c
        n_c = _rodrigues(rng.normal(0, noise, 3)) @ n_c
        v = np.array([n_c[0], n_c[1], n_c[2], d_c + rng.normal(0, noise)])
        return v / np.linalg.norm(v[:3])

    plane_map = np.array([world_coeff(i) for i in range(n_planes)], np.float32).reshape(-1, 4)
    plane_meas = np.array([to_camera(world_coeff(i), 0.004) for i in range(n_planes)], np.float32).reshape(-1, 4)
    # parallel: the same normal at another offset; vertical: an orthogonal wall
    par_map = np.array([world_coeff(i) + np.array([0, 0, 0, 0.8]) for i in range(n_par)], np.float32).reshape(-1, 4)
    par_meas = np.array([to_camera(world_coeff(i), 0.004) for i in range(n_par)], np.float32).reshape(-1, 4)
    ver_map = np.array([world_coeff(i + 1) for i in range(n_ver)], np.float32).reshape(-1, 4)
    ver_meas = np.array([to_camera(world_coeff(i), 0.004) for i in range(n_ver)], np.float32).reshape(-1, 4)

    # ---- initial pose: the truth perturbed (what the motion model would hand over) ----
    dR = _rodrigues(rng.normal(0, rot_pert, 3))
    T0 = np.eye(4)
    T0[:3, :3] = dR @ R_cw
    T0[:3, 3] = t_cw + rng.normal(0, trans_pert, 3)
    return dict(fx=fx, fy=fy, cx=cx, cy=cy, bf=bf, Xw=np.ascontiguousarray(Xw), obs=np.ascontiguousarray(obs),
                inv_sigma2=np.ascontiguousarray(inv_sigma2), line_Xw=np.ascontiguousarray(line_Xw),
                line_obs=np.ascontiguousarray(line_obs), plane_meas=plane_meas, plane_map=plane_map, par_meas=par_meas,
                par_map=par_map, ver_meas=ver_meas, ver_map=ver_map, Tcw0=T0.astype(np.float32), Tcw_true=Tcw_true,
                **PLANE_SETTINGS)


def pose_error(Ta: np.ndarray, Tb: np.ndarray):
    """(rotation angle in rad, translation distance in m) between two 4x4 poses."""
    Ra, Rb = np.asarray(Ta, np.float64)[:3, :3], np.asarray(Tb, np.float64)[:3, :3]
    # chord form: exact 0 for identical inputs even when they are float-rounded (not perfectly orthonormal) rotations,
    # where the arccos(trace) form is ill-conditioned
    ang = 2.0 * np.arcsin(min(1.0, float(np.linalg.norm(Ra - Rb)) / (2.0 * np.sqrt(2.0))))
    return float(ang), float(np.linalg.norm(np.asarray(Ta, np.float64)[:3, 3] - np.asarray(Tb, np.float64)[:3, 3]))
