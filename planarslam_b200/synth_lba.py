"""Seeded synthetic local-bundle-adjustment problems (BASELINE.json config 4: ICL-NUIM-shaped, 20 key frames, ~5000
stereo point edges over ~1700 points, 100 lines x 2 endpoint edges, 30 plane-type edges over 6 planes, pixel noise
sigma = 1 * scale, 5 % outliers) — everything Optimizer::LocalBundleAdjustment reads from the local map
(src/Optimizer.cc:1971-2358), as the plain arrays of pslam_lba_problem (include/pslam_abi.h)."""
from __future__ import annotations

import numpy as np

from . import synth
from .synth_pose import _rodrigues

LBA_SETTINGS = dict(angle_info=0.5, dist_info=50.0, plane_chi=100.0, vp_chi=50.0)     # Examples/RGB-D/ICL.yaml Plane.*
# six world planes "n.x = d": the room corner of synth.py plus three more walls
WORLD_PLANES6 = [(np.array([0.0, 1.0, 0.0]), 1.2), (np.array([1.0, 0.0, 0.0]), -1.6), (np.array([0.0, 0.0, 1.0]), 3.2),
                 (np.array([0.0, 1.0, 0.0]), -1.4), (np.array([1.0, 0.0, 0.0]), 2.2), (np.array([0.0, 0.0, 1.0]), 4.0)]


def make_lba_problem(seed: int, n_kf: int = 20, n_fixed: int = 1, n_points: int = 1700, n_pt_obs: int = 5000, n_lines: int = 100,
                     n_line_obs: int = 100, n_plane_obs=(24, 4, 2), outlier_frac: float = 0.05, mono_frac: float = 0.1,
                     rot_pert: float = 0.004, trans_pert: float = 0.01, pt_pert: float = 0.02, K=synth.ICL_K, bf: float = 40.0,
                     width: int = 640, height: int = 480, kf_stride: int = 2, line_kf_quirk: bool = False, line_norm3: bool = True,
                     plane_outlier_frac: float = 0.0):
    """Returns a dict of numpy arrays laid out like pslam_lba_problem plus the ground truth ('kf_Tcw_true', 'pt_Xw_true').
    Key frames are the frames 0, kf_stride, 2 kf_stride ... of synth.camera_pose; the first n_fixed are fixed."""
    rng = np.random.Generator(np.random.Philox(key=int(seed) * 104729 + 17))
    fx, fy, cx, cy = K
    Rcw, tcw = [], []
    for i in range(n_kf):
        R_wc, t_wc = synth.camera_pose(i * kf_stride, 64)
        Rcw.append(R_wc.T)
        tcw.append(-R_wc.T @ t_wc)
    Rcw, tcw = np.array(Rcw), np.array(tcw)
    T_true = np.tile(np.eye(4), (n_kf, 1, 1))
    T_true[:, :3, :3], T_true[:, :3, 3] = Rcw, tcw

    def sample_world(n):
        """points on the three room-corner planes seen from a random key frame"""
        pts = []
        while len(pts) < n:
            k = int(rng.integers(0, n_kf))
            R_wc, t_wc = Rcw[k].T, -Rcw[k].T @ tcw[k]
            u, v = rng.uniform(8, width - 8), rng.uniform(8, height - 8)
            d = R_wc @ np.array([(u - cx) / fx, (v - cy) / fy, 1.0])
            best = np.inf
            for nrm, dd in WORLD_PLANES6[:3]:
                den = d @ nrm
                if abs(den) > 1e-9:
                    t = (dd - t_wc @ nrm) / den
                    if 0.3 < t < best:
                        best = t
            if np.isfinite(best):
                pts.append(t_wc + best * d)
        return np.array(pts, np.float64).reshape(-1, 3)

    def project(k, Xw):
        Xc = Xw @ Rcw[k].T + tcw[k]
        z = Xc[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            uv = np.stack([fx * Xc[:, 0] / z + cx, fy * Xc[:, 1] / z + cy], 1)
        vis = (z > 0.2) & (uv[:, 0] > 4) & (uv[:, 0] < width - 4) & (uv[:, 1] > 4) & (uv[:, 1] < height - 4)
        return uv, z, vis

    # ---- points and their observations ----
    Xw_true = sample_world(n_points)
    vis_pairs = []
    for k in range(n_kf):
        _, _, vis = project(k, Xw_true)
        vis_pairs.append(np.stack([np.full(vis.sum(), k), np.nonzero(vis)[0]], 1))
    vis_pairs = np.concatenate(vis_pairs)
    # every point gets at least two observations when it is visible twice, then random extra ones up to n_pt_obs
    order = rng.permutation(len(vis_pairs))
    vis_pairs = vis_pairs[order]
    cnt = np.zeros(n_points, int)
    take = np.zeros(len(vis_pairs), bool)
    for i, (k, p) in enumerate(vis_pairs):
        if cnt[p] < 2:
            take[i] = True
            cnt[p] += 1
    rest = np.nonzero(~take)[0]
    extra = max(0, min(n_pt_obs - int(take.sum()), len(rest)))
    take[rest[:extra]] = True
    sel = vis_pairs[take]
    sel = sel[np.lexsort((sel[:, 0], sel[:, 1]))]                 # grouped by point like the reference's creation order
    obs_kf, obs_pt = sel[:, 0].astype(np.int32), sel[:, 1].astype(np.int32)
    n_obs = len(sel)
    uv = np.zeros((n_obs, 2))
    z = np.zeros(n_obs)
    for k in range(n_kf):
        m = obs_kf == k
        uvk, zk, _ = project(k, Xw_true[obs_pt[m]])
        uv[m], z[m] = uvk, zk
    octave = rng.integers(0, 8, n_obs)
    sig = 1.2 ** octave
    uv = uv + rng.normal(0, 1, uv.shape) * sig[:, None]
    n_out = int(outlier_frac * n_obs)
    out_idx = rng.choice(n_obs, n_out, replace=False)
    uv[out_idx] += rng.uniform(15, 60, (n_out, 2)) * rng.choice([-1.0, 1.0], (n_out, 2))
    ur = uv[:, 0] - bf / z + rng.normal(0, 0.5, n_obs)
    ur[rng.random(n_obs) < mono_frac] = -1.0
    pt_obs_uvr = np.concatenate([uv, ur[:, None]], 1).astype(np.float32)
    inv_sigma2 = (1.0 / (np.float32(1.2) ** octave.astype(np.float32)) ** 2).astype(np.float32)

    # ---- lines ----
    A, B = sample_world(n_lines), sample_world(n_lines)
    line_true = np.concatenate([A, B], 1)
    lo_kf, lo_line, lo_l = [], [], []
    tries = 0
    while len(lo_kf) < n_line_obs and n_lines and tries < 50 * n_line_obs:
        tries += 1
        li, k = int(rng.integers(0, n_lines)), int(rng.integers(0, n_kf))
        pa, _, va = project(k, A[li:li + 1])
        pb, _, vb = project(k, B[li:li + 1])
        if not (va[0] and vb[0]) or any(a == k and b == li for a, b in zip(lo_kf, lo_line)):
            continue
        pa = pa[0] + rng.normal(0, 0.8, 2)
        pb = pb[0] + rng.normal(0, 0.8, 2)
        if rng.random() < outlier_frac:
            pa = pa + rng.uniform(-60, 60, 2)                  # a wrong line: both image endpoints move, so the observed lines
            pb = pb + rng.uniform(40, 120, 2)                  # of one map line no longer share a point
        l = np.cross(np.append(pa, 1.0), np.append(pb, 1.0))
        lo_kf.append(n_kf - 1 if line_kf_quirk else k)
        lo_line.append(li)
        # LineSegment::ExtractLineSegment divides by the 3-norm (src/LSDextractor.cpp:37), which scales the pixel distance
        # down by ~|c|; line_norm3=False gives a true pixel distance so that the chi2 gate on lines can fire in tests
        lo_l.append(l / (np.linalg.norm(l) if line_norm3 else np.linalg.norm(l[:2])))
    o = np.lexsort((np.array(lo_kf, int), np.array(lo_line, int))) if lo_kf else np.zeros(0, int)
    line_obs_kf = np.array(lo_kf, np.int32)[o]
    line_obs_line = np.array(lo_line, np.int32)[o]
    line_obs_l = np.array(lo_l, np.float64).reshape(-1, 3)[o]

    # ---- planes ----
    def world_coeff(i):
        nrm, dd = WORLD_PLANES6[i % 6]
        return np.array([nrm[0], nrm[1], nrm[2], -dd])

    def to_camera(k, pw, noise):
        n_c = Rcw[k] @ pw[:3]
        d_c = pw[3] - tcw[k] @ n_c
        n_c = _rodrigues(rng.normal(0, noise, 3)) @ n_c
        v = np.array([n_c[0], n_c[1], n_c[2], d_c + rng.normal(0, noise)])
        return v / np.linalg.norm(v[:3])

    n_planes = 6
    plane_true = np.array([world_coeff(i) for i in range(n_planes)])
    pobs = [[], [], []]
    for t in range(3):
        used = set()
        while len(pobs[t]) < n_plane_obs[t]:
            pl, k = int(rng.integers(0, n_planes)), int(rng.integers(0, n_kf))
            if (pl, k) in used:
                continue
            used.add((pl, k))
            # [0] the plane itself, [1] a wall orthogonal to it, [2] the parallel wall (same normal, other offset)
            src = pl if t == 0 else ((pl + 1) % 6 if t == 1 else (pl + 3) % 6)
            pobs[t].append((pl, k, to_camera(k, world_coeff(src), 0.25 if rng.random() < plane_outlier_frac else 0.004)))
        pobs[t].sort(key=lambda r: (r[0], r[1]))
    plane_obs_kf = [np.array([r[1] for r in pobs[t]], np.int32) for t in range(3)]
    plane_obs_plane = [np.array([r[0] for r in pobs[t]], np.int32) for t in range(3)]
    plane_obs_meas = [np.array([r[2] for r in pobs[t]], np.float32).reshape(-1, 4) for t in range(3)]

    # ---- initial estimates: truth perturbed ----
    kf_fixed = np.zeros(n_kf, np.uint8)
    kf_fixed[:n_fixed] = 1
    T0 = T_true.copy()
    for k in range(n_fixed, n_kf):
        T0[k, :3, :3] = _rodrigues(rng.normal(0, rot_pert, 3)) @ Rcw[k]
        T0[k, :3, 3] = tcw[k] + rng.normal(0, trans_pert, 3)
    pt0 = (Xw_true + rng.normal(0, pt_pert, Xw_true.shape)).astype(np.float32)
    line0 = line_true + rng.normal(0, pt_pert, line_true.shape)
    plane0 = plane_true.copy()
    for i in range(n_planes):
        plane0[i, :3] = _rodrigues(rng.normal(0, 0.01, 3)) @ plane0[i, :3]
        plane0[i, 3] += rng.normal(0, 0.01)
    kf_K = np.tile(np.array([fx, fy, cx, cy, bf], np.float32), (n_kf, 1))
    c = np.ascontiguousarray
    return dict(kf_Tcw=c(T0.astype(np.float32)), kf_fixed=kf_fixed, kf_K=c(kf_K), pt_Xw=c(pt0), pt_obs_kf=c(obs_kf), pt_obs_pt=c(obs_pt),
                pt_obs_uvr=c(pt_obs_uvr), pt_obs_inv_sigma2=c(inv_sigma2), line_Xw=c(line0), line_obs_kf=c(line_obs_kf),
                line_obs_line=c(line_obs_line), line_obs_l=c(line_obs_l), plane_Xw=c(plane0.astype(np.float32)),
                plane_obs_kf=plane_obs_kf, plane_obs_plane=plane_obs_plane, plane_obs_meas=plane_obs_meas,
                kf_Tcw_true=T_true, pt_Xw_true=Xw_true, line_Xw_true=line_true, plane_Xw_true=plane_true, **LBA_SETTINGS)


def restrict_to_local_planes(p: dict) -> dict:
    """Optimizer::LocalBundleAdjustment only walks LOCAL map planes - those a local key frame holds in mvpMapPlanes (src/Optimizer.cc:1917-1935) - and adds a
    plane's vertical / parallel observations inside that walk (:2250-2350): a vertical / parallel observation of a plane nobody matched as a plane is never
    read.  Drops such observations from a make_lba_problem dict in place (the C ABI expects the caller to pass only what the reference would read)."""
    import numpy as np
    has = np.zeros(len(p["plane_Xw"]), bool)
    has[p["plane_obs_plane"][0]] = True
    for t in (1, 2):
        keep = has[p["plane_obs_plane"][t]]
        p["plane_obs_plane"][t] = np.ascontiguousarray(p["plane_obs_plane"][t][keep])
        p["plane_obs_kf"][t] = np.ascontiguousarray(p["plane_obs_kf"][t][keep])
        p["plane_obs_meas"][t] = np.ascontiguousarray(p["plane_obs_meas"][t][keep])
    return p
