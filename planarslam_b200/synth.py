"""Seeded synthetic RGB-D input for the parity tests and bench.py.

The reference ships no data, and there is no network, so every input is generated here:
a textured three-plane "room corner" (floor, left wall, back wall) seen from a smoothly moving pinhole
camera, rendered by exact ray/plane intersection.  Intensity comes from a multi-scale hashed block texture
in plane coordinates (plenty of FAST corners at several scales); depth is z*5000 as uint16 with a
Kinect-like noise model and a fraction of zero "holes" (TUM convention, DepthMapFactor 5000,
Examples/RGB-D/TUM3.yaml:35).

Pure numpy (no cv2) so the same bytes come out in the authoring container and on the GPU box.
"""
from __future__ import annotations

import numpy as np

TUM3_K = (535.4, 539.2, 320.1, 247.6)     # fx, fy, cx, cy  (Examples/RGB-D/TUM3.yaml:8-11)
ICL_K = (481.2, -480.0, 319.5, 239.5)     # Examples/RGB-D/ICL.yaml:8-11
DEPTH_FACTOR = 5000.0


def _hash2(ix: np.ndarray, iy: np.ndarray, salt: int) -> np.ndarray:
    """32-bit integer hash of a lattice cell -> uint8-ish value in [0, 255]."""
    h = (ix.astype(np.int64) * 73856093) ^ (iy.astype(np.int64) * 19349663) ^ (salt * 83492791)
    h &= 0xFFFFFFFF
    h = (h * 2654435761) & 0xFFFFFFFF
    h ^= h >> 15
    h = (h * 2246822519) & 0xFFFFFFFF
    h ^= h >> 13
    return ((h >> 8) & 0xFF).astype(np.float32)


def _texture(u: np.ndarray, v: np.ndarray, salt: int) -> np.ndarray:
    """Multi-scale block texture in metres; returns float32 in about [0, 255]."""
    out = np.zeros(u.shape, np.float32)
    for k, (s, wgt) in enumerate(((0.60, 0.30), (0.21, 0.30), (0.08, 0.25), (0.03, 0.15))):
        out += wgt * _hash2(np.floor(u / s), np.floor(v / s), salt * 7 + k)
    return out


def _blur3(img: np.ndarray) -> np.ndarray:
    """Separable [1 2 1]/4 blur with edge replication (camera-like softening)."""
    p = np.pad(img, 1, mode="edge")
    h = (p[:, :-2] + 2.0 * p[:, 1:-1] + p[:, 2:]) * 0.25
    return (h[:-2] + 2.0 * h[1:-1] + h[2:]) * 0.25


def camera_pose(frame: int, n_frames: int = 64) -> tuple[np.ndarray, np.ndarray]:
    """World-from-camera rotation R_wc (3x3) and camera centre t_wc for a smooth path."""
    s = frame / max(n_frames, 1)
    yaw = 0.25 * np.sin(2 * np.pi * s) + 0.35
    pitch = -(0.10 * np.cos(2 * np.pi * s) + 0.30)
    roll = 0.05 * np.sin(4 * np.pi * s)
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Ry = np.array([[cy, 0, -sy], [0, 1, 0], [sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    Rz = np.array([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]])
    R = Ry @ Rx @ Rz
    t = np.array([0.3 * np.sin(2 * np.pi * s), -0.1 + 0.05 * np.cos(2 * np.pi * s), 0.2 * s])
    return R, t


def render_frame(seed: int, frame: int = 0, width: int = 640, height: int = 480,
                 K: tuple[float, float, float, float] | None = None, n_frames: int = 64,
                 depth_noise: bool = True, hole_frac: float = 0.02):
    """Return (gray uint8 HxW, depth uint16 HxW, z float32 HxW metres (noise-free), (R_wc, t_wc))."""
    if K is None:
        sx = width / 640.0
        K = tuple(k * sx for k in TUM3_K)
    fx, fy, cx, cy = K
    rng = np.random.Generator(np.random.Philox(key=int(seed) * 1000 + int(frame)))
    R, t = camera_pose(frame, n_frames)
    uu, vv = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    dc = np.stack([(uu - cx) / fx, (vv - cy) / fy, np.ones_like(uu)], -1)      # camera rays, z=1
    dw = dc @ R.T
    # planes n.x = d in world coordinates (camera looks along +z, y down)
    planes = [
        (np.array([0.0, 1.0, 0.0]), 1.2, 1),    # floor  y = 1.2
        (np.array([1.0, 0.0, 0.0]), -1.6, 2),   # left wall x = -1.6
        (np.array([0.0, 0.0, 1.0]), 3.2, 3),    # back wall z = 3.2
    ]
    best_t = np.full((height, width), np.inf)
    gray = np.zeros((height, width), np.float32)
    for n, d, salt in planes:
        denom = dw @ n
        with np.errstate(divide="ignore", invalid="ignore"):
            tt = (d - t @ n) / denom
        hit = (tt > 0.05) & (tt < best_t) & np.isfinite(tt)
        P = t + dw * tt[..., None]
        if salt == 1:
            a, b = P[..., 0], P[..., 2]
        elif salt == 2:
            a, b = P[..., 2], P[..., 1]
        else:
            a, b = P[..., 0], P[..., 1]
        tex = _texture(a, b, salt)
        gray = np.where(hit, tex, gray)
        best_t = np.where(hit, tt, best_t)
    z = np.where(np.isfinite(best_t), best_t, 0.0)                # camera-frame depth == t because rays have z=1
    gray = _blur3(gray)
    gray = gray + rng.normal(0.0, 1.5, gray.shape).astype(np.float32)
    gray8 = np.clip(np.rint(gray), 0, 255).astype(np.uint8)
    zn = z.copy()
    if depth_noise:
        sigma = 0.0012 + 0.0019 * (z - 0.4) ** 2
        zn = z + rng.normal(0.0, 1.0, z.shape) * sigma
    d16 = np.clip(np.rint(zn * DEPTH_FACTOR), 0, 65535).astype(np.uint16)
    if hole_frac > 0:
        # Kinect-like dropouts are clustered (specular spots, occlusion shadows), not i.i.d. pixels: paint random
        # discs until about hole_frac of the image is invalid.  (PEAC's INIT_STRICT rejects any 10x10 block that
        # contains a single zero, so i.i.d. holes at 2 % would wipe out 87 % of the blocks.)
        yy, xx = np.mgrid[0:height, 0:width]
        target = hole_frac * width * height
        covered = 0.0
        while covered < target:
            hx, hy = rng.uniform(0, width), rng.uniform(0, height)
            hr = rng.uniform(2.0, 9.0) * width / 640.0
            d16[(xx - hx) ** 2 + (yy - hy) ** 2 < hr * hr] = 0
            covered += np.pi * hr * hr
    d16[z <= 0] = 0
    return gray8, d16, z.astype(np.float32), (R, t)


def render_sequence(seed: int, n: int, width: int = 640, height: int = 480, **kw):
    """Stack of n frames: (gray [n,H,W] uint8, depth [n,H,W] uint16)."""
    g = np.empty((n, height, width), np.uint8)
    d = np.empty((n, height, width), np.uint16)
    for i in range(n):
        g[i], d[i], _, _ = render_frame(seed, i, width, height, n_frames=max(n, 64), **kw)
    return g, d


def _render_job(job):
    seed, i, n, width, height = job
    g, d, _, _ = render_frame(seed, i, width, height, n_frames=max(n, 64))
    return i, g, d


def render_sequence_parallel(seed: int, n: int, width: int = 640, height: int = 480, workers: int = 0):
    """render_sequence over a pool of spawned worker processes (the frames are independent; ~0.4 s of numpy each).  Same bytes as render_sequence."""
    import multiprocessing as mp
    import os
    if workers <= 0:
        workers = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = max(1, min(workers, n))
    if workers == 1:
        return render_sequence(seed, n, width, height)
    g = np.empty((n, height, width), np.uint8)
    d = np.empty((n, height, width), np.uint16)
    with mp.get_context("spawn").Pool(workers) as pool:
        for i, gi, di in pool.imap_unordered(_render_job, [(seed, i, n, width, height) for i in range(n)], chunksize=max(1, n // (4 * workers))):
            g[i], d[i] = gi, di
    return g, d


def polygon_image(seed: int, width: int = 640, height: int = 480, n_poly: int = 40) -> np.ndarray:
    """Config-1 style image: random filled convex-ish polygons + value noise + blur (numpy only)."""
    rng = np.random.Generator(np.random.Philox(key=int(seed)))
    img = np.full((height, width), 96.0, np.float32)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    for _ in range(n_poly):
        c = rng.uniform([0, 0], [width, height])
        r = rng.uniform(20, 140)
        k = int(rng.integers(3, 7))
        ang0 = rng.uniform(0, 2 * np.pi)
        inside = np.ones((height, width), bool)
        for e in range(k):                      # intersection of k half-planes = convex polygon
            a = ang0 + 2 * np.pi * e / k
            nx, ny = np.cos(a), np.sin(a)
            inside &= (xx - c[0]) * nx + (yy - c[1]) * ny < r * rng.uniform(0.5, 1.0)
        img[inside] = rng.uniform(0, 255)
    for octv, amp in ((64, 12.0), (16, 8.0), (4, 4.0)):
        gx, gy = np.floor(xx / octv), np.floor(yy / octv)
        img += amp * (_hash2(gx, gy, 91 + octv) / 255.0 - 0.5) * 2
    img = _blur3(img)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def piecewise_planar_depth(seed: int, width: int = 640, height: int = 480, n_rect: int = 14, noise_mm: float = 3.0, hole_frac: float = 0.01,
                           curved: bool = True):
    """uint16 depth (1/5000 m units) of a wall with `n_rect` tilted rectangular patches in front of it (depth steps at their borders), sensor
    noise, disc-shaped holes and - optionally - one curved (non-planar) patch: more planes, more merges and more rejected blocks per
    frame than render_frame's three-plane room.  For the plane-extractor parity tests."""
    rng = np.random.Generator(np.random.Philox(key=int(seed) * 4099 + 7))
    uu, vv = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    z = 3.0 + 0.0006 * (uu - width / 2) + 0.0003 * (vv - height / 2)
    for _ in range(n_rect):
        w, h = int(rng.integers(60, width // 2)), int(rng.integers(60, height // 2))
        x0, y0 = int(rng.integers(0, width - w)), int(rng.integers(0, height - h))
        zc = rng.uniform(0.8, 2.8)
        a, b = rng.normal(0, 0.0015, 2)
        patch = zc + a * (uu - x0) + b * (vv - y0)
        m = (uu >= x0) & (uu < x0 + w) & (vv >= y0) & (vv < y0 + h) & (patch < z)
        z = np.where(m, patch, z)
    if curved:
        cx, cy, r = rng.uniform(100, width - 100), rng.uniform(100, height - 100), rng.uniform(50, 110)
        d2 = (uu - cx) ** 2 + (vv - cy) ** 2
        z = np.where(d2 < r * r, np.minimum(z, 1.5 - 0.4 * np.sqrt(np.maximum(1 - d2 / (r * r), 0))), z)
    z = z + rng.normal(0, noise_mm * 1e-3, z.shape) * (z / 2.0) ** 2
    d16 = np.clip(np.rint(z * DEPTH_FACTOR), 0, 65535).astype(np.uint16)
    yy, xx = np.mgrid[0:height, 0:width]
    covered, target = 0.0, hole_frac * width * height
    while covered < target:
        hx, hy, hr = rng.uniform(0, width), rng.uniform(0, height), rng.uniform(2.0, 9.0)
        d16[(xx - hx) ** 2 + (yy - hy) ** 2 < hr * hr] = 0
        covered += np.pi * hr * hr
    return d16


def noisy_depth(d16: np.ndarray, seed: int, outlier_frac: float = 0.2, sigma_m: float = 0.01):
    """Corrupt a uint16 depth frame for the 3-D line-fit tests: `outlier_frac` of the pixels scaled by a random factor in [0.7, 1.3] and
    Gaussian noise of `sigma_m` metres everywhere, so that the per-line RANSAC needs several iterations, rejects hypotheses in
    verify3dLine and refits.  Returns uint16 (same 1/5000 m units)."""
    rng = np.random.Generator(np.random.Philox(key=int(seed) * 131 + 9))
    z = d16.astype(np.float64) / DEPTH_FACTOR
    bad = rng.random(z.shape) < outlier_frac
    z = np.where(bad, z * rng.uniform(0.7, 1.3, z.shape), z) + rng.normal(0, sigma_m, z.shape)
    out = np.clip(np.rint(z * DEPTH_FACTOR), 0, 65535).astype(np.uint16)
    out[d16 == 0] = 0
    return out
