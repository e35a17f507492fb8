"""Synthetic inputs for Tracking::TrackManhattanFrame: surface normals and 3-D line directions of a Manhattan world seen from a
camera whose Manhattan rotation is R_true, plus clutter; R_last = R_true perturbed by a few degrees."""
import numpy as np


def _rot(axis, ang):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def make_manhattan(seed: int, n_normals: int = 8560, n_lines: int = 30, weights=(0.45, 0.3, 0.15), clutter: float = 0.1, noise_deg: float = 2.0,
                   perturb_deg: float = 4.0):
    """Returns (R_last float32 3x3, normals float32 [n][3], dirs float64 [m][3], R_true float64 3x3).  Column a of R_true is the
    direction of Manhattan axis a in camera coordinates; weights = share of the normals on each axis (the rest is clutter)."""
    rng = np.random.default_rng(seed)
    R_true = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
    def sample(n, w):
        w = np.asarray(w, np.float64)
        w = np.append(w * (1 - clutter) / w.sum(), clutter)
        which = rng.choice(4, size=n, p=w)
        v = np.zeros((n, 3))
        for a in range(3):
            k = which == a
            sign = rng.choice([-1.0, 1.0], size=int(k.sum())) if a else np.ones(int(k.sum()))      # floor normals point one way, walls both
            v[k] = R_true[:, a] * sign[:, None]
        v[which == 3] = rng.normal(size=(int((which == 3).sum()), 3))
        v += rng.normal(size=v.shape) * np.deg2rad(noise_deg)
        return v / np.linalg.norm(v, axis=1, keepdims=True)
    normals = sample(n_normals, weights).astype(np.float32)
    dirs = sample(n_lines, (1, 1, 1))
    R_last = (R_true @ _rot(rng.normal(size=3), np.deg2rad(perturb_deg))).astype(np.float32)
    return R_last, normals, dirs, R_true
