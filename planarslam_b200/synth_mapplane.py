"""Synthetic inputs of MapPlane::UpdateCoefficientsAndPoints (src/MapPlane.cc:298-365): one world plane seen from several key frames, every observation a
voxel-filtered cloud in its camera frame (what Frame::ComputePlanes leaves in mvPlanePoints) with the camera's inverse pose."""
import numpy as np


def _rot(axis, angle):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


def make_map_plane(seed: int, n_obs: int = 6, pts_per_obs: int = 500, extent: float = 3.0, noise: float = 0.004, with_current: bool = False):
    """Returns a list of (points float32 [k][3] in the camera frame, Twc float64 [4][4]); with_current appends the plane's current world-frame cloud with the
    identity (the (Frame, id) overload)."""
    rng = np.random.Generator(np.random.Philox(key=int(seed) * 4099 + 17))
    n = rng.normal(size=3); n /= np.linalg.norm(n)
    u = np.cross(n, [0.3, 0.5, 0.8]); u /= np.linalg.norm(u)
    v = np.cross(n, u)
    origin = rng.uniform(-2, 2, 3)
    clouds = []
    for k in range(n_obs):
        centre = rng.uniform(-extent / 2, extent / 2, 2)
        ab = centre + rng.uniform(-extent / 3, extent / 3, (pts_per_obs + int(rng.integers(0, 60)), 2))
        world = origin + ab[:, :1] * u + ab[:, 1:] * v + rng.normal(0, noise, (len(ab), 1)) * n
        R = _rot(rng.normal(size=3), rng.uniform(0, 1.0))
        t = rng.uniform(-1.5, 1.5, 3)
        Twc = np.eye(4); Twc[:3, :3] = R; Twc[:3, 3] = t
        cam = (world - t) @ R                       # R^T (X - t)
        clouds.append((cam.astype(np.float32), Twc))
    if with_current:
        ab = rng.uniform(-extent / 2, extent / 2, (300, 2))
        world = origin + ab[:, :1] * u + ab[:, 1:] * v
        clouds.append((world.astype(np.float32), np.eye(4)))
    return clouds
