"""Host-side mirror of the device-resident tracking chain (include/pslam_abi.h pslam_track_*): Tracking::TrackWithMotionModel + TrackLocalMap
(src/Tracking.cc:1739-1859, 1954-2046) of a replayed RGB-D sequence against a POD map snapshot, without host round trips between stages or frames."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Context


class TrackParams(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float), ("depth_factor", C.c_float),
                ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float), ("th_last", C.c_float), ("th_map", C.c_float),
                ("nnratio_map", C.c_float), ("use_motion_model", C.c_int32)]


class MapPointsC(C.Structure):
    _fields_ = [("n", C.c_int32), ("pos", C.c_void_p), ("normal", C.c_void_p), ("max_distance", C.c_void_p), ("min_distance", C.c_void_p),
                ("desc", C.c_void_p), ("skip", C.c_void_p), ("has_obs", C.c_void_p)]


def default_params(K=(535.4, 539.2, 320.1, 247.6), bf=40.0, depth_factor=1.0 / 5000.0, width=640, height=480, use_motion_model=True) -> TrackParams:
    return TrackParams(K[0], K[1], K[2], K[3], bf, np.float32(depth_factor), 0.0, float(width), 0.0, float(height), 15.0, 3.0, 0.8, int(use_motion_model))


class Tracker:
    """tracker = Tracker(ctx); tracker.set_map(map_points_dict); poses, stats = tracker.track(gray [n,H,W] u8, depth [n,H,W] u16, Tcw0)"""

    def __init__(self, ctx: Context, params: TrackParams | None = None):
        self.ctx, self.params = ctx, params or default_params(width=ctx.cfg.width, height=ctx.cfg.height)
        L = ctx.L
        L.pslam_track_set_map.argtypes = [C.c_void_p, C.c_void_p]
        L.pslam_track_sequence.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self._keep = None

    def set_map(self, m: dict):
        a = {k: np.ascontiguousarray(m[k], dt) for k, dt in (("pos", np.float32), ("normal", np.float32), ("max_distance", np.float32), ("min_distance", np.float32),
                                                             ("desc", np.uint8), ("skip", np.uint8), ("has_obs", np.uint8))}
        s = MapPointsC(len(a["skip"]), *[a[k].ctypes.data for k in ("pos", "normal", "max_distance", "min_distance", "desc", "skip", "has_obs")])
        self._keep = a
        self.ctx.check(self.ctx.L.pslam_track_set_map(self.ctx.h, C.byref(s)))

    def track(self, gray: np.ndarray, depth: np.ndarray, Tcw0: np.ndarray):
        g, d = np.ascontiguousarray(gray, np.uint8), np.ascontiguousarray(depth, np.uint16)
        n = len(g)
        T0 = np.ascontiguousarray(Tcw0, np.float32)
        poses, stats = np.zeros((n, 4, 4), np.float32), np.zeros((n, 4), np.int32)
        self.ctx.check(self.ctx.L.pslam_track_sequence(self.ctx.h, g.ctypes.data, d.ctypes.data, n, C.byref(self.params), T0.ctypes.data, poses.ctypes.data,
                                                       stats.ctypes.data))
        return poses, stats
