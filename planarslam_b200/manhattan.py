"""Python mirror of Tracking::TrackManhattanFrame (src/Tracking.cc:963-1137) on top of the C ABI (pslam_track_manhattan_batch)."""
from __future__ import annotations

import numpy as np

from ._lib import Context

MANHATTAN_RESULT_DTYPE = np.dtype([("R", "<f4", (3, 3)), ("density", "<f4", 3), ("found", "<i4", 3), ("n_cone", "<i4", 3), ("n_selected", "<i4", 3),
                                   ("min_num", "<i4"), ("svd_applied", "<i4")])
assert MANHATTAN_RESULT_DTYPE.itemsize == 92


def TrackManhattanFrame(ctx: Context, R_last: np.ndarray, normals, dirs):
    """Batch form.  R_last [nframes][3][3] float32; normals: list of [n_f][3] float32 arrays (SurfaceNormal::normal); dirs: list of [m_f][3]
    float64 arrays (FrameLine::direction).  Returns (MANHATTAN_RESULT_DTYPE [nframes], normal masks list, dir masks list)."""
    nf = len(normals)
    assert len(dirs) == nf
    Rl = np.ascontiguousarray(R_last, np.float32).reshape(nf, 9)
    nn = np.array([len(x) for x in normals], np.int32)
    nd = np.array([len(x) for x in dirs], np.int32)
    mn, md = max(int(nn.max()), 1), max(int(nd.max()), 1)
    N, D = np.zeros((nf, mn, 3), np.float32), np.zeros((nf, md, 3), np.float64)
    for f in range(nf):
        N[f, :nn[f]] = np.asarray(normals[f], np.float32).reshape(-1, 3)
        D[f, :nd[f]] = np.asarray(dirs[f], np.float64).reshape(-1, 3)
    res = np.zeros(nf, MANHATTAN_RESULT_DTYPE)
    nmask, dmask = np.zeros((nf, mn), np.uint8), np.zeros((nf, md), np.uint8)
    ctx.check(ctx.L.pslam_track_manhattan_batch(ctx.h, Rl.ctypes.data, N.ctypes.data, nn.ctypes.data, mn, D.ctypes.data, nd.ctypes.data, md, nf, res.ctypes.data,
                                                nmask.ctypes.data, dmask.ctypes.data))
    return res, [nmask[f, :nn[f]].copy() for f in range(nf)], [dmask[f, :nd[f]].copy() for f in range(nf)]
