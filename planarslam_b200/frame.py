"""Python mirrors of the per-frame Frame:: helpers that sit between feature extraction and matching, on top of the C ABI."""
from __future__ import annotations

import numpy as np

from ._lib import Context


def ComputeStereoFromRGBD(ctx: Context, keys: np.ndarray, n, depth: np.ndarray, depth_factor: float, bf: float, keys_un: np.ndarray | None = None):
    """void Frame::ComputeStereoFromRGBD(const cv::Mat& imDepth) (src/Frame.cc:603-621) for a batch.
    keys: pslam_keypoint records [nframes][cap] (as ORBextractor returns them, padded), n [nframes]; depth uint16 [nframes][h][w] raw.
    Returns (mvuRight, mvDepth), float32 [nframes][cap]."""
    k = np.ascontiguousarray(keys)
    assert k.ndim == 2 and k.dtype.itemsize == 28
    nframes, cap = k.shape
    ku = k if keys_un is None else np.ascontiguousarray(keys_un)
    nn = np.ascontiguousarray(np.broadcast_to(np.asarray(n, np.int32), (nframes,)))
    d = np.ascontiguousarray(depth, np.uint16).reshape(nframes, ctx.cfg.height, ctx.cfg.width)
    ur, dz = np.zeros((nframes, cap), np.float32), np.zeros((nframes, cap), np.float32)
    ctx.check(ctx.L.pslam_compute_stereo_from_rgbd_batch(ctx.h, k.ctypes.data, ku.ctypes.data, nn.ctypes.data, cap, d.ctypes.data, nframes, float(depth_factor), float(bf),
                                                         ur.ctypes.data, dz.ctypes.data))
    return ur, dz
