"""Host-side mirror of the reference's LineSegment (include/LSDextractor.h:349, src/LSDextractor.cpp:13-39) on top of the C ABI.

    LineSegment(ctx).ExtractLineSegment(gray) -> (keylines structured array, line functions [n][3])

ExtractLineSegmentWithDescriptors adds the LBD descriptors (pslam_lines_extract_describe_batch; restated from the published algorithm - no pinnable upstream
implementation here).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Context, E_CAPACITY

KEYLINE_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt", "<f4", 2), ("response", "<f4"), ("size", "<f4"),
                          ("startPointX", "<f4"), ("startPointY", "<f4"), ("endPointX", "<f4"), ("endPointY", "<f4"),
                          ("sPointInOctaveX", "<f4"), ("sPointInOctaveY", "<f4"), ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"),
                          ("lineLength", "<f4"), ("numOfPixels", "<i4")])       # cv::line_descriptor::KeyLine, 68 bytes


class LineSegment:
    def __init__(self, ctx: Context | None = None, width: int = 640, height: int = 480, max_batch: int = 1, device: int = 0):
        self.ctx = ctx or Context(width, height, max_batch, device)

    def set_rect_enumeration(self, mode: int):
        """Pixel enumeration of the NFA validation: 1 OpenCV 4.x's rect_nfa (default), 0 the published LSD rectangle iterator
        (include/pslam_abi.h pslam_lsd_set_rect_enumeration)."""
        self.ctx.check(self.ctx.L.pslam_lsd_set_rect_enumeration(self.ctx.h, int(mode)))

    # cv::LineSegmentDetector::detect (refine: 0 NONE, 1 STD, 2 ADV)
    def detect(self, gray: np.ndarray, refine: int = 2):
        g = np.ascontiguousarray(gray, np.uint8)
        frames = g[None] if g.ndim == 2 else g
        n = len(frames)
        cap = int(self.ctx.L.pslam_lsd_max_segments(self.ctx.h))
        segs, wpn, cnt = np.zeros((n, cap, 4), np.float32), np.zeros((n, cap, 3)), np.zeros(n, np.int32)
        self.ctx.check(self.ctx.L.pslam_lsd_detect_batch(self.ctx.h, frames.ctypes.data, n, refine, segs.ctypes.data, wpn.ctypes.data, cap, cnt.ctypes.data))
        out = [(segs[i, :cnt[i]].copy(), wpn[i, :cnt[i], 0].copy(), wpn[i, :cnt[i], 1].copy(), wpn[i, :cnt[i], 2].copy()) for i in range(n)]
        return out[0] if g.ndim == 2 else out

    # void ExtractLineSegment(const Mat& img, vector<KeyLine>&, Mat& ldesc, vector<Vector3d>& lineFunctions, ...)
    def ExtractLineSegment(self, gray: np.ndarray, max_lines: int = 40):
        g = np.ascontiguousarray(gray, np.uint8)
        frames = g[None] if g.ndim == 2 else g
        n = len(frames)
        kl, lf, cnt = np.zeros((n, max_lines), KEYLINE_DTYPE), np.zeros((n, max_lines, 3)), np.zeros(n, np.int32)
        self.ctx.check(self.ctx.L.pslam_lines_extract_batch(self.ctx.h, frames.ctypes.data, n, max_lines, kl.ctypes.data, lf.ctypes.data, cnt.ctypes.data))
        out = [(kl[i, :cnt[i]].copy(), lf[i, :cnt[i]].copy()) for i in range(n)]
        return out[0] if g.ndim == 2 else out

    # ExtractLineSegment with the LBD descriptors (ldesc / Frame::mLdesc): (keylines, line functions, desc uint8 [n][32], lbd float32 [n][72]) per frame
    def ExtractLineSegmentWithDescriptors(self, gray: np.ndarray, max_lines: int = 40):
        g = np.ascontiguousarray(gray, np.uint8)
        frames = g[None] if g.ndim == 2 else g
        n = len(frames)
        L = self.ctx.L
        L.pslam_lines_extract_describe_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5
        kl, lf, cnt = np.zeros((n, max_lines), KEYLINE_DTYPE), np.zeros((n, max_lines, 3)), np.zeros(n, np.int32)
        desc, lbd = np.zeros((n, max_lines, 32), np.uint8), np.zeros((n, max_lines, 72), np.float32)
        self.ctx.check(L.pslam_lines_extract_describe_batch(self.ctx.h, frames.ctypes.data, n, max_lines, kl.ctypes.data, lf.ctypes.data, desc.ctypes.data,
                                                            lbd.ctypes.data, cnt.ctypes.data))
        out = [(kl[i, :cnt[i]].copy(), lf[i, :cnt[i]].copy(), desc[i, :cnt[i]].copy(), lbd[i, :cnt[i]].copy()) for i in range(n)]
        return out[0] if g.ndim == 2 else out

    def debug_stage(self, frame: int = 0):
        dims = np.zeros(2, np.int32)
        self.ctx.check(self.ctx.L.pslam_lsd_debug_stage(self.ctx.h, frame, dims.ctypes.data, None, None, None, None, None))
        W, H = int(dims[0]), int(dims[1])
        out = dict(scaled=np.zeros((H, W), np.uint8), modgrad=np.zeros((H, W)), angles=np.zeros((H, W)), order=np.zeros(W * H, np.int32))
        no = np.zeros(1, np.int32)
        self.ctx.check(self.ctx.L.pslam_lsd_debug_stage(self.ctx.h, frame, dims.ctypes.data, out["scaled"].ctypes.data, out["modgrad"].ctypes.data,
                                                        out["angles"].ctypes.data, out["order"].ctypes.data, no.ctypes.data))
        out["order"] = out["order"][:int(no[0])]
        return out


LINE3D_DTYPE = np.dtype([("A", "<f8", 3), ("B", "<f8", 3), ("director", "<f8", 3), ("inliers", "<u8"), ("depth", "<f4"), ("n_points", "<i4"),
                         ("n_inliers", "<i4"), ("valid", "<i4")])
assert LINE3D_DTYPE.itemsize == 96


def isLineGood(ctx: Context, keylines: np.ndarray, n_lines, depth: np.ndarray, K, depth_factor: float, seed=1, skip=None):
    """void Frame::isLineGood(const cv::Mat& imGray, const cv::Mat& imDepth, cv::Mat K) (src/Frame.cc:189-267) for a batch of frames.

    keylines: KEYLINE_DTYPE [nframes][max_lines] (as ExtractLineSegment returns them, padded), n_lines [nframes]; depth uint16
    [nframes][h][w] raw (metres = raw * depth_factor); K = (fx, fy, cx, cy); seed / skip: per-frame rand() stream (scalars are
    broadcast).  Returns (LINE3D_DTYPE [nframes][max_lines], n_drawn [nframes])."""
    kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
    assert kl.ndim == 2
    nframes, max_lines = kl.shape
    nl = np.ascontiguousarray(np.broadcast_to(np.asarray(n_lines, np.int32), (nframes,)))
    d = np.ascontiguousarray(depth, np.uint16).reshape(nframes, ctx.cfg.height, ctx.cfg.width)
    cam = np.asarray(K, np.float32)
    sd = np.ascontiguousarray(np.broadcast_to(np.asarray(seed, np.uint32), (nframes,)))
    sk = None if skip is None else np.ascontiguousarray(np.broadcast_to(np.asarray(skip, np.int32), (nframes,)))
    out, drawn = np.zeros((nframes, max_lines), LINE3D_DTYPE), np.zeros(nframes, np.int32)
    ctx.check(ctx.L.pslam_lines3d_batch(ctx.h, kl.ctypes.data, nl.ctypes.data, max_lines, d.ctypes.data, nframes, float(depth_factor), cam.ctypes.data,
                                        sd.ctypes.data, None if sk is None else sk.ctypes.data, out.ctypes.data, drawn.ctypes.data))
    return out, drawn
