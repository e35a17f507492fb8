"""Host-side mirror of Optimizer::LocalBundleAdjustment (include/Optimizer.h:34, src/Optimizer.cc:1853-2678) on top of the
C ABI (pslam_local_bundle_adjustment*, include/pslam_abi.h).

A "problem" is the plain-array view of the local map the reference gathers at :1853-1969 (see pslam_lba_problem);
planarslam_b200.synth_lba.make_lba_problem builds seeded ones.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Context


class LbaProblem(C.Structure):
    """pslam_lba_problem (include/pslam_abi.h)."""
    _fields_ = [("n_kf", C.c_int32), ("kf_Tcw", C.c_void_p), ("kf_fixed", C.c_void_p), ("kf_K", C.c_void_p),
                ("n_points", C.c_int32), ("pt_Xw", C.c_void_p),
                ("n_pt_obs", C.c_int32), ("pt_obs_kf", C.c_void_p), ("pt_obs_pt", C.c_void_p), ("pt_obs_uvr", C.c_void_p),
                ("pt_obs_inv_sigma2", C.c_void_p),
                ("n_lines", C.c_int32), ("line_Xw", C.c_void_p),
                ("n_line_obs", C.c_int32), ("line_obs_kf", C.c_void_p), ("line_obs_line", C.c_void_p), ("line_obs_l", C.c_void_p),
                ("n_planes", C.c_int32), ("plane_Xw", C.c_void_p),
                ("n_plane_obs", C.c_int32 * 3), ("plane_obs_kf", C.c_void_p * 3), ("plane_obs_plane", C.c_void_p * 3),
                ("plane_obs_meas", C.c_void_p * 3),
                ("angle_info", C.c_double), ("dist_info", C.c_double), ("plane_chi", C.c_double), ("vp_chi", C.c_double)]


class LbaResult(C.Structure):
    """pslam_lba_result (include/pslam_abi.h)."""
    _fields_ = [("kf_Tcw", C.c_void_p), ("kf_Tcw_d", C.c_void_p), ("pt_Xw", C.c_void_p), ("pt_Xw_d", C.c_void_p),
                ("line_Xw", C.c_void_p), ("line_Xw_d", C.c_void_p), ("plane_Xw", C.c_void_p), ("plane_Xw_d", C.c_void_p),
                ("erase_pt", C.c_void_p), ("erase_line", C.c_void_p), ("erase_plane", C.c_void_p * 3),
                ("iterations", C.c_int32 * 2), ("trials", C.c_int32 * 2), ("chi2", C.c_double * 2), ("lambda_", C.c_double * 2)]


def _ptr(a: np.ndarray):
    return a.ctypes.data if a.size else None


def problem_struct(p: dict) -> LbaProblem:
    s = LbaProblem()
    s._keep = p
    s.n_kf, s.n_points, s.n_pt_obs = len(p["kf_Tcw"]), len(p["pt_Xw"]), len(p["pt_obs_kf"])
    s.n_lines, s.n_line_obs, s.n_planes = len(p["line_Xw"]), len(p["line_obs_kf"]), len(p["plane_Xw"])
    for k in ("kf_Tcw", "kf_fixed", "kf_K", "pt_Xw", "pt_obs_kf", "pt_obs_pt", "pt_obs_uvr", "pt_obs_inv_sigma2", "line_Xw",
              "line_obs_kf", "line_obs_line", "line_obs_l", "plane_Xw"):
        setattr(s, k, _ptr(p[k]))
    for t in range(3):
        s.n_plane_obs[t] = len(p["plane_obs_kf"][t])
        s.plane_obs_kf[t] = _ptr(p["plane_obs_kf"][t])
        s.plane_obs_plane[t] = _ptr(p["plane_obs_plane"][t])
        s.plane_obs_meas[t] = _ptr(p["plane_obs_meas"][t])
    for k in ("angle_info", "dist_info", "plane_chi", "vp_chi"):
        setattr(s, k, p[k])
    return s


def result_struct(s: LbaProblem):
    """Allocates every output array for one problem; returns (struct, dict of numpy arrays)."""
    o = dict(kf_Tcw=np.zeros((s.n_kf, 4, 4), np.float32), kf_Tcw_d=np.zeros((s.n_kf, 4, 4)),
             pt_Xw=np.zeros((s.n_points, 3), np.float32), pt_Xw_d=np.zeros((s.n_points, 3)),
             line_Xw=np.zeros((s.n_lines, 6)), line_Xw_d=np.zeros((s.n_lines, 6)),
             plane_Xw=np.zeros((s.n_planes, 4), np.float32), plane_Xw_d=np.zeros((s.n_planes, 4)),
             erase_pt=np.zeros(s.n_pt_obs, np.uint8), erase_line=np.zeros(s.n_line_obs, np.uint8),
             erase_plane=[np.zeros(s.n_plane_obs[t], np.uint8) for t in range(3)])
    r = LbaResult()
    r._keep = o
    for k in ("kf_Tcw", "kf_Tcw_d", "pt_Xw", "pt_Xw_d", "line_Xw", "line_Xw_d", "plane_Xw", "plane_Xw_d", "erase_pt", "erase_line"):
        setattr(r, k, _ptr(o[k]))
    for t in range(3):
        r.erase_plane[t] = _ptr(o["erase_plane"][t])
    return r, o


def finish(r: LbaResult, o: dict) -> dict:
    o = dict(o)
    o["iterations"], o["trials"] = list(r.iterations), list(r.trials)
    o["chi2"], o["lambda"] = list(r.chi2), list(r.lambda_)
    return o


class LocalBundleAdjuster:
    def __init__(self, ctx: Context | None = None, device: int = 0):
        self.ctx = ctx or Context(640, 480, 1, device)

    # static void Optimizer::LocalBundleAdjustment(KeyFrame*, bool*, Map*)
    def LocalBundleAdjustment(self, problem: dict) -> dict:
        return self.LocalBundleAdjustmentBatch([problem])[0]

    def LocalBundleAdjustmentBatch(self, problems: list[dict]) -> list[dict]:
        n = len(problems)
        ps = [problem_struct(p) for p in problems]
        rs = [result_struct(s) for s in ps]
        parr = (LbaProblem * n)(*ps)
        rarr = (LbaResult * n)(*[r for r, _ in rs])
        self.ctx.check(self.ctx.L.pslam_local_bundle_adjustment_batch(self.ctx.h, parr, n, rarr))
        return [finish(rarr[i], rs[i][1]) for i in range(n)]

    # split form used by bench.py
    def pack(self, problems: list[dict]):
        self._ps = [problem_struct(p) for p in problems]
        parr = (LbaProblem * len(problems))(*self._ps)
        self.ctx.check(self.ctx.L.pslam_lba_pack(self.ctx.h, parr, len(problems)))

    def run_packed(self):
        self.ctx.check(self.ctx.L.pslam_lba_run_packed(self.ctx.h))

    def fetch(self) -> list[dict]:
        rs = [result_struct(s) for s in self._ps]
        rarr = (LbaResult * len(rs))(*[r for r, _ in rs])
        self.ctx.check(self.ctx.L.pslam_lba_fetch(self.ctx.h, rarr))
        return [finish(rarr[i], rs[i][1]) for i in range(len(rs))]
