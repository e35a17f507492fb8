"""Host-side mirror of the reference's Optimizer (include/Optimizer.h:27-46) for the pose-only path.

    Optimizer(ctx).PoseOptimization(problem_dict) -> (n_inliers, Tcw float32 4x4, outlier flags ...)

A "problem" is the plain-array view of what Optimizer::PoseOptimization reads from a Frame (see pslam_pose_problem in
include/pslam_abi.h); planarslam_b200.synth_pose.make_pose_problem builds seeded ones.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Context, PoseProblem

_ARRS = ("Xw", "obs", "inv_sigma2", "line_Xw", "line_obs", "plane_meas", "plane_map", "par_meas", "par_map", "ver_meas", "ver_map")
_SCAL = ("fx", "fy", "cx", "cy", "bf", "angle_info", "dist_info", "par_info", "ver_info", "plane_chi", "vp_chi")


def _to_struct(p: dict) -> PoseProblem:
    s = PoseProblem()
    for k in _SCAL:
        setattr(s, k, p[k])
    s.n_points, s.n_lines = len(p["Xw"]), len(p["line_Xw"])
    s.n_planes, s.n_par, s.n_ver = len(p["plane_meas"]), len(p["par_meas"]), len(p["ver_meas"])
    for k in _ARRS:
        setattr(s, k, p[k].ctypes.data if p[k].size else None)
    return s


class Optimizer:
    def __init__(self, ctx: Context | None = None, device: int = 0):
        self.ctx = ctx or Context(640, 480, 1, device)

    def pack(self, problems: list[dict], translation_only: bool = False):
        """Pack + upload problems once (device-resident form used by bench.py)."""
        self._keep = problems
        arr = (PoseProblem * len(problems))(*[_to_struct(p) for p in problems])
        T0 = np.ascontiguousarray(np.stack([p["Tcw0"] for p in problems]), np.float32)
        fn = self.ctx.L.pslam_translation_pack if translation_only else self.ctx.L.pslam_pose_pack
        self.ctx.check(fn(self.ctx.h, arr, len(problems), T0.ctypes.data))
        self._counts = [(len(p["Xw"]), len(p["line_Xw"]), len(p["plane_meas"]), len(p["par_meas"]), len(p["ver_meas"])) for p in problems]

    def run_packed(self):
        self.ctx.check(self.ctx.L.pslam_pose_run_packed(self.ctx.h))

    def fetch(self):
        n = len(self._counts)
        tot = np.sum(np.array(self._counts), 0)
        T, Td = np.zeros((n, 4, 4), np.float32), np.zeros((n, 4, 4))
        fl = [np.zeros(max(int(t), 1), np.uint8) for t in tot]
        ninl = np.zeros(n, np.int32)
        ti, td = np.zeros((n, 4, 3), np.int32), np.zeros((n, 4, 2))
        self.ctx.check(self.ctx.L.pslam_pose_fetch(self.ctx.h, T.ctypes.data, Td.ctypes.data, *[f.ctypes.data for f in fl], ninl.ctypes.data,
                                                   ti.ctypes.data, td.ctypes.data))
        out, off = [], np.zeros(5, int)
        for i, cnt in enumerate(self._counts):
            out.append(dict(Tcw=T[i], Tcw_d=Td[i], n_inliers=int(ninl[i]), trace_i=ti[i], trace_d=td[i],
                            outlier_pt=fl[0][off[0]:off[0] + cnt[0]], outlier_line=fl[1][off[1]:off[1] + cnt[1]],
                            outlier_plane=fl[2][off[2]:off[2] + cnt[2]], outlier_par=fl[3][off[3]:off[3] + cnt[3]],
                            outlier_ver=fl[4][off[4]:off[4] + cnt[4]]))
            off += np.array(cnt)
        return out

    def PoseOptimizationBatch(self, problems: list[dict]):
        self.pack(problems)
        self.run_packed()
        return self.fetch()

    # static int TranslationOptimization(Frame*)
    def TranslationOptimization(self, problem: dict):
        self.pack([problem], translation_only=True)
        self.run_packed()
        r = self.fetch()[0]
        return r["n_inliers"], r

    def TranslationOptimizationBatch(self, problems: list[dict]):
        self.pack(problems, translation_only=True)
        self.run_packed()
        return self.fetch()

    # static int PoseOptimization(Frame*): returns the inlier count; pose and outlier flags come back in the dict
    def PoseOptimization(self, problem: dict):
        r = self.PoseOptimizationBatch([problem])[0]
        return r["n_inliers"], r
