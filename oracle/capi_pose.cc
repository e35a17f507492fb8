// TEST INFRASTRUCTURE ONLY — C entry point (ctypes) onto the oracle pose optimisation.
#include "poseopt.h"
#include <cstring>
using namespace oracle;
extern "C" {
struct orc_pose_problem {
    float fx, fy, cx, cy, bf;
    int32_t n_points; const float* Xw; const float* obs; const float* inv_sigma2;
    int32_t n_lines; const double* line_Xw; const double* line_obs;
    int32_t n_planes, n_par, n_ver;
    const float *plane_meas, *plane_map, *par_meas, *par_map, *ver_meas, *ver_map;
    double angle_info, dist_info, par_info, ver_info, plane_chi, vp_chi;
};
// trace: per round {lm_iterations, trials, n_bad} ints [4][3]; chi/lambda doubles [4][2]
static int run_pose(int mode, const orc_pose_problem* p, const float* Tcw_in, float* Tcw_out, double* Tcw_d, uint8_t* o_pt, uint8_t* o_line,
                    uint8_t* o_plane, uint8_t* o_par, uint8_t* o_ver, int32_t* trace_i, double* trace_d) {
    PoseProblem P;
    P.fx = p->fx; P.fy = p->fy; P.cx = p->cx; P.cy = p->cy; P.bf = p->bf;
    P.n_points = p->n_points; P.Xw = p->Xw; P.obs = p->obs; P.inv_sigma2 = p->inv_sigma2;
    P.n_lines = p->n_lines; P.line_Xw = p->line_Xw; P.line_obs = p->line_obs;
    P.n_planes = p->n_planes; P.n_par = p->n_par; P.n_ver = p->n_ver;
    P.plane_meas = p->plane_meas; P.plane_map = p->plane_map; P.par_meas = p->par_meas; P.par_map = p->par_map;
    P.ver_meas = p->ver_meas; P.ver_map = p->ver_map;
    P.angle_info = p->angle_info; P.dist_info = p->dist_info; P.par_info = p->par_info; P.ver_info = p->ver_info;
    P.plane_chi = p->plane_chi; P.vp_chi = p->vp_chi;
    PoseResult R;
    if (mode == 0) pose_optimization(P, Tcw_in, R); else translation_optimization(P, Tcw_in, R);
    std::memcpy(Tcw_out, R.Tcw, sizeof R.Tcw);
    if (Tcw_d) std::memcpy(Tcw_d, R.Tcw_d, sizeof R.Tcw_d);
    if (o_pt && P.n_points) std::memcpy(o_pt, R.outlier_pt.data(), P.n_points);
    if (o_line && P.n_lines) std::memcpy(o_line, R.outlier_line.data(), P.n_lines);
    if (o_plane && P.n_planes) std::memcpy(o_plane, R.outlier_plane.data(), P.n_planes);
    if (o_par && P.n_par) std::memcpy(o_par, R.outlier_par.data(), P.n_par);
    if (o_ver && P.n_ver) std::memcpy(o_ver, R.outlier_ver.data(), P.n_ver);
    for (int r = 0; r < 4; ++r) {
        const bool ok = r < R.n_rounds;
        if (trace_i) { trace_i[3 * r] = ok ? R.rounds[r].lm_iterations : -1; trace_i[3 * r + 1] = ok ? R.rounds[r].trials : -1; trace_i[3 * r + 2] = ok ? R.rounds[r].n_bad : -1; }
        if (trace_d) { trace_d[2 * r] = ok ? R.rounds[r].chi2_final : 0; trace_d[2 * r + 1] = ok ? R.rounds[r].lambda_final : 0; }
    }
    return R.n_inliers;
}
int orc_pose_optimization(const orc_pose_problem* p, const float* Tcw_in, float* Tcw_out, double* Tcw_d, uint8_t* o_pt, uint8_t* o_line,
                          uint8_t* o_plane, uint8_t* o_par, uint8_t* o_ver, int32_t* trace_i, double* trace_d) {
    return run_pose(0, p, Tcw_in, Tcw_out, Tcw_d, o_pt, o_line, o_plane, o_par, o_ver, trace_i, trace_d);
}
int orc_translation_optimization(const orc_pose_problem* p, const float* Tcw_in, float* Tcw_out, double* Tcw_d, uint8_t* o_pt, uint8_t* o_line,
                                 uint8_t* o_plane, uint8_t* o_par, uint8_t* o_ver, int32_t* trace_i, double* trace_d) {
    return run_pose(1, p, Tcw_in, Tcw_out, Tcw_d, o_pt, o_line, o_plane, o_par, o_ver, trace_i, trace_d);
}
}
