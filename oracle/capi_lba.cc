// TEST INFRASTRUCTURE ONLY — C entry point (ctypes) onto the oracle local bundle adjustment.
#include "lba.h"
#include <cstring>
using namespace oracle;
template <class T> static void put(T* dst, const std::vector<T>& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(T)); }
extern "C" {
// same layout as pslam_lba_problem / pslam_lba_result (include/pslam_abi.h) so the tests share one ctypes mirror
struct orc_lba_problem {
    int32_t n_kf; const float* kf_Tcw; const uint8_t* kf_fixed; const float* kf_K;
    int32_t n_points; const float* pt_Xw;
    int32_t n_pt_obs; const int32_t* pt_obs_kf; const int32_t* pt_obs_pt; const float* pt_obs_uvr; const float* pt_obs_inv_sigma2;
    int32_t n_lines; const double* line_Xw;
    int32_t n_line_obs; const int32_t* line_obs_kf; const int32_t* line_obs_line; const double* line_obs_l;
    int32_t n_planes; const float* plane_Xw;
    int32_t n_plane_obs[3]; const int32_t* plane_obs_kf[3]; const int32_t* plane_obs_plane[3]; const float* plane_obs_meas[3];
    double angle_info, dist_info, plane_chi, vp_chi;
};
struct orc_lba_result {
    float* kf_Tcw; double* kf_Tcw_d; float* pt_Xw; double* pt_Xw_d; double* line_Xw; double* line_Xw_d; float* plane_Xw; double* plane_Xw_d;
    uint8_t* erase_pt; uint8_t* erase_line; uint8_t* erase_plane[3];
    int32_t iterations[2], trials[2]; double chi2[2], lambda[2];
};
int orc_local_bundle_adjustment(const orc_lba_problem* p, orc_lba_result* r) {
    LbaProblem P;
    P.n_kf = p->n_kf; P.kf_Tcw = p->kf_Tcw; P.kf_fixed = p->kf_fixed; P.kf_K = p->kf_K;
    P.n_points = p->n_points; P.pt_Xw = p->pt_Xw;
    P.n_pt_obs = p->n_pt_obs; P.pt_obs_kf = p->pt_obs_kf; P.pt_obs_pt = p->pt_obs_pt; P.pt_obs_uvr = p->pt_obs_uvr; P.pt_obs_inv_sigma2 = p->pt_obs_inv_sigma2;
    P.n_lines = p->n_lines; P.line_Xw = p->line_Xw;
    P.n_line_obs = p->n_line_obs; P.line_obs_kf = p->line_obs_kf; P.line_obs_line = p->line_obs_line; P.line_obs_l = p->line_obs_l;
    P.n_planes = p->n_planes; P.plane_Xw = p->plane_Xw;
    for (int t = 0; t < 3; ++t) {
        P.n_plane_obs[t] = p->n_plane_obs[t]; P.plane_obs_kf[t] = p->plane_obs_kf[t]; P.plane_obs_plane[t] = p->plane_obs_plane[t];
        P.plane_obs_meas[t] = p->plane_obs_meas[t];
    }
    P.angle_info = p->angle_info; P.dist_info = p->dist_info; P.plane_chi = p->plane_chi; P.vp_chi = p->vp_chi;
    LbaResult R;
    local_bundle_adjustment(P, R);
    put(r->kf_Tcw, R.kf_Tcw); put(r->kf_Tcw_d, R.kf_Tcw_d); put(r->pt_Xw, R.pt_Xw); put(r->pt_Xw_d, R.pt_Xw_d);
    put(r->line_Xw, R.line_Xw); put(r->line_Xw_d, R.line_Xw_d); put(r->plane_Xw, R.plane_Xw); put(r->plane_Xw_d, R.plane_Xw_d);
    put(r->erase_pt, R.erase_pt); put(r->erase_line, R.erase_line);
    for (int t = 0; t < 3; ++t) put(r->erase_plane[t], R.erase_plane[t]);
    for (int k = 0; k < 2; ++k) { r->iterations[k] = R.iterations[k]; r->trials[k] = R.trials[k]; r->chi2[k] = R.chi2[k]; r->lambda[k] = R.lambda[k]; }
    return 0;
}
}
