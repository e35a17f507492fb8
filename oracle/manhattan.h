// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of the per-frame Manhattan-frame tracking step:
//   Tracking::TrackManhattanFrame   src/Tracking.cc:963-1137
//   Tracking::ProjectSN2Conic       src/Tracking.cc:888-961   (cone test per axis: sin 0.2018 for surface normals, sin 0.1018 for line directions)
//   Tracking::ProjectSN2MF          src/Tracking.cc:763-886   (cone sin 0.2518, tangent-plane coordinates, one mean-shift step, back-projection)
//   Tracking::MeanShift             src/Tracking.cc:1139-1157 (Gaussian kernel c = 20)
// Kept from the reference: `cv::Mat R_cm = R_cm_update;` is a shallow copy (:970), so the columns written for axis 1 are read when
// the cones of axes 2 and 3 are rebuilt in the second pass, and "R_cm_update = R_cm" in the < 2 directions branch (:1051) is a
// no-op: the partially updated matrix is returned without the SVD.
// Third-party arithmetic: cv::SVD on a 3x3 float matrix (oracle/cvsvd.h), cv::gemm float with double accumulators, cv::norm,
// cv::determinant (only through |det + 1| < 0.5).  PARITY UNPINNED: the reference ships no vectors for this path; the input surface
// normals come from PCL, so tests feed synthetic normals.
// Pinned: agrees with the reference's own Tracking::TrackManhattanFrame (src/Tracking.cc compiled unmodified, oracle/ref/track_driver.cc ->
// oracle/_ref/libtrack_ref.so) to 1-3 float ulp on 15 cases (tests/test_oracle_manhattan_ref.py, tests/golden/manhattan_reference.npz).
#pragma once
#include <cstdint>
#include <vector>

namespace oracle {

struct ManhattanResult {
    float R[9];                 // returned R_cm (row-major)
    int found[3];               // directionFound1..3
    float density[3];           // s_j_density per axis (0 when not found)
    int n_cone[3];              // surface normals inside the first-pass cone (numInCone)
    int n_selected[3];          // m_j_selected.size() of the second pass
    int min_num;                // minNumOfSN used
    int svd_applied;            // 0 when fewer than two directions were found
};

// normals: [n][3] float (SurfaceNormal::normal); dirs: [m][3] double (FrameLine::direction); R_last row-major 3x3 float.
// normal_mask [n] / dir_mask [m] (optional): bit a-1 set when the element was appended to vSurfaceNormal{x,y,z} / vVanishingLine{x,y,z}.
void track_manhattan_frame(const float* R_last, const float* normals, int n, const double* dirs, int m, ManhattanResult& res, uint8_t* normal_mask,
                           uint8_t* dir_mask);

}  // namespace oracle
