// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of the reference plane extractor: PlaneDetection::readDepthImage / runPlaneDetection
// (src/PlaneExtractor.cpp:26-65, include/PlaneExtractor.h:18-34) and the vendored PEAC
// (include/peac/AHCPlaneFitter.hpp:211-260 run, :786-972 initGraph, :983-1189 ahCluster, :299-379 refineDetails,
//  :485-587 findBlockMembership, :428-476 floodFill; AHCPlaneSeg.hpp:60-157 Stats, :211-317 ctors, :379-410
//  mergeNbsFrom; AHCParamSet.hpp:68-146; DisjointSet.hpp:64-92).
//
// Parity conventions fixed here (the reference leaves them to the allocator / Eigen; listed in DESIGN.md):
//  * the 3x3 symmetric eigen-decomposition (Eigen::SelfAdjointEigenSolver in eig33sym.hpp:70-74) is a cyclic
//    Jacobi iteration using only + - * / sqrt, so CPU and GPU agree bit-for-bit;
//  * neighbour sets (std::set<PlaneSeg*>, ordered by heap address, AHCPlaneSeg.hpp:188) iterate in node
//    creation order; the min-MSE priority queue is libstdc++'s binary heap (push_heap / pop_heap);
//  * the final size sort (std::sort, :1185) is a stable insertion sort (what libstdc++ does for <= 16 planes);
//  * no fused multiply-add anywhere (the running sums sx += x etc. are plain IEEE double operations).
#pragma once
#include <cstdint>
#include <vector>

namespace oracle {

struct PeacParams {            // compiled-in defaults of the reference (never overridden)
    int win_w = 10, win_h = 10;
    int min_support = 3000;
    int max_step = 100000;
    double depth_sigma = 1.6e-6, std_tol_init = 5, std_tol_merge = 8;
    double z_near = 500, z_far = 4000;
    double angle_near_deg = 15.0, angle_far_deg = 90.0;
    double sim_merge_deg = 60.0, sim_refine_deg = 30.0;
    double depth_alpha = 0.04, depth_change_tol = 0.02;
};

struct PlaneStats { double sx = 0, sy = 0, sz = 0, sxx = 0, syy = 0, szz = 0, sxy = 0, syz = 0, sxz = 0; int N = 0; };

struct PeacPlane {
    double normal[3], center[3], mse, curvature;
    int N, rid;
    PlaneStats stats;
};

struct PeacBlock {             // stage output of the per-block initialisation (for stage-by-stage parity tests)
    PlaneStats stats;
    double center[3], normal[3], mse, curvature;
    int valid;                 // 1 = node kept in the graph
};

struct PeacResult {
    int w = 0, h = 0;
    std::vector<int32_t> labels;                 // membershipImg after refineDetails (raw: negatives are trail counters)
    std::vector<PeacPlane> planes;               // extractedPlanes, sorted by N descending
    std::vector<std::vector<int>> membership;    // plane_vertices_
    std::vector<PeacBlock> blocks;               // per 10x10 block
    std::vector<int32_t> coarse_block_plane;     // blkMap after erosion (plane index before the final merge, or -1)
    int n_coarse_planes = 0;
    int n_queue = 0, n_seeds = 0;               // region-growing queue length (final / initial)
};

void eig33sym_jacobi(const double K[3][3], double s[3], double V[3][3]);

// depth: h x w uint16; scale = metres per unit as float (the reference passes 1/DepthMapFactor as float);
// fx..cy are the float camera matrix entries (K is CV_32F in the reference).
void peac_run(const uint16_t* depth, int w, int h, float fx, float fy, float cx, float cy, float scale,
              const PeacParams& prm, PeacResult& out);

}  // namespace oracle
