"""CPU: the C++ adapter (reference class interfaces over the C ABI) compiles and links against libpslam_b200.so."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "pslam_adapter.hpp"
int main(int argc, char**) {
    if (argc > 100) {   // never executed on the CPU box: construction needs a GPU; this only has to compile and link
        pslam_adapter::ORBextractor ext(1000, 1.2f, 8, 20, 7);
        std::vector<pslam_keypoint> k; std::vector<uint8_t> d;
        pslam_adapter::Image8 img{nullptr, 0, 0, 0};
        ext(img, nullptr, k, d);
        pslam_adapter::PlaneDetection pd; float K[4] = {535.4f, 539.2f, 320.1f, 247.6f};
        pslam_adapter::Image16 dep{nullptr, 0, 0};
        pd.readDepthImage(dep, K, 1.f / 5000.f); pd.runPlaneDetection(480, 640);
        pslam_adapter::Context ctx(640, 480);
        pslam_adapter::Optimizer opt(ctx);
        pslam_pose_problem p{}; float T[16] = {0};
        std::vector<uint8_t> a, b, c, e, f;
        std::vector<pslam_keyline> kl; std::vector<pslam_line3d> l3; int32_t draws = 0;
        pslam_adapter::isLineGood(ctx, kl, dep, 1.f / 5000.f, K, draws, l3);
        std::vector<float> ur, dz, nrm; std::vector<double> dir;
        pslam_adapter::ComputeStereoFromRGBD(ctx, k, k, dep, 1.f / 5000.f, 40.f, ur, dz);
        const float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        pslam_manhattan_result mr = pslam_adapter::TrackManhattanFrame(ctx, R, nrm, dir, a, b);
        pslam_line_frustum_frame ff{}; std::vector<int32_t> lvl; std::vector<float> mx, mn;
        const int nv = pslam_adapter::LinesInFrustum(ctx, ff, dir, dir, mx, mn, 0.6f, a, ur, lvl, dz);
        return opt.PoseOptimization(p, T, a, b, c, e, f) + nv + mr.svd_applied;
    }
    return 0;
}
'''


def test_adapter_compiles_and_links():
    import __graft_entry__ as g
    g.build()
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cc")
        open(src, "w").write(SRC)
        exe = os.path.join(td, "t")
        lib_dir = os.path.join(ROOT, "planarslam_b200")
        subprocess.run(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", lib_dir, "-lpslam_b200",
                        f"-Wl,-rpath,{lib_dir}"], check=True)
        subprocess.run([exe], check=True)
