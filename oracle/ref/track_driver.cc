// TEST INFRASTRUCTURE ONLY.  Tracking::TrackManhattanFrame (src/Tracking.cc:963-1137, with ProjectSN2Conic, ProjectSN2MF, MeanShift, EasyHist) called AS IT IS:
// src/Tracking.cc compiles unmodified against the stand-ins of oracle/ref/shims/ and links against oracle/_ref/libmatch_ref.so (Frame, KeyFrame, Map,
// matchers, Optimizer ...).  The rest of the system that Tracking.cc refers to but TrackManhattanFrame never reaches (local mapping, loop closing, viewer,
// initialiser, PnP solver, mesh viewer) is NOT compiled: the twenty member functions below are link-only stand-ins that abort.  TrackManhattanFrame reads no
// Tracking state except mCurrentFrame (it appends the supporting normals / lines to its vectors), so the driver runs it on raw storage holding a
// default-constructed Frame instead of running Tracking's constructor (which would read a settings file and start the extractors).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <condition_variable>
#include <iomanip>
#include <queue>

#define private public
#define protected public
#include "Tracking.h"
#include "FrameDrawer.h"
#include "Initializer.h"
#include "LocalMapping.h"
#include "LoopClosing.h"
#include "MapDrawer.h"
#include "MeshViewer.h"
#include "PnPsolver.h"
#include "System.h"
#include "Viewer.h"
#undef private
#undef protected

[[noreturn]] static void not_linked(const char* what) { std::fprintf(stderr, "oracle/ref/track_driver.cc: %s is a link-only stand-in\n", what); std::abort(); }

namespace Planar_SLAM {
void FrameDrawer::Update(Tracking*) { not_linked("FrameDrawer::Update"); }
Initializer::Initializer(const Frame&, float, int) { not_linked("Initializer::Initializer"); }
bool LocalMapping::AcceptKeyFrames() { not_linked("LocalMapping::AcceptKeyFrames"); }
void LocalMapping::InsertKeyFrame(KeyFrame*) { not_linked("LocalMapping::InsertKeyFrame"); }
void LocalMapping::InterruptBA() { not_linked("LocalMapping::InterruptBA"); }
void LocalMapping::RequestReset() { not_linked("LocalMapping::RequestReset"); }
bool LocalMapping::SetNotStop(bool) { not_linked("LocalMapping::SetNotStop"); }
bool LocalMapping::isStopped() { not_linked("LocalMapping::isStopped"); }
bool LocalMapping::stopRequested() { not_linked("LocalMapping::stopRequested"); }
void LoopClosing::RequestReset() { not_linked("LoopClosing::RequestReset"); }
void MapDrawer::SetCurrentCameraPose(const cv::Mat&) { not_linked("MapDrawer::SetCurrentCameraPose"); }
void MeshViewer::SaveMeshModel(const string&) { not_linked("MeshViewer::SaveMeshModel"); }
void MeshViewer::print() { not_linked("MeshViewer::print"); }
PnPsolver::PnPsolver(const Frame&, const vector<MapPoint*>&) { not_linked("PnPsolver::PnPsolver"); }
void PnPsolver::SetRansacParameters(double, int, int, int, float, float) { not_linked("PnPsolver::SetRansacParameters"); }
cv::Mat PnPsolver::iterate(int, bool&, vector<bool>&, int&) { not_linked("PnPsolver::iterate"); }
void System::Reset() { not_linked("System::Reset"); }
void Viewer::Release() { not_linked("Viewer::Release"); }
void Viewer::RequestStop() { not_linked("Viewer::RequestStop"); }
bool Viewer::isStopped() { not_linked("Viewer::isStopped"); }
}  // namespace Planar_SLAM

using namespace Planar_SLAM;

// R_last: 3x3 float row-major (mLastRcm); normals [n][3] float (SurfaceNormal::normal); dirs [m][3] double (FrameLine::direction).  R_out: the returned matrix.
extern "C" int ref_track_manhattan_frame(const float* R_last, const float* normals, int n, const double* dirs, int m, float* R_out) {
    void* raw = std::aligned_alloc(alignof(Tracking), ((sizeof(Tracking) + alignof(Tracking) - 1) / alignof(Tracking)) * alignof(Tracking));
    std::memset(raw, 0, sizeof(Tracking));
    Tracking* t = reinterpret_cast<Tracking*>(raw);
    new (&t->mCurrentFrame) Frame();
    cv::Mat last(3, 3, CV_32F);
    for (int i = 0; i < 9; ++i) last.at<float>(i / 3, i % 3) = R_last[i];
    std::vector<SurfaceNormal> sn(n);
    for (int i = 0; i < n; ++i) { sn[i].normal = cv::Point3f(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]); sn[i].cameraPosition = cv::Point3f(0, 0, 1); sn[i].FramePosition = cv::Point2i(0, 0); }
    std::vector<FrameLine> fl(m);
    for (int i = 0; i < m; ++i) { fl[i].direction = cv::Point3d(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]); fl[i].haveDepth = true; }
    cv::Mat R = t->TrackManhattanFrame(last, sn, fl);
    for (int i = 0; i < 9; ++i) R_out[i] = (float)R.getd(i / 3, i % 3);
    t->mCurrentFrame.~Frame();
    std::free(raw);
    return R.rows * 10 + R.cols;
}
