// TEST INFRASTRUCTURE ONLY.  include/MapDrawer.h and include/Viewer.h name Pangolin's matrix type in viewer member declarations; src/Optimizer.cc reaches
// them through Optimizer.h -> LoopClosing.h -> Tracking.h.  Compile-only stand-in: nothing of the viewer is compiled or linked.
#pragma once
namespace pangolin {
struct OpenGlMatrix { double m[16]; void SetIdentity() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; } };
}
