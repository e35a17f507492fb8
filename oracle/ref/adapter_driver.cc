// TEST INFRASTRUCTURE ONLY.  The entry points of match_driver.cc with the PRODUCT's reference-typed adapter (include/pslam_reference_adapter.hpp:
// pslam_adapter::ref::ORBmatcher / LSDmatcher / PlaneMatcher / KeyFrameDatabase / Optimizer, same signatures as the reference's classes) in place of the reference's functions.  The Frame /
// MapPoint / MapPlane / MapLine objects are the reference's own (built by the same code from the same plain arrays, classes linked from libmatch_ref.so); the
// adapter gathers from them, runs the CUDA path through the C ABI (libpslam_b200.so) and writes back into them; the read-back is shared with the reference
// build.  tests/test_reference_adapter_gpu.py calls ref_* and adp_* on the same inputs and compares mvpMapPoints / mvb*Outlier / mTcw.
#define PSLAM_ADAPTER_BUILD
#define DRV(name) adp_##name
#define DRV_ORBMATCHER pslam_adapter::ref::ORBmatcher
#define DRV_LSDMATCHER pslam_adapter::ref::LSDmatcher
#define DRV_PLANEMATCHER pslam_adapter::ref::PlaneMatcher
#define DRV_OPTIMIZER pslam_adapter::ref::Optimizer
#include "match_driver.cc"
