// Host build of planarslam_b200/csrc/manhattan_body.h (the code k_track_manhattan runs, one thread per frame) for tests/test_manhattan_host.py.
#include <cstdint>

#include "manhattan_body.h"

extern "C" void host_track_manhattan(const float* R_last, const float* normals, int n, const double* dirs, int m, MhResult* res, uint8_t* nmask, uint8_t* dmask) {
    mh_track(R_last, normals, n, dirs, m, *res, nmask, dmask);
}
