// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Frame::ComputeStereoFromRGBD  src/Frame.cc:603-621: per key point the depth under the (distorted) key point, truncated coordinates,
// and the virtual right coordinate u_un - bf / d.  depth is the float image of the Frame constructor (raw * depthMapFactor, :80-83).
#pragma once
#include <cstdint>
namespace oracle {
// keys / keys_un: [n][2] float (KeyPoint::pt of mvKeys / mvKeysUn)
void compute_stereo_from_rgbd(int n, const float* keys, const float* keys_un, const float* depth, int w, float bf, float* u_right, float* out_depth);
}  // namespace oracle
