// Packed keypoint word and per-frame status flags shared by the ORB kernels and the ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace pslam {

// packed candidate / keypoint: x (11 bits) | y (11 bits) << 11 | score (8 bits) << 22, coordinates relative to (16,16)
__host__ __device__ inline uint32_t pack_kp(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 11) | ((uint32_t)s << 22); }
__host__ __device__ inline int kp_x(uint32_t p) { return p & 2047; }
__host__ __device__ inline int kp_y(uint32_t p) { return (p >> 11) & 2047; }
__host__ __device__ inline int kp_s(uint32_t p) { return (p >> 22) & 255; }

enum { ST_SLOT_OVERFLOW = 1, ST_CAND_OVERFLOW = 2, ST_NODE_OVERFLOW = 4, ST_OUT_OVERFLOW = 8 };

}  // namespace pslam
