// Minimal TMA probe (debug harness, not product code): loads one box of a u8 tensor into shared memory and checks it against the expected contents
// (zero fill outside the tensor).  usage: tma_min <rank 2|3> <box_w> <box_h> <x> <y> <src: 0 param | 1 global>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../planarslam_b200/csrc/tma_util.cuh"
using namespace pslam;

template <int RANK>
__global__ void k_probe(const __grid_constant__ CUtensorMap pm, const CUtensorMap* gm, int use_global, int box_bytes, int x, int y, int z, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, box_bytes);
        const CUtensorMap* m = use_global ? gm : &pm;
        if (RANK == 3) tma_load_3d(tile, m, x, y, z, &bar);
        else asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(tile)), "l"(m),
                          "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < box_bytes; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
    const int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]), use_global = atoi(argv[6]);
    const int W = 640, H = 480, N = 2, z = 1;
    std::vector<uint8_t> img((size_t)W * H * N);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 24);
    uint8_t *d_img, *d_out; CUtensorMap* d_map;
    cudaMalloc(&d_img, img.size()); cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&d_out, bw * bh); cudaMalloc(&d_map, sizeof(CUtensorMap));
    CUtensorMap m;
    bool ok;
    if (rank == 3) ok = tma_encode_3d(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d_img, W, H, N, W, (size_t)W * H, bw, bh);
    else {
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H * N}, strides[1] = {(cuuint64_t)W};
        const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
        ok = ((encode_fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (!ok) { printf("encode failed\n"); return 2; }
    cudaMemcpy(d_map, &m, sizeof m, cudaMemcpyHostToDevice);
    if (rank == 3) k_probe<3><<<1, 128, bw * bh>>>(m, d_map, use_global, bw * bh, x, y, z, d_out);
    else k_probe<2><<<1, 128, bw * bh>>>(m, d_map, use_global, bw * bh, x, y + z * H, 0, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("FAULT %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint8_t> out(bw * bh);
    cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < bh; ++r)
        for (int c = 0; c < bw; ++c) {
            const int X = x + c, Y = y + r;
            uint8_t exp = 0;
            if (X >= 0 && X < W && Y >= 0 && (rank == 3 ? Y < H : Y + z * H < H * N) && (rank == 3 || Y + z * H >= 0)) exp = img[((size_t)z * H + Y) * W + X];
            bad += out[r * bw + c] != exp;
        }
    printf(bad ? "MISMATCH %d\n" : "OK\n", bad);
    return bad ? 3 : 0;
}
