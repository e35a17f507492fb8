#include <cstdio>
#include <cstdlib>
#include <cstdint>
__global__ void k(const uint8_t* img, int* out, int n) {
    __shared__ __align__(16) uint8_t win[68][72];
    for (int i = threadIdx.x; i < 68*72; i += blockDim.x) (&win[0][0])[i] = img[i];
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        int y = p / 66 + 3, x = p % 66 + 3;
        const int v = win[y][x];
        int d[16];
        d[0] = v - win[y + 3][x];   d[1] = v - win[y + 3][x + 1];  d[2] = v - win[y + 2][x + 2];  d[3] = v - win[y + 1][x + 3];
        d[4] = v - win[y][x + 3];   d[5] = v - win[y - 1][x + 3];  d[6] = v - win[y - 2][x + 2];  d[7] = v - win[y - 3][x + 1];
        d[8] = v - win[y - 3][x];   d[9] = v - win[y - 3][x - 1];  d[10] = v - win[y - 2][x - 2]; d[11] = v - win[y - 1][x - 3];
        d[12] = v - win[y][x - 3];  d[13] = v - win[y + 1][x - 3]; d[14] = v - win[y + 2][x - 2]; d[15] = v - win[y + 3][x - 1];
        int best = 0;
        #pragma unroll
        for (int i = 0; i < 16; ++i) {
            int m = d[i], M = d[i];
            #pragma unroll
            for (int k = 1; k < 9; ++k) { m = min(m, d[(i + k) & 15]); M = max(M, d[(i + k) & 15]); }
            best = max(best, max(m, -M));
            if (p == 1) printf("i=%d m=%d M=%d best=%d\n", i, m, M, best);
        }
        if (p == 1) { printf("d:"); for (int i = 0; i < 16; ++i) printf(" %d", d[i]); printf("\n"); }
        out[p] = best - 1;
    }
}
int main(){
    const int N=68*72; uint8_t* h=(uint8_t*)malloc(N); srand(3); for(int i=0;i<N;++i) h[i]=rand()&255;
    uint8_t* d; int* o; cudaMalloc(&d,N); cudaMalloc(&o,4096*4); cudaMemcpy(d,h,N,cudaMemcpyHostToDevice);
    k<<<1,128>>>(d,o,62*66); int ho[8]; cudaMemcpy(ho,o,8*4,cudaMemcpyDeviceToHost);
    printf("out: %d %d %d %d\n",ho[0],ho[1],ho[2],ho[3]);
}
