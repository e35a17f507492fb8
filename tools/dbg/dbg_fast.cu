// debug: run fast_score on the device over one window and compare with a host evaluation of the same code
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include "../../planarslam_b200/csrc/orb_kernels.cuh"
using namespace pslam;

__device__ int score_brute(const uint8_t (*win)[FAST_WIN_MAX + 4], int x, int y) {
    const int C[16][2] = {{0,3},{1,3},{2,2},{3,1},{3,0},{3,-1},{2,-2},{1,-3},{0,-3},{-1,-3},{-2,-2},{-3,-1},{-3,0},{-3,1},{-2,2},{-1,3}};
    int d[16], v = win[y][x];
    for (int k=0;k<16;++k) d[k]=v-win[y+C[k][1]][x+C[k][0]];
    int best=0;
    for (int s=0;s<16;++s){int mn=255,mx=-255;for(int k=0;k<9;++k){int dv=d[(s+k)&15];mn=min(mn,dv);mx=max(mx,dv);}best=max(best,max(mn,-mx));}
    int sc=best-1; return sc>=7?sc:0;
}
#define IMIN(a,b) ((a)<(b)?(a):(b))
#define IMAX(a,b) ((a)>(b)?(a):(b))
__device__ int score_tern(const uint8_t (*win)[FAST_WIN_MAX + 4], int x, int y) {
    const int v = win[y][x];
    int d[16];
    d[0] = v - win[y + 3][x];   d[1] = v - win[y + 3][x + 1];  d[2] = v - win[y + 2][x + 2];  d[3] = v - win[y + 1][x + 3];
    d[4] = v - win[y][x + 3];   d[5] = v - win[y - 1][x + 3];  d[6] = v - win[y - 2][x + 2];  d[7] = v - win[y - 3][x + 1];
    d[8] = v - win[y - 3][x];   d[9] = v - win[y - 3][x - 1];  d[10] = v - win[y - 2][x - 2]; d[11] = v - win[y - 1][x - 3];
    d[12] = v - win[y][x - 3];  d[13] = v - win[y + 1][x - 3]; d[14] = v - win[y + 2][x - 2]; d[15] = v - win[y + 3][x - 1];
    int mn[16], mx[16], mn4[16], mx4[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { mn[i] = IMIN(d[i], d[(i + 1) & 15]); mx[i] = IMAX(d[i], d[(i + 1) & 15]); }
#pragma unroll
    for (int i = 0; i < 16; ++i) { mn4[i] = IMIN(mn[i], mn[(i + 2) & 15]); mx4[i] = IMAX(mx[i], mx[(i + 2) & 15]); }
    int best = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int m8 = IMIN(mn4[i], mn4[(i + 4) & 15]), M8 = IMAX(mx4[i], mx4[(i + 4) & 15]);
        const int m9 = IMIN(m8, d[(i + 8) & 15]), M9 = IMAX(M8, d[(i + 8) & 15]);
        const int t = IMAX(m9, -M9);
        best = IMAX(best, t);
    }
    const int s = best - 1;
    return s >= 7 ? s : 0;
}
__global__ void k_dbg(const uint8_t* img, int* out) {
    __shared__ __align__(16) uint8_t win[FAST_WIN_MAX][FAST_WIN_MAX + 4];
    for (int i = threadIdx.x; i < FAST_WIN_MAX * (FAST_WIN_MAX + 4); i += blockDim.x) (&win[0][0])[i] = img[i];
    __syncthreads();
    for (int p = threadIdx.x; p < 62 * 66; p += blockDim.x) {
        int y = p / 66 + 3, x = p % 66 + 3;
        out[p] = fast_score(win, x, y, 7); out[p+62*66] = score_brute(win,x,y); out[p+2*62*66]=score_tern(win,x,y);
    }
}
static int host_score(const uint8_t* w, int x, int y) {
    static const int C[16][2] = {{0,3},{1,3},{2,2},{3,1},{3,0},{3,-1},{2,-2},{1,-3},{0,-3},{-1,-3},{-2,-2},{-3,-1},{-3,0},{-3,1},{-2,2},{-1,3}};
    int d[16], v = w[y*72+x];
    for (int k=0;k<16;++k) d[k]=v-w[(y+C[k][1])*72+x+C[k][0]];
    int best=0;
    for (int s=0;s<16;++s){int mn=255,mx=-255;for(int k=0;k<9;++k){int dv=d[(s+k)&15];mn=std::min(mn,dv);mx=std::max(mx,dv);}best=std::max(best,std::max(mn,-mx));}
    int sc=best-1; return sc>=7?sc:0;
}
int main(){
    const int N=FAST_WIN_MAX*(FAST_WIN_MAX+4);
    uint8_t* h=(uint8_t*)malloc(N); srand(3); for(int i=0;i<N;++i) h[i]=rand()&255;
    uint8_t* d; int* o; cudaMalloc(&d,N); cudaMalloc(&o,3*62*66*4); cudaMemcpy(d,h,N,cudaMemcpyHostToDevice);
    k_dbg<<<1,128>>>(d,o); int* ho=(int*)malloc(3*62*66*4); cudaError_t e=cudaMemcpy(ho,o,3*62*66*4,cudaMemcpyDeviceToHost);
    printf("cuda: %s\n", cudaGetErrorString(e));
    int bad=0,nz=0; for(int p=0;p<62*66;++p){int y=p/66+3,x=p%66+3; int hs=host_score(h,x,y); if(hs!=ho[p]){ if(bad<5)printf("(%d,%d) dev %d host %d\n",x,y,ho[p],hs); ++bad;} nz+=hs!=0;}
    printf("bad %d nz %d\n",bad,nz);
    for(int v=1;v<3;++v){int b2=0; for(int p=0;p<62*66;++p){int y=p/66+3,x=p%66+3; if(host_score(h,x,y)!=ho[p+v*62*66]) ++b2;} printf("variant %d bad %d\n",v,b2);}
}
