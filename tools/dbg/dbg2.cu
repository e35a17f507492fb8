#include <cstdio>
#include <cstdlib>
#include <cstdint>
__global__ void k(const uint8_t* img, int* out) {
    __shared__ __align__(16) uint8_t win[68][72];
    for (int i = threadIdx.x; i < 68*72; i += blockDim.x) (&win[0][0])[i] = img[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        int x=4,y=3;
        out[0]=win[y][x]; out[1]=win[y+3][x]; out[2]=win[y+3][x+1]; out[3]=win[y+2][x+2]; out[4]=win[y-3][x]; out[5]=win[y][x-3];
        const uint8_t (*w2)[72] = win;
        out[6]=w2[y][x]; out[7]=w2[y+3][x]; out[8]=w2[y-3][x];
        out[9]=img[y*72+x]; out[10]=img[(y+3)*72+x];
    }
}
int main(){
    const int N=68*72; uint8_t* h=(uint8_t*)malloc(N); srand(3); for(int i=0;i<N;++i) h[i]=rand()&255;
    uint8_t* d; int* o; cudaMalloc(&d,N); cudaMalloc(&o,64*4); cudaMemcpy(d,h,N,cudaMemcpyHostToDevice);
    k<<<1,128>>>(d,o); int ho[64]; cudaMemcpy(ho,o,64*4,cudaMemcpyDeviceToHost);
    int x=4,y=3;
    printf("dev: "); for(int i=0;i<11;++i) printf("%d ",ho[i]); printf("\n");
    printf("host: %d %d %d %d %d %d\n", h[y*72+x], h[(y+3)*72+x], h[(y+3)*72+x+1], h[(y+2)*72+x+2], h[(y-3)*72+x], h[y*72+x-3]);
}
