#include <cstdio>
__device__ __forceinline__ unsigned long long pack(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__global__ void k(float* o, float x, float y, int iters){
  unsigned long long c[16]; for(int j=0;j<16;++j) c[j]=pack(threadIdx.x+j, j);
  unsigned long long a=pack(x,x), b=pack(y,y*0.5f);
  for(int i=0;i<iters;++i){
#pragma unroll
    for(int j=0;j<16;++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c[j]) : "l"(a), "l"(b));
  }
  float s=0; for(int j=0;j<16;++j){ float lo,hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(c[j])); s+=lo+hi;}
  o[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){ float* o; cudaMalloc(&o, 148*8*256*4); cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
 const int iters=20000; k<<<148*4,256>>>(o,1.0001f,0.5f,iters); cudaDeviceSynchronize(); cudaEventRecord(e0); k<<<148*4,256>>>(o,1.0001f,0.5f,iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1);
 double fl=(double)148*4*256*iters*16*2*2; printf("ffma2: %.1f TFLOP/s (%s)\n", fl/ms/1e9, cudaGetErrorString(cudaGetLastError())); }
