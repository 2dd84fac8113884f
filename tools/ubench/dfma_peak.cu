// FP64 vector-pipe peak on B200: 16 independent DFMA chains per thread, register operands.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(double* out, int iters, double x, double y){
  double a[16];
  for (int j=0;j<16;++j) a[j]=threadIdx.x*1e-3+j;
  for (int it=0; it<iters; ++it){
    #pragma unroll
    for (int r=0;r<8;++r){
      #pragma unroll
      for (int j=0;j<16;++j) a[j]=__fma_rn(a[j],x,y);
    }
  }
  double s=0; for (int j=0;j<16;++j) s+=a[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  double* d; cudaMalloc(&d, 148*8*256*8);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int blocks_per_sm : {1,2,4,8}){
    int iters=20000;
    k<<<148*blocks_per_sm,256>>>(d,100,1.0000001,1e-9);
    cudaEventRecord(e0);
    k<<<148*blocks_per_sm,256>>>(d,iters,1.0000001,1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flop = 148.0*blocks_per_sm*256*iters*(128.0*2);
    printf("blocks/SM=%d: %.3f ms  %.2f TFLOP/s\n", blocks_per_sm, ms, flop/ms/1e9);
  }
}
