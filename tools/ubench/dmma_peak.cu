// FP64 tensor-core (DMMA m8n8k4 via mma.sync) peak on B200.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) k(double* out, int iters, double x, double y){
  double c[16];
  for (int j=0;j<16;++j) c[j]=threadIdx.x*1e-3+j;
  double a = x + threadIdx.x*1e-9, b = y;
  for (int it=0; it<iters; ++it){
    #pragma unroll
    for (int r=0;r<8;++r){
      #pragma unroll
      for (int j=0;j<8;++j) dmma(c[2*j], c[2*j+1], a, b);
    }
  }
  double s=0; for (int j=0;j<16;++j) s+=c[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  double* d; cudaMalloc(&d, 148*8*256*8);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int bps : {1,2,4}){
    int iters=5000;
    k<<<148*bps,256>>>(d,100,1e-3,1e-3);
    cudaEventRecord(e0);
    k<<<148*bps,256>>>(d,iters,1e-3,1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flop = 148.0*bps*8*iters*64.0*(8*8*4*2.0);   // warps * iters * mma per iter * flop per mma
    printf("blocks/SM=%d: %.3f ms  %.2f TFLOP/s  %s\n", bps, ms, flop/ms/1e9, cudaGetErrorString(cudaGetLastError()));
  }
}
