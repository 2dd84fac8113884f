// DMMA m8n8k4 throughput on B200 as a function of warps per SM sub-partition and of the number
// of independent accumulator chains per warp: how much instruction-level parallelism a warp of
// the 20-state traversal needs to keep the FP64 tensor pipe busy with 3 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k(double* out, int iters, double x, double y){
  double c[2*CH];
  for (int j=0;j<2*CH;++j) c[j]=threadIdx.x*1e-3+j;
  double a = x + threadIdx.x*1e-9, b = y;
  for (int it=0; it<iters; ++it){
    #pragma unroll
    for (int r=0;r<60/CH;++r){
      #pragma unroll
      for (int j=0;j<CH;++j) dmma(c[2*j], c[2*j+1], a, b);
    }
  }
  double s=0; for (int j=0;j<2*CH;++j) s+=c[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template <int CH> void run(double* d, int warps){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=2000;
  k<CH><<<148,warps*32>>>(d,10,1e-3,1e-3);
  cudaEventRecord(e0);
  k<CH><<<148,warps*32>>>(d,iters,1e-3,1e-3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double n = 148.0*warps*iters*(60/CH)*CH;
  double flop = n*512.0;
  // cycles per DMMA per SMSP at 1.9 GHz nominal: report time per DMMA per scheduler in ns
  printf("warps/SM=%2d chains=%2d: %.3f ms  %.2f TFLOP/s  %.2f ns per DMMA per scheduler\n", warps, CH, ms, flop/ms/1e9,
         ms*1e6/(n/(148.0*4)));
}
int main(){
  double* d; cudaMalloc(&d, 148*1024*8);
  for (int w : {4,8,12,16,32}){
    run<1>(d,w); run<2>(d,w); run<3>(d,w); run<6>(d,w); run<12>(d,w);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
