// Microbenchmark: 20x20 mat-vec products in FP64 where the matrix operand comes from the
// constant bank (kernel parameters / __constant__) as a direct DFMA operand and the rate
// index is warp-uniform and compile-time (switch over warp id).  Question: does the constant
// path sustain the FP64 pipe when 4 warps stream through 25.6 KB of constants?
#include <cstdio>
#include <cuda_runtime.h>
struct P { double m[2][4][400]; };
__constant__ P cP;
template<int K, bool PARAM>
__device__ __forceinline__ void body(const P& p, const double (&cl)[20], const double (&cr)[20], double (&out)[20]) {
  #pragma unroll
  for (int i=0;i<20;++i){
    double a0=0,a1=0,a2=0,a3=0,b0=0,b1=0,b2=0,b3=0;
    #pragma unroll
    for(int b=0;b<5;++b){
      a0=__fma_rn(p.m[0][K][i*20+4*b+0],cl[4*b+0],a0);
      a1=__fma_rn(p.m[0][K][i*20+4*b+1],cl[4*b+1],a1);
      a2=__fma_rn(p.m[0][K][i*20+4*b+2],cl[4*b+2],a2);
      a3=__fma_rn(p.m[0][K][i*20+4*b+3],cl[4*b+3],a3);
      b0=__fma_rn(p.m[1][K][i*20+4*b+0],cr[4*b+0],b0);
      b1=__fma_rn(p.m[1][K][i*20+4*b+1],cr[4*b+1],b1);
      b2=__fma_rn(p.m[1][K][i*20+4*b+2],cr[4*b+2],b2);
      b3=__fma_rn(p.m[1][K][i*20+4*b+3],cr[4*b+3],b3);
    }
    out[i]=__dmul_rn(__dadd_rn(__dadd_rn(a0,a1),__dadd_rn(a2,a3)), __dadd_rn(__dadd_rn(b0,b1),__dadd_rn(b2,b3)));
  }
}
template<bool PARAM>
__global__ void __launch_bounds__(128,2) k(const __grid_constant__ P pp, double* out, int iters){
  const P& p = PARAM ? pp : cP;
  int w = threadIdx.x>>5;
  double cl[20], cr[20], o[20];
  for (int j=0;j<20;++j){cl[j]=1e-3*(threadIdx.x+j); cr[j]=2e-3*(j+1);}
  double acc=0;
  for (int it=0; it<iters; ++it){
    switch(w&3){case 0: body<0,PARAM>(p,cl,cr,o);break;case 1: body<1,PARAM>(p,cl,cr,o);break;case 2: body<2,PARAM>(p,cl,cr,o);break;default: body<3,PARAM>(p,cl,cr,o);}
    for (int j=0;j<20;++j){ acc+=o[j]; cl[j]=o[j]*1e-3+cr[j]; }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}
int main(){
  P* h = new P; for (int a=0;a<2;++a) for(int k=0;k<4;++k) for(int i=0;i<400;++i) h->m[a][k][i]=1.0/(1+i+k);
  cudaMemcpyToSymbol(cP, h, sizeof(P));
  double* d; cudaMalloc(&d, 296*128*8);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode=0; mode<2; ++mode){
    for (int rep=0; rep<2; ++rep){
      int iters=2000;
      cudaEventRecord(e0);
      if (mode==0) k<true><<<296,128>>>(*h,d,iters); else k<false><<<296,128>>>(*h,d,iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms,e0,e1);
      double flop = 296.0*128*iters*(1600.0*2);
      printf("%s: %.3f ms  %.2f TFLOP/s (FMA=2)  err=%s\n", mode==0?"param":"__constant__", ms, flop/ms/1e9, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
