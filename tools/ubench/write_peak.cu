// Developer micro-benchmark: HBM bandwidth of a WRITE-ONLY stream on this GPU (the fused traversal
// kernel's traffic is 98 % writes, so this - not the copy figure - is its roofline) next to a
// read-only stream and a copy.   nvcc -arch=sm_100a -O3 -o write_peak write_peak.cu && ./write_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_fill(double4 * __restrict__ p, size_t n, double v)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    p[i] = make_double4(v, v + 1, v + 2, v + 3);
}
// 128-bit stores, a warp instruction covers 512 contiguous bytes
__global__ void k_fill128(double2 * __restrict__ p, size_t n, double v)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = make_double2(v, v + 1);
}
// 128-bit stores at a 32-byte lane stride: two instructions per 1 KB, each touching half sectors
__global__ void k_fill128_strided(double2 * __restrict__ p, size_t n, double v)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; 2 * i + 1 < n; i += stride)
  {
    p[2 * i] = make_double2(v, v + 1);
    p[2 * i + 1] = make_double2(v + 2, v + 3);
  }
}
// 256-bit stores with the L1::no_allocate hint (what the CLV kernels use)
__global__ void k_fill256_na(double * __restrict__ p, size_t n4, double v)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * i), "d"(v), "d"(v + 1),
                 "d"(v + 2), "d"(v + 3)
                 : "memory");
}
#define FILL_VARIANT(NAME, PTX, WIDTH)                                                            \
  __global__ void NAME(double * __restrict__ p, size_t nvec, double v)                            \
  {                                                                                               \
    const size_t stride = (size_t)gridDim.x * blockDim.x;                                         \
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride)         \
    {                                                                                             \
      if (WIDTH == 2)                                                                             \
        asm volatile(PTX " [%0], {%1,%2};" ::"l"(p + 2 * i), "d"(v), "d"(v + 1) : "memory");     \
      else                                                                                        \
        asm volatile(PTX " [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * i), "d"(v), "d"(v + 1), "d"(v + 2), \
                     "d"(v + 3)                                                                   \
                     : "memory");                                                                 \
    }                                                                                             \
  }
FILL_VARIANT(k_f128_na, "st.global.L1::no_allocate.v2.f64", 2)
FILL_VARIANT(k_f128_cs, "st.global.cs.v2.f64", 2)
FILL_VARIANT(k_f256_cs, "st.global.cs.v4.f64", 4)
FILL_VARIANT(k_f128_wt, "st.global.wt.v2.f64", 2)
FILL_VARIANT(k_f256_wt, "st.global.wt.v4.f64", 4)
FILL_VARIANT(k_f256_na_ef, "st.global.L1::no_allocate.L2::evict_first.v4.f64", 4)

// block-contiguous: every CTA writes its own contiguous 64 KB chunks (256-bit stores)
__global__ void k_fill_chunks(double4 * __restrict__ p, size_t n, double v)
{
  const size_t chunk = 2048; /* double4 per chunk = 64 KB */
  for (size_t c = blockIdx.x; c * chunk < n; c += gridDim.x)
    for (size_t i = threadIdx.x; i < chunk && c * chunk + i < n; i += blockDim.x)
      p[c * chunk + i] = make_double4(v, v + 1, v + 2, v + 3);
}
__global__ void k_read(const double4 * __restrict__ p, size_t n, double * out)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    const double4 v = p[i];
    s += v.x + v.y + v.z + v.w;
  }
  if (s == 12345.678) *out = s;
}
__global__ void k_copy(const double4 * __restrict__ a, double4 * __restrict__ b, size_t n)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = a[i];
}

int main()
{
  const size_t bytes = (size_t)8 << 30, n = bytes / sizeof(double4);
  double4 * a, * b;
  double * out;
  cudaMalloc(&a, bytes);
  cudaMalloc(&b, bytes);
  cudaMalloc(&out, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 8, block = 256;
  for (int which = 0; which < 14; ++which)
  {
    float best = 1e30f;
    for (int rep = 0; rep < 8; ++rep)
    {
      cudaEventRecord(e0);
      if (which == 0) k_fill<<<grid, block>>>(a, n, 1.0 + rep);
      else if (which == 1) cudaMemsetAsync(a, rep, bytes);
      else if (which == 2) k_read<<<grid, block>>>(a, n, out);
      else if (which == 3) k_copy<<<grid, block>>>(a, b, n);
      else if (which == 4) k_fill128<<<grid, block>>>((double2 *)a, 2 * n, 1.0 + rep);
      else if (which == 5) k_fill128_strided<<<grid, block>>>((double2 *)a, 2 * n, 1.0 + rep);
      else if (which == 6) k_fill256_na<<<grid, block>>>((double *)a, n, 1.0 + rep);
      else if (which == 7) k_fill_chunks<<<grid, block>>>(a, n, 1.0 + rep);
      else if (which == 8) k_f128_na<<<grid, block>>>((double *)a, 2 * n, 1.0 + rep);
      else if (which == 9) k_f128_cs<<<grid, block>>>((double *)a, 2 * n, 1.0 + rep);
      else if (which == 10) k_f256_cs<<<grid, block>>>((double *)a, n, 1.0 + rep);
      else if (which == 11) k_f128_wt<<<grid, block>>>((double *)a, 2 * n, 1.0 + rep);
      else if (which == 12) k_f256_wt<<<grid, block>>>((double *)a, n, 1.0 + rep);
      else k_f256_na_ef<<<grid, block>>>((double *)a, n, 1.0 + rep);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const double moved = (which == 3 ? 2.0 : 1.0) * bytes;
    printf("%-28s %8.1f GB/s\n", which == 0 ? "write-only (256-bit stores)" : which == 1 ? "cudaMemsetAsync" : which == 2 ? "read-only" : which == 3 ? "copy (read + write)" : which == 4 ? "write 128-bit coalesced" : which == 5 ? "write 128-bit, 32 B lane stride" : which == 6 ? "write 256-bit L1::no_allocate" : which == 7 ? "write 256-bit, 64 KB per CTA" : which == 8 ? "write 128-bit L1::no_allocate" : which == 9 ? "write 128-bit .cs" : which == 10 ? "write 256-bit .cs" : which == 11 ? "write 128-bit .wt" : which == 12 ? "write 256-bit .wt" : "write 256-bit na + L2::evict_first",
           moved / best / 1e6);
  }
  return 0;
}
