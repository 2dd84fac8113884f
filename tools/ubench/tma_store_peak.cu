// Developer micro-benchmark: HBM bandwidth of a write-only stream issued as TMA bulk stores from shared
// memory (cp.async.bulk.global.shared::cta), per copy size, next to 256-bit st.global from registers.
// Question behind it: would routing the CLV stores of k_traverse_dna through shared memory + TMA lift
// its write roofline (cudaMemsetAsync reaches 7.4 TB/s, kernel-issued stores 6.2-6.4 TB/s)?
//   nvcc -arch=sm_100a -O3 -o tma_store_peak tma_store_peak.cu && ./tma_store_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_st256(double * __restrict__ p, size_t n4, double v)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * i), "d"(v), "d"(v + 1),
                 "d"(v + 2), "d"(v + 3)
                 : "memory");
}

// every warp owns a `bytes` staging buffer in shared memory and stores it over and over to consecutive
// chunks of its CTA's region; at most `depth` copies of a warp in flight
template <int DEPTH>
__global__ void k_tma(unsigned char * __restrict__ out, size_t total, unsigned int bytes, int evict_first)
{
  extern __shared__ __align__(128) unsigned char sm[];
  const unsigned int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  unsigned char * mine = sm + (size_t)warp * bytes;
  for (unsigned int i = lane * 16; i < bytes; i += 32 * 16) *reinterpret_cast<uint4 *>(mine + i) = make_uint4(i, warp, 3, 4);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const size_t chunks = total / bytes;
  const size_t gw = (size_t)blockIdx.x * nw + warp, gstride = (size_t)gridDim.x * nw;
  if (lane == 0)
  {
    for (size_t c = gw; c < chunks; c += gstride)
    {
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(mine);
      if (evict_first)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(out + c * bytes),
                     "r"(src), "r"(bytes), "l"(pol)
                     : "memory");
      else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * bytes), "r"(src), "r"(bytes)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH) : "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main()
{
  const size_t total = (size_t)8 << 30;
  unsigned char * buf;
  cudaMalloc(&buf, total);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  auto time = [&](auto launch) {
    float best = 1e30f;
    for (int r = 0; r < 5; ++r)
    {
      cudaEventRecord(a);
      launch();
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    return total / (best * 1e-3) / 1e9;
  };
  printf("st.global 256-bit L1::no_allocate            %6.0f GB/s\n", time([&] { k_st256<<<148 * 8, 256>>>((double *)buf, total / 32, 1.0); }));
  const unsigned int sizes[] = {640, 1024, 4096, 5120, 10240, 16384};
  for (unsigned int bytes : sizes)
    for (int ef = 0; ef < 2; ++ef)
    {
      const int warps = 12;
      const size_t smem = (size_t)warps * bytes;
      cudaFuncSetAttribute(k_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const double g4 = time([&] { k_tma<4><<<148, warps * 32, smem>>>(buf, total, bytes, ef); });
      const double g1 = time([&] { k_tma<1><<<148, warps * 32, smem>>>(buf, total, bytes, ef); });
      printf("TMA bulk store %5u B per copy, 12 warps/SM%s  depth 4: %6.0f GB/s   depth 1: %6.0f GB/s\n", bytes,
             ef ? ", evict_first" : "             ", g4, g1);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
