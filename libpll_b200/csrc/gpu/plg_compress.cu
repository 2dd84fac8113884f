/*
 * plg_compress.cu - site-pattern compression on the device: the step right
 * before the likelihood path.  Same output as reference src/compress.c:138-286 - unique
 * alignment columns in the reference's sorted order (columns compared as 0-terminated strings
 * of SIGNED encoded bytes, :33-81) with their multiplicities.
 *
 * The reference sorts column strings with a randomised multikey quicksort.  The order of
 * distinct columns is total, so any correct sort gives the same bytes out; here:
 *   1. rows are uploaded as they are ([taxon][site]); one kernel encodes them through the
 *      256-entry code table and transposes to column-major 64-bit words ([site][taxon/8]),
 *      first taxon in the most significant byte, every byte XOR 0x80 so that unsigned word
 *      order == the reference's signed byte order;
 *   2. (only if some column holds a 0 code) bytes after the first 0 of a column are cleared:
 *      the reference stops comparing there;
 *   3. least-significant-word-first radix sort of the column permutation: one stable
 *      cub::DeviceRadixSort::SortPairs on 64-bit keys per 8 taxa (library sort; the kernels
 *      around it are ours);
 *   4. adjacent columns are compared, an exclusive scan numbers the unique ones, run lengths
 *      are the weights;
 *   5. unique columns are decoded through the inverse table and transposed back into the
 *      caller's rows.
 * Cost at 1 000 taxa x 1 M columns: 1 GB up, 125 sorts of 1 M pairs, <= 1 GB down.
 */
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "plg_internal.cuh"

struct ByteTable
{
  unsigned char v[256];
};

/* rows [taxon][site] -> words [site][taxon / 8]; a thread makes one word */
__global__ void k_encode_transpose(const unsigned char * __restrict__ rows, size_t row_pitch,
                                   unsigned int taxa, unsigned int sites, unsigned int words,
                                   unsigned long long * __restrict__ cols, const ByteTable tab,
                                   unsigned int * __restrict__ has_zero)
{
  const unsigned int site = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int w = blockIdx.y;
  if (site >= sites) return;
  unsigned long long word = 0;
  bool zero = false;
#pragma unroll
  for (unsigned int b = 0; b < 8; ++b)
  {
    const unsigned int t = w * 8 + b;
    unsigned int code = 0;
    if (t < taxa)
    {
      code = tab.v[rows[(size_t)t * row_pitch + site]];
      zero |= (code == 0);
    }
    /* padding taxa encode as 0 ^ 0x80 too: equal in every column, so they never decide */
    word = (word << 8) | (unsigned long long)((code ^ 0x80u) & 0xffu);
  }
  cols[(size_t)site * words + w] = word;
  if (zero) atomicOr(has_zero, 1u);
}

/* the reference compares 0-terminated strings: everything after a column's first 0 code is
 * irrelevant; clear it so that whole-word comparisons agree */
__global__ void k_truncate_at_zero(unsigned long long * __restrict__ cols, unsigned int taxa,
                                   unsigned int sites, unsigned int words)
{
  const unsigned int site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= sites) return;
  unsigned long long * c = cols + (size_t)site * words;
  bool ended = false;
  for (unsigned int w = 0; w < words; ++w)
  {
    unsigned long long word = c[w];
    if (ended)
    {
      c[w] = 0x8080808080808080ull;
      continue;
    }
    for (unsigned int b = 0; b < 8 && w * 8 + b < taxa; ++b)
    {
      const unsigned int shift = 56 - 8 * b;
      if (((word >> shift) & 0xffu) == 0x80u)
      {
        /* keep the 0 itself, clear the rest of this word */
        const unsigned long long keep = shift ? (~0ull << shift) : ~0ull;
        word = (word & keep) | (0x8080808080808080ull & ~keep);
        ended = true;
        break;
      }
    }
    c[w] = word;
  }
}

__global__ void k_iota(unsigned int * __restrict__ p, unsigned int n)
{
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

__global__ void k_gather_keys(const unsigned long long * __restrict__ cols, unsigned int words,
                              unsigned int w, const unsigned int * __restrict__ perm,
                              unsigned long long * __restrict__ keys, unsigned int n)
{
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = cols[(size_t)perm[i] * words + w];
}

/* flag[i] = 1 if sorted column i differs from sorted column i-1 */
__global__ void k_flag_new(const unsigned long long * __restrict__ cols, unsigned int words,
                           const unsigned int * __restrict__ perm, unsigned int * __restrict__ flag,
                           unsigned int n)
{
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned int differs = (i == 0);
  if (i)
  {
    const unsigned long long * a = cols + (size_t)perm[i] * words;
    const unsigned long long * b = cols + (size_t)perm[i - 1] * words;
    for (unsigned int w = 0; w < words && !differs; ++w) differs = (a[w] != b[w]);
  }
  flag[i] = differs;
}

/* first[u] = sorted position of the first copy of unique column u; first[unique] = n */
__global__ void k_first_positions(const unsigned int * __restrict__ flag, const unsigned int * __restrict__ rank,
                                  unsigned int * __restrict__ first, unsigned int n, unsigned int unique)
{
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) first[rank[i]] = i;
  if (i == 0) first[unique] = n;
}

__global__ void k_weights(const unsigned int * __restrict__ first, unsigned int * __restrict__ weights,
                          unsigned int unique)
{
  const unsigned int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < unique) weights[u] = first[u + 1] - first[u];
}

/* unique columns -> rows [taxon][unique], decoded */
__global__ void k_decode_transpose(const unsigned long long * __restrict__ cols, unsigned int words,
                                   const unsigned int * __restrict__ perm, const unsigned int * __restrict__ first,
                                   unsigned int taxa, unsigned int unique, unsigned char * __restrict__ rows,
                                   size_t row_pitch, const ByteTable inv)
{
  const unsigned int u = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int w = blockIdx.y;
  if (u >= unique) return;
  const unsigned long long word = cols[(size_t)perm[first[u]] * words + w];
#pragma unroll
  for (unsigned int b = 0; b < 8; ++b)
  {
    const unsigned int t = w * 8 + b;
    if (t < taxa)
      rows[(size_t)t * row_pitch + u] = inv.v[(unsigned int)((word >> (56 - 8 * b)) & 0xffu) ^ 0x80u];
  }
}

#define CMP_CUDA(call)                                                                    \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      plg_set_error("plg_compress_patterns: %s failed: %s", #call, cudaGetErrorString(e_)); \
      rc = PLG_E_CUDA;                                                                    \
      goto done;                                                                          \
    }                                                                                     \
  } while (0)

extern "C" int plg_compress_patterns(int device, unsigned char * const * rows, unsigned int taxa,
                                     size_t length, const unsigned char * code_table,
                                     const unsigned char * inverse_table, unsigned int * weights_out,
                                     size_t * unique_out)
{
  int rc = PLG_OK;
  if (!rows || !taxa || !length || !code_table || !inverse_table || !weights_out || !unique_out ||
      length > 0xfffffff0ull)
  {
    plg_set_error("plg_compress_patterns: invalid argument");
    return PLG_E_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
  {
    cudaGetLastError();
    plg_set_error("plg_compress_patterns: no CUDA device visible (this backend has no CPU fallback)");
    return PLG_E_NODEVICE;
  }
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess)
  {
    plg_set_error("plg_compress_patterns: cannot select device %d", device);
    return PLG_E_NODEVICE;
  }

  const unsigned int n = (unsigned int)length;
  const unsigned int words = (taxa + 7) / 8;
  const size_t pitch = (length + 255) / 256 * 256;
  unsigned char * d_rows = NULL;
  unsigned long long * d_cols = NULL, * d_keys[2] = {NULL, NULL};
  unsigned int * d_perm[2] = {NULL, NULL}, * d_flag = NULL, * d_rank = NULL, * d_first = NULL,
               * d_weights = NULL, * d_zero = NULL;
  void * d_temp = NULL;
  size_t temp_bytes = 0, scan_bytes = 0;
  cudaStream_t st = NULL;
  ByteTable enc, dec;
  memcpy(enc.v, code_table, 256);
  memcpy(dec.v, inverse_table, 256);
  const unsigned int T = 256;
  const unsigned int nb = (n + T - 1) / T;
  unsigned int h_zero = 0, unique = 0, last_flag = 0, last_rank = 0;
  int cur = 0;

  CMP_CUDA(cudaStreamCreate(&st));
  CMP_CUDA(cudaMalloc(&d_rows, pitch * taxa));
  CMP_CUDA(cudaMalloc(&d_cols, (size_t)n * words * sizeof(unsigned long long)));
  CMP_CUDA(cudaMalloc(&d_keys[0], (size_t)n * sizeof(unsigned long long)));
  CMP_CUDA(cudaMalloc(&d_keys[1], (size_t)n * sizeof(unsigned long long)));
  CMP_CUDA(cudaMalloc(&d_perm[0], (size_t)n * sizeof(unsigned int)));
  CMP_CUDA(cudaMalloc(&d_perm[1], (size_t)n * sizeof(unsigned int)));
  CMP_CUDA(cudaMalloc(&d_flag, (size_t)n * sizeof(unsigned int)));
  CMP_CUDA(cudaMalloc(&d_rank, (size_t)n * sizeof(unsigned int)));
  CMP_CUDA(cudaMalloc(&d_first, ((size_t)n + 1) * sizeof(unsigned int)));
  CMP_CUDA(cudaMalloc(&d_weights, (size_t)n * sizeof(unsigned int)));
  CMP_CUDA(cudaMalloc(&d_zero, sizeof(unsigned int)));
  CMP_CUDA(cudaMemsetAsync(d_zero, 0, sizeof(unsigned int), st));
  CMP_CUDA(cub::DeviceRadixSort::SortPairs(NULL, temp_bytes, d_keys[0], d_keys[1], d_perm[0], d_perm[1], (int)n,
                                           0, 64, st));
  CMP_CUDA(cub::DeviceScan::ExclusiveSum(NULL, scan_bytes, d_flag, d_rank, (int)n, st));
  if (scan_bytes > temp_bytes) temp_bytes = scan_bytes;
  CMP_CUDA(cudaMalloc(&d_temp, temp_bytes));

  /* 1. upload + encode + transpose */
  for (unsigned int t = 0; t < taxa; ++t)
    CMP_CUDA(cudaMemcpyAsync(d_rows + (size_t)t * pitch, rows[t], length, cudaMemcpyHostToDevice, st));
  {
    dim3 grid(nb, words);
    k_encode_transpose<<<grid, T, 0, st>>>(d_rows, pitch, taxa, n, words, d_cols, enc, d_zero);
  }
  CMP_CUDA(cudaMemcpyAsync(&h_zero, d_zero, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CMP_CUDA(cudaStreamSynchronize(st));
  /* 2. */
  if (h_zero) k_truncate_at_zero<<<nb, T, 0, st>>>(d_cols, taxa, n, words);

  /* 3. LSD radix sort of the permutation, 8 taxa per pass */
  k_iota<<<nb, T, 0, st>>>(d_perm[0], n);
  for (unsigned int w = words; w-- > 0;)
  {
    k_gather_keys<<<nb, T, 0, st>>>(d_cols, words, w, d_perm[cur], d_keys[0], n);
    CMP_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_keys[0], d_keys[1], d_perm[cur],
                                             d_perm[cur ^ 1], (int)n, 0, 64, st));
    cur ^= 1;
  }

  /* 4. unique columns and their multiplicities */
  k_flag_new<<<nb, T, 0, st>>>(d_cols, words, d_perm[cur], d_flag, n);
  CMP_CUDA(cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_flag, d_rank, (int)n, st));
  CMP_CUDA(cudaMemcpyAsync(&last_flag, d_flag + (n - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CMP_CUDA(cudaMemcpyAsync(&last_rank, d_rank + (n - 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CMP_CUDA(cudaStreamSynchronize(st));
  unique = last_rank + last_flag;
  k_first_positions<<<nb, T, 0, st>>>(d_flag, d_rank, d_first, n, unique);
  k_weights<<<(unique + T - 1) / T, T, 0, st>>>(d_first, d_weights, unique);

  /* 5. decode the unique columns back into rows and bring them home */
  {
    dim3 grid((unique + T - 1) / T, words);
    k_decode_transpose<<<grid, T, 0, st>>>(d_cols, words, d_perm[cur], d_first, taxa, unique, d_rows, pitch, dec);
  }
  CMP_CUDA(cudaGetLastError());
  for (unsigned int t = 0; t < taxa; ++t)
    CMP_CUDA(cudaMemcpyAsync(rows[t], d_rows + (size_t)t * pitch, unique, cudaMemcpyDeviceToHost, st));
  CMP_CUDA(cudaMemcpyAsync(weights_out, d_weights, (size_t)unique * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CMP_CUDA(cudaStreamSynchronize(st));
  *unique_out = unique;

done:
  cudaFree(d_rows); cudaFree(d_cols); cudaFree(d_keys[0]); cudaFree(d_keys[1]);
  cudaFree(d_perm[0]); cudaFree(d_perm[1]); cudaFree(d_flag); cudaFree(d_rank); cudaFree(d_first);
  cudaFree(d_weights); cudaFree(d_zero); cudaFree(d_temp);
  if (st) cudaStreamDestroy(st);
  return rc;
}
