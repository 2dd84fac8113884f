/*
 * plg_pmatrix.cu - transition-probability matrices from the eigendecomposition:
 *   P(t) = I + V^-1 . diag(expm1(lambda_j * r_n * t [/ (1 - pinv)])) . V   per (branch, rate)
 *
 * Replaces pll_core_update_pmatrix and its SIMD rungs:
 *   generic      reference src/core_pmatrix.c:146-250
 *   4x4 AVX      reference src/core_pmatrix_avx.c:42-310   (what the AVX2 flag runs for DNA)
 *   20x20 AVX2   reference src/core_pmatrix_avx2.c:37-284
 *
 * The whole batch of (branch, rate) pairs is one launch.  Operation order follows the
 * reference kernels: x = (lambda*r)*t, optional division by
 * (1-pinv) when pinv > 1e-8, T = V^-1 scaled column-wise by expm1(x), P = T.V summed as
 * (a0+a1)+(a2+a3) [DNA, unfused] or with four FMA lane accumulators over five column
 * blocks [20 states], then + identity.  A zero branch length yields the identity.
 * expm1 is CUDA's (<= 1 ulp from glibc's): P entries agree with the reference to ~1e-16.
 */
#include "plg_internal.cuh"

struct PmatModel
{
  const double * eigenvals;     /* [R][Kp]     */
  const double * eigenvecs;     /* [R][K][Kp]  */
  const double * inv_eigenvecs; /* [R][K][Kp]  */
  const double * rates;         /* [R]         */
  const double * prop_invar;    /* [R]         */
};

__device__ __forceinline__ double pmat_exponent(double eval, double rate, double t, double pinv)
{
  double x = __dmul_rn(__dmul_rn(eval, rate), t);
  if (pinv > PLL_MISC_EPSILON) x = __ddiv_rn(x, __dsub_rn(1.0, pinv));
  return expm1(x);
}

/* DNA: one thread per (branch, rate, row j) */
__global__ void k_pmatrix_dna(double * __restrict__ pmatrix, size_t pmat_len,
                              const unsigned int * __restrict__ matrix_indices,
                              const double * __restrict__ branch_lengths, unsigned int count,
                              unsigned int rate_cats, PmatModel m)
{
  const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int j = tid & 3u;
  const unsigned int n = (tid >> 2) % rate_cats;
  const unsigned int i = tid / (4u * rate_cats);
  if (i >= count) return;

  const double t = branch_lengths[i];
  double * row = pmatrix + (size_t)matrix_indices[i] * pmat_len + n * 16 + j * 4;
  if (t == 0.0)
  {
    row[0] = (j == 0);
    row[1] = (j == 1);
    row[2] = (j == 2);
    row[3] = (j == 3);
    return;
  }
  const double pinv = m.prop_invar[n];
  const double rate = m.rates[n];
  const double * ev = m.eigenvals + n * 4;
  const double * V = m.eigenvecs + n * 16;
  const double * iV = m.inv_eigenvecs + n * 16 + j * 4;

  double T[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) T[q] = __dmul_rn(iV[q], pmat_exponent(ev[q], rate, t, pinv));
#pragma unroll
  for (int c = 0; c < 4; ++c)
  {
    const double s = hsum4(__dmul_rn(T[0], V[0 + c]), __dmul_rn(T[1], V[4 + c]),
                           __dmul_rn(T[2], V[8 + c]), __dmul_rn(T[3], V[12 + c]));
    row[c] = __dadd_rn(s, (c == (int)j) ? 1.0 : 0.0);
  }
}

/* 20 states: one block per (branch, rate) */
__global__ void __launch_bounds__(256)
k_pmatrix_aa(double * __restrict__ pmatrix, size_t pmat_len,
             const unsigned int * __restrict__ matrix_indices,
             const double * __restrict__ branch_lengths, unsigned int rate_cats, PmatModel m)
{
  const unsigned int n = blockIdx.x % rate_cats;
  const unsigned int i = blockIdx.x / rate_cats;
  const double t = branch_lengths[i];
  double * P = pmatrix + (size_t)matrix_indices[i] * pmat_len + (size_t)n * 400;

  if (t == 0.0)
  {
    for (unsigned int q = threadIdx.x; q < 400; q += blockDim.x) P[q] = (q / 20 == q % 20) ? 1.0 : 0.0;
    return;
  }

  __shared__ double expd[20];
  __shared__ double T[400];
  const double * V = m.eigenvecs + (size_t)n * 400;
  const double * iV = m.inv_eigenvecs + (size_t)n * 400;
  if (threadIdx.x < 20)
    expd[threadIdx.x] =
        pmat_exponent(m.eigenvals[n * 20 + threadIdx.x], m.rates[n], t, m.prop_invar[n]);
  __syncthreads();
  for (unsigned int q = threadIdx.x; q < 400; q += blockDim.x)
    T[q] = __dmul_rn(expd[q % 20], iV[q]);
  __syncthreads();
  for (unsigned int q = threadIdx.x; q < 400; q += blockDim.x)
  {
    const unsigned int j = q / 20, c = q % 20;
    const double * Tj = T + j * 20;
    double a0 = __dmul_rn(Tj[0], V[0 * 20 + c]);
    double a1 = __dmul_rn(Tj[1], V[1 * 20 + c]);
    double a2 = __dmul_rn(Tj[2], V[2 * 20 + c]);
    double a3 = __dmul_rn(Tj[3], V[3 * 20 + c]);
#pragma unroll
    for (int b = 1; b < 5; ++b)
    {
      a0 = __fma_rn(Tj[4 * b + 0], V[(4 * b + 0) * 20 + c], a0);
      a1 = __fma_rn(Tj[4 * b + 1], V[(4 * b + 1) * 20 + c], a1);
      a2 = __fma_rn(Tj[4 * b + 2], V[(4 * b + 2) * 20 + c], a2);
      a3 = __fma_rn(Tj[4 * b + 3], V[(4 * b + 3) * 20 + c], a3);
    }
    double s = hsum4(a0, a1, a2, a3);
    if (j == c) s = __dadd_rn(s, 1.0);
    P[q] = s;
  }
}

extern "C" int plg_update_pmatrix(plg_context_t * ctx, const unsigned int * matrix_indices,
                                  const double * branch_lengths, unsigned int count,
                                  const double * rates, const double * prop_invar,
                                  const double * eigenvals, const double * eigenvecs,
                                  const double * inv_eigenvecs)
{
  PLG_CHECK_CTX(ctx);
  if (count == 0) return PLG_OK;
  const unsigned int R = ctx->d.rate_cats, K = ctx->d.states, Kp = ctx->d.states_padded;
  for (unsigned int i = 0; i < count; ++i)
  {
    if (matrix_indices[i] >= ctx->d.prob_matrices)
    {
      plg_set_error("plg_update_pmatrix: matrix index %u out of range", matrix_indices[i]);
      return PLG_E_INVALID;
    }
    if (!(branch_lengths[i] >= 0))
    {
      plg_set_error("plg_update_pmatrix: negative branch length %g", branch_lengths[i]);
      return PLG_E_INVALID;
    }
  }

  /* branches in chunks; each chunk re-stages the (tiny) model so that one contiguous
   * reservation of the staging ring covers everything its kernel reads */
  const unsigned int chunk = 65536;
  const size_t model_bytes = ((size_t)R * Kp + 2 * (size_t)R * K * Kp + 2 * R) * sizeof(double);
  for (unsigned int off = 0; off < count; off += chunk)
  {
    const unsigned int c = (count - off < chunk) ? count - off : chunk;
    if (plg_stage_reserve(ctx, model_bytes + (size_t)c * 12 + 8 * 256)) return PLG_E_CUDA;
    PmatModel m;
    m.eigenvals = (const double *)plg_stage(ctx, eigenvals, (size_t)R * Kp * sizeof(double));
    m.eigenvecs = (const double *)plg_stage(ctx, eigenvecs, (size_t)R * K * Kp * sizeof(double));
    m.inv_eigenvecs =
        (const double *)plg_stage(ctx, inv_eigenvecs, (size_t)R * K * Kp * sizeof(double));
    m.rates = (const double *)plg_stage(ctx, rates, R * sizeof(double));
    m.prop_invar = (const double *)plg_stage(ctx, prop_invar, R * sizeof(double));
    const unsigned int * d_idx =
        (const unsigned int *)plg_stage(ctx, matrix_indices + off, c * sizeof(unsigned int));
    const double * d_bl = (const double *)plg_stage(ctx, branch_lengths + off, c * sizeof(double));
    if (!m.eigenvals || !m.eigenvecs || !m.inv_eigenvecs || !m.rates || !m.prop_invar || !d_idx ||
        !d_bl)
      return PLG_E_CUDA;
    if (K == 4)
    {
      const unsigned int threads = c * R * 4;
      k_pmatrix_dna<<<(threads + 127) / 128, 128, 0, ctx->stream>>>(ctx->pmatrix, ctx->pmat_len,
                                                                     d_idx, d_bl, c, R, m);
    }
    else if (K == 20)
      k_pmatrix_aa<<<c * R, 256, 0, ctx->stream>>>(ctx->pmatrix, ctx->pmat_len, d_idx, d_bl, R, m);
    else
    {
      int rc = plg_gen_pmatrix(ctx, d_idx, d_bl, c, m.eigenvals, m.eigenvecs, m.inv_eigenvecs, m.rates,
                               m.prop_invar);
      if (rc) return rc;
    }
    PLG_LAUNCH_CHECK(ctx);
  }
  return PLG_OK;
}
