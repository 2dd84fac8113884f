/*
 * pll_derivatives.c - sumtable + branch-length derivatives entry points (host wrappers).
 *
 * Mirrors reference src/derivatives.c:164-234 (pll_update_sumtable) and :243-312
 * (pll_compute_likelihood_derivatives).  The small per-call tables that the reference
 * kernels build before their site loop are built here, with the reference's operation order
 * (this file is compiled with -ffp-contract=off; fused steps call fma() explicitly):
 *   inner-inner  W = inv_eigenvecs^T scaled by pi     reference src/core_derivatives_avx.c:86-95
 *   tip-inner    per-code left terms                  reference src/core_derivatives_avx.c:556-579 (DNA)
 *                                                     reference src/core_derivatives_avx2.c:368-410 (20 st.)
 *   diagptable   {e, lk e, (lk)^2 e, 0}               reference src/core_derivatives.c:560-575
 *
 * The `sumtable` argument is only an opaque key: the table stays in HBM (see pll.h).
 */
#include "pll_host.h"

static int want_hostcopy(void)
{
  const char * e = getenv("PLL_GPU_SUMTABLE_HOSTCOPY");
  return e && *e && *e != '0';
}

PLL_EXPORT int pll_update_sumtable(pll_partition_t * partition,
                                   unsigned int parent_clv_index,
                                   unsigned int child_clv_index,
                                   int parent_scaler_index,
                                   int child_scaler_index,
                                   const unsigned int * params_indices,
                                   double * sumtable)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  const unsigned int R = p->rate_cats, K = p->states, Kp = p->states_padded;
  unsigned int r, i, j, n;

  const int pattern_tip = (p->attributes & PLL_ATTRIB_PATTERN_TIP) != 0;
  const int ptip = pattern_tip && parent_clv_index < p->tips;
  const int ctip = pattern_tip && child_clv_index < p->tips;
  if (ptip && ctip)
    return pll_fail(PLL_ERROR_PARAM_INVALID, "pll_update_sumtable: tip-tip edge is not supported");

  for (r = 0; r < R; ++r)
    if (!p->eigen_decomp_valid[params_indices[r]])
      if (!pll_update_eigen(p, params_indices[r])) return PLL_FAILURE;

  const size_t codes = (ptip || ctip) ? (K == 4 ? 16u : p->maxstates) : 0;
  const size_t mat = (size_t)R * K * Kp;
  const size_t left_len = codes ? codes * R * Kp : mat;
  double * evecs = (double *)calloc(mat + left_len, sizeof(double));
  if (!evecs) return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
  double * left = evecs + mat;

  for (r = 0; r < R; ++r)
    memcpy(evecs + (size_t)r * K * Kp, p->eigenvecs[params_indices[r]],
           (size_t)K * Kp * sizeof(double));

  if (!codes)
  {
    for (r = 0; r < R; ++r)
    {
      const double * iv = p->inv_eigenvecs[params_indices[r]];
      const double * f = p->frequencies[params_indices[r]];
      for (j = 0; j < K; ++j)
        for (i = 0; i < K; ++i) left[(size_t)r * K * Kp + (size_t)j * Kp + i] = iv[(size_t)i * Kp + j] * f[i];
    }
  }
  else
  {
    for (n = 0; n < codes; ++n)
    {
      const unsigned int state = (K == 4) ? n : p->tipmap[n];
      for (r = 0; r < R; ++r)
      {
        const double * iv = p->inv_eigenvecs[params_indices[r]];
        const double * f = p->frequencies[params_indices[r]];
        double * out = left + ((size_t)n * R + r) * Kp;
        for (j = 0; j < K; ++j)
        {
          double acc = 0.0;
          for (i = 0; i < K; ++i)
          {
            if (!((state >> i) & 1u)) continue;
            if (K == 4)
              acc = acc + iv[(size_t)i * Kp + j] * f[i]; /* mul, then add */
            else
              acc = fma(iv[(size_t)i * Kp + j], f[i], acc);
          }
          out[j] = acc;
        }
      }
    }
  }

  int rc = pllg_dev_update_sumtable(g, parent_clv_index, child_clv_index, parent_scaler_index,
                               child_scaler_index, evecs, left, sumtable,
                               want_hostcopy() ? sumtable : NULL);
  free(evecs);
  return rc ? pllg_fail(rc, "pll_update_sumtable") : PLL_SUCCESS;
}

PLL_EXPORT int pll_compute_likelihood_derivatives(pll_partition_t * partition,
                                                  int parent_scaler_index,
                                                  int child_scaler_index,
                                                  double branch_length,
                                                  const unsigned int * params_indices,
                                                  const double * sumtable,
                                                  double * d_f,
                                                  double * dd_f)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  const unsigned int R = p->rate_cats, K = p->states, Kp = p->states_padded;
  unsigned int i, j;
  /* site scalers cancel in L'/L; only the ascertainment-bias terms below use them */

  double * buf = (double *)calloc((size_t)R * K * 4 + (size_t)R * Kp + 2 * R, sizeof(double));
  if (!buf) return pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate memory for diagptable");
  double * diagp = buf;
  double * freqs = diagp + (size_t)R * K * 4;
  double * pinv = freqs + (size_t)R * Kp;

  for (i = 0; i < R; ++i)
  {
    const unsigned int pi = params_indices[i];
    const double * ev = p->eigenvals[pi];
    pinv[i] = p->prop_invar[pi];
    memcpy(freqs + (size_t)i * Kp, p->frequencies[pi], Kp * sizeof(double));
    const double ki = p->rates[i] / (1.0 - pinv[i]);
    double * d = diagp + (size_t)i * K * 4;
    for (j = 0; j < K; ++j, d += 4)
    {
      d[0] = exp(ev[j] * ki * branch_length);
      d[1] = ev[j] * ki * d[0];
      d[2] = ev[j] * ki * ev[j] * ki * d[0];
      d[3] = 0;
    }
  }

  /* ascertainment bias (reference src/core_derivatives.c:536-545, 654-727): Stamatakis simply
   * extends the weighted sum over the per-state sites; Lewis / Felsenstein add a term built
   * from the per-state site likelihoods */
  const int ab = (int)(p->attributes & PLL_ATTRIB_AB_MASK);
  int rc = PLG_OK;
  if (ab == PLL_ATTRIB_AB_STAMATAKIS) rc = pllg_dev_set_active_sites(g, p->sites + K);
  if (!rc)
    rc = pllg_dev_likelihood_derivatives(g, sumtable, diagp, p->rate_weights, pinv, freqs, d_f, dd_f);
  if (ab == PLL_ATTRIB_AB_STAMATAKIS) pllg_dev_set_active_sites(g, p->sites);
  int ok = rc ? pllg_fail(rc, "pll_compute_likelihood_derivatives") : PLL_SUCCESS;
  if (ok && ab && ab != PLL_ATTRIB_AB_STAMATAKIS)
    ok = pllg_asc_derivatives(g, parent_scaler_index, child_scaler_index, diagp, sumtable, d_f, dd_f);
  free(buf);
  return ok;
}
