/*
 * pll_ascbias.c - ascertainment-bias correction: the host-side epilogue of
 * the log-likelihood and derivative calls plus the two setters.
 *
 * Storage follows the reference (src/pll.c:492-495): a partition created with any
 * PLL_ATTRIB_AB_* bit holds `sites + states` sites; the trailing `states` "dummy" sites carry,
 * at every tip, the pure state i at position sites + i (src/pll.c:847-856, 942-961,
 * 1028-1043), with weights set by pll_set_asc_state_weights.  The CLV-update kernels simply
 * run over all sites.  The correction terms need only those few trailing sites, so they are
 * read back (states x rate_cats x states_padded doubles per CLV) and combined here with the
 * arithmetic of the reference:
 *   log-likelihood  src/likelihood.c:24-48 (formulas), :50-119 (root), :170-247 (tip-inner),
 *                   :321-414 (inner-inner)
 *   derivatives     src/core_derivatives.c:449-500 (per-site terms), :654-727 (combination)
 */
#include "pll_host.h"

PLL_EXPORT int pll_set_asc_bias_type(pll_partition_t * partition, int asc_bias_type)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  if (!p->asc_bias_alloc)
    return pll_fail(PLL_ERROR_AB_NOSUPPORT, "Partition was not created with ascertainment bias support");
  if (asc_bias_type != 0)
    for (unsigned int i = 0; i < p->rate_matrices; ++i)
      if (p->prop_invar[i] > 0)
        return pll_fail(PLL_ERROR_INVAR_INCOMPAT,
                        "Invariant sites are not compatible with asc bias correction");
  if ((asc_bias_type & PLL_ATTRIB_AB_MASK) != asc_bias_type)
    return pll_fail(PLL_ERROR_AB_INVALIDMETHOD, "Illegal ascertainment bias algorithm \"%d\"",
                    asc_bias_type);
  p->attributes = (p->attributes & ~(unsigned int)PLL_ATTRIB_AB_MASK) | (unsigned int)asc_bias_type;
  return PLL_SUCCESS;
}

PLL_EXPORT void pll_set_asc_state_weights(pll_partition_t * partition, const unsigned int * state_weights)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g || !g->pub.asc_bias_alloc)
  {
    pll_fail(PLL_ERROR_AB_NOSUPPORT, "Partition was not created with ascertainment bias support");
    return;
  }
  pll_partition_t * p = &g->pub;
  memcpy(p->pattern_weights + p->sites, state_weights, p->states * sizeof(unsigned int));
  int rc = pllg_dev_set_pattern_weights(g, p->pattern_weights);
  if (rc) pllg_fail(rc, "pll_set_asc_state_weights");
}

/* ---- helpers ---- */
static double * fetch_clv_tail(pllg_partition_t * g, unsigned int clv_index)
{
  const pll_partition_t * p = &g->pub;
  double * buf = (double *)malloc((size_t)p->states * p->rate_cats * p->states_padded * sizeof(double));
  if (!buf)
  {
    pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  int rc = pllg_dev_get_clv_sites(g, clv_index, p->sites, p->states, buf);
  if (rc)
  {
    pllg_fail(rc, "ascertainment bias correction");
    free(buf);
    return NULL;
  }
  return buf;
}

/* scaler counts of the trailing sites; zeros when there is no scaler.
 *
 * With PLL_ATTRIB_RATE_SCALERS the reference still reads element `sites + n` of the scaler array
 * (src/likelihood.c:91, :233, :378-381, src/core_derivatives.c:684-685) although that array is
 * laid out [site][rate]: what it gets is the count of pattern (sites + n) / R at rate
 * (sites + n) % R.  The value is well defined (the array holds (sites + states) * R counts), so a
 * drop-in returns the same one: the patterns that cover those flat elements are fetched and
 * element `sites + n` picked out of them. */
static int fetch_scaler_tail(pllg_partition_t * g, int scaler_index, unsigned int * out)
{
  const pll_partition_t * p = &g->pub;
  memset(out, 0, p->states * sizeof(unsigned int));
  if (scaler_index == PLL_SCALE_BUFFER_NONE) return 1;
  const unsigned int R = p->rate_cats;
  if (!(p->attributes & PLL_ATTRIB_RATE_SCALERS) || R == 1)
  {
    int rc = pllg_dev_get_scaler_sites(g, (unsigned int)scaler_index, p->sites, p->states, out);
    return rc ? pllg_fail(rc, "ascertainment bias correction") : 1;
  }
  const unsigned int first = p->sites / R, last = (p->sites + p->states - 1) / R;
  const unsigned int n = last - first + 1;
  unsigned int * rows = (unsigned int *)malloc((size_t)n * R * sizeof(unsigned int));
  if (!rows) return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
  int rc = pllg_dev_get_scaler_sites(g, (unsigned int)scaler_index, first, n, rows);
  if (!rc) memcpy(out, rows + (p->sites - (size_t)first * R), p->states * sizeof(unsigned int));
  free(rows);
  return rc ? pllg_fail(rc, "ascertainment bias correction") : 1;
}

/* reference src/likelihood.c:24-48 */
static double correction(double base, unsigned int sum_w, unsigned int sum_w_inv, int type)
{
  switch (type)
  {
    case PLL_ATTRIB_AB_LEWIS: return -(sum_w * log(1 - base));
    case PLL_ATTRIB_AB_STAMATAKIS: return base;
    case PLL_ATTRIB_AB_FELSENSTEIN: return sum_w_inv * log(base);
    default:
      pll_fail(PLL_ERROR_AB_INVALIDMETHOD, "Illegal ascertainment bias algorithm");
      return -INFINITY;
  }
}

/* turns the per-state site likelihoods `term[n]` into the correction term */
static double combine(const pll_partition_t * p, const double * term, const unsigned int * scale_factors)
{
  const int type = (int)(p->attributes & PLL_ATTRIB_AB_MASK);
  const unsigned int * w = p->pattern_weights + p->sites;
  double acc = 0;
  unsigned int sum_w_inv = 0;
  for (unsigned int n = 0; n < p->states; ++n)
  {
    double site_lk;
    sum_w_inv += w[n];
    if (type == PLL_ATTRIB_AB_STAMATAKIS)
    {
      site_lk = log(term[n]) * w[n];
      if (scale_factors[n]) site_lk += scale_factors[n] * log(PLL_SCALE_THRESHOLD);
    }
    else
      site_lk = term[n] * pow(PLL_SCALE_THRESHOLD, scale_factors[n]);
    acc += site_lk;
  }
  return correction(acc, p->pattern_weight_sum, sum_w_inv, type);
}

/* ---- log-likelihood epilogues (called by pll_likelihood.c when an AB type is set) ---- */
double pllg_asc_root(pllg_partition_t * g, unsigned int clv_index, int scaler_index,
                     const unsigned int * freqs_indices)
{
  const pll_partition_t * p = &g->pub;
  const unsigned int K = p->states, Kp = p->states_padded, R = p->rate_cats;
  double * clv = fetch_clv_tail(g, clv_index);
  double * term = (double *)malloc(K * sizeof(double));
  unsigned int * sf = (unsigned int *)malloc(K * sizeof(unsigned int));
  double out = -INFINITY;
  if (clv && term && sf && fetch_scaler_tail(g, scaler_index, sf))
  {
    const double * c = clv;
    for (unsigned int n = 0; n < K; ++n)
    {
      double t = 0;
      for (unsigned int j = 0; j < R; ++j, c += Kp)
      {
        const double * f = p->frequencies[freqs_indices[j]];
        double tr = 0;
        for (unsigned int k = 0; k < K; ++k) tr += c[k] * f[k];
        t += tr * p->rate_weights[j];
      }
      term[n] = t;
    }
    out = combine(p, term, sf);
  }
  free(clv);
  free(term);
  free(sf);
  return out;
}

double pllg_asc_edge(pllg_partition_t * g, unsigned int parent_clv_index, int parent_scaler_index,
                     unsigned int child_clv_index, int child_scaler_index, unsigned int matrix_index,
                     const unsigned int * freqs_indices)
{
  const pll_partition_t * p = &g->pub;
  const unsigned int K = p->states, Kp = p->states_padded, R = p->rate_cats;
  const int pattern_tip = (p->attributes & PLL_ATTRIB_PATTERN_TIP) != 0;
  const int ptip = pattern_tip && parent_clv_index < p->tips;
  const int ctip = pattern_tip && child_clv_index < p->tips;
  /* with one end a pattern tip the inner end plays "parent" and only its scaler counts
   * (reference src/likelihood.c:486-501) */
  const unsigned int a_clv = ptip ? child_clv_index : parent_clv_index;
  const int a_sc = ptip ? child_scaler_index : parent_scaler_index;
  const int tip_edge = ptip || ctip;

  double * pm = (double *)malloc((size_t)R * K * Kp * sizeof(double));
  double * clvp = fetch_clv_tail(g, a_clv);
  double * clvc = tip_edge ? NULL : fetch_clv_tail(g, child_clv_index);
  double * term = (double *)malloc(K * sizeof(double));
  unsigned int * sf = (unsigned int *)malloc(2 * K * sizeof(unsigned int));
  double out = -INFINITY;
  int ok = pm && clvp && (tip_edge || clvc) && term && sf;
  if (ok)
  {
    int rc = plg_get_pmatrix(g->ctx, matrix_index, pm);
    if (rc) ok = pllg_fail(rc, "ascertainment bias correction");
  }
  if (ok) ok = fetch_scaler_tail(g, a_sc, sf);
  if (ok && !tip_edge)
  {
    ok = fetch_scaler_tail(g, child_scaler_index, sf + K);
    for (unsigned int n = 0; ok && n < K; ++n) sf[n] += sf[K + n];
  }
  if (ok)
  {
    const double * cp = clvp, * cc = clvc;
    for (unsigned int n = 0; n < K; ++n)
    {
      const double * m = pm;
      double t = 0;
      for (unsigned int i = 0; i < R; ++i, cp += Kp)
      {
        const double * f = p->frequencies[freqs_indices[i]];
        double tr = 0;
        for (unsigned int j = 0; j < K; ++j, m += Kp)
        {
          if (tip_edge)
            tr += cp[j] * f[j] * m[n]; /* the tip shows state n at dummy site n */
          else
          {
            double tb = 0;
            for (unsigned int k = 0; k < K; ++k) tb += m[k] * cc[k];
            tr += cp[j] * f[j] * tb;
          }
        }
        t += tr * p->rate_weights[i];
        if (!tip_edge) cc += Kp;
      }
      term[n] = t;
    }
    out = combine(p, term, sf);
  }
  free(pm);
  free(clvp);
  free(clvc);
  free(term);
  free(sf);
  return out;
}

/* ---- derivative epilogue: adds the Lewis / Felsenstein terms to d_f, dd_f ---- */
int pllg_asc_derivatives(pllg_partition_t * g, int parent_scaler_index, int child_scaler_index,
                         const double * diagptable, const double * sumtable_key, double * d_f,
                         double * dd_f)
{
  const pll_partition_t * p = &g->pub;
  const unsigned int K = p->states, Kp = p->states_padded, R = p->rate_cats;
  const int type = (int)(p->attributes & PLL_ATTRIB_AB_MASK);
  double * sum = (double *)malloc((size_t)K * R * Kp * sizeof(double));
  unsigned int * sf = (unsigned int *)malloc(2 * K * sizeof(unsigned int));
  int ok = sum && sf;
  if (!ok) pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
  if (ok)
  {
    int rc = pllg_dev_get_sumtable_sites(g, sumtable_key, p->sites, K, sum);
    if (rc) ok = pllg_fail(rc, "pll_compute_likelihood_derivatives");
  }
  if (ok) ok = fetch_scaler_tail(g, parent_scaler_index, sf) && fetch_scaler_tail(g, child_scaler_index, sf + K);
  if (ok)
  {
    double lk[3] = {0.0, 0.0, 0.0};
    unsigned int sum_w_inv = 0;
    const double * s = sum;
    for (unsigned int n = 0; n < K; ++n)
    {
      double site[3] = {0, 0, 0};
      const double * d = diagptable;
      for (unsigned int i = 0; i < R; ++i, s += Kp)
      {
        double cat[3] = {0, 0, 0};
        for (unsigned int j = 0; j < K; ++j, d += 4)
        {
          cat[0] += s[j] * d[0];
          cat[1] += s[j] * d[1];
          cat[2] += s[j] * d[2];
        }
        site[0] += cat[0] * p->rate_weights[i];
        site[1] += cat[1] * p->rate_weights[i];
        site[2] += cat[2] * p->rate_weights[i];
      }
      const double scaling = pow(PLL_SCALE_THRESHOLD, (double)(sf[n] + sf[K + n]));
      lk[0] += site[0] * scaling;
      lk[1] += site[1] * scaling;
      lk[2] += site[2] * scaling;
      sum_w_inv += p->pattern_weights[p->sites + n];
    }
    switch (type)
    {
      case PLL_ATTRIB_AB_LEWIS:
      {
        unsigned int w = 0;
        for (unsigned int n = 0; n < p->sites; ++n) w += p->pattern_weights[n];
        *d_f += w * (lk[1] / (lk[0] - 1.0));
        *dd_f += w * (((lk[0] - 1.0) * lk[2] - lk[1] * lk[1]) / ((lk[0] - 1.0) * (lk[0] - 1.0)));
        break;
      }
      case PLL_ATTRIB_AB_FELSENSTEIN:
        *d_f -= sum_w_inv * (lk[1] / lk[0]);
        *dd_f -= sum_w_inv * (((lk[2] * lk[0]) - lk[1] * lk[1]) / (lk[0] * lk[0]));
        break;
      default: ok = pll_fail(PLL_ERROR_AB_INVALIDMETHOD, "Illegal ascertainment bias algorithm");
    }
  }
  free(sum);
  free(sf);
  return ok;
}
