/*
 * pll_oracle.c - TEST INFRASTRUCTURE ONLY: a plain scalar-C restatement of the reference's
 * likelihood hot path, in the operation order of the path its PLL_ATTRIB_ARCH_AVX2 flag
 * actually executes.  Nothing under libpll_b200/ includes, links or calls
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.
 *
 * Parity pinning: tests/test_oracle_cpu.py checks every function here BIT FOR BIT against the
 * reference's own exported pll_core_* kernels (oracle/_ref, built from /root/reference by
 * oracle/Makefile) on random inputs, and the assembled pipeline against the reference's
 * golden outputs (tests/golden/, derived from reference test/out and examples).
 *
 * Supported: states 4 (the reference's *_4x4_avx kernels), 20 (its AVX2 generic / 20x20
 * kernels) and any other alphabet up to 64 states in the operation order of the reference's
 * plain-C code (src/core_partials.c:510-663, src/core_likelihood.c:163-209, 617-725, 905-1010,
 * src/core_derivatives.c:240-262, 415-440, 449-500, src/core_pmatrix.c:146-250);
 * states_padded == states, any rate_cats, per-site and per-rate scaling.
 * Compile with -ffp-contract=off: every fused multiply-add below is an explicit fma().
 *
 * Signatures mirror the reference's pll_core_* prototypes (reference src/pll.h:829-1027,
 * 1659-1672) so that a test can call both with the same argument list.
 */
#include "pll_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SCALE_FACTOR 0x1p+256
#define ORC_SCALE_THRESHOLD 0x1p-256
#define ORC_RATE_SCALERS (1u << 9) /* PLL_ATTRIB_RATE_SCALERS */
#define ORC_MAXDIFF 4              /* PLL_SCALE_RATE_MAXDIFF */
#define ORC_MISC_EPSILON 1e-8

/* the 4-lane horizontal sum idiom of every AVX kernel (e.g. reference
 * src/core_partials_avx.c:460-471): (a0+a1)+(a2+a3) */
static double hsum4(double a0, double a1, double a2, double a3) { return (a0 + a1) + (a2 + a3); }

/* row . vector over K (multiple of 4) columns with four lane accumulators walking the column
 * blocks; fused = AVX2 fmadd (reference src/core_partials_avx2.c:671-731), unfused = AVX
 * mul+add (reference src/core_partials_avx.c:1236-1286) */
static double row_dot(const double * row, const double * v, unsigned int K, int fused)
{
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  unsigned int b, l;
  if (K != 4 && K != 20)
  {
    /* any other alphabet: the plain sequential sum of the reference's non-SIMD code
     * (e.g. reference src/core_partials.c:630-640) */
    double s = 0.0;
    for (b = 0; b < K; ++b) s += row[b] * v[b];
    return s;
  }
  for (b = 0; b < K; b += 4)
    for (l = 0; l < 4; ++l) a[l] = fused ? fma(row[b + l], v[b + l], a[l]) : a[l] + row[b + l] * v[b + l];
  return hsum4(a[0], a[1], a[2], a[3]);
}

/* ------------------------------------------------------------------------------------ */
/* P-matrices: reference src/core_pmatrix_avx.c:42-310 (4x4), src/core_pmatrix_avx2.c:37-284 */
/* ------------------------------------------------------------------------------------ */
int orc_core_update_pmatrix(double ** pmatrix, unsigned int states, unsigned int rate_cats,
                            const double * rates, const double * branch_lengths,
                            const unsigned int * matrix_indices,
                            const unsigned int * params_indices, const double * prop_invar,
                            double * const * eigenvals, double * const * eigenvecs,
                            double * const * inv_eigenvecs, unsigned int count,
                            unsigned int attrib)
{
  const unsigned int K = states;
  unsigned int i, n, j, c, m;
  double e[64], T[4096];
  (void)attrib;
  if (K > 64) return 0;
  for (i = 0; i < count; ++i)
  {
    double * pmat = pmatrix[matrix_indices[i]];
    for (n = 0; n < rate_cats; ++n, pmat += K * K)
    {
      const double pinv = prop_invar[params_indices[n]];
      const double * V = eigenvecs[params_indices[n]];
      const double * iV = inv_eigenvecs[params_indices[n]];
      const double * ev = eigenvals[params_indices[n]];
      if (!branch_lengths[i])
      {
        for (j = 0; j < K; ++j)
          for (c = 0; c < K; ++c) pmat[j * K + c] = (j == c) ? 1.0 : 0.0;
        continue;
      }
      for (m = 0; m < K; ++m)
      {
        double x = (ev[m] * rates[n]) * branch_lengths[i];
        if (pinv > ORC_MISC_EPSILON) x = x / (1.0 - pinv);
        e[m] = expm1(x);
      }
      for (j = 0; j < K; ++j)
        for (m = 0; m < K; ++m) T[j * K + m] = iV[j * K + m] * e[m];
      for (j = 0; j < K; ++j)
        for (c = 0; c < K; ++c)
        {
          double s;
          if (K == 4)
            s = hsum4(T[j * 4 + 0] * V[0 + c], T[j * 4 + 1] * V[4 + c], T[j * 4 + 2] * V[8 + c],
                      T[j * 4 + 3] * V[12 + c]) + ((j == c) ? 1.0 : 0.0);
          else if (K != 20)
          {
            /* plain loop, identity added first (reference src/core_pmatrix.c:225-236) */
            s = (j == c) ? 1.0 : 0.0;
            for (m = 0; m < K; ++m) s += T[j * K + m] * V[m * K + c];
          }
          else
          {
            /* first column block by mul, the rest by fmadd (reference
             * src/core_pmatrix_avx2.c:24-35), then + 1 on the diagonal (:265-271) */
            double a[4];
            unsigned int b, l;
            for (l = 0; l < 4; ++l) a[l] = T[j * K + l] * V[l * K + c];
            for (b = 4; b < K; b += 4)
              for (l = 0; l < 4; ++l) a[l] = fma(T[j * K + b + l], V[(b + l) * K + c], a[l]);
            s = hsum4(a[0], a[1], a[2], a[3]);
            if (j == c) s += 1.0;
          }
          pmat[j * K + c] = s;
        }
    }
  }
  return 1;
}

/* ------------------------------------------------------------------------------------ */
/* tip tables                                                                            */
/* ------------------------------------------------------------------------------------ */
/* table[code][rate][i] = sum of P_rate[i][m] over the states m in the code's mask.
 * DNA: masked 4-lane sum, absent states add +0.0 (reference src/core_partials_avx.c:944-984);
 * 20 states: sequential sum over set bits via tipmap (reference src/core_partials_avx.c:1140-1177) */
static void tip_table(unsigned int K, unsigned int R, unsigned int codes, const unsigned int * tipmap,
                      const double * pmat, double * table)
{
  unsigned int code, k, i, m;
  for (code = 0; code < codes; ++code)
    for (k = 0; k < R; ++k)
      for (i = 0; i < K; ++i)
      {
        const double * row = pmat + (size_t)k * K * K + i * K;
        double s;
        if (K == 4)
          s = hsum4((code & 1) ? row[0] : 0.0, (code & 2) ? row[1] : 0.0, (code & 4) ? row[2] : 0.0,
                    (code & 8) ? row[3] : 0.0);
        else
        {
          const unsigned int state = tipmap[code];
          s = 0.0;
          for (m = 0; m < K; ++m)
            if ((state >> m) & 1u) s += row[m];
        }
        table[((size_t)code * R + k) * K + i] = s;
      }
}

/* fill_parent_scaler (reference src/core_partials_avx.c:24-46) */
static void fill_scaler(size_t n, unsigned int * parent, const unsigned int * l, const unsigned int * r)
{
  size_t i;
  if (!l && !r) memset(parent, 0, n * sizeof(unsigned int));
  else if (l && r)
  {
    memcpy(parent, l, n * sizeof(unsigned int));
    for (i = 0; i < n; ++i) parent[i] += r[i];
  }
  else
    memcpy(parent, l ? l : r, n * sizeof(unsigned int));
}

/* threshold rescale of one site (reference src/core_partials_avx.c:490-527): per-rate mode
 * looks at each rate block alone, per-site mode needs every entry of the site below 2^-256 */
static void rescale_site(double * clv, unsigned int K, unsigned int R, unsigned int * scaler,
                         unsigned int n, int mode)
{
  unsigned int k, i;
  int all = 1;
  if (!mode) return;
  for (k = 0; k < R; ++k)
  {
    int below = 1;
    for (i = 0; i < K; ++i)
      if (!(clv[k * K + i] < ORC_SCALE_THRESHOLD)) below = 0;
    if (mode == 2)
    {
      if (below)
      {
        for (i = 0; i < K; ++i) clv[k * K + i] *= ORC_SCALE_FACTOR;
        scaler[n * R + k] += 1;
      }
    }
    else
      all &= below;
  }
  if (mode == 1 && all)
  {
    for (i = 0; i < R * K; ++i) clv[i] *= ORC_SCALE_FACTOR;
    scaler[n] += 1;
  }
}

static int scale_mode_of(const unsigned int * parent_scaler, unsigned int attrib)
{
  if (!parent_scaler) return 0;
  return (attrib & ORC_RATE_SCALERS) ? 2 : 1;
}

/* ------------------------------------------------------------------------------------ */
/* CLV updates                                                                           */
/* ------------------------------------------------------------------------------------ */
/* inner-inner: reference src/core_partials_avx.c:366-529 (4x4, unfused),
 * src/core_partials_avx2.c:568-803 (generic, fused) */
void orc_core_update_partial_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                double * parent_clv, unsigned int * parent_scaler,
                                const double * left_clv, const double * right_clv,
                                const double * left_matrix, const double * right_matrix,
                                const unsigned int * left_scaler, const unsigned int * right_scaler,
                                unsigned int attrib)
{
  const unsigned int K = states, R = rate_cats;
  const int mode = scale_mode_of(parent_scaler, attrib);
  unsigned int n, k, i;
  if (mode) fill_scaler(mode == 2 ? (size_t)sites * R : sites, parent_scaler, left_scaler, right_scaler);
  for (n = 0; n < sites; ++n)
  {
    double * p = parent_clv + (size_t)n * R * K;
    for (k = 0; k < R; ++k)
    {
      const double * l = left_clv + ((size_t)n * R + k) * K;
      const double * r = right_clv + ((size_t)n * R + k) * K;
      const double * L = left_matrix + (size_t)k * K * K;
      const double * Rm = right_matrix + (size_t)k * K * K;
      for (i = 0; i < K; ++i)
      {
        double x, y;
        if (K == 4)
        {
          x = hsum4(L[i * 4] * l[0], L[i * 4 + 1] * l[1], L[i * 4 + 2] * l[2], L[i * 4 + 3] * l[3]);
          y = hsum4(Rm[i * 4] * r[0], Rm[i * 4 + 1] * r[1], Rm[i * 4 + 2] * r[2], Rm[i * 4 + 3] * r[3]);
        }
        else
        {
          x = row_dot(L + i * K, l, K, 1);
          y = row_dot(Rm + i * K, r, K, 1);
        }
        p[k * K + i] = x * y;
      }
    }
    rescale_site(p, K, R, parent_scaler, n, mode);
  }
}

/* tip-inner: reference src/core_partials_avx.c:899-1095 (4x4), :1097-1340 (20x20; the AVX
 * kernel is what the AVX2 flag dispatches to, reference src/core_partials.c:427-444) */
void orc_core_update_partial_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                double * parent_clv, unsigned int * parent_scaler,
                                const unsigned char * left_tipchars, const double * right_clv,
                                const double * left_matrix, const double * right_matrix,
                                const unsigned int * right_scaler, const unsigned int * tipmap,
                                unsigned int tipmap_size, unsigned int attrib)
{
  const unsigned int K = states, R = rate_cats;
  const unsigned int codes = (K == 4) ? 16u : tipmap_size;
  const int mode = scale_mode_of(parent_scaler, attrib);
  unsigned int n, k, i;
  double * table = (double *)malloc((size_t)codes * R * K * sizeof(double));
  tip_table(K, R, codes, tipmap, left_matrix, table);
  if (mode) fill_scaler(mode == 2 ? (size_t)sites * R : sites, parent_scaler, NULL, right_scaler);
  for (n = 0; n < sites; ++n)
  {
    double * p = parent_clv + (size_t)n * R * K;
    const double * t = table + (size_t)left_tipchars[n] * R * K;
    for (k = 0; k < R; ++k)
    {
      const double * r = right_clv + ((size_t)n * R + k) * K;
      const double * Rm = right_matrix + (size_t)k * K * K;
      for (i = 0; i < K; ++i)
      {
        const double y = (K == 4)
                             ? hsum4(Rm[i * 4] * r[0], Rm[i * 4 + 1] * r[1], Rm[i * 4 + 2] * r[2],
                                     Rm[i * 4 + 3] * r[3])
                             : row_dot(Rm + i * K, r, K, 0);
        p[k * K + i] = t[k * K + i] * y;
      }
    }
    rescale_site(p, K, R, parent_scaler, n, mode);
  }
  free(table);
}

/* tip-tip: product of the two per-side tables, never scaled, parent scaler zeroed
 * (reference src/core_partials_avx.c:262-364, 531-618, 146-260) */
void orc_core_update_partial_tt(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                double * parent_clv, unsigned int * parent_scaler,
                                const unsigned char * left_tipchars,
                                const unsigned char * right_tipchars, const double * left_matrix,
                                const double * right_matrix, const unsigned int * tipmap,
                                unsigned int tipmap_size, unsigned int attrib)
{
  const unsigned int K = states, R = rate_cats;
  const unsigned int codes = (K == 4) ? 16u : tipmap_size;
  unsigned int n, q;
  double * tl = (double *)malloc((size_t)codes * R * K * sizeof(double));
  double * tr = (double *)malloc((size_t)codes * R * K * sizeof(double));
  tip_table(K, R, codes, tipmap, left_matrix, tl);
  tip_table(K, R, codes, tipmap, right_matrix, tr);
  if (parent_scaler)
    memset(parent_scaler, 0, sizeof(unsigned int) * ((attrib & ORC_RATE_SCALERS) ? (size_t)sites * R : sites));
  for (n = 0; n < sites; ++n)
  {
    const double * a = tl + (size_t)left_tipchars[n] * R * K;
    const double * b = tr + (size_t)right_tipchars[n] * R * K;
    double * p = parent_clv + (size_t)n * R * K;
    for (q = 0; q < R * K; ++q) p[q] = a[q] * b[q];
  }
  free(tl);
  free(tr);
}

/* ------------------------------------------------------------------------------------ */
/* log-likelihood                                                                        */
/* ------------------------------------------------------------------------------------ */
static double scale_minlh(unsigned int d)
{
  double f = 1.0;
  while (d--) f *= ORC_SCALE_THRESHOLD;
  return f;
}

/* per-rate scalers -> (site scaler = min over rates, capped residual per rate)
 * reference src/core_likelihood_avx.c:1136-1154 */
static unsigned int rate_residuals(const unsigned int * ps, const unsigned int * cs, unsigned int n,
                                   unsigned int R, unsigned int * resid)
{
  unsigned int i, mn = UINT_MAX;
  for (i = 0; i < R; ++i)
  {
    resid[i] = (ps ? ps[n * R + i] : 0) + (cs ? cs[n * R + i] : 0);
    if (resid[i] < mn) mn = resid[i];
  }
  for (i = 0; i < R; ++i)
  {
    resid[i] -= mn;
    if (resid[i] > ORC_MAXDIFF) resid[i] = ORC_MAXDIFF;
  }
  return mn;
}

static double finish_site(double term, unsigned int site_scalings, unsigned int weight)
{
  double lk = log(term);
  if (site_scalings) lk += site_scalings * log(ORC_SCALE_THRESHOLD);
  return lk * weight;
}

/* edge, both ends inner: reference src/core_likelihood_avx.c:1079-1266 (4x4),
 * src/core_likelihood_avx2.c:333-547 (generic) */
double orc_core_edge_loglikelihood_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                      const double * parent_clv, const unsigned int * parent_scaler,
                                      const double * child_clv, const unsigned int * child_scaler,
                                      const double * pmatrix, double * const * frequencies,
                                      const double * rate_weights,
                                      const unsigned int * pattern_weights,
                                      const double * invar_proportion, const int * invar_indices,
                                      const unsigned int * freqs_indices, double * persite_lnl,
                                      unsigned int attrib)
{
  const unsigned int K = states, R = rate_cats;
  const int per_rate = (attrib & ORC_RATE_SCALERS) != 0;
  unsigned int n, i, j, b;
  unsigned int resid[64];
  double logl = 0;
  for (n = 0; n < sites; ++n)
  {
    double terma = 0;
    unsigned int site_scalings;
    if (per_rate) site_scalings = rate_residuals(parent_scaler, child_scaler, n, R, resid);
    else site_scalings = (parent_scaler ? parent_scaler[n] : 0) + (child_scaler ? child_scaler[n] : 0);
    for (i = 0; i < R; ++i)
    {
      const double * f = frequencies[freqs_indices[i]];
      const double * p = parent_clv + ((size_t)n * R + i) * K;
      const double * c = child_clv + ((size_t)n * R + i) * K;
      const double * M = pmatrix + (size_t)i * K * K;
      double terma_r;
      if (K == 4)
      {
        double t[4];
        for (j = 0; j < 4; ++j)
          t[j] = (f[j] * hsum4(M[j * 4] * c[0], M[j * 4 + 1] * c[1], M[j * 4 + 2] * c[2], M[j * 4 + 3] * c[3])) * p[j];
        terma_r = hsum4(t[0], t[1], t[2], t[3]);
      }
      else if (K != 20)
      {
        /* reference src/core_likelihood.c:947-958 */
        terma_r = 0;
        for (j = 0; j < K; ++j) terma_r += p[j] * f[j] * row_dot(M + j * K, c, K, 0);
      }
      else
      {
        terma_r = 0;
        for (b = 0; b < K; b += 4)
        {
          double t[4];
          for (j = 0; j < 4; ++j) t[j] = (row_dot(M + (b + j) * K, c, K, 1) * f[b + j]) * p[b + j];
          terma_r += hsum4(t[0], t[1], t[2], t[3]);
        }
      }
      if (per_rate && resid[i] > 0) terma_r *= scale_minlh(resid[i]);
      /* the 4x4 AVX kernel skips non-positive rate terms (reference
       * src/core_likelihood_avx.c:1225); the generic AVX2 kernel does not */
      if (K != 4 || terma_r > 0.)
      {
        const double pinv = invar_proportion ? invar_proportion[freqs_indices[i]] : 0;
        if (pinv > 0)
        {
          const double inv_lk = (invar_indices[n] == -1) ? 0 : f[invar_indices[n]];
          terma += rate_weights[i] * (terma_r * (1 - pinv) + inv_lk * pinv);
        }
        else
          terma += terma_r * rate_weights[i];
      }
    }
    {
      const double lk = finish_site(terma, site_scalings, pattern_weights[n]);
      if (persite_lnl) persite_lnl[n] = lk;
      logl += lk;
    }
  }
  return logl;
}

/* edge, one end a pattern tip: reference src/core_likelihood_avx.c:191-406 (4x4),
 * src/core_likelihood_avx2.c:111-331 (20x20) */
double orc_core_edge_loglikelihood_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                      const double * parent_clv, const unsigned int * parent_scaler,
                                      const unsigned char * tipchars, const unsigned int * tipmap,
                                      unsigned int tipmap_size, const double * pmatrix,
                                      double * const * frequencies, const double * rate_weights,
                                      const unsigned int * pattern_weights,
                                      const double * invar_proportion, const int * invar_indices,
                                      const unsigned int * freqs_indices, double * persite_lnl,
                                      unsigned int attrib)
{
  const unsigned int K = states, R = rate_cats;
  const unsigned int codes = (K == 4) ? 16u : tipmap_size;
  const int per_rate = (attrib & ORC_RATE_SCALERS) != 0 && (K != 4 || parent_scaler);
  unsigned int n, i, j, code;
  unsigned int resid[64];
  double logl = 0;
  double * table = (double *)malloc((size_t)codes * R * K * sizeof(double));
  tip_table(K, R, codes, tipmap, pmatrix, table);
  /* fold pi in: DNA pi * sum (reference src/core_likelihood_avx.c:296-300), 20 states
   * sum * pi (reference src/core_likelihood_avx2.c:228) - the same product */
  for (code = 0; code < codes; ++code)
    for (i = 0; i < R; ++i)
      for (j = 0; j < K; ++j)
        table[((size_t)code * R + i) * K + j] *= frequencies[freqs_indices[i]][j];

  for (n = 0; n < sites; ++n)
  {
    double terma = 0;
    unsigned int site_scalings;
    const double * t = table + (size_t)tipchars[n] * R * K;
    if (per_rate) site_scalings = rate_residuals(parent_scaler, NULL, n, R, resid);
    else site_scalings = parent_scaler ? parent_scaler[n] : 0;
    for (i = 0; i < R; ++i)
    {
      const double * p = parent_clv + ((size_t)n * R + i) * K;
      double terma_r;
      if (K == 4)
        terma_r = hsum4(t[i * 4] * p[0], t[i * 4 + 1] * p[1], t[i * 4 + 2] * p[2], t[i * 4 + 3] * p[3]);
      else
        terma_r = row_dot(t + i * K, p, K, 1);
      if (per_rate && resid[i] > 0) terma_r *= scale_minlh(resid[i]);
      if (K != 4 || terma_r > 0.)
      {
        const double pinv = invar_proportion ? invar_proportion[freqs_indices[i]] : 0;
        if (pinv > 0)
        {
          /* the DNA kernel reads the invariant frequency from the LAST rate category's
           * vector (reference src/core_likelihood_avx.c:274, 370-371) */
          const double * f = frequencies[freqs_indices[K == 4 ? R - 1 : i]];
          const double inv_lk = (invar_indices[n] == -1) ? 0 : f[invar_indices[n]];
          terma += rate_weights[i] * (terma_r * (1 - pinv) + inv_lk * pinv);
        }
        else
          terma += terma_r * rate_weights[i];
      }
    }
    {
      const double lk = finish_site(terma, site_scalings, pattern_weights[n]);
      if (persite_lnl) persite_lnl[n] = lk;
      logl += lk;
    }
  }
  free(table);
  return logl;
}

/* root: reference src/core_likelihood_avx.c:113-189 (4x4), src/core_likelihood_avx2.c:25-109 */
double orc_core_root_loglikelihood(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                   const double * clv, const unsigned int * scaler,
                                   double * const * frequencies, const double * rate_weights,
                                   const unsigned int * pattern_weights,
                                   const double * invar_proportion, const int * invar_indices,
                                   const unsigned int * freqs_indices, double * persite_lnl,
                                   unsigned int attrib)
{
  const unsigned int K = states, R = rate_cats;
  unsigned int n, i;
  double logl = 0;
  (void)attrib;
  for (n = 0; n < sites; ++n)
  {
    double term = 0;
    for (i = 0; i < R; ++i)
    {
      const double * f = frequencies[freqs_indices[i]];
      const double * c = clv + ((size_t)n * R + i) * K;
      const double term_r = (K == 4) ? hsum4(f[0] * c[0], f[1] * c[1], f[2] * c[2], f[3] * c[3])
                                     : row_dot(f, c, K, 1);
      const double pinv = invar_proportion ? invar_proportion[freqs_indices[i]] : 0;
      if (pinv > 0)
      {
        const double inv_lk = (invar_indices[n] == -1) ? 0 : f[invar_indices[n]];
        term += rate_weights[i] * (term_r * (1 - pinv) + inv_lk * pinv);
      }
      else
        term += term_r * rate_weights[i];
    }
    {
      /* the scaler is indexed per site even in per-rate mode (reference
       * src/core_likelihood_avx.c:176-178) */
      const double lk = finish_site(term, scaler ? scaler[n] : 0, pattern_weights[n]);
      if (persite_lnl) persite_lnl[n] = lk;
      logl += lk;
    }
  }
  return logl;
}

/* ------------------------------------------------------------------------------------ */
/* sumtable and derivatives                                                              */
/* ------------------------------------------------------------------------------------ */
/* reference src/core_derivatives_avx.c:25-207 (4x4), src/core_derivatives_avx2.c:24-272 */
int orc_core_update_sumtable_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                const double * parent_clv, const double * child_clv,
                                const unsigned int * parent_scaler, const unsigned int * child_scaler,
                                double * const * eigenvecs, double * const * inv_eigenvecs,
                                double * const * freqs, double * sumtable, unsigned int attrib)
{
  const unsigned int K = states, R = rate_cats;
  const int per_rate = (attrib & ORC_RATE_SCALERS) != 0;
  unsigned int n, i, j, k;
  unsigned int resid[64];
  double * W = (double *)malloc((size_t)R * K * K * sizeof(double));
  for (i = 0; i < R; ++i)
    for (j = 0; j < K; ++j)
      for (k = 0; k < K; ++k) W[((size_t)i * K + j) * K + k] = inv_eigenvecs[i][k * K + j] * freqs[i][k];
  for (n = 0; n < sites; ++n)
  {
    if (per_rate) rate_residuals(parent_scaler, child_scaler, n, R, resid);
    for (i = 0; i < R; ++i)
    {
      const double * p = parent_clv + ((size_t)n * R + i) * K;
      const double * c = child_clv + ((size_t)n * R + i) * K;
      double * s = sumtable + ((size_t)n * R + i) * K;
      for (j = 0; j < K; ++j)
      {
        const double * w = W + ((size_t)i * K + j) * K;
        const double * v = eigenvecs[i] + j * K;
        double l, r;
        if (K == 4)
        {
          l = hsum4(w[0] * p[0], w[1] * p[1], w[2] * p[2], w[3] * p[3]);
          r = hsum4(v[0] * c[0], v[1] * c[1], v[2] * c[2], v[3] * c[3]);
        }
        else
        {
          l = row_dot(w, p, K, 1);
          r = row_dot(v, c, K, 1);
        }
        s[j] = l * r;
        if (per_rate && resid[i] > 0) s[j] *= scale_minlh(resid[i]);
      }
    }
  }
  free(W);
  return 1;
}

/* reference src/core_derivatives_avx.c:462-645 (4x4), src/core_derivatives_avx2.c:274-521 */
int orc_core_update_sumtable_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                const double * parent_clv, const unsigned char * left_tipchars,
                                const unsigned int * parent_scaler, double * const * eigenvecs,
                                double * const * inv_eigenvecs, double * const * freqs,
                                const unsigned int * tipmap, unsigned int tipmap_size,
                                double * sumtable, unsigned int attrib)
{
  const unsigned int K = states, R = rate_cats;
  const unsigned int codes = (K == 4) ? 16u : tipmap_size;
  const int per_rate = (attrib & ORC_RATE_SCALERS) != 0;
  unsigned int n, i, j, k, code;
  unsigned int resid[64];
  double * left = (double *)calloc((size_t)codes * R * K, sizeof(double));
  for (code = 0; code < codes; ++code)
  {
    const unsigned int state = (K == 4) ? code : tipmap[code];
    for (i = 0; i < R; ++i)
      for (j = 0; j < K; ++j)
      {
        double acc = 0.0;
        for (k = 0; k < K; ++k)
          if ((state >> k) & 1u)
            acc = (K == 4) ? acc + inv_eigenvecs[i][k * K + j] * freqs[i][k]
                           : fma(inv_eigenvecs[i][k * K + j], freqs[i][k], acc);
        left[((size_t)code * R + i) * K + j] = acc;
      }
  }
  for (n = 0; n < sites; ++n)
  {
    const double * lt = left + (size_t)left_tipchars[n] * R * K;
    if (per_rate) rate_residuals(parent_scaler, NULL, n, R, resid);
    for (i = 0; i < R; ++i)
    {
      const double * c = parent_clv + ((size_t)n * R + i) * K;
      double * s = sumtable + ((size_t)n * R + i) * K;
      for (j = 0; j < K; ++j)
      {
        const double * v = eigenvecs[i] + j * K;
        double r;
        if (K == 4)
        {
          /* sequential over the child states (reference src/core_derivatives_avx.c:611-620) */
          r = 0.0;
          for (k = 0; k < 4; ++k) r = r + v[k] * c[k];
        }
        else
          r = row_dot(v, c, K, 1);
        s[j] = lt[i * K + j] * r;
        if (per_rate && resid[i] > 0) s[j] *= scale_minlh(resid[i]);
      }
    }
  }
  free(left);
  return 1;
}

/* reference src/core_derivatives.c:501-732 + src/core_derivatives_avx2.c:523-800, including
 * the 4-sites-per-vector accumulation and the backwards tail loop */
int orc_core_likelihood_derivatives(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                    const double * rate_weights, const unsigned int * parent_scaler,
                                    const unsigned int * child_scaler, const int * invariant,
                                    const unsigned int * pattern_weights, double branch_length,
                                    const double * prop_invar, double * const * freqs,
                                    const double * rates, double * const * eigenvals,
                                    const double * sumtable, double * d_f, double * dd_f,
                                    unsigned int attrib)
{
  const unsigned int K = states, R = rate_cats;
  unsigned int n, i, j, b, l;
  int use_pinv = 0, eq_weights = 1;
  double * diagp = (double *)malloc((size_t)R * K * 4 * sizeof(double));
  double * site_lk = (double *)malloc((size_t)sites * 3 * sizeof(double));
  double vdf[4] = {0, 0, 0, 0}, vddf[4] = {0, 0, 0, 0};
  (void)parent_scaler; (void)child_scaler; (void)attrib;

  for (i = 0; i < R; ++i)
  {
    const double ki = rates[i] / (1.0 - prop_invar[i]);
    for (j = 0; j < K; ++j)
    {
      double * d = diagp + ((size_t)i * K + j) * 4;
      d[0] = exp(eigenvals[i][j] * ki * branch_length);
      d[1] = eigenvals[i][j] * ki * d[0];
      d[2] = eigenvals[i][j] * ki * eigenvals[i][j] * ki * d[0];
      d[3] = 0;
    }
    use_pinv |= (prop_invar[i] > 0);
    eq_weights &= (rate_weights[i] == rate_weights[0]);
  }

  for (n = 0; n < sites; ++n)
  {
    double sl[3] = {0, 0, 0};
    for (i = 0; i < R; ++i)
    {
      const double * s = sumtable + ((size_t)n * R + i) * K;
      double cat[3];
      unsigned int x;
      if (K == 4)
      {
        for (x = 0; x < 3; ++x)
        {
          cat[x] = 0.0;
          for (j = 0; j < 4; ++j) cat[x] = fma(s[j], diagp[((size_t)i * 4 + j) * 4 + x], cat[x]);
        }
      }
      else if (K != 20)
      {
        /* reference src/core_derivatives.c:472-480 */
        for (x = 0; x < 3; ++x)
        {
          cat[x] = 0.0;
          for (j = 0; j < K; ++j) cat[x] += s[j] * diagp[((size_t)i * K + j) * 4 + x];
        }
      }
      else
      {
        for (x = 0; x < 3; ++x)
        {
          double a[4];
          for (l = 0; l < 4; ++l) a[l] = s[l] * diagp[((size_t)i * K + l) * 4 + x];
          for (b = 4; b < K; b += 4)
            for (l = 0; l < 4; ++l) a[l] = fma(s[b + l], diagp[((size_t)i * K + b + l) * 4 + x], a[l]);
          cat[x] = hsum4(a[0], a[1], a[2], a[3]);
        }
      }
      if (use_pinv && prop_invar[i] > 0)
      {
        for (x = 0; x < 3; ++x) cat[x] = cat[x] * (1. - prop_invar[i]);
        if (invariant && invariant[n] != -1) cat[0] = cat[0] + freqs[i][invariant[n]] * prop_invar[i];
      }
      for (x = 0; x < 3; ++x)
        sl[x] = eq_weights ? sl[x] + cat[x] : fma(cat[x], rate_weights[i], sl[x]);
    }
    memcpy(site_lk + (size_t)n * 3, sl, sizeof(sl));
  }

  /* four adjacent sites per vector (reference src/core_derivatives_avx2.c:736-768) */
  {
    const unsigned int full = sites / 4 * 4;
    for (n = 0; n < full; n += 4)
    {
      const int unit = (pattern_weights[n] | pattern_weights[n + 1] | pattern_weights[n + 2] |
                        pattern_weights[n + 3]) == 1;
      for (l = 0; l < 4; ++l)
      {
        const double * sl = site_lk + (size_t)(n + l) * 3;
        const double recip = 1. / sl[0];
        const double d1 = sl[1] * recip;
        const double d2 = d1 * d1 - sl[2] * recip;
        if (unit)
        {
          vdf[l] = vdf[l] - d1;
          vddf[l] = vddf[l] + d2;
        }
        else
        {
          vdf[l] = fma(-d1, (double)pattern_weights[n + l], vdf[l]);
          vddf[l] = fma(d2, (double)pattern_weights[n + l], vddf[l]);
        }
      }
    }
    *d_f = *dd_f = 0.;
    /* remainder sites, last one first (reference src/core_derivatives_avx2.c:771-782) */
    for (n = sites; n > full; --n)
    {
      const double * sl = site_lk + (size_t)(n - 1) * 3;
      const double d1 = (-sl[1] / sl[0]);
      const double d2 = (d1 * d1 - (sl[2] / sl[0]));
      *d_f += pattern_weights[n - 1] * d1;
      *dd_f += pattern_weights[n - 1] * d2;
    }
    *d_f += vdf[0] + vdf[1] + vdf[2] + vdf[3];
    *dd_f += vddf[0] + vddf[1] + vddf[2] + vddf[3];
  }
  free(diagp);
  free(site_lk);
  return 1;
}
