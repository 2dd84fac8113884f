/*
 * pll_models.c - substitution-model parameters, eigendecomposition, P-matrix driver and
 * invariant sites for the GPU backend.
 *
 * Mirrors reference src/models.c:
 *   pll_update_eigen            :251-331  (symmetrised rate matrix :182-249, Householder
 *                                          tridiagonalisation :100-178, implicit QL :24-97)
 *   pll_update_prob_matrices    :333-364  -> plg_update_pmatrix (device)
 *   setters                     :366-400
 *   invariant sites             :402-647  (the tip scan runs on the device)
 *
 * The eigendecomposition stays on the host (4x4 or 20x20, microseconds); it uses the same
 * classical algorithm pair as the reference - Householder reduction to tridiagonal form
 * followed by QL iterations with implicit shifts (Numerical Recipes "tred2"/"tqli", EISPACK
 * lineage) - written here for 0-based row-major storage, with the same operation order so the
 * eigenvectors agree with the reference to the last bit and the device P-matrices differ
 * only through expm1.
 */
#include <assert.h>

#include "pll_host.h"

/* ------------------------------------------------------------------------------------ */
/* symmetric eigensolver                                                                 */
/* ------------------------------------------------------------------------------------ */

/* Householder reduction of the symmetric n x n matrix a (row-major, a[r*n+c]) to
 * tridiagonal form.  On return d = diagonal, e = sub-diagonal (e[0] = 0) and a holds the
 * accumulated orthogonal transformation. */
static void householder_tridiag(double * a, unsigned int n, double * d, double * e)
{
#define A(r, c) a[(size_t)(r) * n + (c)]
  unsigned int i, j, k;
  for (i = n - 1; i >= 1; --i)
  {
    const unsigned int l = i; /* number of leading elements of column i that are touched */
    double h = 0.0, scale = 0.0;
    if (l > 1)
    {
      for (k = 0; k < l; ++k) scale += fabs(A(k, i));
      if (scale == 0.0)
        e[i] = A(l - 1, i);
      else
      {
        for (k = 0; k < l; ++k)
        {
          A(k, i) /= scale;
          h += A(k, i) * A(k, i);
        }
        double f = A(l - 1, i);
        double g = (f > 0) ? -sqrt(h) : sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        A(l - 1, i) = f - g;
        f = 0.0;
        for (j = 0; j < l; ++j)
        {
          A(i, j) = A(j, i) / h;
          g = 0.0;
          for (k = 0; k <= j; ++k) g += A(k, j) * A(k, i);
          for (k = j + 1; k < l; ++k) g += A(j, k) * A(k, i);
          e[j] = g / h;
          f += e[j] * A(j, i);
        }
        const double hh = f / (h + h);
        for (j = 0; j < l; ++j)
        {
          f = A(j, i);
          g = e[j] - hh * f;
          e[j] = g;
          for (k = 0; k <= j; ++k) A(k, j) -= (f * e[k] + g * A(k, i));
        }
      }
    }
    else
      e[i] = A(l - 1, i);
    d[i] = h;
  }
  d[0] = 0.0;
  e[0] = 0.0;

  /* accumulate the transformation matrices */
  for (i = 0; i < n; ++i)
  {
    if (d[i] != 0.0)
    {
      for (j = 0; j < i; ++j)
      {
        double g = 0.0;
        for (k = 0; k < i; ++k) g += A(k, i) * A(j, k);
        for (k = 0; k < i; ++k) A(j, k) -= g * A(i, k);
      }
    }
    d[i] = A(i, i);
    A(i, i) = 1.0;
    for (j = 0; j < i; ++j) A(i, j) = A(j, i) = 0.0;
  }
#undef A
}

/* QL with implicit shifts on the tridiagonal (d, e); z (row-major n x n) enters as the
 * Householder transformation and leaves with eigenvector k in ROW k.  Returns 0 if an
 * eigenvalue fails to converge in 30 iterations. */
static int ql_implicit(double * d, double * e, unsigned int n, double * z)
{
#define Z(r, c) z[(size_t)(r) * n + (c)]
  unsigned int m, l, iter, k;
  int i;
  for (l = 1; l < n; ++l) e[l - 1] = e[l];
  e[n - 1] = 0.0;

  for (l = 0; l < n; ++l)
  {
    iter = 0;
    do
    {
      for (m = l; m + 1 < n; ++m)
      {
        const double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) + dd == dd) break;
      }
      if (m != l)
      {
        if (iter++ >= 30) return 0;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = sqrt((g * g) + 1.0);
        g = d[m] - d[l] + e[l] / (g + ((g < 0) ? -fabs(r) : fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        for (i = (int)m - 1; i >= (int)l; --i)
        {
          double f = s * e[i];
          const double b = c * e[i];
          if (fabs(f) >= fabs(g))
          {
            c = g / f;
            r = sqrt((c * c) + 1.0);
            e[i + 1] = f * r;
            c *= (s = 1.0 / r);
          }
          else
          {
            s = f / g;
            r = sqrt((s * s) + 1.0);
            e[i + 1] = g * r;
            s *= (c = 1.0 / r);
          }
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          for (k = 0; k < n; ++k)
          {
            f = Z(i + 1, k);
            Z(i + 1, k) = s * Z(i, k) + c * f;
            Z(i, k) = c * Z(i, k) - s * f;
          }
        }
        d[l] = d[l] - p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return 1;
#undef Z
}

/* sqrt(pi) Q sqrt(pi)^-1, symmetric, normalised to a mean substitution rate of 1
 * (reference src/models.c:182-249) */
static double * symmetric_ratematrix(const double * params, const double * freqs, unsigned int n)
{
  const unsigned int nparams = (n * n - n) / 2;
  unsigned int i, j, k = 0;
  double * q = (double *)calloc((size_t)n * n, sizeof(double));
  double * prm = (double *)malloc(sizeof(double) * (nparams ? nparams : 1));
  if (!q || !prm)
  {
    free(q);
    free(prm);
    return NULL;
  }
  memcpy(prm, params, nparams * sizeof(double));
  if (nparams && prm[nparams - 1] > 0.0)
    for (i = 0; i < nparams; ++i) prm[i] /= prm[nparams - 1];

  for (i = 0; i < n; ++i)
    for (j = i + 1; j < n; ++j)
    {
      const double factor = prm[k++];
      q[i * n + j] = q[j * n + i] = factor * sqrt(freqs[i] * freqs[j]);
      q[i * n + i] -= factor * freqs[j];
      q[j * n + j] -= factor * freqs[i];
    }

  double mean = 0;
  for (i = 0; i < n; ++i) mean += freqs[i] * (-q[i * n + i]);
  for (i = 0; i < n * n; ++i) q[i] /= mean;
  free(prm);
  return q;
}

PLL_EXPORT int pll_update_eigen(pll_partition_t * partition, unsigned int params_index)
{
  pll_partition_t * p = partition;
  const unsigned int n = p->states, np = p->states_padded;
  unsigned int i, j;
  double * evecs = p->eigenvecs[params_index];
  double * ievecs = p->inv_eigenvecs[params_index];
  double * evals = p->eigenvals[params_index];
  const double * freqs = p->frequencies[params_index];

  double * a = symmetric_ratematrix(p->subst_params[params_index], freqs, n);
  double * d = (double *)malloc(n * sizeof(double));
  double * e = (double *)malloc(n * sizeof(double));
  if (!a || !d || !e)
  {
    free(a);
    free(d);
    free(e);
    return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
  }
  householder_tridiag(a, n, d, e);
  if (!ql_implicit(d, e, n, a))
  {
    free(a);
    free(d);
    free(e);
    return pll_fail(PLL_ERROR_PARAM_INVALID, "Eigendecomposition did not converge.");
  }

  /* eigenvector k sits in row k of a; inverse = transpose (orthonormal), then fold sqrt(pi)
   * back in: V <- V * sqrt(pi) (columns), V^-1 <- sqrt(pi)^-1 * V^T (rows)
   * (reference src/models.c:293-320) */
  for (i = 0; i < n; ++i)
  {
    memcpy(evecs + (size_t)i * np, a + (size_t)i * n, n * sizeof(double));
    evals[i] = d[i];
  }
  for (i = 0; i < n; ++i)
    for (j = 0; j < n; ++j) ievecs[(size_t)i * np + j] = evecs[(size_t)j * np + i];
  for (i = 0; i < n; ++i)
    for (j = 0; j < n; ++j) ievecs[(size_t)i * np + j] /= sqrt(freqs[i]);
  for (i = 0; i < n; ++i)
    for (j = 0; j < n; ++j) evecs[(size_t)i * np + j] *= sqrt(freqs[j]);

  p->eigen_decomp_valid[params_index] = 1;
  free(a);
  free(d);
  free(e);
  return PLL_SUCCESS;
}

/* ------------------------------------------------------------------------------------ */
/* P-matrices                                                                            */
/* ------------------------------------------------------------------------------------ */
PLL_EXPORT int pll_update_prob_matrices(pll_partition_t * partition,
                                        const unsigned int * params_indices,
                                        const unsigned int * matrix_indices,
                                        const double * branch_lengths,
                                        unsigned int count)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  const unsigned int R = p->rate_cats, K = p->states, Kp = p->states_padded;
  unsigned int n;

  for (n = 0; n < R; ++n)
    if (!p->eigen_decomp_valid[params_indices[n]])
      if (!pll_update_eigen(p, params_indices[n])) return PLL_FAILURE;

  /* gather the per-rate model the way the reference kernels index it
   * (reference src/core_pmatrix_avx.c:97-100) */
  double * buf = (double *)malloc(((size_t)R * Kp + 2 * (size_t)R * K * Kp + R) * sizeof(double));
  if (!buf) return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
  double * evals = buf;
  double * evecs = evals + (size_t)R * Kp;
  double * ievecs = evecs + (size_t)R * K * Kp;
  double * pinv = ievecs + (size_t)R * K * Kp;
  for (n = 0; n < R; ++n)
  {
    const unsigned int pi = params_indices[n];
    memcpy(evals + (size_t)n * Kp, p->eigenvals[pi], Kp * sizeof(double));
    memcpy(evecs + (size_t)n * K * Kp, p->eigenvecs[pi], (size_t)K * Kp * sizeof(double));
    memcpy(ievecs + (size_t)n * K * Kp, p->inv_eigenvecs[pi], (size_t)K * Kp * sizeof(double));
    pinv[n] = p->prop_invar[pi];
  }
  int rc = pllg_dev_update_pmatrix(g, matrix_indices, branch_lengths, count, p->rates, pinv,
                              evals, evecs, ievecs);
  free(buf);
  return rc ? pllg_fail(rc, "pll_update_prob_matrices") : PLL_SUCCESS;
}

/* ------------------------------------------------------------------------------------ */
/* setters (plain host copies, reference src/models.c:366-400)                           */
/* ------------------------------------------------------------------------------------ */
PLL_EXPORT void pll_set_frequencies(pll_partition_t * partition, unsigned int freqs_index,
                                    const double * frequencies)
{
  memcpy(partition->frequencies[freqs_index], frequencies, partition->states * sizeof(double));
  partition->eigen_decomp_valid[freqs_index] = 0;
}

PLL_EXPORT void pll_set_category_rates(pll_partition_t * partition, const double * rates)
{
  memcpy(partition->rates, rates, partition->rate_cats * sizeof(double));
}

PLL_EXPORT void pll_set_category_weights(pll_partition_t * partition, const double * rate_weights)
{
  memcpy(partition->rate_weights, rate_weights, partition->rate_cats * sizeof(double));
}

PLL_EXPORT void pll_set_subst_params(pll_partition_t * partition, unsigned int params_index,
                                     const double * params)
{
  const unsigned int count = partition->states * (partition->states - 1) / 2;
  memcpy(partition->subst_params[params_index], params, count * sizeof(double));
  partition->eigen_decomp_valid[params_index] = 0;
}

/* ------------------------------------------------------------------------------------ */
/* invariant sites                                                                       */
/* ------------------------------------------------------------------------------------ */
PLL_EXPORT int pll_update_invariant_sites(pll_partition_t * partition)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  if (!p->invariant && !(p->invariant = (int *)malloc((size_t)g->sites_alloc * sizeof(int))))
    return pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate charmap for invariant sites array.");
  int rc = pllg_dev_update_invariant(g, p->invariant);
  return rc ? pllg_fail(rc, "pll_update_invariant_sites") : PLL_SUCCESS;
}

PLL_EXPORT int pll_update_invariant_sites_proportion(pll_partition_t * partition,
                                                     unsigned int params_index,
                                                     double prop_invar)
{
  if (prop_invar != 0.0 && (partition->attributes & PLL_ATTRIB_AB_MASK))
    return pll_fail(PLL_ERROR_INVAR_INCOMPAT,
                    "Invariant sites are not compatible with asc bias correction");
  if (prop_invar < 0 || prop_invar >= 1)
    return pll_fail(PLL_ERROR_INVAR_PROPORTION, "Invalid proportion of invariant sites (%f)",
                    prop_invar);
  if (params_index > partition->rate_matrices)
    return pll_fail(PLL_ERROR_INVAR_PARAMINDEX, "Invalid params index (%d)", params_index);

  if (prop_invar > 0.0 && !partition->invariant)
    if (!pll_update_invariant_sites(partition))
      return pll_fail(PLL_ERROR_INVAR_NONEFOUND, "No invariant sites found");

  partition->prop_invar[params_index] = prop_invar;
  return PLL_SUCCESS;
}

PLL_EXPORT unsigned int pll_count_invariant_sites(pll_partition_t * partition,
                                                  unsigned int * state_inv_count)
{
  /* the reference scans the tips when partition->invariant is absent
   * (reference src/models.c:495-554); here the device scan fills a temporary index */
  pll_partition_t * p = partition;
  unsigned int i, count = 0;
  int * inv = p->invariant;
  int * tmp = NULL;
  if (state_inv_count) memset(state_inv_count, 0, p->states * sizeof(unsigned int));
  if (!inv)
  {
    pllg_partition_t * g = pllg_from(partition);
    if (!g || !(tmp = (int *)malloc((size_t)g->sites_alloc * sizeof(int)))) return 0;
    if (pllg_dev_update_invariant(g, tmp))
    {
      free(tmp);
      return 0;
    }
    inv = tmp;
  }
  for (i = 0; i < p->sites; ++i)
    if (inv[i] > -1)
    {
      assert((unsigned int)inv[i] < p->states);
      count += p->pattern_weights[i];
      if (state_inv_count) state_inv_count[inv[i]]++;
    }
  free(tmp);
  return count;
}
