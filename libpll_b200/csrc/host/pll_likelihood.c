/*
 * pll_likelihood.c - CLV updates and log-likelihood entry points (host wrappers).
 *
 * Mirrors reference src/partials.c:177-213 (pll_update_partials) and
 * src/likelihood.c:121-168, 478-513 (root / edge log-likelihood): unpack the partition and
 * hand the whole request to the device layer.  The tip/inner case analysis the reference
 * does here (src/partials.c:187-206, src/likelihood.c:489-501) lives in plg_update_partials
 * / plg_edge_loglikelihood, next to the batching it feeds.
 */
#include "pll_host.h"

PLL_EXPORT void pll_update_partials(pll_partition_t * partition,
                                    const pll_operation_t * operations,
                                    unsigned int count)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g)
  {
    pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
    return;
  }
  int rc = pllg_dev_update_partials(g, operations, count);
  if (rc)
  {
    pllg_fail(rc, "pll_update_partials");
    return;
  }
  /* PLL_GPU_MIRROR=1: callers written for the reference that read partition->clv[i] /
   * ->scale_buffer[i] directly (e.g. reference test/src/scaling.c:84-113) find every array this
   * call produced downloaded into its host mirror.  A compatibility mode: it moves each result
   * over PCIe and waits for it. */
  if (pll_gpu_mirror_mode())
    for (unsigned int i = 0; i < count; ++i)
    {
      if (!pll_gpu_sync_clv(partition, operations[i].parent_clv_index)) return;
      if (operations[i].parent_scaler_index != PLL_SCALE_BUFFER_NONE &&
          !pll_gpu_sync_scaler(partition, (unsigned int)operations[i].parent_scaler_index))
        return;
    }
}

/* per-rate frequency vectors / invariant proportions selected by freqs_indices, laid out
 * [rate][states_padded] and [rate] */
static double * gather_freqs(const pll_partition_t * p, const unsigned int * freqs_indices,
                             double ** pinv_out)
{
  const unsigned int R = p->rate_cats, Kp = p->states_padded;
  double * buf = (double *)calloc((size_t)R * Kp + R, sizeof(double));
  unsigned int i;
  if (!buf) return NULL;
  for (i = 0; i < R; ++i)
  {
    memcpy(buf + (size_t)i * Kp, p->frequencies[freqs_indices[i]], Kp * sizeof(double));
    buf[(size_t)R * Kp + i] = p->prop_invar[freqs_indices[i]];
  }
  *pinv_out = buf + (size_t)R * Kp;
  return buf;
}

PLL_EXPORT double pll_compute_root_loglikelihood(pll_partition_t * partition,
                                                 unsigned int clv_index,
                                                 int scaler_index,
                                                 const unsigned int * freqs_indices,
                                                 double * persite_lnl)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g)
  {
    pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
    return -INFINITY;
  }
  double * pinv = NULL;
  double * freqs = gather_freqs(&g->pub, freqs_indices, &pinv);
  if (!freqs)
  {
    pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return -INFINITY;
  }
  double logl = -INFINITY;
  int rc = pllg_dev_root_loglikelihood(g, clv_index, scaler_index, freqs, g->pub.rate_weights,
                                  pinv, persite_lnl, &logl);
  free(freqs);
  if (rc)
  {
    pllg_fail(rc, "pll_compute_root_loglikelihood");
    return -INFINITY;
  }
  if (g->pub.attributes & PLL_ATTRIB_AB_MASK)
    logl += pllg_asc_root(g, clv_index, scaler_index, freqs_indices);
  return logl;
}

PLL_EXPORT double pll_compute_edge_loglikelihood(pll_partition_t * partition,
                                                 unsigned int parent_clv_index,
                                                 int parent_scaler_index,
                                                 unsigned int child_clv_index,
                                                 int child_scaler_index,
                                                 unsigned int matrix_index,
                                                 const unsigned int * freqs_indices,
                                                 double * persite_lnl)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g)
  {
    pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
    return -INFINITY;
  }
  double * pinv = NULL;
  double * freqs = gather_freqs(&g->pub, freqs_indices, &pinv);
  if (!freqs)
  {
    pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return -INFINITY;
  }
  double logl = -INFINITY;
  int rc = pllg_dev_edge_loglikelihood(g, parent_clv_index, parent_scaler_index, child_clv_index,
                                  child_scaler_index, matrix_index, freqs, g->pub.rate_weights,
                                  pinv, persite_lnl, &logl);
  free(freqs);
  if (rc)
  {
    pllg_fail(rc, "pll_compute_edge_loglikelihood");
    return -INFINITY;
  }
  if (g->pub.attributes & PLL_ATTRIB_AB_MASK)
    logl += pllg_asc_edge(g, parent_clv_index, parent_scaler_index, child_clv_index, child_scaler_index,
                          matrix_index, freqs_indices);
  return logl;
}
