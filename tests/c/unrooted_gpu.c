/*
 * unrooted_gpu.c - a plain C caller of the pll.h API on the GPU backend: the scenario of the
 * reference's examples/unrooted (4 taxa, 6 sites, GTR with unit rates, Gamma4 alpha=1, then
 * proportions of invariant sites 0.5 and 0.75; reference examples/unrooted/unrooted.c:32-212)
 * followed by the Newton branch-length optimisation of examples/newton (newton.c:31-100).
 * Written against include/pll.h only; the single difference to a libpll program is the
 * PLL_ATTRIB_ARCH_GPU flag (and pll_gpu_sync_* before peeking at device-resident arrays).
 *
 * Expected output (reference values): -33.387713, -34.550204, -36.830297, Newton -> 2.607098.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pll.h"
#include "pll_gpu.h"

static const char * seqs[4] = {"WAAAAB", "CACACD", "AGGACA", "CGTAGT"};

static double evaluate(pll_partition_t * p, const pll_operation_t * ops, const unsigned int * params,
                       const unsigned int * matrices, const double * lengths)
{
  pll_update_prob_matrices(p, params, matrices, lengths, 5);
  pll_update_partials(p, ops, 2);
  return pll_compute_edge_loglikelihood(p, 4, 0, 5, 1, 4, params, NULL);
}

int main(void)
{
  unsigned int i;
  pll_partition_t * p = pll_partition_create(4, 2, 4, 6, 1, 5, 4, 2,
                                             PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP);
  if (!p)
  {
    fprintf(stderr, "pll_partition_create failed (%d): %s\n", pll_errno, pll_errmsg);
    return 2;
  }
  double lengths[5] = {0.2, 0.4, 0.3, 0.5, 0.6};
  double freqs[4] = {0.17, 0.19, 0.25, 0.39};
  double subst[6] = {1, 1, 1, 1, 1, 1};
  unsigned int matrices[5] = {0, 1, 2, 3, 4};
  unsigned int params[4] = {0, 0, 0, 0};
  double rates[4];
  pll_compute_gamma_cats(1.0, 4, rates, PLL_GAMMA_RATES_MEAN);
  pll_set_frequencies(p, 0, freqs);
  pll_set_subst_params(p, 0, subst);
  pll_set_category_rates(p, rates);
  for (i = 0; i < 4; ++i)
    if (!pll_set_tip_states(p, i, pll_map_nt, seqs[i]))
    {
      fprintf(stderr, "pll_set_tip_states failed: %s\n", pll_errmsg);
      return 2;
    }

  pll_operation_t ops[2] = {
      {4, 0, 0, 0, PLL_SCALE_BUFFER_NONE, 1, 1, PLL_SCALE_BUFFER_NONE},
      {5, 1, 2, 2, PLL_SCALE_BUFFER_NONE, 3, 3, PLL_SCALE_BUFFER_NONE},
  };

  printf("Log-L: %f\n", evaluate(p, ops, params, matrices, lengths));
  pll_update_invariant_sites(p);
  pll_update_invariant_sites_proportion(p, 0, 0.5);
  printf("Log-L (Inv+Gamma 0.5): %f\n", evaluate(p, ops, params, matrices, lengths));
  pll_update_invariant_sites_proportion(p, 0, 0.75);
  printf("Log-L (Inv+Gamma 0.75): %f\n", evaluate(p, ops, params, matrices, lengths));

  /* device-resident arrays are visible after an explicit sync */
  pll_gpu_sync_clv(p, 4);
  pll_gpu_sync_scaler(p, 0);
  printf("CLV 4, site 0, rate 0: %.7f %.7f %.7f %.7f\n", p->clv[4][0], p->clv[4][1], p->clv[4][2],
         p->clv[4][3]);

  /* Newton on the edge (4,5), back to the model without invariant sites */
  pll_update_invariant_sites_proportion(p, 0, 0.0);
  evaluate(p, ops, params, matrices, lengths);
  double * sumtable = (double *)pll_aligned_alloc(p->sites * p->rate_cats * p->states_padded * sizeof(double),
                                                  p->alignment);
  pll_update_sumtable(p, 4, 5, 0, 1, params, sumtable);
  double len = lengths[4], d1, d2;
  int it;
  for (it = 0; it < 32; ++it)
  {
    if (!pll_compute_likelihood_derivatives(p, 0, 1, len, params, sumtable, &d1, &d2))
    {
      fprintf(stderr, "derivatives failed: %s\n", pll_errmsg);
      return 2;
    }
    if (fabs(d1) < 1e-5) break;
    len -= d1 / d2;
  }
  printf("Newton: %f after %d iterations\n", len, it + 1);
  pll_aligned_free(sumtable);
  pll_partition_destroy(p);
  return 0;
}
