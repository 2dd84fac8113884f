/* Developer tool: latency of pll_compute_likelihood_derivatives / pll_compute_edge_loglikelihood
 * / pll_update_sumtable called from C (no ctypes in the way): DNA, GTR+G4, a ladder tree.
 *   gcc -O2 -Iinclude tools/newton_c.c -o tools/newton_c -Llibpll_b200 -lpll_b200 -Wl,-rpath,'$ORIGIN/../libpll_b200' -lm */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "pll.h"
#include "pll_gpu.h"

static double now_us(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}

int main(int argc, char ** argv)
{
  const unsigned int tips = argc > 1 ? atoi(argv[1]) : 64, sites = argc > 2 ? atoi(argv[2]) : 1000000;
  const int devices = argc > 3 ? atoi(argv[3]) : 1;
  if (devices > 1) pll_gpu_set_devices(devices);
  pll_partition_t * p = pll_partition_create(tips, tips - 2, 4, sites, 1, 2 * tips - 2, 4, tips - 2,
                                             PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP);
  if (!p) { fprintf(stderr, "create: %s\n", pll_errmsg); return 1; }
  const double freqs[4] = {0.3, 0.2, 0.25, 0.25}, subst[6] = {1.2, 3.1, 0.9, 1.1, 3.3, 1.0};
  double rates[4];
  pll_compute_gamma_cats(0.5, 4, rates, PLL_GAMMA_RATES_MEAN);
  pll_set_frequencies(p, 0, freqs);
  pll_set_subst_params(p, 0, subst);
  pll_set_category_rates(p, rates);
  for (unsigned int t = 0; t < tips; ++t)
    if (!pll_gpu_generate_tip_states(p, t, 43, 0)) { fprintf(stderr, "tips: %s\n", pll_errmsg); return 1; }
  unsigned int params[4] = {0, 0, 0, 0};
  unsigned int * mi = malloc(sizeof(unsigned int) * (2 * tips - 2));
  double * bl = malloc(sizeof(double) * (2 * tips - 2));
  for (unsigned int i = 0; i < 2 * tips - 2; ++i) { mi[i] = i; bl[i] = 0.05 + 0.001 * (i % 50); }
  pll_update_prob_matrices(p, params, mi, bl, 2 * tips - 2);
  pll_operation_t * ops = calloc(tips - 2, sizeof(pll_operation_t));
  unsigned int prev = 0;
  for (unsigned int k = 0; k < tips - 2; ++k)
  {
    ops[k].parent_clv_index = tips + k; ops[k].parent_scaler_index = (int)k;
    ops[k].child1_clv_index = prev; ops[k].child1_matrix_index = prev;
    ops[k].child1_scaler_index = prev >= tips ? (int)(prev - tips) : PLL_SCALE_BUFFER_NONE;
    ops[k].child2_clv_index = k + 1; ops[k].child2_matrix_index = k + 1; ops[k].child2_scaler_index = PLL_SCALE_BUFFER_NONE;
    prev = tips + k;
  }
  pll_update_partials(p, ops, tips - 2);
  const unsigned int pa = prev, ch = prev - 1; /* an inner-inner edge */
  const int sa = (int)(pa - tips), sb = (int)(ch - tips);
  double lnl = pll_compute_edge_loglikelihood(p, pa, sa, ch, sb, ch, params, NULL);
  printf("lnL %.6f (%u tips x %u patterns, %d device slice(s))\n", lnl, tips, sites, pll_gpu_partition_devices(p));
  double * key = pll_aligned_alloc(64, 64);
  const int N = 2000;
  double t0 = now_us();
  for (int i = 0; i < 50; ++i) pll_update_sumtable(p, pa, ch, sa, sb, params, key);
  pll_gpu_synchronize(p);
  printf("pll_update_sumtable            %7.1f us per call\n", (now_us() - t0) / 50);
  double d1 = 0, d2 = 0;
  for (int i = 0; i < 20; ++i) pll_compute_likelihood_derivatives(p, sa, sb, 0.1, params, key, &d1, &d2);
  t0 = now_us();
  for (int i = 0; i < N; ++i) pll_compute_likelihood_derivatives(p, sa, sb, 0.05 + 1e-5 * i, params, key, &d1, &d2);
  printf("pll_compute_likelihood_derivatives %7.1f us per call   (d1 %.3f d2 %.3f)\n", (now_us() - t0) / N, d1, d2);
  t0 = now_us();
  for (int i = 0; i < N; ++i) lnl = pll_compute_edge_loglikelihood(p, pa, sa, ch, sb, ch, params, NULL);
  printf("pll_compute_edge_loglikelihood     %7.1f us per call\n", (now_us() - t0) / N);
  /* a Newton loop as in reference examples/newton/newton.c:64-93, 32 iterations */
  double len = 0.1;
  t0 = now_us();
  for (int i = 0; i < 32; ++i)
  {
    pll_compute_likelihood_derivatives(p, sa, sb, len, params, key, &d1, &d2);
    len -= d1 / d2;
    if (len < 1e-6) len = 1e-6;
  }
  printf("Newton, 32 iterations              %7.3f ms   (t = %.6f)\n", (now_us() - t0) * 1e-3, len);
  pll_partition_destroy(p);
  return 0;
}
