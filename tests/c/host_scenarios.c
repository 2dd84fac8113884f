/*
 * host_scenarios.c - drives the whole pll.h host layer (libpll_b200/csrc/host/*.c) against the
 * byte-touching null device (tests/c/null_device.c) under AddressSanitizer + UBSan: every
 * combination of alphabet size, category count, tip representation, scaler mode, ascertainment
 * bias type and number of pattern slices goes through create -> model -> tips -> P-matrices ->
 * traversal -> log-likelihoods -> sumtable / derivatives -> mirrors -> destroy.  Numbers mean
 * nothing here (the device is a stub); what is checked is that the wrappers hand the device layer
 * buffers of the extents its contract states, survive their error paths and release everything.
 * Built and run by tests/test_sanitizers_cpu.py.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pll.h"
#include "pll_gpu.h"

int null_device_live_contexts(void);
void null_device_fail_create_in(int n);
void null_device_fail_calls(int on);

#define CHECK(cond)                                                                          \
  do                                                                                         \
  {                                                                                          \
    if (!(cond))                                                                             \
    {                                                                                        \
      fprintf(stderr, "%s:%d: CHECK(%s) failed; pll_errno=%d (%s)\n", __FILE__, __LINE__, #cond, \
              pll_errno, pll_errmsg);                                                        \
      exit(3);                                                                               \
    }                                                                                        \
  } while (0)

static unsigned int generic_map[256];
static const unsigned int * map_for(unsigned int states, char * alphabet)
{
  if (states == 4) { strcpy(alphabet, "ACGT"); return pll_map_nt; }
  if (states == 20) { strcpy(alphabet, "ARNDCQEGHILKMFPSTWYV"); return pll_map_aa; }
  if (states == 2) { strcpy(alphabet, "01"); return pll_map_bin; }
  memset(generic_map, 0, sizeof(generic_map));
  for (unsigned int i = 0; i < states; ++i)
  {
    alphabet[i] = (char)('A' + i);
    generic_map['A' + i] = 1u << i;
    generic_map['a' + i] = 1u << i;
  }
  alphabet[states] = 0;
  generic_map['-'] = (states == 32) ? 0xFFFFFFFFu : ((1u << states) - 1);
  return generic_map;
}

static unsigned long scenario(unsigned int states, unsigned int cats, int pattern_tip, int rate_scalers,
                              unsigned int ab, int slices)
{
  const unsigned int tips = 5, inner = 3, sites = 130, matrices = 7;
  unsigned int attrs = PLL_ATTRIB_ARCH_GPU | (pattern_tip ? PLL_ATTRIB_PATTERN_TIP : 0) |
                       (rate_scalers ? PLL_ATTRIB_RATE_SCALERS : 0) | ab;
  char alphabet[64];
  const unsigned int * map = map_for(states, alphabet);

  CHECK(pll_gpu_set_devices(slices));
  pll_partition_t * p = pll_partition_create(tips, inner, states, sites, 2, matrices, cats, inner, attrs);
  pll_gpu_set_devices(0);
  CHECK(p != NULL);
  CHECK(pll_gpu_partition_devices(p) == (slices > 1 ? 3 : 1));   /* 130 patterns (+ states with AB): 64 + 64 + rest */
  CHECK(null_device_live_contexts() == pll_gpu_partition_devices(p));

  /* model */
  double * rates = (double *)malloc(cats * sizeof(double));
  CHECK(pll_compute_gamma_cats(0.7, cats, rates, PLL_GAMMA_RATES_MEAN));
  pll_set_category_rates(p, rates);
  double * w = (double *)malloc(cats * sizeof(double));
  for (unsigned int i = 0; i < cats; ++i) w[i] = 1.0 / cats;
  pll_set_category_weights(p, w);
  const unsigned int n_subst = states * (states - 1) / 2;
  double * subst = (double *)malloc(n_subst * sizeof(double));
  double * freqs = (double *)malloc(states * sizeof(double));
  for (unsigned int m = 0; m < 2; ++m)
  {
    for (unsigned int i = 0; i < n_subst; ++i) subst[i] = 0.5 + ((i * 7 + m) % 5) * 0.4;
    double sum = 0;
    for (unsigned int i = 0; i < states; ++i) sum += freqs[i] = 1.0 + ((i + m) % 3);
    for (unsigned int i = 0; i < states; ++i) freqs[i] /= sum;
    pll_set_subst_params(p, m, subst);
    pll_set_frequencies(p, m, freqs);
  }
  unsigned int * params = (unsigned int *)malloc(cats * sizeof(unsigned int));
  for (unsigned int i = 0; i < cats; ++i) params[i] = i & 1;

  /* tips: valid sequences, then an illegal character */
  char * seq = (char *)malloc(sites + 1);
  for (unsigned int t = 0; t < tips; ++t)
  {
    for (unsigned int i = 0; i < sites; ++i) seq[i] = (i % 11 == 10) ? '-' : alphabet[(i * (t + 1) + t) % states];
    seq[sites] = 0;
    CHECK(pll_set_tip_states(p, t, map, seq));
  }
  seq[17] = '!';
  CHECK(!pll_set_tip_states(p, 0, map, seq) && pll_errno == PLL_ERROR_TIPDATA_ILLEGALSTATE);
  seq[17] = alphabet[0];
  CHECK(pll_set_tip_states(p, 0, map, seq));
  CHECK(!pll_set_tip_states(p, tips, map, seq));
  if (!pattern_tip)
  {
    double * clv = (double *)calloc((size_t)sites * states, sizeof(double));
    for (unsigned int i = 0; i < sites; ++i) clv[(size_t)i * states + i % states] = 1.0;
    CHECK(pll_set_tip_clv(p, 1, clv, 0));
    free(clv);
    double * padded = (double *)calloc((size_t)sites * p->states_padded, sizeof(double));
    CHECK(pll_set_tip_clv(p, 2, padded, 1));
    free(padded);
  }
  else
  {
    double dummy[64] = {0};
    CHECK(!pll_set_tip_clv(p, 1, dummy, 0) && pll_errno == PLL_ERROR_TIPDATA_ILLEGALFUNCTION);
  }
  unsigned int * weights = (unsigned int *)malloc(sites * sizeof(unsigned int));
  for (unsigned int i = 0; i < sites; ++i) weights[i] = 1 + i % 3;
  pll_set_pattern_weights(p, weights);
  if (ab)
  {
    unsigned int * sw = (unsigned int *)malloc(states * sizeof(unsigned int));
    for (unsigned int i = 0; i < states; ++i) sw[i] = 1 + i;
    pll_set_asc_state_weights(p, sw);
    free(sw);
    CHECK(pll_set_asc_bias_type(p, (int)ab));
    CHECK(!pll_set_asc_bias_type(p, 1 << 9));
    CHECK(!pll_update_invariant_sites_proportion(p, 0, 0.2));
  }
  else
  {
    CHECK(!pll_set_asc_bias_type(p, PLL_ATTRIB_AB_LEWIS));
    CHECK(pll_update_invariant_sites(p));
    CHECK(pll_update_invariant_sites_proportion(p, 0, 0.2));
    CHECK(!pll_update_invariant_sites_proportion(p, 0, 1.5));
    unsigned int * per_state = (unsigned int *)malloc(states * sizeof(unsigned int));
    pll_count_invariant_sites(p, per_state);
    free(per_state);
  }

  /* P-matrices (eigendecomposition runs for real on the host) */
  unsigned int mi[7] = {0, 1, 2, 3, 4, 5, 6};
  double bl[7] = {0.1, 0.2, 0.0, 0.4, 1e-9, 2.5, 0.05};
  CHECK(pll_update_prob_matrices(p, params, mi, bl, matrices));
  for (unsigned int i = 0; i < cats; ++i) CHECK(p->eigen_decomp_valid[params[i]]);
  for (unsigned int i = 0; i < states; ++i) CHECK(isfinite(p->eigenvals[0][i]));

  /* traversal ((0,1)5,(2,3)6)7 with tip 4 across the evaluation edge */
  pll_operation_t ops[3] = {{5, 0, 0, 0, PLL_SCALE_BUFFER_NONE, 1, 1, PLL_SCALE_BUFFER_NONE},
                            {6, 1, 2, 2, PLL_SCALE_BUFFER_NONE, 3, 3, PLL_SCALE_BUFFER_NONE},
                            {7, 2, 5, 4, 0, 6, 5, 1}};
  pll_update_partials(p, ops, 3);
  double * persite = (double *)malloc(sites * sizeof(double));
  double l1 = pll_compute_edge_loglikelihood(p, 7, 2, 4, PLL_SCALE_BUFFER_NONE, 6, params, persite);
  double l2 = pll_compute_edge_loglikelihood(p, 7, 2, 6, 1, 6, params, NULL);
  double l3 = pll_compute_root_loglikelihood(p, 7, 2, params, persite);
  /* with ascertainment bias the host epilogue takes logs of the stub's all-zero per-state CLVs */
  if (!ab) CHECK(isfinite(l1) && isfinite(l2) && isfinite(l3));

  /* sumtable + derivatives on an inner-inner and a tip-inner edge */
  const size_t table_len = (size_t)(sites + (ab ? states : 0)) * cats * p->states_padded;
  double * table = (double *)pll_aligned_alloc(table_len * sizeof(double), p->alignment);
  double d1, d2;
  CHECK(pll_update_sumtable(p, 7, 6, 2, 1, params, table));
  CHECK(pll_compute_likelihood_derivatives(p, 2, 1, 0.3, params, table, &d1, &d2));
  CHECK(pll_update_sumtable(p, 7, 4, 2, PLL_SCALE_BUFFER_NONE, params, table));
  CHECK(pll_compute_likelihood_derivatives(p, 2, PLL_SCALE_BUFFER_NONE, 0.3, params, table, &d1, &d2));
  pll_aligned_free(table);

  /* mirrors and printing */
  CHECK(pll_gpu_sync_clv(p, 7) && p->clv[7] != NULL);
  CHECK(pll_gpu_sync_scaler(p, 2) && p->scale_buffer[2] != NULL);
  CHECK(pll_gpu_sync_pmatrix(p, 3));
  CHECK(pll_gpu_push_pmatrix(p, 3));
  CHECK(pll_gpu_push_clv(p, 7));
  CHECK(!pll_gpu_sync_clv(p, tips + inner));
  CHECK(!pll_gpu_sync_scaler(p, inner));
  if (pattern_tip)
  {
    CHECK(pll_gpu_sync_tipchars(p, 0) && p->tipchars[0] != NULL);
    CHECK(!pll_gpu_sync_clv(p, 0));
  }
  else
    CHECK(pll_gpu_sync_clv(p, 0));
  pll_show_pmatrix(p, 0, 4);
  pll_show_clv(p, 7, 2, 4);
  pll_show_clv(p, 0, PLL_SCALE_BUFFER_NONE, 4);
  CHECK(pll_gpu_synchronize(p));
  CHECK(pll_gpu_context(p) != NULL);

  free(persite); free(weights); free(seq); free(params); free(freqs); free(subst); free(w); free(rates);
  pll_partition_destroy(p);
  CHECK(null_device_live_contexts() == 0);
  return 1;
}

int main(void)
{
  const unsigned int states[] = {2, 4, 5, 7, 20, 32};
  const unsigned int cats[] = {1, 4, 5};
  const unsigned int abs_[] = {0, PLL_ATTRIB_AB_LEWIS, PLL_ATTRIB_AB_FELSENSTEIN, PLL_ATTRIB_AB_STAMATAKIS,
                               PLL_ATTRIB_AB_FLAG};
  unsigned long done = 0, refused = 0;
  if (!freopen("/dev/null", "w", stdout)) return 2;
  for (unsigned int s = 0; s < sizeof(states) / sizeof(*states); ++s)
    for (unsigned int c = 0; c < sizeof(cats) / sizeof(*cats); ++c)
      for (int tip = 0; tip < 2; ++tip)
        for (int rs = 0; rs < 2; ++rs)
          for (unsigned int a = 0; a < sizeof(abs_) / sizeof(*abs_); ++a)
            for (int slices = 1; slices <= 3; slices += 2)
            {
              if (abs_[a] == PLL_ATTRIB_AB_FLAG) continue; /* storage only: set type later, covered below */
              if (scenario(states[s], cats[c], tip, rs, abs_[a], slices)) ++done; else ++refused;
            }
  /* device failures: the 2nd of 3 slice contexts cannot be created -> no partition, nothing left
   * behind; then every device call fails -> the wrappers report it through pll_errno and the
   * value-returning ones return -inf / PLL_FAILURE, with nothing pending on any slice */
  {
    CHECK(pll_gpu_set_devices(3));
    null_device_fail_create_in(2);
    pll_partition_t * p = pll_partition_create(5, 3, 4, 400, 1, 7, 4, 3, PLL_ATTRIB_ARCH_GPU);
    CHECK(p == NULL && pll_errno == PLL_ERROR_MEM_ALLOC && null_device_live_contexts() == 0);
    null_device_fail_create_in(0);
    p = pll_partition_create(5, 3, 4, 400, 1, 7, 4, 3, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP);
    pll_gpu_set_devices(0);
    CHECK(p != NULL && pll_gpu_partition_devices(p) == 3);
    unsigned int params[4] = {0, 0, 0, 0};
    pll_operation_t op = {5, 0, 0, 0, PLL_SCALE_BUFFER_NONE, 1, 1, PLL_SCALE_BUFFER_NONE};
    unsigned int w[400];
    for (int i = 0; i < 400; ++i) w[i] = 1;
    null_device_fail_calls(1);
    pll_errno = 0;
    pll_update_partials(p, &op, 1);
    CHECK(pll_errno == PLL_ERROR_GPU_RUNTIME);
    pll_errno = 0;
    pll_set_pattern_weights(p, w);
    CHECK(pll_errno == PLL_ERROR_GPU_RUNTIME);
    CHECK(!pll_update_invariant_sites(p));
    double l = pll_compute_edge_loglikelihood(p, 5, 0, 6, 1, 2, params, NULL);
    CHECK(isinf(l) && l < 0 && pll_errno == PLL_ERROR_GPU_RUNTIME);
    l = pll_compute_root_loglikelihood(p, 5, 0, params, NULL);
    CHECK(isinf(l) && l < 0);
    double table[4], d1, d2;
    CHECK(!pll_compute_likelihood_derivatives(p, 0, 1, 0.1, params, table, &d1, &d2));
    null_device_fail_calls(0);
    l = pll_compute_edge_loglikelihood(p, 5, 0, 6, 1, 2, params, NULL);
    CHECK(l == -400.0);   /* the null device returns -(patterns of the slice): 192 + 192 + 16 summed */
    pll_partition_destroy(p);
    CHECK(null_device_live_contexts() == 0);
  }

  /* the direct-call surface (pll_core_*, host arrays): every array is touched over the extent the
   * reference's callers allocate, for caller paddings equal to and different from the device's */
  {
    const unsigned int cstates[] = {4, 5, 20}, cattr[] = {PLL_ATTRIB_ARCH_AVX2, PLL_ATTRIB_ARCH_CPU, PLL_ATTRIB_ARCH_SSE};
    for (unsigned int s = 0; s < 3; ++s)
      for (unsigned int a = 0; a < 3; ++a)
        for (int rs = 0; rs < 2; ++rs)
        {
          const unsigned int K = cstates[s], R = 3, sites = 37, attrib = cattr[a] | (rs ? PLL_ATTRIB_RATE_SCALERS : 0);
          const unsigned int Kp = (cattr[a] == PLL_ATTRIB_ARCH_CPU) ? K : (cattr[a] == PLL_ATTRIB_ARCH_SSE ? ((K + 1) & ~1u) : ((K + 3) & ~3u));
          const size_t clv_len = (size_t)sites * R * Kp, mat_len = (size_t)R * K * Kp, sc_len = sites * (rs ? R : 1);
          double * clv[3], * mat[3], * ev[3], * iv[3], * fr[3], * el[3];
          double * table = (double *)calloc(clv_len, sizeof(double));
          double * lookup = (double *)calloc(1024 * R + 1024 * (size_t)R * Kp, sizeof(double));
          unsigned int * sc[3];
          for (int i = 0; i < 3; ++i)
          {
            clv[i] = (double *)calloc(clv_len, sizeof(double));
            mat[i] = (double *)calloc(mat_len, sizeof(double));
            ev[i] = (double *)calloc((size_t)K * Kp, sizeof(double));
            iv[i] = (double *)calloc((size_t)K * Kp, sizeof(double));
            fr[i] = (double *)calloc(Kp, sizeof(double));
            el[i] = (double *)calloc(Kp, sizeof(double));
            sc[i] = (unsigned int *)calloc(sc_len, sizeof(unsigned int));
          }
          unsigned char * chars = (unsigned char *)calloc(sites, 1);
          unsigned int tipmap[32], weights[37], idx[3] = {0, 1, 2}, mi[3] = {2, 0, 1};
          int inv[37];
          for (unsigned int i = 0; i < 32; ++i) tipmap[i] = 1u << (i % K);
          for (unsigned int i = 0; i < sites; ++i) { chars[i] = (unsigned char)(1 + i % 3); weights[i] = 1; inv[i] = -1; }
          double rates[3] = {0.5, 1.0, 1.5}, rw[3] = {0.3, 0.3, 0.4}, pinv[3] = {0.1, 0.1, 0.1}, bl[3] = {0.1, 0.2, 0.3}, d1, d2;
          pll_errno = 0;
          pll_core_update_partial_ii(K, sites, R, clv[0], sc[0], clv[1], clv[2], mat[0], mat[1], sc[1], sc[2], attrib);
          pll_core_update_partial_ii(K, sites, R, clv[0], NULL, clv[1], clv[2], mat[0], mat[1], NULL, NULL, attrib);
          pll_core_update_partial_ti(K, sites, R, clv[0], sc[0], chars, clv[2], mat[0], mat[1], sc[2], tipmap, 32, attrib);
          pll_core_create_lookup(K, R, lookup, mat[0], mat[1], tipmap, 32, attrib);
          pll_core_update_partial_tt(K, sites, R, clv[0], sc[0], chars, chars, tipmap, 32, lookup, attrib);
          CHECK(pll_errno == 0);
          CHECK(pll_core_update_pmatrix(mat, K, R, rates, bl, mi, idx, pinv, el, ev, iv, 3, attrib));
          CHECK(pll_core_update_sumtable_ii(K, sites, R, clv[1], clv[2], sc[1], sc[2], ev, iv, fr, table, attrib));
          CHECK(pll_core_update_sumtable_ti(K, sites, R, clv[1], chars, sc[1], ev, iv, fr, tipmap, 32, table, attrib));
          CHECK(pll_core_likelihood_derivatives(K, sites, R, rw, sc[1], sc[2], inv, weights, 0.3, pinv, fr, rates, el,
                                                table, &d1, &d2, attrib));
          double l = pll_core_root_loglikelihood(K, sites, R, clv[1], sc[1], fr, rw, weights, pinv, inv, idx, table, attrib);
          CHECK(l == -(double)sites);
          l = pll_core_edge_loglikelihood_ii(K, sites, R, clv[1], sc[1], clv[2], sc[2], mat[0], fr, rw, weights, NULL,
                                             NULL, idx, NULL, attrib);
          CHECK(l == -(double)sites);
          l = pll_core_edge_loglikelihood_ti(K, sites, R, clv[1], sc[1], chars, tipmap, 32, mat[0], fr, rw, weights,
                                             pinv, inv, idx, table, attrib);
          CHECK(l == -(double)sites);
          /* ascertainment-bias bits belong to the partition API */
          CHECK(!pll_core_update_sumtable_ii(K, sites, R, clv[1], clv[2], sc[1], sc[2], ev, iv, fr, table,
                                             attrib | PLL_ATTRIB_AB_LEWIS) && pll_errno == PLL_ERROR_GPU_UNSUPPORTED);
          for (int i = 0; i < 3; ++i) { free(clv[i]); free(mat[i]); free(ev[i]); free(iv[i]); free(fr[i]); free(el[i]); free(sc[i]); }
          free(table); free(lookup); free(chars);
        }
    CHECK(null_device_live_contexts() > 0);    /* the cached scratch partitions */
    pll_gpu_core_release();
    CHECK(null_device_live_contexts() == 0);
  }

  /* refusals */
  CHECK(pll_partition_create(4, 2, 4, 10, 1, 5, 4, 2, PLL_ATTRIB_ARCH_AVX2) == NULL);
  CHECK(pll_partition_create(4, 2, 4, 0, 1, 5, 4, 2, PLL_ATTRIB_ARCH_GPU) == NULL);
  CHECK(pll_partition_create(4, 2, 4, 10, 1, 5, 4, 2, PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_ARCH_AVX) == NULL);
  CHECK(!pll_gpu_set_devices(-1) && !pll_gpu_set_devices(1000));
  pll_partition_destroy(NULL);
  fprintf(stderr, "scenarios=%lu refused=%lu\n", done, refused);
  return 0;
}
