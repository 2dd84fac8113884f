/*
 * pll_core.c - the reference's direct-call surface (`pll_core_*`, reference src/pll.h:864-1000,
 * :1659-1700) on the device.
 *
 * Callers that bypass pll_partition_t (model-selection tools keep their own CLV / P-matrix arrays
 * and call the kernels directly) hand over plain HOST arrays.  There is no resident state to act
 * on, so every call is upload -> the same kernels the partition API runs -> download, through a
 * scratch GPU partition (2 tips, 3 CLV buffers, 3 scale buffers, one "rate matrix" per rate
 * category so that the caller's per-rate arrays map 1:1) that is cached per thread and reused
 * while the dimensions stay the same.  Nothing is computed on the host: the functions exist for
 * link- and result-compatibility, the fast path is the partition API (data stays in HBM).
 *
 * Array layouts are the reference's for the architecture bits in `attrib`
 * (states_padded = states for PLL_ATTRIB_ARCH_CPU, even for _SSE, a multiple of 4 for _AVX /
 * _AVX2 / _GPU: reference src/pll.c:425-453, src/core_partials.c:534-587); they are re-padded
 * to the device layout on the way in and out.
 *
 * The tip-tip pair: the reference's pll_core_create_lookup (src/core_partials.c:82-186) fills a
 * caller-allocated table that pll_core_update_partial_tt (:188-260) later indexes.  The table's
 * layout is private to that pair; here it carries the two P-matrix sets (2 x rate_cats x states x
 * states_padded doubles - it always fits: the reference sizes the table for maxstates^2 entries),
 * and the tip-tip update builds the device tables from them as pll_update_partials does.
 *
 * Not supported (PLL_ERROR_GPU_UNSUPPORTED): PLL_ATTRIB_AB_* bits in `attrib` - the correction
 * needs the partition's per-state sites; use the partition API.
 */
#include "pll_host.h"

enum { SCR_SITES = 0, SCR_PMAT = 1 };
enum { TIP_A = 0, TIP_B = 1, CLV_P = 2, CLV_A = 3, CLV_B = 4 };

typedef struct
{
  pll_partition_t * p;
  unsigned int states, sites, rate_cats, prob_matrices, rate_scalers;
} scratch_t;

static __thread scratch_t g_scratch[2];

static unsigned int host_padded(unsigned int states, unsigned int attrib)
{
  if (attrib & (PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_ARCH_AVX | PLL_ATTRIB_ARCH_AVX2)) return (states + 3) & ~3u;
  if (attrib & PLL_ATTRIB_ARCH_SSE) return (states + 1) & ~1u;
  return states;
}

PLL_EXPORT void pll_gpu_core_release(void)
{
  for (int i = 0; i < 2; ++i)
  {
    if (g_scratch[i].p) pll_partition_destroy(g_scratch[i].p);
    memset(&g_scratch[i], 0, sizeof(scratch_t));
  }
}

static pll_partition_t * scratch_get(int which, unsigned int states, unsigned int sites, unsigned int rate_cats,
                                     unsigned int prob_matrices, unsigned int attrib)
{
  scratch_t * s = &g_scratch[which];
  const unsigned int rs = attrib & PLL_ATTRIB_RATE_SCALERS;
  if (attrib & PLL_ATTRIB_AB_MASK)
  {
    pll_fail(PLL_ERROR_GPU_UNSUPPORTED, "pll_core_*: ascertainment-bias bits need the partition API.");
    return NULL;
  }
  if (s->p && s->states == states && s->sites == sites && s->rate_cats == rate_cats &&
      s->prob_matrices >= prob_matrices && s->rate_scalers == rs)
    return s->p;
  if (s->p) pll_partition_destroy(s->p);
  memset(s, 0, sizeof(*s));
  unsigned int pm = 2;
  while (pm < prob_matrices) pm *= 2;
  const int slices = pllg_swap_slices(1);
  s->p = pll_partition_create(2, 3, states, sites, rate_cats, pm, rate_cats, 3,
                              PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP | rs);
  pllg_swap_slices(slices);
  if (!s->p) return NULL;
  s->states = states;
  s->sites = sites;
  s->rate_cats = rate_cats;
  s->prob_matrices = pm;
  s->rate_scalers = rs;
  for (unsigned int r = 0; r < rate_cats; ++r) s->p->eigen_decomp_valid[r] = 1;
  return s->p;
}

/* rows of `k` doubles between two paddings; pads of the destination are zeroed */
static void repad(double * dst, unsigned int dst_pad, const double * src, unsigned int src_pad, size_t rows,
                  unsigned int k)
{
  for (size_t i = 0; i < rows; ++i)
  {
    memcpy(dst + i * dst_pad, src + i * src_pad, k * sizeof(double));
    for (unsigned int j = k; j < dst_pad; ++j) dst[i * dst_pad + j] = 0.0;
  }
}

static __thread unsigned int identity[64];
static const unsigned int * ident(unsigned int n)
{
  if (n > 64) return NULL;
  for (unsigned int i = 0; i < n; ++i) identity[i] = i;
  return identity;
}

/* ---- uploads ---- */
static int put_clv(pll_partition_t * p, unsigned int index, const double * clv, unsigned int attrib)
{
  pllg_partition_t * g = pllg_from(p);
  const unsigned int hp = host_padded(p->states, attrib);
  const size_t rows = (size_t)p->sites * p->rate_cats;
  int rc;
  if (hp == p->states_padded)
    rc = pllg_dev_set_clv(g, index, clv);
  else
  {
    double * tmp = (double *)malloc(rows * p->states_padded * sizeof(double));
    if (!tmp) return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    repad(tmp, p->states_padded, clv, hp, rows, p->states);
    rc = pllg_dev_set_clv(g, index, tmp);
    if (!rc) rc = pllg_dev_synchronize(g);
    free(tmp);
  }
  return rc ? pllg_fail(rc, "pll_core: CLV upload") : PLL_SUCCESS;
}

static int get_clv(pll_partition_t * p, unsigned int index, double * clv, unsigned int attrib)
{
  pllg_partition_t * g = pllg_from(p);
  const unsigned int hp = host_padded(p->states, attrib);
  const size_t rows = (size_t)p->sites * p->rate_cats;
  int rc;
  if (hp == p->states_padded)
    rc = pllg_dev_get_clv(g, index, clv);
  else
  {
    double * tmp = (double *)malloc(rows * p->states_padded * sizeof(double));
    if (!tmp) return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    rc = pllg_dev_get_clv(g, index, tmp);
    if (!rc) repad(clv, hp, tmp, p->states_padded, rows, p->states);
    free(tmp);
  }
  return rc ? pllg_fail(rc, "pll_core: CLV download") : PLL_SUCCESS;
}

/* returns the scaler index to use in the call (PLL_SCALE_BUFFER_NONE for a NULL array) or -2 */
static int put_scaler(pll_partition_t * p, int index, const unsigned int * scaler)
{
  if (!scaler) return PLL_SCALE_BUFFER_NONE;
  int rc = plg_set_scaler(pllg_from(p)->ctx, (unsigned int)index, scaler);
  if (rc)
  {
    pllg_fail(rc, "pll_core: scaler upload");
    return -2;
  }
  return index;
}

static int put_pmatrix(pll_partition_t * p, unsigned int index, const double * pmatrix, unsigned int attrib)
{
  repad(p->pmatrix[index], p->states_padded, pmatrix, host_padded(p->states, attrib),
        (size_t)p->rate_cats * p->states, p->states);
  return pll_gpu_push_pmatrix(p, index);
}

static int put_tip(pll_partition_t * p, unsigned int tip, const unsigned char * chars, const unsigned int * tipmap,
                   unsigned int tipmap_size)
{
  pllg_partition_t * g = pllg_from(p);
  int rc = PLG_OK;
  if (p->states != 4)
  {
    /* other alphabets: characters are codes into the caller's tipmap */
    if (!tipmap || tipmap_size == 0 || tipmap_size > PLL_ASCII_SIZE)
      return pll_fail(PLL_ERROR_PARAM_INVALID, "pll_core: invalid tipmap");
    if (!p->tipmap && !(p->tipmap = (unsigned int *)calloc(PLL_ASCII_SIZE, sizeof(unsigned int))))
      return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    memset(p->tipmap, 0, PLL_ASCII_SIZE * sizeof(unsigned int));
    memcpy(p->tipmap, tipmap, tipmap_size * sizeof(unsigned int));
    p->maxstates = tipmap_size;
    rc = pllg_dev_set_tipmap(g, p->tipmap, tipmap_size);
  }
  if (!rc) rc = pllg_dev_set_tipchars(g, tip, chars);
  if (!rc) rc = pllg_dev_synchronize(g);
  return rc ? pllg_fail(rc, "pll_core: tip upload") : PLL_SUCCESS;
}

static int put_sites(pll_partition_t * p, const unsigned int * pattern_weights, const int * invariant)
{
  pllg_partition_t * g = pllg_from(p);
  if (pattern_weights)
    pll_set_pattern_weights(p, pattern_weights);
  else
  {
    for (unsigned int i = 0; i < p->sites; ++i) p->pattern_weights[i] = 1;
    pll_set_pattern_weights(p, p->pattern_weights);
  }
  if (invariant)
  {
    if (!p->invariant && !(p->invariant = (int *)malloc((size_t)p->sites * sizeof(int))))
      return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    memcpy(p->invariant, invariant, (size_t)p->sites * sizeof(int));
    int rc = plg_set_invariant(g->ctx, invariant);
    if (rc) return pllg_fail(rc, "pll_core: invariant upload");
  }
  return PLL_SUCCESS;
}

/* the caller's per-rate model arrays -> "rate matrix" r of the scratch partition (NULL: keep) */
static void put_model(pll_partition_t * p, unsigned int r, const double * eigenvals, const double * eigenvecs,
                      const double * inv_eigenvecs, const double * freqs, unsigned int attrib)
{
  const unsigned int K = p->states, Kp = p->states_padded, hp = host_padded(K, attrib);
  if (eigenvals) repad(p->eigenvals[r], Kp, eigenvals, hp, 1, K);
  if (eigenvecs) repad(p->eigenvecs[r], Kp, eigenvecs, hp, K, K);
  if (inv_eigenvecs) repad(p->inv_eigenvecs[r], Kp, inv_eigenvecs, hp, K, K);
  if (freqs) repad(p->frequencies[r], Kp, freqs, hp, 1, K);
}

static int run_operation(pll_partition_t * p, unsigned int child1, int scaler1, unsigned int child2, int scaler2,
                         double * parent_clv, unsigned int * parent_scaler, unsigned int attrib)
{
  pll_operation_t op;
  op.parent_clv_index = CLV_P;
  op.parent_scaler_index = parent_scaler ? 0 : PLL_SCALE_BUFFER_NONE;
  op.child1_clv_index = child1;
  op.child1_matrix_index = 0;
  op.child1_scaler_index = scaler1;
  op.child2_clv_index = child2;
  op.child2_matrix_index = 1;
  op.child2_scaler_index = scaler2;
  pll_errno = 0;
  pll_update_partials(p, &op, 1);
  if (pll_errno) return PLL_FAILURE;
  if (!get_clv(p, CLV_P, parent_clv, attrib)) return PLL_FAILURE;
  if (parent_scaler)
  {
    int rc = pllg_dev_get_scaler(pllg_from(p), 0, parent_scaler);
    if (rc) return pllg_fail(rc, "pll_core: scaler download");
  }
  return PLL_SUCCESS;
}

/* ------------------------------------------------------------------------------------ */
/* CLV updates: reference src/core_partials.c:188-260 (tt), :262-532 (ti), :534-862 (ii) */
/* ------------------------------------------------------------------------------------ */
PLL_EXPORT void pll_core_update_partial_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                           double * parent_clv, unsigned int * parent_scaler,
                                           const double * left_clv, const double * right_clv,
                                           const double * left_matrix, const double * right_matrix,
                                           const unsigned int * left_scaler, const unsigned int * right_scaler,
                                           unsigned int attrib)
{
  pll_partition_t * p = scratch_get(SCR_SITES, states, sites, rate_cats, 2, attrib);
  if (!p) return;
  const int ls = put_scaler(p, 1, left_scaler), rs = put_scaler(p, 2, right_scaler);
  if (ls == -2 || rs == -2) return;
  if (!put_clv(p, CLV_A, left_clv, attrib) || !put_clv(p, CLV_B, right_clv, attrib) ||
      !put_pmatrix(p, 0, left_matrix, attrib) || !put_pmatrix(p, 1, right_matrix, attrib))
    return;
  run_operation(p, CLV_A, ls, CLV_B, rs, parent_clv, parent_scaler, attrib);
}

PLL_EXPORT void pll_core_update_partial_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                           double * parent_clv, unsigned int * parent_scaler,
                                           const unsigned char * left_tipchars, const double * right_clv,
                                           const double * left_matrix, const double * right_matrix,
                                           const unsigned int * right_scaler, const unsigned int * tipmap,
                                           unsigned int tipmap_size, unsigned int attrib)
{
  pll_partition_t * p = scratch_get(SCR_SITES, states, sites, rate_cats, 2, attrib);
  if (!p) return;
  const int rs = put_scaler(p, 2, right_scaler);
  if (rs == -2) return;
  if (!put_tip(p, TIP_A, left_tipchars, tipmap, tipmap_size) || !put_clv(p, CLV_B, right_clv, attrib) ||
      !put_pmatrix(p, 0, left_matrix, attrib) || !put_pmatrix(p, 1, right_matrix, attrib))
    return;
  run_operation(p, TIP_A, PLL_SCALE_BUFFER_NONE, CLV_B, rs, parent_clv, parent_scaler, attrib);
}

/* reference src/core_partials.c:82-186: here the table only carries the two matrix sets */
PLL_EXPORT void pll_core_create_lookup(unsigned int states, unsigned int rate_cats, double * lookup,
                                       const double * left_matrix, const double * right_matrix,
                                       const unsigned int * tipmap, unsigned int tipmap_size, unsigned int attrib)
{
  const size_t n = (size_t)rate_cats * states * host_padded(states, attrib);
  memcpy(lookup, left_matrix, n * sizeof(double));
  memcpy(lookup + n, right_matrix, n * sizeof(double));
}

PLL_EXPORT void pll_core_update_partial_tt(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                           double * parent_clv, unsigned int * parent_scaler,
                                           const unsigned char * left_tipchars,
                                           const unsigned char * right_tipchars, const unsigned int * tipmap,
                                           unsigned int tipmap_size, const double * lookup, unsigned int attrib)
{
  pll_partition_t * p = scratch_get(SCR_SITES, states, sites, rate_cats, 2, attrib);
  if (!p) return;
  const size_t n = (size_t)rate_cats * states * host_padded(states, attrib);
  if (!put_tip(p, TIP_A, left_tipchars, tipmap, tipmap_size) ||
      !put_tip(p, TIP_B, right_tipchars, tipmap, tipmap_size) || !put_pmatrix(p, 0, lookup, attrib) ||
      !put_pmatrix(p, 1, lookup + n, attrib))
    return;
  run_operation(p, TIP_A, PLL_SCALE_BUFFER_NONE, TIP_B, PLL_SCALE_BUFFER_NONE, parent_clv, parent_scaler, attrib);
}

/* ------------------------------------------------------------------------------------ */
/* P-matrices: reference src/core_pmatrix.c:24-250                                       */
/* ------------------------------------------------------------------------------------ */
PLL_EXPORT int pll_core_update_pmatrix(double ** pmatrix, unsigned int states, unsigned int rate_cats,
                                       const double * rates, const double * branch_lengths,
                                       const unsigned int * matrix_indices, const unsigned int * params_indices,
                                       const double * prop_invar, double * const * eigenvals,
                                       double * const * eigenvecs, double * const * inv_eigenvecs,
                                       unsigned int count, unsigned int attrib)
{
  if (rate_cats > 64) return pll_fail(PLL_ERROR_GPU_UNSUPPORTED, "pll_core_update_pmatrix: more than 64 rate categories");
  if (count == 0) return PLL_SUCCESS;
  pll_partition_t * p = scratch_get(SCR_PMAT, states, 1, rate_cats, count, attrib & ~PLL_ATTRIB_RATE_SCALERS);
  if (!p) return PLL_FAILURE;
  const unsigned int hp = host_padded(states, attrib);
  for (unsigned int r = 0; r < rate_cats; ++r)
  {
    const unsigned int m = params_indices[r];
    put_model(p, r, eigenvals[m], eigenvecs[m], inv_eigenvecs[m], NULL, attrib);
    p->prop_invar[r] = prop_invar[m];
    p->rates[r] = rates[r];
  }
  unsigned int * idx = (unsigned int *)malloc(count * sizeof(unsigned int));
  if (!idx) return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
  for (unsigned int i = 0; i < count; ++i) idx[i] = i;
  int ok = pll_update_prob_matrices(p, ident(rate_cats), idx, branch_lengths, count);
  free(idx);
  for (unsigned int i = 0; ok && i < count; ++i)
  {
    ok = pll_gpu_sync_pmatrix(p, i);
    if (ok) repad(pmatrix[matrix_indices[i]], hp, p->pmatrix[i], p->states_padded, (size_t)rate_cats * states, states);
  }
  for (unsigned int r = 0; r < rate_cats; ++r) p->prop_invar[r] = 0.0;
  return ok;
}

/* ------------------------------------------------------------------------------------ */
/* sumtables and derivatives: reference src/core_derivatives.c:125-446, :501-732         */
/* ------------------------------------------------------------------------------------ */
static int fetch_sumtable(pll_partition_t * p, double * sumtable, unsigned int attrib)
{
  pllg_partition_t * g = pllg_from(p);
  const unsigned int hp = host_padded(p->states, attrib);
  const size_t rows = (size_t)p->sites * p->rate_cats;
  int rc;
  if (hp == p->states_padded)
    rc = pllg_dev_get_sumtable_sites(g, sumtable, 0, p->sites, sumtable);
  else
  {
    double * tmp = (double *)malloc(rows * p->states_padded * sizeof(double));
    if (!tmp) return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    rc = pllg_dev_get_sumtable_sites(g, sumtable, 0, p->sites, tmp);
    if (!rc) repad(sumtable, hp, tmp, p->states_padded, rows, p->states);
    free(tmp);
  }
  pll_gpu_free_sumtable(p, sumtable);
  return rc ? pllg_fail(rc, "pll_core: sumtable download") : PLL_SUCCESS;
}

PLL_EXPORT int pll_core_update_sumtable_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                           const double * parent_clv, const double * child_clv,
                                           const unsigned int * parent_scaler, const unsigned int * child_scaler,
                                           double * const * eigenvecs, double * const * inv_eigenvecs,
                                           double * const * freqs, double * sumtable, unsigned int attrib)
{
  if (rate_cats > 64) return pll_fail(PLL_ERROR_GPU_UNSUPPORTED, "pll_core_update_sumtable_ii: more than 64 rate categories");
  pll_partition_t * p = scratch_get(SCR_SITES, states, sites, rate_cats, 2, attrib);
  if (!p) return PLL_FAILURE;
  const int ps = put_scaler(p, 1, parent_scaler), cs = put_scaler(p, 2, child_scaler);
  if (ps == -2 || cs == -2) return PLL_FAILURE;
  if (!put_clv(p, CLV_A, parent_clv, attrib) || !put_clv(p, CLV_B, child_clv, attrib)) return PLL_FAILURE;
  for (unsigned int r = 0; r < rate_cats; ++r) put_model(p, r, NULL, eigenvecs[r], inv_eigenvecs[r], freqs[r], attrib);
  if (!pll_update_sumtable(p, CLV_A, CLV_B, ps, cs, ident(rate_cats), sumtable)) return PLL_FAILURE;
  return fetch_sumtable(p, sumtable, attrib);
}

PLL_EXPORT int pll_core_update_sumtable_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                           const double * parent_clv, const unsigned char * left_tipchars,
                                           const unsigned int * parent_scaler, double * const * eigenvecs,
                                           double * const * inv_eigenvecs, double * const * freqs,
                                           const unsigned int * tipmap, unsigned int tipmap_size,
                                           double * sumtable, unsigned int attrib)
{
  if (rate_cats > 64) return pll_fail(PLL_ERROR_GPU_UNSUPPORTED, "pll_core_update_sumtable_ti: more than 64 rate categories");
  pll_partition_t * p = scratch_get(SCR_SITES, states, sites, rate_cats, 2, attrib);
  if (!p) return PLL_FAILURE;
  const int ps = put_scaler(p, 1, parent_scaler);
  if (ps == -2) return PLL_FAILURE;
  if (!put_clv(p, CLV_A, parent_clv, attrib) || !put_tip(p, TIP_A, left_tipchars, tipmap, tipmap_size))
    return PLL_FAILURE;
  for (unsigned int r = 0; r < rate_cats; ++r) put_model(p, r, NULL, eigenvecs[r], inv_eigenvecs[r], freqs[r], attrib);
  if (!pll_update_sumtable(p, CLV_A, TIP_A, ps, PLL_SCALE_BUFFER_NONE, ident(rate_cats), sumtable))
    return PLL_FAILURE;
  return fetch_sumtable(p, sumtable, attrib);
}

PLL_EXPORT int pll_core_likelihood_derivatives(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                               const double * rate_weights, const unsigned int * parent_scaler,
                                               const unsigned int * child_scaler, const int * invariant,
                                               const unsigned int * pattern_weights, double branch_length,
                                               const double * prop_invar, double * const * freqs,
                                               const double * rates, double * const * eigenvals,
                                               const double * sumtable, double * d_f, double * dd_f,
                                               unsigned int attrib)
{
  if (rate_cats > 64) return pll_fail(PLL_ERROR_GPU_UNSUPPORTED, "pll_core_likelihood_derivatives: more than 64 rate categories");
  pll_partition_t * p = scratch_get(SCR_SITES, states, sites, rate_cats, 2, attrib);
  if (!p) return PLL_FAILURE;
  pllg_partition_t * g = pllg_from(p);
  const unsigned int hp = host_padded(states, attrib);
  int any_pinv = 0;
  for (unsigned int r = 0; r < rate_cats; ++r)
  {
    put_model(p, r, eigenvals[r], NULL, NULL, freqs[r], attrib);
    p->rates[r] = rates[r];
    p->rate_weights[r] = rate_weights[r];
    p->prop_invar[r] = prop_invar ? prop_invar[r] : 0.0;
    any_pinv |= p->prop_invar[r] > 0.0;
  }
  if (!put_sites(p, pattern_weights, any_pinv ? invariant : NULL)) return PLL_FAILURE;
  int rc;
  if (hp == p->states_padded)
    rc = plg_set_sumtable(g->ctx, sumtable, sumtable);
  else
  {
    const size_t rows = (size_t)sites * rate_cats;
    double * tmp = (double *)malloc(rows * p->states_padded * sizeof(double));
    if (!tmp) return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    repad(tmp, p->states_padded, sumtable, hp, rows, states);
    rc = plg_set_sumtable(g->ctx, sumtable, tmp);
    free(tmp);
  }
  if (rc) return pllg_fail(rc, "pll_core_likelihood_derivatives");
  /* the site scalers cancel in L'/L (they only enter the ascertainment-bias terms) */
  int ok = pll_compute_likelihood_derivatives(p, PLL_SCALE_BUFFER_NONE, PLL_SCALE_BUFFER_NONE, branch_length,
                                              ident(rate_cats), sumtable, d_f, dd_f);
  pll_gpu_free_sumtable(p, sumtable);
  for (unsigned int r = 0; r < rate_cats; ++r) p->prop_invar[r] = 0.0;
  return ok;
}

/* ------------------------------------------------------------------------------------ */
/* log-likelihoods: reference src/core_likelihood.c:25-209 (root), :211-1002 (edge)      */
/* ------------------------------------------------------------------------------------ */
static int put_lnl_model(pll_partition_t * p, double * const * frequencies, const double * rate_weights,
                         const unsigned int * pattern_weights, const double * invar_proportion,
                         const int * invar_indices, const unsigned int * freqs_indices, unsigned int attrib)
{
  int any_pinv = 0;
  for (unsigned int r = 0; r < p->rate_cats; ++r)
  {
    const unsigned int f = freqs_indices[r];
    put_model(p, r, NULL, NULL, NULL, frequencies[f], attrib);
    p->rate_weights[r] = rate_weights[r];
    p->prop_invar[r] = invar_proportion ? invar_proportion[f] : 0.0;
    any_pinv |= p->prop_invar[r] > 0.0;
  }
  return put_sites(p, pattern_weights, any_pinv ? invar_indices : NULL);
}

static void clear_pinv(pll_partition_t * p)
{
  for (unsigned int r = 0; r < p->rate_cats; ++r) p->prop_invar[r] = 0.0;
}

PLL_EXPORT double pll_core_root_loglikelihood(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                              const double * clv, const unsigned int * scaler,
                                              double * const * frequencies, const double * rate_weights,
                                              const unsigned int * pattern_weights,
                                              const double * invar_proportion, const int * invar_indices,
                                              const unsigned int * freqs_indices, double * persite_lnl,
                                              unsigned int attrib)
{
  pll_partition_t * p = rate_cats <= 64 ? scratch_get(SCR_SITES, states, sites, rate_cats, 2, attrib) : NULL;
  if (!p) return -INFINITY;
  const int sc = put_scaler(p, 1, scaler);
  if (sc == -2 || !put_clv(p, CLV_A, clv, attrib) ||
      !put_lnl_model(p, frequencies, rate_weights, pattern_weights, invar_proportion, invar_indices, freqs_indices,
                     attrib))
    return -INFINITY;
  const double logl = pll_compute_root_loglikelihood(p, CLV_A, sc, ident(rate_cats), persite_lnl);
  clear_pinv(p);
  return logl;
}

PLL_EXPORT double pll_core_edge_loglikelihood_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                                 const double * parent_clv, const unsigned int * parent_scaler,
                                                 const double * child_clv, const unsigned int * child_scaler,
                                                 const double * pmatrix, double * const * frequencies,
                                                 const double * rate_weights, const unsigned int * pattern_weights,
                                                 const double * invar_proportion, const int * invar_indices,
                                                 const unsigned int * freqs_indices, double * persite_lnl,
                                                 unsigned int attrib)
{
  pll_partition_t * p = rate_cats <= 64 ? scratch_get(SCR_SITES, states, sites, rate_cats, 2, attrib) : NULL;
  if (!p) return -INFINITY;
  const int ps = put_scaler(p, 1, parent_scaler), cs = put_scaler(p, 2, child_scaler);
  if (ps == -2 || cs == -2 || !put_clv(p, CLV_A, parent_clv, attrib) || !put_clv(p, CLV_B, child_clv, attrib) ||
      !put_pmatrix(p, 0, pmatrix, attrib) ||
      !put_lnl_model(p, frequencies, rate_weights, pattern_weights, invar_proportion, invar_indices, freqs_indices,
                     attrib))
    return -INFINITY;
  const double logl = pll_compute_edge_loglikelihood(p, CLV_A, ps, CLV_B, cs, 0, ident(rate_cats), persite_lnl);
  clear_pinv(p);
  return logl;
}

PLL_EXPORT double pll_core_edge_loglikelihood_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                                 const double * parent_clv, const unsigned int * parent_scaler,
                                                 const unsigned char * tipchars, const unsigned int * tipmap,
                                                 unsigned int tipmap_size, const double * pmatrix,
                                                 double * const * frequencies, const double * rate_weights,
                                                 const unsigned int * pattern_weights,
                                                 const double * invar_proportion, const int * invar_indices,
                                                 const unsigned int * freqs_indices, double * persite_lnl,
                                                 unsigned int attrib)
{
  pll_partition_t * p = rate_cats <= 64 ? scratch_get(SCR_SITES, states, sites, rate_cats, 2, attrib) : NULL;
  if (!p) return -INFINITY;
  const int ps = put_scaler(p, 1, parent_scaler);
  if (ps == -2 || !put_clv(p, CLV_A, parent_clv, attrib) || !put_tip(p, TIP_A, tipchars, tipmap, tipmap_size) ||
      !put_pmatrix(p, 0, pmatrix, attrib) ||
      !put_lnl_model(p, frequencies, rate_weights, pattern_weights, invar_proportion, invar_indices, freqs_indices,
                     attrib))
    return -INFINITY;
  const double logl =
      pll_compute_edge_loglikelihood(p, CLV_A, ps, TIP_A, PLL_SCALE_BUFFER_NONE, 0, ident(rate_cats), persite_lnl);
  clear_pinv(p);
  return logl;
}
