/*
 * null_device.c - a stand-in for the CUDA device layer (include/pll_gpu.h, plg_*) that computes
 * nothing but TOUCHES every byte the real one would: each input array is read and each output
 * array written over exactly the extent the device layer's contract states.  Linked with the host
 * C layer under AddressSanitizer (tests/test_sanitizers_cpu.py) it turns any disagreement
 * about buffer sizes between the pll.h wrappers and the device ABI - and any leak or overflow in
 * the wrappers themselves - into a report, on a machine without a GPU.
 * TEST INFRASTRUCTURE: never linked into libpll_b200.so.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pll.h"
#include "pll_gpu.h"

struct plg_context
{
  plg_dims_t d;
  unsigned int active_sites;
  unsigned int maxstates;
  int deferred;
  double * pending[2];
  unsigned long long calls;
};

static volatile unsigned long long sink;
static int live_contexts;

static void touch_in(const void * p, size_t bytes)
{
  const unsigned char * b = (const unsigned char *)p;
  unsigned long long s = 0;
  for (size_t i = 0; i < bytes; ++i) s += b[i];
  sink += s;
}
static void touch_out(void * p, size_t bytes, int value) { memset(p, value, bytes); }

static size_t span(const plg_context_t * c) { return (size_t)c->d.rate_cats * c->d.states_padded; }
static size_t scaler_len(const plg_context_t * c, size_t sites)
{
  return sites * ((c->d.attributes & PLL_ATTRIB_RATE_SCALERS) ? c->d.rate_cats : 1u);
}
static size_t pmatrix_len(const plg_context_t * c)
{
  return (size_t)c->d.rate_cats * c->d.states * c->d.states_padded;
}

int null_device_live_contexts(void) { return live_contexts; }

/* failure injection: the n-th plg_create from now on fails (0 = never); every other call fails
 * with PLG_E_CUDA while fail_calls is set */
static int fail_create_in, fail_calls;
void null_device_fail_create_in(int n) { fail_create_in = n; }
void null_device_fail_calls(int on) { fail_calls = on; }
#define MAYBE_FAIL() do { if (fail_calls) return PLG_E_CUDA; } while (0)

const char * plg_last_error(void) { return "null device"; }
int plg_device_count(void) { return 2; }

int plg_create(const plg_dims_t * dims, int device, plg_context_t ** out)
{
  if (!dims || !out || dims->sites == 0 || dims->states_padded < dims->states || device >= 2) return PLG_E_INVALID;
  if (fail_create_in && --fail_create_in == 0) return PLG_E_NOMEM;
  plg_context_t * c = (plg_context_t *)calloc(1, sizeof(*c));
  if (!c) return PLG_E_NOMEM;
  c->d = *dims;
  c->active_sites = dims->sites;
  ++live_contexts;
  *out = c;
  return PLG_OK;
}
void plg_destroy(plg_context_t * ctx)
{
  if (!ctx) return;
  --live_contexts;
  free(ctx);
}
int plg_synchronize(plg_context_t * ctx) { return ctx ? PLG_OK : PLG_E_INVALID; }
int plg_set_deferred(plg_context_t * ctx, int enable)
{
  ctx->deferred = enable;
  ctx->pending[0] = ctx->pending[1] = NULL;
  return PLG_OK;
}
int plg_collect(plg_context_t * ctx)
{
  if (ctx->pending[0]) *ctx->pending[0] = -(double)ctx->active_sites;
  if (ctx->pending[1]) *ctx->pending[1] = (double)ctx->active_sites;
  ctx->pending[0] = ctx->pending[1] = NULL;
  return PLG_OK;
}
static int deliver(plg_context_t * ctx, double * a, double * b)
{
  if (ctx->deferred)
  {
    ctx->pending[0] = a;
    ctx->pending[1] = b;
    return PLG_OK;
  }
  if (a) *a = -(double)ctx->active_sites;
  if (b) *b = (double)ctx->active_sites;
  return PLG_OK;
}

int plg_comm_unique_id(unsigned char * id) { return PLG_E_UNSUPPORTED; }
int plg_comm_init(const unsigned char * id, int nranks, int rank, int device) { return PLG_E_UNSUPPORTED; }
int plg_comm_finalize(void) { return PLG_OK; }
int plg_comm_size(void) { return 0; }
/* no peer memory in the null device: the host layer adds the per-slice results itself */
int plg_group_create(plg_context_t * const * members, unsigned int n) { return PLG_E_UNSUPPORTED; }
int plg_group_begin(plg_context_t * leader) { return PLG_E_INVALID; }
int plg_group_collect(plg_context_t * leader, double * out0, double * out1) { return PLG_E_INVALID; }
int plg_group_abort(plg_context_t * leader) { return PLG_OK; }
int plg_set_tipchars(plg_context_t * ctx, unsigned int tip_index, const unsigned char * chars)
{
  if (tip_index >= ctx->d.tips) return PLG_E_INVALID;
  touch_in(chars, ctx->d.sites);
  return PLG_OK;
}
int plg_generate_tipchars(plg_context_t * ctx, unsigned int tip_index, unsigned long long seed,
                          unsigned long long first_site)
{
  return tip_index < ctx->d.tips ? PLG_OK : PLG_E_INVALID;
}
int plg_get_tipchars(plg_context_t * ctx, unsigned int tip_index, unsigned char * chars)
{
  if (tip_index >= ctx->d.tips) return PLG_E_INVALID;
  touch_out(chars, ctx->d.sites, 1);
  return PLG_OK;
}
int plg_set_tipmap(plg_context_t * ctx, const unsigned int * tipmap, unsigned int maxstates)
{
  if (maxstates > PLL_ASCII_SIZE) return PLG_E_INVALID;
  touch_in(tipmap, maxstates * sizeof(unsigned int));
  ctx->maxstates = maxstates;
  return PLG_OK;
}
int plg_set_clv(plg_context_t * ctx, unsigned int clv_index, const double * clv)
{
  if (clv_index >= ctx->d.tips + ctx->d.clv_buffers) return PLG_E_INVALID;
  touch_in(clv, ctx->d.sites * span(ctx) * sizeof(double));
  return PLG_OK;
}
int plg_get_clv(plg_context_t * ctx, unsigned int clv_index, double * clv)
{
  if (clv_index >= ctx->d.tips + ctx->d.clv_buffers) return PLG_E_INVALID;
  touch_out(clv, ctx->d.sites * span(ctx) * sizeof(double), 0);
  return PLG_OK;
}
int plg_set_scaler(plg_context_t * ctx, unsigned int scaler_index, const unsigned int * scaler)
{
  if (scaler_index >= ctx->d.scale_buffers) return PLG_E_INVALID;
  touch_in(scaler, scaler_len(ctx, ctx->d.sites) * sizeof(unsigned int));
  return PLG_OK;
}
int plg_get_scaler(plg_context_t * ctx, unsigned int scaler_index, unsigned int * scaler)
{
  if (scaler_index >= ctx->d.scale_buffers) return PLG_E_INVALID;
  touch_out(scaler, scaler_len(ctx, ctx->d.sites) * sizeof(unsigned int), 0);
  return PLG_OK;
}
int plg_set_pattern_weights(plg_context_t * ctx, const unsigned int * weights)
{
  MAYBE_FAIL();
  touch_in(weights, ctx->d.sites * sizeof(unsigned int));
  return PLG_OK;
}
int plg_update_invariant(plg_context_t * ctx, int * invariant_out)
{
  MAYBE_FAIL();
  if (invariant_out) touch_out(invariant_out, ctx->d.sites * sizeof(int), 0xff); /* -1 everywhere */
  return PLG_OK;
}
int plg_set_invariant(plg_context_t * ctx, const int * invariant)
{
  MAYBE_FAIL();
  touch_in(invariant, ctx->d.sites * sizeof(int));
  return PLG_OK;
}
int plg_set_pmatrix(plg_context_t * ctx, unsigned int matrix_index, const double * pmatrix)
{
  if (matrix_index >= ctx->d.prob_matrices) return PLG_E_INVALID;
  touch_in(pmatrix, pmatrix_len(ctx) * sizeof(double));
  return PLG_OK;
}
int plg_get_pmatrix(plg_context_t * ctx, unsigned int matrix_index, double * pmatrix)
{
  if (matrix_index >= ctx->d.prob_matrices) return PLG_E_INVALID;
  touch_out(pmatrix, pmatrix_len(ctx) * sizeof(double), 0);
  return PLG_OK;
}
int plg_set_active_sites(plg_context_t * ctx, unsigned int sites)
{
  if (sites > ctx->d.sites) return PLG_E_INVALID;
  ctx->active_sites = sites;
  return PLG_OK;
}
int plg_get_clv_sites(plg_context_t * ctx, unsigned int clv_index, unsigned int first_site,
                      unsigned int count, double * out)
{
  if ((size_t)first_site + count > ctx->d.sites) return PLG_E_INVALID;
  touch_out(out, count * span(ctx) * sizeof(double), 0);
  return PLG_OK;
}
int plg_get_scaler_sites(plg_context_t * ctx, unsigned int scaler_index, unsigned int first_site,
                         unsigned int count, unsigned int * out)
{
  if ((size_t)first_site + count > ctx->d.sites) return PLG_E_INVALID;
  touch_out(out, scaler_len(ctx, count) * sizeof(unsigned int), 0);
  return PLG_OK;
}
int plg_get_sumtable_sites(plg_context_t * ctx, const void * key, unsigned int first_site,
                           unsigned int count, double * out)
{
  if ((size_t)first_site + count > ctx->d.sites) return PLG_E_INVALID;
  touch_out(out, count * span(ctx) * sizeof(double), 0);
  return PLG_OK;
}

int plg_update_pmatrix(plg_context_t * ctx, const unsigned int * matrix_indices,
                       const double * branch_lengths, unsigned int count, const double * rates,
                       const double * prop_invar, const double * eigenvals, const double * eigenvecs,
                       const double * inv_eigenvecs)
{
  const size_t R = ctx->d.rate_cats, K = ctx->d.states, Kp = ctx->d.states_padded;
  touch_in(matrix_indices, count * sizeof(unsigned int));
  touch_in(branch_lengths, count * sizeof(double));
  touch_in(rates, R * sizeof(double));
  touch_in(prop_invar, R * sizeof(double));
  touch_in(eigenvals, R * Kp * sizeof(double));
  touch_in(eigenvecs, R * K * Kp * sizeof(double));
  touch_in(inv_eigenvecs, R * K * Kp * sizeof(double));
  for (unsigned int i = 0; i < count; ++i)
    if (matrix_indices[i] >= ctx->d.prob_matrices) return PLG_E_INVALID;
  return PLG_OK;
}
int plg_update_partials(plg_context_t * ctx, const pll_operation_t * operations, unsigned int count)
{
  MAYBE_FAIL();
  touch_in(operations, count * sizeof(pll_operation_t));
  ctx->calls += count;
  return PLG_OK;
}
static void touch_model(plg_context_t * ctx, const double * freqs, const double * rate_weights,
                        const double * prop_invar)
{
  touch_in(freqs, span(ctx) * sizeof(double));
  touch_in(rate_weights, ctx->d.rate_cats * sizeof(double));
  touch_in(prop_invar, ctx->d.rate_cats * sizeof(double));
}
int plg_edge_loglikelihood(plg_context_t * ctx, unsigned int parent_clv_index, int parent_scaler_index,
                           unsigned int child_clv_index, int child_scaler_index,
                           unsigned int matrix_index, const double * freqs, const double * rate_weights,
                           const double * prop_invar, double * persite_lnl, double * logl_out)
{
  MAYBE_FAIL();
  touch_model(ctx, freqs, rate_weights, prop_invar);
  if (persite_lnl) touch_out(persite_lnl, ctx->active_sites * sizeof(double), 0);
  return deliver(ctx, logl_out, NULL);
}
int plg_root_loglikelihood(plg_context_t * ctx, unsigned int clv_index, int scaler_index,
                           const double * freqs, const double * rate_weights, const double * prop_invar,
                           double * persite_lnl, double * logl_out)
{
  MAYBE_FAIL();
  touch_model(ctx, freqs, rate_weights, prop_invar);
  if (persite_lnl) touch_out(persite_lnl, ctx->active_sites * sizeof(double), 0);
  return deliver(ctx, logl_out, NULL);
}
int plg_root_loglikelihood_counts(plg_context_t * ctx, unsigned int clv_index, const unsigned int * site_counts,
                                  const double * freqs, const double * rate_weights, const double * prop_invar,
                                  double * persite_lnl, double * logl_out)
{
  MAYBE_FAIL();
  touch_in(site_counts, ctx->d.sites * sizeof(unsigned int));
  touch_model(ctx, freqs, rate_weights, prop_invar);
  if (persite_lnl) touch_out(persite_lnl, ctx->active_sites * sizeof(double), 0);
  return deliver(ctx, logl_out, NULL);
}
int plg_update_sumtable(plg_context_t * ctx, unsigned int parent_clv_index, unsigned int child_clv_index,
                        int parent_scaler_index, int child_scaler_index, const double * eigenvecs,
                        const double * left_terms, const void * key, double * host_copy)
{
  const size_t R = ctx->d.rate_cats, K = ctx->d.states, Kp = ctx->d.states_padded;
  const int pattern_tips = (ctx->d.attributes & PLL_ATTRIB_PATTERN_TIP) != 0;
  const int ptip = pattern_tips && parent_clv_index < ctx->d.tips;
  const int ctip = pattern_tips && child_clv_index < ctx->d.tips;
  touch_in(eigenvecs, R * K * Kp * sizeof(double));
  if (ptip || ctip)
    touch_in(left_terms, (size_t)(K == 4 ? 16u : ctx->maxstates) * R * Kp * sizeof(double));
  else
    touch_in(left_terms, R * K * Kp * sizeof(double));
  if (host_copy) touch_out(host_copy, ctx->d.sites * span(ctx) * sizeof(double), 0);
  return PLG_OK;
}
int plg_set_sumtable(plg_context_t * ctx, const void * key, const double * table)
{
  MAYBE_FAIL();
  touch_in(table, (size_t)ctx->d.sites * span(ctx) * sizeof(double));
  return PLG_OK;
}
int plg_free_sumtable(plg_context_t * ctx, const void * key) { return PLG_OK; }
int plg_likelihood_derivatives(plg_context_t * ctx, const void * key, const double * diagptable,
                               const double * rate_weights, const double * prop_invar,
                               const double * freqs, double * d_f, double * dd_f)
{
  MAYBE_FAIL();
  touch_in(diagptable, (size_t)ctx->d.rate_cats * ctx->d.states * 4 * sizeof(double));
  touch_model(ctx, freqs, rate_weights, prop_invar);
  return deliver(ctx, d_f, dd_f);
}

int plg_timer_start(plg_context_t * ctx) { return PLG_OK; }
int plg_timer_stop(plg_context_t * ctx, float * elapsed_ms) { *elapsed_ms = 0; return PLG_OK; }
int plg_get_stats(plg_context_t * ctx, plg_stats_t * out) { memset(out, 0, sizeof(*out)); return PLG_OK; }
int plg_set_profiling(plg_context_t * ctx, int enable) { return PLG_OK; }
int plg_reset_stats(plg_context_t * ctx) { return PLG_OK; }
int plg_flush_l2(plg_context_t * ctx) { return PLG_OK; }
int plg_mem_info(plg_context_t * ctx, size_t * free_bytes, size_t * total_bytes)
{
  *free_bytes = *total_bytes = 0;
  return PLG_OK;
}
int plg_compress_patterns(int device, unsigned char * const * rows, unsigned int taxa, size_t length,
                          const unsigned char * code_table, const unsigned char * inverse_table,
                          unsigned int * weights_out, size_t * unique_out)
{
  return PLG_E_NODEVICE;
}
