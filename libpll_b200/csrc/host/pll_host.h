/*
 * pll_host.h - internals of the host C layer (the pll.h API on top of the plg_* device ABI).
 *
 * The public pll_partition_t keeps the reference's exact layout (216 bytes), so the
 * backend's private state is appended behind it: pll_partition_create returns the address of
 * the `pub` member of a pllg_partition_t and every entry point recovers the wrapper from it.
 */
#ifndef PLL_B200_HOST_H_
#define PLL_B200_HOST_H_

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "pll.h"
#include "pll_gpu.h"

#define PLLG_MAGIC 0x706c6c67u /* "pllg" */

#define PLLG_MAX_DEVICES 16

typedef struct pllg_partition
{
  pll_partition_t pub; /* MUST be first */
  unsigned int magic;
  plg_context_t * ctx;       /* = ctxs[0]; the only context unless pll_gpu_set_devices(n > 1) */
  unsigned int sites_alloc;  /* sites (+ states with ascertainment-bias storage) */
  unsigned char * tip_stage; /* sites_alloc bytes: encoding buffer for one tip   */
  /* pattern slices over several devices (pll_devices.c): context d owns [lo[d], lo[d+1]) */
  unsigned int ndev;
  int grouped;               /* 1: scalar results are combined on the devices (plg_group_*) */
  plg_context_t * ctxs[PLLG_MAX_DEVICES];
  unsigned int lo[PLLG_MAX_DEVICES + 1];
  struct pllg_pool * pool;   /* helper threads, one per slice but the first (pll_devices.c); NULL: none */
} pllg_partition_t;

static inline pllg_partition_t * pllg_from(const pll_partition_t * p)
{
  pllg_partition_t * g = (pllg_partition_t *)(void *)p;
  return (g && g->magic == PLLG_MAGIC) ? g : NULL;
}

/* device that pll_partition_create would use in this thread (-1 = current CUDA device) */
int pll_gpu_current_device(void);

/* PLL_GPU_MIRROR set and not "0": keep the host mirrors of CLVs / scalers / tips current */
int pll_gpu_mirror_mode(void);

/* number of pattern slices pll_partition_create would use in this thread (pll_devices.c) */
int pll_gpu_current_slices(void);

int pllg_swap_slices(int count);

/* the plg_* device ABI fanned out over the partition's pattern slices (pll_devices.c): same
 * arguments as the plg_* function of the same name, host arrays indexed by site are offset per
 * slice, scalar results are summed in slice order */
int pllg_dev_create(pllg_partition_t * g, const plg_dims_t * dims, int first_device, int slices);
/* message of a device call that failed on a helper thread (plg_last_error is thread-local), or
 * NULL; reading it clears it */
const char * pllg_pending_error(void);
void pllg_dev_destroy(pllg_partition_t * g);
int pllg_dev_synchronize(pllg_partition_t * g);
int pllg_dev_set_tipmap(pllg_partition_t * g, const unsigned int * tipmap, unsigned int maxstates);
int pllg_dev_set_tipchars(pllg_partition_t * g, unsigned int tip_index, const unsigned char * chars);
int pllg_dev_generate_tipchars(pllg_partition_t * g, unsigned int tip_index, unsigned long long seed,
                               unsigned long long first_site);
int pllg_dev_get_tipchars(pllg_partition_t * g, unsigned int tip_index, unsigned char * chars);
int pllg_dev_set_clv(pllg_partition_t * g, unsigned int clv_index, const double * clv);
int pllg_dev_get_clv(pllg_partition_t * g, unsigned int clv_index, double * clv);
int pllg_dev_get_scaler(pllg_partition_t * g, unsigned int scaler_index, unsigned int * scaler);
int pllg_dev_set_pattern_weights(pllg_partition_t * g, const unsigned int * weights);
int pllg_dev_set_active_sites(pllg_partition_t * g, unsigned int count);
int pllg_dev_get_clv_sites(pllg_partition_t * g, unsigned int clv_index, unsigned int first, unsigned int count,
                           double * out);
int pllg_dev_get_scaler_sites(pllg_partition_t * g, unsigned int scaler_index, unsigned int first,
                              unsigned int count, unsigned int * out);
int pllg_dev_get_sumtable_sites(pllg_partition_t * g, const void * key, unsigned int first, unsigned int count,
                                double * out);
int pllg_dev_update_invariant(pllg_partition_t * g, int * invariant_out);
int pllg_dev_set_pmatrix(pllg_partition_t * g, unsigned int matrix_index, const double * pmatrix);
int pllg_dev_update_pmatrix(pllg_partition_t * g, const unsigned int * matrix_indices,
                            const double * branch_lengths, unsigned int count, const double * rates,
                            const double * prop_invar, const double * eigenvals,
                            const double * eigenvecs, const double * inv_eigenvecs);
int pllg_dev_update_partials(pllg_partition_t * g, const pll_operation_t * operations, unsigned int count);
int pllg_dev_edge_loglikelihood(pllg_partition_t * g, unsigned int parent_clv_index,
                                int parent_scaler_index, unsigned int child_clv_index,
                                int child_scaler_index, unsigned int matrix_index, const double * freqs,
                                const double * rate_weights, const double * prop_invar,
                                double * persite_lnl, double * logl_out);
int pllg_dev_root_loglikelihood(pllg_partition_t * g, unsigned int clv_index, int scaler_index,
                                const double * freqs, const double * rate_weights,
                                const double * prop_invar, double * persite_lnl, double * logl_out);
int pllg_dev_update_sumtable(pllg_partition_t * g, unsigned int parent_clv_index,
                             unsigned int child_clv_index, int parent_scaler_index,
                             int child_scaler_index, const double * eigenvecs,
                             const double * left_terms, const void * key, double * host_copy);
int pllg_dev_likelihood_derivatives(pllg_partition_t * g, const void * key, const double * diagptable,
                                    const double * rate_weights, const double * prop_invar,
                                    const double * freqs, double * d_f, double * dd_f);

/* ascertainment-bias epilogues (pll_ascbias.c) */
double pllg_asc_root(pllg_partition_t * g, unsigned int clv_index, int scaler_index,
                     const unsigned int * freqs_indices);
double pllg_asc_edge(pllg_partition_t * g, unsigned int parent_clv_index, int parent_scaler_index,
                     unsigned int child_clv_index, int child_scaler_index, unsigned int matrix_index,
                     const unsigned int * freqs_indices);
int pllg_asc_derivatives(pllg_partition_t * g, int parent_scaler_index, int child_scaler_index,
                         const double * diagptable, const double * sumtable_key, double * d_f,
                         double * dd_f);

/* sets pll_errno / pll_errmsg from a PLG_E_* code + plg_last_error(); returns PLL_FAILURE */
int pllg_fail(int plg_rc, const char * where);
/* sets pll_errno / pll_errmsg from a format; returns PLL_FAILURE */
int pll_fail(int code, const char * fmt, ...);

#endif
