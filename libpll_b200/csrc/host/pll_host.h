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

typedef struct pllg_partition
{
  pll_partition_t pub; /* MUST be first */
  unsigned int magic;
  plg_context_t * ctx;
  unsigned int sites_alloc;  /* sites (+ states with ascertainment-bias storage) */
  unsigned char * tip_stage; /* sites_alloc bytes: encoding buffer for one tip   */
} pllg_partition_t;

static inline pllg_partition_t * pllg_from(const pll_partition_t * p)
{
  pllg_partition_t * g = (pllg_partition_t *)(void *)p;
  return (g && g->magic == PLLG_MAGIC) ? g : NULL;
}

/* device that pll_partition_create would use in this thread (-1 = current CUDA device) */
int pll_gpu_current_device(void);

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
