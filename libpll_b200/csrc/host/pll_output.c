/*
 * pll_output.c - text dumps of P-matrices and CLVs in the reference's format
 * (reference src/output.c:26-96), used by the reference's golden-output regression programs
 * (test/src/00010_NMDU_lkcalc.c:170-190 and friends).  The arrays they print live in HBM, so
 * both functions first download what they need into the partition's host mirrors
 * (pll_gpu_sync_pmatrix / _clv / _scaler): unmodified callers print what the device computed.
 * The mirrors are caches behind a logically const partition, hence the cast.
 */
#include "pll_host.h"

PLL_EXPORT void pll_show_pmatrix(const pll_partition_t * partition, unsigned int index,
                                 unsigned int float_precision)
{
  if (!pllg_from(partition) || index >= partition->prob_matrices) return;
  if (!pll_gpu_sync_pmatrix((pll_partition_t *)(void *)partition, index))
  {
    printf("[ (P-matrix %u: %s) ]\n", index, pll_errmsg);
    return;
  }
  const unsigned int K = partition->states, Kp = partition->states_padded;
  unsigned int r, i, j;
  for (r = 0; r < partition->rate_cats; ++r)
  {
    const double * m = partition->pmatrix[index] + (size_t)r * K * Kp;
    for (i = 0; i < K; ++i)
    {
      for (j = 0; j < K; ++j) printf("%+2.*f   ", float_precision, m[i * Kp + j]);
      printf("\n");
    }
    printf("\n");
  }
}

PLL_EXPORT void pll_show_clv(const pll_partition_t * partition, unsigned int clv_index,
                             int scaler_index, unsigned int float_precision)
{
  if (!pllg_from(partition)) return;
  const unsigned int K = partition->states, Kp = partition->states_padded;
  const unsigned int R = partition->rate_cats;
  unsigned int n, r, s, t;

  if (clv_index < partition->tips && (partition->attributes & PLL_ATTRIB_PATTERN_TIP)) return;
  pll_partition_t * mirror = (pll_partition_t *)(void *)partition;
  if (!pll_gpu_sync_clv(mirror, clv_index) ||
      (scaler_index != PLL_SCALE_BUFFER_NONE && !pll_gpu_sync_scaler(mirror, (unsigned int)scaler_index)))
  {
    printf("[ (CLV %u: %s) ]\n", clv_index, pll_errmsg);
    return;
  }
  const double * clv = partition->clv[clv_index];
  const unsigned int * scaler =
      (scaler_index == PLL_SCALE_BUFFER_NONE) ? NULL : partition->scale_buffer[scaler_index];

  printf("[ ");
  for (n = 0; n < partition->sites; ++n)
  {
    printf("{");
    for (r = 0; r < R; ++r)
    {
      printf("(");
      for (s = 0; s < K; ++s)
      {
        double prob = clv[(size_t)n * R * Kp + r * Kp + s];
        if (scaler)
          for (t = 0; t < scaler[n]; ++t) prob *= PLL_SCALE_THRESHOLD;
        printf("%.*f%s", float_precision, prob, s + 1 < K ? "," : ")");
      }
      if (r + 1 < R) printf(",");
    }
    printf("} ");
  }
  printf("]\n");
}
