/*
 * pll_partition.c - partition lifecycle, tip data and pattern weights for the GPU backend.
 *
 * Mirrors the interface and error behaviour of reference src/pll.c:
 *   pll_partition_create/destroy   reference src/pll.c:399-823
 *   pll_set_tip_states             reference src/pll.c:966-998 (+ charmap :136-397,
 *                                  set_tipchars_4x4 :825-860, set_tipchars :862-903,
 *                                  set_tipclv :905-964)
 *   pll_set_tip_clv                reference src/pll.c:1001-1045
 *   pll_set_pattern_weights        reference src/pll.c:1047-1059
 * but the big arrays (CLVs, scale buffers, tip characters, P-matrices, weights) are allocated
 * in HBM by plg_create and never live on the host; see pll.h for which pll_partition_t
 * fields stay valid host memory.
 */
#include <stdarg.h>

#include "pll_host.h"

PLL_EXPORT __thread int pll_errno;
PLL_EXPORT __thread char pll_errmsg[200] = {0};

/* device used for partitions created by this thread; < 0 = decide at create time */
static __thread int g_device = -1;

int pll_fail(int code, const char * fmt, ...)
{
  va_list ap;
  pll_errno = code;
  va_start(ap, fmt);
  vsnprintf(pll_errmsg, sizeof(pll_errmsg), fmt, ap);
  va_end(ap);
  return PLL_FAILURE;
}

int pllg_fail(int rc, const char * where)
{
  int code;
  switch (rc)
  {
    case PLG_E_NODEVICE: code = PLL_ERROR_GPU_NODEVICE; break;
    case PLG_E_NOMEM: code = PLL_ERROR_MEM_ALLOC; break;
    case PLG_E_INVALID: code = PLL_ERROR_PARAM_INVALID; break;
    case PLG_E_UNSUPPORTED: code = PLL_ERROR_GPU_UNSUPPORTED; break;
    default: code = PLL_ERROR_GPU_RUNTIME; break;
  }
  const char * message = pllg_pending_error();
  return pll_fail(code, "%s: %s", where, message ? message : plg_last_error());
}

PLL_EXPORT void * pll_aligned_alloc(size_t size, size_t alignment)
{
  void * mem = NULL;
  if (alignment < sizeof(void *)) alignment = sizeof(void *);
  if (posix_memalign(&mem, alignment, size ? size : alignment)) mem = NULL;
  return mem;
}

PLL_EXPORT void pll_aligned_free(void * ptr) { free(ptr); }

PLL_EXPORT int pll_gpu_set_device(int device)
{
  g_device = device;
  return PLL_SUCCESS;
}

PLL_EXPORT int pll_gpu_device_count(void) { return plg_device_count(); }

static int pick_device(void)
{
  const char * e;
  if (g_device >= 0) return g_device;
  if ((e = getenv("PLL_GPU_DEVICE")) && *e) return atoi(e);
  if ((e = getenv("LOCAL_RANK")) && *e)
  {
    int n = plg_device_count();
    return n > 0 ? atoi(e) % n : 0;
  }
  return -1; /* current CUDA device */
}

int pll_gpu_current_device(void) { return pick_device(); }

PLL_EXPORT int pll_gpu_comm_unique_id(unsigned char id[PLL_GPU_COMM_ID_BYTES])
{
  int rc = plg_comm_unique_id(id);
  return rc ? pllg_fail(rc, "pll_gpu_comm_unique_id") : PLL_SUCCESS;
}

PLL_EXPORT int pll_gpu_comm_init(const unsigned char id[PLL_GPU_COMM_ID_BYTES], int nranks, int rank)
{
  int rc = plg_comm_init(id, nranks, rank, pick_device());
  return rc ? pllg_fail(rc, "pll_gpu_comm_init") : PLL_SUCCESS;
}

PLL_EXPORT int pll_gpu_comm_finalize(void)
{
  plg_comm_finalize();
  return PLL_SUCCESS;
}

int pll_gpu_mirror_mode(void)
{
  const char * e = getenv("PLL_GPU_MIRROR");
  return e && *e && strcmp(e, "0") != 0;
}

/* ------------------------------------------------------------------------------------ */
static void * zalloc_aligned(size_t bytes)
{
  void * p = pll_aligned_alloc(bytes, PLL_ALIGNMENT_GPU);
  if (p) memset(p, 0, bytes);
  return p;
}

static void free_host(pllg_partition_t * g)
{
  pll_partition_t * p = &g->pub;
  unsigned int i;
  if (p->clv)
    for (i = 0; i < p->tips + p->clv_buffers; ++i) pll_aligned_free(p->clv[i]);
  free(p->clv);
  if (p->pmatrix) pll_aligned_free(p->pmatrix[0]);
  free(p->pmatrix);
  if (p->scale_buffer)
    for (i = 0; i < p->scale_buffers; ++i) free(p->scale_buffer[i]);
  free(p->scale_buffer);
  if (p->tipchars)
    for (i = 0; i < p->tips; ++i) free(p->tipchars[i]);
  free(p->tipchars);
#define FREE_PER_MATRIX(field)                                                         \
  if (p->field)                                                                        \
    for (i = 0; i < p->rate_matrices; ++i) pll_aligned_free(p->field[i]);              \
  free(p->field)
  FREE_PER_MATRIX(eigenvecs);
  FREE_PER_MATRIX(inv_eigenvecs);
  FREE_PER_MATRIX(eigenvals);
  FREE_PER_MATRIX(subst_params);
  FREE_PER_MATRIX(frequencies);
#undef FREE_PER_MATRIX
  free(p->rates);
  free(p->rate_weights);
  free(p->prop_invar);
  free(p->invariant);
  free(p->pattern_weights);
  free(p->eigen_decomp_valid);
  free(p->charmap);
  free(p->tipmap);
  pll_aligned_free(p->ttlookup);
  free(g->tip_stage);
  pllg_dev_destroy(g);
  g->magic = 0;
  free(g);
}

PLL_EXPORT pll_partition_t * pll_partition_create(unsigned int tips,
                                                  unsigned int clv_buffers,
                                                  unsigned int states,
                                                  unsigned int sites,
                                                  unsigned int rate_matrices,
                                                  unsigned int prob_matrices,
                                                  unsigned int rate_cats,
                                                  unsigned int scale_buffers,
                                                  unsigned int attributes)
{
  unsigned int i;

  /* PLL_GPU_FORCE=1: relink-only drop-in.  A program built against the reference's pll.h asks
   * for PLL_ATTRIB_ARCH_CPU/SSE/AVX/AVX2 (it cannot know the new flag); with this switch its
   * architecture bits are replaced by PLL_ATTRIB_ARCH_GPU, so unmodified callers - the
   * reference's own test/src programs, see tests/test_reference_programs_gpu.py - run on the
   * device.  Still no CPU path: without the switch such a request fails below. */
  {
    const char * force = getenv("PLL_GPU_FORCE");
    if (force && *force && strcmp(force, "0") != 0)
      attributes = (attributes & ~(unsigned int)PLL_ATTRIB_ARCH_MASK) | PLL_ATTRIB_ARCH_GPU;
  }

  if (!(attributes & PLL_ATTRIB_ARCH_GPU))
  {
    pll_fail(PLL_ERROR_PARAM_INVALID,
             "This build implements PLL_ATTRIB_ARCH_GPU only (no CPU/SIMD path, no fallback).");
    return NULL;
  }
  if (attributes & PLL_ATTRIB_ARCH_MASK)
  {
    pll_fail(PLL_ERROR_PARAM_INVALID, "Multiple architecture flags specified.");
    return NULL;
  }
  if (rate_matrices == 0 || rate_cats == 0 || states == 0 || sites == 0)
  {
    pll_fail(PLL_ERROR_PARAM_INVALID, "Invalid partition dimensions.");
    return NULL;
  }

  pllg_partition_t * g = (pllg_partition_t *)calloc(1, sizeof(pllg_partition_t));
  if (!g)
  {
    pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate memory for partition.");
    return NULL;
  }
  g->magic = PLLG_MAGIC;
  pll_partition_t * p = &g->pub;

  p->tips = tips;
  p->clv_buffers = clv_buffers;
  p->states = states;
  p->sites = sites;
  p->pattern_weight_sum = sites;
  p->rate_matrices = rate_matrices;
  p->prob_matrices = prob_matrices;
  p->rate_cats = rate_cats;
  p->scale_buffers = scale_buffers;
  p->attributes = attributes;
  p->alignment = PLL_ALIGNMENT_GPU;
  /* 256-bit device accesses want a multiple of 4 doubles per (site, rate): same padding rule
   * as the AVX layouts (reference src/pll.c:440-453) */
  p->states_padded = (states + 3) & 0xFFFFFFFCu;
  /* ascertainment-bias storage: `states` extra sites (reference src/pll.c:492-495) */
  p->asc_bias_alloc = (attributes & (PLL_ATTRIB_AB_MASK | PLL_ATTRIB_AB_FLAG)) != 0;
  g->sites_alloc = p->asc_bias_alloc ? sites + states : sites;
  const unsigned int Kp = p->states_padded;

  /* ---- device state ---- */
  plg_dims_t dims;
  dims.tips = tips;
  dims.clv_buffers = clv_buffers;
  dims.states = states;
  dims.states_padded = Kp;
  dims.sites = g->sites_alloc;
  dims.rate_cats = rate_cats;
  dims.rate_matrices = rate_matrices;
  dims.prob_matrices = prob_matrices;
  dims.scale_buffers = scale_buffers;
  dims.attributes = attributes & (PLL_ATTRIB_PATTERN_TIP | PLL_ATTRIB_RATE_SCALERS);
  /* one device context per pattern slice (pll_devices.c); a single one unless the caller asked
   * for more with pll_gpu_set_devices / PLL_GPU_DEVICES */
  int slices = pll_gpu_current_slices();
  int rc = pllg_dev_create(g, &dims, pick_device(), slices);
  if (rc != PLG_OK)
  {
    pllg_fail(rc, "pll_partition_create");
    free_host(g);
    return NULL;
  }

  /* ---- small host-side arrays (always valid) and NULL-initialised mirrors ---- */
  int ok = 1;
  ok &= (p->eigen_decomp_valid = (int *)calloc(rate_matrices, sizeof(int))) != NULL;
  ok &= (p->clv = (double **)calloc((size_t)tips + clv_buffers, sizeof(double *))) != NULL;
  ok &= (p->scale_buffer = (unsigned int **)calloc(scale_buffers ? scale_buffers : 1,
                                                   sizeof(unsigned int *))) != NULL;
  ok &= (p->pmatrix = (double **)calloc(prob_matrices ? prob_matrices : 1,
                                        sizeof(double *))) != NULL;
  ok &= (p->eigenvecs = (double **)calloc(rate_matrices, sizeof(double *))) != NULL;
  ok &= (p->inv_eigenvecs = (double **)calloc(rate_matrices, sizeof(double *))) != NULL;
  ok &= (p->eigenvals = (double **)calloc(rate_matrices, sizeof(double *))) != NULL;
  ok &= (p->subst_params = (double **)calloc(rate_matrices, sizeof(double *))) != NULL;
  ok &= (p->frequencies = (double **)calloc(rate_matrices, sizeof(double *))) != NULL;
  ok &= (p->rates = (double *)calloc(rate_cats, sizeof(double))) != NULL;
  ok &= (p->rate_weights = (double *)calloc(rate_cats, sizeof(double))) != NULL;
  ok &= (p->prop_invar = (double *)calloc(rate_matrices, sizeof(double))) != NULL;
  ok &= (p->pattern_weights = (unsigned int *)malloc((size_t)g->sites_alloc *
                                                     sizeof(unsigned int))) != NULL;
  ok &= (g->tip_stage = (unsigned char *)malloc(g->sites_alloc)) != NULL;
  if (ok && prob_matrices)
  {
    /* one contiguous host mirror for all P-matrices (reference src/pll.c:555-573) */
    const size_t per = (size_t)rate_cats * states * Kp;
    p->pmatrix[0] = (double *)zalloc_aligned(prob_matrices * per * sizeof(double));
    ok &= p->pmatrix[0] != NULL;
    for (i = 1; ok && i < prob_matrices; ++i) p->pmatrix[i] = p->pmatrix[i - 1] + per;
  }
  for (i = 0; ok && i < rate_matrices; ++i)
  {
    ok &= (p->eigenvecs[i] = (double *)zalloc_aligned((size_t)states * Kp * sizeof(double))) != NULL;
    ok &= (p->inv_eigenvecs[i] = (double *)zalloc_aligned((size_t)states * Kp * sizeof(double))) != NULL;
    ok &= (p->eigenvals[i] = (double *)zalloc_aligned(Kp * sizeof(double))) != NULL;
    ok &= (p->subst_params[i] =
               (double *)zalloc_aligned(((size_t)states * (states - 1) / 2 + 1) * sizeof(double))) != NULL;
    ok &= (p->frequencies[i] = (double *)zalloc_aligned(Kp * sizeof(double))) != NULL;
  }
  if (!ok)
  {
    free_host(g);
    pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  for (i = 0; i < rate_cats; ++i) p->rate_weights[i] = 1.0 / rate_cats;
  for (i = 0; i < sites; ++i) p->pattern_weights[i] = 1;
  for (i = sites; i < g->sites_alloc; ++i) p->pattern_weights[i] = 0;
  if (p->asc_bias_alloc)
  {
    /* reductions cover the real sites; the per-state sites only feed the correction terms */
    rc = pllg_dev_set_active_sites(g, sites);
    if (!rc) rc = pllg_dev_set_pattern_weights(g, p->pattern_weights);
    if (rc)
    {
      pllg_fail(rc, "pll_partition_create");
      free_host(g);
      return NULL;
    }
  }
  return p;
}

PLL_EXPORT void pll_partition_destroy(pll_partition_t * partition)
{
  pllg_partition_t * g = pllg_from(partition);
  if (g) free_host(g);
}

PLL_EXPORT plg_context_t * pll_gpu_context(const pll_partition_t * partition)
{
  pllg_partition_t * g = pllg_from(partition);
  return g ? g->ctx : NULL;
}

/* ------------------------------------------------------------------------------------ */
/* character map for pattern tips                                                        */
/* ------------------------------------------------------------------------------------ */
/* Assigns one code per distinct state mask, in order of first appearance over the ASCII
 * range, keeping codes handed out by earlier calls: the numbering create_charmap /
 * update_charmap produce (reference src/pll.c:136-263, 272-397). */
static int merge_charmap(pll_partition_t * p, const unsigned int * map)
{
  unsigned int i, j, k = 0;

  if (!p->charmap)
  {
    p->charmap = (unsigned char *)calloc(PLL_ASCII_SIZE, sizeof(unsigned char));
    p->tipmap = (unsigned int *)calloc(PLL_ASCII_SIZE, sizeof(unsigned int));
    if (!p->charmap || !p->tipmap)
      return pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate charmap for tip-tip precomputation.");
  }
  while (k < PLL_ASCII_SIZE && p->tipmap[k]) ++k;

  /* count the masks not yet known; 256 or more codes cannot be stored in a byte */
  unsigned int fresh = 0;
  for (i = 0; i < PLL_ASCII_SIZE; ++i)
  {
    if (!map[i]) continue;
    for (j = 0; j < k; ++j)
      if (p->tipmap[j] == map[i]) break;
    if (j < k) continue;
    for (j = 0; j < i; ++j)
      if (map[j] == map[i]) break;
    if (j == i) ++fresh;
  }
  memset(p->charmap, 0, PLL_ASCII_SIZE);
  if (fresh + k >= PLL_ASCII_SIZE)
  {
    snprintf(pll_errmsg, sizeof(pll_errmsg),
             "Cannot specify 256 or more states with PLL_ATTRIB_PATTERN_TIP.");
    return PLL_FAILURE;
  }

  for (i = 0; i < PLL_ASCII_SIZE; ++i)
  {
    if (!map[i]) continue;
    for (j = 0; j < k; ++j)
      if (p->tipmap[j] == map[i]) break;
    if (j == k) p->tipmap[k++] = map[i];
    p->charmap[i] = (unsigned char)j;
  }

  if (p->states == 4)
  {
    /* DNA stores the raw 4-bit mask as the tip character: table size = largest mask + 1 */
    unsigned int m = 0;
    for (i = 0; i < k; ++i)
      if (p->tipmap[i] > m) m = p->tipmap[i];
    p->maxstates = m + 1;
  }
  else
    p->maxstates = k;
  return PLL_SUCCESS;
}

/* ascertainment-bias storage of a tip CLV: dummy site j is the unit vector of state j in every
 * rate category (reference src/pll.c:942-961, 1028-1043); `clv` is zero-initialised */
static void fill_asc_sites(const pllg_partition_t * g, double * clv)
{
  const pll_partition_t * p = &g->pub;
  const size_t span = (size_t)p->rate_cats * p->states_padded;
  for (unsigned int j = 0; j < g->sites_alloc - p->sites; ++j)
    for (unsigned int r = 0; r < p->rate_cats; ++r)
      clv[(size_t)(p->sites + j) * span + (size_t)r * p->states_padded + j] = 1.0;
}

static int illegal_state(char c)
{
  return pll_fail(PLL_ERROR_TIPDATA_ILLEGALSTATE, "Illegal state code in tip \"%c\"", c);
}

PLL_EXPORT int pll_set_tip_states(pll_partition_t * partition,
                                  unsigned int tip_index,
                                  const unsigned int * map,
                                  const char * sequence)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  unsigned int i, j;
  int rc;

  if (tip_index >= p->tips) return pll_fail(PLL_ERROR_PARAM_INVALID, "Invalid tip index %u", tip_index);

  if (p->attributes & PLL_ATTRIB_PATTERN_TIP)
  {
    if (!merge_charmap(p, map)) return PLL_FAILURE;
    if (!p->tipchars)
    {
      p->tipchars = (unsigned char **)calloc(p->tips, sizeof(unsigned char *));
      if (!p->tipchars)
        return pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate space for storing tip characters.");
    }
    for (i = 0; i < p->sites; ++i)
    {
      const unsigned int c = map[(unsigned char)sequence[i]];
      if (c == 0) return illegal_state(sequence[i]);
      g->tip_stage[i] = (p->states == 4) ? (unsigned char)c
                                         : p->charmap[(unsigned char)sequence[i]];
    }
    /* ascertainment-bias storage: dummy site j shows the pure state j (reference
     * src/pll.c:847-856).  Alphabets other than DNA store charmap codes, so the code whose
     * state set is exactly {j} is looked up; the reference stores an ASCII character there
     * (src/pll.c:885-903), which indexes past its lookup tables - not reproduced. */
    for (j = 0; j < g->sites_alloc - p->sites; ++j)
    {
      if (p->states == 4)
        g->tip_stage[p->sites + j] = (unsigned char)(1u << j);
      else
      {
        unsigned int code;
        for (code = 0; code < p->maxstates; ++code)
          if (p->tipmap[code] == (1u << j)) break;
        if (code == p->maxstates)
          return pll_fail(PLL_ERROR_AB_NOSUPPORT,
                          "The character map has no symbol for state %u alone (needed for ascertainment "
                          "bias correction with PLL_ATTRIB_PATTERN_TIP).", j);
        g->tip_stage[p->sites + j] = (unsigned char)code;
      }
    }
    if ((rc = pllg_dev_set_tipmap(g, p->tipmap, p->states == 4 ? 16u : p->maxstates)))
      return pllg_fail(rc, "pll_set_tip_states");
    if ((rc = pllg_dev_set_tipchars(g, tip_index, g->tip_stage)))
      return pllg_fail(rc, "pll_set_tip_states");
    /* the staging buffer is reused by the next call */
    if ((rc = pllg_dev_synchronize(g))) return pllg_fail(rc, "pll_set_tip_states");
    if (pll_gpu_mirror_mode()) return pll_gpu_sync_tipchars(partition, tip_index);
    return PLL_SUCCESS;
  }

  /* tips as full CLVs: 0/1 entries replicated over the rate categories */
  const size_t span = (size_t)p->rate_cats * p->states_padded;
  double * clv = (double *)calloc((size_t)g->sites_alloc * span, sizeof(double));
  if (!clv) return pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate a host tip CLV.");
  fill_asc_sites(g, clv);
  for (i = 0; i < p->sites; ++i)
  {
    unsigned int c = map[(unsigned char)sequence[i]];
    if (c == 0)
    {
      free(clv);
      return illegal_state(sequence[i]);
    }
    double * site = clv + i * span;
    for (j = 0; j < p->states; ++j, c >>= 1) site[j] = (double)(c & 1u);
    for (j = 1; j < p->rate_cats; ++j)
      memcpy(site + j * p->states_padded, site, p->states * sizeof(double));
  }
  rc = pllg_dev_set_clv(g, tip_index, clv);
  if (!rc) rc = pllg_dev_synchronize(g);
  free(clv);
  if (rc) return pllg_fail(rc, "pll_set_tip_states");
  return pll_gpu_mirror_mode() ? pll_gpu_sync_clv(partition, tip_index) : PLL_SUCCESS;
}

PLL_EXPORT int pll_gpu_generate_tip_states(pll_partition_t * partition, unsigned int tip_index,
                                           unsigned long long seed, unsigned long long first_site)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  int rc;
  if (tip_index >= p->tips) return pll_fail(PLL_ERROR_PARAM_INVALID, "Invalid tip index %u", tip_index);
  if (!(p->attributes & PLL_ATTRIB_PATTERN_TIP) || p->states != 4 || g->sites_alloc != p->sites)
    return pll_fail(PLL_ERROR_PARAM_INVALID,
                    "pll_gpu_generate_tip_states: 4-state pattern-tip partitions without ascertainment bias only");
  if (!merge_charmap(p, pll_map_nt)) return PLL_FAILURE;
  if (!p->tipchars)
  {
    p->tipchars = (unsigned char **)calloc(p->tips, sizeof(unsigned char *));
    if (!p->tipchars)
      return pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate space for storing tip characters.");
  }
  if ((rc = pllg_dev_set_tipmap(g, p->tipmap, 16u))) return pllg_fail(rc, "pll_gpu_generate_tip_states");
  if ((rc = pllg_dev_generate_tipchars(g, tip_index, seed, first_site)))
    return pllg_fail(rc, "pll_gpu_generate_tip_states");
  if (pll_gpu_mirror_mode()) return pll_gpu_sync_tipchars(partition, tip_index);
  return PLL_SUCCESS;
}

PLL_EXPORT int pll_set_tip_clv(pll_partition_t * partition,
                               unsigned int tip_index,
                               const double * clv,
                               int padding)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  unsigned int i, j;

  if (p->attributes & PLL_ATTRIB_PATTERN_TIP)
    return pll_fail(PLL_ERROR_TIPDATA_ILLEGALFUNCTION,
                    "Cannot use pll_set_tip_clv with PLL_ATTRIB_PATTERN_TIP.");
  if (tip_index >= p->tips) return pll_fail(PLL_ERROR_PARAM_INVALID, "Invalid tip index %u", tip_index);

  const size_t span = (size_t)p->rate_cats * p->states_padded;
  double * full = (double *)calloc((size_t)g->sites_alloc * span, sizeof(double));
  if (!full) return pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate a host tip CLV.");
  fill_asc_sites(g, full);
  for (i = 0; i < p->sites; ++i)
  {
    for (j = 0; j < p->rate_cats; ++j)
      memcpy(full + i * span + j * p->states_padded, clv, p->states * sizeof(double));
    clv += padding ? p->states_padded : p->states;
  }
  int rc = pllg_dev_set_clv(g, tip_index, full);
  if (!rc) rc = pllg_dev_synchronize(g);
  free(full);
  if (rc) return pllg_fail(rc, "pll_set_tip_clv");
  return pll_gpu_mirror_mode() ? pll_gpu_sync_clv(partition, tip_index) : PLL_SUCCESS;
}

PLL_EXPORT void pll_set_pattern_weights(pll_partition_t * partition,
                                        const unsigned int * pattern_weights)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return;
  pll_partition_t * p = &g->pub;
  unsigned int i;
  memcpy(p->pattern_weights, pattern_weights, sizeof(unsigned int) * p->sites);
  p->pattern_weight_sum = 0;
  for (i = 0; i < p->sites; ++i) p->pattern_weight_sum += pattern_weights[i];
  int rc = pllg_dev_set_pattern_weights(g, p->pattern_weights);
  if (!rc) rc = pllg_dev_synchronize(g);
  if (rc) pllg_fail(rc, "pll_set_pattern_weights");
}

/* ------------------------------------------------------------------------------------ */
/* host mirrors                                                                          */
/* ------------------------------------------------------------------------------------ */
PLL_EXPORT int pll_gpu_sync_clv(pll_partition_t * partition, unsigned int clv_index)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  if (clv_index >= p->tips + p->clv_buffers ||
      ((p->attributes & PLL_ATTRIB_PATTERN_TIP) && clv_index < p->tips))
    return pll_fail(PLL_ERROR_PARAM_INVALID, "CLV %u has no storage", clv_index);
  const size_t bytes = (size_t)g->sites_alloc * p->rate_cats * p->states_padded * sizeof(double);
  if (!p->clv[clv_index] && !(p->clv[clv_index] = (double *)zalloc_aligned(bytes)))
    return pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate the host mirror of a CLV.");
  int rc = pllg_dev_get_clv(g, clv_index, p->clv[clv_index]);
  return rc ? pllg_fail(rc, "pll_gpu_sync_clv") : PLL_SUCCESS;
}

PLL_EXPORT int pll_gpu_push_clv(pll_partition_t * partition, unsigned int clv_index)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  if (clv_index >= p->tips + p->clv_buffers || !p->clv[clv_index])
    return pll_fail(PLL_ERROR_PARAM_INVALID, "CLV %u has no host mirror", clv_index);
  int rc = pllg_dev_set_clv(g, clv_index, p->clv[clv_index]);
  if (!rc) rc = pllg_dev_synchronize(g);
  return rc ? pllg_fail(rc, "pll_gpu_push_clv") : PLL_SUCCESS;
}

PLL_EXPORT int pll_gpu_sync_scaler(pll_partition_t * partition, unsigned int scaler_index)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  if (scaler_index >= p->scale_buffers)
    return pll_fail(PLL_ERROR_PARAM_INVALID, "Invalid scaler index %u", scaler_index);
  const size_t n = (p->attributes & PLL_ATTRIB_RATE_SCALERS)
                       ? (size_t)g->sites_alloc * p->rate_cats : g->sites_alloc;
  if (!p->scale_buffer[scaler_index] &&
      !(p->scale_buffer[scaler_index] = (unsigned int *)calloc(n, sizeof(unsigned int))))
    return pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate the host mirror of a scale buffer.");
  int rc = pllg_dev_get_scaler(g, scaler_index, p->scale_buffer[scaler_index]);
  return rc ? pllg_fail(rc, "pll_gpu_sync_scaler") : PLL_SUCCESS;
}

PLL_EXPORT int pll_gpu_sync_tipchars(pll_partition_t * partition, unsigned int tip_index)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  pll_partition_t * p = &g->pub;
  if (!(p->attributes & PLL_ATTRIB_PATTERN_TIP) || tip_index >= p->tips || !p->tipchars)
    return pll_fail(PLL_ERROR_PARAM_INVALID, "Tip %u has no character storage", tip_index);
  if (!p->tipchars[tip_index] &&
      !(p->tipchars[tip_index] = (unsigned char *)malloc(g->sites_alloc)))
    return pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate the host mirror of tip characters.");
  int rc = pllg_dev_get_tipchars(g, tip_index, p->tipchars[tip_index]);
  return rc ? pllg_fail(rc, "pll_gpu_sync_tipchars") : PLL_SUCCESS;
}

PLL_EXPORT int pll_gpu_sync_pmatrix(pll_partition_t * partition, unsigned int matrix_index)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  if (matrix_index >= g->pub.prob_matrices)
    return pll_fail(PLL_ERROR_PARAM_INVALID, "Invalid matrix index %u", matrix_index);
  int rc = plg_get_pmatrix(g->ctx, matrix_index, g->pub.pmatrix[matrix_index]);
  return rc ? pllg_fail(rc, "pll_gpu_sync_pmatrix") : PLL_SUCCESS;
}

PLL_EXPORT int pll_gpu_push_pmatrix(pll_partition_t * partition, unsigned int matrix_index)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  if (matrix_index >= g->pub.prob_matrices)
    return pll_fail(PLL_ERROR_PARAM_INVALID, "Invalid matrix index %u", matrix_index);
  int rc = pllg_dev_set_pmatrix(g, matrix_index, g->pub.pmatrix[matrix_index]);
  if (!rc) rc = pllg_dev_synchronize(g);
  return rc ? pllg_fail(rc, "pll_gpu_push_pmatrix") : PLL_SUCCESS;
}

PLL_EXPORT int pll_gpu_synchronize(pll_partition_t * partition)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "Not a GPU partition.");
  int rc = pllg_dev_synchronize(g);
  return rc ? pllg_fail(rc, "pll_gpu_synchronize") : PLL_SUCCESS;
}
