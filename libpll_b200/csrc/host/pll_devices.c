/*
 * pll_devices.c - one pll_partition_t over several GPUs of one process.
 *
 * Site patterns are the independent unit of every kernel on the path (reference
 * src/core_partials_avx.c:421-529, src/core_likelihood_avx.c:1166-1260 and
 * src/core_derivatives_avx2.c:634-766 all loop `for n < sites` with no cross-site dependence;
 * only `logl +=`, `d_f +=`, `dd_f +=` couple sites).  A partition created after
 * pll_gpu_set_devices(n) (or with PLL_GPU_DEVICES=n in the environment) therefore owns n device
 * contexts; context d holds the contiguous pattern slice [lo[d], lo[d+1]) of EVERY CLV, scale
 * buffer, tip row, weight / invariant array and sumtable, while P-matrices, tip maps and the
 * operations list are replicated (KBs; the P-matrix kernel simply runs on every device).
 * Setters and pll_update_partials enqueue on every context's stream and return, so the devices
 * work concurrently; log-likelihoods and derivatives are the sum of the per-device partial
 * results, added on the host in device order (a 1- or 2-double "all-reduce": nothing else ever
 * crosses between GPUs).  Slices are placed round-robin over the visible devices, so n may
 * exceed the device count (the slicing logic is testable on one GPU).
 *
 * Every pllg_dev_* function has the signature of the plg_* function it fans out to, with the
 * partition wrapper in place of the context; host arrays indexed by site are passed with the
 * slice's offset.  With one device they reduce to the single plg_* call.
 */
#include "pll_host.h"
#include <pthread.h>
#include <time.h>

static __thread int g_slices = 0; /* 0 = not set by pll_gpu_set_devices: look at the environment */

PLL_EXPORT int pll_gpu_set_devices(int count)
{
  if (count < 0 || count > PLLG_MAX_DEVICES)
    return pll_fail(PLL_ERROR_PARAM_INVALID, "pll_gpu_set_devices: count must be 0..%d", PLLG_MAX_DEVICES);
  g_slices = count;
  return PLL_SUCCESS;
}

/* sets this thread's slice count and returns the previous raw setting (scratch partitions of
 * pll_core.c live on one device whatever the caller selected for its own partitions) */
int pllg_swap_slices(int count)
{
  const int old = g_slices;
  g_slices = count;
  return old;
}

int pll_gpu_current_slices(void)
{
  if (g_slices > 0) return g_slices;
  const char * e = getenv("PLL_GPU_DEVICES");
  if (e && *e)
  {
    int n = atoi(e);
    if (n >= 1) return n > PLLG_MAX_DEVICES ? PLLG_MAX_DEVICES : n;
  }
  return 1;
}

PLL_EXPORT int pll_gpu_partition_devices(const pll_partition_t * partition)
{
  pllg_partition_t * g = pllg_from(partition);
  return g ? (int)g->ndev : 0;
}

PLL_EXPORT plg_context_t * pll_gpu_context_of(const pll_partition_t * partition, unsigned int slice,
                                              unsigned int * first_site, unsigned int * sites)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g || slice >= g->ndev) return NULL;
  if (first_site) *first_site = g->lo[slice];
  if (sites) *sites = g->lo[slice + 1] - g->lo[slice];
  return g->ctxs[slice];
}

/* Slice boundaries are multiples of 64 patterns (256-bit vector accesses, whole warp tiles):
 * equal slices of ceil(sites / slices) rounded up to 64, the last one takes what is left and
 * slices that would be empty are dropped.  Writes first_site[0 .. n] (first_site[n] = sites) and
 * returns n. */
PLL_EXPORT unsigned int pll_gpu_slice_bounds(unsigned int sites, unsigned int slices,
                                             unsigned int * first_site)
{
  if (slices < 1) slices = 1;
  if (slices > PLLG_MAX_DEVICES) slices = PLLG_MAX_DEVICES;
  unsigned long long per = ((unsigned long long)sites + slices - 1) / slices;
  per = (per + 63u) & ~63ull;
  if (per == 0) per = 64;
  unsigned int n = (unsigned int)(((unsigned long long)sites + per - 1) / per);
  if (n < 1) n = 1;
  for (unsigned int d = 0; d <= n; ++d)
  {
    const unsigned long long at = d * per;
    first_site[d] = at < sites ? (unsigned int)at : sites;
  }
  return n;
}

/* ---- helper threads -------------------------------------------------------------------------
 * A call on a sliced partition is one launch per device.  Issued from one thread they queue up
 * behind each other (~4 us each: a derivative call on 8 B200s took 65 us against 39 us on one);
 * with one helper thread per further slice the launches go out side by side.  A helper spins on
 * a generation counter while calls keep coming (a Newton loop issues one every ~45 us) and backs
 * off to 100 us naps after ~1 ms of silence.  Every device context is only ever touched by one
 * thread at a time: the caller posts a job, runs slice 0 itself and waits for all helpers before
 * it returns.  PLL_GPU_HOST_THREADS=0 keeps everything on the calling thread. */
typedef int (*pllg_job_fn)(pllg_partition_t * g, unsigned int d, void * args);

typedef struct pllg_worker
{
  pthread_t thread;
  struct pllg_pool * pool;
  unsigned int d;
  unsigned long long done; /* last generation this helper has finished */
  int rc;
  char err[256];
} pllg_worker_t;

typedef struct pllg_pool
{
  pllg_partition_t * g;
  unsigned int n; /* helpers: slices 1 .. n */
  unsigned long long gen;
  int stop;
  pllg_job_fn fn;
  void * args;
  pllg_worker_t w[PLLG_MAX_DEVICES];
} pllg_pool_t;

#if defined(__x86_64__) || defined(__i386__)
#define pllg_cpu_relax() __builtin_ia32_pause()
#else
#define pllg_cpu_relax() ((void)0)
#endif

static __thread char g_pending_error[256];
static __thread int g_pending = 0;

const char * pllg_pending_error(void)
{
  if (!g_pending) return NULL;
  g_pending = 0;
  return g_pending_error;
}

static void * worker_main(void * arg)
{
  pllg_worker_t * w = (pllg_worker_t *)arg;
  pllg_pool_t * pool = w->pool;
  unsigned long long seen = 0;
  for (;;)
  {
    unsigned long long cur;
    unsigned int spins = 0;
    while ((cur = __atomic_load_n(&pool->gen, __ATOMIC_ACQUIRE)) == seen)
    {
      if (__atomic_load_n(&pool->stop, __ATOMIC_ACQUIRE)) return NULL;
      if (++spins < 200000u)
        pllg_cpu_relax();
      else
      {
        const struct timespec nap = {0, 100000};
        nanosleep(&nap, NULL);
      }
    }
    seen = cur;
    w->rc = pool->fn(pool->g, w->d, pool->args);
    if (w->rc)
    {
      strncpy(w->err, plg_last_error(), sizeof(w->err) - 1);
      w->err[sizeof(w->err) - 1] = 0;
    }
    __atomic_store_n(&w->done, seen, __ATOMIC_RELEASE);
  }
}

static void pool_destroy(pllg_partition_t * g)
{
  pllg_pool_t * pool = g->pool;
  if (!pool) return;
  __atomic_store_n(&pool->stop, 1, __ATOMIC_RELEASE);
  for (unsigned int i = 0; i < pool->n; ++i) pthread_join(pool->w[i].thread, NULL);
  free(pool);
  g->pool = NULL;
}

static void pool_create(pllg_partition_t * g)
{
  g->pool = NULL;
  const char * e = getenv("PLL_GPU_HOST_THREADS");
  if (g->ndev < 2 || (e && *e == '0')) return;
  pllg_pool_t * pool = (pllg_pool_t *)calloc(1, sizeof(pllg_pool_t));
  if (!pool) return; /* no helpers: the calling thread does it all */
  pool->g = g;
  for (unsigned int i = 0; i + 1 < g->ndev; ++i)
  {
    pool->w[i].pool = pool;
    pool->w[i].d = i + 1;
    if (pthread_create(&pool->w[i].thread, NULL, worker_main, &pool->w[i])) break;
    pool->n = i + 1;
  }
  if (pool->n + 1 < g->ndev)
  {
    /* could not start them all: none */
    g->pool = pool;
    pool_destroy(g);
    return;
  }
  g->pool = pool;
}

/* fn(g, d, args) for every slice d; the first failure (in slice order) is reported */
static int run_on_slices(pllg_partition_t * g, pllg_job_fn fn, void * args)
{
  /* (the device-side reduction hands results from member to member: its calls stay in order) */
  pllg_pool_t * pool = g->grouped ? NULL : g->pool;
  if (!pool)
  {
    for (unsigned int d = 0; d < g->ndev; ++d)
    {
      int rc = fn(g, d, args);
      if (rc) return rc;
    }
    return PLG_OK;
  }
  pool->fn = fn;
  pool->args = args;
  const unsigned long long gen = pool->gen + 1;
  __atomic_store_n(&pool->gen, gen, __ATOMIC_RELEASE);
  int rc = fn(g, 0, args);
  for (unsigned int i = 0; i < pool->n; ++i)
  {
    pllg_worker_t * w = &pool->w[i];
    while (__atomic_load_n(&w->done, __ATOMIC_ACQUIRE) != gen) pllg_cpu_relax();
    if (!rc && w->rc)
    {
      rc = w->rc;
      memcpy(g_pending_error, w->err, sizeof(g_pending_error));
      g_pending = 1;
    }
  }
  return rc;
}

int pllg_dev_create(pllg_partition_t * g, const plg_dims_t * dims, int first_device, int slices)
{
  const unsigned int n = pll_gpu_slice_bounds(dims->sites, (unsigned int)(slices < 1 ? 1 : slices), g->lo);
  int visible = plg_device_count();
  if (visible < 1) visible = 1;

  g->ndev = 0;
  for (unsigned int d = 0; d < n; ++d)
  {
    plg_dims_t slice = *dims;
    slice.sites = g->lo[d + 1] - g->lo[d];
    int device = first_device;
    if (n > 1) device = ((first_device < 0 ? 0 : first_device) + (int)d) % visible;
    int rc = plg_create(&slice, device, &g->ctxs[d]);
    if (rc)
    {
      g->ctxs[d] = NULL;
      pllg_dev_destroy(g);
      return rc;
    }
    g->ndev = d + 1;
  }
  g->ctx = g->ctxs[0];
  /* Scalar results: by default every device writes its partial sums into its own mapped host
   * words and this layer adds them in slice order as their flags arrive (measured on two B200s:
   * 43 us per derivative call).  PLL_GPU_DEVICE_REDUCE=1 lets the devices combine them among
   * themselves instead (plg_group_*: peer-mapped slots on the first device, last arrival adds,
   * ONE host flag) - 50 us: the NVLink round trips of the hand-over cost more than the host's
   * few additions save.  Falls back to the host sum when peer access is missing. */
  g->grouped = 0;
  if (g->ndev > 1)
  {
    const char * e = getenv("PLL_GPU_DEVICE_REDUCE");
    int rc = (e && *e && *e != '0') ? plg_group_create(g->ctxs, g->ndev) : PLG_E_UNSUPPORTED;
    if (rc == PLG_OK) g->grouped = 1;
    else if (rc != PLG_E_UNSUPPORTED)
    {
      pllg_dev_destroy(g);
      return rc;
    }
  }
  pool_create(g);
  return PLG_OK;
}

void pllg_dev_destroy(pllg_partition_t * g)
{
  pool_destroy(g);
  for (unsigned int d = 0; d < g->ndev; ++d)
    if (g->ctxs[d]) plg_destroy(g->ctxs[d]);
  memset(g->ctxs, 0, sizeof(g->ctxs));
  g->ndev = 0;
  g->ctx = NULL;
}

#define FOR_EACH_DEVICE(call)                              \
  do                                                       \
  {                                                        \
    for (unsigned int d = 0; d < g->ndev; ++d)             \
    {                                                      \
      plg_context_t * ctx = g->ctxs[d];                    \
      const size_t lo = g->lo[d];                          \
      (void)lo;                                            \
      int rc_ = (call);                                    \
      if (rc_) return rc_;                                 \
    }                                                      \
    return PLG_OK;                                         \
  } while (0)

/* doubles per pattern of a CLV / sumtable, scaler entries per pattern */
static size_t clv_span(const pllg_partition_t * g)
{
  return (size_t)g->pub.rate_cats * g->pub.states_padded;
}
static size_t scaler_span(const pllg_partition_t * g)
{
  return (g->pub.attributes & PLL_ATTRIB_RATE_SCALERS) ? g->pub.rate_cats : 1u;
}

int pllg_dev_synchronize(pllg_partition_t * g) { FOR_EACH_DEVICE(plg_synchronize(ctx)); }

int pllg_dev_set_tipmap(pllg_partition_t * g, const unsigned int * tipmap, unsigned int maxstates)
{
  FOR_EACH_DEVICE(plg_set_tipmap(ctx, tipmap, maxstates));
}

int pllg_dev_set_tipchars(pllg_partition_t * g, unsigned int tip_index, const unsigned char * chars)
{
  FOR_EACH_DEVICE(plg_set_tipchars(ctx, tip_index, chars + lo));
}

int pllg_dev_generate_tipchars(pllg_partition_t * g, unsigned int tip_index, unsigned long long seed,
                               unsigned long long first_site)
{
  FOR_EACH_DEVICE(plg_generate_tipchars(ctx, tip_index, seed, first_site + lo));
}

int pllg_dev_get_tipchars(pllg_partition_t * g, unsigned int tip_index, unsigned char * chars)
{
  FOR_EACH_DEVICE(plg_get_tipchars(ctx, tip_index, chars + lo));
}

int pllg_dev_set_clv(pllg_partition_t * g, unsigned int clv_index, const double * clv)
{
  const size_t span = clv_span(g);
  FOR_EACH_DEVICE(plg_set_clv(ctx, clv_index, clv + lo * span));
}

int pllg_dev_get_clv(pllg_partition_t * g, unsigned int clv_index, double * clv)
{
  const size_t span = clv_span(g);
  FOR_EACH_DEVICE(plg_get_clv(ctx, clv_index, clv + lo * span));
}

int pllg_dev_get_scaler(pllg_partition_t * g, unsigned int scaler_index, unsigned int * scaler)
{
  const size_t span = scaler_span(g);
  FOR_EACH_DEVICE(plg_get_scaler(ctx, scaler_index, scaler + lo * span));
}

int pllg_dev_set_pattern_weights(pllg_partition_t * g, const unsigned int * weights)
{
  FOR_EACH_DEVICE(plg_set_pattern_weights(ctx, weights + lo));
}

/* The leading `count` patterns of the partition take part in the lnL / derivative reductions
 * (ascertainment bias: the per-state sites behind the last real pattern do not): every slice gets
 * its share of that prefix. */
int pllg_dev_set_active_sites(pllg_partition_t * g, unsigned int count)
{
  for (unsigned int d = 0; d < g->ndev; ++d)
  {
    const unsigned int lo = g->lo[d], hi = g->lo[d + 1];
    const unsigned int mine = count <= lo ? 0u : (count >= hi ? hi - lo : count - lo);
    int rc = plg_set_active_sites(g->ctxs[d], mine);
    if (rc) return rc;
  }
  return PLG_OK;
}

/* A range of patterns [first, first + count) of a CLV / scale buffer / sumtable, wherever its
 * pieces live (the per-state sites of the ascertainment-bias correction sit in the last slice, or
 * straddle the last two) */
#define FOR_EACH_PIECE(call)                                                   \
  do                                                                           \
  {                                                                            \
    for (unsigned int d = 0; d < g->ndev; ++d)                                 \
    {                                                                          \
      const unsigned int a = first > g->lo[d] ? first : g->lo[d];              \
      const unsigned int b = first + count < g->lo[d + 1] ? first + count : g->lo[d + 1]; \
      if (a >= b) continue;                                                    \
      plg_context_t * ctx = g->ctxs[d];                                        \
      const unsigned int local = a - g->lo[d], n = b - a;                      \
      const size_t at = a - first;                                             \
      int rc_ = (call);                                                        \
      if (rc_) return rc_;                                                     \
    }                                                                          \
    return PLG_OK;                                                             \
  } while (0)

int pllg_dev_get_clv_sites(pllg_partition_t * g, unsigned int clv_index, unsigned int first, unsigned int count,
                           double * out)
{
  const size_t span = clv_span(g);
  FOR_EACH_PIECE(plg_get_clv_sites(ctx, clv_index, local, n, out + at * span));
}

int pllg_dev_get_scaler_sites(pllg_partition_t * g, unsigned int scaler_index, unsigned int first,
                              unsigned int count, unsigned int * out)
{
  const size_t span = scaler_span(g);
  FOR_EACH_PIECE(plg_get_scaler_sites(ctx, scaler_index, local, n, out + at * span));
}

int pllg_dev_get_sumtable_sites(pllg_partition_t * g, const void * key, unsigned int first, unsigned int count,
                                double * out)
{
  const size_t span = clv_span(g);
  FOR_EACH_PIECE(plg_get_sumtable_sites(ctx, key, local, n, out + at * span));
}

int pllg_dev_update_invariant(pllg_partition_t * g, int * invariant_out)
{
  FOR_EACH_DEVICE(plg_update_invariant(ctx, invariant_out ? invariant_out + lo : NULL));
}

int pllg_dev_set_pmatrix(pllg_partition_t * g, unsigned int matrix_index, const double * pmatrix)
{
  FOR_EACH_DEVICE(plg_set_pmatrix(ctx, matrix_index, pmatrix));
}

/* the enqueue-only calls of a branch-length loop also go out side by side (their host part -
 * staging, planning, launch - is ~15 us per device) */
struct pmatrix_args
{
  const unsigned int * matrix_indices;
  const double * branch_lengths;
  unsigned int count;
  const double * rates, * prop_invar, * eigenvals, * eigenvecs, * inv_eigenvecs;
};

static int pmatrix_job(pllg_partition_t * g, unsigned int d, void * p)
{
  struct pmatrix_args * a = (struct pmatrix_args *)p;
  return plg_update_pmatrix(g->ctxs[d], a->matrix_indices, a->branch_lengths, a->count, a->rates, a->prop_invar,
                            a->eigenvals, a->eigenvecs, a->inv_eigenvecs);
}

int pllg_dev_update_pmatrix(pllg_partition_t * g, const unsigned int * matrix_indices,
                            const double * branch_lengths, unsigned int count, const double * rates,
                            const double * prop_invar, const double * eigenvals,
                            const double * eigenvecs, const double * inv_eigenvecs)
{
  struct pmatrix_args a = {matrix_indices, branch_lengths, count, rates, prop_invar, eigenvals, eigenvecs, inv_eigenvecs};
  return run_on_slices(g, pmatrix_job, &a);
}

struct partials_args
{
  const pll_operation_t * operations;
  unsigned int count;
};

static int partials_job(pllg_partition_t * g, unsigned int d, void * p)
{
  struct partials_args * a = (struct partials_args *)p;
  return plg_update_partials(g->ctxs[d], a->operations, a->count);
}

int pllg_dev_update_partials(pllg_partition_t * g, const pll_operation_t * operations, unsigned int count)
{
  struct partials_args a = {operations, count};
  return run_on_slices(g, partials_job, &a);
}

/* Device group (plg_group_*): all slices enqueue, the devices add their partial results among
 * themselves and the host waits for one flag. */
static int group_begin(pllg_partition_t * g) { return g->grouped ? plg_group_begin(g->ctxs[0]) : PLG_OK; }
static int group_finish(pllg_partition_t * g, int rc, double * out0, double * out1)
{
  if (rc)
  {
    plg_group_abort(g->ctxs[0]);
    return rc;
  }
  return plg_group_collect(g->ctxs[0], out0, out1);
}

/* Value-returning calls: with several slices the kernels of all devices are enqueued first
 * (plg_set_deferred) and the partial results collected afterwards, so the devices reduce
 * concurrently; partial sums are added in slice order. */
static int begin_deferred(pllg_partition_t * g, unsigned int d)
{
  return g->ndev > 1 ? plg_set_deferred(g->ctxs[d], 1) : PLG_OK;
}

static int collect_all(pllg_partition_t * g, int rc)
{
  if (g->ndev < 2) return rc;
  for (unsigned int d = 0; d < g->ndev; ++d)
  {
    int rc2 = plg_collect(g->ctxs[d]); /* also on failure: nothing may stay pending */
    plg_set_deferred(g->ctxs[d], 0);
    if (!rc) rc = rc2;
  }
  return rc;
}

struct edge_args
{
  unsigned int parent_clv_index, child_clv_index, matrix_index;
  int parent_scaler_index, child_scaler_index;
  const double * freqs, * rate_weights, * prop_invar;
  double * persite_lnl;
  double part[PLLG_MAX_DEVICES];
};

static int edge_job(pllg_partition_t * g, unsigned int d, void * p)
{
  struct edge_args * a = (struct edge_args *)p;
  int rc = g->grouped ? PLG_OK : begin_deferred(g, d);
  if (rc) return rc;
  return plg_edge_loglikelihood(g->ctxs[d], a->parent_clv_index, a->parent_scaler_index, a->child_clv_index,
                                a->child_scaler_index, a->matrix_index, a->freqs, a->rate_weights, a->prop_invar,
                                a->persite_lnl ? a->persite_lnl + g->lo[d] : NULL, &a->part[d]);
}

int pllg_dev_edge_loglikelihood(pllg_partition_t * g, unsigned int parent_clv_index,
                                int parent_scaler_index, unsigned int child_clv_index,
                                int child_scaler_index, unsigned int matrix_index, const double * freqs,
                                const double * rate_weights, const double * prop_invar,
                                double * persite_lnl, double * logl_out)
{
  struct edge_args a = {parent_clv_index, child_clv_index, matrix_index, parent_scaler_index, child_scaler_index,
                        freqs, rate_weights, prop_invar, persite_lnl, {0}};
  int rc = group_begin(g);
  if (rc) return rc;
  rc = run_on_slices(g, edge_job, &a);
  if (g->grouped) return group_finish(g, rc, logl_out, NULL);
  if ((rc = collect_all(g, rc))) return rc;
  double total = a.part[0];
  for (unsigned int d = 1; d < g->ndev; ++d) total += a.part[d];
  *logl_out = total;
  return PLG_OK;
}

struct root_args
{
  unsigned int clv_index;
  int scaler_index;
  const double * freqs, * rate_weights, * prop_invar;
  double * persite_lnl;
  double part[PLLG_MAX_DEVICES];
  const unsigned int * flat; /* see pllg_dev_root_loglikelihood; NULL: the slices' own scale buffers */
};

static int root_job(pllg_partition_t * g, unsigned int d, void * p)
{
  struct root_args * a = (struct root_args *)p;
  int rc = g->grouped ? PLG_OK : begin_deferred(g, d);
  if (rc) return rc;
  if (a->flat)
    return plg_root_loglikelihood_counts(g->ctxs[d], a->clv_index, a->flat + g->lo[d], a->freqs, a->rate_weights,
                                         a->prop_invar, a->persite_lnl ? a->persite_lnl + g->lo[d] : NULL,
                                         &a->part[d]);
  return plg_root_loglikelihood(g->ctxs[d], a->clv_index, a->scaler_index, a->freqs, a->rate_weights, a->prop_invar,
                                a->persite_lnl ? a->persite_lnl + g->lo[d] : NULL, &a->part[d]);
}

int pllg_dev_root_loglikelihood(pllg_partition_t * g, unsigned int clv_index, int scaler_index,
                                const double * freqs, const double * rate_weights,
                                const double * prop_invar, double * persite_lnl, double * logl_out)
{
  struct root_args a = {clv_index, scaler_index, freqs, rate_weights, prop_invar, persite_lnl, {0}, NULL};
  unsigned int * flat = NULL;
  const unsigned int R = g->pub.rate_cats;
  if (g->ndev > 1 && (g->pub.attributes & PLL_ATTRIB_RATE_SCALERS) && R > 1 && scaler_index != PLL_SCALE_BUFFER_NONE)
  {
    /* The reference's root kernels read element n of the per-rate scaler array [site][rate] for
     * pattern n (src/core_likelihood_avx.c:176-178, src/core_likelihood.c:197-198): the count of
     * pattern n / R at rate n % R.  For a slice that starts at pattern lo > 0 those elements belong
     * to an earlier slice, so they are collected here (the first ceil(sites / R) patterns hold them
     * all) and every slice gets the elements [lo, hi) of the flat array. */
    const unsigned int need = (g->sites_alloc + R - 1) / R;
    flat = (unsigned int *)malloc((size_t)need * R * sizeof(unsigned int));
    if (!flat) return PLG_E_NOMEM;
    int grc = pllg_dev_get_scaler_sites(g, (unsigned int)scaler_index, 0, need, flat);
    if (grc)
    {
      free(flat);
      return grc;
    }
    a.flat = flat;
  }
  int rc = group_begin(g);
  if (rc)
  {
    free(flat);
    return rc;
  }
  rc = run_on_slices(g, root_job, &a);
  if (g->grouped)
  {
    rc = group_finish(g, rc, logl_out, NULL);
    free(flat);
    return rc;
  }
  rc = collect_all(g, rc);
  free(flat);
  if (rc) return rc;
  double total = a.part[0];
  for (unsigned int d = 1; d < g->ndev; ++d) total += a.part[d];
  *logl_out = total;
  return PLG_OK;
}

struct sumtable_args
{
  unsigned int parent_clv_index, child_clv_index;
  int parent_scaler_index, child_scaler_index;
  const double * eigenvecs, * left_terms;
  const void * key;
  double * host_copy;
};

static int sumtable_job(pllg_partition_t * g, unsigned int d, void * p)
{
  struct sumtable_args * a = (struct sumtable_args *)p;
  return plg_update_sumtable(g->ctxs[d], a->parent_clv_index, a->child_clv_index, a->parent_scaler_index,
                             a->child_scaler_index, a->eigenvecs, a->left_terms, a->key,
                             a->host_copy ? a->host_copy + g->lo[d] * clv_span(g) : NULL);
}

int pllg_dev_update_sumtable(pllg_partition_t * g, unsigned int parent_clv_index,
                             unsigned int child_clv_index, int parent_scaler_index,
                             int child_scaler_index, const double * eigenvecs,
                             const double * left_terms, const void * key, double * host_copy)
{
  struct sumtable_args a = {parent_clv_index, child_clv_index, parent_scaler_index, child_scaler_index,
                            eigenvecs, left_terms, key, host_copy};
  return run_on_slices(g, sumtable_job, &a);
}

/* Releases the device copy of the sumtable a caller is about to free (the host pointer is only a
 * key, reference src/derivatives.c:164-234 writes through it; see pll.h). */
PLL_EXPORT int pll_gpu_free_sumtable(pll_partition_t * partition, const double * sumtable)
{
  pllg_partition_t * g = pllg_from(partition);
  if (!g) return pll_fail(PLL_ERROR_PARAM_INVALID, "pll_gpu_free_sumtable: not a GPU partition");
  for (unsigned int d = 0; d < g->ndev; ++d)
  {
    int rc = plg_free_sumtable(g->ctxs[d], sumtable);
    if (rc) return pllg_fail(rc, "pll_gpu_free_sumtable");
  }
  return PLL_SUCCESS;
}

struct der_args
{
  const void * key;
  const double * diagptable, * rate_weights, * prop_invar, * freqs;
  double a[PLLG_MAX_DEVICES], b[PLLG_MAX_DEVICES];
};

static int der_job(pllg_partition_t * g, unsigned int d, void * p)
{
  struct der_args * x = (struct der_args *)p;
  int rc = g->grouped ? PLG_OK : begin_deferred(g, d);
  if (rc) return rc;
  return plg_likelihood_derivatives(g->ctxs[d], x->key, x->diagptable, x->rate_weights, x->prop_invar, x->freqs,
                                    &x->a[d], &x->b[d]);
}

int pllg_dev_likelihood_derivatives(pllg_partition_t * g, const void * key, const double * diagptable,
                                    const double * rate_weights, const double * prop_invar,
                                    const double * freqs, double * d_f, double * dd_f)
{
  struct der_args x = {key, diagptable, rate_weights, prop_invar, freqs, {0}, {0}};
  int rc = group_begin(g);
  if (rc) return rc;
  rc = run_on_slices(g, der_job, &x);
  if (g->grouped) return group_finish(g, rc, d_f, dd_f);
  if ((rc = collect_all(g, rc))) return rc;
  double s1 = x.a[0], s2 = x.b[0];
  for (unsigned int d = 1; d < g->ndev; ++d)
  {
    s1 += x.a[d];
    s2 += x.b[d];
  }
  *d_f = s1;
  *dd_f = s2;
  return PLG_OK;
}
