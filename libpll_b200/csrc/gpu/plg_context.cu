/*
 * plg_context.cu - lifetime of the device-resident partition state, uploads/downloads,
 * staging ring, timing and statistics.
 *
 * Replaces the allocation half of pll_partition_create / dealloc_partition_data
 * (reference src/pll.c:31-111, 509-815): instead of one posix_memalign per CLV the device
 * layer makes one HBM slab per array kind and addresses slots by index * stride, which is
 * what lets a whole operation list be described to the kernels by a table of pointers.
 */
#include <cstdarg>
#include <cstdlib>
#include <cmath>

#include "plg_internal.cuh"

/* ------------------------------------------------------------------------------------ */
static thread_local char g_plg_error[512] = "";

void plg_set_error(const char * fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_plg_error, sizeof(g_plg_error), fmt, ap);
  va_end(ap);
}

extern "C" const char * plg_last_error(void) { return g_plg_error; }

extern "C" int plg_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  int usable = 0;
  for (int i = 0; i < n; ++i)
  {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess &&
        major == 10)
      ++usable;
  }
  return usable;
}

static size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

/* ------------------------------------------------------------------------------------ */
extern "C" int plg_create(const plg_dims_t * dims, int device, plg_context_t ** out)
{
  if (!dims || !out)
  {
    plg_set_error("plg_create: NULL argument");
    return PLG_E_INVALID;
  }
  *out = NULL;

  /* 4 and 20 states with 1/2/4/8/16 rate categories run the specialised kernels; any other
   * alphabet / category count runs the generic kernels of plg_generic.cu */
  if (dims->states < 2 || dims->rate_cats == 0)
  {
    plg_set_error("plg_create: states=%u rate_cats=%u is not a valid partition", dims->states,
                  dims->rate_cats);
    return PLG_E_INVALID;
  }
  if (dims->states_padded < dims->states || (dims->states_padded & 3u) ||
      ((dims->states == 4 || dims->states == 20) && dims->states_padded != dims->states))
  {
    plg_set_error("plg_create: states_padded=%u invalid for %u states (multiple of 4, >= states; "
                  "equal to states for 4 and 20)", dims->states_padded, dims->states);
    return PLG_E_INVALID;
  }
  if ((dims->attributes & PLL_ATTRIB_PATTERN_TIP) && dims->states > 32)
  {
    plg_set_error("plg_create: pattern tips need states <= 32 (state sets are 32-bit masks)");
    return PLG_E_UNSUPPORTED;
  }
  if (dims->sites == 0)
  {
    plg_set_error("plg_create: zero sites");
    return PLG_E_INVALID;
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
  {
    cudaGetLastError();
    plg_set_error("plg_create: no CUDA device visible (this backend has no CPU fallback)");
    return PLG_E_NODEVICE;
  }
  if (device < 0) PLG_CUDA(cudaGetDevice(&device));
  if (device >= ndev)
  {
    plg_set_error("plg_create: device %d out of range (%d visible)", device, ndev);
    return PLG_E_NODEVICE;
  }
  int major = 0;
  PLG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10)
  {
    plg_set_error("plg_create: device %d has compute capability %d.x; kernels are built for "
                  "sm_100a only", device, major);
    return PLG_E_NODEVICE;
  }
  PLG_CUDA(cudaSetDevice(device));

  plg_context * ctx = new plg_context();
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  ctx->d = *dims;
  ctx->device = device;
  PLG_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
  ctx->pattern_tip = (dims->attributes & PLL_ATTRIB_PATTERN_TIP) != 0;
  ctx->rate_scalers = (dims->attributes & PLL_ATTRIB_RATE_SCALERS) != 0;
  ctx->active_sites = dims->sites;
  ctx->lnl_scratch = NULL;
  ctx->fused_records = NULL;
  ctx->fused_records_cap = 0;
  ctx->span = (size_t)dims->rate_cats * dims->states_padded;
  ctx->clv_stride = round_up((size_t)dims->sites * ctx->span, 32);
  ctx->scaler_len = ctx->rate_scalers ? (size_t)dims->sites * dims->rate_cats : dims->sites;
  ctx->scaler_stride = round_up(ctx->scaler_len, 64);
  ctx->tip_stride = round_up(dims->sites, 256);
  ctx->pmat_len = (size_t)dims->rate_cats * dims->states * dims->states_padded;
  ctx->clv_first = ctx->pattern_tip ? dims->tips : 0;
  ctx->clv = NULL; ctx->scalers = NULL; ctx->tipchars = NULL; ctx->pmatrix = NULL;
  ctx->weights = NULL; ctx->invariant = NULL; ctx->has_invariant = false;
  ctx->maxstates = 0; ctx->log2_maxstates = 0; ctx->tipmap_epoch = 0; ctx->root_counts = NULL;
  memset(ctx->tipmap, 0, sizeof(ctx->tipmap));
  ctx->stage_host = NULL; ctx->stage_dev = NULL; ctx->stage_size = 0; ctx->stage_off = 0;
  ctx->tables = NULL; ctx->tables_cap = 0; ctx->partials = NULL; ctx->partials_cap = 0;
  ctx->counter = NULL; ctx->result_dev = NULL; ctx->result_host = NULL;
  ctx->deferred = 0; ctx->pending[0] = ctx->pending[1] = NULL;
  ctx->result_seq = 0; ctx->copy_pending = 0; ctx->group = NULL; ctx->group_rank = 0;
  ctx->comm_buf = NULL; ctx->comm_seq = 0;
  ctx->persite_dev = NULL;
  ctx->lnl_table = NULL; ctx->lnl_table_cap = 0;
  ctx->sumtables = new std::unordered_map<const void *, double *>();
  ctx->sumtable_used = new std::unordered_map<const void *, unsigned long long>();
  ctx->sumtable_clock = 0;
  ctx->graphs = new std::unordered_map<uint64_t, plg_graph_entry *>();
  ctx->seen_lists = new std::unordered_map<uint64_t, unsigned int>();
  const char * g = getenv("PLL_GPU_GRAPHS");
  ctx->use_graphs = g ? atoi(g) : 1;
  {
    const char * gc = getenv("PLL_GPU_GRAPH_CACHE");
    int cap = gc ? atoi(gc) : 64;
    ctx->graph_cap = cap < 1 ? 1u : (unsigned int)cap;
  }
  ctx->graph_clock = 0;
  ctx->list_buf = NULL;
  ctx->list_buf_cap = 0;
  const char * ex = getenv("PLL_GPU_AA_EXACT");
  ctx->aa_exact = (ex && *ex && *ex != '0') ? 1 : 0;
  {
    const char * fu = getenv("PLL_GPU_FUSED");
    const char * fs = getenv("PLL_GPU_FUSED_SLOTS");
    ctx->use_fused = fu ? atoi(fu) : 1;
    const char * fa = getenv("PLL_GPU_FUSED_AA");
    ctx->use_fused_aa = fa ? (atoi(fa) ? 1 : 0) : 2; /* unset: the walk for lists that recycle slots */
    ctx->fused_slots = fs ? (unsigned int)atoi(fs) : 3u;
    if (ctx->fused_slots < 1) ctx->fused_slots = 1;
    if (ctx->fused_slots > 3) ctx->fused_slots = 3; /* shared-memory budget of plg_traverse.cu */
  }
  ctx->flush_buf = NULL; ctx->flush_bytes = 0;
  /* L2 residency of the sumtable between derivative passes is opt-in (PLL_GPU_L2_PERSIST=1):
   * setting L2 aside for persisting lines is a DEVICE-wide limit, and with the maximum set aside
   * the traversal kernel of any partition on the device ran at half speed (53 ms instead of
   * 26 ms at BASELINE configs[1]), for a 13 % shorter derivative pass (30.7 us vs 35.3 us). */
  ctx->l2_persist_bytes = 0; ctx->l2_window_max = 0; ctx->l2_pinned = 0;
  {
    const char * lp = getenv("PLL_GPU_L2_PERSIST");
    int max_persist = 0, max_window = 0;
    if (lp && *lp == '1' &&
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device) == cudaSuccess &&
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device) == cudaSuccess &&
        max_persist > 0 && max_window > 0 &&
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) == cudaSuccess)
    {
      ctx->l2_persist_bytes = (size_t)max_persist;
      ctx->l2_window_max = (size_t)max_window;
    }
    cudaGetLastError();
  }
  ctx->stream = NULL; ctx->ev_start = NULL; ctx->ev_stop = NULL;
  ctx->profiling = 0;
  ctx->prof_events = new std::vector<cudaEvent_t>();

#define PLG_CREATE_CUDA(call)                                                          \
  do {                                                                                 \
    cudaError_t err__ = (call);                                                        \
    if (err__ != cudaSuccess) {                                                        \
      plg_set_error("plg_create: %s failed: %s", #call, cudaGetErrorString(err__));    \
      cudaGetLastError();                                                              \
      plg_destroy(ctx);                                                                \
      return (err__ == cudaErrorMemoryAllocation) ? PLG_E_NOMEM : PLG_E_CUDA;          \
    }                                                                                  \
  } while (0)

  PLG_CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  PLG_CREATE_CUDA(cudaEventCreate(&ctx->ev_start));
  PLG_CREATE_CUDA(cudaEventCreate(&ctx->ev_stop));

  const size_t n_clv = (size_t)dims->tips + dims->clv_buffers - ctx->clv_first;
  if (n_clv)
  {
    PLG_CREATE_CUDA(cudaMalloc(&ctx->clv, n_clv * ctx->clv_stride * sizeof(double)));
    PLG_CREATE_CUDA(cudaMemsetAsync(ctx->clv, 0, n_clv * ctx->clv_stride * sizeof(double),
                                    ctx->stream));
  }
  if (dims->scale_buffers)
  {
    size_t bytes = (size_t)dims->scale_buffers * ctx->scaler_stride * sizeof(unsigned int);
    PLG_CREATE_CUDA(cudaMalloc(&ctx->scalers, bytes));
    PLG_CREATE_CUDA(cudaMemsetAsync(ctx->scalers, 0, bytes, ctx->stream));
  }
  if (ctx->pattern_tip && dims->tips)
  {
    size_t bytes = (size_t)dims->tips * ctx->tip_stride;
    PLG_CREATE_CUDA(cudaMalloc(&ctx->tipchars, bytes));
    PLG_CREATE_CUDA(cudaMemsetAsync(ctx->tipchars, 0, bytes, ctx->stream));
  }
  if (dims->prob_matrices)
  {
    size_t bytes = (size_t)dims->prob_matrices * ctx->pmat_len * sizeof(double);
    PLG_CREATE_CUDA(cudaMalloc(&ctx->pmatrix, bytes));
    PLG_CREATE_CUDA(cudaMemsetAsync(ctx->pmatrix, 0, bytes, ctx->stream));
  }
  PLG_CREATE_CUDA(cudaMalloc(&ctx->weights, (size_t)dims->sites * sizeof(unsigned int)));
  PLG_CREATE_CUDA(cudaMalloc(&ctx->invariant, (size_t)dims->sites * sizeof(int)));
  PLG_CREATE_CUDA(cudaMemsetAsync(ctx->invariant, 0xFF, (size_t)dims->sites * sizeof(int),
                                  ctx->stream));

  ctx->stage_size = (size_t)8 << 20;
  PLG_CREATE_CUDA(cudaHostAlloc(&ctx->stage_host, ctx->stage_size, cudaHostAllocDefault));
  PLG_CREATE_CUDA(cudaMalloc(&ctx->stage_dev, ctx->stage_size));
  PLG_CREATE_CUDA(cudaMalloc(&ctx->counter, 64 * sizeof(unsigned int)));
  PLG_CREATE_CUDA(cudaMemsetAsync(ctx->counter, 0, 64 * sizeof(unsigned int), ctx->stream));
  /* scalar results (lnL, d_f, dd_f) are written by the last block straight into mapped pinned
   * host memory: a value-returning call is launch + stream synchronise, no D2H copy */
  PLG_CREATE_CUDA(cudaHostAlloc(&ctx->result_host, 8 * sizeof(double),
                                cudaHostAllocMapped | cudaHostAllocPortable));
  PLG_CREATE_CUDA(cudaHostGetDevicePointer((void **)&ctx->result_dev, ctx->result_host, 0));
  memset(ctx->result_host, 0, 8 * sizeof(double));

  /* one rank of a sharded run (pll_gpu_comm_init): scalar results go through an all-reduce */
  if (plg_comm_covers(device))
  {
    PLG_CREATE_CUDA(cudaMalloc(&ctx->comm_buf, 8 * sizeof(double)));
    PLG_CREATE_CUDA(cudaMemsetAsync(ctx->comm_buf, 0, 8 * sizeof(double), ctx->stream));
  }

  /* pattern weights default to 1 (reference src/pll.c:784) */
  {
    std::vector<unsigned int> ones(dims->sites, 1u);
    PLG_CREATE_CUDA(cudaMemcpyAsync(ctx->weights, ones.data(),
                                    (size_t)dims->sites * sizeof(unsigned int),
                                    cudaMemcpyHostToDevice, ctx->stream));
    PLG_CREATE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
#undef PLG_CREATE_CUDA

  *out = ctx;
  return PLG_OK;
}

static void plg_group_release(plg_context * ctx);

extern "C" void plg_destroy(plg_context_t * ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  plg_group_release(ctx); /* the whole group dissolves with its first member to go */
  cudaSetDevice(ctx->device);
  if (ctx->graphs)
  {
    for (auto & kv : *ctx->graphs)
    {
      if (kv.second->exec) cudaGraphExecDestroy(kv.second->exec);
      cudaFree(kv.second->dev_tables);
      delete kv.second;
    }
    delete ctx->graphs;
    delete ctx->seen_lists;
  }
  if (ctx->sumtables)
  {
    for (auto & kv : *ctx->sumtables) cudaFree(kv.second);
    delete ctx->sumtables;
    delete ctx->sumtable_used;
  }
  cudaFree(ctx->clv);
  cudaFree(ctx->scalers);
  cudaFree(ctx->tipchars);
  cudaFree(ctx->pmatrix);
  cudaFree(ctx->weights);
  cudaFree(ctx->invariant);
  cudaFree(ctx->root_counts);
  cudaFree(ctx->stage_dev);
  cudaFree(ctx->tables);
  cudaFree(ctx->partials);
  cudaFree(ctx->counter);
  cudaFree(ctx->persite_dev);
  cudaFree(ctx->lnl_table);
  cudaFree(ctx->flush_buf);
  if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
  cudaFree(ctx->lnl_scratch);
  cudaFree(ctx->fused_records);
  cudaFree(ctx->list_buf);
  cudaFree(ctx->comm_buf);
  if (ctx->result_host) cudaFreeHost(ctx->result_host);
  if (ctx->prof_events)
  {
    for (cudaEvent_t e : *ctx->prof_events) cudaEventDestroy(e);
    delete ctx->prof_events;
  }
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx;
}

extern "C" int plg_synchronize(plg_context_t * ctx)
{
  PLG_CUDA(cudaStreamSynchronize(ctx->stream));
  return PLG_OK;
}

extern "C" int plg_set_deferred(plg_context_t * ctx, int enable)
{
  PLG_CHECK_CTX(ctx);
  ctx->deferred = enable ? 1 : 0;
  ctx->pending[0] = ctx->pending[1] = NULL;
  return PLG_OK;
}

PlgSink plg_make_sink(plg_context * ctx)
{
  PlgSink s;
  memset(&s, 0, sizeof(s));
  s.result = ctx->result_dev;
  plg_group * g = ctx->group;
  if (g && g->active)
  {
    s.seq = g->seq;
    s.group_slots = g->slots;
    s.group_counter = g->counter;
    s.group_result = g->members[0]->result_dev;
    s.group_size = g->n;
    s.group_rank = ctx->group_rank;
  }
  else
  {
    s.seq = ++ctx->result_seq;
    if (ctx->comm_buf)
    {
      /* sharded over processes: the kernel leaves its sums in device scratch, the all-reduce and
       * the hand-over to the host follow on the stream (plg_finish_result) */
      s.result = ctx->comm_buf;
      s.seq = 0;
      ctx->comm_seq = ctx->result_seq;
    }
  }
  return s;
}

int plg_wait_flag(plg_context * ctx, unsigned long long seq, plg_context * const * watch, unsigned int n_watch)
{
  const unsigned long long * flag = reinterpret_cast<const unsigned long long *>(ctx->result_host + 4);
  /* ~100 us of polling covers a reduction on an idle stream; behind a long traversal the
   * thread sleeps in the driver instead of burning a core */
  for (unsigned int spin = 0; spin < 2500u; ++spin) /* `pause` is ~40 ns on current x86 cores */
  {
    if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) == seq) return PLG_OK;
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  for (unsigned int i = 0; i < n_watch; ++i)
  {
    PLG_CUDA(cudaSetDevice(watch[i]->device));
    PLG_CUDA(cudaStreamSynchronize(watch[i]->stream));
  }
  /* every feeding stream has drained: the last publisher's host writes are on their way (a
   * peer write followed by a system-scope fence), give them a moment */
  for (unsigned int spin = 0; spin < 50000000u; ++spin)
  {
    if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) == seq) return PLG_OK;
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  plg_set_error("the result of a reduction never arrived (sequence %llu, flag %llu)", seq,
                __atomic_load_n(flag, __ATOMIC_ACQUIRE));
  return PLG_E_CUDA;
}

int plg_finish_result(plg_context * ctx, double * out0, double * out1)
{
  if (ctx->group && ctx->group->active)
    return PLG_OK; /* delivered by plg_group_collect */
  if (ctx->comm_seq)
  {
    const unsigned long long seq = ctx->comm_seq;
    ctx->comm_seq = 0;
    int crc = plg_comm_allreduce_publish(ctx, seq);
    if (crc) return crc;
  }
  if (ctx->deferred)
  {
    ctx->pending[0] = out0;
    ctx->pending[1] = out1;
    return PLG_OK;
  }
  if (ctx->copy_pending)
  {
    ctx->copy_pending = 0;
    PLG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  int rc = plg_wait_flag(ctx, ctx->result_seq, &ctx, 1);
  if (rc) return rc;
  if (out0) *out0 = ctx->result_host[0];
  if (out1) *out1 = ctx->result_host[1];
  return PLG_OK;
}

extern "C" int plg_collect(plg_context_t * ctx)
{
  PLG_CHECK_CTX(ctx);
  if (ctx->copy_pending)
  {
    ctx->copy_pending = 0;
    PLG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  int rc = plg_wait_flag(ctx, ctx->result_seq, &ctx, 1);
  if (rc) return rc;
  if (ctx->pending[0]) *ctx->pending[0] = ctx->result_host[0];
  if (ctx->pending[1]) *ctx->pending[1] = ctx->result_host[1];
  ctx->pending[0] = ctx->pending[1] = NULL;
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* device groups: one partition over several GPUs, scalar results combined on the devices  */
/* ------------------------------------------------------------------------------------ */
extern "C" int plg_group_create(plg_context_t * const * members, unsigned int n)
{
  if (!members || n < 2 || n > PLL_GPU_MAX_GROUP)
  {
    plg_set_error("plg_group_create: 2..%d contexts", PLL_GPU_MAX_GROUP);
    return PLG_E_INVALID;
  }
  plg_context * leader = members[0];
  for (unsigned int d = 0; d < n; ++d)
    if (!members[d] || members[d]->group)
    {
      plg_set_error("plg_group_create: context %u is NULL or already grouped", d);
      return PLG_E_INVALID;
    }
  for (unsigned int d = 0; d < n; ++d)
    if (members[d]->comm_buf)
    {
      /* ranks of a sharded run all-reduce every slice's sums; the host adds the slices */
      plg_set_error("plg_group_create: contexts of a cross-process communicator are not grouped");
      return PLG_E_UNSUPPORTED;
    }
  /* every member must be able to write the leader's memory */
  for (unsigned int d = 1; d < n; ++d)
  {
    if (members[d]->device == leader->device) continue;
    int can = 0;
    PLG_CUDA(cudaDeviceCanAccessPeer(&can, members[d]->device, leader->device));
    if (!can)
    {
      plg_set_error("plg_group_create: device %d cannot access device %d", members[d]->device, leader->device);
      return PLG_E_UNSUPPORTED;
    }
    PLG_CUDA(cudaSetDevice(members[d]->device));
    cudaError_t e = cudaDeviceEnablePeerAccess(leader->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
    {
      plg_set_error("plg_group_create: cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return PLG_E_CUDA;
    }
    cudaGetLastError();
  }
  plg_group * g = new plg_group();
  memset(g, 0, sizeof(*g));
  g->n = n;
  PLG_CUDA(cudaSetDevice(leader->device));
  if (cudaMalloc(&g->slots, 2 * PLL_GPU_MAX_GROUP * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&g->counter, 64) != cudaSuccess || cudaMemset(g->counter, 0, 64) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess)
  {
    cudaFree(g->slots);
    cudaFree(g->counter);
    delete g;
    cudaGetLastError();
    plg_set_error("plg_group_create: device allocation failed");
    return PLG_E_NOMEM;
  }
  for (unsigned int d = 0; d < n; ++d)
  {
    g->members[d] = members[d];
    members[d]->group = g;
    members[d]->group_rank = d;
  }
  return PLG_OK;
}

/* called by the leader's plg_destroy (members detach) */
static void plg_group_release(plg_context * ctx)
{
  plg_group * g = ctx->group;
  if (!g) return;
  for (unsigned int d = 0; d < g->n; ++d)
    if (g->members[d]) g->members[d]->group = NULL;
  cudaSetDevice(g->members[0] ? g->members[0]->device : ctx->device);
  cudaFree(g->slots);
  cudaFree(g->counter);
  delete g;
}

extern "C" int plg_group_begin(plg_context_t * leader)
{
  if (!leader || !leader->group || leader->group_rank != 0)
  {
    plg_set_error("plg_group_begin: not the leader of a device group");
    return PLG_E_INVALID;
  }
  plg_group * g = leader->group;
  g->seq = ++leader->result_seq;
  g->active = 1;
  return PLG_OK;
}

extern "C" int plg_group_collect(plg_context_t * leader, double * out0, double * out1)
{
  if (!leader || !leader->group || !leader->group->active)
  {
    plg_set_error("plg_group_collect: no group call in flight");
    return PLG_E_INVALID;
  }
  plg_group * g = leader->group;
  int rc = plg_wait_flag(leader, g->seq, g->members, g->n);
  g->active = 0;
  for (unsigned int d = 0; d < g->n && !rc; ++d)
    if (g->members[d]->copy_pending)
    {
      /* per-pattern values were requested too: their device-to-host copies follow the kernels */
      g->members[d]->copy_pending = 0;
      PLG_CUDA(cudaSetDevice(g->members[d]->device));
      PLG_CUDA(cudaStreamSynchronize(g->members[d]->stream));
    }
  if (rc) return rc;
  if (out0) *out0 = leader->result_host[0];
  if (out1) *out1 = leader->result_host[1];
  return PLG_OK;
}

extern "C" int plg_group_abort(plg_context_t * leader)
{
  if (!leader || !leader->group) return PLG_OK;
  plg_group * g = leader->group;
  g->active = 0;
  for (unsigned int d = 0; d < g->n; ++d)
  {
    cudaSetDevice(g->members[d]->device);
    cudaStreamSynchronize(g->members[d]->stream);
  }
  cudaSetDevice(leader->device);
  cudaMemset(g->counter, 0, 64);
  cudaDeviceSynchronize();
  cudaGetLastError();
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* staging ring                                                                          */
/* ------------------------------------------------------------------------------------ */
void * plg_stage(plg_context * ctx, const void * src, size_t bytes)
{
  const size_t need = round_up(bytes ? bytes : 1, 256);
  if (need > ctx->stage_size)
  {
    plg_set_error("plg_stage: %zu bytes exceed the %zu-byte staging ring", bytes,
                  ctx->stage_size);
    return NULL;
  }
  if (ctx->stage_off + need > ctx->stage_size)
  {
    /* wrap: every consumer of earlier slices must be finished before they are overwritten */
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    {
      plg_set_error("plg_stage: stream synchronize failed");
      return NULL;
    }
    ctx->stage_off = 0;
  }
  char * h = ctx->stage_host + ctx->stage_off;
  char * d = ctx->stage_dev + ctx->stage_off;
  memcpy(h, src, bytes);
  if (cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
  {
    plg_set_error("plg_stage: cudaMemcpyAsync failed: %s",
                  cudaGetErrorString(cudaGetLastError()));
    return NULL;
  }
  ctx->stage_off += need;
  ctx->stats.h2d_bytes += bytes;
  return d;
}

int plg_stage_reserve(plg_context * ctx, size_t bytes)
{
  if (bytes > ctx->stage_size)
  {
    plg_set_error("plg_stage_reserve: %zu bytes exceed the %zu-byte staging ring", bytes,
                  ctx->stage_size);
    return PLG_E_INVALID;
  }
  if (ctx->stage_off + bytes > ctx->stage_size)
  {
    PLG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stage_off = 0;
  }
  return PLG_OK;
}

int plg_ensure_tables(plg_context * ctx, size_t doubles)
{
  if (doubles <= ctx->tables_cap) return PLG_OK;
  PLG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->tables) cudaFree(ctx->tables);
  ctx->tables = NULL;
  ctx->tables_cap = 0;
  size_t cap = round_up(doubles + doubles / 2, 1024);
  PLG_CUDA(cudaMalloc(&ctx->tables, cap * sizeof(double)));
  ctx->tables_cap = cap;
  return PLG_OK;
}

int plg_ensure_partials(plg_context * ctx, size_t doubles)
{
  if (doubles <= ctx->partials_cap) return PLG_OK;
  PLG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->partials) cudaFree(ctx->partials);
  ctx->partials = NULL;
  ctx->partials_cap = 0;
  size_t cap = round_up(doubles * 2, 1024);
  PLG_CUDA(cudaMalloc(&ctx->partials, cap * sizeof(double)));
  ctx->partials_cap = cap;
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* uploads / downloads                                                                   */
/* ------------------------------------------------------------------------------------ */
static int h2d(plg_context * ctx, void * dst, const void * src, size_t bytes)
{
  /* pageable source: cudaMemcpyAsync stages it and returns once the source is consumed */
  PLG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->stats.h2d_bytes += bytes;
  return PLG_OK;
}
static int d2h(plg_context * ctx, void * dst, const void * src, size_t bytes)
{
  PLG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PLG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stats.d2h_bytes += bytes;
  return PLG_OK;
}

extern "C" int plg_set_tipchars(plg_context_t * ctx, unsigned int tip, const unsigned char * chars)
{
  PLG_CHECK_CTX(ctx);
  if (!ctx->pattern_tip || tip >= ctx->d.tips || !chars)
  {
    plg_set_error("plg_set_tipchars: invalid tip %u", tip);
    return PLG_E_INVALID;
  }
  return h2d(ctx, plg_tip_ptr(ctx, tip), chars, ctx->d.sites);
}

extern "C" int plg_get_tipchars(plg_context_t * ctx, unsigned int tip, unsigned char * chars)
{
  PLG_CHECK_CTX(ctx);
  if (!ctx->pattern_tip || tip >= ctx->d.tips || !chars)
  {
    plg_set_error("plg_get_tipchars: invalid tip %u", tip);
    return PLG_E_INVALID;
  }
  return d2h(ctx, chars, plg_tip_ptr(ctx, tip), ctx->d.sites);
}

extern "C" int plg_set_tipmap(plg_context_t * ctx, const unsigned int * tipmap,
                              unsigned int maxstates)
{
  PLG_CHECK_CTX(ctx);
  if (maxstates > PLL_ASCII_SIZE)
  {
    plg_set_error("plg_set_tipmap: maxstates %u > 256", maxstates);
    return PLG_E_INVALID;
  }
  /* captured operation lists carry the map by value: a changed map must not replay them */
  if (maxstates != ctx->maxstates || (tipmap && memcmp(ctx->tipmap, tipmap, maxstates * sizeof(unsigned int)) != 0))
    ctx->tipmap_epoch++;
  memset(ctx->tipmap, 0, sizeof(ctx->tipmap));
  if (tipmap) memcpy(ctx->tipmap, tipmap, maxstates * sizeof(unsigned int));
  ctx->maxstates = maxstates;
  ctx->log2_maxstates = maxstates > 1 ? (unsigned int)ceil(log2((double)maxstates)) : 0;
  return PLG_OK;
}

static int check_clv(plg_context * ctx, unsigned int idx, const char * who)
{
  if (idx < ctx->clv_first || idx >= ctx->d.tips + ctx->d.clv_buffers)
  {
    plg_set_error("%s: CLV index %u has no device storage", who, idx);
    return PLG_E_INVALID;
  }
  return PLG_OK;
}

extern "C" int plg_set_clv(plg_context_t * ctx, unsigned int idx, const double * clv)
{
  PLG_CHECK_CTX(ctx);
  int rc = check_clv(ctx, idx, "plg_set_clv");
  if (rc) return rc;
  return h2d(ctx, plg_clv_ptr(ctx, idx), clv, (size_t)ctx->d.sites * ctx->span * sizeof(double));
}

extern "C" int plg_get_clv(plg_context_t * ctx, unsigned int idx, double * clv)
{
  PLG_CHECK_CTX(ctx);
  int rc = check_clv(ctx, idx, "plg_get_clv");
  if (rc) return rc;
  return d2h(ctx, clv, plg_clv_ptr(ctx, idx), (size_t)ctx->d.sites * ctx->span * sizeof(double));
}

extern "C" int plg_set_scaler(plg_context_t * ctx, unsigned int idx, const unsigned int * s)
{
  PLG_CHECK_CTX(ctx);
  if (idx >= ctx->d.scale_buffers)
  {
    plg_set_error("plg_set_scaler: invalid index %u", idx);
    return PLG_E_INVALID;
  }
  return h2d(ctx, plg_scaler_ptr(ctx, (int)idx), s, ctx->scaler_len * sizeof(unsigned int));
}

extern "C" int plg_get_scaler(plg_context_t * ctx, unsigned int idx, unsigned int * s)
{
  PLG_CHECK_CTX(ctx);
  if (idx >= ctx->d.scale_buffers)
  {
    plg_set_error("plg_get_scaler: invalid index %u", idx);
    return PLG_E_INVALID;
  }
  return d2h(ctx, s, plg_scaler_ptr(ctx, (int)idx), ctx->scaler_len * sizeof(unsigned int));
}

extern "C" int plg_set_pattern_weights(plg_context_t * ctx, const unsigned int * w)
{
  PLG_CHECK_CTX(ctx);
  return h2d(ctx, ctx->weights, w, (size_t)ctx->d.sites * sizeof(unsigned int));
}

extern "C" int plg_set_pmatrix(plg_context_t * ctx, unsigned int idx, const double * p)
{
  PLG_CHECK_CTX(ctx);
  if (idx >= ctx->d.prob_matrices)
  {
    plg_set_error("plg_set_pmatrix: invalid index %u", idx);
    return PLG_E_INVALID;
  }
  return h2d(ctx, plg_pmat_ptr(ctx, idx), p, ctx->pmat_len * sizeof(double));
}

extern "C" int plg_get_pmatrix(plg_context_t * ctx, unsigned int idx, double * p)
{
  PLG_CHECK_CTX(ctx);
  if (idx >= ctx->d.prob_matrices)
  {
    plg_set_error("plg_get_pmatrix: invalid index %u", idx);
    return PLG_E_INVALID;
  }
  return d2h(ctx, p, plg_pmat_ptr(ctx, idx), ctx->pmat_len * sizeof(double));
}

/* ------------------------------------------------------------------------------------ */
/* invariant-site index                                                                  */
/* ------------------------------------------------------------------------------------ */
/* one thread per site: AND of the state masks of all tips, then "exactly one bit" -> index
 * (reference src/models.c:593-645).  use_map selects tipmap[code] (non-DNA alphabets). */
__global__ void k_invariant_tipchars(const unsigned char * __restrict__ tipchars,
                                     size_t tip_stride, unsigned int tips, unsigned int sites,
                                     unsigned int gap_state, int use_map,
                                     const TipmapArg tm, int * __restrict__ invariant)
{
  unsigned int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= sites) return;
  unsigned int state = gap_state;
  for (unsigned int t = 0; t < tips; ++t)
  {
    unsigned int c = tipchars[(size_t)t * tip_stride + n];
    state &= use_map ? tm.map[c] : c;
  }
  invariant[n] = (state == 0 || __popc(state) > 1) ? -1 : (__ffs(state) - 1);
}

/* tips stored as full CLVs: the mask of a tip is the set of states whose entry casts to a
 * non-zero unsigned (reference src/models.c:618-635 reads the first rate category). */
__global__ void k_invariant_tipclv(const double * __restrict__ clv, size_t clv_stride,
                                   size_t span, unsigned int states, unsigned int tips,
                                   unsigned int sites, unsigned int gap_state,
                                   int * __restrict__ invariant)
{
  unsigned int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= sites) return;
  unsigned int state = gap_state;
  for (unsigned int t = 0; t < tips; ++t)
  {
    const double * v = clv + (size_t)t * clv_stride + (size_t)n * span;
    unsigned int m = 0;
    for (unsigned int k = 0; k < states; ++k) m |= ((unsigned int)v[k]) << k;
    state &= m;
  }
  invariant[n] = (state == 0 || __popc(state) > 1) ? -1 : (__ffs(state) - 1);
}

extern "C" int plg_update_invariant(plg_context_t * ctx, int * invariant_out)
{
  PLG_CHECK_CTX(ctx);
  const unsigned int sites = ctx->d.sites;
  unsigned int gap = 0;
  for (unsigned int i = 0; i < ctx->d.states; ++i) gap = (gap << 1) | 1u;
  const int threads = 256;
  const unsigned int blocks = (sites + threads - 1) / threads;
  if (ctx->pattern_tip)
  {
    TipmapArg tm;
    memcpy(tm.map, ctx->tipmap, sizeof(tm.map));
    k_invariant_tipchars<<<blocks, threads, 0, ctx->stream>>>(
        ctx->tipchars, ctx->tip_stride, ctx->d.tips, sites, gap, ctx->d.states != 4, tm,
        ctx->invariant);
  }
  else
  {
    k_invariant_tipclv<<<blocks, threads, 0, ctx->stream>>>(
        ctx->clv, ctx->clv_stride, ctx->span, ctx->d.states, ctx->d.tips, sites, gap,
        ctx->invariant);
  }
  PLG_LAUNCH_CHECK(ctx);
  ctx->has_invariant = true;
  if (invariant_out) return d2h(ctx, invariant_out, ctx->invariant, (size_t)sites * sizeof(int));
  return PLG_OK;
}

/* The caller's own invariant[] (the `invariant` / `invar_indices` argument of the reference's
 * pll_core_* entry points, src/pll.h:943-1000): sites ints, -1 or the state index. */
extern "C" int plg_set_invariant(plg_context_t * ctx, const int * invariant)
{
  PLG_CHECK_CTX(ctx);
  if (!invariant)
  {
    plg_set_error("plg_set_invariant: null array");
    return PLG_E_INVALID;
  }
  int rc = h2d(ctx, ctx->invariant, invariant, (size_t)ctx->d.sites * sizeof(int));
  if (rc) return rc;
  PLG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->has_invariant = true;
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* measurement                                                                           */
/* ------------------------------------------------------------------------------------ */
extern "C" int plg_timer_start(plg_context_t * ctx)
{
  PLG_CHECK_CTX(ctx);
  PLG_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
  return PLG_OK;
}

extern "C" int plg_timer_stop(plg_context_t * ctx, float * elapsed_ms)
{
  PLG_CHECK_CTX(ctx);
  PLG_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
  PLG_CUDA(cudaEventSynchronize(ctx->ev_stop));
  float ms = 0.f;
  PLG_CUDA(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
  if (elapsed_ms) *elapsed_ms = ms;
  return PLG_OK;
}

extern "C" int plg_get_stats(plg_context_t * ctx, plg_stats_t * out)
{
  if (!ctx || !out) return PLG_E_INVALID;
  *out = ctx->stats;
  return PLG_OK;
}

extern "C" int plg_set_profiling(plg_context_t * ctx, int enable)
{
  if (!ctx) return PLG_E_INVALID;
  ctx->profiling = enable ? 1 : 0;
  return PLG_OK;
}

extern "C" int plg_reset_stats(plg_context_t * ctx)
{
  if (!ctx) return PLG_E_INVALID;
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  return PLG_OK;
}

__global__ void k_flush_l2(uint4 * __restrict__ buf, size_t n, unsigned int tag)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) buf[i] = make_uint4(tag, tag, tag, tag);
}

extern "C" int plg_flush_l2(plg_context_t * ctx)
{
  PLG_CHECK_CTX(ctx);
  if (!ctx->flush_buf)
  {
    ctx->flush_bytes = (size_t)256 << 20; /* 2x the 126 MB L2 */
    PLG_CUDA(cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
  }
  static unsigned int tag = 0;
  k_flush_l2<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint4 *)ctx->flush_buf,
                                                         ctx->flush_bytes / sizeof(uint4), ++tag);
  PLG_CUDA(cudaGetLastError()); /* a measurement helper: not counted as a hot-path launch */
  return PLG_OK;
}

extern "C" int plg_mem_info(plg_context_t * ctx, size_t * free_bytes, size_t * total_bytes)
{
  PLG_CHECK_CTX(ctx);
  size_t f = 0, t = 0;
  PLG_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* site ranges: what the ascertainment-bias epilogue of the host layer reads              */
/* ------------------------------------------------------------------------------------ */
extern "C" int plg_set_active_sites(plg_context_t * ctx, unsigned int sites)
{
  PLG_CHECK_CTX(ctx);
  /* 0 is legal: a slice of a multi-device partition that holds per-state sites only */
  if (sites > ctx->d.sites)
  {
    plg_set_error("plg_set_active_sites: %u out of range (0..%u)", sites, ctx->d.sites);
    return PLG_E_INVALID;
  }
  ctx->active_sites = sites;
  return PLG_OK;
}

extern "C" int plg_get_clv_sites(plg_context_t * ctx, unsigned int clv_index, unsigned int first_site,
                                 unsigned int count, double * out)
{
  PLG_CHECK_CTX(ctx);
  if (clv_index < ctx->clv_first || clv_index >= ctx->d.tips + ctx->d.clv_buffers ||
      (size_t)first_site + count > ctx->d.sites || !out)
  {
    plg_set_error("plg_get_clv_sites: index out of range");
    return PLG_E_INVALID;
  }
  return d2h(ctx, out, plg_clv_ptr(ctx, clv_index) + (size_t)first_site * ctx->span,
             (size_t)count * ctx->span * sizeof(double));
}

extern "C" int plg_get_scaler_sites(plg_context_t * ctx, unsigned int scaler_index, unsigned int first_site,
                                    unsigned int count, unsigned int * out)
{
  PLG_CHECK_CTX(ctx);
  const size_t per = ctx->rate_scalers ? ctx->d.rate_cats : 1;
  if (scaler_index >= ctx->d.scale_buffers || (size_t)first_site + count > ctx->d.sites || !out)
  {
    plg_set_error("plg_get_scaler_sites: index out of range");
    return PLG_E_INVALID;
  }
  return d2h(ctx, out, plg_scaler_ptr(ctx, (int)scaler_index) + (size_t)first_site * per,
             (size_t)count * per * sizeof(unsigned int));
}
