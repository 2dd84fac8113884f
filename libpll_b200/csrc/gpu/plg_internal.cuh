/*
 * plg_internal.cuh - shared definitions of the device layer (not part of the public ABI).
 *
 * Data layout in HBM (all arrays 256-byte aligned, one cudaMalloc slab per kind):
 *   clv      [slot][site][rate][state_padded] double   - same order as the reference CLV
 *                                                        (reference src/pll.c:522-542)
 *   scalers  [slot][site] u32, or [slot][site][rate] with PLL_ATTRIB_RATE_SCALERS
 *   tipchars [tip][site] u8                              (only with PLL_ATTRIB_PATTERN_TIP)
 *   pmatrix  [matrix][rate][row][col_padded] double      (reference src/pll.c:555-573)
 *   weights  [site] u32, invariant [site] i32
 */
#ifndef PLG_INTERNAL_CUH_
#define PLG_INTERNAL_CUH_

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "pll_gpu.h"

/* ------------------------------------------------------------------------------------ */
/* error plumbing                                                                        */
/* ------------------------------------------------------------------------------------ */
void plg_set_error(const char * fmt, ...);

#define PLG_CUDA(call)                                                                 \
  do {                                                                                 \
    cudaError_t err__ = (call);                                                        \
    if (err__ != cudaSuccess) {                                                        \
      plg_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,                 \
                    cudaGetErrorString(err__));                                        \
      return (err__ == cudaErrorMemoryAllocation) ? PLG_E_NOMEM : PLG_E_CUDA;          \
    }                                                                                  \
  } while (0)

#define PLG_LAUNCH_CHECK(ctx)                                                          \
  do {                                                                                 \
    cudaError_t err__ = cudaGetLastError();                                            \
    if (err__ != cudaSuccess) {                                                        \
      plg_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,             \
                    cudaGetErrorString(err__));                                        \
      return PLG_E_CUDA;                                                               \
    }                                                                                  \
    (ctx)->stats.kernel_launches++;                                                    \
  } while (0)

/* ------------------------------------------------------------------------------------ */
/* context                                                                               */
/* ------------------------------------------------------------------------------------ */
/* a captured, instantiated CUDA graph of one whole operation list (plg_partials.cu) */
struct plg_graph_entry
{
  cudaGraphExec_t exec;
  std::vector<unsigned char> key_bytes; /* the pll_operation_t list it was built from */
  void * dev_tables;                    /* op / job descriptors, stable for the graph's life */
  unsigned long long kernels;
  unsigned long long levels;
  unsigned long long algorithmic_bytes;
  unsigned long long compulsory_bytes;
  unsigned long long last_used; /* LRU clock of the graph cache */
};

#define PLG_CHECK_CTX(ctx)                                                             \
  do {                                                                                 \
    if (!(ctx)) { plg_set_error("%s: NULL context", __func__); return PLG_E_INVALID; } \
    PLG_CUDA(cudaSetDevice((ctx)->device));                                            \
  } while (0)

/* Where a reduction kernel delivers its 1 or 2 doubles.  A lone context: its own mapped pinned
 * host words ([0], [1] = values, [4] = sequence flag the host polls).  A context of a device
 * group (pll_gpu_set_devices: one partition over several GPUs): a slot in the LEADER device's
 * memory, written over NVLink peer access; the member whose last block arrives last adds the
 * slots in member order and hands the total to the leader's host words - the one cross-device
 * exchange of the path (reference src/core_likelihood_avx.c:1259 `logl +=`,
 * src/core_derivatives_avx2.c:756-765), done on the devices with a single host wake-up. */
struct PlgSink
{
  double * result;
  unsigned long long seq;
  double * group_slots;          /* [group_size][2] on the leader; NULL outside a group call */
  unsigned int * group_counter;  /* arrivals of the current call, on the leader */
  double * group_result;         /* the leader's mapped host words */
  unsigned int group_size, group_rank;
};

struct plg_group
{
  unsigned int n;
  plg_context * members[PLL_GPU_MAX_GROUP];
  double * slots;          /* leader device memory */
  unsigned int * counter;
  unsigned long long seq;  /* of the call in flight */
  int active;              /* between plg_group_begin and plg_group_collect / _abort */
};

struct plg_context
{
  plg_dims_t d;
  int device;
  int sm_count;
  cudaStream_t stream;

  bool pattern_tip;
  bool rate_scalers;
  unsigned int active_sites; /* leading sites the lnL / derivative reductions cover */
  double * lnl_scratch;      /* one CLV-sized scratch (20-state edge lnL), lazily allocated */
  int use_fused;             /* DNA: whole operations list in one kernel (PLL_GPU_FUSED, default 1) */
  int use_fused_aa;          /* 20 states: the same on the tensor cores (plg_walk_aa.cu; PLL_GPU_FUSED_AA): 0 never,
                                1 always, 2 (default) for lists that recycle CLV / scaler slots - there most stores
                                are dead and the walk is 16-22 % faster than the level-by-level kernels, which
                                win by 6 % on lists without recycling */
  unsigned int fused_slots;  /* tiles a warp keeps in shared memory (PLL_GPU_FUSED_SLOTS, default 3) */
  unsigned char * fused_records; /* packed operation records of the non-graph path */
  size_t fused_records_cap;

  size_t span;          /* rate_cats * states_padded (doubles per site of a CLV)          */
  size_t clv_stride;    /* doubles between consecutive CLV slots                          */
  size_t scaler_len;    /* u32 entries per scale buffer (sites or sites*rate_cats)        */
  size_t scaler_stride; /* u32 entries between consecutive scale buffers                  */
  size_t tip_stride;    /* bytes between consecutive tip-character rows                   */
  size_t pmat_len;      /* doubles per P-matrix set: rate_cats*states*states_padded       */
  unsigned int clv_first; /* first CLV index that has device storage                      */

  double * clv;
  unsigned int * scalers;
  unsigned char * tipchars;
  double * pmatrix;
  unsigned int * weights;
  int * invariant;
  bool has_invariant;

  unsigned int maxstates;
  unsigned int log2_maxstates;
  unsigned int * root_counts;      /* per-pattern counts of plg_root_loglikelihood_counts (lazy) */
  unsigned long long tipmap_epoch; /* bumped when the map's content changes: part of the graph-cache key */
  unsigned int tipmap[PLL_ASCII_SIZE];

  /* pinned-host / device staging ring for small per-call constants and op tables */
  char * stage_host;
  char * stage_dev;
  size_t stage_size;
  size_t stage_off;

  /* scratch: per-op tip lookup tables, reduction partials, results */
  double * tables;
  size_t tables_cap; /* doubles */
  double * partials;
  size_t partials_cap; /* doubles */
  unsigned int * counter; /* "last block done" ticket */
  double * result_dev;    /* device alias of result_host (mapped) */
  double * result_host;   /* pinned, 4 doubles */
  /* plg_set_deferred: value-returning calls enqueue only and leave their outputs pending
   * until plg_collect (lets a caller overlap the same call on several devices) */
  int deferred;
  double * pending[2];
  unsigned long long result_seq; /* sequence number of the last value-returning call (flag in result_host[4]) */
  int copy_pending;              /* a D2H copy was enqueued behind the reduction: wait for the stream, not the flag */
  plg_group * group;             /* device group this context belongs to (NULL: none) */
  unsigned int group_rank;
  double * comm_buf;             /* device scratch of the cross-rank all-reduce (plg_comm.cu); NULL: not sharded */
  unsigned long long comm_seq;   /* sequence number of a reduction waiting for its all-reduce (0: none) */
  double * persite_dev;   /* sites doubles, allocated on first use */
  double * lnl_table;     /* pi-weighted tip lookup of the edge-lnL tip-inner kernels */
  size_t lnl_table_cap;   /* doubles */

  /* device-resident sumtables keyed by the caller's host pointer */
  std::unordered_map<const void *, double *> * sumtables;
  std::unordered_map<const void *, unsigned long long> * sumtable_used; /* key -> LRU clock */
  unsigned long long sumtable_clock;

  /* cached CUDA graphs of whole operation lists, keyed by a hash of the list */
  std::unordered_map<uint64_t, plg_graph_entry *> * graphs;
  std::unordered_map<uint64_t, unsigned int> * seen_lists; /* hash -> times seen, lists not (yet) captured */
  int use_graphs;
  unsigned int graph_cap;        /* cached graphs kept per context (PLL_GPU_GRAPH_CACHE, default 64); LRU eviction */
  unsigned long long graph_clock;
  char * list_buf;               /* device home of operation lists too long for the staging ring */
  size_t list_buf_cap;
  int aa_exact; /* 20 states: 1 = bit-exact vector-pipe kernels, 0 = DMMA tensor-core kernels */

  /* L2 residency of a sumtable between the derivative passes of a Newton loop: bytes of L2 set
   * aside for persisting lines (0: off, PLL_GPU_L2_PERSIST=0) and the largest access-policy window */
  size_t l2_persist_bytes;
  size_t l2_window_max;
  int l2_pinned;       /* a derivative pass may have left persisting lines: demote them before other work */

  /* L2 flush buffer (allocated on first plg_flush_l2) */
  char * flush_buf;
  size_t flush_bytes;

  cudaEvent_t ev_start, ev_stop;
  int profiling;
  std::vector<cudaEvent_t> * prof_events;
  plg_stats_t stats;
};

/* cross-process sharding (plg_comm.cu) */
bool plg_comm_covers(int device);
int plg_comm_allreduce_publish(plg_context * ctx, unsigned long long seq);

/* The sink of the next reduction kernel of this context (advances the sequence number). */
PlgSink plg_make_sink(plg_context * ctx);
/* Waits until the mapped host words of `ctx` carry sequence number `seq`: a short spin on the
 * flag (a value-returning call on an idle stream costs no wake-up through the driver), then a
 * blocking stream synchronise.  `watch` are the contexts whose streams feed the result. */
int plg_wait_flag(plg_context * ctx, unsigned long long seq, plg_context * const * watch, unsigned int n_watch);

/* Delivers the 1 or 2 doubles a reduction kernel left in result_host: waits for them and copies
 * them out - or, in deferred mode, remembers where they go until plg_collect. */
int plg_finish_result(plg_context * ctx, double * out0, double * out1);


#define PLG_MAX_DEVICES 64

/* ---- descriptors shared by the CLV-update kernels (plg_partials.cu, plg_generic.cu) ---- */
enum { PLG_KIND_TT = 0, PLG_KIND_TI = 1, PLG_KIND_II = 2 };

struct DevOp
{
  double * parent;
  const double * left;        /* ii: left child CLV                                   */
  const double * right;       /* ii: right child CLV; ti: the inner child's CLV       */
  const unsigned char * ltip; /* tt: left tip chars;  ti: the tip child's chars       */
  const unsigned char * rtip; /* tt: right tip chars                                  */
  const double * lmat;        /* ii: left P-matrix;  ti/tt: lookup table of ltip      */
  const double * rmat;        /* ii/ti: P-matrix of `right`; tt: lookup table of rtip */
  unsigned int * pscale;
  const unsigned int * lscale;
  const unsigned int * rscale;
};

struct TableJob
{
  const double * pmat;
  double * out;
};

struct TipmapArg
{
  unsigned int map[PLL_ASCII_SIZE];
};

/* an operation of the fused traversal (plg_traverse.cu): where the children's tiles live
 * (shared-memory slot of the warp, or -1 = HBM) and where the result tile is kept */
struct FusedOp /* 128 bytes: it travels by TMA bulk copy */
{
  DevOp op;
  int kind;
  int scale_mode;
  int lslot, rslot, pslot;
  unsigned int lbytes, rbytes; /* bytes of the packed left / right block this operation reads */
  int pad;                     /* bit 0: write the result through to HBM (0: dead store, see build_plan);
                                  bit 1: its scaler is read back from HBM later in the list */
  const double * lsrc;         /* P-matrix set of the left / right child (tip or inner) */
  const double * rsrc;
};
/* bytes of one packed operation record: descriptor + left block + right block, each block a
 * bank-conflict-padded P-matrix set ([rate] x 18 doubles) or tip table ([16 codes] x (4R+2)) */
static inline size_t plg_fused_block_bytes(unsigned int R) { return (size_t)16 * (R * 4 + 2) * sizeof(double); }
static inline size_t plg_fused_record_bytes(unsigned int R) { return 128 + 2 * plg_fused_block_bytes(R); }
int plg_launch_fused(plg_context * ctx, const FusedOp * dev_ops, unsigned char * dev_records, unsigned int n_ops,
                     unsigned int nslot);
/* the same for 20 states on the FP64 tensor cores (plg_walk_aa.cu, PLL_GPU_FUSED_AA=1): FusedOp::lbytes = bytes of the
 * operation's record the ring needs, FusedOp::rbytes = first row of its tip table (rows of
 * plg_walk_aa_row_bytes) in the table area behind the records */
#define PLG_WALK_AA_SLOTS 2
int plg_launch_walk_aa(plg_context * ctx, const FusedOp * dev_ops, unsigned char * dev_records, unsigned char * dev_tables,
                       unsigned int n_ops);
size_t plg_walk_aa_record_bytes(unsigned int rate_cats);
size_t plg_walk_aa_row_bytes(unsigned int rate_cats);
bool plg_walk_aa_supported(unsigned int rate_cats, unsigned int ncodes);

/* one operation-shaped launch on the specialised kernels (plg_partials.cu) */
int plg_launch_single_op(plg_context * ctx, int kind, const DevOp & op);

/* the generic (any state count / any number of rate categories) device path, plg_generic.cu */
static inline bool plg_is_pow2(unsigned int v) { return v && !(v & (v - 1)); }
static inline bool plg_fast_path(const plg_context * ctx);
int plg_gen_tables(plg_context * ctx, const TableJob * dev_jobs, unsigned int njobs);
int plg_gen_partials(plg_context * ctx, int kind, int scale_mode, const DevOp * dev_ops, unsigned int count);
int plg_gen_pmatrix(plg_context * ctx, const unsigned int * d_idx, const double * d_bl, unsigned int count,
                    const double * evals, const double * evecs, const double * ievecs, const double * rates,
                    const double * pinv);
struct GenLnl
{
  const double * clvp;
  const double * clvc;       /* ii */
  const unsigned char * tip; /* ti */
  const double * pmat;       /* edge: P-matrix set */
  const unsigned int * pscale;
  const unsigned int * cscale;
  int root;                  /* 1: root lnL (no P-matrix, single CLV) */
};
int plg_gen_loglikelihood(plg_context * ctx, const GenLnl & g, const double * freqs, const double * rate_weights,
                          const double * prop_invar, double * persite_lnl, double * logl_out);
int plg_gen_sumtable(plg_context * ctx, const double * clvp, const double * clvc, const unsigned char * tip,
                     const unsigned int * pscale, const unsigned int * cscale, const double * dev_evecs,
                     const double * dev_left, double * sumtable);
int plg_gen_derivatives(plg_context * ctx, const double * sumtable, const double * diagptable,
                        const double * rate_weights, const double * prop_invar, const double * freqs,
                        double * d_f, double * dd_f);

/* Copies `bytes` of host data into the staging ring and enqueues the H2D copy on the
 * context's stream; returns the device address (256-byte aligned) or NULL on failure. */
void * plg_stage(plg_context * ctx, const void * src, size_t bytes);
/* Guarantees that the next `bytes` (callers add 256 per item for alignment) of plg_stage calls
 * come from one contiguous, not-yet-recycled region of the ring.  Returns 0 on success. */
int plg_stage_reserve(plg_context * ctx, size_t bytes);

/* Hands the L2 lines a Newton loop pinned (the sumtable) back to ordinary replacement. */
static inline void plg_release_l2(plg_context * ctx)
{
  if (ctx->l2_pinned)
  {
    cudaCtxResetPersistingL2Cache();
    ctx->l2_pinned = 0;
  }
}

int plg_ensure_tables(plg_context * ctx, size_t doubles);
int plg_ensure_partials(plg_context * ctx, size_t doubles);

static inline double * plg_clv_ptr(const plg_context * ctx, unsigned int idx)
{
  return ctx->clv + (size_t)(idx - ctx->clv_first) * ctx->clv_stride;
}
static inline unsigned int * plg_scaler_ptr(const plg_context * ctx, int idx)
{
  return (idx == PLL_SCALE_BUFFER_NONE) ? NULL
                                        : ctx->scalers + (size_t)idx * ctx->scaler_stride;
}
static inline unsigned char * plg_tip_ptr(const plg_context * ctx, unsigned int idx)
{
  return ctx->tipchars + (size_t)idx * ctx->tip_stride;
}
static inline double * plg_pmat_ptr(const plg_context * ctx, unsigned int idx)
{
  return ctx->pmatrix + (size_t)idx * ctx->pmat_len;
}
static inline bool plg_is_tip(const plg_context * ctx, unsigned int clv_index)
{
  return ctx->pattern_tip && clv_index < ctx->d.tips;
}
/* specialised kernels exist for 4 and 20 states with 1/2/4/8/16 rate categories; everything
 * else runs on the generic kernels of plg_generic.cu */
static inline bool plg_fast_path(const plg_context * ctx)
{
  return (ctx->d.states == 4 || ctx->d.states == 20) && ctx->d.rate_cats <= 16 &&
         (ctx->d.rate_cats & (ctx->d.rate_cats - 1)) == 0;
}

/* ------------------------------------------------------------------------------------ */
/* device helpers shared by the kernels                                                  */
/* ------------------------------------------------------------------------------------ */
#ifdef __CUDACC__

/* four consecutive states of one (site, rate): one 256-bit global access on sm_100a */
struct __align__(32) d4
{
  double x, y, z, w;
};

/* streaming (read-once / write-once) 256-bit accesses that do not allocate in L1 */
__device__ __forceinline__ d4 ld_stream(const double * p)
{
  d4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
               : "l"(p));
  return r;
}
/* The same access without .nc: for data that THIS launch may have written (a fused traversal
 * reading back a tile it stored earlier; PTX guarantees .nc only for memory that is read-only
 * for the kernel's lifetime). */
__device__ __forceinline__ d4 ld_stream_coherent(const double * p)
{
  d4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ double2 ld_stream_coherent2(const double * p)
{
  double2 r;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ double ld_stream_coherent1(const double * p)
{
  double r;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ unsigned int ld_coherent_u32(const unsigned int * p)
{
  unsigned int r;
  asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_stream(double * p, d4 v)
{
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};"
               :
               : "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w)
               : "memory");
}

/* The reference's 4-lane horizontal sum (unpackhi/unpacklo, add, permute2f128/blend, add;
 * e.g. reference src/core_partials_avx.c:460-471) evaluates (a0+a1)+(a2+a3). */
__device__ __forceinline__ double hsum4(double a0, double a1, double a2, double a3)
{
  return __dadd_rn(__dadd_rn(a0, a1), __dadd_rn(a2, a3));
}

/* unfused row * vector, the DNA kernels' building block (mul, then hsum4) */
__device__ __forceinline__ double dot4_unfused(double m0, double m1, double m2, double m3,
                                               const d4 & c)
{
  return hsum4(__dmul_rn(m0, c.x), __dmul_rn(m1, c.y), __dmul_rn(m2, c.z), __dmul_rn(m3, c.w));
}

/* called by ONE thread of the block that holds the final sums of a reduction */
__device__ __forceinline__ void plg_publish(const PlgSink & s, double r0, double r1)
{
  if (!s.group_slots)
  {
    s.result[0] = r0;
    s.result[1] = r1;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(s.result + 4) = s.seq;
    return;
  }
  volatile double * slot = s.group_slots + 2 * s.group_rank;
  slot[0] = r0;
  slot[1] = r1;
  __threadfence_system();
  if (atomicAdd_system(s.group_counter, 1u) == s.group_size - 1)
  {
    /* every member has published: fixed-order sum (member 0 first, as a host loop would) */
    __threadfence_system();
    const volatile double * all = s.group_slots;
    double t0 = all[0], t1 = all[1];
    for (unsigned int d = 1; d < s.group_size; ++d)
    {
      t0 = __dadd_rn(t0, all[2 * d]);
      t1 = __dadd_rn(t1, all[2 * d + 1]);
    }
    *s.group_counter = 0u;
    s.group_result[0] = t0;
    s.group_result[1] = t1;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(s.group_result + 4) = s.seq;
  }
}

#define PLG_SCALE_THRESHOLD 0x1p-256
#define PLG_SCALE_FACTOR 0x1p+256

/* deterministic block-wide sum: fixed shuffle tree inside each warp, then warp 0 adds the
 * per-warp values in warp order.  Result valid in thread 0. */
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double * smem /* THREADS/32 doubles */)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, off));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double total = 0.0;
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) total = __dadd_rn(total, smem[w]);
  }
  __syncthreads();
  return total;
}

#endif /* __CUDACC__ */

#endif /* PLG_INTERNAL_CUH_ */
