/*
 * pll_gpu.h - the drop-in boundary: a C-ABI over the sm_100a kernels.
 *
 * Two groups of entry points:
 *
 *  (1) plg_*      the device layer.  Plain pointers and sizes, no C++/torch types.  Each
 *                 function replaces one rung of the reference's `pll_core_*` dispatch ladders
 *                 (the reference selects SSE/AVX/AVX2 kernels by `attrib` inside every
 *                 pll_core_* function, e.g. reference src/core_partials.c:534-587); the
 *                 reference interface each one stands in for is cited on the declaration.
 *                 This is what a maintainer of the reference would bind.
 *
 *  (2) pll_gpu_*  optional extensions on a pll_partition_t created with PLL_ATTRIB_ARCH_GPU:
 *                 device selection, host-mirror downloads, stream timing, launch statistics.
 *
 * State model: a plg_context_t owns, in HBM and for its whole lifetime, every CLV slot, scale
 * buffer, tip-character row, P-matrix, the pattern weights and the invariant-site index; all
 * work is enqueued on one CUDA stream per context.  void/setter calls may return after
 * enqueue; value-returning calls synchronise.  All functions return PLG_OK (0) or a PLG_E_*
 * code; plg_last_error() gives the thread-local message.
 */
#ifndef PLL_B200_PLL_GPU_H_
#define PLL_B200_PLL_GPU_H_

#include <stddef.h>
#include "pll.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PLG_OK 0
#define PLG_E_NODEVICE 1    /* no CUDA device, or not compute capability 10.x */
#define PLG_E_NOMEM 2       /* cudaMalloc / cudaHostAlloc failed */
#define PLG_E_CUDA 3        /* any other CUDA runtime / launch failure */
#define PLG_E_INVALID 4     /* bad argument */
#define PLG_E_UNSUPPORTED 5 /* configuration not implemented on the device */

typedef struct plg_context plg_context_t;

/* Shape of a partition as the device layer sees it (mirrors the dimension fields of
 * pll_partition_t, reference src/pll.h:204-217).  `sites` already includes the extra
 * per-state columns when ascertainment-bias storage is requested. */
typedef struct plg_dims
{
  unsigned int tips;
  unsigned int clv_buffers;
  unsigned int states;
  unsigned int states_padded;
  unsigned int sites;
  unsigned int rate_cats;
  unsigned int rate_matrices;
  unsigned int prob_matrices;
  unsigned int scale_buffers;
  unsigned int attributes; /* PLL_ATTRIB_PATTERN_TIP, PLL_ATTRIB_RATE_SCALERS */
} plg_dims_t;

/* Launch / traffic counters since context creation (or the last plg_reset_stats). */
typedef struct plg_stats
{
  unsigned long long kernel_launches;   /* kernels of THIS library launched */
  unsigned long long graph_launches;    /* cudaGraphLaunch calls (each replays many kernels) */
  unsigned long long h2d_bytes;
  unsigned long long d2h_bytes;
  unsigned long long partial_ops;       /* pll_operation_t entries executed */
  unsigned long long partial_levels;    /* dependency levels they were batched into */
  unsigned long long algorithmic_bytes; /* SURVEY.md 8(d) bytes of the partial kernels */
  /* filled only while plg_set_profiling(ctx, 1) is active: per kind of CLV update
   * (0 = tip-tip, 1 = tip-inner, 2 = inner-inner) the device time between CUDA events
   * recorded around each launch on the context's stream, the algorithmic bytes those
   * launches moved and their number */
  unsigned long long kind_ns[3];
  unsigned long long kind_bytes[3];
  unsigned long long kind_launches[3];
  /* bytes the executed CLV-update plans HAD to move over the HBM interface: equal to
   * algorithmic_bytes on the level-by-level kernels; for the single-kernel traversals every
   * observable parent CLV / scaler written once + tip characters + tile-cache misses read back
   * (computed from the plan, not measured) */
  unsigned long long compulsory_bytes;
  unsigned long long graph_evictions;   /* cached operation-list graphs dropped (LRU) */
  unsigned long long collectives;       /* cross-rank all-reduces of scalar results (plg_comm_*) */
} plg_stats_t;

PLL_EXPORT const char * plg_last_error(void);

/* Number of usable sm_100 devices (0 if none); never fails. */
PLL_EXPORT int plg_device_count(void);

/* replaces: the allocation half of pll_partition_create (reference src/pll.c:509-815).
 * device < 0 selects the current CUDA device. */
PLL_EXPORT int plg_create(const plg_dims_t * dims, int device, plg_context_t ** out);
/* replaces: dealloc_partition_data (reference src/pll.c:31-111) */
PLL_EXPORT void plg_destroy(plg_context_t * ctx);

PLL_EXPORT int plg_synchronize(plg_context_t * ctx);

/* Deferred results: while enabled, plg_edge_loglikelihood / plg_root_loglikelihood /
 * plg_likelihood_derivatives only enqueue their kernels and return; plg_collect waits for the
 * stream and stores the pending results through the output pointers of the last such call (which
 * must stay valid until then; persite_lnl likewise).  Used by the multi-device host layer to run
 * the same reduction on all devices at once (libpll_b200/csrc/host/pll_devices.c). */
PLL_EXPORT int plg_set_deferred(plg_context_t * ctx, int enable);
PLL_EXPORT int plg_collect(plg_context_t * ctx);

/* Device groups: the contexts of ONE partition cut into pattern slices over several GPUs
 * (members[0] is the leader).  Between plg_group_begin and plg_group_collect every member's
 * plg_edge_loglikelihood / plg_root_loglikelihood / plg_likelihood_derivatives call only
 * enqueues; the last block of each device publishes its partial sums into a slot of the
 * leader's memory (NVLink peer access), the device that arrives last adds the slots in member
 * order and writes the total to the leader's mapped host words, and plg_group_collect waits for
 * that ONE flag: the scalar all-reduce of reference src/core_likelihood_avx.c:1259 (`logl +=`)
 * and src/core_derivatives_avx2.c:756-765 across devices, with a single host wake-up.
 * plg_group_create fails with PLG_E_UNSUPPORTED when a member cannot reach the leader's memory
 * (the host layer then adds the per-device results itself); plg_group_abort clears a call that
 * failed half way.  The group dissolves when any member is destroyed. */
#define PLL_GPU_MAX_GROUP 16
PLL_EXPORT int plg_group_create(plg_context_t * const * members, unsigned int n);
PLL_EXPORT int plg_group_begin(plg_context_t * leader);
PLL_EXPORT int plg_group_collect(plg_context_t * leader, double * out0, double * out1);
PLL_EXPORT int plg_group_abort(plg_context_t * leader);

/* Cross-process site sharding (one rank per GPU, each owning a pattern slice of every CLV):
 * after plg_comm_init the scalar results of every context created on `device` - edge / root
 * log-likelihood, d_f / dd_f - are summed over the ranks by an ncclAllReduce of 1-2 doubles on
 * the context's stream before they reach the host (the `logl +=` / `d_f +=` coupling of reference
 * src/core_likelihood_avx.c:1259, src/core_derivatives_avx2.c:756-765).  NCCL is dlopen'ed
 * (libnccl.so.2, or $PLL_GPU_NCCL_LIB).  id = 128 bytes from plg_comm_unique_id on rank 0, carried
 * to the other ranks by the caller.  Per-pattern outputs (persite_lnl) stay local. */
#define PLL_GPU_COMM_ID_BYTES 128
PLL_EXPORT int plg_comm_unique_id(unsigned char * id);
PLL_EXPORT int plg_comm_init(const unsigned char * id, int nranks, int rank, int device);
PLL_EXPORT int plg_comm_finalize(void);
PLL_EXPORT int plg_comm_size(void);

/* ---- uploads / downloads of resident state ---------------------------------------- */

/* replaces: the stores of set_tipchars_4x4 / set_tipchars (reference src/pll.c:825-903):
 * `chars` holds ctx->sites bytes (DNA: 4-bit masks; other alphabets: charmap codes). */
PLL_EXPORT int plg_set_tipchars(plg_context_t * ctx, unsigned int tip_index,
                                const unsigned char * chars);
PLL_EXPORT int plg_get_tipchars(plg_context_t * ctx, unsigned int tip_index,
                                unsigned char * chars);
/* Synthetic DNA tip row generated on the device (benchmarks; SURVEY.md 8d "tips generated
 * directly on device"): the state mask of pattern first_site + i of tip `tip_index` is a pure
 * function of (seed, tip_index, first_site + i); libpll_b200/synthetic.py: hash_tip_sequence is
 * the host restatement.  4-state pattern-tip partitions only. */
PLL_EXPORT int plg_generate_tipchars(plg_context_t * ctx, unsigned int tip_index, unsigned long long seed,
                                     unsigned long long first_site);
/* replaces: partition->tipmap / maxstates maintenance (reference src/pll.c:136-397) */
PLL_EXPORT int plg_set_tipmap(plg_context_t * ctx, const unsigned int * tipmap,
                              unsigned int maxstates);
/* replaces: set_tipclv / pll_set_tip_clv stores (reference src/pll.c:905-1045);
 * `clv` is a full host CLV: sites * rate_cats * states_padded doubles. */
PLL_EXPORT int plg_set_clv(plg_context_t * ctx, unsigned int clv_index, const double * clv);
PLL_EXPORT int plg_get_clv(plg_context_t * ctx, unsigned int clv_index, double * clv);
PLL_EXPORT int plg_set_scaler(plg_context_t * ctx, unsigned int scaler_index,
                              const unsigned int * scaler);
PLL_EXPORT int plg_get_scaler(plg_context_t * ctx, unsigned int scaler_index,
                              unsigned int * scaler);
/* replaces: pll_set_pattern_weights memcpy (reference src/pll.c:1047-1059) */
PLL_EXPORT int plg_set_pattern_weights(plg_context_t * ctx, const unsigned int * weights);
/* replaces: the tip scan of pll_update_invariant_sites (reference src/models.c:558-647):
 * computes invariant[] on the device from the resident tips and copies it to `invariant_out`
 * (sites ints; may be NULL to keep it device-only). */
PLL_EXPORT int plg_update_invariant(plg_context_t * ctx, int * invariant_out);
/* The caller's own invariant[] array (sites ints: -1 or the state index) instead of the tip scan;
 * the `invariant` / `invar_indices` argument of the reference's pll_core_* calls
 * (src/pll.h:943-1000). */
PLL_EXPORT int plg_set_invariant(plg_context_t * ctx, const int * invariant);
PLL_EXPORT int plg_set_pmatrix(plg_context_t * ctx, unsigned int matrix_index,
                               const double * pmatrix);
PLL_EXPORT int plg_get_pmatrix(plg_context_t * ctx, unsigned int matrix_index,
                               double * pmatrix);

/* Ascertainment-bias storage (reference src/pll.c:492-495: `sites + states` allocated sites):
 * CLV updates and sumtables always cover all allocated sites; the log-likelihood and
 * derivative reductions cover the first `sites` of plg_set_active_sites (default: all).  The
 * host layer reads the few trailing per-state sites back through the *_sites getters and
 * applies the correction terms (reference src/likelihood.c:24-119,170-247,321-414,
 * src/core_derivatives.c:654-727).  first_site / count are in sites; scaler ranges are
 * count (x rate_cats with per-rate scalers) values. */
PLL_EXPORT int plg_set_active_sites(plg_context_t * ctx, unsigned int sites);
PLL_EXPORT int plg_get_clv_sites(plg_context_t * ctx, unsigned int clv_index, unsigned int first_site,
                                 unsigned int count, double * out);
PLL_EXPORT int plg_get_scaler_sites(plg_context_t * ctx, unsigned int scaler_index,
                                    unsigned int first_site, unsigned int count, unsigned int * out);
PLL_EXPORT int plg_get_sumtable_sites(plg_context_t * ctx, const void * key, unsigned int first_site,
                                      unsigned int count, double * out);

/* ---- the hot path ---------------------------------------------------------------- */

/* replaces: pll_core_update_pmatrix and its _4x4_avx / _20x20_avx2 rungs
 * (reference src/core_pmatrix.c:24-250, src/core_pmatrix_avx.c:42-310,
 * src/core_pmatrix_avx2.c:37-284).  Per-rate model data is passed already gathered by
 * params_indices: eigenvals[r*states_padded + j], eigenvecs / inv_eigenvecs
 * [r*states*states_padded + ...], prop_invar[r]. */
PLL_EXPORT int plg_update_pmatrix(plg_context_t * ctx,
                                  const unsigned int * matrix_indices,
                                  const double * branch_lengths,
                                  unsigned int count,
                                  const double * rates,
                                  const double * prop_invar,
                                  const double * eigenvals,
                                  const double * eigenvecs,
                                  const double * inv_eigenvecs);

/* replaces: the per-operation loop of pll_update_partials and, per operation,
 * pll_core_create_lookup + pll_core_update_partial_tt / _ti / _ii with fill_parent_scaler and
 * the scaling-threshold rescale (reference src/partials.c:177-213,
 * src/core_partials.c:82-862, src/core_partials_avx.c, src/core_partials_avx2.c:568-803).
 * The whole list is analysed for data dependencies (RAW/WAR/WAW on CLV and scaler slots) and
 * enqueued as one batch: for 4 states ONE kernel walks the list tile by tile with the children
 * held on chip (every CLV / scaler whose value is observable after the call is still written to
 * HBM); otherwise one launch per dependency level and kind.  Repeated lists replay a CUDA graph.
 * Results are identical to executing the list in array order. */
PLL_EXPORT int plg_update_partials(plg_context_t * ctx,
                                   const pll_operation_t * operations,
                                   unsigned int count);

/* replaces: pll_core_edge_loglikelihood_ii / _ti / _ti_4x4 (reference
 * src/core_likelihood.c:211-1002, src/core_likelihood_avx.c:191-406,1079-1266,
 * src/core_likelihood_avx2.c:111-547).  Tip-vs-inner is decided from the clv indices when
 * the context has PLL_ATTRIB_PATTERN_TIP; `freqs` is [rate_cats][states_padded] gathered by
 * freqs_indices, `prop_invar` is [rate_cats].  persite_lnl (host, sites doubles) may be NULL. */
PLL_EXPORT int plg_edge_loglikelihood(plg_context_t * ctx,
                                      unsigned int parent_clv_index,
                                      int parent_scaler_index,
                                      unsigned int child_clv_index,
                                      int child_scaler_index,
                                      unsigned int matrix_index,
                                      const double * freqs,
                                      const double * rate_weights,
                                      const double * prop_invar,
                                      double * persite_lnl,
                                      double * logl_out);

/* replaces: pll_core_root_loglikelihood (reference src/core_likelihood.c:25-209,
 * src/core_likelihood_avx.c:113-189, src/core_likelihood_avx2.c:25-109) */
PLL_EXPORT int plg_root_loglikelihood(plg_context_t * ctx,
                                      unsigned int clv_index,
                                      int scaler_index,
                                      const double * freqs,
                                      const double * rate_weights,
                                      const double * prop_invar,
                                      double * persite_lnl,
                                      double * logl_out);

/* plg_root_loglikelihood with one scaler count per pattern supplied by the caller (host array of
 * ctx->sites entries) instead of a scale buffer: with PLL_ATTRIB_RATE_SCALERS the reference's root
 * kernels read element n of the [site][rate] array for pattern n (src/core_likelihood_avx.c:176-178),
 * which on a partition cut into pattern slices lives in another slice - the host layer gathers the
 * counts and hands every slice its own. */
PLL_EXPORT int plg_root_loglikelihood_counts(plg_context_t * ctx,
                                             unsigned int clv_index,
                                             const unsigned int * site_counts,
                                             const double * freqs,
                                             const double * rate_weights,
                                             const double * prop_invar,
                                             double * persite_lnl,
                                             double * logl_out);

/* replaces: pll_core_update_sumtable_ii / _ti (reference src/core_derivatives.c:125-446,
 * src/core_derivatives_avx.c:25-207,462-645, src/core_derivatives_avx2.c:24-521).
 * `eigenvecs` is [rate_cats][states][states_padded] gathered by params_indices.  `left_terms`
 * is the small per-call table the reference kernels precompute before their site loop:
 *   inner-inner: W[r][j][i] = inv_eigenvecs[r][i][j] * freqs[r][i]
 *                (reference src/core_derivatives_avx.c:86-95)
 *   tip-inner  : L[code][r][j] = sum over states i in `code` of inv_eigenvecs[r][i][j]*freqs[r][i]
 *                (reference src/core_derivatives_avx.c:556-579, src/core_derivatives_avx2.c:368-410)
 * The table is written to the device slot associated with `key` (the caller's host sumtable
 * pointer is used as an opaque key); host_copy, if not NULL, additionally receives it. */
PLL_EXPORT int plg_update_sumtable(plg_context_t * ctx,
                                   unsigned int parent_clv_index,
                                   unsigned int child_clv_index,
                                   int parent_scaler_index,
                                   int child_scaler_index,
                                   const double * eigenvecs,
                                   const double * left_terms,
                                   const void * key,
                                   double * host_copy);
/* Uploads a host sumtable (sites x rate_cats x states_padded doubles) into the device slot of
 * `key`: the table argument of the reference's pll_core_likelihood_derivatives (src/pll.h:943-961). */
PLL_EXPORT int plg_set_sumtable(plg_context_t * ctx, const void * key, const double * table);
/* Releases the device slot of `key` (all slots are released by plg_destroy). */
PLL_EXPORT int plg_free_sumtable(plg_context_t * ctx, const void * key);

/* replaces: pll_core_likelihood_derivatives + _avx2 (reference src/core_derivatives.c:501-732,
 * src/core_derivatives_avx2.c:523-800).  `diagptable` is the host-built
 * [rate_cats][states][4] table {e, lk e, (lk)^2 e, 0} (reference src/core_derivatives.c:560-575).
 * Returns d(-lnL)/dt and d2(-lnL)/dt2. */
PLL_EXPORT int plg_likelihood_derivatives(plg_context_t * ctx,
                                          const void * key,
                                          const double * diagptable,
                                          const double * rate_weights,
                                          const double * prop_invar,
                                          const double * freqs,
                                          double * d_f,
                                          double * dd_f);

/* ---- measurement ------------------------------------------------------------------ */
/* CUDA events recorded on the context's own stream (the stream the kernels run on). */
PLL_EXPORT int plg_timer_start(plg_context_t * ctx);
PLL_EXPORT int plg_timer_stop(plg_context_t * ctx, float * elapsed_ms);
PLL_EXPORT int plg_get_stats(plg_context_t * ctx, plg_stats_t * out);
/* Per-launch event timing of plg_update_partials (disables graph replay while on; each call
 * then synchronises).  For roofline measurement inside a benchmark, not for production. */
PLL_EXPORT int plg_set_profiling(plg_context_t * ctx, int enable);
PLL_EXPORT int plg_reset_stats(plg_context_t * ctx);
/* Overwrites a device buffer larger than L2 (126 MB) on the context's stream. */
PLL_EXPORT int plg_flush_l2(plg_context_t * ctx);
/* Free / total device memory in bytes. */
PLL_EXPORT int plg_mem_info(plg_context_t * ctx, size_t * free_bytes, size_t * total_bytes);

/* Site-pattern compression on the device; replaces the sort / unique core of
 * pll_compress_site_patterns (reference src/compress.c:33-81,202-280).  rows[t] are `taxa` host
 * buffers of `length` characters; they are overwritten with the unique columns in the
 * reference's sorted order (first *unique_out characters of every row), decoded through
 * inverse_table; weights_out (room for `length` entries) receives the multiplicities.
 * code_table / inverse_table are the 256-entry byte tables the reference derives from the state
 * map (src/compress.c:83-108,161-175).  No partition is involved; device < 0 = current. */
PLL_EXPORT int plg_compress_patterns(int device, unsigned char * const * rows, unsigned int taxa,
                                     size_t length, const unsigned char * code_table,
                                     const unsigned char * inverse_table, unsigned int * weights_out,
                                     size_t * unique_out);

/* ================= pll_gpu_*: extensions on a GPU partition ========================== */

/* pll_compress_site_patterns (reference src/compress.c:138-286) computed on the device: same
 * arguments, same outputs (compressed 0-terminated sequences in sorted column order, malloc'ed
 * weights, *length updated). */
PLL_EXPORT unsigned int * pll_gpu_compress_site_patterns(char ** sequence,
                                                         const unsigned int * map,
                                                         int count,
                                                         int * length);

/* Device used by subsequent pll_partition_create(PLL_ATTRIB_ARCH_GPU) calls of this thread;
 * default: $PLL_GPU_DEVICE if set, else $LOCAL_RANK if set, else the current CUDA device. */
PLL_EXPORT int pll_gpu_set_device(int device);
PLL_EXPORT int pll_gpu_device_count(void);

/* Several GPUs behind ONE partition, in one process: partitions created by this thread after
 * pll_gpu_set_devices(n) (n = 0 restores the default: $PLL_GPU_DEVICES, else 1) are cut into n contiguous pattern
 * slices (64-pattern aligned; fewer if the alignment is that short), slice d living on device
 * (first device + d) mod pll_gpu_device_count().  Every pll.h call fans out to all slices -
 * setters and pll_update_partials only enqueue, so the devices run concurrently - and
 * log-likelihoods / derivatives are the per-slice partial results added on the host in slice
 * order (the only cross-device exchange on the path: reference src/core_likelihood_avx.c:1259
 * `logl +=`, src/core_derivatives_avx2.c:756-765 are the only statements that couple sites).
 * Nothing in the caller changes (ascertainment-bias correction included: its per-state sites live
 * in the last slice).  Value-returning calls issue their per-device launches from one helper
 * thread per further slice (they spin while calls keep coming, nap when idle; PLL_GPU_HOST_THREADS=0
 * keeps every launch on the calling thread). */
PLL_EXPORT int pll_gpu_set_devices(int count);
/* The slicing rule itself: writes first_site[0 .. n] for `sites` patterns cut into at most
 * `slices` 64-pattern-aligned slices (first_site[n] = sites) and returns n <= slices; first_site
 * needs room for slices + 1 entries. */
PLL_EXPORT unsigned int pll_gpu_slice_bounds(unsigned int sites, unsigned int slices,
                                             unsigned int * first_site);
/* Number of pattern slices of a partition (0 if not a GPU partition). */
PLL_EXPORT int pll_gpu_partition_devices(const pll_partition_t * partition);
/* Context of one slice and the pattern range it owns (NULL if out of range). */
PLL_EXPORT plg_context_t * pll_gpu_context_of(const pll_partition_t * partition, unsigned int slice,
                                              unsigned int * first_site, unsigned int * sites);

/* The device context behind a GPU partition (NULL if not a GPU partition); with several
 * pattern slices, the context of slice 0. */
PLL_EXPORT plg_context_t * pll_gpu_context(const pll_partition_t * partition);

/* Download one array into its host mirror (allocated on first use):
 * partition->clv[i], ->scale_buffer[i], ->tipchars[i], ->pmatrix[i] become valid. */
PLL_EXPORT int pll_gpu_sync_clv(pll_partition_t * partition, unsigned int clv_index);
PLL_EXPORT int pll_gpu_sync_scaler(pll_partition_t * partition, unsigned int scaler_index);
PLL_EXPORT int pll_gpu_sync_tipchars(pll_partition_t * partition, unsigned int tip_index);
PLL_EXPORT int pll_gpu_sync_pmatrix(pll_partition_t * partition, unsigned int matrix_index);
/* Upload the host mirror partition->pmatrix[i] / ->clv[i] to the device (tests, callers that
 * fill P-matrices or CLVs by hand). */
PLL_EXPORT int pll_gpu_push_pmatrix(pll_partition_t * partition, unsigned int matrix_index);
PLL_EXPORT int pll_gpu_push_clv(pll_partition_t * partition, unsigned int clv_index);

PLL_EXPORT int pll_gpu_synchronize(pll_partition_t * partition);

/* One rank per GPU under mpirun / torchrun, every rank holding a pattern slice in its own
 * partition: after pll_gpu_comm_init (same device rule as pll_partition_create) the
 * log-likelihoods and derivatives returned by the pll.h calls of partitions created afterwards
 * are the sums over all ranks (scalar NCCL all-reduce inside the call).  Rank 0 obtains the id
 * with pll_gpu_comm_unique_id and the caller distributes it. */
PLL_EXPORT int pll_gpu_comm_unique_id(unsigned char id[PLL_GPU_COMM_ID_BYTES]);
PLL_EXPORT int pll_gpu_comm_init(const unsigned char id[PLL_GPU_COMM_ID_BYTES], int nranks, int rank);
PLL_EXPORT int pll_gpu_comm_finalize(void);

/* pll_set_tip_states for a synthetic DNA tip whose characters are generated on the device
 * (plg_generate_tipchars); `first_site` is the alignment column of the partition's pattern 0. */
PLL_EXPORT int pll_gpu_generate_tip_states(pll_partition_t * partition, unsigned int tip_index,
                                           unsigned long long seed, unsigned long long first_site);

/* The `sumtable` argument of pll_update_sumtable / pll_compute_likelihood_derivatives is an
 * opaque key under the GPU flag: the table lives in HBM (sites * rate_cats * states_padded
 * doubles per key) until the partition is destroyed.  A caller that frees its host buffer
 * earlier calls this first; tables that were never released are reclaimed least-recently-used
 * first when device memory runs out. */
PLL_EXPORT int pll_gpu_free_sumtable(pll_partition_t * partition, const double * sumtable);

#ifdef __cplusplus
}
#endif

#endif /* PLL_B200_PLL_GPU_H_ */
