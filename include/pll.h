/*
 * pll.h - public API of the B200-native phylogenetic-likelihood library.
 *
 * This header is a fresh, ABI-compatible re-declaration of the part of the reference's public
 * interface (reference src/pll.h) that lies on the likelihood hot path.  Constants, struct field
 * order/types and prototypes are kept so that an existing caller recompiles against it unchanged:
 *
 *   sizeof(pll_partition_t) == 216 and sizeof(pll_operation_t) == 32 on x86-64
 *   (reference src/pll.h:202-244 and :249-259).
 *
 * The only new surface is PLL_ATTRIB_ARCH_GPU (next to the PLL_ATTRIB_ARCH_* flags of reference
 * src/pll.h:106-111) and the optional pll_gpu_* extension calls declared in pll_gpu.h.
 *
 * This build implements the GPU architecture ONLY: there is no CPU fallback.  A partition
 * created without PLL_ATTRIB_ARCH_GPU, or on a machine without a usable CUDA device, fails with
 * pll_errno set.
 *
 * Environment switches for callers that cannot be recompiled (all off by default):
 *   PLL_GPU_FORCE=1     programs built against the reference's header ask for
 *                       PLL_ATTRIB_ARCH_CPU/SSE/AVX/AVX2: replace those bits by PLL_ATTRIB_ARCH_GPU
 *   PLL_GPU_MIRROR=1    keep partition->clv[i] / ->scale_buffer[i] / ->tipchars[i] current for
 *                       callers that read them directly (every result is downloaded)
 *   PLL_GPU_DEVICES=n   spread every partition over n GPUs of this process (pll_gpu.h:
 *                       pll_gpu_set_devices)
 */
#ifndef PLL_B200_PLL_H_
#define PLL_B200_PLL_H_

#include <stddef.h>
#include <stdio.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLL_EXPORT __attribute__((visibility("default")))

/* ---- return codes (reference src/pll.h:75-79) ---- */
#define PLL_FAILURE 0
#define PLL_SUCCESS 1
#define PLL_FALSE 0
#define PLL_TRUE 1

/* ---- alignment of host-side arrays (reference src/pll.h:81-83) ---- */
#define PLL_ALIGNMENT_CPU 8
#define PLL_ALIGNMENT_SSE 16
#define PLL_ALIGNMENT_AVX 32
#define PLL_ALIGNMENT_GPU 32 /* host mirrors; device arrays are 256-byte aligned */

#define PLL_ASCII_SIZE 256

/* ---- numerical scaling (reference src/pll.h:89-97): 2^256 and friends, all exact ---- */
#define PLL_SCALE_FACTOR 0x1p+256
#define PLL_SCALE_THRESHOLD 0x1p-256
#define PLL_SCALE_FACTOR_SQRT 0x1p+128
#define PLL_SCALE_THRESHOLD_SQRT 0x1p-128
#define PLL_SCALE_BUFFER_NONE (-1)
#define PLL_SCALE_RATE_MAXDIFF 4

#define PLL_MISC_EPSILON 1e-8
#define PLL_ONE_EPSILON 1e-15
#define PLL_ONE_MIN (1 - PLL_ONE_EPSILON)
#define PLL_ONE_MAX (1 + PLL_ONE_EPSILON)

/* ---- attribute flags (reference src/pll.h:106-122) ---- */
#define PLL_ATTRIB_ARCH_CPU 0
#define PLL_ATTRIB_ARCH_SSE (1 << 0)
#define PLL_ATTRIB_ARCH_AVX (1 << 1)
#define PLL_ATTRIB_ARCH_AVX2 (1 << 2)
#define PLL_ATTRIB_ARCH_AVX512 (1 << 3)
#define PLL_ATTRIB_ARCH_MASK 0xF

#define PLL_ATTRIB_PATTERN_TIP (1 << 4)

#define PLL_ATTRIB_AB_LEWIS (1 << 5)
#define PLL_ATTRIB_AB_FELSENSTEIN (2 << 5)
#define PLL_ATTRIB_AB_STAMATAKIS (3 << 5)
#define PLL_ATTRIB_AB_MASK (7 << 5)
#define PLL_ATTRIB_AB_FLAG (1 << 8)

#define PLL_ATTRIB_RATE_SCALERS (1 << 9)

/* NEW: the B200 backend.  Lies outside PLL_ATTRIB_ARCH_MASK so the reference's popcount check
 * on the SIMD bits (reference src/pll.c:413-418) is unaffected; combining it with any SIMD
 * bit is rejected explicitly. */
#define PLL_ATTRIB_ARCH_GPU (1 << 10)

/* ---- error codes (reference src/pll.h:137-167); >= 131 are new (CUDA layer) ---- */
#define PLL_ERROR_FILE_OPEN 100
#define PLL_ERROR_FILE_SEEK 101
#define PLL_ERROR_FILE_EOF 102
#define PLL_ERROR_FASTA_ILLEGALCHAR 103
#define PLL_ERROR_FASTA_UNPRINTABLECHAR 104
#define PLL_ERROR_FASTA_INVALIDHEADER 105
#define PLL_ERROR_PHYLIP_SYNTAX 106
#define PLL_ERROR_PHYLIP_LONGSEQ 107
#define PLL_ERROR_PHYLIP_NONALIGNED 108
#define PLL_ERROR_PHYLIP_ILLEGALCHAR 109
#define PLL_ERROR_PHYLIP_UNPRINTABLECHAR 110
#define PLL_ERROR_NEWICK_SYNTAX 111
#define PLL_ERROR_MEM_ALLOC 112
#define PLL_ERROR_PARAM_INVALID 113
#define PLL_ERROR_TIPDATA_ILLEGALSTATE 114
#define PLL_ERROR_TIPDATA_ILLEGALFUNCTION 115
#define PLL_ERROR_TREE_CONVERSION 116
#define PLL_ERROR_INVAR_INCOMPAT 117
#define PLL_ERROR_INVAR_PROPORTION 118
#define PLL_ERROR_INVAR_PARAMINDEX 119
#define PLL_ERROR_INVAR_NONEFOUND 120
#define PLL_ERROR_AB_INVALIDMETHOD 121
#define PLL_ERROR_AB_NOSUPPORT 122
#define PLL_ERROR_EINVAL 130
#define PLL_ERROR_GPU_NODEVICE 131   /* no usable CUDA device / wrong architecture */
#define PLL_ERROR_GPU_RUNTIME 132    /* a CUDA runtime call or kernel launch failed */
#define PLL_ERROR_GPU_UNSUPPORTED 133 /* feature not implemented by the GPU backend */

#define PLL_GAMMA_RATES_MEAN 0
#define PLL_GAMMA_RATES_MEDIAN 1

/* fields printed by pll_utree_show_ascii (reference src/pll.h:171-175) */
#define PLL_UTREE_SHOW_LABEL (1 << 0)
#define PLL_UTREE_SHOW_BRANCH_LENGTH (1 << 1)
#define PLL_UTREE_SHOW_CLV_INDEX (1 << 2)
#define PLL_UTREE_SHOW_SCALER_INDEX (1 << 3)
#define PLL_UTREE_SHOW_PMATRIX_INDEX (1 << 4)

#define PLL_TREE_TRAVERSE_POSTORDER 1
#define PLL_TREE_TRAVERSE_PREORDER 2

/* ---- the partition (field order and types: reference src/pll.h:202-244) ----
 *
 * Under PLL_ATTRIB_ARCH_GPU the big arrays live in HBM for the partition's lifetime:
 *   clv[i], scale_buffer[i], tipchars[i]   are NULL until pll_gpu_sync_clv / _scaler /
 *                                          _tipchars downloads a host mirror (pll_gpu.h), or
 *                                          always current with PLL_GPU_MIRROR=1;
 *   pmatrix[i]                             host mirror, refreshed by pll_gpu_sync_pmatrix;
 *   rates, rate_weights, subst_params, frequencies, prop_invar, eigen*, pattern_weights,
 *   invariant, charmap, tipmap             are ordinary, always-valid host arrays.
 */
typedef struct pll_partition
{
  unsigned int tips;
  unsigned int clv_buffers;
  unsigned int states;
  unsigned int sites;
  unsigned int pattern_weight_sum;
  unsigned int rate_matrices;
  unsigned int prob_matrices;
  unsigned int rate_cats;
  unsigned int scale_buffers;
  unsigned int attributes;

  size_t alignment;
  unsigned int states_padded;

  double ** clv;
  double ** pmatrix;
  double * rates;
  double * rate_weights;
  double ** subst_params;
  unsigned int ** scale_buffer;
  double ** frequencies;
  double * prop_invar;
  int * invariant;
  unsigned int * pattern_weights;

  int * eigen_decomp_valid;
  double ** eigenvecs;
  double ** inv_eigenvecs;
  double ** eigenvals;

  unsigned int maxstates;
  unsigned char ** tipchars;
  unsigned char * charmap;
  double * ttlookup;
  unsigned int * tipmap;

  int asc_bias_alloc;
} pll_partition_t;

/* ---- one CLV update (reference src/pll.h:249-259) ---- */
typedef struct pll_operation
{
  unsigned int parent_clv_index;
  int parent_scaler_index;
  unsigned int child1_clv_index;
  unsigned int child1_matrix_index;
  int child1_scaler_index;
  unsigned int child2_clv_index;
  unsigned int child2_matrix_index;
  int child2_scaler_index;
} pll_operation_t;

/* ---- unrooted tree: a tip is one record (next == NULL); an inner node is a ring of three
 * records linked by `next`, each facing one neighbour through `back`
 * (reference src/pll.h:312-334) ---- */
typedef struct pll_unode_s
{
  char * label;
  double length;
  unsigned int node_index;
  unsigned int clv_index;
  int scaler_index;
  unsigned int pmatrix_index;
  struct pll_unode_s * next;
  struct pll_unode_s * back;
  void * data;
} pll_unode_t;

typedef struct pll_utree_s
{
  unsigned int tip_count;
  unsigned int inner_count;
  unsigned int edge_count;
  pll_unode_t ** nodes; /* tips first, then inner nodes in post-order, the parse root last */
} pll_utree_t;

/* ---- FASTA reader state (reference src/pll.h:85,282-292) ---- */
#define PLL_LINEALLOC 2048
typedef struct pll_fasta
{
  FILE * fp;
  char line[PLL_LINEALLOC];
  const unsigned int * chrstatus;
  long no;
  long filesize;
  long lineno;
  long stripped_count;
  long stripped[256];
} pll_fasta_t;

/* ---- multiple sequence alignment and the PHYLIP reader state (reference
 * src/pll.h:271-278,295-308) ---- */
/* rooted binary tree: plain left / right / parent links (reference src/pll.h:336-361) */
typedef struct pll_rnode_s
{
  char * label;
  double length;
  unsigned int node_index;
  unsigned int clv_index;
  int scaler_index;
  unsigned int pmatrix_index;
  struct pll_rnode_s * left;
  struct pll_rnode_s * right;
  struct pll_rnode_s * parent;
  void * data;
} pll_rnode_t;

typedef struct pll_rtree_s
{
  unsigned int tip_count;
  unsigned int inner_count;
  unsigned int edge_count;
  pll_rnode_t ** nodes; /* tips first, then inner nodes in post-order, the root last */
  pll_rnode_t * root;
} pll_rtree_t;

typedef struct pll_msa_s
{
  int count;
  int length;
  char ** sequence;
  char ** label;
} pll_msa_t;

typedef struct pll_phylip_s
{
  FILE * fp;
  char * line;
  size_t line_size;
  size_t line_maxsize;
  char buffer[PLL_LINEALLOC];
  const unsigned int * chrstatus;
  long no;
  long filesize;
  long lineno;
  long stripped_count;
  long stripped[256];
} pll_phylip_t;

/* ---- thread-local error channel (reference src/pll.h:470-471, src/pll.c:24-25) ---- */
PLL_EXPORT extern __thread int pll_errno;
PLL_EXPORT extern __thread char pll_errmsg[200];

/* ---- character -> state-mask maps (reference src/maps.c:26-110) ---- */
PLL_EXPORT extern const unsigned int pll_map_bin[256];
PLL_EXPORT extern const unsigned int pll_map_nt[256];
PLL_EXPORT extern const unsigned int pll_map_aa[256];

/* ---- character status tables of the file readers: 0 stripped, 1 kept, 2 fatal, 3 silently
 * stripped (reference src/maps.c:117-168) ---- */
PLL_EXPORT extern const unsigned int pll_map_fasta[256];
PLL_EXPORT extern const unsigned int pll_map_phylip[256];

/* ---- empirical amino-acid models (reference src/pll.h:480-522, tables src/maps.c:172-1165):
 * 190 exchangeabilities in pll_set_subst_params order + 20 frequencies each ---- */
PLL_EXPORT extern const double pll_aa_rates_dayhoff[190];
PLL_EXPORT extern const double pll_aa_rates_lg[190];
PLL_EXPORT extern const double pll_aa_rates_dcmut[190];
PLL_EXPORT extern const double pll_aa_rates_jtt[190];
PLL_EXPORT extern const double pll_aa_rates_mtrev[190];
PLL_EXPORT extern const double pll_aa_rates_wag[190];
PLL_EXPORT extern const double pll_aa_rates_rtrev[190];
PLL_EXPORT extern const double pll_aa_rates_cprev[190];
PLL_EXPORT extern const double pll_aa_rates_vt[190];
PLL_EXPORT extern const double pll_aa_rates_blosum62[190];
PLL_EXPORT extern const double pll_aa_rates_mtmam[190];
PLL_EXPORT extern const double pll_aa_rates_mtart[190];
PLL_EXPORT extern const double pll_aa_rates_mtzoa[190];
PLL_EXPORT extern const double pll_aa_rates_pmb[190];
PLL_EXPORT extern const double pll_aa_rates_hivb[190];
PLL_EXPORT extern const double pll_aa_rates_hivw[190];
PLL_EXPORT extern const double pll_aa_rates_jttdcmut[190];
PLL_EXPORT extern const double pll_aa_rates_flu[190];
PLL_EXPORT extern const double pll_aa_rates_stmtrev[190];
PLL_EXPORT extern const double pll_aa_freqs_dayhoff[20];
PLL_EXPORT extern const double pll_aa_freqs_lg[20];
PLL_EXPORT extern const double pll_aa_freqs_dcmut[20];
PLL_EXPORT extern const double pll_aa_freqs_jtt[20];
PLL_EXPORT extern const double pll_aa_freqs_mtrev[20];
PLL_EXPORT extern const double pll_aa_freqs_wag[20];
PLL_EXPORT extern const double pll_aa_freqs_rtrev[20];
PLL_EXPORT extern const double pll_aa_freqs_cprev[20];
PLL_EXPORT extern const double pll_aa_freqs_vt[20];
PLL_EXPORT extern const double pll_aa_freqs_blosum62[20];
PLL_EXPORT extern const double pll_aa_freqs_mtmam[20];
PLL_EXPORT extern const double pll_aa_freqs_mtart[20];
PLL_EXPORT extern const double pll_aa_freqs_mtzoa[20];
PLL_EXPORT extern const double pll_aa_freqs_pmb[20];
PLL_EXPORT extern const double pll_aa_freqs_hivb[20];
PLL_EXPORT extern const double pll_aa_freqs_hivw[20];
PLL_EXPORT extern const double pll_aa_freqs_jttdcmut[20];
PLL_EXPORT extern const double pll_aa_freqs_flu[20];
PLL_EXPORT extern const double pll_aa_freqs_stmtrev[20];
PLL_EXPORT extern const double pll_aa_rates_lg4m[4][190];
PLL_EXPORT extern const double pll_aa_freqs_lg4m[4][20];
PLL_EXPORT extern const double pll_aa_rates_lg4x[4][190];
PLL_EXPORT extern const double pll_aa_freqs_lg4x[4][20];

/* ---- partition lifecycle and tip data (reference src/pll.c:399-1059) ---- */
PLL_EXPORT pll_partition_t * pll_partition_create(unsigned int tips,
                                                  unsigned int clv_buffers,
                                                  unsigned int states,
                                                  unsigned int sites,
                                                  unsigned int rate_matrices,
                                                  unsigned int prob_matrices,
                                                  unsigned int rate_cats,
                                                  unsigned int scale_buffers,
                                                  unsigned int attributes);

PLL_EXPORT void pll_partition_destroy(pll_partition_t * partition);

PLL_EXPORT int pll_set_tip_states(pll_partition_t * partition,
                                  unsigned int tip_index,
                                  const unsigned int * map,
                                  const char * sequence);

PLL_EXPORT int pll_set_tip_clv(pll_partition_t * partition,
                               unsigned int tip_index,
                               const double * clv,
                               int padding);

PLL_EXPORT void pll_set_pattern_weights(pll_partition_t * partition,
                                        const unsigned int * pattern_weights);

PLL_EXPORT void * pll_aligned_alloc(size_t size, size_t alignment);
PLL_EXPORT void pll_aligned_free(void * ptr);

/* ---- model parameters (reference src/models.c:251-647) ---- */
PLL_EXPORT void pll_set_subst_params(pll_partition_t * partition,
                                     unsigned int params_index,
                                     const double * params);

PLL_EXPORT void pll_set_frequencies(pll_partition_t * partition,
                                    unsigned int params_index,
                                    const double * frequencies);

PLL_EXPORT void pll_set_category_rates(pll_partition_t * partition,
                                       const double * rates);

PLL_EXPORT void pll_set_category_weights(pll_partition_t * partition,
                                         const double * rate_weights);

PLL_EXPORT int pll_update_eigen(pll_partition_t * partition,
                                unsigned int params_index);

PLL_EXPORT int pll_update_prob_matrices(pll_partition_t * partition,
                                        const unsigned int * params_indices,
                                        const unsigned int * matrix_indices,
                                        const double * branch_lengths,
                                        unsigned int count);

PLL_EXPORT unsigned int pll_count_invariant_sites(pll_partition_t * partition,
                                                  unsigned int * state_inv_count);

PLL_EXPORT int pll_update_invariant_sites(pll_partition_t * partition);

PLL_EXPORT int pll_update_invariant_sites_proportion(pll_partition_t * partition,
                                                     unsigned int params_index,
                                                     double prop_invar);

/* ---- CLV updates (reference src/partials.c:177-213) ---- */
PLL_EXPORT void pll_update_partials(pll_partition_t * partition,
                                    const pll_operation_t * operations,
                                    unsigned int count);

/* ---- log-likelihood (reference src/likelihood.c:121-168, :478-513) ---- */
PLL_EXPORT double pll_compute_root_loglikelihood(pll_partition_t * partition,
                                                 unsigned int clv_index,
                                                 int scaler_index,
                                                 const unsigned int * freqs_indices,
                                                 double * persite_lnl);

PLL_EXPORT double pll_compute_edge_loglikelihood(pll_partition_t * partition,
                                                 unsigned int parent_clv_index,
                                                 int parent_scaler_index,
                                                 unsigned int child_clv_index,
                                                 int child_scaler_index,
                                                 unsigned int matrix_index,
                                                 const unsigned int * freqs_indices,
                                                 double * persite_lnl);

/* ---- branch-length derivatives (reference src/derivatives.c:164-312) ----
 * `sumtable` is the caller's buffer exactly as in the reference (sites * rate_cats *
 * states_padded doubles).  Under the GPU backend the table itself stays in HBM, keyed by
 * this pointer; the host buffer is only written when PLL_GPU_SUMTABLE_HOSTCOPY=1 is set in
 * the environment (see pll_gpu.h). */
PLL_EXPORT int pll_update_sumtable(pll_partition_t * partition,
                                   unsigned int parent_clv_index,
                                   unsigned int child_clv_index,
                                   int parent_scaler_index,
                                   int child_scaler_index,
                                   const unsigned int * params_indices,
                                   double * sumtable);

PLL_EXPORT int pll_compute_likelihood_derivatives(pll_partition_t * partition,
                                                  int parent_scaler_index,
                                                  int child_scaler_index,
                                                  double branch_length,
                                                  const unsigned int * params_indices,
                                                  const double * sumtable,
                                                  double * d_f,
                                                  double * dd_f);

/* ---- ascertainment-bias correction (reference src/pll.c:1061-1116).  The partition must be
 * created with PLL_ATTRIB_AB_FLAG or one of the PLL_ATTRIB_AB_* types: it then holds `states`
 * extra per-state sites. ---- */
PLL_EXPORT int pll_set_asc_bias_type(pll_partition_t * partition, int asc_bias_type);
PLL_EXPORT void pll_set_asc_state_weights(pll_partition_t * partition,
                                          const unsigned int * state_weights);

/* ---- discrete Gamma rates (reference src/gamma.c:220-292) ---- */
PLL_EXPORT int pll_compute_gamma_cats(double alpha,
                                      unsigned int categories,
                                      double * output_rates,
                                      int rates_mode);

/* ---- site-pattern compression (reference src/compress.c:138-286) ---- */
PLL_EXPORT unsigned int * pll_compress_site_patterns(char ** sequence,
                                                     const unsigned int * map,
                                                     int count,
                                                     int * length);

/* ---- FASTA reader (reference src/fasta.c:39-323, prototypes src/pll.h:668-681) ---- */
PLL_EXPORT pll_fasta_t * pll_fasta_open(const char * filename, const unsigned int * map);
PLL_EXPORT int pll_fasta_getnext(pll_fasta_t * fd, char ** head, long * head_len, char ** seq,
                                 long * seq_len, long * seqno);
PLL_EXPORT void pll_fasta_close(pll_fasta_t * fd);
PLL_EXPORT long pll_fasta_getfilesize(const pll_fasta_t * fd);
PLL_EXPORT long pll_fasta_getfilepos(pll_fasta_t * fd);
PLL_EXPORT int pll_fasta_rewind(pll_fasta_t * fd);

/* ---- PHYLIP reader, sequential and interleaved (reference src/phylip.c:249-730, prototypes
 * src/pll.h:768-779) ---- */
PLL_EXPORT void pll_msa_destroy(pll_msa_t * msa);
PLL_EXPORT pll_phylip_t * pll_phylip_open(const char * filename, const unsigned int * map);
PLL_EXPORT int pll_phylip_rewind(pll_phylip_t * fd);
PLL_EXPORT void pll_phylip_close(pll_phylip_t * fd);
PLL_EXPORT pll_msa_t * pll_phylip_parse_interleaved(pll_phylip_t * fd);
PLL_EXPORT pll_msa_t * pll_phylip_parse_sequential(pll_phylip_t * fd);

/* ---- unrooted trees: Newick reader, traversal, traversal -> operations
 * (reference src/parse_utree.y:71-524, src/utree.c:217-442, prototypes src/pll.h:702-760).
 * Hand-written and non-recursive here (no flex/bison, no call-stack depth on deep trees). ---- */
PLL_EXPORT pll_utree_t * pll_utree_parse_newick(const char * filename);
PLL_EXPORT pll_utree_t * pll_utree_parse_newick_string(const char * s);
PLL_EXPORT void pll_utree_destroy(pll_utree_t * tree, void (*cb_destroy)(void *));
PLL_EXPORT void pll_utree_graph_destroy(pll_unode_t * root, void (*cb_destroy)(void *));
PLL_EXPORT void pll_utree_reset_template_indices(pll_unode_t * node, unsigned int tip_count);
PLL_EXPORT pll_utree_t * pll_utree_wraptree(pll_unode_t * root, unsigned int tip_count);
PLL_EXPORT char * pll_utree_export_newick(const pll_unode_t * root,
                                          char * (*cb_serialize)(const pll_unode_t *));
/* ASCII drawing on stdout (reference src/utree.c:122-157); options = PLL_UTREE_SHOW_* */
PLL_EXPORT void pll_utree_show_ascii(const pll_unode_t * root, int options);
PLL_EXPORT int pll_utree_traverse(pll_unode_t * root,
                                  int traversal,
                                  int (*cbtrav)(pll_unode_t *),
                                  pll_unode_t ** outbuffer,
                                  unsigned int * trav_size);
PLL_EXPORT int pll_utree_every(pll_utree_t * tree, int (*cb)(pll_unode_t *));
PLL_EXPORT int pll_utree_every_const(const pll_utree_t * tree, int (*cb)(const pll_unode_t *));
PLL_EXPORT int pll_utree_check_integrity(const pll_utree_t * tree);
PLL_EXPORT pll_unode_t * pll_utree_graph_clone(const pll_unode_t * root);
PLL_EXPORT pll_utree_t * pll_utree_clone(const pll_utree_t * tree);
PLL_EXPORT void pll_utree_create_operations(pll_unode_t * const * trav_buffer,
                                            unsigned int trav_buffer_size,
                                            double * branches,
                                            unsigned int * pmatrix_indices,
                                            pll_operation_t * ops,
                                            unsigned int * matrix_count,
                                            unsigned int * ops_count);

/* ---- rooted trees: Newick reader (strictly binary, also at the root), index template,
 * traversal, traversal -> operations for pll_compute_root_loglikelihood, export, drawing
 * (reference src/parse_rtree.y:46-400, src/rtree.c:24-330, prototypes src/pll.h:685-810).
 * Iterative like the unrooted code. ---- */
PLL_EXPORT pll_rtree_t * pll_rtree_parse_newick(const char * filename);
PLL_EXPORT pll_rtree_t * pll_rtree_parse_newick_string(const char * s);
PLL_EXPORT void pll_rtree_destroy(pll_rtree_t * tree, void (*cb_destroy)(void *));
PLL_EXPORT void pll_rtree_graph_destroy(pll_rnode_t * root, void (*cb_destroy)(void *));
PLL_EXPORT void pll_rtree_reset_template_indices(pll_rnode_t * root, unsigned int tip_count);
PLL_EXPORT pll_rtree_t * pll_rtree_wraptree(pll_rnode_t * root, unsigned int tip_count);
PLL_EXPORT void pll_rtree_show_ascii(const pll_rnode_t * root, int options);
PLL_EXPORT char * pll_rtree_export_newick(const pll_rnode_t * root,
                                          char * (*cb_serialize)(const pll_rnode_t *));
PLL_EXPORT int pll_rtree_traverse(pll_rnode_t * root,
                                  int traversal,
                                  int (*cbtrav)(pll_rnode_t *),
                                  pll_rnode_t ** outbuffer,
                                  unsigned int * trav_size);
PLL_EXPORT void pll_rtree_create_operations(pll_rnode_t * const * trav_buffer,
                                            unsigned int trav_buffer_size,
                                            double * branches,
                                            unsigned int * pmatrix_indices,
                                            pll_operation_t * ops,
                                            unsigned int * matrix_count,
                                            unsigned int * ops_count);

/* NEW (no reference counterpart): operation list of a FULL traversal towards the edge
 * (root, root->back) that recycles CLV / scaler slots, for partitions created with fewer
 * clv_buffers / scale_buffers than inner nodes (at 5 000 taxa x 10 M patterns one CLV is
 * 1.28 GB).  Subtrees are evaluated larger-demand-first, so `max_slots` >= the tree's Strahler
 * number + 1 (<= log2(tips) + 2) always suffices; *slots_used receives the exact need and the
 * call fails with PLL_ERROR_PARAM_INVALID if it exceeds max_slots.  Inner node results go to
 * CLV tip_count + slot and scaler slot; tips keep clv_index / PLL_SCALE_BUFFER_NONE;
 * P-matrix indices and branches are the tree's own (all 2T-3 of them are listed).
 * edge_clv[2] / edge_scaler[2] receive where the two ends of the evaluation edge ended up
 * ([0] = root, [1] = root->back).  The tree must carry the default index template
 * (pll_utree_parse_newick* or pll_utree_reset_template_indices). */
PLL_EXPORT int pll_utree_create_operations_recycled(pll_unode_t * root,
                                                    unsigned int tip_count,
                                                    unsigned int max_slots,
                                                    double * branches,
                                                    unsigned int * pmatrix_indices,
                                                    pll_operation_t * ops,
                                                    unsigned int * matrix_count,
                                                    unsigned int * ops_count,
                                                    unsigned int * edge_clv,
                                                    int * edge_scaler,
                                                    unsigned int * slots_used);

/* ---- printing helpers (reference src/output.c:26-96); they download the arrays they print
 * from HBM into the host mirrors themselves ---- */
PLL_EXPORT void pll_show_pmatrix(const pll_partition_t * partition,
                                 unsigned int index,
                                 unsigned int float_precision);

PLL_EXPORT void pll_show_clv(const pll_partition_t * partition,
                             unsigned int clv_index,
                             int scaler_index,
                             unsigned int float_precision);

/* ---- the direct-call surface (reference src/pll.h:864-1000, :1659-1700): plain HOST arrays, no
 * partition.  Each call uploads its arguments, runs the kernels the partition API runs and downloads
 * the result (libpll_b200/csrc/host/pll_core.c) - for callers that keep their own arrays; data
 * that should stay in HBM belongs in a partition.  Array layouts follow the architecture bits in
 * `attrib` as in the reference (states_padded = states / even / a multiple of 4 for
 * PLL_ATTRIB_ARCH_CPU / _SSE / _AVX, _AVX2, _GPU).  The `lookup` table of the tip-tip pair is
 * opaque (same size as the reference's).  PLL_ATTRIB_AB_* bits are not supported here. ---- */
/* replaces reference src/core_partials.c:82-186 */
PLL_EXPORT void pll_core_create_lookup(unsigned int states, unsigned int rate_cats, double * lookup,
                                       const double * left_matrix, const double * right_matrix,
                                       const unsigned int * tipmap, unsigned int tipmap_size,
                                       unsigned int attrib);
/* replaces reference src/core_partials.c:188-260 */
PLL_EXPORT void pll_core_update_partial_tt(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                           double * parent_clv, unsigned int * parent_scaler,
                                           const unsigned char * left_tipchars,
                                           const unsigned char * right_tipchars,
                                           const unsigned int * tipmap, unsigned int tipmap_size,
                                           const double * lookup, unsigned int attrib);
/* replaces reference src/core_partials.c:262-532 */
PLL_EXPORT void pll_core_update_partial_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                           double * parent_clv, unsigned int * parent_scaler,
                                           const unsigned char * left_tipchars, const double * right_clv,
                                           const double * left_matrix, const double * right_matrix,
                                           const unsigned int * right_scaler, const unsigned int * tipmap,
                                           unsigned int tipmap_size, unsigned int attrib);
/* replaces reference src/core_partials.c:534-862 */
PLL_EXPORT void pll_core_update_partial_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                           double * parent_clv, unsigned int * parent_scaler,
                                           const double * left_clv, const double * right_clv,
                                           const double * left_matrix, const double * right_matrix,
                                           const unsigned int * left_scaler,
                                           const unsigned int * right_scaler, unsigned int attrib);
/* replaces reference src/core_pmatrix.c:24-250 */
PLL_EXPORT int pll_core_update_pmatrix(double ** pmatrix, unsigned int states, unsigned int rate_cats,
                                       const double * rates, const double * branch_lengths,
                                       const unsigned int * matrix_indices,
                                       const unsigned int * params_indices, const double * prop_invar,
                                       double * const * eigenvals, double * const * eigenvecs,
                                       double * const * inv_eigenvecs, unsigned int count,
                                       unsigned int attrib);
/* replace reference src/core_derivatives.c:125-296 (ii), :298-446 (ti), :501-732 */
PLL_EXPORT int pll_core_update_sumtable_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                           const double * parent_clv, const double * child_clv,
                                           const unsigned int * parent_scaler,
                                           const unsigned int * child_scaler, double * const * eigenvecs,
                                           double * const * inv_eigenvecs, double * const * freqs,
                                           double * sumtable, unsigned int attrib);
PLL_EXPORT int pll_core_update_sumtable_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                           const double * parent_clv, const unsigned char * left_tipchars,
                                           const unsigned int * parent_scaler, double * const * eigenvecs,
                                           double * const * inv_eigenvecs, double * const * freqs,
                                           const unsigned int * tipmap, unsigned int tipmap_size,
                                           double * sumtable, unsigned int attrib);
PLL_EXPORT int pll_core_likelihood_derivatives(unsigned int states, unsigned int sites,
                                               unsigned int rate_cats, const double * rate_weights,
                                               const unsigned int * parent_scaler,
                                               const unsigned int * child_scaler, const int * invariant,
                                               const unsigned int * pattern_weights, double branch_length,
                                               const double * prop_invar, double * const * freqs,
                                               const double * rates, double * const * eigenvals,
                                               const double * sumtable, double * d_f, double * dd_f,
                                               unsigned int attrib);
/* replace reference src/core_likelihood.c:25-209 (root), :211-596 (edge ti), :598-1002 (edge ii) */
PLL_EXPORT double pll_core_root_loglikelihood(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                              const double * clv, const unsigned int * scaler,
                                              double * const * frequencies, const double * rate_weights,
                                              const unsigned int * pattern_weights,
                                              const double * invar_proportion, const int * invar_indices,
                                              const unsigned int * freqs_indices, double * persite_lnl,
                                              unsigned int attrib);
PLL_EXPORT double pll_core_edge_loglikelihood_ti(unsigned int states, unsigned int sites,
                                                 unsigned int rate_cats, const double * parent_clv,
                                                 const unsigned int * parent_scaler,
                                                 const unsigned char * tipchars, const unsigned int * tipmap,
                                                 unsigned int tipmap_size, const double * pmatrix,
                                                 double * const * frequencies, const double * rate_weights,
                                                 const unsigned int * pattern_weights,
                                                 const double * invar_proportion, const int * invar_indices,
                                                 const unsigned int * freqs_indices, double * persite_lnl,
                                                 unsigned int attrib);
PLL_EXPORT double pll_core_edge_loglikelihood_ii(unsigned int states, unsigned int sites,
                                                 unsigned int rate_cats, const double * parent_clv,
                                                 const unsigned int * parent_scaler, const double * child_clv,
                                                 const unsigned int * child_scaler, const double * pmatrix,
                                                 double * const * frequencies, const double * rate_weights,
                                                 const unsigned int * pattern_weights,
                                                 const double * invar_proportion, const int * invar_indices,
                                                 const unsigned int * freqs_indices, double * persite_lnl,
                                                 unsigned int attrib);
/* frees the calling thread's scratch device state of the pll_core_* calls (optional) */
PLL_EXPORT void pll_gpu_core_release(void);

#ifdef __cplusplus
}
#endif

#endif /* PLL_B200_PLL_H_ */
