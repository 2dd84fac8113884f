/*
 * lg4_gpu.c - a data-driven C caller of the pll.h API on the GPU backend, the scenario of the
 * reference's examples/lg4 (examples/lg4/lg4.c:64-430): read an unrooted binary tree (Newick)
 * and an amino-acid alignment (FASTA or PHYLIP), compress the site patterns on the device,
 * build the operation list of a full traversal - once with one CLV per inner node, once with
 * recycled slots - and evaluate the log-likelihood under LG4M and LG4X.
 *
 *   usage: lg4_gpu <tree.newick> <alignment.fas | alignment.phy>
 *
 * Written against include/pll.h / pll_gpu.h only.  Differences to a libpll program: the
 * PLL_ATTRIB_ARCH_GPU flag; optional pll_gpu_compress_site_patterns and
 * pll_utree_create_operations_recycled.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pll.h"
#include "pll_gpu.h"

static void die(const char * what)
{
  fprintf(stderr, "%s (%d): %s\n", what, pll_errno, pll_errmsg);
  exit(2);
}

static int take_all(pll_unode_t * node)
{
  (void)node;
  return 1;
}

static int ends_with(const char * s, const char * suffix)
{
  const size_t n = strlen(s), m = strlen(suffix);
  return n >= m && !strcmp(s + n - m, suffix);
}

static void set_lg4(pll_partition_t * p, const double rates[4][190], const double freqs[4][20])
{
  for (unsigned int i = 0; i < 4; ++i)
  {
    pll_set_frequencies(p, i, freqs[i]);
    pll_set_subst_params(p, i, rates[i]);
  }
}

int main(int argc, char ** argv)
{
  if (argc != 3)
  {
    fprintf(stderr, "usage: %s <tree.newick> <alignment.fas|.phy>\n", argv[0]);
    return 1;
  }
  pll_utree_t * tree = pll_utree_parse_newick(argv[1]);
  if (!tree) die("pll_utree_parse_newick");
  const unsigned int tips = tree->tip_count, inner = tree->inner_count, branches = tree->edge_count;
  for (unsigned int i = 0; i < tips + inner; ++i) /* missing lengths as in lg4.c:37-62 */
  {
    pll_unode_t * n = tree->nodes[i];
    if (!n->length) n->length = 0.000001;
    if (n->next)
    {
      if (!n->next->length) n->next->length = 0.000001;
      if (!n->next->next->length) n->next->next->length = 0.000001;
    }
  }

  /* ---- alignment: rows ordered like the tree's tips ---- */
  char ** rows = (char **)calloc(tips, sizeof(char *));
  int sites = -1;
  unsigned int found = 0;
  if (ends_with(argv[2], ".phy"))
  {
    pll_phylip_t * fd = pll_phylip_open(argv[2], pll_map_phylip);
    if (!fd) die("pll_phylip_open");
    pll_msa_t * msa = pll_phylip_parse_sequential(fd);
    if (!msa) die("pll_phylip_parse_sequential");
    sites = msa->length;
    for (int s = 0; s < msa->count; ++s)
      for (unsigned int t = 0; t < tips; ++t)
        if (!strcmp(tree->nodes[t]->label, msa->label[s]) && !rows[tree->nodes[t]->clv_index])
        {
          rows[tree->nodes[t]->clv_index] = strdup(msa->sequence[s]);
          ++found;
        }
    pll_msa_destroy(msa);
    pll_phylip_close(fd);
  }
  else
  {
    pll_fasta_t * fd = pll_fasta_open(argv[2], pll_map_fasta);
    if (!fd) die("pll_fasta_open");
    char * head, * seq;
    long hl, sl, no;
    while (pll_fasta_getnext(fd, &head, &hl, &seq, &sl, &no))
    {
      if (sites != -1 && sites != (int)sl) die("sequences of different length");
      sites = (int)sl;
      for (unsigned int t = 0; t < tips; ++t)
        if (!strcmp(tree->nodes[t]->label, head) && !rows[tree->nodes[t]->clv_index])
        {
          rows[tree->nodes[t]->clv_index] = seq;
          seq = NULL;
          ++found;
        }
      free(head);
      free(seq);
    }
    if (pll_errno != PLL_ERROR_FILE_EOF) die("pll_fasta_getnext");
    pll_fasta_close(fd);
  }
  if (found != tips || sites <= 0) die("the alignment does not cover the tree's taxa");
  printf("%u taxa, %d sites\n", tips, sites);

  /* ---- site patterns, on the device ---- */
  int patterns = sites;
  unsigned int * weights = pll_gpu_compress_site_patterns(rows, pll_map_aa, (int)tips, &patterns);
  if (!weights) die("pll_gpu_compress_site_patterns");
  printf("%d patterns\n", patterns);

  /* ---- operations: one CLV per inner node, and recycled slots ---- */
  pll_unode_t * root = tree->nodes[tips + inner - 1];
  pll_unode_t ** trav = (pll_unode_t **)malloc((tips + inner) * sizeof(pll_unode_t *));
  double * lengths = (double *)malloc(branches * sizeof(double));
  unsigned int * matrices = (unsigned int *)malloc(branches * sizeof(unsigned int));
  pll_operation_t * ops = (pll_operation_t *)malloc(inner * sizeof(pll_operation_t));
  unsigned int trav_size, n_mat, n_ops, edge_clv[2], slots = 0;
  int edge_scaler[2];
  const unsigned int params[4] = {0, 1, 2, 3};
  const double lg4x_weights[4] = {0.209224645, 0.224707726, 0.277599198, 0.288468431};
  const double lg4x_rates[4] = {0.498991136, 0.563680734, 0.808264032, 1.887769458};
  double gamma[4];
  pll_compute_gamma_cats(1.0, 4, gamma, PLL_GAMMA_RATES_MEAN);

  for (int recycled = 0; recycled < 2; ++recycled)
  {
    unsigned int buffers = inner;
    if (recycled)
    {
      if (!pll_utree_create_operations_recycled(root, tips, 64, lengths, matrices, ops, &n_mat, &n_ops,
                                                edge_clv, edge_scaler, &slots))
        die("pll_utree_create_operations_recycled");
      buffers = slots;
    }
    else
    {
      if (!pll_utree_traverse(root, PLL_TREE_TRAVERSE_POSTORDER, take_all, trav, &trav_size))
        die("pll_utree_traverse");
      pll_utree_create_operations(trav, trav_size, lengths, matrices, ops, &n_mat, &n_ops);
      edge_clv[0] = root->clv_index;
      edge_scaler[0] = root->scaler_index;
      edge_clv[1] = root->back->clv_index;
      edge_scaler[1] = root->back->scaler_index;
    }
    pll_partition_t * p = pll_partition_create(tips, buffers, 20, (unsigned int)patterns, 4, branches, 4, buffers,
                                               PLL_ATTRIB_ARCH_GPU | PLL_ATTRIB_PATTERN_TIP);
    if (!p) die("pll_partition_create");
    for (unsigned int t = 0; t < tips; ++t)
      if (!pll_set_tip_states(p, t, pll_map_aa, rows[t])) die("pll_set_tip_states");
    pll_set_pattern_weights(p, weights);

    pll_set_category_rates(p, gamma);
    set_lg4(p, pll_aa_rates_lg4m, pll_aa_freqs_lg4m);
    pll_update_prob_matrices(p, params, matrices, lengths, n_mat);
    pll_update_partials(p, ops, n_ops);
    double logl = pll_compute_edge_loglikelihood(p, edge_clv[0], edge_scaler[0], edge_clv[1], edge_scaler[1],
                                                 root->pmatrix_index, params, NULL);
    printf("[%s, %u CLV buffers] Log-L (LG4M): %.6f\n", recycled ? "recycled" : "plain", buffers, logl);

    set_lg4(p, pll_aa_rates_lg4x, pll_aa_freqs_lg4x);
    pll_set_category_rates(p, lg4x_rates);
    pll_set_category_weights(p, lg4x_weights);
    pll_update_prob_matrices(p, params, matrices, lengths, n_mat);
    pll_update_partials(p, ops, n_ops);
    logl = pll_compute_edge_loglikelihood(p, edge_clv[0], edge_scaler[0], edge_clv[1], edge_scaler[1],
                                          root->pmatrix_index, params, NULL);
    printf("[%s, %u CLV buffers] Log-L (LG4X): %.6f\n", recycled ? "recycled" : "plain", buffers, logl);
    pll_partition_destroy(p);
  }

  for (unsigned int t = 0; t < tips; ++t) free(rows[t]);
  free(rows);
  free(weights);
  free(trav);
  free(lengths);
  free(matrices);
  free(ops);
  pll_utree_destroy(tree, NULL);
  return 0;
}
