/*
 * fuzz_frontend.c - mutation fuzzing of the callers'-side host code under AddressSanitizer and
 * UndefinedBehaviorSanitizer: Newick reader + tree utilities, FASTA and PHYLIP readers, site
 * pattern compression.  Built and run by tests/test_sanitizers_cpu.py from the host C
 * sources directly (no CUDA objects); the few device-layer symbols those files reference are
 * stubbed below.  Deterministic (xorshift seed on the command line).
 *
 *   fuzz_frontend <iterations> <seed> <scratch-dir>
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pll.h"
#include "pll_gpu.h"

/* ---- what pll_partition.c / the device layer would provide ---- */
__thread int pll_errno;
__thread char pll_errmsg[200];
int pll_fail(int code, const char * fmt, ...)
{
  va_list ap;
  pll_errno = code;
  va_start(ap, fmt);
  vsnprintf(pll_errmsg, sizeof(pll_errmsg), fmt, ap);
  va_end(ap);
  return PLL_FAILURE;
}
int pllg_fail(int rc, const char * where) { return pll_fail(PLL_ERROR_GPU_RUNTIME, "%s: %d", where, rc); }
int pll_gpu_current_device(void) { return -1; }
int plg_compress_patterns(int device, unsigned char * const * rows, unsigned int taxa, size_t length,
                          const unsigned char * code_table, const unsigned char * inverse_table,
                          unsigned int * weights_out, size_t * unique_out)
{
  return PLG_E_NODEVICE;
}
void * pll_aligned_alloc(size_t size, size_t alignment)
{
  void * mem = NULL;
  if (posix_memalign(&mem, alignment < sizeof(void *) ? sizeof(void *) : alignment, size ? size : 1)) mem = NULL;
  return mem;
}
void pll_aligned_free(void * ptr) { free(ptr); }

/* ---- deterministic RNG ---- */
static unsigned long long rng_state;
static unsigned int rnd(void)
{
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return (unsigned int)(rng_state >> 16);
}

static const char * const NEWICK_SEEDS[] = {
  "((a:0.1,b:0.2):0.05,c:0.3,d:0.4);",
  "(a,b,(c,d));",
  "((a:1e-3,b:2.5E+1)x:0.1,(c:0.2,d:0.3)y:0.4,(e,f)'z w':1)root;",
  "(((((a,b),c),d),e),f,g);",
  "(a:0.1,b:0.2,(c:0.3,(d:0.4,(e:0.5,(f:0.6,g:0.7):0.8):0.9):1.0):1.1);",
  "((a,b),(c,d));",            /* rooted: must be rejected or unrooted as the reference does */
  "(a,b);", "a;", "();", "((a,b,c),d,e);", "(a:0.1,b:0.2,c:0.3)",
};
static const char * const ROOTED_SEEDS[] = {
  "((a:0.1,b:0.2):0.05,(c:0.3,d:0.4):0.1);",
  "(a,b);",
  "(((a,b)x:1e-3,c)'y z':2.5E+1,(d,(e,f)));",
  "((((a,b),c),d),e)root:0.1;",
  "(a,b,c);", "((a,b),(c,d))", "a;", "();",
};
static const char * const FASTA_SEEDS[] = {
  ">s1\nACGTACGT\n>s2\nACGTTCGA\n>s3 description\nAC-TNCGT\n",
  ">a\nAC\nGT\n\n>b\nTTTT\n",
  ">only header\n",
  "ACGT\n>late\nAC\n",
  ">x\nAC*T|9\n",
};
static const char * const PHYLIP_SEEDS[] = {
  "3 8\ns1 ACGTACGT\ns2 ACGTTCGA\ns3 AC-TNCGT\n",
  " 2 12\nt1 ACGTAC\nt2 TTGTAC\n\nGTACGT\nGTACGA\n",
  "2 4\nlongname_without_space\nACGT\nb\nAC\nGT\n",
  "0 0\n", "3 4\na ACGT\nb ACG\n", "2 3 extra\na AAA\nb CCC\n",
};
static const char FUZZ_ALPHABET[] = "(),:;'\"[] \t\n>ACGTN-acgt0123456789.eE+-_*|\\\x01\xff";

static char * mutate(const char * seed, size_t * out_len)
{
  size_t n = strlen(seed);
  size_t cap = n + 64;
  char * s = (char *)malloc(cap + 1);
  memcpy(s, seed, n);
  unsigned int edits = rnd() % 6;
  for (unsigned int e = 0; e < edits; ++e)
  {
    unsigned int kind = rnd() % 5;
    size_t at = n ? rnd() % n : 0;
    char c = FUZZ_ALPHABET[rnd() % (sizeof(FUZZ_ALPHABET) - 1)];
    if (kind == 0 && n) s[at] = c;                                  /* replace */
    else if (kind == 1 && n + 1 < cap) { memmove(s + at + 1, s + at, n - at); s[at] = c; ++n; } /* insert */
    else if (kind == 2 && n) { memmove(s + at, s + at + 1, n - at - 1); --n; }                 /* delete */
    else if (kind == 3 && n) n = at;                                /* truncate */
    else if (kind == 4 && n)                                        /* duplicate a chunk */
    {
      size_t len = 1 + rnd() % 8;
      if (at + len > n) len = n - at;
      if (n + len < cap) { memmove(s + at + len, s + at, n - at); n += len; }
    }
  }
  s[n] = 0;
  *out_len = n;
  return s;
}

static void write_file(const char * path, const char * data, size_t len)
{
  FILE * f = fopen(path, "wb");
  if (!f) { perror(path); exit(2); }
  fwrite(data, 1, len, f);
  fclose(f);
}

static int visit_all(pll_unode_t * node) { return node != NULL; }
static int count_cb(pll_unode_t * node) { return node ? 1 : 0; }
/* labels are exported verbatim (as the reference does, src/utree.c:217-282), so a tree only has
 * to parse back when no label needs quoting */
static int plain_labels;
static int label_cb(pll_unode_t * node)
{
  if (node->label && strpbrk(node->label, "(),:;'\"[] \t\n\r"))
    plain_labels = 0;
  if (node->label && !*node->label) plain_labels = 0;
  return 1;
}

static int rvisit_all(pll_rnode_t * node) { return node != NULL; }

static unsigned long exercise_rooted(pll_rtree_t * tree)
{
  const unsigned int T = tree->tip_count, I = tree->inner_count;
  if (I != T - 1 || tree->edge_count != 2 * T - 2 || tree->nodes[T + I - 1] != tree->root)
  {
    fprintf(stderr, "rooted tree: inconsistent counts\n");
    exit(3);
  }
  pll_rnode_t ** buf = (pll_rnode_t **)malloc((size_t)(T + I) * sizeof(*buf));
  double * br = (double *)malloc((size_t)(T + I) * sizeof(double));
  unsigned int * mi = (unsigned int *)malloc((size_t)(T + I) * sizeof(unsigned int));
  pll_operation_t * ops = (pll_operation_t *)malloc((size_t)I * sizeof(*ops));
  unsigned int n = 0, mc = 0, oc = 0;
  for (int order = PLL_TREE_TRAVERSE_POSTORDER; order <= PLL_TREE_TRAVERSE_PREORDER; ++order)
  {
    if (!pll_rtree_traverse(tree->root, order, rvisit_all, buf, &n) || n != T + I)
    {
      fprintf(stderr, "rooted traversal visits %u of %u nodes\n", n, T + I);
      exit(3);
    }
  }
  pll_rtree_traverse(tree->root, PLL_TREE_TRAVERSE_POSTORDER, rvisit_all, buf, &n);
  pll_rtree_create_operations(buf, n, br, mi, ops, &mc, &oc);
  if (oc != I || mc != 2 * T - 2) { fprintf(stderr, "rooted operations: %u ops, %u matrices\n", oc, mc); exit(3); }
  free(buf); free(br); free(mi); free(ops);
  char * text = pll_rtree_export_newick(tree->root, NULL);
  free(text);
  return oc;
}

static unsigned long exercise_tree(pll_utree_t * tree)
{
  unsigned long work = 0;
  const unsigned int T = tree->tip_count, I = tree->inner_count;
  if (!pll_utree_check_integrity(tree)) { fprintf(stderr, "parsed tree fails its integrity check\n"); exit(3); }
  pll_unode_t * root = tree->nodes[T + I - 1];
  pll_unode_t ** buf = (pll_unode_t **)malloc((size_t)(T + I) * sizeof(*buf));
  unsigned int n = 0;
  if (pll_utree_traverse(root, PLL_TREE_TRAVERSE_POSTORDER, visit_all, buf, &n))
  {
    double * br = (double *)malloc((size_t)(2 * T) * sizeof(double));
    unsigned int * mi = (unsigned int *)malloc((size_t)(2 * T) * sizeof(unsigned int));
    pll_operation_t * ops = (pll_operation_t *)malloc((size_t)(I + 1) * sizeof(*ops));
    unsigned int mc = 0, oc = 0;
    pll_utree_create_operations(buf, n, br, mi, ops, &mc, &oc);
    if (oc != I || mc != 2 * T - 3) { fprintf(stderr, "full traversal: %u ops, %u matrices for %u tips\n", oc, mc, T); exit(3); }
    unsigned int eclv[2], used = 0;
    int escal[2];
    if (pll_utree_create_operations_recycled(root, T, T, br, mi, ops, &mc, &oc, eclv, escal, &used))
      if (oc != I || used > I) { fprintf(stderr, "recycled traversal inconsistent\n"); exit(3); }
    work += oc;
    free(br); free(mi); free(ops);
  }
  free(buf);
  char * text = pll_utree_export_newick(root, NULL);
  plain_labels = 1;
  pll_utree_every(tree, label_cb);
  if (text && !plain_labels) { free(text); text = NULL; }
  if (text)
  {
    pll_utree_t * again = pll_utree_parse_newick_string(text);
    if (!again || again->tip_count != T) { fprintf(stderr, "exported tree does not parse back: %s\n", text); exit(3); }
    pll_utree_destroy(again, NULL);
    free(text);
  }
  pll_utree_t * copy = pll_utree_clone(tree);
  if (copy)
  {
    pll_utree_every(copy, count_cb);
    pll_utree_destroy(copy, NULL);
  }
  return work;
}

int main(int argc, char ** argv)
{
  if (argc < 4) { fprintf(stderr, "usage: %s iterations seed scratch-dir\n", argv[0]); return 2; }
  const long iterations = atol(argv[1]);
  rng_state = 0x9E3779B97F4A7C15ull ^ (unsigned long long)atoll(argv[2]);
  char path[4096];
  snprintf(path, sizeof(path), "%s/fuzz_input.txt", argv[3]);
  unsigned long trees = 0, rooted_trees = 0, records = 0, alignments = 0, compressed = 0, work = 0;

  for (long it = 0; it < iterations; ++it)
  {
    size_t len;
    /* Newick, from memory and (every 8th) from a file */
    char * s = mutate(NEWICK_SEEDS[rnd() % (sizeof(NEWICK_SEEDS) / sizeof(*NEWICK_SEEDS))], &len);
    pll_utree_t * tree;
    if (it % 8 == 0) { write_file(path, s, len); tree = pll_utree_parse_newick(path); }
    else tree = pll_utree_parse_newick_string(s);
    if (tree) { ++trees; work += exercise_tree(tree); pll_utree_destroy(tree, NULL); }
    free(s);

    /* rooted Newick */
    if (it % 2 == 1)
    {
      s = mutate(ROOTED_SEEDS[rnd() % (sizeof(ROOTED_SEEDS) / sizeof(*ROOTED_SEEDS))], &len);
      pll_rtree_t * rooted;
      if (it % 16 == 1) { write_file(path, s, len); rooted = pll_rtree_parse_newick(path); }
      else rooted = pll_rtree_parse_newick_string(s);
      if (rooted) { ++rooted_trees; work += exercise_rooted(rooted); pll_rtree_destroy(rooted, NULL); }
      free(s);
    }

    /* FASTA */
    if (it % 4 == 1)
    {
      s = mutate(FASTA_SEEDS[rnd() % (sizeof(FASTA_SEEDS) / sizeof(*FASTA_SEEDS))], &len);
      write_file(path, s, len);
      pll_fasta_t * fd = pll_fasta_open(path, pll_map_fasta);
      if (fd)
      {
        char * head = NULL, * seq = NULL;
        long hl, sl, no;
        while (pll_fasta_getnext(fd, &head, &hl, &seq, &sl, &no))
        {
          if ((long)strlen(head) != hl || (long)strlen(seq) != sl) { fprintf(stderr, "fasta lengths disagree\n"); exit(3); }
          ++records; free(head); free(seq);
        }
        pll_fasta_getfilesize(fd); pll_fasta_getfilepos(fd); pll_fasta_rewind(fd);
        pll_fasta_close(fd);
      }
      free(s);
    }

    /* PHYLIP, both flavours */
    if (it % 4 == 2)
    {
      s = mutate(PHYLIP_SEEDS[rnd() % (sizeof(PHYLIP_SEEDS) / sizeof(*PHYLIP_SEEDS))], &len);
      write_file(path, s, len);
      for (int flavour = 0; flavour < 2; ++flavour)
      {
        pll_phylip_t * fd = pll_phylip_open(path, pll_map_phylip);
        if (!fd) continue;
        pll_msa_t * msa = flavour ? pll_phylip_parse_interleaved(fd) : pll_phylip_parse_sequential(fd);
        if (msa)
        {
          for (int i = 0; i < msa->count; ++i)
            if ((int)strlen(msa->sequence[i]) != msa->length) { fprintf(stderr, "phylip row length\n"); exit(3); }
          /* compress what was read */
          int length = msa->length;
          if (msa->count > 0 && length > 0)
          {
            unsigned int * w = pll_compress_site_patterns(msa->sequence, (rnd() & 1) ? pll_map_nt : pll_map_aa,
                                                          msa->count, &length);
            if (w)
            {
              unsigned long total = 0;
              for (int i = 0; i < length; ++i) total += w[i];
              if (total != (unsigned long)msa->length) { fprintf(stderr, "compression lost columns\n"); exit(3); }
              ++compressed; free(w);
            }
          }
          ++alignments;
          pll_msa_destroy(msa);
        }
        pll_phylip_close(fd);
      }
      free(s);
    }

    /* compression of random columns with many duplicates */
    if (it % 16 == 3)
    {
      int taxa = 1 + (int)(rnd() % 9), length = 1 + (int)(rnd() % 200), orig = length;
      char ** rows = (char **)malloc((size_t)taxa * sizeof(char *));
      for (int t = 0; t < taxa; ++t)
      {
        rows[t] = (char *)malloc((size_t)length + 1);
        for (int i = 0; i < length; ++i) rows[t][i] = "ACGT-NRY"[rnd() % ((rnd() & 3) ? 2 : 8)];
        rows[t][length] = 0;
      }
      unsigned int * w = pll_compress_site_patterns(rows, pll_map_nt, taxa, &length);
      if (!w) { fprintf(stderr, "compression failed: %s\n", pll_errmsg); exit(3); }
      unsigned long total = 0;
      for (int i = 0; i < length; ++i) total += w[i];
      if (total != (unsigned long)orig) { fprintf(stderr, "compression lost columns\n"); exit(3); }
      ++compressed; free(w);
      for (int t = 0; t < taxa; ++t) free(rows[t]);
      free(rows);
    }
  }
  remove(path);
  printf("iterations=%ld trees=%lu fasta_records=%lu alignments=%lu compressed=%lu rooted=%lu work=%lu\n", iterations,
         trees, records, alignments, compressed, rooted_trees, work);
  return 0;
}
