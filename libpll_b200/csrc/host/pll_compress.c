/*
 * pll_compress.c - site-pattern compression (host).
 *
 * Mirrors the interface and the exact output of reference src/compress.c:138-286
 * (pll_compress_site_patterns): the alignment columns are encoded through the state map,
 * sorted, made unique and written back over the input sequences IN SORTED ORDER together with
 * the multiplicity of every unique column; the sequences are decoded back to characters with
 * the inverse map, in which the LAST ASCII character mapping to a code wins (reference
 * :173-175 - upper-case input comes back lower-case for the nucleotide map).
 *
 * The reference sorts with a randomised multikey quicksort (:33-81) whose comparisons are on
 * `char` (signed on x86-64), over 0-terminated column strings.  Because the order of distinct
 * columns under that comparison is total, any correct sort yields the same output; this
 * implementation sorts column indices with a most-significant-byte-first three-way radix
 * quicksort written independently, using the same signed-byte order and the same
 * "a zero byte ends the column" rule.
 */
#include "pll_host.h"

/* encoded alignment, column-major: column i occupies [i*(count+1), (i+1)*(count+1)) and is
 * 0-terminated */
typedef struct
{
  const signed char * data;
  size_t stride;
} columns_t;

static inline int byte_at(const columns_t * c, unsigned int col, size_t depth)
{
  return (int)c->data[(size_t)col * c->stride + depth];
}

static void swap_u(unsigned int * a, unsigned int * b)
{
  unsigned int t = *a;
  *a = *b;
  *b = t;
}

/* three-way partition on the byte at `depth` (Bentley-Sedgewick), iterative on the middle
 * part to bound recursion depth by the number of distinct prefixes */
static void sort_columns(const columns_t * c, unsigned int * idx, size_t n, size_t depth)
{
  while (n > 1)
  {
    /* median of three as pivot byte */
    const int p0 = byte_at(c, idx[0], depth), p1 = byte_at(c, idx[n / 2], depth),
              p2 = byte_at(c, idx[n - 1], depth);
    const int pivot = (p0 < p1) ? ((p1 < p2) ? p1 : (p0 < p2 ? p2 : p0))
                                : ((p0 < p2) ? p0 : (p1 < p2 ? p2 : p1));
    size_t lt = 0, i = 0, gt = n;
    while (i < gt)
    {
      const int b = byte_at(c, idx[i], depth);
      if (b < pivot) swap_u(&idx[lt++], &idx[i++]);
      else if (b > pivot) swap_u(&idx[i], &idx[--gt]);
      else ++i;
    }
    sort_columns(c, idx, lt, depth);
    sort_columns(c, idx + gt, n - gt, depth);
    /* equal part: next byte, unless the columns ended here */
    if (pivot == 0) return;
    idx += lt;
    n = gt - lt;
    ++depth;
  }
}

static int same_column(const columns_t * c, unsigned int a, unsigned int b)
{
  return strcmp((const char *)c->data + (size_t)a * c->stride,
                (const char *)c->data + (size_t)b * c->stride) == 0;
}

/* code table (state map squeezed into a byte) and its inverse, as the reference builds them */
static void build_tables(const unsigned int * map, unsigned char * charmap, unsigned char * inv_charmap)
{
  unsigned int i, maxcode = 0;
  /* states that do not fit a byte are renumbered 1..k in order of first appearance
   * (reference src/compress.c:83-108, 161-169) */
  for (i = 0; i < PLL_ASCII_SIZE; ++i)
    if (map[i] > maxcode) maxcode = map[i];
  if (maxcode >= PLL_ASCII_SIZE)
  {
    unsigned char k = 1;
    unsigned int seen[PLL_ASCII_SIZE];
    unsigned int nseen = 0, s;
    memset(charmap, 0, PLL_ASCII_SIZE);
    for (i = 0; i < PLL_ASCII_SIZE; ++i)
    {
      if (!map[i]) continue;
      for (s = 0; s < nseen; ++s)
        if (map[seen[s]] == map[i]) break;
      if (s == nseen)
      {
        seen[nseen++] = i;
        charmap[i] = k++;
      }
      else
        charmap[i] = charmap[seen[s]];
    }
  }
  else
    for (i = 0; i < PLL_ASCII_SIZE; ++i) charmap[i] = (unsigned char)map[i];

  memset(inv_charmap, 0, PLL_ASCII_SIZE);
  for (i = 0; i < PLL_ASCII_SIZE; ++i)
    if (map[i]) inv_charmap[charmap[i]] = (unsigned char)i;
}

PLL_EXPORT unsigned int * pll_compress_site_patterns(char ** sequence,
                                                     const unsigned int * map,
                                                     int count,
                                                     int * length)
{
  unsigned char charmap[PLL_ASCII_SIZE];
  unsigned char inv_charmap[PLL_ASCII_SIZE];
  unsigned int i;
  int j;

  if (!count || !map || map[0]) return NULL;
  const size_t len = (size_t)*length;
  if (!len) return NULL;

  build_tables(map, charmap, inv_charmap);

  const size_t stride = (size_t)count + 1;
  signed char * data = (signed char *)malloc(len * stride);
  unsigned int * idx = (unsigned int *)malloc(len * sizeof(unsigned int));
  unsigned int * weight = (unsigned int *)malloc(len * sizeof(unsigned int));
  if (!data || !idx || !weight)
  {
    free(data);
    free(idx);
    free(weight);
    pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate space for matrix columns.");
    return NULL;
  }

  /* encode and transpose: one 0-terminated string per alignment column */
  for (j = 0; j < count; ++j)
  {
    const unsigned char * row = (const unsigned char *)sequence[j];
    for (i = 0; i < len; ++i) data[(size_t)i * stride + j] = (signed char)charmap[row[i]];
  }
  for (i = 0; i < len; ++i)
  {
    data[(size_t)i * stride + count] = 0;
    idx[i] = i;
  }

  columns_t cols = {data, stride};
  sort_columns(&cols, idx, len, 0);

  /* unique columns, in sorted order, with multiplicities; decoded straight back */
  size_t unique = 0;
  for (i = 0; i < len; ++i)
  {
    if (i && same_column(&cols, idx[i], idx[i - 1]))
    {
      weight[unique - 1]++;
      continue;
    }
    weight[unique] = 1;
    for (j = 0; j < count; ++j)
      sequence[j][unique] = (char)inv_charmap[(unsigned char)data[(size_t)idx[i] * stride + j]];
    ++unique;
  }
  for (j = 0; j < count; ++j) sequence[j][unique] = 0;

  free(data);
  free(idx);
  unsigned int * shrunk = (unsigned int *)realloc(weight, unique * sizeof(unsigned int));
  *length = (int)unique;
  return shrunk ? shrunk : weight;
}

/* NEW: the same function on the device (gpu/plg_compress.cu).  Identical output - unique columns
 * in the reference's sorted order, decoded through the inverse map, 0-terminated rows, weights
 * in a malloc'ed array the caller frees - for alignments where the host sort is the set-up
 * bottleneck (10 M columns). */
PLL_EXPORT unsigned int * pll_gpu_compress_site_patterns(char ** sequence,
                                                         const unsigned int * map,
                                                         int count,
                                                         int * length)
{
  unsigned char charmap[PLL_ASCII_SIZE];
  unsigned char inv_charmap[PLL_ASCII_SIZE];
  if (!count || !map || map[0] || !length || *length <= 0) return NULL;
  build_tables(map, charmap, inv_charmap);
  unsigned int * weight = (unsigned int *)malloc((size_t)*length * sizeof(unsigned int));
  if (!weight)
  {
    pll_fail(PLL_ERROR_MEM_ALLOC, "Cannot allocate space for pattern weights.");
    return NULL;
  }
  size_t unique = 0;
  int rc = plg_compress_patterns(pll_gpu_current_device(), (unsigned char * const *)sequence, (unsigned int)count,
                                 (size_t)*length, charmap, inv_charmap, weight, &unique);
  if (rc)
  {
    free(weight);
    pllg_fail(rc, "pll_gpu_compress_site_patterns");
    return NULL;
  }
  for (int j = 0; j < count; ++j) sequence[j][unique] = 0;
  unsigned int * shrunk = (unsigned int *)realloc(weight, unique * sizeof(unsigned int));
  *length = (int)unique;
  return shrunk ? shrunk : weight;
}
