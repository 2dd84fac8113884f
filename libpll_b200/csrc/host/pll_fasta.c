/*
 * pll_fasta.c - FASTA reader (reference src/fasta.c:39-323), the on-disk format on the input
 * side of the likelihood path.
 *
 * Same interface and observable behaviour as the reference: records are returned one at a
 * time in freshly malloc'ed buffers the caller frees; every sequence character is classified
 * through the 256-entry status table given to pll_fasta_open (0 = stripped and counted,
 * 1 = kept, 2 = fatal, 3 = silently stripped); end of input is reported as PLL_FAILURE with
 * pll_errno == PLL_ERROR_FILE_EOF.
 */
#include "pll_host.h"

/* status table for FASTA / PHYLIP payload characters (reference src/maps.c:117-168): tab..CR
 * silently dropped, other control characters and '.' fatal, letters (except the two lower-case
 * letters j and o, as in the reference), digits, '-' and '?' kept, everything else stripped */
#define FILE_STATUS_TABLE                                                                 \
  {                                                                                       \
    [0 ... 8] = 2, [9 ... 13] = 3, [14 ... 31] = 2, ['-'] = 1, ['.'] = 2, ['0' ... '9'] = 1, \
    ['?'] = 1, ['A' ... 'Z'] = 1, ['a' ... 'i'] = 1, ['k' ... 'n'] = 1, ['p' ... 'z'] = 1,  \
  }
PLL_EXPORT const unsigned int pll_map_fasta[256] = FILE_STATUS_TABLE;
PLL_EXPORT const unsigned int pll_map_phylip[256] = FILE_STATUS_TABLE;

static int cache_line(pll_fasta_t * fd)
{
  fd->line[0] = 0;
  return fgets(fd->line, PLL_LINEALLOC, fd->fp) != NULL;
}

static void reset_counts(pll_fasta_t * fd)
{
  fd->stripped_count = 0;
  memset(fd->stripped, 0, sizeof(fd->stripped));
}

PLL_EXPORT pll_fasta_t * pll_fasta_open(const char * filename, const unsigned int * map)
{
  pll_fasta_t * fd = (pll_fasta_t *)malloc(sizeof(pll_fasta_t));
  if (!fd)
  {
    pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  fd->lineno = 0;
  fd->no = -1;
  fd->chrstatus = map;
  fd->fp = fopen(filename, "r");
  if (!fd->fp)
  {
    pll_fail(PLL_ERROR_FILE_OPEN, "Unable to open file (%s)", filename);
    free(fd);
    return NULL;
  }
  if (fseek(fd->fp, 0, SEEK_END))
  {
    pll_fail(PLL_ERROR_FILE_SEEK, "Unable to seek in file (%s)", filename);
    fclose(fd->fp);
    free(fd);
    return NULL;
  }
  fd->filesize = ftell(fd->fp);
  rewind(fd->fp);
  reset_counts(fd);
  if (!cache_line(fd))
  {
    pll_fail(PLL_ERROR_FILE_SEEK, "Unable to read file (%s)", filename);
    fclose(fd->fp);
    free(fd);
    return NULL;
  }
  fd->lineno = 1;
  return fd;
}

PLL_EXPORT int pll_fasta_rewind(pll_fasta_t * fd)
{
  rewind(fd->fp);
  reset_counts(fd);
  if (!cache_line(fd)) return pll_fail(PLL_ERROR_FILE_SEEK, "Unable to rewind and cache data");
  fd->lineno = 1;
  return PLL_SUCCESS;
}

PLL_EXPORT void pll_fasta_close(pll_fasta_t * fd)
{
  fclose(fd->fp);
  free(fd);
}

PLL_EXPORT long pll_fasta_getfilesize(const pll_fasta_t * fd) { return fd->filesize; }
PLL_EXPORT long pll_fasta_getfilepos(pll_fasta_t * fd) { return ftell(fd->fp); }

PLL_EXPORT int pll_fasta_getnext(pll_fasta_t * fd, char ** head, long * head_len, char ** seq,
                                 long * seq_len, long * seqno)
{
  *head_len = 0;
  *seq_len = 0;
  if (!fd->line[0])
  {
    *head = *seq = NULL;
    return pll_fail(PLL_ERROR_FILE_EOF, "End of file\n");
  }
  if (fd->line[0] != '>')
  {
    *head = *seq = NULL;
    return pll_fail(PLL_ERROR_FASTA_INVALIDHEADER, "Illegal header line in query fasta file");
  }
  /* header: the rest of the cached line up to CR or LF */
  const char * h = fd->line + 1;
  size_t hl = strcspn(h, strchr(h, '\r') ? "\r" : "\n");
  size_t cap = 4096, len = 0;
  char * header = (char *)malloc(hl + 1 > cap ? hl + 1 : cap);
  char * data = (char *)malloc(cap);
  if (!header || !data)
  {
    free(header);
    free(data);
    *head = *seq = NULL;
    return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
  }
  memcpy(header, h, hl);
  header[hl] = 0;
  cache_line(fd);
  fd->lineno++;

  while (fd->line[0] && fd->line[0] != '>')
  {
    for (const char * p = fd->line; *p; ++p)
    {
      const unsigned char c = (unsigned char)*p;
      switch (fd->chrstatus[c])
      {
        case 0:
          fd->stripped_count++;
          fd->stripped[c]++;
          break;
        case 1:
          if (len + 2 > cap)
          {
            char * grown = (char *)realloc(data, cap += 4096);
            if (!grown)
            {
              free(header);
              free(data);
              *head = *seq = NULL;
              return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
            }
            data = grown;
          }
          data[len++] = (char)c;
          break;
        case 2:
          free(header);
          free(data);
          *head = *seq = NULL;
          if (c >= 32)
            return pll_fail(PLL_ERROR_FASTA_ILLEGALCHAR,
                            "illegal character '%c' on line %ld in the fasta file", c, fd->lineno);
          return pll_fail(PLL_ERROR_FASTA_UNPRINTABLECHAR,
                          "illegal unprintable character %#.2x (hexadecimal) on line %ld in the fasta file",
                          c, fd->lineno);
        default: break; /* 3: silently stripped */
      }
    }
    cache_line(fd);
    fd->lineno++;
  }
  data[len] = 0;
  *head = header;
  *head_len = (long)hl;
  *seq = data;
  *seq_len = (long)len;
  *seqno = ++fd->no;
  return PLL_SUCCESS;
}
