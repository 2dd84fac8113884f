/*
 * pll_phylip.c - PHYLIP alignment reader, sequential and interleaved (reference
 * src/phylip.c:249-730), the other on-disk alignment format on the input side of the path
 * (SURVEY.md row f3).  Same interface, results and error codes as the reference; written
 * independently: the file is read once into memory and walked with a cursor, instead of the
 * reference's growing fgets() line buffer.
 *
 * Format rules reproduced from the reference:
 *   header      two positive integers (sequences, sites) and nothing but blanks after them
 *   label       first token of a sequence's first line: it ends at the first ' ' if the line
 *               has one, else at the first TAB, else at CR / end of line
 *   data        every character goes through the 256-entry status table given to
 *               pll_phylip_open: 0 stripped (and counted), 1 kept, 2 fatal, 3 silently dropped
 *   sequential  a sequence continues over as many lines as it needs
 *   interleaved every block carries the same number of characters for every sequence; data
 *               may start on the line after the label; blank lines separate blocks
 */
#include "pll_host.h"

static int is_blank(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }

/* loads the whole file into fd->line; fd->line_size is the read cursor */
static int load(pll_phylip_t * fd)
{
  rewind(fd->fp);
  if (!fd->line)
  {
    fd->line_maxsize = (size_t)fd->filesize + 1;
    if (!(fd->line = (char *)malloc(fd->line_maxsize)))
      return pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
  }
  const size_t got = fread(fd->line, 1, (size_t)fd->filesize, fd->fp);
  fd->line[got] = 0;
  fd->filesize = (long)got;
  fd->line_size = 0;
  fd->lineno = 0;
  fd->no = -1;
  fd->stripped_count = 0;
  memset(fd->stripped, 0, sizeof(fd->stripped));
  return 1;
}

/* next line, 0-terminated in place (the '\n' is consumed); NULL at end of input */
static char * next_line(pll_phylip_t * fd)
{
  if ((long)fd->line_size >= fd->filesize) return NULL;
  char * start = fd->line + fd->line_size;
  char * nl = memchr(start, '\n', (size_t)fd->filesize - fd->line_size);
  if (nl)
  {
    *nl = 0;
    fd->line_size = (size_t)(nl - fd->line) + 1;
  }
  else
    fd->line_size = (size_t)fd->filesize;
  fd->lineno++;
  return start;
}

PLL_EXPORT pll_phylip_t * pll_phylip_open(const char * filename, const unsigned int * map)
{
  pll_phylip_t * fd = (pll_phylip_t *)calloc(1, sizeof(pll_phylip_t));
  if (!fd)
  {
    pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  fd->chrstatus = map;
  if (!(fd->fp = fopen(filename, "r")))
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
  if (fd->filesize <= 0 || !load(fd))
  {
    /* an empty file has no header line (the reference returns NULL here as well) */
    fclose(fd->fp);
    free(fd->line);
    free(fd);
    return NULL;
  }
  return fd;
}

PLL_EXPORT int pll_phylip_rewind(pll_phylip_t * fd)
{
  fseek(fd->fp, 0, SEEK_END);
  fd->filesize = ftell(fd->fp);
  if ((size_t)fd->filesize + 1 > fd->line_maxsize)
  {
    free(fd->line);
    fd->line = NULL;
  }
  if (fd->filesize <= 0 || !load(fd)) return pll_fail(PLL_ERROR_FILE_SEEK, "Unable to rewind and cache data");
  return PLL_SUCCESS;
}

PLL_EXPORT void pll_phylip_close(pll_phylip_t * fd)
{
  fclose(fd->fp);
  free(fd->line);
  free(fd);
}

PLL_EXPORT void pll_msa_destroy(pll_msa_t * msa)
{
  if (!msa) return;
  for (int i = 0; i < msa->count; ++i)
  {
    if (msa->label) free(msa->label[i]);
    if (msa->sequence) free(msa->sequence[i]);
  }
  free(msa->label);
  free(msa->sequence);
  free(msa);
}

/* header line -> an msa with empty, 0-terminated sequences */
static pll_msa_t * new_msa(pll_phylip_t * fd)
{
  char * header = next_line(fd);
  int count = 0, length = 0, used = 0;
  if (!header || sscanf(header, "%d%n", &count, &used) < 1 || !used || !count)
  {
    pll_fail(PLL_ERROR_PHYLIP_SYNTAX, "Invalid number of sequences in header");
    return NULL;
  }
  header += used;
  used = 0;
  if (sscanf(header, "%d%n", &length, &used) < 1 || !used || !length)
  {
    pll_fail(PLL_ERROR_PHYLIP_SYNTAX, "Invalid sequence length in header");
    return NULL;
  }
  header += used;
  while (*header && is_blank(*header)) ++header;
  if (*header || count < 0 || length < 0)
  {
    pll_fail(PLL_ERROR_PHYLIP_SYNTAX, "Unexpected characters after the dimensions in the header");
    return NULL;
  }
  pll_msa_t * msa = (pll_msa_t *)calloc(1, sizeof(pll_msa_t));
  if (msa)
  {
    msa->count = count;
    msa->length = length;
    msa->sequence = (char **)calloc((size_t)count, sizeof(char *));
    msa->label = (char **)calloc((size_t)count, sizeof(char *));
  }
  int ok = msa && msa->sequence && msa->label;
  for (int i = 0; ok && i < count; ++i)
  {
    ok = (msa->sequence[i] = (char *)malloc((size_t)length + 1)) != NULL;
    if (ok) msa->sequence[i][length] = 0;
  }
  if (!ok)
  {
    pll_msa_destroy(msa);
    pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  return msa;
}

/* appends the legal characters of `text` to sequence `seqno` at `offset`; returns how many were
 * appended, or -1 with pll_errno set */
static int take_data(pll_phylip_t * fd, pll_msa_t * msa, const char * text, int seqno, int offset)
{
  int n = 0;
  for (; *text; ++text)
  {
    const unsigned char c = (unsigned char)*text;
    switch (fd->chrstatus[c])
    {
      case 0:
        fd->stripped_count++;
        fd->stripped[c]++;
        break;
      case 1:
        if (offset + n >= msa->length)
        {
          pll_fail(PLL_ERROR_PHYLIP_LONGSEQ, "Sequence %d (%.100s) longer than expected", seqno + 1,
                   msa->label[seqno]);
          return -1;
        }
        msa->sequence[seqno][offset + n++] = (char)c;
        break;
      case 2:
        if (c >= 32)
          pll_fail(PLL_ERROR_PHYLIP_ILLEGALCHAR, "illegal character '%c' on line %ld in the phylip file", c,
                   fd->lineno);
        else
          pll_fail(PLL_ERROR_PHYLIP_UNPRINTABLECHAR,
                   "illegal unprintable character %#.2x (hexadecimal) on line %ld in the phylip file", c,
                   fd->lineno);
        return -1;
      default: break;
    }
  }
  return n;
}

/* cuts the label off the front of a sequence's first line; returns the rest of the line */
static char * take_label(pll_msa_t * msa, int seqno, char * p)
{
  size_t n;
  if (strchr(p, ' ')) n = strcspn(p, " ");
  else if (strchr(p, '\t')) n = strcspn(p, "\t");
  else n = strcspn(p, "\r");
  if (!(msa->label[seqno] = (char *)malloc(n + 1)))
  {
    pll_fail(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  memcpy(msa->label[seqno], p, n);
  msa->label[seqno][n] = 0;
  return p + n;
}

/* next non-blank line with leading blanks skipped; NULL at end of input */
static char * next_content_line(pll_phylip_t * fd)
{
  char * p;
  while ((p = next_line(fd)))
  {
    while (*p && is_blank(*p)) ++p;
    if (*p) return p;
  }
  return NULL;
}

PLL_EXPORT pll_msa_t * pll_phylip_parse_sequential(pll_phylip_t * fd)
{
  pll_msa_t * msa = new_msa(fd);
  if (!msa) return NULL;
  int seqno = 0;
  char * p;
  while ((p = next_content_line(fd)))
  {
    if (seqno == msa->count)
    {
      pll_fail(PLL_ERROR_PHYLIP_SYNTAX, "Found at least %d sequences but expected %d", seqno + 1, msa->count);
      goto fail;
    }
    if (!(p = take_label(msa, seqno, p))) goto fail;
    int have = 0;
    for (;;)
    {
      const int n = take_data(fd, msa, p, seqno, have);
      if (n < 0) goto fail;
      have += n;
      if (have == msa->length) break;
      if (!(p = next_line(fd)))
      {
        pll_fail(PLL_ERROR_PHYLIP_SYNTAX, "Sequence %d (%.100s) has %d characters but expected %d", seqno + 1,
                 msa->label[seqno], have, msa->length);
        goto fail;
      }
    }
    ++seqno;
  }
  if (seqno != msa->count)
  {
    pll_fail(PLL_ERROR_PHYLIP_SYNTAX, "Found %d sequence(s) but expected %d", seqno, msa->count);
    goto fail;
  }
  return msa;
fail:
  pll_msa_destroy(msa);
  return NULL;
}

/* one sequence's share of a block: the first line from `p` on that carries data.  Returns 1 and
 * sets *got, 0 at end of input (then *got == 0), -1 on error. */
static int take_block_line(pll_phylip_t * fd, pll_msa_t * msa, char * p, int seqno, int offset, int * got)
{
  *got = 0;
  while (p)
  {
    const int n = take_data(fd, msa, p, seqno, offset);
    if (n < 0) return -1;
    if (n)
    {
      *got = n;
      return 1;
    }
    p = next_line(fd);
  }
  return 0;
}

PLL_EXPORT pll_msa_t * pll_phylip_parse_interleaved(pll_phylip_t * fd)
{
  pll_msa_t * msa = new_msa(fd);
  if (!msa) return NULL;
  int seqno = 0, block_len = 0, total = 0, block = 1, got = 0, rc = 1;
  char * p;

  /* first block: label + data per sequence */
  while (seqno < msa->count && (p = next_content_line(fd)))
  {
    if (!(p = take_label(msa, seqno, p))) goto fail;
    rc = take_block_line(fd, msa, p, seqno, 0, &got);
    if (rc < 0) goto fail;
    if (rc == 0) break;
    if (block_len && got != block_len) goto misaligned;
    block_len = got;
    ++seqno;
  }
  if (seqno != msa->count)
  {
    pll_fail(PLL_ERROR_PHYLIP_SYNTAX, "Found %d sequence(s) but expected %d", seqno, msa->count);
    goto fail;
  }
  total = block_len;

  /* the remaining blocks: data only, sequences in the same order */
  seqno = 0;
  block_len = 0;
  block = 2;
  for (;;)
  {
    rc = take_block_line(fd, msa, next_line(fd), seqno, total, &got);
    if (rc < 0) goto fail;
    if (rc == 0) break;
    if (block_len && got != block_len) goto misaligned;
    block_len = got;
    if (++seqno == msa->count)
    {
      seqno = 0;
      total += block_len;
      block_len = 0;
      ++block;
    }
  }
  if (seqno)
  {
    pll_fail(PLL_ERROR_PHYLIP_SYNTAX, "Found %d sequences in block %d but expected %d", seqno, block, msa->count);
    goto fail;
  }
  if (total != msa->length)
  {
    pll_fail(PLL_ERROR_PHYLIP_SYNTAX, "Sequence length is %d but expected %d", total, msa->length);
    goto fail;
  }
  return msa;

misaligned:
  pll_fail(PLL_ERROR_PHYLIP_NONALIGNED, "Sequence %d (%.100s) data out of alignment", seqno + 1, msa->label[seqno]);
fail:
  pll_msa_destroy(msa);
  return NULL;
}
