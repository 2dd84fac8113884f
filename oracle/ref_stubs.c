/* TEST INFRASTRUCTURE ONLY.
 *
 * The reference keeps two plain-C helpers inside the epilogue of its bison grammar
 * (reference src/parse_utree.y:71-101 and :395-445).  bison/flex are not available in this
 * image, so the grammar cannot be generated; src/utree.c still references the two symbols.
 * The likelihood hot path never reaches them, so the oracle build satisfies the linker with
 * stubs that abort loudly if anything ever does call them.
 */
#include <stdio.h>
#include <stdlib.h>

void pll_utree_graph_destroy(void * root, void (*cb_destroy)(void *))
{
  (void)root; (void)cb_destroy;
  fprintf(stderr, "oracle/_ref: pll_utree_graph_destroy is a stub (parser not built)\n");
  abort();
}

void * pll_utree_wraptree(void * root, unsigned int tip_count)
{
  (void)root; (void)tip_count;
  fprintf(stderr, "oracle/_ref: pll_utree_wraptree is a stub (parser not built)\n");
  abort();
  return NULL;
}
