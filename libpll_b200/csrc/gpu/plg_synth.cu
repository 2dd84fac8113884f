/*
 * plg_synth.cu - synthetic DNA tip rows generated directly in HBM (the
 * 5 000-taxon x 10 M-pattern configuration cannot be fed from host-generated characters).
 *
 * Counter-based: the character of (tip, global site) is a pure function of a seed, so any
 * slice of any tip can be produced independently on any device, and the host restatement
 * (libpll_b200/synthetic.py: hash_tip_sequence) gives the CPU arm and the tests the same
 * alignment.  Recipe of SURVEY 8d: a root sequence uniform over ACGT; a tip shows the root
 * state with probability 0.7, otherwise a uniform state; 1 % of the characters become the full
 * ambiguity N, 0.5 % a two-state ambiguity (R = A|G or Y = C|T).  The row is stored as the
 * 4-bit state masks pll_set_tip_states would store (reference src/pll.c:825-860).
 */
#include "plg_internal.cuh"

__host__ __device__ static inline unsigned long long splitmix64(unsigned long long x)
{
  x += 0x9E3779B97F4A7C15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void k_generate_tip_dna(unsigned char * __restrict__ row, unsigned int sites, unsigned long long seed,
                                   unsigned long long tip_key, unsigned long long first_site)
{
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sites) return;
  const unsigned long long site = first_site + i;
  const unsigned int root = (unsigned int)(splitmix64(seed ^ (site * 0xD1342543DE82EF95ull)) & 3u);
  const unsigned long long h = splitmix64(tip_key ^ site);
  const unsigned int u = (unsigned int)(h & 0xFFFFu);
  const unsigned int alt = (unsigned int)((h >> 16) & 3u);
  const unsigned int u2 = (unsigned int)((h >> 32) & 0xFFFFu);
  unsigned int mask = 1u << ((u < 45875u) ? root : alt); /* 0.7 * 65536 */
  if (u2 < 655u) mask = 15u;                              /* 1 %: N */
  else if (u2 < 983u) mask = (alt & 1u) ? 10u : 5u;       /* 0.5 %: Y = C|T or R = A|G */
  row[i] = (unsigned char)mask;
}

extern "C" int plg_generate_tipchars(plg_context_t * ctx, unsigned int tip_index, unsigned long long seed,
                                     unsigned long long first_site)
{
  PLG_CHECK_CTX(ctx);
  if (!ctx->pattern_tip || ctx->d.states != 4 || tip_index >= ctx->d.tips)
  {
    plg_set_error("plg_generate_tipchars: needs a 4-state pattern-tip partition and a valid tip index");
    return PLG_E_INVALID;
  }
  const unsigned long long tip_key = splitmix64(seed + 0x632BE59BD9B4E019ull * (tip_index + 1ull));
  const unsigned int n = ctx->d.sites;
  k_generate_tip_dna<<<(n + 255) / 256, 256, 0, ctx->stream>>>(plg_tip_ptr(ctx, tip_index), n, seed, tip_key, first_site);
  PLG_LAUNCH_CHECK(ctx);
  return PLG_OK;
}
