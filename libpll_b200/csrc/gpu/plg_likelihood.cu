/*
 * plg_likelihood.cu - per-site log-likelihood at an edge or at a root CLV and its
 * pattern-weighted sum.
 *
 * Replaces (the AVX2-flag rungs are the parity spec):
 *   pll_core_edge_loglikelihood_ii  4x4: reference src/core_likelihood_avx.c:1079-1266
 *                                   gen: reference src/core_likelihood_avx2.c:333-547
 *   pll_core_edge_loglikelihood_ti  4x4: reference src/core_likelihood_avx.c:191-406
 *                                 20x20: reference src/core_likelihood_avx2.c:111-331
 *   pll_core_root_loglikelihood     4x4: reference src/core_likelihood_avx.c:113-189
 *                                   gen: reference src/core_likelihood_avx2.c:25-109
 *
 * Mapping: one thread per (site, rate); the R lanes of a site are adjacent, lane 0 of the
 * group collects the per-rate terms by shuffle and adds them in rate order exactly like the
 * reference's scalar loop, takes one log per site and applies scalers and pattern weight.
 * The sum over sites is a fixed-shape tree: block partials (deterministic shuffle tree) are
 * written to HBM and the last block to finish adds them in block order - one launch, no
 * atomics on doubles, bit-reproducible from run to run.
 */
#include <cmath>
#include <vector>

#include "plg_internal.cuh"

#define PLG_LNL_THREADS 256
#define PLG_MAX_RATES 16

struct LnlParams
{
  double freqs[PLG_MAX_RATES * 20]; /* [rate][states_padded] gathered by freqs_indices */
  double rate_weights[PLG_MAX_RATES];
  double prop_invar[PLG_MAX_RATES];
  double log_threshold; /* log(2^-256) from the host libm, as the reference computes it */
  int any_pinv;
};

struct LnlArgs
{
  const double * clvp;          /* the inner ("parent") CLV                         */
  const double * clvc;          /* ii: the other CLV                                */
  const unsigned char * tip;    /* ti: the tip's characters                         */
  const double * pmat;          /* ii: P-matrix set; ti: pi-weighted lookup table   */
  const unsigned int * pscale;
  const unsigned int * cscale;
  const unsigned int * weights;
  const int * invariant;
  double * persite;             /* may be NULL */
  double * partials;
  unsigned int * counter;
  PlgSink sink;
  unsigned int nelem;           /* sites * R */
  int per_rate_scaling;
};

/* ------------------------------------------------------------------------------------ */
/* site epilogue shared by all variants                                                  */
/* ------------------------------------------------------------------------------------ */
/* Gathers the R per-rate terms of a site, combines them as the reference does and returns
 * the weighted site log-likelihood in lane 0 of each R-group (0.0 in the other lanes).
 *   GUARD_POSITIVE : the DNA AVX kernels add a rate's term only if it is > 0
 *   TI_DNA_FREQS   : the DNA tip-inner kernel reads the invariant-site frequency from the
 *                    last rate category's frequency vector (reference
 *                    src/core_likelihood_avx.c:274,370-371) */
template <int R, int K, bool GUARD_POSITIVE, bool TI_DNA_FREQS>
__device__ __forceinline__ double site_epilogue(double term_r, bool valid, unsigned int e,
                                                const LnlArgs & a, const LnlParams & P)
{
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int k = lane & (R - 1);
  const unsigned int gbase = lane & ~(unsigned int)(R - 1);

  unsigned int site_scalings = 0;
  if (a.per_rate_scaling)
  {
    /* per-rate scalers: site scaler = min over rates, residual (capped) applied to the term
     * (reference src/core_likelihood_avx.c:1136-1154,1219-1223) */
    unsigned int rs = 0;
    if (valid)
    {
      if (a.pscale) rs += a.pscale[e];
      if (a.cscale) rs += a.cscale[e];
    }
    unsigned int mn = rs;
#pragma unroll
    for (int off = 1; off < R; off <<= 1)
    {
      const unsigned int o = __shfl_xor_sync(0xffffffffu, mn, off);
      mn = o < mn ? o : mn;
    }
    site_scalings = mn;
    unsigned int diff = rs - mn;
    if (diff > PLL_SCALE_RATE_MAXDIFF) diff = PLL_SCALE_RATE_MAXDIFF;
    if (diff > 0)
    {
      double f = 1.0;
      for (unsigned int q = 0; q < diff; ++q) f = __dmul_rn(f, PLG_SCALE_THRESHOLD);
      term_r = __dmul_rn(term_r, f);
    }
  }
  else if (valid && k == 0)
  {
    const unsigned int n = e / R;
    if (a.pscale) site_scalings += a.pscale[n];
    if (a.cscale) site_scalings += a.cscale[n];
  }

  double term = 0.0;
  int inv = -1;
  if (P.any_pinv && valid && k == 0) inv = a.invariant[e / R];
#pragma unroll
  for (int kk = 0; kk < R; ++kk)
  {
    const double v = __shfl_sync(0xffffffffu, term_r, gbase + kk);
    if (k == 0)
    {
      if (!GUARD_POSITIVE || v > 0.0)
      {
        const double pinv = P.prop_invar[kk];
        if (pinv > 0.0)
        {
          const int fr = TI_DNA_FREQS ? (R - 1) : kk;
          const double inv_lk = (inv == -1) ? 0.0 : P.freqs[fr * K + inv];
          const double mix = __dadd_rn(__dmul_rn(v, __dsub_rn(1.0, pinv)), __dmul_rn(inv_lk, pinv));
          term = __dadd_rn(term, __dmul_rn(P.rate_weights[kk], mix));
        }
        else
          term = __dadd_rn(term, __dmul_rn(v, P.rate_weights[kk]));
      }
    }
  }

  double site_lk = 0.0;
  if (valid && k == 0)
  {
    const unsigned int n = e / R;
    site_lk = log(term);
    if (site_scalings) site_lk = __dadd_rn(site_lk, __dmul_rn((double)site_scalings, P.log_threshold));
    site_lk = __dmul_rn(site_lk, (double)a.weights[n]);
    if (a.persite) a.persite[n] = site_lk;
  }
  return site_lk;
}

/* block partial -> HBM; last block adds all partials in block order */
template <int THREADS>
__device__ __forceinline__ void finish_sum(double v, const LnlArgs & a)
{
  __shared__ double red[THREADS / 32];
  __shared__ bool is_last;
  const double bsum = block_sum<THREADS>(v, red);
  if (threadIdx.x == 0)
  {
    a.partials[blockIdx.x] = bsum;
    __threadfence();
    const unsigned int ticket = atomicAdd(a.counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last)
  {
    __threadfence();
    /* each thread adds a contiguous run of partials in order, then a fixed tree */
    const unsigned int nb = gridDim.x;
    const unsigned int per = (nb + THREADS - 1) / THREADS;
    const unsigned int lo = threadIdx.x * per;
    unsigned int hi = lo + per;
    if (hi > nb) hi = nb;
    double s = 0.0;
    for (unsigned int b = lo; b < hi; ++b) s = __dadd_rn(s, __ldcg(a.partials + b));
    const double total = block_sum<THREADS>(s, red);
    if (threadIdx.x == 0)
    {
      plg_publish(a.sink, total, 0.0);
      *a.counter = 0u;
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* DNA kernels                                                                           */
/* ------------------------------------------------------------------------------------ */
/* One thread per pattern (a DNA pattern is only R x 32 bytes per CLV): all R rate terms, the
 * scaler bookkeeping and the single log of a pattern stay in one thread - no shuffles, no
 * idle lanes around the log.  Persistent grid of one resident wave; per-thread running sum,
 * one block tree at the end.  MODE 0: edge, both ends inner; 1: edge with a pattern tip;
 * 2: root.  Arithmetic order per rate as the reference's 4x4 AVX kernels:
 *   edge ii  src/core_likelihood_avx.c:1166-1217   (P_row . c) * pi * p, then (t0+t1)+(t2+t3)
 *   edge ti  src/core_likelihood_avx.c:296-300,330-345   (pi * masked row sum) * p
 *   root     src/core_likelihood_avx.c:145-156 */
template <int R, int MODE>
__global__ void __launch_bounds__(PLG_LNL_THREADS, 2)
k_lnl_dna(const LnlArgs a, const __grid_constant__ LnlParams P)
{
  constexpr int RC = (R < 4) ? R : 4; /* rates whose loads are issued together */
  __shared__ __align__(32) double Ms[(MODE == 2) ? 4 : ((MODE == 0) ? R * 16 : 16 * R * 4)];
  if (MODE == 0)
    for (unsigned int t = threadIdx.x; t < R * 16; t += PLG_LNL_THREADS) Ms[t] = __ldg(a.pmat + t);
  if (MODE == 1)
    for (unsigned int t = threadIdx.x; t < 16 * R * 4; t += PLG_LNL_THREADS) Ms[t] = __ldg(a.pmat + t);
  __syncthreads();

  const unsigned int sites = a.nelem / R;
  const unsigned int stride = gridDim.x * PLG_LNL_THREADS;
  double sum = 0.0;
  for (unsigned int n = blockIdx.x * PLG_LNL_THREADS + threadIdx.x; n < sites; n += stride)
  {
    const unsigned int weight = __ldg(a.weights + n);
    const int inv = P.any_pinv ? __ldg(a.invariant + n) : -1;
    unsigned int code = 0;
    if (MODE == 1) code = __ldg(a.tip + n);

    /* scaler counts: per-site sum, or per-rate with the site scaler = min over rates and
     * capped residuals (reference src/core_likelihood_avx.c:1136-1154,1219-1223) */
    unsigned int site_scalings = 0;
    unsigned int resid[R];
    if (a.per_rate_scaling)
    {
      unsigned int mn = 0xffffffffu;
#pragma unroll
      for (int r = 0; r < R; ++r)
      {
        unsigned int rs = 0;
        if (a.pscale) rs += __ldg(a.pscale + (size_t)n * R + r);
        if (a.cscale) rs += __ldg(a.cscale + (size_t)n * R + r);
        resid[r] = rs;
        mn = rs < mn ? rs : mn;
      }
      site_scalings = mn;
#pragma unroll
      for (int r = 0; r < R; ++r)
      {
        const unsigned int d = resid[r] - mn;
        resid[r] = d > PLL_SCALE_RATE_MAXDIFF ? PLL_SCALE_RATE_MAXDIFF : d;
      }
    }
    else
    {
      if (a.pscale) site_scalings += __ldg(a.pscale + n);
      if (a.cscale) site_scalings += __ldg(a.cscale + n);
    }

    double term = 0.0;
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += RC)
    {
      d4 p[RC], c[RC];
#pragma unroll
      for (int q = 0; q < RC; ++q)
      {
        p[q] = ld_stream(a.clvp + ((size_t)n * R + r0 + q) * 4);
        if (MODE == 0) c[q] = ld_stream(a.clvc + ((size_t)n * R + r0 + q) * 4);
      }
#pragma unroll
      for (int q = 0; q < RC; ++q)
      {
        const int r = r0 + q;
        const double * f = P.freqs + r * 4;
        double v;
        if (MODE == 0)
        {
          const double * M = Ms + r * 16;
          const double t0 = __dmul_rn(__dmul_rn(f[0], dot4_unfused(M[0], M[1], M[2], M[3], c[q])), p[q].x);
          const double t1 = __dmul_rn(__dmul_rn(f[1], dot4_unfused(M[4], M[5], M[6], M[7], c[q])), p[q].y);
          const double t2 = __dmul_rn(__dmul_rn(f[2], dot4_unfused(M[8], M[9], M[10], M[11], c[q])), p[q].z);
          const double t3 = __dmul_rn(__dmul_rn(f[3], dot4_unfused(M[12], M[13], M[14], M[15], c[q])), p[q].w);
          v = hsum4(t0, t1, t2, t3);
        }
        else if (MODE == 1)
        {
          const d4 l = *reinterpret_cast<const d4 *>(Ms + ((size_t)code * R + r) * 4);
          v = hsum4(__dmul_rn(l.x, p[q].x), __dmul_rn(l.y, p[q].y), __dmul_rn(l.z, p[q].z), __dmul_rn(l.w, p[q].w));
        }
        else
          v = hsum4(__dmul_rn(f[0], p[q].x), __dmul_rn(f[1], p[q].y), __dmul_rn(f[2], p[q].z), __dmul_rn(f[3], p[q].w));

        if (a.per_rate_scaling && resid[r] > 0)
        {
          double sc = 1.0;
          for (unsigned int i = 0; i < resid[r]; ++i) sc = __dmul_rn(sc, PLG_SCALE_THRESHOLD);
          v = __dmul_rn(v, sc);
        }
        /* the edge kernels add a rate's term only if it is positive (reference
         * src/core_likelihood_avx.c:1225); the root kernel adds it unconditionally */
        if (MODE == 2 || v > 0.0)
        {
          const double pinv = P.prop_invar[r];
          if (pinv > 0.0)
          {
            /* the tip-inner kernel reads the invariant frequency from the LAST rate's vector
             * (reference src/core_likelihood_avx.c:274,370-371) */
            const int fr = (MODE == 1) ? (R - 1) : r;
            const double inv_lk = (inv == -1) ? 0.0 : P.freqs[fr * 4 + inv];
            const double mix = __dadd_rn(__dmul_rn(v, __dsub_rn(1.0, pinv)), __dmul_rn(inv_lk, pinv));
            term = __dadd_rn(term, __dmul_rn(P.rate_weights[r], mix));
          }
          else
            term = __dadd_rn(term, __dmul_rn(v, P.rate_weights[r]));
        }
      }
    }
    double site_lk = log(term);
    if (site_scalings) site_lk = __dadd_rn(site_lk, __dmul_rn((double)site_scalings, P.log_threshold));
    site_lk = __dmul_rn(site_lk, (double)weight);
    if (a.persite) a.persite[n] = site_lk;
    sum = __dadd_rn(sum, site_lk);
  }
  finish_sum<PLG_LNL_THREADS>(sum, a);
}

template <int R, int MODE>
static int launch_lnl_dna(plg_context * ctx, LnlArgs & a, const LnlParams & P)
{
  /* one resident wave; the grid depends only on the device => reproducible sums */
  static int per_sm = 0;
  if (!per_sm)
    PLG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lnl_dna<R, MODE>, PLG_LNL_THREADS, 0));
  const unsigned int sites = a.nelem / R;
  unsigned int nblocks = (sites + PLG_LNL_THREADS - 1) / PLG_LNL_THREADS;
  const unsigned int cap = (unsigned int)(ctx->sm_count * (per_sm > 0 ? per_sm : 1));
  if (nblocks > cap) nblocks = cap;
  if (nblocks == 0) nblocks = 1;
  k_lnl_dna<R, MODE><<<nblocks, PLG_LNL_THREADS, 0, ctx->stream>>>(a, P);
  return PLG_OK;
}

/* lookup[code][rate][i] = pi_i * (masked row sum of P_rate[i][:])
 * reference src/core_likelihood_avx.c:261-309 */
__global__ void k_edge_ti_table_dna(const double * __restrict__ pmat, double * __restrict__ out,
                                    unsigned int rate_cats, const __grid_constant__ LnlParams P)
{
  const unsigned int entries = 16u * rate_cats * 4u;
  for (unsigned int t = threadIdx.x; t < entries; t += blockDim.x)
  {
    const unsigned int i = t & 3u;
    const unsigned int k = (t >> 2) % rate_cats;
    const unsigned int code = t / (rate_cats * 4u);
    const double * row = pmat + (size_t)k * 16 + i * 4;
    const double a0 = (code & 1u) ? row[0] : 0.0;
    const double a1 = (code & 2u) ? row[1] : 0.0;
    const double a2 = (code & 4u) ? row[2] : 0.0;
    const double a3 = (code & 8u) ? row[3] : 0.0;
    out[t] = __dmul_rn(P.freqs[k * 4 + i], hsum4(a0, a1, a2, a3));
  }
}

/* ------------------------------------------------------------------------------------ */
/* 20-state kernels                                                                      */
/* ------------------------------------------------------------------------------------ */
__device__ __forceinline__ void load20s(const double * p, double (&c)[20])
{
#pragma unroll
  for (int b = 0; b < 5; ++b)
  {
    const d4 v = ld_stream(p + 4 * b);
    c[4 * b + 0] = v.x;
    c[4 * b + 1] = v.y;
    c[4 * b + 2] = v.z;
    c[4 * b + 3] = v.w;
  }
}

template <int R>
__global__ void __launch_bounds__(PLG_LNL_THREADS)
k_edge_lnl_ii_aa(const LnlArgs a, const __grid_constant__ LnlParams P)
{
  extern __shared__ __align__(16) double Ms[]; /* [R][20][20] */
  for (unsigned int t = threadIdx.x; t < R * 400; t += PLG_LNL_THREADS) Ms[t] = __ldg(a.pmat + t);
  __syncthreads();

  const unsigned int k = threadIdx.x & (R - 1);
  const unsigned int e = blockIdx.x * PLG_LNL_THREADS + threadIdx.x;
  const bool valid = e < a.nelem;
  double term_r = 0.0;
  if (valid)
  {
    double c[20], p[20];
    load20s(a.clvc + (size_t)e * 20, c);
    load20s(a.clvp + (size_t)e * 20, p);
    const double * Mk = Ms + k * 400;
    const double * f = P.freqs + k * 20;
    /* reference src/core_likelihood_avx2.c:431-502 */
#pragma unroll
    for (int jb = 0; jb < 5; ++jb)
    {
      double t[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
      {
        const double * row = Mk + (4 * jb + q) * 20;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int b = 0; b < 5; ++b)
        {
          a0 = __fma_rn(row[4 * b + 0], c[4 * b + 0], a0);
          a1 = __fma_rn(row[4 * b + 1], c[4 * b + 1], a1);
          a2 = __fma_rn(row[4 * b + 2], c[4 * b + 2], a2);
          a3 = __fma_rn(row[4 * b + 3], c[4 * b + 3], a3);
        }
        t[q] = __dmul_rn(__dmul_rn(hsum4(a0, a1, a2, a3), f[4 * jb + q]), p[4 * jb + q]);
      }
      term_r = __dadd_rn(term_r, hsum4(t[0], t[1], t[2], t[3]));
    }
  }
  const double site_lk = site_epilogue<R, 20, false, false>(term_r, valid, e, a, P);
  finish_sum<PLG_LNL_THREADS>(site_lk, a);
}

struct TipmapArgL
{
  unsigned int map[PLL_ASCII_SIZE];
};

/* lookup[code][rate][i] = (sequential sum of P_rate[i][m] over states m in tipmap[code]) * pi_i
 * reference src/core_likelihood_avx2.c:193-233 */
__global__ void k_edge_ti_table_aa(const double * __restrict__ pmat, double * __restrict__ out,
                                   unsigned int rate_cats, unsigned int maxstates,
                                   const TipmapArgL tm, const __grid_constant__ LnlParams P)
{
  const unsigned int entries = maxstates * rate_cats * 20u;
  for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < entries;
       t += gridDim.x * blockDim.x)
  {
    const unsigned int i = t % 20u;
    const unsigned int k = (t / 20u) % rate_cats;
    const unsigned int code = t / (rate_cats * 20u);
    const unsigned int state = tm.map[code];
    const double * row = pmat + (size_t)k * 400 + i * 20;
    double s = 0.0;
    for (unsigned int m = 0; m < 20u; ++m)
      if ((state >> m) & 1u) s = __dadd_rn(s, row[m]);
    out[t] = __dmul_rn(s, P.freqs[k * 20 + i]);
  }
}

/* 20 states, one thread per pattern (as k_lnl_dna): MODE 1 edge with a pattern tip (pi-weighted
 * lookup table in a.pmat, [code][rate][20]), MODE 2 root (frequencies), MODE 3 plain sum of a
 * scratch CLV that already holds pi_i p_i (P c)_i (second half of the two-step inner-inner edge).
 * Per rate: four FMA lane accumulators over the five blocks of four states, then the AVX
 * horizontal sum (reference src/core_likelihood_avx2.c:57-77, 266-286). */
template <int R, int MODE>
__global__ void __launch_bounds__(PLG_LNL_THREADS, 2)
k_lnl_aa(const LnlArgs a, const __grid_constant__ LnlParams P)
{
  const unsigned int sites = a.nelem / R;
  const unsigned int stride = gridDim.x * PLG_LNL_THREADS;
  double sum = 0.0;
  for (unsigned int n = blockIdx.x * PLG_LNL_THREADS + threadIdx.x; n < sites; n += stride)
  {
    const unsigned int weight = __ldg(a.weights + n);
    const int inv = P.any_pinv ? __ldg(a.invariant + n) : -1;
    unsigned int code = 0;
    if (MODE == 1) code = __ldg(a.tip + n);

    unsigned int site_scalings = 0;
    unsigned int resid[R];
    if (a.per_rate_scaling)
    {
      unsigned int mn = 0xffffffffu;
#pragma unroll
      for (int r = 0; r < R; ++r)
      {
        unsigned int rs = 0;
        if (a.pscale) rs += __ldg(a.pscale + (size_t)n * R + r);
        if (a.cscale) rs += __ldg(a.cscale + (size_t)n * R + r);
        resid[r] = rs;
        mn = rs < mn ? rs : mn;
      }
      site_scalings = mn;
#pragma unroll
      for (int r = 0; r < R; ++r)
      {
        const unsigned int d = resid[r] - mn;
        resid[r] = d > PLL_SCALE_RATE_MAXDIFF ? PLL_SCALE_RATE_MAXDIFF : d;
      }
    }
    else
    {
      if (a.pscale) site_scalings += __ldg(a.pscale + n);
      if (a.cscale) site_scalings += __ldg(a.cscale + n);
    }

    double term = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r)
    {
      double c[20];
      load20s(a.clvp + ((size_t)n * R + r) * 20, c);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      if (MODE == 3)
      {
#pragma unroll
        for (int b = 0; b < 5; ++b)
        {
          a0 = __dadd_rn(a0, c[4 * b + 0]);
          a1 = __dadd_rn(a1, c[4 * b + 1]);
          a2 = __dadd_rn(a2, c[4 * b + 2]);
          a3 = __dadd_rn(a3, c[4 * b + 3]);
        }
      }
      else
      {
        const double * tab = a.pmat + ((size_t)code * R + r) * 20;
#pragma unroll
        for (int b = 0; b < 5; ++b)
        {
          d4 l;
          if (MODE == 1)
            l = *reinterpret_cast<const d4 *>(tab + 4 * b);
          else
            l = d4{P.freqs[r * 20 + 4 * b], P.freqs[r * 20 + 4 * b + 1], P.freqs[r * 20 + 4 * b + 2],
                   P.freqs[r * 20 + 4 * b + 3]};
          a0 = __fma_rn(l.x, c[4 * b + 0], a0);
          a1 = __fma_rn(l.y, c[4 * b + 1], a1);
          a2 = __fma_rn(l.z, c[4 * b + 2], a2);
          a3 = __fma_rn(l.w, c[4 * b + 3], a3);
        }
      }
      double v = hsum4(a0, a1, a2, a3);
      if (a.per_rate_scaling && resid[r] > 0)
      {
        double sc = 1.0;
        for (unsigned int i = 0; i < resid[r]; ++i) sc = __dmul_rn(sc, PLG_SCALE_THRESHOLD);
        v = __dmul_rn(v, sc);
      }
      const double pinv = P.prop_invar[r];
      if (pinv > 0.0)
      {
        const double inv_lk = (inv == -1) ? 0.0 : P.freqs[r * 20 + inv];
        const double mix = __dadd_rn(__dmul_rn(v, __dsub_rn(1.0, pinv)), __dmul_rn(inv_lk, pinv));
        term = __dadd_rn(term, __dmul_rn(P.rate_weights[r], mix));
      }
      else
        term = __dadd_rn(term, __dmul_rn(v, P.rate_weights[r]));
    }
    double site_lk = log(term);
    if (site_scalings) site_lk = __dadd_rn(site_lk, __dmul_rn((double)site_scalings, P.log_threshold));
    site_lk = __dmul_rn(site_lk, (double)weight);
    if (a.persite) a.persite[n] = site_lk;
    sum = __dadd_rn(sum, site_lk);
  }
  finish_sum<PLG_LNL_THREADS>(sum, a);
}

template <int R, int MODE>
static int launch_lnl_aa(plg_context * ctx, LnlArgs & a, const LnlParams & P)
{
  static int per_sm = 0;
  if (!per_sm)
    PLG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lnl_aa<R, MODE>, PLG_LNL_THREADS, 0));
  const unsigned int sites = a.nelem / R;
  unsigned int nblocks = (sites + PLG_LNL_THREADS - 1) / PLG_LNL_THREADS;
  const unsigned int cap = (unsigned int)(ctx->sm_count * (per_sm > 0 ? per_sm : 1));
  if (nblocks > cap) nblocks = cap;
  if (nblocks == 0) nblocks = 1;
  k_lnl_aa<R, MODE><<<nblocks, PLG_LNL_THREADS, 0, ctx->stream>>>(a, P);
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* host side                                                                             */
/* ------------------------------------------------------------------------------------ */
static int fill_params(plg_context * ctx, const double * freqs, const double * rate_weights,
                       const double * prop_invar, LnlParams & P)
{
  const unsigned int R = ctx->d.rate_cats, Kp = ctx->d.states_padded;
  memset(&P, 0, sizeof(P));
  memcpy(P.freqs, freqs, (size_t)R * Kp * sizeof(double));
  memcpy(P.rate_weights, rate_weights, R * sizeof(double));
  P.any_pinv = 0;
  for (unsigned int i = 0; i < R; ++i)
  {
    P.prop_invar[i] = prop_invar ? prop_invar[i] : 0.0;
    if (P.prop_invar[i] > 0) P.any_pinv = 1;
  }
  P.log_threshold = log(PLL_SCALE_THRESHOLD);
  if (P.any_pinv && !ctx->has_invariant)
  {
    plg_set_error("log-likelihood with prop_invar > 0 needs the invariant-site index "
                  "(pll_update_invariant_sites)");
    return PLG_E_INVALID;
  }
  return PLG_OK;
}

static int common_args(plg_context * ctx, LnlArgs & a, double * persite_lnl, unsigned int * nblocks)
{
  const unsigned int nelem = ctx->active_sites * ctx->d.rate_cats;
  *nblocks = (nelem + PLG_LNL_THREADS - 1) / PLG_LNL_THREADS;
  if (*nblocks == 0) *nblocks = 1; /* no active pattern: one block still delivers the (zero) sum */
  int rc = plg_ensure_partials(ctx, *nblocks);
  if (rc) return rc;
  if (persite_lnl && !ctx->persite_dev)
    PLG_CUDA(cudaMalloc(&ctx->persite_dev, (size_t)ctx->active_sites * sizeof(double)));
  a.weights = ctx->weights;
  a.invariant = ctx->invariant;
  a.persite = persite_lnl ? ctx->persite_dev : NULL;
  a.partials = ctx->partials;
  a.counter = ctx->counter;
  a.sink = plg_make_sink(ctx);
  a.nelem = nelem;
  a.per_rate_scaling = ctx->rate_scalers ? 1 : 0;
  return PLG_OK;
}

static int fetch_result(plg_context * ctx, double * persite_lnl, double * logl_out)
{
  ctx->stats.d2h_bytes += sizeof(double);
  if (persite_lnl)
  {
    PLG_CUDA(cudaMemcpyAsync(persite_lnl, ctx->persite_dev, (size_t)ctx->active_sites * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    ctx->copy_pending = 1;
    ctx->stats.d2h_bytes += (size_t)ctx->active_sites * sizeof(double);
  }
  return plg_finish_result(ctx, logl_out, NULL);
}

#define PLG_DISPATCH_R(R_, ...)                                                         \
  switch (R_)                                                                           \
  {                                                                                     \
    case 1: { constexpr int RR = 1; __VA_ARGS__; } break;                                      \
    case 2: { constexpr int RR = 2; __VA_ARGS__; } break;                                      \
    case 4: { constexpr int RR = 4; __VA_ARGS__; } break;                                      \
    case 8: { constexpr int RR = 8; __VA_ARGS__; } break;                                      \
    case 16: { constexpr int RR = 16; __VA_ARGS__; } break;                                    \
    default: plg_set_error("rate_cats=%u unsupported", R_); return PLG_E_UNSUPPORTED;   \
  }

extern "C" int plg_edge_loglikelihood(plg_context_t * ctx, unsigned int parent_clv_index,
                                      int parent_scaler_index, unsigned int child_clv_index,
                                      int child_scaler_index, unsigned int matrix_index,
                                      const double * freqs, const double * rate_weights,
                                      const double * prop_invar, double * persite_lnl,
                                      double * logl_out)
{
  PLG_CHECK_CTX(ctx);
  const unsigned int n_clv = ctx->d.tips + ctx->d.clv_buffers;
  if (parent_clv_index >= n_clv || child_clv_index >= n_clv ||
      matrix_index >= ctx->d.prob_matrices || parent_scaler_index >= (int)ctx->d.scale_buffers ||
      child_scaler_index >= (int)ctx->d.scale_buffers || !logl_out)
  {
    plg_set_error("plg_edge_loglikelihood: index out of range");
    return PLG_E_INVALID;
  }
  const bool ptip = plg_is_tip(ctx, parent_clv_index);
  const bool ctip = plg_is_tip(ctx, child_clv_index);
  if (ptip && ctip)
  {
    plg_set_error("plg_edge_loglikelihood: edge between two pattern tips is not supported "
                  "(nor by the reference, src/likelihood.c:489-501)");
    return PLG_E_UNSUPPORTED;
  }
  if (!plg_fast_path(ctx))
  {
    GenLnl g;
    memset(&g, 0, sizeof(g));
    g.pmat = plg_pmat_ptr(ctx, matrix_index);
    if (ptip || ctip)
    {
      g.clvp = plg_clv_ptr(ctx, ptip ? child_clv_index : parent_clv_index);
      g.tip = plg_tip_ptr(ctx, ptip ? parent_clv_index : child_clv_index);
      g.pscale = plg_scaler_ptr(ctx, ptip ? child_scaler_index : parent_scaler_index);
    }
    else
    {
      g.clvp = plg_clv_ptr(ctx, parent_clv_index);
      g.clvc = plg_clv_ptr(ctx, child_clv_index);
      g.pscale = plg_scaler_ptr(ctx, parent_scaler_index);
      g.cscale = plg_scaler_ptr(ctx, child_scaler_index);
    }
    return plg_gen_loglikelihood(ctx, g, freqs, rate_weights, prop_invar, persite_lnl, logl_out);
  }

  LnlParams P;
  int rc = fill_params(ctx, freqs, rate_weights, prop_invar, P);
  if (rc) return rc;
  LnlArgs a;
  memset(&a, 0, sizeof(a));
  unsigned int nblocks = 0;
  rc = common_args(ctx, a, persite_lnl, &nblocks);
  if (rc) return rc;

  const unsigned int R = ctx->d.rate_cats, K = ctx->d.states;
  if (ptip || ctip)
  {
    /* the inner node plays "parent", its scaler is the only one used
     * (reference src/likelihood.c:486-501) */
    const unsigned int inner = ptip ? child_clv_index : parent_clv_index;
    const unsigned int tip = ptip ? parent_clv_index : child_clv_index;
    const int inner_sc = ptip ? child_scaler_index : parent_scaler_index;
    a.clvp = plg_clv_ptr(ctx, inner);
    a.tip = plg_tip_ptr(ctx, tip);
    a.pscale = plg_scaler_ptr(ctx, inner_sc);
    a.cscale = NULL;
    const size_t tab_len = (size_t)(K == 4 ? 16u : ctx->maxstates) * R * K;
    if (tab_len > ctx->lnl_table_cap)
    {
      PLG_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->lnl_table);
      ctx->lnl_table = NULL;
      ctx->lnl_table_cap = 0;
      PLG_CUDA(cudaMalloc(&ctx->lnl_table, tab_len * sizeof(double)));
      ctx->lnl_table_cap = tab_len;
    }
    double * scratch = ctx->lnl_table;
    a.pmat = scratch;
    if (K == 4)
    {
      k_edge_ti_table_dna<<<1, 256, 0, ctx->stream>>>(plg_pmat_ptr(ctx, matrix_index), scratch, R, P);
      PLG_LAUNCH_CHECK(ctx);
      PLG_DISPATCH_R(R, { int lrc = launch_lnl_dna<RR, 1>(ctx, a, P); if (lrc) return lrc; });
    }
    else
    {
      TipmapArgL tm;
      memcpy(tm.map, ctx->tipmap, sizeof(tm.map));
      k_edge_ti_table_aa<<<8, 256, 0, ctx->stream>>>(plg_pmat_ptr(ctx, matrix_index), scratch, R,
                                                     ctx->maxstates, tm, P);
      PLG_LAUNCH_CHECK(ctx);
      PLG_DISPATCH_R(R, { int lrc = launch_lnl_aa<RR, 1>(ctx, a, P); if (lrc) return lrc; });
    }
    PLG_LAUNCH_CHECK(ctx);
  }
  else
  {
    a.clvp = plg_clv_ptr(ctx, parent_clv_index);
    a.clvc = plg_clv_ptr(ctx, child_clv_index);
    a.pmat = plg_pmat_ptr(ctx, matrix_index);
    a.pscale = plg_scaler_ptr(ctx, parent_scaler_index);
    a.cscale = plg_scaler_ptr(ctx, child_scaler_index);
    if (K == 4)
    {
      PLG_DISPATCH_R(R, { int lrc = launch_lnl_dna<RR, 0>(ctx, a, P); if (lrc) return lrc; });
    }
    else if (!ctx->aa_exact)
    {
      /* two steps on the tensor cores: the inner-inner update kernel with L = diag(pi) and
       * R = P writes pi_i p_i (P c)_i into a scratch CLV (same shape as a CLV update, no
       * scaling), then one streaming pass sums it per rate and finishes the pattern.  3.4x
       * faster than the shared-memory mat-vec kernel below, which stays as the bit-exact
       * (PLL_GPU_AA_EXACT) path. */
      if (!ctx->lnl_scratch)
        PLG_CUDA(cudaMalloc(&ctx->lnl_scratch, (size_t)ctx->d.sites * ctx->span * sizeof(double)));
      std::vector<double> diag((size_t)R * 400, 0.0);
      for (unsigned int r = 0; r < R; ++r)
        for (unsigned int i = 0; i < 20; ++i) diag[(size_t)r * 400 + i * 20 + i] = freqs[(size_t)r * 20 + i];
      if (plg_stage_reserve(ctx, diag.size() * sizeof(double) + 2048)) return PLG_E_CUDA;
      const double * d_diag = (const double *)plg_stage(ctx, diag.data(), diag.size() * sizeof(double));
      if (!d_diag) return PLG_E_CUDA;
      DevOp op;
      memset(&op, 0, sizeof(op));
      op.parent = ctx->lnl_scratch;
      op.left = a.clvp;
      op.right = a.clvc;
      op.lmat = d_diag;
      op.rmat = a.pmat;
      rc = plg_launch_single_op(ctx, PLG_KIND_II, op);
      if (rc) return rc;
      a.clvp = ctx->lnl_scratch;
      a.clvc = NULL;
      PLG_DISPATCH_R(R, { int lrc = launch_lnl_aa<RR, 3>(ctx, a, P); if (lrc) return lrc; });
    }
    else
    {
      const size_t smem = (size_t)R * 400 * sizeof(double);
      PLG_DISPATCH_R(R, {
        static bool attr_done[PLG_MAX_DEVICES] = {}; /* a function attribute is per device */
        if (!attr_done[ctx->device % PLG_MAX_DEVICES])
        {
          PLG_CUDA(cudaFuncSetAttribute(k_edge_lnl_ii_aa<RR>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          attr_done[ctx->device % PLG_MAX_DEVICES] = true;
        }
        k_edge_lnl_ii_aa<RR><<<nblocks, PLG_LNL_THREADS, smem, ctx->stream>>>(a, P);
      });
    }
    PLG_LAUNCH_CHECK(ctx);
  }
  return fetch_result(ctx, persite_lnl, logl_out);
}

/* root log-likelihood with the scaler counts at `pscale` (device; NULL: none) */
static int root_loglikelihood(plg_context * ctx, unsigned int clv_index, const unsigned int * pscale,
                              const double * freqs, const double * rate_weights, const double * prop_invar,
                              double * persite_lnl, double * logl_out)
{
  if (!plg_fast_path(ctx))
  {
    GenLnl g;
    memset(&g, 0, sizeof(g));
    g.clvp = plg_clv_ptr(ctx, clv_index);
    g.pscale = pscale;
    g.root = 1;
    return plg_gen_loglikelihood(ctx, g, freqs, rate_weights, prop_invar, persite_lnl, logl_out);
  }
  LnlParams P;
  int rc = fill_params(ctx, freqs, rate_weights, prop_invar, P);
  if (rc) return rc;
  LnlArgs a;
  memset(&a, 0, sizeof(a));
  unsigned int nblocks = 0;
  rc = common_args(ctx, a, persite_lnl, &nblocks);
  if (rc) return rc;
  a.clvp = plg_clv_ptr(ctx, clv_index);
  a.pscale = pscale;
  /* the root kernels index the scaler per site even in per-rate mode
   * (reference src/core_likelihood_avx.c:176-178) */
  a.per_rate_scaling = 0;
  const unsigned int R = ctx->d.rate_cats;
  if (ctx->d.states == 4)
  {
    PLG_DISPATCH_R(R, { int lrc = launch_lnl_dna<RR, 2>(ctx, a, P); if (lrc) return lrc; });
  }
  else
  {
    PLG_DISPATCH_R(R, { int lrc = launch_lnl_aa<RR, 2>(ctx, a, P); if (lrc) return lrc; });
  }
  PLG_LAUNCH_CHECK(ctx);
  return fetch_result(ctx, persite_lnl, logl_out);
}

extern "C" int plg_root_loglikelihood(plg_context_t * ctx, unsigned int clv_index,
                                      int scaler_index, const double * freqs,
                                      const double * rate_weights, const double * prop_invar,
                                      double * persite_lnl, double * logl_out)
{
  PLG_CHECK_CTX(ctx);
  if (clv_index < ctx->clv_first || clv_index >= ctx->d.tips + ctx->d.clv_buffers ||
      scaler_index >= (int)ctx->d.scale_buffers || !logl_out)
  {
    plg_set_error("plg_root_loglikelihood: index out of range");
    return PLG_E_INVALID;
  }
  return root_loglikelihood(ctx, clv_index, plg_scaler_ptr(ctx, scaler_index), freqs, rate_weights, prop_invar,
                            persite_lnl, logl_out);
}

/* The same with one scaler count per pattern handed over by the caller (host array, sites
 * entries).  With per-rate scalers the reference's root kernels read element n of the
 * [site][rate] array for pattern n (src/core_likelihood_avx.c:176-178); on a partition cut into
 * pattern slices that element lives in another slice, so the host layer collects the counts
 * (libpll_b200/csrc/host/pll_devices.c) and passes each slice its own. */
extern "C" int plg_root_loglikelihood_counts(plg_context_t * ctx, unsigned int clv_index,
                                             const unsigned int * site_counts, const double * freqs,
                                             const double * rate_weights, const double * prop_invar,
                                             double * persite_lnl, double * logl_out)
{
  PLG_CHECK_CTX(ctx);
  if (clv_index < ctx->clv_first || clv_index >= ctx->d.tips + ctx->d.clv_buffers || !site_counts || !logl_out)
  {
    plg_set_error("plg_root_loglikelihood_counts: invalid argument");
    return PLG_E_INVALID;
  }
  if (!ctx->root_counts)
    PLG_CUDA(cudaMalloc(&ctx->root_counts, (size_t)ctx->d.sites * sizeof(unsigned int)));
  /* pageable source: consumed when the call returns */
  PLG_CUDA(cudaMemcpyAsync(ctx->root_counts, site_counts, (size_t)ctx->d.sites * sizeof(unsigned int),
                           cudaMemcpyHostToDevice, ctx->stream));
  ctx->stats.h2d_bytes += (size_t)ctx->d.sites * sizeof(unsigned int);
  return root_loglikelihood(ctx, clv_index, ctx->root_counts, freqs, rate_weights, prop_invar, persite_lnl,
                            logl_out);
}
