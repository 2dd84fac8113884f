/*
 * plg_derivatives.cu - sumtable and first/second derivatives of -lnL w.r.t. a branch length
 * (the inner loop of Newton branch-length optimisation).
 *
 * Replaces (AVX2-flag rungs = parity spec):
 *   pll_core_update_sumtable_ii   4x4: reference src/core_derivatives_avx.c:25-207
 *                                 gen: reference src/core_derivatives_avx2.c:24-272
 *   pll_core_update_sumtable_ti   4x4: reference src/core_derivatives_avx.c:462-645
 *                                 gen: reference src/core_derivatives_avx2.c:274-521
 *   pll_core_likelihood_derivatives(_avx2)  reference src/core_derivatives.c:501-732,
 *                                           src/core_derivatives_avx2.c:523-800
 *
 * The sumtable never leaves HBM: it is stored in a device slot keyed by the caller's host
 * buffer address (SURVEY.md 8b "Sumtable on GPU").  Host-side precomputation of the small
 * per-call tables (pi-weighted transposed inverse eigenvectors, per-code left terms,
 * diagptable) is done by the C layer with the reference's operation order; the kernels
 * consume them from the staging ring.
 */
#include <cmath>

#include "plg_internal.cuh"

#define PLG_DER_THREADS 256
#define PLG_MAX_RATES 16

struct SumArgs
{
  const double * clvp;       /* ii: parent CLV; ti: the inner CLV                       */
  const double * clvc;       /* ii: child CLV                                            */
  const unsigned char * tip; /* ti                                                       */
  const double * left;       /* ii: W[R][K][K] = inv_eigenvecs^T * pi; ti: [codes][R][K] */
  const double * right;      /* eigenvecs [R][K][K]                                      */
  const unsigned int * pscale;
  const unsigned int * cscale;
  double * sumtable;
  unsigned int nelem;
  int per_rate_scaling;
};

/* per-rate residual scaling of a sumtable element (only in PLL_ATTRIB_RATE_SCALERS mode;
 * reference src/core_derivatives_avx.c:101-118,187-191) */
template <int R>
__device__ __forceinline__ double rate_residual(bool valid, unsigned int e, const SumArgs & a)
{
  if (!a.per_rate_scaling) return 1.0;
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
  unsigned int diff = rs - mn;
  if (diff > PLL_SCALE_RATE_MAXDIFF) diff = PLL_SCALE_RATE_MAXDIFF;
  double f = 1.0;
  for (unsigned int q = 0; q < diff; ++q) f = __dmul_rn(f, PLG_SCALE_THRESHOLD);
  return f;
}

/* ------------------------------------------------------------------------------------ */
/* DNA sumtables                                                                         */
/* ------------------------------------------------------------------------------------ */
/* Each thread owns PLG_SUM_ITEMS elements (site, rate) of one rate category, 256 apart, and
 * keeps that category's 4x4 left / right matrices in registers: all streaming loads of the
 * thread are issued before the arithmetic. */
#define PLG_SUM_ITEMS 4
#define PLG_SUM_ITEMS_II 2

template <int R>
__global__ void __launch_bounds__(PLG_DER_THREADS, 2) k_sumtable_ii_dna(const SumArgs a)
{
  const unsigned int k = threadIdx.x & (R - 1);
  const unsigned int base = blockIdx.x * (PLG_DER_THREADS * PLG_SUM_ITEMS_II) + threadIdx.x;
  double W[16], V[16];
#pragma unroll
  for (int i = 0; i < 16; i += 4)
  {
    const d4 w = *reinterpret_cast<const d4 *>(a.left + k * 16 + i);
    const d4 v = *reinterpret_cast<const d4 *>(a.right + k * 16 + i);
    W[i] = w.x; W[i + 1] = w.y; W[i + 2] = w.z; W[i + 3] = w.w;
    V[i] = v.x; V[i + 1] = v.y; V[i + 2] = v.z; V[i + 3] = v.w;
  }
  d4 p[PLG_SUM_ITEMS_II], c[PLG_SUM_ITEMS_II];
#pragma unroll
  for (int u = 0; u < PLG_SUM_ITEMS_II; ++u)
  {
    const unsigned int e = base + u * PLG_DER_THREADS;
    if (e < a.nelem)
    {
      p[u] = ld_stream(a.clvp + (size_t)e * 4);
      c[u] = ld_stream(a.clvc + (size_t)e * 4);
    }
  }
#pragma unroll
  for (int u = 0; u < PLG_SUM_ITEMS_II; ++u)
  {
    const unsigned int e = base + u * PLG_DER_THREADS;
    const bool valid = e < a.nelem; /* warp-uniform up to the last warp; shuffles inside */
    const double f = rate_residual<R>(valid, e, a);
    if (!valid) continue;
    d4 s;
    /* left_j = (W_j . p), right_j = (V_j . c), both unfused with (a0+a1)+(a2+a3)
     * reference src/core_derivatives_avx.c:131-185 */
    s.x = __dmul_rn(dot4_unfused(W[0], W[1], W[2], W[3], p[u]), dot4_unfused(V[0], V[1], V[2], V[3], c[u]));
    s.y = __dmul_rn(dot4_unfused(W[4], W[5], W[6], W[7], p[u]), dot4_unfused(V[4], V[5], V[6], V[7], c[u]));
    s.z = __dmul_rn(dot4_unfused(W[8], W[9], W[10], W[11], p[u]), dot4_unfused(V[8], V[9], V[10], V[11], c[u]));
    s.w = __dmul_rn(dot4_unfused(W[12], W[13], W[14], W[15], p[u]), dot4_unfused(V[12], V[13], V[14], V[15], c[u]));
    if (f != 1.0)
    {
      s.x = __dmul_rn(s.x, f);
      s.y = __dmul_rn(s.y, f);
      s.z = __dmul_rn(s.z, f);
      s.w = __dmul_rn(s.w, f);
    }
    st_stream(a.sumtable + (size_t)e * 4, s);
  }
}

template <int R>
__global__ void __launch_bounds__(PLG_DER_THREADS) k_sumtable_ti_dna(const SumArgs a)
{
  const unsigned int k = threadIdx.x & (R - 1);
  const unsigned int base = blockIdx.x * (PLG_DER_THREADS * PLG_SUM_ITEMS) + threadIdx.x;
  double V[16];
#pragma unroll
  for (int i = 0; i < 16; i += 4)
  {
    const d4 v = *reinterpret_cast<const d4 *>(a.right + k * 16 + i);
    V[i] = v.x; V[i + 1] = v.y; V[i + 2] = v.z; V[i + 3] = v.w;
  }
  d4 cl[PLG_SUM_ITEMS];
  unsigned int code[PLG_SUM_ITEMS];
#pragma unroll
  for (int u = 0; u < PLG_SUM_ITEMS; ++u)
  {
    const unsigned int e = base + u * PLG_DER_THREADS;
    code[u] = 0;
    if (e < a.nelem)
    {
      cl[u] = ld_stream(a.clvp + (size_t)e * 4);
      code[u] = __ldg(a.tip + e / R);
    }
  }
#pragma unroll
  for (int u = 0; u < PLG_SUM_ITEMS; ++u)
  {
    const unsigned int e = base + u * PLG_DER_THREADS;
    const bool valid = e < a.nelem;
    const double f = rate_residual<R>(valid, e, a);
    if (!valid) continue;
    const d4 c = cl[u];
    const d4 L = *reinterpret_cast<const d4 *>(a.left + ((size_t)code[u] * R + k) * 4);
    /* right_j accumulates sequentially over the child states (reference
     * src/core_derivatives_avx.c:611-620: broadcast clvc[k] times column k of V^T) */
    double r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
      double acc = 0.0;
      acc = __dadd_rn(acc, __dmul_rn(V[j * 4 + 0], c.x));
      acc = __dadd_rn(acc, __dmul_rn(V[j * 4 + 1], c.y));
      acc = __dadd_rn(acc, __dmul_rn(V[j * 4 + 2], c.z));
      acc = __dadd_rn(acc, __dmul_rn(V[j * 4 + 3], c.w));
      r[j] = acc;
    }
    d4 s;
    s.x = __dmul_rn(L.x, r[0]);
    s.y = __dmul_rn(L.y, r[1]);
    s.z = __dmul_rn(L.z, r[2]);
    s.w = __dmul_rn(L.w, r[3]);
    if (f != 1.0)
    {
      s.x = __dmul_rn(s.x, f);
      s.y = __dmul_rn(s.y, f);
      s.z = __dmul_rn(s.z, f);
      s.w = __dmul_rn(s.w, f);
    }
    st_stream(a.sumtable + (size_t)e * 4, s);
  }
}

/* ------------------------------------------------------------------------------------ */
/* 20-state sumtables                                                                    */
/* ------------------------------------------------------------------------------------ */
__device__ __forceinline__ void load20d(const double * p, double (&c)[20])
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

__device__ __forceinline__ double row20_fma(const double * __restrict__ row, const double (&c)[20])
{
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int b = 0; b < 5; ++b)
  {
    a0 = __fma_rn(row[4 * b + 0], c[4 * b + 0], a0);
    a1 = __fma_rn(row[4 * b + 1], c[4 * b + 1], a1);
    a2 = __fma_rn(row[4 * b + 2], c[4 * b + 2], a2);
    a3 = __fma_rn(row[4 * b + 3], c[4 * b + 3], a3);
  }
  return hsum4(a0, a1, a2, a3);
}

template <int R>
__global__ void __launch_bounds__(PLG_DER_THREADS) k_sumtable_ii_aa(const SumArgs a)
{
  extern __shared__ __align__(16) double sm[];
  double * Ws = sm;           /* [R][20][20] */
  double * Vs = sm + R * 400; /* [R][20][20] */
  for (unsigned int t = threadIdx.x; t < R * 400; t += PLG_DER_THREADS)
  {
    Ws[t] = __ldg(a.left + t);
    Vs[t] = __ldg(a.right + t);
  }
  __syncthreads();
  const unsigned int k = threadIdx.x & (R - 1);
  const unsigned int e = blockIdx.x * PLG_DER_THREADS + threadIdx.x;
  const bool valid = e < a.nelem;
  const double f = rate_residual<R>(valid, e, a);
  if (!valid) return;
  double p[20], c[20];
  load20d(a.clvp + (size_t)e * 20, p);
  load20d(a.clvc + (size_t)e * 20, c);
  double * out = a.sumtable + (size_t)e * 20;
  /* reference src/core_derivatives_avx2.c:158-258 */
#pragma unroll
  for (int jb = 0; jb < 5; ++jb)
  {
    d4 s;
    s.x = __dmul_rn(row20_fma(Ws + k * 400 + (4 * jb + 0) * 20, p), row20_fma(Vs + k * 400 + (4 * jb + 0) * 20, c));
    s.y = __dmul_rn(row20_fma(Ws + k * 400 + (4 * jb + 1) * 20, p), row20_fma(Vs + k * 400 + (4 * jb + 1) * 20, c));
    s.z = __dmul_rn(row20_fma(Ws + k * 400 + (4 * jb + 2) * 20, p), row20_fma(Vs + k * 400 + (4 * jb + 2) * 20, c));
    s.w = __dmul_rn(row20_fma(Ws + k * 400 + (4 * jb + 3) * 20, p), row20_fma(Vs + k * 400 + (4 * jb + 3) * 20, c));
    if (f != 1.0)
    {
      s.x = __dmul_rn(s.x, f);
      s.y = __dmul_rn(s.y, f);
      s.z = __dmul_rn(s.z, f);
      s.w = __dmul_rn(s.w, f);
    }
    *reinterpret_cast<d4 *>(out + 4 * jb) = s;
  }
}

template <int R>
__global__ void __launch_bounds__(PLG_DER_THREADS) k_sumtable_ti_aa(const SumArgs a)
{
  extern __shared__ __align__(16) double sm[];
  double * Vs = sm; /* [R][20][20] */
  for (unsigned int t = threadIdx.x; t < R * 400; t += PLG_DER_THREADS) Vs[t] = __ldg(a.right + t);
  __syncthreads();
  const unsigned int k = threadIdx.x & (R - 1);
  const unsigned int e = blockIdx.x * PLG_DER_THREADS + threadIdx.x;
  const bool valid = e < a.nelem;
  const double f = rate_residual<R>(valid, e, a);
  if (!valid) return;
  double c[20];
  load20d(a.clvp + (size_t)e * 20, c);
  const unsigned int code = __ldg(a.tip + e / R);
  const double * L = a.left + ((size_t)code * R + k) * 20;
  double * out = a.sumtable + (size_t)e * 20;
  /* reference src/core_derivatives_avx2.c:441-512 */
#pragma unroll
  for (int jb = 0; jb < 5; ++jb)
  {
    const d4 l = *reinterpret_cast<const d4 *>(L + 4 * jb);
    d4 s;
    s.x = __dmul_rn(l.x, row20_fma(Vs + k * 400 + (4 * jb + 0) * 20, c));
    s.y = __dmul_rn(l.y, row20_fma(Vs + k * 400 + (4 * jb + 1) * 20, c));
    s.z = __dmul_rn(l.z, row20_fma(Vs + k * 400 + (4 * jb + 2) * 20, c));
    s.w = __dmul_rn(l.w, row20_fma(Vs + k * 400 + (4 * jb + 3) * 20, c));
    if (f != 1.0)
    {
      s.x = __dmul_rn(s.x, f);
      s.y = __dmul_rn(s.y, f);
      s.z = __dmul_rn(s.z, f);
      s.w = __dmul_rn(s.w, f);
    }
    *reinterpret_cast<d4 *>(out + 4 * jb) = s;
  }
}

/* ------------------------------------------------------------------------------------ */
/* derivatives                                                                           */
/* ------------------------------------------------------------------------------------ */
struct DerParams
{
  double invar_lk[PLG_MAX_RATES * 20]; /* [rate][state] = freqs * prop_invar */
  double rate_weights[PLG_MAX_RATES];
  double prop_invar[PLG_MAX_RATES];
  int use_pinv;
  int eq_weights;
};

struct DerArgs
{
  const double * sumtable;
  const double * diagp; /* DNA: [R][4][4] {e, e', e'', 0}; 20 states: [R][3][20] */
  const unsigned int * weights;
  const int * invariant;
  double * partials; /* 2 * gridDim.x */
  unsigned int * counter;
  PlgSink sink; /* d_f, dd_f */
  unsigned int nelem;
};

/* ------------------------------------------------------------------------------------ */
/* DNA derivatives: one thread per pattern                                               */
/* ------------------------------------------------------------------------------------ */
/* The (site, rate)-per-thread mapping above spends most of its instructions on shuffles and
 * predication (measured: issue-bound at 2.7 TB/s).  For 4 states a whole pattern is only
 * R x 32 bytes, so one thread takes a pattern: R 256-bit loads, 12 R FMAs against diagptable
 * entries that are immediate constant-bank operands (kernel parameter), one division.  Same
 * operation order as above / the reference (src/core_derivatives_avx2.c:634-766). */
struct DerDnaParams
{
  double diag[PLG_MAX_RATES * 12]; /* [rate][state][3] */
  double invar_lk[PLG_MAX_RATES * 4];
  double rate_weights[PLG_MAX_RATES];
  double prop_invar[PLG_MAX_RATES];
  int use_pinv;
  int eq_weights;
};

#ifndef PLG_DERDNA_U
#define PLG_DERDNA_U 1
#endif
#ifndef PLG_DERDNA_MINB
#define PLG_DERDNA_MINB 2
#endif
template <int R>
__global__ void __launch_bounds__(PLG_DER_THREADS, (R <= 4) ? PLG_DERDNA_MINB : 1)
k_derivatives_dna(const DerArgs a, const __grid_constant__ DerDnaParams P)
{
  constexpr int U = (R <= 4) ? PLG_DERDNA_U : 1;
  const unsigned int sites = a.nelem / R;
  const unsigned int stride = gridDim.x * PLG_DER_THREADS;
  double df = 0.0, ddf = 0.0;
  for (unsigned int n0 = blockIdx.x * PLG_DER_THREADS + threadIdx.x; n0 < sites; n0 += stride * U)
  {
    d4 sv[U][R];
    unsigned int wgt[U];
    int inv[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const unsigned int n = n0 + u * stride;
      wgt[u] = 0;
      inv[u] = -1;
      if (n < sites)
      {
#pragma unroll
        for (int r = 0; r < R; ++r) sv[u][r] = ld_stream(a.sumtable + ((size_t)n * R + r) * 4);
        wgt[u] = __ldg(a.weights + n);
        if (P.use_pinv && a.invariant) inv[u] = __ldg(a.invariant + n);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      if (n0 + u * stride >= sites) continue;
      double l0 = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r)
      {
        const double s[4] = {sv[u][r].x, sv[u][r].y, sv[u][r].z, sv[u][r].w};
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          v0 = __fma_rn(s[j], P.diag[r * 12 + j * 3 + 0], v0);
          v1 = __fma_rn(s[j], P.diag[r * 12 + j * 3 + 1], v1);
          v2 = __fma_rn(s[j], P.diag[r * 12 + j * 3 + 2], v2);
        }
        if (P.use_pinv && P.prop_invar[r] > 0.0)
        {
          const double q = __dsub_rn(1.0, P.prop_invar[r]);
          v0 = __dmul_rn(v0, q);
          v1 = __dmul_rn(v1, q);
          v2 = __dmul_rn(v2, q);
          if (inv[u] != -1) v0 = __dadd_rn(v0, P.invar_lk[r * 4 + inv[u]]);
        }
        if (P.eq_weights)
        {
          l0 = __dadd_rn(l0, v0);
          l1 = __dadd_rn(l1, v1);
          l2 = __dadd_rn(l2, v2);
        }
        else
        {
          l0 = __fma_rn(v0, P.rate_weights[r], l0);
          l1 = __fma_rn(v1, P.rate_weights[r], l1);
          l2 = __fma_rn(v2, P.rate_weights[r], l2);
        }
      }
      const double recip = __ddiv_rn(1.0, l0);
      const double d1 = __dmul_rn(l1, recip);
      const double d2 = __dsub_rn(__dmul_rn(d1, d1), __dmul_rn(l2, recip));
      const double w = (double)wgt[u];
      df = __dsub_rn(df, __dmul_rn(d1, w));
      ddf = __dadd_rn(ddf, __dmul_rn(d2, w));
    }
  }

  __shared__ double red[PLG_DER_THREADS / 32];
  __shared__ bool is_last;
  const double s1 = block_sum<PLG_DER_THREADS>(df, red);
  const double s2 = block_sum<PLG_DER_THREADS>(ddf, red);
  if (threadIdx.x == 0)
  {
    a.partials[2 * blockIdx.x + 0] = s1;
    a.partials[2 * blockIdx.x + 1] = s2;
    __threadfence();
    is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last)
  {
    __threadfence();
    const unsigned int nb = gridDim.x;
    const unsigned int per = (nb + PLG_DER_THREADS - 1) / PLG_DER_THREADS;
    const unsigned int lo = threadIdx.x * per;
    const unsigned int hi = (lo + per < nb) ? lo + per : nb;
    double t1 = 0.0, t2 = 0.0;
    for (unsigned int b = lo; b < hi; ++b)
    {
      t1 = __dadd_rn(t1, __ldcg(a.partials + 2 * b));
      t2 = __dadd_rn(t2, __ldcg(a.partials + 2 * b + 1));
    }
    const double r1 = block_sum<PLG_DER_THREADS>(t1, red);
    const double r2 = block_sum<PLG_DER_THREADS>(t2, red);
    if (threadIdx.x == 0)
    {
      plg_publish(a.sink, r1, r2);
      *a.counter = 0u;
    }
  }
}

/* Launch configuration of a derivative pass: the sumtable is re-read by every pass of a Newton
 * loop (reference examples/newton/newton.c:64-93: up to 32 calls per branch), so the launch asks
 * L2 to keep as much of it as the device lets persist; the rest streams. */
struct DerLaunch
{
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  DerLaunch(plg_context * ctx, unsigned int nblocks, const double * sumtable, size_t table_bytes)
  {
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(nblocks);
    cfg.blockDim = dim3(PLG_DER_THREADS);
    cfg.stream = ctx->stream;
    cfg.attrs = attr;
    cfg.numAttrs = 0;
    if (ctx->l2_persist_bytes && table_bytes)
    {
      const size_t window = table_bytes < ctx->l2_window_max ? table_bytes : ctx->l2_window_max;
      const double ratio = (double)ctx->l2_persist_bytes / (double)window;
      attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
      attr[0].val.accessPolicyWindow.base_ptr = const_cast<double *>(sumtable);
      attr[0].val.accessPolicyWindow.num_bytes = window;
      attr[0].val.accessPolicyWindow.hitRatio = (float)(ratio < 1.0 ? ratio : 1.0);
      attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cfg.numAttrs = 1;
      ctx->l2_pinned = 1;
    }
  }
};

template <int R>
static int launch_derivatives_dna(plg_context * ctx, const DerArgs & a, const DerDnaParams & P)
{
  /* one resident wave: the grid is fixed for a given device, so results are reproducible */
  static int per_sm = 0;
  if (!per_sm)
    PLG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_derivatives_dna<R>, PLG_DER_THREADS, 0));
  const unsigned int sites = a.nelem / R;
  unsigned int nblocks = (sites + PLG_DER_THREADS - 1) / PLG_DER_THREADS;
  const unsigned int cap = (unsigned int)(ctx->sm_count * (per_sm > 0 ? per_sm : 1));
  if (nblocks > cap) nblocks = cap;
  if (nblocks == 0) nblocks = 1; /* no active pattern: one block still delivers the (zero) sums */
  int rc = plg_ensure_partials(ctx, 2 * (size_t)nblocks);
  if (rc) return rc;
  DerArgs b = a;
  b.partials = ctx->partials;
  DerLaunch launch(ctx, nblocks, b.sumtable, (size_t)b.nelem * 4 * sizeof(double));
  PLG_CUDA(cudaLaunchKernelEx(&launch.cfg, k_derivatives_dna<R>, b, P));
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* 20-state derivatives: one thread per pattern                                          */
/* ------------------------------------------------------------------------------------ */
/* Same mapping as k_derivatives_dna; the transposed diagptable ([rate][3][20], reference
 * src/core_derivatives_avx2.c:598-609) sits in shared memory, read as broadcast.  Per rate and
 * per derivative order: first block by mul, the other four by fma, then the AVX horizontal sum
 * (reference src/core_derivatives_avx2.c:656-702); rates combined in order (:704-729). */
template <int R>
__global__ void __launch_bounds__(PLG_DER_THREADS, 2)
k_derivatives_aa(const DerArgs a, const __grid_constant__ DerParams P)
{
  __shared__ __align__(32) double sd[R * 60];
  for (unsigned int t = threadIdx.x; t < R * 60; t += PLG_DER_THREADS) sd[t] = __ldg(a.diagp + t);
  __syncthreads();
  const unsigned int sites = a.nelem / R;
  const unsigned int stride = gridDim.x * PLG_DER_THREADS;
  double df = 0.0, ddf = 0.0;
  for (unsigned int n = blockIdx.x * PLG_DER_THREADS + threadIdx.x; n < sites; n += stride)
  {
    const unsigned int wgt = __ldg(a.weights + n);
    int inv = -1;
    if (P.use_pinv && a.invariant) inv = __ldg(a.invariant + n);
    double l0 = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r)
    {
      double s[20];
      load20d(a.sumtable + ((size_t)n * R + r) * 20, s);
      double v[3];
#pragma unroll
      for (int x = 0; x < 3; ++x)
      {
        const double * d = sd + r * 60 + x * 20;
        double acc[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) acc[l] = __dmul_rn(s[l], d[l]);
#pragma unroll
        for (int b = 1; b < 5; ++b)
#pragma unroll
          for (int l = 0; l < 4; ++l) acc[l] = __fma_rn(s[4 * b + l], d[4 * b + l], acc[l]);
        v[x] = hsum4(acc[0], acc[1], acc[2], acc[3]);
      }
      if (P.use_pinv && P.prop_invar[r] > 0.0)
      {
        const double q = __dsub_rn(1.0, P.prop_invar[r]);
        v[0] = __dmul_rn(v[0], q);
        v[1] = __dmul_rn(v[1], q);
        v[2] = __dmul_rn(v[2], q);
        if (inv != -1) v[0] = __dadd_rn(v[0], P.invar_lk[r * 20 + inv]);
      }
      if (P.eq_weights)
      {
        l0 = __dadd_rn(l0, v[0]);
        l1 = __dadd_rn(l1, v[1]);
        l2 = __dadd_rn(l2, v[2]);
      }
      else
      {
        l0 = __fma_rn(v[0], P.rate_weights[r], l0);
        l1 = __fma_rn(v[1], P.rate_weights[r], l1);
        l2 = __fma_rn(v[2], P.rate_weights[r], l2);
      }
    }
    const double recip = __ddiv_rn(1.0, l0);
    const double d1 = __dmul_rn(l1, recip);
    const double d2 = __dsub_rn(__dmul_rn(d1, d1), __dmul_rn(l2, recip));
    const double w = (double)wgt;
    df = __dsub_rn(df, __dmul_rn(d1, w));
    ddf = __dadd_rn(ddf, __dmul_rn(d2, w));
  }

  __shared__ double red[PLG_DER_THREADS / 32];
  __shared__ bool is_last;
  const double s1 = block_sum<PLG_DER_THREADS>(df, red);
  const double s2 = block_sum<PLG_DER_THREADS>(ddf, red);
  if (threadIdx.x == 0)
  {
    a.partials[2 * blockIdx.x + 0] = s1;
    a.partials[2 * blockIdx.x + 1] = s2;
    __threadfence();
    is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last)
  {
    __threadfence();
    const unsigned int nb = gridDim.x;
    const unsigned int per = (nb + PLG_DER_THREADS - 1) / PLG_DER_THREADS;
    const unsigned int lo = threadIdx.x * per;
    const unsigned int hi = (lo + per < nb) ? lo + per : nb;
    double t1 = 0.0, t2 = 0.0;
    for (unsigned int b = lo; b < hi; ++b)
    {
      t1 = __dadd_rn(t1, __ldcg(a.partials + 2 * b));
      t2 = __dadd_rn(t2, __ldcg(a.partials + 2 * b + 1));
    }
    const double r1 = block_sum<PLG_DER_THREADS>(t1, red);
    const double r2 = block_sum<PLG_DER_THREADS>(t2, red);
    if (threadIdx.x == 0)
    {
      plg_publish(a.sink, r1, r2);
      *a.counter = 0u;
    }
  }
}

template <int R>
static int launch_derivatives_aa(plg_context * ctx, const DerArgs & a, const DerParams & P)
{
  static int per_sm = 0;
  if (!per_sm)
    PLG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_derivatives_aa<R>, PLG_DER_THREADS, 0));
  const unsigned int sites = a.nelem / R;
  unsigned int nblocks = (sites + PLG_DER_THREADS - 1) / PLG_DER_THREADS;
  const unsigned int cap = (unsigned int)(ctx->sm_count * (per_sm > 0 ? per_sm : 1));
  if (nblocks > cap) nblocks = cap;
  if (nblocks == 0) nblocks = 1; /* no active pattern: one block still delivers the (zero) sums */
  int rc = plg_ensure_partials(ctx, 2 * (size_t)nblocks);
  if (rc) return rc;
  DerArgs b = a;
  b.partials = ctx->partials;
  DerLaunch launch(ctx, nblocks, b.sumtable, (size_t)b.nelem * 20 * sizeof(double));
  PLG_CUDA(cudaLaunchKernelEx(&launch.cfg, k_derivatives_aa<R>, b, P));
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* host side                                                                             */
/* ------------------------------------------------------------------------------------ */
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

/* Device slot of the sumtable the caller identifies by `key` (its host pointer).  Slots live
 * until plg_free_sumtable / plg_destroy; a caller that allocates and frees a host sumtable per
 * branch without telling us leaves slots behind, so when HBM runs out the least recently used
 * slots are reclaimed before giving up. */
static int sumtable_slot(plg_context * ctx, const void * key, double ** out)
{
  (*ctx->sumtable_used)[key] = ++ctx->sumtable_clock;
  auto it = ctx->sumtables->find(key);
  if (it != ctx->sumtables->end())
  {
    *out = it->second;
    return PLG_OK;
  }
  double * dev = NULL;
  const size_t bytes = (size_t)ctx->d.sites * ctx->span * sizeof(double);
  cudaError_t err = cudaMalloc(&dev, bytes);
  while (err == cudaErrorMemoryAllocation && !ctx->sumtables->empty())
  {
    cudaGetLastError();
    const void * victim = NULL;
    unsigned long long oldest = ~0ull;
    for (const auto & kv : *ctx->sumtables)
    {
      const unsigned long long used = (*ctx->sumtable_used)[kv.first];
      if (used < oldest) { oldest = used; victim = kv.first; }
    }
    PLG_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree((*ctx->sumtables)[victim]);
    ctx->sumtables->erase(victim);
    ctx->sumtable_used->erase(victim);
    err = cudaMalloc(&dev, bytes);
  }
  if (err != cudaSuccess)
  {
    cudaGetLastError();
    plg_set_error("plg_update_sumtable: cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(err));
    return err == cudaErrorMemoryAllocation ? PLG_E_NOMEM : PLG_E_CUDA;
  }
  (*ctx->sumtables)[key] = dev;
  *out = dev;
  return PLG_OK;
}

extern "C" int plg_update_sumtable(plg_context_t * ctx, unsigned int parent_clv_index,
                                   unsigned int child_clv_index, int parent_scaler_index,
                                   int child_scaler_index, const double * eigenvecs,
                                   const double * left_table, const void * key,
                                   double * host_copy)
{
  PLG_CHECK_CTX(ctx);
  plg_release_l2(ctx);
  const unsigned int n_clv = ctx->d.tips + ctx->d.clv_buffers;
  if (parent_clv_index >= n_clv || child_clv_index >= n_clv ||
      parent_scaler_index >= (int)ctx->d.scale_buffers ||
      child_scaler_index >= (int)ctx->d.scale_buffers || !key)
  {
    plg_set_error("plg_update_sumtable: index out of range");
    return PLG_E_INVALID;
  }
  const bool ptip = plg_is_tip(ctx, parent_clv_index);
  const bool ctip = plg_is_tip(ctx, child_clv_index);
  if (ptip && ctip)
  {
    plg_set_error("plg_update_sumtable: tip-tip edge (the reference asserts, "
                  "src/derivatives.c:195)");
    return PLG_E_UNSUPPORTED;
  }
  const unsigned int R = ctx->d.rate_cats, K = ctx->d.states;
  if (!plg_fast_path(ctx))
  {
    const unsigned int Kp = ctx->d.states_padded;
    double * table = NULL;
    int grc = sumtable_slot(ctx, key, &table);
    if (grc) return grc;
    const bool gti = ptip || ctip;
    const size_t gmat = (size_t)R * K * Kp * sizeof(double);
    const size_t gleft = gti ? (size_t)(K == 4 ? 16u : ctx->maxstates) * R * Kp * sizeof(double) : gmat;
    if (plg_stage_reserve(ctx, gmat + gleft + 1024)) return PLG_E_CUDA;
    const double * d_evecs = (const double *)plg_stage(ctx, eigenvecs, gmat);
    const double * d_left = (const double *)plg_stage(ctx, left_table, gleft);
    if (!d_evecs || !d_left) return PLG_E_CUDA;
    if (gti)
      grc = plg_gen_sumtable(ctx, plg_clv_ptr(ctx, ptip ? child_clv_index : parent_clv_index), NULL,
                             plg_tip_ptr(ctx, ptip ? parent_clv_index : child_clv_index),
                             plg_scaler_ptr(ctx, ptip ? child_scaler_index : parent_scaler_index), NULL,
                             d_evecs, d_left, table);
    else
      grc = plg_gen_sumtable(ctx, plg_clv_ptr(ctx, parent_clv_index), plg_clv_ptr(ctx, child_clv_index), NULL,
                             plg_scaler_ptr(ctx, parent_scaler_index), plg_scaler_ptr(ctx, child_scaler_index),
                             d_evecs, d_left, table);
    if (grc) return grc;
    if (host_copy)
    {
      const size_t bytes = (size_t)ctx->d.sites * ctx->span * sizeof(double);
      PLG_CUDA(cudaMemcpyAsync(host_copy, table, bytes, cudaMemcpyDeviceToHost, ctx->stream));
      PLG_CUDA(cudaStreamSynchronize(ctx->stream));
      ctx->stats.d2h_bytes += bytes;
    }
    return PLG_OK;
  }
  const unsigned int nelem = ctx->d.sites * R;
  const unsigned int nblocks = (nelem + PLG_DER_THREADS - 1) / PLG_DER_THREADS;
  const unsigned int nblocks_dna =
      (nelem + PLG_DER_THREADS * PLG_SUM_ITEMS - 1) / (PLG_DER_THREADS * PLG_SUM_ITEMS);

  SumArgs a;
  memset(&a, 0, sizeof(a));
  int rc = sumtable_slot(ctx, key, &a.sumtable);
  if (rc) return rc;
  a.nelem = nelem;
  a.per_rate_scaling = ctx->rate_scalers ? 1 : 0;

  const size_t mat_bytes = (size_t)R * K * K * sizeof(double);
  const bool ti = ptip || ctip;
  const size_t codes = (K == 4) ? 16u : ctx->maxstates;
  const size_t left_bytes = ti ? codes * R * K * sizeof(double) : mat_bytes;
  if (plg_stage_reserve(ctx, mat_bytes + left_bytes + 2048)) return PLG_E_CUDA;
  a.right = (const double *)plg_stage(ctx, eigenvecs, mat_bytes);
  a.left = (const double *)plg_stage(ctx, left_table, left_bytes);
  if (!a.right || !a.left) return PLG_E_CUDA;

  /* Without per-rate residuals the table is one CLV-update-shaped operation: DNA inner-inner
   * is bit-identical to the streaming kernel's arithmetic; 20 states go to the tensor-core
   * kernels (unless the bit-exact vector path was asked for).  The DNA tip-inner table keeps
   * its own kernel: the reference sums the child states sequentially there. */
  if (!ctx->rate_scalers && ((K == 4 && !ti) || (K == 20 && !ctx->aa_exact)))
  {
    DevOp op;
    memset(&op, 0, sizeof(op));
    op.parent = a.sumtable;
    if (ti)
    {
      op.right = plg_clv_ptr(ctx, ptip ? child_clv_index : parent_clv_index);
      op.ltip = plg_tip_ptr(ctx, ptip ? parent_clv_index : child_clv_index);
    }
    else
    {
      op.left = plg_clv_ptr(ctx, parent_clv_index);
      op.right = plg_clv_ptr(ctx, child_clv_index);
    }
    op.lmat = a.left;
    op.rmat = a.right;
    rc = plg_launch_single_op(ctx, ti ? PLG_KIND_TI : PLG_KIND_II, op);
    if (rc) return rc;
    if (host_copy)
    {
      const size_t bytes = (size_t)ctx->d.sites * ctx->span * sizeof(double);
      PLG_CUDA(cudaMemcpyAsync(host_copy, a.sumtable, bytes, cudaMemcpyDeviceToHost, ctx->stream));
      PLG_CUDA(cudaStreamSynchronize(ctx->stream));
      ctx->stats.d2h_bytes += bytes;
    }
    return PLG_OK;
  }

  if (ti)
  {
    const unsigned int inner = ptip ? child_clv_index : parent_clv_index;
    const unsigned int tip = ptip ? parent_clv_index : child_clv_index;
    a.clvp = plg_clv_ptr(ctx, inner);
    a.tip = plg_tip_ptr(ctx, tip);
    a.pscale = plg_scaler_ptr(ctx, ptip ? child_scaler_index : parent_scaler_index);
    if (K == 4)
    {
      PLG_DISPATCH_R(R, (k_sumtable_ti_dna<RR><<<nblocks_dna, PLG_DER_THREADS, 0, ctx->stream>>>(a)));
    }
    else
    {
      const size_t smem = (size_t)R * 400 * sizeof(double);
      PLG_DISPATCH_R(R, {
        static bool attr_done[PLG_MAX_DEVICES] = {}; /* a function attribute is per device */
        if (!attr_done[ctx->device % PLG_MAX_DEVICES])
        {
          PLG_CUDA(cudaFuncSetAttribute(k_sumtable_ti_aa<RR>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          attr_done[ctx->device % PLG_MAX_DEVICES] = true;
        }
        k_sumtable_ti_aa<RR><<<nblocks, PLG_DER_THREADS, smem, ctx->stream>>>(a);
      });
    }
  }
  else
  {
    a.clvp = plg_clv_ptr(ctx, parent_clv_index);
    a.clvc = plg_clv_ptr(ctx, child_clv_index);
    a.pscale = plg_scaler_ptr(ctx, parent_scaler_index);
    a.cscale = plg_scaler_ptr(ctx, child_scaler_index);
    if (K == 4)
    {
      const unsigned int nblocks_ii =
          (nelem + PLG_DER_THREADS * PLG_SUM_ITEMS_II - 1) / (PLG_DER_THREADS * PLG_SUM_ITEMS_II);
      PLG_DISPATCH_R(R, (k_sumtable_ii_dna<RR><<<nblocks_ii, PLG_DER_THREADS, 0, ctx->stream>>>(a)));
    }
    else
    {
      const size_t smem = (size_t)2 * R * 400 * sizeof(double);
      PLG_DISPATCH_R(R, {
        static bool attr_done[PLG_MAX_DEVICES] = {}; /* a function attribute is per device */
        if (!attr_done[ctx->device % PLG_MAX_DEVICES])
        {
          PLG_CUDA(cudaFuncSetAttribute(k_sumtable_ii_aa<RR>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          attr_done[ctx->device % PLG_MAX_DEVICES] = true;
        }
        k_sumtable_ii_aa<RR><<<nblocks, PLG_DER_THREADS, smem, ctx->stream>>>(a);
      });
    }
  }
  PLG_LAUNCH_CHECK(ctx);

  if (host_copy)
  {
    const size_t bytes = (size_t)ctx->d.sites * ctx->span * sizeof(double);
    PLG_CUDA(cudaMemcpyAsync(host_copy, a.sumtable, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PLG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += bytes;
  }
  return PLG_OK;
}

/* A sumtable the caller holds on the host (sites x rate_cats x states_padded doubles, the layout
 * plg_update_sumtable writes) becomes the device slot of `key`: the input side of the reference's
 * pll_core_likelihood_derivatives (src/pll.h:943-961), which receives the table as an argument. */
extern "C" int plg_set_sumtable(plg_context_t * ctx, const void * key, const double * table)
{
  PLG_CHECK_CTX(ctx);
  if (!key || !table)
  {
    plg_set_error("plg_set_sumtable: null argument");
    return PLG_E_INVALID;
  }
  plg_release_l2(ctx);
  double * dev = NULL;
  int rc = sumtable_slot(ctx, key, &dev);
  if (rc) return rc;
  const size_t bytes = (size_t)ctx->d.sites * ctx->span * sizeof(double);
  PLG_CUDA(cudaMemcpyAsync(dev, table, bytes, cudaMemcpyHostToDevice, ctx->stream));
  PLG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stats.h2d_bytes += bytes;
  return PLG_OK;
}

extern "C" int plg_free_sumtable(plg_context_t * ctx, const void * key)
{
  PLG_CHECK_CTX(ctx);
  auto it = ctx->sumtables->find(key);
  if (it == ctx->sumtables->end()) return PLG_OK;
  PLG_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(it->second);
  ctx->sumtables->erase(it);
  ctx->sumtable_used->erase(key);
  return PLG_OK;
}

extern "C" int plg_likelihood_derivatives(plg_context_t * ctx, const void * key,
                                          const double * diagptable, const double * rate_weights,
                                          const double * prop_invar, const double * freqs,
                                          double * d_f, double * dd_f)
{
  PLG_CHECK_CTX(ctx);
  auto it = ctx->sumtables->find(key);
  if (it == ctx->sumtables->end())
  {
    plg_set_error("plg_likelihood_derivatives: no sumtable was computed for this buffer "
                  "(call pll_update_sumtable first)");
    return PLG_E_INVALID;
  }
  (*ctx->sumtable_used)[key] = ++ctx->sumtable_clock;
  if (!plg_fast_path(ctx))
    return plg_gen_derivatives(ctx, it->second, diagptable, rate_weights, prop_invar, freqs, d_f, dd_f);
  const unsigned int R = ctx->d.rate_cats, K = ctx->d.states;
  const unsigned int nelem = ctx->active_sites * R;
  unsigned int nblocks = (nelem + PLG_DER_THREADS - 1) / PLG_DER_THREADS;
  if (nblocks == 0) nblocks = 1; /* no active pattern: one block still delivers the (zero) sums */
  /* persistent: exactly the CTAs that are resident at once (one wave) */
  const unsigned int cap = (unsigned int)ctx->sm_count * 2u;
  if (nblocks > cap) nblocks = cap;
  int rc = plg_ensure_partials(ctx, 2 * (size_t)nblocks);
  if (rc) return rc;

  DerParams P;
  memset(&P, 0, sizeof(P));
  P.use_pinv = 0;
  P.eq_weights = 1;
  for (unsigned int i = 0; i < R; ++i)
  {
    P.rate_weights[i] = rate_weights[i];
    P.prop_invar[i] = prop_invar[i];
    if (prop_invar[i] > 0) P.use_pinv = 1;
    if (rate_weights[i] != rate_weights[0]) P.eq_weights = 0;
    for (unsigned int s = 0; s < K; ++s)
      P.invar_lk[i * K + s] = freqs[i * ctx->d.states_padded + s] * prop_invar[i];
  }
  if (P.use_pinv && !ctx->has_invariant)
  {
    plg_set_error("derivatives with prop_invar > 0 need the invariant-site index");
    return PLG_E_INVALID;
  }

  if (K == 4)
  {
    /* everything the kernel needs besides the sumtable travels as kernel parameters */
    DerDnaParams D;
    memset(&D, 0, sizeof(D));
    D.use_pinv = P.use_pinv;
    D.eq_weights = P.eq_weights;
    for (unsigned int i = 0; i < R; ++i)
    {
      D.rate_weights[i] = P.rate_weights[i];
      D.prop_invar[i] = P.prop_invar[i];
      for (unsigned int j = 0; j < 4; ++j)
      {
        D.invar_lk[i * 4 + j] = P.invar_lk[i * 4 + j];
        for (unsigned int x = 0; x < 3; ++x) D.diag[i * 12 + j * 3 + x] = diagptable[(size_t)i * 16 + j * 4 + x];
      }
    }
    DerArgs da;
    da.sumtable = it->second;
    da.diagp = NULL;
    da.weights = ctx->weights;
    da.invariant = ctx->has_invariant ? ctx->invariant : NULL;
    da.partials = ctx->partials;
    da.counter = ctx->counter;
    da.sink = plg_make_sink(ctx);
    da.nelem = nelem;
    PLG_DISPATCH_R(R, { int lrc = launch_derivatives_dna<RR>(ctx, da, D); if (lrc) return lrc; });
    PLG_LAUNCH_CHECK(ctx);
    ctx->stats.d2h_bytes += 2 * sizeof(double);
    return plg_finish_result(ctx, d_f, dd_f);
  }

  /* device layout of diagp: 20 states is transposed to [R][3][20] exactly like the
   * reference's t_diagp (src/core_derivatives_avx2.c:598-609) */
  double diag_host[PLG_MAX_RATES * 60];
  const size_t diag_len = (size_t)R * 60;
  for (unsigned int i = 0; i < R; ++i)
    for (unsigned int j = 0; j < K; ++j)
      for (unsigned int x = 0; x < 3; ++x)
        diag_host[i * 60 + x * 20 + j] = diagptable[(size_t)i * K * 4 + j * 4 + x];

  DerArgs a;
  a.sumtable = it->second;
  a.diagp = (const double *)plg_stage(ctx, diag_host, diag_len * sizeof(double));
  if (!a.diagp) return PLG_E_CUDA;
  a.weights = ctx->weights;
  a.invariant = ctx->has_invariant ? ctx->invariant : NULL;
  a.partials = ctx->partials;
  a.counter = ctx->counter;
  a.sink = plg_make_sink(ctx);
  a.nelem = nelem;

  PLG_DISPATCH_R(R, { int lrc = launch_derivatives_aa<RR>(ctx, a, P); if (lrc) return lrc; });
  PLG_LAUNCH_CHECK(ctx);

  ctx->stats.d2h_bytes += 2 * sizeof(double);
  return plg_finish_result(ctx, d_f, dd_f);
}

extern "C" int plg_get_sumtable_sites(plg_context_t * ctx, const void * key, unsigned int first_site,
                                      unsigned int count, double * out)
{
  PLG_CHECK_CTX(ctx);
  auto it = ctx->sumtables->find(key);
  if (it == ctx->sumtables->end() || (size_t)first_site + count > ctx->d.sites || !out)
  {
    plg_set_error("plg_get_sumtable_sites: unknown sumtable or range out of bounds");
    return PLG_E_INVALID;
  }
  const size_t bytes = (size_t)count * ctx->span * sizeof(double);
  PLG_CUDA(cudaMemcpyAsync(out, it->second + (size_t)first_site * ctx->span, bytes, cudaMemcpyDeviceToHost,
                           ctx->stream));
  PLG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stats.d2h_bytes += bytes;
  return PLG_OK;
}
