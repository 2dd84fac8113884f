/*
 * plg_generic.cu - the likelihood path for ANY number of states and rate categories.
 *
 * The specialised kernels (plg_partials.cu, plg_likelihood.cu, plg_derivatives.cu,
 * plg_pmatrix.cu) cover 4 and 20 states with 1/2/4/8/16 rate categories - the BASELINE
 * configurations.  Everything else the reference accepts (binary, odd state counts such as
 * the 5- and 7-state cases of reference test/src/00012_NMOU_lkcalc.c and
 * derivatives-oddstates.c, codon-sized alphabets, 3 or 5 categories ...) runs here: one thread
 * per alignment pattern, plain loops over rates and states, matrices read through L1.  These
 * kernels follow the operation order of the reference's generic C code (the semantic
 * definition of the path: reference src/core_partials.c:510-663, src/core_likelihood.c:163-209
 * and :940-1000, src/core_derivatives.c:240-262,449-732, src/core_pmatrix.c:146-250); padded
 * states (states..states_padded-1) are written as exact zeros.  Correct first, not tuned.
 */
#include <cmath>

#include "plg_internal.cuh"

#define PLG_GEN_THREADS 128

/* ------------------------------------------------------------------------------------ */
__global__ void k_gen_pmatrix(double * __restrict__ pmatrix, size_t pmat_len,
                              const unsigned int * __restrict__ matrix_indices,
                              const double * __restrict__ branch_lengths, unsigned int count,
                              unsigned int R, unsigned int K, unsigned int Kp,
                              const double * __restrict__ evals, const double * __restrict__ evecs,
                              const double * __restrict__ ievecs, const double * __restrict__ rates,
                              const double * __restrict__ pinvs)
{
  /* one thread per (branch, rate, row j, column c) */
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int c = tid % K, j = (tid / K) % K, n = (tid / ((size_t)K * K)) % R;
  const size_t i = tid / ((size_t)K * K * R);
  if (i >= count) return;
  const double t = branch_lengths[i];
  double * out = pmatrix + (size_t)matrix_indices[i] * pmat_len + (size_t)n * K * Kp + (size_t)j * Kp + c;
  if (t == 0.0)
  {
    *out = (j == c) ? 1.0 : 0.0;
    return;
  }
  const double pinv = pinvs[n];
  double s = (j == c) ? 1.0 : 0.0;
  for (unsigned int m = 0; m < K; ++m)
  {
    double x = __dmul_rn(__dmul_rn(evals[n * Kp + m], rates[n]), t);
    if (pinv > PLL_MISC_EPSILON) x = __ddiv_rn(x, __dsub_rn(1.0, pinv));
    const double tmp = __dmul_rn(ievecs[(size_t)n * K * Kp + j * Kp + m], expm1(x));
    s = __dadd_rn(s, __dmul_rn(tmp, evecs[(size_t)n * K * Kp + m * Kp + c]));
  }
  *out = s;
}

int plg_gen_pmatrix(plg_context * ctx, const unsigned int * d_idx, const double * d_bl, unsigned int count,
                    const double * evals, const double * evecs, const double * ievecs, const double * rates,
                    const double * pinv)
{
  const unsigned int R = ctx->d.rate_cats, K = ctx->d.states, Kp = ctx->d.states_padded;
  const size_t threads = (size_t)count * R * K * K;
  k_gen_pmatrix<<<(unsigned int)((threads + 127) / 128), 128, 0, ctx->stream>>>(
      ctx->pmatrix, ctx->pmat_len, d_idx, d_bl, count, R, K, Kp, evals, evecs, ievecs, rates, pinv);
  PLG_LAUNCH_CHECK(ctx);
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* table[code][rate][i] = sum over the states m in the code's mask of P_rate[i][m]
 * (pattern tips; DNA codes are the masks themselves, other alphabets go through tipmap) */
__global__ void k_gen_tables(const TableJob * __restrict__ jobs, unsigned int R, unsigned int K,
                             unsigned int Kp, unsigned int codes, int use_map, const TipmapArg tm)
{
  const TableJob job = jobs[blockIdx.x];
  const unsigned int entries = codes * R * Kp;
  for (unsigned int t = threadIdx.x; t < entries; t += blockDim.x)
  {
    const unsigned int i = t % Kp, k = (t / Kp) % R, code = t / (R * Kp);
    double s = 0.0;
    if (i < K)
    {
      const unsigned int state = use_map ? tm.map[code] : code;
      const double * row = job.pmat + (size_t)k * K * Kp + (size_t)i * Kp;
      for (unsigned int m = 0; m < K; ++m)
        if ((state >> m) & 1u) s = __dadd_rn(s, row[m]);
    }
    job.out[t] = s;
  }
}

int plg_gen_tables(plg_context * ctx, const TableJob * dev_jobs, unsigned int njobs)
{
  TipmapArg tm;
  memcpy(tm.map, ctx->tipmap, sizeof(tm.map));
  const unsigned int codes = (ctx->d.states == 4) ? 16u : ctx->maxstates;
  k_gen_tables<<<njobs, 256, 0, ctx->stream>>>(dev_jobs, ctx->d.rate_cats, ctx->d.states,
                                               ctx->d.states_padded, codes, ctx->d.states != 4, tm);
  PLG_LAUNCH_CHECK(ctx);
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* CLV update, one thread per pattern (reference src/core_partials.c:604-662 and the tt/ti
 * variants :82-200, :354-508) */
__global__ void __launch_bounds__(PLG_GEN_THREADS)
k_gen_partial(const DevOp * __restrict__ ops, unsigned int sites, unsigned int R, unsigned int K,
              unsigned int Kp, int kind, int scale_mode)
{
  const DevOp op = ops[blockIdx.y];
  const unsigned int n = blockIdx.x * PLG_GEN_THREADS + threadIdx.x;
  if (n >= sites) return;
  const size_t span = (size_t)R * Kp;
  double * parent = op.parent + (size_t)n * span;
  const double * l = (kind == PLG_KIND_II) ? op.left + (size_t)n * span : nullptr;
  const double * r = (kind != PLG_KIND_TT) ? op.right + (size_t)n * span : nullptr;
  const double * tl = (kind != PLG_KIND_II) ? op.lmat + (size_t)op.ltip[n] * span : nullptr;
  const double * tr = (kind == PLG_KIND_TT) ? op.rmat + (size_t)op.rtip[n] * span : nullptr;

  bool site_below = true;
  for (unsigned int k = 0; k < R; ++k)
  {
    bool rate_below = true;
    for (unsigned int i = 0; i < K; ++i)
    {
      double x, y;
      if (kind == PLG_KIND_II)
      {
        const double * row = op.lmat + (size_t)k * K * Kp + (size_t)i * Kp;
        x = 0.0;
        for (unsigned int j = 0; j < K; ++j) x = __dadd_rn(x, __dmul_rn(row[j], l[k * Kp + j]));
      }
      else
        x = tl[k * Kp + i];
      if (kind == PLG_KIND_TT)
        y = tr[k * Kp + i];
      else
      {
        const double * row = op.rmat + (size_t)k * K * Kp + (size_t)i * Kp;
        y = 0.0;
        for (unsigned int j = 0; j < K; ++j) y = __dadd_rn(y, __dmul_rn(row[j], r[k * Kp + j]));
      }
      const double p = __dmul_rn(x, y);
      parent[k * Kp + i] = p;
      rate_below = rate_below && (p < PLG_SCALE_THRESHOLD);
    }
    for (unsigned int i = K; i < Kp; ++i) parent[k * Kp + i] = 0.0;
    if (kind != PLG_KIND_TT && scale_mode == 2)
    {
      unsigned int s = rate_below ? 1u : 0u;
      const size_t e = (size_t)n * R + k;
      if (op.lscale) s += op.lscale[e];
      if (op.rscale) s += op.rscale[e];
      op.pscale[e] = s;
      if (rate_below)
        for (unsigned int i = 0; i < K; ++i) parent[k * Kp + i] = __dmul_rn(parent[k * Kp + i], PLG_SCALE_FACTOR);
    }
    site_below = site_below && rate_below;
  }
  if (kind == PLG_KIND_TT)
  {
    /* tip-tip never scales and zeroes the parent scaler (reference src/core_partials.c:113-116) */
    if (scale_mode == 1) op.pscale[n] = 0u;
    else if (scale_mode == 2)
      for (unsigned int k = 0; k < R; ++k) op.pscale[(size_t)n * R + k] = 0u;
  }
  else if (scale_mode == 1)
  {
    unsigned int s = site_below ? 1u : 0u;
    if (op.lscale) s += op.lscale[n];
    if (op.rscale) s += op.rscale[n];
    op.pscale[n] = s;
    if (site_below)
      for (unsigned int k = 0; k < R; ++k)
        for (unsigned int i = 0; i < K; ++i) parent[k * Kp + i] = __dmul_rn(parent[k * Kp + i], PLG_SCALE_FACTOR);
  }
}

int plg_gen_partials(plg_context * ctx, int kind, int scale_mode, const DevOp * dev_ops, unsigned int count)
{
  dim3 grid((ctx->d.sites + PLG_GEN_THREADS - 1) / PLG_GEN_THREADS, count);
  k_gen_partial<<<grid, PLG_GEN_THREADS, 0, ctx->stream>>>(dev_ops, ctx->d.sites, ctx->d.rate_cats,
                                                          ctx->d.states, ctx->d.states_padded, kind, scale_mode);
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* deterministic sum of one (or two) per-thread values over the grid: block tree, then the
 * last block to finish adds the block partials in order */
template <int NV>
__device__ __forceinline__ void gen_finish(double v0, double v1, double * partials, unsigned int * counter,
                                           const PlgSink & sink)
{
  __shared__ double red[PLG_GEN_THREADS / 32];
  __shared__ bool is_last;
  const double s0 = block_sum<PLG_GEN_THREADS>(v0, red);
  const double s1 = (NV == 2) ? block_sum<PLG_GEN_THREADS>(v1, red) : 0.0;
  if (threadIdx.x == 0)
  {
    partials[NV * blockIdx.x] = s0;
    if (NV == 2) partials[NV * blockIdx.x + 1] = s1;
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last)
  {
    __threadfence();
    const unsigned int nb = gridDim.x, per = (nb + PLG_GEN_THREADS - 1) / PLG_GEN_THREADS;
    const unsigned int lo = threadIdx.x * per;
    const unsigned int hi = (lo + per < nb) ? lo + per : nb;
    double t0 = 0.0, t1 = 0.0;
    for (unsigned int b = lo; b < hi; ++b)
    {
      t0 = __dadd_rn(t0, __ldcg(partials + NV * b));
      if (NV == 2) t1 = __dadd_rn(t1, __ldcg(partials + NV * b + 1));
    }
    const double r0 = block_sum<PLG_GEN_THREADS>(t0, red);
    const double r1 = (NV == 2) ? block_sum<PLG_GEN_THREADS>(t1, red) : 0.0;
    if (threadIdx.x == 0)
    {
      plg_publish(sink, r0, (NV == 2) ? r1 : 0.0);
      *counter = 0u;
    }
  }
}

struct GenLnlDev
{
  GenLnl g;
  const double * freqs;        /* [R][Kp] */
  const double * rate_weights; /* [R] */
  const double * prop_invar;   /* [R] */
  const unsigned int * weights;
  const int * invariant;
  const unsigned int * tipmap; /* device copy, [256] */
  double * persite;
  double * partials;
  unsigned int * counter;
  PlgSink sink;
  double log_threshold;
  unsigned int sites, R, K, Kp;
  int per_rate, use_map;
};

/* edge (ii / ti) and root lnL, one thread per pattern (reference src/core_likelihood.c:163-209,
 * :320-410 generic ti, :940-1000 generic ii) */
__global__ void __launch_bounds__(PLG_GEN_THREADS) k_gen_lnl(const GenLnlDev a)
{
  const unsigned int n = blockIdx.x * PLG_GEN_THREADS + threadIdx.x;
  double site_lk = 0.0;
  if (n < a.sites)
  {
    const unsigned int R = a.R, K = a.K, Kp = a.Kp;
    const size_t span = (size_t)R * Kp;
    /* per-rate scalers: site scaler = min over rates, capped residuals per rate */
    unsigned int site_scalings = 0;
    if (a.per_rate && !a.g.root)
    {
      unsigned int mn = 0xffffffffu;
      for (unsigned int k = 0; k < R; ++k)
      {
        const unsigned int s = (a.g.pscale ? a.g.pscale[(size_t)n * R + k] : 0u) +
                               (a.g.cscale ? a.g.cscale[(size_t)n * R + k] : 0u);
        mn = s < mn ? s : mn;
      }
      site_scalings = mn;
    }
    else
      site_scalings = (a.g.pscale ? a.g.pscale[n] : 0u) + (a.g.cscale ? a.g.cscale[n] : 0u);

    const unsigned int tipstate = a.g.tip ? (a.use_map ? a.tipmap[a.g.tip[n]] : a.g.tip[n]) : 0u;
    double term = 0.0;
    for (unsigned int k = 0; k < R; ++k)
    {
      const double * f = a.freqs + (size_t)k * Kp;
      const double * p = a.g.clvp + (size_t)n * span + (size_t)k * Kp;
      double term_r = 0.0;
      for (unsigned int j = 0; j < K; ++j)
      {
        double tb;
        if (a.g.root)
          tb = 1.0;
        else
        {
          const double * row = a.g.pmat + (size_t)k * K * Kp + (size_t)j * Kp;
          tb = 0.0;
          if (a.g.tip)
          {
            for (unsigned int m = 0; m < K; ++m)
              if ((tipstate >> m) & 1u) tb = __dadd_rn(tb, row[m]);
          }
          else
          {
            const double * c = a.g.clvc + (size_t)n * span + (size_t)k * Kp;
            for (unsigned int m = 0; m < K; ++m) tb = __dadd_rn(tb, __dmul_rn(row[m], c[m]));
          }
        }
        term_r = a.g.root ? __dadd_rn(term_r, __dmul_rn(p[j], f[j]))
                          : __dadd_rn(term_r, __dmul_rn(__dmul_rn(p[j], f[j]), tb));
      }
      if (a.per_rate && !a.g.root)
      {
        unsigned int d = (a.g.pscale ? a.g.pscale[(size_t)n * R + k] : 0u) +
                         (a.g.cscale ? a.g.cscale[(size_t)n * R + k] : 0u) - site_scalings;
        if (d > PLL_SCALE_RATE_MAXDIFF) d = PLL_SCALE_RATE_MAXDIFF;
        for (unsigned int q = 0; q < d; ++q) term_r = __dmul_rn(term_r, PLG_SCALE_THRESHOLD);
      }
      const double pinv = a.prop_invar[k];
      if (pinv > 0.0)
      {
        const int inv = a.invariant[n];
        const double inv_lk = (inv == -1) ? 0.0 : f[inv];
        term = __dadd_rn(term, __dmul_rn(a.rate_weights[k],
                                         __dadd_rn(__dmul_rn(term_r, __dsub_rn(1.0, pinv)), __dmul_rn(inv_lk, pinv))));
      }
      else
        term = __dadd_rn(term, __dmul_rn(term_r, a.rate_weights[k]));
    }
    site_lk = log(term);
    if (site_scalings) site_lk = __dadd_rn(site_lk, __dmul_rn((double)site_scalings, a.log_threshold));
    site_lk = __dmul_rn(site_lk, (double)a.weights[n]);
    if (a.persite) a.persite[n] = site_lk;
  }
  gen_finish<1>(site_lk, 0.0, a.partials, a.counter, a.sink);
}

int plg_gen_loglikelihood(plg_context * ctx, const GenLnl & g, const double * freqs, const double * rate_weights,
                          const double * prop_invar, double * persite_lnl, double * logl_out)
{
  const unsigned int R = ctx->d.rate_cats, Kp = ctx->d.states_padded;
  const unsigned int nblocks = ctx->active_sites ? (ctx->active_sites + PLG_GEN_THREADS - 1) / PLG_GEN_THREADS : 1u;
  int rc = plg_ensure_partials(ctx, nblocks);
  if (rc) return rc;
  bool any_pinv = false;
  for (unsigned int i = 0; i < R; ++i) any_pinv |= (prop_invar && prop_invar[i] > 0);
  if (any_pinv && !ctx->has_invariant)
  {
    plg_set_error("log-likelihood with prop_invar > 0 needs the invariant-site index");
    return PLG_E_INVALID;
  }
  if (persite_lnl && !ctx->persite_dev)
    PLG_CUDA(cudaMalloc(&ctx->persite_dev, (size_t)ctx->active_sites * sizeof(double)));
  std::vector<double> pinv(R, 0.0);
  if (prop_invar) pinv.assign(prop_invar, prop_invar + R);
  if (plg_stage_reserve(ctx, ((size_t)R * Kp + 2 * R) * 8 + 1024 + 4 * 256)) return PLG_E_CUDA;
  GenLnlDev a;
  memset(&a, 0, sizeof(a));
  a.g = g;
  a.freqs = (const double *)plg_stage(ctx, freqs, (size_t)R * Kp * sizeof(double));
  a.rate_weights = (const double *)plg_stage(ctx, rate_weights, R * sizeof(double));
  a.prop_invar = (const double *)plg_stage(ctx, pinv.data(), R * sizeof(double));
  a.tipmap = (const unsigned int *)plg_stage(ctx, ctx->tipmap, sizeof(ctx->tipmap));
  if (!a.freqs || !a.rate_weights || !a.prop_invar || !a.tipmap) return PLG_E_CUDA;
  a.weights = ctx->weights;
  a.invariant = ctx->invariant;
  a.persite = persite_lnl ? ctx->persite_dev : NULL;
  a.partials = ctx->partials;
  a.counter = ctx->counter;
  a.sink = plg_make_sink(ctx);
  a.log_threshold = log(PLL_SCALE_THRESHOLD);
  a.sites = ctx->active_sites;
  a.R = R;
  a.K = ctx->d.states;
  a.Kp = Kp;
  a.per_rate = ctx->rate_scalers ? 1 : 0;
  a.use_map = ctx->d.states != 4;
  k_gen_lnl<<<nblocks, PLG_GEN_THREADS, 0, ctx->stream>>>(a);
  PLG_LAUNCH_CHECK(ctx);
  if (persite_lnl)
  {
      PLG_CUDA(cudaMemcpyAsync(persite_lnl, ctx->persite_dev, (size_t)ctx->active_sites * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    ctx->copy_pending = 1;
  }
  ctx->stats.d2h_bytes += sizeof(double) + (persite_lnl ? (size_t)ctx->active_sites * sizeof(double) : 0);
  return plg_finish_result(ctx, logl_out, NULL);
}

/* ------------------------------------------------------------------------------------ */
/* sumtable: sum[n][k][j] = left_j * right_j (reference src/core_derivatives.c:240-262,
 * :400-440); `left` is the host-built W (ii) or per-code table (ti), see pll_derivatives.c */
__global__ void __launch_bounds__(PLG_GEN_THREADS)
k_gen_sumtable(const double * __restrict__ clvp, const double * __restrict__ clvc,
               const unsigned char * __restrict__ tip, const unsigned int * __restrict__ pscale,
               const unsigned int * __restrict__ cscale, const double * __restrict__ evecs,
               const double * __restrict__ left, double * __restrict__ sumtable, unsigned int sites,
               unsigned int R, unsigned int K, unsigned int Kp, int per_rate)
{
  const unsigned int n = blockIdx.x * PLG_GEN_THREADS + threadIdx.x;
  if (n >= sites) return;
  const size_t span = (size_t)R * Kp;
  unsigned int mn = 0xffffffffu;
  if (per_rate)
    for (unsigned int k = 0; k < R; ++k)
    {
      const unsigned int s = (pscale ? pscale[(size_t)n * R + k] : 0u) + (cscale ? cscale[(size_t)n * R + k] : 0u);
      mn = s < mn ? s : mn;
    }
  for (unsigned int k = 0; k < R; ++k)
  {
    double f = 1.0;
    if (per_rate)
    {
      unsigned int d = (pscale ? pscale[(size_t)n * R + k] : 0u) + (cscale ? cscale[(size_t)n * R + k] : 0u) - mn;
      if (d > PLL_SCALE_RATE_MAXDIFF) d = PLL_SCALE_RATE_MAXDIFF;
      for (unsigned int q = 0; q < d; ++q) f = __dmul_rn(f, PLG_SCALE_THRESHOLD);
    }
    const double * p = clvp + (size_t)n * span + (size_t)k * Kp;
    for (unsigned int j = 0; j < Kp; ++j)
    {
      double s = 0.0;
      if (j < K)
      {
        double lt, rt = 0.0;
        const double * v = evecs + (size_t)k * K * Kp + (size_t)j * Kp;
        if (tip)
        {
          lt = left[((size_t)tip[n] * R + k) * Kp + j];
          for (unsigned int i = 0; i < K; ++i) rt = __dadd_rn(rt, __dmul_rn(v[i], p[i]));
        }
        else
        {
          const double * w = left + (size_t)k * K * Kp + (size_t)j * Kp;
          const double * c = clvc + (size_t)n * span + (size_t)k * Kp;
          lt = 0.0;
          for (unsigned int i = 0; i < K; ++i)
          {
            lt = __dadd_rn(lt, __dmul_rn(w[i], p[i]));
            rt = __dadd_rn(rt, __dmul_rn(v[i], c[i]));
          }
        }
        s = __dmul_rn(lt, rt);
        if (f != 1.0) s = __dmul_rn(s, f);
      }
      sumtable[(size_t)n * span + (size_t)k * Kp + j] = s;
    }
  }
}

int plg_gen_sumtable(plg_context * ctx, const double * clvp, const double * clvc, const unsigned char * tip,
                     const unsigned int * pscale, const unsigned int * cscale, const double * dev_evecs,
                     const double * dev_left, double * sumtable)
{
  const unsigned int nblocks = (ctx->d.sites + PLG_GEN_THREADS - 1) / PLG_GEN_THREADS;
  k_gen_sumtable<<<nblocks, PLG_GEN_THREADS, 0, ctx->stream>>>(clvp, clvc, tip, pscale, cscale, dev_evecs, dev_left,
                                                              sumtable, ctx->d.sites, ctx->d.rate_cats, ctx->d.states,
                                                              ctx->d.states_padded, ctx->rate_scalers ? 1 : 0);
  PLG_LAUNCH_CHECK(ctx);
  return PLG_OK;
}

/* ------------------------------------------------------------------------------------ */
/* derivatives (reference src/core_derivatives.c:449-500, :660-690) */
__global__ void __launch_bounds__(PLG_GEN_THREADS)
k_gen_derivatives(const double * __restrict__ sumtable, const double * __restrict__ diagp /* [R][K][4] */,
                  const double * __restrict__ rate_weights, const double * __restrict__ prop_invar,
                  const double * __restrict__ freqs, const unsigned int * __restrict__ weights,
                  const int * __restrict__ invariant, unsigned int sites, unsigned int R, unsigned int K,
                  unsigned int Kp, double * partials, unsigned int * counter, const PlgSink sink)
{
  const unsigned int n = blockIdx.x * PLG_GEN_THREADS + threadIdx.x;
  double df = 0.0, ddf = 0.0;
  if (n < sites)
  {
    double l0 = 0.0, l1 = 0.0, l2 = 0.0;
    for (unsigned int k = 0; k < R; ++k)
    {
      const double * s = sumtable + ((size_t)n * R + k) * Kp;
      double c0 = 0.0, c1 = 0.0, c2 = 0.0;
      for (unsigned int j = 0; j < K; ++j)
      {
        const double * d = diagp + ((size_t)k * K + j) * 4;
        c0 = __dadd_rn(c0, __dmul_rn(s[j], d[0]));
        c1 = __dadd_rn(c1, __dmul_rn(s[j], d[1]));
        c2 = __dadd_rn(c2, __dmul_rn(s[j], d[2]));
      }
      const double pinv = prop_invar[k];
      if (pinv > 0.0)
      {
        const int inv = invariant ? invariant[n] : -1;
        const double inv_lk = (inv == -1) ? 0.0 : __dmul_rn(freqs[(size_t)k * Kp + inv], pinv);
        const double q = __dsub_rn(1.0, pinv);
        c0 = __dadd_rn(__dmul_rn(c0, q), inv_lk);
        c1 = __dmul_rn(c1, q);
        c2 = __dmul_rn(c2, q);
      }
      l0 = __dadd_rn(l0, __dmul_rn(c0, rate_weights[k]));
      l1 = __dadd_rn(l1, __dmul_rn(c1, rate_weights[k]));
      l2 = __dadd_rn(l2, __dmul_rn(c2, rate_weights[k]));
    }
    const double d1 = -__ddiv_rn(l1, l0);
    const double d2 = __dsub_rn(__dmul_rn(d1, d1), __ddiv_rn(l2, l0));
    df = __dmul_rn((double)weights[n], d1);
    ddf = __dmul_rn((double)weights[n], d2);
  }
  gen_finish<2>(df, ddf, partials, counter, sink);
}

int plg_gen_derivatives(plg_context * ctx, const double * sumtable, const double * diagptable,
                        const double * rate_weights, const double * prop_invar, const double * freqs,
                        double * d_f, double * dd_f)
{
  const unsigned int R = ctx->d.rate_cats, K = ctx->d.states, Kp = ctx->d.states_padded;
  const unsigned int nblocks = ctx->active_sites ? (ctx->active_sites + PLG_GEN_THREADS - 1) / PLG_GEN_THREADS : 1u;
  int rc = plg_ensure_partials(ctx, 2 * (size_t)nblocks);
  if (rc) return rc;
  bool any_pinv = false;
  for (unsigned int i = 0; i < R; ++i) any_pinv |= prop_invar[i] > 0;
  if (any_pinv && !ctx->has_invariant)
  {
    plg_set_error("derivatives with prop_invar > 0 need the invariant-site index");
    return PLG_E_INVALID;
  }
  if (plg_stage_reserve(ctx, ((size_t)R * K * 4 + 2 * R + (size_t)R * Kp) * 8 + 4 * 256)) return PLG_E_CUDA;
  const double * d_diag = (const double *)plg_stage(ctx, diagptable, (size_t)R * K * 4 * sizeof(double));
  const double * d_rw = (const double *)plg_stage(ctx, rate_weights, R * sizeof(double));
  const double * d_pinv = (const double *)plg_stage(ctx, prop_invar, R * sizeof(double));
  const double * d_freqs = (const double *)plg_stage(ctx, freqs, (size_t)R * Kp * sizeof(double));
  if (!d_diag || !d_rw || !d_pinv || !d_freqs) return PLG_E_CUDA;
  k_gen_derivatives<<<nblocks, PLG_GEN_THREADS, 0, ctx->stream>>>(
      sumtable, d_diag, d_rw, d_pinv, d_freqs, ctx->weights, ctx->has_invariant ? ctx->invariant : NULL,
      ctx->active_sites, R, K, Kp, ctx->partials, ctx->counter, plg_make_sink(ctx));
  PLG_LAUNCH_CHECK(ctx);
  ctx->stats.d2h_bytes += 2 * sizeof(double);
  return plg_finish_result(ctx, d_f, dd_f);
}
