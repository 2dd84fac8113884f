/*
 * plg_partials.cu - conditional-likelihood-vector (CLV) updates: the >99 % hot loop.
 *
 * Replaces, for a whole pll_operation_t list at once:
 *   pll_update_partials' per-node loop          reference src/partials.c:177-213
 *   pll_core_create_lookup (+_4x4_avx,_20x20)   reference src/core_partials_avx.c:146-364
 *   pll_core_update_partial_tt                   reference src/core_partials_avx.c:531-618
 *   pll_core_update_partial_ti                   reference src/core_partials_avx.c:899-1340
 *   pll_core_update_partial_ii                   reference src/core_partials_avx.c:366-529,
 *                                                          src/core_partials_avx2.c:568-803
 *   fill_parent_scaler + threshold rescale       reference src/core_partials_avx.c:24-46,490-527
 *
 * Work mapping.  One thread owns one (site, rate) element of the parent CLV: for DNA that is
 * one 256-bit load per child and one 256-bit store, so a warp streams 1 KB contiguous per
 * array per instruction.  blockIdx.y selects the operation inside a batch of mutually
 * independent operations (one dependency level), so a level of the tree is one launch.  The
 * per-site "all entries below 2^-256" decision is a warp ballot over the rate lanes of a
 * site; the lane of rate 0 does the integer scaler arithmetic.
 *
 * Arithmetic follows the order of the reference's AVX2-flag path exactly:
 * DNA uses unfused multiply/add with the (a0+a1)+(a2+a3) horizontal sum; 20-state
 * inner-inner uses 4 FMA lane accumulators per row, 20-state tip-inner the unfused form.
 * The file is compiled with -fmad=false so nothing is contracted behind our back.
 */
#include <algorithm>
#include <array>
#include <cmath>

#include <unordered_map>

#include "plg_internal.cuh"
#include "plg_async.cuh"
#include "plg_dmma.cuh"

/* ------------------------------------------------------------------------------------ */
/* descriptors                                                                           */
/* ------------------------------------------------------------------------------------ */
/* ------------------------------------------------------------------------------------ */
/* tip lookup tables                                                                     */
/* ------------------------------------------------------------------------------------ */
/* DNA: table[code][rate][i] = sum over the states m in `code` of P_rate[i][m], summed as
 * (a0+a1)+(a2+a3) with absent states contributing +0.0 - the masked-load form of reference
 * src/core_partials_avx.c:944-984 (tip-inner) and :280-353 (tip-tip, per side). */
__global__ void k_tip_tables_dna(const TableJob * __restrict__ jobs, unsigned int rate_cats)
{
  const TableJob job = jobs[blockIdx.x];
  const unsigned int entries = 16u * rate_cats * 4u;
  for (unsigned int t = threadIdx.x; t < entries; t += blockDim.x)
  {
    const unsigned int i = t & 3u;
    const unsigned int k = (t >> 2) % rate_cats;
    const unsigned int code = t / (rate_cats * 4u);
    const double * row = job.pmat + (size_t)k * 16 + i * 4;
    const double a0 = (code & 1u) ? row[0] : 0.0;
    const double a1 = (code & 2u) ? row[1] : 0.0;
    const double a2 = (code & 4u) ? row[2] : 0.0;
    const double a3 = (code & 8u) ? row[3] : 0.0;
    job.out[t] = hsum4(a0, a1, a2, a3);
  }
}

/* 20 states: table[code][rate][i] = sequential sum, in increasing state order, of
 * P_rate[i][m] over the states m present in tipmap[code] (reference
 * src/core_partials_avx.c:1140-1177 and :177-220). */
__global__ void k_tip_tables_aa(const TableJob * __restrict__ jobs, unsigned int rate_cats,
                                unsigned int maxstates, const TipmapArg tm)
{
  const TableJob job = jobs[blockIdx.x];
  const unsigned int entries = maxstates * rate_cats * 20u;
  for (unsigned int t = threadIdx.x; t < entries; t += blockDim.x)
  {
    const unsigned int i = t % 20u;
    const unsigned int k = (t / 20u) % rate_cats;
    const unsigned int code = t / (rate_cats * 20u);
    const unsigned int state = tm.map[code];
    const double * row = job.pmat + (size_t)k * 400 + i * 20;
    double s = 0.0;
    for (unsigned int m = 0; m < 20u; ++m)
      if ((state >> m) & 1u) s = __dadd_rn(s, row[m]);
    job.out[t] = s;
  }
}

/* ------------------------------------------------------------------------------------ */
/* scaling helpers                                                                       */
/* ------------------------------------------------------------------------------------ */
/* scale_mode: 0 = parent has no scaler (no scaling at all), 1 = per-site, 2 = per-rate
 * (reference src/core_partials_avx.c:397-410). */

/* Returns true if this element must be multiplied by 2^256, and performs the scaler
 * bookkeeping.  `below` = all states of this (site, rate) element are < 2^-256.
 * Must be called by all 32 lanes of the warp. */
template <int R>
__device__ __forceinline__ bool scale_decision(bool valid, bool below, int scale_mode,
                                               unsigned int e, const DevOp & op)
{
  const unsigned int lane = threadIdx.x & 31u;
  if (scale_mode == 1)
  {
    const unsigned int b = __ballot_sync(0xffffffffu, valid && below);
    const unsigned int full = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
    const unsigned int grp = (b >> (lane & ~(unsigned int)(R - 1))) & full;
    const bool scale = (grp == full);
    if (valid && (lane & (R - 1)) == 0)
    {
      const unsigned int n = e / R;
      unsigned int s = scale ? 1u : 0u;
      if (op.lscale) s += op.lscale[n];
      if (op.rscale) s += op.rscale[n];
      op.pscale[n] = s;
    }
    return scale;
  }
  if (scale_mode == 2)
  {
    if (valid)
    {
      unsigned int s = below ? 1u : 0u;
      if (op.lscale) s += op.lscale[e];
      if (op.rscale) s += op.rscale[e];
      op.pscale[e] = s;
    }
    return below;
  }
  return false;
}

__device__ __forceinline__ bool all_below(const d4 & p)
{
  return (p.x < PLG_SCALE_THRESHOLD) & (p.y < PLG_SCALE_THRESHOLD) &
         (p.z < PLG_SCALE_THRESHOLD) & (p.w < PLG_SCALE_THRESHOLD);
}

__device__ __forceinline__ d4 scale_up(const d4 & p)
{
  d4 r;
  r.x = __dmul_rn(p.x, PLG_SCALE_FACTOR);
  r.y = __dmul_rn(p.y, PLG_SCALE_FACTOR);
  r.z = __dmul_rn(p.z, PLG_SCALE_FACTOR);
  r.w = __dmul_rn(p.w, PLG_SCALE_FACTOR);
  return r;
}

/* ------------------------------------------------------------------------------------ */
/* DNA kernels (4 states)                                                                */
/* ------------------------------------------------------------------------------------ */
#define PLG_DNA_THREADS 256

/* tip-tip: parent[n][k][i] = tableL[l[n]][k][i] * tableR[r[n]][k][i]; never scales and
 * zeroes the parent scaler (reference src/core_partials_avx.c:581-618, :262-364: the
 * reference materialises the 16x16 product table, the products are the same numbers). */
template <int R, int ITEMS>
__global__ void __launch_bounds__(PLG_DNA_THREADS)
k_partial_tt_dna(const DevOp * __restrict__ ops, unsigned int nelem, int scale_mode)
{
  /* one thread per HALF element (two states, 16 bytes): a warp's store instruction then
   * covers 512 contiguous bytes in full 32-byte sectors */
  const DevOp op = ops[blockIdx.y];
  __shared__ double2 tabl[16 * R * 2];
  __shared__ double2 tabr[16 * R * 2];
  for (unsigned int t = threadIdx.x; t < 16 * R * 2; t += PLG_DNA_THREADS)
  {
    tabl[t] = *reinterpret_cast<const double2 *>(op.lmat + (size_t)t * 2);
    tabr[t] = *reinterpret_cast<const double2 *>(op.rmat + (size_t)t * 2);
  }
  __syncthreads();

  const unsigned int nhalf = nelem * 2u;
  const unsigned int base = blockIdx.x * (PLG_DNA_THREADS * ITEMS) + threadIdx.x;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
  {
    const unsigned int hidx = base + j * PLG_DNA_THREADS; /* half-element index */
    if (hidx < nhalf)
    {
      const unsigned int e = hidx >> 1;
      const unsigned int kh = hidx & (2 * R - 1); /* (rate, half) within the site */
      const unsigned int n = e / R;
      const unsigned int lc = __ldg(op.ltip + n);
      const unsigned int rc = __ldg(op.rtip + n);
      const double2 a = tabl[lc * 2 * R + kh];
      const double2 b = tabr[rc * 2 * R + kh];
      double2 p;
      p.x = __dmul_rn(a.x, b.x);
      p.y = __dmul_rn(a.y, b.y);
      asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(op.parent + (size_t)hidx * 2),
                   "d"(p.x), "d"(p.y)
                   : "memory");
      if (scale_mode == 1)
      {
        if (kh == 0) op.pscale[n] = 0u;
      }
      else if (scale_mode == 2)
      {
        if ((hidx & 1u) == 0) op.pscale[e] = 0u;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* DNA streaming kernels: persistent CTAs + TMA bulk-copy pipeline                       */
/* ------------------------------------------------------------------------------------ */
/*
 * The tip-inner and inner-inner updates are pure streaming (0.6 flop/B): what limits them is
 * the number of bytes in flight per SM, not arithmetic.  These kernels therefore decouple the
 * memory pipeline from the register file:
 *   - the grid is persistent (a few CTAs per SM); the (operation, tile) space of a whole
 *     dependency level is flattened and cut into one contiguous chunk per CTA;
 *   - one elected thread feeds a STAGES-deep ring of shared-memory tiles with 1-D TMA
 *     bulk copies (cp.async.bulk, completion on "full" mbarriers); consumers release a stage
 *     through an "empty" mbarrier as soon as they have pulled their 32-byte element into
 *     registers, so copies for the next tiles are always outstanding;
 *   - every thread owns one (site, rate) element per tile, with its rate's P-matrix rows in
 *     registers for the whole chunk (reloaded only when the chunk crosses into the next
 *     operation);
 *   - results leave with 256-bit streaming stores straight from registers (a shared-memory
 *     ring + TMA bulk stores was measured 8 % slower: it costs a CTA barrier per tile).
 * Arithmetic, voting and scaler bookkeeping are exactly those of the simple kernels above.
 */
#define PLG_STREAM_THREADS 256
#define PLG_II_TILE 256     /* elements per tile: 8 KB per CLV array per stage */
#ifndef PLG_TI_TILE
#define PLG_TI_TILE 256
#endif
#ifndef PLG_II_STAGES
#define PLG_II_STAGES 4 /* x 16 KB */
#endif
#ifndef PLG_TI_STAGES
#define PLG_TI_STAGES 4     /* x 16 KB */
#endif
#ifndef PLG_II_MINB
#define PLG_II_MINB 2 /* resident CTAs per SM */
#endif
#ifndef PLG_TI_MINB
#define PLG_TI_MINB 3
#endif

template <int R, int KIND, int STAGES>
struct StreamSmem
{
  static constexpr int TILE_E = (KIND == PLG_KIND_II) ? PLG_II_TILE : PLG_TI_TILE;
  d4 in_r[STAGES][TILE_E];
  d4 in_l[KIND == PLG_KIND_II ? STAGES : 1][KIND == PLG_KIND_II ? TILE_E : 1];
  double2 tab[KIND == PLG_KIND_TI ? 16 * R * 2 : 1];
  unsigned char tips[KIND == PLG_KIND_TI ? STAGES : 1][TILE_E];
  unsigned int sc_l[KIND == PLG_KIND_II ? STAGES : 1][TILE_E]; /* child scalers of the tile */
  unsigned int sc_r[STAGES][TILE_E];
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
};

/* scaler bookkeeping for the half-element mapping: the 2R lanes of a site vote */
template <int R>
__device__ __forceinline__ bool scale_decision_half(bool valid, bool below, int scale_mode,
                                                    unsigned int hidx, unsigned int child_sum,
                                                    const DevOp & op)
{
  const unsigned int lane = threadIdx.x & 31u;
  if (scale_mode == 1)
  {
    constexpr unsigned int W = 2 * R; /* lanes per site */
    const unsigned int b = __ballot_sync(0xffffffffu, valid && below);
    const unsigned int full = (W >= 32) ? 0xffffffffu : ((1u << W) - 1u);
    const unsigned int grp = (b >> (lane & ~(W - 1))) & full;
    const bool scale = (grp == full);
    if (valid && (lane & (W - 1)) == 0)
    {
      op.pscale[hidx / W] = child_sum + (scale ? 1u : 0u);
    }
    return scale;
  }
  if (scale_mode == 2)
  {
    /* per-rate: the two halves of an element are adjacent lanes */
    const unsigned int b = __ballot_sync(0xffffffffu, valid && below);
    const bool both = ((b >> (lane & ~1u)) & 3u) == 3u;
    if (valid && (lane & 1u) == 0) op.pscale[hidx >> 1] = child_sum + (both ? 1u : 0u);
    return both;
  }
  return false;
}

__device__ __forceinline__ void st_stream2(double * p, double x, double y)
{
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}

/*
 * One thread per HALF element: (site, rate, state pair).  Each thread keeps the two rows of
 * its rate's P-matrices that it needs in registers, reads the full 4-state child vectors
 * from the shared-memory stage and emits one 128-bit store, so that every store instruction
 * of a warp covers 512 contiguous bytes in whole 32-byte sectors (measured +7 % write
 * bandwidth over 256-bit per-thread stores, which the LSU splits into half-sector passes).
 */
template <int R, int KIND, int STAGES, int MINB>
__global__ void __launch_bounds__(PLG_STREAM_THREADS, MINB)
k_partial_stream_dna(const DevOp * __restrict__ ops, unsigned int nelem, unsigned int ntiles,
                     unsigned int total_tiles, int scale_mode)
{
  using namespace plg_async;
  constexpr unsigned int TILE = StreamSmem<R, KIND, STAGES>::TILE_E;
  constexpr int EPT = 2 * TILE / PLG_STREAM_THREADS; /* half elements per thread per tile */
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StreamSmem<R, KIND, STAGES> & sm = *reinterpret_cast<StreamSmem<R, KIND, STAGES> *>(smem_raw);

  const unsigned int tid = threadIdx.x;
  const unsigned int h = tid & 1u;              /* which state pair */
  const unsigned int k = (tid >> 1) & (R - 1);  /* which rate       */

  /* this CTA's contiguous chunk of the flattened (operation, tile) space */
  const unsigned int per = (total_tiles + gridDim.x - 1) / gridDim.x;
  const unsigned int q0 = blockIdx.x * per;
  if (q0 >= total_tiles) return;
  const unsigned int q1 = (q0 + per < total_tiles) ? q0 + per : total_tiles;
  const unsigned int n = q1 - q0;

  if (tid == 0)
  {
#pragma unroll
    for (int s = 0; s < STAGES; ++s)
    {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], PLG_STREAM_THREADS / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  /* producer: TMA copies for the j-th tile of the chunk into stage j % STAGES */
  auto issue = [&](unsigned int j) {
    const unsigned int q = q0 + j;
    const unsigned int o = q / ntiles;
    const unsigned int e0 = (q - o * ntiles) * TILE;
    const unsigned int cnt = (nelem - e0 < TILE) ? nelem - e0 : TILE;
    const unsigned int s = j % STAGES;
    const DevOp * op = ops + o;
    const unsigned int bytes = cnt * 32u;
    /* child scalers of the tile ride along (no global loads in the consumer loop) */
    const unsigned int sc_off = (scale_mode == 2) ? e0 : e0 / R;
    const unsigned int sc_bytes = (((scale_mode == 2) ? cnt : cnt / R) * 4u + 15u) & ~15u;
    const unsigned int * lsc = (scale_mode != 0 && KIND == PLG_KIND_II) ? op->lscale : nullptr;
    const unsigned int * rsc = (scale_mode != 0) ? op->rscale : nullptr;
    const unsigned int sc_total = (lsc ? sc_bytes : 0u) + (rsc ? sc_bytes : 0u);
    if (KIND == PLG_KIND_II)
    {
      mbar_arrive_expect_tx(&sm.full[s], 2 * bytes + sc_total);
      bulk_g2s(sm.in_l[s], op->left + (size_t)e0 * 4, bytes, &sm.full[s]);
      bulk_g2s(sm.in_r[s], op->right + (size_t)e0 * 4, bytes, &sm.full[s]);
      if (lsc) bulk_g2s(sm.sc_l[s], lsc + sc_off, sc_bytes, &sm.full[s]);
    }
    else
    {
      const unsigned int tip_bytes = ((cnt / R) + 15u) & ~15u;
      mbar_arrive_expect_tx(&sm.full[s], bytes + tip_bytes + sc_total);
      bulk_g2s(sm.in_r[s], op->right + (size_t)e0 * 4, bytes, &sm.full[s]);
      bulk_g2s(sm.tips[s], op->ltip + e0 / R, tip_bytes, &sm.full[s]);
    }
    if (rsc) bulk_g2s(sm.sc_r[s], rsc + sc_off, sc_bytes, &sm.full[s]);
  };

  if (tid == 0)
  {
    const unsigned int pre = n < STAGES ? n : STAGES;
    for (unsigned int j = 0; j < pre; ++j) issue(j);
  }

  double L[8], Rm[8]; /* rows 2h and 2h+1 of this thread's rate */
  DevOp op;
  unsigned int cur_op = 0xffffffffu;

  for (unsigned int j = 0; j < n; ++j)
  {
    const unsigned int q = q0 + j;
    const unsigned int o = q / ntiles;
    const unsigned int tile = q - o * ntiles;
    if (o != cur_op)
    {
      /* chunk crossed into the next operation (block-uniform): new pointers and matrices */
      cur_op = o;
      op = ops[o];
#pragma unroll
      for (int i = 0; i < 2; ++i)
      {
        const d4 rr = *reinterpret_cast<const d4 *>(op.rmat + k * 16 + (2 * h + i) * 4);
        Rm[4 * i + 0] = rr.x; Rm[4 * i + 1] = rr.y; Rm[4 * i + 2] = rr.z; Rm[4 * i + 3] = rr.w;
        if (KIND == PLG_KIND_II)
        {
          const d4 ll = *reinterpret_cast<const d4 *>(op.lmat + k * 16 + (2 * h + i) * 4);
          L[4 * i + 0] = ll.x; L[4 * i + 1] = ll.y; L[4 * i + 2] = ll.z; L[4 * i + 3] = ll.w;
        }
      }
      if (KIND == PLG_KIND_TI)
      {
        __syncthreads(); /* nobody still reads the previous operation's table */
        for (unsigned int t = tid; t < 16 * R * 2; t += PLG_STREAM_THREADS)
          sm.tab[t] = *reinterpret_cast<const double2 *>(op.lmat + (size_t)t * 2);
        __syncthreads();
      }
    }

    const unsigned int s = j % STAGES;
    mbar_wait(&sm.full[s], (j / STAGES) & 1u);
    d4 r[EPT], l[EPT];
    unsigned int code[EPT], csum[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i)
    {
      const unsigned int el = (tid + i * PLG_STREAM_THREADS) >> 1; /* element within the tile */
      r[i] = sm.in_r[s][el];
      if (KIND == PLG_KIND_II) l[i] = sm.in_l[s][el];
      /* lanes past the end of a short tile see stale bytes: keep the index in range */
      else code[i] = sm.tips[s][el / R] & 15u;
      csum[i] = 0u;
      if (scale_mode != 0)
      {
        const unsigned int si = (scale_mode == 2) ? el : el / R;
        if (KIND == PLG_KIND_II && op.lscale) csum[i] += sm.sc_l[s][si];
        if (op.rscale) csum[i] += sm.sc_r[s][si];
      }
    }
    __syncwarp();
    if ((tid & 31u) == 0) mbar_arrive(&sm.empty[s]);

    if (tid == 0 && j >= 1 && j - 1 + STAGES < n)
    {
      /* stage of the previous tile: every warp released it an iteration ago */
      mbar_wait(&sm.empty[(j - 1) % STAGES], ((j - 1) / STAGES) & 1u);
      issue(j - 1 + STAGES);
    }

#pragma unroll
    for (int i = 0; i < EPT; ++i)
    {
      const unsigned int hidx = tile * (2 * TILE) + tid + i * PLG_STREAM_THREADS;
      const bool valid = hidx < 2 * nelem;
      const double y0 = dot4_unfused(Rm[0], Rm[1], Rm[2], Rm[3], r[i]);
      const double y1 = dot4_unfused(Rm[4], Rm[5], Rm[6], Rm[7], r[i]);
      double x0, x1;
      if (KIND == PLG_KIND_II)
      {
        x0 = dot4_unfused(L[0], L[1], L[2], L[3], l[i]);
        x1 = dot4_unfused(L[4], L[5], L[6], L[7], l[i]);
      }
      else
      {
        const double2 t = sm.tab[(code[i] * R + k) * 2 + h];
        x0 = t.x;
        x1 = t.y;
      }
      double p0 = __dmul_rn(x0, y0), p1 = __dmul_rn(x1, y1);
      const bool below = valid && (p0 < PLG_SCALE_THRESHOLD) & (p1 < PLG_SCALE_THRESHOLD);
      if (scale_decision_half<R>(valid, below, scale_mode, hidx, csum[i], op))
      {
        p0 = __dmul_rn(p0, PLG_SCALE_FACTOR);
        p1 = __dmul_rn(p1, PLG_SCALE_FACTOR);
      }
      if (valid) st_stream2(op.parent + (size_t)hidx * 2, p0, p1);
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* 20-state kernels                                                                      */
/* ------------------------------------------------------------------------------------ */
#define PLG_AA_THREADS 128
/* shared-memory stride of one rate's 20x20 matrix, in doubles: 404 (not 400) so that the R
 * matrices the lanes of a warp read from start 8 banks apart instead of all in the same bank */
#define PLG_AA_MSTRIDE 404

/* rows [4*ib, 4*ib+4) of M (20x20, row-major in shared memory) times c, with the AVX2
 * accumulation order: per row four lane accumulators over the five column blocks, FMA
 * (reference src/core_partials_avx2.c:671-731) or mul+add (reference
 * src/core_partials_avx.c:1236-1286), then the (a0+a1)+(a2+a3) sum. */
template <bool FUSED>
__device__ __forceinline__ double row20(const double * __restrict__ Mrow, const double (&c)[20])
{
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int b = 0; b < 5; ++b)
  {
    const double2 m01 = *reinterpret_cast<const double2 *>(Mrow + 4 * b);
    const double2 m23 = *reinterpret_cast<const double2 *>(Mrow + 4 * b + 2);
    if (FUSED)
    {
      a0 = __fma_rn(m01.x, c[4 * b + 0], a0);
      a1 = __fma_rn(m01.y, c[4 * b + 1], a1);
      a2 = __fma_rn(m23.x, c[4 * b + 2], a2);
      a3 = __fma_rn(m23.y, c[4 * b + 3], a3);
    }
    else
    {
      a0 = __dadd_rn(a0, __dmul_rn(m01.x, c[4 * b + 0]));
      a1 = __dadd_rn(a1, __dmul_rn(m01.y, c[4 * b + 1]));
      a2 = __dadd_rn(a2, __dmul_rn(m23.x, c[4 * b + 2]));
      a3 = __dadd_rn(a3, __dmul_rn(m23.y, c[4 * b + 3]));
    }
  }
  return hsum4(a0, a1, a2, a3);
}

__device__ __forceinline__ void load20(const double * p, double (&c)[20])
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

__device__ __forceinline__ void rescale20(double * p)
{
#pragma unroll
  for (int b = 0; b < 5; ++b)
  {
    d4 v = *reinterpret_cast<d4 *>(p + 4 * b);
    *reinterpret_cast<d4 *>(p + 4 * b) = scale_up(v);
  }
}

/* inner-inner, 20 states (reference src/core_partials_avx2.c:632-802) */
template <int R>
__global__ void __launch_bounds__(PLG_AA_THREADS)
k_partial_ii_aa(const DevOp * __restrict__ ops, unsigned int nelem, int scale_mode)
{
  extern __shared__ __align__(16) double smem[];
  double * Ls = smem;            /* [R][20][20] */
  double * Rs = smem + R * PLG_AA_MSTRIDE; /* [R][20][20] */
  const DevOp op = ops[blockIdx.y];
  for (unsigned int t = threadIdx.x; t < R * 400; t += PLG_AA_THREADS)
  {
    Ls[(t / 400) * PLG_AA_MSTRIDE + t % 400] = __ldg(op.lmat + t);
    Rs[(t / 400) * PLG_AA_MSTRIDE + t % 400] = __ldg(op.rmat + t);
  }
  __syncthreads();

  const unsigned int k = threadIdx.x & (R - 1);
  const unsigned int e = blockIdx.x * PLG_AA_THREADS + threadIdx.x;
  const bool valid = e < nelem;
  bool below = true;
  double * out = op.parent + (size_t)e * 20;
  if (valid)
  {
    double l[20], r[20];
    load20(op.left + (size_t)e * 20, l);
    load20(op.right + (size_t)e * 20, r);
    const double * Lk = Ls + k * PLG_AA_MSTRIDE;
    const double * Rk = Rs + k * PLG_AA_MSTRIDE;
#pragma unroll
    for (int ib = 0; ib < 5; ++ib)
    {
      d4 p;
      p.x = __dmul_rn(row20<true>(Lk + (4 * ib + 0) * 20, l), row20<true>(Rk + (4 * ib + 0) * 20, r));
      p.y = __dmul_rn(row20<true>(Lk + (4 * ib + 1) * 20, l), row20<true>(Rk + (4 * ib + 1) * 20, r));
      p.z = __dmul_rn(row20<true>(Lk + (4 * ib + 2) * 20, l), row20<true>(Rk + (4 * ib + 2) * 20, r));
      p.w = __dmul_rn(row20<true>(Lk + (4 * ib + 3) * 20, l), row20<true>(Rk + (4 * ib + 3) * 20, r));
      below = below && all_below(p);
      *reinterpret_cast<d4 *>(out + 4 * ib) = p;
    }
  }
  const bool scale = scale_decision<R>(valid, valid && below, scale_mode, e, op);
  if (valid && scale) rescale20(out);
}

/* tip-inner, 20 states (reference src/core_partials_avx.c:1204-1338: the AVX kernel is what
 * the AVX2 flag dispatches to, reference src/core_partials.c:427-444) */
template <int R>
__global__ void __launch_bounds__(PLG_AA_THREADS)
k_partial_ti_aa(const DevOp * __restrict__ ops, unsigned int nelem, int scale_mode)
{
  extern __shared__ __align__(16) double smem[];
  double * Rs = smem; /* [R][20][20] */
  const DevOp op = ops[blockIdx.y];
  for (unsigned int t = threadIdx.x; t < R * 400; t += PLG_AA_THREADS)
    Rs[(t / 400) * PLG_AA_MSTRIDE + t % 400] = __ldg(op.rmat + t);
  __syncthreads();

  const unsigned int k = threadIdx.x & (R - 1);
  const unsigned int e = blockIdx.x * PLG_AA_THREADS + threadIdx.x;
  const bool valid = e < nelem;
  bool below = true;
  double * out = op.parent + (size_t)e * 20;
  if (valid)
  {
    double r[20];
    load20(op.right + (size_t)e * 20, r);
    const unsigned int code = __ldg(op.ltip + e / R);
    const double * tab = op.lmat + ((size_t)code * R + k) * 20;
    const double * Rk = Rs + k * PLG_AA_MSTRIDE;
#pragma unroll
    for (int ib = 0; ib < 5; ++ib)
    {
      const d4 a = *reinterpret_cast<const d4 *>(tab + 4 * ib);
      d4 p;
      p.x = __dmul_rn(a.x, row20<false>(Rk + (4 * ib + 0) * 20, r));
      p.y = __dmul_rn(a.y, row20<false>(Rk + (4 * ib + 1) * 20, r));
      p.z = __dmul_rn(a.z, row20<false>(Rk + (4 * ib + 2) * 20, r));
      p.w = __dmul_rn(a.w, row20<false>(Rk + (4 * ib + 3) * 20, r));
      below = below && all_below(p);
      *reinterpret_cast<d4 *>(out + 4 * ib) = p;
    }
  }
  const bool scale = scale_decision<R>(valid, valid && below, scale_mode, e, op);
  if (valid && scale) rescale20(out);
}

/* tip-tip, 20 states (reference src/core_partials_avx.c:531-579, :225-256): one thread per
 * 16-byte chunk (two states) of the parent CLV, so every store instruction of a warp covers 512
 * contiguous bytes; the two per-side tables stay hot in L1 */
#define PLG_AA_TT_ITEMS 4
template <int R>
__global__ void __launch_bounds__(256)
k_partial_tt_aa(const DevOp * __restrict__ ops, unsigned int nelem, int scale_mode)
{
  const DevOp op = ops[blockIdx.y];
  const unsigned long long nchunks = (unsigned long long)nelem * 10ull; /* 16-byte chunks */
  const unsigned long long base = (unsigned long long)blockIdx.x * (256 * PLG_AA_TT_ITEMS) + threadIdx.x;
#pragma unroll
  for (int j = 0; j < PLG_AA_TT_ITEMS; ++j)
  {
    const unsigned long long c = base + (unsigned long long)j * 256;
    if (c >= nchunks) break;
    const unsigned int n = (unsigned int)(c / (R * 10));  /* site                          */
    const unsigned int w = (unsigned int)(c % (R * 10));  /* chunk within the site: k*10+p */
    const unsigned int lc = __ldg(op.ltip + n);
    const unsigned int rc = __ldg(op.rtip + n);
    const double2 a = __ldg(reinterpret_cast<const double2 *>(op.lmat + (size_t)lc * R * 20) + w);
    const double2 b = __ldg(reinterpret_cast<const double2 *>(op.rmat + (size_t)rc * R * 20) + w);
    st_stream2(op.parent + c * 2, __dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y));
    if (scale_mode == 1)
    {
      if (w == 0) op.pscale[n] = 0u;
    }
    else if (scale_mode == 2)
    {
      if (w % 10 == 0) op.pscale[(size_t)n * R + w / 10] = 0u;
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* 20-state kernels on the FP64 tensor cores (DMMA m8n8k4)                               */
/* ------------------------------------------------------------------------------------ */
/*
 * The (sites x rates) x 20 x 20 products are the one GEMM-shaped piece of the path.  On the
 * vector pipe every FMA needs a P-matrix operand replicated to 32 lanes through shared memory
 * (LDS-bound: measured 1.8 TB/s, 17 % of the FP64 peak); DMMA keeps the matrix distributed
 * across the lanes of a warp as A fragments and reads each CLV entry exactly once as a B
 * fragment straight from HBM (8 full 32-byte sectors per request), so the loop is
 * tensor-/HBM-bound instead.  Measured peaks on this B200: DFMA 33.8 TFLOP/s, DMMA 37.1.
 *
 *   Y[s][i] = sum_j c[s][k][j] * P_k[i][j]      M = 8 sites per warp-tile,
 *                                               N = parent states (20, padded to 3 x 8), K = 20
 *   The 20 child states are fed in 5 k-steps in the order of plg_dmma.cuh (shared with the
 *   single-kernel traversal, so both produce the same bits):
 *   A fragment (ks)    : lane (g, q) holds c[site g][j(ks, q)] - two 16-byte pieces (states
 *                        2q, 2q+1 and 8+2q, 9+2q) and one 8-byte piece of the CLV row
 *   B fragment (nt, ks): lane holds P[8nt + g][j(ks, q)]             (shared memory, per op)
 *   D fragment (nt)    : lane holds Y[site g][8nt + 2q + {0,1}]      -> 128-bit stores
 *
 * A warp owns 8 sites and loops over the rates; results are stored as they are produced and
 * re-read for the (rare) x 2^256 rescale, exactly like the reference's per-site loop
 * (reference src/core_partials_avx2.c:788-801).  The summation order inside a DMMA differs
 * from the AVX2 lane order, so CLVs agree with the reference to ~1e-16 relative instead of bit
 * for bit; PLL_GPU_AA_EXACT=1 selects the vector-pipe kernels above, which are bit-exact.
 */
#define PLG_DMMA_THREADS 256

__device__ __forceinline__ double ldg_stream64(const double * p)
{
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

/* the five A-fragment values of one (site, rate), in k-step order (plg_dmma.cuh) */
struct afrag5
{
  double2 lo0, lo1;
  double hi;
};

__device__ __forceinline__ void cp_async16(void * dst_smem, const void * src_gmem)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem))),
               "l"(src_gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

/* Each warp streams its own 8-site units through a private 2-deep ring in shared memory
 * (cp.async, 16-byte chunks, rows padded to PLG_DMMA_PITCH doubles so that the A-fragment
 * reads are bank-conflict free): ~20 KB in flight per warp, independent of register use. */
/* inner-inner: 12 warps, one ring slot each (A fragments of a whole unit are pulled into
 * registers first, so the slot is refilled a full unit ahead); tip-inner: 8 warps x 2 slots,
 * two CTAs per SM */
#define PLG_DMMA_WARPS_II 12
#define PLG_DMMA_WARPS_TI 8
template <int KIND>
struct dmma_cfg
{
  static constexpr int WARPS = (KIND == PLG_KIND_II) ? PLG_DMMA_WARPS_II : PLG_DMMA_WARPS_TI;
  static constexpr int NSLOT = (KIND == PLG_KIND_II) ? 1 : 2;
};
template <int R>
struct dmma_geom
{
  static constexpr int ROW = R * 20;        /* doubles per site                       */
  /* The 8 sites of a unit sit in the ring as two contiguous halves (sites 0..3, sites 4..7), the
   * second half 64 bytes further modulo 128.  Each half is ONE bulk copy (two per child instead of
   * eight row copies: less issue work per unit, no padded rows; measured neutral in time - the
   * copies were not request-bound, tools/ubench/tma_store_peak.cu) and DMMA row g works on site
   * (g & 1) * 4 + (g >> 1): the two sites of a quarter warp then start 64 bytes apart modulo 128
   * and the four lanes of a site read 64 contiguous bytes - conflict free. */
  static constexpr int HALF = 4 * ROW + 8;  /* doubles from the first half to the second */
  static constexpr int UNIT = 8 * ROW + 8;  /* doubles per child per unit (8 sites)      */
};

/* KIND: PLG_KIND_II (two matrix products) or PLG_KIND_TI (tip table x one matrix product) */
template <int R, int KIND>
__global__ void __launch_bounds__(dmma_cfg<KIND>::WARPS * 32, 1)
k_partial_dmma_aa(const DevOp * __restrict__ ops, unsigned int n_ops, unsigned int sites, int scale_mode)
{
  extern __shared__ __align__(16) double smem_d[];
  constexpr int NCHILD = (KIND == PLG_KIND_II) ? 2 : 1;
  constexpr int PLG_DMMA_WARPS = dmma_cfg<KIND>::WARPS;
  constexpr int NSLOT = dmma_cfg<KIND>::NSLOT;
  using G = dmma_geom<R>;
  double * bfrag = smem_d;                                   /* [child][rate][nt][ks][lane] */
  const unsigned int lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned int g = lane >> 2, q = lane & 3u;
  const unsigned int hi_state = dmma_child_state(4, q);
  double * ring = smem_d + NCHILD * R * 15 * 32 + (size_t)warp * NSLOT * NCHILD * G::UNIT; /* [slot][child][8][PITCH] */
  uint64_t * full = reinterpret_cast<uint64_t *>(smem_d + NCHILD * R * 15 * 32 +
                                                 (size_t)PLG_DMMA_WARPS * NSLOT * NCHILD * G::UNIT) + 2 * warp;
  const unsigned int units = (sites + 7) / 8;
  if (lane == 0)
  {
    plg_async::mbar_init(&full[0], 1);
    plg_async::mbar_init(&full[1], 1);
    plg_async::fence_barrier_init();
  }
  __syncwarp();
  unsigned int it = 0; /* this warp's unit counter: ring slot = it % NSLOT, barrier phase = (it / NSLOT) & 1 */

  /* the (operation, unit) space of the whole launch is flattened and cut into one contiguous
   * chunk per CTA (no partial second wave); a chunk is walked one operation segment at a time */
  const unsigned long long total_units = (unsigned long long)n_ops * units;
  const unsigned long long per = (total_units + gridDim.x - 1) / gridDim.x;
  const unsigned long long f_begin = blockIdx.x * per;
  const unsigned long long f_end = (f_begin + per < total_units) ? f_begin + per : total_units;

  for (unsigned long long f_seg = f_begin; f_seg < f_end;)
  {
  const unsigned int op_idx = (unsigned int)(f_seg / units);
  const unsigned int u_seg = (unsigned int)(f_seg - (unsigned long long)op_idx * units);
  const unsigned long long seg_len_ll = ((unsigned long long)units - u_seg < f_end - f_seg)
                                            ? (unsigned long long)units - u_seg : f_end - f_seg;
  const unsigned int u_stop = u_seg + (unsigned int)seg_len_ll; /* units [u_seg, u_stop) of op_idx */
  f_seg += seg_len_ll;
  const DevOp op = ops[op_idx];

  __syncthreads(); /* previous segment's fragments no longer in use */
  /* P matrices -> B fragments (parent states 20..23 of the third N tile are zero padding) */
  for (unsigned int t = threadIdx.x; t < NCHILD * R * 15 * 32; t += blockDim.x)
  {
    const unsigned int l = t & 31u, f = (t >> 5) % 15, kc = (t >> 5) / 15; /* kc = child*R + k */
    const double * M = (NCHILD == 2 && kc < (unsigned)R) ? op.lmat : op.rmat;
    const unsigned int k = kc % R;
    bfrag[t] = dmma_bfrag_value(M + (size_t)k * 400, f, l);
  }
  __syncthreads();
  const double * BL = bfrag;                               /* left child (ii only) */
  const double * BR = bfrag + (NCHILD - 1) * R * 15 * 32;  /* right / inner child  */

  const unsigned int stride = PLG_DMMA_WARPS;
  const unsigned int u_first = u_seg + warp;

  /* TMA copies of unit u (two 4-site halves x NCHILD children) into ring slot `slot`; completion
   * is signalled on the warp's own mbarrier */
  auto fetch = [&](unsigned int u, unsigned int slot) {
    if (u < u_stop)
    {
      const unsigned int first_site = 8 * u;
      const unsigned int nrows = (sites - first_site < 8) ? sites - first_site : 8;
      if (lane == 0) plg_async::mbar_arrive_expect_tx(&full[slot], nrows * NCHILD * G::ROW * 8u);
      __syncwarp();
      if (lane < 2 * NCHILD)
      {
        const unsigned int c = lane >> 1, h = lane & 1u;
        if (nrows > 4 * h)
        {
          const unsigned int rows = (nrows - 4 * h < 4) ? nrows - 4 * h : 4;
          const double * src = ((NCHILD == 2 && c == 0) ? op.left : op.right) + (size_t)(first_site + 4 * h) * G::ROW;
          double * dst = ring + ((size_t)slot * NCHILD + c) * G::UNIT + h * G::HALF;
          plg_async::bulk_g2s(dst, src, rows * G::ROW * 8u, &full[slot]);
        }
      }
    }
  };

  fetch(u_first, it % NSLOT);
  if (NSLOT == 2) fetch(u_first + stride, (it + 1) % NSLOT);
  for (unsigned int u = u_first; u < u_stop; u += stride, ++it)
  {
    const unsigned int slot = it % NSLOT;
    const unsigned int srow = (g & 1u) * 4u + (g >> 1);
    const unsigned int site = 8 * u + srow; /* this lane's site: A rows and D rows alike */
    const bool ok = site < sites;
    const size_t site_off = (size_t)site * G::ROW;

    unsigned int child_sum = 0, code = 0;
    if (scale_mode == 1 && q == 0 && ok)
    {
      if (KIND == PLG_KIND_II && op.lscale) child_sum += op.lscale[site];
      if (op.rscale) child_sum += op.rscale[site];
    }
    if (KIND == PLG_KIND_TI && ok) code = __ldg(op.ltip + site);

    plg_async::mbar_wait(&full[slot], (it / NSLOT) & 1u); /* this unit's rows have landed */
    const unsigned int srow_off = (g & 1u) * G::HALF + (g >> 1) * G::ROW;
    const double * rowL = ring + ((size_t)slot * NCHILD + 0) * G::UNIT + srow_off;
    const double * rowR = ring + ((size_t)slot * NCHILD + (NCHILD - 1)) * G::UNIT + srow_off;

    /* single-slot ring: pull the whole unit's A fragments into registers, then hand the slot
     * straight back to the TMA for the next unit */
    afrag5 pre_r[NSLOT == 1 ? R : 1], pre_l[NSLOT == 1 ? R : 1];
    if (NSLOT == 1)
    {
#pragma unroll
      for (int k = 0; k < R; ++k)
      {
        pre_r[k].lo0 = *reinterpret_cast<const double2 *>(rowR + k * 20 + 2 * q);
        pre_r[k].lo1 = *reinterpret_cast<const double2 *>(rowR + k * 20 + 8 + 2 * q);
        pre_r[k].hi = rowR[k * 20 + hi_state];
        if (KIND == PLG_KIND_II)
        {
          pre_l[k].lo0 = *reinterpret_cast<const double2 *>(rowL + k * 20 + 2 * q);
          pre_l[k].lo1 = *reinterpret_cast<const double2 *>(rowL + k * 20 + 8 + 2 * q);
          pre_l[k].hi = rowL[k * 20 + hi_state];
        }
      }
      __syncwarp();
      fetch(u + stride, slot);
    }

    bool below_site = true;
#pragma unroll
    for (int k = 0; k < R; ++k)
    {
      afrag5 ar, al = {};
      if (NSLOT == 1)
      {
        ar = pre_r[k];
        if (KIND == PLG_KIND_II) al = pre_l[k];
      }
      else
      {
        ar.lo0 = *reinterpret_cast<const double2 *>(rowR + k * 20 + 2 * q);
        ar.lo1 = *reinterpret_cast<const double2 *>(rowR + k * 20 + 8 + 2 * q);
        ar.hi = rowR[k * 20 + hi_state];
      }

      /* six independent accumulator chains (3 N tiles x {left, right}) advance one k-step
       * at a time, so that consecutive DMMAs never depend on each other */
      double y[3][2], x[3][2];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) y[nt][0] = y[nt][1] = x[nt][0] = x[nt][1] = 0.0;
      const double av_r[5] = {ar.lo0.x, ar.lo0.y, ar.lo1.x, ar.lo1.y, ar.hi};
      const double av_l[5] = {al.lo0.x, al.lo0.y, al.lo1.x, al.lo1.y, al.hi};
#pragma unroll
      for (int ks = 0; ks < 5; ++ks)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
        {
          dmma884(y[nt][0], y[nt][1], av_r[ks], BR[((size_t)k * 15 + nt * 5 + ks) * 32 + lane]);
          if (KIND == PLG_KIND_II)
            dmma884(x[nt][0], x[nt][1], av_l[ks], BL[((size_t)k * 15 + nt * 5 + ks) * 32 + lane]);
        }

      double p[3][2];
      bool below = true;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        const unsigned int row = 8 * nt + 2 * q; /* first of this lane's two parent states */
        if (KIND == PLG_KIND_TI && row < 20)
        {
          const double2 t = __ldg(reinterpret_cast<const double2 *>(op.lmat + ((size_t)code * R + k) * 20 + row));
          x[nt][0] = t.x;
          x[nt][1] = t.y;
        }
        p[nt][0] = __dmul_rn(x[nt][0], y[nt][0]);
        p[nt][1] = __dmul_rn(x[nt][1], y[nt][1]);
        if (row < 20)
        {
          below = below && (p[nt][0] < PLG_SCALE_THRESHOLD) && (p[nt][1] < PLG_SCALE_THRESHOLD);
          if (ok) st_stream2(op.parent + site_off + k * 20 + row, p[nt][0], p[nt][1]);
        }
      }
      if (scale_mode == 2)
      {
        /* per-rate: all 20 states of (site, rate) below -> rescale that block now */
        const unsigned int m = __ballot_sync(0xffffffffu, below);
        const bool sc = ((m >> (lane & ~3u)) & 0xFu) == 0xFu;
        if (sc && ok)
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
          {
            const unsigned int row = 8 * nt + 2 * q;
            if (row < 20)
              st_stream2(op.parent + site_off + k * 20 + row, __dmul_rn(p[nt][0], PLG_SCALE_FACTOR),
                         __dmul_rn(p[nt][1], PLG_SCALE_FACTOR));
          }
        if (q == 0 && ok)
        {
          const size_t e = (size_t)site * R + k;
          op.pscale[e] = (sc ? 1u : 0u) + ((KIND == PLG_KIND_II && op.lscale) ? op.lscale[e] : 0u) +
                         (op.rscale ? op.rscale[e] : 0u);
        }
      }
      below_site = below_site && below;
    }

    /* the slot is free again: start fetching the unit after next into it */
    if (NSLOT == 2)
    {
      __syncwarp();
      fetch(u + 2 * stride, slot);
    }

    if (scale_mode == 1)
    {
      /* per-site: every entry (all rates, all states) of the site below the threshold; the
       * four lanes of a site hold all of them */
      const unsigned int m = __ballot_sync(0xffffffffu, below_site);
      const bool sc = ((m >> (lane & ~3u)) & 0xFu) == 0xFu;
      if (sc && ok)
      {
        /* rare: re-read what this lane stored and scale it (reference rescales in place too) */
#pragma unroll 1
        for (int k = 0; k < R; ++k)
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
          {
            const unsigned int row = 8 * nt + 2 * q;
            if (row < 20)
            {
              double2 * dst = reinterpret_cast<double2 *>(op.parent + site_off + k * 20 + row);
              double2 v = *dst;
              v.x = __dmul_rn(v.x, PLG_SCALE_FACTOR);
              v.y = __dmul_rn(v.y, PLG_SCALE_FACTOR);
              *dst = v;
            }
          }
      }
      if (q == 0 && ok) op.pscale[site] = child_sum + (sc ? 1u : 0u);
    }
  }
  __syncwarp();
  } /* segment */
}

template <int R, int KIND>
static constexpr size_t dmma_smem_bytes()
{
  constexpr int NCHILD = (KIND == PLG_KIND_II) ? 2 : 1;
  return ((size_t)NCHILD * R * 15 * 32 +
          (size_t)dmma_cfg<KIND>::WARPS * dmma_cfg<KIND>::NSLOT * NCHILD * dmma_geom<R>::UNIT) * sizeof(double) +
         (size_t)dmma_cfg<KIND>::WARPS * 2 * sizeof(uint64_t);
}

/* ------------------------------------------------------------------------------------ */
/* host side: levelisation, batching, launch                                             */
/* ------------------------------------------------------------------------------------ */
struct Group
{
  int kind;
  int scale_mode;
  unsigned int first; /* index into the sorted DevOp array */
  unsigned int count;
  unsigned long long bytes; /* algorithmic bytes of the group */
};

struct Plan
{
  std::vector<DevOp> ops;      /* sorted by (level, kind, scale_mode) */
  std::vector<TableJob> jobs;  /* tip lookup tables to build first */
  std::vector<Group> groups;
  std::vector<FusedOp> fused;  /* DNA: the same list for the single-kernel traversal */
  unsigned int fused_hits, fused_misses;
  unsigned int fused_nslot;    /* tile-cache slots per warp the plan was made for */
  size_t fused_scratch_bytes;  /* device scratch of the fused path: packed records (+ tip tables) */
  size_t fused_table_offset;   /* 20-state walk: where the tip tables start inside that scratch */
  unsigned long long levels;
  unsigned long long algorithmic_bytes;
  /* bytes that MUST cross the HBM interface for this list on the path that runs it: level by
   * level that is the algorithmic figure; the fused traversal writes every observable parent
   * CLV / scaler once and reads tip characters and tile-cache misses only */
  unsigned long long compulsory_bytes;
  size_t table_doubles;
};

static int build_plan(plg_context * ctx, const pll_operation_t * operations, unsigned int count,
                      Plan & plan)
{
  const unsigned int n_clv = ctx->d.tips + ctx->d.clv_buffers;
  const unsigned int n_sc = ctx->d.scale_buffers;
  const unsigned int K = ctx->d.states;
  const unsigned int R = ctx->d.rate_cats;
  const size_t table_len = (size_t)(K == 4 ? 16u : ctx->maxstates) * R * ctx->d.states_padded;

  /* dependency levels: level(op) > level of every earlier op it has a RAW, WAR or WAW
   * relation with, on CLV slots and on scaler slots.  Executing levels in increasing order
   * is therefore equivalent to the reference's strictly sequential loop. */
  std::vector<int> clv_w(n_clv, -1), clv_r(n_clv, -1), sc_w(n_sc, -1), sc_r(n_sc, -1);
  /* the operation (list index) that last wrote a slot, and per operation the producers of what it
   * reads: lets the reordering below be checked against ALL read-after-write relations, also
   * those that exist through a scaler only */
  std::vector<int> clv_writer(n_clv, -1), sc_writer(n_sc, -1);
  std::vector<std::array<int, 4>> producers(count);

  struct Item
  {
    int level, kind, scale_mode;
    unsigned long long bytes;
    DevOp op;
    const double * src_l; /* P-matrix sets behind the left / right term (fused path) */
    const double * src_r;
  };
  std::vector<Item> items(count);
  size_t n_tables = 0;
  int max_level = -1;
  plan.algorithmic_bytes = 0;
  bool recycled = false; /* some CLV / scaler slot is written after an earlier read or write */

  const size_t span_bytes = ctx->span * sizeof(double);
  const size_t scaler_unit = ctx->rate_scalers ? 4u * R : 4u;

  for (unsigned int i = 0; i < count; ++i)
  {
    const pll_operation_t & o = operations[i];
    if (o.parent_clv_index >= n_clv || o.child1_clv_index >= n_clv ||
        o.child2_clv_index >= n_clv || o.parent_clv_index < ctx->clv_first ||
        o.child1_matrix_index >= ctx->d.prob_matrices ||
        o.child2_matrix_index >= ctx->d.prob_matrices ||
        o.parent_scaler_index >= (int)n_sc || o.child1_scaler_index >= (int)n_sc ||
        o.child2_scaler_index >= (int)n_sc)
    {
      plg_set_error("plg_update_partials: operation %u has an index out of range", i);
      return PLG_E_INVALID;
    }
    const bool t1 = plg_is_tip(ctx, o.child1_clv_index);
    const bool t2 = plg_is_tip(ctx, o.child2_clv_index);
    Item & it = items[i];
    memset(&it.op, 0, sizeof(DevOp));
    it.kind = (t1 && t2) ? PLG_KIND_TT : ((t1 || t2) ? PLG_KIND_TI : PLG_KIND_II);
    it.op.parent = plg_clv_ptr(ctx, o.parent_clv_index);
    it.op.pscale = plg_scaler_ptr(ctx, o.parent_scaler_index);
    it.scale_mode = it.op.pscale ? (ctx->rate_scalers ? 2 : 1) : 0;

    int level = 0;
    auto after = [&](int l) { if (l + 1 > level) level = l + 1; };
    /* parent slot: WAW and WAR */
    after(clv_w[o.parent_clv_index]);
    after(clv_r[o.parent_clv_index]);
    if (clv_w[o.parent_clv_index] >= 0 || clv_r[o.parent_clv_index] >= 0) recycled = true;
    if (it.op.pscale)
    {
      after(sc_w[o.parent_scaler_index]);
      after(sc_r[o.parent_scaler_index]);
      if (sc_w[o.parent_scaler_index] >= 0 || sc_r[o.parent_scaler_index] >= 0) recycled = true;
    }

    int read_clv[2] = {-1, -1}, read_sc[2] = {-1, -1};
    size_t bytes = span_bytes; /* parent store */
    if (it.kind == PLG_KIND_II)
    {
      it.op.left = plg_clv_ptr(ctx, o.child1_clv_index);
      it.op.right = plg_clv_ptr(ctx, o.child2_clv_index);
      it.op.lmat = plg_pmat_ptr(ctx, o.child1_matrix_index);
      it.op.rmat = plg_pmat_ptr(ctx, o.child2_matrix_index);
      it.src_l = it.op.lmat;
      it.src_r = it.op.rmat;
      read_clv[0] = (int)o.child1_clv_index;
      read_clv[1] = (int)o.child2_clv_index;
      if (it.op.pscale)
      {
        /* children scalers are only consulted when the parent has one
         * (reference src/core_partials_avx.c:397-410) */
        it.op.lscale = plg_scaler_ptr(ctx, o.child1_scaler_index);
        it.op.rscale = plg_scaler_ptr(ctx, o.child2_scaler_index);
        if (it.op.lscale) read_sc[0] = o.child1_scaler_index;
        if (it.op.rscale) read_sc[1] = o.child2_scaler_index;
      }
      bytes += 2 * span_bytes;
    }
    else if (it.kind == PLG_KIND_TI)
    {
      /* which child is the tip (reference src/partials.c:91-112); the tip's scaler index
       * is ignored */
      const unsigned int tip = t1 ? o.child1_clv_index : o.child2_clv_index;
      const unsigned int inner = t1 ? o.child2_clv_index : o.child1_clv_index;
      const unsigned int tip_mat = t1 ? o.child1_matrix_index : o.child2_matrix_index;
      const unsigned int inner_mat = t1 ? o.child2_matrix_index : o.child1_matrix_index;
      const int inner_sc = t1 ? o.child2_scaler_index : o.child1_scaler_index;
      it.op.ltip = plg_tip_ptr(ctx, tip);
      it.op.right = plg_clv_ptr(ctx, inner);
      it.op.rmat = plg_pmat_ptr(ctx, inner_mat);
      it.op.lmat = (const double *)(uintptr_t)(n_tables * table_len); /* offset, fixed up below */
      plan.jobs.push_back(TableJob{plg_pmat_ptr(ctx, tip_mat), (double *)(uintptr_t)(n_tables * table_len)});
      ++n_tables;
      it.src_l = plg_pmat_ptr(ctx, tip_mat);
      it.src_r = it.op.rmat;
      read_clv[0] = (int)inner;
      if (it.op.pscale)
      {
        it.op.rscale = plg_scaler_ptr(ctx, inner_sc);
        if (it.op.rscale) read_sc[0] = inner_sc;
      }
      bytes += span_bytes + 1;
    }
    else
    {
      it.op.ltip = plg_tip_ptr(ctx, o.child1_clv_index);
      it.op.rtip = plg_tip_ptr(ctx, o.child2_clv_index);
      it.op.lmat = (const double *)(uintptr_t)(n_tables * table_len);
      plan.jobs.push_back(TableJob{plg_pmat_ptr(ctx, o.child1_matrix_index), (double *)(uintptr_t)(n_tables * table_len)});
      ++n_tables;
      it.op.rmat = (const double *)(uintptr_t)(n_tables * table_len);
      plan.jobs.push_back(TableJob{plg_pmat_ptr(ctx, o.child2_matrix_index), (double *)(uintptr_t)(n_tables * table_len)});
      ++n_tables;
      it.src_l = plg_pmat_ptr(ctx, o.child1_matrix_index);
      it.src_r = plg_pmat_ptr(ctx, o.child2_matrix_index);
      bytes += 2;
    }
    if (it.op.pscale) bytes += scaler_unit;
    if (it.op.lscale) bytes += scaler_unit;
    if (it.op.rscale) bytes += scaler_unit;
    it.bytes = (unsigned long long)bytes * ctx->d.sites;
    plan.algorithmic_bytes += it.bytes;

    for (int c = 0; c < 2; ++c)
    {
      if (read_clv[c] >= 0) after(clv_w[read_clv[c]]);
      if (read_sc[c] >= 0) after(sc_w[read_sc[c]]);
      producers[i][c] = read_clv[c] >= 0 ? clv_writer[read_clv[c]] : -1;
      producers[i][2 + c] = read_sc[c] >= 0 ? sc_writer[read_sc[c]] : -1;
    }
    clv_writer[o.parent_clv_index] = (int)i;
    if (it.op.pscale) sc_writer[o.parent_scaler_index] = (int)i;
    /* `after(l)` left level = 1 + max over predecessors (0 when there is none) */
    it.level = level;

    clv_w[o.parent_clv_index] = level;
    if (it.op.pscale) sc_w[o.parent_scaler_index] = level;
    for (int c = 0; c < 2; ++c)
    {
      if (read_clv[c] >= 0 && clv_r[read_clv[c]] < level) clv_r[read_clv[c]] = level;
      if (read_sc[c] >= 0 && sc_r[read_sc[c]] < level) sc_r[read_sc[c]] = level;
    }
    if (level > max_level) max_level = level;
  }
  plan.levels = (unsigned long long)(max_level + 1);
  plan.table_doubles = n_tables * table_len;
  plan.compulsory_bytes = plan.algorithmic_bytes;

  /* stable sort by (level, kind, scale_mode) -> contiguous groups */
  std::vector<unsigned int> order(count);
  for (unsigned int i = 0; i < count; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](unsigned int a, unsigned int b) {
    if (items[a].level != items[b].level) return items[a].level < items[b].level;
    if (items[a].kind != items[b].kind) return items[a].kind < items[b].kind;
    return items[a].scale_mode < items[b].scale_mode;
  });
  plan.ops.resize(count);
  for (unsigned int i = 0; i < count; ++i)
  {
    const Item & it = items[order[i]];
    plan.ops[i] = it.op;
    const bool new_group = plan.groups.empty() || i == 0 ||
                           items[order[i - 1]].level != it.level ||
                           items[order[i - 1]].kind != it.kind ||
                           items[order[i - 1]].scale_mode != it.scale_mode ||
                           plan.groups.back().count >= 65535u;
    if (new_group) plan.groups.push_back(Group{it.kind, it.scale_mode, i, 0, 0});
    plan.groups.back().count++;
    plan.groups.back().bytes += it.bytes;
  }

  /* ---- the single-kernel traversal (plg_traverse.cu): execution order + tile cache ---- */
  plan.fused.clear();
  plan.fused_hits = plan.fused_misses = 0;
  /* 20 states: the tensor-core walk (plg_walk_aa.cu) - 1, 2 or 4 rate categories, per-site scalers
   * or none, not in bit-exact mode */
  const bool aa_walk = K == 20 && (ctx->use_fused_aa == 1 || (ctx->use_fused_aa == 2 && recycled)) &&
                       !ctx->aa_exact && !ctx->rate_scalers &&
                       plg_walk_aa_supported(R, ctx->pattern_tip ? ctx->maxstates : 1u);
  const unsigned int aa_slots =
      aa_walk ? (ctx->fused_slots < PLG_WALK_AA_SLOTS ? ctx->fused_slots : PLG_WALK_AA_SLOTS) : 0;
  plan.fused_scratch_bytes = 0;
  plan.fused_table_offset = 0;
  if (ctx->use_fused && (K == 4 || aa_slots > 0) && plg_fast_path(ctx) && count >= 2)
  {
    /* Execution order.  A list without slot recycling is a forest: walk it depth-first, the
     * larger subtree first, so that few results are waiting for their parent at any time.
     * A list that recycles slots is executed as given (its order is part of its meaning). */
    std::vector<unsigned int> exec;
    exec.reserve(count);
    if (recycled)
      for (unsigned int i = 0; i < count; ++i) exec.push_back(i);
    else
    {
      std::vector<int> producer(n_clv, -1), kid0(count, -1), kid1(count, -1);
      std::vector<unsigned int> size(count, 1);
      std::vector<char> consumed(count, 0);
      for (unsigned int i = 0; i < count; ++i) producer[operations[i].parent_clv_index] = (int)i;
      for (unsigned int i = 0; i < count; ++i)
      {
        /* children are produced EARLIER in the list (it is a valid sequential program) */
        const int a = producer[operations[i].child1_clv_index], b = producer[operations[i].child2_clv_index];
        if (a >= 0 && (unsigned int)a < i && !plg_is_tip(ctx, operations[i].child1_clv_index)) kid0[i] = a;
        if (b >= 0 && (unsigned int)b < i && !plg_is_tip(ctx, operations[i].child2_clv_index)) kid1[i] = b;
        size[i] = 1 + (kid0[i] >= 0 ? size[kid0[i]] : 0) + (kid1[i] >= 0 ? size[kid1[i]] : 0);
        if (kid0[i] >= 0) consumed[kid0[i]] = 1;
        if (kid1[i] >= 0) consumed[kid1[i]] = 1;
      }
      std::vector<std::pair<unsigned int, int>> stack;
      for (unsigned int root = 0; root < count; ++root)
      {
        if (consumed[root]) continue;
        stack.push_back({root, 0});
        while (!stack.empty())
        {
          auto [n, state] = stack.back();
          stack.pop_back();
          if (state == 1)
          {
            exec.push_back(n);
            continue;
          }
          stack.push_back({n, 1});
          int first = kid0[n], second = kid1[n];
          if (second >= 0 && (first < 0 || size[second] > size[first])) std::swap(first, second);
          if (second >= 0) stack.push_back({(unsigned int)second, 0});
          if (first >= 0) stack.push_back({(unsigned int)first, 0}); /* popped next: runs first */
        }
      }
      /* The walk follows CLV edges only.  Keep it only if it is a permutation (no CLV consumed
       * twice) in which every operation still runs after the producers of all it reads - a list
       * may also chain operations through a scaler alone; otherwise execute the list as given. */
      bool valid = exec.size() == count;
      if (valid)
      {
        std::vector<unsigned int> position(count, count);
        for (unsigned int x = 0; x < count; ++x) position[exec[x]] = x;
        for (unsigned int i = 0; i < count && valid; ++i)
        {
          if (position[i] == count) valid = false;
          for (int c = 0; c < 4 && valid; ++c)
            if (producers[i][c] >= 0 && position[producers[i][c]] > position[i]) valid = false;
        }
      }
      if (!valid)
      {
        exec.clear();
        for (unsigned int i = 0; i < count; ++i) exec.push_back(i);
      }
    }

    /* Tile cache of a warp, simulated here: slot tags are (CLV address, scaler address). */
    const unsigned int nslot = (K == 4) ? ctx->fused_slots : aa_slots;
    plan.fused_nslot = nslot;
    std::vector<const double *> tag_clv(nslot, nullptr);
    std::vector<const unsigned int *> tag_sc(nslot, nullptr);
    std::vector<unsigned long long> born(nslot, 0);
    unsigned long long clock = 0;
    auto lookup = [&](const double * clv, const unsigned int * sc) -> int {
      if (!clv) return -1;
      for (unsigned int q = 0; q < nslot; ++q)
        if (tag_clv[q] == clv && (!sc || tag_sc[q] == sc))
        {
          tag_clv[q] = nullptr; /* consumed: the slot is free again (also for this op's result) */
          tag_sc[q] = nullptr;
          return (int)q;
        }
      return -1;
    };
    /* Register forwarding: in depth-first order the child finished LAST is the operation right
     * before its parent (always so for the inner child of a tip-inner operation and for one
     * child of an inner-inner operation); the kernel hands that tile over in registers
     * (slot -2), it never enters the shared-memory cache. */
    std::vector<signed char> forward(count, 0); /* 1: left child is the previous result, 2: right */
    for (unsigned int x = 1; x < count; ++x)
    {
      Item & it = items[exec[x]];
      const Item & pv = items[exec[x - 1]];
      if (it.kind == PLG_KIND_II && it.op.left == pv.op.parent && (!it.op.lscale || it.op.lscale == pv.op.pscale))
      {
        forward[x] = 1;
        if (K == 20)
        {
          /* the tensor-core traversal has ONE straight-line inner-inner path (the right child
           * arrives in registers): the two sides of the product commute bit for bit, so the
           * children simply change places */
          std::swap(it.op.left, it.op.right);
          std::swap(it.op.lmat, it.op.rmat);
          std::swap(it.op.lscale, it.op.rscale);
          std::swap(it.src_l, it.src_r);
          forward[x] = 2;
        }
      }
      else if (it.kind != PLG_KIND_TT && it.op.right == pv.op.parent &&
               (!it.op.rscale || it.op.rscale == pv.op.pscale))
        forward[x] = 2;
    }
    /* Dead stores: with slot recycling a CLV / scaler buffer is overwritten several times within
     * one list.  Only the LAST value of a buffer is observable after the call; an earlier one
     * must reach HBM only if some operation of the list reads it from there (tile-cache miss).
     * `pad` = 1 marks the operations whose result is written through. */
    std::unordered_map<const void *, unsigned int> last_clv, last_sc;
    /* a child read back from HBM needs BOTH its CLV and its scaler there, and with slot
     * recycling the two may have been written by different operations */
    auto need_hbm = [&](const void * clv_buffer, const void * sc_buffer) {
      auto w = last_clv.find(clv_buffer);
      if (w != last_clv.end()) plan.fused[w->second].pad |= 1;
      if (sc_buffer)
      {
        /* bit 1: this scaler is read back from HBM later in the list (the 20-state kernel then
         * lets every rate warp store it, so that each warp re-reads its own store) */
        auto v = last_sc.find(sc_buffer);
        if (v != last_sc.end()) plan.fused[v->second].pad |= 3;
      }
    };
    plan.fused.resize(count);
    for (unsigned int x = 0; x < count; ++x)
    {
      const Item & it = items[exec[x]];
      FusedOp & f = plan.fused[x];
      memset(&f, 0, sizeof(f));
      f.op = it.op;
      f.kind = it.kind;
      f.scale_mode = it.scale_mode;
      {
        const unsigned int mat_bytes = R * 18u * 8u, table_bytes = (unsigned int)plg_fused_block_bytes(R);
        f.lbytes = (it.kind == PLG_KIND_II) ? mat_bytes : table_bytes;
        f.rbytes = (it.kind == PLG_KIND_TT) ? table_bytes : mat_bytes;
        f.lsrc = it.src_l;
        f.rsrc = it.src_r;
      }
      f.lslot = (forward[x] == 1) ? -2 : ((it.kind == PLG_KIND_II) ? lookup(it.op.left, it.op.lscale) : -1);
      f.rslot = (forward[x] == 2) ? -2 : ((it.kind != PLG_KIND_TT) ? lookup(it.op.right, it.op.rscale) : -1);
      if (it.kind == PLG_KIND_II) (f.lslot != -1 ? plan.fused_hits : plan.fused_misses)++;
      if (it.kind != PLG_KIND_TT) (f.rslot != -1 ? plan.fused_hits : plan.fused_misses)++;
      if (it.kind == PLG_KIND_II && f.lslot == -1) need_hbm(it.op.left, it.op.lscale);
      if (it.kind != PLG_KIND_TT && f.rslot == -1) need_hbm(it.op.right, it.op.rscale);
      last_clv[it.op.parent] = x;
      if (it.op.pscale) last_sc[it.op.pscale] = x;
      /* stale copies of what this operation overwrites */
      for (unsigned int q = 0; q < nslot; ++q)
        if (tag_clv[q] == it.op.parent || (it.op.pscale && tag_sc[q] == it.op.pscale))
        {
          tag_clv[q] = nullptr;
          tag_sc[q] = nullptr;
        }
      if (x + 1 < count && forward[x + 1])
      {
        f.pslot = -1; /* handed to the next operation in registers */
        continue;
      }
      int dst = -1;
      for (unsigned int q = 0; q < nslot && dst < 0; ++q)
        if (!tag_clv[q]) dst = (int)q;
      if (dst < 0)
      {
        dst = 0; /* full: drop the oldest waiting tile (it is in HBM anyway) */
        for (unsigned int q = 1; q < nslot; ++q)
          if (born[q] < born[dst]) dst = (int)q;
      }
      f.pslot = dst;
      tag_clv[dst] = it.op.parent;
      tag_sc[dst] = it.op.pscale;
      born[dst] = ++clock;
    }
    for (const auto & kv : last_clv) plan.fused[kv.second].pad |= 1;
    for (const auto & kv : last_sc) plan.fused[kv.second].pad |= 1;

    unsigned long long per_site = 0;
    for (const FusedOp & f : plan.fused)
    {
      if (f.pad & 1) per_site += span_bytes + (f.op.pscale ? scaler_unit : 0);
      if (f.kind == PLG_KIND_II && f.lslot == -1) per_site += span_bytes + (f.op.lscale ? scaler_unit : 0);
      if (f.kind != PLG_KIND_TT && f.rslot == -1) per_site += span_bytes + (f.op.rscale ? scaler_unit : 0);
      per_site += (f.kind == PLG_KIND_TT) ? 2 : (f.kind == PLG_KIND_TI ? 1 : 0);
    }
    plan.compulsory_bytes = per_site * ctx->d.sites;

    if (aa_walk)
    {
      const unsigned int mat = 3200u * R;
      unsigned long long rows = 0;
      const unsigned long long nc = ctx->maxstates;
      for (FusedOp & f : plan.fused)
      {
        f.lbytes = 128u + (f.kind == PLG_KIND_II ? 2u * mat : (f.kind == PLG_KIND_TI ? mat : 0u));
        if (rows > 0xffffffffull)
        {
          plg_set_error("plg_update_partials: tip tables of this list exceed the 20-state walk's index range");
          return PLG_E_UNSUPPORTED;
        }
        f.rbytes = (unsigned int)rows;
        rows += f.kind == PLG_KIND_TT ? nc * nc : (f.kind == PLG_KIND_TI ? nc : 0ull);
      }
      /* the ring of the walk is refilled by whichever warp frees a stage last: the size of the
       * record two operations ahead travels in this operation's descriptor (upper half) */
      const size_t nf = plan.fused.size();
      for (size_t x = 0; x < nf; ++x) plan.fused[x].lbytes |= (plan.fused[(x + 2) % nf].lbytes & 0xffffu) << 16;
      plan.fused_table_offset = plan.fused.size() * plg_walk_aa_record_bytes(R);
      plan.fused_scratch_bytes = plan.fused_table_offset + (size_t)rows * plg_walk_aa_row_bytes(R);
    }
    else
      plan.fused_scratch_bytes = plan.fused.size() * plg_fused_record_bytes(R);
  }
  return PLG_OK;
}

template <int R>
static void launch_group(plg_context * ctx, const Group & g, const DevOp * dev_ops,
                         unsigned int nelem)
{
  const DevOp * ops = dev_ops + g.first;
  if (ctx->d.states == 4)
  {
    if (g.kind == PLG_KIND_II || g.kind == PLG_KIND_TI)
    {
      unsigned int blocks =
          (unsigned int)ctx->sm_count * (g.kind == PLG_KIND_II ? PLG_II_MINB : PLG_TI_MINB);
      const unsigned int tile_e = (g.kind == PLG_KIND_II) ? PLG_II_TILE : PLG_TI_TILE;
      const unsigned int ntiles = (nelem + tile_e - 1) / tile_e;
      const unsigned long long total = (unsigned long long)ntiles * g.count;
      if (total < blocks) blocks = (unsigned int)total;
      if (g.kind == PLG_KIND_II)
      {
        const size_t smem = sizeof(StreamSmem<R, PLG_KIND_II, PLG_II_STAGES>);
        k_partial_stream_dna<R, PLG_KIND_II, PLG_II_STAGES, PLG_II_MINB>
            <<<blocks, PLG_STREAM_THREADS, smem, ctx->stream>>>(ops, nelem, ntiles, (unsigned int)total,
                                                               g.scale_mode);
      }
      else
      {
        const size_t smem = sizeof(StreamSmem<R, PLG_KIND_TI, PLG_TI_STAGES>);
        k_partial_stream_dna<R, PLG_KIND_TI, PLG_TI_STAGES, PLG_TI_MINB>
            <<<blocks, PLG_STREAM_THREADS, smem, ctx->stream>>>(ops, nelem, ntiles, (unsigned int)total,
                                                               g.scale_mode);
      }
    }
    else
    {
      constexpr int ITEMS = 8;
      dim3 grid((2 * nelem + PLG_DNA_THREADS * ITEMS - 1) / (PLG_DNA_THREADS * ITEMS), g.count);
      k_partial_tt_dna<R, ITEMS><<<grid, PLG_DNA_THREADS, 0, ctx->stream>>>(ops, nelem, g.scale_mode);
    }
  }
  else if (!ctx->aa_exact && g.kind != PLG_KIND_TT && dmma_smem_bytes<R, PLG_KIND_II>() <= 227 * 1024)
  {
    /* tensor-core path: persistent CTAs over the flattened (operation, 8-site unit) space;
     * resident CTAs per SM by shared memory: 1 (inner-inner) or 2 (tip-inner) */
    const unsigned int units = (ctx->d.sites + 7) / 8;
    const unsigned long long total = (unsigned long long)units * g.count;
    const unsigned int warps = (g.kind == PLG_KIND_II) ? PLG_DMMA_WARPS_II : PLG_DMMA_WARPS_TI;
    unsigned long long blocks = (unsigned long long)ctx->sm_count * (g.kind == PLG_KIND_II ? 1u : 2u);
    const unsigned long long want = (total + warps - 1) / warps;
    if (want < blocks) blocks = want ? want : 1;
    if (g.kind == PLG_KIND_II)
      k_partial_dmma_aa<R, PLG_KIND_II><<<(unsigned int)blocks, warps * 32, dmma_smem_bytes<R, PLG_KIND_II>(),
                                          ctx->stream>>>(ops, g.count, ctx->d.sites, g.scale_mode);
    else
      k_partial_dmma_aa<R, PLG_KIND_TI><<<(unsigned int)blocks, warps * 32, dmma_smem_bytes<R, PLG_KIND_TI>(),
                                          ctx->stream>>>(ops, g.count, ctx->d.sites, g.scale_mode);
  }
  else
  {
    dim3 grid((nelem + PLG_AA_THREADS - 1) / PLG_AA_THREADS, g.count);
    if (g.kind == PLG_KIND_II)
    {
      const size_t smem = (size_t)2 * R * PLG_AA_MSTRIDE * sizeof(double);
      k_partial_ii_aa<R><<<grid, PLG_AA_THREADS, smem, ctx->stream>>>(ops, nelem, g.scale_mode);
    }
    else if (g.kind == PLG_KIND_TI)
    {
      const size_t smem = (size_t)R * PLG_AA_MSTRIDE * sizeof(double);
      k_partial_ti_aa<R><<<grid, PLG_AA_THREADS, smem, ctx->stream>>>(ops, nelem, g.scale_mode);
    }
    else
    {
      const unsigned long long nchunks = (unsigned long long)nelem * 10ull;
      dim3 gtt((unsigned int)((nchunks + 256 * PLG_AA_TT_ITEMS - 1) / (256 * PLG_AA_TT_ITEMS)), g.count);
      k_partial_tt_aa<R><<<gtt, 256, 0, ctx->stream>>>(ops, nelem, g.scale_mode);
    }
  }
}

template <int R>
static int set_smem_limits()
{
  /* opt in to > 48 KB of dynamic shared memory: the DNA streaming rings and the 20-state
   * kernels' P-matrix sets (2*R*3200 bytes) */
  static bool done[PLG_MAX_DEVICES] = {}; /* function attributes are per device */
  int dev = 0;
  PLG_CUDA(cudaGetDevice(&dev));
  if (done[dev % PLG_MAX_DEVICES]) return PLG_OK;
  PLG_CUDA(cudaFuncSetAttribute(k_partial_stream_dna<R, PLG_KIND_II, PLG_II_STAGES, PLG_II_MINB>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(StreamSmem<R, PLG_KIND_II, PLG_II_STAGES>)));
  PLG_CUDA(cudaFuncSetAttribute(k_partial_stream_dna<R, PLG_KIND_TI, PLG_TI_STAGES, PLG_TI_MINB>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(StreamSmem<R, PLG_KIND_TI, PLG_TI_STAGES>)));
  if (dmma_smem_bytes<R, PLG_KIND_II>() <= 227 * 1024)
  {
    PLG_CUDA(cudaFuncSetAttribute(k_partial_dmma_aa<R, PLG_KIND_II>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)dmma_smem_bytes<R, PLG_KIND_II>()));
    PLG_CUDA(cudaFuncSetAttribute(k_partial_dmma_aa<R, PLG_KIND_TI>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)dmma_smem_bytes<R, PLG_KIND_TI>()));
  }
  PLG_CUDA(cudaFuncSetAttribute(k_partial_ii_aa<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                2 * R * PLG_AA_MSTRIDE * (int)sizeof(double)));
  PLG_CUDA(cudaFuncSetAttribute(k_partial_ti_aa<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                R * PLG_AA_MSTRIDE * (int)sizeof(double)));
  done[dev % PLG_MAX_DEVICES] = true;
  return PLG_OK;
}

static int enqueue_plan(plg_context * ctx, const Plan & plan, const void * dev_payload,
                        const TableJob * dev_jobs, unsigned long long * kernels, bool fused,
                        unsigned char * dev_records)
{
  const DevOp * dev_ops = (const DevOp *)dev_payload;
  const unsigned int R = ctx->d.rate_cats;
  const unsigned int nelem = ctx->d.sites * R;
  unsigned long long launched = 0;

  const bool fast = plg_fast_path(ctx);
  if (!plan.jobs.empty() && !fused)
  {
    if (!fast)
    {
      int rc = plg_gen_tables(ctx, dev_jobs, (unsigned int)plan.jobs.size());
      if (rc) return rc;
    }
    else if (ctx->d.states == 4)
      k_tip_tables_dna<<<(unsigned int)plan.jobs.size(), 256, 0, ctx->stream>>>(dev_jobs, R);
    else
    {
      TipmapArg tm;
      memcpy(tm.map, ctx->tipmap, sizeof(tm.map));
      k_tip_tables_aa<<<(unsigned int)plan.jobs.size(), 256, 0, ctx->stream>>>(dev_jobs, R,
                                                                                ctx->maxstates, tm);
    }
    ++launched;
  }
  if (fused)
  {
    /* the whole list in one kernel (plg_traverse.cu) */
    int frc = (ctx->d.states == 4)
                  ? plg_launch_fused(ctx, (const FusedOp *)dev_payload, dev_records, (unsigned int)plan.fused.size(),
                                     plan.fused_nslot)
                  : plg_launch_walk_aa(ctx, (const FusedOp *)dev_payload, dev_records,
                                       dev_records + plan.fused_table_offset, (unsigned int)plan.fused.size());
    if (frc) return frc;
    cudaError_t ferr = cudaGetLastError();
    if (ferr != cudaSuccess)
    {
      plg_set_error("plg_update_partials: kernel launch failed: %s", cudaGetErrorString(ferr));
      return PLG_E_CUDA;
    }
    *kernels = launched + 2; /* pack + traverse */
    return PLG_OK;
  }
  const bool prof = ctx->profiling != 0;
  if (prof)
  {
    while (ctx->prof_events->size() < plan.groups.size() + 1)
    {
      cudaEvent_t ev;
      if (cudaEventCreate(&ev) != cudaSuccess)
      {
        plg_set_error("plg_update_partials: cudaEventCreate failed");
        return PLG_E_CUDA;
      }
      ctx->prof_events->push_back(ev);
    }
  }
  size_t gi = 0;
  for (const Group & g : plan.groups)
  {
    if (prof) cudaEventRecord((*ctx->prof_events)[gi], ctx->stream);
    ++gi;
    if (!fast)
      plg_gen_partials(ctx, g.kind, g.scale_mode, dev_ops + g.first, g.count);
    else switch (R)
    {
      case 1: launch_group<1>(ctx, g, dev_ops, nelem); break;
      case 2: launch_group<2>(ctx, g, dev_ops, nelem); break;
      case 4: launch_group<4>(ctx, g, dev_ops, nelem); break;
      case 8: launch_group<8>(ctx, g, dev_ops, nelem); break;
      case 16: launch_group<16>(ctx, g, dev_ops, nelem); break;
      default:
        plg_set_error("plg_update_partials: rate_cats=%u unsupported", R);
        return PLG_E_UNSUPPORTED;
    }
    ++launched;
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess)
  {
    plg_set_error("plg_update_partials: kernel launch failed: %s", cudaGetErrorString(err));
    return PLG_E_CUDA;
  }
  if (prof)
  {
    cudaEventRecord((*ctx->prof_events)[gi], ctx->stream);
    PLG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < plan.groups.size(); ++i)
    {
      float ms = 0.f;
      PLG_CUDA(cudaEventElapsedTime(&ms, (*ctx->prof_events)[i], (*ctx->prof_events)[i + 1]));
      const int kind = plan.groups[i].kind;
      ctx->stats.kind_ns[kind] += (unsigned long long)((double)ms * 1e6);
      ctx->stats.kind_bytes[kind] += plan.groups[i].bytes;
      ctx->stats.kind_launches[kind] += 1;
    }
  }
  *kernels = launched;
  return PLG_OK;
}

static uint64_t fnv1a(const void * data, size_t bytes)
{
  const unsigned char * p = (const unsigned char *)data;
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < bytes; ++i)
  {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

/* One CLV-update-shaped operation outside an operations list: out = (L . a) o (R . b) per
 * (site, rate), no scaling.  The sumtable of an inner-inner edge has exactly this shape
 * (L = pi-weighted inverse eigenvectors, R = eigenvectors; reference
 * src/core_derivatives_avx.c:131-185, src/core_derivatives_avx2.c:158-258), and that of a
 * tip-inner edge the tip-inner shape, so plg_update_sumtable reuses the streaming / tensor-core
 * kernels through this entry.  `op` holds device pointers. */
int plg_launch_single_op(plg_context * ctx, int kind, const DevOp & op)
{
  const unsigned int R = ctx->d.rate_cats;
  int rc = PLG_OK;
  switch (R)
  {
    case 1: rc = set_smem_limits<1>(); break;
    case 2: rc = set_smem_limits<2>(); break;
    case 4: rc = set_smem_limits<4>(); break;
    case 8: rc = set_smem_limits<8>(); break;
    case 16: rc = set_smem_limits<16>(); break;
    default: plg_set_error("rate_cats=%u unsupported", R); return PLG_E_UNSUPPORTED;
  }
  if (rc) return rc;
  const DevOp * dev = (const DevOp *)plg_stage(ctx, &op, sizeof(DevOp));
  if (!dev) return PLG_E_CUDA;
  Group g{kind, 0, 0, 1, 0};
  const unsigned int nelem = ctx->d.sites * R;
  switch (R)
  {
    case 1: launch_group<1>(ctx, g, dev, nelem); break;
    case 2: launch_group<2>(ctx, g, dev, nelem); break;
    case 4: launch_group<4>(ctx, g, dev, nelem); break;
    case 8: launch_group<8>(ctx, g, dev, nelem); break;
    default: launch_group<16>(ctx, g, dev, nelem); break;
  }
  PLG_LAUNCH_CHECK(ctx);
  ctx->stats.kernel_launches++;
  return PLG_OK;
}

extern "C" int plg_update_partials(plg_context_t * ctx, const pll_operation_t * operations,
                                   unsigned int count)
{
  PLG_CHECK_CTX(ctx);
  if (count == 0) return PLG_OK;
  plg_release_l2(ctx);
  if (!operations)
  {
    plg_set_error("plg_update_partials: NULL operations");
    return PLG_E_INVALID;
  }
  {
    int rc = PLG_OK;
    switch (ctx->d.rate_cats)
    {
      case 1: rc = set_smem_limits<1>(); break;
      case 2: rc = set_smem_limits<2>(); break;
      case 4: rc = set_smem_limits<4>(); break;
      case 8: rc = set_smem_limits<8>(); break;
      case 16: rc = set_smem_limits<16>(); break;
    }
    if (rc) return rc;
  }

  const size_t key_bytes = (size_t)count * sizeof(pll_operation_t);

  /* ---- replay a cached graph of this exact list ---- */
  uint64_t key = 0;
  const bool try_graph = ctx->use_graphs && !ctx->profiling && count >= 2;
  if (try_graph)
  {
    key = fnv1a(operations, key_bytes) ^ ((uint64_t)ctx->maxstates << 56) ^ (ctx->tipmap_epoch * 0x9E3779B97F4A7C15ull);
    auto hit = ctx->graphs->find(key);
    if (hit != ctx->graphs->end() && hit->second->key_bytes.size() == key_bytes &&
        memcmp(hit->second->key_bytes.data(), operations, key_bytes) == 0)
    {
      plg_graph_entry * ge = hit->second;
      ge->last_used = ++ctx->graph_clock;
      PLG_CUDA(cudaGraphLaunch(ge->exec, ctx->stream));
      ctx->stats.graph_launches++;
      ctx->stats.kernel_launches += ge->kernels;
      ctx->stats.partial_ops += count;
      ctx->stats.partial_levels += ge->levels;
      ctx->stats.algorithmic_bytes += ge->algorithmic_bytes;
      ctx->stats.compulsory_bytes += ge->compulsory_bytes;
      return PLG_OK;
    }
  }

  Plan plan;
  int rc = build_plan(ctx, operations, count, plan);
  if (rc) return rc;

  if (plan.table_doubles > ctx->tables_cap)
  {
    /* the scratch is about to move: cached graphs hold its old address */
    for (auto & kv : *ctx->graphs)
    {
      if (kv.second->exec) cudaGraphExecDestroy(kv.second->exec);
      cudaFree(kv.second->dev_tables);
      delete kv.second;
    }
    ctx->graphs->clear();
    rc = plg_ensure_tables(ctx, plan.table_doubles);
    if (rc) return rc;
  }
  /* turn table offsets into addresses */
  for (TableJob & j : plan.jobs) j.out = ctx->tables + (uintptr_t)j.out;
  for (size_t gi = 0; gi < plan.groups.size(); ++gi)
  {
    const Group & g = plan.groups[gi];
    for (unsigned int i = g.first; i < g.first + g.count; ++i)
    {
      DevOp & op = plan.ops[i];
      if (g.kind != PLG_KIND_II) op.lmat = ctx->tables + (uintptr_t)op.lmat;
      if (g.kind == PLG_KIND_TT) op.rmat = ctx->tables + (uintptr_t)op.rmat;
    }
  }

  const bool fused = !plan.fused.empty() && !ctx->profiling;
  for (FusedOp & f : plan.fused)
  {
    if (f.kind != PLG_KIND_II) f.op.lmat = ctx->tables + (uintptr_t)f.op.lmat;
    if (f.kind == PLG_KIND_TT) f.op.rmat = ctx->tables + (uintptr_t)f.op.rmat;
  }
  const void * ops_src = fused ? (const void *)plan.fused.data() : (const void *)plan.ops.data();
  const size_t ops_bytes = fused ? plan.fused.size() * sizeof(FusedOp) : plan.ops.size() * sizeof(DevOp);
  const size_t jobs_bytes = plan.jobs.size() * sizeof(TableJob);
  const size_t ops_bytes_al = (ops_bytes + 255) / 256 * 256;
  unsigned long long kernels = 0;

  /* a cached graph owns its descriptors (and, fused path, its packed records: 4.7 KB per
   * operation); very long lists are not worth pinning that much memory per distinct list */
  const bool graph_fits = plan.fused_scratch_bytes <= ((size_t)(ctx->d.states == 4 ? 64 : 512) << 20);
  /* capture on the SECOND sighting of a list: one-off lists (partial traversals during a tree
   * search) do not pay for cudaGraphInstantiate */
  bool capture = try_graph && graph_fits;
  if (capture)
  {
    if (ctx->seen_lists->size() > 4096) ctx->seen_lists->clear();
    capture = ++(*ctx->seen_lists)[key] >= 2;
  }
  if (capture && ctx->graphs->size() >= ctx->graph_cap && ctx->graphs->find(key) == ctx->graphs->end())
  {
    /* cache full: the least recently replayed list makes room (a tree search keeps issuing new
     * partial traversals; the lists of the current neighbourhood stay, old ones go) */
    auto victim = ctx->graphs->begin();
    for (auto it = ctx->graphs->begin(); it != ctx->graphs->end(); ++it)
      if (it->second->last_used < victim->second->last_used) victim = it;
    PLG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (victim->second->exec) cudaGraphExecDestroy(victim->second->exec);
    cudaFree(victim->second->dev_tables);
    ctx->seen_lists->erase(victim->first);
    delete victim->second;
    ctx->graphs->erase(victim);
    ctx->stats.graph_evictions++;
  }
  if (capture)
  {
    /* descriptors get a stable home, then the whole list is captured once */
    void * dev = NULL;
    const size_t rec_bytes = fused ? plan.fused_scratch_bytes : 0;
    const size_t jobs_bytes_al = (jobs_bytes + 255) / 256 * 256;
    PLG_CUDA(cudaMalloc(&dev, ops_bytes_al + jobs_bytes_al + rec_bytes + 256));
    PLG_CUDA(cudaMemcpyAsync(dev, ops_src, ops_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (jobs_bytes)
      PLG_CUDA(cudaMemcpyAsync((char *)dev + ops_bytes_al, plan.jobs.data(), jobs_bytes,
                               cudaMemcpyHostToDevice, ctx->stream));
    PLG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes += ops_bytes + jobs_bytes;

    cudaGraph_t graph = NULL;
    PLG_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    rc = enqueue_plan(ctx, plan, dev, (const TableJob *)((char *)dev + ops_bytes_al), &kernels, fused,
                      (unsigned char *)dev + ops_bytes_al + jobs_bytes_al);
    cudaError_t cerr = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc || cerr != cudaSuccess)
    {
      if (graph) cudaGraphDestroy(graph);
      cudaFree(dev);
      if (!rc)
      {
        plg_set_error("plg_update_partials: graph capture failed: %s", cudaGetErrorString(cerr));
        rc = PLG_E_CUDA;
      }
      return rc;
    }
    cudaGraphExec_t exec = NULL;
    cerr = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (cerr != cudaSuccess)
    {
      cudaFree(dev);
      plg_set_error("plg_update_partials: graph instantiate failed: %s", cudaGetErrorString(cerr));
      return PLG_E_CUDA;
    }
    plg_graph_entry * ge = new plg_graph_entry();
    ge->exec = exec;
    ge->key_bytes.assign((const unsigned char *)operations, (const unsigned char *)operations + key_bytes);
    ge->dev_tables = dev;
    ge->kernels = kernels;
    ge->levels = plan.levels;
    ge->last_used = ++ctx->graph_clock;
    ge->algorithmic_bytes = plan.algorithmic_bytes;
    ge->compulsory_bytes = fused ? plan.compulsory_bytes : plan.algorithmic_bytes;
    auto old = ctx->graphs->find(key);
    if (old != ctx->graphs->end())
    {
      /* hash collision with a different list: replace */
      PLG_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaGraphExecDestroy(old->second->exec);
      cudaFree(old->second->dev_tables);
      delete old->second;
    }
    (*ctx->graphs)[key] = ge;
    PLG_CUDA(cudaGraphLaunch(exec, ctx->stream));
    ctx->stats.graph_launches++;
  }
  else
  {
    const void * dev_ops = NULL;
    const TableJob * dev_jobs = NULL;
    if (ops_bytes + jobs_bytes + 1024 > ctx->stage_size / 2)
    {
      /* a list too long for the staging ring (the reference accepts any count) gets a
       * grow-only device buffer of its own; copies and kernels are ordered by the stream */
      const size_t need = ops_bytes_al + jobs_bytes + 256;
      if (need > ctx->list_buf_cap)
      {
        PLG_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->list_buf);
        ctx->list_buf = NULL;
        ctx->list_buf_cap = 0;
        PLG_CUDA(cudaMalloc(&ctx->list_buf, need + need / 2));
        ctx->list_buf_cap = need + need / 2;
      }
      PLG_CUDA(cudaMemcpyAsync(ctx->list_buf, ops_src, ops_bytes, cudaMemcpyHostToDevice, ctx->stream));
      dev_ops = ctx->list_buf;
      if (jobs_bytes)
      {
        PLG_CUDA(cudaMemcpyAsync(ctx->list_buf + ops_bytes_al, plan.jobs.data(), jobs_bytes,
                                 cudaMemcpyHostToDevice, ctx->stream));
        dev_jobs = (const TableJob *)(ctx->list_buf + ops_bytes_al);
      }
      ctx->stats.h2d_bytes += ops_bytes + jobs_bytes;
    }
    else
    {
      int src = plg_stage_reserve(ctx, ops_bytes + jobs_bytes + 1024);
      if (src) return src;
      dev_ops = plg_stage(ctx, ops_src, ops_bytes);
      if (!dev_ops) return PLG_E_CUDA;
      if (jobs_bytes)
      {
        dev_jobs = (const TableJob *)plg_stage(ctx, plan.jobs.data(), jobs_bytes);
        if (!dev_jobs) return PLG_E_CUDA;
      }
    }
    unsigned char * records = NULL;
    if (fused)
    {
      const size_t need = plan.fused_scratch_bytes;
      if (need > ctx->fused_records_cap)
      {
        PLG_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->fused_records);
        ctx->fused_records = NULL;
        ctx->fused_records_cap = 0;
        PLG_CUDA(cudaMalloc(&ctx->fused_records, need));
        ctx->fused_records_cap = need;
      }
      records = ctx->fused_records;
    }
    rc = enqueue_plan(ctx, plan, dev_ops, dev_jobs, &kernels, fused, records);
    if (rc) return rc;
  }
  ctx->stats.kernel_launches += kernels;
  ctx->stats.partial_ops += count;
  ctx->stats.partial_levels += plan.levels;
  ctx->stats.algorithmic_bytes += plan.algorithmic_bytes;
  ctx->stats.compulsory_bytes += fused ? plan.compulsory_bytes : plan.algorithmic_bytes;
  return PLG_OK;
}
