/*
 * plg_traverse.cu - the whole operations list of pll_update_partials in ONE kernel (DNA).
 *
 * Every kernel of plg_partials.cu streams the children of an operation from HBM and the
 * parent back: 264 KB of traffic per pattern for a 1 000-taxon traversal, which is what the
 * level-by-level path is bound by.  But patterns are independent, and a CLV element
 * (site, rate) of a parent depends only on the SAME element of its children.  So a warp can
 * take a small tile of elements and walk the entire list for it, keeping the tiles it has just
 * produced in shared memory until their parent consumes them:
 *
 *   - no CTA-wide synchronisation at all: each lane owns EPT elements of the tile and only
 *     ever reads back what it wrote itself (the only cross-lane step is the per-site rescaling
 *     vote, a warp ballot among the R lanes of a site);
 *   - the host orders the list depth-first, larger subtree first, and simulates a tiny cache of
 *     NSLOT tiles (CLV + scaler counts) per warp: every operation is told where its children
 *     live (slot or HBM) and where to keep its result.  For a random 1 000-taxon tree 4 slots
 *     catch 99 % of the inner-child reads (Strahler number ~5);
 *   - every parent is still written through to HBM (CLVs and scalers are outputs of the API),
 *     so what remains is ~132 KB of writes per pattern plus the tip characters: half the
 *     traffic of the level-by-level path.
 *
 * Arithmetic per element is that of the level-by-level kernels (same device functions, same
 * order), so CLVs and scaler counts are bit-identical to them and to the reference
 * (src/core_partials_avx.c:262-364, 366-529, 581-618, 899-1095); lists with slot recycling
 * (WAR / WAW hazards) run in their original order, which a sequential walk honours trivially.
 */
#include "plg_internal.cuh"
#include "plg_async.cuh"

#ifndef PLG_FUSED_EPT
#define PLG_FUSED_EPT 4
#endif
#ifndef PLG_FUSED_WARPS
#define PLG_FUSED_WARPS 11 /* + 1 producer warp = 12: registers are allocated per 4 warps */
#endif

__device__ __forceinline__ d4 fmatvec(const d4 (&M)[4], const d4 & c)
{
  d4 y;
  y.x = dot4_unfused(M[0].x, M[0].y, M[0].z, M[0].w, c);
  y.y = dot4_unfused(M[1].x, M[1].y, M[1].z, M[1].w, c);
  y.z = dot4_unfused(M[2].x, M[2].y, M[2].z, M[2].w, c);
  y.w = dot4_unfused(M[3].x, M[3].y, M[3].z, M[3].w, c);
  return y;
}

/* same, the 4x4 matrix of this rate read row by row from shared memory (operation ring) */
__device__ __forceinline__ d4 fmatvec_s(const double * M, const d4 & c)
{
  const double2 * m = reinterpret_cast<const double2 *>(M);
  const double2 r0a = m[0], r0b = m[1], r1a = m[2], r1b = m[3], r2a = m[4], r2b = m[5], r3a = m[6], r3b = m[7];
  struct { double x, y, z, w; } r0 = {r0a.x, r0a.y, r0b.x, r0b.y}, r1 = {r1a.x, r1a.y, r1b.x, r1b.y},
                                 r2 = {r2a.x, r2a.y, r2b.x, r2b.y}, r3 = {r3a.x, r3a.y, r3b.x, r3b.y};
  d4 y;
  y.x = dot4_unfused(r0.x, r0.y, r0.z, r0.w, c);
  y.y = dot4_unfused(r1.x, r1.y, r1.z, r1.w, c);
  y.z = dot4_unfused(r2.x, r2.y, r2.z, r2.w, c);
  y.w = dot4_unfused(r3.x, r3.y, r3.z, r3.w, c);
  return y;
}

__device__ __forceinline__ d4 fmul4(const d4 & a, const d4 & b)
{
  d4 r;
  r.x = __dmul_rn(a.x, b.x);
  r.y = __dmul_rn(a.y, b.y);
  r.z = __dmul_rn(a.z, b.z);
  r.w = __dmul_rn(a.w, b.w);
  return r;
}

/* shared-memory layout of one warp: [slot][item][half][lane] double2, then
 * [slot][item][lane] u32 - every lane reads and writes only its own column, 16-byte accesses
 * of a warp are contiguous (no bank conflicts) */
template <int EPT>
struct WarpCache
{
  double2 * clv;      /* NSLOT * EPT * 2 * 32 */
  unsigned int * sc;  /* NSLOT * EPT * 32     */
  __device__ __forceinline__ d4 load(int slot, int j, unsigned int lane) const
  {
    const double2 a = clv[((slot * EPT + j) * 2 + 0) * 32 + lane];
    const double2 b = clv[((slot * EPT + j) * 2 + 1) * 32 + lane];
    d4 r;
    r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
    return r;
  }
  __device__ __forceinline__ void store(int slot, int j, unsigned int lane, const d4 & v, unsigned int s)
  {
    clv[((slot * EPT + j) * 2 + 0) * 32 + lane] = make_double2(v.x, v.y);
    clv[((slot * EPT + j) * 2 + 1) * 32 + lane] = make_double2(v.z, v.w);
    sc[(slot * EPT + j) * 32 + lane] = s;
  }
  __device__ __forceinline__ unsigned int scaler(int slot, int j, unsigned int lane) const
  {
    return sc[(slot * EPT + j) * 32 + lane];
  }
};

#ifndef PLG_FUSED_MINB
#define PLG_FUSED_MINB 1
#endif
#ifndef PLG_FUSED_SLEEP
#define PLG_FUSED_SLEEP 64
#endif
#ifndef PLG_FUSED_STAGES
#define PLG_FUSED_STAGES 8
#endif

/* Operation data (descriptor + the two matrices / tip tables) does not depend on the tile, so a
 * producer warp streams it through a ring of shared-memory stages with TMA bulk copies while the
 * compute warps walk the list: a compute warp never waits for a global load of operation data.
 * The compute warps of a CTA consume the ring in lockstep order (not in lockstep time: the ring
 * gives PLG_FUSED_STAGES operations of slack); every warp makes the same number of passes. */
/* One packed operation record = one ring stage: descriptor, left block, right block.  A block
 * is either this operation's P-matrix set with rows of one rate 18 doubles apart, or a tip
 * table with the 16 codes (4R + 2) doubles apart: the 2-double pads rotate the shared-memory
 * banks so that the four rates of a site (matrix) and neighbouring codes (table) do not collide. */
/* ring depth: the records of 8 / 16 rate categories are larger (8.8 / 17 KB), the ring shorter */
__host__ __device__ constexpr int fused_stages(int R) { return R <= 4 ? PLG_FUSED_STAGES : 4; }

template <int R>
struct FusedStage
{
  static constexpr int MPITCH = 18;        /* doubles between the matrices of two rates */
  static constexpr int TPITCH = 4 * R + 2; /* doubles between the table rows of two codes */
  FusedOp desc;
  double L[16 * TPITCH];
  double Rr[16 * TPITCH];
};

__device__ __forceinline__ d4 lds_d4(const double * p)
{
  const double2 a = *reinterpret_cast<const double2 *>(p);
  const double2 b = *reinterpret_cast<const double2 *>(p + 2);
  d4 r;
  r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
  return r;
}

/* Builds the packed records of an operations list from the resident P-matrices: one block per
 * operation.  Tip tables are the masked row sums of reference src/core_partials_avx.c:944-984
 * ((a0+a1)+(a2+a3), absent states contribute +0.0), exactly as k_tip_tables_dna computes them. */
template <int R>
__global__ void k_fused_pack(const FusedOp * __restrict__ ops, unsigned char * __restrict__ records)
{
  using Stage = FusedStage<R>;
  const FusedOp f = ops[blockIdx.x];
  Stage * rec = reinterpret_cast<Stage *>(records + (size_t)blockIdx.x * sizeof(Stage));
  if (threadIdx.x < sizeof(FusedOp) / 8)
    reinterpret_cast<unsigned long long *>(&rec->desc)[threadIdx.x] =
        reinterpret_cast<const unsigned long long *>(ops + blockIdx.x)[threadIdx.x];
  for (int side = 0; side < 2; ++side)
  {
    const double * src = side ? f.rsrc : f.lsrc;
    double * dst = side ? rec->Rr : rec->L;
    const bool table = side ? (f.kind == PLG_KIND_TT) : (f.kind != PLG_KIND_II);
    if (table)
    {
      for (unsigned int t = threadIdx.x; t < 16u * R * 4u; t += blockDim.x)
      {
        const unsigned int i = t & 3u, k = (t >> 2) % R, code = t / (R * 4u);
        const double * row = src + (size_t)k * 16 + i * 4;
        const double a0 = (code & 1u) ? row[0] : 0.0;
        const double a1 = (code & 2u) ? row[1] : 0.0;
        const double a2 = (code & 4u) ? row[2] : 0.0;
        const double a3 = (code & 8u) ? row[3] : 0.0;
        dst[code * Stage::TPITCH + k * 4 + i] = hsum4(a0, a1, a2, a3);
      }
    }
    else
      for (unsigned int t = threadIdx.x; t < R * 16u; t += blockDim.x)
        dst[(t / 16u) * Stage::MPITCH + (t & 15u)] = src[t];
  }
}

/* rescaling vote + bookkeeping of one element; returns the scaler count to keep with the tile.
 * MODE = scale_mode of the operation (0 none, 1 per site, 2 per rate). */
template <int R, int MODE>
__device__ __forceinline__ unsigned int finish_element(d4 & p, bool valid, unsigned int child_scalers,
                                                       unsigned int gshift, unsigned int full_mask)
{
  if (MODE == 0) return 0;
  const bool below = valid && (p.x < PLG_SCALE_THRESHOLD) && (p.y < PLG_SCALE_THRESHOLD) &&
                     (p.z < PLG_SCALE_THRESHOLD) && (p.w < PLG_SCALE_THRESHOLD);
  bool scale = below;
  if (MODE == 1)
  {
    const unsigned int b = __ballot_sync(0xffffffffu, below);
    scale = (((b >> gshift) & full_mask) == full_mask);
  }
  if (scale)
  {
    p.x = __dmul_rn(p.x, PLG_SCALE_FACTOR);
    p.y = __dmul_rn(p.y, PLG_SCALE_FACTOR);
    p.z = __dmul_rn(p.z, PLG_SCALE_FACTOR);
    p.w = __dmul_rn(p.w, PLG_SCALE_FACTOR);
  }
  return child_scalers + (scale ? 1u : 0u);
}

__device__ __forceinline__ void load_matrix(const double * M, d4 (&out)[4])
{
#pragma unroll
  for (int r = 0; r < 4; ++r) out[r] = lds_d4(M + r * 4);
}

/* Child terms of the EPT elements of this lane: handed over in registers by the previous
 * operation (slot -2), from the warp's tile cache (slot >= 0), or - rare - from HBM (slot -1).
 * The source is warp-uniform, so it is decided once per operation, outside the element loop;
 * `use(j, x)` consumes element j's child vector. */
template <int R, int EPT, int MODE, bool FULL, typename Use>
__device__ __forceinline__ void for_each_child(const WarpCache<EPT> & cache, int slot, unsigned int lane,
                                               const double * clv, const unsigned int * scaler, unsigned int e0,
                                               unsigned int nelem, const d4 (&prev)[EPT],
                                               const unsigned int (&prev_sc)[EPT], unsigned int (&sc)[EPT], Use use)
{
  const bool count = MODE != 0 && scaler != nullptr;
  if (slot == -2)
  {
#pragma unroll
    for (int j = 0; j < EPT; ++j)
    {
      if (count) sc[j] += prev_sc[j];
      use(j, prev[j]);
    }
  }
  else if (slot >= 0)
  {
#pragma unroll
    for (int j = 0; j < EPT; ++j)
    {
      if (count) sc[j] += cache.scaler(slot, j, lane);
      use(j, cache.load(slot, j, lane));
    }
  }
  else
  {
    /* coherent loads: the tile may have been stored earlier in THIS launch (by this lane; a
     * per-site scaler by the rate-0 lane of the site, ordered by the __syncwarp that ends every
     * operation) - .nc / __ldg are only defined for data the kernel never writes */
#pragma unroll
    for (int j = 0; j < EPT; ++j)
    {
      const unsigned int e = e0 + j * 32;
      const bool valid = FULL || e < nelem;
      if (count && valid) sc[j] += ld_coherent_u32(scaler + (MODE == 2 ? e : e / R));
      use(j, valid ? ld_stream_coherent(clv + (size_t)e * 4) : d4{0.0, 0.0, 0.0, 0.0});
    }
  }
}

/* One operation on the EPT elements of this lane.  KIND, MODE (scaling), FULL (no element of the
 * tile is past the end) and RIGHT_FIRST are compile-time.  `p` / `psc` hold the result tile of
 * the previous operation on entry and this operation's on exit; a child handed over that way is
 * transformed IN PLACE, so its side is evaluated first (the product of the two sides commutes
 * bit for bit).  One matrix is resident in registers at a time. */
template <int R, int EPT, int KIND, int MODE, bool FULL, bool RIGHT_FIRST>
__device__ __forceinline__ void run_op(const FusedStage<R> & st, WarpCache<EPT> & cache, unsigned int lane,
                                       unsigned int k, unsigned int e0, unsigned int nelem,
                                       const unsigned char * lcode, const unsigned char * rcode,
                                       unsigned int gshift, unsigned int full_mask, d4 (&p)[EPT],
                                       unsigned int (&psc)[EPT])
{
  const int lslot = st.desc.lslot, rslot = st.desc.rslot, pslot = st.desc.pslot;
  const unsigned int * lscale = st.desc.op.lscale;
  const unsigned int * rscale = st.desc.op.rscale;
  double * parent = st.desc.op.parent + (size_t)e0 * 4;
  unsigned int * pscale = st.desc.op.pscale;
  /* 0: a dead store - the buffer is overwritten later in this list and nobody reads this value
   * from HBM (slot-recycling lists; see build_plan) */
  const bool write_back = (st.desc.pad & 1) != 0;
  unsigned int sc[EPT];

  if (KIND == PLG_KIND_TT)
  {
#pragma unroll
    for (int j = 0; j < EPT; ++j)
    {
      sc[j] = 0;
      p[j] = fmul4(lds_d4(st.L + lcode[j * (32 / R)] * FusedStage<R>::TPITCH + k * 4),
                   lds_d4(st.Rr + rcode[j * (32 / R)] * FusedStage<R>::TPITCH + k * 4));
    }
  }
  else if (RIGHT_FIRST)
  {
    /* right child = previous result: p <- R.p, then times the left term */
    {
      d4 Rm[4];
      load_matrix(st.Rr + k * FusedStage<R>::MPITCH, Rm);
#pragma unroll
      for (int j = 0; j < EPT; ++j)
      {
        sc[j] = (MODE != 0 && rscale) ? psc[j] : 0u;
        p[j] = fmatvec(Rm, p[j]);
      }
    }
    if (KIND == PLG_KIND_II)
    {
      d4 Lm[4];
      load_matrix(st.L + k * FusedStage<R>::MPITCH, Lm);
      /* lslot is never -2 here (the right child was the one handed over) */
      for_each_child<R, EPT, MODE, FULL>(cache, lslot, lane, st.desc.op.left, lscale, e0, nelem, p, psc, sc,
                                         [&](int j, const d4 & x) { p[j] = fmul4(fmatvec(Lm, x), p[j]); });
    }
    else
    {
#pragma unroll
      for (int j = 0; j < EPT; ++j) p[j] = fmul4(lds_d4(st.L + lcode[j * (32 / R)] * FusedStage<R>::TPITCH + k * 4), p[j]);
    }
  }
  else
  {
    /* left term first (in place if the left child is the previous result), then the right */
    if (KIND == PLG_KIND_II)
    {
      d4 Lm[4];
      load_matrix(st.L + k * FusedStage<R>::MPITCH, Lm);
#pragma unroll
      for (int j = 0; j < EPT; ++j) sc[j] = 0;
      for_each_child<R, EPT, MODE, FULL>(cache, lslot, lane, st.desc.op.left, lscale, e0, nelem, p, psc, sc,
                                         [&](int j, const d4 & x) { p[j] = fmatvec(Lm, x); });
    }
    else
    {
#pragma unroll
      for (int j = 0; j < EPT; ++j)
      {
        sc[j] = 0;
        p[j] = lds_d4(st.L + lcode[j * (32 / R)] * FusedStage<R>::TPITCH + k * 4);
      }
    }
    d4 Rm[4];
    load_matrix(st.Rr + k * FusedStage<R>::MPITCH, Rm);
    /* rslot is never -2 here */
    for_each_child<R, EPT, MODE, FULL>(cache, rslot, lane, st.desc.op.right, rscale, e0, nelem, p, psc, sc,
                                       [&](int j, const d4 & y) { p[j] = fmul4(p[j], fmatvec(Rm, y)); });
  }
  /* rescaling votes (tip-tip never rescales and zeroes the scaler, reference
   * src/core_partials_avx.c:113-116), then the write-back */
#pragma unroll
  for (int j = 0; j < EPT; ++j)
  {
    const unsigned int e = e0 + j * 32;
    const bool valid = FULL || e < nelem;
    const unsigned int sv =
        (KIND == PLG_KIND_TT) ? 0u : finish_element<R, MODE>(p[j], valid, sc[j], gshift, full_mask);
    if (pslot >= 0) cache.store(pslot, j, lane, p[j], sv);
    psc[j] = sv;
    if (valid && write_back)
    {
      st_stream(parent + j * 128, p[j]);
      if (MODE == 2) pscale[e] = sv;
      else if (MODE == 1 && k == 0) pscale[e / R] = sv;
    }
  }
}

template <int R, int EPT, bool FULL>
__device__ __forceinline__ void dispatch_op(const FusedStage<R> & st, WarpCache<EPT> & cache, unsigned int lane,
                                            unsigned int k, unsigned int e0, unsigned int nelem,
                                            const unsigned char * lcode, const unsigned char * rcode,
                                            unsigned int gshift, unsigned int full_mask, d4 (&prev)[EPT],
                                            unsigned int (&prev_sc)[EPT])
{
  const int kind = st.desc.kind, mode = st.desc.scale_mode;
#define PLG_RUN(K_, M_)                                                                                   \
  do {                                                                                                    \
    if (K_ != PLG_KIND_TT && st.desc.rslot == -2)                                                         \
      run_op<R, EPT, K_, M_, FULL, true>(st, cache, lane, k, e0, nelem, lcode, rcode, gshift, full_mask,  \
                                         prev, prev_sc);                                                  \
    else                                                                                                  \
      run_op<R, EPT, K_, M_, FULL, false>(st, cache, lane, k, e0, nelem, lcode, rcode, gshift, full_mask, \
                                          prev, prev_sc);                                                 \
  } while (0)
  if (kind == PLG_KIND_TT)
  {
    if (mode == 0) PLG_RUN(PLG_KIND_TT, 0);
    else if (mode == 1) PLG_RUN(PLG_KIND_TT, 1);
    else PLG_RUN(PLG_KIND_TT, 2);
  }
  else if (kind == PLG_KIND_TI)
  {
    if (mode == 1) PLG_RUN(PLG_KIND_TI, 1);
    else if (mode == 0) PLG_RUN(PLG_KIND_TI, 0);
    else PLG_RUN(PLG_KIND_TI, 2);
  }
  else
  {
    if (mode == 1) PLG_RUN(PLG_KIND_II, 1);
    else if (mode == 0) PLG_RUN(PLG_KIND_II, 0);
    else PLG_RUN(PLG_KIND_II, 2);
  }
#undef PLG_RUN
}

template <int R, int EPT>
__global__ void __launch_bounds__((PLG_FUSED_WARPS + 1) * 32, PLG_FUSED_MINB)
k_traverse_dna(const unsigned char * __restrict__ records, unsigned int n_ops, unsigned int nelem,
               unsigned int nslot)
{
  using namespace plg_async;
  constexpr int S = fused_stages(R);
  constexpr int NW = PLG_FUSED_WARPS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FusedStage<R> * stages = reinterpret_cast<FusedStage<R> *>(smem_raw);
  uint64_t * full = reinterpret_cast<uint64_t *>(smem_raw + S * sizeof(FusedStage<R>));
  uint64_t * empty = full + S;
  unsigned char * cache_base = smem_raw + S * sizeof(FusedStage<R>) + 128;

  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
  {
    for (int s = 0; s < S; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NW);
    }
    fence_barrier_init();
  }
  __syncthreads();

  constexpr unsigned int TILE = 32u * EPT;
  const unsigned int ntiles = (nelem + TILE - 1) / TILE;
  /* Every CTA owns a contiguous run of tiles, the runs differing by at most one tile, and walks
   * it NW tiles per pass.  All CTAs then make the same number of passes and the remainder is a
   * LIGHT last pass everywhere (a few active warps per CTA, which finish sooner) instead of a full
   * extra pass on a few SMs while the others idle. */
  const unsigned int base_tiles = ntiles / gridDim.x, extra_tiles = ntiles % gridDim.x;
  const unsigned int my_tiles = base_tiles + (blockIdx.x < extra_tiles ? 1u : 0u);
  const unsigned int first_tile = blockIdx.x * base_tiles + (blockIdx.x < extra_tiles ? blockIdx.x : extra_tiles);
  const unsigned int passes = (my_tiles + NW - 1) / NW;

  if (warp == NW)
  {
    /* ---- producer: one lane feeds the ring ---- */
    if (lane == 0)
    {
      /* the two block sizes of an operation are requested PF operations ahead of their use, so
       * that their L2 latency never sits between two bulk copies */
      constexpr int PF = 4;
      unsigned int q_lb[PF], q_rb[PF];
      const unsigned int total = passes * n_ops;
      auto rec_of = [&](unsigned int i) { return reinterpret_cast<const FusedStage<R> *>(records) + i; };
#pragma unroll
      for (int q = 0; q < PF; ++q)
      {
        const FusedStage<R> * src = rec_of((unsigned int)q % n_ops);
        q_lb[q] = src->desc.lbytes;
        q_rb[q] = src->desc.rbytes;
      }
      for (unsigned int base = 0; base < total; base += PF)
      {
#pragma unroll
        for (int q = 0; q < PF; ++q)
        {
          const unsigned int it = base + q;
          if (it >= total) break;
          const int s = it % S;
          const FusedStage<R> * src = rec_of(it % n_ops);
          const unsigned int lb = q_lb[q], rb = q_rb[q];
          {
            const FusedStage<R> * nxt = rec_of((it + PF) % n_ops);
            q_lb[q] = nxt->desc.lbytes;
            q_rb[q] = nxt->desc.rbytes;
          }
          /* the producer is usually a ring ahead: wait politely, its spinning would take issue
           * slots from the compute warps of its scheduler */
          if (it >= (unsigned int)S)
            while (!mbar_try_wait(&empty[s], ((it / S) - 1) & 1u)) __nanosleep(PLG_FUSED_SLEEP);
          mbar_arrive_expect_tx(&full[s], (unsigned int)sizeof(FusedOp) + lb + rb);
          /* descriptor and left block are adjacent in the record */
          bulk_g2s(&stages[s].desc, &src->desc, (unsigned int)sizeof(FusedOp) + lb, &full[s]);
          bulk_g2s(stages[s].Rr, src->Rr, rb, &full[s]);
        }
      }
    }
    return;
  }

  /* ---- compute warps ---- */
  const unsigned int k = lane & (R - 1);
  const size_t per_warp = (size_t)nslot * EPT * 32 * (32 + 4);
  WarpCache<EPT> cache;
  cache.clv = reinterpret_cast<double2 *>(cache_base + warp * per_warp);
  cache.sc = reinterpret_cast<unsigned int *>(cache_base + warp * per_warp + (size_t)nslot * EPT * 32 * 32);
  const unsigned int full_mask = (R >= 32) ? 0xffffffffu : ((1u << R) - 1u);
  const unsigned int gshift = lane & ~(unsigned int)(R - 1);

  /* Tip characters are the one per-tile input that still comes from HBM / L2.  A tile covers
   * TILE / R consecutive sites, i.e. TILE / R consecutive bytes of a tip row (rows are padded to
   * 256 bytes, a tile never straddles the padding): while operation i is computed, a few lanes
   * copy the bytes operation i+1 needs into a double-buffered strip of shared memory with
   * cp.async - no register holds them in the meantime. */
  constexpr unsigned int TIP_BYTES = TILE / R;           /* per tip row and tile */
  constexpr unsigned int TIP_LANES = TIP_BYTES / 4;      /* 4-byte cp.async each */
  unsigned char * codes = cache_base + (size_t)NW * per_warp + (size_t)warp * (4 * TIP_BYTES);
  auto prefetch_codes = [&](const FusedStage<R> & st, unsigned int tile_n, bool haven, unsigned int buf)
  {
    const int kind = st.desc.kind;
    if (kind != PLG_KIND_II && haven)
    {
      const unsigned int chunks = ((kind == PLG_KIND_TT) ? 2u : 1u) * TIP_LANES;
      for (unsigned int c = lane; c < chunks; c += 32)
      {
        const unsigned int side = c / TIP_LANES, q = c % TIP_LANES;
        const unsigned char * src = (side ? st.desc.op.rtip : st.desc.op.ltip) + (size_t)tile_n * TIP_BYTES + q * 4;
        unsigned char * dst = codes + (buf * 2 + side) * TIP_BYTES + q * 4;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  d4 prev[EPT];                /* result tile of the previous operation (register forwarding) */
  unsigned int prev_sc[EPT];
#pragma unroll
  for (int j = 0; j < EPT; ++j)
  {
    prev[j] = d4{0.0, 0.0, 0.0, 0.0};
    prev_sc[j] = 0;
  }
  unsigned int it = 0;
  const unsigned int total_its = passes * n_ops;
  if (total_its)
  {
    mbar_wait(&full[0], 0);
    prefetch_codes(stages[0], first_tile + warp, warp < my_tiles, 0);
  }
  for (unsigned int pass = 0; pass < passes; ++pass)
  {
    const unsigned int tile = first_tile + pass * NW + warp;
    const bool have = pass * NW + warp < my_tiles;
    const unsigned int e0 = tile * TILE + lane;
    const bool tile_full = (tile + 1) * TILE <= nelem;
    const unsigned int tile_next = tile + NW;
    const bool have_next = (pass + 1) * NW + warp < my_tiles;
    for (unsigned int i = 0; i < n_ops; ++i, ++it)
    {
      /* stage `it` is known to be full: it was waited for when its tip codes were requested */
      const int s = it % S;
      const unsigned int buf = it & 1u;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      const unsigned char * lcode = codes + (buf * 2 + 0) * TIP_BYTES + lane / R;
      const unsigned char * rcode = codes + (buf * 2 + 1) * TIP_BYTES + lane / R;
      if (it + 1 < total_its)
      {
        const unsigned int itn = it + 1;
        mbar_wait(&full[itn % S], (itn / S) & 1u);
        if (i + 1 == n_ops)
          prefetch_codes(stages[itn % S], tile_next, have_next, buf ^ 1u);
        else
          prefetch_codes(stages[itn % S], tile, have, buf ^ 1u);
      }
      if (have)
      {
        const FusedStage<R> & st = stages[s];
        if (tile_full)
          dispatch_op<R, EPT, true>(st, cache, lane, k, e0, nelem, lcode, rcode, gshift, full_mask, prev, prev_sc);
        else
          dispatch_op<R, EPT, false>(st, cache, lane, k, e0, nelem, lcode, rcode, gshift, full_mask, prev, prev_sc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

/* ------------------------------------------------------------------------------------ */
template <int R>
static int launch_fused(plg_context * ctx, const FusedOp * dev_ops, unsigned char * dev_records, unsigned int n_ops,
                        unsigned int nslot)
{
  constexpr int EPT = PLG_FUSED_EPT;
  static_assert(sizeof(FusedOp) == 128, "descriptor must be 128 bytes");
  const size_t smem = fused_stages(R) * sizeof(FusedStage<R>) + 128 +
                      (size_t)PLG_FUSED_WARPS * nslot * EPT * 32 * (32 + 4) +
                      (size_t)PLG_FUSED_WARPS * 4 * (32 * EPT / R);
  static size_t configured[PLG_MAX_DEVICES] = {}; /* function attributes are per device */
  if (smem > configured[ctx->device % PLG_MAX_DEVICES])
  {
    PLG_CUDA(cudaFuncSetAttribute(k_traverse_dna<R, EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[ctx->device % PLG_MAX_DEVICES] = smem;
  }
  int per_sm = 0;
  PLG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_traverse_dna<R, EPT>,
                                                        (PLG_FUSED_WARPS + 1) * 32, smem));
  if (per_sm < 1) per_sm = 1;
  const unsigned int nelem = ctx->d.sites * R;
  const unsigned int ntiles = (nelem + 32 * EPT - 1) / (32 * EPT);
  unsigned int blocks = (unsigned int)(ctx->sm_count * per_sm);
  const unsigned int want = (ntiles + PLG_FUSED_WARPS - 1) / PLG_FUSED_WARPS;
  if (want < blocks) blocks = want;
  k_fused_pack<R><<<n_ops, 128, 0, ctx->stream>>>(dev_ops, dev_records);
  k_traverse_dna<R, EPT><<<blocks, (PLG_FUSED_WARPS + 1) * 32, smem, ctx->stream>>>(dev_records, n_ops, nelem, nslot);
  return PLG_OK;
}

int plg_launch_fused(plg_context * ctx, const FusedOp * dev_ops, unsigned char * dev_records, unsigned int n_ops,
                     unsigned int nslot)
{
  if (plg_fused_record_bytes(ctx->d.rate_cats) != sizeof(FusedStage<4>) && ctx->d.rate_cats == 4)
  {
    plg_set_error("fused traversal: record size mismatch");
    return PLG_E_INVALID;
  }
  switch (ctx->d.rate_cats)
  {
    case 1: return launch_fused<1>(ctx, dev_ops, dev_records, n_ops, nslot);
    case 2: return launch_fused<2>(ctx, dev_ops, dev_records, n_ops, nslot);
    case 4: return launch_fused<4>(ctx, dev_ops, dev_records, n_ops, nslot);
    case 8: return launch_fused<8>(ctx, dev_ops, dev_records, n_ops, nslot);
    case 16: return launch_fused<16>(ctx, dev_ops, dev_records, n_ops, nslot);
    default: plg_set_error("fused traversal: rate_cats=%u unsupported", ctx->d.rate_cats); return PLG_E_UNSUPPORTED;
  }
}
