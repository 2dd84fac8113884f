/*
 * plg_traverse_aa.cu - the whole operations list of pll_update_partials in ONE kernel, 20 states.
 *
 * The level-by-level tensor-core kernel (plg_partials.cu: k_partial_dmma_aa) streams both
 * children of every operation from HBM and the parent back: 1 932 B per pattern for an
 * inner-inner update of which only the 644 B parent store is compulsory.  As in the DNA
 * traversal (plg_traverse.cu) patterns are independent, so a tile of patterns can walk the
 * entire list with the children it has just produced kept on chip.  What differs from DNA is
 * that the update is a (sites x rates) x 20 x 20 product on the FP64 tensor cores
 * (reference src/core_partials_avx2.c:568-803, src/core_partials_avx.c:1097-1340) and that a
 * tile is five times larger per pattern:
 *
 *   - work split by RATE: warp w of a CTA owns rate k = w % R of the 32 patterns of team
 *     w / R.  A warp therefore needs only the 20 x 20 matrices of ITS rate, which it pulls into
 *     registers as DMMA B fragments once per operation (15 doubles per lane and child) and
 *     reuses for its four 8-pattern groups: no shared-memory operand read per DMMA at all;
 *   - the result of an operation stays in registers as DMMA D fragments.  With the child-state
 *     order of plg_dmma.cuh a lane's D values ARE its A values for the next product (one
 *     64-bit shuffle moves the two odd ones), so a parent consumes the child finished right
 *     before it - every inner child of a tip-inner operation and one child of every
 *     inner-inner one in depth-first order - without touching shared memory;
 *   - the other child waits in a per-warp tile cache in shared memory in A-fragment layout,
 *     lane-private columns (a lane only reads back what it wrote: no synchronisation); a miss
 *     re-reads the tile from HBM with coherent loads;
 *   - the only cross-warp step is the per-site rescaling vote (all R x 20 entries of a site
 *     below 2^-256, reference src/core_partials_avx2.c:788-801): the R warps of a team exchange
 *     one bit mask per 8-pattern group through shared memory around a named barrier;
 *   - operation data (descriptor + the two matrix sets as B fragments, or tip tables) is one
 *     fixed-size packed record per operation, built on the device by k_fused_pack_aa and
 *     streamed through a 3-stage shared-memory ring by TMA bulk copies.  There is no producer
 *     warp: the warp that draws the last ticket of a stage issues the copy of the record three
 *     operations ahead;
 *   - every parent CLV / scaler observable after the call is written through to HBM (same
 *     dead-store rule as the DNA traversal, build_plan in plg_partials.cu).
 *
 * The common case - complete tile, per-site scalers, right child in registers, left child of an
 * inner-inner operation in the tile cache - runs a straight-line path (fast_op_aa); everything
 * else (partial tiles, per-rate scalers, no scalers, cache misses) the general run_op_aa.
 *
 * Arithmetic: the same DMMA chains, in the same order, as k_partial_dmma_aa - CLVs and scaler
 * counts are bit-identical to the level-by-level path (tests/test_fused_traversal_aa_gpu.py).
 */
#include "plg_internal.cuh"
#include "plg_async.cuh"
#include "plg_dmma.cuh"

#define AAF_W PLG_AAF_WARPS
#define AAF_NSG 4   /* 8-pattern groups per warp tile: 32 patterns */
#define AAF_S 3     /* ring stages */
#define AAF_BLOCK_DOUBLES 480 /* per rate: 15 fragments x 32 lanes = 24 codes x 20 states */

/* doubles between the table rows of two tip codes: padded for four categories so that the rows
 * of different codes start in different banks (a row of 4 x 20 doubles is a multiple of 32
 * banks: unpadded, the eight patterns of a group would hit the same 16 banks) */
__host__ __device__ constexpr int aa_table_pitch(int R) { return R == 4 ? R * 20 + 2 : R * 20; }
/* tip codes a block has room for */
__host__ __device__ constexpr int aa_table_codes(int R) { return R * AAF_BLOCK_DOUBLES / aa_table_pitch(R); }

template <int R>
struct AaStage
{
  FusedOp desc;
  double L[R * AAF_BLOCK_DOUBLES];  /* B fragments (k_fused_pack_aa) or tip table [code][pitch] */
  double Rr[R * AAF_BLOCK_DOUBLES];
};

/* ------------------------------------------------------------------------------------ */
/* packed records                                                                        */
/* ------------------------------------------------------------------------------------ */
/* One block per operation.  Tip tables: sequential sum, in increasing state order, of
 * P_rate[i][m] over the states m in tipmap[code] - exactly k_tip_tables_aa (reference
 * src/core_partials_avx.c:1140-1177, :177-220). */
template <int R>
__global__ void k_fused_pack_aa(const FusedOp * __restrict__ ops, unsigned char * __restrict__ records,
                                unsigned int maxstates, const TipmapArg tm)
{
  using Stage = AaStage<R>;
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
      constexpr unsigned int TP = aa_table_pitch(R);
      for (unsigned int t = threadIdx.x; t < (unsigned int)R * AAF_BLOCK_DOUBLES; t += blockDim.x)
      {
        const unsigned int code = t / TP, within = t % TP;
        const unsigned int i = within % 20u, k = within / 20u;
        double s = 0.0;
        if (code < maxstates && k < (unsigned int)R)
        {
          const unsigned int state = tm.map[code];
          const double * row = src + (size_t)k * 400 + i * 20;
          for (unsigned int m = 0; m < 20u; ++m)
            if ((state >> m) & 1u) s = __dadd_rn(s, row[m]);
        }
        dst[t] = s;
      }
    }
    else
      /* B fragments: [rate][pair j < 7][lane][2] holds fragments 2j and 2j + 1 (one 16-byte load
       * per pair), then [lane] fragment 14 */
      for (unsigned int t = threadIdx.x; t < (unsigned int)R * AAF_BLOCK_DOUBLES; t += blockDim.x)
      {
        const unsigned int k = t / AAF_BLOCK_DOUBLES, within = t % AAF_BLOCK_DOUBLES;
        const unsigned int f = within < 448u ? (within >> 6) * 2u + (within & 1u) : 14u;
        const unsigned int l = within < 448u ? (within >> 1) & 31u : within - 448u;
        dst[t] = dmma_bfrag_value(src + (size_t)k * 400, f, l);
      }
  }
}

/* ------------------------------------------------------------------------------------ */
/* device helpers                                                                        */
/* ------------------------------------------------------------------------------------ */
__device__ __forceinline__ void st_stream2_aa(double * p, double x, double y)
{
#ifdef AAF_EXP_NOSTG
  if (x == 1.2345e-300) /* timing experiment (tools/exp_variants.sh): stores never happen */
#endif
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}

/* wrapping ticket counter: returns the old value, stores old >= wrap ? 0 : old + 1 (relaxed: the
 * ordering comes from the mbarrier arrival / wait around it) */
__device__ __forceinline__ unsigned int atom_inc_smem(unsigned int * p, unsigned int wrap)
{
  unsigned int old;
  asm volatile("atom.relaxed.cta.shared::cta.inc.u32 %0, [%1], %2;"
               : "=r"(old)
               : "r"(plg_async::smem_addr(p)), "r"(wrap)
               : "memory");
  return old;
}

__device__ __forceinline__ void named_barrier(unsigned int id, unsigned int threads)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

/* per-warp tile cache: A-fragment layout, every lane owns a column.  The pointers are already
 * offset by the lane. */
struct AaCache
{
  double2 * v2;      /* [slot][group][2][32] */
  double * v1;       /* [slot][group][32]    */
  unsigned int * sc; /* [slot][group][32]    */
  __device__ __forceinline__ void load(int slot, int sg, double (&a)[5]) const
  {
    const double2 x = v2[((slot * AAF_NSG + sg) * 2 + 0) * 32];
    const double2 y = v2[((slot * AAF_NSG + sg) * 2 + 1) * 32];
    a[0] = x.x; a[1] = x.y; a[2] = y.x; a[3] = y.y;
    a[4] = v1[(slot * AAF_NSG + sg) * 32];
  }
  __device__ __forceinline__ void store(int slot, int sg, const double (&a)[5], unsigned int s) const
  {
#ifdef AAF_EXP_NOCACHE
    if (a[0] != 1.2345e-300) return;
#endif
    v2[((slot * AAF_NSG + sg) * 2 + 0) * 32] = make_double2(a[0], a[1]);
    v2[((slot * AAF_NSG + sg) * 2 + 1) * 32] = make_double2(a[2], a[3]);
    v1[(slot * AAF_NSG + sg) * 32] = a[4];
    sc[(slot * AAF_NSG + sg) * 32] = s;
  }
  __device__ __forceinline__ unsigned int scaler(int slot, int sg) const
  {
    return sc[(slot * AAF_NSG + sg) * 32];
  }
};
#define AAF_SLOT_BYTES (AAF_NSG * (2 * 32 * 16 + 32 * 8 + 32 * 4)) /* 5632 */

/* D fragments of a finished tile -> the A fragments of the next product */
__device__ __forceinline__ void afrag_from_tile(const double (&t)[3][2], unsigned int lane, unsigned int q,
                                                double (&a)[5])
{
  a[0] = t[0][0]; a[1] = t[0][1]; a[2] = t[1][0]; a[3] = t[1][1];
  const double other = __shfl_sync(0xffffffffu, t[2][1], (lane & ~3u) | (q & 1u));
  a[4] = (q < 2u) ? t[2][0] : other;
}

__device__ __forceinline__ void load_bfrag(const double * __restrict__ blk, unsigned int k, unsigned int lane,
                                           double (&B)[PLG_DMMA_FRAGS])
{
  const double * mine = blk + (size_t)k * AAF_BLOCK_DOUBLES;
  const double2 * b2 = reinterpret_cast<const double2 *>(mine) + lane;
#pragma unroll
  for (int j = 0; j < 7; ++j)
  {
    const double2 v = b2[j * 32];
    B[2 * j] = v.x;
    B[2 * j + 1] = v.y;
  }
  B[14] = mine[448 + lane];
}

/* two 8-pattern groups at once: six independent accumulator chains, k-step outer - the chain of
 * one N tile is the chain k_partial_dmma_aa runs (same bits) */
__device__ __forceinline__ void mma_pair(const double (&B)[PLG_DMMA_FRAGS], const double (&a0)[5],
                                         const double (&a1)[5], double (&d0)[3][2], double (&d1)[3][2])
{
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) d0[nt][0] = d0[nt][1] = d1[nt][0] = d1[nt][1] = 0.0;
#pragma unroll
  for (int ks = 0; ks < 5; ++ks)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
    {
#ifdef AAF_EXP_NOMMA
      d0[nt][0] += a0[ks] * B[nt * 5 + ks];
      d1[nt][0] += a1[ks] * B[nt * 5 + ks];
#else
      dmma884(d0[nt][0], d0[nt][1], a0[ks], B[nt * 5 + ks]);
      dmma884(d1[nt][0], d1[nt][1], a1[ks], B[nt * 5 + ks]);
#endif
    }
}

__device__ __forceinline__ void mma_one(const double (&B)[PLG_DMMA_FRAGS], const double (&a)[5], double (&d)[3][2])
{
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) d[nt][0] = d[nt][1] = 0.0;
#pragma unroll
  for (int ks = 0; ks < 5; ++ks)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#ifdef AAF_EXP_NOMMA
      d[nt][0] += a[ks] * B[nt * 5 + ks];
#else
      dmma884(d[nt][0], d[nt][1], a[ks], B[nt * 5 + ks]);
#endif
}

/* the same chains with schedulable DMMAs (fast path: straight-line, all lanes active) */
__device__ __forceinline__ void mma_pair_free(const double (&B)[PLG_DMMA_FRAGS], const double (&a0)[5],
                                              const double (&a1)[5], double (&d0)[3][2], double (&d1)[3][2])
{
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) d0[nt][0] = d0[nt][1] = d1[nt][0] = d1[nt][1] = 0.0;
#pragma unroll
  for (int ks = 0; ks < 5; ++ks)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
    {
      dmma884_free(d0[nt][0], d0[nt][1], a0[ks], B[nt * 5 + ks]);
      dmma884_free(d1[nt][0], d1[nt][1], a1[ks], B[nt * 5 + ks]);
    }
}

__device__ __forceinline__ void mma_one_free(const double (&B)[PLG_DMMA_FRAGS], const double (&a)[5],
                                             double (&d)[3][2])
{
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) d[nt][0] = d[nt][1] = 0.0;
#pragma unroll
  for (int ks = 0; ks < 5; ++ks)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) dmma884_free(d[nt][0], d[nt][1], a[ks], B[nt * 5 + ks]);
}

/* streaming 16-byte store the compiler may schedule (st.global.cs) */
__device__ __forceinline__ void st_cs2(double * p, double x, double y)
{
#ifdef AAF_EXP_NOSTG
  if (x == 1.2345e-300)
#endif
  __stcs(reinterpret_cast<double2 *>(p), make_double2(x, y));
}

/* A fragments (and scaler count) of a child that is not handed over in registers: from the
 * warp's tile cache (slot >= 0) or - a miss - from HBM */
template <int R>
__device__ __forceinline__ void child_afrag(const AaCache & cache, int slot, int sg, unsigned int k,
                                            unsigned int q, unsigned int hi_state, const double * clv,
                                            const unsigned int * scaler, int mode, unsigned int site, bool ok,
                                            double (&a)[5], unsigned int & sc)
{
  if (slot >= 0)
  {
    cache.load(slot, sg, a);
    if (mode != 0 && scaler) sc += cache.scaler(slot, sg);
    return;
  }
  if (ok)
  {
    /* coherent loads: the tile may have been stored earlier in this launch - by this very lane
     * (a per-site scaler that is read back is stored by every rate warp, FusedOp::pad bit 1, so
     * no warp depends on another warp's store) */
    const double * r = clv + (size_t)site * (R * 20) + k * 20;
    const double2 x = ld_stream_coherent2(r + 2 * q);
    const double2 y = ld_stream_coherent2(r + 8 + 2 * q);
    a[0] = x.x; a[1] = x.y; a[2] = y.x; a[3] = y.y;
    a[4] = ld_stream_coherent1(r + hi_state);
    if (mode != 0 && scaler) sc += ld_coherent_u32(scaler + (mode == 2 ? (size_t)site * R + k : (size_t)site));
  }
  else
    a[0] = a[1] = a[2] = a[3] = a[4] = 0.0;
}

/* D-layout values of a tip-table row: table[code][k][8 nt + 2q .. +1] */
template <int R>
__device__ __forceinline__ double2 table_pair(const double * __restrict__ tab, unsigned int code, unsigned int k,
                                              int nt, unsigned int q)
{
  if (nt == 2 && q >= 2u) return make_double2(0.0, 0.0);
#ifdef AAF_EXP_NOTABLE
  return make_double2(0.25 + code, 0.5);
#endif
  return *reinterpret_cast<const double2 *>(tab + (size_t)code * aa_table_pitch(R) + k * 20 + 8 * nt + 2 * q);
}

__device__ __forceinline__ unsigned int tip_code(const unsigned char * codes, int sg, unsigned int g, unsigned int limit)
{
  return min((unsigned int)codes[sg * 8 + g], limit);
}

/* per-warp state of the walk */
struct AaWarp
{
  unsigned int lane, g, q, k, team, hi_state;
  unsigned int first_site; /* of this warp's tile */
  unsigned int sites;
  bool full;               /* no pattern of the tile is past the end */
  uint4 * votes;           /* [2][W] */
  unsigned int vbuf;
  size_t clv_off;          /* doubles from a CLV's base to this lane's first value of the tile */
};

/* per-site rescaling decision of the team: vote[sg] bit 4g = all 20 states of (pattern g of group
 * sg, this rate) below the threshold; ANDed over the R rate warps of the team */
template <int R>
__device__ __forceinline__ void team_vote(AaWarp & w, unsigned int (&vote)[AAF_NSG])
{
  if (R == 1) return;
#ifdef AAF_EXP_NOVOTE
  return;
#endif
  uint4 * mine = w.votes + (size_t)w.vbuf * AAF_W + w.team * R;
  if (w.lane == 0) mine[w.k] = make_uint4(vote[0], vote[1], vote[2], vote[3]);
  named_barrier(1 + w.team, R * 32);
#pragma unroll
  for (int kk = 0; kk < R; ++kk)
  {
    const uint4 o = mine[kk];
    vote[0] &= o.x; vote[1] &= o.y; vote[2] &= o.z; vote[3] &= o.w;
  }
  w.vbuf ^= 1u;
}

/* the two halves of team_vote: publish this warp's bits; later, meet the team and combine */
template <int R>
__device__ __forceinline__ void team_vote_post(AaWarp & w, const unsigned int (&vote)[AAF_NSG])
{
  if (R == 1) return;
#ifdef AAF_EXP_NOVOTE
  return;
#endif
  uint4 * mine = w.votes + (size_t)w.vbuf * AAF_W + w.team * R;
  if (w.lane == 0) mine[w.k] = make_uint4(vote[0], vote[1], vote[2], vote[3]);
}
template <int R>
__device__ __forceinline__ void team_vote_collect(AaWarp & w, unsigned int (&vote)[AAF_NSG])
{
  if (R == 1) return;
#ifdef AAF_EXP_NOVOTE
  return;
#endif
  const uint4 * mine = w.votes + (size_t)w.vbuf * AAF_W + w.team * R;
  named_barrier(1 + w.team, R * 32);
#pragma unroll
  for (int kk = 0; kk < R; ++kk)
  {
    const uint4 o = mine[kk];
    vote[0] &= o.x; vote[1] &= o.y; vote[2] &= o.z; vote[3] &= o.w;
  }
  w.vbuf ^= 1u;
}

__device__ __forceinline__ unsigned int below_votes(const double (&t)[3][2], unsigned int q)
{
  bool below = (t[0][0] < PLG_SCALE_THRESHOLD) && (t[0][1] < PLG_SCALE_THRESHOLD) &&
               (t[1][0] < PLG_SCALE_THRESHOLD) && (t[1][1] < PLG_SCALE_THRESHOLD);
  if (q < 2u) below = below && (t[2][0] < PLG_SCALE_THRESHOLD) && (t[2][1] < PLG_SCALE_THRESHOLD);
  unsigned int b = __ballot_sync(0xffffffffu, below);
  b &= b >> 1;
  b &= b >> 2;
  return b & 0x11111111u; /* bit 4g: all 20 states of (pattern g, this rate) below */
}

/* ------------------------------------------------------------------------------------ */
/* the general path                                                                      */
/* ------------------------------------------------------------------------------------ */
/* One operation on the 32 patterns x 1 rate of this warp.  KIND and FWD (0: no child handed
 * over in registers, 1: the left one, 2: the right / inner one) are compile-time; scaling mode
 * and tile fullness are warp-uniform run-time values.  `p` / `psc` hold the result tile of the
 * previous operation on entry and this operation's on exit.  `release` hands the ring stage
 * back as soon as its last byte has been read. */
template <int R, int KIND, int FWD, typename Release>
__device__ __forceinline__ void run_op_aa(const AaStage<R> & st, const AaCache & cache, AaWarp & w,
                                          const unsigned char * lcode, const unsigned char * rcode,
                                          double (&p)[AAF_NSG][3][2], unsigned int (&psc)[AAF_NSG],
                                          Release release)
{
  double * const parent = st.desc.op.parent;
  unsigned int * const pscale = st.desc.op.pscale;
  const double * const left = st.desc.op.left;
  const double * const right = st.desc.op.right;
  const unsigned int * const lscale = st.desc.op.lscale;
  const unsigned int * const rscale = st.desc.op.rscale;
  const int lslot = st.desc.lslot, rslot = st.desc.rslot, pslot = st.desc.pslot;
  const int mode = st.desc.scale_mode;
  const int pad = st.desc.pad;
  constexpr unsigned int CODE_MAX = aa_table_codes(R) - 1;

  const unsigned int lane = w.lane, g = w.g, q = w.q, k = w.k;
  unsigned int sc[AAF_NSG];
  bool ok[AAF_NSG];
#pragma unroll
  for (int sg = 0; sg < AAF_NSG; ++sg)
  {
    sc[sg] = 0;
    ok[sg] = w.full || (w.first_site + sg * 8 + g < w.sites);
  }
  double B[PLG_DMMA_FRAGS];

  if (KIND == PLG_KIND_TT)
  {
#pragma unroll
    for (int sg = 0; sg < AAF_NSG; ++sg)
    {
      const unsigned int lc = tip_code(lcode, sg, g, CODE_MAX), rc = tip_code(rcode, sg, g, CODE_MAX);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        const double2 x = table_pair<R>(st.L, lc, k, nt, q);
        const double2 y = table_pair<R>(st.Rr, rc, k, nt, q);
        p[sg][nt][0] = __dmul_rn(x.x, y.x);
        p[sg][nt][1] = __dmul_rn(x.y, y.y);
      }
    }
    release();
  }
  else if (FWD == 2)
  {
    /* the right (for tip-inner: the inner) child is the previous result: p <- R . p in place */
    load_bfrag(st.Rr, k, lane, B);
#pragma unroll
    for (int sg = 0; sg < AAF_NSG; ++sg)
    {
      double a[5], d[3][2];
      afrag_from_tile(p[sg], lane, q, a);
      mma_one(B, a, d);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) { p[sg][nt][0] = d[nt][0]; p[sg][nt][1] = d[nt][1]; }
      if (mode != 0 && rscale) sc[sg] = psc[sg];
    }
    if (KIND == PLG_KIND_II)
    {
      load_bfrag(st.L, k, lane, B);
      release();
#pragma unroll
      for (int sg = 0; sg < AAF_NSG; ++sg)
      {
        double a[5], d[3][2];
        child_afrag<R>(cache, lslot, sg, k, q, w.hi_state, left, lscale, mode, w.first_site + sg * 8 + g, ok[sg],
                       a, sc[sg]);
        mma_one(B, a, d);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
        {
          p[sg][nt][0] = __dmul_rn(d[nt][0], p[sg][nt][0]);
          p[sg][nt][1] = __dmul_rn(d[nt][1], p[sg][nt][1]);
        }
      }
    }
    else
    {
#pragma unroll
      for (int sg = 0; sg < AAF_NSG; ++sg)
      {
        const unsigned int lc = tip_code(lcode, sg, g, CODE_MAX);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
        {
          const double2 x = table_pair<R>(st.L, lc, k, nt, q);
          p[sg][nt][0] = __dmul_rn(x.x, p[sg][nt][0]);
          p[sg][nt][1] = __dmul_rn(x.y, p[sg][nt][1]);
        }
      }
      release();
    }
  }
  else
  {
    /* left term first (in place if the left child is the previous result), then the right */
    if (KIND == PLG_KIND_II)
    {
      load_bfrag(st.L, k, lane, B);
#pragma unroll
      for (int sg = 0; sg < AAF_NSG; ++sg)
      {
        double a[5], d[3][2];
        if (FWD == 1)
        {
          afrag_from_tile(p[sg], lane, q, a);
          if (mode != 0 && lscale) sc[sg] = psc[sg];
        }
        else
          child_afrag<R>(cache, lslot, sg, k, q, w.hi_state, left, lscale, mode, w.first_site + sg * 8 + g, ok[sg],
                         a, sc[sg]);
        mma_one(B, a, d);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) { p[sg][nt][0] = d[nt][0]; p[sg][nt][1] = d[nt][1]; }
      }
    }
    else
    {
#pragma unroll
      for (int sg = 0; sg < AAF_NSG; ++sg)
      {
        const unsigned int lc = tip_code(lcode, sg, g, CODE_MAX);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
        {
          const double2 x = table_pair<R>(st.L, lc, k, nt, q);
          p[sg][nt][0] = x.x;
          p[sg][nt][1] = x.y;
        }
      }
    }
    load_bfrag(st.Rr, k, lane, B);
    release();
#pragma unroll
    for (int sg = 0; sg < AAF_NSG; ++sg)
    {
      double a[5], d[3][2];
      child_afrag<R>(cache, rslot, sg, k, q, w.hi_state, right, rscale, mode, w.first_site + sg * 8 + g, ok[sg], a,
                     sc[sg]);
      mma_one(B, a, d);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        p[sg][nt][0] = __dmul_rn(p[sg][nt][0], d[nt][0]);
        p[sg][nt][1] = __dmul_rn(p[sg][nt][1], d[nt][1]);
      }
    }
  }

  /* ---- rescaling vote (tip-tip never rescales and zeroes the scaler) ---- */
  unsigned int vote[AAF_NSG];
#pragma unroll
  for (int sg = 0; sg < AAF_NSG; ++sg) vote[sg] = 0;
  if (KIND != PLG_KIND_TT && mode != 0)
  {
#pragma unroll
    for (int sg = 0; sg < AAF_NSG; ++sg) vote[sg] = below_votes(p[sg], q);
    if (mode == 1) team_vote<R>(w, vote); /* per-site scaling: all R rates of the pattern must agree */
  }

  /* ---- scale, write through, keep ---- */
#pragma unroll
  for (int sg = 0; sg < AAF_NSG; ++sg)
  {
    const bool scale = (vote[sg] >> (4u * g)) & 1u;
    if (scale)
    {
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        p[sg][nt][0] = __dmul_rn(p[sg][nt][0], PLG_SCALE_FACTOR);
        p[sg][nt][1] = __dmul_rn(p[sg][nt][1], PLG_SCALE_FACTOR);
      }
    }
    const unsigned int sv = (KIND == PLG_KIND_TT || mode == 0) ? 0u : sc[sg] + (scale ? 1u : 0u);
    psc[sg] = sv;
    const unsigned int site = w.first_site + sg * 8 + g;
    if ((pad & 1) && ok[sg])
    {
      double * out = parent + w.clv_off + sg * (8 * R * 20);
      st_stream2_aa(out, p[sg][0][0], p[sg][0][1]);
      st_stream2_aa(out + 8, p[sg][1][0], p[sg][1][1]);
      if (q < 2u) st_stream2_aa(out + 16, p[sg][2][0], p[sg][2][1]);
      if (q == 0u)
      {
        if (mode == 2) pscale[(size_t)site * R + k] = sv;
        else if (mode == 1 && (k == 0u || (pad & 2))) pscale[site] = sv;
      }
    }
    if (pslot >= 0)
    {
      double a[5];
      afrag_from_tile(p[sg], lane, q, a);
      cache.store(pslot, sg, a, sv);
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* the common case, straight-line: the tile is complete, per-site scalers (or a tip-tip    */
/* operation), the inner / right child arrives in registers and the left child of an       */
/* inner-inner operation waits in the tile cache.                                          */
/* ------------------------------------------------------------------------------------ */
template <int R, int KIND, typename Release>
__device__ __forceinline__ void fast_op_aa(const AaStage<R> & st, const AaCache & cache, AaWarp & w,
                                           const unsigned char * lcode, const unsigned char * rcode,
                                           double (&p)[AAF_NSG][3][2], unsigned int (&psc)[AAF_NSG],
                                           Release release)
{
  double * const parent = st.desc.op.parent;
  unsigned int * const pscale = st.desc.op.pscale;
  const bool has_l = st.desc.op.lscale != nullptr, has_r = st.desc.op.rscale != nullptr;
  const int lslot = st.desc.lslot, pslot = st.desc.pslot;
  const int pad = st.desc.pad;
  const unsigned int lane = w.lane, g = w.g, q = w.q, k = w.k;
  constexpr unsigned int CODE_MAX = aa_table_codes(R) - 1;
  unsigned int sc[AAF_NSG];
  double B[PLG_DMMA_FRAGS];

  if (KIND == PLG_KIND_TT)
  {
#pragma unroll
    for (int sg = 0; sg < AAF_NSG; ++sg)
    {
      const unsigned int lc = tip_code(lcode, sg, g, CODE_MAX), rc = tip_code(rcode, sg, g, CODE_MAX);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        const double2 x = table_pair<R>(st.L, lc, k, nt, q);
        const double2 y = table_pair<R>(st.Rr, rc, k, nt, q);
        p[sg][nt][0] = __dmul_rn(x.x, y.x);
        p[sg][nt][1] = __dmul_rn(x.y, y.y);
      }
      sc[sg] = 0;
    }
    release();
  }
  else
  {
    /* p <- R . p in place (the right / inner child is the previous result) */
    load_bfrag(st.Rr, k, lane, B);
#pragma unroll
    for (int sg = 0; sg < AAF_NSG; sg += 2)
    {
      double a0[5], a1[5], d0[3][2], d1[3][2];
      afrag_from_tile(p[sg], lane, q, a0);
      afrag_from_tile(p[sg + 1], lane, q, a1);
      mma_pair_free(B, a0, a1, d0, d1);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        p[sg][nt][0] = d0[nt][0]; p[sg][nt][1] = d0[nt][1];
        p[sg + 1][nt][0] = d1[nt][0]; p[sg + 1][nt][1] = d1[nt][1];
      }
    }
#pragma unroll
    for (int sg = 0; sg < AAF_NSG; ++sg) sc[sg] = has_r ? psc[sg] : 0u;
    if (KIND == PLG_KIND_II)
    {
      load_bfrag(st.L, k, lane, B);
      release();
      /* one 8-pattern group at a time (three accumulator chains): the right-hand term of all
       * four groups stays in registers meanwhile, a pair would not fit */
#pragma unroll
      for (int sg = 0; sg < AAF_NSG; ++sg)
      {
        double a[5], d[3][2];
        cache.load(lslot, sg, a);
        if (has_l) sc[sg] += cache.scaler(lslot, sg);
        mma_one_free(B, a, d);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
        {
          p[sg][nt][0] = __dmul_rn(d[nt][0], p[sg][nt][0]);
          p[sg][nt][1] = __dmul_rn(d[nt][1], p[sg][nt][1]);
        }
      }
    }
    else
    {
#pragma unroll
      for (int sg = 0; sg < AAF_NSG; ++sg)
      {
        const unsigned int lc = tip_code(lcode, sg, g, CODE_MAX);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
        {
          const double2 x = table_pair<R>(st.L, lc, k, nt, q);
          p[sg][nt][0] = __dmul_rn(x.x, p[sg][nt][0]);
          p[sg][nt][1] = __dmul_rn(x.y, p[sg][nt][1]);
        }
      }
      release();
    }
  }

  /* ---- epilogue: the stores do not wait for the rescaling vote.  Rescaling is rare, so the
   * tile is written through (and kept) as computed while the team's votes are collected, and
   * the few tiles in which some pattern does rescale are multiplied and stored once more. ---- */
  unsigned int vote[AAF_NSG];
#pragma unroll
  for (int sg = 0; sg < AAF_NSG; ++sg) vote[sg] = 0;
  if (KIND != PLG_KIND_TT)
  {
#pragma unroll
    for (int sg = 0; sg < AAF_NSG; ++sg) vote[sg] = below_votes(p[sg], q);
    team_vote_post<R>(w, vote);
  }
  double * const out = parent + w.clv_off;
  unsigned int * const so = pscale + w.first_site + g;
  const bool store_sc = pscale != nullptr && (k == 0u || (pad & 2)) && q == 0u;
  auto write_out = [&]()
  {
    if (pad & 1)
    {
#pragma unroll
      for (int sg = 0; sg < AAF_NSG; ++sg)
      {
        st_cs2(out + sg * (8 * R * 20), p[sg][0][0], p[sg][0][1]);
        st_cs2(out + sg * (8 * R * 20) + 8, p[sg][1][0], p[sg][1][1]);
        if (q < 2u) st_cs2(out + sg * (8 * R * 20) + 16, p[sg][2][0], p[sg][2][1]);
      }
      if (store_sc)
      {
#pragma unroll
        for (int sg = 0; sg < AAF_NSG; ++sg) so[sg * 8] = sc[sg];
      }
    }
    if (pslot >= 0)
    {
#pragma unroll
      for (int sg = 0; sg < AAF_NSG; ++sg)
      {
        double a[5];
        afrag_from_tile(p[sg], lane, q, a);
        cache.store(pslot, sg, a, sc[sg]);
      }
    }
  };
  write_out();
  if (KIND != PLG_KIND_TT)
  {
    team_vote_collect<R>(w, vote);
    if ((vote[0] | vote[1] | vote[2] | vote[3]) != 0u)
    {
      /* rare (warp-uniform): some pattern of the tile is rescaled */
#pragma unroll
      for (int sg = 0; sg < AAF_NSG; ++sg)
      {
        const bool scale = (vote[sg] >> (4u * g)) & 1u;
        const double f = scale ? PLG_SCALE_FACTOR : 1.0;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
        {
          p[sg][nt][0] = __dmul_rn(p[sg][nt][0], f);
          p[sg][nt][1] = __dmul_rn(p[sg][nt][1], f);
        }
        sc[sg] += scale ? 1u : 0u;
      }
      write_out();
    }
  }
#pragma unroll
  for (int sg = 0; sg < AAF_NSG; ++sg) psc[sg] = sc[sg];
}

/* ------------------------------------------------------------------------------------ */
template <int R>
__global__ void __launch_bounds__(AAF_W * 32, 1)
k_traverse_aa(const unsigned char * __restrict__ records, unsigned int n_ops, unsigned int sites, unsigned int nslot)
{
  using namespace plg_async;
  using Stage = AaStage<R>;
  constexpr int TEAMS = AAF_W / R;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Stage * stages = reinterpret_cast<Stage *>(smem_raw);
  unsigned char * after = smem_raw + AAF_S * sizeof(Stage);
  uint64_t * full = reinterpret_cast<uint64_t *>(after);                /* AAF_S */
  uint64_t * empty = full + AAF_S;                                      /* AAF_S */
  unsigned int * ticket = reinterpret_cast<unsigned int *>(after + 48); /* AAF_S */
  uint4 * votes = reinterpret_cast<uint4 *>(after + 64);                /* [2][AAF_W] */
  unsigned char * codes_base = after + 64 + 2 * AAF_W * 16;             /* [AAF_W][2 bufs][2 sides][32] */
  unsigned char * cache_base = codes_base + AAF_W * 128;

  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int warp = threadIdx.x >> 5;
  const unsigned int ntiles = (sites + 31u) / 32u;
  const unsigned int tiles_per_pass = gridDim.x * TEAMS;
  const unsigned int passes = (ntiles + tiles_per_pass - 1) / tiles_per_pass;
  const unsigned int total_its = passes * n_ops;
  const Stage * recs = reinterpret_cast<const Stage *>(records);

  if (threadIdx.x == 0)
  {
    for (int s = 0; s < AAF_S; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], AAF_W);
      ticket[s] = 0;
    }
    fence_barrier_init();
    for (unsigned int it0 = 0; it0 < (unsigned int)AAF_S && it0 < total_its; ++it0)
    {
      mbar_arrive_expect_tx(&full[it0], (unsigned int)sizeof(Stage));
      bulk_g2s(&stages[it0], recs + (it0 % n_ops), (unsigned int)sizeof(Stage), &full[it0]);
    }
  }
  __syncthreads();

  AaWarp w;
  w.lane = lane;
  w.g = lane >> 2;
  w.q = lane & 3u;
  w.k = warp % R;
  w.team = warp / R;
  w.hi_state = dmma_child_state(4, w.q);
  w.sites = sites;
  w.votes = votes;
  w.vbuf = 0;
  w.first_site = 0;
  w.full = false;
  w.clv_off = 0;

  AaCache cache;
  {
    unsigned char * mine = cache_base + (size_t)warp * nslot * AAF_SLOT_BYTES;
    cache.v2 = reinterpret_cast<double2 *>(mine) + lane;
    cache.v1 = reinterpret_cast<double *>(mine + (size_t)nslot * AAF_NSG * 2 * 32 * 16) + lane;
    cache.sc = reinterpret_cast<unsigned int *>(mine + (size_t)nslot * AAF_NSG * (2 * 32 * 16 + 32 * 8)) + lane;
  }

  /* tip characters of the tile: 32 bytes per tip row, fetched one operation ahead by cp.async
   * into a double-buffered strip of this warp (16-byte chunks: lanes 0-1 left row, 2-3 right) */
  unsigned char * codes = codes_base + warp * 128;
  auto prefetch_codes = [&](const Stage & st, unsigned int tile_n, bool haven, unsigned int buf)
  {
    const int kind = st.desc.kind;
    if (kind != PLG_KIND_II && haven && lane < ((kind == PLG_KIND_TT) ? 4u : 2u))
    {
      const unsigned int side = lane >> 1, half = lane & 1u;
      const unsigned char * src = (side ? st.desc.op.rtip : st.desc.op.ltip) + (size_t)tile_n * 32 + half * 16;
      unsigned char * dst = codes + (buf * 2 + side) * 32 + half * 16;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  double p[AAF_NSG][3][2];
  unsigned int psc[AAF_NSG];
#pragma unroll
  for (int sg = 0; sg < AAF_NSG; ++sg)
  {
    psc[sg] = 0;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) p[sg][nt][0] = p[sg][nt][1] = 0.0;
  }

  unsigned int it = 0;
  {
    mbar_wait(&full[0], 0);
    const unsigned int tile0 = blockIdx.x * TEAMS + w.team;
    prefetch_codes(stages[0], tile0, tile0 < ntiles, 0);
  }
  for (unsigned int pass = 0; pass < passes; ++pass)
  {
    const unsigned int tile = (pass * gridDim.x + blockIdx.x) * TEAMS + w.team;
    const bool have = tile < ntiles;
    const unsigned int tile_next = ((pass + 1) * gridDim.x + blockIdx.x) * TEAMS + w.team;
    w.first_site = tile * 32u;
    w.full = (tile + 1) * 32u <= sites;
    w.clv_off = (size_t)(w.first_site + w.g) * (R * 20) + w.k * 20 + 2 * w.q;
    for (unsigned int i = 0; i < n_ops; ++i, ++it)
    {
      /* stage `it` is known to be full: it was waited for when its tip codes were requested */
      const unsigned int s = it % AAF_S;
      const unsigned int buf = it & 1u;
      /* record index of operation it + AAF_S (the one a release may have to fetch) */
      unsigned int rec_ahead = i + AAF_S;
      while (rec_ahead >= n_ops) rec_ahead -= n_ops;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      const unsigned char * lcode = codes + (buf * 2 + 0) * 32;
      const unsigned char * rcode = codes + (buf * 2 + 1) * 32;
      bool next_requested = (it + 1 >= total_its);
      auto request_next = [&](bool blocking)
      {
        if (next_requested) return;
        const unsigned int itn = it + 1;
        uint64_t * bar = &full[itn % AAF_S];
        const unsigned int parity = (itn / AAF_S) & 1u;
        if (blocking) mbar_wait(bar, parity);
        else if (!mbar_try_wait(bar, parity)) return;
        if (i + 1 == n_ops) prefetch_codes(stages[itn % AAF_S], tile_next, tile_next < ntiles, buf ^ 1u);
        else prefetch_codes(stages[itn % AAF_S], tile, have, buf ^ 1u);
        next_requested = true;
      };
      request_next(false);

      bool released = false;
      auto release = [&]()
      {
        /* the warp is done with stage s (release: its reads precede the arrival).  Every warp
         * arrives BEFORE it draws its ticket, so the warp that draws the last one knows that all
         * have arrived: it acquires the completed phase and refills the stage with the record
         * AAF_S operations ahead. */
        released = true;
        __syncwarp();
        if (lane == 0)
        {
          mbar_arrive(&empty[s]);
          if (atom_inc_smem(&ticket[s], AAF_W - 1) == AAF_W - 1 && it + AAF_S < total_its)
          {
            mbar_wait(&empty[s], (it / AAF_S) & 1u);
            mbar_arrive_expect_tx(&full[s], (unsigned int)sizeof(Stage));
            bulk_g2s(&stages[s], recs + rec_ahead, (unsigned int)sizeof(Stage), &full[s]);
          }
        }
      };

      if (have)
      {
        const Stage & st = stages[s];
        const int kind = st.desc.kind;
        const int fwd = (kind == PLG_KIND_TT) ? 0 : (st.desc.rslot == -2 ? 2 : (st.desc.lslot == -2 ? 1 : 0));
        const bool fast = w.full && ((kind == PLG_KIND_TT && st.desc.scale_mode != 2) ||
                                     (st.desc.scale_mode == 1 && fwd == 2 &&
                                      (kind == PLG_KIND_TI || st.desc.lslot >= 0)));
        if (fast)
        {
          if (kind == PLG_KIND_TT) fast_op_aa<R, PLG_KIND_TT>(st, cache, w, lcode, rcode, p, psc, release);
          else if (kind == PLG_KIND_TI) fast_op_aa<R, PLG_KIND_TI>(st, cache, w, lcode, rcode, p, psc, release);
          else fast_op_aa<R, PLG_KIND_II>(st, cache, w, lcode, rcode, p, psc, release);
        }
        else if (kind == PLG_KIND_TT) run_op_aa<R, PLG_KIND_TT, 0>(st, cache, w, lcode, rcode, p, psc, release);
        else if (kind == PLG_KIND_TI)
        {
          if (fwd == 2) run_op_aa<R, PLG_KIND_TI, 2>(st, cache, w, lcode, rcode, p, psc, release);
          else run_op_aa<R, PLG_KIND_TI, 0>(st, cache, w, lcode, rcode, p, psc, release);
        }
        else
        {
          if (fwd == 2) run_op_aa<R, PLG_KIND_II, 2>(st, cache, w, lcode, rcode, p, psc, release);
          else if (fwd == 1) run_op_aa<R, PLG_KIND_II, 1>(st, cache, w, lcode, rcode, p, psc, release);
          else run_op_aa<R, PLG_KIND_II, 0>(st, cache, w, lcode, rcode, p, psc, release);
        }
      }
      if (!released) release();
      request_next(true);
      __syncwarp();
    }
  }
}

/* ------------------------------------------------------------------------------------ */
template <int R>
static size_t aa_fused_smem(unsigned int nslot)
{
  return AAF_S * sizeof(AaStage<R>) + 64 + 2 * AAF_W * 16 + AAF_W * 128 + (size_t)AAF_W * nslot * AAF_SLOT_BYTES;
}

unsigned int plg_fused_aa_slots(unsigned int rate_cats, unsigned int wanted)
{
  size_t stage = 0;
  switch (rate_cats)
  {
    case 1: stage = sizeof(AaStage<1>); break;
    case 2: stage = sizeof(AaStage<2>); break;
    case 4: stage = sizeof(AaStage<4>); break;
    default: return 0;
  }
  const size_t fixed = AAF_S * stage + 64 + 2 * AAF_W * 16 + AAF_W * 128;
  const size_t budget = 227 * 1024;
  if (fixed + (size_t)AAF_W * AAF_SLOT_BYTES > budget) return 0;
  unsigned int fit = (unsigned int)((budget - fixed) / ((size_t)AAF_W * AAF_SLOT_BYTES));
  return wanted < fit ? wanted : fit;
}

unsigned int plg_fused_aa_max_codes(unsigned int rate_cats)
{
  return rate_cats == 4 ? aa_table_codes(4) : rate_cats == 2 ? aa_table_codes(2) : aa_table_codes(1);
}

size_t plg_fused_aa_record_bytes(unsigned int rate_cats)
{
  return sizeof(FusedOp) + 2 * (size_t)rate_cats * AAF_BLOCK_DOUBLES * sizeof(double);
}

template <int R>
static int launch_fused_aa(plg_context * ctx, const FusedOp * dev_ops, unsigned char * dev_records, unsigned int n_ops,
                           unsigned int nslot)
{
  static_assert(sizeof(AaStage<R>) == sizeof(FusedOp) + 2 * R * AAF_BLOCK_DOUBLES * sizeof(double), "record layout");
  static_assert(AAF_W % R == 0, "a team is R warps");
  const size_t smem = aa_fused_smem<R>(nslot);
  static size_t configured[PLG_MAX_DEVICES] = {};
  if (smem > configured[ctx->device % PLG_MAX_DEVICES])
  {
    PLG_CUDA(cudaFuncSetAttribute(k_traverse_aa<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[ctx->device % PLG_MAX_DEVICES] = smem;
  }
  const unsigned int ntiles = (ctx->d.sites + 31u) / 32u;
  constexpr unsigned int TEAMS = AAF_W / R;
  unsigned int blocks = (unsigned int)ctx->sm_count;
  const unsigned int want = (ntiles + TEAMS - 1) / TEAMS;
  if (want < blocks) blocks = want;
  TipmapArg tm;
  memcpy(tm.map, ctx->tipmap, sizeof(tm.map));
  k_fused_pack_aa<R><<<n_ops, 256, 0, ctx->stream>>>(dev_ops, dev_records, ctx->maxstates, tm);
  k_traverse_aa<R><<<blocks, AAF_W * 32, smem, ctx->stream>>>(dev_records, n_ops, ctx->d.sites, nslot);
  return PLG_OK;
}

int plg_launch_fused_aa(plg_context * ctx, const FusedOp * dev_ops, unsigned char * dev_records, unsigned int n_ops,
                        unsigned int nslot)
{
  switch (ctx->d.rate_cats)
  {
    case 1: return launch_fused_aa<1>(ctx, dev_ops, dev_records, n_ops, nslot);
    case 2: return launch_fused_aa<2>(ctx, dev_ops, dev_records, n_ops, nslot);
    case 4: return launch_fused_aa<4>(ctx, dev_ops, dev_records, n_ops, nslot);
    default: plg_set_error("fused 20-state traversal: rate_cats=%u unsupported", ctx->d.rate_cats); return PLG_E_UNSUPPORTED;
  }
}
