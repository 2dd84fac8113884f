/*
 * plg_walk_aa.cu - the whole operations list of pll_update_partials in ONE kernel, 20 states,
 * second design (PLL_GPU_FUSED_AA=2; plg_traverse_aa.cu is the first).
 *
 * What the first kernel taught: the tensor pipe is saturated by ONE warp
 * per scheduler, what costs time is everything around the DMMAs - twelve scattered 128-bit
 * global stores per group, tip-table gathers with bank conflicts, a cross-warp vote, and a
 * load/store unit that is as busy as the tensor pipe.  This kernel moves all global traffic to
 * the TMA unit and keeps the instruction stream of a math warp close to "fragment loads, DMMA,
 * multiply, shared-memory store":
 *
 *   - a TILE is 16 patterns x all rates (two 8-pattern DMMA groups).  One tile team per
 *     scheduler (4 per SM) walks the whole list for its tile; a team is one warp, or (SPLIT = 2)
 *     two warps on the same scheduler that own half of the rate categories each.  The B
 *     fragments of a (child, rate) matrix are pulled into registers once per operation and
 *     reused for both groups;
 *   - every result tile is written to a shared-memory SLOT of the team in the natural CLV layout
 *     (rows of R x 20 doubles, padded so that fragment reads and writes are bank-conflict free)
 *     and leaves for HBM as one TMA bulk store per pattern row (cp.async.bulk.global.shared::cta):
 *     no compute lane issues a global store.  The same slot is the tile cache: a parent reads
 *     its children's A fragments from their slots.  Four slots per team: two that the host
 *     planner (build_plan, plg_partials.cu) manages as the tile cache, two that alternate as
 *     the home of results which the very next operation consumes;
 *   - tip-tip operations do no arithmetic at all: the pack kernel multiplies the two tip tables
 *     into a pair table [left code][right code][rate][state] (L2 resident), the walk gathers
 *     one row per pattern into the slot by TMA and stores it from there.  Tip-inner operations
 *     gather the tip's table rows into the result slot the same way and multiply in place;
 *   - the rescaling vote (all R x 20 entries of a pattern below 2^-256, reference
 *     src/core_partials_avx2.c:788-801) is local to the team; a tile in which some pattern
 *     rescales (rare) is fixed up in its slot before the store is issued;
 *   - operation data (descriptor + the matrix sets as B fragments, exactly 3 200 bytes per
 *     matrix) streams through a 2-stage ring filled by a producer warp with TMA bulk copies.
 *
 * Arithmetic: the same DMMA chains in the same order as k_partial_dmma_aa (plg_dmma.cuh): CLVs
 * and scaler counts are bit-identical to the level-by-level path.
 * Replaces the loop of reference src/partials.c:184-212 over src/core_partials_avx2.c:568-803 /
 * src/core_partials_avx.c:1097-1340, :531-579.
 */
#include "plg_internal.cuh"
#include "plg_async.cuh"
#include "plg_dmma.cuh"

#define AW_TILE 16
#define AW_TEAMS 4
#define AW_STAGES 2
#define AW_SLOTS 4
#define AW_CODEBUFS 4

template <int R>
struct AwGeom
{
  static constexpr int ROWB = R * 160; /* bytes of one pattern row */
  /* byte offset of row s inside a slot: the two rows of a quarter warp must start 64 bytes apart
   * modulo 128 (four lanes of a pattern touch 64 contiguous bytes) */
  __host__ __device__ static constexpr int row_off(int s)
  {
    return R == 4 ? s * 640 + ((s + 1) >> 1) * 64 : (R == 2 ? s * 320 : s * 192);
  }
  static constexpr int ROWS_BYTES = row_off(AW_TILE);
  static constexpr int SLOT_BYTES = ROWS_BYTES + 64; /* + per-pattern scaler counts */
  static constexpr int MAT_BYTES = 3200;             /* one 20 x 20 matrix as B fragments */
  static constexpr int REC_BYTES = 128 + 2 * R * MAT_BYTES;
};

/* ------------------------------------------------------------------------------------ */
/* packed records and tip tables                                                         */
/* ------------------------------------------------------------------------------------ */
/* One block per operation.  Record: descriptor | right matrix set | left matrix set, a matrix as
 * [pair j < 5][lane][2] (fragments 2j, 2j + 1 of N tiles 0 and 1), [pair j < 2][lane < 16][2]
 * (fragments 10..13: N tile 2 has parent states 16..19 only), [lane < 16] (fragment 14).
 * Tip tables (rows of R x 20 doubles at tables + FusedOp::rbytes rows): sequential sum, in
 * increasing state order, of P_rate[i][m] over the states m in tipmap[code] - exactly
 * k_tip_tables_aa (reference src/core_partials_avx.c:1140-1177, :177-220); a tip-tip operation
 * gets the products of all code pairs (reference src/core_partials_avx.c:531-579). */
template <int R>
__global__ void k_walk_pack_aa(const FusedOp * __restrict__ ops, unsigned char * __restrict__ records,
                               unsigned char * __restrict__ tables, unsigned int ncodes, const TipmapArg tm)
{
  using G = AwGeom<R>;
  extern __shared__ double side_tab[]; /* [2][ncodes][R * 20] */
  const FusedOp f = ops[blockIdx.x];
  unsigned char * rec = records + (size_t)blockIdx.x * G::REC_BYTES;
  if (threadIdx.x < sizeof(FusedOp) / 8)
    reinterpret_cast<unsigned long long *>(rec)[threadIdx.x] =
        reinterpret_cast<const unsigned long long *>(ops + blockIdx.x)[threadIdx.x];
  for (int side = 0; side < 2; ++side) /* 0: right (first in the record), 1: left */
  {
    const bool matrix = side == 0 ? (f.kind != PLG_KIND_TT) : (f.kind == PLG_KIND_II);
    if (!matrix) continue;
    const double * src = side == 0 ? f.rsrc : f.lsrc;
    double * dst = reinterpret_cast<double *>(rec + 128 + (size_t)side * R * G::MAT_BYTES);
    for (unsigned int t = threadIdx.x; t < (unsigned int)R * 400u; t += blockDim.x)
    {
      const unsigned int k = t / 400u, w = t % 400u;
      unsigned int frag, lane;
      if (w < 320u) { frag = (w >> 6) * 2u + (w & 1u); lane = (w >> 1) & 31u; }
      else if (w < 384u) { frag = 10u + ((w - 320u) >> 5) * 2u + (w & 1u); lane = ((w - 320u) >> 1) & 15u; }
      else { frag = 14u; lane = w - 384u; }
      dst[t] = dmma_bfrag_value(src + (size_t)k * 400, frag, lane);
    }
  }
  if (f.kind == PLG_KIND_II) return;
  const unsigned int row_doubles = R * 20u;
  const unsigned int nsides = f.kind == PLG_KIND_TT ? 2u : 1u;
  for (unsigned int t = threadIdx.x; t < nsides * ncodes * row_doubles; t += blockDim.x)
  {
    const unsigned int side = t / (ncodes * row_doubles), e = t % (ncodes * row_doubles);
    const unsigned int code = e / row_doubles, within = e % row_doubles;
    const unsigned int k = within / 20u, i = within % 20u;
    const double * row = (side ? f.rsrc : f.lsrc) + (size_t)k * 400 + i * 20;
    const unsigned int state = tm.map[code];
    double s = 0.0;
    for (unsigned int m = 0; m < 20u; ++m)
      if ((state >> m) & 1u) s = __dadd_rn(s, row[m]);
    side_tab[t] = s;
  }
  __syncthreads();
  double * out = reinterpret_cast<double *>(tables + (size_t)f.rbytes * G::ROWB);
  if (f.kind == PLG_KIND_TI)
    for (unsigned int t = threadIdx.x; t < ncodes * row_doubles; t += blockDim.x) out[t] = side_tab[t];
  else
    for (unsigned int t = threadIdx.x; t < ncodes * ncodes * row_doubles; t += blockDim.x)
    {
      const unsigned int lc = t / (ncodes * row_doubles), rem = t % (ncodes * row_doubles);
      const unsigned int rc = rem / row_doubles, e = rem % row_doubles;
      out[t] = __dmul_rn(side_tab[lc * row_doubles + e], side_tab[(ncodes + rc) * row_doubles + e]);
    }
}

/* ------------------------------------------------------------------------------------ */
/* device helpers                                                                        */
/* ------------------------------------------------------------------------------------ */
/* One elected lane of a converged warp (the same lane every time: its bulk async-groups are the
 * ones the warp waits on).  The TMA instructions take warp-uniform operands: issued under this
 * predicate with uniform addresses they are single instructions, issued per lane with lane-
 * dependent addresses the compiler serialises them into a 30-instruction loop per copy. */
__device__ __forceinline__ bool aw_elect()
{
  unsigned int p;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
  return p != 0;
}

/* B fragments of one matrix from its packed block (smem) */
__device__ __forceinline__ void aw_load_bfrag(const unsigned char * __restrict__ blk, unsigned int lane,
                                              double (&B)[PLG_DMMA_FRAGS])
{
  const double2 * b2 = reinterpret_cast<const double2 *>(blk) + lane;
#pragma unroll
  for (int j = 0; j < 5; ++j)
  {
    const double2 v = b2[j * 32];
    B[2 * j] = v.x;
    B[2 * j + 1] = v.y;
  }
  if (lane < 16u)
  {
    const double2 * c2 = reinterpret_cast<const double2 *>(blk + 2560) + lane;
    const double2 v0 = c2[0], v1 = c2[16];
    B[10] = v0.x; B[11] = v0.y; B[12] = v1.x; B[13] = v1.y;
    B[14] = reinterpret_cast<const double *>(blk + 3072)[lane];
  }
  else
    B[10] = B[11] = B[12] = B[13] = B[14] = 0.0;
}

/* A fragments of (pattern row, rate) from a slot: `p` points at the row's rate block + 16 q */
__device__ __forceinline__ void aw_afrag_smem(const unsigned char * p, unsigned int hi_off, double (&a)[5])
{
  const double2 x = *reinterpret_cast<const double2 *>(p);
  const double2 y = *reinterpret_cast<const double2 *>(p + 64);
  a[0] = x.x; a[1] = x.y; a[2] = y.x; a[3] = y.y;
  a[4] = *reinterpret_cast<const double *>(p + hi_off);
}

/* the same from HBM (tile-cache miss): coherent loads - the tile may have been stored earlier
 * in this launch (by this warp's own TMA stores, completed and fenced by the caller) */
__device__ __forceinline__ void aw_afrag_hbm(const double * r, unsigned int q, unsigned int hi_state, bool ok,
                                             double (&a)[5])
{
  if (ok)
  {
    const double2 x = ld_stream_coherent2(r + 2 * q);
    const double2 y = ld_stream_coherent2(r + 8 + 2 * q);
    a[0] = x.x; a[1] = x.y; a[2] = y.x; a[3] = y.y;
    a[4] = ld_stream_coherent1(r + hi_state);
  }
  else
    a[0] = a[1] = a[2] = a[3] = a[4] = 0.0;
}

/* per-lane constants of a math warp */
struct AwLane
{
  unsigned int lane, g, q, hi_state, hi_off;
  unsigned int roff[2]; /* byte offset of this lane's quarter in its pattern row of group 0 / 1 */
};

/* The arithmetic of one tip-inner / inner-inner operation on a tile: products into the result
 * slot, `below` = every entry this lane produced is under the rescaling threshold.  FAST: both
 * inner children sit in slots and the tip rows were gathered into the result slot (the common
 * case, no run-time source decisions).  The result slot may only be touched once `ready` has
 * completed (the DMA warp has drained the store that last read it / landed the tip rows). */
template <int R, int KIND, bool FAST>
__device__ __forceinline__ void aw_compute(const unsigned char * stage, const unsigned char * slots, unsigned char * oslot,
                                           int lphys, int rphys, const double * left, const double * right,
                                           const unsigned char * tip_rows, const unsigned int (&tcode)[2], bool gather,
                                           uint64_t * ready, unsigned int ready_parity, bool & ready_seen, const AwLane & L,
                                           unsigned int first_site, const bool (&ok)[2], bool (&below)[2])
{
  using G = AwGeom<R>;
  const unsigned int q = L.q;
#pragma unroll
  for (int k = 0; k < R; ++k)
  {
    double BR[PLG_DMMA_FRAGS], BL[PLG_DMMA_FRAGS];
    aw_load_bfrag(stage + 128 + k * G::MAT_BYTES, L.lane, BR);
    if (KIND == PLG_KIND_II) aw_load_bfrag(stage + 128 + (R + k) * G::MAT_BYTES, L.lane, BL);
    /* all fragment loads of this rate category before its first store: the result slot may be
     * one of the children's (in place), and the loads then overlap instead of trailing the stores */
    double arr[2][5], all_[2][5];
#pragma unroll
    for (int sg = 0; sg < 2; ++sg)
    {
      if (FAST || rphys >= 0) aw_afrag_smem(slots + rphys * G::SLOT_BYTES + L.roff[sg] + k * 160u, L.hi_off, arr[sg]);
      else aw_afrag_hbm(right + (size_t)(first_site + sg * 8 + L.g) * (R * 20) + k * 20, q, L.hi_state, ok[sg], arr[sg]);
      if (KIND == PLG_KIND_II)
      {
        if (FAST || lphys >= 0) aw_afrag_smem(slots + lphys * G::SLOT_BYTES + L.roff[sg] + k * 160u, L.hi_off, all_[sg]);
        else aw_afrag_hbm(left + (size_t)(first_site + sg * 8 + L.g) * (R * 20) + k * 20, q, L.hi_state, ok[sg], all_[sg]);
      }
    }
    double y[2][3][2], x[2][3][2];
#pragma unroll
    for (int sg = 0; sg < 2; ++sg)
    {
      const double (&ar)[5] = arr[sg];
      const double (&al)[5] = all_[sg];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) y[sg][nt][0] = y[sg][nt][1] = x[sg][nt][0] = x[sg][nt][1] = 0.0;
      if (KIND == PLG_KIND_II)
      {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
          {
            dmma884_free(y[sg][nt][0], y[sg][nt][1], ar[ks], BR[nt * 5 + ks]);
            dmma884_free(x[sg][nt][0], x[sg][nt][1], al[ks], BL[nt * 5 + ks]);
          }
      }
      else
      {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) dmma884_free(y[sg][nt][0], y[sg][nt][1], ar[ks], BR[nt * 5 + ks]);
      }
    }
    if (!ready_seen)
    {
      plg_async::mbar_wait(ready, ready_parity);
      ready_seen = true;
    }
#pragma unroll
    for (int sg = 0; sg < 2; ++sg)
    {
      if (KIND == PLG_KIND_TI)
      {
        if (FAST || gather)
        {
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
            if (nt < 2 || q < 2u)
            {
              const double2 t = *reinterpret_cast<const double2 *>(oslot + L.roff[sg] + k * 160u + nt * 64);
              x[sg][nt][0] = t.x;
              x[sg][nt][1] = t.y;
            }
        }
        else
        {
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
            if (nt < 2 || q < 2u)
            {
              const double2 t = __ldg(reinterpret_cast<const double2 *>(tip_rows + (size_t)tcode[sg] * G::ROWB + k * 160u +
                                                                         nt * 64 + 16u * q));
              x[sg][nt][0] = t.x;
              x[sg][nt][1] = t.y;
            }
        }
      }
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
        if (nt < 2 || q < 2u)
        {
          const double p0 = __dmul_rn(x[sg][nt][0], y[sg][nt][0]);
          const double p1 = __dmul_rn(x[sg][nt][1], y[sg][nt][1]);
          below[sg] = below[sg] && (p0 < PLG_SCALE_THRESHOLD) && (p1 < PLG_SCALE_THRESHOLD);
          *reinterpret_cast<double2 *>(oslot + L.roff[sg] + k * 160u + nt * 64) = make_double2(p0, p1);
        }
    }
  }
}

/* where the tiles of an operation live.  Both warps of a team run this on the same descriptors
 * and therefore agree: a result the planner keeps (pslot >= 0) goes to that cache slot, any other
 * to one of the two staging slots, which alternate each time one is used - so the staging slot
 * of a tip-tip operation is normally idle while the operation before it runs (gather look-ahead) */
struct AwSlots
{
  int prev_out;
  unsigned int toggle;
  int out, lphys, rphys;
  __device__ __forceinline__ void place(int lslot, int rslot, int pslot)
  {
    out = pslot >= 0 ? pslot : 2 + (int)toggle;
    lphys = lslot == -2 ? prev_out : lslot;
    rphys = rslot == -2 ? prev_out : rslot;
  }
  __device__ __forceinline__ void advance(int pslot)
  {
    if (pslot < 0) toggle ^= 1u;
    prev_out = out;
  }
};

__device__ __forceinline__ unsigned int aw_atom_inc(unsigned int * p, unsigned int wrap)
{
  unsigned int old;
  asm volatile("atom.relaxed.cta.shared::cta.inc.u32 %0, [%1], %2;"
               : "=r"(old)
               : "r"(plg_async::smem_addr(p)), "r"(wrap)
               : "memory");
  return old;
}

/* TMA bulk copies with an L2 eviction-priority hint: tables and records are re-read by every
 * tile and should survive the 64 GB write stream (evict_last), result rows are never read again
 * by this launch (evict_first) */
__device__ __forceinline__ unsigned long long aw_policy_keep()
{
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long aw_policy_stream()
{
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void aw_g2s(void * dst_smem, const void * src_gmem, uint32_t bytes, uint64_t * bar,
                                       unsigned long long policy)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(plg_async::smem_addr(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(plg_async::smem_addr(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void aw_s2g(void * dst_gmem, const void * src_smem, uint32_t bytes, unsigned long long policy)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(plg_async::smem_addr(src_smem)), "r"(bytes), "l"(policy)
               : "memory");
}

/* ------------------------------------------------------------------------------------ */
/* the walk                                                                              */
/* ------------------------------------------------------------------------------------ */
#define AW_CONSUMERS (2 * AW_TEAMS)

template <int R>
__global__ void __launch_bounds__(2 * AW_TEAMS * 32, 1)
k_walk_aa(const FusedOp * __restrict__ ops, const unsigned char * __restrict__ records,
          const unsigned char * __restrict__ tables, unsigned int n_ops, unsigned int sites, unsigned int ncodes)
{
  using namespace plg_async;
  using G = AwGeom<R>;

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char * stage_base = smem;                                             /* AW_STAGES records */
  unsigned char * slot_base = smem + AW_STAGES * G::REC_BYTES;                   /* [team][slot] */
  unsigned char * code_base = slot_base + AW_TEAMS * AW_SLOTS * G::SLOT_BYTES;   /* [team][buf][side][16] */
  uint64_t * full = reinterpret_cast<uint64_t *>(code_base + AW_TEAMS * AW_CODEBUFS * 32);
  uint64_t * empty = full + AW_STAGES;
  uint64_t * ready_all = empty + AW_STAGES;                                       /* [team][2] */
  uint64_t * done_all = ready_all + AW_TEAMS * 2;                                 /* [team][2] */
  unsigned int * ticket = reinterpret_cast<unsigned int *>(done_all + AW_TEAMS * 2); /* [AW_STAGES] */

  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int warp = threadIdx.x >> 5;
  const unsigned int ntiles = (sites + AW_TILE - 1) / AW_TILE;
  const unsigned int tiles_per_pass = gridDim.x * AW_TEAMS;
  const unsigned int passes = (ntiles + tiles_per_pass - 1) / tiles_per_pass;
  const unsigned int total_its = passes * n_ops;
  const unsigned long long keep = aw_policy_keep();

  if (threadIdx.x == 0)
  {
    for (int s = 0; s < AW_STAGES; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], AW_CONSUMERS);
      ticket[s] = 0;
    }
    for (int w = 0; w < AW_TEAMS * 2; ++w)
    {
      mbar_init(&ready_all[w], 1);
      mbar_init(&done_all[w], 1);
    }
    fence_barrier_init();
    for (unsigned int it0 = 0; it0 < (unsigned int)AW_STAGES && it0 < total_its; ++it0)
    {
      const unsigned int i0 = it0 % n_ops;
      const unsigned int bytes = ops[i0].lbytes & 0xffffu;
      mbar_arrive_expect_tx(&full[it0], bytes);
      aw_g2s(stage_base + it0 * G::REC_BYTES, records + (size_t)i0 * G::REC_BYTES, bytes, &full[it0], keep);
    }
  }
  __syncthreads();

  const unsigned int team = warp % AW_TEAMS;
  const bool is_math = warp < AW_TEAMS;
  unsigned char * const slots = slot_base + team * AW_SLOTS * G::SLOT_BYTES;
  unsigned char * const codes = code_base + team * AW_CODEBUFS * 32;
  uint64_t * const ready = ready_all + team * 2;
  uint64_t * const done = done_all + team * 2;

  /* hands ring stage s back; the last of the consumers to do so refills it with the record two
   * operations ahead, whose size travels in this operation's descriptor */
  auto release = [&](unsigned int s, unsigned int it, unsigned int i, unsigned int ahead_bytes)
  {
    __syncwarp();
    if (lane == 0)
    {
      mbar_arrive(&empty[s]);
      if (aw_atom_inc(&ticket[s], AW_CONSUMERS - 1) == AW_CONSUMERS - 1 && it + AW_STAGES < total_its)
      {
        mbar_wait(&empty[s], (it / AW_STAGES) & 1u);
        unsigned int i2 = i + AW_STAGES;
        while (i2 >= n_ops) i2 -= n_ops;
        mbar_arrive_expect_tx(&full[s], ahead_bytes);
        aw_g2s(stage_base + s * G::REC_BYTES, records + (size_t)i2 * G::REC_BYTES, ahead_bytes, &full[s], keep);
      }
    }
  };

  AwSlots sl;
  sl.prev_out = -1;
  sl.toggle = 0;
  unsigned int it = 0;

  if (is_math)
  {
    /* ================= math warp: fragments, DMMA, products, vote ================= */
    AwLane L;
    L.lane = lane;
    L.g = lane >> 2;
    L.q = lane & 3u;
    L.hi_state = dmma_child_state(4, L.q);
    L.hi_off = 8u * L.hi_state - 16u * L.q; /* from a row's rate block + 16 q to its ks = 4 state */
    L.roff[0] = (unsigned int)G::row_off((int)L.g) + 16u * L.q;
    L.roff[1] = (unsigned int)G::row_off((int)L.g + 8) + 16u * L.q;
    const unsigned int g = L.g, q = L.q;

    for (unsigned int pass = 0; pass < passes; ++pass)
    {
      const unsigned int tile = (pass * gridDim.x + blockIdx.x) * AW_TEAMS + team;
      const bool have = tile < ntiles;
      const unsigned int first_site = tile * AW_TILE;
      const unsigned int nrows = have ? min((unsigned int)AW_TILE, sites - first_site) : 0u;
      const bool ok[2] = {first_site + g < sites && have, first_site + 8 + g < sites && have};

      for (unsigned int i = 0; i < n_ops; ++i, ++it)
      {
        const unsigned int s = it % AW_STAGES;
        mbar_wait(&full[s], (it / AW_STAGES) & 1u);
        const unsigned char * stage = stage_base + s * G::REC_BYTES;
        const FusedOp & d = *reinterpret_cast<const FusedOp *>(stage);
        const unsigned int ahead = d.lbytes >> 16;
        if (!have)
        {
          release(s, it, i, ahead);
          continue;
        }
        const int kind = d.kind;
        const int mode = d.scale_mode;
        const int pad = d.pad;
        sl.place(d.lslot, d.rslot, d.pslot);
        const int out = sl.out, lphys = sl.lphys, rphys = sl.rphys;
        sl.advance(d.pslot);
        uint64_t * const rb = &ready[it & 1u];
        const unsigned int rpar = (it >> 1) & 1u;
        bool ready_seen = false;

        if (kind == PLG_KIND_TT)
        {
          /* nothing to compute: the DMA warp gathers the rows of the pair table, zeroes the
           * scaler counts and stores the tile; the next reader of the slot must see it landed */
          release(s, it, i, ahead);
          mbar_wait(rb, rpar);
          __syncwarp();
          if (lane == 0) mbar_arrive(&done[it & 1u]);
          continue;
        }

        unsigned int * const pscale = d.op.pscale;
        const double * const left = d.op.left;
        const double * const right = d.op.right;
        const unsigned int * const lscale = d.op.lscale;
        const unsigned int * const rscale = d.op.rscale;
        unsigned char * const oslot = slots + out * G::SLOT_BYTES;
        const unsigned char * tip_rows = tables + (size_t)d.rbytes * G::ROWB;
        const bool gather = kind == PLG_KIND_TI && out != rphys;
        const bool miss = (kind == PLG_KIND_II && lphys < 0) || rphys < 0;

        bool below[2] = {true, true};
        unsigned int tcode[2] = {0u, 0u};
        if (miss || (kind == PLG_KIND_TI && !gather))
        {
          /* children read back from HBM need this team's stores landed, in-place tip rows need
           * the tip characters: both are behind the DMA warp's `ready` */
          mbar_wait(rb, rpar);
          ready_seen = true;
          if (kind == PLG_KIND_TI && !gather)
          {
            const unsigned char * cbuf = codes + (it % AW_CODEBUFS) * 32;
            tcode[0] = min((unsigned int)cbuf[g], ncodes - 1u);
            tcode[1] = min((unsigned int)cbuf[8 + g], ncodes - 1u);
          }
        }
        if (kind == PLG_KIND_II)
        {
          if (lphys >= 0 && rphys >= 0)
            aw_compute<R, PLG_KIND_II, true>(stage, slots, oslot, lphys, rphys, left, right, tip_rows, tcode, gather, rb, rpar,
                                             ready_seen, L, first_site, ok, below);
          else
            aw_compute<R, PLG_KIND_II, false>(stage, slots, oslot, lphys, rphys, left, right, tip_rows, tcode, gather, rb,
                                              rpar, ready_seen, L, first_site, ok, below);
        }
        else
        {
          if (rphys >= 0 && gather)
            aw_compute<R, PLG_KIND_TI, true>(stage, slots, oslot, lphys, rphys, left, right, tip_rows, tcode, gather, rb, rpar,
                                             ready_seen, L, first_site, ok, below);
          else
            aw_compute<R, PLG_KIND_TI, false>(stage, slots, oslot, lphys, rphys, left, right, tip_rows, tcode, gather, rb,
                                              rpar, ready_seen, L, first_site, ok, below);
        }
        /* the ring stage has been read (matrices are in registers, the descriptor in locals) */
        release(s, it, i, ahead);

        unsigned int vote[2] = {0u, 0u};
        if (mode == 1)
        {
#pragma unroll
          for (int sg = 0; sg < 2; ++sg)
          {
            unsigned int b = __ballot_sync(0xffffffffu, below[sg]);
            b &= b >> 1;
            b &= b >> 2;
            vote[sg] = b & 0x11111111u; /* bit 4g: every entry of pattern g is below the threshold */
          }
          if ((vote[0] | vote[1]) != 0u)
          {
            /* rare: some pattern of the tile is rescaled - in its slot, before the store leaves */
#pragma unroll
            for (int sg = 0; sg < 2; ++sg)
              if ((vote[sg] >> (4u * g)) & 1u)
#pragma unroll 1
                for (int k = 0; k < R; ++k)
#pragma unroll
                  for (int nt = 0; nt < 3; ++nt)
                    if (nt < 2 || q < 2u)
                    {
                      double2 * ptr = reinterpret_cast<double2 *>(oslot + L.roff[sg] + k * 160u + nt * 64);
                      double2 t = *ptr;
                      t.x = __dmul_rn(t.x, PLG_SCALE_FACTOR);
                      t.y = __dmul_rn(t.y, PLG_SCALE_FACTOR);
                      *ptr = t;
                    }
          }
          /* scaler counts of the tile, kept next to it in the slot */
          if (q == 0u)
          {
            unsigned int * ostrip = reinterpret_cast<unsigned int *>(oslot + G::ROWS_BYTES);
#pragma unroll
            for (int sg = 0; sg < 2; ++sg)
            {
              unsigned int sv = (vote[sg] >> (4u * g)) & 1u;
              if (kind == PLG_KIND_II && lscale)
                sv += lphys >= 0 ? reinterpret_cast<const unsigned int *>(slots + lphys * G::SLOT_BYTES + G::ROWS_BYTES)[sg * 8 + g]
                                 : (ok[sg] ? ld_coherent_u32(lscale + first_site + sg * 8 + g) : 0u);
              if (rscale)
                sv += rphys >= 0 ? reinterpret_cast<const unsigned int *>(slots + rphys * G::SLOT_BYTES + G::ROWS_BYTES)[sg * 8 + g]
                                 : (ok[sg] ? ld_coherent_u32(rscale + first_site + sg * 8 + g) : 0u);
              ostrip[sg * 8 + g] = sv;
              if ((pad & 1) && nrows < AW_TILE && ok[sg]) pscale[first_site + sg * 8 + g] = sv; /* ragged last tile */
            }
          }
        }
        /* the tile is complete: over to the DMA warp (its TMA reads come after this fence) */
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[it & 1u]);
      }
    }
    return;
  }

  /* ================= DMA warp: tip characters, gathers, stores ================= */
  const unsigned long long stream = aw_policy_stream();
  int last_store_slot = -1; /* source slot of the most recently committed store group */

  /* tip characters three operations ahead (cp.async of 16 bytes per tip row); what that needs of
   * the descriptor - kind and tip pointers - is loaded one operation earlier still: `pend`
   * describes operation it + 3 when iteration it starts */
  struct Pend { int kind; unsigned long long ltip, rtip; unsigned int tile; bool valid; } pend;
  auto load_pend = [&](unsigned int i_t, unsigned int pass_t)
  {
    while (i_t >= n_ops) { i_t -= n_ops; ++pass_t; }
    pend.valid = pass_t < passes;
    pend.tile = (pass_t * gridDim.x + blockIdx.x) * AW_TEAMS + team;
    if (pend.valid)
    {
      pend.kind = __ldg(&ops[i_t].kind);
      pend.ltip = __ldg(reinterpret_cast<const unsigned long long *>(&ops[i_t].op.ltip));
      pend.rtip = __ldg(reinterpret_cast<const unsigned long long *>(&ops[i_t].op.rtip));
    }
  };
  auto issue_codes = [&](unsigned int buf)
  {
    if (pend.valid && pend.kind != PLG_KIND_II && pend.tile < ntiles && lane < (pend.kind == PLG_KIND_TT ? 2u : 1u))
    {
      const unsigned char * src = reinterpret_cast<const unsigned char *>(lane ? pend.rtip : pend.ltip) + (size_t)pend.tile * AW_TILE;
      unsigned char * dst = codes + buf * 32 + lane * 16;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_pend(0, 0);
  issue_codes(0);
  load_pend(1, 0);
  issue_codes(1);
  load_pend(2, 0);
  issue_codes(2);
  load_pend(3, 0);

  /* what must happen before the math warp may touch the result slot of operation `it_x`: the
   * store that last read the slot has drained, tip rows are on their way (completion = bytes on
   * `ready`), a tile-cache miss finds this team's stores landed */
  auto prepare = [&](const FusedOp & dx, unsigned int it_x, int out_x, int rphys_x, bool miss_x, unsigned int nrows_x)
  {
    if (out_x == last_store_slot) bulk_wait_read<0>();
    else bulk_wait_read<1>();
    uint64_t * rb = &ready[it_x & 1u];
    unsigned char * oslot = slots + out_x * G::SLOT_BYTES;
    const int kind_x = dx.kind;
    const bool gather = kind_x == PLG_KIND_TT || (kind_x == PLG_KIND_TI && out_x != rphys_x);
    if (miss_x)
    {
      bulk_wait<0>();
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (gather)
    {
      const unsigned char * cbuf = codes + (it_x % AW_CODEBUFS) * 32;
      /* lane r works out the table row of pattern r; the elected lane needs the 16 byte offsets
       * as warp-uniform values: one shuffle each */
      unsigned int my_off = 0;
      if (lane < AW_TILE)
      {
        unsigned int row = min((unsigned int)cbuf[lane], ncodes - 1u);
        if (kind_x == PLG_KIND_TT)
        {
          row = row * ncodes + min((unsigned int)cbuf[16 + lane], ncodes - 1u);
          if (dx.scale_mode == 1) reinterpret_cast<unsigned int *>(oslot + G::ROWS_BYTES)[lane] = 0u;
        }
        my_off = row * (unsigned int)G::ROWB;
      }
      unsigned int offs[AW_TILE];
#pragma unroll
      for (int r = 0; r < AW_TILE; ++r) offs[r] = __shfl_sync(0xffffffffu, my_off, r);
      const unsigned char * src0 = tables + (size_t)dx.rbytes * G::ROWB;
      fence_proxy_async_smem(); /* the zeroed counts: read by the TMA store of this tile */
      __syncwarp();
      if (aw_elect())
      {
        mbar_arrive_expect_tx(rb, nrows_x * G::ROWB);
        if (nrows_x == AW_TILE)
        {
#pragma unroll
          for (int r = 0; r < AW_TILE; ++r) aw_g2s(oslot + G::row_off(r), src0 + offs[r], G::ROWB, rb, keep);
        }
        else
        {
#pragma unroll
          for (int r = 0; r < AW_TILE; ++r)
            if ((unsigned int)r < nrows_x) aw_g2s(oslot + G::row_off(r), src0 + offs[r], G::ROWB, rb, keep);
        }
      }
    }
    else
    {
      __syncwarp();
      if (lane == 0) mbar_arrive(rb);
    }
    __syncwarp();
  };

  bool prepared = false; /* the current operation was prepared during the previous one */
  for (unsigned int pass = 0; pass < passes; ++pass)
  {
    const unsigned int tile = (pass * gridDim.x + blockIdx.x) * AW_TEAMS + team;
    const bool have = tile < ntiles;
    const unsigned int first_site = tile * AW_TILE;
    const unsigned int nrows = have ? min((unsigned int)AW_TILE, sites - first_site) : 0u;

    for (unsigned int i = 0; i < n_ops; ++i, ++it)
    {
      const unsigned int s = it % AW_STAGES;
      /* tip characters: request those of it + 3; those of it and it + 1 have landed */
      issue_codes((it + 3u) % AW_CODEBUFS);
      load_pend(i + 4, pass);
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      __syncwarp();

      mbar_wait(&full[s], (it / AW_STAGES) & 1u);
      const FusedOp & d = *reinterpret_cast<const FusedOp *>(stage_base + s * G::REC_BYTES);
      const unsigned int ahead = d.lbytes >> 16;
      if (!have)
      {
        release(s, it, i, ahead);
        continue;
      }
      const int kind = d.kind;
      const int pad = d.pad;
      const int mode = d.scale_mode;
      double * const parent = d.op.parent;
      unsigned int * const pscale = d.op.pscale;
      const int pslot = d.pslot;
      sl.place(d.lslot, d.rslot, pslot);
      const int out = sl.out, lphys = sl.lphys, rphys = sl.rphys;
      if (!prepared)
      {
        const bool miss = (kind == PLG_KIND_II && lphys < 0) || (kind != PLG_KIND_TT && rphys < 0);
        prepare(d, it, out, rphys, miss, nrows);
      }
      prepared = false;
      sl.advance(pslot);
      /* everything this warp needs of the record is in registers: hand the stage back early, the
       * refill then has a whole operation to arrive */
      release(s, it, i, ahead);

      /* The next operation of this tile.  While the math warp is still busy with this one, tip
       * rows can already be gathered if nobody uses their slot; once it has finished, any slot
       * but this operation's own result is free.  Either way the math warp finds its next
       * operation ready before this one's stores are even issued. */
      const bool next_here = i + 1 < n_ops;
      const FusedOp & dn = *reinterpret_cast<const FusedOp *>(stage_base + (s ^ 1u) * G::REC_BYTES);
      const unsigned int next_parity = ((it + 1u) / AW_STAGES) & 1u;
      if (next_here && mbar_try_wait(&full[s ^ 1u], next_parity))
      {
        AwSlots sn = sl;
        sn.place(dn.lslot, dn.rslot, dn.pslot);
        const bool gather_n = dn.kind == PLG_KIND_TT || (dn.kind == PLG_KIND_TI && sn.out != sn.rphys && sn.rphys >= 0);
        if (gather_n && sn.out != out && sn.out != lphys && sn.out != rphys)
        {
          prepare(dn, it + 1u, sn.out, sn.rphys, false, nrows);
          prepared = true;
        }
      }

      /* the tile is complete when its rows have landed (tip-tip) or the math warp says so */
      if (kind == PLG_KIND_TT) mbar_wait(&ready[it & 1u], (it >> 1) & 1u);
      else mbar_wait(&done[it & 1u], (it >> 1) & 1u);

      if (next_here && !prepared)
      {
        mbar_wait(&full[s ^ 1u], next_parity);
        AwSlots sn = sl;
        sn.place(dn.lslot, dn.rslot, dn.pslot);
        const bool miss_n = (dn.kind == PLG_KIND_II && sn.lphys < 0) || (dn.kind != PLG_KIND_TT && sn.rphys < 0);
        /* a child read back from HBM may be this very result: its store comes first */
        if (!miss_n && sn.out != out)
        {
          prepare(dn, it + 1u, sn.out, sn.rphys, false, nrows);
          prepared = true;
        }
      }

      unsigned char * const oslot = slots + out * G::SLOT_BYTES;
      if (pad & 1)
      {
        if (kind == PLG_KIND_TT && mode == 1 && nrows < AW_TILE && lane < nrows) pscale[first_site + lane] = 0u;
        __syncwarp();
        if (aw_elect())
        {
          double * dst0 = parent + (size_t)first_site * (R * 20);
          if (nrows == AW_TILE)
          {
#pragma unroll
            for (int r = 0; r < AW_TILE; ++r) aw_s2g(dst0 + (size_t)r * (R * 20), oslot + G::row_off(r), G::ROWB, stream);
            if (mode == 1) aw_s2g(pscale + first_site, oslot + G::ROWS_BYTES, 64, stream);
          }
          else
          {
#pragma unroll
            for (int r = 0; r < AW_TILE; ++r)
              if ((unsigned int)r < nrows) aw_s2g(dst0 + (size_t)r * (R * 20), oslot + G::row_off(r), G::ROWB, stream);
          }
        }
        last_store_slot = out;
      }
      else
        last_store_slot = -1;
      bulk_commit();
    }
  }
  /* the last stores must have left shared memory before the CTA goes away */
  bulk_wait<0>();
}

/* ------------------------------------------------------------------------------------ */
template <int R>
static size_t walk_smem()
{
  using G = AwGeom<R>;
  return (size_t)AW_STAGES * G::REC_BYTES + (size_t)AW_TEAMS * AW_SLOTS * G::SLOT_BYTES + AW_TEAMS * AW_CODEBUFS * 32 +
         (2 * AW_STAGES + 4 * AW_TEAMS) * sizeof(uint64_t) + AW_STAGES * sizeof(unsigned int) + 8;
}

size_t plg_walk_aa_record_bytes(unsigned int rate_cats) { return 128 + 2 * (size_t)rate_cats * 3200; }
size_t plg_walk_aa_row_bytes(unsigned int rate_cats) { return (size_t)rate_cats * 160; }

bool plg_walk_aa_supported(unsigned int rate_cats, unsigned int ncodes)
{
  return (rate_cats == 1 || rate_cats == 2 || rate_cats == 4) && ncodes >= 1 &&
         2 * (size_t)ncodes * rate_cats * 20 * sizeof(double) <= 96 * 1024;
}

template <int R>
static int launch_walk(plg_context * ctx, const FusedOp * dev_ops, unsigned char * dev_records, unsigned char * dev_tables,
                       unsigned int n_ops)
{
  const size_t smem = walk_smem<R>();
  const size_t pack_smem = 2 * (size_t)ctx->maxstates * R * 20 * sizeof(double);
  static bool configured[PLG_MAX_DEVICES] = {};
  if (!configured[ctx->device % PLG_MAX_DEVICES])
  {
    PLG_CUDA(cudaFuncSetAttribute(k_walk_aa<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PLG_CUDA(cudaFuncSetAttribute(k_walk_pack_aa<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured[ctx->device % PLG_MAX_DEVICES] = true;
  }
  const unsigned int ntiles = (ctx->d.sites + AW_TILE - 1) / AW_TILE;
  unsigned int blocks = (unsigned int)ctx->sm_count;
  const unsigned int want = (ntiles + AW_TEAMS - 1) / AW_TEAMS;
  if (want < blocks) blocks = want;
  TipmapArg tm;
  memcpy(tm.map, ctx->tipmap, sizeof(tm.map));
  k_walk_pack_aa<R><<<n_ops, 256, pack_smem, ctx->stream>>>(dev_ops, dev_records, dev_tables, ctx->maxstates, tm);
  k_walk_aa<R><<<blocks, 2 * AW_TEAMS * 32, smem, ctx->stream>>>(dev_ops, dev_records, dev_tables, n_ops, ctx->d.sites,
                                                                 ctx->maxstates);
  return PLG_OK;
}

int plg_launch_walk_aa(plg_context * ctx, const FusedOp * dev_ops, unsigned char * dev_records, unsigned char * dev_tables,
                       unsigned int n_ops)
{
  switch (ctx->d.rate_cats)
  {
    case 1: return launch_walk<1>(ctx, dev_ops, dev_records, dev_tables, n_ops);
    case 2: return launch_walk<2>(ctx, dev_ops, dev_records, dev_tables, n_ops);
    case 4: return launch_walk<4>(ctx, dev_ops, dev_records, dev_tables, n_ops);
    default: plg_set_error("20-state walk: rate_cats=%u unsupported", ctx->d.rate_cats); return PLG_E_UNSUPPORTED;
  }
}
