/*
 * plg_walk_aa.cu - the whole operations list of pll_update_partials in ONE kernel, 20 states,
 * the default for lists that recycle CLV / scaler slots (most of their stores are dead: 16-22 %
 * faster than the level-by-level kernels of plg_partials.cu), PLL_GPU_FUSED_AA=1 for every list
 * (at BASELINE configs[2], one slot per node, it takes 24.4 ms against their 22.3-23.5 ms:
 * DESIGN.md section 3).
 *
 * Patterns are independent, so a tile of patterns can walk the entire list with the children it
 * has just produced kept on chip; only results leave for HBM (64 GB instead of 128 GB at
 * configs[2]).  The kernel is a warp-specialised pipeline, 12 warps per SM:
 *
 *   - a TILE is 16 patterns x all rate categories (two 8-pattern DMMA groups).  Four tile teams
 *     per SM walk the list; a team is two MATH warps, each owning half of the rate categories,
 *     and one DMA warp.  The B fragments of a (child, rate) matrix are pulled into registers once
 *     per operation and reused for both groups.  The two math warps that share a scheduler belong
 *     to different teams; registers move from the DMA to the math warps (setmaxnreg: 64 / 216);
 *   - every result tile is written to a shared-memory SLOT of the team in the CLV layout (two
 *     contiguous 8-row halves, see AwGeom) and leaves for HBM as two TMA bulk stores issued by the
 *     DMA warp: no compute lane issues a global store.  The same slot is the tile cache: a parent
 *     reads its children's A fragments from their slots.  Four slots per team: two that the host
 *     planner (build_plan, plg_partials.cu) manages as the tile cache, two that alternate as the
 *     home of results which the very next operation consumes;
 *   - tip-tip operations do no arithmetic at all: the pack kernel multiplies the two tip tables
 *     into a pair table [left code][right code][rate][state] (L2 resident, evict_last), the DMA
 *     warp gathers one row per pattern into the slot with 16-byte asynchronous copies - ahead of
 *     time whenever the slot is idle - and stores the tile from there.  Tip-inner operations read
 *     the tip's table rows with read-only loads issued before the DMMAs of a category;
 *   - math and DMA warps meet on mbarriers only: `ready` (the result slot may be written: the
 *     store that last read it has drained / the rows have landed), `done` (the tile is complete),
 *     `voted` (the two halves of a team exchange their rescaling votes: all R x 20 entries of a
 *     pattern below 2^-256, reference src/core_partials_avx2.c:788-801; a tile in which some
 *     pattern rescales - rare - is fixed up in its slot before the store is issued);
 *   - operation data (descriptor + the matrix sets as B fragments, exactly 3 200 bytes per
 *     matrix) streams through a 2-stage ring of TMA bulk copies, fed by the DMA warp of team 0
 *     whenever it would otherwise spin.
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
#define AW_CODEBUFS 8

template <int R>
struct AwGeom
{
  static constexpr int ROWB = R * 160; /* bytes of one pattern row */
  /* A slot holds the 16 pattern rows of a tile as two contiguous halves (rows 0..7, rows 8..15:
   * each half is ONE bulk copy to HBM instead of eight row copies the DMA warp would have to issue),
   * the second half 64 bytes further modulo 128.  A DMMA group takes four rows of each half
   * (AwLane::row): the two patterns of a quarter warp then sit 64 bytes apart modulo 128 and the
   * four lanes of a pattern touch 64 contiguous bytes - fragment loads and stores are
   * bank-conflict free. */
  __host__ __device__ static constexpr int row_off(int s) { return s < 8 ? s * ROWB : 8 * ROWB + 64 + (s - 8) * ROWB; }
  static constexpr int HALF_BYTES = 8 * ROWB;
  static constexpr int ROWS_BYTES = 16 * ROWB + 64;
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
  unsigned int row[2];  /* the tile row (pattern) this lane works on in group 0 / 1 */
  unsigned int roff[2]; /* byte offset of this lane's quarter in that row */
  unsigned int k0;      /* first rate category of this warp */
};

/* The arithmetic of one tip-inner / inner-inner operation on a tile: products into the result
 * slot, `below` = every entry this lane produced is under the rescaling threshold.  FAST: both
 * inner children sit in slots and the tip rows were gathered into the result slot (the common
 * case, no run-time source decisions).  The result slot may only be touched once `ready` has
 * completed (the DMA warp has drained the store that last read it / landed the tip rows). */
template <int R, int NRW, int KIND, bool FAST>
__device__ __forceinline__ void aw_compute(const unsigned char * stage, const unsigned char * slots, unsigned char * oslot,
                                           int lphys, int rphys, const double * left, const double * right,
                                           const unsigned char * tip_rows, const unsigned int (&tcode)[2],
                                           uint64_t * ready, unsigned int ready_parity, bool & ready_seen, const AwLane & L,
                                           unsigned int first_site, const bool (&ok)[2], bool (&below)[2])
{
  using G = AwGeom<R>;
  const unsigned int q = L.q;
  /* fragments of rate category k + 1 are requested between the DMMAs of category k and its
   * epilogue: the compiler cannot move shared-memory loads above the epilogue's stores (the
   * result slot may alias a child's), in this order their latency hides behind the epilogue */
  double BR[PLG_DMMA_FRAGS], BL[PLG_DMMA_FRAGS];
  double arr[2][5], all_[2][5];
  auto load_frags = [&](int kk)
  {
    const unsigned int k = L.k0 + kk;
    aw_load_bfrag(stage + 128 + k * G::MAT_BYTES, L.lane, BR);
    if (KIND == PLG_KIND_II) aw_load_bfrag(stage + 128 + (R + k) * G::MAT_BYTES, L.lane, BL);
#pragma unroll
    for (int sg = 0; sg < 2; ++sg)
    {
      if (FAST || rphys >= 0) aw_afrag_smem(slots + rphys * G::SLOT_BYTES + L.roff[sg] + k * 160u, L.hi_off, arr[sg]);
      else aw_afrag_hbm(right + (size_t)(first_site + L.row[sg]) * (R * 20) + k * 20, q, L.hi_state, ok[sg], arr[sg]);
      if (KIND == PLG_KIND_II)
      {
        if (FAST || lphys >= 0) aw_afrag_smem(slots + lphys * G::SLOT_BYTES + L.roff[sg] + k * 160u, L.hi_off, all_[sg]);
        else aw_afrag_hbm(left + (size_t)(first_site + L.row[sg]) * (R * 20) + k * 20, q, L.hi_state, ok[sg], all_[sg]);
      }
    }
  };
  load_frags(0);
#pragma unroll
  for (int kk = 0; kk < NRW; ++kk)
  {
    const unsigned int k = L.k0 + kk;
    double y[2][3][2], x[2][3][2];
    if (KIND == PLG_KIND_TI)
    {
      /* the tip's table rows (L2 / L1 resident, read-only for this launch) are requested before
       * the DMMAs of the category and arrive behind them */
#pragma unroll
      for (int sg = 0; sg < 2; ++sg)
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
#pragma unroll
    for (int sg = 0; sg < 2; ++sg)
    {
      const double (&ar)[5] = arr[sg];
      const double (&al)[5] = all_[sg];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
      {
        y[sg][nt][0] = y[sg][nt][1] = 0.0;
        if (KIND == PLG_KIND_II) x[sg][nt][0] = x[sg][nt][1] = 0.0;
      }
      if (KIND == PLG_KIND_II)
      {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
          {
#ifdef AW_EXP_NOMMA
            y[sg][nt][0] += ar[ks] * BR[nt * 5 + ks];
            x[sg][nt][0] += al[ks] * BL[nt * 5 + ks];
#else
            dmma884_free(y[sg][nt][0], y[sg][nt][1], ar[ks], BR[nt * 5 + ks]);
            dmma884_free(x[sg][nt][0], x[sg][nt][1], al[ks], BL[nt * 5 + ks]);
#endif
          }
      }
      else
      {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
#ifdef AW_EXP_NOMMA
            y[sg][nt][0] += ar[ks] * BR[nt * 5 + ks];
#else
            dmma884_free(y[sg][nt][0], y[sg][nt][1], ar[ks], BR[nt * 5 + ks]);
#endif
      }
    }
    if (kk + 1 < NRW) load_frags(kk + 1);
    if (!ready_seen)
    {
      plg_async::mbar_wait(ready, ready_parity);
      ready_seen = true;
    }
#pragma unroll
    for (int sg = 0; sg < 2; ++sg)
    {
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
#define AW_MATH_WARPS (2 * AW_TEAMS)           /* two per team: half of the rate categories each */
#define AW_WARPS (3 * AW_TEAMS)                /* + one DMA warp per team */
#define AW_CONSUMERS AW_WARPS
#define AW_REGS_DMA 64
#define AW_REGS_MATH 216                       /* 8 x 32 x 216 + 4 x 32 x 64 <= 12 x 32 x 168 */

template <int R>
__global__ void __launch_bounds__(AW_WARPS * 32, 1)
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
  uint64_t * voted_all = done_all + AW_TEAMS * 2;                                 /* [team][2] */
  unsigned int * votes_all = reinterpret_cast<unsigned int *>(voted_all + AW_TEAMS * 2); /* [team][2][half][2] */

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
    }
    for (int w = 0; w < AW_TEAMS * 2; ++w)
    {
      mbar_init(&ready_all[w], 32);
      mbar_init(&done_all[w], 2);
      mbar_init(&voted_all[w], 2);
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

  /* the two math warps of a scheduler (warps w and w + 4) belong to DIFFERENT teams: the halves
   * of one team move in lockstep - same phase, nothing to overlap - while two teams drift apart
   * and one's DMMAs fill the other's epilogue */
  const bool is_math = warp < AW_MATH_WARPS;
  const unsigned int team = (warp + (is_math ? warp / AW_TEAMS : 0u)) % AW_TEAMS;
  /* the two math warps of a team sit on the same scheduler (warps w and w + 4): while one waits
   * or runs its epilogue the other keeps the tensor pipe busy.  Registers move from the DMA
   * warps to the math warps (both warpgroup-wide) */

  unsigned char * const slots = slot_base + team * AW_SLOTS * G::SLOT_BYTES;
  unsigned char * const codes = code_base + team * AW_CODEBUFS * 32;
  uint64_t * const ready = ready_all + team * 2;
  uint64_t * const done = done_all + team * 2;
  uint64_t * const voted = voted_all + team * 2;
  unsigned int * const votes = votes_all + team * 8;

  /* hands ring stage s back */
  auto release = [&](unsigned int s)
  {
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  };

  AwSlots sl;
  sl.prev_out = -1;
  sl.toggle = 0;
  unsigned int it = 0;

  if (is_math)
  {
    /* ================= math warp: fragments, DMMA, products, vote ================= */
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AW_REGS_MATH));
    AwLane L;
    L.lane = lane;
    L.g = lane >> 2;
    L.q = lane & 3u;
    L.hi_state = dmma_child_state(4, L.q);
    L.hi_off = 8u * L.hi_state - 16u * L.q; /* from a row's rate block + 16 q to its ks = 4 state */
    L.row[0] = (L.g & 1u) * 8u + (L.g >> 1);
    L.row[1] = L.row[0] + 4u;
    L.roff[0] = (unsigned int)G::row_off((int)L.row[0]) + 16u * L.q;
    L.roff[1] = (unsigned int)G::row_off((int)L.row[1]) + 16u * L.q;
    constexpr int NRW = R >= 2 ? R / 2 : 1;
    const unsigned int half = warp / AW_TEAMS;
    const bool ghost = R < 2 && half == 1; /* one category: the second warp only keeps the protocol */
    L.k0 = ghost ? 0u : half * NRW;
    const unsigned int g = L.g, q = L.q;
    unsigned int vote_phase = 0;
#ifdef AW_EXP_TIMING
    long long tb[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* full, tt, compute, vote wait, epilogue, total */
    long long t_prev = clock64();
    const long long t_begin = t_prev;
#define AW_TICK(b) { const long long t_now = clock64(); tb[b] += t_now - t_prev; t_prev = t_now; }
#else
#define AW_TICK(b)
#endif

    for (unsigned int pass = 0; pass < passes; ++pass)
    {
      const unsigned int tile = (pass * gridDim.x + blockIdx.x) * AW_TEAMS + team;
      const bool have = tile < ntiles;
      const unsigned int first_site = tile * AW_TILE;
      const unsigned int nrows = have ? min((unsigned int)AW_TILE, sites - first_site) : 0u;
      const bool ok[2] = {first_site + L.row[0] < sites && have, first_site + L.row[1] < sites && have};

      for (unsigned int i = 0; i < n_ops; ++i, ++it)
      {
        const unsigned int s = it % AW_STAGES;
        AW_TICK(6)
        mbar_wait(&full[s], (it / AW_STAGES) & 1u);
        AW_TICK(0)
        const unsigned char * stage = stage_base + s * G::REC_BYTES;
        const FusedOp & d = *reinterpret_cast<const FusedOp *>(stage);
        if (!have)
        {
          release(s);
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
          release(s);
          mbar_wait(rb, rpar);
          __syncwarp();
          if (lane == 0) mbar_arrive(&done[it & 1u]);
          AW_TICK(1)
          continue;
        }

        unsigned int * const pscale = d.op.pscale;
        const double * const left = d.op.left;
        const double * const right = d.op.right;
        const unsigned int * const lscale = d.op.lscale;
        const unsigned int * const rscale = d.op.rscale;
        unsigned char * const oslot = slots + out * G::SLOT_BYTES;
        const unsigned char * tip_rows = tables + (size_t)d.rbytes * G::ROWB;
        const bool miss = (kind == PLG_KIND_II && lphys < 0) || rphys < 0;

        bool below[2] = {true, true};
        unsigned int tcode[2] = {0u, 0u};
        if (kind == PLG_KIND_TI)
        {
          /* tip characters of this operation: landed before the DMA warp signalled `ready` for the
           * previous one (it waits for those of three operations at the top of an iteration) */
          const unsigned char * cbuf = codes + (it % AW_CODEBUFS) * 32;
          tcode[0] = min((unsigned int)cbuf[L.row[0]], ncodes - 1u);
          tcode[1] = min((unsigned int)cbuf[L.row[1]], ncodes - 1u);
        }
        if (miss)
        {
          /* children read back from HBM need this team's stores landed: behind `ready` */
          mbar_wait(rb, rpar);
          ready_seen = true;
        }
        if (ghost) { }
        else if (kind == PLG_KIND_II)
        {
          if (lphys >= 0 && rphys >= 0)
            aw_compute<R, NRW, PLG_KIND_II, true>(stage, slots, oslot, lphys, rphys, left, right, tip_rows, tcode, rb, rpar,
                                             ready_seen, L, first_site, ok, below);
          else
            aw_compute<R, NRW, PLG_KIND_II, false>(stage, slots, oslot, lphys, rphys, left, right, tip_rows, tcode, rb,
                                              rpar, ready_seen, L, first_site, ok, below);
        }
        else
        {
          if (rphys >= 0)
            aw_compute<R, NRW, PLG_KIND_TI, true>(stage, slots, oslot, lphys, rphys, left, right, tip_rows, tcode, rb, rpar,
                                             ready_seen, L, first_site, ok, below);
          else
            aw_compute<R, NRW, PLG_KIND_TI, false>(stage, slots, oslot, lphys, rphys, left, right, tip_rows, tcode, rb,
                                              rpar, ready_seen, L, first_site, ok, below);
        }
        /* the ring stage has been read (matrices are in registers, the descriptor in locals) */
        release(s);
        AW_TICK(2)

        unsigned int vote[2] = {0u, 0u};
        if (mode == 1)
        {
#pragma unroll
          for (int sg = 0; sg < 2; ++sg)
          {
            unsigned int b = __ballot_sync(0xffffffffu, below[sg]);
            b &= b >> 1;
            b &= b >> 2;
            vote[sg] = b & 0x11111111u; /* bit 4g: every entry of pattern g (this warp's categories) is below */
          }
          /* per-site scaling: both halves of the categories must agree */
          unsigned int * v = votes + (it & 1u) * 4;
          if (lane < 2) v[half * 2 + lane] = lane ? vote[1] : vote[0];
          __syncwarp();
          if (lane == 0) mbar_arrive(&voted[it & 1u]);
          AW_TICK(4)
          mbar_wait(&voted[it & 1u], (vote_phase >> (it & 1u)) & 1u);
          AW_TICK(3)
          vote_phase ^= 1u << (it & 1u); /* not every operation votes: the barriers keep their own phase */
          vote[0] &= v[(half ^ 1u) * 2 + 0];
          vote[1] &= v[(half ^ 1u) * 2 + 1];
          if ((vote[0] | vote[1]) != 0u && !ghost)
          {
            /* rare: some pattern of the tile is rescaled - in its slot, before the store leaves */
#pragma unroll
            for (int sg = 0; sg < 2; ++sg)
              if ((vote[sg] >> (4u * g)) & 1u)
#pragma unroll 1
                for (int kk = 0; kk < NRW; ++kk)
#pragma unroll
                  for (int nt = 0; nt < 3; ++nt)
                    if (nt < 2 || q < 2u)
                    {
                      double2 * ptr = reinterpret_cast<double2 *>(oslot + L.roff[sg] + (L.k0 + kk) * 160u + nt * 64);
                      double2 t = *ptr;
                      t.x = __dmul_rn(t.x, PLG_SCALE_FACTOR);
                      t.y = __dmul_rn(t.y, PLG_SCALE_FACTOR);
                      *ptr = t;
                    }
          }
          /* scaler counts of the tile, kept next to it in the slot */
          if (q == 0u && half == 0)
          {
            unsigned int * ostrip = reinterpret_cast<unsigned int *>(oslot + G::ROWS_BYTES);
#pragma unroll
            for (int sg = 0; sg < 2; ++sg)
            {
              unsigned int sv = (vote[sg] >> (4u * g)) & 1u;
              if (kind == PLG_KIND_II && lscale)
                sv += lphys >= 0 ? reinterpret_cast<const unsigned int *>(slots + lphys * G::SLOT_BYTES + G::ROWS_BYTES)[L.row[sg]]
                                 : (ok[sg] ? ld_coherent_u32(lscale + first_site + L.row[sg]) : 0u);
              if (rscale)
                sv += rphys >= 0 ? reinterpret_cast<const unsigned int *>(slots + rphys * G::SLOT_BYTES + G::ROWS_BYTES)[L.row[sg]]
                                 : (ok[sg] ? ld_coherent_u32(rscale + first_site + L.row[sg]) : 0u);
              ostrip[L.row[sg]] = sv;
            }
          }
        }
        /* the tile is complete: over to the DMA warp (its TMA reads come after this fence) */
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[it & 1u]);
        AW_TICK(4)
      }
    }
#ifdef AW_EXP_TIMING
    if (blockIdx.x == 3 && lane == 0)
      printf("warp %u: total %lld | ring %lld  tip-tip %lld  compute(+ready) %lld  vote-wait %lld  epilogue %lld  loop %lld\n", warp,
             clock64() - t_begin, tb[0], tb[1], tb[2], tb[3], tb[4], tb[6]);
#endif
    return;
  }

  /* ================= DMA warp: tip characters, gathers, stores ================= */
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AW_REGS_DMA));
  const unsigned long long stream = aw_policy_stream();
  /* The DMA warp of team 0 also feeds the operation ring: whenever it has to wait it checks
   * whether all consumers have handed back the stage of the next record to load (the record of
   * operation fill_it, two ahead of the one whose descriptor carried its size). */
  unsigned int fill_it = AW_STAGES, fill_i = AW_STAGES % n_ops;
  unsigned int fill_b0 = 0u, fill_b1 = 0u; /* AW_STAGES == 2 */
  unsigned int known_it = 0; /* sizes are known for operations < known_it + AW_STAGES */
  auto producer_poll = [&]()
  {
    if (team != 0 || fill_it >= total_its || fill_it >= known_it + AW_STAGES) return;
    const unsigned int fs = fill_it % AW_STAGES;
    if (!mbar_test_wait(&empty[fs], ((fill_it / AW_STAGES) - 1u) & 1u)) return;
    if (lane == 0)
    {
#ifdef AW_EXP_NORING
      const unsigned int fb = 128u;
#else
      const unsigned int fb = fs ? fill_b1 : fill_b0;
#endif
      mbar_arrive_expect_tx(&full[fs], fb);
      aw_g2s(stage_base + fs * G::REC_BYTES, records + (size_t)fill_i * G::REC_BYTES, fb, &full[fs], keep);
    }
    ++fill_it;
    if (++fill_i == n_ops) fill_i = 0;
  };
  auto wait_poll = [&](uint64_t * bar, unsigned int parity)
  {
    while (!mbar_test_wait(bar, parity)) producer_poll();
  };
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
  /* (an empty group after each: the loop commits two groups per iteration and counts on it) */
  load_pend(0, 0);
  issue_codes(0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  load_pend(1, 0);
  issue_codes(1);
  asm volatile("cp.async.commit_group;" ::: "memory");
  load_pend(2, 0);
  issue_codes(2);
  asm volatile("cp.async.commit_group;" ::: "memory");
  load_pend(3, 0);
  issue_codes(3);
  asm volatile("cp.async.commit_group;" ::: "memory");
  load_pend(4, 0);

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
    const bool gather = kind_x == PLG_KIND_TT;
    if (miss_x)
    {
      bulk_wait<0>();
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (gather)
    {
      const unsigned char * cbuf = codes + (it_x % AW_CODEBUFS) * 32;
      /* lane r works out the table row of pattern r */
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
      /* 16-byte asynchronous copies through the load/store path (the rows are scattered over
       * the table: twenty warp-wide cp.async instead of sixteen single-lane TMA requests).  A row is
       * CH chunks: whole warps take its first 32 k chunks, the tails of several rows share a warp
       * instruction.  Completion: every lane arrives on `ready` when its copies have landed. */
      const unsigned char * src0 = tables + (size_t)dx.rbytes * G::ROWB;
#ifdef AW_EXP_NOGATHER
      nrows_x = 0;
#endif
      constexpr int CH = G::ROWB / 16; /* 40, 20, 10 */
      constexpr int FULLW = CH / 32, TAIL = CH % 32, RPT = TAIL ? 32 / TAIL : 1; /* rows per tail instruction */
#pragma unroll
      for (int r = 0; r < AW_TILE; ++r)
      {
        const unsigned int off_r = __shfl_sync(0xffffffffu, my_off, r);
        if ((unsigned int)r < nrows_x)
#pragma unroll
          for (int w = 0; w < FULLW; ++w)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(oslot + G::row_off(r) + (w * 32 + lane) * 16)),
                         "l"(src0 + off_r + (w * 32 + lane) * 16)
                         : "memory");
      }
      if (TAIL)
      {
#pragma unroll
        for (int r0 = 0; r0 < AW_TILE; r0 += RPT)
        {
          const unsigned int rr = r0 + lane / TAIL, c = FULLW * 32 + lane % TAIL;
          const unsigned int off_r = __shfl_sync(0xffffffffu, my_off, rr & 15u);
          if (lane < RPT * TAIL && rr < nrows_x)
          {
            const unsigned int dst_off = (rr < 8u ? rr * G::ROWB : 8u * G::ROWB + 64u + (rr - 8u) * G::ROWB) + c * 16u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(oslot + dst_off)), "l"(src0 + off_r + c * 16u)
                         : "memory");
          }
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr(rb)) : "memory");
    }
    else
      mbar_arrive(rb);
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
      /* tip characters: request those of it + 4; those of it .. it + 2 must have landed */
      issue_codes((it + 4u) % AW_CODEBUFS);
      load_pend(i + 5, pass);
      /* two groups are committed per iteration (tip characters, then whatever gathers the
       * iteration issued): all but the last four complete = the characters up to it + 2 */
      asm volatile("cp.async.wait_group 4;" ::: "memory");
      __syncwarp();

      wait_poll(&full[s], (it / AW_STAGES) & 1u);
      const FusedOp & d = *reinterpret_cast<const FusedOp *>(stage_base + s * G::REC_BYTES);
      if (s) fill_b1 = d.lbytes >> 16; /* size of the record of operation it + AW_STAGES */
      else fill_b0 = d.lbytes >> 16;
      known_it = it + 1u;
      if (!have)
      {
        asm volatile("cp.async.commit_group;" ::: "memory");
        release(s);
        producer_poll();
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
      release(s);

      /* The next operation of this tile.  While the math warp is still busy with this one, tip
       * rows can already be gathered if nobody uses their slot; once it has finished, any slot
       * but this operation's own result is free.  Either way the math warp finds its next
       * operation ready before this one's stores are even issued. */
      const bool next_here = i + 1 < n_ops;
      const FusedOp & dn = *reinterpret_cast<const FusedOp *>(stage_base + (s ^ 1u) * G::REC_BYTES);
      const unsigned int next_parity = ((it + 1u) / AW_STAGES) & 1u;
      bool prepared_blocked = false; /* the next operation's slot is in use by this one: after `done` */
      auto try_early = [&]()
      {
        if (prepared || !next_here || !mbar_test_wait(&full[s ^ 1u], next_parity)) return;
        AwSlots sn = sl;
        sn.place(dn.lslot, dn.rslot, dn.pslot);
        const bool miss_n = (dn.kind == PLG_KIND_II && sn.lphys < 0) || (dn.kind != PLG_KIND_TT && sn.rphys < 0);
        if (!miss_n && sn.out != out && sn.out != lphys && sn.out != rphys)
        {
          prepare(dn, it + 1u, sn.out, sn.rphys, false, nrows);
          prepared = true;
        }
        else
          prepared_blocked = true;
      };
      try_early();

      /* the tile is complete when its rows have landed (tip-tip) or the math warps say so */
      {
        uint64_t * bar = kind == PLG_KIND_TT ? &ready[it & 1u] : &done[it & 1u];
        while (!mbar_test_wait(bar, (it >> 1) & 1u))
        {
          producer_poll();
          if (!prepared_blocked) try_early();
        }
      }

      if (next_here && !prepared)
      {
        wait_poll(&full[s ^ 1u], next_parity);
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

      asm volatile("cp.async.commit_group;" ::: "memory"); /* this iteration's gathers */
      unsigned char * const oslot = slots + out * G::SLOT_BYTES;
      if (pad & 1)
      {
        if (mode == 1 && lane < nrows) pscale[first_site + lane] = reinterpret_cast<const unsigned int *>(oslot + G::ROWS_BYTES)[lane];
        fence_proxy_async_smem(); /* rows gathered by this warp's cp.async copies (tip-tip): TMA reads them next */
        __syncwarp();
        if (aw_elect())
        {
          double * dst0 = parent + (size_t)first_site * (R * 20);
          const unsigned int first_half = min(nrows, 8u);
#ifdef AW_EXP_NOSTORE
          if (nrows > 99u)
#endif
          {
          aw_s2g(dst0, oslot, first_half * G::ROWB, stream);
          if (nrows > 8u) aw_s2g(dst0 + 8 * (R * 20), oslot + G::HALF_BYTES + 64, (nrows - 8u) * G::ROWB, stream);
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
         (2 * AW_STAGES + 6 * AW_TEAMS) * sizeof(uint64_t) + AW_TEAMS * 8 * sizeof(unsigned int);
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
  k_walk_aa<R><<<blocks, AW_WARPS * 32, smem, ctx->stream>>>(dev_ops, dev_records, dev_tables, n_ops, ctx->d.sites,
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
