/*
 * plg_dmma.cuh - the FP64 tensor-core building block of the 20-state kernels
 * (mma.sync.aligned.m8n8k4.f64, SASS DMMA.8x8x4), shared by the level-by-level kernel
 * (plg_partials.cu: k_partial_dmma_aa) and the single-kernel walk (plg_walk_aa.cu).
 *
 *   Y[s][i] = sum_j c[s][j] * P[i][j]      M = 8 sites (rows of A), N = parent states in three
 *                                          tiles of 8 (20 -> 24, rows 20..23 of P are zero), K = 20
 *                                          child states in five k-steps of 4
 *
 * Fragment ownership of lane (g, q) = (lane >> 2, lane & 3):
 *   A (ks)      c[site g][child_state(ks, q)]
 *   B (nt, ks)  P[8 nt + g][child_state(ks, q)]
 *   D (nt)      Y[site g][8 nt + 2q], Y[site g][8 nt + 2q + 1]
 *
 * The order in which the 20 child states are fed is free as long as A and B agree.  It is
 * chosen so that a lane's five A values are (almost) the D values it already owns:
 *   ks 0..3 -> states 2q, 2q+1, 8+2q, 9+2q      (the lane's own D values of tiles 0 and 1)
 *   ks 4    -> 16 + 2(q & 1) + (q >> 1)          (q = 0, 1: own first value of tile 2;
 *                                                 q = 2, 3: the SECOND value of lane q - 2)
 * so the result tile of one operation becomes the A operand of the next with one 64-bit shuffle,
 * without a trip through shared memory; from memory the same five values are two 16-byte loads
 * and one 8-byte load.  Both kernels use the same order and the same accumulation chains
 * (ks outer, one chain per N tile), hence produce bit-identical CLVs.
 */
#ifndef PLG_DMMA_CUH_
#define PLG_DMMA_CUH_

__host__ __device__ __forceinline__ unsigned int dmma_child_state(unsigned int ks, unsigned int q)
{
  return (ks < 2) ? 2u * q + ks : (ks < 4) ? 6u + 2u * q + ks : 16u + 2u * (q & 1u) + (q >> 1);
}

__device__ __forceinline__ void dmma884(double & d0, double & d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

/* The same instruction without `volatile`: the compiler may interleave other work (stores, loads,
 * the epilogue of the previous group) between the DMMAs of straight-line, convergent code. */
__device__ __forceinline__ void dmma884_free(double & d0, double & d1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

/* doubles of one P-matrix set as B fragments: [rate][nt * 5 + ks][lane] */
#define PLG_DMMA_FRAGS 15
__host__ __device__ constexpr unsigned int dmma_bfrag_doubles(unsigned int R) { return R * PLG_DMMA_FRAGS * 32u; }

/* value of B fragment f = nt * 5 + ks for `lane`, from a row-major 20 x 20 matrix */
__device__ __forceinline__ double dmma_bfrag_value(const double * __restrict__ M, unsigned int f, unsigned int lane)
{
  const unsigned int nt = f / 5u, ks = f % 5u;
  const unsigned int row = 8u * nt + (lane >> 2);
  const unsigned int col = dmma_child_state(ks, lane & 3u);
  return (row < 20u) ? M[row * 20u + col] : 0.0;
}

#endif /* PLG_DMMA_CUH_ */
