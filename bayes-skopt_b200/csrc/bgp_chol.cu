// K2: batched-theta FP64 Cholesky + forward solve + LML + log-prior, one CTA (or a cluster of
// 2-8 CTAs when the batch leaves SMs idle) per theta.  Left-looking blocked factorisation,
// 32-column panels:
//   * the Gram matrix arrives in the L2-resident factor slab, written by the K1 kernel of the
//     same stream (bgp_gram.cu) in exactly the tiled layout consumed here;
//   * trailing updates and the panel triangular solve are DMMA.8x8x4 GEMMs whose A operand
//     streams from the slab through a per-warp cp.async ring, B operand is staged in shared
//     memory;
//   * y rides along as one extra row, so z = L^-1 y (and y^T K^-1 y = |z|^2) falls out of the
//     same panel solves; in factorise mode identity rows ride along too and come out as
//     L^-T, which the candidate sweep consumes.
// Replaces sklearn:_gpr.py:583-617 + bask/bayesgpr.py:351-379 (logprob mode) and
// bask/bayesgpr.py:200-217 (factorise mode).
#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

constexpr int KCH = 15;        // previous panels staged in shared memory at once
constexpr int LS = 40;         // row stride of the factored diagonal block read by DMMA (== 8 mod 16)
constexpr int PS = 33;         // row stride of the diagonal block being factored (odd: conflict-free)

template <int NW>
struct CholSmem {
  DevProgram prog;
  ThetaParams tp;
  alignas(16) double Dblk[32 * PS];  // diagonal block being factored
  alignas(16) double Lt[32 * LS];    // factored block L_kk, DMMA-friendly stride
  alignas(16) double Wd[4][8 * 8];   // inverses of its four 8x8 diagonal sub-blocks
  double red[NW];
  int fail;
};

// ---- warp-level Cholesky of the 32x32 diagonal block + inverses of its 8x8 diagonal blocks --
// Right-looking over 8-column blocks with the block held in DMMA accumulator layout (lane (r,q)
// owns [8t+r][8u+2q..2q+1] of sub-tile (t,u)):
//   1. every lane reads the current 8x8 diagonal sub-block from shared memory (broadcast) and
//      factors it redundantly in registers -- no shuffles on the rsqrt -> scale -> fma chain;
//   2. lane c (mod 8) derives column c of W = L_bb^-1 from its register copy;
//   3. the sub-tiles below become X = C W^T and the trailing sub-tiles C -= X X^T on DMMA; the
//      accumulator layout doubles as the A and the B operand (k-permutation), so nothing moves.
// The panel solve only needs L_kk (DMMA-friendly copy Lt) and the four 8x8 inverses Wd.
// Returns 0 or failing local column + 1 (LAPACK dpotrf: pivot <= 0 or NaN).
__device__ __forceinline__ int warp_potrf32(double* __restrict__ D, double* __restrict__ Lt,
                                         double* __restrict__ Wd, int lane, double& logdet, int ncols_real,
                                            long long* ts = nullptr) {
  const int r = lane >> 2, q = lane & 3, ci = lane & 7;
#define POTRF_TS(i) do { if (ts && lane == 0) ts[i] = clock64(); } while (0)
  POTRF_TS(0);
  int fail = 0;
  double C[4][4][2];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u <= t; ++u) {
      C[t][u][0] = D[(8 * t + r) * PS + 8 * u + 2 * q];
      C[t][u][1] = D[(8 * t + r) * PS + 8 * u + 2 * q + 1];
    }
  double invprod = 1.0;   // lane c < 8: product of 1/L_jj over its columns j = c (mod 8)
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    if (b > 0) {
      D[(8 * b + r) * PS + 8 * b + 2 * q] = C[b][b][0];
      D[(8 * b + r) * PS + 8 * b + 2 * q + 1] = C[b][b][1];
      __syncwarp();
    }
    POTRF_TS(1 + 6 * b);
    double a[8][8], inv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) a[i][j] = D[(8 * b + i) * PS + 8 * b + j];
    POTRF_TS(2 + 6 * b);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      double ajj = a[c][c];
      if (!(ajj > 0.0)) { if (!fail) fail = 8 * b + c + 1; ajj = 1.0; }
      inv[c] = rsqrt(ajj);
      a[c][c] = ajj * inv[c];
#pragma unroll
      for (int i = c + 1; i < 8; ++i) a[i][c] *= inv[c];
#pragma unroll
      for (int j = c + 1; j < 8; ++j)
#pragma unroll
        for (int i = j; i < 8; ++i) a[i][j] = fma(-a[i][c], a[j][c], a[i][j]);
      if (lane == c && 8 * b + c < ncols_real) invprod *= inv[c];
    }
    POTRF_TS(3 + 6 * b);
    // column ci of W_bb by forward substitution on the register copy
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double sacc = (i == ci) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; ++k) sacc = fma(-a[i][k], x[k], sacc);
      x[i] = (i >= ci) ? sacc * inv[i] : 0.0;
    }
    POTRF_TS(4 + 6 * b);
    if (lane < 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) Wd[b * 64 + i * 8 + ci] = x[i];
    }
    __syncwarp();
    POTRF_TS(5 + 6 * b);
    {
      // X = C W^T for the sub-tiles of this block column; for the diagonal sub-tile itself this
      // reproduces L_bb (C_bb = L_bb L_bb^T, only the lower part of C_bb enters the lower part
      // of the product) in accumulator layout, which is what the Lt copy is written from
      const double2 wv = *reinterpret_cast<const double2*>(Wd + b * 64 + r * 8 + 2 * q);
#pragma unroll
      for (int t = b; t < 4; ++t) {
        double o[2] = {0.0, 0.0};
        dmma(o, C[t][b][0], wv.x);
        dmma(o, C[t][b][1], wv.y);
        C[t][b][0] = o[0]; C[t][b][1] = o[1];
      }
      if (2 * q > r) C[b][b][0] = 0.0;
      if (2 * q + 1 > r) C[b][b][1] = 0.0;
#pragma unroll
      for (int u = b + 1; u < 4; ++u)
#pragma unroll
        for (int t = u; t < 4; ++t) {
          dmma(C[t][u], C[t][b][0], -C[u][b][0]);
          dmma(C[t][u], C[t][b][1], -C[u][b][1]);
        }
    }
  }
  POTRF_TS(25);
  if (lane < 8) logdet -= log(invprod);
  POTRF_TS(26);
  // DMMA-friendly copy of L_kk, zeros above the diagonal
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<double2*>(Lt + (8 * t + r) * LS + 8 * u + 2 * q) =
          u <= t ? make_double2(C[t][u][0], C[t][u][1]) : make_double2(0.0, 0.0);
  __syncwarp();
  POTRF_TS(27);
#undef POTRF_TS
  return fail;
}

// stage rows [c0, c0+32) x the columns of panels [j0, j0+kc) of L into shared memory with
// cp.async (16-byte LDGSTS, L2 only): every copy of a thread is in flight before the first wait,
// so the stage costs about one L2 round trip instead of one per loop iteration
__device__ __forceinline__ void stage_block_row_issue(const double* slab, const SlabGeom& G, double* Bs, int bstride,
                                                      int c0, int j0, int kc, int tid, int nthreads) {
  for (int idx = tid; idx < 512 * kc; idx += nthreads) {
    const int c2 = idx & 15, row = (idx >> 4) & 31, jj = idx >> 9;
    const double* src = slab + G.off(j0 + jj) + (size_t)(c0 + row - 32 * (j0 + jj)) * 32 + 2 * c2;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(Bs + (size_t)row * bstride + 32 * jj + 2 * c2);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void stage_block_row(const double* slab, const SlabGeom& G, double* Bs, int bstride,
                                                int c0, int j0, int kc, int tid, int nthreads) {
  stage_block_row_issue(slab, G, Bs, bstride, c0, j0, kc, tid, nthreads);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

__device__ __align__(16) const double g_zero16[2] = {0.0, 0.0};

struct TileSet {
  int rb[4];     // first storage row of each 8-row tile
  int kind[4];   // 0 training rows, 1 the y row, 2 identity rows, 3 unused slot
  int js[4];     // first previous panel with non-zero entries (identity rows only)
};

// acc[t][u] += A_t (8 x 32*kc, streamed from the slab) * B_u^T (shared memory) for the NTL tiles
// of this warp.  NTL is a template parameter on purpose: predicated-off DMMAs still occupy the
// FP64 pipe, so inactive tiles must not appear in the instruction stream at all.
//
// The A operand comes from L2 (~1000 cycles away, one 8-column step is only NTL x 128 cycles of
// DMMA issue), so it has to run several steps ahead.  A register ring does not survive ptxas (it
// loads into a temporary and copies into the ring slot right away, i.e. waits for every load in
// the step that issued it), so the ring lives in shared memory and is filled with cp.async:
// every lane copies its own 16-byte fragment into its own slot and cp.async.wait_group gives the
// exact "all but the newest N" wait that scoreboards cannot express.  ring: this warp's
// RING_BYTES of shared memory.
constexpr int RING_BYTES = 8192;
template <int NTL>
__device__ __forceinline__ void k_chunk(double (&acc)[4][4][2], const TileSet& TS, const double* slab,
                                        const SlabGeom& G, const double* Bs, int bstride, int j0, int kc,
                                        int r, int q, unsigned ring) {
  static_assert(NTL == 1 || NTL == 2, "ring sized for one or two tiles per warp");
  constexpr int ST = NTL == 1 ? 16 : 8;   // ring stages: ST * NTL * 512 B == RING_BYTES, multiple of 4
  int jmin = j0 + kc;
#pragma unroll
  for (int t = 0; t < NTL; ++t) jmin = min(jmin, TS.js[t]);
  const int jbeg = max(j0, jmin), jend = j0 + kc;
  if (jbeg >= jend) return;
  const int steps = 4 * (jend - jbeg);
  const double* tb[NTL];
#pragma unroll
  for (int t = 0; t < NTL; ++t) tb[t] = slab + (size_t)(TS.rb[t] + r) * 32 + 2 * q;
  // B operand: rows 8u + r of the staged block row, 8 columns (64 bytes) per step
  unsigned bsa[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
    bsa[u] = (unsigned)__cvta_generic_to_shared(Bs + (size_t)(8 * u + r) * bstride + 32 * (jbeg - j0) + 2 * q);
  // copy the fragments of step sn into stage `stage`; always commits a (possibly empty) group so
  // that the group count per step is uniform
  auto issue = [&](int sn, int stage) {
    if (sn < steps) {
      const int jn = jbeg + (sn >> 2);
      const int o = 32 * jn * (G.R - 16 * (jn + 1)) + 8 * (sn & 3);
#pragma unroll
      for (int t = 0; t < NTL; ++t) {
        const double* src = jn >= TS.js[t] ? tb[t] + o : g_zero16;   // identity rows: zero before js
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(ring + (stage * NTL + t) * 512), "l"(src));
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
#define BGP_LDS2(dst, addr) \
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"((dst).x), "=d"((dst).y) : "r"(addr))
#pragma unroll
  for (int s = 0; s < ST - 1; ++s) issue(s, s);
  // operands of the current / next step ping-pong between two statically indexed register sets
  // (a copy "cur = next" would make ptxas wait for the loads in the step that issued them)
  double2 av[2][NTL], bv[2][4];
  asm volatile("cp.async.wait_group %0;\n" ::"n"(ST - 2));
#pragma unroll
  for (int t = 0; t < NTL; ++t) BGP_LDS2(av[0][t], ring + t * 512);
#pragma unroll
  for (int u = 0; u < 4; ++u) BGP_LDS2(bv[0][u], bsa[u]);
  for (int st0 = 0; st0 < steps; st0 += ST) {
#pragma unroll
    for (int s = 0; s < ST; ++s) {
      const int st = st0 + s;
      if (st < steps) {
        // Everything in the step is volatile asm, so the order below is the issue order: the
        // loads for step st+1 (its fragments have landed once at most ST-3 younger groups are
        // pending) and the refill of the stage consumed in step st-1 sit between the DMMAs of
        // step st, whose 16-cycle issue slots hide them.
        asm volatile("cp.async.wait_group %0;\n" ::"n"(ST - 3));
#pragma unroll
        for (int i = 0; i < 8 * NTL; ++i) {
          const int h = i / (4 * NTL), t = (i >> 2) % NTL, u = i & 3;
          constexpr int c = 0;
          dmma(acc[t][u], h ? av[s & 1][t].y : av[s & 1][t].x, h ? bv[s & 1][u].y : bv[s & 1][u].x);
          if (i < NTL) BGP_LDS2(av[(s & 1) ^ 1][i], ring + (((s + 1) % ST) * NTL + i) * 512);
          if (i == NTL) issue(st + ST - 1, (s + ST - 1) % ST);
          // (not past the last step: the columns behind the K range are being written by the block-row pushes)
          if (i > NTL && i <= NTL + 4 && st + 1 < steps) BGP_LDS2(bv[(s & 1) ^ 1][i - NTL - 1], bsa[i - NTL - 1] + 64 * (st + 1));
          (void)c;
        }
      }
    }
  }
#undef BGP_LDS2
  asm volatile("cp.async.wait_all;\n" ::);
}

// Gram values of a warp's tiles for panel k, requested before the K-loop so that their L2 round
// trip is over when finish_tiles needs them (volatile asm: the loads stay where they are written)
template <int NTL>
__device__ __forceinline__ void prefetch_gram(double2 (&ginit)[2][4], const TileSet& TS, const double* slab,
                                              const SlabGeom& G, int k, int npad, int r, int q) {
  const int c0 = 32 * k;
#pragma unroll
  for (int t = 0; t < NTL; ++t) {
    const int rowc = TS.kind[t] == 0 ? min(TS.rb[t] + r, npad - 1) : c0;
    const double* src = slab + G.off(k) + (size_t)(rowc - c0) * 32 + 2 * q;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];\n"
                   : "=d"(ginit[t][u].x), "=d"(ginit[t][u].y)
                   : "l"(src + 8 * u));
  }
}

// C = init - acc;  X = C * L_kk^-T (block forward substitution on DMMA);  store X into panel k.
template <int NTL>
__device__ __forceinline__ void finish_tiles(double (&acc)[4][4][2], const TileSet& TS, const CholArgs& A,
                                             const double* Lt, const double* Wd, double* slab,
                                             const SlabGeom& G, int k, int n, int npad, int b, int r, int q,
                                             double& zz, const double2 (&ginit)[2][4], double* Bs_push, int bstride,
                                             int cs) {
  const int c0 = 32 * k;
#pragma unroll
  for (int t = 0; t < NTL; ++t) {
    const int row = TS.rb[t] + r;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int cl = 8 * u + 2 * q, col = c0 + cl;
      double2 v0 = make_double2(0.0, 0.0);
      if (TS.kind[t] == 0) {
        if (A.dense) {
          if (row < n && col < n) v0.x = A.dense[(size_t)row * A.ldd + col];
          if (row < n && col + 1 < n) v0.y = A.dense[(size_t)row * A.ldd + col + 1];
        } else {
          v0 = ginit[t][u];
          if (row >= n || col >= n) v0.x = 0.0;
          if (row >= n || col + 1 >= n) v0.y = 0.0;
        }
      } else if (TS.kind[t] == 1) {
        if (row == G.Rz && A.y) {
          if (col < n) v0.x = A.y[col];
          if (col + 1 < n) v0.y = A.y[col + 1];
        }
      } else {
        v0.x = (row - G.Ra == col) ? 1.0 : 0.0;
        v0.y = (row - G.Ra == col + 1) ? 1.0 : 0.0;
      }
      acc[t][u][0] = v0.x - acc[t][u][0];
      acc[t][u][1] = v0.y - acc[t][u][1];
    }
  }
  // X_b = (C_b - sum_{a<b} X_a L_ba^T) W_bb^T,  b = 0..3  (8-column blocks)
#pragma unroll
  for (int ub = 0; ub < 4; ++ub) {
#pragma unroll
    for (int ua = 0; ua < ub; ++ua) {
      const double2 lv = *reinterpret_cast<const double2*>(Lt + (8 * ub + r) * LS + 8 * ua + 2 * q);
#pragma unroll
      for (int t = 0; t < NTL; ++t) {
        dmma(acc[t][ub], acc[t][ua][0], -lv.x);
        dmma(acc[t][ub], acc[t][ua][1], -lv.y);
      }
    }
    const double2 wv = *reinterpret_cast<const double2*>(Wd + ub * 64 + r * 8 + 2 * q);
#pragma unroll
    for (int t = 0; t < NTL; ++t) {
      double o[2] = {0.0, 0.0};
      dmma(o, acc[t][ub][0], wv.x);
      dmma(o, acc[t][ub][1], wv.y);
      acc[t][ub][0] = o[0]; acc[t][ub][1] = o[1];
    }
  }
#pragma unroll
  for (int t = 0; t < NTL; ++t) {
    const int row = TS.rb[t] + r;
    double* dst = slab + G.off(k) + (size_t)(row - c0) * 32 + 2 * q;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<double2*>(dst + 8 * u) = make_double2(acc[t][u][0], acc[t][u][1]);
    // rows of block row k+1: this is column block k of the NEXT panel's B operand -- written straight
    // into the staging buffer of every CTA of the cluster (columns [32k, 32k+32) are not in use during
    // panel k), so the next panel does not start with an L2 round trip; the end-of-panel barrier
    // (release / acquire at cluster scope) publishes it
    if (Bs_push && TS.kind[t] == 0 && TS.rb[t] < c0 + 64) {
      double* lp = Bs_push + (size_t)(TS.rb[t] - (c0 + 32) + r) * bstride + c0 + 2 * q;
      if (cs == 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) *reinterpret_cast<double2*>(lp + 8 * u) = make_double2(acc[t][u][0], acc[t][u][1]);
      } else {
        const unsigned la = (unsigned)__cvta_generic_to_shared(lp);
        for (int pr = 0; pr < cs; ++pr) {
          unsigned ra;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(la), "r"(pr));
#pragma unroll
          for (int u = 0; u < 4; ++u)
            asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};\n" ::"r"(ra + 64 * u), "d"(acc[t][u][0]),
                         "d"(acc[t][u][1]) : "memory");
        }
      }
    }
    if (TS.kind[t] == 1 && r == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        zz = fma(acc[t][u][0], acc[t][u][0], zz);
        zz = fma(acc[t][u][1], acc[t][u][1], zz);
        if (A.z_out) {
          const int col = c0 + 8 * u + 2 * q;
          if (col < n) A.z_out[(size_t)b * n + col] = acc[t][u][0];
          if (col + 1 < n) A.z_out[(size_t)b * n + col + 1] = acc[t][u][1];
        }
      }
    }
  }
}

// The Gram values of the diagonal block a thread assembles (elements tid, tid + nthreads, ...),
// requested at the top of the panel so that their L2 round trip overlaps the stage + diagonal GEMM
template <int NE>
__device__ __forceinline__ void prefetch_diag(double (&g)[NE], const CholArgs& A, const double* slab,
                                              const SlabGeom& G, int kk, int n, int t0, int nthreads) {
  const int c0 = 32 * kk;
#pragma unroll
  for (int s = 0; s < NE; ++s) {
    const int e = t0 + s * nthreads;
    const int rl = (e >> 5) & 31, cl = e & 31;
    const int row = min(c0 + rl, n - 1), col = min(c0 + cl, n - 1);
    if (A.dense) {
      g[s] = A.dense[(size_t)row * A.ldd + col];
    } else {
      asm volatile("ld.global.cg.f64 %0, [%1];\n"
                   : "=d"(g[s])
                   : "l"(slab + G.off(kk) + (size_t)(row - c0) * 32 + (col - c0)));
    }
  }
}

// Dblk = init(diagonal block kk) - sum of the NP K-split partial products (when kk > 0); a partial
// holds the ten lower 8x8 sub-tiles in fragment order: [sub-tile][lane][2]
constexpr int PART_DOUBLES = 10 * 64;
template <int NE, int NP>
__device__ __forceinline__ void assemble_diag(const double (&g)[NE], const CholArgs& A, const double* Part,
                                              double* Dblk, int kk, int n, bool with_part, int t0, int nthreads) {
  const int c0 = 32 * kk;
#pragma unroll
  for (int s = 0; s < NE; ++s) {
    const int e = t0 + s * nthreads;
    if (e >= 1024) continue;
    const int rl = e >> 5, cl = e & 31;
    double v = 0.0;
    if (cl <= rl) {
      const int row = c0 + rl, col = c0 + cl;
      if (row < n && col < n) {
        v = g[s] + ((A.dense && row == col) ? A.jitter : 0.0);
        if (with_part) {
          const int t = rl >> 3, u = cl >> 3;
          const int idx = (t * (t + 1) / 2 + u) * 64 + (4 * (rl & 7) + ((cl & 7) >> 1)) * 2 + (cl & 1);
          double sum = 0.0;
#pragma unroll
          for (int w = 0; w < NP; ++w) sum += Part[w * PART_DOUBLES + idx];
          v -= sum;
        }
      } else {
        v = (row == col) ? 1.0 : 0.0;
      }
    }
    Dblk[rl * PS + cl] = v;
  }
}

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// CS = CTAs per theta.  With CS > 1 (used when the batch leaves SMs idle: CS * batch <= #SMs) a
// thread-block cluster of CS CTAs shares one matrix: all run the short serial chain (diagonal
// update, potrf) redundantly -- it is deterministic, so no exchange is needed -- and split the
// row tiles of the trailing update / panel solve; a cluster barrier per panel publishes the rows
// each of them wrote to the (L2-resident) slab.
constexpr int MAXT = 2;   // row tiles per warp and round

template <int NW, int CS>
__global__ void __launch_bounds__(NW * 32, 1) chol_lml_kernel(CholArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CholSmem<NW>& S = *reinterpret_cast<CholSmem<NW>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int crank = 0;
  if (CS > 1) asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int r = lane >> 2, q = lane & 3;
  const int n = A.n;
  const SlabGeom G = SlabGeom::make(n, A.aug != 0);
  const int P = G.P, npad = 32 * P;
  const int kch = P - 1 < KCH ? (P - 1 > 0 ? P - 1 : 1) : KCH;
  const int bstride = 32 * kch + 8;
  double* Bs = reinterpret_cast<double*>(smem_raw + ((sizeof(CholSmem<NW>) + 15) & ~size_t(15)));  // 32 x bstride
  // after Bs: the K-split partial sums of the diagonal update (phase 1) and the warps' cp.async
  // rings of the trailing update (phase 2) share one region
  double* region = Bs + (size_t)32 * bstride;
  double* Part = region;   // NW partials of PART_DOUBLES
  const unsigned ring = (unsigned)__cvta_generic_to_shared(region) + warp * RING_BYTES + lane * 16;
  if (A.prog) {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&S.prog);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += NW * 32) dst[i] = src[i];
  } else if (tid == 0) {
    S.prog.n_ops = 0; S.prog.n_theta = 0; S.prog.n_leaves = 0; S.prog.d = 0; S.prog.fast_kind = 0;
  }
  __syncthreads();
  // every CTA of the cluster is running before anyone stores into a peer's shared memory (the block-row pushes)
  if (CS > 1) cluster_barrier();
  const DevProgram& PR = S.prog;

  for (int b = blockIdx.x / CS; b < A.batch; b += gridDim.x / CS) {
    const double* theta = A.theta + (size_t)b * PR.n_theta;
    double* slab = A.slabs + (size_t)(A.slab_per_block ? blockIdx.x / CS : b) * G.doubles();
    if (tid == 0) S.fail = 0;
    double logdet = 0.0, zz = 0.0;   // meaningful in warp 0 / z-row owners
    __syncthreads();

#define BGP_STAMP(slot) do { if (A.dbg && blockIdx.x == 0 && tid == A.dbg_tid) A.dbg[k * 12 + (slot)] = clock64(); } while (0)
#define BGP_STAMP_ADD(slot, t0) do { if (A.dbg && blockIdx.x == 0 && tid == A.dbg_tid) A.dbg[k * 12 + (slot)] += clock64() - (t0); } while (0)
    for (int k = 0; k < P; ++k) {
      const int c0 = 32 * k;
      BGP_STAMP(0);
      // ------------------------------------------------ phase 1: diagonal block
      constexpr int NE = 1024 / (NW * 32);
      double gd[NE];
      prefetch_diag<NE>(gd, A, slab, G, k, n, tid, NW * 32);
      // D -= sum_j L[k,j] L[k,j]^T: every warp takes every NW-th 8-column step of the staged block
      // row and accumulates the ten lower 8x8 sub-tiles (the fragments of the four row tiles are
      // both the A and the B operands)
      double dacc[10][2];
#pragma unroll
      for (int i = 0; i < 10; ++i) dacc[i][0] = dacc[i][1] = 0.0;
      const int nchunks = (k + kch - 1) / kch;
      // pre_staged: the previous panel left this panel's block row in Bs -- columns of panels < k-1 by
      // cp.async issued after its last K-loop, column block k-1 pushed by the warps that solved those rows
      const bool pre_staged = k > 0 && k <= kch && NW > 1;
      if (pre_staged) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      for (int ch = 0; ch < nchunks; ++ch) {
        const int j0 = ch * kch, kc = min(kch, k - j0);
        __syncthreads();
        if (!pre_staged) {
          stage_block_row(slab, G, Bs, bstride, c0, j0, kc, tid, NW * 32);
          __syncthreads();
        }
        for (int c8 = warp; c8 < 4 * kc; c8 += NW) {
          double2 f[4];
#pragma unroll
          for (int t = 0; t < 4; ++t)
            f[t] = *reinterpret_cast<const double2*>(Bs + (size_t)(8 * t + r) * bstride + 8 * c8 + 2 * q);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
              for (int u = 0; u <= t; ++u)
                dmma(dacc[t * (t + 1) / 2 + u], h ? f[t].y : f[t].x, h ? f[u].y : f[u].x);
        }
      }
      if (k > 0) {
#pragma unroll
        for (int i = 0; i < 10; ++i)
          *reinterpret_cast<double2*>(Part + warp * PART_DOUBLES + i * 64 + lane * 2) = make_double2(dacc[i][0], dacc[i][1]);
      }
      __syncthreads();   // also makes the Gram panel (written earlier) visible to the whole CTA
      BGP_STAMP(1);
      assemble_diag<NE, NW>(gd, A, Part, S.Dblk, k, n, k > 0, tid, NW * 32);
      __syncthreads();
      BGP_STAMP(2);
      // Warp 0 factors the diagonal block.  When K is not chunked the other warps do not wait for
      // it: the trailing-update GEMM of their first round only needs previous panels, so they run
      // it now and meet warp 0 at named barrier 2 right before the panel solve.
      const bool overlap = nchunks <= 1 && NW > 1;
      // the next panel is pre-staged by this one (see pre_staged above)
      const bool next_pre = overlap && k + 1 < P && k + 1 <= kch;
      if (warp == 0) {
        int f = warp_potrf32(S.Dblk, S.Lt, &S.Wd[0][0], lane, logdet, min(32, n - c0));
        if (f && lane == 0) S.fail = c0 + f;
        // L_kk -> slab (diag group rows of panel k)
        if (crank == 0)
#pragma unroll
          for (int e = lane; e < 512; e += 32)
            *reinterpret_cast<double2*>(slab + G.off(k) + 2 * e) =
                *reinterpret_cast<const double2*>(S.Lt + (e >> 4) * LS + 2 * (e & 15));
        BGP_STAMP(3);
        if (overlap) asm volatile("bar.sync 2, %0;" ::"r"(NW * 32) : "memory");
        if (next_pre) {
          asm volatile("bar.sync 3, %0;" ::"r"(NW * 32) : "memory");   // every K-loop of this panel is done with Bs
          stage_block_row_issue(slab, G, Bs, bstride, c0 + 32, 0, k, tid, NW * 32);
        }
      } else if (A.dbg && warp == (A.dbg_tid >> 5)) {
        BGP_STAMP(3);
      }
      if (!overlap) {
        __syncthreads();
        if (S.fail) break;
      }
      BGP_STAMP(4);

      // ------------------------------------------------ phase 2: rows below the block, 8-row
      // tiles dealt evenly to the warps (at most MAXT per warp and round)
      // 32-row groups below the diagonal and identity-row groups; with CS == 2 groups alternate
      // between the two CTAs of the cluster and the y tile belongs to rank 0
      // tiles of this panel: the y tile, the 4 (P-1-k) tiles below the block, the identity-row
      // tiles; the CTAs of a cluster take contiguous, equally long parts of that list
      const int nmt = 4 * (P - 1 - k), nat = A.aug ? 4 * (k + 1) : 0;
      const int Tg = 1 + nmt + nat;
      // (rounded up, so that the y tile -- whose |z|^2 the epilogue of rank 0 sums -- stays on rank 0)
      const int tlo = (Tg * crank + CS - 1) / CS, T = (Tg * (crank + 1) + CS - 1) / CS - tlo;
      // warp w owns the contiguous tiles [w0, w0 + mine); every warp runs the same number of
      // rounds (barriers inside when K is chunked), each with at most MAXT of its tiles
      // worker warps (warp 0 is busy when overlapping).  The deal starts with the warps that have
      // an SM sub-partition to themselves and ends with warp NW/2, which shares its FP64 pipe with
      // the potrf warp: with few tiles left the K-loops run alone on their pipes and the
      // dependent DFMA chain of the potrf does not queue behind DMMAs.
      const int nwk = overlap ? NW - 1 : NW;
      const int wrk = !overlap ? warp
                      : NW == 8 ? (warp == 0 ? -1 : warp == 4 ? 6 : warp < 4 ? warp - 1 : warp - 2)
                                : warp - 1;
      // From panel HEAVY_K on the first K-loop round outlasts the potrf, which then has slack: warp 4 takes
      // its share of the FP64 pipe of sub-partition 0 (about a sixth of the tiles -- it runs alone there, at
      // two thirds of the issue rate of a pair) and the other six warps split the rest, extras going to the
      // three shared sub-partitions in turn.
      constexpr int HEAVY_K = 6;
      const bool heavy = overlap && NW == 8 && k >= HEAVY_K;
      const int t4 = heavy ? (T + 3) / 6 : 0;
      const int Tr = T - t4, nwr = heavy ? nwk - 1 : nwk;
      const int tq = Tr / nwr, trm = Tr % nwr;
      const bool is4 = heavy && warp == 4;
      const int mine = wrk < 0 ? 0 : is4 ? t4 : tq + (wrk < trm ? 1 : 0);
      const int w0 = wrk < 0 ? 0 : is4 ? Tr : wrk * tq + min(wrk, trm);
      const int rounds = max(1, (max(tq + (trm ? 1 : 0), t4) + MAXT - 1) / MAXT);
      const int per = (mine + rounds - 1) / rounds;
      for (int rd = 0; rd < rounds && wrk >= 0; ++rd) {
        const int first = w0 + rd * per;
        const int ntl = max(0, min(w0 + mine, first + per) - first);   // tiles of this warp and round
        TileSet TS;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int ti = first + t;
          TS.rb[t] = c0; TS.kind[t] = 3; TS.js[t] = 0;   // kind 3: padding slot of the tile set
          if (t < ntl) {
            const int gi = tlo + ti;
            if (gi == 0) {
              TS.rb[t] = G.Rz; TS.kind[t] = 1;
            } else if (gi <= nmt) {
              TS.rb[t] = 32 * (k + 1) + 8 * (gi - 1); TS.kind[t] = 0;
            } else {
              const int l2 = gi - 1 - nmt;
              TS.rb[t] = G.Ra + 8 * l2; TS.kind[t] = 2; TS.js[t] = l2 >> 2;
            }
          }
        }
        long long tph = clock64();
        double acc[4][4][2];
        double2 ginit[2][4];
        if (!A.dense) {
          if (ntl == 2) prefetch_gram<2>(ginit, TS, slab, G, k, npad, r, q);
          else if (ntl == 1) prefetch_gram<1>(ginit, TS, slab, G, k, npad, r, q);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
        for (int ch = 0; ch < nchunks; ++ch) {
          const int j0 = ch * kch, kc = min(kch, k - j0);
          if (nchunks > 1) {
            __syncthreads();
            stage_block_row(slab, G, Bs, bstride, c0, j0, kc, tid, NW * 32);
            __syncthreads();
          }
          switch (ntl) {
            case 2: k_chunk<2>(acc, TS, slab, G, Bs, bstride, j0, kc, r, q, ring); break;
            case 1: k_chunk<1>(acc, TS, slab, G, Bs, bstride, j0, kc, r, q, ring); break;
            default: break;
          }
        }
        BGP_STAMP_ADD(7, tph); tph = clock64();
        if (overlap && rd == 0) asm volatile("bar.sync 2, %0;" ::"r"(NW * 32) : "memory");
        if (next_pre && rd == rounds - 1) {
          asm volatile("bar.sync 3, %0;" ::"r"(NW * 32) : "memory");   // every K-loop of this panel is done with Bs
          stage_block_row_issue(slab, G, Bs, bstride, c0 + 32, 0, k, tid, NW * 32);
        }
        if (S.fail) continue;
        double zpart = 0.0;
        double* push = next_pre ? Bs : nullptr;
        switch (ntl) {
          case 2: finish_tiles<2>(acc, TS, A, S.Lt, &S.Wd[0][0], slab, G, k, n, npad, b, r, q, zpart, ginit, push, bstride, CS); break;
          case 1: finish_tiles<1>(acc, TS, A, S.Lt, &S.Wd[0][0], slab, G, k, n, npad, b, r, q, zpart, ginit, push, bstride, CS); break;
          default: break;
        }
        zz += zpart;
        BGP_STAMP_ADD(8, tph);
      }
      BGP_STAMP(5);
      if (CS > 1) cluster_barrier(); else __syncthreads();
      BGP_STAMP(6);
      if (S.fail) break;
    }

    asm volatile("cp.async.wait_group 0;\n" ::: "memory");   // a pre-stage issued before a failed panel
    // ------------------------------------------------------------------ epilogue
    zz = warp_sum(zz);
    if (warp == 0) logdet = warp_sum(logdet);
    if (lane == 0) S.red[warp] = zz;
    __syncthreads();
    if (tid == 0 && crank == 0) {
      double ztz = 0.0;
      for (int w = 0; w < NW; ++w) ztz += S.red[w];
      double lml, lp;
      if (S.fail) {
        lml = -INFINITY; lp = -INFINITY;
      } else {
        lml = -0.5 * ztz - logdet - 0.5 * n * 1.8378770664093453;
        lp = lml;
        if (A.priors) lp += log_prior(A.priors, A.n_priors, theta);
        if (A.lp_extra) lp += A.lp_extra[b];
        if (!isfinite(lp)) lp = -INFINITY;
      }
      if (A.lml) A.lml[b] = lml;
      if (A.lp) A.lp[b] = lp;
      if (A.info) A.info[b] = S.fail;
    }
    __syncthreads();
  }
}

static int pick_nw(int n) { return n <= 64 ? 4 : 8; }

static size_t chol_smem_bytes(int n) {
  const int P = (n + 31) / 32;
  const int kch = P - 1 < KCH ? (P - 1 > 0 ? P - 1 : 1) : KCH;
  size_t base = pick_nw(n) == 8 ? sizeof(CholSmem<8>) : sizeof(CholSmem<4>);
  base = (base + 15) & ~size_t(15);
  const size_t part = sizeof(double) * pick_nw(n) * PART_DOUBLES, rings = (size_t)pick_nw(n) * RING_BYTES;
  return base + sizeof(double) * (size_t)32 * (32 * kch + 8) + (part > rings ? part : rings);
}

// largest portable cluster size that still fits the batch on the chip
static int pick_cluster(int n, int grid_thetas, int sms) {
  if (pick_nw(n) != 8) return 1;
  for (int cs = 8; cs > 1; cs >>= 1)
    if (cs * grid_thetas <= sms) return cs;
  return 1;
}

template <int CS>
static cudaError_t launch_cluster(const CholArgs& A, int grid, size_t smem, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS * grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, chol_lml_kernel<8, CS>, A);
}

// opt-in to the large dynamic shared-memory carve-out (must happen outside stream capture).  The
// attribute is per function, not per handle: it is raised to the hardware maximum once for every
// variant, so handles with different n in one process cannot lower each other's limit.
constexpr int CHOL_SMEM_OPTIN = 227 * 1024;
cudaError_t prepare_chol(int n) {
  if (chol_smem_bytes(n) > (size_t)CHOL_SMEM_OPTIN) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(chol_lml_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_OPTIN);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(chol_lml_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_OPTIN);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(chol_lml_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_OPTIN);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(chol_lml_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_OPTIN);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(chol_lml_kernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_OPTIN);
  return e;
}

cudaError_t launch_chol(const CholArgs& A, int grid, int sms, cudaStream_t stream) {
  const size_t smem = chol_smem_bytes(A.n);
  if (pick_nw(A.n) == 4) {
    chol_lml_kernel<4, 1><<<grid, 128, smem, stream>>>(A);
    return cudaGetLastError();
  }
  switch (pick_cluster(A.n, grid, sms)) {
    case 8: return launch_cluster<8>(A, grid, smem, stream);
    case 4: return launch_cluster<4>(A, grid, smem, stream);
    case 2: return launch_cluster<2>(A, grid, smem, stream);
    default: chol_lml_kernel<8, 1><<<grid, 256, smem, stream>>>(A); return cudaGetLastError();
  }
}

}  // namespace bgp
