// K1+K2: batched-theta Gram build + FP64 Cholesky + forward solve + LML + log-prior,
// one CTA per theta.  Left-looking blocked factorisation, 32-column panels:
//   * the Gram matrix arrives in the L2-resident factor slab, written by the K1 kernel of the
//     same stream (bgp_gram.cu) in exactly the tiled layout consumed here;
//   * trailing updates and the panel triangular solve are DMMA.8x8x4 GEMMs whose A operand
//     streams from the slab in 16-byte loads, B operand is staged in shared memory;
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
constexpr int PP = 34;         // row stride of the K-split partial blocks

template <int NW>
struct CholSmem {
  DevProgram prog;
  ThetaParams tp;
  double Dblk[32 * PS];              // diagonal block being factored
  double Lt[32 * LS];                // factored block L_kk, DMMA-friendly stride
  double Wd[4][8 * 8];               // inverses of its four 8x8 diagonal sub-blocks
  double Part[4][32 * PP];           // K-split partial sums of the diagonal update
  double red[NW];
  int fail;
};

__device__ __forceinline__ double log_prior(const bgp_prior_t* pr, int n, const double* theta) {
  double lp = 0.0;
  for (int k = 0; k < n; ++k) {
    const double x = theta[k];
    const double* p = pr[k].p;
    switch (pr[k].kind) {
      case BGP_PRIOR_HALFNORMAL_SQRT:
        lp += -0.22579135264472744 /* 0.5*log(2/pi) */ - log(p[0]) - exp(x) / (2.0 * p[0] * p[0]) +
              0.5 * x - 0.6931471805599453;
        break;
      case BGP_PRIOR_ROUNDFLAT:
        lp += -2.0 * (exp(-2.0 * p[2] * (x - log(p[0]))) + exp(2.0 * p[3] * (x - log(p[1])))) -
              p[4] + x;
        break;
      case BGP_PRIOR_INVGAMMA:
        lp += p[0] * log(p[1]) - lgamma(p[0]) - (p[0] + 1.0) * x - p[1] * exp(-x) + x;
        break;
      case BGP_PRIOR_NORMAL: {
        double t = (x - p[0]) / p[1];
        lp += -0.5 * t * t - log(p[1]) - 0.9189385332046727;
      } break;
      default: break;
    }
  }
  return lp;
}

// ---- warp-level Cholesky of the 32x32 diagonal block + inverses of its 8x8 diagonal blocks --
// Lane i owns row i.  Columns are processed in blocks of eight held in registers: a rolled
// left-looking update from the previous column blocks (shared memory), then an unrolled
// in-register factorisation of the eight columns with shuffles, so the per-column critical
// path is shuffle -> rsqrt -> scale -> shuffle -> fma.  The panel solve only needs the four
// 8x8 inverses (block forward substitution on DMMA), computed at the end by all lanes at once.
// Returns 0 or failing local column + 1 (LAPACK dpotrf: pivot <= 0 or NaN).
__device__ __noinline__ int warp_potrf32(double* __restrict__ D, double* __restrict__ Lt,
                                         double* __restrict__ Wd, int lane, double& logdet, int ncols_real) {
  int fail = 0;
  double* __restrict__ myrow = D + lane * PS;
  double my_inv = 1.0, w[8];
  for (int b = 0; b < 4; ++b) {
#pragma unroll
    for (int c = 0; c < 8; ++c) w[c] = myrow[8 * b + c];
    for (int k = 0; k < 8 * b; ++k) {
      const double lik = myrow[k];
#pragma unroll
      for (int c = 0; c < 8; ++c) w[c] = fma(-lik, D[(8 * b + c) * PS + k], w[c]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = 8 * b + c;
      double ajj = __shfl_sync(0xffffffffu, w[c], j);
      if (!(ajj > 0.0)) { if (!fail) fail = j + 1; ajj = 1.0; }
      const double inv = rsqrt(ajj);
      const double lij = lane > j ? w[c] * inv : (lane == j ? ajj * inv : 0.0);
      if (lane == j) my_inv = inv;
      w[c] = lij;
#pragma unroll
      for (int c2 = c + 1; c2 < 8; ++c2) {
        const double lkj = __shfl_sync(0xffffffffu, lij, 8 * b + c2);
        w[c2] = fma(-lij, lkj, w[c2]);
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) myrow[8 * b + c] = w[c];
    __syncwarp();
  }
  if (lane < ncols_real) logdet -= log(my_inv);
  // after the loop w[] holds this lane's entries of column block 3; reload the diagonal block
  // row of this lane's own 8x8 sub-block: lanes 8g..8g+7 hold L_gg row-wise
  const int g = lane >> 3, ci = lane & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) w[c] = myrow[8 * g + c];
  // column ci of W_gg = L_gg^-1 by forward substitution; L_gg[i][k] lives in lane 8g+i, w[k]
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double sacc = (i == ci) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) {
      const double lik = __shfl_sync(0xffffffffu, w[k], i, 8);
      sacc = fma(-lik, x[k], sacc);
    }
    const double dinv = __shfl_sync(0xffffffffu, my_inv, i, 8);
    x[i] = (i >= ci) ? sacc * dinv : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) Wd[g * 64 + i * 8 + ci] = x[i];
  // DMMA-friendly copy of L_kk (zeros above the diagonal)
  for (int e = lane; e < 1024; e += 32) {
    const int rr = e >> 5, cc = e & 31;
    Lt[rr * LS + cc] = cc <= rr ? D[rr * PS + cc] : 0.0;
  }
  __syncwarp();
  return fail;
}

// stage rows [c0, c0+32) x the columns of panels [j0, j0+kc) of L into shared memory with
// cp.async (16-byte LDGSTS, L2 only): every copy of a thread is in flight before the first wait,
// so the stage costs about one L2 round trip instead of one per loop iteration
__device__ __forceinline__ void stage_block_row(const double* slab, const SlabGeom& G, double* Bs, int bstride,
                                                int c0, int j0, int kc, int tid, int nthreads) {
  for (int idx = tid; idx < 512 * kc; idx += nthreads) {
    const int c2 = idx & 15, row = (idx >> 4) & 31, jj = idx >> 9;
    const double* src = slab + G.off(j0 + jj) + (size_t)(c0 + row - 32 * (j0 + jj)) * 32 + 2 * c2;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(Bs + (size_t)row * bstride + 32 * jj + 2 * c2);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

struct TileSet {
  int rb[4];     // first storage row of each 8-row tile
  int kind[4];   // 0 training rows, 1 the y row, 2 identity rows, 3 unused slot
  int js[4];     // first previous panel with non-zero entries (identity rows only)
};

// acc[t][u] += A_t (8 x 32*kc, streamed from the slab) * B_u^T (shared memory) for the NTL tiles
// of this warp.  NTL is a template parameter on purpose: predicated-off DMMAs still occupy the
// FP64 pipe, so inactive tiles must not appear in the instruction stream at all.
template <int NTL>
__device__ __forceinline__ void k_chunk(double (&acc)[4][4][2], const TileSet& TS, const double* slab,
                                        const SlabGeom& G, const double* Bs, int bstride, int j0, int kc,
                                        int r, int q) {
  // L2 round trips cost ~1000 cycles here while one 8-column step is only NTL x 128 cycles of DMMA
  // issue, so the A operand runs PF steps ahead in a register ring (statically indexed)
  constexpr int PF = NTL == 1 ? 8 : (NTL == 2 ? 4 : 3);
  int jmin = j0 + kc;
#pragma unroll
  for (int t = 0; t < NTL; ++t) jmin = min(jmin, TS.js[t]);
  const int jbeg = max(j0, jmin), jend = j0 + kc;
  if (jbeg >= jend) return;
  const int steps = 4 * (jend - jbeg);
  double2 ring[PF][NTL];
#pragma unroll
  for (int s = 0; s < PF; ++s) {
    const int jj = jbeg + (s >> 2);
#pragma unroll
    for (int t = 0; t < NTL; ++t) {
      ring[s][t] = make_double2(0.0, 0.0);
      if (s < steps && jj >= TS.js[t])
        ring[s][t] = *reinterpret_cast<const double2*>(
            slab + G.off(jj) + (size_t)(TS.rb[t] + r - 32 * jj) * 32 + 8 * (s & 3) + 2 * q);
    }
  }
  for (int st0 = 0; st0 < steps; st0 += PF) {
#pragma unroll
    for (int s = 0; s < PF; ++s) {
      const int st = st0 + s;
      if (st < steps) {
        const int jcur = jbeg + (st >> 2);
        double2 av[NTL];
#pragma unroll
        for (int t = 0; t < NTL; ++t) av[t] = ring[s][t];
        const int sn = st + PF, jn = jbeg + (sn >> 2);
        if (sn < steps) {
#pragma unroll
          for (int t = 0; t < NTL; ++t)
            ring[s][t] = (jn >= TS.js[t]) ? *reinterpret_cast<const double2*>(
                                                slab + G.off(jn) + (size_t)(TS.rb[t] + r - 32 * jn) * 32 +
                                                8 * (sn & 3) + 2 * q)
                                          : make_double2(0.0, 0.0);
        }
        const int bcol = 32 * (jcur - j0) + 8 * (st & 3) + 2 * q;
        double2 bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          bv[u] = *reinterpret_cast<const double2*>(Bs + (size_t)(8 * u + r) * bstride + bcol);
#pragma unroll
        for (int t = 0; t < NTL; ++t) {
#pragma unroll
          for (int u = 0; u < 4; ++u) dmma(acc[t][u], av[t].x, bv[u].x);
#pragma unroll
          for (int u = 0; u < 4; ++u) dmma(acc[t][u], av[t].y, bv[u].y);
        }
      }
    }
  }
}

// C = init - acc;  X = C * L_kk^-T (block forward substitution on DMMA);  store X into panel k.
template <int NTL>
__device__ __forceinline__ void finish_tiles(double (&acc)[4][4][2], const TileSet& TS, const CholArgs& A,
                                             const double* Lt, const double* Wd, double* slab,
                                             const SlabGeom& G, int k, int n, int npad, int b, int r, int q,
                                             double& zz) {
  const int c0 = 32 * k;
  // Gram values of my tiles for this panel: all loads issued together, branch-free
  double2 ginit[NTL][4];
  if (!A.dense) {
#pragma unroll
    for (int t = 0; t < NTL; ++t)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rowc = TS.kind[t] == 0 ? min(TS.rb[t] + r, npad - 1) : c0;
        ginit[t][u] = *reinterpret_cast<const double2*>(slab + G.off(k) + (size_t)(rowc - c0) * 32 + 8 * u + 2 * q);
      }
  }
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

// Dblk = init(diagonal block kk) - sum of the four partial products (when kk > 0)
__device__ __forceinline__ void assemble_diag(const CholArgs& A, const double* slab, const SlabGeom& G,
                                              const double (*Part)[32 * PP], double* Dblk, int kk, int n,
                                              bool with_part, int t0, int nthreads) {
  const int c0 = 32 * kk;
  // all loads of a thread are issued before the first use (clamped, branch-free addresses)
  for (int e0 = t0; e0 < 1024; e0 += 8 * nthreads) {
    double g[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int e = e0 + s * nthreads;
      const int rl = (e >> 5) & 31, cl = e & 31;
      const int row = min(c0 + rl, n - 1), col = min(c0 + cl, n - 1);
      g[s] = (e < 1024) ? (A.dense ? A.dense[(size_t)row * A.ldd + col]
                                   : __ldcg(slab + G.off(kk) + (size_t)(row - c0) * 32 + (col - c0)))
                        : 0.0;
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int e = e0 + s * nthreads;
      if (e >= 1024) continue;
      const int rl = e >> 5, cl = e & 31;
      double v = 0.0;
      if (cl <= rl) {
        const int row = c0 + rl, col = c0 + cl;
        if (row < n && col < n) {
          v = g[s] + ((A.dense && row == col) ? A.jitter : 0.0);
          if (with_part) v -= (Part[0][rl * PP + cl] + Part[1][rl * PP + cl]) +
                              (Part[2][rl * PP + cl] + Part[3][rl * PP + cl]);
        } else {
          v = (row == col) ? 1.0 : 0.0;
        }
      }
      Dblk[rl * PS + cl] = v;
    }
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
constexpr int MAXT = 3;   // row tiles per warp and round

template <int NW, int CS>
__global__ void __launch_bounds__(NW * 32, 1) chol_lml_kernel(CholArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CholSmem<NW>& S = *reinterpret_cast<CholSmem<NW>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int crank = 0;
  if (CS > 1) asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int r = lane >> 2, q = lane & 3;
  const int n = A.n, d = A.d;
  const SlabGeom G = SlabGeom::make(n, A.aug != 0);
  const int P = G.P, npad = 32 * P;
  const int kch = P - 1 < KCH ? (P - 1 > 0 ? P - 1 : 1) : KCH;
  const int bstride = 32 * kch + 8;
  double* Bs = reinterpret_cast<double*>(smem_raw + ((sizeof(CholSmem<NW>) + 15) & ~size_t(15)));  // 32 x bstride
  if (A.prog) {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&S.prog);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += NW * 32) dst[i] = src[i];
  } else if (tid == 0) {
    S.prog.n_ops = 0; S.prog.n_theta = 0; S.prog.n_leaves = 0; S.prog.d = 0; S.prog.fast_kind = 0;
  }
  __syncthreads();
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
      double acc[4][4][2];
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
      const int nchunks = (k + kch - 1) / kch;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int j0 = ch * kch, kc = min(kch, k - j0);
        __syncthreads();
        stage_block_row(slab, G, Bs, bstride, c0, j0, kc, tid, NW * 32);
        __syncthreads();
        if (warp < 4) {
          for (int c8 = warp; c8 < 4 * kc; c8 += 4) {
            double2 f[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
              f[t] = *reinterpret_cast<const double2*>(Bs + (size_t)(8 * t + r) * bstride + 8 * c8 + 2 * q);
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                dmma(acc[t][u], f[t].x, f[u].x);
                dmma(acc[t][u], f[t].y, f[u].y);
              }
          }
        }
      }
      if (warp < 4) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            S.Part[warp][(8 * t + r) * PP + 8 * u + 2 * q] = acc[t][u][0];
            S.Part[warp][(8 * t + r) * PP + 8 * u + 2 * q + 1] = acc[t][u][1];
          }
      }
      __syncthreads();   // also makes the Gram panel (written earlier) visible to the whole CTA
      BGP_STAMP(1);
      assemble_diag(A, slab, G, S.Part, S.Dblk, k, n, k > 0, tid, NW * 32);
      __syncthreads();
      BGP_STAMP(2);
      // Warp 0 factors the diagonal block.  When K is not chunked the other warps do not wait for
      // it: the trailing-update GEMM of their first round only needs previous panels, so they run
      // it now and meet warp 0 at named barrier 2 right before the panel solve.
      const bool overlap = nchunks <= 1 && NW > 1;
      if (warp == 0) {
        int f = warp_potrf32(S.Dblk, S.Lt, &S.Wd[0][0], lane, logdet, min(32, n - c0));
        if (f && lane == 0) S.fail = c0 + f;
        // L_kk -> slab (diag group rows of panel k)
        if (crank == 0)
          for (int e = lane; e < 1024; e += 32) slab[G.off(k) + e] = S.Lt[(e >> 5) * LS + (e & 31)];
        BGP_STAMP(3);
        if (overlap) asm volatile("bar.sync 2, %0;" ::"r"(NW * 32) : "memory");
      } else if (A.dbg && warp == (A.dbg_tid >> 5)) {
        BGP_STAMP(3);
      }
      if (!overlap) {
        __syncthreads();
        if (S.fail) break;
      }
      BGP_STAMP(4);

      // ------------------------------------------------ phase 2: rows below the block, 8-row
      // tiles dealt evenly to the warps (up to four per warp and round)
      // 32-row groups below the diagonal and identity-row groups; with CS == 2 groups alternate
      // between the two CTAs of the cluster and the y tile belongs to rank 0
      const int nmg = P - 1 - k, nag = A.aug ? k + 1 : 0;
      const int my_mg = (nmg - crank + CS - 1) / CS;                 // groups crank, crank + CS, ...
      const int a0 = ((crank - nmg) % CS + CS) % CS;                  // continue the deal over identity groups
      const int my_ag = nag > a0 ? (nag - a0 + CS - 1) / CS : 0;
      const int has_z = (CS == 1 || crank == 0) ? 1 : 0;
      const int n_main_t = 4 * my_mg;
      const int T = n_main_t + has_z + 4 * my_ag;
      // warp w owns the contiguous tiles [w0, w0 + mine); every warp runs the same number of
      // rounds (barriers inside when K is chunked), each with at most four of its tiles
      const int nwk = overlap ? NW - 1 : NW;          // worker warps (warp 0 is busy when overlapping)
      const int wrk = overlap ? warp - 1 : warp;
      const int tq = T / nwk, trm = T % nwk;
      const int mine = wrk < 0 ? 0 : tq + (wrk < trm ? 1 : 0);
      const int w0 = wrk < 0 ? 0 : wrk * tq + min(wrk, trm);
      const int rounds = max(1, (tq + (trm ? 1 : 0) + MAXT - 1) / MAXT);
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
            if (ti < n_main_t) {
              const int g = CS * (ti >> 2) + crank;
              TS.rb[t] = 32 * (k + 1) + 32 * g + 8 * (ti & 3); TS.kind[t] = 0;
            } else if (has_z && ti == n_main_t) {
              TS.rb[t] = G.Rz; TS.kind[t] = 1;
            } else {
              const int l2 = ti - n_main_t - has_z;
              const int a = a0 + CS * (l2 >> 2);
              TS.rb[t] = G.Ra + 32 * a + 8 * (l2 & 3); TS.kind[t] = 2; TS.js[t] = a;
            }
          }
        }
        long long tph = clock64();
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
            case 3: k_chunk<3>(acc, TS, slab, G, Bs, bstride, j0, kc, r, q); break;
            case 2: k_chunk<2>(acc, TS, slab, G, Bs, bstride, j0, kc, r, q); break;
            case 1: k_chunk<1>(acc, TS, slab, G, Bs, bstride, j0, kc, r, q); break;
            default: break;
          }
        }
        BGP_STAMP_ADD(7, tph); tph = clock64();
        if (overlap && rd == 0) asm volatile("bar.sync 2, %0;" ::"r"(NW * 32) : "memory");
        if (S.fail) continue;
        double zpart = 0.0;
        switch (ntl) {
          case 3: finish_tiles<3>(acc, TS, A, S.Lt, &S.Wd[0][0], slab, G, k, n, npad, b, r, q, zpart); break;
          case 2: finish_tiles<2>(acc, TS, A, S.Lt, &S.Wd[0][0], slab, G, k, n, npad, b, r, q, zpart); break;
          case 1: finish_tiles<1>(acc, TS, A, S.Lt, &S.Wd[0][0], slab, G, k, n, npad, b, r, q, zpart); break;
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
  return base + sizeof(double) * (size_t)32 * (32 * kch + 8);
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

// opt-in to the large dynamic shared-memory carve-out (must happen outside stream capture)
cudaError_t prepare_chol(int n) {
  const size_t smem = chol_smem_bytes(n);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  if (pick_nw(n) == 4)
    return cudaFuncSetAttribute(chol_lml_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaError_t e = cudaFuncSetAttribute(chol_lml_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(chol_lml_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(chol_lml_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(chol_lml_kernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
