// K1+K2: batched-theta Gram build + FP64 Cholesky + forward solve + LML + log-prior,
// one CTA per theta.  Left-looking blocked factorisation, 32-column panels:
//   * the Gram panel is generated from X by the CTA itself right before it is factored
//     (fused K1) into the L2-resident factor slab -- the n x n matrix never exists in HBM as
//     an input;
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
constexpr int WS = 40;         // row stride of the 32x32 inverse block (== 8 mod 16)
constexpr int PS = 33;         // row stride of the diagonal block (odd: conflict-free columns)
constexpr int PP = 34;         // row stride of the K-split partial blocks

template <int NW>
struct CholSmem {
  DevProgram prog;
  ThetaParams tp;
  double Dblk[32 * PS];              // diagonal block being factored
  double Ws[32 * WS];                // its inverse
  double Part[4][32 * PP];           // K-split partial sums of the diagonal update
  double red[NW];
  double inv_diag[32];
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

// ---- warp-level Cholesky + inverse of the 32x32 diagonal block in shared memory ----------
// Compact (rolled) code on purpose: it runs once per panel on one warp, so straight-line
// unrolled code would be instruction-fetch bound.  Returns 0 or failing local column + 1
// (LAPACK dpotrf: pivot <= 0 or NaN).
__device__ __noinline__ int warp_potrf_inv32(double* D, double* Ws, double* inv_diag, int lane,
                                             double& logdet, int ncols_real) {
  int fail = 0;
  for (int j = 0; j < 32; ++j) {
    double ajj = D[j * PS + j];
    if (!(ajj > 0.0)) { if (!fail) fail = j + 1; ajj = 1.0; }
    const double ljj = sqrt(ajj);
    const double inv = 1.0 / ljj;
    const double v = D[lane * PS + j];
    const double lij = lane > j ? v * inv : (lane == j ? ljj : 0.0);
    __syncwarp();
    D[lane * PS + j] = lij;
    if (lane == j) inv_diag[j] = inv;
    __syncwarp();
#pragma unroll 4
    for (int k = j + 1; k < 32; ++k) {
      const double lkj = D[k * PS + j];
      if (lane >= k) D[lane * PS + k] = fma(-lij, lkj, D[lane * PS + k]);
    }
    __syncwarp();
  }
  if (lane < ncols_real) logdet += log(D[lane * PS + lane]);
  // W = L^-1: lane c solves column c by forward substitution (4 partial sums break the chain)
  for (int i = 0; i < 32; ++i) {
    double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int k = 0;
    for (; k + 4 <= i; k += 4) {
      s0 = fma(-D[i * PS + k], Ws[k * WS + lane], s0);
      s1 = fma(-D[i * PS + k + 1], Ws[(k + 1) * WS + lane], s1);
      s2 = fma(-D[i * PS + k + 2], Ws[(k + 2) * WS + lane], s2);
      s3 = fma(-D[i * PS + k + 3], Ws[(k + 3) * WS + lane], s3);
    }
    for (; k < i; ++k) s0 = fma(-D[i * PS + k], Ws[k * WS + lane], s0);
    Ws[i * WS + lane] = (i >= lane) ? ((s0 + s1) + (s2 + s3)) * inv_diag[i] : 0.0;
  }
  __syncwarp();
  return fail;
}

// Gram entries of panel k (rows [32k, n) x 32 columns, lower part) -> slab.  Each warp takes
// four rows at a time, lanes are the 32 columns; scaled coordinates come from the transposed
// shared-memory copy Xt[leaf][dim][npad] (conflict-free for lanes, broadcast for rows).
template <int NW>
__device__ __forceinline__ void gram_panel(const DevProgram& PR, const ThetaParams& TP, const double* Xt,
                                           int npad, const double* __restrict__ alpha, double* slab,
                                           const SlabGeom& G, int k, int n, int d, int warp, int lane) {
  const int c0 = 32 * k, col = c0 + lane;
  double* base = slab + G.off(k);
  for (int r0 = c0 + 4 * warp; r0 < n; r0 += 4 * NW) {
    double r2[4][BGP_MAX_LEAVES];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int l = 0; l < BGP_MAX_LEAVES; ++l) r2[a][l] = 0.0;
#pragma unroll
    for (int l = 0; l < BGP_MAX_LEAVES; ++l) {
      if (l < PR.n_leaves) {
        const double* xl = Xt + (size_t)l * d * npad;
        for (int kk = 0; kk < d; ++kk) {
          const double* xr = xl + (size_t)kk * npad;
          const double xc = xr[col];
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const double t = xr[min(r0 + a, npad - 1)] - xc;
            r2[a][l] = fma(t, t, r2[a][l]);
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = r0 + a;
      if (row < n && col <= row) {
        const bool same = row == col;
        if (same) {
#pragma unroll
          for (int l = 0; l < BGP_MAX_LEAVES; ++l) r2[a][l] = 0.0;
        }
        double v = eval_program(PR, TP, r2[a], same, true);
        if (same) v += alpha[row];
        base[(size_t)(row - c0) * 32 + lane] = v;
      }
    }
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) chol_lml_kernel(CholArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CholSmem<NW>& S = *reinterpret_cast<CholSmem<NW>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int n = A.n, d = A.d;
  const SlabGeom G = SlabGeom::make(n, A.aug != 0);
  const int P = G.P, npad = 32 * P;
  const int kch = P - 1 < KCH ? (P - 1 > 0 ? P - 1 : 1) : KCH;
  const int bstride = 32 * kch + 8;
  double* Bs = reinterpret_cast<double*>(smem_raw + ((sizeof(CholSmem<NW>) + 15) & ~size_t(15)));  // 32 x bstride
  // [leaf][dim][npad] scaled training inputs (Gram mode): shared memory when it fits, else a
  // per-CTA global scratch (same layout, L1/L2 cached)
  double* Xt = A.xt_scratch ? A.xt_scratch + (size_t)blockIdx.x * A.xt_stride
                            : Bs + (size_t)32 * bstride;
  if (A.prog) {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&S.prog);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += NW * 32) dst[i] = src[i];
  } else if (tid == 0) {
    S.prog.n_ops = 0; S.prog.n_theta = 0; S.prog.n_leaves = 0; S.prog.d = 0;
  }
  __syncthreads();
  const DevProgram& PR = S.prog;

  for (int b = blockIdx.x; b < A.batch; b += gridDim.x) {
    const double* theta = A.theta + (size_t)b * PR.n_theta;
    double* slab = A.slabs + (size_t)(A.slab_per_block ? blockIdx.x : b) * G.doubles();
    if (!A.dense) resolve_theta(PR, theta, A.fixed_ls, S.tp, tid, NW * 32);
    if (tid == 0) S.fail = 0;
    double logdet = 0.0, zz = 0.0;   // meaningful in warp 0 / z-row owners
    __syncthreads();
    if (!A.dense) {
      for (int e = tid; e < PR.n_leaves * d * npad; e += NW * 32) {
        const int l = e / (d * npad), rem = e - l * d * npad, kk = rem / npad, i = rem - kk * npad;
        Xt[e] = (i < n) ? A.X[(size_t)i * d + kk] * S.tp.inv_ls[l][kk] : 0.0;
      }
      __syncthreads();
    }

    for (int k = 0; k < P; ++k) {
      const int c0 = 32 * k;
      if (!A.dense) gram_panel<NW>(PR, S.tp, Xt, npad, A.alpha, slab, G, k, n, d, warp, lane);
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
        // stage rows [c0, c0+32) x columns of panels [j0, j0+kc) of L
        for (int e = tid; e < 32 * kc * 16; e += NW * 32) {
          int row = e / (kc * 16), rem = e - row * kc * 16, jj = rem >> 4, c2 = rem & 15;
          const double2 v = *reinterpret_cast<const double2*>(
              slab + G.off(j0 + jj) + (size_t)(c0 + row - 32 * (j0 + jj)) * 32 + 2 * c2);
          *reinterpret_cast<double2*>(Bs + (size_t)row * bstride + 32 * jj + 2 * c2) = v;
        }
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
      __syncthreads();   // also makes the Gram panel written above visible to the whole CTA
      for (int e = tid; e < 1024; e += NW * 32) {
        const int rl = e >> 5, cl = e & 31;
        double v = 0.0;
        if (cl <= rl) {
          const int row = c0 + rl, col = c0 + cl;
          if (row < n && col < n) {
            v = A.dense ? A.dense[(size_t)row * A.ldd + col] + (row == col ? A.jitter : 0.0)
                        : slab[G.off(k) + (size_t)rl * 32 + cl];
            if (k > 0) v -= (S.Part[0][rl * PP + cl] + S.Part[1][rl * PP + cl]) +
                            (S.Part[2][rl * PP + cl] + S.Part[3][rl * PP + cl]);
          } else {
            v = (row == col) ? 1.0 : 0.0;
          }
        }
        S.Dblk[rl * PS + cl] = v;
      }
      __syncthreads();
      if (warp == 0) {
        int f = warp_potrf_inv32(S.Dblk, S.Ws, S.inv_diag, lane, logdet, min(32, n - c0));
        if (f && lane == 0) S.fail = c0 + f;
        // L_kk -> slab (diag group rows of panel k)
        for (int e = lane; e < 1024; e += 32)
          slab[G.off(k) + e] = S.Dblk[(e >> 5) * PS + (e & 31)];
      }
      __syncthreads();
      if (S.fail) break;

      // ------------------------------------------------ phase 2: rows below the block
      const int n_main = P - 1 - k;
      const int n_groups = n_main + 1 + (A.aug ? k + 1 : 0);
      const int rounds = (n_groups + NW - 1) / NW;
      for (int rd = 0; rd < rounds; ++rd) {
        const int gi = rd * NW + warp;
        const bool valid = gi < n_groups;
        int rb = 0, kind = 0, jstart = 0;   // kind 0 main, 1 z, 2 identity rows
        if (valid) {
          if (gi < n_main) { rb = 32 * (k + 1 + gi); kind = 0; }
          else if (gi == n_main) { rb = G.Rz; kind = 1; }
          else { int a = gi - n_main - 1; rb = G.Ra + 32 * a; kind = 2; jstart = a; }
        }
        const int mt = (kind == 1) ? 1 : 4;
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
        for (int ch = 0; ch < nchunks; ++ch) {
          const int j0 = ch * kch, kc = min(kch, k - j0);
          if (nchunks > 1) {
            __syncthreads();
            for (int e = tid; e < 32 * kc * 16; e += NW * 32) {
              int row = e / (kc * 16), rem = e - row * kc * 16, jj = rem >> 4, c2 = rem & 15;
              const double2 v = *reinterpret_cast<const double2*>(
                  slab + G.off(j0 + jj) + (size_t)(c0 + row - 32 * (j0 + jj)) * 32 + 2 * c2);
              *reinterpret_cast<double2*>(Bs + (size_t)row * bstride + 32 * jj + 2 * c2) = v;
            }
            __syncthreads();
          }
          if (!valid) continue;
          const int jbeg = max(j0, jstart), jend = j0 + kc;
          if (jbeg >= jend) continue;
          // stream A (my 32 rows of the previous panels) from the slab, one 8-column step ahead
          const int steps = 4 * (jend - jbeg);
          const double* ap[4];
          double2 nxt[4];
          int j = jbeg, c8p = 0;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            ap[t] = slab + G.off(j) + (size_t)(rb + 8 * t + r - 32 * j) * 32 + 2 * q;
            nxt[t] = (t < mt) ? *reinterpret_cast<const double2*>(ap[t]) : make_double2(0, 0);
          }
          for (int st = 0; st < steps; ++st) {
            double2 av[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) av[t] = nxt[t];
            const int bcol = 32 * (j - j0) + 8 * c8p + 2 * q;
            // advance + prefetch
            if (++c8p == 4) {
              c8p = 0; ++j;
#pragma unroll
              for (int t = 0; t < 4; ++t)
                ap[t] = slab + G.off(j) + (size_t)(rb + 8 * t + r - 32 * j) * 32 + 2 * q;
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t) ap[t] += 8;
            }
            if (st + 1 < steps) {
#pragma unroll
              for (int t = 0; t < 4; ++t)
                if (t < mt) nxt[t] = *reinterpret_cast<const double2*>(ap[t]);
            }
            double2 bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              bv[u] = *reinterpret_cast<const double2*>(Bs + (size_t)(8 * u + r) * bstride + bcol);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (t < mt) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  dmma(acc[t][u], av[t].x, bv[u].x);
                  dmma(acc[t][u], av[t].y, bv[u].y);
                }
              }
            }
          }
        }
        if (!valid) continue;
        // C = init - acc, in accumulator layout (init: Gram panel in the slab / y / identity)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (t >= mt) continue;
          const int row = rb + 8 * t + r;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int cl = 8 * u + 2 * q, col = c0 + cl;
            double2 v0 = make_double2(0.0, 0.0);
            if (kind == 0) {
              if (row < n) {
                if (A.dense) {
                  if (col < n) v0.x = A.dense[(size_t)row * A.ldd + col];
                  if (col + 1 < n) v0.y = A.dense[(size_t)row * A.ldd + col + 1];
                } else {
                  v0 = *reinterpret_cast<const double2*>(slab + G.off(k) + (size_t)(row - c0) * 32 + cl);
                  if (col >= n) v0.x = 0.0;
                  if (col + 1 >= n) v0.y = 0.0;
                }
              }
            } else if (kind == 1) {
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
        // X = C * W^T through DMMA, in place (descending output tile)
#pragma unroll
        for (int uo = 3; uo >= 0; --uo) {
          double o[4][2];
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t][0] = o[t][1] = 0.0;
#pragma unroll
          for (int ui = 0; ui <= uo; ++ui) {
            const double2 wv = *reinterpret_cast<const double2*>(S.Ws + (8 * uo + r) * WS + 8 * ui + 2 * q);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (t < mt) {
                dmma(o[t], acc[t][ui][0], wv.x);
                dmma(o[t], acc[t][ui][1], wv.y);
              }
            }
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) { acc[t][uo][0] = o[t][0]; acc[t][uo][1] = o[t][1]; }
        }
        // store the finished rows of panel k
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (t >= mt) continue;
          const int row = rb + 8 * t + r;
          double* dst = slab + G.off(k) + (size_t)(row - c0) * 32 + 2 * q;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<double2*>(dst + 8 * u) = make_double2(acc[t][u][0], acc[t][u][1]);
        }
        if (kind == 1 && r == 0) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            zz = fma(acc[0][u][0], acc[0][u][0], zz);
            zz = fma(acc[0][u][1], acc[0][u][1], zz);
            if (A.z_out) {
              int col = c0 + 8 * u + 2 * q;
              if (col < n) A.z_out[(size_t)b * n + col] = acc[0][u][0];
              if (col + 1 < n) A.z_out[(size_t)b * n + col + 1] = acc[0][u][1];
            }
          }
        }
      }
      __syncthreads();
    }

    // ------------------------------------------------------------------ epilogue
    zz = warp_sum(zz);
    if (warp == 0) logdet = warp_sum(logdet);
    if (lane == 0) S.red[warp] = zz;
    __syncthreads();
    if (tid == 0) {
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

static int pick_nw(int n) { return n <= 64 ? 4 : 16; }

static size_t smem_base(int n) {
  const int P = (n + 31) / 32;
  const int kch = P - 1 < KCH ? (P - 1 > 0 ? P - 1 : 1) : KCH;
  size_t base = pick_nw(n) == 16 ? sizeof(CholSmem<16>) : sizeof(CholSmem<4>);
  base = (base + 15) & ~size_t(15);
  return base + sizeof(double) * (size_t)32 * (32 * kch + 8);
}

// doubles of per-CTA global scratch the scaled inputs need when they do not fit in shared memory
size_t chol_xt_doubles(int n, int d, int n_leaves) {
  return (size_t)(n_leaves > 0 ? n_leaves : 1) * d * (32 * ((n + 31) / 32));
}
size_t chol_xt_scratch_doubles(int n, int d, int n_leaves) {
  const size_t xt = chol_xt_doubles(n, d, n_leaves);
  return smem_base(n) + sizeof(double) * xt <= 226 * 1024 ? 0 : xt;
}

static size_t chol_smem_bytes(int n, int d, int n_leaves, bool dense) {
  if (dense || chol_xt_scratch_doubles(n, d, n_leaves)) return smem_base(n);
  return smem_base(n) + sizeof(double) * chol_xt_doubles(n, d, n_leaves);
}

// opt-in to the large dynamic shared-memory carve-out (must happen outside stream capture)
cudaError_t prepare_chol(int n, int d, int n_leaves, bool dense) {
  const size_t smem = chol_smem_bytes(n, d, n_leaves, dense);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  return pick_nw(n) == 4
             ? cudaFuncSetAttribute(chol_lml_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
             : cudaFuncSetAttribute(chol_lml_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t launch_chol(const CholArgs& A, int grid, int n_leaves, cudaStream_t stream) {
  const size_t smem = chol_smem_bytes(A.n, A.d, n_leaves, A.dense != nullptr);
  if (pick_nw(A.n) == 4) chol_lml_kernel<4><<<grid, 128, smem, stream>>>(A);
  else chol_lml_kernel<16><<<grid, 512, smem, stream>>>(A);
  return cudaGetLastError();
}

}  // namespace bgp
