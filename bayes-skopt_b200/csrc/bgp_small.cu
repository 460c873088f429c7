// K1+K2 for small n, fused: one CTA per theta, the whole matrix resident in shared memory.
//   Gram (ARD scaled differences, Matern/RBF program, +alpha on the diagonal) is built straight into
//   8x8 tiles of the lower triangle, factored right-looking over 8-column blocks, with y riding along
//   as one extra tile row so that z = L^-1 y falls out of the same panel solves; LML + typed
//   log-priors in the epilogue.  One launch, nothing n x n ever leaves the SM.
// Per 8-column block:
//   1. every warp reads the diagonal tile and factors it redundantly in registers (no exchange, no
//      barrier between "factor" and "use");
//   2. panel solve X = C L_bb^-T by forward substitution, one lane per tile row (four tiles per warp
//      at a time) -- the dependent chain is 8 short steps, cheaper than forming an inverse;
//   3. trailing update C_ij -= X_i X_j^T on DMMA.8x8x4, tiles stored in fragment order so that both
//      operands are one conflict-free 8-byte load per lane.
// Shared memory bounds n: (T(T+1)/2 + T) tiles of 512 B with T = ceil(n/8), i.e. n <= 224 at d = 6;
// several CTAs share an SM when n is small (n = 100: four).  Larger n take the blocked L2-slab
// path of bgp_chol.cu.
// Replaces sklearn:_gpr.py:583-617 + bask/bayesgpr.py:351-379 for the configurations that live in
// this regime (BASELINE configs[0] n=20 and configs[1] n=100).
#include <cstdlib>

#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

struct SmallSmem {
  DevProgram prog;
  ThetaParams tp;
  double red[8];
  double lprior;
  int fail;
  int pad_[1];
};

// position of element (r, c) inside a tile: halves of four columns, [half][row][col & 3] -- the
// DMMA A/B fragment order (lane l = 4 r + q reads positions l and 32 + l)
__device__ __forceinline__ int tpos(int r, int c) { return (c >> 2) * 32 + r * 4 + (c & 3); }
__device__ __forceinline__ int tidx(int i, int j) { return i * (i + 1) / 2 + j; }


// advance a (tile row, tile column) position of the lower triangle by `step` tiles in row-major order
__device__ __forceinline__ void tri_advance(int& i, int& j, int step) {
  j += step;
  while (j > i) { j -= i + 1; ++i; }
}

// Gram tiles of the lower triangle, fast path  c * k(r) + white  with the stationary kind known at compile
// time: thread = (position inside a tile, tile slot); four tiles per thread and iteration, i.e. four
// independent sqrt/exp chains in flight -- at <= 8 warps per SM latency, not throughput, bounds this phase
template <int KIND, int NT>
__device__ __forceinline__ void gram_fast(double* tiles, const double* Xs, const double* alpha, double cval,
                                          double wval, int T, int np, int n, int d, int tid) {
  constexpr int STEP = NT / 64;
  const int pos = tid & 63, sub = tid >> 6;
  const int rr = (pos & 31) >> 2, cc = ((pos >> 5) << 2) + (pos & 3);
  const int nlow = T * (T + 1) / 2;
  int i = 0, j = 0;
  tri_advance(i, j, sub);
  for (int t = sub; t < nlow; t += 4 * STEP) {
    int row[4], col[4], rowc[4], colc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      row[u] = 8 * i + rr; col[u] = 8 * j + cc;
      rowc[u] = min(row[u], n - 1); colc[u] = min(col[u], n - 1);
      tri_advance(i, j, STEP);
    }
    double r2[4] = {0.0, 0.0, 0.0, 0.0};
    for (int kk = 0; kk < d; ++kk) {
      const double* xk = Xs + kk * np;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double t2 = xk[rowc[u]] - xk[colc[u]];
        r2[u] = fma(t2, t2, r2[u]);
      }
    }
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = cval * stationary_value(KIND, r2[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double dg = (cval + wval) + alpha[rowc[u]];            // k(x, x) = c exactly, then + white, + alpha
      const bool pad = row[u] >= n || col[u] >= n;
      v[u] = pad ? (row[u] == col[u] ? 1.0 : 0.0) : (col[u] > row[u] ? 0.0 : (row[u] == col[u] ? dg : v[u]));
      if (t + u * STEP < nlow) tiles[(size_t)(t + u * STEP) * 64 + pos] = v[u];
    }
  }
}

// the same for an arbitrary covariance program (interpreter, one tile per thread and iteration)
template <int NT>
__device__ __forceinline__ void gram_general(double* tiles, const double* Xs, const double* alpha,
                                             const DevProgram& PR, const ThetaParams& tp, int T, int np, int n,
                                             int d, int tid) {
  constexpr int STEP = NT / 64;
  const int pos = tid & 63, sub = tid >> 6;
  const int rr = (pos & 31) >> 2, cc = ((pos >> 5) << 2) + (pos & 3);
  const int nlow = T * (T + 1) / 2;
  int i = 0, j = 0;
  tri_advance(i, j, sub);
  for (int t = sub; t < nlow; t += STEP) {
    const int row = 8 * i + rr, col = 8 * j + cc;
    double v;
    if (row >= n || col >= n) {
      v = (row == col) ? 1.0 : 0.0;                  // identity padding
    } else if (col > row) {
      v = 0.0;                                        // strictly upper part of a diagonal tile: never read
    } else {
      double r2[BGP_MAX_LEAVES];
#pragma unroll
      for (int l = 0; l < BGP_MAX_LEAVES; ++l) {
        r2[l] = 0.0;
        if (l < PR.n_leaves && row != col) {
          const double* xl = Xs + (size_t)l * d * np;
          double acc = 0.0;
          for (int kk = 0; kk < d; ++kk) {
            const double t2 = xl[kk * np + row] - xl[kk * np + col];
            acc = fma(t2, t2, acc);
          }
          r2[l] = acc;
        }
      }
      v = eval_program(PR, tp, r2, row == col, true);
      if (row == col) v += alpha[row];
    }
    tiles[(size_t)t * 64 + pos] = v;
    tri_advance(i, j, STEP);
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) small_lml_kernel(CholArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmallSmem& S = *reinterpret_cast<SmallSmem*>(smem_raw);
  constexpr int NT = NW * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int n = A.n, d = A.d, T = (n + 7) / 8, np = 8 * T;
  double* tiles = reinterpret_cast<double*>(smem_raw + ((sizeof(SmallSmem) + 127) & ~size_t(127)));
  const int ntiles = T * (T + 1) / 2 + T;           // lower triangle + the y row (tile row T)
  double* Xs = tiles + (size_t)ntiles * 64;          // [leaf][dim][np] scaled (and warped) inputs
  {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&S.prog);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += NT) dst[i] = src[i];
  }
  __syncthreads();
  const DevProgram& PR = S.prog;

  for (int b = blockIdx.x; b < A.batch; b += gridDim.x) {
    const double* theta = A.theta + (size_t)b * PR.n_theta;
#define SMALL_STAMP(i) do { if (A.dbg && blockIdx.x == 0 && tid == A.dbg_tid) A.dbg[i] = clock64(); } while (0)
    SMALL_STAMP(0);
    __syncthreads();                                  // previous theta's epilogue is done with S
    resolve_theta(PR, theta, A.fixed_ls, S.tp, tid, NT);
    if (tid == 0) S.fail = 0;
    if (warp == NW - 1) {
      // typed log-priors, one per lane (exp / log / lgamma chains of ~1k cycles each, so not in the
      // single-thread epilogue); summed in table order by a fixed shuffle tree
      double lpv = 0.0;
      if (A.priors)
        for (int k = lane; k < A.n_priors; k += 32) lpv += log_prior(A.priors + k, 1, theta + k);
      lpv = warp_sum(lpv);
      if (lane == 0) S.lprior = lpv;
    }
    __syncthreads();
    for (int e = tid; e < PR.n_leaves * d * np; e += NT) {
      const int l = e / (d * np), rem = e - l * d * np, kk = rem / np, i = rem - kk * np;
      Xs[e] = (i < n) ? bgp_warp_coord(PR, theta, kk, A.X[(size_t)i * d + kk]) * S.tp.inv_ls[l][kk] : 0.0;
    }
    __syncthreads();
    SMALL_STAMP(1);
    // ---- Gram: lower-triangle tiles, then the y row (tile row T, y in its first row)
    {
      const double cval = S.tp.opval[PR.fast_const], wval = S.tp.opval[PR.fast_white];
      switch (PR.fast_kind) {
        case BGP_OP_MATERN52: gram_fast<BGP_OP_MATERN52, NT>(tiles, Xs, A.alpha, cval, wval, T, np, n, d, tid); break;
        case BGP_OP_MATERN32: gram_fast<BGP_OP_MATERN32, NT>(tiles, Xs, A.alpha, cval, wval, T, np, n, d, tid); break;
        case BGP_OP_MATERN12: gram_fast<BGP_OP_MATERN12, NT>(tiles, Xs, A.alpha, cval, wval, T, np, n, d, tid); break;
        case BGP_OP_RBF: gram_fast<BGP_OP_RBF, NT>(tiles, Xs, A.alpha, cval, wval, T, np, n, d, tid); break;
        default: gram_general<NT>(tiles, Xs, A.alpha, PR, S.tp, T, np, n, d, tid); break;
      }
      double* yrow = tiles + (size_t)tidx(T, 0) * 64;
      for (int e = tid; e < T * 64; e += NT) {
        const int pos = e & 63, rr = (pos & 31) >> 2, col = 8 * (e >> 6) + ((pos >> 5) << 2) + (pos & 3);
        yrow[e] = (rr == 0 && col < n) ? A.y[col] : 0.0;
      }
    }
    __syncthreads();
    SMALL_STAMP(2);

    // ---- factorisation
    double logdet = 0.0;                               // warp 0, lane 0
    int fail = 0;
    for (int kb = 0; kb < T; ++kb) {
#define KB_STAMP(i) do { if (A.dbg && blockIdx.x == 0 && tid == A.dbg_tid && (kb == 0 || kb == T / 2)) A.dbg[8 + (kb ? 8 : 0) + (i)] = clock64(); } while (0)
      KB_STAMP(0);
      // 1. diagonal tile, redundantly in every lane of the solver warps -- one warp per scheduler, so the
      // FP64 pipe serves one dependent rsqrt -> scale -> fma chain at a time
      constexpr int NSOLVE = NW < 4 ? NW : 4;
      if (warp < NSOLVE) {
        const double* Dt = tiles + (size_t)tidx(kb, kb) * 64;
        double a[8][8], inv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) a[i][j] = Dt[tpos(i, j)];
        double invprod = 1.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          double ajj = a[c][c];
          if (!(ajj > 0.0)) { if (!fail) fail = 8 * kb + c + 1; ajj = 1.0; }
          inv[c] = rsqrt(ajj);
#pragma unroll
          for (int i = c + 1; i < 8; ++i) a[i][c] *= inv[c];
#pragma unroll
          for (int j = c + 1; j < 8; ++j)
#pragma unroll
            for (int i = j; i < 8; ++i) a[i][j] = fma(-a[i][c], a[j][c], a[i][j]);
          if (8 * kb + c < n) invprod *= inv[c];
        }
        if (tid == 0) logdet -= log(invprod);
        KB_STAMP(1);
        // 2. panel solve: tile rows kb+1 .. T of tile column kb, one lane per tile row, 4 tiles per warp
        const int npanel = T - kb;
        for (int t0 = 4 * warp; t0 < npanel; t0 += 4 * NSOLVE) {
          const int t = t0 + (lane >> 3), rr = lane & 7;
          if (t < npanel) {
            double* Ct = tiles + (size_t)tidx(kb + 1 + t, kb) * 64;
            const double4 lo = *reinterpret_cast<const double4*>(Ct + rr * 4);
            const double4 hi = *reinterpret_cast<const double4*>(Ct + 32 + rr * 4);
            double x[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              double sacc = x[c];
#pragma unroll
              for (int k = 0; k < c; ++k) sacc = fma(-x[k], a[c][k], sacc);
              x[c] = sacc * inv[c];
            }
            *reinterpret_cast<double4*>(Ct + rr * 4) = make_double4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<double4*>(Ct + 32 + rr * 4) = make_double4(x[4], x[5], x[6], x[7]);
          }
        }
      }
      KB_STAMP(2);
      __syncthreads();
      KB_STAMP(3);
      // 3. trailing update: tiles (i, j), kb < j <= i (then the y row, j < T) as one flat sequence dealt
      // round-robin to the warps; the operands of a warp's next tile are loaded before the DMMAs of the
      // current one
      {
        const int M = T - kb - 1;                      // trailing tile rows 0..M-1, row M = the y row
        const int total = M * (M + 1) / 2 + M;
        const int cp = (q >> 1) * 32 + r * 4 + 2 * (q & 1);     // accumulator layout: [r][2q], [r][2q+1]
        int ip = 0, jp = 0;
        tri_advance(ip, jp, warp);                     // row M (the y row) is simply the next, shorter, row
        double* Ct = nullptr;
        double2 cv = make_double2(0.0, 0.0);
        double xi0 = 0.0, xi1 = 0.0, xj0 = 0.0, xj1 = 0.0;
        auto load = [&](int t) {
          if (t < total) {
            const int i = kb + 1 + ip, j = kb + 1 + jp;
            Ct = tiles + (size_t)tidx(i, j) * 64 + cp;
            const double* Xi = tiles + (size_t)tidx(i, kb) * 64;
            const double* Xj = tiles + (size_t)tidx(j, kb) * 64;
            cv = *reinterpret_cast<const double2*>(Ct);
            xi0 = Xi[lane]; xi1 = Xi[32 + lane];
            xj0 = Xj[lane]; xj1 = Xj[32 + lane];
            tri_advance(ip, jp, NW);
          }
        };
        load(warp);
        for (int t = warp; t < total; t += NW) {
          double* Cc = Ct;
          double c2[2] = {cv.x, cv.y};
          const double a0 = xi0, a1 = xi1, b0 = -xj0, b1 = -xj1;
          load(t + NW);
          dmma(c2, a0, b0);
          dmma(c2, a1, b1);
          *reinterpret_cast<double2*>(Cc) = make_double2(c2[0], c2[1]);
        }
      }
      KB_STAMP(4);
      __syncthreads();
      KB_STAMP(5);
#undef KB_STAMP
    }

    SMALL_STAMP(3);
    // ---- epilogue: |z|^2 from the first row of the y tiles, LML, log-prior
    double zz = 0.0;
    for (int e = tid; e < np; e += NT) {
      const double zv = tiles[(size_t)tidx(T, e >> 3) * 64 + tpos(0, e & 7)];
      zz = fma(zv, zv, zz);
      if (A.z_out && e < n) A.z_out[(size_t)b * n + e] = zv;
    }
    zz = warp_sum(zz);
    if (lane == 0) S.red[warp] = zz;
    if (fail && tid == 0) S.fail = fail;
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
        if (A.priors) lp += S.lprior;
        if (A.lp_extra) lp += A.lp_extra[b];
        if (!isfinite(lp)) lp = -INFINITY;
      }
      if (A.lml) A.lml[b] = lml;
      if (A.lp) A.lp[b] = lp;
      if (A.info) A.info[b] = S.fail;
    }
    SMALL_STAMP(4);
#undef SMALL_STAMP
  }
}

static int small_nw(int n) { return n <= 32 ? 2 : (n <= 96 ? 4 : 8); }

size_t small_smem_bytes(int n, int d, int n_leaves) {
  const int T = (n + 7) / 8;
  const size_t ntiles = (size_t)T * (T + 1) / 2 + T;
  const size_t base = (sizeof(SmallSmem) + 127) & ~size_t(127);
  return base + 8 * (ntiles * 64 + (size_t)(n_leaves > 0 ? n_leaves : 1) * d * 8 * T);
}

constexpr size_t SMALL_SMEM_OPTIN = 227 * 1024;

// BGP_NO_SMALL (read per call: the parity tests flip it) forces the blocked path at any n
bool small_path_fits(int n, int d, int n_leaves) {
  return std::getenv("BGP_NO_SMALL") == nullptr && small_smem_bytes(n, d, n_leaves) <= SMALL_SMEM_OPTIN;
}

cudaError_t prepare_small() {
  cudaError_t e = cudaFuncSetAttribute(small_lml_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_OPTIN);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(small_lml_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_OPTIN);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(small_lml_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_OPTIN);
  return e;
}

// grid: one CTA per theta up to the number of CTAs the chip holds at this shared-memory size
cudaError_t launch_small(const CholArgs& A, int n_leaves, int sms, cudaStream_t stream) {
  const size_t smem = small_smem_bytes(A.n, A.d, n_leaves);
  if (smem > SMALL_SMEM_OPTIN) return cudaErrorInvalidValue;
  const int nw = small_nw(A.n);
  int per_sm = 1;
  cudaError_t e = nw == 2   ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, small_lml_kernel<2>, 64, smem)
                  : nw == 4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, small_lml_kernel<4>, 128, smem)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, small_lml_kernel<8>, 256, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  const int grid = A.batch < per_sm * sms ? A.batch : per_sm * sms;
  switch (nw) {
    case 2: small_lml_kernel<2><<<grid, 64, smem, stream>>>(A); break;
    case 4: small_lml_kernel<4><<<grid, 128, smem, stream>>>(A); break;
    default: small_lml_kernel<8><<<grid, 256, smem, stream>>>(A); break;
  }
  return cudaGetLastError();
}

}  // namespace bgp
