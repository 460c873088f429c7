// K1: batched-theta kernel-matrix construction.  One launch builds the lower triangle of
// K_theta(X, X) + diag(alpha) for every theta of the batch, panel by panel, directly in the
// tiled factor-slab layout the factorisation kernel consumes (so the matrix lives in L2 between
// the two kernels and is never read from HBM as an input).  ARD scaled differences use the
// difference form (no ||a||^2 + ||b||^2 - 2ab cancellation); the diagonal is exact.
// Replaces kernel(self.X_train_) of sklearn:_gpr.py:586 / bask/bayesgpr.py:203.
//
// Why a separate kernel: with W/2 = 64 thetas per MCMC half step only 64 of the 148 SMs run a
// factorisation; the Gram build is embarrassingly parallel FP64 ALU work (sqrt + exp per entry)
// and spreads over all SMs here instead of stretching each factorisation CTA's critical path.
#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

// Xt[b][leaf][dim][nx] = X[i][dim] / length_scale(theta_b, leaf, dim), zero padded to nx = 32 P + 64
// rows so that the 64-row chunks of gram_kernel can read their eight rows without clamping
__host__ __device__ inline int gram_nx(int n) { return 32 * ((n + 31) / 32) + 64; }

__global__ void __launch_bounds__(256) scale_x_kernel(GramArgs A) {
  __shared__ DevProgram PR;
  __shared__ ThetaParams TP;
  const int tid = threadIdx.x, b = blockIdx.y;
  {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&PR);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  resolve_theta(PR, A.theta + (size_t)b * PR.n_theta, A.fixed_ls, TP, tid, 256);
  __syncthreads();
  const int npad = gram_nx(A.n), d = A.d;
  double* Xt = A.xt + (size_t)b * A.xt_stride;
  for (int e = blockIdx.x * 256 + tid; e < PR.n_leaves * d * npad; e += gridDim.x * 256) {
    const int l = e / (d * npad), rem = e - l * d * npad, kk = rem / npad, i = rem - kk * npad;
    Xt[e] = (i < A.n) ? bgp_warp_coord(PR, A.theta + (size_t)b * PR.n_theta, kk, A.X[(size_t)i * d + kk]) *
                            TP.inv_ls[l][kk]
                      : 0.0;
  }
  // the resolved constants of this theta (exp(theta) per op) follow the scaled inputs, so that the
  // CTAs of gram_kernel read two doubles instead of resolving theta again
  if (blockIdx.x == 0 && tid < BGP_MAX_OPS)
    Xt[(size_t)(PR.n_leaves > 0 ? PR.n_leaves : 1) * d * npad + tid] = tid < PR.n_ops ? TP.opval[tid] : 0.0;
}

// CTA (32-row chunk of panel k, theta b): 32 x 32 entries of the lower triangle.  A warp takes
// four rows (four independent sqrt/exp chains; 64 registers -- at 40 the kernel spilled ~700k instructions per launch), lanes are the 32 columns.  Chunks of all panels
// are enumerated along blockIdx.x so that every CTA has the same amount of work.
constexpr int GRAM_RB = 4;                 // rows per warp
constexpr int GRAM_ROWS = 8 * GRAM_RB;     // rows per CTA
__host__ __device__ inline int gram_chunks(int n, int k) { return (n - 32 * k + GRAM_ROWS - 1) / GRAM_ROWS; }

__global__ void __launch_bounds__(256, 4) gram_kernel(GramArgs A) {
  __shared__ DevProgram PR;
  __shared__ ThetaParams TP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, n = A.n, d = A.d;
  int k = 0, chunk = blockIdx.x;
  while (chunk >= gram_chunks(n, k)) { chunk -= gram_chunks(n, k); ++k; }
  const int fast_kind = A.prog->fast_kind, n_leaves = A.prog->n_leaves;
  const double* opv = A.xt + (size_t)b * A.xt_stride + (size_t)(n_leaves > 0 ? n_leaves : 1) * d * gram_nx(n);
  if (!fast_kind) {
    // interpreter path: program and resolved constants into shared memory (the length scales are
    // already folded into Xt)
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&PR);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += 256) dst[i] = src[i];
    if (tid < BGP_MAX_OPS) TP.opval[tid] = opv[tid];
    __syncthreads();
  }
  const SlabGeom G = SlabGeom::make(n, A.aug != 0);
  const int npad = gram_nx(n);
  const double* Xt = A.xt + (size_t)b * A.xt_stride;
  double* base = A.slabs + (size_t)b * G.doubles() + G.off(k);
  constexpr int RB = GRAM_RB;
  const int c0 = 32 * k, col = c0 + lane;
  {
    const int r0 = c0 + GRAM_ROWS * chunk + RB * warp;
    if (fast_kind) {
      double r2[RB];
#pragma unroll
      for (int a = 0; a < RB; ++a) r2[a] = 0.0;
      // the eight rows of a warp are contiguous in Xt: four 16-byte loads per dimension, no
      // per-element address arithmetic (r0 is a multiple of 4, the row stride a multiple of 32)
      const double* xr = Xt + r0;
      const double* xcp = Xt + col;
      for (int kk = 0; kk < d; ++kk, xr += npad, xcp += npad) {
        const double xc = *xcp;
        double2 rv[RB / 2];
#pragma unroll
        for (int a = 0; a < RB / 2; ++a) rv[a] = reinterpret_cast<const double2*>(xr)[a];
#pragma unroll
        for (int a = 0; a < RB / 2; ++a) {
          const double t0 = rv[a].x - xc, t1 = rv[a].y - xc;
          r2[2 * a] = fma(t0, t0, r2[2 * a]);
          r2[2 * a + 1] = fma(t1, t1, r2[2 * a + 1]);
        }
      }
      const double cval = opv[A.prog->fast_const], wval = opv[A.prog->fast_white];
      double v[RB];
#pragma unroll
      for (int a = 0; a < RB; ++a) {
        const bool same = (r0 + a) == col;
        v[a] = cval * stationary_value(fast_kind, same ? 0.0 : r2[a]);
        if (same) v[a] += wval + A.alpha[min(r0 + a, n - 1)];
      }
#pragma unroll
      for (int a = 0; a < RB; ++a) {
        const int row = r0 + a;
        if (row < n && col <= row) base[(size_t)(row - c0) * 32 + lane] = v[a];
      }
    } else {
      for (int a = 0; a < RB; ++a) {
        const int row = r0 + a;
        if (row >= n || col > row) continue;
        double r2[BGP_MAX_LEAVES];
#pragma unroll
        for (int l = 0; l < BGP_MAX_LEAVES; ++l) {
          r2[l] = 0.0;
          if (l < PR.n_leaves && row != col) {
            const double* xl = Xt + (size_t)l * d * npad;
            double acc2 = 0.0;
            for (int kk = 0; kk < d; ++kk) {
              const double t = xl[(size_t)kk * npad + row] - xl[(size_t)kk * npad + col];
              acc2 = fma(t, t, acc2);
            }
            r2[l] = acc2;
          }
        }
        double v = eval_program(PR, TP, r2, row == col, true);
        if (row == col) v += A.alpha[row];
        base[(size_t)(row - c0) * 32 + lane] = v;
      }
    }
  }
}

// Fast-path kernel without input warping (the default bask kernel, c * stationary(r) + white): ONE launch.
// A CTA resolves its theta once (exp of the constant / white / length-scale entries) and scales ALL n points by
// the inverse length scales into shared memory, transposed ([d][n_pad]: 24 KB at n = 500, d = 6) -- no scaled
// copy of X in global memory, no second launch, and nothing to synchronise on afterwards: the warps then walk
// independently over a contiguous range of 32 x 32 chunks of the lower triangle, four rows per warp (four
// independent sqrt / exp chains), lanes on the columns.  The exact-diagonal / white-noise / alpha handling
// only exists in the diagonal chunk of a panel (a warp-uniform branch).  gridDim.x CTAs share a theta's
// chunks, sized so that the grid is one resident wave.
constexpr int GRAM_FUSED_MAX_SMEM = 56 * 1024;   // 4 CTAs per SM stay resident; larger problems take two launches
__host__ __device__ inline int gram_fused_stride(int n) { return 32 * ((n + 31) / 32) + 32; }
bool gram_fused_fits(int n, int d) { return sizeof(double) * (size_t)d * gram_fused_stride(n) <= (size_t)GRAM_FUSED_MAX_SMEM; }

__global__ void __launch_bounds__(256, 4) gram_fused_kernel(GramArgs A) {
  __shared__ DevProgram PR;
  __shared__ ThetaParams TP;
  extern __shared__ double xs[];   // [d][stride] scaled inputs, zero beyond n
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, n = A.n, d = A.d;
  {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&PR);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  resolve_theta(PR, A.theta + (size_t)b * PR.n_theta, A.fixed_ls, TP, tid, 256);
  __syncthreads();
  const int stride = gram_fused_stride(n);
  for (int kk = 0; kk < d; ++kk) {
    const double il = TP.inv_ls[0][kk];
    for (int i = tid; i < stride; i += 256) xs[kk * stride + i] = i < n ? A.X[(size_t)i * d + kk] * il : 0.0;
  }
  __syncthreads();
  const int fast_kind = PR.fast_kind;
  const double cval = TP.opval[PR.fast_const], wval = TP.opval[PR.fast_white];
  const SlabGeom G = SlabGeom::make(n, A.aug != 0);
  const int P = (n + 31) / 32;
  double* slab = A.slabs + (size_t)b * G.doubles();
  constexpr int RB = GRAM_RB;
  int total = 0;
  for (int j = 0; j < P; ++j) total += gram_chunks(n, j);
  const int first = (int)((long long)total * blockIdx.x / gridDim.x), last = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
  int k = 0, kbase = 0;   // panel of the current chunk, chunks before that panel
  for (int chunk = first; chunk < last; ++chunk) {
    while (chunk >= kbase + gram_chunks(n, k)) { kbase += gram_chunks(n, k); ++k; }
    const int c0 = 32 * k, R0 = c0 + GRAM_ROWS * (chunk - kbase);
    const double* xc = xs + c0 + lane;
    const double* xr = xs + R0 + RB * warp;
    double r2[RB];
#pragma unroll
    for (int a = 0; a < RB; ++a) r2[a] = 0.0;
    for (int kk = 0; kk < d; ++kk) {
      const double cv = xc[kk * stride];
      const double2 r01 = *reinterpret_cast<const double2*>(xr + kk * stride);
      const double2 r23 = *reinterpret_cast<const double2*>(xr + kk * stride + 2);
      const double t0 = r01.x - cv, t1 = r01.y - cv, t2 = r23.x - cv, t3 = r23.y - cv;
      r2[0] = fma(t0, t0, r2[0]); r2[1] = fma(t1, t1, r2[1]);
      r2[2] = fma(t2, t2, r2[2]); r2[3] = fma(t3, t3, r2[3]);
    }
    const int r0 = R0 + RB * warp, col = c0 + lane;
    double* base = slab + G.off(k) + (size_t)(r0 - c0) * 32 + lane;
    if (R0 != c0) {
      // strictly below the diagonal block: every column of the chunk is left of every row
#pragma unroll
      for (int a = 0; a < RB; ++a) {
        const double v = cval * stationary_value(fast_kind, r2[a]);
        if (r0 + a < n && col < n) base[a * 32] = v;
      }
    } else {
#pragma unroll
      for (int a = 0; a < RB; ++a) {
        const int row = r0 + a;
        const bool same = row == col;
        double v = cval * stationary_value(fast_kind, same ? 0.0 : r2[a]);
        if (same && row < n) v += wval + A.alpha[row];
        if (row < n && col <= row) base[a * 32] = v;
      }
    }
  }
}

// out[s][i][kk] = warp_{theta_s}(X[i][kk]): the per-theta warped copy of a point set that the
// sweep / posterior-covariance kernels read instead of the raw points when warping is on
__global__ void __launch_bounds__(256) warp_points_kernel(const double* __restrict__ X, int npts, int d,
                                                          const double* __restrict__ theta, const DevProgram* prog,
                                                          double* __restrict__ out) {
  const int s = blockIdx.y;
  const DevProgram& PR = *prog;
  const double* th = theta + (size_t)s * PR.n_theta;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < npts * d; e += gridDim.x * 256)
    out[(size_t)s * npts * d + e] = bgp_warp_coord(PR, th, e % d, X[e]);
}

cudaError_t launch_warp_points(const double* X, int npts, int d, const double* theta, int S, const DevProgram* prog,
                               double* out, cudaStream_t stream) {
  int blocks = (npts * d + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  warp_points_kernel<<<dim3(blocks, S), 256, 0, stream>>>(X, npts, d, theta, prog, out);
  return cudaGetLastError();
}

size_t gram_xt_doubles(int n, int d, int n_leaves) {
  return (size_t)(n_leaves > 0 ? n_leaves : 1) * d * gram_nx(n) + 32;   // + resolved op constants
}

cudaError_t launch_scale_x(const GramArgs& A, cudaStream_t stream) {
  const int per_theta = (int)((gram_xt_doubles(A.n, A.d, 1) * 4 + 255) / 256);
  dim3 gs(per_theta < 1 ? 1 : (per_theta > 32 ? 32 : per_theta), A.batch);
  scale_x_kernel<<<gs, 256, 0, stream>>>(A);
  return cudaGetLastError();
}

cudaError_t prepare_gram() {
  return cudaFuncSetAttribute(gram_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_FUSED_MAX_SMEM);
}

cudaError_t launch_gram(const GramArgs& A, cudaStream_t stream) {
  const int P = (A.n + 31) / 32;
  int chunks = 0;
  for (int k = 0; k < P; ++k) chunks += gram_chunks(A.n, k);
  if (A.fused_ok && gram_fused_fits(A.n, A.d)) {
    // one resident wave: 4 CTAs per SM shared evenly by the thetas of the batch
    int per_theta = (4 * (A.sms > 0 ? A.sms : 148)) / A.batch;
    per_theta = per_theta < 1 ? 1 : (per_theta > chunks ? chunks : per_theta);
    gram_fused_kernel<<<dim3(per_theta, A.batch), 256, sizeof(double) * A.d * gram_fused_stride(A.n), stream>>>(A);
    return cudaGetLastError();
  }
  cudaError_t e0 = launch_scale_x(A, stream);
  if (e0 != cudaSuccess) return e0;
  dim3 gg(chunks, A.batch);
  gram_kernel<<<gg, 256, 0, stream>>>(A);
  return cudaGetLastError();
}

}  // namespace bgp
