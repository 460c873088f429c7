// Joint posterior over a candidate set: cov = (K(X*,X*) - V V^T) y_std^2 with V = L^-1 K(X,X*)
// from the sweep, and draws mean + chol(cov) eps.  Replaces the return_cov branch of skopt's
// predict and sklearn sample_y (sklearn:_gpr.py:502-539) as used by BayesGPR.sample_y
// (bask/bayesgpr.py:637-718), ThompsonSampling and PVRS (bask/acquisition.py:270-274, 320-327).
#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

struct PostSmem {
  DevProgram prog;
  ThetaParams tp;
};

// 64 x 64 output tile per CTA (4 warps, 32 x 32 each); V V^T on DMMA with both operands
// streamed from L2 in 16-byte loads; lower tiles only, mirrored on store.
__global__ void __launch_bounds__(128) postcov_kernel(PostCovArgs A) {
  __shared__ PostSmem S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r = lane >> 2, q = lane & 3;
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&S.prog);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += 128) dst[i] = src[i];
  }
  __syncthreads();
  resolve_theta(S.prog, A.theta, A.fixed_ls, S.tp, tid, 128);
  __syncthreads();
  const int i0 = 64 * ti + 32 * (warp >> 1), j0 = 64 * tj + 32 * (warp & 1);
  double acc[4][4][2];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
  const int npad = 32 * ((A.n + 31) / 32);
  const double* pa[4];
  const double* pb[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int ia = min(i0 + 8 * t + r, A.m - 1), jb = min(j0 + 8 * t + r, A.m - 1);
    pa[t] = A.v + (size_t)ia * A.v_ld + 2 * q;
    pb[t] = A.v + (size_t)jb * A.v_ld + 2 * q;
  }
  for (int k = 0; k < npad; k += 8) {
    double2 av[4], bv[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      av[t] = *reinterpret_cast<const double2*>(pa[t] + k);
      bv[t] = *reinterpret_cast<const double2*>(pb[t] + k);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dmma(acc[t][u], av[t].x, bv[u].x);
        dmma(acc[t][u], av[t].y, bv[u].y);
      }
  }
  const double s2 = A.y_std * A.y_std;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = i0 + 8 * t + r, j = j0 + 8 * u + 2 * q + e;
        if (i >= A.m || j >= A.m || j > i) continue;
        double r2[BGP_MAX_LEAVES];
#pragma unroll
        for (int l = 0; l < BGP_MAX_LEAVES; ++l) {
          r2[l] = 0.0;
          if (l < S.prog.n_leaves && i != j) {
            double a2 = 0.0;
            for (int kk = 0; kk < A.d; ++kk) {
              const double tt = (A.Xc[(size_t)i * A.d + kk] - A.Xc[(size_t)j * A.d + kk]) * S.tp.inv_ls[l][kk];
              a2 = fma(tt, tt, a2);
            }
            r2[l] = a2;
          }
        }
        const double kij = eval_program(S.prog, S.tp, r2, i == j, A.noise_off == 0);
        const double c = (kij - acc[t][u][e]) * s2;
        A.cov[(size_t)i * A.ldc + j] = c;
        A.cov[(size_t)j * A.ldc + i] = c;
      }
}

cudaError_t launch_postcov(const PostCovArgs& A, cudaStream_t stream) {
  const int nt = (A.m + 63) / 64;
  dim3 grid(nt, nt);
  postcov_kernel<<<grid, 128, 0, stream>>>(A);
  return cudaGetLastError();
}

// out[i][s] = mean[i] + sum_{c <= i} L[i][c] E[c][s]     (L in slab layout, E row-major m x ns)
__global__ void slab_trmm_kernel(const double* __restrict__ slab, int m, const double* __restrict__ E, int ns,
                                 const double* __restrict__ mean, double* __restrict__ out) {
  const SlabGeom G = SlabGeom::make(m, false);
  const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __shared__ double red[8][32];
  for (int s0 = 0; s0 < ns; s0 += 8 * nw) {
    const int s = s0 + warp;
    double acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0;
    // each warp takes 8 consecutive samples, lanes stride over columns
    for (int c = lane; c <= i; c += 32) {
      const int j = c >> 5;
      const double l = slab[G.off(j) + (size_t)(i - 32 * j) * 32 + (c & 31)];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ss = s0 + warp * 8 + e;
        if (ss < ns) acc[e] = fma(l, E[(size_t)c * ns + ss], acc[e]);
      }
    }
    (void)s;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      double v = warp_sum(acc[e]);
      const int ss = s0 + warp * 8 + e;
      if (lane == 0 && ss < ns) out[(size_t)i * ns + ss] = v + (mean ? mean[i] : 0.0);
    }
  }
  (void)red;
}

cudaError_t launch_slab_trmm(const double* slab, int m, const double* E, int ns, const double* mean,
                             double* out, cudaStream_t stream) {
  slab_trmm_kernel<<<m, 128, 0, stream>>>(slab, m, E, ns, mean, out);
  return cudaGetLastError();
}

// PVRS epilogue (Schur form of bask/acquisition.py:328-339):
//   out[i] = sum_t |v_t|^2 + (k(xt_t, xc_i) - dots[t][i])^2 / s_i
__global__ void __launch_bounds__(256) pvrs_combine_kernel(CombineArgs A) {
  __shared__ PostSmem S;
  __shared__ double base_s;
  __shared__ double red[8];
  const int tid = threadIdx.x;
  {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&S.prog);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  resolve_theta(S.prog, A.theta, A.fixed_ls, S.tp, tid, 256);
  double b = 0.0;
  for (int e = tid; e < A.R * A.n; e += 256) b = fma(A.vt[e], A.vt[e], b);
  b = warp_sum(b);
  if ((tid & 31) == 0) red[tid >> 5] = b;
  __syncthreads();
  if (tid == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += red[w]; base_s = t; }
  __syncthreads();
  const int i = blockIdx.x * 256 + tid;
  if (i >= A.m) return;
  double acc = base_s;
  const double si = A.s[i];
  for (int t = 0; t < A.R; ++t) {
    double r2[BGP_MAX_LEAVES];
#pragma unroll
    for (int l = 0; l < BGP_MAX_LEAVES; ++l) {
      r2[l] = 0.0;
      if (l < S.prog.n_leaves) {
        double a2 = 0.0;
        for (int kk = 0; kk < A.d; ++kk) {
          const double tt = A.Xt[(size_t)t * A.d + kk] * S.tp.inv_ls[l][kk] - A.Xc[(size_t)i * A.d + kk] * S.tp.inv_ls[l][kk];
          a2 = fma(tt, tt, a2);
        }
        r2[l] = a2;
      }
    }
    const double c = eval_program(S.prog, S.tp, r2, false, true) - A.dots[(size_t)t * A.m + i];
    acc += c * c / si;
  }
  A.out[i] = acc;
}

// VR epilogue: out[i] = sum_t (k0 - C[t][t]) + sum_t C[t][i]^2 / s_i   (C: noise-free posterior cov)
__global__ void __launch_bounds__(256) vr_base_kernel(CombineArgs A, double* base_out) {
  __shared__ PostSmem S;
  __shared__ double red[8];
  const int tid = threadIdx.x;
  {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&S.prog);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  resolve_theta(S.prog, A.theta, A.fixed_ls, S.tp, tid, 256);
  __syncthreads();
  double r2[BGP_MAX_LEAVES] = {0, 0, 0, 0};
  const double k0 = eval_program(S.prog, S.tp, r2, true, false);
  double b = 0.0;
  for (int t = tid; t < A.m; t += 256) b += k0 - A.cov[(size_t)t * A.ldc + t];
  b = warp_sum(b);
  if ((tid & 31) == 0) red[tid >> 5] = b;
  __syncthreads();
  if (tid == 0) { double t = 0.0; for (int w = 0; w < 8; ++w) t += red[w]; *base_out = t; }
}
__global__ void __launch_bounds__(256) vr_cols_kernel(CombineArgs A, const double* base) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= A.m) return;
  double acc = 0.0;
  for (int t = 0; t < A.m; ++t) { const double c = A.cov[(size_t)t * A.ldc + i]; acc = fma(c, c, acc); }
  A.out[i] = *base + acc / A.s[i];
}

cudaError_t launch_pvrs_combine(const CombineArgs& A, cudaStream_t stream) {
  pvrs_combine_kernel<<<(A.m + 255) / 256, 256, 0, stream>>>(A);
  return cudaGetLastError();
}
cudaError_t launch_vr_combine(const CombineArgs& A, double* scratch, cudaStream_t stream) {
  vr_base_kernel<<<1, 256, 0, stream>>>(A, scratch);
  vr_cols_kernel<<<(A.m + 255) / 256, 256, 0, stream>>>(A, scratch);
  return cudaGetLastError();
}

}  // namespace bgp
