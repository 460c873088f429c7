// Analytic gradient of the log-marginal likelihood with respect to the kernel's log hyper-parameters:
//   dLML/dtheta_k = 1/2 tr((alpha alpha^T - K^-1) dK/dtheta_k)
// (sklearn:_gpr.py:619-651, driven by the L-BFGS-B MAP search of the skopt fit, bask/bayesgpr.py:607).
// dK/dtheta_k is evaluated per matrix entry by running the covariance program on dual numbers
// (value, derivative with respect to the selected theta): leaves follow sklearn:kernels.py --
// ConstantKernel :1284-1296, WhiteKernel :1407-1419, RBF :1571-1587, Matern :1744-1786 --
// Sum / Product / Exponentiation the sum, product and power rules (:856-873, :957-973, :1104-1118).
// Inputs: alpha_ and K_inv_ of the factorisation at theta (bgp_factor_extract) and the scaled,
// transposed training inputs that the Gram kernel's scale_x pass leaves in the handle's scratch.
// Only the lower triangle is visited (weight 2 off the diagonal); per-block partial sums are
// combined in block order by a second kernel, so the result is deterministic.
#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

__host__ __device__ inline int grad_nx(int n) { return 32 * ((n + 31) / 32) + 64; }   // == gram_nx

// d k(r) / d log(length scale) for one stationary leaf: s2 = squared scaled difference along the
// selected dimension (ARD) or r2 itself (isotropic length scale)
__device__ __forceinline__ double stationary_dlogl(int code, double r2, double s2) {
  if (code == BGP_OP_RBF) return s2 * exp(-0.5 * r2);
  if (code == BGP_OP_MATERN52) {
    const double t = sqrt(5.0 * r2);
    return (5.0 / 3.0) * s2 * (t + 1.0) * exp(-t);
  }
  if (code == BGP_OP_MATERN32) return 3.0 * s2 * exp(-sqrt(3.0 * r2));
  const double r = sqrt(r2);                       // Matern 1/2: K * D / r, 0 where r = 0
  return r > 0.0 ? s2 * exp(-r) / r : 0.0;
}

// postfix program on (value, d value / d theta_own); `own` = op that holds the selected theta
__device__ __forceinline__ double eval_program_dual(const DevProgram& P, const double* opval, const double* r2,
                                                    double s2, int own, bool same_point) {
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0, d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#define BGP_PUSH2(v, dv) { v3 = v2; v2 = v1; v1 = v0; v0 = (v); d3 = d2; d2 = d1; d1 = d0; d0 = (dv); }
  for (int o = 0; o < P.n_ops; ++o) {
    const int code = P.ops[o].code;
    switch (code) {
      case BGP_OP_CONST: BGP_PUSH2(opval[o], o == own ? opval[o] : 0.0); break;
      case BGP_OP_WHITE: {
        const double w = same_point ? opval[o] : 0.0;
        BGP_PUSH2(w, o == own ? w : 0.0);
      } break;
      case BGP_OP_RBF: case BGP_OP_MATERN12: case BGP_OP_MATERN32: case BGP_OP_MATERN52: {
        const double rr = pick_leaf(r2, P.leaf_of_op[o]);
        BGP_PUSH2(stationary_value(code, rr), o == own ? stationary_dlogl(code, rr, s2) : 0.0);
      } break;
      case BGP_OP_ADD: { v0 = v1 + v0; d0 = d1 + d0; v1 = v2; v2 = v3; d1 = d2; d2 = d3; } break;
      case BGP_OP_MUL: { d0 = d1 * v0 + v1 * d0; v0 = v1 * v0; v1 = v2; v2 = v3; d1 = d2; d2 = d3; } break;
      case BGP_OP_POW: { d0 = opval[o] * pow(v0, opval[o] - 1.0) * d0; v0 = pow(v0, opval[o]); } break;
      default: break;
    }
  }
#undef BGP_PUSH2
  return d0;
}

constexpr int GT = 16;   // 16 x 16 entries per block

__global__ void __launch_bounds__(GT * GT) lml_grad_kernel(GradArgs A) {
  __shared__ DevProgram PR;
  __shared__ double opv[BGP_MAX_OPS];
  __shared__ double red[GT * GT / 32];
  const int tid = threadIdx.y * GT + threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&PR);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += GT * GT) dst[i] = src[i];
  }
  const int n = A.n, d = A.d, npad = grad_nx(n);
  const int nl = A.n_leaves > 0 ? A.n_leaves : 1;
  if (tid < BGP_MAX_OPS) opv[tid] = A.xt[(size_t)nl * d * npad + tid];
  __syncthreads();
  int bi = 0, bj = blockIdx.x;
  while (bj > bi) { bj -= bi + 1; ++bi; }
  const int row = GT * bi + threadIdx.y, col = GT * bj + threadIdx.x;
  const bool valid = row < n && col <= row;
  const int rc = min(row, n - 1), cc = min(col, n - 1);
  double r2[BGP_MAX_LEAVES];
#pragma unroll
  for (int l = 0; l < BGP_MAX_LEAVES; ++l) {
    r2[l] = 0.0;
    if (l < PR.n_leaves && rc != cc) {
      const double* xl = A.xt + (size_t)l * d * npad;
      double acc = 0.0;
      for (int kk = 0; kk < d; ++kk) {
        const double t = xl[(size_t)kk * npad + rc] - xl[(size_t)kk * npad + cc];
        acc = fma(t, t, acc);
      }
      r2[l] = acc;
    }
  }
  const double w = valid ? (row == col ? 0.5 : 1.0) * (A.alpha_vec[rc] * A.alpha_vec[cc] - A.kinv[(size_t)rc * n + cc])
                         : 0.0;
  for (int k = 0; k < A.p_kernel; ++k) {
    // op that owns theta_k and, for an ARD length scale, its dimension (uniform over the block)
    int own = -1, dim = -1;
    for (int o = 0; o < PR.n_ops; ++o) {
      const bgp_op_t& op = PR.ops[o];
      if (op.theta_idx < 0) continue;
      const bool stat = op.code >= BGP_OP_RBF && op.code <= BGP_OP_MATERN52;
      const int width = stat ? op.n_ls : 1;
      if (k >= op.theta_idx && k < op.theta_idx + width) {
        own = o;
        dim = (stat && op.n_ls > 1) ? k - op.theta_idx : -1;
      }
    }
    double s2 = 0.0;
    if (own >= 0 && dim >= 0) {
      const double* xl = A.xt + (size_t)PR.leaf_of_op[own] * d * npad;
      const double t = xl[(size_t)dim * npad + rc] - xl[(size_t)dim * npad + cc];
      s2 = t * t;
    } else if (own >= 0) {
      s2 = pick_leaf(r2, PR.leaf_of_op[own]);       // isotropic (only read for stationary owners)
    }
    double v = (valid && own >= 0) ? w * eval_program_dual(PR, opv, r2, s2, own, row == col) : 0.0;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int i = 0; i < GT * GT / 32; ++i) s += red[i];
      A.partial[(size_t)blockIdx.x * A.p_kernel + k] = s;
    }
    __syncthreads();
  }
}

__global__ void lml_grad_reduce_kernel(const double* __restrict__ partial, int nblocks, int p, double* __restrict__ grad) {
  const int k = blockIdx.x;
  // fixed order: thread t sums blocks t, t + 256, ...; the 256 partials are then added in index order
  __shared__ double part[256];
  double s = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 256) s += partial[(size_t)b * p + k];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 256; ++i) t += part[i];
    grad[k] = t;
  }
}

size_t grad_partial_doubles(int n, int p) {
  const int nt = (n + GT - 1) / GT;
  return (size_t)nt * (nt + 1) / 2 * p;
}

cudaError_t launch_lml_grad(const GradArgs& A, cudaStream_t stream) {
  const int nt = (A.n + GT - 1) / GT, nblocks = nt * (nt + 1) / 2;
  lml_grad_kernel<<<nblocks, dim3(GT, GT), 0, stream>>>(A);
  lml_grad_reduce_kernel<<<A.p_kernel, 256, 0, stream>>>(A.partial, nblocks, A.p_kernel, A.grad);
  return cudaGetLastError();
}

}  // namespace bgp
