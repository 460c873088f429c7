// K4: candidate sweep.  For one theta and a tile of candidates the CTA
//   1. builds the cross-covariances k*(x_c, X) in shared memory (fused K1, never in HBM),
//   2. computes the whitened vectors v = L^-1 k* as a triangular GEMM on DMMA.8x8x4, streaming
//      L^-1 (the L^-T rows of the factor slab) from L2 with 16-byte loads four steps ahead,
//   3. reduces  var = k(x,x) - |v|^2  and  mean = v . z  (z = L^-1 y)  in the epilogue, so
//      only 16 bytes per (theta, candidate) reach HBM.
// Replaces skopt GaussianProcessRegressor.predict as called by bask/acquisition.py:121-129
// (einsum "ki,kj,ij->k" with the explicit K_inv_) and the cho_solve loops of PVRS / VR
// (bask/acquisition.py:285-339) through the optional extra right-hand sides.
#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

constexpr int SW1_NW = 16;

struct Sweep1Smem {
  DevProgram prog;
  ThetaParams tp;
  double kss[32];                 // k(x_c, x_c)
  double sums[SW1_NW][32][2];      // per-warp partial |v|^2 and v.z
};

template <int NT>
__global__ void __launch_bounds__(SW1_NW * 32, 1) sweep_kernel_v1(SweepArgs A) {
  constexpr int NC = 8 * NT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sweep1Smem& S = *reinterpret_cast<Sweep1Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int n = A.n, d = A.d, s = blockIdx.y, c0 = blockIdx.x * NC;
  const SlabGeom G = SlabGeom::make(n, true);
  const int P = G.P, kstride = 32 * P + 8;
  double* Xs = reinterpret_cast<double*>(smem_raw + ((sizeof(Sweep1Smem) + 15) & ~size_t(15)));
  double* Ks = Xs + (size_t)BGP_MAX_LEAVES * NC * d;   // NC x kstride
  double* dotx = Ks + (size_t)NC * kstride;            // R x NC
  {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&S.prog);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += SW1_NW * 32) dst[i] = src[i];
  }
  __syncthreads();
  const DevProgram& PR = S.prog;
  resolve_theta(PR, A.theta + (size_t)s * PR.n_theta, A.fixed_ls, S.tp, tid, SW1_NW * 32);
  for (int e = tid; e < SW1_NW * 32 * 2; e += SW1_NW * 32) (&S.sums[0][0][0])[e] = 0.0;
  for (int e = tid; e < A.R * NC; e += SW1_NW * 32) dotx[e] = 0.0;
  __syncthreads();
  // scaled candidate coordinates per stationary leaf
  for (int e = tid; e < PR.n_leaves * NC * d; e += SW1_NW * 32) {
    int l = e / (NC * d), rem = e - l * NC * d, c = rem / d, kk = rem - c * d;
    int ci = c0 + c;
    Xs[e] = (ci < A.m) ? A.Xc[(size_t)s * A.xc_stride + (size_t)ci * d + kk] * S.tp.inv_ls[l][kk] : 0.0;
  }
  __syncthreads();
  // cross-covariance tile
  for (int e = tid; e < NC * 32 * P; e += SW1_NW * 32) {
    const int c = e / (32 * P), i = e - c * 32 * P;
    double v = 0.0;
    if (i < n && c0 + c < A.m) {
      double r2[BGP_MAX_LEAVES];
#pragma unroll
      for (int l = 0; l < BGP_MAX_LEAVES; ++l) {
        r2[l] = 0.0;
        if (l < PR.n_leaves) {
          const double* xc = Xs + (size_t)(l * NC + c) * d;
          double acc = 0.0;
          for (int kk = 0; kk < d; ++kk) {
            double t = __ldg(A.X + (size_t)s * A.x_stride + (size_t)i * d + kk) * S.tp.inv_ls[l][kk] - xc[kk];
            acc = fma(t, t, acc);
          }
          r2[l] = acc;
        }
      }
      v = eval_program(PR, S.tp, r2, false, true);
    }
    Ks[(size_t)c * kstride + i] = v;
  }
  if (tid < NC) {
    double r2[BGP_MAX_LEAVES] = {0, 0, 0, 0};
    S.kss[tid] = eval_program(PR, S.tp, r2, true, A.noise_off == 0);
  }
  __syncthreads();

  const double* slab = A.slabs + (size_t)s * G.doubles();
  const double* z = A.z + (size_t)s * n;
  // Work unit = half a row panel of L^-1 (16 rows = two 8-row DMMA tiles); unit u needs the
  // first 16(u+1) columns (triangular), so units are dealt to warps in zig-zag pairs (w, 2W-1-w):
  // every warp gets the same number of 8-column steps whatever P is.
  const int U = 2 * P;
  for (int base = 0; base < U; base += 2 * SW1_NW) {
    for (int side = 0; side < 2; ++side) {
      const int u = side == 0 ? base + warp : base + 2 * SW1_NW - 1 - warp;
      if (u >= U) continue;
      const int j = u >> 1, hf = u & 1;
      double acc[2][NT][2];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int v = 0; v < NT; ++v) acc[t][v][0] = acc[t][v][1] = 0.0;
      // A operand: (L^-1)[32j + 16hf + 2r + t][i] lives at aug_base(j) + 32 i + 16hf + 2r + t
      const double* ap = slab + G.aug_base(j) + 16 * hf + 2 * r + (size_t)32 * (2 * q);
      const int steps = 2 * (u + 1);
      constexpr int PF = 4;
      double2 ring[PF][2];
#pragma unroll
      for (int s2 = 0; s2 < PF; ++s2) {
        ring[s2][0] = ring[s2][1] = make_double2(0.0, 0.0);
        if (s2 < steps) {
          ring[s2][0] = *reinterpret_cast<const double2*>(ap + (size_t)256 * s2);
          ring[s2][1] = *reinterpret_cast<const double2*>(ap + (size_t)256 * s2 + 32);
        }
      }
      for (int st0 = 0; st0 < steps; st0 += PF) {
#pragma unroll
        for (int s2 = 0; s2 < PF; ++s2) {
          const int st = st0 + s2;
          if (st < steps) {
            const double2 a0 = ring[s2][0], a1 = ring[s2][1];   // k = 8st+2q (rows 2r, 2r+1), k+1
            if (st + PF < steps) {
              ring[s2][0] = *reinterpret_cast<const double2*>(ap + (size_t)256 * (st + PF));
              ring[s2][1] = *reinterpret_cast<const double2*>(ap + (size_t)256 * (st + PF) + 32);
            }
            double2 bv[NT];
#pragma unroll
            for (int v = 0; v < NT; ++v)
              bv[v] = *reinterpret_cast<const double2*>(Ks + (size_t)(8 * v + r) * kstride + 8 * st + 2 * q);
#pragma unroll
            for (int v = 0; v < NT; ++v) { dmma(acc[0][v], a0.x, bv[v].x); dmma(acc[1][v], a0.y, bv[v].x); }
#pragma unroll
            for (int v = 0; v < NT; ++v) { dmma(acc[0][v], a1.x, bv[v].y); dmma(acc[1][v], a1.y, bv[v].y); }
          }
        }
      }
      // epilogue of this unit: rows 32j + 16hf + 2r + t
      const int row0 = 32 * j + 16 * hf + 2 * r;
      double zr[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) zr[t] = (row0 + t < n) ? z[row0 + t] : 0.0;
#pragma unroll
      for (int v = 0; v < NT; ++v) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double pv = 0.0, pm = 0.0;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            pv = fma(acc[t][v][e], acc[t][v][e], pv);
            pm = fma(acc[t][v][e], zr[t], pm);
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            pv += __shfl_xor_sync(0xffffffffu, pv, o);
            pm += __shfl_xor_sync(0xffffffffu, pm, o);
          }
          if (r == 0) {
            S.sums[warp][8 * v + 2 * q + e][0] += pv;
            S.sums[warp][8 * v + 2 * q + e][1] += pm;
          }
        }
      }
      if (A.R > 0) {
        for (int rr = 0; rr < A.R; ++rr) {
          const double* ze = A.zextra + ((size_t)s * A.R + rr) * n;
          double ez[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) ez[t] = (row0 + t < n) ? ze[row0 + t] : 0.0;
#pragma unroll
          for (int v = 0; v < NT; ++v)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              double pe = fma(acc[0][v][e], ez[0], acc[1][v][e] * ez[1]);
#pragma unroll
              for (int o = 4; o < 32; o <<= 1) pe += __shfl_xor_sync(0xffffffffu, pe, o);
              if (r == 0) atomicAdd(&dotx[rr * NC + 8 * v + 2 * q + e], pe);
            }
        }
      }
      if (A.v_out) {
#pragma unroll
        for (int v = 0; v < NT; ++v)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int ci = c0 + 8 * v + 2 * q + e;
            if (ci < A.m) {
              double* dst = A.v_out + ((size_t)s * A.m + ci) * A.v_ld + row0;
              *reinterpret_cast<double2*>(dst) = make_double2(acc[0][v][e], acc[1][v][e]);
            }
          }
      }
    }
  }
  __syncthreads();
  if (tid < NC && c0 + tid < A.m) {
    double vv = 0.0, mm = 0.0;
    for (int w = 0; w < SW1_NW; ++w) { vv += S.sums[w][tid][0]; mm += S.sums[w][tid][1]; }
    double var = S.kss[tid] - vv;
    if (var < 0.0) var = 0.0;
    const size_t o = (size_t)s * A.m + c0 + tid;
    A.mu[o] = A.y_std * mm + A.y_mean;
    A.sd[o] = sqrt(var * A.y_std * A.y_std);
  }
  for (int e = tid; e < A.R * NC; e += SW1_NW * 32) {
    const int rr = e / NC, c = e - rr * NC;
    if (c0 + c < A.m) A.dots[((size_t)s * A.R + rr) * A.m + c0 + c] = dotx[e];
  }
}

static size_t sweep1_smem(int nc, int n, int d, int R) {
  const int P = (n + 31) / 32;
  size_t base = (sizeof(Sweep1Smem) + 15) & ~size_t(15);
  return base + sizeof(double) * ((size_t)BGP_MAX_LEAVES * nc * d + (size_t)nc * (32 * P + 8) + (size_t)R * nc);
}

template <int NT>
static cudaError_t launch1_nt(const SweepArgs& A, cudaStream_t stream) {
  const size_t smem = sweep1_smem(8 * NT, A.n, A.d, A.R);
  cudaError_t e = cudaFuncSetAttribute(sweep_kernel_v1<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((A.m + 8 * NT - 1) / (8 * NT), A.S);
  sweep_kernel_v1<NT><<<grid, SW1_NW * 32, smem, stream>>>(A);
  return cudaGetLastError();
}

cudaError_t launch_sweep_v1(const SweepArgs& A, cudaStream_t stream) {
  const size_t cap = 200 * 1024;
  if (sweep1_smem(32, A.n, A.d, A.R) <= cap) return launch1_nt<4>(A, stream);
  if (sweep1_smem(16, A.n, A.d, A.R) <= cap) return launch1_nt<2>(A, stream);
  if (sweep1_smem(8, A.n, A.d, A.R) <= cap) return launch1_nt<1>(A, stream);
  return cudaErrorInvalidValue;
}

}  // namespace bgp
