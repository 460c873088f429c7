// Dense views of a factor slab, computed on demand for the estimator attributes the
// reference exposes (L_, K_inv_, alpha_: bask/bayesgpr.py:116-137, 200-217).  Not on the
// MCMC / sweep hot path.
#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

__global__ void extract_L_kernel(const double* __restrict__ slab, double* __restrict__ out, int n) {
  const SlabGeom G = SlabGeom::make(n, true);
  const int row = blockIdx.x;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    double v = 0.0;
    if (c <= row) {
      const int j = c >> 5;
      v = slab[G.off(j) + (size_t)(row - 32 * j) * 32 + (c & 31)];
    }
    out[(size_t)row * n + c] = v;
  }
}

// out[c][i] = (L^-1)[c][i]
__global__ void extract_Linv_kernel(const double* __restrict__ slab, double* __restrict__ out, int n) {
  const SlabGeom G = SlabGeom::make(n, true);
  const int c = blockIdx.x, j = c >> 5;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double v = 0.0;
    if (i <= c) v = slab[G.aug_base(j) + (size_t)32 * i + (c & 31)];
    out[(size_t)c * n + i] = v;
  }
}

// alpha[i] = sum_c (L^-1)[c][i] z[c]
__global__ void alpha_kernel(const double* __restrict__ slab, const double* __restrict__ z,
                             double* __restrict__ out, int n) {
  const SlabGeom G = SlabGeom::make(n, true);
  const int i = blockIdx.x;
  double s = 0.0;
  for (int c = i + threadIdx.x; c < n; c += blockDim.x) {
    const int j = c >> 5;
    s = fma(slab[G.aug_base(j) + (size_t)32 * i + (c & 31)], z[c], s);
  }
  s = warp_sum(s);
  __shared__ double red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    out[i] = t;
  }
}

// K_inv[a][b] = sum_{c >= max(a,b)} Linv[c][a] Linv[c][b]   (Linv dense, n x n)
__global__ void kinv_kernel(const double* __restrict__ Linv, double* __restrict__ out, int n) {
  __shared__ double As[16][17], Bs[16][17];
  const int a = blockIdx.y * 16 + threadIdx.y, b = blockIdx.x * 16 + threadIdx.x;
  double s = 0.0;
  const int cstart = (min(blockIdx.x, blockIdx.y) * 16) & ~15;
  for (int c0 = cstart; c0 < n; c0 += 16) {
    const int ca = c0 + threadIdx.x;
    As[threadIdx.y][threadIdx.x] = (ca < n && a < n) ? Linv[(size_t)ca * n + a] : 0.0;   // [a][c]
    const int cb = c0 + threadIdx.y;
    Bs[threadIdx.y][threadIdx.x] = (cb < n && b < n) ? Linv[(size_t)cb * n + b] : 0.0;   // [c][b]
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) s = fma(As[threadIdx.y][k], Bs[k][threadIdx.x], s);
    __syncthreads();
  }
  if (a < n && b < n) out[(size_t)a * n + b] = s;
}

cudaError_t launch_extract(const ExtractArgs& A, double* scratch, cudaStream_t stream) {
  const int n = A.n;
  switch (A.what) {
    case BGP_EXTRACT_L: extract_L_kernel<<<n, 128, 0, stream>>>(A.slab, A.out, n); break;
    case BGP_EXTRACT_LINV: extract_Linv_kernel<<<n, 128, 0, stream>>>(A.slab, A.out, n); break;
    case BGP_EXTRACT_ALPHA: alpha_kernel<<<n, 128, 0, stream>>>(A.slab, A.z, A.out, n); break;
    case BGP_EXTRACT_KINV: {
      if (!scratch) return cudaErrorInvalidValue;
      extract_Linv_kernel<<<n, 128, 0, stream>>>(A.slab, scratch, n);
      dim3 grid((n + 15) / 16, (n + 15) / 16), block(16, 16);
      kinv_kernel<<<grid, block, 0, stream>>>(scratch, A.out, n);
    } break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace bgp
