// Chip-wide dense FP64 Cholesky for ONE large SPD matrix (the m x m posterior covariance behind joint
// draws: ThompsonSampling / PVRS / sample_y at thousands of candidates, bask/bayesgpr.py:637-718 ->
// sklearn:_gpr.py:502-539, where numpy factors the covariance with an SVD).
//
// Blocked right-looking, 256-column blocks, in place in the caller's row-major matrix (lower triangle):
//   1. the 256 x 256 diagonal block is factored by this library's own DMMA kernel (bgp_chol.cu, dense mode,
//      a cluster of 8 CTAs) and written back as L_JJ;
//   2. the block column below it is one triangular solve, the trailing matrix one symmetric rank-256 update --
//      plain library BLAS-3 (cuBLAS dtrsm / dsyrk), which is where m^3/3 of the flops are and what a
//      whole-chip kernel is for; the cluster kernel alone keeps a single matrix on 8 of the 148 SMs.
// cuBLAS is bound lazily with dlopen (no link-time dependency: the hot path of the library never needs it,
// and a process that already holds a libcublas.so.12 -- PyTorch's -- shares that instance).
// Row-major lower L is column-major upper U = L^T, so the calls below are the upper-Cholesky forms.
#include <cublas_v2.h>
#include <dlfcn.h>

#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

struct CublasApi {
  void* lib = nullptr;
  cublasStatus_t (*create)(cublasHandle_t*) = nullptr;
  cublasStatus_t (*destroy)(cublasHandle_t) = nullptr;
  cublasStatus_t (*set_stream)(cublasHandle_t, cudaStream_t) = nullptr;
  cublasStatus_t (*dtrsm)(cublasHandle_t, cublasSideMode_t, cublasFillMode_t, cublasOperation_t, cublasDiagType_t, int,
                          int, const double*, const double*, int, double*, int) = nullptr;
  cublasStatus_t (*dsyrk)(cublasHandle_t, cublasFillMode_t, cublasOperation_t, int, int, const double*, const double*,
                          int, const double*, double*, int) = nullptr;
  cublasStatus_t (*dtrmm)(cublasHandle_t, cublasSideMode_t, cublasFillMode_t, cublasOperation_t, cublasDiagType_t, int,
                          int, const double*, const double*, int, const double*, int, double*, int) = nullptr;
};

static CublasApi* cublas_api() {
  static CublasApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libcublas.so.12", "libcublas.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.create = (decltype(api.create))dlsym(api.lib, "cublasCreate_v2");
      api.destroy = (decltype(api.destroy))dlsym(api.lib, "cublasDestroy_v2");
      api.set_stream = (decltype(api.set_stream))dlsym(api.lib, "cublasSetStream_v2");
      api.dtrsm = (decltype(api.dtrsm))dlsym(api.lib, "cublasDtrsm_v2");
      api.dsyrk = (decltype(api.dsyrk))dlsym(api.lib, "cublasDsyrk_v2");
      api.dtrmm = (decltype(api.dtrmm))dlsym(api.lib, "cublasDtrmm_v2");
      if (!api.create || !api.destroy || !api.set_stream || !api.dtrsm || !api.dsyrk || !api.dtrmm) api.lib = nullptr;
    }
  }
  return api.lib ? &api : nullptr;
}

// L_JJ out of the factor slab of the diagonal block into the lower triangle of the caller's matrix
__global__ void block_L_writeback_kernel(const double* __restrict__ slab, double* __restrict__ a, long long lda, int nb) {
  const SlabGeom G = SlabGeom::make(nb, false);
  const int row = blockIdx.x;
  for (int c = threadIdx.x; c <= row; c += blockDim.x) {
    const int j = c >> 5;
    a[(size_t)row * lda + c] = slab[G.off(j) + (size_t)(row - 32 * j) * 32 + (c & 31)];
  }
}

// info = LAPACK's dpotrf convention over the whole matrix: first failing column + 1, or 0
__global__ void big_info_kernel(const int32_t* __restrict__ block_info, int nblocks, int nb, int32_t* __restrict__ info) {
  int v = 0;
  for (int b = 0; b < nblocks && v == 0; ++b)
    if (block_info[b] != 0) v = b * nb + block_info[b];
  info[0] = v;
}

__global__ void add_mean_kernel(double* __restrict__ out, const double* __restrict__ mean, int m, int ns) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (size_t)m * ns) out[e] += mean[e / ns];
}

constexpr int BIG_NB = 256;

size_t big_workspace_bytes(int m) {
  const int nblocks = (m + BIG_NB - 1) / BIG_NB;
  return sizeof(double) * (size_t)SlabGeom::make(BIG_NB, false).doubles() + sizeof(int32_t) * (size_t)(nblocks + 2);
}

const char* big_cholesky(void** cublas_handle, double* a, int m, long long lda, double jitter, int32_t* info,
                         void* workspace, int sms, cudaStream_t stream) {
  CublasApi* B = cublas_api();
  if (!B) return "libcublas.so.12 could not be loaded (needed only for the chip-wide dense Cholesky)";
  if (lda > 0x7fffffffLL) return "leading dimension too large";
  cublasHandle_t hb = (cublasHandle_t)*cublas_handle;
  if (!hb) {
    if (B->create(&hb) != CUBLAS_STATUS_SUCCESS) return "cublasCreate failed";
    *cublas_handle = hb;
  }
  if (B->set_stream(hb, stream) != CUBLAS_STATUS_SUCCESS) return "cublasSetStream failed";
  if (prepare_chol(BIG_NB) != cudaSuccess) return "shared-memory opt-in of the factorisation kernel failed";
  double* slab = static_cast<double*>(workspace);
  int32_t* binfo = reinterpret_cast<int32_t*>(slab + SlabGeom::make(BIG_NB, false).doubles());
  const int nblocks = (m + BIG_NB - 1) / BIG_NB;
  const double one = 1.0, minus_one = -1.0;
  for (int jb = 0; jb < nblocks; ++jb) {
    const int J = jb * BIG_NB, nb = m - J < BIG_NB ? m - J : BIG_NB, rem = m - J - nb;
    double* ajj = a + (size_t)J * lda + J;
    CholArgs A;
    memset(&A, 0, sizeof(A));
    A.slabs = slab; A.info = binfo + jb; A.n = nb; A.d = 1; A.batch = 1;
    A.dense = ajj; A.ldd = lda; A.jitter = jitter;
    if (launch_chol(A, 1, sms, stream) != cudaSuccess) return "diagonal-block factorisation failed to launch";
    block_L_writeback_kernel<<<nb, 128, 0, stream>>>(slab, ajj, lda, nb);
    if (rem > 0) {
      double* below = a + (size_t)(J + nb) * lda + J;           // column-major view: U_{J, rest}, nb x rem
      double* trail = a + (size_t)(J + nb) * lda + (J + nb);
      if (B->dtrsm(hb, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, nb, rem, &one, ajj,
                   (int)lda, below, (int)lda) != CUBLAS_STATUS_SUCCESS) return "cublasDtrsm failed";
      if (B->dsyrk(hb, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, rem, nb, &minus_one, below, (int)lda, &one, trail,
                   (int)lda) != CUBLAS_STATUS_SUCCESS) return "cublasDsyrk failed";
    }
  }
  big_info_kernel<<<1, 1, 0, stream>>>(binfo, nblocks, BIG_NB, info);
  return cudaGetLastError() == cudaSuccess ? nullptr : "kernel launch failed";
}

// out[i][s] = mean[i] + sum_c L[i][c] e[c][s]; L row-major lower in `l` (lda), e / out row-major m x ns
const char* big_trmm(void** cublas_handle, const double* l, int m, long long lda, const double* e, int ns,
                     const double* mean, double* out, cudaStream_t stream) {
  CublasApi* B = cublas_api();
  if (!B) return "libcublas.so.12 could not be loaded (needed only for the chip-wide dense Cholesky)";
  cublasHandle_t hb = (cublasHandle_t)*cublas_handle;
  if (!hb) {
    if (B->create(&hb) != CUBLAS_STATUS_SUCCESS) return "cublasCreate failed";
    *cublas_handle = hb;
  }
  if (B->set_stream(hb, stream) != CUBLAS_STATUS_SUCCESS) return "cublasSetStream failed";
  const double one = 1.0;
  if (B->dtrmm(hb, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, ns, m, &one, l, (int)lda,
               e, ns, out, ns) != CUBLAS_STATUS_SUCCESS) return "cublasDtrmm failed";
  if (mean) {
    const size_t total = (size_t)m * ns;
    add_mean_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(out, mean, m, ns);
  }
  return cudaGetLastError() == cudaSuccess ? nullptr : "kernel launch failed";
}

void big_release(void** cublas_handle) {
  CublasApi* B = cublas_api();
  if (B && *cublas_handle) B->destroy((cublasHandle_t)*cublas_handle);
  *cublas_handle = nullptr;
}

}  // namespace bgp
