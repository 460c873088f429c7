// NCCL binding for the multi-GPU C entry point (bgp_acq_sweep_nccl in bgp_api.cu) and its small helper
// kernels.  libnccl.so.2 is bound lazily with dlopen: the library has no link-time NCCL dependency, and a
// process that already holds an NCCL (PyTorch's) shares that instance, so the caller's ncclComm_t and the
// calls made here come from the same library.
#include <dlfcn.h>

#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

// the handful of NCCL declarations used (nccl.h is not needed at build time)
enum { kNcclFloat64 = 8, kNcclInt32 = 2, kNcclSum = 0, kNcclMax = 2, kNcclMin = 3 };

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* lib = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (lib) break;
    }
    if (lib) {
      api.all_reduce = (decltype(api.all_reduce))dlsym(lib, "ncclAllReduce");
      api.all_gather = (decltype(api.all_gather))dlsym(lib, "ncclAllGather");
      api.error_string = (decltype(api.error_string))dlsym(lib, "ncclGetErrorString");
      api.ok = api.all_reduce && api.all_gather;
    }
  }
  return api.ok ? &api : nullptr;
}

int nccl_dtype_f64() { return kNcclFloat64; }
int nccl_dtype_i32() { return kNcclInt32; }
int nccl_op_min() { return kNcclMin; }
int nccl_op_max() { return kNcclMax; }

// dst[s][0..m_loc) <- src[s][0..m_loc) with row strides: packs a rank's (S x m_loc) block into the padded
// (S x m_max) send buffer of the equal-size all-gather
__global__ void pad_rows_kernel(const double* __restrict__ src, int S, int m_loc, double* __restrict__ dst, int m_max) {
  const int s = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m_max; i += gridDim.x * blockDim.x)
    dst[(size_t)s * m_max + i] = i < m_loc ? src[(size_t)s * m_loc + i] : 0.0;
}

// gathered[r][s][i] (world x S x m_max, rank r's block has sizes m_total / world (+1 for the first
// m_total % world ranks)) -> out[s][lo_r + i] (S x m_total)
__global__ void unpack_rows_kernel(const double* __restrict__ gathered, int world, int S, int m_max, int m_total,
                                   double* __restrict__ out) {
  const int s = blockIdx.y, q = m_total / world, rem = m_total % world;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < m_total; g += gridDim.x * blockDim.x) {
    // rank that owns global candidate g under the contiguous-block rule
    const int split = rem * (q + 1);
    const int r = g < split ? g / (q + 1) : rem + (g - split) / (q > 0 ? q : 1);
    const int lo = r * q + (r < rem ? r : rem);
    out[(size_t)s * m_total + g] = gathered[((size_t)r * S + s) * m_max + (g - lo)];
  }
}

__global__ void column0_kernel(const double* __restrict__ stats, int S, int stride, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S) out[s] = stats[(size_t)s * stride];
}

// allref[r][s] = {EI max, global index, mu, sd} of rank r's shard -> best[s]: largest EI, ties to the
// smallest global index (numpy's argmax)
__global__ void ttei_pick_kernel(const double* __restrict__ allref, int world, int S, double* __restrict__ best) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  int br = 0;
  for (int r = 1; r < world; ++r) {
    const double* c = allref + ((size_t)r * S + s) * 4;
    const double* b = allref + ((size_t)br * S + s) * 4;
    if (c[0] > b[0] || (c[0] == b[0] && c[1] < b[1])) br = r;
  }
  for (int k = 0; k < 4; ++k) best[(size_t)s * 4 + k] = allref[((size_t)br * S + s) * 4 + k];
}

cudaError_t launch_pad_rows(const double* src, int S, int m_loc, double* dst, int m_max, cudaStream_t st) {
  pad_rows_kernel<<<dim3((m_max + 255) / 256 > 256 ? 256 : (m_max + 255) / 256, S), 256, 0, st>>>(src, S, m_loc, dst, m_max);
  return cudaGetLastError();
}
cudaError_t launch_unpack_rows(const double* gathered, int world, int S, int m_max, int m_total, double* out,
                               cudaStream_t st) {
  unpack_rows_kernel<<<dim3((m_total + 255) / 256 > 256 ? 256 : (m_total + 255) / 256, S), 256, 0, st>>>(
      gathered, world, S, m_max, m_total, out);
  return cudaGetLastError();
}
cudaError_t launch_column0(const double* stats, int S, int stride, double* out, cudaStream_t st) {
  column0_kernel<<<(S + 127) / 128, 128, 0, st>>>(stats, S, stride, out);
  return cudaGetLastError();
}
cudaError_t launch_ttei_pick(const double* allref, int world, int S, double* best, cudaStream_t st) {
  ttei_pick_kernel<<<(S + 127) / 128, 128, 0, st>>>(allref, world, S, best);
  return cudaGetLastError();
}

}  // namespace bgp
