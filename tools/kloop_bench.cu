// Developer tooling: cycles per 8-column step of the trailing-update K-loop (k_chunk<NTL>) in
// isolation, as a function of active warps per CTA, tiles per warp and CTAs on the chip.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/kloop_bench tools/kloop_bench.cu
#include <cstdio>
#include <vector>
#include "../bayes-skopt_b200/csrc/bgp_chol.cu"

template <int NTL>
__global__ void __launch_bounds__(256, 1) kloop_kernel(const double* slabs, int n, int k, int active_warps,
                                                       long long* clk, double* sink) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r = lane >> 2, q = lane & 3;
  const SlabGeom G = SlabGeom::make(n, false);
  const int bstride = 32 * k + 8;
  double* Bs = reinterpret_cast<double*>(smem_raw);
  double* region = Bs + (size_t)32 * bstride;
  const unsigned ring = (unsigned)__cvta_generic_to_shared(region) + warp * bgp::RING_BYTES + lane * 16;
  const double* slab = slabs + (size_t)blockIdx.x * G.doubles();
  for (int e = tid; e < 32 * bstride; e += 256) Bs[e] = 1e-3 * (e & 7);
  __syncthreads();
  double acc[4][4][2];
  for (int t = 0; t < 4; ++t) for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
  bgp::TileSet TS;
  for (int t = 0; t < 4; ++t) { TS.rb[t] = 32 * (k + 1) + 8 * ((warp * NTL + t) % (4 * (G.P - 1 - k))); TS.kind[t] = 0; TS.js[t] = 0; }
  long long t0 = clock64();
  if (warp < active_warps) bgp::k_chunk<NTL>(acc, TS, slab, G, Bs, bstride, 0, k, r, q, ring);
  long long t1 = clock64();
  double s = 0;
  for (int t = 0; t < NTL; ++t) for (int u = 0; u < 4; ++u) s += acc[t][u][0] + acc[t][u][1];
  sink[blockIdx.x * 256 + tid] = s;
  if (blockIdx.x == 0 && lane == 0) clk[warp] = t1 - t0;
}

int main() {
  const int n = 512, P = 16;
  const SlabGeom G = SlabGeom::make(n, false);
  const int maxb = 128;
  double* slabs; long long* dclk; double* sink;
  cudaMalloc(&slabs, sizeof(double) * G.doubles() * maxb);
  cudaMemset(slabs, 0, sizeof(double) * G.doubles() * maxb);
  cudaMalloc(&dclk, 64); cudaMalloc(&sink, 8 * 256 * maxb);
  const int k = 12;
  const size_t smem = sizeof(double) * 32 * (32 * k + 8) + 8 * bgp::RING_BYTES;
  cudaFuncSetAttribute(kloop_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(kloop_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("k=%d panels (%d steps); cycles per step of warp 0 / last active warp\n", k, 4 * k);
  for (int grid : {1, 64, 128}) {
    for (int ntl = 1; ntl <= 2; ++ntl) {
      for (int aw : {1, 2, 4, 5, 7, 8}) {
        long long clk[8];
        for (int rep = 0; rep < 2; ++rep) {
          if (ntl == 1) kloop_kernel<1><<<grid, 256, smem>>>(slabs, n, k, aw, dclk, sink);
          if (ntl == 2) kloop_kernel<2><<<grid, 256, smem>>>(slabs, n, k, aw, dclk, sink);
          cudaDeviceSynchronize();
        }
        cudaMemcpy(clk, dclk, 64, cudaMemcpyDeviceToHost);
        printf("grid %3d NTL %d warps %d : %6.1f / %6.1f clk per step  (%.0f%% of the DMMA rate for the busiest SMSP)\n", grid, ntl, aw,
               clk[0] / (4.0 * k), clk[aw - 1] / (4.0 * k),
               100.0 * (ntl * 128.0 * ((aw + 3) / 4)) / (clk[aw - 1] / (4.0 * k)));
      }
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  (void)P;
  return 0;
}
