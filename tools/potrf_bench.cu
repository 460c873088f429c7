// Developer tooling: latency of the warp-level 32x32 potrf and of the FP64 primitives its
// critical path is made of (single warp, nothing else on the SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/potrf_bench tools/potrf_bench.cu
#include <cstdio>
#include <vector>
#include "../bayes-skopt_b200/csrc/bgp_chol.cu"

__global__ void lat_kernel(double* out, long long* clk, double seed) {
  double x = seed, y = seed * 0.5;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = fma(x, y, 1e-9);
  }
  long long t1 = clock64();
  double z = seed + 1.0;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) z = rsqrt(z) + 1.5;
  }
  long long t2 = clock64();
  double w = seed + 1.0;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) w = 1.0 / w + 1.5;
  }
  long long t3 = clock64();
  float f = (float)seed + 1.0f;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f = fmaf(f, 0.999f, 1e-3f);
  }
  long long t4 = clock64();
  double s = seed + 1.0;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) s = sqrt(s) + 1.5;
  }
  long long t5 = clock64();
  if (threadIdx.x == 0) {
    clk[0] = (t1 - t0) / 1024; clk[1] = (t2 - t1) / 256; clk[2] = (t3 - t2) / 256; clk[3] = (t4 - t3) / 1024;
    clk[4] = (t5 - t4) / 256;
  }
  out[threadIdx.x] = x + z + w + f + s;
}

__global__ void potrf_kernel(const double* Ain, double* Lout, long long* clk, int reps) {
  __shared__ double D[32 * bgp::PS];
  __shared__ __align__(16) double Lt[32 * bgp::LS];
  __shared__ __align__(16) double Wd[4 * 64];
  const int lane = threadIdx.x;
  long long best = 1LL << 60;
  for (int rep = 0; rep < reps; ++rep) {
    for (int e = lane; e < 1024; e += 32) {
      int r = e >> 5, c = e & 31;
      D[r * bgp::PS + c] = c <= r ? Ain[e] : 0.0;
    }
    __syncwarp();
    double logdet = 0.0;
    long long t0 = clock64();
    int f = bgp::warp_potrf32(D, Lt, Wd, lane, logdet, 32, clk + 8);
    long long t1 = clock64();
    if (t1 - t0 < best) best = t1 - t0;
    if (f) clk[2] = f;
  }
  if (lane == 0) clk[0] = best;
  for (int e = lane; e < 1024; e += 32) Lout[e] = Lt[(e >> 5) * bgp::LS + (e & 31)];
}

int main() {
  std::vector<double> A(1024), L(1024);
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) A[i * 32 + j] = (i == j ? 3.0 : 0.0) + exp(-0.05 * (i - j) * (i - j));
  double *dA, *dL, *dout; long long* dclk;
  cudaMalloc(&dA, 8192); cudaMalloc(&dL, 8192); cudaMalloc(&dout, 8192); cudaMalloc(&dclk, 512);
  cudaMemset(dclk, 0, 64);
  cudaMemcpy(dA, A.data(), 8192, cudaMemcpyHostToDevice);
  long long clk[64];
  lat_kernel<<<1, 32>>>(dout, dclk, 1.0);
  cudaMemcpy(clk, dclk, 64, cudaMemcpyDeviceToHost);
  printf("dependent latency (clk): DFMA %lld  rsqrt(double)+add %lld  1/x+add %lld  FFMA %lld  sqrt+add %lld\n", clk[0], clk[1], clk[2], clk[3], clk[4]);
  cudaMemset(dclk, 0, 512);
  potrf_kernel<<<1, 32>>>(dA, dL, dclk, 5);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(clk, dclk, 512, cudaMemcpyDeviceToHost);
  for (int i = 1; i < 28; ++i) printf("%s%lld", i % 6 == 1 ? "\n  " : " ", clk[8 + i] - clk[8 + i - 1]);
  printf("\n");
  cudaMemcpy(L.data(), dL, 8192, cudaMemcpyDeviceToHost);
  // check against a host Cholesky
  std::vector<double> R(A);
  for (int j = 0; j < 32; ++j) {
    for (int k = 0; k < j; ++k) for (int i = j; i < 32; ++i) R[i * 32 + j] -= R[i * 32 + k] * R[j * 32 + k];
    double d = sqrt(R[j * 32 + j]);
    for (int i = j; i < 32; ++i) R[i * 32 + j] /= d;
  }
  double err = 0;
  for (int i = 0; i < 32; ++i) for (int j = 0; j <= i; ++j) err = fmax(err, fabs(R[i * 32 + j] - L[i * 32 + j]));
  printf("warp_potrf32: %lld clk (best of 5), fail=%lld, max |L - L_host| = %.3e, cuda: %s\n", clk[0], clk[2], err, cudaGetErrorString(e));
  return 0;
}
