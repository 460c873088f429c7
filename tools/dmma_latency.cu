// Dependent-issue latency of DMMA.8x8x4 (accumulator chain, and result -> A operand chain), developer microbenchmark.
// nvcc -arch=sm_100a tools/dmma_latency.cu -o /tmp/dl && /tmp/dl
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__global__ void lat(long long* out, const double* in) {
  double c[2] = {in[threadIdx.x], in[threadIdx.x + 32]};
  const double a = in[threadIdx.x + 64], b = in[threadIdx.x + 96];
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) dmma(c, a, b);
  long long t1 = clock64();
  // result feeds the A operand of the next one
  double x = c[0];
#pragma unroll
  for (int i = 0; i < 64; ++i) { double o[2] = {0.0, 0.0}; dmma(o, x, b); x = o[0]; }
  long long t2 = clock64();
  // 4 independent accumulators
  double d[4][2] = {{x, x}, {a, a}, {b, b}, {c[0], c[1]}};
#pragma unroll
  for (int i = 0; i < 64; ++i) dmma(d[i & 3], a, b);
  long long t3 = clock64();
  if (threadIdx.x == 0) { out[0] = (t1 - t0) / 64; out[1] = (t2 - t1) / 64; out[2] = (t3 - t2) / 64; }
  if (x + d[0][0] + d[1][0] + d[2][0] + d[3][0] == 1.2345) out[3] = 1;
}
int main() {
  double* in; long long* out; cudaMalloc(&in, 1024); cudaMemset(in, 0, 1024); cudaMalloc(&out, 64);
  lat<<<1, 32>>>(out, in); lat<<<1, 32>>>(out, in);
  long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
  printf("DMMA.8x8x4 cycles per instruction: accumulator chain %lld, result->A chain %lld, 4 independent accumulators %lld\n", h[0], h[1], h[2]);
  return 0;
}
