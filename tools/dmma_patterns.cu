// DMMA.8x8x4 issue-rate under realistic operand patterns (developer microbenchmark).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// 4x4 register tile, distinct A/B registers, two dependent k-steps back to back (as chol v4)
template <int MODE>
__global__ void __launch_bounds__(256) k44(double* out, const double* in, int iters) {
  double acc[4][4][2];
  double a[4][2], b[4][2];
  for (int t = 0; t < 4; ++t) { a[t][0] = in[threadIdx.x + 32 * t]; a[t][1] = in[threadIdx.x + 7 * t];
                                b[t][0] = in[threadIdx.x + 64 + t]; b[t][1] = in[threadIdx.x + 99 + t]; }
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) { dmma(acc[t][u], a[t][0], b[u][0]); dmma(acc[t][u], a[t][1], b[u][1]); }
    } else {
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int u = 0; u < 4; ++u) dmma(acc[t][u], a[t][e], b[u][e]);
    }
    if (MODE == 2) {  // perturb operands so they cannot sit in the reuse cache
#pragma unroll
      for (int t = 0; t < 4; ++t) { a[t][0] += 1e-9; b[t][1] -= 1e-9; }
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u) s += acc[t][u][0] + acc[t][u][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
  double *out, *in; cudaMalloc(&out, 8 * sms * 512); cudaMalloc(&in, 8 * 1024); cudaMemset(in, 0, 8 * 1024);
  const int iters = 4000;
  for (int warps : {1, 2, 4, 8}) {
    double fl = 2.0 * 256 * 32 * (double)iters * warps * sms;
    float m0 = time_ms([&] { k44<0><<<sms, warps * 32>>>(out, in, iters); });
    float m1 = time_ms([&] { k44<1><<<sms, warps * 32>>>(out, in, iters); });
    float m2 = time_ms([&] { k44<2><<<sms, warps * 32>>>(out, in, iters); });
    printf("warps/SM %d: pairs-back-to-back %.2f TF | k-step-major %.2f TF | k-step-major + operand updates %.2f TF\n",
           warps, fl / m0 / 1e9, fl / m1 / 1e9, fl / m2 / 1e9);
  }
  return 0;
}
