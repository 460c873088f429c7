// Measures the FP64 denominators the rooflines in DESIGN.md use: DMMA.8x8x4 and DFMA issue
// rates (register-resident, no memory traffic) plus latency of a dependent DMMA chain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(512) dmma_rate(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(512) dfma_rate(double* out, int iters, double a, double b) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dexp_rate(double* out, int iters, double a) {
  double x = a + threadIdx.x * 1e-3, s = 0;
  for (int it = 0; it < iters; ++it) { s += exp(-x); x += 1e-6; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dsqrt_rate(double* out, int iters, double a) {
  double x = a + threadIdx.x * 1e-3, s = 0;
  for (int it = 0; it < iters; ++it) { s += sqrt(x); x += 1e-6; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32 > 512 ? 512 : warps * 32; int blocks = sms * (warps * 32 / threads);
    float ms = time_ms([&] { dmma_rate<16><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
    double fl = 2.0 * 256 * 16 * (double)iters * warps * sms;
    printf(", \"dmma_tflops_w%d\": %.2f", warps, fl / ms / 1e9);
  }
  {
    float ms = time_ms([&] { dmma_rate<1><<<sms, 32>>>(out, iters, 1.0000001, 0.9999999); });
    printf(", \"dmma_dep_chain_ns\": %.2f", ms * 1e6 / iters);
    ms = time_ms([&] { dmma_rate<4><<<sms, 32>>>(out, iters, 1.0000001, 0.9999999); });
    printf(", \"dmma_1warp_4acc_ns_per_mma\": %.2f", ms * 1e6 / iters / 4);
    ms = time_ms([&] { dmma_rate<16><<<sms, 32>>>(out, iters, 1.0000001, 0.9999999); });
    printf(", \"dmma_1warp_16acc_ns_per_mma\": %.2f", ms * 1e6 / iters / 16);
  }
  for (int warps : {8, 16, 32}) {
    int threads = 512 < warps * 32 ? 512 : warps * 32; int blocks = sms * (warps * 32 / threads);
    float ms = time_ms([&] { dfma_rate<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 16 * (double)iters * warps * 32 * sms;
    printf(", \"dfma_tflops_w%d\": %.2f", warps, fl / ms / 1e9);
  }
  {
    float ms = time_ms([&] { dfma_rate<1><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf(", \"dfma_dep_chain_ns\": %.2f", ms * 1e6 / iters);
    ms = time_ms([&] { dexp_rate<<<sms * 2, 512>>>(out, 2000, 0.5); });
    printf(", \"dexp_gops\": %.1f", 2000.0 * sms * 2 * 512 / ms / 1e6);
    ms = time_ms([&] { dsqrt_rate<<<sms * 2, 512>>>(out, 2000, 0.5); });
    printf(", \"dsqrt_gops\": %.1f", 2000.0 * sms * 2 * 512 / ms / 1e6);
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf(", \"clock_khz\": %d}\n", clk);
  return 0;
}
