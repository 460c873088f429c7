// Microbenchmark of the DMMA inner loop shape used by libbgp (A streamed from global/L2 with
// 16-byte loads, B from shared memory, 4x4 accumulator tile) -- developer tooling.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_nv(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// MODE 0: A from global (prefetch 1), B from smem.  MODE 1: A and B from smem.  MODE 2: like 0 but
// non-volatile asm.  MODE 3: A global, B smem, no prefetch (load-use in the same step)
template <int MODE, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) loopk(double* out, const double* gA, int steps, int bstride) {
  extern __shared__ double Bs[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r = lane >> 2, q = lane & 3;
  for (int i = tid; i < 32 * bstride; i += NWARPS * 32) Bs[i] = 1e-3 * (i % 7);
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
  const double* ap[4];
#pragma unroll
  for (int t = 0; t < 4; ++t)
    ap[t] = gA + ((size_t)blockIdx.x * NWARPS + warp) * 32 * 512 + (size_t)(8 * t + r) * 512 + 2 * q;
  double2 nxt[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) nxt[t] = *reinterpret_cast<const double2*>(ap[t]);
  for (int st = 0; st < steps; ++st) {
    double2 av[4];
    const int c = (st & 63) * 8;
    if (MODE == 1) {
#pragma unroll
      for (int t = 0; t < 4; ++t) av[t] = *reinterpret_cast<const double2*>(Bs + (size_t)(8 * t + r) * bstride + ((c + 64) & 255) + 2 * q);
    } else if (MODE == 3) {
#pragma unroll
      for (int t = 0; t < 4; ++t) av[t] = *reinterpret_cast<const double2*>(ap[t] + c);
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) av[t] = nxt[t];
      const int cn = ((st + 1) & 63) * 8;
#pragma unroll
      for (int t = 0; t < 4; ++t) nxt[t] = *reinterpret_cast<const double2*>(ap[t] + cn);
    }
    double2 bv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) bv[u] = *reinterpret_cast<const double2*>(Bs + (size_t)(8 * u + r) * bstride + (c & 255) + 2 * q);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
#pragma unroll
      for (int u = 0; u < 4; ++u) { if (MODE == 2) dmma_nv(acc[t][u], av[t].x, bv[u].x); else dmma(acc[t][u], av[t].x, bv[u].x); }
#pragma unroll
      for (int u = 0; u < 4; ++u) { if (MODE == 2) dmma_nv(acc[t][u], av[t].y, bv[u].y); else dmma(acc[t][u], av[t].y, bv[u].y); }
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u) s += acc[t][u][0] + acc[t][u][1];
  out[blockIdx.x * NWARPS * 32 + tid] = s;
}
template <class F> float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int MODE, int NW> void run(const char* name, double* out, double* gA, int sms, int bstride) {
  const int steps = 4096; size_t smem = sizeof(double) * 32 * bstride;
  cudaFuncSetAttribute(loopk<MODE, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float ms = time_ms([&] { loopk<MODE, NW><<<sms, NW * 32, smem, 0>>>(out, gA, steps, bstride); });
  double fl = 2.0 * 256 * 32 * (double)steps * NW * sms;
  printf("%-44s warps %2d  bstride %4d: %6.2f TF\n", name, NW, bstride, fl / ms / 1e9);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
  double *out, *gA; cudaMalloc(&out, 8 * sms * 512); cudaMalloc(&gA, sizeof(double) * (size_t)sms * 16 * 32 * 512);
  cudaMemset(gA, 0, sizeof(double) * (size_t)sms * 16 * 32 * 512);
  run<0, 8>("A global (prefetch 1), B smem", out, gA, sms, 264);
  run<0, 8>("A global (prefetch 1), B smem", out, gA, sms, 488);
  run<0, 16>("A global (prefetch 1), B smem", out, gA, sms, 264);
  run<0, 4>("A global (prefetch 1), B smem", out, gA, sms, 264);
  run<1, 8>("A smem, B smem", out, gA, sms, 264);
  run<2, 8>("A global, B smem, non-volatile asm", out, gA, sms, 264);
  run<3, 8>("A global no prefetch, B smem", out, gA, sms, 264);
  run<0, 8>("A global (prefetch 1), B smem, stride 256", out, gA, sms, 256);
  return 0;
}
