// Accuracy of fast_exp_neg / fast_sqrt (bgp_common.cuh) against the CUDA library routines, in ulps
// (developer tool; build and run on the GPU box:
//  nvcc -arch=sm_100a -I bayes-skopt_b200/csrc -I include tools/fastmath_check.cu -o /tmp/fm && /tmp/fm)
#include <cstdio>
#include <cmath>
#include "bgp_common.cuh"

__device__ double ulps(double a, double b) {
  if (a == b) return 0.0;
  const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
  return fabs((double)(ia - ib));
}
__global__ void check(int n, double* out) {
  double me = 0.0, ms = 0.0, xe = 0.0, xs = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Philox4 r = philox4x32_10(12345ull, (uint32_t)i, 7u, 0u, 0u);
    const double u = u01_from(r.c[0], r.c[1]), v = u01_from(r.c[2], r.c[3]);
    // exp argument: log-uniform magnitudes 1e-12 .. 690; sqrt argument: 1e-280 .. 1e+8
    const double x = -exp(-27.6 + 34.1 * u);
    const double s = exp(-644.0 + 662.0 * v);
    const double de = ulps(fast_exp_neg(x), exp(x)), ds = ulps(fast_sqrt(s), sqrt(s));
    if (de > me) { me = de; xe = x; }
    if (ds > ms) { ms = ds; xs = s; }
  }
  // block max via atomics on the bit patterns (non-negative doubles order like integers)
  atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(me));
  atomicMax(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)__double_as_longlong(ms));
  if (me >= 2.0) out[2] = xe;
  if (ms >= 2.0) out[3] = xs;
}
__global__ void edge(double* out) {
  out[0] = fast_exp_neg(0.0); out[1] = fast_exp_neg(-0.0); out[2] = fast_exp_neg(-745.0); out[3] = fast_exp_neg(-1e300);
  out[4] = fast_sqrt(0.0); out[5] = fast_sqrt(4.0); out[6] = fast_sqrt(1e-320); out[7] = fast_exp_neg(-699.9) / exp(-699.9);
}
int main() {
  double* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  check<<<592, 256>>>(1 << 26, d);
  double h[8]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("max ulp error: exp %.0f (at %g)  sqrt %.0f (at %g)\n", h[0], h[2], h[1], h[3]);
  edge<<<1, 1>>>(d); cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("exp(0)=%.17g exp(-0)=%.17g exp(-745)=%g exp(-1e300)=%g sqrt(0)=%g sqrt(4)=%.17g sqrt(1e-320)=%g ratio(-699.9)=%.17g\n",
         h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  return (h[0] == 1.0 && h[5] == 2.0) ? 0 : 1;
}
