// K3: the affine-invariant stretch move (Goodman & Weare; emcee 3.1.6 EnsembleSampler with
// StretchMove, as driven by bask/bayesgpr.py:510-530) on device.  Red/blue split, proposals
// and accept tests use a counter-based Philox-4x32-10 stream keyed by (seed; step, purpose,
// walker), so every rank of a multi-GPU job regenerates identical proposals without a
// broadcast and a CUDA graph of the whole run can be replayed with a new seed.
// The ensemble state is tiny (W x p doubles), so each of these kernels is a single CTA; the
// work is in the batched log-posterior kernel launched between propose and accept.
#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

constexpr uint32_t TAG_SPLIT = 0x53504c54u, TAG_PROP = 0x50524f50u, TAG_ACC = 0x41434350u;

__device__ __forceinline__ uint64_t seed_of(uint64_t seed, const uint64_t* seed_ptr) {
  return seed_ptr ? *seed_ptr : seed;
}

// colour[i] = 0 for a uniformly random subset of ceil(W/2) walkers, 1 for the rest
// (emcee: inds = arange(W) % 2; random.shuffle(inds)).
__global__ void split_kernel(int W, uint64_t seed, const uint64_t* seed_ptr, int step,
                             int32_t* __restrict__ colour) {
  extern __shared__ uint64_t keys[];
  const uint64_t sd = seed_of(seed, seed_ptr);
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    Philox4 r = philox4x32_10(sd, (uint32_t)step, TAG_SPLIT, (uint32_t)i, 0u);
    keys[i] = ((uint64_t)r.c[0] << 32) | r.c[1];
  }
  __syncthreads();
  const int n0 = (W + 1) / 2;
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    const uint64_t k = keys[i];
    int rank = 0;
    for (int j = 0; j < W; ++j) rank += (keys[j] < k) || (keys[j] == k && j < i);
    colour[i] = rank < n0 ? 0 : 1;
  }
}

// q_k = c - (c - s_k) z,  z = ((a-1)u+1)^2 / a,  c = random walker of the other colour;
// factors_k = (p-1) log z.  movers[k] = index of the k-th walker of colour `half`.
__global__ void propose_kernel(const double* __restrict__ pos, const int32_t* __restrict__ colour, int W,
                               int p, int half, double a, uint64_t seed, const uint64_t* seed_ptr,
                               int step, double* __restrict__ q, double* __restrict__ factors,
                               int32_t* __restrict__ movers) {
  extern __shared__ int32_t lists[];   // movers[W] | others[W] | counts[2]
  int32_t* mv = lists;
  int32_t* ot = lists + W;
  int32_t* cnt = lists + 2 * W;
  if (threadIdx.x == 0) { cnt[0] = 0; cnt[1] = 0; }
  __syncthreads();
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    const int ci = colour[i];
    int before = 0;
    for (int j = 0; j < i; ++j) before += (colour[j] == ci);
    if (ci == half) { mv[before] = i; atomicAdd(&cnt[0], 1); }
    else { ot[before] = i; atomicAdd(&cnt[1], 1); }
  }
  __syncthreads();
  const int ns = cnt[0], nc = cnt[1];
  const uint64_t sd = seed_of(seed, seed_ptr);
  for (int k = threadIdx.x; k < ns; k += blockDim.x) {
    Philox4 r = philox4x32_10(sd, (uint32_t)step, TAG_PROP + (uint32_t)half, (uint32_t)k, 0u);
    const double u = u01_from(r.c[0], r.c[1]);
    const double t = (a - 1.0) * u + 1.0;
    const double zz = t * t / a;
    int partner = (int)(u01_from(r.c[2], r.c[3]) * nc);
    if (partner >= nc) partner = nc - 1;
    const double* s = pos + (size_t)mv[k] * p;
    const double* c = pos + (size_t)ot[partner] * p;
    for (int e = 0; e < p; ++e) q[(size_t)k * p + e] = c[e] - (c[e] - s[e]) * zz;
    factors[k] = (p - 1.0) * log(zz);
    movers[k] = mv[k];
  }
  for (int k = ns + threadIdx.x; k < W; k += blockDim.x) movers[k] = -1;
}

// accept k iff factors_k + new_lp_k - lp[movers_k] > log u   (emcee RedBlueMove.propose)
__global__ void accept_kernel(double* __restrict__ pos, double* __restrict__ lp, const double* __restrict__ q,
                              const double* __restrict__ factors, const double* __restrict__ new_lp,
                              const int32_t* __restrict__ movers, int W, int p, int half, uint64_t seed,
                              const uint64_t* seed_ptr, int step, int32_t* __restrict__ accepted,
                              double* __restrict__ chain_step, double* __restrict__ lp_step) {
  const uint64_t sd = seed_of(seed, seed_ptr);
  for (int k = threadIdx.x; k < W; k += blockDim.x) {
    const int i = movers[k];
    if (i < 0) continue;
    Philox4 r = philox4x32_10(sd, (uint32_t)step, TAG_ACC + (uint32_t)half, (uint32_t)k, 0u);
    const double lnpdiff = factors[k] + new_lp[k] - lp[i];
    if (lnpdiff > log(u01_from(r.c[0], r.c[1]))) {
      for (int e = 0; e < p; ++e) pos[(size_t)i * p + e] = q[(size_t)k * p + e];
      lp[i] = new_lp[k];
      if (accepted) accepted[i] += 1;
    }
  }
  if (chain_step) {
    __syncthreads();
    for (int e = threadIdx.x; e < W * p; e += blockDim.x) chain_step[e] = pos[e];
    if (lp_step) for (int e = threadIdx.x; e < W; e += blockDim.x) lp_step[e] = lp[e];
  }
}

cudaError_t launch_split(int W, uint64_t seed, const uint64_t* seed_ptr, int step, int32_t* colour,
                         cudaStream_t stream) {
  split_kernel<<<1, 256, W * sizeof(uint64_t), stream>>>(W, seed, seed_ptr, step, colour);
  return cudaGetLastError();
}
cudaError_t launch_propose(const double* pos, const int32_t* colour, int W, int p, int half, double a,
                           uint64_t seed, const uint64_t* seed_ptr, int step, double* q, double* factors,
                           int32_t* movers, cudaStream_t stream) {
  propose_kernel<<<1, 256, (2 * W + 2) * sizeof(int32_t), stream>>>(pos, colour, W, p, half, a, seed,
                                                                      seed_ptr, step, q, factors, movers);
  return cudaGetLastError();
}
cudaError_t launch_accept(double* pos, double* lp, const double* q, const double* factors,
                          const double* new_lp, const int32_t* movers, int W, int p, int half,
                          uint64_t seed, const uint64_t* seed_ptr, int step, int32_t* accepted,
                          double* chain_step, double* lp_step, cudaStream_t stream) {
  accept_kernel<<<1, 256, 0, stream>>>(pos, lp, q, factors, new_lp, movers, W, p, half, seed, seed_ptr,
                                        step, accepted, chain_step, lp_step);
  return cudaGetLastError();
}

}  // namespace bgp
