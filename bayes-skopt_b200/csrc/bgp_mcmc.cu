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
// (emcee: inds = arange(W) % 2; random.shuffle(inds)): walker i gets colour 0 when the rank of its
// Philox key among all keys is below ceil(W/2).  Every CTA regenerates all W keys (cheap) and ranks
// SPLIT_WALKERS of them, four threads per walker each scanning a quarter of the keys, so the
// quadratic comparison count is spread over W / 64 CTAs (it was one CTA: 25 us at W = 1024).
constexpr int SPLIT_WALKERS = 64;
// step0 + blockIdx.y is the step, colour + blockIdx.y * W its row: one launch can colour every step of a run
__global__ void __launch_bounds__(256) split_kernel(int W, uint64_t seed, const uint64_t* seed_ptr, int step0,
                                                    int32_t* __restrict__ colour) {
  extern __shared__ uint64_t keys[];
  const uint64_t sd = seed_of(seed, seed_ptr);
  const int step = step0 + blockIdx.y;
  colour += (size_t)blockIdx.y * W;
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    Philox4 r = philox4x32_10(sd, (uint32_t)step, TAG_SPLIT, (uint32_t)i, 0u);
    keys[i] = ((uint64_t)r.c[0] << 32) | r.c[1];
  }
  __syncthreads();
  const int n0 = (W + 1) / 2;
  const int i = blockIdx.x * SPLIT_WALKERS + (threadIdx.x >> 2), part = threadIdx.x & 3;
  int rank = 0;
  if (i < W) {
    const uint64_t k = keys[i];
    for (int j = part; j < W; j += 4) rank += (keys[j] < k) || (keys[j] == k && j < i);
  }
  rank += __shfl_xor_sync(0xffffffffu, rank, 1);
  rank += __shfl_xor_sync(0xffffffffu, rank, 2);
  if (i < W && part == 0) colour[i] = rank < n0 ? 0 : 1;
}

// q_k = c - (c - s_k) z,  z = ((a-1)u+1)^2 / a,  c = random walker of the other colour;
// factors_k = (p-1) log z.  movers[k] = index of the k-th walker of colour `half`.
__device__ void propose_body(const double* __restrict__ pos, const int32_t* __restrict__ colour, int W,
                             int p, int half, double a, uint64_t sd, int step, double* __restrict__ q,
                             double* __restrict__ factors, int32_t* __restrict__ movers, int32_t* lists) {
  // lists: movers[W] | others[W] | counts[2]
  int32_t* mv = lists;
  int32_t* ot = lists + W;
  int32_t* cnt = lists + 2 * W;
  // position of walker i among the walkers of its colour: ballot prefix inside the warp, warp totals
  // through shared memory, a running base per block-sized chunk of walkers (O(W) instead of O(W^2))
  __shared__ int wz[32], wo[32];
  if (threadIdx.x == 0) { cnt[0] = 0; cnt[1] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int c0 = 0; c0 < W; c0 += blockDim.x) {
    const int i = c0 + threadIdx.x;
    const int ci = i < W ? colour[i] : -1;
    const unsigned m0 = __ballot_sync(0xffffffffu, ci == 0), m1 = __ballot_sync(0xffffffffu, ci == 1);
    if (lane == 0) { wz[wid] = __popc(m0); wo[wid] = __popc(m1); }
    __syncthreads();
    int bz = cnt[0], bo = cnt[1], tz = 0, to = 0;   // cnt[0]: walkers of colour 0 so far, cnt[1]: colour 1
    for (int w = 0; w < nwarps; ++w) {
      if (w < wid) { bz += wz[w]; bo += wo[w]; }
      tz += wz[w]; to += wo[w];
    }
    const unsigned lt = (1u << lane) - 1u;
    if (ci >= 0) {
      const int before = ci == 0 ? bz + __popc(m0 & lt) : bo + __popc(m1 & lt);
      if (ci == half) mv[before] = i; else ot[before] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) { cnt[0] += tz; cnt[1] += to; }
    __syncthreads();
  }
  const int ns = cnt[half], nc = cnt[1 - half];
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

__global__ void propose_kernel(const double* __restrict__ pos, const int32_t* __restrict__ colour, int W,
                               int p, int half, double a, uint64_t seed, const uint64_t* seed_ptr,
                               int step, double* __restrict__ q, double* __restrict__ factors,
                               int32_t* __restrict__ movers) {
  extern __shared__ int32_t lists[];
  propose_body(pos, colour, W, p, half, a, seed_of(seed, seed_ptr), step, q, factors, movers, lists);
}

// accept k iff factors_k + new_lp_k - lp[movers_k] > log u   (emcee RedBlueMove.propose)
__device__ void accept_body(double* __restrict__ pos, double* __restrict__ lp, const double* __restrict__ q,
                            const double* __restrict__ factors, const double* __restrict__ new_lp,
                            const int32_t* __restrict__ movers, int W, int p, int half, uint64_t sd, int step,
                            int32_t* __restrict__ accepted, double* __restrict__ chain_step,
                            double* __restrict__ lp_step) {
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

__global__ void accept_kernel(double* __restrict__ pos, double* __restrict__ lp, const double* __restrict__ q,
                              const double* __restrict__ factors, const double* __restrict__ new_lp,
                              const int32_t* __restrict__ movers, int W, int p, int half, uint64_t seed,
                              const uint64_t* seed_ptr, int step, int32_t* __restrict__ accepted,
                              double* __restrict__ chain_step, double* __restrict__ lp_step) {
  accept_body(pos, lp, q, factors, new_lp, movers, W, p, half, seed_of(seed, seed_ptr), step, accepted, chain_step,
              lp_step);
}

// The graph of a whole run (bgp_mcmc_run) uses these instead: the accept test of half step (step, half) and the
// proposals of the next one in ONE single-CTA launch (the colours of every step come from one split launch up
// front), i.e. 2 T + 2 small launches per run instead of 5 T.  Same arithmetic, same Philox keys: the chain is
// the one the separate kernels produce.
struct NextHalf {
  const int32_t* colour;   // row of the next half step's step; null: no further proposal
  int half, step;
  double a;
};
__global__ void accept_propose_kernel(double* __restrict__ pos, double* __restrict__ lp, double* __restrict__ q,
                                      double* __restrict__ factors, const double* __restrict__ new_lp,
                                      int32_t* __restrict__ movers, int W, int p, int half,
                                      const uint64_t* seed_ptr, int step, int32_t* __restrict__ accepted,
                                      double* __restrict__ chain_step, double* __restrict__ lp_step, NextHalf nx) {
  extern __shared__ int32_t lists[];
  const uint64_t sd = *seed_ptr;
  accept_body(pos, lp, q, factors, new_lp, movers, W, p, half, sd, step, accepted, chain_step, lp_step);
  if (nx.colour) {
    __syncthreads();
    propose_body(pos, nx.colour, W, p, nx.half, nx.a, sd, nx.step, q, factors, movers, lists);
  }
}

cudaError_t prepare_mcmc_xchg();

// split_kernel keeps W 64-bit sort keys in dynamic shared memory: opt in above the 48 KB default so
// that the W <= 8192 the API accepts really launches (outside stream capture, from bgp_create)
cudaError_t prepare_mcmc() {
  cudaError_t e = cudaFuncSetAttribute(split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * (int)sizeof(uint64_t));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(propose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (2 * 8192 + 2) * (int)sizeof(int32_t));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(accept_propose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (2 * 8192 + 2) * (int)sizeof(int32_t));
  if (e == cudaSuccess) e = prepare_mcmc_xchg();
  return e;
}

cudaError_t launch_split(int W, uint64_t seed, const uint64_t* seed_ptr, int step, int32_t* colour,
                         cudaStream_t stream) {
  split_kernel<<<(W + SPLIT_WALKERS - 1) / SPLIT_WALKERS, 256, W * sizeof(uint64_t), stream>>>(W, seed, seed_ptr, step, colour);
  return cudaGetLastError();
}
// colours of steps [0, T): colour[t * W + i]
cudaError_t launch_split_all(int W, int T, const uint64_t* seed_ptr, int32_t* colour, cudaStream_t stream) {
  if (T <= 0) return cudaSuccess;
  split_kernel<<<dim3((W + SPLIT_WALKERS - 1) / SPLIT_WALKERS, T), 256, W * sizeof(uint64_t), stream>>>(W, 0, seed_ptr, 0, colour);
  return cudaGetLastError();
}
cudaError_t launch_accept_propose(double* pos, double* lp, double* q, double* factors, const double* new_lp,
                                  int32_t* movers, int W, int p, int half, const uint64_t* seed_ptr, int step,
                                  int32_t* accepted, double* chain_step, double* lp_step, const int32_t* next_colour,
                                  int next_half, int next_step, double a, cudaStream_t stream) {
  NextHalf nx{next_colour, next_half, next_step, a};
  accept_propose_kernel<<<1, W <= 256 ? 256 : 512, (2 * W + 2) * sizeof(int32_t), stream>>>(
      pos, lp, q, factors, new_lp, movers, W, p, half, seed_ptr, step, accepted, chain_step, lp_step, nx);
  return cudaGetLastError();
}
cudaError_t launch_propose(const double* pos, const int32_t* colour, int W, int p, int half, double a,
                           uint64_t seed, const uint64_t* seed_ptr, int step, double* q, double* factors,
                           int32_t* movers, cudaStream_t stream) {
  propose_kernel<<<1, W <= 256 ? 256 : 512, (2 * W + 2) * sizeof(int32_t), stream>>>(pos, colour, W, p, half, a, seed,
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

// ---------------------------------------------------------------------------- multi-GPU
// Walker sharding without the host: the log-probabilities of a half step are exchanged by
// plain stores into every peer's exchange block (cudaIpc-mapped, NVLink P2P through NVSwitch) plus a
// release flag per source rank; the accept kernel acquires on the flags.  No NCCL call and no host
// round trip sits between propose and accept, so the whole sharded run is one CUDA graph per rank.
//
// Exchange block of a rank (bgp_peer_export allocates it with cudaMalloc; every peer maps it):
//   [0, 2 cap)      doubles   two value buffers (epoch parity): slot i = log-prob of proposal i
//   then 8 + 8 u64            flags[src] = last epoch whose slice from rank `src` is complete;
//                             word 8 = this rank's own epoch counter, word 9 = time-out flag,
//                             words 10 / 11 = nanoseconds spent in exchanges / their number (tooling)
// A rank may run at most one epoch ahead of a peer (it cannot pass the wait of epoch e + 1 before the
// peer has published e + 1, which the peer does after it finished reading the buffer of epoch e), so
// two buffers are enough.
// Replaces the `ncclAllGather of (W/2G) log-probs` of SURVEY.md 8(e) / emcee's serial map over walkers
// (bask/bayesgpr.py:510-530).


__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}

// Publishes src[0, cnt) as slots [lo, lo + cnt) of the current epoch to every rank, waits until every
// rank's slice has arrived here, and returns this rank's complete buffer.  Called by all threads of a
// single-CTA kernel.
__device__ const double* peer_exchange(const PeerXchg& X, const double* __restrict__ src, int lo, int cnt) {
  __shared__ unsigned long long ep_s;
  const int tid = threadIdx.x, nt = blockDim.x;
  unsigned long long* my_words = reinterpret_cast<unsigned long long*>(X.block[X.rank] + 2 * (size_t)X.cap);
  unsigned long long t_enter = 0;
  if (tid == 0) { ep_s = my_words[8] + 1; t_enter = global_ns(); }
  __syncthreads();
  const unsigned long long ep = ep_s;
  const size_t par = (size_t)(ep & 1) * X.cap;
  for (int peer = 0; peer < X.world; ++peer) {
    double* dst = X.block[peer] + par + lo;
    for (int i = tid; i < cnt; i += nt) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < X.world) {
    unsigned long long* peer_words = reinterpret_cast<unsigned long long*>(X.block[tid] + 2 * (size_t)X.cap);
    st_release_sys(peer_words + X.rank, ep);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(my_words + tid) < ep) {
      if (global_ns() - t0 > 10000000000ULL) { my_words[9] = 1; break; }   // 10 s: a peer died; do not hang the GPU
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (tid == 0) {
    my_words[8] = ep;
    my_words[10] += global_ns() - t_enter;   // developer counters: time spent in exchanges, their number
    my_words[11] += 1;
  }
  return X.block[X.rank] + par;
}

// all W initial log-probs: lp[i] = exchanged value i
__global__ void xchg_gather_kernel(PeerXchg X, const double* __restrict__ src, int lo, int cnt, int total,
                                   double* __restrict__ out) {
  const double* all = peer_exchange(X, src, lo, cnt);
  for (int i = threadIdx.x; i < total; i += blockDim.x) out[i] = all[i];
}

// accept with the exchange in front (new_lp_local holds this rank's slice [lo, lo + cnt) of the half step), then
// the next half step's proposals
__global__ void accept_xchg_propose_kernel(PeerXchg X, double* __restrict__ pos, double* __restrict__ lp,
                                           double* __restrict__ q, double* __restrict__ factors,
                                           const double* __restrict__ new_lp_local, int lo, int cnt,
                                           int32_t* __restrict__ movers, int W, int p, int half,
                                           const uint64_t* seed_ptr, int step, int32_t* __restrict__ accepted,
                                           double* __restrict__ chain_step, double* __restrict__ lp_step, NextHalf nx) {
  extern __shared__ int32_t lists[];
  const uint64_t sd = *seed_ptr;
  const double* new_lp = peer_exchange(X, new_lp_local, lo, cnt);
  accept_body(pos, lp, q, factors, new_lp, movers, W, p, half, sd, step, accepted, chain_step, lp_step);
  if (nx.colour) {
    __syncthreads();
    propose_body(pos, nx.colour, W, p, nx.half, nx.a, sd, nx.step, q, factors, movers, lists);
  }
}
cudaError_t prepare_mcmc_xchg() {
  return cudaFuncSetAttribute(accept_xchg_propose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (2 * 8192 + 2) * (int)sizeof(int32_t));
}
cudaError_t launch_accept_xchg_propose(const PeerXchg& X, double* pos, double* lp, double* q, double* factors,
                                       const double* new_lp_local, int lo, int cnt, int32_t* movers, int W, int p,
                                       int half, const uint64_t* seed_ptr, int step, int32_t* accepted,
                                       double* chain_step, double* lp_step, const int32_t* next_colour, int next_half,
                                       int next_step, double a, cudaStream_t stream) {
  NextHalf nx{next_colour, next_half, next_step, a};
  accept_xchg_propose_kernel<<<1, W <= 256 ? 256 : 512, (2 * W + 2) * sizeof(int32_t), stream>>>(
      X, pos, lp, q, factors, new_lp_local, lo, cnt, movers, W, p, half, seed_ptr, step, accepted, chain_step, lp_step, nx);
  return cudaGetLastError();
}

cudaError_t launch_xchg_gather(const PeerXchg& X, const double* src, int lo, int cnt, int total, double* out,
                               cudaStream_t stream) {
  xchg_gather_kernel<<<1, 256, 0, stream>>>(X, src, lo, cnt, total, out);
  return cudaGetLastError();
}


}  // namespace bgp
