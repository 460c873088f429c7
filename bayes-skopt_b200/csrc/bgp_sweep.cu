// K4: candidate sweep.  A persistent CTA takes (theta, 32-candidate) tiles and, per tile,
//   1. builds the cross-covariances k*(x_c, X) -- the B operand, 32 x n_pad -- either resident in
//      shared memory (n_pad <= ~544) or, for larger n, in an L2-resident scratch of its own in
//      128-column windows that come back through a 3-deep ring of TMA bulk copies
//      (cp.async.bulk + mbarrier complete_tx; the warp that finishes a window last re-arms it);
//   2. computes the whitened vectors v = L^-1 k* as a triangular GEMM on DMMA.8x8x4: a warp owns
//      a 16-row unit of L^-1 (rows of the L^-T part of the factor slab) and streams it from L2
//      through its private cp.async ring, 16-row units are dealt in passes of 16 units whose
//      direction alternates, so every warp gets the same number of 8-column steps whatever n is;
//   3. reduces  var = k(x,x) - |v|^2  and  mean = v . z  (z = L^-1 y)  in the epilogue, so only
//      16 bytes per (theta, candidate) reach HBM.
// Replaces skopt GaussianProcessRegressor.predict as called by bask/acquisition.py:121-129
// (einsum "ki,kj,ij->k" with the explicit K_inv_) and the cho_solve loops of PVRS / VR
// (bask/acquisition.py:285-339) through the optional extra right-hand sides.
#include <cstdlib>

#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace bgp {

constexpr int SW_NW = 16;                 // warps per CTA
constexpr int SW_NT = 4;                  // 8-candidate column tiles per CTA tile
constexpr int SW_NC = 8 * SW_NT;          // candidates per CTA tile
constexpr int SW_KC = 128;                // columns per window of the k* tile (windowed mode)
constexpr int SW_WS = SW_KC + 8;          // row stride of a window (== 8 mod 16: conflict-free LDS.128)
constexpr int SW_NB = 3;                  // window buffers
constexpr int SW_ST = 4;                  // stages of a warp's L^-T ring
constexpr int SW_STAGE_BYTES = 1024;      // 2 fragments x 32 lanes x 16 bytes
constexpr int SW_WIN_BYTES = SW_NC * SW_WS * 8;

struct SweepSmem {
  DevProgram prog;
  ThetaParams tp;
  double kss[SW_NC];                 // k(x_c, x_c)
  double sums[SW_NW][SW_NC][2];      // per-warp partial |v|^2 and v.z
  unsigned long long full[SW_NB];    // mbarriers: window buffer b holds the chunk of sequence number q, q % NB == b
  int cnt[SW_NB];                    // warps that are done with buffer b
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// one thread: arm the barrier of window buffer `buf` and start the bulk copy of chunk `cidx` of this
// CTA's k* scratch into it
__device__ __forceinline__ void window_load(SweepSmem& S, double* wins, const double* scratch, int buf, int cidx) {
  const unsigned bar = smem_u32(&S.full[buf]);
  const unsigned dst = smem_u32(wins + (size_t)buf * SW_NC * SW_WS);
  const double* src = scratch + (size_t)cidx * SW_NC * SW_WS;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(SW_WIN_BYTES) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(SW_WIN_BYTES), "r"(bar)
               : "memory");
}

// passes of 16 units from the bottom of L^-1: pass p covers units [lo, hi), hi = U - 16 p
__device__ __forceinline__ int pass_hi(int U, int p) { return U - SW_NW * p; }
__device__ __forceinline__ int pass_chunks(int U, int p) { return (16 * pass_hi(U, p) + SW_KC - 1) / SW_KC; }
// chunk of the k* tile that the window sequence number q (tile-local) carries
__device__ __forceinline__ int chunk_of_seq(int U, int q) {
  for (int p = 0;; ++p) {
    const int c = pass_chunks(U, p);
    if (q < c) return q;
    q -= c;
  }
}

template <bool WIN>
__global__ void __launch_bounds__(SW_NW * 32, 1) sweep_kernel(SweepArgs A) {
  constexpr int NT = SW_NT, NC = SW_NC, ST = SW_ST;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SweepSmem& S = *reinterpret_cast<SweepSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int n = A.n, d = A.d;
  const SlabGeom G = SlabGeom::make(n, true);
  const int P = G.P, npad = 32 * P, kstride = npad + 8, U = 2 * P;
  const int npass = (U + SW_NW - 1) / SW_NW;
  int tile_chunks = 0;                                     // windows per tile (windowed mode)
  if (WIN) for (int p = 0; p < npass; ++p) tile_chunks += pass_chunks(U, p);
  // dynamic shared memory: struct | window buffers or the resident k* tile | rings | Xs | dotx
  size_t off = (sizeof(SweepSmem) + 127) & ~size_t(127);
  double* Ks = reinterpret_cast<double*>(smem_raw + off);
  off += WIN ? (size_t)SW_NB * SW_WIN_BYTES : (size_t)NC * kstride * 8;
  off = (off + 127) & ~size_t(127);
  const unsigned ring = smem_u32(smem_raw + off) + warp * (ST * SW_STAGE_BYTES) + lane * 16;
  off += (size_t)SW_NW * ST * SW_STAGE_BYTES;
  double* Xs = reinterpret_cast<double*>(smem_raw + off);  // [leaf][dim][candidate]
  double* dotx = Xs + (size_t)(A.n_leaves > 0 ? A.n_leaves : 1) * NC * d;   // R x NC
  double* scratch = WIN ? A.ks_scratch + (size_t)blockIdx.x * A.ks_scratch_stride : nullptr;
  {
    const int* src = reinterpret_cast<const int*>(A.prog);
    int* dst = reinterpret_cast<int*>(&S.prog);
    for (int i = tid; i < (int)(sizeof(DevProgram) / 4); i += SW_NW * 32) dst[i] = src[i];
  }
  if (WIN && tid == 0) {
    for (int b = 0; b < SW_NB; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&S.full[b])) : "memory");
      S.cnt[b] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const DevProgram& PR = S.prog;
  const int tpt = (A.m + NC - 1) / NC;      // tiles per theta
  int s_loaded = -1;
  int seq_base = 0;                         // window sequence numbers used by earlier tiles

  for (int tile = blockIdx.x; tile < A.S * tpt; tile += gridDim.x) {
    const int s = tile / tpt, c0 = (tile - s * tpt) * NC;
    __syncthreads();                        // previous tile's epilogue has read sums / dotx / kss
    if (s != s_loaded) {
      resolve_theta(PR, A.theta + (size_t)s * PR.n_theta, A.fixed_ls, S.tp, tid, SW_NW * 32);
      s_loaded = s;
    }
    for (int e = tid; e < SW_NW * NC * 2; e += SW_NW * 32) (&S.sums[0][0][0])[e] = 0.0;
    for (int e = tid; e < A.R * NC; e += SW_NW * 32) dotx[e] = 0.0;
    __syncthreads();
    // scaled candidate coordinates per stationary leaf, candidate index fastest
    for (int e = tid; e < PR.n_leaves * d * NC; e += SW_NW * 32) {
      const int c = e % NC, kk = (e / NC) % d, l = e / (NC * d);
      const int ci = c0 + c;
      Xs[e] = (ci < A.m) ? A.Xc[(size_t)s * A.xc_stride + (size_t)ci * d + kk] * S.tp.inv_ls[l][kk] : 0.0;
    }
    __syncthreads();
    // ---- phase 0: cross-covariance tile.  A thread owns training row i and walks the candidates
    // four at a time (four independent sqrt/exp chains), its scaled coordinates come from L1
    const double* Xtr = A.X + (size_t)s * A.x_stride;
    for (int i = tid; i < npad; i += SW_NW * 32) {
      double* dst = WIN ? scratch + (size_t)(i / SW_KC) * NC * SW_WS + (i % SW_KC) : Ks + i;
      const int dstride = WIN ? SW_WS : kstride;
      if (i >= n) {
        for (int c = 0; c < NC; ++c) dst[(size_t)c * dstride] = 0.0;
        continue;
      }
      if (PR.n_leaves == 1) {
        const double* il = S.tp.inv_ls[0];
        for (int cb = 0; cb < NC; cb += 4) {
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          for (int kk = 0; kk < d; ++kk) {
            const double xi = __ldg(Xtr + (size_t)i * d + kk) * il[kk];
            const double4 xc = *reinterpret_cast<const double4*>(Xs + (size_t)kk * NC + cb);
            const double t0 = xi - xc.x, t1 = xi - xc.y, t2 = xi - xc.z, t3 = xi - xc.w;
            a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1); a2 = fma(t2, t2, a2); a3 = fma(t3, t3, a3);
          }
          double r2[BGP_MAX_LEAVES] = {0, 0, 0, 0};
          r2[0] = a0; const double v0 = eval_program(PR, S.tp, r2, false, true);
          r2[0] = a1; const double v1 = eval_program(PR, S.tp, r2, false, true);
          r2[0] = a2; const double v2 = eval_program(PR, S.tp, r2, false, true);
          r2[0] = a3; const double v3 = eval_program(PR, S.tp, r2, false, true);
          dst[(size_t)(cb + 0) * dstride] = (c0 + cb + 0 < A.m) ? v0 : 0.0;
          dst[(size_t)(cb + 1) * dstride] = (c0 + cb + 1 < A.m) ? v1 : 0.0;
          dst[(size_t)(cb + 2) * dstride] = (c0 + cb + 2 < A.m) ? v2 : 0.0;
          dst[(size_t)(cb + 3) * dstride] = (c0 + cb + 3 < A.m) ? v3 : 0.0;
        }
      } else {
        for (int c = 0; c < NC; ++c) {
          double r2[BGP_MAX_LEAVES];
#pragma unroll
          for (int l = 0; l < BGP_MAX_LEAVES; ++l) {
            r2[l] = 0.0;
            if (l < PR.n_leaves) {
              double acc = 0.0;
              for (int kk = 0; kk < d; ++kk) {
                const double t = __ldg(Xtr + (size_t)i * d + kk) * S.tp.inv_ls[l][kk] - Xs[((size_t)l * d + kk) * NC + c];
                acc = fma(t, t, acc);
              }
              r2[l] = acc;
            }
          }
          dst[(size_t)c * dstride] = (c0 + c < A.m) ? eval_program(PR, S.tp, r2, false, true) : 0.0;
        }
      }
    }
    if (tid < NC) {
      double r2[BGP_MAX_LEAVES] = {0, 0, 0, 0};
      S.kss[tid] = eval_program(PR, S.tp, r2, true, A.noise_off == 0);
    }
    if (WIN) asm volatile("fence.proxy.async;\n" ::: "memory");   // scratch writes -> the bulk copies below
    __syncthreads();
    if (WIN && tid == 0) {
      for (int b = 0; b < SW_NB && b < tile_chunks; ++b)
        window_load(S, Ks, scratch, (seq_base + b) % SW_NB, chunk_of_seq(U, b));
    }

    // ---- phase 1: v = L^-1 k*, units of 16 rows in passes of 16 units
    const double* slab = A.slabs + (size_t)s * G.doubles();
    const double* z = A.z + (size_t)s * n;
    int seq = seq_base;                      // window sequence number at the start of the pass
    for (int p = 0; p < npass; ++p) {
      const int hi = pass_hi(U, p), lo = hi - SW_NW > 0 ? hi - SW_NW : 0;
      const int u = (p & 1) ? lo + warp : hi - 1 - warp;
      const bool active = u >= lo && u < hi;
      const int steps = active ? 2 * (u + 1) : 0;
      const int nch = WIN ? pass_chunks(U, p) : 1;
      const int j = u >> 1, hf = u & 1;
      double acc[2][NT][2];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int v = 0; v < NT; ++v) acc[t][v][0] = acc[t][v][1] = 0.0;
      // A operand: (L^-1)[32j + 16hf + 2r + t][i] lives at aug_base(j) + 32 i + 16hf + 2r + t; a step
      // (8 columns) takes two 16-byte fragments per lane: columns 8st + 2q and 8st + 2q + 1
      const double* ap = slab + (active ? G.aug_base(j) + 16 * hf + 2 * r + 64 * q : 0);
      auto issue = [&](int sn, int stage) {
        if (sn < steps) {
          const double* src = ap + (size_t)256 * sn;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(ring + stage * SW_STAGE_BYTES), "l"(src));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(ring + stage * SW_STAGE_BYTES + 512),
                       "l"(src + 32));
        }
        asm volatile("cp.async.commit_group;\n" ::);
      };
#define BGP_LDS2(dst, addr) \
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"((dst).x), "=d"((dst).y) : "r"(addr))
      double2 av[2][2], bv[2][NT];
      if (active) {
#pragma unroll
        for (int s2 = 0; s2 < ST - 1; ++s2) issue(s2, s2);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(ST - 2));
        BGP_LDS2(av[0][0], ring);
        BGP_LDS2(av[0][1], ring + 512);
      }
      for (int ch = 0; ch < nch; ++ch) {
        // steps of this warp that fall into the window (all of them when the tile is resident)
        const int sb = WIN ? (SW_KC / 8) * ch : 0;
        const int se = WIN ? min(steps, (SW_KC / 8) * (ch + 1)) : steps;
        unsigned bb;                                   // B operand: row 8v + r of the window / tile
        if (WIN) {
          const int sq = seq + ch;
          mbar_wait(smem_u32(&S.full[sq % SW_NB]), (unsigned)((sq / SW_NB) & 1));
          bb = smem_u32(Ks + (size_t)(sq % SW_NB) * NC * SW_WS + (size_t)r * SW_WS + 2 * q);
        } else {
          bb = smem_u32(Ks + (size_t)r * kstride + 2 * q);
        }
        constexpr int BROW = 8 * 8 * (WIN ? SW_WS : 0);   // bytes between column tiles (windowed)
        const unsigned brow = WIN ? (unsigned)BROW : (unsigned)(64 * kstride);
        if (sb < se) {
#pragma unroll
          for (int v = 0; v < NT; ++v) BGP_LDS2(bv[0][v], bb + v * brow);   // first step of the window (sb is even)
          for (int st0 = sb; st0 < se; st0 += ST) {
#pragma unroll
            for (int s2 = 0; s2 < ST; ++s2) {
              const int st = st0 + s2;
              if (st < se) {
                // everything below is volatile asm, so this is the issue order: the loads of step
                // st + 1 and the refill of the stage consumed in step st - 1 sit between the DMMAs
                asm volatile("cp.async.wait_group %0;\n" ::"n"(ST - 3));
                const bool more_a = st + 1 < steps, more_b = st + 1 < se;
#pragma unroll
                for (int i = 0; i < 4 * NT; ++i) {
                  const int h = i / (2 * NT), t = (i / NT) & 1, v = i % NT;
                  dmma(acc[t][v], t ? av[s2 & 1][h].y : av[s2 & 1][h].x, h ? bv[s2 & 1][v].y : bv[s2 & 1][v].x);
                  if (i == 0 && more_a) BGP_LDS2(av[(s2 & 1) ^ 1][0], ring + ((s2 + 1) % ST) * SW_STAGE_BYTES);
                  if (i == 1 && more_a) BGP_LDS2(av[(s2 & 1) ^ 1][1], ring + ((s2 + 1) % ST) * SW_STAGE_BYTES + 512);
                  if (i == 2) issue(st + ST - 1, (s2 + ST - 1) % ST);
                  if (i >= 3 && i < 3 + NT && more_b)
                    BGP_LDS2(bv[(s2 & 1) ^ 1][i - 3], bb + (i - 3) * brow + 64 * (st + 1 - sb));
                }
              }
            }
          }
        }
        if (WIN) {
          // done with this window: the last warp to get here re-arms the buffer with the chunk that
          // is NB windows ahead
          __syncwarp();
          if (lane == 0) {
            const int sq = seq + ch, buf = sq % SW_NB;
            if (atomicAdd(&S.cnt[buf], 1) == SW_NW - 1) {
              S.cnt[buf] = 0;
              const int nq = sq - seq_base + SW_NB;
              if (nq < tile_chunks) {
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                window_load(S, Ks, scratch, buf, chunk_of_seq(U, nq));
              }
            }
          }
        }
      }
#undef BGP_LDS2
      asm volatile("cp.async.wait_all;\n" ::);
      seq += nch;
      if (!active) continue;
      // epilogue of this unit: rows 32j + 16hf + 2r + t
      const int row0 = 16 * u + 2 * r;
      double zr[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) zr[t] = (row0 + t < n) ? z[row0 + t] : 0.0;
#pragma unroll
      for (int v = 0; v < NT; ++v) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double pv = 0.0, pm = 0.0;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            pv = fma(acc[t][v][e], acc[t][v][e], pv);
            pm = fma(acc[t][v][e], zr[t], pm);
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            pv += __shfl_xor_sync(0xffffffffu, pv, o);
            pm += __shfl_xor_sync(0xffffffffu, pm, o);
          }
          if (r == 0) {
            S.sums[warp][8 * v + 2 * q + e][0] += pv;
            S.sums[warp][8 * v + 2 * q + e][1] += pm;
          }
        }
      }
      if (A.R > 0) {
        for (int rr = 0; rr < A.R; ++rr) {
          const double* ze = A.zextra + ((size_t)s * A.R + rr) * n;
          double ez[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) ez[t] = (row0 + t < n) ? ze[row0 + t] : 0.0;
#pragma unroll
          for (int v = 0; v < NT; ++v)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              double pe = fma(acc[0][v][e], ez[0], acc[1][v][e] * ez[1]);
#pragma unroll
              for (int o = 4; o < 32; o <<= 1) pe += __shfl_xor_sync(0xffffffffu, pe, o);
              if (r == 0) atomicAdd(&dotx[rr * NC + 8 * v + 2 * q + e], pe);
            }
        }
      }
      if (A.v_out) {
#pragma unroll
        for (int v = 0; v < NT; ++v)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int ci = c0 + 8 * v + 2 * q + e;
            if (ci < A.m) {
              double* dstv = A.v_out + ((size_t)s * A.m + ci) * A.v_ld + row0;
              *reinterpret_cast<double2*>(dstv) = make_double2(acc[0][v][e], acc[1][v][e]);
            }
          }
      }
    }
    seq_base += tile_chunks;
    __syncthreads();
    if (tid < NC && c0 + tid < A.m) {
      double vv = 0.0, mm = 0.0;
      for (int w = 0; w < SW_NW; ++w) { vv += S.sums[w][tid][0]; mm += S.sums[w][tid][1]; }
      double var = S.kss[tid] - vv;
      if (var < 0.0) var = 0.0;
      const size_t o = (size_t)s * A.m + c0 + tid;
      A.mu[o] = A.y_std * mm + A.y_mean;
      A.sd[o] = sqrt(var * A.y_std * A.y_std);
    }
    for (int e = tid; e < A.R * NC; e += SW_NW * 32) {
      const int rr = e / NC, c = e - rr * NC;
      if (c0 + c < A.m) A.dots[((size_t)s * A.R + rr) * A.m + c0 + c] = dotx[e];
    }
  }
}

static size_t sweep_smem(bool win, int n, int d, int R, int n_leaves) {
  const int P = (n + 31) / 32;
  size_t off = (sizeof(SweepSmem) + 127) & ~size_t(127);
  off += win ? (size_t)SW_NB * SW_WIN_BYTES : (size_t)SW_NC * (32 * P + 8) * 8;
  off = (off + 127) & ~size_t(127);
  off += (size_t)SW_NW * SW_ST * SW_STAGE_BYTES;
  off += sizeof(double) * ((size_t)(n_leaves > 0 ? n_leaves : 1) * SW_NC * d + (size_t)R * SW_NC);
  return off;
}

constexpr size_t SWEEP_SMEM_OPTIN = 227 * 1024;

// windowed mode keeps the k* tile of every resident CTA in a scratch of this many doubles
size_t sweep_scratch_doubles(int n) {
  const int P = (n + 31) / 32;
  return (size_t)((32 * P + SW_KC - 1) / SW_KC) * SW_NC * SW_WS;
}

bool sweep_is_windowed(int n, int d, int R, int n_leaves) {
  // BGP_SWEEP_WINDOWED forces the windowed path at any n (read per call: the parity tests flip it)
  return std::getenv("BGP_SWEEP_WINDOWED") != nullptr || sweep_smem(false, n, d, R, n_leaves) > SWEEP_SMEM_OPTIN;
}

cudaError_t prepare_sweep() {
  cudaError_t e = cudaFuncSetAttribute(sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM_OPTIN);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM_OPTIN);
  return e;
}

cudaError_t launch_sweep(const SweepArgs& A, int sms, cudaStream_t stream) {
  const bool win = sweep_is_windowed(A.n, A.d, A.R, A.n_leaves);
  const size_t smem = sweep_smem(win, A.n, A.d, A.R, A.n_leaves);
  if (smem > SWEEP_SMEM_OPTIN) return cudaErrorInvalidValue;
  if (win && !A.ks_scratch) return cudaErrorInvalidValue;
  const long long tiles = (long long)A.S * ((A.m + SW_NC - 1) / SW_NC);
  const int grid = (int)(tiles < sms ? tiles : sms);
  if (win) sweep_kernel<true><<<grid, SW_NW * 32, smem, stream>>>(A);
  else sweep_kernel<false><<<grid, SW_NW * 32, smem, stream>>>(A);
  return cudaGetLastError();
}

}  // namespace bgp
