// Acquisition epilogues on the (theta-sample x candidate) predictive moments and the
// finite-guarded mean over theta samples.  Replaces bask/acquisition.py:112-141 (loop +
// `if np.all(np.isfinite(tmp_out)): acq_output[j] += tmp_out / n_samples`) and the
// UncertaintyAcquisition classes: ExpectedImprovement (:154-172), TopTwoEI (:175-194),
// Expectation (:197-201), LCB (:204-216), MaxValueSearch (:219-267).
#include "bgp_common.cuh"
#include "bgp_internal.h"
#include "bgp_mes_table.inc"

namespace bgp {

constexpr int MES_PTS = 192;      // trial points evaluated per refinement round (3 x 64)
constexpr int MES_CH = 512;       // candidates per block in the quantile search
constexpr int MES_KL = 8;         // lanes that share one candidate in the MES epilogue
constexpr int MES_PCH = 8;        // trial points per block (blockIdx.z)
constexpr int MES_NEWTON = 9;        // safeguarded Newton from a 1/63 bracket: quadratic, converged after ~5
constexpr int ST = 8;             // doubles of per-theta statistics
constexpr int MS = 16;            // doubles of per-theta MES search state

__device__ __forceinline__ double norm_pdf(double x) { return 0.3989422804014327 * exp(-0.5 * x * x); }
__device__ __forceinline__ double norm_cdf(double x) {
  return x > 0.0 ? 1.0 - 0.5 * erfc(x * 0.7071067811865476) : 0.5 * erfc(-x * 0.7071067811865476);
}
__device__ __forceinline__ double log_ndtr(double x) {
  if (x > 0.0) return log1p(-0.5 * erfc(x * 0.7071067811865476));
  const double t = -x * 0.7071067811865476;
  if (t < 1.0) return log(0.5 * erfc(t));
  return log(0.5 * erfcx(t)) - t * t;
}
// phi(x)/Phi(x), stable in the lower tail
__device__ __forceinline__ double hazard_lower(double x) {
  if (x > 0.0) return norm_pdf(x) / (1.0 - 0.5 * erfc(x * 0.7071067811865476));
  return 0.7978845608028654 / erfcx(-x * 0.7071067811865476);
}
__device__ __forceinline__ double ei_f(double x) { return x * norm_cdf(x) + norm_pdf(x); }

__device__ __forceinline__ bool better_max(double v, long long i, double bv, long long bi) {
  // numpy argmax: NaN wins, first occurrence wins ties
  const bool vn = isnan(v), bn = isnan(bv);
  if (vn != bn) return vn;
  if (vn) return i < bi;
  return v > bv || (v == bv && i < bi);
}

__device__ void block_argmax(double& v, long long& i, double* sv, long long* si) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    long long oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (better_max(ov, oi, v, i)) { v = ov; i = oi; }
  }
  if (lane == 0) { sv[warp] = v; si[warp] = i; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    v = lane < nw ? sv[lane] : -INFINITY;
    i = lane < nw ? si[lane] : (1LL << 62);
    if (lane >= nw) v = -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, v, o);
      long long oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (better_max(ov, oi, v, i)) { v = ov; i = oi; }
    }
  }
}

// stats[s] = {min mu, left, right, all-finite(mu,sd)}  (left/right: bask/acquisition.py:239-241)
__global__ void acq_stats_kernel(const double* __restrict__ mu, const double* __restrict__ sd, int m,
                                 double* __restrict__ stats, double* __restrict__ pts) {
  const int s = blockIdx.x;
  double mn = INFINITY, lf = INFINITY, rt = -INFINITY;
  int fin = 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double a = mu[(size_t)s * m + i], b = sd[(size_t)s * m + i];
    mn = fmin(mn, a); lf = fmin(lf, -a - 3.0 * b); rt = fmax(rt, -a + 5.0 * b);
    fin &= (isfinite(a) && isfinite(b));
  }
  __shared__ double r0[32], r1[32], r2[32];
  __shared__ int r3[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    lf = fmin(lf, __shfl_xor_sync(0xffffffffu, lf, o));
    rt = fmax(rt, __shfl_xor_sync(0xffffffffu, rt, o));
    fin &= __shfl_xor_sync(0xffffffffu, fin, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { r0[warp] = mn; r1[warp] = lf; r2[warp] = rt; r3[warp] = fin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      mn = fmin(mn, r0[w]); lf = fmin(lf, r1[w]); rt = fmax(rt, r2[w]); fin &= r3[w];
    }
    stats[s * ST + 0] = mn; stats[s * ST + 1] = lf; stats[s * ST + 2] = rt; stats[s * ST + 3] = fin;
    r0[0] = lf; r1[0] = rt;
  }
  __syncthreads();
  if (pts) {
    const double l = r0[0], h = r1[0];
    for (int p = threadIdx.x; p < 64; p += blockDim.x)
      pts[s * MES_PTS + p] = p == 63 ? h : l + (h - l) * (p / 63.0);
  }
}

// y_opt: explicit scalar (y_opt_in not NaN), else per-theta array yopt (stride ystride doubles)
__global__ void ei_kernel(const double* __restrict__ mu, const double* __restrict__ sd, int m,
                          const double* __restrict__ yopt, int ystride, double y_opt_in,
                          double* __restrict__ out) {
  const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double y_opt = isnan(y_opt_in) ? yopt[s * ystride] : y_opt_in;
  const double a = mu[(size_t)s * m + i], b = sd[(size_t)s * m + i];
  out[(size_t)s * m + i] = (b > 0.0) ? ei_f((y_opt - a) / b) * b : 0.0;
}

__global__ void argmax_rows_kernel(const double* __restrict__ v, int m, long long* __restrict__ idx) {
  const int s = blockIdx.x;
  double bv = -INFINITY;
  long long bi = 1LL << 62;
  for (long long i = threadIdx.x; i < m; i += blockDim.x) {
    const double x = v[(size_t)s * m + i];
    if (better_max(x, i, bv, bi)) { bv = x; bi = i; }
  }
  __shared__ double sv[32];
  __shared__ long long si[32];
  block_argmax(bv, bi, sv, si);
  if (threadIdx.x == 0) idx[s] = bi;
}

// ref[s] = {ei max, global index, mu, sd} of the EI-maximising candidate of theta s
__global__ void ttei_ref_kernel(const double* __restrict__ mu, const double* __restrict__ sd,
                                const double* __restrict__ ei, int m, const long long* __restrict__ imax,
                                long long index_offset, double* __restrict__ ref) {
  const int s = threadIdx.x;
  const long long j = imax[s];
  ref[s * 4 + 0] = ei[(size_t)s * m + j];
  ref[s * 4 + 1] = (double)(j + index_offset);
  ref[s * 4 + 2] = mu[(size_t)s * m + j];
  ref[s * 4 + 3] = sd[(size_t)s * m + j];
}

__global__ void ttei_kernel(const double* __restrict__ mu, const double* __restrict__ sd, int m,
                            const double* __restrict__ ref, double* __restrict__ out) {
  const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double a = mu[(size_t)s * m + i], b = sd[(size_t)s * m + i];
  const double aj = ref[s * 4 + 2], bj = ref[s * 4 + 3];
  double v = 0.0;
  if (b > 0.0) {
    const double outer = sqrt(b * b + bj * bj);
    v = outer * ei_f((aj - a) / outer);
  }
  out[(size_t)s * m + i] = v;
}

__global__ void mean_lcb_kernel(const double* __restrict__ mu, const double* __restrict__ sd, int m,
                                int kind, double alpha, double* __restrict__ out) {
  const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double a = mu[(size_t)s * m + i], b = sd[(size_t)s * m + i];
  double v;
  if (kind == BGP_ACQ_MEAN) v = -a;
  else v = isinf(alpha) ? b : alpha * b - a;
  out[(size_t)s * m + i] = v;
}

// One safeguarded Newton step of the three quantile searches of theta s (lanes 0..2 of one warp), from the block
// partial sums of mes_eval_kernel, summed in block order; `final`: also the Gumbel fit
// (bask/acquisition.py:251-252).  Runs in the last CTA of a theta to finish its evaluation (the partial sums of
// the other CTAs are read past L1).
__device__ void mes_newton_step(int s, int nblk, int lane, double* __restrict__ pts, const double* gpart,
                                const double* dpart, double* __restrict__ st, bool final, double* __restrict__ fit) {
  double* S = st + s * MS;   // lo[0..2], hi[3..5], x[6..8]
  const double tau[3] = {-1.3862943611198906, -0.6931471805599453, -0.2876820724517809};
  if (lane < 3) {
    const int q = lane;
    double g = 0.0, dg = 0.0;
#pragma unroll 8
    for (int k = 0; k < nblk; ++k) {
      g += __ldcg(gpart + ((size_t)s * nblk + k) * MES_PTS + q);
      dg += __ldcg(dpart + ((size_t)s * nblk + k) * MES_PTS + q);
    }
    double lo = S[q], hi = S[3 + q];
    const double x = S[6 + q];
    if (g <= tau[q]) lo = x; else hi = x;
    double xn = x + (tau[q] - g) / dg;
    if (!(xn >= lo && xn <= hi)) xn = 0.5 * (lo + hi);   // NaN or outside the bracket: bisect
    S[q] = lo; S[3 + q] = hi; S[6 + q] = xn;
    pts[s * MES_PTS + q] = xn;
  }
  __syncwarp();
  if (final && lane == 0) {
    const double q1 = S[6], med = S[7], q2 = S[8];
    const double beta = (q1 - q2) / (log(log(4.0 / 3.0)) - log(log(4.0)));
    const double alpha = med + beta * log(log(2.0));
    S[9] = alpha; S[10] = beta;
    if (fit) { fit[s * 5 + 0] = alpha; fit[s * 5 + 1] = beta; fit[s * 5 + 2] = q1; fit[s * 5 + 3] = med; fit[s * 5 + 4] = q2; }
  }
}

// g(x) = sum_i log Phi((x + mu_i)/sd_i) (and optionally g') at npts trial points per theta;
// deterministic two-stage reduction: part[s][blk][p]
// newton: 0 = just the partial sums; 1 / 2 = the last CTA of a theta (a ticket counter per theta) also does
// the Newton step (2: and the Gumbel fit), so a refinement round is one launch
__global__ void mes_eval_kernel(const double* __restrict__ mu, const double* __restrict__ sd, int m,
                                double* __restrict__ pts, int npts, int with_grad,
                                double* __restrict__ gpart, double* __restrict__ dpart, int newton,
                                double* __restrict__ st, double* __restrict__ fit, int* __restrict__ ticket) {
  const int s = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  constexpr int PER = MES_CH / 256;
  // log Phi by the piecewise polynomials of bgp_mes_table.inc (2e-16 relative): exp(-t^2/2) R2(t) on the upper
  // side, -t^2/2 + L(-t) on the lower side -- a branch-free exp and a Horner scheme instead of erfc + log1p /
  // erfcx + log, and one reciprocal per candidate instead of a division per trial point
  __shared__ double tabp[BGP_MES_TAB_INTERVALS * BGP_MES_TAB_COEFS], tabn[BGP_MES_TAB_INTERVALS * BGP_MES_TAB_COEFS];
  for (int j = threadIdx.x; j < BGP_MES_TAB_INTERVALS * BGP_MES_TAB_COEFS; j += 256) {
    tabp[j] = BGP_NLOGCDF_POS_TAB[j];
    tabn[j] = BGP_LOGCDF_NEG_TAB[j];
  }
  double mean[PER], sdv[PER], inv[PER];
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int i = blk * MES_CH + e * 256 + threadIdx.x;
    if (i < m) { mean[e] = -mu[(size_t)s * m + i]; sdv[e] = sd[(size_t)s * m + i]; }
    else { mean[e] = 0.0; sdv[e] = -1.0; }   // sd < 0 marks padding
    inv[e] = 1.0 / sdv[e];
    if (!(sdv[e] > 1e-300 && isfinite(inv[e]))) inv[e] = 0.0;   // 0 marks "divide" (sd = 0: the reference's inf / NaN)
  }
  __shared__ double rg[MES_PCH][8], rd[MES_PCH][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  // the trial points are spread over blockIdx.z in chunks of MES_PCH: with m = 10^4 a (block, theta) grid alone
  // is 200 CTAs and leaves most of the chip idle for 192 sequential points
  const int p_begin = blockIdx.z * MES_PCH, p_end = min(npts, p_begin + MES_PCH);
  // partial sums of all (<= MES_PCH) points of this CTA stay in registers; one block reduction at the end
  double g[MES_PCH], dg[MES_PCH];
#pragma unroll
  for (int pp = 0; pp < MES_PCH; ++pp) {
    g[pp] = 0.0; dg[pp] = 0.0;
    const int p = p_begin + pp;
    if (p >= p_end) continue;
    const double x = pts[s * MES_PTS + p];
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      if (sdv[e] >= 0.0) {
        const double t = inv[e] != 0.0 ? (x - mean[e]) * inv[e] : (x - mean[e]) / sdv[e];
        const double u = fabs(t);
        if (u < 0.5 * BGP_MES_TAB_INTERVALS) {
          const int iv = (int)(u * 2.0);
          const double xx = fma(4.0, u, -(double)(2 * iv + 1));
          const double* c = (t >= 0.0 ? tabp : tabn) + iv * BGP_MES_TAB_COEFS;
          double R = c[0];
#pragma unroll
          for (int j = 1; j < BGP_MES_TAB_COEFS; ++j) R = fma(R, xx, c[j]);
          g[pp] += t >= 0.0 ? -R * fast_exp_neg(-0.5 * t * t) : fma(-0.5 * t, t, R);
        } else {
          g[pp] += log_ndtr(t);
        }
        if (with_grad) dg[pp] += hazard_lower(t) / sdv[e];
      }
    }
  }
#pragma unroll
  for (int pp = 0; pp < MES_PCH; ++pp) {
    g[pp] = warp_sum(g[pp]);
    if (with_grad) dg[pp] = warp_sum(dg[pp]);
    if (lane == 0) { rg[pp][warp] = g[pp]; rd[pp][warp] = dg[pp]; }
  }
  __syncthreads();
  if (threadIdx.x < MES_PCH && p_begin + (int)threadIdx.x < p_end) {
    const int pp = threadIdx.x, p = p_begin + pp;
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) { a += rg[pp][w]; b += rd[pp][w]; }
    gpart[((size_t)s * nblk + blk) * MES_PTS + p] = a;
    if (with_grad) dpart[((size_t)s * nblk + blk) * MES_PTS + p] = b;
  }
  if (newton) {   // (gridDim.z == 1 here: three points)
    __shared__ int last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int t = atomicAdd(ticket + s, 1);
      last = t == nblk - 1;
      if (last) ticket[s] = 0;
    }
    __syncthreads();
    if (last && warp == 0) {
      __threadfence();
      mes_newton_step(s, nblk, lane, pts, gpart, dpart, st, newton == 2, fit);
    }
  }
}

// brackets the three quantiles on the 64-point grid and starts the Newton iteration at the bracket midpoints
__global__ void mes_bracket_kernel(int nblk, double* __restrict__ pts, const double* __restrict__ gpart,
                                   double* __restrict__ st, int* __restrict__ ticket) {
  const int s = blockIdx.x, lane = threadIdx.x;
  __shared__ double g[64], x0[64];
  for (int p = lane; p < 64; p += 32) {
    double a = 0.0;
#pragma unroll 8
    for (int k = 0; k < nblk; ++k) a += gpart[((size_t)s * nblk + k) * MES_PTS + p];   // block order
    g[p] = a; x0[p] = pts[s * MES_PTS + p];
  }
  __syncwarp();
  double* S = st + s * MS;   // lo[0..2], hi[3..5], x[6..8]
  const double tau[3] = {-1.3862943611198906, -0.6931471805599453, -0.2876820724517809};
  if (lane < 3) {
    const int q = lane;
    int lo = 0;
    for (int p = 0; p < 64; ++p) if (g[p] <= tau[q]) lo = p;   // g is non-decreasing
    if (lo > 62) lo = 62;
    const double xl = x0[lo], xh = x0[lo + 1];
    S[q] = xl; S[3 + q] = xh; S[6 + q] = 0.5 * (xl + xh);
  }
  __syncwarp();
  if (lane < 3) pts[s * MES_PTS + lane] = S[6 + lane];
  if (lane == 0) ticket[s] = 0;
}

// mean_k [ gamma phi(gamma) / (2 Phi(gamma)) - log Phi(gamma) ],  gamma = (maxv_k + mu)/sd
//
// A CTA sorts the K Gumbel variates of its theta once (bitonic, shared memory) and then walks over chunks of
// 32 candidates.  With the draws in ascending order the terms of a candidate fall monotonically once
// gamma > 1 (gamma phi(gamma) and 1 - Phi(gamma) both decrease), super-exponentially so: a lane stops as soon
// as a term is below 1e-18 of its running sum -- everything it skips adds less than K * 1e-18 relative -- and
// adjacent lanes (adjacent draws) take the same branch.  A NaN sum never stops early (the comparison is false),
// an infinite one stays infinite, and the most negative gammas, where the reference's non-finite results come
// from, are visited first -- a row that is non-finite in the reference is non-finite here.
constexpr int MES_EPI_THREADS = 256;
__global__ void __launch_bounds__(MES_EPI_THREADS) mes_epilogue_kernel(
    const double* __restrict__ mu, const double* __restrict__ sd, int m, const float* __restrict__ gumbel, int K,
    int Kp, const double* __restrict__ fit, double* __restrict__ out) {
  extern __shared__ float gs[];   // Kp = K rounded up to a power of two, padded with +inf
  __shared__ double tab[BGP_MES_TAB_INTERVALS * BGP_MES_TAB_COEFS];
  const int s = blockIdx.y, tid = threadIdx.x;
  for (int j = tid; j < BGP_MES_TAB_INTERVALS * BGP_MES_TAB_COEFS; j += MES_EPI_THREADS) tab[j] = BGP_MES_TAB[j];
  const double alpha = fit[s * 5 + 0], beta = fit[s * 5 + 1];
  for (int k = tid; k < Kp; k += MES_EPI_THREADS) gs[k] = k < K ? gumbel[(size_t)s * K + k] : INFINITY;
  __syncthreads();
  for (int size = 2; size <= Kp; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int idx = tid; idx < Kp / 2; idx += MES_EPI_THREADS) {
        const int lo = 2 * idx - (idx & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const float a = gs[lo], b = gs[hi];
        if ((a > b) == up) { gs[lo] = b; gs[hi] = a; }
      }
      __syncthreads();
    }
  const bool sorted_up = beta > 0.0;   // maxv = beta g + alpha is ascending with g
  // a candidate's K draws are split over MES_KL adjacent lanes (more warps in flight for the long FP64
  // special-function chains), partial sums combined by a fixed shuffle tree
  const int kl = tid % MES_KL;
  for (int base = blockIdx.x * (MES_EPI_THREADS / MES_KL); base < m; base += gridDim.x * (MES_EPI_THREADS / MES_KL)) {
    const int i = base + tid / MES_KL;
    const bool live = i < m;
    const double mean = live ? -mu[(size_t)s * m + i] : 0.0, b = live ? sd[(size_t)s * m + i] : 1.0;
    // one reciprocal per candidate instead of one division per draw; gamma moves by at most an ulp.  sd = 0 (a
    // candidate on a noise-free training point) keeps the division: 0/0 and x/0 must come out as the
    // reference's NaN / inf
    const double inv_b = 1.0 / b;
    const bool use_inv = b > 1e-300 && isfinite(inv_b);
    double acc = 0.0;
    for (int k = kl; k < K; k += MES_KL) {
      const double maxv = fma((double)gs[k], beta, alpha);
      const double gam = use_inv ? (maxv - mean) * inv_b : (maxv - mean) / b;
      double term;
      if (gam > 0.0) {
        // T(gamma) = exp(-gamma^2/2) R(gamma), R by the piecewise degree-12 polynomials of bgp_mes_table.inc
        // (relative error 2e-16, tools/gen_mes_table.py): one branch-free exp and a Horner scheme instead of
        // exp + erfcx + a division + log1p; beyond the table exp(-gamma^2/2) has underflowed (inf * 0 keeps
        // the NaN of an infinite gamma)
        if (gam < 0.5 * BGP_MES_TAB_INTERVALS) {
          const int iv = (int)(gam * 2.0);
          const double x = fma(4.0, gam, -(double)(2 * iv + 1));
          const double* c = tab + iv * BGP_MES_TAB_COEFS;
          double R = c[0];
#pragma unroll
          for (int j = 1; j < BGP_MES_TAB_COEFS; ++j) R = fma(R, x, c[j]);
          term = R * fast_exp_neg(-0.5 * gam * gam);
        } else {
          term = gam * 0.0;
        }
      } else {
        const double t = -gam * 0.7071067811865476;
        if (t < 26.0) {
          const double ex = erfcx(t);   // Phi = 0.5 ex exp(-t^2), phi = exp(-t^2)/sqrt(2 pi)
          term = gam * 0.3989422804014327 / ex - (log(0.5 * ex) - t * t);
        } else {
          // the reference's naive ratio underflows here (cdf -> 0): keep its non-finite result
          const double cdf = 0.5 * erfc(t), pdf = norm_pdf(gam);
          term = gam * pdf / (2.0 * cdf) - (log(0.5 * erfcx(t)) - t * t);
        }
      }
      acc += term;
      if (sorted_up && gam > 1.0 && term <= 1e-18 * acc) break;
    }
#pragma unroll
    for (int o = 1; o < MES_KL; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (live && kl == 0) out[(size_t)s * m + i] = acc / K;
  }
}

__global__ void finite_rows_kernel(const double* __restrict__ v, int m, const double* __restrict__ stats,
                                   int32_t* __restrict__ skipped) {
  const int s = blockIdx.x;
  int fin = 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) fin &= isfinite(v[(size_t)s * m + i]) ? 1 : 0;
  fin = __syncthreads_and(fin);
  if (threadIdx.x == 0) skipped[s] = fin ? 0 : 1;
}

__global__ void combine_kernel(const double* __restrict__ v, int S, int m, const int32_t* __restrict__ skipped,
                               double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double acc = 0.0;
  for (int s = 0; s < S; ++s)
    if (!skipped[s]) acc += v[(size_t)s * m + i] / S;
  out[i] = acc;
}

// mes_epilogue_kernel keeps the K Gumbel variates (padded to a power of two, floats) in dynamic shared memory
constexpr int MES_MAX_K = 32768;
cudaError_t prepare_acq() {
  return cudaFuncSetAttribute(mes_epilogue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              MES_MAX_K * (int)sizeof(float));   // + 7.9 KB static for the term table
}

size_t acq_scratch_doubles(int S, int m) {
  const size_t nblk = (m + MES_CH - 1) / MES_CH;
  return (size_t)S * (ST + MS + MES_PTS + 8 + 5 + 4) + 2 * (size_t)S * nblk * MES_PTS + (size_t)S * m + 64;
}

struct AcqScratch {
  double *stats, *st, *pts, *fit, *ref, *gpart, *dpart, *tmp;
  long long* imax;
};
static AcqScratch carve(double* base, int S, int m) {
  const size_t nblk = (m + MES_CH - 1) / MES_CH;
  AcqScratch w;
  w.stats = base;
  w.st = w.stats + (size_t)S * ST;
  w.pts = w.st + (size_t)S * MS;
  w.imax = reinterpret_cast<long long*>(w.pts + (size_t)S * MES_PTS);
  w.fit = w.pts + (size_t)S * MES_PTS + (size_t)S * 8;
  w.ref = w.fit + (size_t)S * 5;
  w.gpart = w.ref + (size_t)S * 4;
  w.dpart = w.gpart + (size_t)S * nblk * MES_PTS;
  w.tmp = w.dpart + (size_t)S * nblk * MES_PTS;
  return w;
}

// stats_out (S x 4): min mu, min(-mu - 3 sd), max(-mu + 5 sd), all-finite flag
cudaError_t launch_acq_stats(const double* mu, const double* sd, int S, int m, double* stats_out,
                             double* scratch, cudaStream_t stream) {
  AcqScratch w = carve(scratch, S, m);
  acq_stats_kernel<<<S, 1024, 0, stream>>>(mu, sd, m, w.stats, w.pts);
  if (stats_out) cudaMemcpy2DAsync(stats_out, 4 * sizeof(double), w.stats, ST * sizeof(double), 4 * sizeof(double), S,
                                   cudaMemcpyDeviceToDevice, stream);
  return cudaGetLastError();
}

// Gumbel fit of the max-value distribution (needs a preceding launch_acq_stats on the same data)
cudaError_t launch_mes_fit(const double* mu, const double* sd, int S, int m, double* fit_out, double* scratch,
                           cudaStream_t stream) {
  AcqScratch w = carve(scratch, S, m);
  const int nblk = (m + MES_CH - 1) / MES_CH;
  auto grid = [&](int npts) { return dim3(nblk, S, (npts + MES_PCH - 1) / MES_PCH); };
  int* ticket = reinterpret_cast<int*>(w.imax + S);   // S ints behind the S argmax slots (8 S doubles are reserved)
  mes_eval_kernel<<<grid(64), 256, 0, stream>>>(mu, sd, m, w.pts, 64, 0, w.gpart, w.dpart, 0, nullptr, nullptr, nullptr);
  mes_bracket_kernel<<<S, 32, 0, stream>>>(nblk, w.pts, w.gpart, w.st, ticket);
  for (int it = 0; it < MES_NEWTON; ++it)
    mes_eval_kernel<<<grid(3), 256, 0, stream>>>(mu, sd, m, w.pts, 3, 1, w.gpart, w.dpart, it == MES_NEWTON - 1 ? 2 : 1, w.st,
                                                 fit_out ? fit_out : w.fit, ticket);
  return cudaGetLastError();
}

// local EI maximiser per theta: ref_out (S x 4) = {ei max, global index, mu, sd}
cudaError_t launch_ei_best(const double* mu, const double* sd, int S, int m, double p0, const double* yopt,
                           long long index_offset, double* ref_out, double* scratch, cudaStream_t stream) {
  if (S > 1024) return cudaErrorInvalidValue;
  AcqScratch w = carve(scratch, S, m);
  dim3 ge((m + 255) / 256, S);
  ei_kernel<<<ge, 256, 0, stream>>>(mu, sd, m, yopt ? yopt : w.stats, yopt ? 1 : ST, p0, w.tmp);
  argmax_rows_kernel<<<S, 1024, 0, stream>>>(w.tmp, m, w.imax);
  ttei_ref_kernel<<<1, S, 0, stream>>>(mu, sd, w.tmp, m, w.imax, index_offset, ref_out ? ref_out : w.ref);
  return cudaGetLastError();
}

cudaError_t launch_acq_per_theta(const AcqArgs& A, cudaStream_t stream) {
  const int S = A.S, m = A.m;
  AcqScratch w = carve(A.scratch, S, m);
  dim3 ge((m + 255) / 256, S);
  switch (A.kind) {
    case BGP_ACQ_EI:
      ei_kernel<<<ge, 256, 0, stream>>>(A.mu, A.sd, m, A.yopt ? A.yopt : w.stats, A.yopt ? 1 : ST, A.p0, A.per_theta);
      break;
    case BGP_ACQ_TTEI:
      ttei_kernel<<<ge, 256, 0, stream>>>(A.mu, A.sd, m, A.ref ? A.ref : w.ref, A.per_theta);
      break;
    case BGP_ACQ_MEAN:
    case BGP_ACQ_LCB:
      mean_lcb_kernel<<<ge, 256, 0, stream>>>(A.mu, A.sd, m, A.kind, A.p0, A.per_theta);
      break;
    case BGP_ACQ_MES: {
      if (!A.u32 || A.K <= 0 || A.K > MES_MAX_K) return cudaErrorInvalidValue;
      int Kp = 2;
      while (Kp < A.K) Kp <<= 1;
      // every CTA sorts its theta's variates once, then walks over chunks of 32 candidates: about 6 CTAs per SM
      // in all, fewer when there are not that many chunks
      const int chunks = (m + MES_EPI_THREADS / MES_KL - 1) / (MES_EPI_THREADS / MES_KL);
      int per_theta = (6 * 148 + S - 1) / S;
      per_theta = per_theta > chunks ? chunks : per_theta;
      dim3 gm(per_theta, S);
      mes_epilogue_kernel<<<gm, MES_EPI_THREADS, Kp * sizeof(float), stream>>>(A.mu, A.sd, m, A.u32, A.K, Kp,
                                                                              A.mes_fit ? A.mes_fit : w.fit, A.per_theta);
    } break;
    default: return cudaErrorInvalidValue;
  }
  finite_rows_kernel<<<S, 1024, 0, stream>>>(A.per_theta, m, w.stats, A.skipped);
  return cudaGetLastError();
}

cudaError_t launch_acq_combine(const double* per_theta, int S, int m, const int32_t* skipped, double* out,
                               cudaStream_t stream) {
  combine_kernel<<<(m + 255) / 256, 256, 0, stream>>>(per_theta, S, m, skipped, out);
  return cudaGetLastError();
}

// single-GPU composition: stats -> (fit | EI argmax) -> per-theta values -> finite-guarded mean
cudaError_t launch_acq(const AcqArgs& A, cudaStream_t stream) {
  if (A.S <= 0) return cudaSuccess;
  cudaError_t e = launch_acq_stats(A.mu, A.sd, A.S, A.m, nullptr, A.scratch, stream);
  if (e != cudaSuccess) return e;
  AcqArgs B = A;
  if (A.kind == BGP_ACQ_MES && !A.mes_fit_given) {
    e = launch_mes_fit(A.mu, A.sd, A.S, A.m, A.mes_fit, A.scratch, stream);
    if (e != cudaSuccess) return e;
  }
  if (A.kind == BGP_ACQ_TTEI && !A.ref) {
    e = launch_ei_best(A.mu, A.sd, A.S, A.m, A.p0, A.yopt, 0, nullptr, A.scratch, stream);
    if (e != cudaSuccess) return e;
  }
  e = launch_acq_per_theta(B, stream);
  if (e != cudaSuccess) return e;
  return launch_acq_combine(A.per_theta, A.S, A.m, A.skipped, A.out, stream);
}

__global__ void argmax_final_kernel(const double* __restrict__ v, int m, long long* __restrict__ idx) {
  double bv = -INFINITY;
  long long bi = 1LL << 62;
  for (long long i = threadIdx.x; i < m; i += blockDim.x) {
    const double x = v[i];
    if (better_max(x, i, bv, bi)) { bv = x; bi = i; }
  }
  __shared__ double sv[32];
  __shared__ long long si[32];
  block_argmax(bv, bi, sv, si);
  if (threadIdx.x == 0) idx[0] = bi;
}

cudaError_t launch_argmax(const double* v, int m, long long* idx, double*, cudaStream_t stream) {
  argmax_final_kernel<<<1, 1024, 0, stream>>>(v, m, idx);
  return cudaGetLastError();
}

}  // namespace bgp
