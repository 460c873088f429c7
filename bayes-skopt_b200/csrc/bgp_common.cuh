// Shared device-side pieces of libbgp: covariance-program interpreter, DMMA wrapper,
// Philox-4x32-10, tiled factor-slab layout.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/bgp.h"

#define BGP_NB 32  // panel width (columns factored together)

// ---------------------------------------------------------------------------- layout
// Factor slab ("panel-major lower"): panel j holds columns [32j, 32j+32) for storage rows
// [32j, R) as a row-major (R-32j) x 32 block.  Storage rows:
//   [0, 32P)            training rows (rows >= n are identity padding)
//   32P                 the y row: after the factorisation it holds z = L^-1 y
//   [32P+32, 64P+32)    identity rows: after the factorisation they hold L^-T, i.e. the
//                       block of panel j, read as [i][c], is (L^-1)[32j+c][i]
struct SlabGeom {
  int n, P, R, Rz, Ra;
  __host__ __device__ static SlabGeom make(int n, bool aug) {
    SlabGeom g;
    g.n = n;
    g.P = (n + BGP_NB - 1) / BGP_NB;
    g.Rz = BGP_NB * g.P;
    g.Ra = g.Rz + BGP_NB;
    g.R = aug ? g.Ra + BGP_NB * g.P : g.Ra;
    return g;
  }
  __host__ __device__ long long off(int j) const {
    return 32LL * ((long long)j * R - 16LL * j * (j - 1));
  }
  __host__ __device__ long long doubles() const { return off(P); }
  // first element of the L^-T block of panel j
  __host__ __device__ long long aug_base(int j) const { return off(j) + (long long)(Ra - 32 * j) * 32; }
};

// ------------------------------------------------------------------- covariance program
struct DevProgram {
  int n_ops, n_theta, n_leaves, d;
  // fast path for the shape  c * stationary(r) + white  (the default bask kernel): opcode of the
  // stationary leaf (0 = use the interpreter) and the op slots of the constant / white levels
  int fast_kind, fast_const, fast_white, fast_white_zeroable;
  // input warping (bask warp_inputs=True): theta rows carry 2 * n_warp extra entries after the
  // kernel's own, log a_1..a_d then log b_1..b_d of the per-dimension Beta CDF warps
  // (bask/bayesgpr.py:351-365); n_theta is the full row length
  int n_warp, warp_off;
  int reserved_[2];   // keeps sizeof(DevProgram) a multiple of 16: shared-memory structs that embed it
                      // access their later members with 16-byte loads
  bgp_op_t ops[BGP_MAX_OPS];
  int leaf_of_op[BGP_MAX_OPS];
};

static_assert(sizeof(DevProgram) % 16 == 0, "DevProgram must keep 16-byte alignment of what follows it");

// per-theta resolved parameters (shared memory)
struct ThetaParams {
  double opval[BGP_MAX_OPS];                     // exp(theta) or the fixed value / exponent
  double inv_ls[BGP_MAX_LEAVES][BGP_MAX_DIM];    // 1 / length scale per stationary leaf
};

__device__ __forceinline__ void resolve_theta(const DevProgram& P, const double* __restrict__ theta,
                                              const double* __restrict__ fixed_ls,
                                              ThetaParams& out, int tid, int nthreads) {
  for (int o = tid; o < P.n_ops; o += nthreads) {
    const bgp_op_t& op = P.ops[o];
    double v = op.value;
    if ((op.code == BGP_OP_CONST || op.code == BGP_OP_WHITE) && op.theta_idx >= 0)
      v = exp(theta[op.theta_idx]);
    out.opval[o] = v;
  }
  for (int o = 0; o < P.n_ops; ++o) {
    const bgp_op_t& op = P.ops[o];
    if (op.code < BGP_OP_RBF || op.code > BGP_OP_MATERN52) continue;
    int leaf = P.leaf_of_op[o];
    for (int k = tid; k < P.d; k += nthreads) {
      double ls;
      if (op.theta_idx >= 0)
        ls = exp(theta[op.theta_idx + (op.n_ls > 1 ? k : 0)]);
      else
        ls = (op.n_ls > 1) ? fixed_ls[op.fixed_ls_offset + k] : op.value;
      out.inv_ls[leaf][k] = 1.0 / ls;
    }
  }
}

// Evaluates the postfix program for one pair of points.  r2[leaf] = squared scaled
// distance per stationary leaf; same_point selects the White contribution (Gram diagonal
// or k(x,x)); white_on=false switches the zeroable White leaf off (noise_set_to_zero).
__device__ __forceinline__ double pick_leaf(const double* r2, int leaf) {
  double v = r2[0];
#pragma unroll
  for (int l = 1; l < BGP_MAX_LEAVES; ++l) v = (leaf == l) ? r2[l] : v;
  return v;
}

// exp(x) for finite x <= 0 and sqrt(x) for x >= 0, written for the Gram / k* inner loops: the library
// routines spend three quarters of their issue slots on things these loops do not need (range and
// special-case branches, register moves, constants materialised per use -- 166 instructions per
// matrix entry of which 44 were FP64).  Here every constant is a direct constant-bank operand, there is no
// branch, and the results stay within ~1 ulp of the library's (tools/fastmath_check.cu).
//   exp: x = k ln2 + r (Cody-Waite, k by the 1.5 * 2^52 trick), Taylor polynomial of degree 13 in r
//        (|r| <= ln2 / 2: remainder 4e-18), scaling by an integer add to the exponent; below -700 (1e-304,
//        where the add would reach the denormals) the result is flushed to zero.
//   sqrt: MUFU.RSQ64H seed, two coupled Goldschmidt iterations for sqrt(x) and 1 / (2 sqrt(x)), one
//        residual correction; x = 0 (seed = inf) is selected at the end.
__constant__ double BGP_EXP_TAYLOR[14] = {
    1.0, 1.0, 0.5, 1.6666666666666666e-01, 4.1666666666666664e-02, 8.3333333333333332e-03,
    1.3888888888888889e-03, 1.9841269841269841e-04, 2.4801587301587302e-05, 2.7557319223985893e-06,
    2.7557319223985888e-07, 2.5052108385441720e-08, 2.0876756987868100e-09, 1.6059043836821613e-10};

__device__ __forceinline__ double fast_exp_neg(double x) {
  double kd = fma(x, 1.4426950408889634, 6755399441055744.0);
  const int ki = __double2loint(kd);
  kd -= 6755399441055744.0;
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = BGP_EXP_TAYLOR[13];
#pragma unroll
  for (int i = 12; i >= 0; --i) p = fma(p, r, BGP_EXP_TAYLOR[i]);
  p = __hiloint2double(__double2hiint(p) + (ki << 20), __double2loint(p));
  return x < -700.0 ? 0.0 : p;
}

__device__ __forceinline__ double fast_sqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double e = fma(-g, h, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-g, h, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  g = fma(fma(-g, g, x), h, g);
  return x > 1e-290 ? g : 0.0;
}

__device__ __forceinline__ double stationary_value(int code, double r2) {
  if (code == BGP_OP_MATERN52) {
    const double t = fast_sqrt(r2) * 2.23606797749979;
    return (1.0 + t + t * t * 0.3333333333333333) * fast_exp_neg(-t);
  }
  if (code == BGP_OP_RBF) return fast_exp_neg(-0.5 * r2);
  if (code == BGP_OP_MATERN32) {
    const double t = fast_sqrt(r2) * 1.7320508075688772;
    return (1.0 + t) * fast_exp_neg(-t);
  }
  return fast_exp_neg(-fast_sqrt(r2));
}

__device__ __forceinline__ double eval_program(const DevProgram& P, const ThetaParams& T,
                                               const double* r2, bool same_point, bool white_on) {
  if (P.fast_kind) {
    double v = T.opval[P.fast_const] * stationary_value(P.fast_kind, r2[0]);
    if (same_point && (white_on || !P.fast_white_zeroable)) v += T.opval[P.fast_white];
    return v;
  }
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#define BGP_PUSH(v) { s3 = s2; s2 = s1; s1 = s0; s0 = (v); }
#define BGP_BIN(expr) { double _a = s1, _b = s0; s0 = (expr); s1 = s2; s2 = s3; }
  for (int o = 0; o < P.n_ops; ++o) {
    const int code = P.ops[o].code;
    switch (code) {
      case BGP_OP_CONST: BGP_PUSH(T.opval[o]); break;
      case BGP_OP_WHITE: {
        bool on = same_point && (white_on || !(P.ops[o].flags & BGP_FLAG_ZEROABLE_WHITE));
        BGP_PUSH(on ? T.opval[o] : 0.0);
      } break;
      case BGP_OP_RBF: BGP_PUSH(exp(-0.5 * pick_leaf(r2, P.leaf_of_op[o]))); break;
      case BGP_OP_MATERN12: BGP_PUSH(exp(-sqrt(pick_leaf(r2, P.leaf_of_op[o])))); break;
      case BGP_OP_MATERN32: {
        double t = sqrt(pick_leaf(r2, P.leaf_of_op[o])) * 1.7320508075688772;
        BGP_PUSH((1.0 + t) * exp(-t));
      } break;
      case BGP_OP_MATERN52: {
        double t = sqrt(pick_leaf(r2, P.leaf_of_op[o])) * 2.23606797749979;
        BGP_PUSH((1.0 + t + t * t / 3.0) * exp(-t));
      } break;
      case BGP_OP_ADD: BGP_BIN(_a + _b); break;
      case BGP_OP_MUL: BGP_BIN(_a * _b); break;
      case BGP_OP_POW: s0 = pow(s0, T.opval[o]); break;
      default: break;
    }
  }
#undef BGP_PUSH
#undef BGP_BIN
  return s0;
}

// -------------------------------------------------------------------------- input warp
// Regularised incomplete beta function I_x(a, b) = Beta(a, b).cdf(x) (scipy.stats.beta.cdf in
// bask/bayesgpr.py:298-316): continued fraction (modified Lentz) on the side where it converges
// fast, prefactor through lgamma.
__device__ inline double bgp_betacf(double a, double b, double x) {
  const double TINY = 1e-300, EPS = 1e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < TINY) d = TINY;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 300; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < TINY) d = TINY;
    c = 1.0 + aa / c; if (fabs(c) < TINY) c = TINY;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < TINY) d = TINY;
    c = 1.0 + aa / c; if (fabs(c) < TINY) c = TINY;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < EPS) break;
  }
  return h;
}
__device__ inline double bgp_beta_cdf(double x, double a, double b) {
  if (!(x > 0.0)) return 0.0;
  if (x >= 1.0) return 1.0;
  const double front = exp(a * log(x) + b * log1p(-x) - (lgamma(a) + lgamma(b) - lgamma(a + b)));
  if (x < (a + 1.0) / (a + b + 2.0)) return front * bgp_betacf(a, b, x) / a;
  return 1.0 - front * bgp_betacf(b, a, 1.0 - x) / b;
}
// coordinate kk of a point in the (possibly warped) input space of theta row `theta`
__device__ __forceinline__ double bgp_warp_coord(const DevProgram& P, const double* __restrict__ theta, int kk,
                                                 double x) {
  if (P.n_warp == 0) return x;
  return bgp_beta_cdf(x, exp(theta[P.warp_off + kk]), exp(theta[P.warp_off + P.n_warp + kk]));
}

// ------------------------------------------------------------------------------ priors
// typed log-priors of a theta row (bask/utils.py:68-124, bask/priors.py:7-57), summed in table order
__device__ __forceinline__ double log_prior(const bgp_prior_t* pr, int n, const double* theta) {
  double lp = 0.0;
  for (int k = 0; k < n; ++k) {
    const double x = theta[k];
    const double* p = pr[k].p;
    switch (pr[k].kind) {
      case BGP_PRIOR_HALFNORMAL_SQRT:
        lp += -0.22579135264472744 /* 0.5*log(2/pi) */ - log(p[0]) - exp(x) / (2.0 * p[0] * p[0]) +
              0.5 * x - 0.6931471805599453;
        break;
      case BGP_PRIOR_ROUNDFLAT:
        lp += -2.0 * (exp(-2.0 * p[2] * (x - log(p[0]))) + exp(2.0 * p[3] * (x - log(p[1])))) -
              p[4] + x;
        break;
      case BGP_PRIOR_INVGAMMA:
        lp += p[0] * log(p[1]) - lgamma(p[0]) - (p[0] + 1.0) * x - p[1] * exp(-x) + x;
        break;
      case BGP_PRIOR_NORMAL: {
        double t = (x - p[0]) / p[1];
        lp += -0.5 * t * t - log(p[1]) - 0.9189385332046727;
      } break;
      default: break;
    }
  }
  return lp;
}

// ------------------------------------------------------------------------------- DMMA
// D(8x8) += A(8x4, row) * B(4x8, col).  Lane l holds A[l>>2][l&3], B[l&3][l>>2],
// D[l>>2][2(l&3)], D[l>>2][2(l&3)+1].  SASS: DMMA.8x8x4 (tcgen05 has no f64 kind).
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------ Philox
struct Philox4 {
  uint32_t c[4];
};
__device__ __forceinline__ Philox4 philox4x32_10(uint64_t key, uint32_t c0, uint32_t c1, uint32_t c2,
                                                 uint32_t c3) {
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
    uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
    x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 o; o.c[0] = x0; o.c[1] = x1; o.c[2] = x2; o.c[3] = x3;
  return o;
}
// uniform in (0,1) with 53 random bits (never 0, never 1)
__device__ __forceinline__ double u01_from(uint32_t hi, uint32_t lo) {
  uint64_t bits = (((uint64_t)hi << 32) | lo) >> 11;  // 53 bits
  return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
}

// -------------------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
