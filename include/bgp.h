/* libbgp -- C ABI of the B200-native fully Bayesian GP hot path (sm_100a).
 *
 * The reference (kiudee/bayes-skopt) has no native boundary: its hot path is Python on top of
 * scikit-learn / scikit-optimize / emcee / LAPACK.  Each entry point below replaces one of
 * those Python call sites; the reference file:line it replaces is cited on the declaration
 * (paths relative to the reference tree; "sklearn:" = scikit-learn 1.7.2
 * sklearn/gaussian_process/).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - Plain C symbols, no C++/torch types.  Every call returns 0 on success, <0 on an
 *     argument / CUDA error (bgp_last_error() gives the text).  Nothing throws.
 *   - All matrices are FP64, row-major.  theta is in log space, scikit-learn order
 *     (k1.theta ++ k2.theta, sklearn:kernels.py:738-766).
 *   - Pointers named *_dev are DEVICE pointers owned by the caller; the handle owns only
 *     its private workspace.  `stream` is a cudaStream_t passed as void* (0 = default).
 *   - Numerical failures are per item: a non positive definite Gram matrix gives
 *     log-prob = -inf and info[b] = failing column + 1 (sklearn:_gpr.py:592-593,
 *     bask/bayesgpr.py:373-378).  A batch is never aborted.
 *   - One handle per (device, stream); thread-compatible, not thread-safe.
 */
#ifndef BGP_H_
#define BGP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bgp_handle_s* bgp_handle_t;

/* ---- covariance program: a kernel tree compiled to postfix (host side does the walk) ---- */
enum bgp_opcode {
  BGP_OP_CONST = 1,    /* push c                          sklearn:kernels.py:1244-1296 */
  BGP_OP_WHITE = 2,    /* push sigma^2 iff same point     sklearn:kernels.py:1374-1419 */
  BGP_OP_RBF = 3,      /* push exp(-r^2/2)                sklearn:kernels.py:1530-1587 */
  BGP_OP_MATERN12 = 4, /* push exp(-r)                    sklearn:kernels.py:1685-1786 */
  BGP_OP_MATERN32 = 5, /* push (1+sqrt3 r) exp(-sqrt3 r) */
  BGP_OP_MATERN52 = 6, /* push (1+sqrt5 r+5r^2/3) exp(-sqrt5 r) */
  BGP_OP_ADD = 7,      /* Sum                             sklearn:kernels.py:838-873 */
  BGP_OP_MUL = 8,      /* Product                         sklearn:kernels.py:936-973 */
  BGP_OP_POW = 9       /* Exponentiation (fixed exponent) sklearn:kernels.py:1060-1110 */
};

#define BGP_FLAG_ZEROABLE_WHITE 1 /* the White leaf noise_set_to_zero() switches off
                                     (bask/bayesgpr.py:318-336) */
#define BGP_MAX_OPS 24
#define BGP_MAX_LEAVES 4  /* stationary (RBF/Matern) leaves per program */
#define BGP_MAX_DIM 64
#define BGP_MAX_THETA 80

typedef struct {
  int32_t code;      /* enum bgp_opcode */
  int32_t theta_idx; /* first theta slot of this leaf, or -1 when the leaf is fixed */
  int32_t n_ls;      /* stationary leaves: 1 (isotropic) or d (ARD); else 0 */
  int32_t flags;
  double value;      /* fixed value (const/white/isotropic length scale) or the exponent */
  int32_t fixed_ls_offset; /* ARD leaf with fixed length scales: offset into fixed_ls[] */
  int32_t reserved;
} bgp_op_t;

/* ---- log-prior program, one entry per theta slot (bask/utils.py:68-124, bask/priors.py) ---- */
enum bgp_prior_kind {
  BGP_PRIOR_NONE = 0,
  BGP_PRIOR_HALFNORMAL_SQRT = 1, /* p[0]=scale: halfnorm on sqrt(exp x), + x/2 - log 2 */
  BGP_PRIOR_ROUNDFLAT = 2,       /* p[0..3]=lo,hi,steep_lo,steep_hi  p[4]=log Z; + x */
  BGP_PRIOR_INVGAMMA = 3,        /* p[0]=a p[1]=scale on exp x; + x */
  BGP_PRIOR_NORMAL = 4           /* p[0]=loc p[1]=scale directly on x */
};
typedef struct {
  int32_t kind;
  int32_t reserved;
  double p[6];
} bgp_prior_t;

/* ---- life cycle ---- */
int bgp_create(bgp_handle_t* out, int device);
int bgp_destroy(bgp_handle_t h);
const char* bgp_last_error(void);
int bgp_version(void);
/* kernels one bgp_logprob_batched wave launches for the current model and data: 1 (fused small-n kernel),
 * 2 (Gram kernel that scales its inputs + factorisation) or 3 (scale_x + Gram + factorisation); for hosts that
 * report launch counts */
int bgp_logprob_launches(bgp_handle_t h);

/* kernel_ = user kernel (+ WhiteKernel): replaces kernel.clone_with_theta + kernel.__call__
 * (sklearn:_gpr.py:577-586).  fixed_ls: concatenated fixed ARD length scales (may be NULL). */
int bgp_set_kernel(bgp_handle_t h, const bgp_op_t* ops, int n_ops, int n_theta,
                   const double* fixed_ls, int n_fixed_ls);
/* warp_inputs=True (bask/bayesgpr.py:219-316, 351-365): every theta row then carries 2*n_warp extra
 * entries behind the kernel's own -- log a_1..a_d, log b_1..b_d of the per-dimension Beta-CDF
 * warps -- and every point set (training inputs, candidates) is warped per theta on the device.
 * n_warp = d enables, 0 disables; call after bgp_set_kernel (which resets it) and before
 * bgp_set_priors (whose table covers the extended row). */
int bgp_set_warp(bgp_handle_t h, int n_warp);
/* priors=None of bask/bayesgpr.py:459-460 -> guess_priors, or the typed user priors. */
int bgp_set_priors(bgp_handle_t h, const bgp_prior_t* priors, int n_priors);
/* X_train_/y_train_/alpha of the estimator (bask/bayesgpr.py:469-488).  Copied into the
 * handle (device-to-device, async on `stream`). */
int bgp_set_data(bgp_handle_t h, const double* X_dev, const double* y_dev,
                 const double* alpha_dev, int n, int d, void* stream);

/* ---- K1+K2: batched log posterior --------------------------------------------------
 * lp[b] = sum_k prior_k(theta[b,k]) + LML(theta[b]) (+ lp_extra[b]); replaces
 * BayesGPR._log_prob_fn (bask/bayesgpr.py:351-379) and log_marginal_likelihood
 * (sklearn:_gpr.py:541-617) called W*(1+T) times by emcee.  lml_dev / lp_extra_dev may be
 * NULL.  Non-finite results are mapped to -inf (bask/bayesgpr.py:377-378). */
int bgp_logprob_batched(bgp_handle_t h, const double* theta_dev, int batch,
                        const double* lp_extra_dev, double* lp_dev, double* lml_dev,
                        int32_t* info_dev, void* stream);

/* ---- factorisation at S thetas (the `theta` setter, bask/bayesgpr.py:200-217) -------
 * Writes, per theta, an opaque factor slab (L and L^-1 in the library's tiled layout;
 * bgp_factor_slab_doubles() each), z = L^-1 y (S x n) and the LML.  The dense attributes
 * of the estimator are extracted on demand by bgp_factor_extract. */
int64_t bgp_factor_slab_doubles(bgp_handle_t h);
int bgp_factorize_batched(bgp_handle_t h, const double* theta_dev, int S, double* slabs_dev,
                          double* z_dev, double* lml_dev, int32_t* info_dev, void* stream);
enum bgp_extract_what {
  BGP_EXTRACT_L = 1,     /* L_      (n x n, zeros above the diagonal) */
  BGP_EXTRACT_LINV = 2,  /* L^-1    (n x n) */
  BGP_EXTRACT_KINV = 3,  /* K_inv_ = L^-T L^-1 (n x n)   bask/bayesgpr.py:207-208 */
  BGP_EXTRACT_ALPHA = 4  /* alpha_ = K^-1 y  (n)         bask/bayesgpr.py:217 */
};
int bgp_factor_extract(bgp_handle_t h, const double* slab_dev, const double* z_dev, int what,
                       double* out_dev, void* stream);

/* ---- analytic LML gradient (MAP start of fit) ---------------------------------------
 * grad[k] = 1/2 tr((alpha alpha^T - K^-1) dK/dtheta_k) over the kernel's own log hyper-parameters
 * (input-warp entries of a theta row are not differentiated); replaces
 * log_marginal_likelihood(theta, eval_gradient=True) of sklearn:_gpr.py:619-651 as driven by the
 * L-BFGS-B search of the skopt fit (bask/bayesgpr.py:607).  alpha_dev (n) and kinv_dev (n x n) are
 * bgp_factor_extract(BGP_EXTRACT_ALPHA / BGP_EXTRACT_KINV) of the factorisation at theta_dev (one row). */
int bgp_lml_gradient(bgp_handle_t h, const double* theta_dev, const double* alpha_dev,
                     const double* kinv_dev, double* grad_dev, void* stream);

/* ---- K4: candidate sweep ------------------------------------------------------------
 * For S thetas and m candidates: mu[s,i] = y_std * k*(x_i)^T K^-1 y + y_mean and
 * sd[s,i] = sqrt(max(0, k(x_i,x_i) - k*^T K^-1 k*) * y_std^2); replaces gpr.predict
 * (bask/acquisition.py:129 -> skopt GaussianProcessRegressor.predict) after
 * `gpr.theta = chain_[i]` (bask/acquisition.py:121).  noise_off=1 evaluates k(x,x)
 * without the zeroable White leaf (noise_set_to_zero).  Optional extras:
 *   zextra_dev (S x R x n): dots_dev[s,r,i] = (L^-1 k*_i) . zextra[s,r]   (PVRS / VR)
 *   v_dev (S x m x n_pad):  the whitened cross-covariances L^-1 k*_i     (joint draws)    */
int bgp_predict_batched(bgp_handle_t h, const double* theta_dev, int S, const double* slabs_dev,
                        const double* z_dev, const double* Xc_dev, int m, int noise_off,
                        double y_mean, double y_std, double* mu_dev, double* sd_dev,
                        const double* zextra_dev, int R, double* dots_dev, double* v_dev,
                        int64_t v_ld, void* stream);

/* ---- joint posterior over a candidate set -------------------------------------------
 * cov[i,j] = (k(x_i,x_j) - v_i . v_j) * y_std^2 from the whitened cross-covariances v written
 * by bgp_predict_batched (v_dev, one theta); replaces the return_cov branch of skopt's predict
 * inside sklearn sample_y (sklearn:_gpr.py:502-539) as used by BayesGPR.sample_y
 * (bask/bayesgpr.py:637-718).  bgp_dense_cholesky factors cov + jitter*I into a slab of
 * bgp_dense_slab_doubles(m) doubles (info = failing column + 1, as LAPACK dpotrf), and
 * bgp_slab_trmm draws out[i,s] = mean[i] + sum_c L[i,c] e[c,s] (e: m x ns standard normals). */
int bgp_posterior_cov(bgp_handle_t h, const double* theta_dev, const double* v_dev,
                      const double* Xc_dev, int m, int64_t v_ld, int noise_off, double y_std,
                      double* cov_dev, int64_t ldc, void* stream);
int64_t bgp_dense_slab_doubles(int m);
int bgp_dense_cholesky(bgp_handle_t h, const double* a_dev, int m, int64_t lda, double jitter,
                       double* slab_dev, int32_t* info_dev, void* stream);
/* Chip-wide variant for one LARGE matrix (thousands of candidates): factors a_dev + jitter*I in place
 * (row-major, lower triangle read and overwritten with L; the strictly upper part is left alone) by
 * 256-column blocks -- diagonal blocks on the library's own DMMA kernel, block-column solve and trailing
 * update as cuBLAS dtrsm / dsyrk (bound lazily; every other entry point is free of library calls) --
 * and bgp_dense_trmm draws out = mean + L e from it. */
int bgp_dense_cholesky_inplace(bgp_handle_t h, double* a_dev, int m, int64_t lda, double jitter,
                               int32_t* info_dev, void* stream);
int bgp_dense_trmm(bgp_handle_t h, const double* l_dev, int m, int64_t lda, const double* e_dev, int ns,
                   const double* mean_dev, double* out_dev, void* stream);
int bgp_slab_trmm(bgp_handle_t h, const double* slab_dev, int m, const double* e_dev, int ns,
                  const double* mean_dev, double* out_dev, void* stream);

/* ---- full-GP acquisitions in Schur-complement form ------------------------------------
 * The reference refactorises an (n+1)x(n+1) Gram matrix per candidate (PVRS
 * bask/acquisition.py:328-339, VarianceReduction :285-300).  With v = L^-1 k(X, .) at the
 * current theta (noise ON) the same number is
 *   out[i] = sum_t |v_t|^2 + sum_t (k(t, x_i) - v_t . v_i)^2 / s_i,   s_i = k(x_i,x_i) - |v_i|^2
 * PVRS: t runs over R Thompson points (dots_dev = v_t . v_i from bgp_predict_batched's extra
 * right-hand sides, vt_dev = their whitened vectors, R x n).  VR: t runs over all candidates,
 * with k(t,x_i) - v_t.v_i read from the noise-free posterior covariance (bgp_posterior_cov). */
int bgp_pvrs_combine(bgp_handle_t h, const double* theta_dev, const double* xt_dev, int R,
                     const double* xc_dev, int m, const double* dots_dev, const double* vt_dev,
                     const double* s_dev, double* out_dev, void* stream);
int bgp_vr_combine(bgp_handle_t h, const double* cov_dev, int m, int64_t ldc, const double* xc_dev,
                   const double* theta_dev, const double* s_dev, double* out_dev, void* stream);

/* ---- acquisition epilogues on (mu, sd), S x m -> m (bask/acquisition.py:112-141) ----
 * out[i] = (1/S) * sum over thetas whose row is entirely finite of acq(mu[s,:], sd[s,:])[i].
 * kind: registry strings of bask/optimizer.py:23-32.  p0: EI/TTEI y_opt (NaN = default
 * mu.min(), bask/acquisition.py:166-167); LCB alpha (inf = "inf" mode).  MES takes K
 * standard Gumbel variates per theta (g32_dev, float32, S x K): -log(-log(u)) of the float32
 * uniforms the reference draws from the global numpy RNG, computed in float32 by the host exactly
 * as bask/acquisition.py:253-257 does; mes_fit_dev (S x 5: a, b, q1, med, q2)
 * may be NULL. */
enum bgp_acq_kind { BGP_ACQ_EI = 1, BGP_ACQ_TTEI = 2, BGP_ACQ_MEAN = 3, BGP_ACQ_LCB = 4,
                    BGP_ACQ_MES = 5 };
int bgp_acq_sweep(bgp_handle_t h, int kind, const double* mu_dev, const double* sd_dev, int S,
                  int m, double p0, const float* g32_dev, int K, double* per_theta_dev,
                  double* out_dev, int32_t* skipped_dev, double* mes_fit_dev, void* stream);
/* The same epilogues in stages, for hosts that shard the candidates over several GPUs and
 * exchange the per-theta scalars in between (small all-reduces / all-gathers):
 *   bgp_acq_stats     stats[s] = {min mu, min(-mu-3sd), max(-mu+5sd), all-finite}        (S x 4)
 *   bgp_mes_fit       Gumbel fit {a, b, q1, med, q2} of the max-value distribution from the
 *                     moments of ALL candidates (bask/acquisition.py:235-252)              (S x 5)
 *   bgp_ei_best       ref[s] = {max EI, global index, mu, sd} of this shard's EI maximiser  (S x 4)
 *   bgp_acq_per_theta per-theta values of this shard given the global y_opt (yopt_dev, S), the
 *                     global EI maximiser (ref_dev, TTEI) or the global Gumbel fit (fit_dev, MES);
 *                     skipped[s] = 1 when a value of theta s is non-finite on this shard
 *   bgp_acq_combine   out[i] = (1/S) sum_{s not skipped} per_theta[s, i]                          */
int bgp_acq_stats(bgp_handle_t h, const double* mu_dev, const double* sd_dev, int S, int m,
                  double* stats_dev, void* stream);
int bgp_mes_fit(bgp_handle_t h, const double* mu_dev, const double* sd_dev, int S, int m,
                double* fit_dev, void* stream);
int bgp_ei_best(bgp_handle_t h, const double* mu_dev, const double* sd_dev, int S, int m, double p0,
                const double* yopt_dev, int64_t index_offset, double* ref_dev, void* stream);
int bgp_acq_per_theta(bgp_handle_t h, int kind, const double* mu_dev, const double* sd_dev, int S,
                      int m, double p0, const double* yopt_dev, const double* ref_dev,
                      const float* g32_dev, int K, const double* fit_dev, double* per_theta_dev,
                      int32_t* skipped_dev, void* stream);
int bgp_acq_combine(bgp_handle_t h, const double* per_theta_dev, int S, int m,
                    const int32_t* skipped_dev, double* out_dev, void* stream);

/* argmax with numpy tie-breaking (first maximum), bask/optimizer.py:374-376 */
int bgp_argmax(bgp_handle_t h, const double* v_dev, int m, int64_t* idx_dev, void* stream);

/* ---- multi-GPU sweep for C callers ---------------------------------------------------
 * One call per rank (one process / handle per GPU), `nccl_comm` an ncclComm_t over the `world` ranks created by
 * the caller.  Every rank passes the same thetas (with their factorisations, bgp_factorize_batched), the FULL
 * candidate array (m_total x d) and, for MaxValueSearch, the same Gumbel variates; it sweeps only its contiguous
 * block of candidates (the first m_total % world ranks take one more) and exchanges what evaluate_acquisitions
 * (bask/acquisition.py:48-147) needs globally: the per-theta min mean (EI's default y_opt, all-reduce MIN), the
 * per-theta EI maximiser (TopTwoEI, all-gather), the moments for the Gumbel fit (MaxValueSearch, all-gather), the
 * non-finite flags (all-reduce MAX) and finally the values.  out_dev (m_total) and argmax_dev (may be NULL)
 * are identical on every rank and equal to the single-GPU bgp_acq_sweep on the whole candidate set.  NCCL is
 * bound at run time (dlopen "libnccl.so.2"), so the communicator and these calls share one NCCL instance with
 * the host application.  The walker-sharded MCMC needs no NCCL at all: bgp_peer_export / bgp_peer_connect /
 * bgp_mcmc_run_sharded exchange log-probabilities by peer stores. */
int bgp_acq_sweep_nccl(bgp_handle_t h, void* nccl_comm, int rank, int world, int kind,
                       const double* theta_dev, int S, const double* slabs_dev, const double* z_dev,
                       const double* Xc_all_dev, int m_total, double p0, const float* g32_dev, int K,
                       double y_mean, double y_std, double* out_dev, int64_t* argmax_dev, void* stream);

/* ---- K3: emcee-equivalent stretch move on device (bask/bayesgpr.py:510-530) ---------
 * pos (W x p) in/out, lp (W) out; chain (T x W x p) and lp_chain (T x W) step-major like
 * EnsembleSampler.get_chain; accepted (W) counts.  Philox-4x32-10 keyed by `seed`; the
 * whole run is one CUDA graph launch.  a = stretch scale (emcee default 2.0). */
int bgp_mcmc_run(bgp_handle_t h, double* pos_dev, double* lp_dev, int W, int T, double a,
                 uint64_t seed, double* chain_dev, double* lp_chain_dev, int32_t* accepted_dev,
                 void* stream);
/* building blocks of the same move, for hosts that interleave their own work (Python
 * callable priors, multi-GPU all-gather of log-probs) between propose and accept. */
int bgp_mcmc_split(bgp_handle_t h, int W, uint64_t seed, int step, int32_t* colour_dev,
                   void* stream);
int bgp_mcmc_propose(bgp_handle_t h, const double* pos_dev, const int32_t* colour_dev, int W,
                     int half, double a, uint64_t seed, int step, double* q_dev,
                     double* factors_dev, int32_t* movers_dev, void* stream);
int bgp_mcmc_accept(bgp_handle_t h, double* pos_dev, double* lp_dev, const double* q_dev,
                    const double* factors_dev, const double* new_lp_dev,
                    const int32_t* movers_dev, int W, int half, uint64_t seed, int step,
                    int32_t* accepted_dev, double* chain_step_dev, double* lp_step_dev,
                    void* stream);
/* Stepped entry points above: when seed_dev is not NULL the Philox key is read from that device
 * word instead of the by-value `seed` argument, so a caller can capture the stepped loop (with its
 * own collectives between propose and accept) in a CUDA graph and replay it with a new seed.
 * NULL restores by-value seeds. */
int bgp_mcmc_seed_source(bgp_handle_t h, const uint64_t* seed_dev);

/* ---- multi-GPU walker sharding (one process per GPU, walkers split over the ranks) ----
 * Replaces emcee's serial map over the walkers of a half step (bask/bayesgpr.py:510-530) across GPUs
 * without NCCL and without the host: every rank evaluates its slice of the proposals and stores the
 * log-probabilities straight into every peer's exchange block (cudaIpc mapping, NVLink P2P) followed by
 * a release flag; the accept kernel acquires on the flags.  The whole run is one CUDA graph per rank and
 * every rank ends with the same chain as bgp_mcmc_run on one GPU.
 *   bgp_peer_export   allocates this rank's exchange block for up to max_walkers walkers and returns its
 *                     cudaIpc handle (BGP_IPC_HANDLE_BYTES bytes; the host all-gathers them)
 *   bgp_peer_connect  maps the blocks of all `world` ranks (ipc_handles: world x BGP_IPC_HANDLE_BYTES,
 *                     rank order); world <= 8, one node
 *   bgp_peer_close    unmaps them (before the process group is torn down)
 *   bgp_peer_status   timed_out = 1 when an exchange ever waited longer than 10 s for a peer
 *   bgp_peer_counters out[0] = nanoseconds this rank has spent inside exchanges (peer stores, fence, waiting for
 *                     the slowest rank), out[1] = number of exchanges -- where a sharded run's time goes */
#define BGP_IPC_HANDLE_BYTES 64
int bgp_peer_export(bgp_handle_t h, int max_walkers, void* ipc_handle_out);
int bgp_peer_connect(bgp_handle_t h, const void* ipc_handles, int rank, int world);
int bgp_peer_close(bgp_handle_t h);
int bgp_peer_status(bgp_handle_t h, int* timed_out);
int bgp_peer_counters(bgp_handle_t h, unsigned long long* out);
int bgp_mcmc_run_sharded(bgp_handle_t h, double* pos_dev, double* lp_dev, int W, int T, double a,
                         uint64_t seed, double* chain_dev, double* lp_chain_dev, int32_t* accepted_dev,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BGP_H_ */
