// Host-side internals shared by the translation units of libbgp (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bgp.h"

struct DevProgram;

namespace bgp {

struct CholArgs {
  const double* X;       // n x d
  const double* y;       // n
  const double* alpha;   // n
  const double* theta;   // batch x p
  const double* lp_extra;
  double* lp;
  double* lml;
  int32_t* info;
  double* slabs;         // scratch (slab_per_block=1) or per-theta output slabs
  double* z_out;         // batch x n or null
  const DevProgram* prog;
  const double* fixed_ls;
  const bgp_prior_t* priors;
  int n_priors;
  int n, d, batch, aug, slab_per_block;
  const double* dense;   // when set: factor this dense SPD matrix (lower part read) instead of a Gram
  long long ldd;
  double jitter;
  long long* dbg;        // optional clock64 stamps of CTA 0 (12 per panel), developer tooling
  int dbg_tid;           // thread that records them
};
cudaError_t prepare_chol(int n);
cudaError_t prepare_mcmc();   // one-time opt-ins to large dynamic shared memory (outside capture)
cudaError_t prepare_acq();
cudaError_t launch_chol(const CholArgs& A, int grid, int sms, cudaStream_t stream);
// small n: Gram + factorisation + LML fused, the matrix resident in shared memory (bgp_small.cu)
cudaError_t prepare_small();
bool small_path_fits(int n, int d, int n_leaves);
cudaError_t launch_small(const CholArgs& A, int n_leaves, int sms, cudaStream_t stream);

struct GramArgs {
  const double* X;       // n x d
  const double* alpha;   // n
  const double* theta;   // batch x p
  double* slabs;         // batch factor slabs (the lower triangle of each is written)
  double* xt;            // scratch: batch x xt_stride scaled, transposed inputs
  long long xt_stride;
  const DevProgram* prog;
  const double* fixed_ls;
  int n, d, batch, aug;
  int fused_ok;          // host: fast-path kernel without input warping -> one launch, inputs scaled in the Gram CTAs
  int sms;
};
size_t gram_xt_doubles(int n, int d, int n_leaves);
bool gram_fused_fits(int n, int d);   // the one-launch Gram kernel keeps all scaled inputs in shared memory
cudaError_t prepare_gram();
cudaError_t launch_gram(const GramArgs& A, cudaStream_t stream);
cudaError_t launch_scale_x(const GramArgs& A, cudaStream_t stream);   // the first half of launch_gram

// analytic LML gradient (bgp_grad.cu)
struct GradArgs {
  const double* xt;         // scaled transposed inputs + resolved op constants of ONE theta (launch_scale_x)
  const double* alpha_vec;  // alpha_ = K^-1 y (n)
  const double* kinv;       // K_inv_ (n x n)
  double* partial;          // grad_partial_doubles(n, p_kernel) workspace
  double* grad;             // p_kernel outputs
  const DevProgram* prog;
  int n, d, n_leaves, p_kernel;
};
size_t grad_partial_doubles(int n, int p);
cudaError_t launch_lml_grad(const GradArgs& A, cudaStream_t stream);
cudaError_t launch_warp_points(const double* X, int npts, int d, const double* theta, int S, const DevProgram* prog,
                               double* out, cudaStream_t stream);

struct SweepArgs {
  const double* X;        // n x d
  const double* theta;    // S x p
  const double* slabs;    // S factor slabs (aug)
  const double* z;        // S x n
  const double* Xc;       // m x d
  long long x_stride;     // 0, or n*d / m*d when X / Xc hold one (warped) copy per theta
  long long xc_stride;
  const double* zextra;   // S x R x n or null
  double* mu;             // S x m
  double* sd;             // S x m
  double* dots;           // S x R x m or null
  double* v_out;          // S x m x v_ld or null
  long long v_ld;
  const DevProgram* prog;
  const double* fixed_ls;
  double y_mean, y_std;
  int n, d, S, m, R, noise_off;
  int n_leaves;           // stationary leaves of the program (sizes the scaled-candidate block)
  double* ks_scratch;     // windowed mode: one k* tile per resident CTA (sweep_scratch_doubles each)
  long long ks_scratch_stride;
};
cudaError_t prepare_sweep();
size_t sweep_scratch_doubles(int n);
bool sweep_is_windowed(int n, int d, int R, int n_leaves);
cudaError_t launch_sweep(const SweepArgs& A, int sms, cudaStream_t stream);

struct ExtractArgs {
  const double* slab;
  const double* z;
  double* out;
  int n, what;
};
cudaError_t launch_extract(const ExtractArgs& A, double* scratch, cudaStream_t stream);

struct AcqArgs {
  int kind;
  const double* mu;
  const double* sd;
  int S, m;
  double p0;
  const float* u32;    // float32 standard Gumbel variates, S x K (MES)
  int K;
  double* per_theta;   // S x m output
  double* out;         // m
  int32_t* skipped;    // S
  double* mes_fit;     // S x 5 (a, b, q1, med, q2): output of the search, or input when mes_fit_given
  double* scratch;     // workspace (see acq_scratch_doubles)
  const double* yopt;  // optional per-theta y_opt (S) overriding the local min mu (EI / TTEI)
  const double* ref;   // optional per-theta {ei, index, mu, sd} of the global EI maximiser (TTEI)
  int mes_fit_given;
};
size_t acq_scratch_doubles(int S, int m);
cudaError_t launch_acq(const AcqArgs& A, cudaStream_t stream);
cudaError_t launch_acq_stats(const double* mu, const double* sd, int S, int m, double* stats_out, double* scratch,
                             cudaStream_t stream);
cudaError_t launch_mes_fit(const double* mu, const double* sd, int S, int m, double* fit_out, double* scratch,
                           cudaStream_t stream);
cudaError_t launch_ei_best(const double* mu, const double* sd, int S, int m, double p0, const double* yopt,
                           long long index_offset, double* ref_out, double* scratch, cudaStream_t stream);
cudaError_t launch_acq_per_theta(const AcqArgs& A, cudaStream_t stream);
cudaError_t launch_acq_combine(const double* per_theta, int S, int m, const int32_t* skipped, double* out,
                               cudaStream_t stream);
struct PostCovArgs {
  const double* X;       // unused (kept for symmetry)
  const double* theta;   // 1 x p
  const double* v;       // m x v_ld whitened cross-covariances
  const double* Xc;      // m x d
  double* cov;           // m x ldc
  long long v_ld, ldc;
  const DevProgram* prog;
  const double* fixed_ls;
  double y_std;
  int n, d, m, noise_off;
};
cudaError_t launch_postcov(const PostCovArgs& A, cudaStream_t stream);
struct CombineArgs {
  const DevProgram* prog;
  const double* fixed_ls;
  const double* theta;   // 1 x p
  const double* Xt;      // R x d Thompson points
  const double* Xc;      // m x d candidates
  const double* dots;    // R x m
  const double* vt;      // R x n
  const double* s;       // m Schur complements
  const double* cov;     // m x ldc (VR)
  double* out;           // m
  long long ldc;
  int R, m, n, d;
};
cudaError_t launch_pvrs_combine(const CombineArgs& A, cudaStream_t stream);
cudaError_t launch_vr_combine(const CombineArgs& A, double* scratch, cudaStream_t stream);
cudaError_t launch_slab_trmm(const double* slab, int m, const double* E, int ns, const double* mean,
                             double* out, cudaStream_t stream);
// chip-wide dense Cholesky of one large matrix + the draw that uses it (bgp_big.cu); nullptr or an error text
size_t big_workspace_bytes(int m);
const char* big_cholesky(void** cublas_handle, double* a, int m, long long lda, double jitter, int32_t* info,
                         void* workspace, int sms, cudaStream_t stream);
const char* big_trmm(void** cublas_handle, const double* l, int m, long long lda, const double* e, int ns,
                     const double* mean, double* out, cudaStream_t stream);
void big_release(void** cublas_handle);
// lazily bound NCCL (bgp_nccl.cu) for the multi-GPU C entry point
struct NcclApi {
  int (*all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*all_gather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*error_string)(int) = nullptr;
  bool ok = false;
};
NcclApi* nccl_api();
int nccl_dtype_f64();
int nccl_dtype_i32();
int nccl_op_min();
int nccl_op_max();
cudaError_t launch_pad_rows(const double* src, int S, int m_loc, double* dst, int m_max, cudaStream_t st);
cudaError_t launch_unpack_rows(const double* gathered, int world, int S, int m_max, int m_total, double* out,
                               cudaStream_t st);
cudaError_t launch_column0(const double* stats, int S, int stride, double* out, cudaStream_t st);
cudaError_t launch_ttei_pick(const double* allref, int world, int S, double* best, cudaStream_t st);
cudaError_t launch_argmax(const double* v, int m, long long* idx, double* scratch, cudaStream_t stream);

cudaError_t launch_split(int W, uint64_t seed, const uint64_t* seed_ptr, int step, int32_t* colour,
                         cudaStream_t stream);
cudaError_t launch_propose(const double* pos, const int32_t* colour, int W, int p, int half, double a,
                           uint64_t seed, const uint64_t* seed_ptr, int step, double* q, double* factors,
                           int32_t* movers, cudaStream_t stream);
cudaError_t launch_accept(double* pos, double* lp, const double* q, const double* factors,
                          const double* new_lp, const int32_t* movers, int W, int p, int half,
                          uint64_t seed, const uint64_t* seed_ptr, int step, int32_t* accepted,
                          double* chain_step, double* lp_step, cudaStream_t stream);

cudaError_t launch_split_all(int W, int T, const uint64_t* seed_ptr, int32_t* colour, cudaStream_t stream);
// accept test of one half step + proposals of the next one (next_colour null: none) in one launch
cudaError_t launch_accept_propose(double* pos, double* lp, double* q, double* factors, const double* new_lp,
                                  int32_t* movers, int W, int p, int half, const uint64_t* seed_ptr, int step,
                                  int32_t* accepted, double* chain_step, double* lp_step, const int32_t* next_colour,
                                  int next_half, int next_step, double a, cudaStream_t stream);

// peer exchange of walker log-probs (bgp_peer.cu): block[r] = rank r's exchange block as mapped here
struct PeerXchg {
  double* block[8];
  int rank, world, cap;
};
cudaError_t launch_xchg_gather(const PeerXchg& X, const double* src, int lo, int cnt, int total, double* out,
                               cudaStream_t stream);
cudaError_t launch_accept_xchg_propose(const PeerXchg& X, double* pos, double* lp, double* q, double* factors,
                                       const double* new_lp_local, int lo, int cnt, int32_t* movers, int W, int p,
                                       int half, const uint64_t* seed_ptr, int step, int32_t* accepted,
                                       double* chain_step, double* lp_step, const int32_t* next_colour, int next_half,
                                       int next_step, double a, cudaStream_t stream);
}  // namespace bgp
