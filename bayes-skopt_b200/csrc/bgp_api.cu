// C ABI of libbgp (include/bgp.h): handle, argument checking, workspace, CUDA-graph cache.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bgp_common.cuh"
#include "bgp_internal.h"

namespace {

thread_local std::string g_err;

int fail(const char* what, cudaError_t e = cudaSuccess) {
  g_err = what;
  if (e != cudaSuccess) { g_err += ": "; g_err += cudaGetErrorString(e); }
  return -1;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T* as() { return static_cast<T*>(p); }
};

struct GraphKey {
  const void *pos, *lp, *chain, *lpc, *acc;
  int W, T, n, d, p;
  int rank, world;   // 0, 1 for the single-GPU run
  double a;
  bool operator==(const GraphKey& o) const { return std::memcmp(this, &o, sizeof(GraphKey)) == 0; }
};

}  // namespace

struct bgp_handle_s {
  int device = 0, sms = 0;
  DevProgram host_prog;
  bool have_prog = false, have_priors = false, have_data = false;
  int n = 0, d = 0, n_priors = 0;
  DevBuf prog, fixed_ls, priors, X, y, alpha;
  bgp_prior_t priors_host[BGP_MAX_THETA] = {};   // what the device table holds (bgp_set_priors skips identical tables)
  DevBuf slabs_scratch;      // one factor slab per resident CTA (logprob mode)
  DevBuf xt_scratch;         // scaled inputs per resident CTA when they exceed shared memory
  DevBuf acq_scratch, extract_scratch;
  DevBuf sweep_scratch;      // windowed sweep: one k* tile per resident CTA
  DevBuf nccl_scratch;       // multi-GPU C entry point: moments, per-theta values, gather buffers
  DevBuf big_scratch;        // chip-wide dense Cholesky: slab of one diagonal block + per-block info
  void* cublas = nullptr;    // cublasHandle_t, created on first use of the chip-wide dense Cholesky
  DevBuf warp_x, warp_xc, warp_xt;   // per-theta warped copies of X / candidates / Thompson points
  DevBuf mc_colour, mc_movers, mc_q, mc_factors, mc_newlp, mc_seed;
  // multi-GPU walker sharding: this rank's exchange block and the peers' blocks as mapped here
  DevBuf xchg;
  bgp::PeerXchg peers;
  bool have_peers = false;
  uint64_t* seed_pinned = nullptr;
  cudaGraphExec_t graph = nullptr;
  GraphKey key;
  bool have_graph = false;
  const uint64_t* step_seed_dev = nullptr;   // stepped MCMC entry points: seed read from here when set
  long long* dbg = nullptr;   // developer tooling: clock stamps of the factorisation kernel
  int dbg_tid = 0;
};

#define CHECK_H(h) if (!(h)) return fail("null handle")
#define CUDA_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(#expr, _e); } while (0)

extern "C" {

const char* bgp_last_error(void) { return g_err.c_str(); }
int bgp_version(void) { return 100; }

int bgp_create(bgp_handle_t* out, int device) {
  if (!out) return fail("null out");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail("no CUDA device: libbgp has no CPU fallback", e);
  if (device < 0 || device >= count) return fail("bad device index");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("libbgp is built for sm_100a (B200) only");
  bgp_handle_s* h = new bgp_handle_s();
  h->device = device;
  h->sms = prop.multiProcessorCount;
  std::memset(&h->key, 0, sizeof(h->key));
  CUDA_TRY(bgp::prepare_mcmc());
  CUDA_TRY(bgp::prepare_acq());
  CUDA_TRY(bgp::prepare_gram());
  CUDA_TRY(bgp::prepare_sweep());
  CUDA_TRY(bgp::prepare_small());
  CUDA_TRY(cudaMallocHost((void**)&h->seed_pinned, sizeof(uint64_t)));
  CUDA_TRY(h->mc_seed.ensure(sizeof(uint64_t)));
  *out = h;
  return 0;
}

int bgp_destroy(bgp_handle_t h) {
  CHECK_H(h);
  cudaSetDevice(h->device);
  if (h->graph) cudaGraphExecDestroy(h->graph);
  bgp_peer_close(h);
  bgp::big_release(&h->cublas);
  h->big_scratch.release();
  h->nccl_scratch.release();
  h->xchg.release();
  for (DevBuf* b : {&h->prog, &h->fixed_ls, &h->priors, &h->X, &h->y, &h->alpha, &h->slabs_scratch, &h->xt_scratch,
                    &h->acq_scratch, &h->extract_scratch, &h->sweep_scratch, &h->warp_x, &h->warp_xc, &h->warp_xt, &h->mc_colour, &h->mc_movers, &h->mc_q,
                    &h->mc_factors, &h->mc_newlp, &h->mc_seed})
    b->release();
  if (h->seed_pinned) cudaFreeHost(h->seed_pinned);
  delete h;
  return 0;
}

int bgp_set_kernel(bgp_handle_t h, const bgp_op_t* ops, int n_ops, int n_theta, const double* fixed_ls,
                   int n_fixed_ls) {
  CHECK_H(h);
  if (!ops || n_ops <= 0 || n_ops > BGP_MAX_OPS) return fail("bad op count");
  if (n_theta < 0 || n_theta > BGP_MAX_THETA) return fail("bad theta count");
  CUDA_TRY(cudaSetDevice(h->device));
  DevProgram& P = h->host_prog;
  std::memset(&P, 0, sizeof(P));
  P.n_ops = n_ops; P.n_theta = n_theta;
  int leaves = 0, depth = 0;
  for (int i = 0; i < n_ops; ++i) {
    P.ops[i] = ops[i];
    const int c = ops[i].code;
    if (c >= BGP_OP_RBF && c <= BGP_OP_MATERN52) {
      if (leaves >= BGP_MAX_LEAVES) return fail("too many stationary leaves");
      P.leaf_of_op[i] = leaves++;
      ++depth;
    } else if (c == BGP_OP_CONST || c == BGP_OP_WHITE) {
      ++depth;
    } else if (c == BGP_OP_ADD || c == BGP_OP_MUL) {
      if (depth < 2) return fail("malformed postfix program");
      --depth;
    } else if (c == BGP_OP_POW) {
      if (depth < 1) return fail("malformed postfix program");
    } else {
      return fail("unknown opcode");
    }
    if (depth > 4) return fail("kernel tree too deep (stack > 4)");
  }
  if (depth != 1) return fail("malformed postfix program");
  P.n_leaves = leaves;
  if (n_ops == 5 && ops[0].code == BGP_OP_CONST && ops[1].code >= BGP_OP_RBF && ops[1].code <= BGP_OP_MATERN52 &&
      ops[2].code == BGP_OP_MUL && ops[3].code == BGP_OP_WHITE && ops[4].code == BGP_OP_ADD) {
    P.fast_kind = ops[1].code; P.fast_const = 0; P.fast_white = 3;
    P.fast_white_zeroable = (ops[3].flags & BGP_FLAG_ZEROABLE_WHITE) ? 1 : 0;
  }
  P.d = h->d;
  if (n_fixed_ls > 0) {
    CUDA_TRY(h->fixed_ls.ensure(sizeof(double) * n_fixed_ls));
    CUDA_TRY(cudaMemcpy(h->fixed_ls.p, fixed_ls, sizeof(double) * n_fixed_ls, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(h->prog.ensure(sizeof(DevProgram)));
  CUDA_TRY(cudaMemcpy(h->prog.p, &P, sizeof(DevProgram), cudaMemcpyHostToDevice));
  h->have_prog = true;
  h->have_graph = false;
  if (h->have_data) {
    cudaError_t e = bgp::prepare_chol(h->n);
    if (e != cudaSuccess) return fail("n/d too large for the shared-memory plan of the factorisation kernel", e);
  }
  return 0;
}

int bgp_set_warp(bgp_handle_t h, int n_warp) {
  CHECK_H(h);
  if (!h->have_prog) return fail("bgp_set_kernel has not been called");
  if (n_warp < 0 || n_warp > BGP_MAX_DIM) return fail("bad warp dimension count");
  CUDA_TRY(cudaSetDevice(h->device));
  DevProgram& P = h->host_prog;
  const int p_kernel = P.n_warp ? P.warp_off : P.n_theta;
  if (p_kernel + 2 * n_warp > BGP_MAX_THETA) return fail("too many hyper-parameters with input warping");
  P.n_warp = n_warp;
  P.warp_off = p_kernel;
  P.n_theta = p_kernel + 2 * n_warp;
  CUDA_TRY(cudaMemcpy(h->prog.p, &P, sizeof(DevProgram), cudaMemcpyHostToDevice));
  h->have_graph = false;
  return 0;
}

// per-theta warped copy of a point set (identity when warping is off: returns the input)
static int warped(bgp_handle_t h, DevBuf& buf, const double* pts_dev, int npts, const double* theta_dev, int S,
                  cudaStream_t st, const double** out, long long* stride) {
  if (h->host_prog.n_warp == 0) { *out = pts_dev; *stride = 0; return 0; }
  if (h->host_prog.n_warp != h->d) return fail("input warping needs one (a, b) pair per input dimension");
  CUDA_TRY(buf.ensure(sizeof(double) * (size_t)S * npts * h->d));
  CUDA_TRY(bgp::launch_warp_points(pts_dev, npts, h->d, theta_dev, S, h->prog.as<DevProgram>(), buf.as<double>(), st));
  *out = buf.as<double>();
  *stride = (long long)npts * h->d;
  return 0;
}

/* kernels one batched log-posterior call launches for the current model and data (waves of <= #SMs thetas each):
 * 1 on the fused small-n path, 2 when the Gram kernel scales its inputs itself (default kernel shape, no input
 * warping), else 3 (scale_x + gram + chol) -- lets a host keep an honest launch count */
int bgp_logprob_launches(bgp_handle_t h) {
  CHECK_H(h);
  if (!h->have_prog || !h->have_data) return 3;
  if (bgp::small_path_fits(h->n, h->d, h->host_prog.n_leaves)) return 1;
  return (h->host_prog.fast_kind != 0 && h->host_prog.n_warp == 0 && bgp::gram_fused_fits(h->n, h->d)) ? 2 : 3;
}

int bgp_set_priors(bgp_handle_t h, const bgp_prior_t* priors, int n_priors) {
  CHECK_H(h);
  CUDA_TRY(cudaSetDevice(h->device));
  if (n_priors <= 0 || !priors) {
    if (h->have_priors) h->have_graph = false;
    h->have_priors = false; h->n_priors = 0;
    return 0;
  }
  if (n_priors > BGP_MAX_THETA) return fail("too many priors");
  // the captured MCMC graph holds the table's address and length, not its contents: it stays valid
  // unless one of those changes; an identical table (every sample() call sets it) costs nothing
  if (h->have_priors && h->n_priors == n_priors &&
      std::memcmp(h->priors_host, priors, sizeof(bgp_prior_t) * n_priors) == 0)
    return 0;
  void* before = h->priors.p;
  CUDA_TRY(h->priors.ensure(sizeof(bgp_prior_t) * n_priors));
  CUDA_TRY(cudaMemcpy(h->priors.p, priors, sizeof(bgp_prior_t) * n_priors, cudaMemcpyHostToDevice));
  std::memset(h->priors_host, 0, sizeof(h->priors_host));
  std::memcpy(h->priors_host, priors, sizeof(bgp_prior_t) * n_priors);
  if (!h->have_priors || h->n_priors != n_priors || h->priors.p != before) h->have_graph = false;
  h->n_priors = n_priors;
  h->have_priors = true;
  return 0;
}

int bgp_set_data(bgp_handle_t h, const double* X_dev, const double* y_dev, const double* alpha_dev, int n,
                 int d, void* stream) {
  CHECK_H(h);
  if (!X_dev || !y_dev || !alpha_dev || n <= 0 || d <= 0 || d > BGP_MAX_DIM) return fail("bad data arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(h->X.ensure(sizeof(double) * n * d));
  CUDA_TRY(h->y.ensure(sizeof(double) * n));
  CUDA_TRY(h->alpha.ensure(sizeof(double) * n));
  CUDA_TRY(cudaMemcpyAsync(h->X.p, X_dev, sizeof(double) * n * d, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(h->y.p, y_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(h->alpha.p, alpha_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
  if (h->n != n || h->d != d) h->have_graph = false;
  h->n = n; h->d = d;
  {
    cudaError_t e = bgp::prepare_chol(n);
    if (e != cudaSuccess) return fail("n/d too large for the shared-memory plan of the factorisation kernel", e);
  }
  if (h->have_prog && h->host_prog.d != d) {
    h->host_prog.d = d;
    CUDA_TRY(cudaMemcpyAsync(h->prog.p, &h->host_prog, sizeof(DevProgram), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  h->have_data = true;
  return 0;
}

static int ready(bgp_handle_t h) {
  if (!h->have_prog) return fail("bgp_set_kernel has not been called");
  if (!h->have_data) return fail("bgp_set_data has not been called");
  if (h->host_prog.d != h->d) return fail("internal: program dimension mismatch");
  return 0;
}

static int slots_for(bgp_handle_t h) { return h->n <= 64 ? 2 * h->sms : h->sms; }

static int ensure_logprob_workspace(bgp_handle_t h) {
  const SlabGeom G = SlabGeom::make(h->n, false);
  const int slots = slots_for(h);
  CUDA_TRY(h->slabs_scratch.ensure(sizeof(double) * (size_t)G.doubles() * slots));
  CUDA_TRY(h->xt_scratch.ensure(sizeof(double) * bgp::gram_xt_doubles(h->n, h->d, h->host_prog.n_leaves) * slots));
  return 0;
}

// K1 (Gram, all SMs) then K2 (factorisation, one CTA per theta), in waves of one slab per SM
static int logprob_impl(bgp_handle_t h, const double* theta_dev, int batch, const double* lp_extra_dev,
                        double* lp_dev, double* lml_dev, int32_t* info_dev, cudaStream_t st) {
  if (bgp::small_path_fits(h->n, h->d, h->host_prog.n_leaves)) {
    // small n: one fused launch, the matrix never leaves shared memory
    bgp::CholArgs A;
    std::memset(&A, 0, sizeof(A));
    A.X = h->X.as<double>(); A.y = h->y.as<double>(); A.alpha = h->alpha.as<double>();
    A.theta = theta_dev; A.lp_extra = lp_extra_dev; A.lp = lp_dev; A.lml = lml_dev; A.info = info_dev;
    A.prog = h->prog.as<DevProgram>(); A.fixed_ls = h->fixed_ls.as<double>();
    A.priors = h->have_priors ? h->priors.as<bgp_prior_t>() : nullptr; A.n_priors = h->n_priors;
    A.n = h->n; A.d = h->d; A.batch = batch;
    A.dbg = h->dbg; A.dbg_tid = h->dbg_tid;
    CUDA_TRY(bgp::launch_small(A, h->host_prog.n_leaves, h->sms, st));
    return 0;
  }
  if (ensure_logprob_workspace(h)) return -1;
  const int slots = slots_for(h), p = h->host_prog.n_theta;
  const size_t xt = bgp::gram_xt_doubles(h->n, h->d, h->host_prog.n_leaves);
  for (int b0 = 0; b0 < batch; b0 += slots) {
    const int nb = batch - b0 < slots ? batch - b0 : slots;
    bgp::GramArgs Gm{h->X.as<double>(), h->alpha.as<double>(), theta_dev + (size_t)b0 * p,
                     h->slabs_scratch.as<double>(), h->xt_scratch.as<double>(), (long long)xt,
                     h->prog.as<DevProgram>(), h->fixed_ls.as<double>(), h->n, h->d, nb, 0,
                     h->host_prog.fast_kind != 0 && h->host_prog.n_warp == 0, h->sms};
    CUDA_TRY(bgp::launch_gram(Gm, st));
    bgp::CholArgs A;
    std::memset(&A, 0, sizeof(A));
    A.X = h->X.as<double>(); A.y = h->y.as<double>(); A.alpha = h->alpha.as<double>();
    A.theta = theta_dev + (size_t)b0 * p;
    A.lp_extra = lp_extra_dev ? lp_extra_dev + b0 : nullptr;
    A.lp = lp_dev ? lp_dev + b0 : nullptr;
    A.lml = lml_dev ? lml_dev + b0 : nullptr;
    A.info = info_dev ? info_dev + b0 : nullptr;
    A.slabs = h->slabs_scratch.as<double>();
    A.prog = h->prog.as<DevProgram>(); A.fixed_ls = h->fixed_ls.as<double>();
    A.priors = h->have_priors ? h->priors.as<bgp_prior_t>() : nullptr; A.n_priors = h->n_priors;
    A.n = h->n; A.d = h->d; A.batch = nb; A.aug = 0; A.slab_per_block = 0;
    A.dbg = h->dbg; A.dbg_tid = h->dbg_tid;
    CUDA_TRY(bgp::launch_chol(A, nb, h->sms, st));
  }
  return 0;
}

int bgp_logprob_batched(bgp_handle_t h, const double* theta_dev, int batch, const double* lp_extra_dev,
                        double* lp_dev, double* lml_dev, int32_t* info_dev, void* stream) {
  CHECK_H(h);
  if (ready(h)) return -1;
  if (!theta_dev || batch <= 0 || !lp_dev) return fail("bad logprob arguments");
  if (h->have_priors && h->n_priors != h->host_prog.n_theta) return fail("prior count != theta count");
  CUDA_TRY(cudaSetDevice(h->device));
  return logprob_impl(h, theta_dev, batch, lp_extra_dev, lp_dev, lml_dev, info_dev, (cudaStream_t)stream);
}

/* developer tooling (not in include/bgp.h): device buffer of 8 clock64 stamps per panel */
int bgp_debug_set_stamps(bgp_handle_t h, long long* stamps_dev, int tid) {
  CHECK_H(h);
  h->dbg = stamps_dev;
  h->dbg_tid = tid;
  return 0;
}

int64_t bgp_factor_slab_doubles(bgp_handle_t h) {
  if (!h || !h->have_data) return -1;
  return SlabGeom::make(h->n, true).doubles();
}

int bgp_factorize_batched(bgp_handle_t h, const double* theta_dev, int S, double* slabs_dev, double* z_dev,
                          double* lml_dev, int32_t* info_dev, void* stream) {
  CHECK_H(h);
  if (ready(h)) return -1;
  if (!theta_dev || S <= 0 || !slabs_dev || !z_dev) return fail("bad factorize arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t xt = bgp::gram_xt_doubles(h->n, h->d, h->host_prog.n_leaves);
  CUDA_TRY(h->xt_scratch.ensure(sizeof(double) * xt * (size_t)(S > slots_for(h) ? S : slots_for(h))));
  bgp::GramArgs Gm{h->X.as<double>(), h->alpha.as<double>(), theta_dev, slabs_dev, h->xt_scratch.as<double>(),
                   (long long)xt, h->prog.as<DevProgram>(), h->fixed_ls.as<double>(), h->n, h->d, S, 1,
                   h->host_prog.fast_kind != 0 && h->host_prog.n_warp == 0, h->sms};
  CUDA_TRY(bgp::launch_gram(Gm, st));
  bgp::CholArgs A;
  std::memset(&A, 0, sizeof(A));
  A.X = h->X.as<double>(); A.y = h->y.as<double>(); A.alpha = h->alpha.as<double>();
  A.theta = theta_dev; A.lml = lml_dev; A.info = info_dev; A.slabs = slabs_dev; A.z_out = z_dev;
  A.prog = h->prog.as<DevProgram>(); A.fixed_ls = h->fixed_ls.as<double>();
  A.n = h->n; A.d = h->d; A.batch = S; A.aug = 1; A.slab_per_block = 0;
  CUDA_TRY(bgp::launch_chol(A, S, h->sms, st));
  return 0;
}

int bgp_factor_extract(bgp_handle_t h, const double* slab_dev, const double* z_dev, int what, double* out_dev,
                       void* stream) {
  CHECK_H(h);
  if (ready(h)) return -1;
  if (!slab_dev || !out_dev) return fail("bad extract arguments");
  if (what == BGP_EXTRACT_ALPHA && !z_dev) return fail("alpha extraction needs z");
  CUDA_TRY(cudaSetDevice(h->device));
  double* scratch = nullptr;
  if (what == BGP_EXTRACT_KINV) {
    CUDA_TRY(h->extract_scratch.ensure(sizeof(double) * (size_t)h->n * h->n));
    scratch = h->extract_scratch.as<double>();
  }
  bgp::ExtractArgs A{slab_dev, z_dev, out_dev, h->n, what};
  CUDA_TRY(bgp::launch_extract(A, scratch, (cudaStream_t)stream));
  return 0;
}

int bgp_lml_gradient(bgp_handle_t h, const double* theta_dev, const double* alpha_dev, const double* kinv_dev,
                     double* grad_dev, void* stream) {
  CHECK_H(h);
  if (ready(h)) return -1;
  if (!theta_dev || !alpha_dev || !kinv_dev || !grad_dev) return fail("bad gradient arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const DevProgram& P = h->host_prog;
  const int p_kernel = P.n_warp ? P.warp_off : P.n_theta;
  if (p_kernel <= 0) return 0;
  const size_t xt = bgp::gram_xt_doubles(h->n, h->d, P.n_leaves);
  CUDA_TRY(h->xt_scratch.ensure(sizeof(double) * xt * (size_t)slots_for(h)));
  CUDA_TRY(h->extract_scratch.ensure(sizeof(double) * ((size_t)h->n * h->n + bgp::grad_partial_doubles(h->n, p_kernel))));
  bgp::GramArgs Gm{h->X.as<double>(), h->alpha.as<double>(), theta_dev, nullptr, h->xt_scratch.as<double>(),
                   (long long)xt, h->prog.as<DevProgram>(), h->fixed_ls.as<double>(), h->n, h->d, 1, 0, 0, h->sms};
  CUDA_TRY(bgp::launch_scale_x(Gm, st));
  // the partial sums live behind the n x n region that bgp_factor_extract(K_INV) uses as its own scratch
  bgp::GradArgs A{h->xt_scratch.as<double>(), alpha_dev, kinv_dev, h->extract_scratch.as<double>() + (size_t)h->n * h->n,
                  grad_dev, h->prog.as<DevProgram>(), h->n, h->d, P.n_leaves, p_kernel};
  CUDA_TRY(bgp::launch_lml_grad(A, st));
  return 0;
}

int bgp_predict_batched(bgp_handle_t h, const double* theta_dev, int S, const double* slabs_dev,
                        const double* z_dev, const double* Xc_dev, int m, int noise_off, double y_mean,
                        double y_std, double* mu_dev, double* sd_dev, const double* zextra_dev, int R,
                        double* dots_dev, double* v_dev, int64_t v_ld, void* stream) {
  CHECK_H(h);
  if (ready(h)) return -1;
  if (!theta_dev || S <= 0 || !slabs_dev || !z_dev || !Xc_dev || m <= 0 || !mu_dev || !sd_dev)
    return fail("bad predict arguments");
  if (R < 0 || (R > 0 && (!zextra_dev || !dots_dev))) return fail("bad extra right-hand sides");
  if (v_dev && (v_ld < 32 * ((h->n + 31) / 32) || (v_ld & 3))) return fail("bad v_ld");
  CUDA_TRY(cudaSetDevice(h->device));
  bgp::SweepArgs A;
  if (warped(h, h->warp_x, h->X.as<double>(), h->n, theta_dev, S, (cudaStream_t)stream, &A.X, &A.x_stride)) return -1;
  if (warped(h, h->warp_xc, Xc_dev, m, theta_dev, S, (cudaStream_t)stream, &A.Xc, &A.xc_stride)) return -1;
  A.theta = theta_dev; A.slabs = slabs_dev; A.z = z_dev;
  A.zextra = zextra_dev; A.mu = mu_dev; A.sd = sd_dev; A.dots = dots_dev; A.v_out = v_dev; A.v_ld = v_ld;
  A.prog = h->prog.as<DevProgram>(); A.fixed_ls = h->fixed_ls.as<double>();
  A.y_mean = y_mean; A.y_std = y_std; A.n = h->n; A.d = h->d; A.S = S; A.m = m; A.R = R;
  A.noise_off = noise_off;
  A.n_leaves = h->host_prog.n_leaves; A.ks_scratch = nullptr; A.ks_scratch_stride = 0;
  if (bgp::sweep_is_windowed(h->n, h->d, R, h->host_prog.n_leaves)) {
    const size_t per_cta = bgp::sweep_scratch_doubles(h->n);
    CUDA_TRY(h->sweep_scratch.ensure(sizeof(double) * per_cta * (size_t)h->sms));
    A.ks_scratch = h->sweep_scratch.as<double>(); A.ks_scratch_stride = (long long)per_cta;
  }
  cudaError_t e = bgp::launch_sweep(A, h->sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail("sweep launch (n too large for the shared-memory resident tile?)", e);
  return 0;
}

int bgp_acq_sweep(bgp_handle_t h, int kind, const double* mu_dev, const double* sd_dev, int S, int m, double p0,
                  const float* g32_dev, int K, double* per_theta_dev, double* out_dev, int32_t* skipped_dev,
                  double* mes_fit_dev, void* stream) {
  CHECK_H(h);
  if (!mu_dev || !sd_dev || S <= 0 || S > 1024 || m <= 0 || !per_theta_dev || !out_dev || !skipped_dev)
    return fail("bad acquisition arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(h->acq_scratch.ensure(sizeof(double) * bgp::acq_scratch_doubles(S, m)));
  bgp::AcqArgs A{kind, mu_dev, sd_dev, S, m, p0, g32_dev, K, per_theta_dev, out_dev, skipped_dev, mes_fit_dev,
                 h->acq_scratch.as<double>(), nullptr, nullptr, 0};
  CUDA_TRY(bgp::launch_acq(A, (cudaStream_t)stream));
  return 0;
}

int bgp_acq_stats(bgp_handle_t h, const double* mu_dev, const double* sd_dev, int S, int m, double* stats_dev,
                  void* stream) {
  CHECK_H(h);
  if (!mu_dev || !sd_dev || S <= 0 || S > 1024 || m <= 0 || !stats_dev) return fail("bad stats arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(h->acq_scratch.ensure(sizeof(double) * bgp::acq_scratch_doubles(S, m)));
  CUDA_TRY(bgp::launch_acq_stats(mu_dev, sd_dev, S, m, stats_dev, h->acq_scratch.as<double>(), (cudaStream_t)stream));
  return 0;
}

int bgp_mes_fit(bgp_handle_t h, const double* mu_dev, const double* sd_dev, int S, int m, double* fit_dev,
                void* stream) {
  CHECK_H(h);
  if (!mu_dev || !sd_dev || S <= 0 || S > 1024 || m <= 0 || !fit_dev) return fail("bad mes-fit arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(h->acq_scratch.ensure(sizeof(double) * bgp::acq_scratch_doubles(S, m)));
  CUDA_TRY(bgp::launch_acq_stats(mu_dev, sd_dev, S, m, nullptr, h->acq_scratch.as<double>(), (cudaStream_t)stream));
  CUDA_TRY(bgp::launch_mes_fit(mu_dev, sd_dev, S, m, fit_dev, h->acq_scratch.as<double>(), (cudaStream_t)stream));
  return 0;
}

int bgp_ei_best(bgp_handle_t h, const double* mu_dev, const double* sd_dev, int S, int m, double p0,
                const double* yopt_dev, int64_t index_offset, double* ref_dev, void* stream) {
  CHECK_H(h);
  if (!mu_dev || !sd_dev || S <= 0 || S > 1024 || m <= 0 || !ref_dev) return fail("bad ei-best arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(h->acq_scratch.ensure(sizeof(double) * bgp::acq_scratch_doubles(S, m)));
  if (!yopt_dev)
    CUDA_TRY(bgp::launch_acq_stats(mu_dev, sd_dev, S, m, nullptr, h->acq_scratch.as<double>(), (cudaStream_t)stream));
  CUDA_TRY(bgp::launch_ei_best(mu_dev, sd_dev, S, m, p0, yopt_dev, index_offset, ref_dev,
                               h->acq_scratch.as<double>(), (cudaStream_t)stream));
  return 0;
}

int bgp_acq_per_theta(bgp_handle_t h, int kind, const double* mu_dev, const double* sd_dev, int S, int m,
                      double p0, const double* yopt_dev, const double* ref_dev, const float* g32_dev, int K,
                      const double* fit_dev, double* per_theta_dev, int32_t* skipped_dev, void* stream) {
  CHECK_H(h);
  if (!mu_dev || !sd_dev || S <= 0 || S > 1024 || m <= 0 || !per_theta_dev || !skipped_dev)
    return fail("bad per-theta acquisition arguments");
  if (kind == BGP_ACQ_MES && !fit_dev) return fail("MES needs the Gumbel fit (bgp_mes_fit)");
  if (kind == BGP_ACQ_TTEI && !ref_dev) return fail("TTEI needs the EI maximiser (bgp_ei_best)");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(h->acq_scratch.ensure(sizeof(double) * bgp::acq_scratch_doubles(S, m)));
  if (kind == BGP_ACQ_EI && !yopt_dev && isnan(p0))
    CUDA_TRY(bgp::launch_acq_stats(mu_dev, sd_dev, S, m, nullptr, h->acq_scratch.as<double>(), (cudaStream_t)stream));
  bgp::AcqArgs A{kind, mu_dev, sd_dev, S, m, p0, g32_dev, K, per_theta_dev, nullptr, skipped_dev,
                 const_cast<double*>(fit_dev), h->acq_scratch.as<double>(), yopt_dev, ref_dev, 1};
  CUDA_TRY(bgp::launch_acq_per_theta(A, (cudaStream_t)stream));
  return 0;
}

int bgp_acq_combine(bgp_handle_t h, const double* per_theta_dev, int S, int m, const int32_t* skipped_dev,
                    double* out_dev, void* stream) {
  CHECK_H(h);
  if (!per_theta_dev || S <= 0 || m <= 0 || !skipped_dev || !out_dev) return fail("bad combine arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(bgp::launch_acq_combine(per_theta_dev, S, m, skipped_dev, out_dev, (cudaStream_t)stream));
  return 0;
}

/* ---- multi-GPU sweep for C callers: candidates sharded over the ranks of an ncclComm_t ---- */
#define NCCL_TRY(expr) do { int _r = (expr); if (_r != 0) { \
    g_err = std::string(#expr) + ": " + (N->error_string ? N->error_string(_r) : "NCCL error"); return -1; } } while (0)

int bgp_acq_sweep_nccl(bgp_handle_t h, void* nccl_comm, int rank, int world, int kind, const double* theta_dev, int S,
                       const double* slabs_dev, const double* z_dev, const double* Xc_all_dev, int m_total, double p0,
                       const float* g32_dev, int K, double y_mean, double y_std, double* out_dev, int64_t* argmax_dev,
                       void* stream) {
  CHECK_H(h);
  if (ready(h)) return -1;
  if (!nccl_comm || world <= 0 || rank < 0 || rank >= world) return fail("bad communicator arguments");
  if (!theta_dev || S <= 0 || S > 1024 || !slabs_dev || !z_dev || !Xc_all_dev || m_total < world || !out_dev)
    return fail("bad sharded-sweep arguments (every rank needs at least one candidate)");
  if (kind < BGP_ACQ_EI || kind > BGP_ACQ_MES) return fail("unknown acquisition");
  if (kind == BGP_ACQ_MES && (!g32_dev || K <= 0)) return fail("MES needs the float32 Gumbel variates");
  bgp::NcclApi* N = bgp::nccl_api();
  if (!N) return fail("libnccl.so.2 could not be loaded");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int q = m_total / world, rem = m_total % world;
  const int lo = rank * q + (rank < rem ? rank : rem), m_loc = q + (rank < rem ? 1 : 0), m_max = q + (rem ? 1 : 0);
  const int F64 = bgp::nccl_dtype_f64(), I32 = bgp::nccl_dtype_i32();
  // workspace: mu, sd, per-theta (S x m_loc each) | send (S x m_max) | gathered (world x S x m_max) |
  // mu_all, sd_all (S x m_total each, MES) | stats S x 4 | yopt S | fit S x 5 | ref S x 4 | allref world x S x 4 |
  // local values m_loc | skipped S (int32)
  const size_t Sm = (size_t)S * m_loc, Smax = (size_t)S * m_max, Sall = (size_t)S * m_total;
  size_t need = 3 * Sm + Smax + (size_t)world * Smax + 2 * Sall + (size_t)S * (4 + 1 + 5 + 4) + (size_t)world * S * 4 +
                m_loc + S + 16;
  CUDA_TRY(h->nccl_scratch.ensure(sizeof(double) * need));
  double* w = h->nccl_scratch.as<double>();
  double *mu = w, *sd = mu + Sm, *per = sd + Sm, *send = per + Sm, *gath = send + Smax, *mu_all = gath + (size_t)world * Smax,
         *sd_all = mu_all + Sall, *stats = sd_all + Sall, *yopt = stats + (size_t)S * 4, *fit = yopt + S,
         *ref = fit + (size_t)S * 5, *allref = ref + (size_t)S * 4, *vals = allref + (size_t)world * S * 4;
  int32_t* skipped = reinterpret_cast<int32_t*>(vals + m_loc);

  if (bgp_predict_batched(h, theta_dev, S, slabs_dev, z_dev, Xc_all_dev + (size_t)lo * h->d, m_loc, 1, y_mean, y_std, mu, sd,
                          nullptr, 0, nullptr, nullptr, 0, stream)) return -1;
  auto gather_rows = [&](const double* src, int rows, double* dst_all) -> int {
    // (rows x m_loc) blocks of every rank -> (rows x m_total) on every rank
    CUDA_TRY(bgp::launch_pad_rows(src, rows, m_loc, send, m_max, st));
    NCCL_TRY(N->all_gather(send, gath, (size_t)rows * m_max, F64, nccl_comm, st));
    CUDA_TRY(bgp::launch_unpack_rows(gath, world, rows, m_max, m_total, dst_all, st));
    return 0;
  };
  const double* yopt_p = nullptr;
  const double* ref_p = nullptr;
  const double* fit_p = nullptr;
  if ((kind == BGP_ACQ_EI || kind == BGP_ACQ_TTEI) && isnan(p0)) {
    // y_opt = min over ALL candidates of the posterior mean, per theta (bask/acquisition.py:166-167)
    if (bgp_acq_stats(h, mu, sd, S, m_loc, stats, stream)) return -1;
    CUDA_TRY(bgp::launch_column0(stats, S, 4, yopt, st));
    NCCL_TRY(N->all_reduce(yopt, yopt, (size_t)S, F64, bgp::nccl_op_min(), nccl_comm, st));
    yopt_p = yopt;
  }
  if (kind == BGP_ACQ_TTEI) {
    if (bgp_ei_best(h, mu, sd, S, m_loc, p0, yopt_p, (int64_t)lo, ref, stream)) return -1;
    NCCL_TRY(N->all_gather(ref, allref, (size_t)S * 4, F64, nccl_comm, st));
    CUDA_TRY(bgp::launch_ttei_pick(allref, world, S, ref, st));
    ref_p = ref;
  }
  if (kind == BGP_ACQ_MES) {
    // the Gumbel fit needs the moments of all candidates: gathered, then fitted on every rank (same bits)
    if (gather_rows(mu, S, mu_all) || gather_rows(sd, S, sd_all)) return -1;
    if (bgp_mes_fit(h, mu_all, sd_all, S, m_total, fit, stream)) return -1;
    fit_p = fit;
  }
  if (bgp_acq_per_theta(h, kind, mu, sd, S, m_loc, p0, yopt_p, ref_p, g32_dev, K, fit_p, per, skipped, stream)) return -1;
  // a theta is dropped if ANY candidate anywhere is non-finite (bask/acquisition.py:140-141)
  NCCL_TRY(N->all_reduce(skipped, skipped, (size_t)S, I32, bgp::nccl_op_max(), nccl_comm, st));
  if (bgp_acq_combine(h, per, S, m_loc, skipped, vals, stream)) return -1;
  if (gather_rows(vals, 1, out_dev)) return -1;
  if (argmax_dev && bgp_argmax(h, out_dev, m_total, argmax_dev, stream)) return -1;
  return 0;
}

int bgp_argmax(bgp_handle_t h, const double* v_dev, int m, int64_t* idx_dev, void* stream) {
  CHECK_H(h);
  if (!v_dev || m <= 0 || !idx_dev) return fail("bad argmax arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(bgp::launch_argmax(v_dev, m, reinterpret_cast<long long*>(idx_dev), nullptr, (cudaStream_t)stream));
  return 0;
}

int bgp_posterior_cov(bgp_handle_t h, const double* theta_dev, const double* v_dev, const double* Xc_dev, int m,
                      int64_t v_ld, int noise_off, double y_std, double* cov_dev, int64_t ldc, void* stream) {
  CHECK_H(h);
  if (ready(h)) return -1;
  if (!theta_dev || !v_dev || !Xc_dev || m <= 0 || !cov_dev || ldc < m || (v_ld & 1)) return fail("bad posterior-cov arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  bgp::PostCovArgs A;
  long long unused_stride = 0;
  if (warped(h, h->warp_xc, Xc_dev, m, theta_dev, 1, (cudaStream_t)stream, &A.Xc, &unused_stride)) return -1;
  A.X = h->X.as<double>(); A.theta = theta_dev; A.v = v_dev; A.cov = cov_dev; A.v_ld = v_ld;
  A.ldc = ldc; A.prog = h->prog.as<DevProgram>(); A.fixed_ls = h->fixed_ls.as<double>(); A.y_std = y_std;
  A.n = h->n; A.d = h->d; A.m = m; A.noise_off = noise_off;
  CUDA_TRY(bgp::launch_postcov(A, (cudaStream_t)stream));
  return 0;
}

int64_t bgp_dense_slab_doubles(int m) { return m > 0 ? SlabGeom::make(m, false).doubles() : -1; }

int bgp_dense_cholesky(bgp_handle_t h, const double* a_dev, int m, int64_t lda, double jitter, double* slab_dev,
                       int32_t* info_dev, void* stream) {
  CHECK_H(h);
  if (!a_dev || m <= 0 || lda < m || !slab_dev || !info_dev) return fail("bad dense-cholesky arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  if (bgp::prepare_chol(m) != cudaSuccess) return fail("m too large for the shared-memory plan of the factorisation kernel");
  bgp::CholArgs A;
  std::memset(&A, 0, sizeof(A));
  A.slabs = slab_dev; A.info = info_dev; A.n = m; A.d = 1; A.batch = 1; A.aug = 0; A.slab_per_block = 0;
  A.dense = a_dev; A.ldd = lda; A.jitter = jitter;
  CUDA_TRY(bgp::launch_chol(A, 1, h->sms, (cudaStream_t)stream));
  return 0;
}

int bgp_dense_cholesky_inplace(bgp_handle_t h, double* a_dev, int m, int64_t lda, double jitter, int32_t* info_dev,
                               void* stream) {
  CHECK_H(h);
  if (!a_dev || m <= 0 || lda < m || !info_dev) return fail("bad dense-cholesky arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(h->big_scratch.ensure(bgp::big_workspace_bytes(m)));
  const char* err = bgp::big_cholesky(&h->cublas, a_dev, m, lda, jitter, info_dev, h->big_scratch.p, h->sms,
                                      (cudaStream_t)stream);
  return err ? fail(err) : 0;
}

int bgp_dense_trmm(bgp_handle_t h, const double* l_dev, int m, int64_t lda, const double* e_dev, int ns,
                   const double* mean_dev, double* out_dev, void* stream) {
  CHECK_H(h);
  if (!l_dev || m <= 0 || lda < m || !e_dev || ns <= 0 || !out_dev) return fail("bad trmm arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  const char* err = bgp::big_trmm(&h->cublas, l_dev, m, lda, e_dev, ns, mean_dev, out_dev, (cudaStream_t)stream);
  return err ? fail(err) : 0;
}

int bgp_slab_trmm(bgp_handle_t h, const double* slab_dev, int m, const double* e_dev, int ns,
                  const double* mean_dev, double* out_dev, void* stream) {
  CHECK_H(h);
  if (!slab_dev || m <= 0 || !e_dev || ns <= 0 || !out_dev) return fail("bad trmm arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(bgp::launch_slab_trmm(slab_dev, m, e_dev, ns, mean_dev, out_dev, (cudaStream_t)stream));
  return 0;
}

int bgp_pvrs_combine(bgp_handle_t h, const double* theta_dev, const double* xt_dev, int R, const double* xc_dev,
                     int m, const double* dots_dev, const double* vt_dev, const double* s_dev, double* out_dev,
                     void* stream) {
  CHECK_H(h);
  if (ready(h)) return -1;
  if (!theta_dev || !xt_dev || R <= 0 || !xc_dev || m <= 0 || !dots_dev || !vt_dev || !s_dev || !out_dev)
    return fail("bad pvrs arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  long long unused_stride = 0;
  if (warped(h, h->warp_xt, xt_dev, R, theta_dev, 1, (cudaStream_t)stream, &xt_dev, &unused_stride)) return -1;
  if (warped(h, h->warp_xc, xc_dev, m, theta_dev, 1, (cudaStream_t)stream, &xc_dev, &unused_stride)) return -1;
  bgp::CombineArgs A{h->prog.as<DevProgram>(), h->fixed_ls.as<double>(), theta_dev, xt_dev, xc_dev, dots_dev,
                     vt_dev, s_dev, nullptr, out_dev, 0, R, m, h->n, h->d};
  CUDA_TRY(bgp::launch_pvrs_combine(A, (cudaStream_t)stream));
  return 0;
}

int bgp_vr_combine(bgp_handle_t h, const double* cov_dev, int m, int64_t ldc, const double* xc_dev,
                   const double* theta_dev, const double* s_dev, double* out_dev, void* stream) {
  CHECK_H(h);
  if (ready(h)) return -1;
  if (!cov_dev || m <= 0 || ldc < m || !theta_dev || !s_dev || !out_dev) return fail("bad vr arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(h->acq_scratch.ensure(sizeof(double) * 64));
  long long unused_stride = 0;
  if (xc_dev && warped(h, h->warp_xc, xc_dev, m, theta_dev, 1, (cudaStream_t)stream, &xc_dev, &unused_stride)) return -1;
  bgp::CombineArgs A{h->prog.as<DevProgram>(), h->fixed_ls.as<double>(), theta_dev, nullptr, xc_dev, nullptr,
                     nullptr, s_dev, cov_dev, out_dev, ldc, 0, m, h->n, h->d};
  CUDA_TRY(bgp::launch_vr_combine(A, h->acq_scratch.as<double>(), (cudaStream_t)stream));
  return 0;
}

int bgp_mcmc_split(bgp_handle_t h, int W, uint64_t seed, int step, int32_t* colour_dev, void* stream) {
  CHECK_H(h);
  if (W < 2 || W > 8192 || !colour_dev) return fail("bad split arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(bgp::launch_split(W, seed, h->step_seed_dev, step, colour_dev, (cudaStream_t)stream));
  return 0;
}

int bgp_mcmc_propose(bgp_handle_t h, const double* pos_dev, const int32_t* colour_dev, int W, int half,
                     double a, uint64_t seed, int step, double* q_dev, double* factors_dev,
                     int32_t* movers_dev, void* stream) {
  CHECK_H(h);
  if (!h->have_prog) return fail("bgp_set_kernel has not been called");
  if (!pos_dev || !colour_dev || W < 2 || W > 8192 || !q_dev || !factors_dev || !movers_dev)
    return fail("bad propose arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(bgp::launch_propose(pos_dev, colour_dev, W, h->host_prog.n_theta, half, a, seed, h->step_seed_dev, step,
                               q_dev, factors_dev, movers_dev, (cudaStream_t)stream));
  return 0;
}

int bgp_mcmc_accept(bgp_handle_t h, double* pos_dev, double* lp_dev, const double* q_dev,
                    const double* factors_dev, const double* new_lp_dev, const int32_t* movers_dev, int W,
                    int half, uint64_t seed, int step, int32_t* accepted_dev, double* chain_step_dev,
                    double* lp_step_dev, void* stream) {
  CHECK_H(h);
  if (!h->have_prog) return fail("bgp_set_kernel has not been called");
  if (!pos_dev || !lp_dev || !q_dev || !factors_dev || !new_lp_dev || !movers_dev) return fail("bad accept arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(bgp::launch_accept(pos_dev, lp_dev, q_dev, factors_dev, new_lp_dev, movers_dev, W,
                              h->host_prog.n_theta, half, seed, h->step_seed_dev, step, accepted_dev, chain_step_dev,
                              lp_step_dev, (cudaStream_t)stream));
  return 0;
}

int bgp_mcmc_seed_source(bgp_handle_t h, const uint64_t* seed_dev) {
  CHECK_H(h);
  h->step_seed_dev = seed_dev;
  return 0;
}

// slice [lo, hi) of `count` items owned by `rank` (the first count % world ranks get one extra)
static void shard(int count, int world, int rank, int* lo, int* cnt) {
  const int q = count / world, r = count % world;
  *lo = rank * q + (rank < r ? rank : r);
  *cnt = q + (rank < r ? 1 : 0);
}

// sharded = every rank evaluates its slice of the proposals and the accept kernel starts with the peer
// exchange (bgp_mcmc.cu); the chain is identical on every rank and identical to the single-GPU one
static int mcmc_enqueue(bgp_handle_t h, double* pos, double* lp, int W, int T, double a, double* chain,
                        double* lpc, int32_t* acc, bool sharded, cudaStream_t st) {
  const int p = h->host_prog.n_theta;
  const uint64_t* sp = h->mc_seed.as<uint64_t>();
  const int world = sharded ? h->peers.world : 1, rank = sharded ? h->peers.rank : 0;
  int lo = 0, cnt = W;
  if (acc) CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(int32_t) * W, st));
  if (sharded) {
    shard(W, world, rank, &lo, &cnt);
    if (cnt > 0 && logprob_impl(h, pos + (size_t)lo * p, cnt, nullptr, h->mc_newlp.as<double>(), nullptr, nullptr, st))
      return -1;
    CUDA_TRY(bgp::launch_xchg_gather(h->peers, h->mc_newlp.as<double>(), lo, cnt, W, lp, st));
  } else if (logprob_impl(h, pos, W, nullptr, lp, nullptr, nullptr, st)) {
    return -1;
  }
  // colours of every step in one launch, the first proposals, then per half step: batched log-posterior of the
  // proposals + ONE launch for the accept test and the next half step's proposals
  if (T <= 0) return 0;
  int32_t* colours = h->mc_colour.as<int32_t>();
  CUDA_TRY(bgp::launch_split_all(W, T, sp, colours, st));
  CUDA_TRY(bgp::launch_propose(pos, colours, W, p, 0, a, 0, sp, 0, h->mc_q.as<double>(), h->mc_factors.as<double>(),
                               h->mc_movers.as<int32_t>(), st));
  for (int t = 0; t < T; ++t) {
    for (int half = 0; half < 2; ++half) {
      const int ns = half == 0 ? (W + 1) / 2 : W / 2;
      double* chain_t = (half == 1 && chain) ? chain + (size_t)t * W * p : nullptr;
      double* lpc_t = (half == 1 && lpc) ? lpc + (size_t)t * W : nullptr;
      const int nt = half == 0 ? t : t + 1, nh = 1 - half;   // the half step after this one
      const int32_t* next_colour = nt < T ? colours + (size_t)nt * W : nullptr;
      if (sharded) {
        shard(ns, world, rank, &lo, &cnt);
        if (cnt > 0 && logprob_impl(h, h->mc_q.as<double>() + (size_t)lo * p, cnt, nullptr, h->mc_newlp.as<double>(),
                                    nullptr, nullptr, st))
          return -1;
        CUDA_TRY(bgp::launch_accept_xchg_propose(h->peers, pos, lp, h->mc_q.as<double>(), h->mc_factors.as<double>(),
                                                 h->mc_newlp.as<double>(), lo, cnt, h->mc_movers.as<int32_t>(), W, p,
                                                 half, sp, t, acc, chain_t, lpc_t, next_colour, nh, nt, a, st));
      } else {
        if (logprob_impl(h, h->mc_q.as<double>(), ns, nullptr, h->mc_newlp.as<double>(), nullptr, nullptr, st))
          return -1;
        CUDA_TRY(bgp::launch_accept_propose(pos, lp, h->mc_q.as<double>(), h->mc_factors.as<double>(),
                                            h->mc_newlp.as<double>(), h->mc_movers.as<int32_t>(), W, p, half, sp, t,
                                            acc, chain_t, lpc_t, next_colour, nh, nt, a, st));
      }
    }
  }
  return 0;
}

static int mcmc_run_impl(bgp_handle_t h, double* pos_dev, double* lp_dev, int W, int T, double a, uint64_t seed,
                         double* chain_dev, double* lp_chain_dev, int32_t* accepted_dev, bool sharded, void* stream);

int bgp_mcmc_run(bgp_handle_t h, double* pos_dev, double* lp_dev, int W, int T, double a, uint64_t seed,
                 double* chain_dev, double* lp_chain_dev, int32_t* accepted_dev, void* stream) {
  CHECK_H(h);
  return mcmc_run_impl(h, pos_dev, lp_dev, W, T, a, seed, chain_dev, lp_chain_dev, accepted_dev, false, stream);
}

int bgp_mcmc_run_sharded(bgp_handle_t h, double* pos_dev, double* lp_dev, int W, int T, double a, uint64_t seed,
                         double* chain_dev, double* lp_chain_dev, int32_t* accepted_dev, void* stream) {
  CHECK_H(h);
  if (!h->have_peers) return fail("bgp_peer_connect has not been called");
  if (W > h->peers.cap) return fail("more walkers than the exchange block was exported for");
  if (stream == nullptr) return fail("the sharded run needs a non-default stream (it is a CUDA graph)");
  return mcmc_run_impl(h, pos_dev, lp_dev, W, T, a, seed, chain_dev, lp_chain_dev, accepted_dev, true, stream);
}

/* ---- peer exchange blocks (cudaIpc) ---- */
int bgp_peer_export(bgp_handle_t h, int max_walkers, void* ipc_handle_out) {
  CHECK_H(h);
  if (max_walkers < 2 || max_walkers > 8192 || !ipc_handle_out) return fail("bad peer-export arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->have_peers) return fail("peers are already connected (bgp_peer_close first)");
  const size_t bytes = sizeof(double) * 2 * (size_t)max_walkers + sizeof(unsigned long long) * 16;
  h->xchg.release();
  CUDA_TRY(h->xchg.ensure(bytes));
  CUDA_TRY(cudaMemset(h->xchg.p, 0, bytes));
  cudaIpcMemHandle_t ih;
  CUDA_TRY(cudaIpcGetMemHandle(&ih, h->xchg.p));
  static_assert(sizeof(cudaIpcMemHandle_t) == BGP_IPC_HANDLE_BYTES, "IPC handle size");
  std::memcpy(ipc_handle_out, &ih, sizeof(ih));
  h->peers.cap = max_walkers;
  return 0;
}

int bgp_peer_connect(bgp_handle_t h, const void* ipc_handles, int rank, int world) {
  CHECK_H(h);
  if (!ipc_handles || world < 1 || world > 8 || rank < 0 || rank >= world) return fail("bad peer-connect arguments");
  if (!h->xchg.p) return fail("bgp_peer_export has not been called");
  CUDA_TRY(cudaSetDevice(h->device));
  const int cap = h->peers.cap;
  std::memset(&h->peers, 0, sizeof(h->peers));
  h->peers.cap = cap; h->peers.rank = rank; h->peers.world = world;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { h->peers.block[r] = h->xchg.as<double>(); continue; }
    cudaIpcMemHandle_t ih;
    std::memcpy(&ih, static_cast<const char*>(ipc_handles) + (size_t)r * sizeof(ih), sizeof(ih));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int k = 0; k < r; ++k) if (k != rank && h->peers.block[k]) cudaIpcCloseMemHandle(h->peers.block[k]);
      return fail("cudaIpcOpenMemHandle (no peer access between the GPUs of this box?)", e);
    }
    h->peers.block[r] = static_cast<double*>(ptr);
  }
  h->have_peers = true;
  h->have_graph = false;
  return 0;
}

int bgp_peer_close(bgp_handle_t h) {
  CHECK_H(h);
  if (!h->have_peers) return 0;
  cudaSetDevice(h->device);
  if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
  h->have_graph = false;
  for (int r = 0; r < h->peers.world; ++r)
    if (r != h->peers.rank && h->peers.block[r]) cudaIpcCloseMemHandle(h->peers.block[r]);
  h->have_peers = false;
  return 0;
}

/* developer counters of the peer exchange: out[0] = nanoseconds this rank spent inside exchanges (stores, fence,
 * waiting for the slowest peer), out[1] = number of exchanges, since bgp_peer_export */
int bgp_peer_counters(bgp_handle_t h, unsigned long long* out) {
  CHECK_H(h);
  if (!out) return fail("null out");
  out[0] = out[1] = 0;
  if (!h->xchg.p) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpy(out, reinterpret_cast<const char*>(h->xchg.p) + sizeof(double) * 2 * (size_t)h->peers.cap +
                               sizeof(unsigned long long) * 10, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

/* 1 when a peer exchange of this handle ever timed out (a rank died or never launched): the chain of that run
 * is invalid */
int bgp_peer_status(bgp_handle_t h, int* timed_out) {
  CHECK_H(h);
  if (!timed_out) return fail("null out");
  *timed_out = 0;
  if (!h->xchg.p) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  unsigned long long w = 0;
  CUDA_TRY(cudaMemcpy(&w, reinterpret_cast<const char*>(h->xchg.p) + sizeof(double) * 2 * (size_t)h->peers.cap +
                              sizeof(unsigned long long) * 9, sizeof(w), cudaMemcpyDeviceToHost));
  *timed_out = w != 0;
  return 0;
}

static int mcmc_run_impl(bgp_handle_t h, double* pos_dev, double* lp_dev, int W, int T, double a, uint64_t seed,
                         double* chain_dev, double* lp_chain_dev, int32_t* accepted_dev, bool sharded, void* stream) {
  if (ready(h)) return -1;
  const int p = h->host_prog.n_theta;
  if (!pos_dev || !lp_dev || W < 2 || W > 8192 || T < 0) return fail("bad mcmc arguments");
  if (W < 2 * p) return fail("fewer walkers than twice the number of dimensions");
  if (h->have_priors && h->n_priors != p) return fail("prior count != theta count");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(h->mc_colour.ensure(sizeof(int32_t) * W * (size_t)(T > 0 ? T : 1)));   // one row per step
  CUDA_TRY(h->mc_movers.ensure(sizeof(int32_t) * W));
  CUDA_TRY(h->mc_q.ensure(sizeof(double) * W * p));
  CUDA_TRY(h->mc_factors.ensure(sizeof(double) * W));
  CUDA_TRY(h->mc_newlp.ensure(sizeof(double) * W));
  if (ensure_logprob_workspace(h)) return -1;   // must exist before capture starts
  *h->seed_pinned = seed;
  CUDA_TRY(cudaMemcpyAsync(h->mc_seed.p, h->seed_pinned, sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  if (st == nullptr) {  // the legacy default stream cannot be captured: run eagerly
    return mcmc_enqueue(h, pos_dev, lp_dev, W, T, a, chain_dev, lp_chain_dev, accepted_dev, sharded, st);
  }
  GraphKey key;
  std::memset(&key, 0, sizeof(key));   // compared with memcmp: padding bytes must be defined
  key.pos = pos_dev; key.lp = lp_dev; key.chain = chain_dev; key.lpc = lp_chain_dev; key.acc = accepted_dev;
  key.W = W; key.T = T; key.n = h->n; key.d = h->d; key.p = p; key.a = a;
  key.rank = sharded ? h->peers.rank : 0; key.world = sharded ? h->peers.world : 1;
  if (!(h->have_graph && h->key == key)) {
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
    h->have_graph = false;
    CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = mcmc_enqueue(h, pos_dev, lp_dev, W, T, a, chain_dev, lp_chain_dev, accepted_dev, sharded, st);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return -1; }
    if (e != cudaSuccess) return fail("cudaStreamEndCapture", e);
    e = cudaGraphInstantiate(&h->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail("cudaGraphInstantiate", e);
    h->key = key;
    h->have_graph = true;
  }
  CUDA_TRY(cudaGraphLaunch(h->graph, st));
  return 0;
}

}  // extern "C"
