"""BayesGPR: Gaussian-process regressor whose kernel hyper-parameters are inferred in a fully
Bayesian way -- the reference's public surface (bask/bayesgpr.py:18-718) over libbgp.

What runs where
  * device (libbgp, sm_100a): every Gram build, Cholesky, triangular solve, LML, log-prior,
    the whole ensemble MCMC, predictive moments, joint posterior draws;
  * host (this file): argument handling, y normalisation, the kernel *object* bookkeeping
    (scikit-learn kernel objects are used as parameter containers exactly like the reference
    uses skopt's subclasses of them), L-BFGS-B's iteration logic for the MAP start, the
    geometric median of the chain.

Deliberate differences from the reference, all additive or documented:
  * ``warp_inputs=True`` warps every point set per theta row on the device (Beta-CDF device
    function); the host keeps the reference's ``warp``/``unwarp``/``create_warpers`` surface;
  * the MCMC random stream is Philox on device, so chains agree with emcee in distribution,
    not draw by draw (BASELINE.json north_star);
  * ``sample``/``fit`` take an extra ``n_walkers`` alias and ``with_hyperparam(theta)``
    is a new context manager (the north star asks for it; the reference has no such method);
  * ``L_``, ``K_inv_`` and ``alpha_`` are materialised lazily from the device factor.
"""
import warnings
from contextlib import contextmanager

import numpy as np
import scipy.optimize
from sklearn.base import clone
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, WhiteKernel
from sklearn.utils import check_random_state

from . import _lib
from ._engine import Engine, find_zeroable_white, nvtx_range
from .priors import as_device_priors
from .utils import geometric_median, guess_priors, validate_zeroone

__all__ = ["BayesGPR"]

_INT32_MAX = np.iinfo(np.int32).max


def _walkers_independent(coords):
    """emcee 3.1.6 ``walkers_independent``: the centred, column-normalised ensemble must have a
    condition number <= 1e8, otherwise run_mcmc refuses the initial state."""
    if not np.all(np.isfinite(coords)):
        return False
    C = coords - np.mean(coords, axis=0)[None, :]
    C_colmax = np.amax(np.abs(C), axis=0)
    if np.any(C_colmax == 0):
        return False
    C = C / C_colmax
    C_colsum = np.sqrt(np.sum(C ** 2, axis=0))
    C = C / C_colsum
    return np.linalg.cond(C.astype(float)) <= 1e8


def p_eager():
    """BGP_EAGER_SAMPLE=1 makes sample() read its results back before returning (debugging aid)."""
    import os
    return os.environ.get("BGP_EAGER_SAMPLE", "0") == "1"


class BayesGPR:
    """Drop-in for ``bask.BayesGPR`` on one B200.  See the reference docstring
    (bask/bayesgpr.py:19-146) for the meaning of every parameter and attribute."""

    def __init__(self, kernel=None, alpha=1e-10, optimizer="fmin_l_bfgs_b", n_restarts_optimizer=0,
                 normalize_y=False, warp_inputs=False, copy_X_train=True, random_state=None,
                 noise="gaussian", device=None):
        self._kernel = None if kernel is None else kernel.clone_with_theta(kernel.theta)
        self.kernel = kernel
        self.alpha = alpha
        self._alpha = self.alpha
        self.optimizer = optimizer
        self.n_restarts_optimizer = n_restarts_optimizer
        self.normalize_y = normalize_y
        self.warp_inputs = bool(warp_inputs)
        self.copy_X_train = copy_X_train
        self.random_state = check_random_state(random_state)
        self.noise = noise
        self.noise_ = None
        self._pending = None         # device-side result of the last sample() not yet read back
        self.chain_ = None
        self.pos_ = None
        self.kernel_ = None
        self._device = device
        self._engine = None
        self._factor = None          # device factorisation at the current theta
        self._dense = {}             # lazily extracted / user-assigned L_, K_inv_, alpha_
        self._mc_buffers = None
        self._prior_key = None
        self.chain_generation_ = 0   # bumped whenever chain_ is replaced (cache keys of the diagnostics)
        self.timings_ = {}           # CUDA-event milliseconds of the last sample(): "mcmc", "factorize"

    # ------------------------------------------------------------------ deferred read-back
    # sample() on the device path returns as soon as the MCMC graph is enqueued.  chain_, pos_,
    # kernel_ (whose theta becomes the geometric median of the chain) and everything derived from
    # them are read back at their first use (`_materialize`): a caller that goes straight on to
    # evaluate_acquisitions enqueues the sweep behind the chain (theta rows gathered on the
    # device) and the read-back, the median and the point-estimate factorisation then overlap it.
    def _lazy(name):   # noqa: N805
        store = "_lz_" + name

        def get(self):
            if self._pending is not None:
                self._materialize()
            return getattr(self, store, None)

        def put(self, v):
            if getattr(self, "_pending", None) is not None:
                self._materialize()
            setattr(self, store, v)
        return property(get, put)

    chain_ = _lazy("chain_")
    pos_ = _lazy("pos_")
    kernel_ = _lazy("kernel_")
    timings_ = _lazy("timings_")          # CUDA-event milliseconds of the last sample()
    _acceptance = _lazy("_acceptance")    # per-walker acceptance fraction of the last sample()
    del _lazy

    def _n_kernel_theta(self):
        """len(kernel_.theta) without walking the kernel tree on every call (~0.1 ms in sklearn)."""
        k = self.kernel_
        c = getattr(self, "_nk_cache", None)
        if c is None or c[0] is not k:
            c = self._nk_cache = (k, len(k.theta))
        return c[1]

    def _chain_len(self):
        p = self._pending
        return p["kept"] * p["W"] if p is not None else len(self.chain_)

    def _chain_rows_dev(self, idx):
        """Rows ``idx`` of chain_ as a device tensor, without forcing the read-back."""
        e = self._eng()
        p = self._pending
        if p is None:
            return e.to_dev(self.chain_[idx])
        import torch
        idx = np.asarray(idx, dtype=np.int64)
        flat = (p["first"] + p["n_thin"] * (idx // p["W"])) * p["W"] + idx % p["W"]
        sel = e.to_dev(flat, dtype=torch.int64)
        with torch.cuda.stream(e.stream):
            return p["buf"]["chain"].reshape(-1, p["n_dim"]).index_select(0, sel)

    def _materialize(self):
        p, self._pending = self._pending, None
        if p is None:
            return
        e = self._eng()
        buf, ev = p["buf"], p["ev"]
        chain_steps, pos_out, acc = e.fetch_after(ev[1], buf["chain"], buf["pos"], buf["acc"])
        self._lz_timings_["mcmc_ms"] = ev[0].elapsed_time(ev[1])
        self._lz__acceptance = acc / max(p["n_samples"], 1)
        self._finish_sample(chain_steps, pos_out, p["n_burnin"], p["n_thin"], p["n_dim"], p["n_kernel"],
                            p["added_dims"], p["add"])

    # ------------------------------------------------------------------ engine plumbing
    def _eng(self):
        if self._engine is None:
            self._engine = Engine(self._device)
        return self._engine

    def _upload_model(self, structure_changed):
        e = self._eng()
        if structure_changed:
            e.set_kernel(self.kernel_, n_warp=self._X_train.shape[1] if self.warp_inputs else 0)
            self._prior_key = None
        alpha = self.alpha
        e.set_data(self._X_train, self.y_train_, alpha)

    def _theta_for_device(self, theta=None):
        """Device theta row for kernel hyper-parameters ``theta`` (default: the point estimate):
        with input warping the current warp parameters are appended (log a, then log b; zeros,
        i.e. the identity warp, before the first ``create_warpers``)."""
        with np.errstate(divide="ignore"):
            th = np.array(self.kernel_.theta if theta is None else theta, dtype=np.float64)
        if self.warp_inputs:
            d = self._X_train.shape[1]
            a = getattr(self, "warp_alphas_", np.zeros(d))
            b = getattr(self, "warp_betas_", np.zeros(d))
            th = np.concatenate([th, a, b], axis=-1) if th.ndim == 1 else \
                np.concatenate([th, np.tile(a, (len(th), 1)), np.tile(b, (len(th), 1))], axis=1)
        return th

    def _refactor(self, lazy=False):
        """Factorises at the current theta.  ``lazy`` only enqueues the work: the positive-
        definiteness check (and its LinAlgError) then happens at the first use of the factor."""
        e = self._eng()
        self._factor_pending = e.factorize(self._theta_for_device()[None, :])
        self._factor_value = None
        self._dense = {}
        self._lml_value = None
        if not lazy:
            self._factor  # noqa: B018  (materialises: synchronises and checks)

    @property
    def _factor(self):
        if self._pending is not None:
            self._materialize()
        f = getattr(self, "_factor_pending", None)
        if f is not None:
            self._factor_pending = None
            info = int(self._eng().to_host(f.info)[0])
            if info != 0:
                raise np.linalg.LinAlgError(
                    "The kernel, %s, is not returning a positive definite matrix. Try gradually "
                    "increasing the 'alpha' parameter of your GaussianProcessRegressor estimator."
                    % self.kernel_, f"{info}-th leading minor of the array is not positive definite")
            self._factor_value = f
        return getattr(self, "_factor_value", None)

    @_factor.setter
    def _factor(self, f):
        if getattr(self, "_pending", None) is not None:
            self._materialize()
        self._factor_pending = None
        self._factor_value = f

    @property
    def log_marginal_likelihood_value_(self):
        """LML at the current theta (sklearn attribute); read back from the device on first use."""
        if getattr(self, "_lml_value", None) is None and getattr(self, "_lml_from_factor", False):
            f = self._factor
            if f is not None:
                self._lml_value = float(self._eng().to_host(f.lml)[0])
        return getattr(self, "_lml_value", None)

    @log_marginal_likelihood_value_.setter
    def log_marginal_likelihood_value_(self, v):
        self._lml_from_factor = False
        self._lml_value = v

    def _dense_attr(self, name, what):
        if name in self._dense:
            return self._dense[name]
        if self._factor is None:
            raise AttributeError(f"{name} is only available after fit")
        e = self._eng()
        self._dense[name] = e.to_host(e.extract(self._factor, 0, what))
        return self._dense[name]

    L_ = property(lambda s: s._dense_attr("L_", _lib.EXTRACT_L),
                  lambda s, v: s._dense.__setitem__("L_", v))
    K_inv_ = property(lambda s: s._dense_attr("K_inv_", _lib.EXTRACT_KINV),
                      lambda s, v: s._dense.__setitem__("K_inv_", v))
    alpha_ = property(lambda s: s._dense_attr("alpha_", _lib.EXTRACT_ALPHA),
                      lambda s, v: s._dense.__setitem__("alpha_", v))

    # ------------------------------------------------------------------ reference surface
    @property
    def theta(self):
        """Current point estimate of the hyper-parameters, log space (bask/bayesgpr.py:182-198)."""
        if self.kernel_ is not None:
            with np.errstate(divide="ignore"):
                return np.copy(self.kernel_.theta)
        return None

    @theta.setter
    def theta(self, theta):
        """Sets the hyper-parameters and re-factorises on device (bask/bayesgpr.py:200-217)."""
        self.kernel_.theta = np.asarray(theta, dtype=np.float64)
        self._refactor()

    @property
    def X_train_(self):
        """Training inputs; the warped instances when ``warp_inputs=True`` and warpers exist
        (bask/bayesgpr.py:219-247).  The device always holds the original inputs and warps them
        per theta itself."""
        X = getattr(self, "_X_train", None)
        if X is not None and self.warp_inputs and hasattr(self, "warpers_"):
            if getattr(self, "_X_train_warped", None) is None:
                self._X_train_warped = self.warp(X)
            return self._X_train_warped
        return X

    @X_train_.setter
    def X_train_(self, X_train):
        self._X_train = np.copy(X_train) if self.copy_X_train else X_train
        self._X_train_warped = None

    def warp(self, X):
        """Beta-CDF warp of X with the current warpers (bask/bayesgpr.py:249-264)."""
        if self.warp_inputs and hasattr(self, "warpers_"):
            X = np.asarray(X, dtype=np.float64)
            X_warped = np.empty_like(X)
            for col, warper in enumerate(self.warpers_):
                X_warped[:, col] = warper(X[:, col])
            X = X_warped
        return X

    def unwarp(self, X):
        """Inverse of ``warp`` (bask/bayesgpr.py:266-283)."""
        if self.warp_inputs and hasattr(self, "warpers_"):
            X = np.asarray(X, dtype=np.float64)
            X_orig = np.empty_like(X)
            for col, unwarper in enumerate(self.unwarpers_):
                X_orig[:, col] = unwarper(X[:, col])
            X = X_orig
        return X

    def rewarp(self):
        """Re-applies the warpers to the stored training inputs (bask/bayesgpr.py:285-297)."""
        self._X_train_warped = None

    def create_warpers(self, alphas, betas):
        """Beta CDFs / inverse CDFs from log-space parameters (bask/bayesgpr.py:298-316)."""
        if self.warp_inputs:
            import scipy.stats as st
            self.warpers_, self.unwarpers_ = [], []
            self.warp_alphas_ = np.copy(alphas)
            self.warp_betas_ = np.copy(betas)
            for a_log, b_log in zip(alphas, betas, strict=True):
                dist = st.beta(a=np.exp(a_log), b=np.exp(b_log))
                self.warpers_.append(dist.cdf)
                self.unwarpers_.append(dist.ppf)
            self._X_train_warped = None

    @contextmanager
    def noise_set_to_zero(self):
        """Predictions inside the context exclude the observation noise: the White kernel found
        by skopt's search is swapped for WhiteKernel(0) while alpha_/L_/K_inv_ stay untouched
        (bask/bayesgpr.py:318-336)."""
        current_theta = self.theta
        try:
            _white, name = find_zeroable_white(self.kernel_)
            if name is None:
                name = "_"
            self.kernel_.set_params(**{name: WhiteKernel(noise_level=0.0)})
            yield self
        finally:
            self.kernel_.theta = current_theta

    @contextmanager
    def with_hyperparam(self, theta):
        """Evaluate at ``theta`` (log space), then restore the previous point estimate and its
        device factor.  New, additive API (BASELINE.json north_star); parity is defined against
        ``gp.theta = theta`` of the reference."""
        backup_theta, backup_factor, backup_dense = self.theta, self._factor, self._dense
        try:
            self.theta = theta
            yield self
        finally:
            self.kernel_.theta = backup_theta
            self._factor, self._dense = backup_factor, backup_dense

    def _apply_noise_vector(self, n_instances, noise_vector):
        """bask/bayesgpr.py:338-349"""
        if noise_vector is not None:
            if not np.iterable(self.alpha):
                alpha = np.ones(n_instances) * self.alpha
            elif not np.iterable(self._alpha):
                alpha = np.ones(n_instances) * self._alpha
            alpha[: len(noise_vector)] += noise_vector
            self.alpha = alpha

    # ------------------------------------------------------------------ log-probabilities
    def log_marginal_likelihood(self, theta=None, eval_gradient=False, clone_kernel=True):
        """LML of theta on device (sklearn:_gpr.py:541-656); with ``eval_gradient`` also its analytic
        gradient 1/2 tr((alpha alpha^T - K^-1) dK/dtheta) (``bgp_lml_gradient``), which drives the
        L-BFGS-B MAP start of ``fit``."""
        if theta is None:
            if eval_gradient:
                raise ValueError("Gradient can only be evaluated for theta!=None")
            return self.log_marginal_likelihood_value_
        theta = np.asarray(theta, dtype=np.float64)
        if not clone_kernel:
            self.kernel_.theta = theta
        e = self._eng()
        if not eval_gradient:
            _, lml, _ = e.logprob(self._theta_for_device(theta)[None, :])
            return float(lml[0])
        lml, grad, info = e.lml_gradient(self._theta_for_device(theta))
        if info != 0 or not np.isfinite(lml):
            return -np.inf, np.zeros_like(theta)
        return lml, grad

    def _prior_table(self, priors, warp_priors, n_kernel):
        """(device prior table, host part) over a full theta row: the kernel's priors followed,
        with input warping, by the warp priors -- warp_priors[0] on every log a_k, warp_priors[1]
        on every log b_k, default Normal(0, 0.3) (bask/bayesgpr.py:351-372, 462-466)."""
        table, host_fn = as_device_priors(priors, n_kernel)
        if not self.warp_inputs:
            return table, host_fn
        d_in = self._X_train.shape[1]
        if warp_priors is None:
            from .priors import NormalPrior
            warp_priors = (NormalPrior(0.0, 0.3), NormalPrior(0.0, 0.3))
        if callable(warp_priors) and not isinstance(warp_priors, (list, tuple)):
            wtable = [(_lib.PRIOR_NONE, ())] * (2 * d_in)
            wfn = lambda w, f=warp_priors: float(sum(f(w[k], w[d_in + k]) for k in range(d_in)))  # noqa: E731
        else:
            wtable, wfn = as_device_priors([warp_priors[0]] * d_in + [warp_priors[1]] * d_in, 2 * d_in)
        if host_fn is not None or wfn is not None:
            kf, wf = host_fn, wfn
            host_fn = lambda th: (kf(th[:n_kernel]) if kf else 0.0) + (wf(th[n_kernel:]) if wf else 0.0)  # noqa: E731
        return table + wtable, host_fn

    def _log_prob_fn(self, x, priors, warp_priors=None):
        """Log posterior of one theta (or a (B, p) batch) -- bask/bayesgpr.py:351-379.  With input
        warping the rows are ``kernel theta ++ log a ++ log b``."""
        x = np.asarray(x, dtype=np.float64)
        single = x.ndim == 1
        X = np.atleast_2d(x)
        n_kernel = X.shape[1] - (2 * self._X_train.shape[1] if self.warp_inputs else 0)
        table, host_fn = self._prior_table(priors, warp_priors, n_kernel)
        e = self._eng()
        e.set_priors(table)
        self._prior_key = None
        extra = None if host_fn is None else np.array([host_fn(t) for t in X])
        lp, _, _ = e.logprob(X, extra)
        return float(lp[0]) if single else lp

    # ------------------------------------------------------------------ MCMC
    def sample(self, X=None, y=None, noise_vector=None, n_threads=1, n_desired_samples=100, n_burnin=0,
               n_thin=1, n_walkers_per_thread=100, progress=False, priors=None, warp_priors=None,
               position=None, add=False, n_walkers=None, process_group=None, **kwargs):
        """Ensemble MCMC over the hyper-posterior, entirely on device (bask/bayesgpr.py:381-548).
        ``**kwargs`` are the emcee.EnsembleSampler keywords of the reference; ``a`` (stretch
        scale) is honoured, ``vectorize``/``threads`` are meaningless here and ignored."""
        if X is None and self.X_train_ is None or self.kernel_ is None:
            raise ValueError(
                "It looks like you are trying to sample from the GP posterior without data. "
                "Pass X and y, or ensure that you call fit before sample.")
        if priors is None:
            priors = guess_priors(self.kernel_)
        data_changed = False
        if X is not None:
            y = np.asarray(y, dtype=np.float64)
            if self.normalize_y:
                self._y_train_mean = np.mean(y, axis=0)
                self._y_train_std = np.std(y, axis=0)
            else:
                self._y_train_mean = np.zeros(1)
                self._y_train_std = 1
            self.y_train_std_ = self._y_train_std
            self.y_train_mean_ = self._y_train_mean
            y = (y - self.y_train_mean_) / self.y_train_std_
            if noise_vector is not None:
                noise_vector = np.array(noise_vector) / np.power(self.y_train_std_, 2)
            self.X_train_ = np.array(X, dtype=np.float64)
            self.y_train_ = np.copy(y) if self.copy_X_train else y
            data_changed = True
        self._apply_noise_vector(len(self.y_train_), noise_vector)
        if data_changed or noise_vector is not None:
            self._upload_model(structure_changed=False)

        n_dim = n_kernel = self._n_kernel_theta()
        if n_walkers is None:
            n_walkers = n_threads * n_walkers_per_thread
        n_samples = int(np.ceil(n_desired_samples / n_walkers) + n_burnin)
        pos = None
        if position is not None:
            pos = position
        elif self.pos_ is not None:
            pos = self.pos_
        added_dims = 0
        if self.warp_inputs:
            added_dims = self._X_train.shape[1] * 2
            n_dim += added_dims
        if pos is None:
            theta = self.theta
            theta[np.isinf(theta)] = np.log(self.noise_)
            if self.warp_inputs:
                theta = np.concatenate([theta, np.zeros(added_dims)])
            pos = [theta + 1e-2 * self.random_state.randn(n_dim) for _ in range(n_walkers)]
        pos = np.array(pos, dtype=np.float64)
        if pos.shape != (n_walkers, n_dim):
            raise ValueError("incompatible input dimensions")
        unknown = set(kwargs) - {"a", "vectorize", "threads", "pool", "backend", "blobs_dtype"}
        if unknown:
            raise NotImplementedError(f"EnsembleSampler options {sorted(unknown)} are not supported "
                                      "by the device sampler (only the default StretchMove)")
        a = float(kwargs.get("a", 2.0) or 2.0)
        seed = int(self.random_state.randint(0, _INT32_MAX))
        if n_walkers < 2 * n_dim:
            raise RuntimeError("It is unadvisable to use a red-blue move with fewer walkers than twice "
                               "the number of dimensions.")

        def check_initial_state():
            if not np.all(np.isfinite(pos)):
                raise ValueError("At least one parameter value was infinite or NaN")
            if not _walkers_independent(pos):
                raise ValueError("Initial state has a large condition number. Make sure that your walkers "
                                 "are linearly independent for the best performance")

        e = self._eng()
        table, host_fn = self._prior_table(priors, warp_priors, n_kernel)
        e.set_priors(table)
        if host_fn is None and process_group is not None and \
                __import__("torch").distributed.get_world_size(process_group) > 1:
            # walkers sharded over the ranks of the group (one all-gather of W/2 log-probs per half
            # step); every rank ends up with the identical chain
            from .distributed import sharded_mcmc
            check_initial_state()
            chain_steps, pos_out, accepted = sharded_mcmc(e, pos, n_samples, seed, a, process_group)
            self._acceptance = accepted / max(n_samples, 1)
        elif host_fn is None:
            import torch
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            with torch.cuda.nvtx.range("bgp.sample.mcmc"):
                ev[0].record(e.stream)
                buf = e.mcmc(pos, n_samples, seed, a=a, buffers=self._mc_buffers)
                ev[1].record(e.stream)
            self._mc_buffers = buf
            # emcee's checks of the initial state (a 128 x p SVD among them) run while the device is
            # already sampling; a rejected state raises exactly as before and the run is discarded
            check_initial_state()
            self._lz_timings_["mcmc_logprob_evals"] = int(n_walkers * (1 + n_samples))
            first = n_burnin + n_thin - 1
            self._pending = dict(buf=buf, ev=ev, n_samples=n_samples, n_burnin=n_burnin, n_thin=n_thin,
                                 n_dim=n_dim, n_kernel=n_kernel, added_dims=added_dims, add=add, W=n_walkers,
                                 first=first, kept=len(range(first, n_samples, n_thin)))
            self.chain_generation_ += 1
            if add or self.warp_inputs or p_eager():
                self._materialize()
            return
        else:
            check_initial_state()
            chain_steps, pos_out = self._host_stepped_mcmc(pos, n_samples, seed, a, host_fn)
        self.chain_generation_ += 1
        self._finish_sample(chain_steps, pos_out, n_burnin, n_thin, n_dim, n_kernel, added_dims, add)

    def _finish_sample(self, chain_steps, pos_out, n_burnin, n_thin, n_dim, n_kernel, added_dims, add):
        """Host tail of sample(): chain bookkeeping, geometric median as the point estimate
        (bask/bayesgpr.py:531-548)."""
        if not np.all(np.isfinite(chain_steps)):
            raise ValueError("At least one parameter value was infinite or NaN")
        chain = chain_steps[n_burnin + n_thin - 1:: n_thin].reshape(-1, n_dim)
        if add and self.chain_ is not None:
            self.chain_ = np.concatenate([self.chain_, chain])
        else:
            self.chain_ = chain
        # point estimate: the factorisation at the geometric median is only enqueued here; its
        # LinAlgError check and the LML read-back happen at the first use (no host round trip)
        median = geometric_median(self.chain_)
        if self.warp_inputs:
            warp_params = median[n_kernel:]
            self.create_warpers(warp_params[: added_dims // 2], warp_params[added_dims // 2:])
            self.rewarp()
        self.kernel_.theta = median[:n_kernel]
        self._refactor(lazy=True)
        self._lml_from_factor = True
        self.pos_ = pos_out

    def _host_stepped_mcmc(self, pos, n_steps, seed, a, host_fn):
        """Same move, but the log-prior of untyped Python callables is evaluated on the host
        between propose and accept (one D2H/H2D of W/2 x p doubles per half step).  The GP
        numerics stay on device."""
        import ctypes as C

        import torch
        e = self._eng()
        W, p = pos.shape
        lib, h, st = e.lib, e.h, e._st
        P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        d_pos = e.to_dev(pos)
        extra = e.to_dev(np.array([host_fn(t) for t in pos]))
        d_lp, _, _ = e.logprob_dev(d_pos, extra)
        colour = e.empty(W, dtype=torch.int32)
        movers = e.empty(W, dtype=torch.int32)
        q, fac = e.empty(W, p), e.empty(W)
        acc = e.zeros(W, dtype=torch.int32)
        chain = e.empty(n_steps, W, p)
        lpc = e.empty(n_steps, W)
        sd = C.c_uint64(seed)
        for t in range(n_steps):
            _lib.check(lib.bgp_mcmc_split(h, W, sd, t, P(colour), st), "bgp_mcmc_split")
            for half in (0, 1):
                ns = (W + 1) // 2 if half == 0 else W // 2
                _lib.check(lib.bgp_mcmc_propose(h, P(d_pos), P(colour), W, half, a, sd, t, P(q), P(fac),
                                                P(movers), st), "bgp_mcmc_propose")
                qh = e.to_host(q[:ns])
                if not np.all(np.isfinite(qh)):
                    raise ValueError("At least one parameter value was infinite or NaN")
                ex = e.to_dev(np.array([host_fn(v) for v in qh]))
                nlp, _, _ = e.logprob_dev(q[:ns], ex)
                _lib.check(lib.bgp_mcmc_accept(h, P(d_pos), P(d_lp), P(q), P(fac), P(nlp), P(movers), W, half,
                                               sd, t, P(acc), P(chain[t]) if half == 1 else None,
                                               P(lpc[t]) if half == 1 else None, st), "bgp_mcmc_accept")
                e.launches += 3
        e.sync()
        return chain.cpu().numpy(), d_pos.cpu().numpy()

    # ------------------------------------------------------------------ fit
    def fit(self, X, y, noise_vector=None, n_threads=1, n_desired_samples=100, n_burnin=10,
            n_walkers_per_thread=100, progress=True, priors=None, warp_priors=None, position=None,
            process_group=None, **kwargs):
        """MAP start (L-BFGS-B over device LML evaluations) followed by ``sample``
        (bask/bayesgpr.py:550-620 -> skopt/sklearn fit, sklearn:_gpr.py:233-368)."""
        self.kernel = self._kernel
        if self.normalize_y and noise_vector is not None:
            y_std = np.std(y, axis=0)
            noise_vector = np.array(noise_vector) / np.power(y_std, 2)
        self._apply_noise_vector(len(y), noise_vector)
        with nvtx_range("bgp.fit.map"):
            self._fit_map(X, y)
        self.sample(n_threads=n_threads, n_desired_samples=n_desired_samples, n_burnin=n_burnin,
                    n_walkers_per_thread=n_walkers_per_thread, progress=progress, priors=priors,
                    warp_priors=warp_priors, position=position, add=False, process_group=process_group,
                    **kwargs)
        return self

    def _fit_map(self, X, y):
        kernel = self.kernel
        if kernel is None:
            kernel = ConstantKernel(1.0, constant_value_bounds="fixed") * RBF(1.0, length_scale_bounds="fixed")
        if self.noise == "gaussian":
            kernel = kernel + WhiteKernel()
        elif self.noise:
            kernel = kernel + WhiteKernel(noise_level=self.noise, noise_level_bounds="fixed")
        self.kernel = kernel
        self.kernel_ = clone(kernel)
        self._rng = check_random_state(self.random_state)
        X = np.array(X, dtype=np.float64)
        y = np.array(y, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError(f"Expected 2D array, got {X.ndim}D array instead")
        if y.ndim != 1 or len(y) != len(X):
            raise ValueError("y must be 1-D with one target per row of X (multi-output GPs are outside "
                             "the hot path)")
        if not (np.all(np.isfinite(X)) and np.all(np.isfinite(y))):
            raise ValueError("Input contains NaN or infinity")
        if self.normalize_y:
            self._y_train_mean = np.mean(y, axis=0)
            std = np.std(y, axis=0)
            self._y_train_std = 1.0 if std < 10 * np.finfo(np.float64).eps else std
            y = (y - self._y_train_mean) / self._y_train_std
        else:
            self._y_train_mean = np.zeros(shape=1)
            self._y_train_std = np.ones(shape=1)
        if np.iterable(self.alpha) and self.alpha.shape[0] != y.shape[0]:
            if self.alpha.shape[0] == 1:
                self.alpha = self.alpha[0]
            else:
                raise ValueError("alpha must be a scalar or an array with same number of entries as y. "
                                 f"({self.alpha.shape[0]} != {y.shape[0]})")
        self.X_train_ = X
        self.y_train_ = np.copy(y) if self.copy_X_train else y
        self.pos_ = self.pos_  # untouched: the reference keeps a warm start across fits
        self._upload_model(structure_changed=True)

        if self.optimizer is not None and self.kernel_.n_dims > 0:
            def obj_func(theta, eval_gradient=True):
                if eval_gradient:
                    lml, grad = self.log_marginal_likelihood(theta, eval_gradient=True, clone_kernel=False)
                    return -lml, -grad
                return -self.log_marginal_likelihood(theta, clone_kernel=False)

            optima = [self._constrained_optimization(obj_func, self.kernel_.theta, self.kernel_.bounds)]
            if self.n_restarts_optimizer > 0:
                if not np.isfinite(self.kernel_.bounds).all():
                    raise ValueError("Multiple optimizer restarts (n_restarts_optimizer>0) requires that "
                                     "all bounds are finite.")
                bounds = self.kernel_.bounds
                for _ in range(self.n_restarts_optimizer):
                    theta_initial = self._rng.uniform(bounds[:, 0], bounds[:, 1])
                    optima.append(self._constrained_optimization(obj_func, theta_initial, bounds))
            lml_values = [o[1] for o in optima]
            self.kernel_.theta = optima[int(np.argmin(lml_values))][0]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.kernel_._check_bounds_params()
            map_lml = -np.min(lml_values)
        else:
            map_lml = self.log_marginal_likelihood(self.kernel_.theta, clone_kernel=False)
        self._refactor()
        self.log_marginal_likelihood_value_ = map_lml      # (the refactorisation clears the cached value)
        self.noise_ = None
        if self.noise:
            if isinstance(self.kernel_, WhiteKernel):
                self.kernel_.set_params(noise_level=0.0)
            else:
                white, name = find_zeroable_white(self.kernel_)
                if white is not None:
                    self.noise_ = white.noise_level
                    n_free = self.kernel_.n_dims
                    self.kernel_.set_params(**{name: WhiteKernel(noise_level=0.0)})
                    if self.kernel_.n_dims != n_free:
                        # noise=<float>: skopt's fixed White leaf has just become a free hyper-parameter
                        # (bask samples it, starting from log(noise_)): the device program gets the new slot
                        self._eng().set_kernel(self.kernel_,
                                               n_warp=self._X_train.shape[1] if self.warp_inputs else 0)
                        self._prior_key = None
        self.y_train_std_ = self._y_train_std
        self.y_train_mean_ = self._y_train_mean

    def _constrained_optimization(self, obj_func, initial_theta, bounds):
        if self.optimizer == "fmin_l_bfgs_b":
            res = scipy.optimize.minimize(obj_func, initial_theta, method="L-BFGS-B", jac=True, bounds=bounds)
            return res.x, res.fun
        if callable(self.optimizer):
            return self.optimizer(obj_func, initial_theta, bounds=bounds)
        raise ValueError(f"Unknown optimizer {self.optimizer}.")

    # ------------------------------------------------------------------ prediction
    def _moments_dev(self, X, factor=None, thetas=None, want_v=False, zextra=None):
        e = self._eng()
        X = np.asarray(X, dtype=np.float64)
        if X.ndim != 2 or X.shape[1] != self.X_train_.shape[1]:
            raise ValueError(f"X has {X.shape[-1] if X.ndim else 0} features, but BayesGPR is expecting "
                             f"{self.X_train_.shape[1]} features as input.")
        Xd = e.to_dev(X)
        f = self._factor if factor is None else factor
        th = e.to_dev(self._theta_for_device()[None, :]) if thetas is None else thetas
        y_mean = float(np.atleast_1d(self.y_train_mean_)[0])
        y_std = float(np.atleast_1d(self.y_train_std_)[0])
        return e.predict(f, Xd, thetas_dev=th, noise_off=False, y_mean=y_mean, y_std=y_std, zextra=zextra,
                         want_v=want_v) + (Xd, th, y_std)

    def predict(self, X, return_std=False, return_cov=False, return_mean_grad=False, return_std_grad=False):
        """Posterior mean (and std or covariance) at X with the CURRENT kernel_ -- i.e. including
        the noise level unless inside ``noise_set_to_zero`` (bask/bayesgpr.py:622-635 -> skopt
        predict)."""
        if return_std and return_cov:
            raise RuntimeError("Not returning standard deviation of predictions when returning full covariance.")
        if return_mean_grad or return_std_grad:
            raise NotImplementedError("gradient outputs are only used by skopt's own acquisition "
                                      "optimisers and are outside the hot path")
        X = np.asarray(X, dtype=np.float64)
        if self.warp_inputs:
            validate_zeroone(X)
        if self.X_train_ is None or self._factor is None:   # GP prior (skopt predict, unfitted branch)
            k = self.kernel if self.kernel is not None else ConstantKernel(1.0) * RBF(1.0)
            y_mean = np.zeros(X.shape[0])
            if return_cov:
                return y_mean, k(X)
            if return_std:
                return y_mean, np.sqrt(k.diag(X))
            return y_mean
        e = self._eng()
        mu, sd, _, v, Xd, th, y_std = self._moments_dev(X, want_v=return_cov)
        if return_cov:
            m = X.shape[0]
            cov = e.empty(m, m)
            _lib.check(e.lib.bgp_posterior_cov(e.h, th.data_ptr(), v.data_ptr(), Xd.data_ptr(), m, v.shape[2], 0,
                                               y_std, cov.data_ptr(), m, e._st), "bgp_posterior_cov")
            e.launches += 1
            return e.to_host(mu)[0], e.to_host(cov)
        if return_std:
            e.sync()
            return mu.cpu().numpy()[0], sd.cpu().numpy()[0]
        return e.to_host(mu)[0]

    # ------------------------------------------------------------------ joint draws
    # joint draws over more points than this factor the posterior covariance with the chip-wide blocked
    # Cholesky (bgp_dense_cholesky_inplace) instead of the one-cluster kernel
    _JOINT_DRAW_BIG_M = 1024

    def _joint_draws_dev(self, Xd, thetas_dev, factor, eps, noise):
        """eps: (S, m, ns) standard normals on device -> (S, m, ns) draws."""
        import torch
        e = self._eng()
        S, m, ns = eps.shape
        y_mean = float(np.atleast_1d(self.y_train_mean_)[0])
        y_std = float(np.atleast_1d(self.y_train_std_)[0])
        mu, _sd, _, v = e.predict(factor, Xd, thetas_dev=thetas_dev, noise_off=not noise, y_mean=y_mean,
                                  y_std=y_std, want_v=True)
        out = e.empty(S, m, ns)
        cov = e.empty(m, m)
        big = m > self._JOINT_DRAW_BIG_M      # one large matrix: whole-chip blocked factorisation, in place
        slab = None if big else e.empty(int(e.lib.bgp_dense_slab_doubles(m)))
        info = e.empty(1, dtype=torch.int32)
        for s in range(S):
            def build_cov():
                _lib.check(e.lib.bgp_posterior_cov(e.h, thetas_dev[s].data_ptr(), v[s].data_ptr(), Xd.data_ptr(), m,
                                                   v.shape[2], 0 if noise else 1, y_std, cov.data_ptr(), m, e._st),
                           "bgp_posterior_cov")
                e.launches += 1
            build_cov()
            with torch.cuda.stream(e.stream):   # ordered with the kernel that wrote cov
                scale = float(torch.diagonal(cov).abs().max().item()) or 1.0
            for attempt, jit in enumerate((1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4)):
                if big:
                    if attempt > 0:
                        build_cov()             # the failed in-place attempt overwrote the lower triangle
                    _lib.check(e.lib.bgp_dense_cholesky_inplace(e.h, cov.data_ptr(), m, m, jit * scale,
                                                                info.data_ptr(), e._st), "bgp_dense_cholesky_inplace")
                    e.launches += 2 * ((m + 255) // 256) + 1
                else:
                    _lib.check(e.lib.bgp_dense_cholesky(e.h, cov.data_ptr(), m, m, jit * scale, slab.data_ptr(),
                                                        info.data_ptr(), e._st), "bgp_dense_cholesky")
                    e.launches += 1
                if int(e.to_host(info)[0]) == 0:
                    break
            else:
                raise np.linalg.LinAlgError("posterior covariance is not positive definite even with jitter")
            if big:
                _lib.check(e.lib.bgp_dense_trmm(e.h, cov.data_ptr(), m, m, eps[s].data_ptr(), ns, mu[s].data_ptr(),
                                                out[s].data_ptr(), e._st), "bgp_dense_trmm")
            else:
                _lib.check(e.lib.bgp_slab_trmm(e.h, slab.data_ptr(), m, eps[s].data_ptr(), ns, mu[s].data_ptr(),
                                               out[s].data_ptr(), e._st), "bgp_slab_trmm")
            e.launches += 2
        return out, v

    def sample_y(self, X, sample_mean=False, noise=False, n_samples=1, random_state=0):
        """Joint posterior function draws at X, shape (n_points, n_samples)
        (bask/bayesgpr.py:637-718).  The host RandomState is consumed exactly like the reference
        consumes it (choice of chain indices, then n_samples x n_points standard normals per
        draw); the draws themselves use a Cholesky factor of the posterior covariance instead
        of numpy's SVD, so they agree in distribution."""
        rng = check_random_state(random_state)
        e = self._eng()
        X = np.asarray(X, dtype=np.float64)
        m = X.shape[0]
        Xd = e.to_dev(X)
        if sample_mean:
            eps = rng.standard_normal(size=(n_samples, m)).T
            th = e.to_dev(self._theta_for_device()[None, :])
            out, _ = self._joint_draws_dev(Xd, th, self._factor, e.to_dev(eps[None]), noise)
            return e.to_host(out)[0]
        ind = rng.choice(len(self.chain_), size=n_samples, replace=True)
        eps = np.stack([rng.standard_normal(size=(1, m)).T for _ in ind])       # (S, m, 1)
        th = e.to_dev(self.chain_[ind])
        f = e.factorize(th)
        if np.any(e.to_host(f.info) != 0):
            raise np.linalg.LinAlgError("The kernel is not returning a positive definite matrix for a "
                                        "sampled hyper-parameter vector.")
        out, _ = self._joint_draws_dev(Xd, th, f, e.to_dev(eps), noise)
        return e.to_host(out)[:, :, 0].T
