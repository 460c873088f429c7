"""TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of the reference's fully Bayesian GP
numerics (not part of the product; the product never imports this).

What it restates, and from where:

* kernel evaluation -- scikit-learn 1.7.2 (pinned uv.lock:2411-2412; 1.9.0 installed here)
  ``sklearn/gaussian_process/kernels.py``: Sum/Product ``__call__``/``diag`` (:838-873,
  :936-973), Exponentiation (:1060-1110), ConstantKernel (:1244-1296), WhiteKernel
  (:1374-1419: sigma^2*I only for K(X,X), zeros for cross kernels), RBF (:1530-1587),
  Matern nu in {1/2, 3/2, 5/2, inf} (:1685-1786), theta ordering k1.theta ++ k2.theta in
  log space (:738-766).
* log marginal likelihood -- ``sklearn/gaussian_process/_gpr.py:541-656`` (Alg. 2.1 GPML).
* theta setter / explicit K^-1 -- bask/bayesgpr.py:200-217.
* predictive mean/std/cov -- skopt 0.10.2 ``GaussianProcessRegressor.predict`` (off-tree;
  restated in oracle/skopt_port.py) as called from bask/bayesgpr.py:622-635; the noise-free
  variant is bask/bayesgpr.py:318-336 (White kernel set to 0, alpha_/L_/K_inv_ NOT recomputed).
* joint posterior draws -- ``sklearn/gaussian_process/_gpr.py:502-539`` as called from
  bask/bayesgpr.py:637-718.
* priors -- bask/utils.py:68-124 (half-normal on sqrt of a variance, round-flat on a length
  scale, both with the log-space Jacobian), bask/priors.py:7-57.
* log posterior of one theta -- bask/bayesgpr.py:351-379.
* geometric median -- bask/utils.py:21-65.

Pinning: tests/test_oracle_golden.py checks every function here against vectors produced by
the UNMODIFIED reference (+ sklearn) in the build container (tests/golden/make_golden.py,
committed with its outputs), and tests/test_oracle_ref_golden.py re-runs the reference's
own golden tests.

Kernel description ("spec"): a nested tuple tree
    ("sum", a, b) | ("product", a, b) | ("exp", a, exponent)
    ("const", value, fixed) | ("white", level, fixed)
    ("rbf", length_scale, fixed) | ("matern", length_scale, nu, fixed)
``length_scale`` is a float (isotropic) or a 1-D array (ARD).  ``fixed`` leaves take no
slot in theta.  ``spec_from_sklearn`` converts a scikit-learn kernel object.
"""
import math

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.spatial.distance import cdist, pdist, squareform
from scipy.stats import halfnorm, invgamma, norm

LOG_2PI = math.log(2.0 * math.pi)


# ------------------------------------------------------------------ kernel specs
def spec_from_sklearn(k):
    """Duck-typed conversion of a scikit-learn (or skopt) kernel object into a spec."""
    name = type(k).__name__
    if name == "Sum":
        return ("sum", spec_from_sklearn(k.k1), spec_from_sklearn(k.k2))
    if name == "Product":
        return ("product", spec_from_sklearn(k.k1), spec_from_sklearn(k.k2))
    if name == "Exponentiation":
        return ("exp", spec_from_sklearn(k.kernel), float(k.exponent))
    if name == "ConstantKernel":
        return ("const", float(k.constant_value), _is_fixed(k.constant_value_bounds))
    if name == "WhiteKernel":
        return ("white", float(k.noise_level), _is_fixed(k.noise_level_bounds))
    if name == "RBF":
        return ("rbf", _ls(k.length_scale), _is_fixed(k.length_scale_bounds))
    if name == "Matern":
        return ("matern", _ls(k.length_scale), float(k.nu), _is_fixed(k.length_scale_bounds))
    raise NotImplementedError(f"kernel {name} is outside the restated path")


def _is_fixed(bounds):
    return isinstance(bounds, str) and bounds == "fixed"


def _ls(v):
    return float(v) if np.ndim(v) == 0 else np.asarray(v, dtype=np.float64).copy()


def n_theta(spec):
    tag = spec[0]
    if tag in ("sum", "product"):
        return n_theta(spec[1]) + n_theta(spec[2])
    if tag == "exp":
        return n_theta(spec[1])
    if spec[-1]:
        return 0
    if tag in ("rbf", "matern"):
        return int(np.size(spec[1]))
    return 1


def get_theta(spec):
    """log of every free hyper-parameter, sklearn order."""
    tag = spec[0]
    if tag in ("sum", "product"):
        return np.concatenate([get_theta(spec[1]), get_theta(spec[2])])
    if tag == "exp":
        return get_theta(spec[1])
    if spec[-1]:
        return np.zeros(0)
    with np.errstate(divide="ignore"):
        return np.log(np.atleast_1d(np.asarray(spec[1], dtype=np.float64)))


def with_theta(spec, theta):
    """Returns a new spec whose free hyper-parameters are exp(theta)."""
    out, used = _with_theta(spec, np.asarray(theta, dtype=np.float64), 0)
    if used != len(theta):
        raise ValueError(f"theta has {len(theta)} entries, kernel takes {used}")
    return out


def _with_theta(spec, theta, at):
    tag = spec[0]
    if tag in ("sum", "product"):
        a, at = _with_theta(spec[1], theta, at)
        b, at = _with_theta(spec[2], theta, at)
        return (tag, a, b), at
    if tag == "exp":
        a, at = _with_theta(spec[1], theta, at)
        return (tag, a, spec[2]), at
    if spec[-1]:
        return spec, at
    if tag in ("rbf", "matern"):
        k = int(np.size(spec[1]))
        vals = np.exp(theta[at:at + k])
        ls = float(vals[0]) if np.ndim(spec[1]) == 0 else vals
        return (tag, ls) + tuple(spec[2:]), at + k
    return (tag, float(np.exp(theta[at]))) + tuple(spec[2:]), at + 1


def zero_white(spec):
    """skopt fit / noise_set_to_zero: the White kernel that is a direct child of a (nested)
    Sum is replaced by WhiteKernel(0) (oracle/skopt_port.py:_param_for_white_kernel_in_Sum).
    Only the first one found, depth-first k1 before k2."""
    done = [False]

    def walk(s):
        if s[0] == "sum" and not done[0]:
            kids = []
            for child in (s[1], s[2]):
                if done[0]:
                    kids.append(child)
                elif child[0] == "white":
                    done[0] = True
                    kids.append(("white", 0.0, child[2]))
                else:
                    kids.append(walk(child))
            return ("sum", kids[0], kids[1])
        return s

    return walk(spec)


def white_level(spec):
    """noise level of the White kernel found by the same search as ``zero_white``."""
    if spec[0] == "sum":
        for child in (spec[1], spec[2]):
            if child[0] == "white":
                return child[1]
            v = white_level(child)
            if v is not None:
                return v
    return None


# ------------------------------------------------------------- kernel evaluation
def _stationary(tag, spec, X, Y):
    ls = spec[1]
    Xs = X / ls
    if tag == "rbf":
        if Y is None:
            K = squareform(np.exp(-0.5 * pdist(Xs, metric="sqeuclidean")))
            np.fill_diagonal(K, 1)
            return K
        return np.exp(-0.5 * cdist(Xs, Y / ls, metric="sqeuclidean"))
    nu = spec[2]
    d = pdist(Xs, metric="euclidean") if Y is None else cdist(Xs, Y / ls, metric="euclidean")
    if nu == 0.5:
        K = np.exp(-d)
    elif nu == 1.5:
        K = d * math.sqrt(3)
        K = (1.0 + K) * np.exp(-K)
    elif nu == 2.5:
        K = d * math.sqrt(5)
        K = (1.0 + K + K ** 2 / 3.0) * np.exp(-K)
    elif nu == np.inf:
        K = np.exp(-(d ** 2) / 2.0)
    else:
        raise NotImplementedError("general-nu Matern (Bessel kv) is outside the path")
    if Y is None:
        K = squareform(K)
        np.fill_diagonal(K, 1)
    return K


def kernel_matrix(spec, X, Y=None):
    """k(X, X) when Y is None, else the cross kernel k(X, Y)."""
    tag = spec[0]
    X = np.atleast_2d(X)
    if tag == "sum":
        return kernel_matrix(spec[1], X, Y) + kernel_matrix(spec[2], X, Y)
    if tag == "product":
        return kernel_matrix(spec[1], X, Y) * kernel_matrix(spec[2], X, Y)
    if tag == "exp":
        return kernel_matrix(spec[1], X, Y) ** spec[2]
    ny = X.shape[0] if Y is None else np.atleast_2d(Y).shape[0]
    if tag == "const":
        return np.full((X.shape[0], ny), spec[1], dtype=np.float64)
    if tag == "white":
        if Y is None:
            return spec[1] * np.eye(X.shape[0])
        return np.zeros((X.shape[0], ny))
    return _stationary(tag, spec, X, None if Y is None else np.atleast_2d(Y))


def kernel_diag(spec, X):
    tag = spec[0]
    m = np.atleast_2d(X).shape[0]
    if tag == "sum":
        return kernel_diag(spec[1], X) + kernel_diag(spec[2], X)
    if tag == "product":
        return kernel_diag(spec[1], X) * kernel_diag(spec[2], X)
    if tag == "exp":
        return kernel_diag(spec[1], X) ** spec[2]
    if tag in ("const", "white"):
        return np.full(m, spec[1], dtype=np.float64)
    return np.ones(m)


# ------------------------------------------------------- kernel gradient (MAP start of fit)
def _stationary_gradient(tag, spec, X):
    """d k(X, X) / d log(length scale): sklearn kernels.py:1571-1587 (RBF) and :1744-1786 (Matern).
    D[i, j, k] = ((x_ik - x_jk) / l_k)^2; an isotropic length scale gets the sum over k."""
    ls = spec[1]
    fixed = spec[-1]
    n = X.shape[0]
    if fixed:
        return np.zeros((n, n, 0))
    iso = np.ndim(ls) == 0
    D = (X[:, None, :] - X[None, :, :]) ** 2 / (np.asarray(ls, dtype=np.float64) ** 2)
    r2 = D.sum(-1)
    if tag == "rbf" or spec[2] == np.inf:
        K = np.exp(-0.5 * r2)
        G = D * K[..., None]
    else:
        nu = spec[2]
        if nu == 0.5:
            K = np.exp(-np.sqrt(r2))
            with np.errstate(divide="ignore", invalid="ignore"):
                G = K[..., None] * D / np.sqrt(r2)[..., None]
            G[~np.isfinite(G)] = 0
        elif nu == 1.5:
            G = 3 * D * np.exp(-np.sqrt(3 * r2))[..., None]
        elif nu == 2.5:
            t = np.sqrt(5 * r2)[..., None]
            G = 5.0 / 3.0 * D * (t + 1) * np.exp(-t)
        else:
            raise NotImplementedError("general-nu Matern (Bessel kv) is outside the path")
    return G.sum(-1, keepdims=True) if iso else G


def kernel_gradient(spec, X):
    """(K, dK/dtheta) with dK of shape (n, n, n_theta(spec)), theta in log space, sklearn order --
    Sum :856-873, Product :957-973, Exponentiation :1104-1118, ConstantKernel :1284-1296,
    WhiteKernel :1407-1419 of sklearn/gaussian_process/kernels.py."""
    tag = spec[0]
    X = np.atleast_2d(X)
    n = X.shape[0]
    if tag == "sum":
        K1, G1 = kernel_gradient(spec[1], X)
        K2, G2 = kernel_gradient(spec[2], X)
        return K1 + K2, np.dstack((G1, G2))
    if tag == "product":
        K1, G1 = kernel_gradient(spec[1], X)
        K2, G2 = kernel_gradient(spec[2], X)
        return K1 * K2, np.dstack((G1 * K2[:, :, None], G2 * K1[:, :, None]))
    if tag == "exp":
        K, G = kernel_gradient(spec[1], X)
        return K ** spec[2], G * (spec[2] * K[:, :, None] ** (spec[2] - 1))
    K = kernel_matrix(spec, X)
    if tag == "const":
        return K, (np.zeros((n, n, 0)) if spec[2] else np.full((n, n, 1), spec[1], dtype=np.float64))
    if tag == "white":
        return K, (np.zeros((n, n, 0)) if spec[2] else (spec[1] * np.eye(n))[:, :, None])
    return K, _stationary_gradient(tag, spec, X)


def lml_gradient(spec, theta, X, y, alpha):
    """(LML, dLML/dtheta): sklearn _gpr.py:583-651 --
    0.5 * einsum("ijl,jik->kl", alpha alpha^T - K^-1, dK) for the single-output case."""
    K, dK = kernel_gradient(with_theta(spec, theta), X)
    K = K.copy()
    K[np.diag_indices_from(K)] += alpha
    try:
        L = cholesky(K, lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        return -np.inf, np.zeros(len(theta))
    a = cho_solve((L, True), y, check_finite=False)
    lml = float(-0.5 * np.dot(y, a) - np.log(np.diag(L)).sum() - K.shape[0] / 2 * LOG_2PI)
    inner = np.outer(a, a) - cho_solve((L, True), np.eye(K.shape[0]), check_finite=False)
    return lml, 0.5 * np.einsum("ij,jik->k", inner, dK)


# ------------------------------------------------------------------------ priors
class HalfNormalOnSqrt:
    """log-density of theta = log v when sqrt(v) ~ half-normal(scale): the prior
    bask/utils.py:95-99 puts on every ConstantKernel / WhiteKernel level."""

    def __init__(self, scale=2.0):
        self.scale = float(scale)

    def __call__(self, x):
        return halfnorm(scale=self.scale).logpdf(np.sqrt(np.exp(x))) + x / 2.0 - np.log(2.0)


class RoundFlat:
    """bask/priors.py:7-57 evaluated at exp(theta) plus the Jacobian theta
    (bask/utils.py:113-120)."""

    def __init__(self, lower_bound=0.1, upper_bound=0.6, lower_steepness=2.0,
                 upper_steepness=8.0, integration_bounds=(0.0, 10.0)):
        from scipy.integrate import quad
        self.lo, self.hi = float(lower_bound), float(upper_bound)
        self.slo, self.shi = float(lower_steepness), float(upper_steepness)
        with np.errstate(divide="ignore", over="ignore"):
            self.norm = quad(lambda v: np.exp(self._raw(v)), integration_bounds[0],
                             integration_bounds[1])[0]

    def _raw(self, v):
        return -2 * ((v / self.lo) ** (-2 * self.slo) + (v / self.hi) ** (2 * self.shi))

    def density_log(self, v):
        return self._raw(v) - np.log(self.norm)

    def __call__(self, x):
        return self.density_log(np.exp(x)) + x


class InvGammaOnValue:
    """theta = log v, v ~ inverse-gamma(a, scale) (tests/test_acquisition.py:36)."""

    def __init__(self, a, scale=1.0):
        self.a, self.scale = float(a), float(scale)

    def __call__(self, x):
        return invgamma(a=self.a, scale=self.scale).logpdf(np.exp(x)) + x


class NormalOnTheta:
    """theta ~ N(loc, scale) directly in log space (the default warp prior form,
    bask/bayesgpr.py:462-466)."""

    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = float(loc), float(scale)

    def __call__(self, x):
        return norm(loc=self.loc, scale=self.scale).logpdf(x)


def guess_priors(spec):
    """bask/utils.py:68-124,154-179: one prior per free hyper-parameter, theta order."""
    out = []

    def walk(s):
        tag = s[0]
        if tag == "exp":
            walk(s[1])
        elif tag in ("sum", "product"):
            walk(s[1])
            walk(s[2])
        elif tag in ("const", "white"):
            if not s[-1]:
                out.append(HalfNormalOnSqrt(2.0))
        elif tag in ("rbf", "matern"):
            # NB the reference adds length-scale priors even for fixed bounds (utils.py:100-120)
            rf = RoundFlat(0.1, 0.6, 2.0, 8.0)
            out.extend([rf] * int(np.size(s[1])))
        else:
            raise NotImplementedError(tag)

    walk(spec)
    return out


# ---------------------------------------------------------------- GP at one theta
def gram(spec, theta, X, alpha):
    K = kernel_matrix(with_theta(spec, theta), X)
    K[np.diag_indices_from(K)] += alpha
    return K


def log_marginal_likelihood(spec, theta, X, y, alpha):
    """sklearn _gpr.py:583-617: -1/2 y^T K^-1 y - sum log diag L - n/2 log 2 pi; -inf when
    the Cholesky factorisation fails (:592-593)."""
    K = gram(spec, theta, X, alpha)
    try:
        L = cholesky(K, lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        return -np.inf
    a = cho_solve((L, True), y, check_finite=False)
    return float(-0.5 * np.dot(y, a) - np.log(np.diag(L)).sum() - K.shape[0] / 2 * LOG_2PI)


def log_prob(spec, theta, X, y, alpha, priors):
    """bask/bayesgpr.py:351-379 without input warping."""
    lp = 0
    if callable(priors):
        lp += priors(theta)
    else:
        if len(priors) != len(theta):
            raise ValueError("zip() argument lengths differ")  # zip(strict=True)
        for prior, val in zip(priors, theta):
            lp += prior(val)
    try:
        lp = lp + log_marginal_likelihood(spec, theta, X, y, alpha)
    except ValueError:
        return -np.inf
    if not np.isfinite(lp):
        return -np.inf
    return float(lp)


def warp_inputs(X, a_log, b_log):
    """Beta-CDF warp of every input dimension (bask/bayesgpr.py:249-264, 298-316): column k is
    mapped through scipy.stats.beta(exp(a_log[k]), exp(b_log[k])).cdf."""
    import scipy.stats as st
    X = np.asarray(X, dtype=np.float64)
    out = np.empty_like(X)
    for k in range(X.shape[1]):
        out[:, k] = st.beta(a=np.exp(a_log[k]), b=np.exp(b_log[k])).cdf(X[:, k])
    return out


def log_prob_warped(spec, theta_full, X, y, alpha, priors, warp_priors):
    """bask/bayesgpr.py:351-379 with warp_inputs=True: theta_full = kernel theta ++ log a ++ log b;
    warp_priors is the (prior_a, prior_b) pair applied per dimension."""
    d = X.shape[1]
    theta_full = np.asarray(theta_full, dtype=np.float64)
    x_gp, a_log, b_log = theta_full[:-2 * d], theta_full[-2 * d:-d], theta_full[-d:]
    lp = 0
    for a, b in zip(a_log, b_log):
        lp += warp_priors[0](a) + warp_priors[1](b)
    return lp + log_prob(spec, x_gp, warp_inputs(X, a_log, b_log), y, alpha, priors) if np.isfinite(lp) else -np.inf


def factorize(spec, theta, X, y, alpha):
    """bask/bayesgpr.py:200-217: L_, K_inv_ = L^-T L^-1, alpha_ = K^-1 y."""
    K = gram(spec, theta, X, alpha)
    L = cholesky(K, lower=True)
    L_inv = solve_triangular(L.T, np.eye(L.shape[0]))
    K_inv = L_inv.dot(L_inv.T)
    a = cho_solve((L, True), y)
    return L, K_inv, a


def predict(spec, theta, X, Xs, K_inv, a, y_mean=0.0, y_std=1.0, noise_zero=True):
    """oracle/skopt_port.py predict(return_std=True) at theta; ``noise_zero`` evaluates the
    kernel with the White level at 0 while K_inv/alpha_ stay those of the noisy fit
    (bask/bayesgpr.py:318-336, bask/acquisition.py:122-129)."""
    k = with_theta(spec, theta)
    if noise_zero:
        k = zero_white(k)
    K_trans = kernel_matrix(k, Xs, X)
    mu = y_std * K_trans.dot(a) + y_mean
    var = kernel_diag(k, Xs)
    var = var - np.einsum("ki,kj,ij->k", K_trans, K_trans, K_inv)
    var[var < 0] = 0.0
    return mu, np.sqrt(var * y_std ** 2)


def predict_cov(spec, theta, X, Xs, L, a, y_mean=0.0, y_std=1.0, noise_zero=True):
    """oracle/skopt_port.py predict(return_cov=True)."""
    k = with_theta(spec, theta)
    if noise_zero:
        k = zero_white(k)
    K_trans = kernel_matrix(k, Xs, X)
    mu = y_std * K_trans.dot(a) + y_mean
    v = cho_solve((L, True), K_trans.T)
    cov = kernel_matrix(k, Xs) - K_trans.dot(v)
    return mu, cov * y_std ** 2


def sample_y(spec, theta, X, Xs, L, a, rng, n_samples=1, y_mean=0.0, y_std=1.0,
             noise_zero=True):
    """sklearn _gpr.py:502-539 on top of the skopt predict: numpy's SVD-based
    ``multivariate_normal``; returns (m, n_samples)."""
    mu, cov = predict_cov(spec, theta, X, Xs, L, a, y_mean, y_std, noise_zero)
    return rng.multivariate_normal(mu, cov, n_samples).T


def geometric_median(P, eps=1e-5):
    """Weiszfeld iteration, bask/utils.py:21-65."""
    P = np.asarray(P, dtype=np.float64)
    y = np.mean(P, 0)
    while True:
        D = cdist(P, [y])
        nz = (D != 0)[:, 0]
        Dinv = 1 / D[nz]
        Dinvs = np.sum(Dinv)
        T = np.sum((Dinv / Dinvs) * P[nz], 0)
        zeros = len(P) - np.sum(nz)
        if zeros == 0:
            y1 = T
        elif zeros == len(P):
            return y
        else:
            R = (T - y) * Dinvs
            r = np.linalg.norm(R)
            rinv = 0 if r == 0 else zeros / r
            y1 = max(0, 1 - rinv) * T + min(1, rinv) * y
        if np.sqrt(np.sum((y - y1) ** 2)) < eps:
            return y1
        y = y1
