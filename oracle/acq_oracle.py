"""TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of the reference's acquisition sweep
(not part of the product; the product never imports this).

Restates bask/acquisition.py: ``evaluate_acquisitions`` (:48-147), ``_ei_f`` (:150-151),
ExpectedImprovement (:154-172), TopTwoEI (:175-194), Expectation (:197-201), LCB (:204-216),
MaxValueSearch (:219-267), ThompsonSampling (:270-274), VarianceReduction (:277-300) and
PVRS (:303-339), as plain functions of (mu, std) or of a ``GPState``.

Pinned by tests/test_oracle_golden.py against vectors produced by the unmodified reference
(tests/golden/make_golden.py).
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.stats as st
from scipy.linalg import cho_solve, cholesky
from scipy.optimize import brentq

from . import gp_oracle as G

UNCERTAINTY = ("ei", "ttei", "mean", "lcb", "mes")
SAMPLE = ("ts",)
FULL_GP = ("vr", "pvrs")


def _ei_f(x):
    return x * st.norm.cdf(x) + st.norm.pdf(x)


def expected_improvement(mu, std, y_opt=None, **_):
    if y_opt is None:
        y_opt = mu.min()
    values = np.zeros_like(mu)
    mask = std > 0
    inner = (y_opt - mu[mask]) / std[mask]
    values[mask] = _ei_f(inner) * std[mask]
    return values


def top_two_ei(mu, std, y_opt=None, **_):
    ei = expected_improvement(mu, std, y_opt=y_opt)
    values = np.zeros_like(mu)
    best = np.argmax(ei)
    mask = std > 0
    outer = np.sqrt(np.power(std[mask], 2) + np.power(std[best], 2))
    inner = (mu[best] - mu[mask]) / outer
    values[mask] = outer * _ei_f(inner)
    return values


def expectation(mu, std, **_):
    return -mu


def lcb(mu, std, alpha=1.96, **_):
    if alpha == "inf":
        return std
    return alpha * std - mu


def mes_gumbel_fit(mu, std):
    """The three brentq roots and the Gumbel (location a, scale b) of
    bask/acquisition.py:235-252.  Returns (a, b, (q1, med, q2))."""
    mean = -mu

    def probf(x):
        return np.exp(np.sum(st.norm.logcdf((x - mean) / std), axis=0))

    left = np.min(mean - 3 * std)
    right = np.max(mean + 5 * std)
    q1, med, q2 = [brentq(lambda x, val=val: probf(x) - val, left, right)
                   for val in (0.25, 0.5, 0.75)]
    beta = (q1 - q2) / (np.log(np.log(4.0 / 3.0)) - np.log(np.log(4.0)))
    alpha = med + beta * np.log(np.log(2.0))
    return alpha, beta, (q1, med, q2)


def max_value_search(mu, std, n_min_samples=1000, uniforms=None, **_):
    """bask/acquisition.py:234-267.  ``uniforms`` injects the float32 U(0,1) variates the
    reference takes from the GLOBAL numpy RNG (:253-254); None draws them the same way."""
    alpha, beta, _q = mes_gumbel_fit(mu, std)
    if uniforms is None:
        uniforms = np.random.rand(n_min_samples).astype(np.float32)
    max_values = -np.log(-np.log(uniforms)) * beta + alpha
    mean = -mu
    gamma = (max_values[None, :] - mean[:, None]) / std[:, None]
    return np.sum(gamma * st.norm().pdf(gamma) / (2.0 * st.norm().cdf(gamma))
                  - st.norm().logcdf(gamma), axis=1) / n_min_samples


def thompson(gp_sample, **_):
    return -gp_sample


UNCERTAINTY_FN = {"ei": expected_improvement, "ttei": top_two_ei, "mean": expectation,
                  "lcb": lcb, "mes": max_value_search}


# ------------------------------------------------------------------ GP container
@dataclass
class GPState:
    """What the acquisition sweep reads from a fitted BayesGPR (bask/bayesgpr.py:116-137)."""
    spec: tuple                 # kernel tree WITH the White kernel (levels are placeholders)
    X: np.ndarray
    y: np.ndarray               # normalised targets (y_train_)
    alpha: object               # scalar or (n,) jitter + per-point noise (gp.alpha)
    chain: np.ndarray           # (N, p) hyper-posterior samples (chain_)
    theta: np.ndarray           # current point estimate; its White entry may be -inf
    y_mean: float = 0.0
    y_std: float = 1.0
    L: np.ndarray = field(default=None)
    K_inv: np.ndarray = field(default=None)
    a: np.ndarray = field(default=None)

    def set_theta(self, theta):
        """the ``theta`` setter, bask/bayesgpr.py:200-217"""
        self.theta = np.array(theta, dtype=np.float64)
        self.L, self.K_inv, self.a = G.factorize(self.spec, self.theta, self.X, self.y, self.alpha)


def variance_reduction(Xc, gp: GPState, points=None):
    """bask/acquisition.py:285-300 (points=None -> all candidates) and the loop of PVRS
    (:328-339): one fresh (n+1)x(n+1) Cholesky per candidate, kernel at gp.theta with noise ON."""
    k = G.with_theta(gp.spec, gp.theta)
    pts = Xc if points is None else points
    covs = np.empty(len(Xc))
    for i in range(len(Xc)):
        X_aug = np.concatenate([gp.X, [Xc[i]]])
        K = G.kernel_matrix(k, X_aug)
        if np.iterable(gp.alpha):
            K[np.diag_indices_from(K)] += np.concatenate([gp.alpha, [0.0]])
        L = cholesky(K, lower=True)
        K_trans = G.kernel_matrix(k, pts, X_aug)
        v = cho_solve((L, True), K_trans.T)
        covs[i] = np.diag(K_trans.dot(v)).sum()
    return covs


def pvrs(Xc, gp: GPState, n_thompson=10, random_state=None, thompson_sample=None, **_):
    """bask/acquisition.py:316-339.  ``thompson_sample`` (m, n_thompson) injects the joint
    draws; None draws them like ``gp.sample_y(sample_mean=True)`` (noise-free kernel)."""
    if thompson_sample is None:
        rng = random_state if hasattr(random_state, "multivariate_normal") else \
            np.random.RandomState(random_state)
        thompson_sample = G.sample_y(gp.spec, gp.theta, gp.X, Xc, gp.L, gp.a, rng,
                                     n_samples=n_thompson, y_mean=gp.y_mean, y_std=gp.y_std)
    points = np.array(Xc)[np.argmin(thompson_sample, axis=0)]
    return variance_reduction(Xc, gp, points=points)


def evaluate_acquisitions(Xc, gp: GPState, acquisitions, n_samples=10, random_state=None,
                          mes_uniforms=None, **kwargs):
    """bask/acquisition.py:48-147.  ``acquisitions`` is a sequence of the registry strings of
    bask/optimizer.py:23-32.  Returns (n_acq, m)."""
    m = len(Xc)
    out = np.zeros((len(acquisitions), m))
    rng = random_state if hasattr(random_state, "choice") else np.random.RandomState(random_state)
    picks = rng.choice(len(gp.chain), replace=False, size=n_samples)
    theta_backup = np.copy(gp.theta)
    for j, name in enumerate(acquisitions):
        if name in FULL_GP:
            if name == "vr":
                val = variance_reduction(Xc, gp)
            else:
                val = pvrs(Xc, gp, random_state=rng, **kwargs)
            if np.all(np.isfinite(val)):
                out[j] = val
    for i in picks:
        gp.set_theta(gp.chain[i])
        mu = std = sample = None
        for j, name in enumerate(acquisitions):
            if name in UNCERTAINTY:
                if mu is None:
                    mu, std = G.predict(gp.spec, gp.theta, gp.X, Xc, gp.K_inv, gp.a,
                                        gp.y_mean, gp.y_std, noise_zero=True)
                kw = dict(kwargs)
                if name == "mes" and mes_uniforms is not None:
                    kw["uniforms"] = mes_uniforms
                with np.errstate(all="ignore"):
                    tmp = UNCERTAINTY_FN[name](mu, std, **kw)
            elif name in SAMPLE:
                if sample is None:
                    # bask/bayesgpr.py:679: sample_y draws its OWN chain index
                    pick = rng.choice(len(gp.chain), size=1, replace=True)[0]
                    L, _Ki, a = G.factorize(gp.spec, gp.chain[pick], gp.X, gp.y, gp.alpha)
                    sample = G.sample_y(gp.spec, gp.chain[pick], gp.X, Xc, L, a, rng, 1,
                                        gp.y_mean, gp.y_std).flatten()
                tmp = thompson(sample)
            else:
                continue
            if np.all(np.isfinite(tmp)):
                out[j] += tmp / n_samples
    gp.set_theta(theta_backup)
    return out
