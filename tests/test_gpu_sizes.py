"""GPU: sizes and shapes around the kernels' internal boundaries, against the numpy oracle
computed on the fly (sizes the oracle finishes in seconds), plus size-independent properties at
the headline size."""
import numpy as np
import pytest

import bench_workloads as W
from oracle import gp_oracle as G

pytestmark = pytest.mark.gpu


def make_engine(X, y, alpha, d):
    import bask_b200  # noqa: F401
    from bask_b200._engine import Engine
    from bask_b200.priors import as_device_priors
    from bask_b200.utils import construct_default_kernel, guess_priors
    from sklearn.gaussian_process.kernels import WhiteKernel
    e = Engine()
    k = construct_default_kernel(list(range(d))) + WhiteKernel()
    e.set_kernel(k)
    e.set_priors(as_device_priors(guess_priors(k), e.p)[0])
    e.set_data(X, y, alpha)
    return e


def spec_for(d):
    return ("sum", ("product", ("const", 1.0, False), ("matern", 0.3 * np.ones(d), 2.5, False)),
            ("white", 1.0, False))


# 1, 2: degenerate; 31/32/33, 64/65: panel and kernel-variant boundaries; 300: clustered CTAs;
# 545: K chunked in shared memory (P-1 > 15); 1100: three K chunks
@pytest.mark.parametrize("n,d", [(1, 1), (2, 3), (31, 2), (32, 2), (33, 2), (64, 4), (65, 4), (300, 6),
                                 (545, 3), (1100, 5)])
def test_logprob_and_moments_across_sizes(n, d):
    r = np.random.RandomState(n)
    X = r.uniform(size=(n, d))
    y = np.sin(3 * X.sum(1)) + 0.1 * r.randn(n)
    y = (y - y.mean()) / (y.std() if n > 1 else 1.0)
    alpha = 1e-10 + 0.01 * r.uniform(size=n)          # ragged per-point noise
    e = make_engine(X, y, alpha, d)
    spec = spec_for(d)
    priors = G.guess_priors(spec)
    B = 5 if n > 600 else 9                            # odd batch: exercises every cluster size
    thetas = W.centre_theta(d) + 0.2 * r.randn(B, d + 2)
    lp, lml, info = e.logprob(thetas)
    ref_lml = [G.log_marginal_likelihood(spec, t, X, y, alpha) for t in thetas]
    ref_lp = [G.log_prob(spec, t, X, y, alpha, priors) for t in thetas]
    assert (info == 0).all()
    np.testing.assert_allclose(lml, ref_lml, rtol=1e-8)
    np.testing.assert_allclose(lp, ref_lp, rtol=1e-8)
    # factorise + sweep on a ragged candidate count
    S, m = 2, 37
    Xc = r.uniform(size=(m, d))
    f = e.factorize(thetas[:S])
    mu, sd, _, _ = e.predict(f, e.to_dev(Xc), noise_off=True, y_mean=0.3, y_std=1.7)
    mu, sd = e.to_host(mu), e.to_host(sd)
    for s in range(S):
        L, Ki, a = G.factorize(spec, thetas[s], X, y, alpha)
        rm, rs = G.predict(spec, thetas[s], X, Xc, Ki, a, 0.3, 1.7)
        np.testing.assert_allclose(mu[s], rm, rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(sd[s], rs, rtol=1e-6, atol=1e-8)


def test_non_positive_definite_gives_minus_inf_not_an_abort():
    """sklearn:_gpr.py:592-593: LinAlgError -> -inf for that theta only."""
    r = np.random.RandomState(0)
    X = r.uniform(size=(40, 2))
    X[7] = X[3]                                       # duplicate point, no jitter, no noise -> singular
    y = r.randn(40)
    e = make_engine(X, y, 0.0, 2)
    good = W.centre_theta(2)
    bad = good.copy()
    bad[-1] = -800.0                                  # exp underflows to 0 noise
    lp, lml, info = e.logprob(np.array([good, bad, good]))
    assert np.isfinite(lp[0]) and np.isfinite(lp[2]) and lp[0] == lp[2]
    assert lp[1] == -np.inf and lml[1] == -np.inf and info[1] > 0
    assert info[0] == 0 and info[2] == 0


def test_headline_size_properties():
    """n=500, d=6 (config 3): properties that do not need the oracle at full size."""
    w = W.config3(m=4096)
    y = (w.y - w.y.mean()) / w.y.std()
    e = make_engine(w.X, y, 1e-10, w.d)
    r = np.random.RandomState(1)
    thetas = W.centre_theta(w.d) + 0.1 * r.randn(130, w.d + 2)     # > one wave of clustered CTAs
    lp, lml, info = e.logprob(thetas)
    assert (info == 0).all() and np.isfinite(lp).all()
    # batching / wave / cluster-size independence: same thetas one at a time, bit for bit
    for i in (0, 64, 129):
        lp1, lml1, _ = e.logprob(thetas[i:i + 1])
        assert lml1[0] == lml[i] and lp1[0] == lp[i]
    # L L^T reproduces the Gram matrix; L^-1 L = I; alpha = K^-1 y
    from bask_b200 import _lib
    f = e.factorize(thetas[:1])
    L = e.to_host(e.extract(f, 0, _lib.EXTRACT_L))
    Li = e.to_host(e.extract(f, 0, _lib.EXTRACT_LINV))
    a = e.to_host(e.extract(f, 0, _lib.EXTRACT_ALPHA))
    K = G.gram(spec_for(w.d), thetas[0], w.X, 1e-10)
    np.testing.assert_allclose(L @ L.T, K, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(Li @ L, np.eye(w.n), atol=1e-9)
    np.testing.assert_allclose(K @ a, y, rtol=1e-6, atol=1e-8)
    # predictions at the training inputs interpolate them when the noise is tiny, and the sweep is
    # permutation equivariant in the candidates
    Xc = w.candidates
    mu, sd, _, _ = e.predict(f, e.to_dev(Xc), noise_off=True)
    perm = r.permutation(len(Xc))
    mu2, sd2, _, _ = e.predict(f, e.to_dev(Xc[perm]), noise_off=True)
    np.testing.assert_array_equal(e.to_host(mu)[0][perm], e.to_host(mu2)[0])
    np.testing.assert_array_equal(e.to_host(sd)[0][perm], e.to_host(sd2)[0])
    assert (e.to_host(sd) >= 0).all()


def test_kernel_zoo_on_device(g4):
    """Every kernel form of tests/golden/g4 through the compiled program (interpreter path)."""
    import bask_b200  # noqa: F401
    from bask_b200._engine import Engine
    from bask_b200.priors import as_device_priors
    from bask_b200.utils import guess_priors
    from sklearn.gaussian_process.kernels import (RBF, ConstantKernel, Exponentiation, Matern, WhiteKernel)
    zoo = {
        "const_plus_matern15_iso": ConstantKernel(1.0) + Matern(0.4, nu=1.5),
        "const_times_rbf_ard": ConstantKernel(1.5) * RBF([0.3, 0.5, 0.7]),
        "matern05_ard_fixedconst": ConstantKernel(2.0, "fixed") * Matern([0.5, 0.4, 0.3], nu=0.5),
        "exp2_of_sum": Exponentiation(ConstantKernel(0.5) * Matern(0.6, nu=2.5) + RBF([1.0, 1.0, 1.0]), 2.0),
        "matern_inf_iso": ConstantKernel(1.0) * Matern(0.5, nu=np.inf),
        "product_of_stationary": ConstantKernel(1.0) * RBF(0.8) * Matern([0.9, 0.8, 0.7], nu=2.5),
    }
    for name, base in zoo.items():
        k = base + WhiteKernel()
        e = Engine()
        e.set_kernel(k)
        e.set_priors(as_device_priors(guess_priors(k), e.p)[0])
        e.set_data(g4["X"], g4[f"{name}__y_train"], 1e-10)
        thetas = g4[f"{name}__thetas"]
        lp, lml, info = e.logprob(thetas)
        np.testing.assert_allclose(lml, g4[f"{name}__lml"], rtol=1e-8, err_msg=name)
        np.testing.assert_allclose(lp, g4[f"{name}__logprob"], rtol=1e-8, err_msg=name)
        f = e.factorize(thetas[:3])
        mu, sd, _, _ = e.predict(f, e.to_dev(g4["Xc"]), noise_off=True, y_mean=float(g4[f"{name}__y_mean"][0]),
                                 y_std=float(g4[f"{name}__y_std"][0]))
        np.testing.assert_allclose(e.to_host(mu), g4[f"{name}__mu"], rtol=1e-8, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(e.to_host(sd), g4[f"{name}__std"], rtol=1e-6, atol=1e-8, err_msg=name)


# the fused shared-memory path (bgp_small.cu) covers n <= ~224; 7/8/9, 16/17: tile boundaries of its 8-wide
# blocks; 96/97: 4- vs 8-warp variant; 224: the largest n whose tiles fit; 225: first n on the blocked path
@pytest.mark.parametrize("n,d", [(3, 1), (7, 2), (8, 2), (9, 2), (16, 3), (17, 3), (20, 2), (96, 6), (97, 6),
                                 (100, 6), (160, 20), (224, 6), (225, 6)])
def test_small_fused_path_matches_oracle_and_blocked_path(n, d, monkeypatch):
    r = np.random.RandomState(1000 + n)
    X = r.uniform(size=(n, d))
    y = np.sin(3 * X.sum(1)) + 0.1 * r.randn(n)
    y = (y - y.mean()) / y.std()
    alpha = 1e-10 + 0.01 * r.uniform(size=n)
    e = make_engine(X, y, alpha, d)
    spec = spec_for(d)
    priors = G.guess_priors(spec)
    B = 333                                            # more thetas than resident CTAs at the larger n
    thetas = W.centre_theta(d) + 0.2 * r.randn(B, d + 2)
    thetas[5, -1] = -800.0                            # zero noise; with duplicate-free X still PD (alpha > 0)
    lp, lml, info = e.logprob(thetas)
    monkeypatch.setenv("BGP_NO_SMALL", "1")
    lp_b, lml_b, info_b = e.logprob(thetas)
    monkeypatch.delenv("BGP_NO_SMALL")
    np.testing.assert_array_equal(info, info_b)
    np.testing.assert_allclose(lml, lml_b, rtol=1e-11)
    np.testing.assert_allclose(lp, lp_b, rtol=1e-11)
    pick = r.choice(B, size=6, replace=False)
    ref_lml = [G.log_marginal_likelihood(spec, t, X, y, alpha) for t in thetas[pick]]
    ref_lp = [G.log_prob(spec, t, X, y, alpha, priors) for t in thetas[pick]]
    np.testing.assert_allclose(lml[pick], ref_lml, rtol=1e-8)
    np.testing.assert_allclose(lp[pick], ref_lp, rtol=1e-8)

