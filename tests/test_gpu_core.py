"""GPU parity of the libbgp kernels (through the C ABI) against the reference's golden vectors
and the numpy oracle.  Tolerance: 1e-8 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

import bench_workloads as W

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def default_kernel(d):
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel
    return ConstantKernel(1.0, (0.1, 2.0)) * Matern([0.3] * d, (0.2, 0.5), nu=2.5) + WhiteKernel()


@pytest.fixture(scope="module")
def engine_factory():
    import bask_b200  # noqa: F401
    from bask_b200._engine import Engine
    from bask_b200.priors import as_device_priors
    from bask_b200.utils import guess_priors

    def make(g, d):
        e = Engine()
        k = default_kernel(d)
        e.set_kernel(k)
        table, host = as_device_priors(guess_priors(k), e.p)
        assert host is None
        e.set_priors(table)
        e.set_data(g["X"], g["y_train"], g["alpha_vec"])
        return e
    return make


@pytest.mark.parametrize("name,d", [("g1", 2), ("g2", 6), ("g3", 6)])
def test_logprob_matches_reference(name, d, request, engine_factory):
    g = request.getfixturevalue(name)
    e = engine_factory(g, d)
    lp, lml, info = e.logprob(g["thetas"])
    assert (info == 0).all()
    np.testing.assert_allclose(lml, g["lml"], rtol=RTOL)
    np.testing.assert_allclose(lp, g["logprob"], rtol=RTOL)


@pytest.mark.parametrize("name,d", [("g1", 2), ("g2", 6), ("g3", 6)])
def test_predict_matches_reference(name, d, request, engine_factory):
    g = request.getfixturevalue(name)
    e = engine_factory(g, d)
    S = len(g["mu"])
    f = e.factorize(g["thetas"][:S])
    Xc = e.to_dev(g["Xc"])
    mu, sd, _, _ = e.predict(f, Xc, noise_off=True, y_mean=float(g["y_mean"][0]), y_std=float(g["y_std"][0]))
    mu, sd = e.to_host(mu), e.to_host(sd)
    assert (e.to_host(f.info) == 0).all()
    np.testing.assert_allclose(e.to_host(f.lml), g["lml"][:S], rtol=RTOL)
    np.testing.assert_allclose(mu, g["mu"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(sd, g["std"], rtol=1e-7, atol=1e-9)
    f1 = e.factorize(g["thetas"][:1])
    mu1, sd1, _, _ = e.predict(f1, Xc, noise_off=False, y_mean=float(g["y_mean"][0]), y_std=float(g["y_std"][0]))
    np.testing.assert_allclose(e.to_host(sd1)[0], g["std_noisy"], rtol=RTOL)
    from bask_b200 import _lib
    alpha = e.to_host(e.extract(f, 0, _lib.EXTRACT_ALPHA))
    np.testing.assert_allclose(alpha, g["alpha_"][0], rtol=1e-7, atol=1e-9)


def test_dense_attributes(g1, engine_factory):
    from bask_b200 import _lib
    e = engine_factory(g1, 2)
    f = e.factorize(g1["theta_median"][None, :])
    L = e.to_host(e.extract(f, 0, _lib.EXTRACT_L))
    Ki = e.to_host(e.extract(f, 0, _lib.EXTRACT_KINV))
    a = e.to_host(e.extract(f, 0, _lib.EXTRACT_ALPHA))
    np.testing.assert_allclose(L, g1["L_median"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(Ki, g1["K_inv_median"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(a, g1["alpha_median"], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("name,d", [("g1", 2), ("g2", 6), ("g3", 6)])
def test_acquisitions_match_reference(name, d, request, engine_factory):
    from bask_b200 import _lib
    g = request.getfixturevalue(name)
    e = engine_factory(g, d)
    S = len(g["mu"])
    mu, sd = e.to_dev(g["mu"]), e.to_dev(g["std"])   # reference moments in: isolates the epilogues
    for kind, key, p0 in [(_lib.ACQ_EI, "ei", float("nan")), (_lib.ACQ_TTEI, "ttei", float("nan")),
                          (_lib.ACQ_LCB, "lcb", 1.96), (_lib.ACQ_MEAN, "mean", 0.0)]:
        out, per, skipped, _ = e.acq(kind, mu, sd, p0=p0)
        per = e.to_host(per)
        np.testing.assert_allclose(per, g[key], rtol=RTOL, atol=1e-200)
        np.testing.assert_allclose(e.to_host(out), g[key].mean(axis=0), rtol=RTOL, atol=1e-200)
        assert (e.to_host(skipped) == 0).all()
    with np.errstate(all="ignore"):
        g32 = -np.log(-np.log(g["mes_uniforms"][:S]))
    assert g32.dtype == np.float32
    out, per, skipped, fit = e.acq(_lib.ACQ_MES, mu, sd, gumbel32=e.to_dev(g32, dtype=__import__("torch").float32),
                                   want_fit=True)
    per = e.to_host(per)
    finite = np.isfinite(g["mes"]).all(axis=1)
    np.testing.assert_array_equal(e.to_host(skipped) == 0, finite)
    np.testing.assert_allclose(per[finite], g["mes"][finite], rtol=RTOL, atol=1e-200)
    for s in range(S):
        assert np.argmax(per[s]) == np.argmax(g["mes"][s])


def test_mcmc_runs_and_moves(g2, engine_factory):
    e = engine_factory(g2, 6)
    pos0 = g2["pos"]
    buf = e.mcmc(pos0, 20, seed=1234)
    chain = e.to_host(buf["chain"])
    lpc = e.to_host(buf["lpc"])
    acc = e.to_host(buf["acc"])
    assert chain.shape == (20, 64, 8) and np.isfinite(chain).all()
    frac = acc.sum() / (20 * 64)
    assert 0.1 < frac < 0.9
    # stored log-probs are the log-posterior of the stored positions
    lp, _, _ = e.logprob(chain[-1])
    np.testing.assert_allclose(lp, lpc[-1], rtol=1e-10)
    # replaying the same seed reproduces the chain bit for bit (graph replay path)
    buf2 = e.mcmc(pos0, 20, seed=1234, buffers=buf)
    np.testing.assert_array_equal(e.to_host(buf2["chain"]), chain)
    buf3 = e.mcmc(pos0, 20, seed=99, buffers=buf)
    assert not np.array_equal(e.to_host(buf3["chain"]), chain)
