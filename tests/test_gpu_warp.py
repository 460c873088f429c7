"""GPU: input warping (warp_inputs=True) against vectors produced by the unmodified
reference (tests/golden/make_golden.py g6): the Beta-CDF warp runs on the device per theta row
(theta = kernel theta ++ log a ++ log b)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def _kernel(d):
    from bask_b200.utils import construct_default_kernel
    from sklearn.gaussian_process.kernels import WhiteKernel
    return construct_default_kernel(list(range(d))) + WhiteKernel()


def _engine(g, d):
    import bask_b200  # noqa: F401
    from bask_b200._engine import Engine
    from bask_b200.priors import NormalPrior, as_device_priors
    from bask_b200.utils import guess_priors
    e = Engine()
    k = _kernel(d)
    e.set_kernel(k, n_warp=d)
    table, host = as_device_priors(guess_priors(k), e.p_kernel)
    wt, whost = as_device_priors([NormalPrior(0.0, 0.3)] * (2 * d), 2 * d)
    assert host is None and whost is None
    e.set_priors(table + wt)
    e.set_data(g["X"], g["y_train"], g["alpha_vec"])
    return e


def test_logprob_with_warp(g6):
    e = _engine(g6, 2)
    lp, lml, info = e.logprob(g6["thetas"])
    assert (info == 0).all()
    np.testing.assert_allclose(lml, g6["lml"], rtol=RTOL)
    np.testing.assert_allclose(lp, g6["logprob"], rtol=RTOL)


def test_predict_with_warp(g6):
    e = _engine(g6, 2)
    th = e.to_dev(g6["thetas"])
    f = e.factorize(th)
    assert (e.to_host(f.info) == 0).all()
    mu, sd, _, _ = e.predict(f, e.to_dev(g6["Xc"]), noise_off=True, y_mean=float(g6["y_mean"][0]),
                             y_std=float(g6["y_std"][0]))
    np.testing.assert_allclose(e.to_host(mu), g6["mu"], rtol=RTOL, atol=1e-10)
    # sigma^2 = k** - k* K^-1 k*^T cancels near the training points and K is ill-conditioned
    # (alpha = 1e-10): the ~1e-16 differences between the two incomplete-beta implementations in
    # the warped coordinates show up at 1e-6 relative in a few small sigmas -- in BOTH codes
    sd_h = e.to_host(sd)
    np.testing.assert_allclose(sd_h, g6["std"], rtol=1e-5, atol=1e-9)
    assert np.mean(np.abs(sd_h - g6["std"]) <= 1e-7 * np.abs(g6["std"]) + 1e-9) > 0.97


def _fitted_like_reference(g6):
    """Estimator in the state the reference reached after fit(): same chain, point estimate and warps."""
    import bask_b200
    from bask_b200.utils import construct_default_kernel
    gp = bask_b200.BayesGPR(kernel=construct_default_kernel([0, 1]), normalize_y=True, warp_inputs=True,
                            random_state=3)
    gp.fit(g6["X"], g6["y_raw"], noise_vector=g6["noise_vector"], n_desired_samples=100, n_burnin=2,
           n_walkers_per_thread=100, progress=False)
    gp.chain_ = g6["chain"].copy()
    gp.create_warpers(g6["warp_alphas"], g6["warp_betas"])
    gp.rewarp()
    gp.theta = g6["theta_median"]
    return gp


def test_estimator_surface_with_warp(g6):
    import bask_b200
    gp = _fitted_like_reference(g6)
    assert gp.warp_inputs and gp.chain_.shape[1] == 4 + 2 * 2
    np.testing.assert_allclose(gp.X_train_, g6["X_train_warped"], rtol=1e-12)
    np.testing.assert_allclose(gp.warp(g6["Xc"]), g6["warp_of_Xc"], rtol=1e-12)
    np.testing.assert_allclose(gp.unwarp(gp.warp(g6["Xc"][:50])), g6["unwarp_roundtrip"], rtol=1e-9)
    mu, std = gp.predict(g6["Xc"], return_std=True)
    np.testing.assert_allclose(mu, g6["mu_median"], rtol=RTOL, atol=1e-10)
    np.testing.assert_allclose(std, g6["std_median"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(gp.log_marginal_likelihood(gp.theta), g6["lml_at_median"][0], rtol=RTOL)
    lp = gp._log_prob_fn(g6["thetas"], bask_b200.guess_priors(gp.kernel_), None)
    np.testing.assert_allclose(lp, g6["logprob"], rtol=RTOL)
    with pytest.raises(ValueError):
        gp.predict(np.array([[0.5, 1.5]]))
    # swept acquisitions: same theta picks (random_state) and the same global-RNG Gumbel draws
    np.random.seed(2)
    vals = bask_b200.evaluate_acquisitions(
        g6["Xc"], gp, [bask_b200.ExpectedImprovement(), bask_b200.LCB(), bask_b200.MaxValueSearch()],
        n_samples=10, random_state=1)
    for v, name in zip(vals, ("ei", "lcb", "mes")):
        ref = g6[f"sweep_{name}"]
        big = np.abs(ref) > 1e-6 * np.abs(ref).max()
        np.testing.assert_allclose(v[big], ref[big], rtol=1e-7)
        assert int(np.argmax(v)) == int(np.argmax(ref))
    # full-GP acquisitions work in the warped space (Thompson indices injected: the joint draws
    # themselves agree with the reference in distribution only)
    vr = bask_b200.VarianceReduction()(g6["Xc"][:100], gp)
    np.testing.assert_allclose(vr, g6["vr"], rtol=1e-6)
    pv = bask_b200.PVRS()(g6["Xc"], gp, thompson_idx=g6["pvrs_thompson_idx"])
    np.testing.assert_allclose(pv, g6["pvrs"], rtol=1e-6)
    assert np.argmax(pv) == np.argmax(g6["pvrs"])
    draws = gp.sample_y(g6["Xc"][:40], n_samples=3, random_state=1)
    assert draws.shape == (40, 3) and np.all(np.isfinite(draws))


def test_optimizer_with_warp():
    import bask_b200
    opt = bask_b200.Optimizer(dimensions=[(0.0, 1.0), (0.0, 1.0)], n_points=200, n_initial_points=6,
                              acq_func="mes", gp_kwargs=dict(warp_inputs=True), random_state=2)
    f = lambda x: (x[0] - 0.3) ** 2 + (x[1] - 0.6) ** 2   # noqa: E731
    for _ in range(6):
        x = opt.ask()
        opt.tell(x, f(x), fit=False)
    opt.tell([0.5, 0.5], f([0.5, 0.5]), n_samples=2, gp_samples=100, gp_burnin=2)
    nxt = opt.ask()
    assert len(nxt) == 2 and all(0.0 <= v <= 1.0 for v in nxt)
    assert opt.gp.chain_.shape[1] == 4 + 4


def test_sampling_with_warp_runs_and_moves_warp_parameters(g6):
    import bask_b200
    from bask_b200.utils import construct_default_kernel
    gp = bask_b200.BayesGPR(kernel=construct_default_kernel([0, 1]), normalize_y=True, warp_inputs=True,
                            random_state=0)
    gp.fit(g6["X"], g6["y_raw"], noise_vector=g6["noise_vector"], n_desired_samples=3000, n_burnin=30,
           n_walkers_per_thread=100, progress=False)
    assert gp.chain_.shape == (3000, 8) and np.all(np.isfinite(gp.chain_))
    # prior N(0, 0.3) on the log warp parameters keeps them near the identity warp; compare the
    # posterior spread with the reference chain's (distributional, Philox vs emcee streams)
    ref = g6["chain"][:, 4:]
    ours = gp.chain_[:, 4:]
    assert np.all(np.abs(ours.mean(0) - ref.mean(0)) < 0.35)
    assert np.all(ours.std(0) < 0.6) and np.all(ours.std(0) > 0.05)
    assert hasattr(gp, "warp_alphas_") and len(gp.warpers_) == 2
