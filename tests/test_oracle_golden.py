"""CPU: the numpy oracle (oracle/*.py) against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Tolerance 1e-10 relative -- the oracle repeats the reference's
own library calls, so differences are re-association noise only."""
import numpy as np
import pytest

import bench_workloads as W
from oracle import acq_oracle as A
from oracle import gp_oracle as G
from conftest import load_golden

RTOL = 1e-10


def default_spec(d):
    return ("sum", ("product", ("const", 1.0, False), ("matern", 0.3 * np.ones(d), 2.5, False)),
            ("white", 1.0, False))


def state_from(g, d):
    return A.GPState(spec=default_spec(d), X=g["X"], y=g["y_train"], alpha=g["alpha_vec"],
                     chain=g["chain"], theta=g["theta_median"].copy(),
                     y_mean=float(g["y_mean"][0]), y_std=float(g["y_std"][0]))


@pytest.mark.parametrize("name,d", [("g1", 2), ("g2", 6), ("g3", 6)])
def test_lml_and_logprob(name, d, request):
    g = request.getfixturevalue(name)
    spec = default_spec(d)
    priors = G.guess_priors(spec)
    lml = [G.log_marginal_likelihood(spec, t, g["X"], g["y_train"], g["alpha_vec"]) for t in g["thetas"]]
    lp = [G.log_prob(spec, t, g["X"], g["y_train"], g["alpha_vec"], priors) for t in g["thetas"]]
    np.testing.assert_allclose(lml, g["lml"], rtol=RTOL)
    np.testing.assert_allclose(lp, g["logprob"], rtol=RTOL)


@pytest.mark.parametrize("name,d", [("g1", 2), ("g2", 6), ("g3", 6)])
def test_predict_and_uncertainty_acquisitions(name, d, request):
    g = request.getfixturevalue(name)
    gp = state_from(g, d)
    for s in range(min(3, len(g["mu"]))):
        gp.set_theta(g["thetas"][s])
        np.testing.assert_allclose(gp.a, g["alpha_"][s], rtol=1e-9, atol=1e-12)
        mu, std = G.predict(gp.spec, gp.theta, gp.X, g["Xc"], gp.K_inv, gp.a, gp.y_mean, gp.y_std)
        np.testing.assert_allclose(mu, g["mu"][s], rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(std, g["std"][s], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(A.expected_improvement(mu, std), g["ei"][s], rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(A.top_two_ei(mu, std), g["ttei"][s], rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(A.lcb(mu, std), g["lcb"][s], rtol=1e-9)
        np.testing.assert_allclose(A.expectation(mu, std), g["mean"][s], rtol=1e-9)
        with np.errstate(all="ignore"):
            mes = A.max_value_search(mu, std, uniforms=g["mes_uniforms"][s])
        np.testing.assert_allclose(mes, g["mes"][s], rtol=1e-9, atol=1e-300)
    gp.set_theta(g["thetas"][0])
    mu, std = G.predict(gp.spec, gp.theta, gp.X, g["Xc"], gp.K_inv, gp.a, gp.y_mean, gp.y_std,
                        noise_zero=False)
    np.testing.assert_allclose(std, g["std_noisy"], rtol=1e-9)


def test_theta_setter_matrices(g1):
    gp = state_from(g1, 2)
    gp.set_theta(g1["theta_median"])
    np.testing.assert_allclose(gp.L, g1["L_median"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(gp.K_inv, g1["K_inv_median"], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(gp.a, g1["alpha_median"], rtol=1e-8, atol=1e-10)


def test_sweep_matches_reference_rng_flow(g1):
    """evaluate_acquisitions incl. theta sub-sampling and the global-RNG MES draws."""
    gp = state_from(g1, 2)
    gp.set_theta(g1["theta_median"])
    np.random.seed(W.config1().mes_seed)
    with np.errstate(all="ignore"):
        out = A.evaluate_acquisitions(g1["Xc"], gp, ("ei", "ttei", "lcb", "mean", "mes"),
                                      n_samples=10, random_state=1)
    for j, n in enumerate(("ei", "ttei", "lcb", "mean", "mes")):
        np.testing.assert_allclose(out[j], g1[f"sweep_{n}"], rtol=1e-8, atol=1e-300)
        assert np.argmax(out[j]) == np.argmax(g1[f"sweep_{n}"])


def test_full_gp_acquisitions(g1, g2):
    gp = state_from(g1, 2)
    gp.set_theta(g1["theta_median"])
    np.testing.assert_allclose(A.variance_reduction(g1["Xc"][:200], gp), g1["vr"], rtol=1e-9)
    pts = g1["Xc"][g1["pvrs_thompson_idx"]]
    np.testing.assert_allclose(A.variance_reduction(g1["Xc"], gp, points=pts), g1["pvrs"], rtol=1e-9)
    # identical RNG flow -> identical Thompson points -> identical values
    np.testing.assert_allclose(A.pvrs(g1["Xc"], gp, 10, np.random.RandomState(5)), g1["pvrs"], rtol=1e-9)
    gp2 = state_from(g2, 6)
    gp2.set_theta(g2["theta_median"])
    pts = g2["Xc"][g2["pvrs_thompson_idx"]]
    np.testing.assert_allclose(A.variance_reduction(g2["Xc"][:60], gp2, points=pts),
                               g2["pvrs"][:60], rtol=1e-9)


def test_joint_posterior(g1):
    gp = state_from(g1, 2)
    gp.set_theta(g1["theta_median"])
    mu, cov = G.predict_cov(gp.spec, gp.theta, gp.X, g1["Xc"][:48], gp.L, gp.a, gp.y_mean, gp.y_std)
    np.testing.assert_allclose(mu, g1["post_mean48"], rtol=1e-10)
    np.testing.assert_allclose(cov, g1["post_cov48"], rtol=1e-9, atol=1e-12)


ZOO = {
    "const_plus_matern15_iso": ("sum", ("sum", ("const", 1.0, False), ("matern", 0.4, 1.5, False)), ("white", 1.0, False)),
    "const_times_rbf_ard": ("sum", ("product", ("const", 1.5, False), ("rbf", np.array([0.3, 0.5, 0.7]), False)), ("white", 1.0, False)),
    "matern05_ard_fixedconst": ("sum", ("product", ("const", 2.0, True), ("matern", np.array([0.5, 0.4, 0.3]), 0.5, False)), ("white", 1.0, False)),
    "exp2_of_sum": ("sum", ("exp", ("sum", ("product", ("const", 0.5, False), ("matern", 0.6, 2.5, False)), ("rbf", np.ones(3), False)), 2.0), ("white", 1.0, False)),
    "matern_inf_iso": ("sum", ("product", ("const", 1.0, False), ("matern", 0.5, np.inf, False)), ("white", 1.0, False)),
    "product_of_stationary": ("sum", ("product", ("product", ("const", 1.0, False), ("rbf", 0.8, False)), ("matern", np.array([0.9, 0.8, 0.7]), 2.5, False)), ("white", 1.0, False)),
}


@pytest.mark.parametrize("name", sorted(ZOO))
def test_kernel_zoo(name, g4):
    spec = ZOO[name]
    X, Xc = g4["X"], g4["Xc"]
    y = g4[f"{name}__y_train"]
    priors = G.guess_priors(spec)
    thetas = g4[f"{name}__thetas"]
    assert thetas.shape[1] == G.n_theta(spec)
    lml = [G.log_marginal_likelihood(spec, t, X, y, 1e-10) for t in thetas]
    np.testing.assert_allclose(lml, g4[f"{name}__lml"], rtol=RTOL)
    lp = [G.log_prob(spec, t, X, y, 1e-10, priors) for t in thetas]
    np.testing.assert_allclose(lp, g4[f"{name}__logprob"], rtol=RTOL)
    for s in range(3):
        L, Ki, a = G.factorize(spec, thetas[s], X, y, 1e-10)
        mu, std = G.predict(spec, thetas[s], X, Xc, Ki, a, float(g4[f"{name}__y_mean"][0]),
                            float(g4[f"{name}__y_std"][0]))
        np.testing.assert_allclose(mu, g4[f"{name}__mu"][s], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(std, g4[f"{name}__std"][s], rtol=1e-7, atol=1e-9)


def test_geometric_median(g4):
    np.testing.assert_allclose(G.geometric_median(g4["geomedian_in"]), g4["geomedian_out"], rtol=1e-12)


def test_known_answer_priors():
    """tests/test_utils.py:20-40 of the reference."""
    spec = ("exp", ("sum", ("sum", ("product", ("const", 1.0, True), ("matern", 1.0, 1.5, False)),
                            ("white", 1.0, False)), ("rbf", np.ones(2), False)), 2.0)
    pr = G.guess_priors(spec)
    assert len(pr) == 4
    for p, v in zip(pr, [-0.02116327824572739, -2.112906921232193, -0.02116327824572739,
                         -0.02116327824572739]):
        assert abs(p(-0.9) - v) < 1e-7


def test_warped_log_prob(g6):
    """Input warping: oracle log-posterior / LML / predictions at full theta rows
    (kernel theta ++ log a ++ log b) against the reference with warp_inputs=True."""
    import scipy.stats as st
    g, d = g6, 2
    spec = default_spec(d)
    priors = G.guess_priors(spec)
    wp = (st.norm(loc=0.0, scale=0.3).logpdf, st.norm(loc=0.0, scale=0.3).logpdf)
    lp = [G.log_prob_warped(spec, t, g["X"], g["y_train"], g["alpha_vec"], priors, wp) for t in g["thetas"]]
    np.testing.assert_allclose(lp, g["logprob"], rtol=RTOL)
    for s in (0, 13):
        t = g["thetas"][s]
        Xw = G.warp_inputs(g["X"], t[-2 * d:-d], t[-d:])
        lml = G.log_marginal_likelihood(spec, t[:-2 * d], Xw, g["y_train"], g["alpha_vec"])
        np.testing.assert_allclose(lml, g["lml"][s], rtol=RTOL)
        L, K_inv, a = G.factorize(spec, t[:-2 * d], Xw, g["y_train"], g["alpha_vec"])
        mu, std = G.predict(spec, t[:-2 * d], Xw, G.warp_inputs(g["Xc"], t[-2 * d:-d], t[-d:]), K_inv, a,
                            float(g["y_mean"][0]), float(g["y_std"][0]))
        np.testing.assert_allclose(mu, g["mu"][s], rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(std, g["std"][s], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(G.warp_inputs(g["X"], g["warp_alphas"], g["warp_betas"]), g["X_train_warped"],
                               rtol=1e-12)


def test_lml_gradient_restatement():
    """N3: oracle gradient (sklearn _gpr.py:619-651 + kernel gradients) against the reference's
    log_marginal_likelihood(eval_gradient=True) on g1/g2 data and on the kernel zoo."""
    g10 = load_golden("g10_lml_gradients.npz")
    for tag, name, d in (("g1", "g1_branin_n20.npz", 2), ("g2", "g2_hartmann6_n100.npz", 6)):
        g = load_golden(name)
        spec = default_spec(d)
        for t, lml_ref, grad_ref in list(zip(g["thetas"], g10[f"{tag}__lml"], g10[f"{tag}__grad"]))[:6]:
            lml, grad = G.lml_gradient(spec, t, g["X"], g["y_train"], g["alpha_vec"])
            np.testing.assert_allclose(lml, lml_ref, rtol=RTOL)
            np.testing.assert_allclose(grad, grad_ref, rtol=1e-9, atol=1e-9 * np.abs(grad_ref).max())
    g4 = load_golden("g4_kernel_zoo.npz")
    zoo = {
        "const_plus_matern15_iso": ("sum", ("const", 1.0, False), ("matern", 0.4, 1.5, False)),
        "matern05_ard_fixedconst": ("product", ("const", 2.0, True), ("matern", np.array([0.5, 0.4, 0.3]), 0.5, False)),
        "exp2_of_sum": ("exp", ("sum", ("product", ("const", 0.5, False), ("matern", 0.6, 2.5, False)),
                                ("rbf", np.ones(3), False)), 2.0),
    }
    for name, base in zoo.items():
        spec = ("sum", base, ("white", 1.0, False))
        for t, grad_ref in zip(g4[f"{name}__thetas"], g10[f"zoo_{name}__grad"]):
            _, grad = G.lml_gradient(spec, t, g4["X"], g4[f"{name}__y_train"], 1e-10 * np.ones(len(g4["X"])))
            np.testing.assert_allclose(grad, grad_ref, rtol=1e-9, atol=1e-9 * np.abs(grad_ref).max(), err_msg=name)
