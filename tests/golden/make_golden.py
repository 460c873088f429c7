"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference/bask,
loaded through oracle/ref_loader.py) together with the installed scikit-learn.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

Every array is float64 unless noted.  What each file pins is listed in tests/golden/README.md.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle.ref_loader import load_reference  # noqa: E402

bask = load_reference()
from bask.acquisition import (LCB, PVRS, Expectation, ExpectedImprovement,  # noqa: E402
                              MaxValueSearch, ThompsonSampling, TopTwoEI, VarianceReduction,
                              evaluate_acquisitions)
from bask.bayesgpr import BayesGPR  # noqa: E402
from bask.utils import construct_default_kernel, geometric_median, guess_priors  # noqa: E402
from sklearn.gaussian_process.kernels import (RBF, ConstantKernel, Exponentiation,  # noqa: E402
                                              Matern, WhiteKernel)

import bench_workloads as W  # noqa: E402

warnings.simplefilter("ignore")


def fitted_reference_gp(w, n_desired=None, n_burnin=None, seed=0, kernel=None):
    gp = BayesGPR(kernel=kernel if kernel is not None else construct_default_kernel(list(range(w.d))),
                  normalize_y=True, random_state=seed)
    t0 = time.time()
    gp.fit(w.X, w.y, noise_vector=w.noise_vector,
           n_desired_samples=w.n_desired_samples if n_desired is None else n_desired,
           n_burnin=w.n_burnin if n_burnin is None else n_burnin,
           n_walkers_per_thread=w.n_walkers, progress=False)
    print(f"  reference fit: {time.time() - t0:.1f}s, chain {gp.chain_.shape}")
    return gp


def per_theta_vectors(gp, thetas, Xc, n_acq_theta, mes_seed):
    priors = guess_priors(gp.kernel_)
    out = {}
    out["lml"] = np.array([gp.log_marginal_likelihood(t) for t in thetas])
    out["logprob"] = np.array([gp._log_prob_fn(t, priors=priors, warp_priors=None) for t in thetas])
    theta_backup = gp.theta
    mus, stds, alphas, eis, tteis, lcbs, means, mess, qs = ([] for _ in range(9))
    np.random.seed(mes_seed)
    uniforms = []
    for t in thetas[:n_acq_theta]:
        gp.theta = t
        alphas.append(gp.alpha_.copy())
        with gp.noise_set_to_zero():
            mu, std = gp.predict(Xc, return_std=True)
        mus.append(mu)
        stds.append(std)
        eis.append(ExpectedImprovement()(mu, std))
        tteis.append(TopTwoEI()(mu, std))
        lcbs.append(LCB()(mu, std))
        means.append(Expectation()(mu, std))
        state = np.random.get_state()
        uniforms.append(np.random.rand(1000).astype(np.float32))
        np.random.set_state(state)
        with np.errstate(all="ignore"):
            mess.append(MaxValueSearch()(mu, std))
    # predictive std WITH the noise kernel on, first theta
    gp.theta = thetas[0]
    mu_n, std_n = gp.predict(Xc, return_std=True)
    gp.theta = theta_backup
    out.update(alpha_=np.array(alphas), mu=np.array(mus), std=np.array(stds), ei=np.array(eis),
               ttei=np.array(tteis), lcb=np.array(lcbs), mean=np.array(means),
               mes=np.array(mess), mes_uniforms=np.array(uniforms),
               mu_noisy=mu_n, std_noisy=std_n)
    return out


def common(gp, w):
    return dict(X=w.X, y_raw=w.y, noise_vector=w.noise_vector, y_train=gp.y_train_,
                y_mean=np.atleast_1d(gp.y_train_mean_), y_std=np.atleast_1d(gp.y_train_std_),
                alpha_vec=np.asarray(gp.alpha, dtype=np.float64) * np.ones(w.n),
                chain=gp.chain_, pos=np.asarray(gp.pos_), theta_median=gp.theta,
                lml_at_median=np.atleast_1d(gp.log_marginal_likelihood_value_),
                noise_=np.atleast_1d(gp.noise_))


def sweep_vectors(gp, Xc, acqs, names, n_samples, seed, mes_seed, **kw):
    np.random.seed(mes_seed)
    vals = evaluate_acquisitions(Xc, gp, acqs, n_samples=n_samples, random_state=seed, **kw)
    return {f"sweep_{n}": v for n, v in zip(names, vals)}


def g1():
    print("G1: config 1 (Branin n=20, m=500)")
    w = W.config1()
    gp = fitted_reference_gp(w)
    d = common(gp, w)
    d["thetas"] = gp.chain_[:16].copy()
    d["Xc"] = w.candidates
    d.update(per_theta_vectors(gp, d["thetas"], w.candidates, 16, w.mes_seed))
    gp.theta = d["theta_median"]
    d["L_median"], d["K_inv_median"], d["alpha_median"] = gp.L_.copy(), gp.K_inv_.copy(), gp.alpha_.copy()
    d.update(sweep_vectors(gp, w.candidates,
                           [ExpectedImprovement(), TopTwoEI(), LCB(), Expectation(), MaxValueSearch()],
                           ["ei", "ttei", "lcb", "mean", "mes"], 10, 1, w.mes_seed))
    # full-GP acquisitions at the median theta, noise ON (bask/acquisition.py:106-111)
    d["vr"] = VarianceReduction()(w.candidates[:200], gp)
    rs = np.random.RandomState(5)
    ts = gp.sample_y(w.candidates, sample_mean=True, n_samples=10, random_state=rs)
    d["pvrs_thompson_idx"] = np.argmin(ts, axis=0).astype(np.int64)
    d["pvrs"] = PVRS()(w.candidates, gp, n_thompson=10, random_state=np.random.RandomState(5))
    # joint posterior of a small candidate block at the median theta (noise-free kernel)
    with gp.noise_set_to_zero():
        mu_c, cov_c = gp.predict(w.candidates[:48], return_cov=True)
    d["post_mean48"], d["post_cov48"] = mu_c, cov_c
    np.savez_compressed(os.path.join(HERE, "g1_branin_n20.npz"), **d)


def g2():
    print("G2: config 2 (Hartmann-6 n=100, m=1000)")
    w = W.config2()
    gp = fitted_reference_gp(w)
    d = common(gp, w)
    d["thetas"] = gp.chain_[:16].copy()
    d["Xc"] = w.candidates
    d.update(per_theta_vectors(gp, d["thetas"], w.candidates, 8, w.mes_seed))
    gp.theta = d["theta_median"]
    rs = np.random.RandomState(7)
    ts = gp.sample_y(w.candidates, sample_mean=True, n_samples=10, random_state=rs)
    d["pvrs_thompson_idx"] = np.argmin(ts, axis=0).astype(np.int64)
    t0 = time.time()
    d["pvrs"] = PVRS()(w.candidates, gp, n_thompson=10, random_state=np.random.RandomState(7))
    print(f"  reference PVRS over {len(w.candidates)} candidates: {time.time() - t0:.1f}s")
    d["vr"] = VarianceReduction()(w.candidates[:100], gp)
    np.savez_compressed(os.path.join(HERE, "g2_hartmann6_n100.npz"), **d)


def g3():
    print("G3: config 3 data (n=500, d=6), candidates cut to 1500, 4 theta")
    w = W.config3(m=1500)
    gp = fitted_reference_gp(w, n_desired=128, n_burnin=1)
    d = common(gp, w)
    d["thetas"] = gp.chain_[:16].copy()
    d["Xc"] = w.candidates
    d.update(per_theta_vectors(gp, d["thetas"], w.candidates, 4, w.mes_seed))
    gp.theta = d["theta_median"]
    d.update(sweep_vectors(gp, w.candidates, [MaxValueSearch(), ExpectedImprovement()],
                           ["mes", "ei"], 3, 1, w.mes_seed))
    d["alpha_median"] = gp.alpha_.copy()
    d["L_median_diag"] = np.diag(gp.L_).copy()
    np.savez_compressed(os.path.join(HERE, "g3_wavy6_n500.npz"), **d)


def g4():
    """Kernel zoo: every kernel form the compiled kernel program must reproduce."""
    print("G4: kernel zoo (n=40, d=3)")
    r = np.random.RandomState(123)
    X = r.uniform(size=(40, 3))
    y = np.sin(4 * X[:, 0]) + X[:, 1] ** 2 - 0.5 * X[:, 2] + 0.05 * r.randn(40)
    Xc = r.uniform(size=(64, 3))
    zoo = {
        "const_plus_matern15_iso": ConstantKernel(1.0) + Matern(0.4, nu=1.5),
        "const_times_rbf_ard": ConstantKernel(1.5) * RBF([0.3, 0.5, 0.7]),
        "matern05_ard_fixedconst": ConstantKernel(2.0, "fixed") * Matern([0.5, 0.4, 0.3], nu=0.5),
        "exp2_of_sum": Exponentiation(ConstantKernel(0.5) * Matern(0.6, nu=2.5) + RBF([1.0, 1.0, 1.0]), 2.0),
        "matern_inf_iso": ConstantKernel(1.0) * Matern(0.5, nu=np.inf),
        "product_of_stationary": ConstantKernel(1.0) * RBF(0.8) * Matern([0.9, 0.8, 0.7], nu=2.5),
    }
    d = dict(X=X, y_raw=y, Xc=Xc)
    for name, kern in zoo.items():
        gp = BayesGPR(kernel=kern, normalize_y=True, random_state=3)
        gp.fit(X, y, n_desired_samples=2 * max(8, 2 * (len(kern.theta) + 1)), n_burnin=0,
               n_walkers_per_thread=max(8, 2 * (len(kern.theta) + 1)), progress=False)
        thetas = gp.chain_[:6].copy()
        priors = guess_priors(gp.kernel_)
        d[f"{name}__thetas"] = thetas
        d[f"{name}__lml"] = np.array([gp.log_marginal_likelihood(t) for t in thetas])
        d[f"{name}__logprob"] = np.array([gp._log_prob_fn(t, priors=priors, warp_priors=None) for t in thetas])
        mus, stds = [], []
        for t in thetas[:3]:
            gp.theta = t
            with gp.noise_set_to_zero():
                mu, std = gp.predict(Xc, return_std=True)
            mus.append(mu)
            stds.append(std)
        d[f"{name}__mu"], d[f"{name}__std"] = np.array(mus), np.array(stds)
        d[f"{name}__y_train"] = gp.y_train_
        d[f"{name}__y_mean"] = np.atleast_1d(gp.y_train_mean_)
        d[f"{name}__y_std"] = np.atleast_1d(gp.y_train_std_)
    d["geomedian_in"] = r.randn(50, 4)
    d["geomedian_out"] = geometric_median(d["geomedian_in"])
    np.savez_compressed(os.path.join(HERE, "g4_kernel_zoo.npz"), **d)


def g5():
    """Long reference MCMC on config 1: posterior moments of theta for the distributional check."""
    print("G5: long reference chain on config 1 (100 walkers x 2000 steps after 200 burn-in)")
    w = W.config1()
    gp = fitted_reference_gp(w, n_desired=100 * 2000, n_burnin=200, seed=11)
    d = dict(chain_mean=gp.chain_.mean(axis=0), chain_std=gp.chain_.std(axis=0),
             chain_len=np.array([len(gp.chain_)]), chain_thin=gp.chain_[::200].copy())
    # per-step ensemble statistics (the chain is step-major: 2000 steps x 100 walkers): what the
    # autocorrelation-aware standard errors of the GPU test are computed from, and the acceptance rate
    steps = gp.chain_.reshape(-1, w.n_walkers, gp.chain_.shape[1])
    d["step_means"] = steps.mean(axis=1)
    d["step_sqdev"] = ((steps - d["chain_mean"]) ** 2).mean(axis=1)
    d["acceptance"] = np.atleast_1d(np.mean(np.any(np.diff(steps, axis=0) != 0.0, axis=2)))
    np.savez_compressed(os.path.join(HERE, "g5_branin_long_chain.npz"), **d)


def g6():
    """Input warping (warp_inputs=True) on config 1: log-posterior / LML at full
    theta rows (kernel theta ++ log a ++ log b), predictions and a swept EI / LCB / MES with the
    per-theta warps, the fitted point estimate and its warp parameters."""
    print("G6: config 1 with warp_inputs=True")
    import scipy.stats as st
    w = W.config1()
    gp = BayesGPR(kernel=construct_default_kernel(list(range(w.d))), normalize_y=True, warp_inputs=True,
                  random_state=3)
    gp.fit(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=w.n_desired_samples, n_burnin=w.n_burnin,
           n_walkers_per_thread=w.n_walkers, progress=False)
    d = dict(X=w.X, y_raw=w.y, noise_vector=w.noise_vector, y_train=gp.y_train_,
             y_mean=np.atleast_1d(gp.y_train_mean_), y_std=np.atleast_1d(gp.y_train_std_),
             alpha_vec=np.asarray(gp.alpha, dtype=np.float64) * np.ones(w.n),
             chain=gp.chain_, pos=np.asarray(gp.pos_), theta_median=gp.theta,
             warp_alphas=gp.warp_alphas_.copy(), warp_betas=gp.warp_betas_.copy(),
             lml_at_median=np.atleast_1d(gp.log_marginal_likelihood_value_), noise_=np.atleast_1d(gp.noise_),
             Xc=w.candidates, X_train_warped=gp.X_train_.copy())
    priors = guess_priors(gp.kernel_)
    wp = (st.norm(loc=0.0, scale=0.3).logpdf, st.norm(loc=0.0, scale=0.3).logpdf)
    a_bak, b_bak, th_bak = gp.warp_alphas_.copy(), gp.warp_betas_.copy(), gp.theta
    # a spread of full rows: chain rows and exaggerated warps
    rows = gp.chain_[:12].copy()
    rs = np.random.RandomState(9)
    big = gp.chain_[12:16].copy()
    big[:, -2 * w.d:] = rs.uniform(-1.2, 1.2, size=(4, 2 * w.d))
    rows = np.vstack([rows, big])
    d["thetas"] = rows
    d["logprob"] = np.array([gp._log_prob_fn(t, priors=priors, warp_priors=wp) for t in rows])
    lml, mus, stds = [], [], []
    nk = len(th_bak)
    for t in rows:
        gp.create_warpers(t[nk:nk + w.d], t[nk + w.d:])
        gp.rewarp()
        lml.append(gp.log_marginal_likelihood(t[:nk]))
        gp.theta = t[:nk]
        with gp.noise_set_to_zero():
            mu, std = gp.predict(w.candidates, return_std=True)
        mus.append(mu)
        stds.append(std)
    d.update(lml=np.array(lml), mu=np.array(mus), std=np.array(stds))
    gp.create_warpers(a_bak, b_bak)
    gp.rewarp()
    gp.theta = th_bak
    d["mu_median"], d["std_median"] = gp.predict(w.candidates, return_std=True)
    d["warp_of_Xc"] = gp.warp(w.candidates)
    d["unwarp_roundtrip"] = gp.unwarp(gp.warp(w.candidates[:50]))
    d.update(sweep_vectors(gp, w.candidates, [ExpectedImprovement(), LCB(), MaxValueSearch()],
                           ["ei", "lcb", "mes"], 10, 1, w.mes_seed))
    d["vr"] = VarianceReduction()(w.candidates[:100], gp)
    d["pvrs"] = PVRS()(w.candidates, gp, n_thompson=10, random_state=np.random.RandomState(5))
    rs5 = np.random.RandomState(5)
    ts = gp.sample_y(gp.warp(w.candidates) if False else w.candidates, sample_mean=True, n_samples=10,
                     random_state=rs5)
    d["pvrs_thompson_idx"] = np.argmin(ts, axis=0).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "g6_branin_warp.npz"), **d)


def _light_reference_gp(w, n_walkers, n_steps, seed=0):
    """Reference BayesGPR fitted with a SHORT MCMC (the numbers pinned below are evaluated at fixed
    thetas, so the chain only has to supply plausible hyper-parameters)."""
    gp = BayesGPR(kernel=construct_default_kernel(list(range(w.d))), normalize_y=True, random_state=seed)
    t0 = time.time()
    gp.fit(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=n_walkers, n_burnin=n_steps - 1,
           n_walkers_per_thread=n_walkers, progress=False)
    print(f"  reference fit: {time.time() - t0:.1f}s, chain {gp.chain_.shape}")
    return gp


def g7():
    """BASELINE config 5 (Ackley-20, n=2000, d=20, p=22): LML / log-posterior at 16 thetas, moments,
    EI and MES at 4 thetas x 2000 candidates, and a 4-theta evaluate_acquisitions sweep (EI + MES)."""
    print("G7: config 5 (Ackley-20 n=2000), 2000 candidates")
    w = W.config5(m=2000)
    gp = _light_reference_gp(w, 64, 3)
    d = common(gp, w)
    rs = np.random.RandomState(17)
    base = gp.chain_[rs.choice(len(gp.chain_), 16, replace=False)].copy()
    base[8:] += 0.15 * rs.randn(8, base.shape[1])      # a wider spread than a 3-step chain has
    d["thetas"] = base
    d["Xc"] = w.candidates
    t0 = time.time()
    d.update(per_theta_vectors(gp, d["thetas"], w.candidates, 4, w.mes_seed))
    print(f"  per-theta vectors: {time.time() - t0:.1f}s")
    gp.theta = d["theta_median"]
    t0 = time.time()
    d.update(sweep_vectors(gp, w.candidates, [ExpectedImprovement(), MaxValueSearch()], ["ei", "mes"],
                           4, 1, w.mes_seed))
    print(f"  sweep: {time.time() - t0:.1f}s")
    d["alpha_median"] = gp.alpha_.copy()
    d["L_median_diag"] = np.diag(gp.L_).copy()
    for k in ("alpha_",):      # 4 x 2000 -- keep; drop nothing else: the file stays < 1 MB
        pass
    np.savez_compressed(os.path.join(HERE, "g7_ackley20_n2000.npz"), **d)


def g8():
    """BASELINE config 4 at the large sizes: reference LML / log-posterior at 16 thetas for n=2048 and
    n=4096 (d=6, the C4 generator).  Only scalars are stored; the inputs are regenerated from
    bench_workloads.config4 by the test."""
    print("G8: config 4 LML at n=2048 / 4096")
    d = {}
    for n in (2048, 4096):
        w, thetas = W.config4(n, 16)
        gp = BayesGPR(kernel=construct_default_kernel(list(range(w.d))), normalize_y=True, random_state=0,
                      optimizer=None)
        # no MAP search, no MCMC: the skopt/sklearn fit with optimizer=None only installs the data
        from skopt.learning import GaussianProcessRegressor as SkoptGPR
        SkoptGPR.fit(gp, w.X, w.y)
        priors = guess_priors(gp.kernel_)
        t0 = time.time()
        d[f"n{n}__thetas"] = thetas
        d[f"n{n}__lml"] = np.array([gp.log_marginal_likelihood(t) for t in thetas])
        d[f"n{n}__logprob"] = np.array([gp._log_prob_fn(t, priors=priors, warp_priors=None) for t in thetas])
        d[f"n{n}__y_train"] = gp.y_train_
        print(f"  n={n}: {time.time() - t0:.1f}s")
    np.savez_compressed(os.path.join(HERE, "g8_lml_large_n.npz"), **d)


def g9():
    """Headline size, un-cut: the reference's evaluate_acquisitions over all 10 000 candidates of config 3
    with S=10 thetas (MES + EI), theta picks random_state=1, Gumbel draws np.random.seed(2)."""
    print("G9: config 3 full sweep (n=500, m=10000, S=10)")
    w = W.config3()
    gp = fitted_reference_gp(w, n_desired=128, n_burnin=1)
    d = common(gp, w)
    t0 = time.time()
    d.update(sweep_vectors(gp, w.candidates, [MaxValueSearch(), ExpectedImprovement()],
                           ["mes", "ei"], 10, 1, w.mes_seed))
    print(f"  sweep: {time.time() - t0:.1f}s")
    d = {k: v for k, v in d.items() if k not in ("X", "y_raw", "noise_vector")}   # regenerated by the test
    np.savez_compressed(os.path.join(HERE, "g9_wavy6_full_sweep.npz"), **d)


def g10():
    """LML gradients and the MAP start: sklearn's
    log_marginal_likelihood(theta, eval_gradient=True) -- what the L-BFGS-B search of the skopt fit
    drives (bask/bayesgpr.py:607 -> sklearn _gpr.py:299-344, 583-651) -- at the 16 thetas of g1/g2/g3 and at
    the kernel-zoo thetas of g4, plus the MAP point of that search (theta with the noise level in the White
    slot, noise_, LML) on the g1/g2/g3 data."""
    print("G10: LML gradients + MAP points")
    from skopt.learning import GaussianProcessRegressor as SkoptGPR
    d = {}
    for tag, name, dim in (("g1", "g1_branin_n20.npz", 2), ("g2", "g2_hartmann6_n100.npz", 6),
                           ("g3", "g3_wavy6_n500.npz", 6)):
        g = np.load(os.path.join(HERE, name))
        t0 = time.time()
        gp = BayesGPR(kernel=construct_default_kernel(list(range(dim))), normalize_y=True, random_state=0,
                      optimizer=None, alpha=g["alpha_vec"])
        SkoptGPR.fit(gp, g["X"], g["y_raw"])
        assert np.allclose(gp.y_train_, g["y_train"], rtol=0, atol=1e-13)
        vals = [gp.log_marginal_likelihood(t, eval_gradient=True) for t in g["thetas"]]
        d[f"{tag}__lml"] = np.array([v[0] for v in vals])
        d[f"{tag}__grad"] = np.array([v[1] for v in vals])
        gm = BayesGPR(kernel=construct_default_kernel(list(range(dim))), normalize_y=True, random_state=0,
                      alpha=g["alpha_vec"])
        SkoptGPR.fit(gm, g["X"], g["y_raw"])
        th = gm.kernel_.theta.copy()
        th[np.isinf(th)] = np.log(gm.noise_)
        d[f"{tag}__map_theta"] = th
        d[f"{tag}__map_noise"] = np.atleast_1d(gm.noise_)
        d[f"{tag}__map_lml"] = np.atleast_1d(gm.log_marginal_likelihood_value_)
        print(f"  {tag}: {time.time() - t0:.1f}s  MAP theta {np.round(th, 3)}")
    g4d = np.load(os.path.join(HERE, "g4_kernel_zoo.npz"))
    zoo = {
        "const_plus_matern15_iso": ConstantKernel(1.0) + Matern(0.4, nu=1.5),
        "const_times_rbf_ard": ConstantKernel(1.5) * RBF([0.3, 0.5, 0.7]),
        "matern05_ard_fixedconst": ConstantKernel(2.0, "fixed") * Matern([0.5, 0.4, 0.3], nu=0.5),
        "exp2_of_sum": Exponentiation(ConstantKernel(0.5) * Matern(0.6, nu=2.5) + RBF([1.0, 1.0, 1.0]), 2.0),
        "matern_inf_iso": ConstantKernel(1.0) * Matern(0.5, nu=np.inf),
        "product_of_stationary": ConstantKernel(1.0) * RBF(0.8) * Matern([0.9, 0.8, 0.7], nu=2.5),
    }
    for name, kern in zoo.items():
        gp = BayesGPR(kernel=kern, normalize_y=True, random_state=3, optimizer=None)
        SkoptGPR.fit(gp, g4d["X"], g4d["y_raw"])
        vals = [gp.log_marginal_likelihood(t, eval_gradient=True) for t in g4d[f"{name}__thetas"]]
        d[f"zoo_{name}__grad"] = np.array([v[1] for v in vals])
    np.savez_compressed(os.path.join(HERE, "g10_lml_gradients.npz"), **d)


if __name__ == "__main__":
    which = sys.argv[1:] or ["g1", "g2", "g3", "g4", "g5", "g6", "g7", "g8", "g9", "g10"]
    for name in which:
        globals()[name]()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
