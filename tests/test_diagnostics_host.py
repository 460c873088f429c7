"""CPU: host-side pieces of the optimiser diagnostics: expected_minimum and hdi.
The reference takes both from third-party packages (skopt.utils.expected_minimum, arviz.hdi)."""
import numpy as np
import pytest

import bask_b200.space as S


class _Bowl:
    def predict(self, X):
        return np.array([((X[0, 0] - 0.3) ** 2 + (X[0, 1] - 0.6) ** 2)])


def test_expected_minimum_finds_the_surrogate_minimum():
    sp = S.normalize_dimensions([(-2.0, 3.0), (0.0, 10.0)])
    res = S.create_result([[0.0, 1.0], [1.0, 5.0]], [1.0, 0.5], sp, np.random.RandomState(0), models=[_Bowl()])
    x, fun = S.expected_minimum(res, n_random_starts=5, random_state=0)
    np.testing.assert_allclose(x, [-2.0 + 0.3 * 5.0, 0.6 * 10.0], atol=1e-4)     # minimum in the original space
    assert fun < 1e-8
    # same oracle restatement (App. A3 of SURVEY.md): identical starts, identical answer
    from oracle import skopt_port
    sp2 = skopt_port.normalize_dimensions([(-2.0, 3.0), (0.0, 10.0)])
    res2 = skopt_port.create_result([[0.0, 1.0], [1.0, 5.0]], [1.0, 0.5], sp2, np.random.RandomState(0),
                                    models=[_Bowl()])
    x2, fun2 = skopt_port.expected_minimum(res2, n_random_starts=5, random_state=0)
    np.testing.assert_allclose(x, x2, rtol=1e-12)
    np.testing.assert_allclose(fun, fun2, atol=1e-15)


def test_hdi_unimodal_and_multimodal():
    rs = np.random.RandomState(0)
    x = rs.randn(20000)
    lo, hi = S.hdi(x, 0.95)
    assert abs(lo + 1.96) < 0.08 and abs(hi - 1.96) < 0.08
    # shortest interval: no other window with the same number of samples is shorter
    xs = np.sort(x)
    inc = int(np.floor(0.95 * len(xs)))
    assert hi - lo <= np.min(xs[inc:] - xs[: len(xs) - inc]) + 1e-12
    two = np.concatenate([rs.randn(5000) * 0.1 - 2, rs.randn(5000) * 0.1 + 2])
    modes = S.hdi(two, 0.9, multimodal=True)
    assert modes.shape == (2, 2) and modes[0, 1] < 0 < modes[1, 0]
    one = S.hdi(x, 0.95, multimodal=True)
    assert one.shape == (1, 2)
    assert S.hdi(np.array([0.5]), 0.95).tolist() == [0.5, 0.5]


def test_sb_sequence_matches_reference():
    """Steinerberger initial design: same points as the reference's sb_sequence for
    the same RandomState (checked against the unmodified reference when it is available, i.e. in
    the build container; structural checks everywhere)."""
    import os

    import bask_b200
    pts = bask_b200.sb_sequence(6, 2, random_state=3, restarts=5)
    assert pts.shape == (6, 2) and np.all((pts >= 0) & (pts <= 1))
    d = np.abs(pts[:, None, :] - pts[None, :, :]).sum(-1)
    assert np.min(d[np.triu_indices(6, 1)]) > 0.05                 # well separated
    more = bask_b200.sb_sequence(8, 2, existing_points=pts, random_state=4, restarts=5)
    np.testing.assert_array_equal(more[:6], pts)
    with np.testing.assert_raises(ValueError):
        bask_b200.sb_sequence(6, 2, existing_points=pts)
    if os.path.isdir("/root/reference/bask"):
        from oracle.ref_loader import load_reference
        load_reference()
        from bask.init import sb_sequence as ref_sb
        np.testing.assert_allclose(pts, ref_sb(6, 2, random_state=3, restarts=5), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(more, ref_sb(8, 2, existing_points=pts, random_state=4, restarts=5),
                                   rtol=1e-9, atol=1e-12)


def test_prior_table_with_input_warping():
    """Host logic of warp_inputs=True: the device prior table covers kernel theta ++ log a ++ log b
    (bask/bayesgpr.py:351-372); untyped callables end up in the host part with the right slices."""
    import bask_b200
    from bask_b200 import _lib
    from bask_b200.priors import NormalPrior
    from sklearn.gaussian_process.kernels import WhiteKernel
    gp = bask_b200.BayesGPR(kernel=bask_b200.construct_default_kernel([0, 1]), warp_inputs=True)
    gp._X_train = np.zeros((5, 2))
    k = bask_b200.construct_default_kernel([0, 1]) + WhiteKernel()
    table, host = gp._prior_table(bask_b200.guess_priors(k), None, 4)
    assert len(table) == 4 + 4 and host is None
    assert all(kind == _lib.PRIOR_NORMAL and tuple(par)[:2] == (0.0, 0.3) for kind, par in table[4:])
    # a joint callable warp prior f(a_log, b_log) is summed over the dimensions on the host
    table, host = gp._prior_table(bask_b200.guess_priors(k), lambda a, b: -(a * a + b * b), 4)
    assert all(kind == _lib.PRIOR_NONE for kind, _ in table[4:])
    th = np.array([0.0, 0.1, 0.2, 0.3, 1.0, 2.0, 3.0, 4.0])     # a = (1, 2), b = (3, 4)
    np.testing.assert_allclose(host(th), -(1 + 9) - (4 + 16))
    # typed per-parameter pair
    table, host = gp._prior_table(bask_b200.guess_priors(k), (NormalPrior(0.0, 1.0), NormalPrior(1.0, 2.0)), 4)
    assert host is None and [tuple(p)[:2] for _, p in table[4:]] == [(0.0, 1.0)] * 2 + [(1.0, 2.0)] * 2
    # identity warp before any warpers exist; full theta rows get zeros appended
    gp.kernel_ = k
    assert gp._theta_for_device().shape == (4 + 4,) and np.all(gp._theta_for_device()[4:] == 0.0)
    np.testing.assert_array_equal(gp.warp(np.full((3, 2), 0.25)), np.full((3, 2), 0.25))
    gp.create_warpers(np.log([2.0, 1.0]), np.log([1.0, 1.0]))
    np.testing.assert_allclose(gp.warp(np.full((3, 2), 0.5))[:, 0], 0.25)          # Beta(2,1).cdf(x) = x^2
    np.testing.assert_allclose(gp.unwarp(gp.warp(np.full((3, 2), 0.3))), 0.3, rtol=1e-12)
    np.testing.assert_allclose(gp._theta_for_device()[4:], np.log([2.0, 1.0, 1.0, 1.0]))


def test_diagnostics_arithmetic_with_a_stub_gp():
    """The optimiser diagnostics on a stub GP whose joint draws are known: the probabilities and the
    gap must equal a direct re-computation from the same draws (bask/optimizer.py:505-525, 610-620)."""
    import bask_b200

    class StubGP:
        warp_inputs = False
        chain_ = None
        kernel_ = True

        def predict(self, X):
            return np.array([np.sum((np.asarray(X) - 0.4) ** 2)])

        def sample_y(self, X, n_samples=1, sample_mean=True, random_state=None):
            rs = np.random.RandomState(1234)
            base = np.sum((np.asarray(X) - 0.4) ** 2, axis=1)
            return base[:, None] + 0.05 * rs.randn(len(X), n_samples)

    opt = bask_b200.Optimizer([(0.0, 1.0), (0.0, 1.0)], n_initial_points=0, random_state=0)
    opt.gp = StubGP()
    opt.Xi = [[0.1, 0.2], [0.5, 0.4], [0.9, 0.9]]
    opt.yi = [0.13, 0.01, 0.5]
    kw = dict(n_space_samples=60, n_gp_samples=40, n_random_starts=2, random_state=7)
    ps = opt.probability_of_optimality([0.0, 0.5, 1.0, 3.0], **kw)
    # direct re-computation
    x0 = opt._expected_optimum(2, 7)
    pts = [x0] + opt.space.rvs(n_samples=60, random_state=7)
    draws = StubGP().sample_y(opt.space.transform(pts), n_samples=40)
    std = np.std(draws, axis=0)
    want = [float((((draws[0][None, :] - draws) / std - eps).max(axis=0) < 0.0).mean()) for eps in (0.0, 0.5, 1.0, 3.0)]
    np.testing.assert_allclose(ps, want)
    assert ps == sorted(ps) and opt.probability_of_optimality(3.0, **kw) == ps[-1]
    raw = opt.probability_of_optimality(0.02, normalized_scores=False, **kw)
    assert raw == float(((draws[0][None, :] - draws - 0.02).max(axis=0) < 0.0).mean())
    gap = opt.expected_optimality_gap(n_probabilities=15, n_space_samples=60, n_gp_samples=40, n_random_starts=2,
                                      random_state=3)
    assert np.isfinite(gap) and 0.0 <= gap <= max(opt.yi) - min(opt.yi)


def test_bayes_search_cv_search_space_forms():
    """BayesSearchCV (bask/searchcv.py) host logic: the three accepted forms of search_spaces, sorted-name
    dimension order, scikit-learn estimator protocol (get_params / clone).  No device work."""
    from sklearn.base import clone
    from sklearn.linear_model import Ridge

    from bask_b200.searchcv import BayesSearchCV, dimensions_aslist, point_asdict
    one = {"alpha": (1e-3, 1e2, "log-uniform"), "fit_intercept": [True, False]}
    s = BayesSearchCV(Ridge(), one, n_iter=7)
    assert s.total_iterations == 7 and clone(s).n_iter == 7
    assert BayesSearchCV(Ridge(), [one, one], n_iter=4).total_iterations == 8
    assert BayesSearchCV(Ridge(), [(one, 3), (one, 2)]).total_iterations == 5
    assert dimensions_aslist(one) == [one["alpha"], one["fit_intercept"]]
    assert point_asdict(one, [0.5, False]) == {"alpha": 0.5, "fit_intercept": False}
    for bad in ([], [(one, 0)], [{}], 5):
        with pytest.raises((TypeError, ValueError)):
            BayesSearchCV(Ridge(), bad).total_iterations  # noqa: B018
