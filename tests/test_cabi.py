"""CPU: the C-ABI library loads and exports every symbol include/bgp.h declares; host-side
logic (kernel compilation, priors, space, error paths) works without a GPU."""
import os
import re

import numpy as np
import pytest

import bask_b200
from bask_b200 import _lib
from bask_b200._engine import compile_kernel, find_zeroable_white
from bask_b200.priors import (HalfNormalSqrtPrior, InvGammaPrior, NormalPrior, RoundFlatPrior,
                              as_device_priors, make_roundflat)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(REPO, "include", "bgp.h")).read()
    declared = set(re.findall(r"\b(bgp_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/bgp.h but not exported"
    for name in _lib.EXPORTED:
        assert name in declared, f"{name} is bound by ctypes but not declared in include/bgp.h"
    assert lib.bgp_version() >= 100


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.BgpError):
        bask_b200.BayesGPR(kernel=bask_b200.construct_default_kernel([0, 1])).fit(np.zeros((3, 2)), np.zeros(3))
    import ctypes as C
    h = C.c_void_p()
    assert _lib.load().bgp_create(C.byref(h), 0) != 0
    assert b"no CUDA device" in _lib.load().bgp_last_error()


def test_kernel_program_compilation():
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Exponentiation, Matern, WhiteKernel
    k = ConstantKernel(1.0, (0.1, 2.0)) * Matern([0.3] * 3, (0.2, 0.5), nu=2.5) + WhiteKernel()
    ops, fixed_ls, p = compile_kernel(k)
    assert [o.code for o in ops] == [_lib.OP_CONST, _lib.OP_MATERN52, _lib.OP_MUL, _lib.OP_WHITE, _lib.OP_ADD]
    assert [o.theta_idx for o in ops] == [0, 1, -1, 4, -1] and p == 5 == len(k.theta)
    assert ops[3].flags == _lib.FLAG_ZEROABLE_WHITE
    k2 = Exponentiation(ConstantKernel(constant_value_bounds="fixed") * Matern() + WhiteKernel()
                        + RBF(length_scale=(1.0, 1.0)), 2.0)
    ops2, _, p2 = compile_kernel(k2)
    assert p2 == len(k2.theta) == 4 and ops2[-1].code == _lib.OP_POW
    assert find_zeroable_white(k2.kernel)[1] == "k1__k2"
    with pytest.raises(NotImplementedError):
        compile_kernel(ConstantKernel() * Matern(nu=0.7))
    fx = ConstantKernel(2.0, "fixed") * Matern([0.5, 0.4], "fixed", nu=1.5)
    ops3, fl3, p3 = compile_kernel(fx)
    assert p3 == 0 and fl3 == [0.5, 0.4]


def test_typed_priors_match_reference_known_answers():
    """tests/test_utils.py:20-40 and tests/test_priors.py:8-11 of the reference."""
    from scipy.integrate import quad
    from scipy.stats import halfnorm, invgamma, norm
    assert abs(RoundFlatPrior()(-0.9) - -0.02116327824572739) < 1e-7
    assert abs(HalfNormalSqrtPrior(2.0)(-0.9) - -2.112906921232193) < 1e-12
    prior = make_roundflat()
    assert abs(quad(lambda x: np.exp(prior(x)), 0.0, 10.0)[0] - 1.0) < 1e-7
    for x in (-3.0, -0.5, 0.0, 1.3):
        ref = halfnorm(scale=1.0).logpdf(np.sqrt(np.exp(x))) + x / 2 - np.log(2)
        assert abs(HalfNormalSqrtPrior(1.0)(x) - ref) < 1e-12
        assert abs(InvGammaPrior(5.0, 1.0)(x) - (invgamma(a=5.0, scale=1.0).logpdf(np.exp(x)) + x)) < 1e-11
        assert abs(NormalPrior(0.0, 0.3)(x) - norm(0, 0.3).logpdf(x)) < 1e-12
    k = bask_b200.construct_default_kernel([0, 1])
    pr = bask_b200.guess_priors(k)
    assert len(pr) == 3
    table, host = as_device_priors(pr, 3)
    assert host is None
    assert [t[0] for t in table] == [_lib.PRIOR_HALFNORMAL_SQRT, _lib.PRIOR_ROUNDFLAT, _lib.PRIOR_ROUNDFLAT]
    table, host = as_device_priors([pr[0], lambda x: -x * x, pr[2]], 3)
    assert table[1][0] == _lib.PRIOR_NONE and host(np.array([0.0, 2.0, 0.0])) == -4.0
    with pytest.raises(ValueError):
        as_device_priors(pr, 4)


def test_space_and_optimizer_bookkeeping():
    from bask_b200.space import Space
    sp = Space([(-2.0, 2.0), (1, 5), ["a", "b", "c"], (1e-3, 1e1, "log-uniform")])
    pts = sp.rvs(5, random_state=0)
    Xt = sp.transform(pts)
    assert Xt.shape == (5, 4) and Xt.min() >= 0 and Xt.max() <= 1
    back = sp.inverse_transform(Xt)
    for a, b in zip(pts, back):
        assert a[1] == b[1] and a[2] == b[2] and abs(a[0] - b[0]) < 1e-12 and abs(a[3] / b[3] - 1) < 1e-9
    opt = bask_b200.Optimizer([(-2.0, 2.0)], n_initial_points=3, init_strategy="r2", random_state=0)
    x = opt.ask()
    assert not isinstance(x[0], list)
    opt.tell([x], [0.0])
    assert opt._n_initial_points == 2 and opt.gp.chain_ is None
    with pytest.raises(NotImplementedError):
        opt.ask(n_points=2)
    with pytest.raises(ValueError):
        opt.tell([0.1], 0.0, noise_vector=[0.1])
    with pytest.raises(ValueError):
        opt.tell([[0.1], [0.2]], [0.0, 1.0], noise_vector=[0.1])
    np.testing.assert_allclose(bask_b200.r2_sequence(3, 1)[:, 0],
                               (0.5 + np.arange(1, 4) / 1.6180339887498949) % 1)
    assert len(bask_b200.construct_default_kernel([0, 1]).theta) == 3
    med = bask_b200.geometric_median(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.0, 0.0]]))
    assert np.linalg.norm(med) < 1e-4


def test_estimator_argument_errors():
    gp = bask_b200.BayesGPR(kernel=bask_b200.construct_default_kernel([0]))
    with pytest.raises(ValueError):
        gp.sample()
    gw = bask_b200.BayesGPR(warp_inputs=True)      # input warping: identity until warpers exist
    assert gw.warp_inputs and gw.warp(np.ones((1, 1)))[0, 0] == 1.0
    assert gp.theta is None and gp.chain_ is None and gp.pos_ is None
    mu, sd = gp.predict(np.zeros((2, 1)), return_std=True)    # GP prior before any fit
    np.testing.assert_allclose(mu, 0.0)
    np.testing.assert_allclose(sd, 1.0)
