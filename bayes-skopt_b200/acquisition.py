"""Acquisition functions and the theta-averaged candidate sweep -- the reference's
bask/acquisition.py surface (same class names, call protocol and error behaviour) on libbgp.

``evaluate_acquisitions`` batches what the reference does one theta at a time
(bask/acquisition.py:112-141): all sampled thetas are factorised in one launch, the
(theta x candidate) predictive moments come from one fused sweep, and every built-in
acquisition is a device epilogue followed by the finite-guarded mean over thetas.  User
subclasses of the three protocol classes keep working: they receive host copies of the
device-computed moments / draws."""
from abc import ABC, abstractmethod

import numpy as np
import torch
from sklearn.utils import check_random_state

from . import _lib
from ._engine import Engine, nvtx_range

__all__ = ["argmax_acquisition", "evaluate_acquisitions", "ExpectedImprovement", "TopTwoEI", "Expectation", "LCB",
           "MaxValueSearch", "ThompsonSampling", "VarianceReduction", "PVRS"]

_util_engine = None


def _utility_engine():
    """Engine used when an acquisition object is called directly on host arrays."""
    global _util_engine
    if _util_engine is None:
        _util_engine = Engine()
    return _util_engine


class Acquisition(ABC):
    @abstractmethod
    def __call__(self, *args, **kwargs):
        pass


class UncertaintyAcquisition(Acquisition, ABC):
    @abstractmethod
    def __call__(self, mu, std, *args, **kwargs):
        pass


class SampleAcquisition(Acquisition, ABC):
    @abstractmethod
    def __call__(self, gp_sample, *args, **kwargs):
        pass


class FullGPAcquisition(Acquisition, ABC):
    @abstractmethod
    def __call__(self, X, gp, *args, **kwargs):
        pass


def gumbel32_like_reference(n_min_samples):
    """The reference draws its Gumbel variates from the GLOBAL numpy RNG as float32 uniforms and
    takes -log(-log(u)) in float32 (bask/acquisition.py:253-257).  Done here on the host, with
    the same numpy calls, so a seeded run consumes the global stream identically."""
    with np.errstate(divide="ignore"):
        return -np.log(-np.log(np.random.rand(n_min_samples).astype(np.float32)))


class _DeviceUncertainty(UncertaintyAcquisition):
    """Built-in (mu, std) acquisitions: evaluated by libbgp's epilogue kernels."""
    kind = None

    def _params(self, kwargs):
        return float("nan"), None

    def device_eval(self, engine, mu_dev, sd_dev, kwargs, gumbel=None):
        p0, K = self._params(kwargs)
        g = None
        if K is not None:
            if gumbel is None:
                gumbel = np.stack([gumbel32_like_reference(K) for _ in range(mu_dev.shape[0])])
            g = engine.to_dev(gumbel, dtype=torch.float32)
        return engine.acq(self.kind, mu_dev, sd_dev, p0=p0, gumbel32=g)

    def __call__(self, mu, std, *args, **kwargs):
        e = _utility_engine()
        mu_d = e.to_dev(np.asarray(mu, dtype=np.float64)[None, :])
        sd_d = e.to_dev(np.asarray(std, dtype=np.float64)[None, :])
        _out, per, _skipped, _ = self.device_eval(e, mu_d, sd_d, kwargs)
        return e.to_host(per)[0]


class ExpectedImprovement(_DeviceUncertainty):
    """Expected improvement over ``y_opt`` (default: the minimum predicted mean of the candidate
    set); 0 where std == 0.  bask/acquisition.py:154-172."""
    kind = _lib.ACQ_EI

    def _params(self, kwargs):
        y_opt = kwargs.get("y_opt")
        return (float("nan") if y_opt is None else float(y_opt)), None


class TopTwoEI(ExpectedImprovement):
    """Expected improvement over the EI-maximising candidate.  bask/acquisition.py:175-194."""
    kind = _lib.ACQ_TTEI


class Expectation(_DeviceUncertainty):
    """Lowest predicted mean.  bask/acquisition.py:197-201."""
    kind = _lib.ACQ_MEAN


class LCB(_DeviceUncertainty):
    """alpha * std - mu (alpha=1.96; alpha="inf" returns std).  bask/acquisition.py:204-216."""
    kind = _lib.ACQ_LCB

    def _params(self, kwargs):
        alpha = kwargs.get("alpha", 1.96)
        return (float("inf") if alpha == "inf" else float(alpha)), None


class MaxValueSearch(_DeviceUncertainty):
    """Max-value entropy search (Wang & Jegelka 2017) with a Gumbel fit to the max-value
    distribution.  bask/acquisition.py:219-267."""
    kind = _lib.ACQ_MES

    def _params(self, kwargs):
        return float("nan"), int(kwargs.get("n_min_samples", 1000))


class ThompsonSampling(SampleAcquisition):
    """Negated joint posterior draw.  bask/acquisition.py:270-274."""

    def __call__(self, gp_sample, *args, **kwargs):
        return -gp_sample


def _variance_reduction_dev(gp, Xd, points_idx):
    """Schur-complement form of the reference's per-candidate (n+1)x(n+1) refactorisation
    (bask/acquisition.py:285-300, 328-339; identity checked in SURVEY.md row A12):
        covs[i] = sum_t |v_t|^2 + sum_t (k(t, x_i) - v_t . v_i)^2 / (k(x_i,x_i) + noise - |v_i|^2)
    with v = L^-1 k(X, .) at the current theta, noise ON.  points_idx=None -> all candidates."""
    e = gp._eng()
    th = e.to_dev(gp._theta_for_device()[None, :])
    f = gp._factor
    m = Xd.shape[0]
    _mu, sd, _, v = e.predict(f, Xd, thetas_dev=th, noise_off=False, y_mean=0.0, y_std=1.0, want_v=True)
    with torch.cuda.stream(e.stream):   # torch ops must be ordered with the library's kernels
        s_i = sd[0] * sd[0]
    if points_idx is None:
        cov = e.empty(m, m)
        _lib.check(e.lib.bgp_posterior_cov(e.h, th.data_ptr(), v[0].data_ptr(), Xd.data_ptr(), m, v.shape[2], 1,
                                           1.0, cov.data_ptr(), m, e._st), "bgp_posterior_cov")
        out = e.empty(m)
        _lib.check(e.lib.bgp_vr_combine(e.h, cov.data_ptr(), m, m, Xd.data_ptr(), th.data_ptr(), s_i.data_ptr(),
                                        out.data_ptr(), e._st), "bgp_vr_combine")
        e.launches += 2
        return out
    idx = e.to_dev(np.asarray(points_idx, dtype=np.int64), dtype=torch.int64)   # on the engine's stream
    with torch.cuda.stream(e.stream):
        vt = v[0].index_select(0, idx)[:, : e.n].contiguous()         # (R, n) whitened Thompson points
        Xt = Xd.index_select(0, idx).contiguous()
        assert vt[None].is_contiguous()
    R = vt.shape[0]
    _mu2, _sd2, dots, _ = e.predict(f, Xd, thetas_dev=th, noise_off=False, y_mean=0.0, y_std=1.0,
                                    zextra=vt[None])
    out = e.empty(m)
    _lib.check(e.lib.bgp_pvrs_combine(e.h, th.data_ptr(), Xt.data_ptr(), R, Xd.data_ptr(), m, dots.data_ptr(),
                                      vt.data_ptr(), s_i.data_ptr(), out.data_ptr(), e._st), "bgp_pvrs_combine")
    e.launches += 1
    return out


class VarianceReduction(FullGPAcquisition):
    """Trace of the explained covariance over ALL candidates after adding each candidate
    (active learning).  bask/acquisition.py:277-300."""

    def __call__(self, X, gp, *args, **kwargs):
        e = gp._eng()
        Xd = e.to_dev(np.asarray(X, dtype=np.float64))
        return e.to_host(_variance_reduction_dev(gp, Xd, None))


class PVRS(FullGPAcquisition):
    """Predictive variance reduction search (Nguyen et al. 2017): variance reduction at the
    minimisers of ``n_thompson`` joint posterior draws.  bask/acquisition.py:303-339."""

    def __call__(self, X, gp, *args, n_thompson=10, random_state=None, thompson_idx=None, **kwargs):
        e = gp._eng()
        X = np.asarray(X, dtype=np.float64)
        if thompson_idx is None:
            thompson_sample = gp.sample_y(X, sample_mean=True, n_samples=n_thompson, random_state=random_state)
            thompson_idx = np.argmin(thompson_sample, axis=0)
        Xd = e.to_dev(X)
        return e.to_host(_variance_reduction_dev(gp, Xd, thompson_idx))


def evaluate_acquisitions(X, gpr, acquisition_functions=None, n_samples=10, progress=False,
                          random_state=None, process_group=None, **kwargs):
    with nvtx_range("bgp.evaluate_acquisitions"):
        return _evaluate_acquisitions(X, gpr, acquisition_functions, n_samples, progress, random_state,
                                      process_group, **kwargs)


evaluate_acquisitions.__doc__ = """Evaluates acquisition functions on candidate points, averaged over ``n_samples`` draws from the
hyper-posterior chain (bask/acquisition.py:48-147); see ``_evaluate_acquisitions``."""


def _evaluate_acquisitions(X, gpr, acquisition_functions=None, n_samples=10, progress=False,
                           random_state=None, process_group=None, **kwargs):
    """Evaluates acquisition functions on candidate points, averaged over ``n_samples`` draws
    from the hyper-posterior chain.  Same arguments, RNG consumption and output as
    bask/acquisition.py:48-147; returns ``(len(acquisition_functions), len(X))`` float64.

    ``process_group`` (additive): a torch.distributed group with one rank per GPU.  Every rank
    passes identical arguments; the built-in (mu, std) acquisitions are then swept over
    candidates sharded across the ranks (bask_b200/distributed.py) and every rank receives the
    full result.  Full-GP and sample acquisitions are computed redundantly (replicas only)."""
    X = np.asarray(X, dtype=np.float64)
    n_cand_points = len(X)
    n_acqs = len(acquisition_functions)
    acq_output = np.zeros((n_acqs, n_cand_points))
    random_state = check_random_state(random_state)
    trace_sample_i = random_state.choice(gpr._chain_len(), replace=False, size=n_samples)
    e = gpr._eng()
    for i_acq, acq in enumerate(acquisition_functions):
        if isinstance(acq, FullGPAcquisition):
            out = acq(X, gpr, random_state=random_state, **kwargs)
            if np.all(np.isfinite(out)):
                acq_output[i_acq] = out
    S = len(trace_sample_i)
    has_unc = any(isinstance(a, UncertaintyAcquisition) for a in acquisition_functions)
    has_smp = any(isinstance(a, SampleAcquisition) for a in acquisition_functions)
    if S == 0 or not (has_unc or has_smp):
        return acq_output
    sharded = process_group is not None and torch.distributed.get_world_size(process_group) > 1
    Xd = e.to_dev(X)
    y_mean = float(np.atleast_1d(gpr.y_train_mean_)[0])
    y_std = float(np.atleast_1d(gpr.y_train_std_)[0])
    mu = sd = pd_info = None

    def _check_pd():
        nonlocal pd_info
        if pd_info is not None:
            info, pd_info = e.to_host(pd_info), None
            if np.any(info != 0):
                raise np.linalg.LinAlgError(
                    "The kernel, %s, is not returning a positive definite matrix. Try gradually increasing "
                    "the 'alpha' parameter of your GaussianProcessRegressor estimator." % gpr.kernel_)

    all_builtin = all(type(a).__call__ is _DeviceUncertainty.__call__ for a in acquisition_functions
                      if isinstance(a, UncertaintyAcquisition))
    if sharded and not all_builtin:
        raise NotImplementedError("user-defined UncertaintyAcquisition classes are not supported with a "
                                  "process_group (they need the moments of all candidates on one rank)")
    if has_unc and not sharded:
        th = gpr._chain_rows_dev(trace_sample_i)
        f = e.factorize(th)
        # the positive-definiteness flags are read back after the sweep has been enqueued (see
        # _check_pd below): a host round trip here would leave the GPU idle
        pd_info = f.info
        mu, sd, _, _ = e.predict(f, Xd, noise_off=True, y_mean=y_mean, y_std=y_std)
    # host-RNG consumption in the reference's order: per theta, MES draws (global numpy RNG) in
    # acquisition order; the first SampleAcquisition triggers one sample_y (random_state)
    gumbels = {j: [] for j, a in enumerate(acquisition_functions) if isinstance(a, MaxValueSearch)}
    draws = []
    for _s in range(S):
        sample_drawn = False
        for j, acq in enumerate(acquisition_functions):
            if j in gumbels:
                gumbels[j].append(gumbel32_like_reference(acq._params(kwargs)[1]))
            elif isinstance(acq, SampleAcquisition) and not sample_drawn:
                sample_drawn = True
                ind = random_state.choice(gpr._chain_len(), size=1, replace=True)
                draws.append((int(ind[0]), random_state.standard_normal(size=(1, n_cand_points)).T))
    samples = None
    if has_smp:
        th_s = e.to_dev(gpr.chain_[[d[0] for d in draws]])
        f_s = e.factorize(th_s)
        if np.any(e.to_host(f_s.info) != 0):
            raise np.linalg.LinAlgError("The kernel is not returning a positive definite matrix.")
        eps = e.to_dev(np.stack([d[1] for d in draws]))
        out, _ = gpr._joint_draws_dev(Xd, th_s, f_s, eps, noise=False)
        samples = e.to_host(out)[:, :, 0]                       # (S, m)
    mu_h = sd_h = None
    if sharded and has_unc:
        from .distributed import DeviceBackend, ShardedSweep
        idx = [j for j, a in enumerate(acquisition_functions) if isinstance(a, _DeviceUncertainty)]
        spec = [(acquisition_functions[j].kind, acquisition_functions[j]._params(kwargs)[0]) for j in idx]
        gmb = {i: np.stack(gumbels[j]) for i, j in enumerate(idx) if j in gumbels}
        vals = ShardedSweep(DeviceBackend(gpr), process_group).evaluate(X, gpr.chain_[trace_sample_i], spec, gmb)
        for i, j in enumerate(idx):
            acq_output[j] += vals[i]
    for j, acq in enumerate(acquisition_functions):
        if sharded and isinstance(acq, UncertaintyAcquisition):
            continue
        if isinstance(acq, _DeviceUncertainty) and type(acq).__call__ is _DeviceUncertainty.__call__:
            g = np.stack(gumbels[j]) if j in gumbels else None
            out, _per, _skipped, _ = acq.device_eval(e, mu, sd, kwargs, gumbel=g)
            gpr._materialize()   # a deferred sample() read-back overlaps the sweep just enqueued
            res = e.to_host(out)
            _check_pd()
            acq_output[j] += res
        elif isinstance(acq, UncertaintyAcquisition):
            if mu_h is None:
                mu_h, sd_h = e.to_host(mu), e.to_host(sd)
                _check_pd()
            for s in range(S):
                tmp = acq(mu_h[s], sd_h[s], **kwargs)
                if np.all(np.isfinite(tmp)):
                    acq_output[j] += tmp / n_samples
        elif isinstance(acq, SampleAcquisition):
            for s in range(S):
                tmp = acq(samples[s], **kwargs)
                if np.all(np.isfinite(tmp)):
                    acq_output[j] += tmp / n_samples
    _check_pd()
    return acq_output


def argmax_acquisition(X, gpr, acq, n_samples=10, random_state=None, process_group=None, **kwargs):
    """Index of the candidate that maximises ONE acquisition function -- the tail of ``Optimizer.tell``
    (bask/optimizer.py:365-380: evaluate_acquisitions(...).flatten() followed by np.argmax).

    For the built-in (mu, std) acquisitions on one GPU nothing but the index leaves the device: thetas are
    picked and MaxValueSearch's Gumbel variates drawn on the host in the reference's order, then
    factorise -> sweep -> acquisition epilogue -> finite-guarded theta-mean -> ``bgp_argmax`` (numpy's
    first-maximum tie rule) run back to back on the engine's stream and 8 bytes come back.  Everything else
    (full-GP and sample acquisitions, user-defined classes, a process group, n_samples == 0) goes through
    ``evaluate_acquisitions`` and ``np.argmax`` -- same values, same index."""
    builtin = isinstance(acq, _DeviceUncertainty) and type(acq).__call__ is _DeviceUncertainty.__call__
    sharded = process_group is not None and torch.distributed.get_world_size(process_group) > 1
    if not builtin or sharded or n_samples <= 0:
        vals = evaluate_acquisitions(X, gpr, (acq,), n_samples=n_samples, progress=False, random_state=random_state,
                                     process_group=process_group, **kwargs)
        return int(np.argmax(vals.flatten()))
    X = np.asarray(X, dtype=np.float64)
    random_state = check_random_state(random_state)
    picks = random_state.choice(gpr._chain_len(), replace=False, size=n_samples)
    e = gpr._eng()
    gumbel = None
    if isinstance(acq, MaxValueSearch):
        gumbel = np.stack([gumbel32_like_reference(acq._params(kwargs)[1]) for _ in picks])
    with nvtx_range("bgp.argmax_acquisition"):
        f = e.factorize(gpr._chain_rows_dev(picks))
        mu, sd, _, _ = e.predict(f, e.to_dev(X), noise_off=True, y_mean=float(np.atleast_1d(gpr.y_train_mean_)[0]),
                                 y_std=float(np.atleast_1d(gpr.y_train_std_)[0]))
        out, _per, _skipped, _ = acq.device_eval(e, mu, sd, kwargs, gumbel=gumbel)
        idx = e.argmax(out)
        gpr._materialize()   # a deferred sample() read-back overlaps the sweep just enqueued
        e.sync()
    if np.any(f.info.cpu().numpy() != 0):
        raise np.linalg.LinAlgError(
            "The kernel, %s, is not returning a positive definite matrix. Try gradually increasing "
            "the 'alpha' parameter of your GaussianProcessRegressor estimator." % gpr.kernel_)
    return int(idx.cpu().numpy()[0])
