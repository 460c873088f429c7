"""Log-priors on log-space hyper-parameters.

Same call protocol as the reference (a prior is a callable ``f(theta_k) -> float``,
bask/bayesgpr.py:368-372) but *typed*, so that the closed forms the reference builds from
scipy.stats lambdas (bask/utils.py:95-120, bask/priors.py:7-57) can be evaluated inside the
CUDA log-posterior kernel.  Calling one of these objects on the host evaluates the same
closed form with plain ``math`` (used for user inspection and the host-stepped MCMC mode).
"""
import math

import numpy as np
from scipy.integrate import quad

from . import _lib

__all__ = ["make_roundflat", "RoundFlatPrior", "HalfNormalSqrtPrior", "InvGammaPrior",
           "NormalPrior", "as_device_priors"]


_NORM_CACHE = {}


def make_roundflat(lower_bound=0.1, upper_bound=0.6, lower_steepness=2.0, upper_steepness=8.0,
                   integration_bounds=(0.0, 10.0)):
    """Round-flat log-density on the ORIGINAL scale (bask/priors.py:7-57): roughly flat inside
    (lower_bound, upper_bound), normalised by quadrature over ``integration_bounds``."""
    def roundflat(x):
        return -2 * ((x / lower_bound) ** (-2 * lower_steepness)
                     + (x / upper_bound) ** (2 * upper_steepness))

    key = (lower_bound, upper_bound, lower_steepness, upper_steepness, tuple(integration_bounds))
    if key not in _NORM_CACHE:   # the reference re-integrates on every guess_priors call (~0.5 ms)
        with np.errstate(divide="ignore", over="ignore"):
            _NORM_CACHE[key] = quad(lambda x: np.exp(roundflat(x)), integration_bounds[0],
                                    integration_bounds[1])[0]
    value = _NORM_CACHE[key]

    def prior(x):
        return roundflat(x) - np.log(value)

    prior.log_norm = float(np.log(value))
    return prior


class _TypedPrior:
    kind = _lib.PRIOR_NONE

    def params(self):
        return ()


class HalfNormalSqrtPrior(_TypedPrior):
    """theta = log v with sqrt(v) ~ half-normal(scale): what guess_priors puts on the signal
    variance and the noise level (bask/utils.py:95-99)."""
    kind = _lib.PRIOR_HALFNORMAL_SQRT

    def __init__(self, scale=2.0):
        self.scale = float(scale)

    def params(self):
        return (self.scale,)

    def __call__(self, x):
        return (0.5 * math.log(2.0 / math.pi) - math.log(self.scale)
                - math.exp(x) / (2.0 * self.scale ** 2) + x / 2.0 - math.log(2.0))


class RoundFlatPrior(_TypedPrior):
    """theta = log l with l ~ round-flat(lo, hi) (bask/utils.py:113-120)."""
    kind = _lib.PRIOR_ROUNDFLAT

    def __init__(self, lower_bound=0.1, upper_bound=0.6, lower_steepness=2.0, upper_steepness=8.0,
                 integration_bounds=(0.0, 10.0)):
        self.lo, self.hi = float(lower_bound), float(upper_bound)
        self.slo, self.shi = float(lower_steepness), float(upper_steepness)
        self.log_norm = make_roundflat(lower_bound, upper_bound, lower_steepness, upper_steepness,
                                       integration_bounds).log_norm

    def params(self):
        return (self.lo, self.hi, self.slo, self.shi, self.log_norm)

    def __call__(self, x):
        return (-2.0 * (math.exp(-2.0 * self.slo * (x - math.log(self.lo)))
                        + math.exp(2.0 * self.shi * (x - math.log(self.hi)))) - self.log_norm + x)


class InvGammaPrior(_TypedPrior):
    """theta = log v with v ~ inverse-gamma(a, scale)."""
    kind = _lib.PRIOR_INVGAMMA

    def __init__(self, a, scale=1.0):
        self.a, self.scale = float(a), float(scale)

    def params(self):
        return (self.a, self.scale)

    def __call__(self, x):
        return (self.a * math.log(self.scale) - math.lgamma(self.a) - (self.a + 1.0) * x
                - self.scale * math.exp(-x) + x)


class NormalPrior(_TypedPrior):
    """theta ~ N(loc, scale) directly in log space."""
    kind = _lib.PRIOR_NORMAL

    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = float(loc), float(scale)

    def params(self):
        return (self.loc, self.scale)

    def __call__(self, x):
        t = (x - self.loc) / self.scale
        return -0.5 * t * t - math.log(self.scale) - 0.5 * math.log(2.0 * math.pi)


def as_device_priors(priors, n_theta):
    """Splits a reference-style ``priors`` argument into (device prior table, host part).

    Returns ``(table, host_fn)``: ``table`` is a list of n_theta (kind, params) entries the
    CUDA kernel sums; ``host_fn`` is None when everything is typed, else a callable
    ``theta(p,) -> float`` adding up the untyped Python callables (evaluated on the host by
    the host-stepped sampler -- only the prior, never the GP numerics)."""
    if priors is None:
        return [(_lib.PRIOR_NONE, ())] * n_theta, None
    if callable(priors) and not isinstance(priors, (list, tuple)):
        return [(_lib.PRIOR_NONE, ())] * n_theta, (lambda th, f=priors: float(f(th)))
    priors = list(priors)
    if len(priors) != n_theta:
        raise ValueError(f"zip() argument 2 is {'longer' if n_theta > len(priors) else 'shorter'} "
                         f"than argument 1: {len(priors)} priors for {n_theta} hyperparameters")
    table, host = [], []
    for k, pr in enumerate(priors):
        if isinstance(pr, _TypedPrior):
            table.append((pr.kind, pr.params()))
        else:
            table.append((_lib.PRIOR_NONE, ()))
            host.append((k, pr))
    if not host:
        return table, None

    def host_fn(th, host=host):
        return float(sum(f(th[k]) for k, f in host))

    return table, host_fn
