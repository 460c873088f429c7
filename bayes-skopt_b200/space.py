"""Search-space transforms used by Optimizer: a minimal stand-in for the parts of
``skopt.space`` / ``skopt.utils`` the reference calls (bask/optimizer.py:7-13,144,359-380;
skopt 0.10.2 is not vendored in the reference and is not a dependency here).  Every dimension
uses skopt's "normalize" transform, i.e. the GP always sees inputs in [0, 1]^d."""
import numbers

import numpy as np
from scipy.optimize import OptimizeResult
from sklearn.utils import check_random_state

__all__ = ["Real", "Integer", "Categorical", "Space", "normalize_dimensions", "create_result", "expected_minimum", "hdi",
           "is_listlike", "is_2Dlistlike"]

_ONE_PLUS = np.nextafter(1.0, 2.0)


class Dimension:
    name = None
    prior = None
    transformed_size = 1

    def rvs(self, n_samples=1, random_state=None):
        rng = check_random_state(random_state)
        # scipy.stats.uniform(0, nextafter(1, 2)).rvs == loc + scale * random_sample
        return self.inverse_transform(rng.uniform(size=n_samples) * _ONE_PLUS)

    @property
    def transformed_bounds(self):
        return (0.0, 1.0)


class Real(Dimension):
    def __init__(self, low, high, prior="uniform", base=10, transform="normalize", name=None, dtype=float):
        if high <= low:
            raise ValueError(f"the lower bound {low} has to be less than the upper bound {high}")
        if prior not in ("uniform", "log-uniform"):
            raise ValueError(f"prior should be 'uniform' or 'log-uniform', got {prior}")
        self.low, self.high, self.prior, self.base, self.name = float(low), float(high), prior, base, name

    def _fwd(self, x):
        return np.log10(x) / np.log10(self.base) if self.prior == "log-uniform" else x

    def transform(self, X):
        X = np.asarray(X, dtype=float)
        lo, hi = self._fwd(self.low), self._fwd(self.high)
        return (self._fwd(X) - lo) / (hi - lo)

    def inverse_transform(self, Xt):
        Xt = np.asarray(Xt, dtype=float)
        lo, hi = self._fwd(self.low), self._fwd(self.high)
        x = Xt * (hi - lo) + lo
        if self.prior == "log-uniform":
            x = self.base ** x
        return np.clip(x, self.low, self.high).astype(float).tolist()

    @property
    def bounds(self):
        return (self.low, self.high)

    def __repr__(self):
        return f"Real(low={self.low}, high={self.high}, prior='{self.prior}', transform='normalize')"


class Integer(Dimension):
    def __init__(self, low, high, prior="uniform", base=10, transform="normalize", name=None, dtype=np.int64):
        if high <= low:
            raise ValueError(f"the lower bound {low} has to be less than the upper bound {high}")
        self.low, self.high, self.prior, self.name = int(low), int(high), prior, name

    def transform(self, X):
        return (np.asarray(X, dtype=float) - self.low) / (self.high - self.low)

    def inverse_transform(self, Xt):
        x = np.round(np.asarray(Xt, dtype=float) * (self.high - self.low) + self.low)
        return np.clip(x, self.low, self.high).astype(np.int64).tolist()

    @property
    def bounds(self):
        return (self.low, self.high)

    def __repr__(self):
        return f"Integer(low={self.low}, high={self.high}, transform='normalize')"


class Categorical(Dimension):
    def __init__(self, categories, prior=None, transform="normalize", name=None):
        self.categories, self.prior, self.name = tuple(categories), prior, name

    def rvs(self, n_samples=1, random_state=None):
        """Categories drawn with probability ``prior`` (uniform when None) by inverting the discrete
        CDF on one uniform per sample -- what skopt's rv_discrete-based Categorical does.  (The
        inherited rounding of a uniform on [0, 1] would halve the mass of the first and last category.)"""
        rng = check_random_state(random_state)
        k = len(self.categories)
        p = np.full(k, 1.0 / k) if self.prior is None else np.asarray(self.prior, dtype=float)
        idx = np.minimum(np.searchsorted(np.cumsum(p), rng.uniform(size=n_samples), side="left"), k - 1)
        return [self.categories[i] for i in idx]

    def transform(self, X):
        idx = np.array([self.categories.index(x) for x in X], dtype=float)
        return idx / max(len(self.categories) - 1, 1)

    def inverse_transform(self, Xt):
        k = max(len(self.categories) - 1, 1)
        idx = np.clip(np.round(np.asarray(Xt, dtype=float) * k), 0, len(self.categories) - 1).astype(int)
        return [self.categories[i] for i in idx]

    @property
    def bounds(self):
        return self.categories

    def __repr__(self):
        return f"Categorical(categories={self.categories})"


def _as_dimension(d):
    if isinstance(d, Dimension):
        return d
    if isinstance(d, (list, tuple)):
        if len(d) == 2 and all(isinstance(v, numbers.Integral) and not isinstance(v, bool) for v in d):
            return Integer(*d)
        if len(d) == 2 and all(isinstance(v, numbers.Real) and not isinstance(v, bool) for v in d):
            return Real(*d)
        if len(d) == 3 and isinstance(d[2], str) and all(isinstance(v, numbers.Real) for v in d[:2]):
            return Real(d[0], d[1], prior=d[2])
        if len(d) >= 1:
            return Categorical(d)
    raise ValueError(f"Invalid dimension {d!r}. Read the documentation for supported types.")


class Space:
    def __init__(self, dimensions):
        self.dimensions = [_as_dimension(d) for d in dimensions]

    n_dims = property(lambda self: len(self.dimensions))
    transformed_n_dims = property(lambda self: sum(d.transformed_size for d in self.dimensions))
    is_partly_categorical = property(lambda self: any(isinstance(d, Categorical) for d in self.dimensions))
    bounds = property(lambda self: [d.bounds for d in self.dimensions])
    transformed_bounds = property(lambda self: [d.transformed_bounds for d in self.dimensions])

    def rvs(self, n_samples=1, random_state=None):
        """Draws dimension by dimension (so the RandomState is consumed column-major, as skopt
        does) and returns a list of points."""
        rng = check_random_state(random_state)
        cols = [d.rvs(n_samples=n_samples, random_state=rng) for d in self.dimensions]
        return [list(r) for r in zip(*cols)]

    def transform(self, X):
        cols = [np.asarray(d.transform([x[i] for x in X])).reshape((len(X), -1))
                for i, d in enumerate(self.dimensions)]
        return np.hstack(cols)

    def inverse_transform(self, Xt):
        Xt = np.asarray(Xt)
        cols = [d.inverse_transform(Xt[:, i]) for i, d in enumerate(self.dimensions)]
        return [list(r) for r in zip(*cols)]

    def __repr__(self):
        return "Space([" + ",\n       ".join(map(repr, self.dimensions)) + "])"


def normalize_dimensions(dimensions):
    return Space(dimensions)


def is_listlike(x):
    return isinstance(x, (list, tuple))


def is_2Dlistlike(x):
    return np.all([is_listlike(xi) for xi in x])


def create_result(Xi, yi, space=None, rng=None, specs=None, models=None):
    res = OptimizeResult()
    yi = np.asarray(yi)
    best = np.argmin(yi)
    res.x = Xi[best]
    res.fun = yi[best]
    res.func_vals = yi
    res.x_iters = Xi
    res.models = models
    res.space = space
    res.random_state = rng
    res.specs = specs
    return res


def expected_minimum(res, n_random_starts=20, random_state=None):
    """Minimum of the surrogate mean (skopt.utils.expected_minimum as used by
    bask/optimizer.py:490-504): L-BFGS-B on ``res.models[-1].predict`` in the original space, started
    from the best observed point and ``n_random_starts`` random points.  Returns ``(x, fun)``."""
    from scipy.optimize import minimize

    def func(x):
        xt = res.space.transform(np.asarray(x).reshape(1, -1))
        return float(res.models[-1].predict(xt.reshape(1, -1))[0])

    xs = [res.x]
    if n_random_starts > 0:
        xs.extend(res.space.rvs(n_random_starts, random_state=random_state))
    best_x, best_fun = None, np.inf
    for x0 in xs:
        r = minimize(func, x0=x0, bounds=res.space.bounds)
        if r.fun < best_fun:
            best_x, best_fun = r.x, r.fun
    return [v for v in best_x], best_fun


def hdi(x, hdi_prob=0.95, multimodal=False):
    """Highest density interval(s) of a 1-D sample (the role arviz.hdi plays in
    bask/optimizer.py:681-689).  Unimodal: the shortest interval holding ``hdi_prob`` of the
    sample, shape (2,).  Multimodal: the region where a Gaussian KDE exceeds the density level that
    encloses ``hdi_prob`` of its mass, as an (n_modes, 2) array."""
    x = np.sort(np.asarray(x, dtype=np.float64).ravel())
    n = len(x)
    if n == 0:
        raise ValueError("hdi needs at least one sample")
    if not multimodal or n < 3 or x[0] == x[-1]:
        inc = int(np.floor(hdi_prob * n))
        n_intervals = n - inc
        if n_intervals < 1 or inc < 1:
            return np.array([x[0], x[-1]])
        widths = x[inc:] - x[:n_intervals]
        i = int(np.argmin(widths))
        out = np.array([x[i], x[i + inc]])
        return out[None, :] if multimodal else out
    from scipy.stats import gaussian_kde
    grid = np.linspace(x[0], x[-1], 512)
    dens = gaussian_kde(x)(grid)
    dens = dens / dens.sum()
    order = np.argsort(dens)[::-1]
    k = int(np.searchsorted(np.cumsum(dens[order]), hdi_prob)) + 1
    inside = np.zeros(len(grid), dtype=bool)
    inside[order[:k]] = True
    edges = np.flatnonzero(np.diff(np.concatenate([[0], inside.astype(np.int8), [0]])))
    return np.array([[grid[a], grid[b - 1]] for a, b in zip(edges[::2], edges[1::2])])
