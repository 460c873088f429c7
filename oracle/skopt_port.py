"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the scikit-optimize 0.10.2 pieces
the reference's hot path calls (not part of the product).

scikit-optimize 0.10.2 is pinned by the reference (uv.lock:2455-2456), is not
vendored under /root/reference and is not installed in this image.  This module
restates the published behaviour of

* ``skopt.learning.gaussian_process.gpr.GaussianProcessRegressor.{__init__,fit,predict}``
  and ``_param_for_white_kernel_in_Sum`` -- call sites bask/bayesgpr.py:9-11,165-174,
  328-333,607,633,675,709;
* ``skopt.space.{Real,Integer,Categorical,Space}`` with the "normalize" transform and
  ``skopt.utils.{normalize_dimensions,create_result,expected_minimum,is_listlike,
  is_2Dlistlike}`` -- call sites bask/optimizer.py:7-13,144,149,210,218-221,359-363,
  374-380;

on top of the scikit-learn that IS installed (1.9.0; the reference pins 1.7.2, the GP
code paths used are numerically identical).  Pinned by tests/test_oracle_ref_golden.py,
which runs the unmodified reference sources (and the reference's own golden values) on
top of it in the build container.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this.
"""
import numbers
import warnings

import numpy as np
from scipy.linalg import cho_solve, solve_triangular
from scipy.optimize import OptimizeResult
from scipy.optimize import minimize as _sp_minimize
from scipy.stats import uniform as _uniform
from sklearn.gaussian_process import GaussianProcessRegressor as _SkGPR
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Sum, WhiteKernel
from sklearn.utils import check_array, check_random_state


# --------------------------------------------------------------------------- gpr
def _param_for_white_kernel_in_Sum(kernel, kernel_str=""):
    """Locate a WhiteKernel that is a direct child of a (nested) Sum.

    Returns ``(found, "k1__k2"-style parameter name)``.
    """
    if kernel_str != "":
        kernel_str = kernel_str + "__"
    if isinstance(kernel, Sum):
        for param, child in kernel.get_params(deep=False).items():
            if isinstance(child, WhiteKernel):
                return True, kernel_str + param
            present, child_str = _param_for_white_kernel_in_Sum(child, kernel_str + param)
            if present:
                return True, child_str
    return False, "_"


class GaussianProcessRegressor(_SkGPR):
    """skopt's GPR: sklearn's GPR + an automatically added WhiteKernel whose fitted level
    is moved to ``noise_`` (and zeroed in ``kernel_``), an explicit ``K_inv_``, and a
    ``predict`` whose variance is ``diag(K**) - diag(K* K^-1 K*^T)`` through ``K_inv_``."""

    def __init__(self, kernel=None, alpha=1e-10, optimizer="fmin_l_bfgs_b",
                 n_restarts_optimizer=0, normalize_y=False, copy_X_train=True,
                 random_state=None, noise=None):
        self.noise = noise
        super().__init__(kernel=kernel, alpha=alpha, optimizer=optimizer,
                         n_restarts_optimizer=n_restarts_optimizer,
                         normalize_y=normalize_y, copy_X_train=copy_X_train,
                         random_state=random_state)

    def fit(self, X, y):
        if self.kernel is None:
            self.kernel = ConstantKernel(1.0, constant_value_bounds="fixed") * RBF(
                1.0, length_scale_bounds="fixed")
        if self.noise == "gaussian":
            self.kernel = self.kernel + WhiteKernel()
        elif self.noise:
            self.kernel = self.kernel + WhiteKernel(
                noise_level=self.noise, noise_level_bounds="fixed")
        super().fit(X, y)
        self.noise_ = None
        if self.noise:
            if isinstance(self.kernel_, WhiteKernel):
                self.kernel_.set_params(noise_level=0.0)
            else:
                present, name = _param_for_white_kernel_in_Sum(self.kernel_)
                if present:
                    self.noise_ = self.kernel_.get_params()[name].noise_level
                    self.kernel_.set_params(**{name: WhiteKernel(noise_level=0.0)})
        L_inv = solve_triangular(self.L_.T, np.eye(self.L_.shape[0]))
        self.K_inv_ = L_inv.dot(L_inv.T)
        self.y_train_std_ = self._y_train_std
        self.y_train_mean_ = self._y_train_mean
        return self

    def predict(self, X, return_std=False, return_cov=False, return_mean_grad=False,
                return_std_grad=False):
        if return_std and return_cov:
            raise RuntimeError("Not returning standard deviation of predictions when "
                               "returning full covariance.")
        if return_mean_grad or return_std_grad:
            raise NotImplementedError("gradient outputs are off the restated path")
        X = check_array(X)
        if not hasattr(self, "X_train_") or self.X_train_ is None:
            y_mean = np.zeros(X.shape[0])
            if return_cov:
                return y_mean, self.kernel(X)
            if return_std:
                return y_mean, np.sqrt(self.kernel.diag(X))
            return y_mean
        K_trans = self.kernel_(X, self.X_train_)
        y_mean = self.y_train_std_ * K_trans.dot(self.alpha_) + self.y_train_mean_
        if return_cov:
            v = cho_solve((self.L_, True), K_trans.T)
            y_cov = self.kernel_(X) - K_trans.dot(v)
            return y_mean, y_cov * self.y_train_std_ ** 2
        if return_std:
            y_var = self.kernel_.diag(X)
            y_var -= np.einsum("ki,kj,ij->k", K_trans, K_trans, self.K_inv_)
            negative = y_var < 0
            if np.any(negative):
                warnings.warn("Predicted variances smaller than 0. Setting those "
                              "variances to 0.")
                y_var[negative] = 0.0
            return y_mean, np.sqrt(y_var * self.y_train_std_ ** 2)
        return y_mean


# ------------------------------------------------------------------------- space
class Dimension:
    prior = None
    name = None
    transformed_size = 1

    def rvs(self, n_samples=1, random_state=None):
        rng = check_random_state(random_state)
        return self.inverse_transform(self._rvs.rvs(size=n_samples, random_state=rng))

    @property
    def transformed_bounds(self):
        return (0.0, 1.0)


class Real(Dimension):
    """Real dimension under the "normalize" transform: (x-low)/(high-low), with a
    log10 stage first when ``prior == "log-uniform"``."""

    def __init__(self, low, high, prior="uniform", base=10, transform="normalize",
                 name=None, dtype=float):
        if high <= low:
            raise ValueError(f"the lower bound {low} has to be less than the upper bound {high}")
        self.low, self.high = float(low), float(high)
        self.prior, self.base, self.name = prior, base, name
        self._rvs = _uniform(0.0, np.nextafter(1.0, 2.0))

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


class Integer(Dimension):
    def __init__(self, low, high, prior="uniform", base=10, transform="normalize",
                 name=None, dtype=np.int64):
        if high <= low:
            raise ValueError(f"the lower bound {low} has to be less than the upper bound {high}")
        self.low, self.high = int(low), int(high)
        self.prior, self.base, self.name = prior, base, name
        self._rvs = _uniform(0.0, np.nextafter(1.0, 2.0))

    def transform(self, X):
        X = np.asarray(X, dtype=float)
        return (X - self.low) / (self.high - self.low)

    def inverse_transform(self, Xt):
        Xt = np.asarray(Xt, dtype=float)
        x = np.round(Xt * (self.high - self.low) + self.low)
        return np.clip(x, self.low, self.high).astype(np.int64).tolist()

    @property
    def bounds(self):
        return (self.low, self.high)


class Categorical(Dimension):
    """Label-encode then normalise to [0, 1] (skopt transform="normalize")."""

    def __init__(self, categories, prior=None, transform="normalize", name=None):
        self.categories = tuple(categories)
        self.name = name
        self.prior = prior
        self._rvs = _uniform(0.0, np.nextafter(1.0, 2.0))

    def transform(self, X):
        idx = np.array([self.categories.index(x) for x in X], dtype=float)
        return idx / max(len(self.categories) - 1, 1)

    def inverse_transform(self, Xt):
        Xt = np.asarray(Xt, dtype=float)
        idx = np.clip(np.round(Xt * max(len(self.categories) - 1, 1)), 0,
                      len(self.categories) - 1).astype(int)
        return [self.categories[i] for i in idx]

    @property
    def bounds(self):
        return self.categories


def _as_dimension(d):
    if isinstance(d, Dimension):
        return d
    if isinstance(d, (list, tuple)):
        if len(d) == 2 and all(isinstance(v, numbers.Integral) and not isinstance(v, bool)
                               for v in d):
            return Integer(*d)
        if len(d) == 2 and all(isinstance(v, numbers.Real) and not isinstance(v, bool)
                               for v in d):
            return Real(*d)
        if len(d) == 3 and isinstance(d[2], str) and all(
                isinstance(v, numbers.Real) for v in d[:2]):
            return Real(d[0], d[1], prior=d[2])
        return Categorical(d)
    raise ValueError(f"Invalid dimension {d!r}")


class Space:
    def __init__(self, dimensions):
        self.dimensions = [_as_dimension(d) for d in dimensions]

    @property
    def n_dims(self):
        return len(self.dimensions)

    @property
    def transformed_n_dims(self):
        return sum(d.transformed_size for d in self.dimensions)

    @property
    def is_partly_categorical(self):
        return any(isinstance(d, Categorical) for d in self.dimensions)

    @property
    def bounds(self):
        return [d.bounds for d in self.dimensions]

    @property
    def transformed_bounds(self):
        return [d.transformed_bounds for d in self.dimensions]

    def rvs(self, n_samples=1, random_state=None):
        """Dimension-by-dimension draws (column-major RNG consumption), rows out."""
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


# ------------------------------------------------------------------------- utils
def is_listlike(x):
    return isinstance(x, (list, tuple))


def is_2Dlistlike(x):
    return np.all([is_listlike(xi) for xi in x])


def normalize_dimensions(dimensions):
    return Space(dimensions)


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
    def func(x):
        reg = res.models[-1]
        xt = res.space.transform(x.reshape(1, -1))
        return reg.predict(xt.reshape(1, -1))[0]

    xs = [res.x]
    if n_random_starts > 0:
        xs.extend(res.space.rvs(n_random_starts, random_state=random_state))
    best_x, best_fun = None, np.inf
    for x0 in xs:
        r = _sp_minimize(func, x0=x0, bounds=res.space.bounds)
        if r.fun < best_fun:
            best_x, best_fun = r.x, r.fun
    return [v for v in best_x], best_fun


def bench1(x):
    return x[0] ** 2
