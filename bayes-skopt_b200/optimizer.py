"""Ask/tell driver with the reference's surface (``Optimizer.__init__/ask/tell/run`` and the optimality
diagnostics, bask/optimizer.py:120-689) on the B200 path.  ``ask`` does no maths: it hands out the
initial design (R2 or Steinerberger sequence, bask/init.py) and afterwards the point computed at the tail
of the last ``tell`` (hyper-posterior on device -> candidates -> device acquisition sweep -> device argmax).
Behaviour -- argument meaning, RNG consumption order, exception types -- follows the reference; the code
is organised around the device path (see ``Optimizer``)."""
import warnings

import numpy as np
from sklearn.utils import check_random_state

from . import acquisition
from .acquisition import evaluate_acquisitions
from .bayesgpr import BayesGPR
from .space import (create_result, expected_minimum, hdi, is_2Dlistlike, is_listlike,
                    normalize_dimensions)
from .utils import construct_default_kernel

__all__ = ["Optimizer", "r2_sequence"]

ACQUISITION_FUNC = {
    "ei": acquisition.ExpectedImprovement(),
    "lcb": acquisition.LCB(),
    "mean": acquisition.Expectation(),
    "mes": acquisition.MaxValueSearch(),
    "pvrs": acquisition.PVRS(),
    "ts": acquisition.ThompsonSampling(),
    "ttei": acquisition.TopTwoEI(),
    "vr": acquisition.VarianceReduction(),
}


def _phi(d, n_iter=10):
    if d == 1:
        return 1.61803398874989484820458683436563
    if d == 2:
        return 1.32471795724474602596090885447809
    x = 2.0
    for _ in range(n_iter):
        x = (1 + x) ** (1.0 / (d + 1.0))
    return x


def r2_sequence(n, d, seed=0.5):
    """Additive-recurrence low-discrepancy sequence (Roberts' R_d), the reference's
    ``init_strategy="r2"`` (bask/init.py:103-128)."""
    g = _phi(d)
    alpha = np.array([(1.0 / g) ** (j + 1) % 1 for j in range(d)])
    return (seed + alpha[None, :] * (np.arange(n)[:, None] + 1)) % 1


def _sb_energy(x, pts):
    """Steinerberger's energy of a trial point against the points placed so far:
    sum_j prod_k (1 - log(2 sin(pi |x_k - p_jk|))); +inf when the trial point sits on one of them."""
    diff = np.abs(np.asarray(x)[None, :] - pts)
    with np.errstate(divide="ignore", invalid="ignore"):
        terms = 1.0 - np.log(2.0 * np.sin(np.pi * diff))
    val = np.sum(np.prod(terms, axis=-1))
    return val if np.isfinite(val) else np.inf


def sb_sequence(n, d, existing_points=None, random_state=None, restarts=20):
    """Greedy low-discrepancy sequence of Steinerberger (2019) in [0, 1]^d, the reference's
    ``init_strategy="sb"`` (bask/init.py:26-100): every new point minimises the energy above,
    searched by L-BFGS-B from ``restarts`` uniform starting points.  Consumes ``random_state`` in
    the reference's order (one uniform(d) for the first point when nothing exists, then one
    uniform((restarts, d)) block per added point).  Host-only: it runs before any GP exists."""
    from scipy.optimize import minimize
    rng = check_random_state(random_state)
    if existing_points is None:
        pts = [rng.uniform(size=d)]
    else:
        pts = [np.asarray(p, dtype=np.float64) for p in existing_points]
        if len(pts) >= n:
            raise ValueError("No more points left to generate.")
    for _ in range(n - len(pts)):
        starts = rng.uniform(size=(restarts, d))
        best_val, best_pt = np.inf, starts[0]
        placed = np.array(pts)
        for x0 in starts:
            with np.errstate(invalid="ignore"):
                res = minimize(_sb_energy, x0=x0, bounds=[(0.0, 1.0)] * d, args=(placed,))
            if res.fun < best_val:
                best_val, best_pt = res.fun, res.x
        pts.append(best_pt)
    return np.array(pts)


class _InitialDesign:
    """Where the points come from while no surrogate exists yet (bask/optimizer.py:140-152, 197-219):
    "r2" -- a pre-computed additive-recurrence sequence, handed out back to front; "sb" -- Steinerberger's
    greedy sequence, extended by one point against everything told so far; anything else -- uniform draws."""

    def __init__(self, strategy, space, n_points, rng):
        self.strategy, self.space = strategy, space
        self._r2 = None
        self._rng = None
        if strategy == "r2":
            self._r2 = space.inverse_transform(r2_sequence(n=max(n_points, 1), d=space.n_dims))
        elif strategy == "sb":
            self._rng = np.random.RandomState(rng.randint(2 ** 31))   # its own stream, seeded once

    def point(self, n_remaining, told):
        if self.strategy == "r2":
            return self._r2[n_remaining - 1]
        if self.strategy == "sb":
            placed = self.space.transform(told) if len(told) > 0 else None
            seq = sb_sequence(n=len(told) + 1, d=self.space.transformed_n_dims, existing_points=placed,
                              random_state=self._rng.randint(2 ** 31))
            return self.space.inverse_transform(np.atleast_2d(seq[len(told)]))[0]
        return self.space.rvs()[0]


class _Observations:
    """The told points: inputs (original space), targets and per-point noise variances."""

    def __init__(self):
        self.X, self.y, self.noise = [], [], []

    def clear(self):
        del self.X[:], self.y[:], self.noise[:]

    def add(self, x, y, noise):
        """One observation (x a point, y a number) or a batch (x a list of points, y a list); returns how
        many were added.  Raises ValueError for mismatched shapes, like bask/optimizer.py:287-321."""
        if is_listlike(y) and is_2Dlistlike(x):
            if noise is None:
                noise = [0.0] * len(y)
            elif not is_listlike(noise) or len(noise) != len(y):
                raise ValueError("Vector of noise variances needs to be of equal length as `y`.")
            self.X.extend(x), self.y.extend(y), self.noise.extend(noise)
            return len(y)
        if is_listlike(x):
            if is_listlike(noise):
                raise ValueError("Vector of noise variances is a list, while tell only received one datapoint.")
            self.X.append(x), self.y.append(y), self.noise.append(0.0 if noise is None else noise)
            return 1
        raise ValueError(f"Type of arguments `x` ({type(x)}) and `y` ({type(y)}) not compatible.")


class Optimizer:
    """Stepwise Bayesian optimisation with a fully Bayesian GP -- the reference's ask/tell surface
    (bask/optimizer.py:35-445; parameters as documented there :35-119), organised around four steps:
    an initial-design provider, the observation store, ``_refit`` (hyper-posterior on device) and
    ``_propose`` (candidates -> device acquisition sweep -> device argmax)."""

    def __init__(self, dimensions, n_points=500, n_initial_points=10, init_strategy="sb", gp_kernel=None,
                 gp_kwargs=None, gp_priors=None, acq_func="pvrs", acq_func_kwargs=None, random_state=None,
                 **kwargs):
        self.rng = check_random_state(random_state)
        self.acq_func = acq_func if callable(acq_func) else ACQUISITION_FUNC[acq_func]
        self.acq_func_kwargs = dict(acq_func_kwargs or {})
        self.space = normalize_dimensions(dimensions)
        self.n_points = n_points
        self.init_strategy = init_strategy
        self.n_initial_points_ = n_initial_points          # as configured
        self._n_initial_points = n_initial_points          # still to be told before the first fit
        self._design = _InitialDesign(init_strategy, self.space, n_initial_points, self.rng)
        kernel = gp_kernel if gp_kernel is not None else \
            construct_default_kernel(list(range(self.space.transformed_n_dims)))
        self.gp = BayesGPR(kernel=kernel, random_state=self.rng.randint(0, np.iinfo(np.int32).max),
                           **(gp_kwargs or {}))
        self.gp_priors = gp_priors
        self._obs = _Observations()
        self._next_x = None

    # the reference exposes the told data as plain lists
    Xi = property(lambda self: self._obs.X, lambda self, v: setattr(self._obs, "X", list(v)))
    yi = property(lambda self: self._obs.y, lambda self, v: setattr(self._obs, "y", list(v)))
    noisei = property(lambda self: self._obs.noise, lambda self, v: setattr(self._obs, "noise", list(v)))

    def ask(self, n_points=1):
        """Next point to evaluate: from the initial design while it lasts, afterwards the maximiser computed
        at the tail of the last ``tell`` (bask/optimizer.py:177-226)."""
        if n_points > 1:
            raise NotImplementedError("Returning multiple points is not implemented yet.")
        if self._n_initial_points > 0:
            return self._design.point(self._n_initial_points, self._obs.X)
        if not self.gp.kernel_:
            raise RuntimeError("Initialization is finished, but no model has been fit.")
        return self._next_x

    def tell(self, x, y, noise_vector=None, fit=True, replace=False, n_samples=0, gp_samples=100,
             gp_burnin=10, progress=False):
        """Records observations; once the initial design is used up (and ``fit``) re-samples the hyper-
        posterior and computes the next proposal (bask/optimizer.py:228-380).  Returns an OptimizeResult."""
        if replace:
            self._obs.clear()
            self._n_initial_points = self.n_initial_points_
        self._n_initial_points -= self._obs.add(x, y, noise_vector)
        if fit and self._n_initial_points <= 0:
            self._refit(gp_samples, gp_burnin, progress, from_scratch=replace)
            self._next_x = self._propose(n_samples)
        return self.get_result()

    def _refit(self, gp_samples, gp_burnin, progress, from_scratch):
        """First call (or ``replace``): MAP start + MCMC (``fit``); later calls continue the ensemble from
        its last positions on the grown data set (``sample``) -- bask/optimizer.py:323-351."""
        if self.gp_priors is not None and len(self.gp_priors) != self.space.transformed_n_dims + 2:
            raise ValueError("The number of priors does not match the number of dimensions + 2.")
        step = self.gp.fit if (self.gp.pos_ is None or from_scratch) else self.gp.sample
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            step(self.space.transform(self._obs.X), self._obs.y, noise_vector=np.array(self._obs.noise),
                 priors=self.gp_priors, n_desired_samples=gp_samples, n_burnin=gp_burnin, progress=progress)

    def _candidates(self):
        """``n_points`` random candidates in the transformed space; with input warping they are uniform in
        the WARPED space and mapped back (bask/optimizer.py:353-363)."""
        if self.gp.warp_inputs:
            return self.gp.unwarp(self.rng.uniform(size=(self.n_points, self.space.transformed_n_dims)))
        return self.space.transform(self.space.rvs(n_samples=self.n_points, random_state=self.rng))

    def _propose(self, n_samples):
        """Candidate with the largest acquisition value, back in the original space.  For the built-in
        (mu, std) acquisitions values and argmax stay on the device; only the index comes back."""
        X = self._candidates()
        best = acquisition.argmax_acquisition(X, self.gp, self.acq_func, n_samples=n_samples,
                                              random_state=self.rng.randint(0, np.iinfo(np.int32).max),
                                              **self.acq_func_kwargs)
        return self.space.inverse_transform(X[best].reshape((1, -1)))[0]

    def get_result(self):
        return create_result(self._obs.X, self._obs.y, self.space, self.rng, models=[self.gp])

    def run(self, func, n_iter=1, replace=False, n_samples=5, gp_samples=100, gp_burnin=10):
        """``n_iter`` rounds of ask -> func -> tell; ``func`` returns a value or a (value, noise variance)
        pair (bask/optimizer.py:382-445)."""
        for it in range(n_iter):
            x = self.ask()
            out = func(x)
            val, noise = out if hasattr(out, "__len__") else (out, 0.0)
            self.tell(x, val, noise_vector=noise, n_samples=n_samples, gp_samples=gp_samples, gp_burnin=gp_burnin,
                      replace=replace and it == 0)
        return self.get_result()

    # ------------------------------------------------------------------ diagnostics
    def _expected_optimum(self, n_random_starts, random_state):
        """expected_minimum of the current surrogate; identical calls (same data, same integer
        seed -- what expected_optimality_gap issues dozens of times) are answered from a cache."""
        key = None
        if isinstance(random_state, (int, np.integer)):
            key = (len(self.Xi), getattr(self.gp, "chain_generation_", id(self.gp.chain_)), int(random_state), int(n_random_starts))
            if getattr(self, "_expected_optimum_cache", (None, None))[0] == key:
                return self._expected_optimum_cache[1]
        result = self.get_result()
        x = expected_minimum(result, random_state=random_state, n_random_starts=n_random_starts)[0]
        if key is not None:
            self._expected_optimum_cache = (key, x)
        return x

    def _optimality_margins(self, n_space_samples, n_gp_samples, n_random_starts, use_mean_gp, normalized_scores,
                            random_state):
        """Per joint posterior draw: by how much the best of ``n_space_samples`` random points beats the
        expected optimum (row 0 of the draw), optionally in units of the draw's standard deviation."""
        pts = [self._expected_optimum(n_random_starts, random_state)]
        pts += self.space.rvs(n_samples=n_space_samples, random_state=random_state)
        draws = self.gp.sample_y(self.space.transform(pts), n_samples=n_gp_samples, sample_mean=use_mean_gp,
                                 random_state=random_state)                       # (1 + n_space, n_gp)
        gain = draws[0][None, :] - draws                                           # > 0: the point is better
        if normalized_scores:
            gain = gain / np.std(draws, axis=0)
        return gain.max(axis=0)

    def probability_of_optimality(self, threshold, n_space_samples=500, n_gp_samples=200, n_random_starts=100,
                                  use_mean_gp=True, normalized_scores=True, random_state=None):
        """Probability that the current expected optimum cannot be improved by more than
        ``threshold`` (a float, or a list -> list of probabilities): the share of joint posterior
        draws in which no random point beats it by that margin (bask/optimizer.py:447-525).  The
        draws come from the device ``sample_y`` (Cholesky instead of numpy's SVD: same
        distribution, other variates)."""
        best_gain = self._optimality_margins(n_space_samples, n_gp_samples, n_random_starts, use_mean_gp,
                                             normalized_scores, random_state)
        many = is_listlike(threshold)
        probs = [float(np.mean(best_gain - eps < 0.0)) for eps in (threshold if many else [threshold])]
        return probs if many and len(probs) != 1 else probs[0]

    def expected_optimality_gap(self, max_tries=3, n_probabilities=50, n_space_samples=500, n_gp_samples=200,
                                n_random_starts=100, tol=0.01, use_mean_gp=True, normalized_scores=True,
                                random_state=None):
        """Expected optimality gap of the current optimum w.r.t. sampled consistent optima
        (bask/optimizer.py:527-620): the threshold beyond which the optimum is almost surely
        optimal is located with a bounded scalar search, then the gap distribution on
        ``n_probabilities`` thresholds below it is integrated."""
        from scipy.optimize import minimize_scalar
        seed = check_random_state(random_state).randint(0, 2 ** 32 - 1, dtype=np.int64)
        common = dict(n_random_starts=n_random_starts, n_gp_samples=n_gp_samples, n_space_samples=n_space_samples,
                      use_mean_gp=use_mean_gp, normalized_scores=normalized_scores, random_state=seed)

        def objective(eps):
            return (self.probability_of_optimality(threshold=eps, **common) - 1.0) ** 2 + 1e-3 * eps ** 2

        span = np.max(self.yi) - np.min(self.yi)
        upper = None
        for _ in range(max_tries):
            try:
                upper = minimize_scalar(objective, bounds=(0.0, span), method="bounded", options={"xatol": tol}).x
                break
            except ValueError:
                continue
        if upper is None:
            raise ValueError("Determining the upper threshold was not possible.")
        grid = np.linspace(0, upper, num=n_probabilities)
        cdf = np.asarray(self.probability_of_optimality(list(grid), **common), dtype=np.float64).reshape(-1)
        return float(np.sum(np.diff(cdf) * grid[1:]))

    def optimum_intervals(self, hdi_prob=0.95, multimodal=True, opt_samples=200, space_samples=500,
                          only_mean=True, random_state=None):
        """Highest density intervals of the optimum's location per dimension, from Thompson
        samples of the optimum (bask/optimizer.py:622-689).  Returns a list of (n_modes, 2)
        arrays in the original space."""
        if self.space.is_partly_categorical:
            raise NotImplementedError("Highest density interval not implemented for categorical parameters.")
        X = self.space.rvs(n_samples=space_samples, random_state=random_state)
        X = self.space.transform(X)
        optimum_samples = self.gp.sample_y(X, sample_mean=only_mean, n_samples=opt_samples,
                                           random_state=random_state)
        X_opt = X[np.argmin(optimum_samples, axis=0)]
        intervals = []
        for i, col in enumerate(X_opt.T):
            raw_interval = hdi(col, hdi_prob=hdi_prob, multimodal=multimodal)
            intervals.append(np.asarray(self.space.dimensions[i].inverse_transform(raw_interval)))
        return intervals
