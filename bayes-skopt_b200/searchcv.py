"""BayesSearchCV: cross-validated hyper-parameter search of a scikit-learn estimator, driven by the
fully Bayesian ``Optimizer`` (the reference wraps skopt.BayesSearchCV: bask/searchcv.py:8-354; skopt is
not a dependency here, so the search loop sits directly on scikit-learn's ``BaseSearchCV``).

Off the GPU hot path: the estimator fits run wherever scikit-learn runs them; only the surrogate
(``Optimizer.tell``: hyper-posterior MCMC + acquisition sweep) uses the device.

search_spaces: a dict {parameter name: dimension}, a list of such dicts, or a list of (dict, n_iter)
pairs; a dimension is a ``space.Real/Integer/Categorical`` or anything ``normalize_dimensions`` accepts
((low, high) tuples, lists of categories).  Scores are maximised, so the optimiser is told ``-score``."""
import numpy as np
from sklearn.model_selection._search import BaseSearchCV

from .optimizer import Optimizer

__all__ = ["BayesSearchCV", "dimensions_aslist", "point_asdict"]


def dimensions_aslist(search_space):
    """dict of dimensions -> list ordered by parameter name (skopt.utils.dimensions_aslist)."""
    return [search_space[k] for k in sorted(search_space.keys())]


def point_asdict(search_space, point_as_list):
    """list of values (in sorted-name order) -> {name: value} (skopt.utils.point_asdict)."""
    return {k: v for k, v in zip(sorted(search_space.keys()), point_as_list)}


class BayesSearchCV(BaseSearchCV):
    """See bask/searchcv.py:8-243 for the parameters.  ``optimizer_kwargs`` go to ``Optimizer`` except
    ``n_samples`` / ``gp_samples`` / ``gp_burnin``, which are passed to every ``tell`` (defaults 0 / 100 / 5);
    the acquisition defaults to "pvrs" (bask/searchcv.py:283-289)."""

    def __init__(self, estimator, search_spaces, optimizer_kwargs=None, n_iter=50, return_policy="best_setting",
                 scoring=None, fit_params=None, n_jobs=1, n_points=1, iid=True, refit=True, cv=None, verbose=0,
                 pre_dispatch="2*n_jobs", random_state=None, error_score="raise", return_train_score=False):
        self.search_spaces = search_spaces
        self.n_iter = n_iter
        self.n_points = n_points
        self.random_state = random_state
        self.optimizer_kwargs = optimizer_kwargs
        self.return_policy = return_policy
        self.fit_params = fit_params
        self.iid = iid
        super().__init__(estimator=estimator, scoring=scoring, n_jobs=n_jobs, refit=refit, cv=cv, verbose=verbose,
                         pre_dispatch=pre_dispatch, error_score=error_score, return_train_score=return_train_score)

    # ------------------------------------------------------------------ search spaces
    def _spaces(self):
        """-> list of (dict, n_iter); validates the three accepted forms (skopt BayesSearchCV._check_search_space)."""
        sp = self.search_spaces
        if isinstance(sp, dict):
            sp = [sp]
        if not isinstance(sp, list) or len(sp) == 0:
            raise TypeError("search_spaces must be a dict, a list of dicts or a list of (dict, int) tuples")
        out = []
        for entry in sp:
            if isinstance(entry, tuple):
                if len(entry) != 2 or not isinstance(entry[0], dict) or int(entry[1]) <= 0:
                    raise ValueError("a search-space tuple must be (dict, number of iterations > 0)")
                out.append((entry[0], int(entry[1])))
            elif isinstance(entry, dict):
                out.append((entry, self.n_iter))
            else:
                raise TypeError(f"search space {entry!r} is neither a dict nor a (dict, int) tuple")
        for space, _ in out:
            if len(space) == 0:
                raise ValueError("empty search space")
        return out

    @property
    def total_iterations(self):
        return sum(n for _, n in self._spaces())

    def _make_optimizer(self, params_space):
        """``Optimizer`` over the dimensions in sorted-name order; unnamed dimensions take the parameter names
        (bask/searchcv.py:292-318)."""
        kwargs = dict(self.optimizer_kwargs or {})
        for k in ("n_samples", "gp_samples", "gp_burnin"):
            kwargs.pop(k, None)
        kwargs.setdefault("acq_func", "pvrs")
        kwargs.setdefault("random_state", self.random_state)
        kwargs["dimensions"] = dimensions_aslist(params_space)
        optimizer = Optimizer(**kwargs)
        for dim, name in zip(optimizer.space.dimensions, sorted(params_space.keys())):
            if getattr(dim, "name", None) is None:
                dim.name = name
        return optimizer

    def _step(self, search_space, optimizer, evaluate_candidates):
        """ask -> cross-validate -> tell(-mean test score) (bask/searchcv.py:320-354)."""
        okw = self.optimizer_kwargs or {}
        params = [[np.array(v).item() for v in optimizer.ask(n_points=1)]]
        all_results = evaluate_candidates([point_asdict(search_space, p) for p in params])
        score_name = "mean_test_score" if "mean_test_score" in all_results else \
            f"mean_test_{self.refit}" if isinstance(self.refit, str) else \
            next(k for k in all_results if k.startswith("mean_test_"))
        scores = all_results[score_name][-len(params):]
        return optimizer.tell(params, [-float(s) for s in scores], n_samples=okw.get("n_samples", 0),
                              gp_samples=okw.get("gp_samples", 100), gp_burnin=okw.get("gp_burnin", 5), progress=False)

    def _run_search(self, evaluate_candidates):
        self.optimizer_results_ = []
        for search_space, n_iter in self._spaces():
            optimizer = self._make_optimizer(search_space)
            result = None
            for _ in range(n_iter):
                result = self._step(search_space, optimizer, evaluate_candidates)
            self.optimizer_results_.append(result)

    def fit(self, X, y=None, *, groups=None, **fit_params):
        if self.fit_params:
            fit_params = {**self.fit_params, **fit_params}
        return super().fit(X, y, groups=groups, **fit_params)
