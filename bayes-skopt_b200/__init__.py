"""bask_b200 -- B200-native (sm_100a) implementation of bayes-skopt's fully Bayesian GP hot
path behind the reference's own Python surface (bask/__init__.py:12-35).

Host code is Python/PyTorch (device memory, streams); every numeric step of the path runs in
hand-written CUDA behind the C ABI of ``libbgp.so`` (include/bgp.h).  No CPU fallback."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .acquisition import (LCB, PVRS, Expectation, ExpectedImprovement, MaxValueSearch,  # noqa: F401
                          ThompsonSampling, TopTwoEI, VarianceReduction, evaluate_acquisitions)
from .bayesgpr import BayesGPR  # noqa: F401
from .optimizer import Optimizer, r2_sequence, sb_sequence  # noqa: F401
from .searchcv import BayesSearchCV  # noqa: F401
from .utils import construct_default_kernel, geometric_median, guess_priors  # noqa: F401

__all__ = ["BayesGPR", "Optimizer", "guess_priors", "construct_default_kernel", "geometric_median",
           "evaluate_acquisitions", "ExpectedImprovement", "TopTwoEI", "Expectation", "LCB",
           "MaxValueSearch", "ThompsonSampling", "VarianceReduction", "PVRS", "r2_sequence", "sb_sequence",
           "BayesSearchCV"]
