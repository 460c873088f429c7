"""Host-side helpers of the hot path: default kernel, guessed priors, geometric median.
Mirrors bask/utils.py (same names, argument meaning and error behaviour)."""
import collections.abc

import numpy as np
from sklearn.gaussian_process.kernels import ConstantKernel, Matern

from .priors import HalfNormalSqrtPrior, RoundFlatPrior

__all__ = ["geometric_median", "guess_priors", "construct_default_kernel", "validate_zeroone"]


def geometric_median(X, eps=1e-5):
    """Weiszfeld iteration for the point minimising the summed L2 distance to the rows of X
    (bask/utils.py:21-65).  O(N p) per iteration on a (N, p) chain: stays on the host."""
    X = np.asarray(X, dtype=np.float64)
    y = np.mean(X, 0)
    N = len(X)
    while True:
        diff = X - y
        D = np.sqrt(np.einsum("ij,ij->i", diff, diff))
        nz = D != 0
        if nz.all():
            Dinv = 1.0 / D
            y1 = (Dinv / Dinv.sum()) @ X
        else:
            num_zeros = N - int(nz.sum())
            if num_zeros == N:
                return y
            Dinv = 1.0 / D[nz]
            Dinvs = Dinv.sum()
            T = (Dinv / Dinvs) @ X[nz]
            R = (T - y) * Dinvs
            r = np.linalg.norm(R)
            rinv = 0 if r == 0 else num_zeros / r
            y1 = max(0, 1 - rinv) * T + min(1, rinv) * y
        d = y - y1
        if np.sqrt(d @ d) < eps:
            return y1
        y = y1


def _recursive_priors(kernel, prior_list):
    if hasattr(kernel, "kernel"):  # Exponentiation
        _recursive_priors(kernel.kernel, prior_list)
    elif hasattr(kernel, "k1"):  # Sum / Product
        _recursive_priors(kernel.k1, prior_list)
        _recursive_priors(kernel.k2, prior_list)
    elif hasattr(kernel, "kernels"):  # CompoundKernel
        for k in kernel.kernels:
            _recursive_priors(k, prior_list)
    else:
        name = type(kernel).__name__
        if name in ["ConstantKernel", "WhiteKernel"]:
            if name == "ConstantKernel" and kernel.constant_value_bounds == "fixed":
                return
            if name == "WhiteKernel" and kernel.noise_level_bounds == "fixed":
                return
            prior_list.append(HalfNormalSqrtPrior(scale=2.0))
        elif name in ["Matern", "RBF"]:
            if isinstance(kernel.length_scale, (collections.abc.Sequence, np.ndarray)):
                n_priors = len(kernel.length_scale)
            else:
                n_priors = 1
            roundflat = RoundFlatPrior(lower_bound=0.1, upper_bound=0.6, lower_steepness=2.0,
                                       upper_steepness=8.0)
            for _ in range(n_priors):
                prior_list.append(roundflat)
        else:
            raise NotImplementedError(f"Unable to guess priors for this kernel: {kernel}.")


def guess_priors(kernel):
    """Half-normal(0, 2) priors on every Constant/White level, round-flat(0.1, 0.6) on every
    length scale, each with the log-space Jacobian, in ``kernel.theta`` order
    (bask/utils.py:154-179).  The returned objects are callable like the reference's lambdas
    AND typed, so the CUDA kernel can evaluate them."""
    priors = []
    _recursive_priors(kernel, priors)
    return priors


def construct_default_kernel(dimensions):
    """Constant(1.0, (0.1, 2.0)) * Matern-5/2 ARD(0.3, (0.2, 0.5)) -- bask/utils.py:127-151."""
    n_parameters = len(dimensions)
    return ConstantKernel(constant_value=1.0, constant_value_bounds=(0.1, 2.0)) * Matern(
        length_scale=[0.3] * n_parameters, length_scale_bounds=(0.2, 0.5), nu=2.5)


def validate_zeroone(arr):
    """Raises ValueError unless every entry lies in [0, 1] (bask/utils.py:212-228)."""
    if not isinstance(arr, np.ndarray):
        arr = np.array(arr)
    if np.any(arr < 0) or np.any(arr > 1):
        raise ValueError("Not all values of the array are between 0 and 1.")
