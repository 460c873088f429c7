"""Seeded synthetic workloads for the five BASELINE.json configs.

Pure data generation (numpy only): used by bench.py, the tests and the golden-vector
generator so that every leg sees byte-identical inputs.  Inputs live in [0,1]^d (what
skopt's "normalize" transform hands the GP), targets are raw (the GP normalises them).
"""
import math
from dataclasses import dataclass

import numpy as np


def branin01(X):
    """Branin on x1 in [-5,10], x2 in [0,15], evaluated on the unit square."""
    x1 = 15.0 * X[:, 0] - 5.0
    x2 = 15.0 * X[:, 1]
    b, c, r, s, t = 5.1 / (4 * math.pi ** 2), 5.0 / math.pi, 6.0, 10.0, 1.0 / (8 * math.pi)
    return (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s


_H6_A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14],
                  [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]], dtype=np.float64)
_H6_P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                         [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]],
                        dtype=np.float64)
_H6_ALPHA = np.array([1.0, 1.2, 3.0, 3.2])


def hartmann6(X):
    inner = np.einsum("kj,ikj->ik", _H6_A, (X[:, None, :] - _H6_P[None]) ** 2)
    return -np.exp(-inner).dot(_H6_ALPHA)


def wavy6(X):
    return np.sin(3.0 * X.sum(1)) + 0.5 * np.cos(5.0 * X[:, 0])


def ackley01(X):
    Z = (2.0 * X - 1.0) * 32.768
    d = Z.shape[1]
    return (-20.0 * np.exp(-0.2 * np.sqrt((Z ** 2).sum(1) / d))
            - np.exp(np.cos(2 * math.pi * Z).sum(1) / d) + 20.0 + math.e)


@dataclass
class Workload:
    name: str
    X: np.ndarray            # (n, d) in [0,1]^d
    y: np.ndarray            # (n,) raw targets
    noise_vector: np.ndarray  # (n,) per-point noise variances handed to tell() (zeros)
    candidates: np.ndarray   # (m, d)
    n_walkers: int
    n_desired_samples: int
    n_burnin: int
    acquisition: str         # registry string of bask/optimizer.py:23-32
    n_theta_samples: int     # n_samples of evaluate_acquisitions
    acq_kwargs: dict
    mes_seed: int = 2        # seed of the GLOBAL numpy RNG MaxValueSearch reads

    @property
    def n(self):
        return self.X.shape[0]

    @property
    def d(self):
        return self.X.shape[1]

    @property
    def n_steps(self):
        return int(math.ceil(self.n_desired_samples / self.n_walkers) + self.n_burnin)

    @property
    def n_logprob_evals(self):
        return self.n_walkers * (1 + self.n_steps)


def _make(name, f, n, d, m, seed_x, seed_c, noise, **kw):
    r = np.random.RandomState(seed_x)
    X = r.uniform(size=(n, d))
    y = f(X) + noise * r.randn(n)
    Xc = np.random.RandomState(seed_c).uniform(size=(m, d))
    return Workload(name=name, X=X, y=y, noise_vector=np.zeros(n), candidates=Xc, **kw)


def config1():
    """Branin 2-D, 20 observations, EI over 500 candidates (the reference's CPU case)."""
    return _make("C1-branin-n20-ei-m500", branin01, 20, 2, 500, 0, 1, 0.05, n_walkers=100,
                 n_desired_samples=100, n_burnin=10, acquisition="ei", n_theta_samples=10,
                 acq_kwargs={})


def config2():
    """Hartmann-6, 100 observations, PVRS, 64 walkers, 1k candidates."""
    return _make("C2-hartmann6-n100-pvrs-m1000", hartmann6, 100, 6, 1000, 10, 11, 0.05,
                 n_walkers=64, n_desired_samples=64, n_burnin=10, acquisition="pvrs",
                 n_theta_samples=0, acq_kwargs={"n_thompson": 10})


def config3(n=500, m=10000, n_walkers=128):
    """6-D synthetic, 500 observations, MES, 128 walkers, 10k candidates (headline)."""
    return _make(f"C3-wavy6-n{n}-mes-m{m}", wavy6, n, 6, m, 20, 21, 0.05, n_walkers=n_walkers,
                 n_desired_samples=n_walkers, n_burnin=10, acquisition="mes",
                 n_theta_samples=10, acq_kwargs={"n_min_samples": 1000})


def config4(n, batch, d=6):
    """batched-theta LML sweep: same generator as C3, theta ~ centre + 0.1 N(0, I)."""
    w = _make(f"C4-lml-n{n}-B{batch}", wavy6, n, d, 8, 30, 31, 0.05, n_walkers=batch,
              n_desired_samples=batch, n_burnin=0, acquisition="ei", n_theta_samples=1,
              acq_kwargs={})
    centre = np.concatenate([[0.0], np.log(0.3) * np.ones(d), [np.log(0.05)]])
    thetas = centre + 0.1 * np.random.RandomState(3).randn(batch, d + 2)
    return w, thetas


def config5(n=2000, m=100000, n_walkers=256, acquisition="ei"):
    """20-D Ackley, 2000 observations, 256 walkers, 100k candidates (multi-GPU sweep)."""
    return _make(f"C5-ackley20-n{n}-{acquisition}-m{m}", ackley01, n, 20, m, 40, 4, 0.05,
                 n_walkers=n_walkers, n_desired_samples=n_walkers, n_burnin=10,
                 acquisition=acquisition, n_theta_samples=16,
                 acq_kwargs={"n_min_samples": 1000} if acquisition == "mes" else {})


def centre_theta(d):
    """[log c, log l_1..l_d, log sigma^2] of the default kernel near its prior mode."""
    return np.concatenate([[0.0], np.log(0.3) * np.ones(d), [np.log(0.05)]])
