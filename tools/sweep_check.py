"""Developer check (GPU): the sweep kernel against the round-1 kernel and, forced, its windowed mode,
over a range of n; prints max relative differences and timings."""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bask_b200  # noqa: E402,F401
import bench_workloads as W  # noqa: E402
from bask_b200._engine import Engine  # noqa: E402
from bask_b200.utils import construct_default_kernel  # noqa: E402
from sklearn.gaussian_process.kernels import WhiteKernel  # noqa: E402
import torch  # noqa: E402


def run(n, d, m, S, mode):
    for k in ("BGP_SWEEP_V1", "BGP_SWEEP_WINDOWED"):
        os.environ.pop(k, None)
    if mode == "v1":
        os.environ["BGP_SWEEP_V1"] = "1"
    if mode == "win":
        os.environ["BGP_SWEEP_WINDOWED"] = "1"
    w = W.config3(n=n, m=m) if d == 6 else W.config5(n=n, m=m)
    e = Engine()
    e.set_kernel(construct_default_kernel(list(range(w.d))) + WhiteKernel())
    y = (w.y - w.y.mean()) / w.y.std()
    e.set_data(w.X, y, 1e-10 * np.ones(n))
    th = W.centre_theta(w.d) + 0.1 * np.random.RandomState(3).randn(S, w.d + 2)
    f = e.factorize(th)
    Xc = e.to_dev(w.candidates)
    mu, sd, _, _ = e.predict(f, Xc, noise_off=True, y_mean=0.2, y_std=1.3)
    e.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(e.stream)
    for _ in range(3):
        e.predict(f, Xc, noise_off=True, y_mean=0.2, y_std=1.3)
    b.record(e.stream)
    e.sync()
    return e.to_host(mu), e.to_host(sd), a.elapsed_time(b) / 3


if __name__ == "__main__":
    cases = [(20, 6, 70, 2), (64, 6, 333, 3), (500, 6, 10000, 10), (544, 6, 4096, 4), (600, 6, 4096, 4),
             (1100, 6, 4096, 4), (2000, 20, 8192, 4)]
    for n, d, m, S in cases:
        ref = run(n, d, m, S, "v1")
        row = [f"n={n} d={d} m={m} S={S}: v1 {ref[2]:.3f} ms"]
        for mode in ("new", "win"):
            t0 = time.time()
            out = run(n, d, m, S, mode)
            dmu = np.max(np.abs(out[0] - ref[0]) / (np.abs(ref[0]) + 1e-12))
            dsd = np.max(np.abs(out[1] - ref[1]) / (np.abs(ref[1]) + 1e-12))
            flops = S * m * (float(n) ** 2 + n * (3 * d + 14))
            row.append(f"{mode} {out[2]:.3f} ms ({flops / out[2] / 1e9:.1f} TF) dmu {dmu:.1e} dsd {dsd:.1e}")
        print(" | ".join(row), flush=True)
