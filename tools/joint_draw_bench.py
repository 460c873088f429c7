"""Joint posterior draws (sample_y, the work behind ThompsonSampling / PVRS) over m candidates at config-3 data:
chip-wide blocked Cholesky (bgp_dense_cholesky_inplace) vs the one-cluster kernel.  usage: joint_draw_bench.py m..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench_workloads as W
import bask_b200
w = W.config3(m=64)
gp = bask_b200.BayesGPR(kernel=bask_b200.construct_default_kernel(list(range(w.d))), normalize_y=True, random_state=0)
gp.fit(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=128, n_burnin=2, n_walkers_per_thread=128, progress=False)
for m in [int(v) for v in (sys.argv[1:] or ["1000", "2000", "4000", "10000"])]:
    X = np.random.RandomState(m).uniform(size=(m, w.d))
    row = f"m={m:6d}"
    for name, thr in (("chip-wide", 0), ("one cluster", 10 ** 9)):
        if name == "one cluster" and m > 6000:
            continue
        gp._JOINT_DRAW_BIG_M = thr
        gp.sample_y(X, sample_mean=True, n_samples=10, random_state=0)
        t0 = time.perf_counter()
        d = gp.sample_y(X, sample_mean=True, n_samples=10, random_state=0)
        row += f"   {name}: {1e3 * (time.perf_counter() - t0):9.1f} ms"
        assert np.all(np.isfinite(d))
    print(row, flush=True)
