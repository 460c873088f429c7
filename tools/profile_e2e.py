"""cProfile of the end-to-end (host API) cycle at the headline config."""
import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bask_b200, bench_workloads as W
from bask_b200.utils import construct_default_kernel
w = W.config3()
gp = bask_b200.BayesGPR(kernel=construct_default_kernel(list(range(w.d))), normalize_y=True, random_state=0)
gp.fit(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=128, n_burnin=10, n_walkers_per_thread=128, progress=False)
mes = bask_b200.MaxValueSearch()
def cycle(seed):
    gp.sample(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=128, n_burnin=10, n_walkers_per_thread=128)
    v = bask_b200.evaluate_acquisitions(w.candidates, gp, (mes,), n_samples=10, random_state=seed, n_min_samples=1000)[0]
    return int(np.argmax(v))
for i in range(3): cycle(i)
pr = cProfile.Profile(); pr.enable()
for i in range(10): cycle(10 + i)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
