"""Host-side timeline of the end-to-end cycle: where the GPU waits for Python (developer tooling)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bask_b200, bench_workloads as W
from bask_b200.utils import construct_default_kernel
from bask_b200._engine import Engine
w = W.config3()
gp = bask_b200.BayesGPR(kernel=construct_default_kernel(list(range(w.d))), normalize_y=True, random_state=0)
gp.fit(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=128, n_burnin=10, n_walkers_per_thread=128, progress=False)
mes = bask_b200.MaxValueSearch()
e = gp._eng()
marks = []
def wrap(name):
    f = getattr(Engine, name)
    def g(self, *a, **k):
        t0 = time.perf_counter(); r = f(self, *a, **k); marks.append((name, t0, time.perf_counter())); return r
    setattr(Engine, name, g)
for nm in ["mcmc", "sync", "factorize", "predict", "acq", "to_dev", "to_host", "set_data", "set_priors", "set_kernel", "argmax", "fetch_after"]:
    if hasattr(Engine, nm): wrap(nm)
def cycle(seed):
    marks.clear()
    t0 = time.perf_counter()
    gp.sample(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=128, n_burnin=10, n_walkers_per_thread=128)
    t1 = time.perf_counter()
    v = bask_b200.evaluate_acquisitions(w.candidates, gp, (mes,), n_samples=10, random_state=seed, n_min_samples=1000)[0]
    b = int(np.argmax(v))
    t2 = time.perf_counter()
    return t0, t1, t2
for i in range(5): cycle(i)
torch.cuda.synchronize()
t0, t1, t2 = cycle(9)
print(f"sample {1e3*(t1-t0):.3f} ms, ask {1e3*(t2-t1):.3f} ms")
for nm, a, b in marks:
    print(f"{nm:12s} start {1e3*(a-t0):8.3f}  dur {1e3*(b-a):7.3f}")
