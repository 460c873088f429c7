"""ncu driver for the small-n configurations (C1: n=20, C2: n=100): a short device MCMC each.
usage: python tools/profile_small.py [c1|c2]   (developer tooling)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workloads as W
import bask_b200
from bask_b200._engine import Engine
from bask_b200.priors import as_device_priors
from bask_b200.utils import construct_default_kernel, guess_priors
from sklearn.gaussian_process.kernels import WhiteKernel

for tag in (sys.argv[1:] or ["c1", "c2"]):
    w = W.config1() if tag == "c1" else W.config2()
    e = Engine()
    k = construct_default_kernel(list(range(w.d))) + WhiteKernel()
    e.set_kernel(k)
    e.set_priors(as_device_priors(guess_priors(k), e.p)[0])
    e.set_data(w.X, (w.y - w.y.mean()) / w.y.std(), 1e-10)
    pos = W.centre_theta(w.d) + 0.05 * np.random.RandomState(0).randn(w.n_walkers, w.d + 2)
    b = None
    for rep in range(3):
        e.sync(); t0 = time.perf_counter()
        b = e.mcmc(pos, 11, 100 + rep, buffers=b)
        e.sync(); dt = time.perf_counter() - t0
    print(tag, "mcmc 11 steps: %.3f ms wall" % (1e3 * dt))
    th = e.to_dev(pos[: (w.n_walkers + 1) // 2])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e.logprob_dev(th); e.sync()
    ev[0].record(e.stream)
    for _ in range(20):
        e.logprob_dev(th)
    ev[1].record(e.stream); e.sync()
    print(tag, "logprob of %d thetas: %.1f us" % (len(th), 1e3 * ev[0].elapsed_time(ev[1]) / 20))
