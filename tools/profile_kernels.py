"""Small driver for ncu: runs the headline-shape kernels a few times (C3: n=500, d=6).
usage: python tools/profile_kernels.py [chol|sweep|mes|all]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workloads as W
import bask_b200
from bask_b200 import _lib
from bask_b200._engine import Engine
from bask_b200.priors import as_device_priors
from bask_b200.utils import construct_default_kernel, guess_priors
from sklearn.gaussian_process.kernels import WhiteKernel

what = sys.argv[1] if len(sys.argv) > 1 else "all"
w = W.config3()
e = Engine()
k = construct_default_kernel(list(range(w.d))) + WhiteKernel()
e.set_kernel(k)
e.set_priors(as_device_priors(guess_priors(k), e.p)[0])
y = (w.y - w.y.mean()) / w.y.std()
e.set_data(w.X, y, 1e-10)
th = W.centre_theta(w.d) + 0.05 * np.random.RandomState(0).randn(64, w.d + 2)
thd = e.to_dev(th)
reps = 3
if what in ("chol", "all"):
    for _ in range(reps):
        e.logprob_dev(thd)
if what in ("sweep", "mes", "all"):
    f = e.factorize(thd[:10].contiguous())
    Xc = e.to_dev(w.candidates)
    for _ in range(reps):
        mu, sd, _, _ = e.predict(f, Xc, noise_off=True)
    if what in ("mes", "all"):
        g = np.stack([bask_b200.acquisition.gumbel32_like_reference(1000) for _ in range(10)])
        gd = e.to_dev(g, dtype=torch.float32)
        for _ in range(reps):
            e.acq(_lib.ACQ_MES, mu, sd, gumbel32=gd)
e.sync()
print("done")
