"""Small run through every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
fused small-n path, blocked path with clusters, factorise + sweep (resident and windowed), MES epilogue,
a two-step device MCMC, the LML gradient.  Checks nothing numerically -- the parity tests do that."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workloads as W
import bask_b200
from bask_b200 import _lib
from bask_b200._engine import Engine
from bask_b200.priors import as_device_priors
from bask_b200.utils import construct_default_kernel, guess_priors
from sklearn.gaussian_process.kernels import WhiteKernel

for n, d in ((40, 2), (300, 6)):
    r = np.random.RandomState(n)
    X = r.uniform(size=(n, d)); y = np.sin(3 * X.sum(1)) + 0.1 * r.randn(n); y = (y - y.mean()) / y.std()
    e = Engine()
    k = construct_default_kernel(list(range(d))) + WhiteKernel()
    e.set_kernel(k); e.set_priors(as_device_priors(guess_priors(k), e.p)[0]); e.set_data(X, y, 1e-10)
    th = W.centre_theta(d) + 0.1 * r.randn(12, d + 2)
    lp, lml, info = e.logprob(th)
    assert np.isfinite(lp).all()
    f = e.factorize(th[:3])
    Xc = e.to_dev(r.uniform(size=(200, d)))
    for windowed in (False, True):
        if windowed:
            os.environ["BGP_SWEEP_WINDOWED"] = "1"
        mu, sd, _, _ = e.predict(f, Xc, noise_off=True)
        os.environ.pop("BGP_SWEEP_WINDOWED", None)
    g = e.to_dev(np.stack([bask_b200.acquisition.gumbel32_like_reference(64) for _ in range(3)]), dtype=torch.float32)
    for kind in (_lib.ACQ_EI, _lib.ACQ_TTEI, _lib.ACQ_LCB, _lib.ACQ_MES):
        e.acq(kind, mu, sd, gumbel32=g if kind == _lib.ACQ_MES else None)
    e.lml_gradient(th[0])
    b = e.mcmc(W.centre_theta(d) + 0.05 * r.randn(4 * (d + 2), d + 2), 2, 7)
    e.sync()
    assert np.isfinite(b["chain"].cpu().numpy()).all()
print("sanitize target ok")
