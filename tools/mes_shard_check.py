"""One-GPU emulation of the candidate-sharded MES sweep (developer tooling): are the theta-sharded
Gumbel fit and the per-block epilogue bit-identical to the single-GPU composition?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workloads as W
import bask_b200
from bask_b200 import _lib
from bask_b200.distributed import DeviceBackend, shard_bounds
from bask_b200.utils import construct_default_kernel

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = W.config5(m=20000, acquisition="mes") if cfg == "c5" else W.config3(m=20000)
S, K, world = w.n_theta_samples, 1000, 2
gp = bask_b200.BayesGPR(kernel=construct_default_kernel(list(range(w.d))), normalize_y=True, random_state=0)
gp.fit(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=w.n_walkers, n_burnin=2,
       n_walkers_per_thread=w.n_walkers, progress=False)
e = gp._eng()
b = DeviceBackend(gp)
th = e.to_dev(gp.chain_[np.random.RandomState(1).choice(len(gp.chain_), replace=False, size=S)])
Xc = e.to_dev(w.candidates)
g32 = e.to_dev(np.stack([bask_b200.acquisition.gumbel32_like_reference(K) for _ in range(S)]), dtype=torch.float32)
with torch.cuda.stream(e.stream):
    mu, sd = b.moments(th, Xc)
    out1, pt1, sk1, fit1 = e.acq(_lib.ACQ_MES, mu, sd, gumbel32=g32, want_fit=True)
    # sharded emulation
    blocks = [shard_bounds(len(w.candidates), world, r) for r in range(world)]
    mus, sds = zip(*[b.moments(th, Xc[lo:hi]) for lo, hi in blocks])
    mu_all, sd_all = torch.cat(mus, 1).contiguous(), torch.cat(sds, 1).contiguous()
    fits = [b.mes_fit(mu_all[lo:hi].contiguous(), sd_all[lo:hi].contiguous()) for lo, hi in
            [shard_bounds(S, world, r) for r in range(world)]]
    fit2 = torch.cat(fits, 0).contiguous()
    vals = [b.per_theta(_lib.ACQ_MES, m_, s_, float("nan"), gumbel=g32, fit=fit2) for m_, s_ in zip(mus, sds)]
    sk = torch.stack([v[1] for v in vals]).max(0).values.contiguous()
    out2 = torch.cat([b.combine(v[0], sk) for v in vals])
e.sync()
H = lambda t: t.cpu().numpy()
print("mu bit-identical:", np.array_equal(H(mu), H(mu_all)), " sd:", np.array_equal(H(sd), H(sd_all)))
print("fit max abs diff:", np.abs(H(fit1) - H(fit2)).max(), "\nfit1[0]", H(fit1)[0], "\nfit2[0]", H(fit2)[0])
a, c = H(out1), H(out2)
rel = np.abs(a - c) / np.maximum(np.abs(a), 1e-300)
i = int(np.argmax(rel))
print("out max rel diff:", rel.max(), "at", i, a[i], c[i], " skipped:", H(sk1), H(sk))
print("values range:", a.min(), a.max(), " rel diff above 1e-12*max:", rel[a > 1e-12 * a.max()].max())
