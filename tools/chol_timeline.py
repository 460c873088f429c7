"""Per-panel clock-cycle timeline of CTA 0 of the factorisation kernel (developer tooling)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workloads as W
import bask_b200
from bask_b200._engine import Engine
from bask_b200.priors import as_device_priors
from bask_b200.utils import construct_default_kernel, guess_priors
from sklearn.gaussian_process.kernels import WhiteKernel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
w = W.config3(n=n, m=64)
e = Engine()
k = construct_default_kernel(list(range(w.d))) + WhiteKernel()
e.set_kernel(k); e.set_priors(as_device_priors(guess_priors(k), e.p)[0])
e.set_data(w.X, (w.y - w.y.mean()) / w.y.std(), 1e-10)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
th = e.to_dev(W.centre_theta(w.d) + 0.05 * np.random.RandomState(0).randn(B, w.d + 2))
P = (n + 31) // 32
st = torch.zeros(P * 12, dtype=torch.int64, device=e.device)
e.lib.bgp_debug_set_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
stid = int(sys.argv[2]) if len(sys.argv) > 2 else 0
e.logprob_dev(th); e.sync()
e.lib.bgp_debug_set_stamps(e.h, C.c_void_p(st.data_ptr()), stid)
e.logprob_dev(th); e.sync()
s = st.cpu().numpy().reshape(P, 12)
names = ["stage+diagGEMM", "Dblk asm", "potrf+inv", "wait others", "phase2", "end barrier", "p2:Kloop", "p2:init", "p2:trsm", "p2:store"]
print("panel " + " ".join(f"{x:>14s}" for x in names) + "   total")
tot = np.zeros(10)
step = max(1, P // 16)
for k in range(P):
    d = [s[k,1]-s[k,0], s[k,2]-s[k,1], s[k,3]-s[k,2], s[k,4]-s[k,3], s[k,5]-s[k,4], s[k,6]-s[k,5], s[k,7], s[k,8], s[k,9], s[k,10]]
    tot += d
    if k % step == 0:
        print(f"{k:5d} " + " ".join(f"{x:14d}" for x in d) + f" {s[k,6]-s[k,0]:8d}")
print("sum   " + " ".join(f"{int(x):14d}" for x in tot) + f" {int(tot[:6].sum()):8d}")
