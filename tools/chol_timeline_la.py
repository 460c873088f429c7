"""Timeline of the look-ahead factorisation kernel (CTA 0): tid 0 = diagonal team, tid 128 = bulk warp."""
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
th = e.to_dev(W.centre_theta(w.d) + 0.05 * np.random.RandomState(0).randn(64, w.d + 2))
P = (n + 31) // 32
e.lib.bgp_debug_set_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
e.logprob_dev(th); e.sync()
for tid, names in ((0, ["stage", "own tile", "diag gemm", "assemble", "potrf", "wait end"]), (128, ["stage", "bulk tiles", "-", "-", "-", "wait end"])):
    st = torch.zeros(P * 12, dtype=torch.int64, device=e.device)
    e.lib.bgp_debug_set_stamps(e.h, C.c_void_p(st.data_ptr()), tid)
    e.logprob_dev(th); e.sync()
    s = st.cpu().numpy().reshape(P, 12)
    print(f"tid {tid}: " + " ".join(f"{x:>10s}" for x in names) + "      total")
    tot = np.zeros(6)
    for k_ in range(P):
        if tid == 0:
            d = [s[k_,1]-s[k_,0], s[k_,2]-s[k_,1], s[k_,3]-s[k_,2], s[k_,4]-s[k_,3], s[k_,5]-s[k_,4], s[k_,6]-s[k_,5]]
        else:
            d = [s[k_,1]-s[k_,0], s[k_,5]-s[k_,1], 0, 0, 0, s[k_,6]-s[k_,5]]
        d = [int(x) if abs(x) < 1e9 else 0 for x in d]
        tot += d
        print(f"  {k_:5d} " + " ".join(f"{x:10d}" for x in d) + f" {s[k_,6]-s[k_,0]:10d}")
    print("  sum   " + " ".join(f"{int(x):10d}" for x in tot) + f" {int(tot.sum()):10d}")
