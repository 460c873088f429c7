"""Phase clock stamps of CTA 0 of the fused small-n kernel (developer tooling).  usage: small_timeline.py n d B"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workloads as W
import bask_b200
from bask_b200._engine import Engine
from bask_b200.priors import as_device_priors
from bask_b200.utils import construct_default_kernel, guess_priors
from sklearn.gaussian_process.kernels import WhiteKernel
n, d, B = (int(v) for v in (sys.argv[1:4] + ["100", "6", "32"][len(sys.argv) - 1:]))
r = np.random.RandomState(0)
X = r.uniform(size=(n, d)); y = r.randn(n)
e = Engine()
k = construct_default_kernel(list(range(d))) + WhiteKernel()
e.set_kernel(k); e.set_priors(as_device_priors(guess_priors(k), e.p)[0]); e.set_data(X, y, 1e-10)
th = e.to_dev(W.centre_theta(d) + 0.05 * r.randn(B, d + 2))
st = torch.zeros(32, dtype=torch.int64, device=e.device)
e.lib.bgp_debug_set_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
e.logprob_dev(th); e.sync()
e.lib.bgp_debug_set_stamps(e.h, C.c_void_p(st.data_ptr()), int(sys.argv[4]) if len(sys.argv) > 4 else 0)
e.logprob_dev(th); e.sync()
s = st.cpu().numpy()
print(f"n={n} d={d} B={B}: scale {s[1]-s[0]}  gram {s[2]-s[1]}  factor {s[3]-s[2]}  epilogue {s[4]-s[3]}  total {s[4]-s[0]} cycles")
for name, o in (("kb=0", 8), ("kb=T/2", 16)):
    v = s[o:o + 6]
    print(f"   {name}: potrf8 {v[1]-v[0]}  panel {v[2]-v[1]}  barrier {v[3]-v[2]}  trailing {v[4]-v[3]}  barrier {v[5]-v[4]}")
