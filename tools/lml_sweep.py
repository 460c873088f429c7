"""BASELINE.json configs[3]: batched-theta LML + Cholesky throughput sweep, n = 128..4096 x batch 64..1024
(d = 6, default kernel + White, guess_priors).  Prints one JSON line per (n, B) and a table."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workloads as W
import bask_b200
from bask_b200._engine import Engine
from bask_b200.priors import as_device_priors
from bask_b200.utils import construct_default_kernel, guess_priors
from sklearn.gaussian_process.kernels import WhiteKernel

def flops_lml(n, d): return n ** 3 / 3.0 + 2.0 * n ** 2 + 0.5 * n * (n - 1) * (3 * d + 12)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "profiles", "fp64_peaks_r01.json")))["dmma_tflops_w8"]
ns = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [128, 256, 512, 1024, 2048, 4096]
Bs = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 256, 1024]
rows = []
for n in ns:
    for B in Bs:
        w, thetas = W.config4(n, B)
        e = Engine()
        k = construct_default_kernel(list(range(w.d))) + WhiteKernel()
        e.set_kernel(k); e.set_priors(as_device_priors(guess_priors(k), e.p)[0])
        e.set_data(w.X, (w.y - w.y.mean()) / w.y.std(), 1e-10)
        th = e.to_dev(thetas)
        lp, _, info = e.logprob_dev(th); e.sync()
        reps = 2 if n >= 4096 else (3 if n >= 2048 else 10)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(e.stream)
        for _ in range(reps):
            e.logprob_dev(th)
        e1.record(e.stream); e.sync()
        ms = e0.elapsed_time(e1) / reps
        tf = B * flops_lml(n, w.d) / (ms * 1e-3) / 1e12
        row = {"n": n, "batch": B, "ms": ms, "lml_evals_per_s": B / (ms * 1e-3), "tflops": tf, "frac_of_dmma_peak": tf / peak,
               "all_finite": bool(torch.isfinite(lp).all().item()), "info_zero": bool((info == 0).all().item())}
        rows.append(row); print(json.dumps(row), flush=True)
        del e
print(f"{'n':>6s} {'batch':>6s} {'ms':>10s} {'evals/s':>12s} {'TFLOP/s':>9s} {'of DMMA peak':>13s}")
for r in rows:
    print(f"{r['n']:6d} {r['batch']:6d} {r['ms']:10.3f} {r['lml_evals_per_s']:12.0f} {r['tflops']:9.2f} {100*r['frac_of_dmma_peak']:12.1f}%")
