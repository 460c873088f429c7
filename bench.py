"""Benchmark of the fully Bayesian GP hot path on B200 (contract: see the task brief).

Workload (BASELINE.json configs[2], the headline): 6-D synthetic, n=500 observations,
Constant*Matern-5/2 ARD + White, 128 walkers x 11 stretch-move steps (1 536 log-posterior
evaluations, warm-started like Optimizer.tell's 2nd+ call), then the MaxValueSearch sweep over
10 theta samples x 10 000 candidates and the argmax.  One "step" = one such sample()+ask() cycle.

value  : LML evaluations per second over the whole cycle, inputs resident in HBM
         (1 536 * n_gpus_cycles / cycle time; the sweep's time is charged to it on purpose, so
         the ratio against the reference arm is the cycle speed-up the north star asks for).
e2e    : the same through the public API (BayesGPR.sample + evaluate_acquisitions + argmax) with
         numpy inputs: H2D of X, y, noise, candidates and Gumbel variates, D2H of the chain,
         walker positions and acquisition values inside the timed region.

--impl reference times the CPU restatement of the reference path (oracle/, kind "port":
scikit-optimize and emcee are not installable here) on the host cores, on a bounded sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

# Libraries (NCCL's version banner, torchrun notices) may write to stdout; the contract is ONE JSON
# line there, so file descriptor 1 is pointed at stderr for the whole run and the line is written to
# the saved descriptor at the end.
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import bench_workloads as W  # noqa: E402

METRIC = "GP LML evals/sec (batched theta) over one BayesGPR.sample + ask() cycle at n=500 obs, 10k candidates"
UNIT = "LML evals/s"


def flops_lml(n, d):
    """potrf + two trsv + Gram (exp/sqrt counted as 1)."""
    return n ** 3 / 3.0 + 2.0 * n ** 2 + 0.5 * n * (n - 1) * (3 * d + 12)


def flops_sweep(n, d):
    return float(n) ** 2 + n * (3 * d + 14)


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self, t0=None, t1=None):
        """median SM clock / throttle reasons of the samples that arrived inside [t0, t1] (the timed
        region); nvidia-smi needs ~0.2 s to start, so the sampler is started before the warm-up"""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r[1:] for r in self.rows if t0 is None or (t0 - 0.02 <= r[0] <= t1 + 0.12)]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_reference_cycle(w, lml_evals=48, sweep_thetas=1, sweep_cands=400, seed=0):
    """Times the oracle (CPU port of the reference path) on a bounded sample of the workload and
    extrapolates to the full cycle.  Returns (cycle_seconds_estimate, detail dict)."""
    from oracle import acq_oracle as A
    from oracle import gp_oracle as G
    spec = ("sum", ("product", ("const", 1.0, False), ("matern", 0.3 * np.ones(w.d), 2.5, False)),
            ("white", 1.0, False))
    priors = G.guess_priors(spec)
    y = (w.y - w.y.mean()) / w.y.std()
    alpha = 1e-10 * np.ones(w.n) + w.noise_vector
    rng = np.random.RandomState(seed)
    thetas = W.centre_theta(w.d) + 0.05 * rng.randn(lml_evals, w.d + 2)
    t0 = time.perf_counter()
    for t in thetas:
        G.log_prob(spec, t, w.X, y, alpha, priors)
    t_lml = (time.perf_counter() - t0) / lml_evals
    gp = A.GPState(spec=spec, X=w.X, y=y, alpha=alpha, chain=thetas, theta=thetas[0].copy(),
                   y_mean=float(w.y.mean()), y_std=float(w.y.std()))
    Xc = w.candidates[:sweep_cands]
    t0 = time.perf_counter()
    for s in range(sweep_thetas):
        gp.set_theta(thetas[s])
    t_setter = (time.perf_counter() - t0) / sweep_thetas
    t0 = time.perf_counter()
    for s in range(sweep_thetas):
        mu, sd = G.predict(spec, thetas[s], w.X, Xc, gp.K_inv, gp.a, gp.y_mean, gp.y_std)
    t_pred = (time.perf_counter() - t0) / sweep_thetas / sweep_cands
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        if w.acquisition == "mes":
            A.max_value_search(mu, sd, n_min_samples=w.acq_kwargs.get("n_min_samples", 1000),
                               uniforms=rng.rand(w.acq_kwargs.get("n_min_samples", 1000)).astype(np.float32))
        else:
            A.UNCERTAINTY_FN[w.acquisition](mu, sd)
    t_acq = (time.perf_counter() - t0) / sweep_cands
    m, S = len(w.candidates), w.n_theta_samples
    t_sample = w.n_logprob_evals * t_lml
    t_ask = S * (t_setter + m * (t_pred + t_acq))
    detail = {"lml_eval_ms": 1e3 * t_lml, "theta_setter_ms": 1e3 * t_setter,
              "predict_us_per_candidate": 1e6 * t_pred, "acq_us_per_candidate": 1e6 * t_acq,
              "sample_s_extrapolated": t_sample, "ask_s_extrapolated": t_ask}
    return t_sample + t_ask, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = W.config3()
    times, detail = [], None
    # bounded sample per step, sized so that the whole run takes about 75 s of CPU time
    # (BGP_BENCH_REF_BUDGET_S overrides the total, used by the CPU test of the output contract)
    budget = float(os.environ.get("BGP_BENCH_REF_BUDGET_S", "75")) / max(1, args.warmup + args.steps)
    n_lml = int(min(256, max(16, budget * 0.4 / 0.006)))
    n_cand = int(min(2000, max(200, budget * 0.5 / 0.00045)))
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cyc, detail = cpu_reference_cycle(w, lml_evals=n_lml, sweep_thetas=1, sweep_cands=n_cand, seed=i)
        if i >= args.warmup:
            times.append(cyc)
        detail["sample_wall_s"] = time.perf_counter() - t0
    cyc = float(np.mean(times))
    value = w.n_logprob_evals / cyc
    sample = (f"per step: {n_lml} log-posterior evals + 1 theta-setter + predict/MES over {n_cand} candidates at "
              "n=500, extrapolated linearly to 1536 evals + 10 thetas x 10000 candidates")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cyc, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w.name, "n_obs": w.n, "dims": w.d, "walkers": w.n_walkers,
                       "mcmc_steps": w.n_steps, "theta_samples": w.n_theta_samples,
                       "candidates": len(w.candidates), "acquisition": w.acquisition},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": sample, "detail": detail},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import bask_b200
    from bask_b200 import _lib
    from bask_b200.utils import construct_default_kernel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = W.config3()
    m_local = len(w.candidates)
    # weak scaling: the candidate set grows with the number of GPUs (10k per rank) and is sharded by
    # bask_b200.distributed; the (tiny) MCMC is replicated with identical Philox streams, so the only
    # collectives are the sweep's per-theta scalars and the final all-gather of acquisition values
    cands = np.random.RandomState(21).uniform(size=(m_local * world, w.d)) if world > 1 else w.candidates
    gp = bask_b200.BayesGPR(kernel=construct_default_kernel(list(range(w.d))), normalize_y=True,
                            random_state=0, device=local)
    pg = dist.group.WORLD if world > 1 else None
    Wk = w.n_walkers * world       # weak scaling: 128 walkers per GPU, sharded by bask_b200.distributed
    gp.fit(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=Wk, n_burnin=w.n_burnin,
           n_walkers_per_thread=Wk, progress=False, process_group=pg)
    e = gp._eng()
    mes = bask_b200.MaxValueSearch()
    K = w.acq_kwargs["n_min_samples"]
    S, T = w.n_theta_samples, w.n_steps

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=e.device)

    # ---------------- device-resident cycle (inputs already in HBM)
    pos_dev = e.to_dev(gp.pos_)
    Xc_dev = e.to_dev(cands)
    sweep = None
    if world > 1:
        from bask_b200.distributed import DeviceBackend, ShardedSweep
        sweep = ShardedSweep(DeviceBackend(gp), dist.group.WORLD, keep_on_device=True)
    picks = np.random.RandomState(1).choice(len(gp.chain_), replace=False, size=S)
    g32 = np.stack([bask_b200.acquisition.gumbel32_like_reference(K) for _ in range(S)])
    g32_dev = e.to_dev(g32, dtype=torch.float32)
    y_mean, y_std = float(gp.y_train_mean_), float(gp.y_train_std_)
    bufs = {"mc": None}

    def device_cycle(seed):
        with torch.cuda.stream(e.stream):
            flush.zero_()
        if world > 1:
            from bask_b200.distributed import sharded_mcmc
            chain_h, _pos_h, _acc = sharded_mcmc(e, gp.pos_, T, seed, 2.0, pg)
            th = e.to_dev(chain_h[-1][picks % Wk])
        else:
            b = e.mcmc(pos_dev, T, seed, buffers=bufs["mc"])
            bufs["mc"] = b
            with torch.cuda.stream(e.stream):
                th = b["chain"][-1][torch.as_tensor(picks % Wk, device=e.device)].contiguous()
        if sweep is not None:
            out = sweep.evaluate(Xc_dev, th, [(_lib.ACQ_MES, float("nan"))], {0: g32_dev})[0]
            return e.argmax(out.contiguous())
        f = e.factorize(th)
        mu, sd, _, _ = e.predict(f, Xc_dev, noise_off=True, y_mean=y_mean, y_std=y_std)
        out, _, _, _ = e.acq(_lib.ACQ_MES, mu, sd, gumbel32=g32_dev)
        return e.argmax(out)

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        for i in range(args.warmup):
            device_cycle(100 + i)
        barrier()
        l0 = e.launches
        t_begin = time.time()
        ev0.record(e.stream)
        for i in range(args.steps):
            device_cycle(200 + i)
        ev1.record(e.stream)
        barrier()
        t_end = time.time()
        time.sleep(0.15)
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    launches = (e.launches - l0) // max(args.steps, 1)
    clocks = clk.summary(t_begin, t_end)

    # ---------------- end to end through the public API (host buffers in, host results out)
    def e2e_cycle(seed):
        with torch.cuda.stream(e.stream):
            flush.zero_()
        gp.sample(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=Wk, n_burnin=w.n_burnin,
                  n_walkers_per_thread=Wk, process_group=pg)
        vals = bask_b200.evaluate_acquisitions(cands, gp, (mes,), n_samples=S, random_state=seed,
                                               process_group=dist.group.WORLD if world > 1 else None,
                                               **w.acq_kwargs)[0]
        return int(np.argmax(vals))

    for i in range(args.warmup):
        e2e_cycle(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_cycle(10 + i)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    h2d = 8 * (w.X.size + 2 * w.n + Wk * (w.d + 2) + cands.size // world + S * (w.d + 2)) + 4 * S * K
    d2h = 8 * (T * Wk * (w.d + 2) + Wk * (w.d + 2) + m_local * world + 1 + S)

    # ---------------- roofline of the dominant kernel, timed alone on its launch stream
    nb = (w.n_walkers + 1) // 2
    th64 = e.to_dev(gp.chain_[:nb])
    for _ in range(3):
        e.logprob_dev(th64)
    reps = 20
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e.sync()
    k0.record(e.stream)
    for _ in range(reps):
        e.logprob_dev(th64)
    k1.record(e.stream)
    e.sync()
    chol_ms = k0.elapsed_time(k1) / reps
    peaks = json.load(open(os.path.join(REPO, "profiles", "fp64_peaks_r01.json")))
    peak_tf = float(peaks["dmma_tflops_w8"])
    try:    # DRAM bytes per launch of the factorisation kernel from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic_r01.json")))["chol_lml_kernel"]
    except Exception:
        traffic = None
    chol_tf = nb * flops_lml(w.n, w.d) / (chol_ms * 1e-3) / 1e12
    # sweep kernel
    th = e.to_dev(gp.chain_[picks])
    f = e.factorize(th)
    Xc_dev = Xc_dev[:m_local].contiguous()
    for _ in range(2):
        e.predict(f, Xc_dev, noise_off=True, y_mean=y_mean, y_std=y_std)
    e.sync()
    k0.record(e.stream)
    for _ in range(5):
        e.predict(f, Xc_dev, noise_off=True, y_mean=y_mean, y_std=y_std)
    k1.record(e.stream)
    e.sync()
    sweep_ms = k0.elapsed_time(k1) / 5
    sweep_tf = S * m_local * flops_sweep(w.n, w.d) / (sweep_ms * 1e-3) / 1e12

    times = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=e.device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = [float(v) for v in times.cpu()]

    if rank == 0:
        evals = Wk * (1 + T)          # distinct log-posterior evaluations per cycle, all ranks together
        line = {"metric": METRIC, "value": evals / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w.name, "n_obs": w.n, "dims": w.d, "walkers": Wk, "walkers_per_gpu": w.n_walkers, "mcmc_steps": T,
                           "theta_samples": S, "candidates_per_gpu": m_local, "acquisition": w.acquisition,
                           "l2": "256 MiB memset between steps (inside the timed region)",
                           "multi_gpu": "weak scaling: 128 walkers and 10k candidates per GPU; walkers sharded with one all-gather of W/2 log-probs per half step, candidates sharded with per-theta scalar exchanges"},
                "clocks": clocks,
                "e2e": {"value": evals / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches),
                "roofline": {"kernel": "scale_x + gram_kernel + chol_lml_kernel<8,2> (K1+K2: Gram, Cholesky, LML; 64 thetas, n=500)",
                             "bound": "tensor", "achieved": chol_tf, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": chol_tf / peak_tf, "traffic": traffic,
                             "traffic_note": "dram read+write bytes per launch from one ncu --set full capture (cold L2: ncu "
                                             "flushes caches, in the pipeline the slabs are L2 hits); algorithmic bytes ~30 KB",
                             "peak_source": "measured FP64 DMMA.8x8x4 issue peak on this pool (profiles/fp64_peaks_r01.json; "
                                            "MEASURED_PEAKS.json has no FP64 entry; cuBLAS DGEMM 8192^3 = 35.5)",
                             "launch_ms": chol_ms},
                "roofline_sweep": {"kernel": "sweep_kernel<4> (10 thetas x 10k candidates, n=500)", "bound": "tensor",
                                   "achieved": sweep_tf, "peak": peak_tf, "unit": "TFLOP/s",
                                   "frac": sweep_tf / peak_tf, "launch_ms": sweep_ms}}
        if world == 1 and not args.no_cpu_baseline:
            cyc, detail = cpu_reference_cycle(w, lml_evals=48, sweep_thetas=1, sweep_cands=400)
            line["cpu_baseline"] = {
                "value": w.n_logprob_evals / cyc, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                "sample": "48 log-posterior evals + 1 theta-setter + predict/MES over 400 candidates at n=500, "
                          "extrapolated linearly to the full cycle (1536 evals + 10 x 10000 candidates)",
                "cycle_s_extrapolated": cyc, "detail": detail}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
