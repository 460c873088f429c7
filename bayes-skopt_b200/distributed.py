"""Multi-GPU orchestration: one process per GPU (torch.distributed for the plumbing; NCCL over NVLink
on the B200 box, gloo in the CPU tests).

What shards:

* the WALKERS of the ensemble MCMC: every rank regenerates the identical red/blue split, proposals and
  accept draws from the shared Philox stream, evaluates the log-posterior of its slice of each half
  step's proposals and stores the results straight into every peer's exchange block (cudaIpc-mapped
  memory, NVLink P2P) -- no NCCL call and no host round trip between propose and accept, so the whole
  run is ONE CUDA graph per rank (``Engine.mcmc_sharded`` -> ``bgp_mcmc_run_sharded``).  When the peer
  mapping is not available the same move runs host-stepped with one NCCL all-gather per half step.
* the CANDIDATES of the acquisition sweep: every rank holds the (tiny) training set, factorises the same
  S sampled thetas, sweeps only its contiguous block of candidates and exchanges per-theta scalars:

  EI        all-reduce(MIN) of the per-theta minimum mean                      S doubles
  TopTwoEI  all-gather of each rank's best (EI, index, mu, sd) per theta       4 S doubles / rank
  MES       all-gather of the (S x m_local) moments; the Gumbel fit is then split over thetas (every
            rank fits its share on all candidates -- the same bits as a one-GPU fit) and the five fit
            parameters per theta are gathered
  all       all-reduce(MAX) of the per-theta "non-finite" flags (bask/acquisition.py:140-141
            skips a theta if ANY candidate is non-finite), then all-gather of the m_local values.

  Every kernel and collective of a sweep is enqueued on the engine's stream; the host synchronises once,
  at the end.

What does not shard: the joint posterior draw behind ThompsonSampling / PVRS (one m x m factorisation)
-- replicas only.

The numerical steps are injected through a small backend protocol so that the orchestration (this file)
is exercised on CPU with gloo, using the oracle as the stand-in compute."""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib

__all__ = ["shard_bounds", "ShardedSweep", "DeviceBackend", "sharded_mcmc", "sharded_mcmc_dev"]


def shard_bounds(m, world, rank):
    """Contiguous block [lo, hi) of rank `rank`: the first m % world ranks get one extra."""
    q, r = divmod(m, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class DeviceBackend:
    """libbgp-backed compute steps on this rank's GPU (tensors live on the engine's device)."""

    def __init__(self, gpr):
        self.gpr, self.e = gpr, gpr._eng()
        self.device = self.e.device

    def stream_ctx(self):
        """collectives are enqueued on the engine's stream so they order with its kernels"""
        return torch.cuda.stream(self.e.stream)

    def _run(self, fn, *args):
        _lib.check(fn(self.e.h, *args, self.e._st), fn.__name__)

    def bind_group(self, group):
        """ShardedSweep tells the backend which ranks it runs over (used to share the factorisations)."""
        self.group = group

    # factor slabs above this many bytes in total no longer sit in the 126 MB L2 while S of them are
    # factorised side by side: the ranks then factorise S / world thetas each and exchange the slabs
    SHARE_FACTORS_ABOVE_BYTES = 96 << 20

    def _factorize_shared(self, th):
        """Factorisations of all S thetas, computed S / world per rank and all-gathered (slabs, z, LML, info)
        over NVLink.  At n = 2000 (config 5) a slab is 50 MB: sixteen of them factorised on one GPU stream
        through HBM (13.8 ms), two stay in L2 (6.7 ms), and the gather of 0.8 GB costs about 1.3 ms -- the
        replicated factorisation is the term that does not shrink with the candidate block."""
        from ._engine import Factor
        e, group = self.e, getattr(self, "group", None)
        world = dist.get_world_size(group) if group is not None else 1
        S = th.shape[0]
        slab = int(e.lib.bgp_factor_slab_doubles(e.h))
        if world == 1 or S < world or 8 * slab * S <= self.SHARE_FACTORS_ABOVE_BYTES:
            return e.factorize(th)
        rank = dist.get_rank(group)
        per = -(-S // world)                                   # thetas per rank, the last ranks may repeat one
        idx = torch.clamp(torch.arange(rank * per, (rank + 1) * per, device=th.device), max=S - 1)
        mine = e.factorize(th.index_select(0, idx).contiguous())
        # three collectives: the slabs (the 0.8 GB one), z, and (LML, info) -- each rank's piece is contiguous
        # in the result, which the sweep reads as plain (S, ...) arrays
        def gather(t):
            out = torch.empty((world * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t.contiguous(), group=group)
            return out[:S]
        misc = gather(torch.stack([mine.lml, mine.info.to(torch.float64)], dim=1))
        return Factor(th, gather(mine.slabs), gather(mine.z), misc[:, 0].contiguous(),
                      misc[:, 1].to(torch.int32).contiguous())

    def moments(self, thetas, X_block):
        e = self.e
        th = thetas if torch.is_tensor(thetas) else e.to_dev(thetas)
        f = self._factorize_shared(th)
        self._pending_info = f.info       # read back once, in finalize(): no host round trip in the sweep
        y_mean = float(np.atleast_1d(self.gpr.y_train_mean_)[0])
        y_std = float(np.atleast_1d(self.gpr.y_train_std_)[0])
        Xd = X_block.contiguous() if torch.is_tensor(X_block) else e.to_dev(X_block)
        mu, sd, _, _ = e.predict(f, Xd, noise_off=True, y_mean=y_mean, y_std=y_std)
        return mu, sd

    def finalize(self):
        """The single host synchronisation of a sweep: positive-definiteness flags of the factorisations."""
        info, self._pending_info = getattr(self, "_pending_info", None), None
        if info is not None and np.any(self.e.to_host(info) != 0):
            raise np.linalg.LinAlgError("The kernel is not returning a positive definite matrix.")

    def min_mu(self, mu, sd):
        S, m = mu.shape
        stats = self.e.empty(S, 4)
        self._run(self.e.lib.bgp_acq_stats, mu.data_ptr(), sd.data_ptr(), S, m, stats.data_ptr())
        return stats[:, 0].contiguous()

    def mes_fit(self, mu_all, sd_all):
        S, m = mu_all.shape
        fit = self.e.empty(S, 5)
        self._run(self.e.lib.bgp_mes_fit, mu_all.data_ptr(), sd_all.data_ptr(), S, m, fit.data_ptr())
        return fit

    def ei_best(self, mu, sd, p0, yopt, index_offset):
        S, m = mu.shape
        ref = self.e.empty(S, 4)
        self._run(self.e.lib.bgp_ei_best, mu.data_ptr(), sd.data_ptr(), S, m, float(p0),
                  None if yopt is None else yopt.data_ptr(), int(index_offset), ref.data_ptr())
        return ref

    def per_theta(self, kind, mu, sd, p0, yopt=None, ref=None, gumbel=None, fit=None):
        S, m = mu.shape
        vals = self.e.empty(S, m)
        skipped = self.e.empty(S, dtype=torch.int32)
        g = None if gumbel is None else (gumbel if torch.is_tensor(gumbel) else self.e.to_dev(gumbel, dtype=torch.float32))
        self._run(self.e.lib.bgp_acq_per_theta, kind, mu.data_ptr(), sd.data_ptr(), S, m, float(p0),
                  None if yopt is None else yopt.data_ptr(), None if ref is None else ref.data_ptr(),
                  None if g is None else g.data_ptr(), 0 if g is None else g.shape[1],
                  None if fit is None else fit.data_ptr(), vals.data_ptr(), skipped.data_ptr())
        return vals, skipped

    def combine(self, vals, skipped):
        S, m = vals.shape
        out = self.e.empty(m)
        self._run(self.e.lib.bgp_acq_combine, vals.data_ptr(), S, m, skipped.data_ptr(), out.data_ptr())
        return out


class ShardedSweep:
    """evaluate the built-in (mu, std) acquisitions over candidates sharded across the ranks of
    `group`.  Every rank passes the same arguments and gets the same (n_acq, m) result."""

    def __init__(self, backend, group=None, keep_on_device=False):
        self.b, self.group, self.keep_on_device = backend, group, keep_on_device
        if hasattr(backend, "bind_group"):
            backend.bind_group(group)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def _allgather_cols(self, t, sizes):
        """t: (S, m_local) -> (S, m) with the ranks' blocks side by side (ragged sizes allowed)."""
        S = t.shape[0]
        mmax = max(sizes)
        pad = torch.zeros(S, mmax, dtype=t.dtype, device=t.device)
        pad[:, : t.shape[1]] = t
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        with self.b.stream_ctx():
            dist.all_gather(parts, pad, group=self.group)
        return torch.cat([p[:, :sz] for p, sz in zip(parts, sizes)], dim=1)

    def _same_on_all_ranks(self, g):
        """The Gumbel variates of MaxValueSearch come from each process's GLOBAL numpy RNG (the reference's
        convention, bask/acquisition.py:253-257), which the ranks have no reason to have seeded alike:
        rank 0's draws are broadcast (S x K float32) so that every candidate block sees the same max-value
        samples."""
        dev = torch.device(getattr(self.b, "device", "cpu"))
        t = g if torch.is_tensor(g) else torch.from_numpy(np.ascontiguousarray(g)).to(dev)
        t = t.contiguous()
        with self.b.stream_ctx():
            dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                           group=self.group)
        return t if dev.type == "cuda" else t.numpy()

    def evaluate(self, X, thetas, acquisitions, gumbels=None):
        """See _evaluate; every tensor op and collective is issued on the backend's stream."""
        with self.b.stream_ctx():
            out = self._evaluate(X, thetas, acquisitions, gumbels)
        fin = getattr(self.b, "finalize", None)
        if fin is not None and not self.keep_on_device:
            fin()
        return out

    def _evaluate(self, X, thetas, acquisitions, gumbels=None):
        """X: (m, d) all candidates (same on every rank); thetas: (S, p) sampled hyper-parameters;
        acquisitions: list of (kind, p0) with kind in _lib.ACQ_*; gumbels: {acq index: (S, K)
        float32}.  Returns a (len(acquisitions), m) float64 numpy array."""
        m = len(X)
        sizes = [shard_bounds(m, self.world, r)[1] - shard_bounds(m, self.world, r)[0] for r in range(self.world)]
        lo, hi = shard_bounds(m, self.world, self.rank)
        mu, sd = self.b.moments(thetas, X[lo:hi])
        S = mu.shape[0]
        out = np.zeros((len(acquisitions), m))
        out_dev = []
        yopt = None
        fit = None
        for j, (kind, p0) in enumerate(acquisitions):
            kw = {}
            if kind in (_lib.ACQ_EI, _lib.ACQ_TTEI) and np.isnan(p0):
                if yopt is None:
                    yopt = self.b.min_mu(mu, sd)
                    with self.b.stream_ctx():
                        dist.all_reduce(yopt, op=dist.ReduceOp.MIN, group=self.group)
                kw["yopt"] = yopt
            if kind == _lib.ACQ_TTEI:
                ref = self.b.ei_best(mu, sd, p0, kw.get("yopt"), lo)
                refs = [torch.empty_like(ref) for _ in range(self.world)]
                with self.b.stream_ctx():
                    dist.all_gather(refs, ref, group=self.group)
                allref = torch.stack(refs)                         # (world, S, 4)
                # global EI maximiser per theta: largest EI, ties -> smallest global index (np.argmax)
                best = allref[0].clone()
                for w in range(1, self.world):
                    cand = allref[w]
                    take = (cand[:, 0] > best[:, 0]) | ((cand[:, 0] == best[:, 0]) & (cand[:, 1] < best[:, 1]))
                    best[take] = cand[take]
                kw["ref"] = best.contiguous()
            if kind == _lib.ACQ_MES:
                if fit is None:
                    # the Gumbel fit needs the moments of ALL candidates, but only per theta: every
                    # rank fits its share of the thetas (same kernel, same data -> the same bits
                    # as a fit on one GPU) and the 5 fit parameters per theta are gathered, so the
                    # fit cost per rank does not grow with the number of ranks
                    mu_all = self._allgather_cols(mu, sizes)
                    sd_all = self._allgather_cols(sd, sizes)
                    S_all = mu_all.shape[0]
                    s_lo, s_hi = shard_bounds(S_all, self.world, self.rank)
                    rows = [shard_bounds(S_all, self.world, r)[1] - shard_bounds(S_all, self.world, r)[0]
                            for r in range(self.world)]
                    rmax = max(rows)
                    mine = torch.zeros(rmax, 5, dtype=mu_all.dtype, device=mu_all.device)
                    if s_hi > s_lo:
                        mine[: s_hi - s_lo] = self.b.mes_fit(mu_all[s_lo:s_hi].contiguous(),
                                                             sd_all[s_lo:s_hi].contiguous())
                    parts = [torch.empty_like(mine) for _ in range(self.world)]
                    with self.b.stream_ctx():
                        dist.all_gather(parts, mine, group=self.group)
                    fit = torch.cat([p_[:n_] for p_, n_ in zip(parts, rows)], dim=0).contiguous()
                kw["fit"] = fit
                kw["gumbel"] = self._same_on_all_ranks(gumbels[j])
            vals, skipped = self.b.per_theta(kind, mu, sd, p0, **kw)
            with self.b.stream_ctx():
                dist.all_reduce(skipped, op=dist.ReduceOp.MAX, group=self.group)
            local = self.b.combine(vals, skipped)
            full = self._allgather_cols(local[None, :], sizes)[0]
            if self.keep_on_device:
                out_dev.append(full)
            else:
                out[j] = full.cpu().numpy()
        return out_dev if self.keep_on_device else out


def sharded_mcmc_dev(engine, pos, n_steps, seed, a, group, buffers=None):
    """Walker-sharded stretch move, device resident: returns the engine's MCMC buffers (pos, lp, chain,
    lpc, acc tensors), identical on every rank.  One CUDA graph per rank with peer stores for the
    log-prob exchange; falls back to the host-stepped NCCL variant when the exchange blocks cannot be
    mapped (no P2P between the GPUs)."""
    e = engine
    if getattr(e, "_no_peer_access", False):
        return _sharded_mcmc_nccl(e, pos, n_steps, seed, a, group)
    try:
        return e.mcmc_sharded(pos, n_steps, seed, group, a=a, buffers=buffers)
    except _lib.BgpError as exc:
        if "cudaIpc" not in str(exc):
            raise
        import warnings
        warnings.warn(f"peer mapping unavailable ({exc}); the sharded MCMC runs host-stepped over NCCL")
        e._no_peer_access = True
        return _sharded_mcmc_nccl(e, pos, n_steps, seed, a, group)


def sharded_mcmc(engine, pos, n_steps, seed, a, group):
    """Host-facing wrapper of ``sharded_mcmc_dev``: (chain_steps (T, W, p), final pos (W, p), acceptance
    counts (W,)) as numpy arrays."""
    b = sharded_mcmc_dev(engine, pos, n_steps, seed, a, group, buffers=getattr(engine, "_mc_buffers_sharded", None))
    engine._mc_buffers_sharded = b
    engine.sync()
    if getattr(engine, "_peers", None) is not None and engine.peer_timed_out():
        raise RuntimeError("a rank of the process group did not answer within 10 s during the sharded MCMC")
    return b["chain"].cpu().numpy(), b["pos"].cpu().numpy(), b["acc"].cpu().numpy()


def _sharded_mcmc_nccl(engine, pos, n_steps, seed, a, group):
    """The same move with the log-probs exchanged by one all-gather of W/2 doubles per half step (host
    stepped: 5 launches + 1 collective per half step)."""
    import ctypes as C
    e = engine
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    W, p = pos.shape
    lib, h, st = e.lib, e.h, e._st
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    sd = C.c_uint64(int(seed) & (2 ** 64 - 1))
    with torch.cuda.stream(e.stream):
        d_pos = (pos if torch.is_tensor(pos) else e.to_dev(pos)).clone()
        lo, hi = shard_bounds(W, world, rank)
        sizes = [shard_bounds(W, world, r)[1] - shard_bounds(W, world, r)[0] for r in range(world)]
        lp_loc, _, _ = e.logprob_dev(d_pos[lo:hi].contiguous())
        d_lp = _allgather_1d(lp_loc, sizes, group)
        colour = e.empty(W, dtype=torch.int32)
        movers = e.empty(W, dtype=torch.int32)
        q, fac = e.empty(W, p), e.empty(W)
        acc = e.zeros(W, dtype=torch.int32)
        chain = e.empty(n_steps, W, p)
        lpc = e.empty(n_steps, W)
        for t in range(n_steps):
            _lib.check(lib.bgp_mcmc_split(h, W, sd, t, P(colour), st), "bgp_mcmc_split")
            for half in (0, 1):
                ns = (W + 1) // 2 if half == 0 else W // 2
                _lib.check(lib.bgp_mcmc_propose(h, P(d_pos), P(colour), W, half, float(a), sd, t, P(q), P(fac),
                                                P(movers), st), "bgp_mcmc_propose")
                lo, hi = shard_bounds(ns, world, rank)
                sizes = [shard_bounds(ns, world, r)[1] - shard_bounds(ns, world, r)[0] for r in range(world)]
                nlp_loc, _, _ = e.logprob_dev(q[lo:hi])
                nlp = _allgather_1d(nlp_loc, sizes, group)
                _lib.check(lib.bgp_mcmc_accept(h, P(d_pos), P(d_lp), P(q), P(fac), P(nlp), P(movers), W, half, sd,
                                               t, P(acc), P(chain[t]) if half == 1 else None,
                                               P(lpc[t]) if half == 1 else None, st), "bgp_mcmc_accept")
                e.launches += 5
    return dict(pos=d_pos, lp=d_lp, chain=chain, lpc=lpc, acc=acc)


def _allgather_1d(t, sizes, group):
    """ragged all-gather of 1-D tensors (sizes[r] elements from rank r) -> concatenation"""
    mmax = max(sizes)
    if min(sizes) == mmax:     # equal shards (the usual case): one collective straight into the result
        out = torch.empty(mmax * len(sizes), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    pad = torch.zeros(mmax, dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p_[:sz] for p_, sz in zip(parts, sizes)]).contiguous()
