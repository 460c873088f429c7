"""CPU, world_size 2, gloo: the candidate-sharding orchestration (bask_b200/distributed.py) must
reproduce the single-process result.  The numeric steps are supplied by the numpy oracle, which
is exactly what the orchestration is checked against."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Stand-in for DeviceBackend with the same protocol, computing with oracle/*.py on CPU."""

    def __init__(self, g, d):
        from contextlib import nullcontext
        from oracle import acq_oracle as A
        from oracle import gp_oracle as G
        self.A, self.G, self.g = A, G, g
        self.spec = ("sum", ("product", ("const", 1.0, False), ("matern", 0.3 * np.ones(d), 2.5, False)),
                     ("white", 1.0, False))
        self.stream_ctx = nullcontext

    def moments(self, thetas, X_block):
        g, G = self.g, self.G
        mus, sds = [], []
        for t in thetas:
            L, Ki, a = G.factorize(self.spec, t, g["X"], g["y_train"], g["alpha_vec"])
            mu, sd = G.predict(self.spec, t, g["X"], X_block, Ki, a, float(g["y_mean"][0]), float(g["y_std"][0]))
            mus.append(mu)
            sds.append(sd)
        return torch.tensor(np.array(mus)), torch.tensor(np.array(sds))

    def min_mu(self, mu, sd):
        return mu.min(dim=1).values.contiguous()

    def mes_fit(self, mu_all, sd_all):
        rows = []
        for m_, s_ in zip(mu_all.numpy(), sd_all.numpy()):
            a, b, (q1, med, q2) = self.A.mes_gumbel_fit(m_, s_)
            rows.append([a, b, q1, med, q2])
        return torch.tensor(rows)

    def ei_best(self, mu, sd, p0, yopt, index_offset):
        rows = []
        for s in range(mu.shape[0]):
            y = float(yopt[s]) if yopt is not None else (None if np.isnan(p0) else p0)
            ei = self.A.expected_improvement(mu[s].numpy(), sd[s].numpy(), y_opt=y)
            j = int(np.argmax(ei))
            rows.append([ei[j], j + index_offset, float(mu[s, j]), float(sd[s, j])])
        return torch.tensor(rows)

    def per_theta(self, kind, mu, sd, p0, yopt=None, ref=None, gumbel=None, fit=None):
        from bask_b200 import _lib
        import scipy.stats as st
        out = []
        for s in range(mu.shape[0]):
            m_, s_ = mu[s].numpy(), sd[s].numpy()
            if kind == _lib.ACQ_EI:
                v = self.A.expected_improvement(m_, s_, y_opt=float(yopt[s]) if yopt is not None else p0)
            elif kind == _lib.ACQ_TTEI:
                v = np.zeros_like(m_)
                mask = s_ > 0
                outer = np.sqrt(s_[mask] ** 2 + float(ref[s, 3]) ** 2)
                v[mask] = outer * self.A._ei_f((float(ref[s, 2]) - m_[mask]) / outer)
            elif kind == _lib.ACQ_LCB:
                v = self.A.lcb(m_, s_, alpha=p0)
            elif kind == _lib.ACQ_MEAN:
                v = -m_
            else:
                a, b = float(fit[s, 0]), float(fit[s, 1])
                maxv = gumbel[s].astype(np.float64) * b + a
                gam = (maxv[None, :] + m_[:, None]) / s_[:, None]
                with np.errstate(all="ignore"):
                    v = np.sum(gam * st.norm.pdf(gam) / (2 * st.norm.cdf(gam)) - st.norm.logcdf(gam), axis=1) / len(maxv)
            out.append(v)
        vals = torch.tensor(np.array(out))
        skipped = (~torch.isfinite(vals).all(dim=1)).to(torch.int32)
        return vals, skipped

    def combine(self, vals, skipped):
        keep = (skipped == 0).to(vals.dtype)[:, None]
        return (torch.nan_to_num(vals, nan=0.0, posinf=0.0, neginf=0.0) * keep).sum(dim=0) / vals.shape[0]


def _worker(rank, world, port, q, n_theta=3):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bask_b200  # noqa: F401
    from bask_b200 import _lib
    from bask_b200.distributed import ShardedSweep, shard_bounds
    g = dict(np.load(os.path.join(REPO, "tests", "golden", "g1_branin_n20.npz")))
    be = OracleBackend(g, 2)
    X = g["Xc"][:203]          # ragged split: 102 + 101
    thetas = g["thetas"][:n_theta]
    with np.errstate(all="ignore"):
        gum = -np.log(-np.log(g["mes_uniforms"][:n_theta]))
    acqs = [(_lib.ACQ_EI, float("nan")), (_lib.ACQ_TTEI, float("nan")), (_lib.ACQ_LCB, 1.96),
            (_lib.ACQ_MEAN, 0.0), (_lib.ACQ_MES, float("nan"))]
    out = ShardedSweep(be, None).evaluate(X, thetas, acqs, {4: gum})
    lo, hi = shard_bounds(len(X), world, rank)
    q.put((rank, out, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sweep_equals_single_process():
    sys.path.insert(0, REPO)
    from oracle import acq_oracle as A
    from oracle import gp_oracle as G
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][2] == (0, 102) and res[1][2] == (102, 203)
    np.testing.assert_array_equal(res[0][1], res[1][1])      # every rank gets the same answer
    # single-process oracle over all candidates
    g = dict(np.load(os.path.join(REPO, "tests", "golden", "g1_branin_n20.npz")))
    be = OracleBackend(g, 2)
    X, thetas = g["Xc"][:203], g["thetas"][:3]
    mu, sd = be.moments(thetas, X)
    mu, sd = mu.numpy(), sd.numpy()
    with np.errstate(all="ignore"):
        expect = [np.mean([A.expected_improvement(mu[s], sd[s]) for s in range(3)], axis=0),
                  np.mean([A.top_two_ei(mu[s], sd[s]) for s in range(3)], axis=0),
                  np.mean([A.lcb(mu[s], sd[s]) for s in range(3)], axis=0),
                  np.mean([A.expectation(mu[s], sd[s]) for s in range(3)], axis=0),
                  np.mean([A.max_value_search(mu[s], sd[s], uniforms=g["mes_uniforms"][s]) for s in range(3)], axis=0)]
    for j, name in enumerate(["ei", "ttei", "lcb", "mean", "mes"]):
        np.testing.assert_allclose(res[0][1][j], expect[j], rtol=1e-9, atol=1e-300, err_msg=name)
        assert np.argmax(res[0][1][j]) == np.argmax(expect[j])


def test_sharded_sweep_more_ranks_than_thetas():
    """World 3, two thetas: one rank owns no theta of the theta-sharded MES fit and the candidate
    blocks are ragged (68 + 68 + 67); every rank must still return the single-process answer."""
    sys.path.insert(0, REPO)
    from oracle import acq_oracle as A
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 3, port, q, 2)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[2] for r in res] == [(0, 68), (68, 136), (136, 203)]
    np.testing.assert_array_equal(res[0][1], res[1][1])
    np.testing.assert_array_equal(res[0][1], res[2][1])
    g = dict(np.load(os.path.join(REPO, "tests", "golden", "g1_branin_n20.npz")))
    be = OracleBackend(g, 2)
    mu, sd = be.moments(g["thetas"][:2], g["Xc"][:203])
    mu, sd = mu.numpy(), sd.numpy()
    with np.errstate(all="ignore"):
        mes = np.mean([A.max_value_search(mu[s], sd[s], uniforms=g["mes_uniforms"][s]) for s in range(2)], axis=0)
        ei = np.mean([A.expected_improvement(mu[s], sd[s]) for s in range(2)], axis=0)
    np.testing.assert_allclose(res[0][1][4], mes, rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(res[0][1][0], ei, rtol=1e-9, atol=1e-300)


def test_shard_bounds_cover_everything():
    sys.path.insert(0, REPO)
    import bask_b200  # noqa: F401
    from bask_b200.distributed import shard_bounds
    for m in (1, 7, 8, 100000, 100003):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(m, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
