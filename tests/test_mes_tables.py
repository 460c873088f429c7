"""The polynomial tables the MES kernels use (csrc/bgp_mes_table.inc) against scipy, on the CPU: the committed file
is what tools/gen_mes_table.py generates, and a float64 Horner evaluation of it reproduces
gamma phi / (2 Phi) - log Phi and log Phi (bask/acquisition.py:236-267) to 1e-12 over the whole range (the tables themselves are 2e-16 from the 60-digit values; numpy's exp of -gamma^2/2 carries the rest)."""
import os
import re

import numpy as np
from scipy.special import log_ndtr

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(REPO, "bayes-skopt_b200", "csrc", "bgp_mes_table.inc")


def _tables():
    txt = open(INC).read()
    ni = int(re.search(r"#define BGP_MES_TAB_INTERVALS (\d+)", txt).group(1))
    nc = int(re.search(r"#define BGP_MES_TAB_COEFS (\d+)", txt).group(1))
    out = {}
    for name in ("BGP_MES_TAB", "BGP_NLOGCDF_POS_TAB", "BGP_LOGCDF_NEG_TAB"):
        body = re.search(name + r"\[[^\]]*\] = \{(.*?)\};", txt, re.S).group(1)
        vals = np.array([float(v) for v in body.replace("\n", " ").split(",") if v.strip()])
        out[name] = vals.reshape(ni, nc)
    return ni, nc, out


def _horner(tab, u):
    iv = np.minimum((u * 2.0).astype(int), tab.shape[0] - 1)
    x = 4.0 * u - (2 * iv + 1)
    r = tab[iv, 0]
    for j in range(1, tab.shape[1]):
        r = r * x + tab[iv, j]
    return r


def test_tables_reproduce_the_special_functions():
    ni, nc, t = _tables()
    assert (ni, nc) == (76, 13)
    g = np.concatenate([np.linspace(0.0, 37.999, 20001), np.arange(0, 76) * 0.5 + 1e-12, np.arange(1, 77) * 0.5 - 1e-12])
    lcdf = log_ndtr(g)
    term = g * np.exp(-0.5 * g * g - 0.9189385332046727 - lcdf) / 2.0 - lcdf
    got = _horner(t["BGP_MES_TAB"], g) * np.exp(-0.5 * g * g)
    ok = term > 1e-300
    np.testing.assert_allclose(got[ok], term[ok], rtol=1e-12)
    np.testing.assert_allclose(-_horner(t["BGP_NLOGCDF_POS_TAB"], g)[ok] * np.exp(-0.5 * g[ok] ** 2), lcdf[ok], rtol=1e-12)
    np.testing.assert_allclose(-0.5 * g * g + _horner(t["BGP_LOGCDF_NEG_TAB"], g), log_ndtr(-g), rtol=1e-12)


def test_committed_tables_are_the_generated_ones(tmp_path, monkeypatch):
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_mes_table", os.path.join(REPO, "tools", "gen_mes_table.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    rows = gen.table(gen.R, "R")
    _, _, t = _tables()
    assert np.array_equal(np.array(rows), t["BGP_MES_TAB"])


def test_sorted_early_exit_of_the_mes_mean_is_below_rounding():
    """The MES epilogue visits a candidate's max-value draws in ascending order and leaves the loop once a term
    (gamma > 1) is below 1e-18 of the running sum (csrc/bgp_acq.cu).  Restated in numpy: the truncated mean equals
    the full mean of bask/acquisition.py:259-267 to 1e-15 relative over wide and narrow gamma ranges, and the rule
    never fires on a NaN sum."""
    r = np.random.RandomState(3)
    K = 1000
    g = np.sort(-np.log(-np.log(r.uniform(size=K).astype(np.float32))).astype(np.float64))

    def term(gam):
        lc = log_ndtr(gam)
        return gam * np.exp(-0.5 * gam * gam - 0.9189385332046727 - lc) / 2.0 - lc

    skipped_any = False
    for beta, mean, sd in [(0.3, 1.0, 0.01), (1.0, 0.5, 0.05), (0.2, 2.5, 1.5), (0.8, -3.0, 0.3), (0.05, 2.0, 2.0)]:
        gam = (g * beta + 2.0 - mean) / sd
        t = term(gam)
        full = t.sum()
        acc, used = 0.0, 0
        for k in range(K):       # one lane walking all draws (eight lanes each walk a strided subset the same way)
            acc += t[k]
            used += 1
            if gam[k] > 1.0 and t[k] <= 1e-18 * acc:
                break
        skipped_any |= used < K
        assert abs(acc - full) <= 1e-15 * abs(full)
    assert skipped_any
    assert not (1.0 <= 1e-18 * np.nan)     # a NaN sum (the reference's non-finite rows) never satisfies the rule
