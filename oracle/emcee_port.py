"""TEST INFRASTRUCTURE ONLY -- CPU restatement of emcee 3.1.6 (not part of the product).

Restates the published algorithm of ``emcee.EnsembleSampler`` with the default
``StretchMove(a=2.0)`` (Goodman & Weare 2010; emcee 3.1.6 is pinned by the
reference at uv.lock:697-698 and is NOT vendored under /root/reference, nor
installed in this image).  The reference's only call site is
bask/bayesgpr.py:510-530 (construct, inject ``random_state``, ``run_mcmc``,
``get_chain``).

The order in which the ``numpy.random.RandomState`` stream is consumed is part of
the specification: it is what makes the reference's golden tests
(tests/test_acquisition.py:42-70, tests/test_optimizer.py:85-140) reproduce.
Pinned by: tests/test_oracle_ref_golden.py (run in the build container, where
/root/reference exists) which executes the UNMODIFIED reference sources on top of
this module and checks the reference's own golden values.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this.
"""
import numpy as np

__version__ = "3.1.6-restated"


class EnsembleSampler:
    """Affine-invariant ensemble sampler, red/blue stretch move.

    Per step the random stream is consumed in this order (emcee 3.1.6,
    ``EnsembleSampler.sample`` -> ``RedBlueMove.propose`` -> ``StretchMove.get_proposal``):
      1. ``choice(moves, p=weights)``      -- one uniform, even with one move
      2. ``shuffle(arange(W) % 2)``        -- red/blue assignment
      3. for each half: ``rand(Ns)`` (stretch z), ``randint(Nc, size=Ns)`` (partner),
         log-prob of proposals, then one scalar ``rand()`` per walker of the half
         for the accept test ``(p-1) log z + lp_new - lp_old > log u``.
    Accepted rows are written into the live state before the second half moves.
    """

    def __init__(self, nwalkers, ndim, log_prob_fn, kwargs=None, threads=None,
                 vectorize=False, a=2.0, **_ignored):
        self.nwalkers = int(nwalkers)
        self.ndim = int(ndim)
        self.a = float(a)
        self.vectorize = vectorize
        self._fn = log_prob_fn
        self._kwargs = kwargs or {}
        self._random = np.random.mtrand.RandomState()
        self._random.set_state(np.random.get_state())
        self._steps = []
        self._log_prob = []
        self.n_log_prob_evals = 0
        self.naccepted = np.zeros(self.nwalkers)

    @property
    def random_state(self):
        return self._random.get_state()

    @random_state.setter
    def random_state(self, state):
        self._random.set_state(state)

    def compute_log_prob(self, coords):
        if np.any(np.isinf(coords)):
            raise ValueError("At least one parameter value was infinite")
        if np.any(np.isnan(coords)):
            raise ValueError("At least one parameter value was NaN")
        self.n_log_prob_evals += len(coords)
        if self.vectorize:
            lp = np.asarray(self._fn(coords, **self._kwargs), dtype=np.float64)
        else:
            lp = np.array([float(self._fn(coords[i], **self._kwargs))
                           for i in range(len(coords))])
        if np.any(np.isnan(lp)):
            raise ValueError("Probability function returned NaN")
        return lp

    def run_mcmc(self, initial_state, nsteps, progress=False):
        x = np.array(initial_state, dtype=np.float64, copy=True)
        if x.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions")
        if self.nwalkers < 2 * self.ndim:
            raise RuntimeError(
                "It is unadvisable to use a red-blue move with fewer walkers "
                "than twice the number of dimensions.")
        lp = self.compute_log_prob(x)
        rnd = self._random
        everyone = np.arange(self.nwalkers)
        for _ in range(int(nsteps)):
            rnd.choice([0], p=[1.0])
            colour = everyone % 2
            rnd.shuffle(colour)
            for split in (0, 1):
                mine = colour == split
                s = x[mine]
                c = x[~mine]
                ns, nc = len(s), len(c)
                zz = ((self.a - 1.0) * rnd.rand(ns) + 1.0) ** 2.0 / self.a
                factors = (self.ndim - 1.0) * np.log(zz)
                partner = rnd.randint(nc, size=(ns,))
                q = c[partner] - (c[partner] - s) * zz[:, None]
                new_lp = self.compute_log_prob(q)
                movers = everyone[mine]
                accepted = np.zeros(ns, dtype=bool)
                for k in range(ns):
                    lnpdiff = factors[k] + new_lp[k] - lp[movers[k]]
                    accepted[k] = lnpdiff > np.log(rnd.rand())
                idx = movers[accepted]
                x[idx] = q[accepted]
                lp[idx] = new_lp[accepted]
                self.naccepted[idx] += 1
            self._steps.append(x.copy())
            self._log_prob.append(lp.copy())
        return x, lp, self.random_state

    def get_chain(self, flat=False, discard=0, thin=1):
        v = np.array(self._steps).reshape(-1, self.nwalkers, self.ndim)
        v = v[discard + thin - 1::thin]
        return v.reshape(-1, self.ndim) if flat else v

    def get_log_prob(self, flat=False, discard=0, thin=1):
        v = np.array(self._log_prob).reshape(-1, self.nwalkers)
        v = v[discard + thin - 1::thin]
        return v.reshape(-1) if flat else v
