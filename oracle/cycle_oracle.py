"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's MCMC driver and ask/tell tail
(not part of the product; the product never imports this).

Restates ``BayesGPR.sample`` (bask/bayesgpr.py:381-548: walker count, step count, start
ball, RNG hand-over to emcee, chain extraction, geometric-median point estimate, warm start)
on top of oracle/emcee_port.py and oracle/gp_oracle.py, and the candidate/argmax tail of
``Optimizer.tell`` (bask/optimizer.py:353-376).  It is also what ``bench.py`` times on the
host cores as the CPU baseline (kind "port").
"""
import numpy as np

from . import acq_oracle as A
from . import gp_oracle as G
from .emcee_port import EnsembleSampler


def start_ball(theta, noise_level, n_walkers, random_state):
    """bask/bayesgpr.py:501-509: theta + 1e-2 * randn per walker; -inf entries (the zeroed
    White level left behind by fit) become log(noise_)."""
    theta = np.array(theta, dtype=np.float64)
    theta[np.isinf(theta)] = np.log(noise_level)
    return np.array([theta + 1e-2 * random_state.randn(len(theta)) for _ in range(n_walkers)])


def sample(gp: A.GPState, priors, random_state, n_desired_samples=100, n_burnin=0, n_thin=1,
           n_walkers=100, position=None, noise_level=None, add=False, vectorize=False):
    """Returns (pos, log_prob, n_log_prob_evals); updates gp.chain / gp.theta / factors in
    place like the reference does."""
    n_steps = int(np.ceil(n_desired_samples / n_walkers) + n_burnin)
    if position is None:
        position = start_ball(gp.theta, noise_level, n_walkers, random_state)

    if vectorize:
        def fn(T):
            return np.array([G.log_prob(gp.spec, t, gp.X, gp.y, gp.alpha, priors) for t in T])
    else:
        def fn(t):
            return G.log_prob(gp.spec, t, gp.X, gp.y, gp.alpha, priors)

    sampler = EnsembleSampler(n_walkers, np.shape(position)[1], fn, vectorize=vectorize)
    rng = np.random.RandomState(random_state.randint(0, np.iinfo(np.int32).max))
    sampler.random_state = rng.get_state()
    pos, lp, _ = sampler.run_mcmc(position, n_steps)
    chain = sampler.get_chain(flat=True, discard=n_burnin, thin=n_thin)
    gp.chain = np.concatenate([gp.chain, chain]) if (add and gp.chain is not None) else chain
    gp.set_theta(G.geometric_median(gp.chain))
    return pos, lp, sampler.n_log_prob_evals


def ask_tail(Xc, gp: A.GPState, acquisition, n_samples, random_state, **kwargs):
    """bask/optimizer.py:364-376: sweep one acquisition, return (argmax, values)."""
    vals = A.evaluate_acquisitions(Xc, gp, (acquisition,), n_samples=n_samples,
                                   random_state=random_state, **kwargs).flatten()
    return int(np.argmax(vals)), vals
