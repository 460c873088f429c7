"""GPU, 2 ranks over NCCL (needs a box with >= 2 GPUs: `gpurun --gpus 2`): the candidate-sharded
sweep equals the single-GPU sweep, value for value and argmax for argmax."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import bask_b200
    import bench_workloads as W
    g = dict(np.load(os.path.join(REPO, "tests", "golden", "g1_branin_n20.npz")))
    w = W.config1()
    gp = bask_b200.BayesGPR(kernel=bask_b200.construct_default_kernel([0, 1]), normalize_y=True, random_state=0,
                            device=rank)
    gp.fit(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=100, n_burnin=0, n_walkers_per_thread=100,
           progress=False)
    gp.chain_ = g["chain"].copy()
    gp.theta = g["theta_median"]
    acqs = [bask_b200.ExpectedImprovement(), bask_b200.TopTwoEI(), bask_b200.LCB(), bask_b200.Expectation(),
            bask_b200.MaxValueSearch()]
    np.random.seed(w.mes_seed)
    out = bask_b200.evaluate_acquisitions(g["Xc"][:499], gp, acqs, n_samples=10, random_state=1,
                                          process_group=dist.group.WORLD)
    # the same with the factorisations shared (every rank factorises S / world thetas and the slabs are
    # all-gathered): forced here, by default only large problems (config 5) take this route
    from bask_b200.distributed import DeviceBackend
    DeviceBackend.SHARE_FACTORS_ABOVE_BYTES = 0
    np.random.seed(w.mes_seed)
    out_shared = bask_b200.evaluate_acquisitions(g["Xc"][:499], gp, acqs, n_samples=9, random_state=1,
                                                 process_group=dist.group.WORLD)
    DeviceBackend.SHARE_FACTORS_ABOVE_BYTES = 96 << 20
    np.random.seed(w.mes_seed)
    single9 = bask_b200.evaluate_acquisitions(g["Xc"][:499], gp, acqs, n_samples=9, random_state=1)
    np.random.seed(w.mes_seed)
    single = bask_b200.evaluate_acquisitions(g["Xc"][:499], gp, acqs, n_samples=10, random_state=1)
    # walker-sharded MCMC replays the same Philox stream as the single-GPU graph: identical chains
    gp2 = bask_b200.BayesGPR(kernel=bask_b200.construct_default_kernel([0, 1]), normalize_y=True,
                             random_state=5, device=rank)
    gp2.fit(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=200, n_burnin=3, n_walkers_per_thread=100,
            progress=False, process_group=dist.group.WORLD)
    gp3 = bask_b200.BayesGPR(kernel=bask_b200.construct_default_kernel([0, 1]), normalize_y=True,
                             random_state=5, device=rank)
    gp3.fit(w.X, w.y, noise_vector=w.noise_vector, n_desired_samples=200, n_burnin=3, n_walkers_per_thread=100,
            progress=False)
    q.put((rank, out, single, gp2.chain_, gp3.chain_, out_shared, single9))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_gpu():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    np.testing.assert_array_equal(res[0][1], res[1][1])
    for j, name in enumerate(["ei", "ttei", "lcb", "mean", "mes"]):
        np.testing.assert_allclose(res[0][1][j], res[0][2][j], rtol=1e-12, atol=1e-300, err_msg=name)
        assert np.argmax(res[0][1][j]) == np.argmax(res[0][2][j])
        np.testing.assert_allclose(res[0][5][j], res[0][6][j], rtol=1e-12, atol=1e-300, err_msg=name + " (shared factors)")
    np.testing.assert_array_equal(res[0][5], res[1][5])
    np.testing.assert_array_equal(res[0][3], res[1][3])          # same chain on both ranks
    np.testing.assert_allclose(res[0][3], res[0][4], rtol=1e-12)  # and the same as the one-GPU graph
