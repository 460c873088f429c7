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
    nccl_ok = _c_entry_point_over_a_raw_nccl_communicator(rank, world, gp, g["Xc"][:499])
    q.put((rank, out, single, gp2.chain_, gp3.chain_, out_shared, single9, nccl_ok))
    dist.barrier()
    dist.destroy_process_group()


def _c_entry_point_over_a_raw_nccl_communicator(rank, world, gp, Xc):
    """bgp_acq_sweep_nccl with an ncclComm_t created through NCCL's own C API (ctypes), as a C host would:
    values and argmax must equal the single-GPU bgp_acq_sweep over all candidates, bit for bit."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import bask_b200
    from bask_b200 import _lib

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_char * 128)]

    nccl = C.CDLL("libnccl.so.2")
    uid = UniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=f"cuda:{rank}")
    dist.broadcast(t, src=0)
    C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    e = gp._eng()
    S, K, m = 6, 500, len(Xc)
    th = e.to_dev(gp.chain_[np.random.RandomState(2).choice(len(gp.chain_), S, replace=False)])
    f = e.factorize(th)
    Xd = e.to_dev(Xc)
    g32 = e.to_dev(-np.log(-np.log(np.random.RandomState(4).rand(S, K).astype(np.float32))), dtype=torch.float32)
    mu, sd, _, _ = e.predict(f, Xd, noise_off=True, y_mean=0.25, y_std=1.5)
    ok = True
    for kind in (_lib.ACQ_EI, _lib.ACQ_TTEI, _lib.ACQ_LCB, _lib.ACQ_MEAN, _lib.ACQ_MES):
        p0 = 1.96 if kind == _lib.ACQ_LCB else float("nan")
        ref, _, _, _ = e.acq(kind, mu, sd, p0=p0, gumbel32=g32 if kind == _lib.ACQ_MES else None)
        ref_idx = e.argmax(ref)
        out = e.empty(m)
        idx = e.empty(1, dtype=torch.int64)
        _lib.check(e.lib.bgp_acq_sweep_nccl(e.h, comm, rank, world, kind, th.data_ptr(), S, f.slabs.data_ptr(),
                                            f.z.data_ptr(), Xd.data_ptr(), m, p0,
                                            g32.data_ptr() if kind == _lib.ACQ_MES else None,
                                            K if kind == _lib.ACQ_MES else 0, 0.25, 1.5, out.data_ptr(), idx.data_ptr(),
                                            e._st), "bgp_acq_sweep_nccl")
        a, b = e.to_host(out), e.to_host(ref)
        ok = ok and np.array_equal(a, b) and int(e.to_host(idx)[0]) == int(e.to_host(ref_idx)[0])
    e.sync()
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    return bool(ok)


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
    assert res[0][7] and res[1][7]            # C entry point over a raw ncclComm_t == single GPU
    np.testing.assert_array_equal(res[0][3], res[1][3])          # same chain on both ranks
    np.testing.assert_allclose(res[0][3], res[0][4], rtol=1e-12)  # and the same as the one-GPU graph
