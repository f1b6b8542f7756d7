"""The N > 1 PRODUCT paths under `pytest -m gpu` on a one-GPU box: two ranks (gloo rendezvous on
127.0.0.1) share cuda:0 and run the package's own CUDA code -- sharded acquisition arg-max
(src/abstractMFGP.py:124-129), restart-parallel fit on the GPU objective
(section 8f rank 1, src/abstractMFGP.py:137), broadcast of the factorised state and sharded MC prediction
(section 8e row 1) -- each compared bit for bit with the single-process result.  (NCCL itself refuses two
ranks on one device; the NCCL runs of the same paths are bench.py --gpus N and tools/dist_check.py.)"""
import socket

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

THETA_C = np.array([1.0, 0.8, 1.0, 0.5, 0.1, 0.3, 1e-3])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    rs = np.random.RandomState(10)
    X_lf, X_hf = rs.uniform(size=(100, 2)), rs.uniform(size=(12, 2))
    cands = np.random.default_rng(0).uniform(size=(200001, 2))
    Xt = np.random.default_rng(3).uniform(size=(4000, 2))
    return X_lf, X_hf, cands, Xt


def _single_process():
    import multifidelity_datafusion_gps_b200 as pkg
    X_lf, X_hf, cands, Xt = _case()
    np.random.seed(7)
    fit = pkg.NARGP(2, util.hf_2d, util.lf_2d)
    fit.fit(X_hf)
    m = pkg.NARGP(2, util.hf_2d, None, lf_X=X_lf, lf_Y=util.lf_2d(X_lf))
    m.lf_model._set_params(np.array([1.5, 0.4, 1e-3]))
    m.fit(X_hf, theta=THETA_C)
    return dict(theta=fit.hf_model.param_array, runs=[f for _, f in fit.hf_model.optimization_runs],
                argmax=m.acquisition_argmax(cands), argmax_fit=fit.acquisition_argmax(cands),
                mc=m.predict_mc(Xt, n_samples=16, seed=5))


def _rank(rank, world, port, q):
    import torch
    import torch.distributed as tdist
    import multifidelity_datafusion_gps_b200 as pkg
    from multifidelity_datafusion_gps_b200 import dist
    torch.cuda.set_device(0)
    tdist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    X_lf, X_hf, cands, Xt = _case()
    np.random.seed(7)                                   # same RNG state on every rank (documented requirement)
    fit = pkg.NARGP(2, util.hf_2d, util.lf_2d)
    fit.parallel_restarts = True
    fit.fit(X_hf)                                       # each rank runs its share of the 6 restarts on the GPU
    m = pkg.NARGP(2, util.hf_2d, None, lf_X=X_lf, lf_Y=util.lf_2d(X_lf))
    if rank == 0:
        m.lf_model._set_params(np.array([1.5, 0.4, 1e-3]))
        m.fit(X_hf, theta=THETA_C)
    m.broadcast_state(src=0)                            # W, alpha of both levels + metadata
    lo, hi = dist.shard_range(len(Xt), rank, world)
    mean, var = m.predict_mc(Xt[lo:hi], n_samples=16, seed=5, m0=lo)
    q.put(dict(rank=rank, theta=fit.hf_model.param_array, runs=[f for _, f in fit.hf_model.optimization_runs],
               argmax=m.acquisition_argmax(cands, distributed=True),
               argmax_fit=fit.acquisition_argmax(cands, distributed=True), lo=lo, hi=hi, mean=mean, var=var))
    tdist.barrier()
    tdist.destroy_process_group()


def test_two_ranks_on_one_gpu_match_the_single_process_results(gpu):
    import torch.multiprocessing as mp
    ref = _single_process()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in res:
        assert np.array_equal(r["theta"], ref["theta"]) and r["runs"] == ref["runs"]     # restart-parallel == serial
        assert r["argmax"] == ref["argmax"] and r["argmax_fit"] == ref["argmax_fit"]       # sharded == single GPU
        assert np.array_equal(r["mean"], ref["mc"][0][r["lo"]:r["hi"]])                    # shard of the MC sweep
        assert np.array_equal(r["var"], ref["mc"][1][r["lo"]:r["hi"]])
    assert sorted((r["lo"], r["hi"]) for r in res) == [(0, 2000), (2000, 4000)]
