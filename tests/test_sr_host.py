"""SR natural gradient (jqmc_b200.sr; reference jqmc/jqmc_mcmc.py:2960-3330): every branch (primal direct / CG, dual direct /
CG) against a plain NumPy solve of the same equations, on one rank and with the samples split over two ``gloo`` ranks."""

import socket

import numpy as np
import pytest
import torch

from jqmc_b200.sr import sr_natural_gradient


def _data(n, K, seed=0):
    rng = np.random.default_rng(seed)
    w = rng.uniform(0.5, 1.5, size=n)
    O = rng.normal(size=(n, K)) * rng.uniform(0.1, 10.0, size=K)[None, :]  # parameters of very different scales
    e = -17.0 + 0.3 * rng.normal(size=n) + 0.05 * O[:, 0] / np.abs(O[:, 0]).max()
    return w, e, O


def _numpy_sr(w, e, O, eps):
    W = w.sum()
    Ob, eb = (w @ O) / W, (w @ e) / W
    X = ((O - Ob) * np.sqrt(w)[:, None] / np.sqrt(W)).T
    F = -2.0 * np.sqrt(w) * (e - eb) / np.sqrt(W)
    dS = np.einsum("kj,kj->k", X, X)
    Xs = X / np.sqrt(dS)[:, None]
    S = Xs @ Xs.T + eps * np.eye(len(dS))
    return np.linalg.solve(S, Xs @ F) / np.sqrt(dS), X @ F


@pytest.mark.parametrize("n,K", [(400, 30), (60, 150)])
def test_sr_branches_single_rank(n, K):
    w, e, O = _data(n, K)
    ref, f_ref = _numpy_sr(w, e, O, 1e-3)
    tw, te, tO = (torch.from_numpy(x) for x in (w, e, O))
    for kw in (dict(), dict(use_cg=True), dict(force_dual=True), dict(force_dual=True, use_cg=True), dict(force_dual=False)):
        theta, info = sr_natural_gradient(tw, te, tO, epsilon=1e-3, cg_tol=1e-13, **kw)
        np.testing.assert_allclose(theta.numpy(), ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max(), err_msg=str((kw, info["method"])))
        np.testing.assert_allclose(info["f"].numpy(), f_ref, rtol=1e-10, atol=1e-13)
    # default branch selection follows the reference: primal if K < samples else dual
    assert sr_natural_gradient(tw, te, tO)[1]["method"] == ("primal-direct" if K < n else "dual-direct")
    # a parameter that never varies is frozen
    O2 = O.copy()
    O2[:, 3] = 0.7
    theta, info = sr_natural_gradient(tw, te, torch.from_numpy(O2), epsilon=1e-3)
    assert info["frozen"] == 1 and theta[3] == 0.0 and torch.isfinite(theta).all()


def _worker(rank, world, port, q):
    import os

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        for n, K in ((400, 30), (60, 150)):
            w, e, O = _data(n, K)
            sl = slice(rank * n // world, (rank + 1) * n // world)
            args = tuple(torch.from_numpy(np.ascontiguousarray(x[sl])) for x in (w, e, O))
            for name, kw in (("direct", {}), ("cg", dict(use_cg=True, cg_tol=1e-13))):
                theta, info = sr_natural_gradient(*args, epsilon=1e-3, **kw)
                out[(n, K, name)] = (theta.numpy().copy(), info["method"])
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sr_two_ranks_gloo():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=250) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for n, K in ((400, 30), (60, 150)):
        ref, _ = _numpy_sr(*_data(n, K), 1e-3)
        for name in ("direct", "cg"):
            for rank in range(2):
                theta, method = res[rank][(n, K, name)]
                assert method.startswith("primal" if K < n else "dual") and method.endswith(name)
                np.testing.assert_allclose(theta, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
