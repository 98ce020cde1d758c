"""The N > 1 data plane on real hardware: two NCCL ranks (one process per GPU) run the LRDMC driver -- fused projection kernel,
packed one-collective reconfiguration over NVLink, comb, walker copy-out -- and must reproduce the two-rank run emulated
sequentially with plain oracle calls (tests/test_gfmc_host.py::_reference_run, the check the gloo CPU test applies to the
host logic).  Skipped on a box with a single GPU; run it with `gpurun --gpus 2`."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from jqmc_b200.gfmc import GFMC_n
        from tests import test_gfmc_host as T

        H = T._system()
        g = GFMC_n(H, num_walkers=T.NW, num_mcmc_per_measurement=T.NMPM, num_gfmc_collect_steps=1, mcmc_seed=T.SEED, E_scf=T.E_SCF, alat=T.ALAT)
        g.run(T.STEPS)
        q.put((rank, g.bare_w_L.copy(), g.e_L.copy(), g.e_L2.copy(), g.num_survived_walkers, g.latest_r_up_carts.copy(),
               g.latest_r_dn_carts.copy(), g.jax_PRNG_key_list.copy(), g.engine.launch_count()))  # fmt: skip
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_gfmc_n_two_nccl_ranks_match_emulated_oracle_run():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    from tests import test_gfmc_host as T

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in procs:
        res = q.get(timeout=800)
        out[res[0]] = res[1:]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    hist, ranks = T._reference_run(T._system(), 2)
    for r in range(2):
        w, e, e2, ns, ru, rd, keys, launches = out[r]
        assert launches > 0
        np.testing.assert_allclose(w[:, 0], [h[0] for h in hist], rtol=1e-9)
        np.testing.assert_allclose(e[:, 0], [h[1] for h in hist][1:], rtol=1e-8)
        np.testing.assert_allclose(e2[:, 0], [h[2] for h in hist][1:], rtol=1e-8)
        assert ns == sum(h[3] for h in hist)  # branching decisions bit-exact
        np.testing.assert_allclose(ru, ranks[r]["r_up"], rtol=0, atol=1e-10)  # the same walkers survived on the same slots
        np.testing.assert_allclose(rd, ranks[r]["r_dn"], rtol=0, atol=1e-10)
        assert [tuple(int(x) for x in k) for k in keys] == ranks[r]["keys"]
