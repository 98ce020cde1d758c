"""Host-side LRDMC driver (jqmc_b200.gfmc.GFMC_n) on CPU: the step loop, the per-step reductions and the walker
reconfiguration run through the real driver with an oracle-backed engine double, single process and with two
``gloo`` ranks (the N > 1 path: all_reduce of the weighted sums, all_gather of weights and coordinates, identical
comb on every rank).  Reference behaviour: jqmc/jqmc_gfmc.py:5774-6417; the reference's own two-rank test is
tests/test_jqmc_gfmc_bra.py run under ``mpirun -np 2``."""

import os
import socket

import numpy as np
import pytest
import torch

from jqmc_b200 import rng_host
from jqmc_b200.data import Jastrow_data, Jastrow_two_body_data
from jqmc_b200.gfmc import GFMC_n, GFMC_t, compute_G_L, jackknife_E_scf
from jqmc_b200.mcmc import generate_init_electron_configurations
from oracle import drivers as OD
from tests.conftest import load_system
from tests.oracle_engine import OracleEngine

NW, NMPM, STEPS, SEED, ALAT, E_SCF = 2, 2, 3, 3446, 0.3, -1.0


def _system():
    H = load_system("H2_ae_ccpvdz_cart")
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.75))
    return H


def _reference_run(H, world):
    """The same branching steps for `world` ranks emulated sequentially with plain oracle calls."""
    gem, cp = H.wavefunction_data.geminal_data, H.coulomb_potential_data
    ranks = []
    for r in range(world):
        seed = SEED * (r + 1)
        keys = [tuple(int(x) for x in k) for k in rng_host.split(rng_host.PRNGKey(seed), NW)]
        np.random.seed(seed)
        r_up, r_dn, _, _ = generate_init_electron_configurations(
            gem.num_electron_up, gem.num_electron_dn, NW, cp.effective_charges, H.structure_data.positions
        )
        ranks.append(dict(keys=keys, r_up=r_up, r_dn=r_dn))
    zeta_rng = np.random.RandomState(SEED)
    hist = []
    for _ in range(STEPS):
        W, S = [], np.zeros(5)
        for st in ranks:
            w_new, Vd, Vn = [], [], []
            for i in range(NW):
                _, Ginv = OD.geminal_inv(gem, st["r_up"][i], st["r_dn"][i])
                w, ru, rd, _, key, RT, _, _ = OD.lrdmc_projection(
                    H, 1.0, st["r_up"][i], st["r_dn"][i], Ginv, st["keys"][i], E_SCF, NMPM, True, "tmove", ALAT
                )
                st["r_up"][i], st["r_dn"][i], st["keys"][i] = ru, rd, key
                d, n = OD.lrdmc_V_elements(H, ru, rd, RT, "tmove", ALAT)
                w_new.append(w), Vd.append(d), Vn.append(n)
            W.append(np.array(w_new))
            S += OD.lrdmc_collect(w_new, Vd, Vn, E_SCF)
        chosen, ns = OD.lrdmc_branch_indices(W, zeta_rng.random_sample())
        up_all = np.concatenate([st["r_up"] for st in ranks])
        dn_all = np.concatenate([st["r_dn"] for st in ranks])
        for r, st in enumerate(ranks):
            st["r_up"] = up_all[chosen[r * NW : (r + 1) * NW]].copy()
            st["r_dn"] = dn_all[chosen[r * NW : (r + 1) * NW]].copy()
        hist.append((S[1] / S[0], S[3] / S[2], S[4] / S[2], ns))
    return hist, ranks


def _driver_run(H):
    g = GFMC_n(H, num_walkers=NW, num_mcmc_per_measurement=NMPM, num_gfmc_collect_steps=1, mcmc_seed=SEED, E_scf=E_SCF, alat=ALAT,
               engine=OracleEngine(H))  # fmt: skip
    g.run(STEPS)
    return g


def _check(g, hist, ranks, rank):
    np.testing.assert_allclose(g.bare_w_L[:, 0], [h[0] for h in hist], rtol=1e-12)
    np.testing.assert_allclose(g.e_L[:, 0], [h[1] for h in hist][1:], rtol=1e-12)
    np.testing.assert_allclose(g.e_L2[:, 0], [h[2] for h in hist][1:], rtol=1e-12)
    assert g.num_survived_walkers == sum(h[3] for h in hist)
    np.testing.assert_array_equal(g.latest_r_up_carts, ranks[rank]["r_up"])
    np.testing.assert_array_equal(g.latest_r_dn_carts, ranks[rank]["r_dn"])
    assert [tuple(int(x) for x in k) for k in g.jax_PRNG_key_list] == ranks[rank]["keys"]


def test_gfmc_n_single_rank():
    H = _system()
    hist, ranks = _reference_run(H, 1)
    g = _driver_run(H)
    _check(g, hist, ranks, 0)
    assert g.mcmc_counter == STEPS - 1 and g.w_L.shape == (STEPS - 1, 1)


def _worker(rank, world, port, q, kind="n"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H = _system()
        g = _driver_run(H) if kind == "n" else _driver_run_t(H)
        q.put((rank, g.bare_w_L.copy(), g.e_L.copy(), g.e_L2.copy(), g.num_survived_walkers, g.latest_r_up_carts.copy(),
               g.latest_r_dn_carts.copy(), g.jax_PRNG_key_list.copy()))  # fmt: skip
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("kind", ["n", "t"])
def test_gfmc_two_ranks_gloo(kind):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, kind)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in procs:
        res = q.get(timeout=500)
        out[res[0]] = res[1:]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    H = _system()
    hist, ranks = _reference_run(H, 2) if kind == "n" else _reference_run_t(H, 2)
    for rank in range(2):
        w, e, e2, nsv, ru, rd, keys = out[rank]
        np.testing.assert_allclose(w[:, 0], [h[0] for h in hist], rtol=1e-12)
        np.testing.assert_allclose(e[:, 0], [h[1] for h in hist][1:], rtol=1e-12)
        np.testing.assert_allclose(e2[:, 0], [h[2] for h in hist][1:], rtol=1e-12)
        assert nsv == sum(h[3] for h in hist)
        np.testing.assert_array_equal(ru, ranks[rank]["r_up"])
        np.testing.assert_array_equal(rd, ranks[rank]["r_dn"])
        assert [tuple(int(x) for x in k) for k in keys] == ranks[rank]["keys"]


# ---- GFMC_t (jqmc/jqmc_gfmc.py:646-2391) ---------------------------------------------------------------------------------
TAU = 0.02


def _reference_run_t(H, world):
    """GFMC_t branching steps for `world` ranks emulated sequentially with plain oracle calls: the projection loop of a
    rank runs all of its walkers until the slowest one is out of time."""
    gem, cp = H.wavefunction_data.geminal_data, H.coulomb_potential_data
    ranks = []
    for r in range(world):
        seed = SEED * (r + 1)
        keys = np.array(rng_host.split(rng_host.PRNGKey(seed), NW), dtype=np.uint32)
        np.random.seed(seed)
        r_up, r_dn, _, _ = generate_init_electron_configurations(
            gem.num_electron_up, gem.num_electron_dn, NW, cp.effective_charges, H.structure_data.positions
        )
        ranks.append(dict(keys=keys, r_up=r_up, r_dn=r_dn, pc=[]))
    zeta_rng = np.random.RandomState(SEED)
    hist = []
    for _ in range(STEPS):
        W, S = [], np.zeros(4)
        for st in ranks:
            Ginv = np.array([OD.geminal_inv(gem, st["r_up"][i], st["r_dn"][i])[1] for i in range(NW)])
            e, pc, w, ru, rd, _, k2, _, _ = OD.lrdmc_projection_t_loop(H, np.ones(NW), st["r_up"], st["r_dn"], Ginv, st["keys"], TAU, True, "tmove", ALAT)
            st["r_up"], st["r_dn"], st["keys"] = ru, rd, k2
            st["pc"].append(np.mean(pc))
            W.append(w)
            S += OD.lrdmc_collect_t(w, e)
        chosen, ns = OD.lrdmc_branch_indices(W, zeta_rng.random_sample())
        up_all = np.concatenate([st["r_up"] for st in ranks])
        dn_all = np.concatenate([st["r_dn"] for st in ranks])
        for r, st in enumerate(ranks):
            st["r_up"] = up_all[chosen[r * NW : (r + 1) * NW]].copy()
            st["r_dn"] = dn_all[chosen[r * NW : (r + 1) * NW]].copy()
        hist.append((S[1] / S[0], S[2] / S[1], S[3] / S[1], ns))
    for st in ranks:
        st["keys"] = [tuple(int(x) for x in k) for k in st["keys"]]
    return hist, ranks


def _driver_run_t(H):
    g = GFMC_t(H, num_walkers=NW, num_gfmc_collect_steps=1, mcmc_seed=SEED, tau=TAU, alat=ALAT, engine=OracleEngine(H))
    g.run(STEPS)
    return g


def test_gfmc_t_single_rank():
    H = _system()
    hist, ranks = _reference_run_t(H, 1)
    g = _driver_run_t(H)
    _check(g, hist, ranks, 0)
    np.testing.assert_allclose(g.average_projection_counter, ranks[0]["pc"])
    assert g.mcmc_counter == STEPS - 1 and g.w_L.shape == (STEPS - 1, 1) and g.tau == TAU
    assert np.all(g.average_projection_counter >= 1.0)


def test_projection_t_oracle_properties():
    """The continuous-time projection of the oracle: time bookkeeping, the no-move rule at tau_left <= 0 and the key
    schedule of the vmapped while_loop (every walker splits three keys per iteration until the slowest is done)."""
    H = _system()
    gem = H.wavefunction_data.geminal_data
    rng = np.random.default_rng(3)
    nw = 3
    r_up = rng.normal(scale=0.8, size=(nw, 1, 3))
    r_dn = rng.normal(scale=0.8, size=(nw, 1, 3))
    Ginv = np.array([OD.geminal_inv(gem, r_up[i], r_dn[i])[1] for i in range(nw)])
    keys = np.array([[0, 5 + i] for i in range(nw)], dtype=np.uint32)
    trace = [[] for _ in range(nw)]
    e, pc, w, ru, rd, gi, k2, RT, n_it = OD.lrdmc_projection_t_loop(H, np.ones(nw), r_up, r_dn, Ginv, keys, 0.05, True, "tmove", ALAT, trace)
    assert n_it == max(pc) and min(pc) >= 1
    for i in range(nw):
        tr = trace[i]
        assert len(tr) == n_it
        np.testing.assert_allclose(sum(t["tau_update"] for t in tr), 0.05, rtol=1e-12)
        assert [t["moved"] for t in tr] == [True] * (pc[i] - 1) + [False] * (n_it - pc[i] + 1)
        np.testing.assert_allclose(w[i], np.exp(-sum(t["tau_update"] * t["e_L"] for t in tr)), rtol=1e-12)
        key = (int(keys[i, 0]), int(keys[i, 1]))
        for _ in range(3 * n_it):
            key, _sub = rng_host_split(key)
        assert tuple(int(x) for x in k2[i]) == key
        # the inverse carried by Sherman-Morrison is the inverse at the final configuration
        np.testing.assert_allclose(gi[i], OD.geminal_inv(gem, ru[i], rd[i])[1], rtol=1e-8, atol=1e-10)
    # fixed mesh: rotation is the identity
    out = OD.lrdmc_projection_t_loop(H, np.ones(nw), r_up, r_dn, Ginv, keys, 0.01, False, "tmove", ALAT)
    np.testing.assert_array_equal(out[7], np.broadcast_to(np.eye(3), (nw, 3, 3)))


def rng_host_split(key):
    from oracle import jaxrng as R

    return R.split(key)


def test_G_L_and_E_scf_jackknife():
    rng = np.random.default_rng(0)
    w = rng.uniform(0.9, 1.1, size=(40, 1))
    G = compute_G_L(w, 5)
    assert G.shape == (35, 1)
    np.testing.assert_allclose(G[0, 0], np.prod(w[0:5, 0]))
    np.testing.assert_allclose(G[-1, 0], np.prod(w[34:39, 0]))
    e = rng.normal(-17.0, 0.1, size=35)
    E, s = jackknife_E_scf(G[:, 0], G[:, 0] * e, 10)
    assert abs(E - np.sum(G[:, 0] * e) / np.sum(G[:, 0])) < 1e-3 and 0 < s < 0.1


def test_branch_oracle_properties():
    rng = np.random.default_rng(1)
    w = [rng.uniform(0.5, 1.5, size=64) for _ in range(3)]
    chosen, ns = OD.lrdmc_branch_indices(w, 0.37)
    assert chosen.shape == (192,) and np.all(np.diff(chosen) >= 0) and 0 <= chosen.min() and chosen.max() < 192
    assert ns == len(np.unique(chosen))
    # equal weights: the comb picks every walker exactly once
    chosen, ns = OD.lrdmc_branch_indices([np.ones(8), np.ones(8)], 0.5)
    np.testing.assert_array_equal(chosen, np.arange(16))
    # a walker with (almost) all the weight takes every slot
    ww = np.full(8, 1e-14)
    ww[5] = 1.0
    chosen, ns = OD.lrdmc_branch_indices([ww], 0.2)
    assert ns == 1 and np.all(chosen == 5)
    assert isinstance(torch.tensor(chosen), torch.Tensor)


# ---- atomic forces in the LRDMC drivers (jqmc_gfmc.py:5840-6051, 6694-6990) -------------------------------------------------
def _force_driver(kind, H, deriv, **kw):
    if kind == "n":
        return GFMC_n(H, num_walkers=3, num_mcmc_per_measurement=NMPM, num_gfmc_collect_steps=1, mcmc_seed=SEED, E_scf=E_SCF, alat=ALAT,
                      comput_position_deriv=deriv, engine=OracleEngine(H), **kw)  # fmt: skip
    return GFMC_t(H, num_walkers=3, num_gfmc_collect_steps=1, mcmc_seed=SEED, tau=TAU, alat=ALAT, comput_position_deriv=deriv,
                  engine=OracleEngine(H), **kw)  # fmt: skip


@pytest.mark.parametrize("kind", ["n", "t"])
def test_gfmc_force_histories(kind, tmp_path):
    """The force terms ride on the chain without changing it; with SWCT the per-step sums over the atoms of F_HF and F_PP
    vanish (translation invariance: sum_alpha omega_alpha = 1), the stored averages equal a direct per-walker evaluation at the
    pre-branching walkers, and the histories survive a checkpoint round trip."""
    from jqmc_b200 import checkpoint
    from jqmc_b200.forces import ForceEvaluator

    H = _system()
    plain = _force_driver(kind, H, False)
    plain.run(4)
    g = _force_driver(kind, H, True, use_swct=True, epsilon_PW=0.4)
    seen = []
    fs0 = g._force_sums

    def spy(r_up, r_dn, RTs, weight, e_L, world):  # record the arguments of every step
        seen.append([x.detach().cpu().numpy().copy() for x in (r_up, r_dn, RTs, weight, e_L)])
        return fs0(r_up, r_dn, RTs, weight, e_L, world)

    g._force_sums = spy
    g.run(4)
    np.testing.assert_array_equal(g.bare_w_L, plain.bare_w_L)
    np.testing.assert_array_equal(g.e_L, plain.e_L)
    np.testing.assert_array_equal(g.latest_r_up_carts, plain.latest_r_up_carts)
    assert g.force_HF.shape == (3, 1, 2, 3) and g.force_PP.shape == (3, 1, 2, 3) and g.E_L_force_PP.shape == (3, 1, 2, 3)
    scale = np.abs(g.force_HF).max()
    assert np.isfinite(scale) and scale > 0
    np.testing.assert_allclose(g.force_HF.sum(axis=2), 0.0, atol=2e-5 * max(1.0, scale))
    np.testing.assert_allclose(g.force_PP.sum(axis=2), 0.0, atol=2e-6 * max(1.0, np.abs(g.force_PP).max()))
    # direct evaluation of the last step: per-walker products, Pathak-Wagner factor and weighted averages written out
    r_up, r_dn, RTs, weight, e_L = seen[-1]
    fe = ForceEvaluator(H, OracleEngine(H), lattice=(ALAT, "tmove"))
    f_hf, f_pp, e0 = (x.numpy() for x in fe.force_products(r_up, r_dn, RTs, True))
    if kind == "n":
        np.testing.assert_allclose(e0, e_L, rtol=1e-12)  # V_diag + V_nondiag at the base point
    gl = fe._last
    gn2 = (gl["dln_Psi_dr_up"].numpy() ** 2).sum(axis=(1, 2)) + (gl["dln_Psi_dr_dn"].numpy() ** 2).sum(axis=(1, 2))
    t = 1.0 / np.sqrt(gn2) / 0.4
    f_eps = np.where(t < 1.0, 7.0 * t**6 - 15.0 * t**4 + 9.0 * t**2, 1.0)
    assert np.any(t < 1.0) or np.all(f_eps == 1.0)
    den = weight.sum()
    np.testing.assert_allclose(g.force_HF[-1, 0], np.einsum("i,ijk->jk", weight * f_eps, f_hf) / den, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(g.force_PP[-1, 0], np.einsum("i,ijk->jk", weight * f_eps, f_pp) / den, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(g.E_L_force_PP[-1, 0], np.einsum("i,ijk->jk", weight * f_eps * e_L, f_pp) / den, rtol=1e-10, atol=1e-12)
    # checkpoint round trip keeps the histories and the switch
    full = str(tmp_path / "restart.h5")
    checkpoint.save_checkpoint(g, full, tmp_pattern=str(tmp_path / "._restart_rank{rank}.h5"))
    g2 = type(g).load_from_hdf5(full, rank=0, engine=OracleEngine(H))
    assert g2.comput_position_deriv and g2._forces is not None
    np.testing.assert_array_equal(g2.force_HF, g.force_HF)
    np.testing.assert_array_equal(g2.E_L_force_PP, g.E_L_force_PP)


def test_gfmc_get_aF_matches_definition():
    """get_aF against the estimator written out bin by bin (jqmc_gfmc.py:6712-6812): leave-one-bin-out ratios of G_L-weighted sums."""
    H = _system()
    g = _force_driver("n", H, False)
    rng = np.random.default_rng(5)
    A, c, nb, warm = 31, 1, 5, 3
    g._num_gfmc_collect_steps = c
    g._mcmc_counter = A
    g._stored_w_L = rng.uniform(0.8, 1.2, size=(A, 1))
    g._stored_e_L = rng.normal(-1.1, 0.1, size=(A, 1))
    g._stored_force_HF = rng.normal(size=(A, 1, 2, 3))
    g._stored_force_PP = rng.normal(size=(A, 1, 2, 3))
    g._stored_E_L_force_PP = g._stored_e_L[..., None, None] * g._stored_force_PP + 0.01 * rng.normal(size=(A, 1, 2, 3))
    mean, std = g.get_aF(num_mcmc_warmup_steps=warm, num_mcmc_bin_blocks=nb)
    G = compute_G_L(g._stored_w_L, c)[warm:, 0]
    e, fh, fp, ef = (x[c:][warm:, 0] for x in (g._stored_e_L, g._stored_force_HF, g._stored_force_PP, g._stored_E_L_force_PP))
    idx = np.array_split(np.arange(len(G)), nb)
    est = []
    for j in range(nb):
        keep = np.concatenate([idx[k] for k in range(nb) if k != j])
        W = G[keep].sum()
        hf = np.einsum("i,ijk->jk", G[keep], fh[keep]) / W
        pp = np.einsum("i,ijk->jk", G[keep], fp[keep]) / W
        epp = np.einsum("i,ijk->jk", G[keep], ef[keep]) / W
        E = (G[keep] * e[keep]).sum() / W
        est.append(-hf - 2.0 * (epp - E * pp))
    est = np.array(est)
    np.testing.assert_allclose(mean, est.mean(axis=0), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(std, np.sqrt((nb - 1) * ((est - est.mean(axis=0)) ** 2).mean(axis=0)), rtol=1e-10)
    with pytest.raises(ValueError):
        _force_driver("n", H, False).get_aF()
