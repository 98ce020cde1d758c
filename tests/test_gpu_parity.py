"""GPU parity tests: every kernel family of SURVEY.md §8(a) through the C ABI against the CPU oracle on
the same seeded inputs.  Tolerances: fp64, 1e-10 relative (north_star) unless a looser bound is argued
in the test; integer results (accept/reject counts, keys) bit-exact."""

import copy

import numpy as np
import pytest

from jqmc_b200.data import Jastrow_data, Jastrow_one_body_data, Jastrow_two_body_data
from oracle import drivers as OD
from oracle import jaxrng as R
from oracle import physics as P
from tests.conftest import load_system, random_walkers

pytestmark = pytest.mark.gpu

RTOL = 1e-10
SYSTEMS = ["water_ccecp_ccpvqz", "N2_ecp_ccpvtz_cart", "H2_ae_ccpvdz_cart", "Li_ae_ccpvdz_cart", "H2_ecp_ccpvtz", "H_ecp_ccpvqz"]


def _engine(H, **kw):
    from jqmc_b200.engine import WalkerEngine

    return WalkerEngine(H, **kw)


def _with_jastrow(H, kind):
    H = copy.deepcopy(H)
    cp = H.coulomb_potential_data
    core = tuple(cp.z_cores) if cp.ecp_flag else tuple(0 for _ in H.structure_data.atomic_numbers)
    if kind == "j2pade":
        jd = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.8, jastrow_2b_type="pade"))
    elif kind == "j1exp_j2exp":
        jd = Jastrow_data(
            jastrow_one_body_data=Jastrow_one_body_data(jastrow_1b_param=0.9, jastrow_1b_type="exp", structure_data=H.structure_data, core_electrons=core),
            jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.6, jastrow_2b_type="exp"),
        )
    elif kind == "j1pade_j2pade":
        jd = Jastrow_data(
            jastrow_one_body_data=Jastrow_one_body_data(jastrow_1b_param=1.3, jastrow_1b_type="pade", structure_data=H.structure_data, core_electrons=core),
            jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=1.1, jastrow_2b_type="pade"),
        )
    else:
        jd = Jastrow_data()
    H.wavefunction_data.jastrow_data = jd
    return H


@pytest.mark.parametrize("name", SYSTEMS)
def test_ao_and_mo_value_grad_lap(name):
    """kernels 1-2: AO / MO value, gradient, Laplacian (a2-a4) vs the closed-form oracle."""
    H = load_system(name)
    eng = _engine(H)
    mos = H.wavefunction_data.geminal_data.orb_data_up_spin
    rng = np.random.default_rng(1)
    Rn = np.asarray(H.structure_data.positions)
    r = Rn[rng.integers(0, len(Rn), 40)] + rng.normal(scale=0.9, size=(40, 3))
    r[0] = Rn[0]  # a point exactly on a nucleus
    ref_ao = np.stack(P.compute_AOs_value_grad_lap(mos.aos_data, r))
    got_ao = eng.eval_orbitals("up", "ao", r).cpu().numpy()
    scale = np.abs(ref_ao).max(axis=(1, 2), keepdims=True)
    np.testing.assert_allclose(got_ao, ref_ao, rtol=RTOL, atol=1e-12 * scale.max())
    ref_mo = np.stack(P.compute_orb_value_grad_lap(mos, r))
    got_mo = eng.eval_orbitals("up", "orb", r).cpu().numpy()
    np.testing.assert_allclose(got_mo, ref_mo, rtol=RTOL, atol=1e-12 * np.abs(ref_mo).max())


@pytest.mark.parametrize("name", SYSTEMS)
def test_geminal_inverse_and_ln_wavefunction(name):
    """kernel 3 (a5-a7, a21): G, Ginv, ln|Psi|."""
    H = _with_jastrow(load_system(name), "j1exp_j2exp")
    eng = _engine(H)
    nw = 6
    r_up, r_dn = random_walkers(H, nw, 21)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    ln, sg = eng.ln_wavefunction(r_up, r_dn)
    G, Ginv, ln, sg = (x.cpu().numpy() for x in (G, Ginv, ln, sg))
    gem = H.wavefunction_data.geminal_data
    for w in range(nw):
        Gr = P.compute_geminal_all_elements(gem, r_up[w], r_dn[w])
        np.testing.assert_allclose(G[w], Gr, rtol=RTOL, atol=1e-13 * np.abs(Gr).max())
        Gir = P.geminal_inv_svd(Gr)
        cond = np.linalg.cond(Gr)
        np.testing.assert_allclose(Ginv[w], Gir, rtol=0, atol=1e-14 * cond * np.abs(Gir).max())
        ref = P.evaluate_ln_wavefunction(H.wavefunction_data, r_up[w], r_dn[w])
        np.testing.assert_allclose(ln[w], ref, rtol=RTOL, atol=1e-11)
        assert sg[w] == np.sign(np.linalg.det(Gr))


@pytest.mark.parametrize("name,jas", [("water_ccecp_ccpvqz", "none"), ("water_ccecp_ccpvqz", "j2pade"), ("water_ccecp_ccpvqz", "j1exp_j2exp"),
                                      ("N2_ecp_ccpvtz_cart", "j1pade_j2pade"), ("H2_ae_ccpvdz_cart", "j1exp_j2exp"),
                                      ("Li_ae_ccpvdz_cart", "j2pade"), ("H2_ecp_ccpvtz", "j1pade_j2pade"),
                                      ("H_ecp_ccpvqz", "j1exp_j2exp")])  # fmt: skip
def test_local_energy_and_parts(name, jas):
    """kernels 4-5 (a10, a13-a16, a19, a23-a27): e_L, per-electron kinetic energies, potential pieces."""
    H = _with_jastrow(load_system(name), jas)
    eng = _engine(H)
    nw = 5
    r_up, r_dn = random_walkers(H, nw, 33)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    keys = np.array([[3, 100 + i] for i in range(nw)], dtype=np.uint32)
    RT = eng.generate_RTs(keys)
    e_L, T, V = eng.e_L_fast(r_up, r_dn, RT, Ginv, return_parts=True)
    e_L, T, V, RT, Ginv = (x.cpu().numpy() for x in (e_L, T, V, RT, Ginv))
    wf, cp = H.wavefunction_data, H.coulomb_potential_data
    for w in range(nw):
        RTw = OD.generate_rotation_matrix((3, 100 + w))
        np.testing.assert_allclose(RT[w], RTw, rtol=0, atol=1e-14)
        Tu, Td = P.compute_kinetic_energy_all_elements(wf, r_up[w], r_dn[w], Ginv[w])
        Tref = np.concatenate([Tu, Td])
        np.testing.assert_allclose(T[w], Tref, rtol=RTOL, atol=1e-10 * np.abs(Tref).max())
        vb = P.compute_bare_coulomb_potential(cp, r_up[w], r_dn[w])
        np.testing.assert_allclose(V[w, 0], vb, rtol=RTOL)
        if cp.ecp_flag:
            np.testing.assert_allclose(V[w, 1], P.compute_ecp_local_parts(cp, r_up[w], r_dn[w]), rtol=RTOL, atol=1e-12)
            vnl = P.compute_ecp_non_local_parts_nearest_neighbors(cp, wf, r_up[w], r_dn[w], RTw, NN=1, Nv=6, Ginv=Ginv[w])[3]
            np.testing.assert_allclose(V[w, 2], vnl, rtol=1e-9, atol=1e-11)
        ref = P.compute_local_energy(H, r_up[w], r_dn[w], RTw, Ginv=Ginv[w])
        np.testing.assert_allclose(e_L[w], ref, rtol=RTOL, atol=1e-10 * np.abs(Tref).max())


def test_local_energy_turborvb_golden(water):
    """The GPU local energy reproduces the TurboRVB known answers of the reference's own test
    (tests/test_comparison_with_turborvb_ECP.py:105-129, 231-274) directly."""
    from tests.test_oracle_golden import DN_A, DN_B, NEW_UP2, UP_A, UP_B

    for H, up, dn, kin, vpot in (
        (water, UP_A, DN_A, 14.6961809426982, -17.0152290468758 + 0.328893830058865),
        (_with_jastrow(water, "none"), UP_A, DN_A, 14.6961809426982, -17.0152290468758 + 0.328893830058865),
    ):
        eng = _engine(H)
        new_up = up.copy()
        new_up[2] = NEW_UP2
        G, Ginv = eng.geminal_inv_batched(new_up[None], dn[None])
        e_L, T, V = eng.e_L_fast(new_up[None], dn[None], np.eye(3)[None], Ginv, return_parts=True)
        np.testing.assert_almost_equal(T.sum().item(), kin, decimal=6)
        np.testing.assert_almost_equal(V[0, :3].sum().item(), vpot, decimal=5)
    H = copy.deepcopy(water)
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.676718854150191))
    eng = _engine(H)
    new_up = UP_B.copy()
    new_up[2] = NEW_UP2
    G, Ginv = eng.geminal_inv_batched(new_up[None], DN_B[None])
    e_L, T, V = eng.e_L_fast(new_up[None], DN_B[None], np.eye(3)[None], Ginv, return_parts=True)
    np.testing.assert_almost_equal(T.sum().item(), 11.1237599317225, decimal=6)
    np.testing.assert_almost_equal(V[0, :3].sum().item(), -27.03387193107 + 0.243517439611676, decimal=5)
    # WF ratio^2 of the golden move through the move-ratio entry
    G0, Ginv0 = eng.geminal_inv_batched(UP_B[None], DN_B[None])
    dr, jr = eng.move_ratios(UP_B[None], DN_B[None], Ginv0, [2], np.array(NEW_UP2)[None, None, :])
    np.testing.assert_almost_equal(((dr * jr) ** 2).item(), 0.881124604511419, decimal=6)


@pytest.mark.parametrize("name,jas", [("water_ccecp_ccpvqz", "j1exp_j2exp"), ("Li_ae_ccpvdz_cart", "j1pade_j2pade"), ("N2_ecp_ccpvtz_cart", "j2pade")])
def test_move_ratios(name, jas):
    """a9 / a17: single-electron determinant and Jastrow ratios vs brute-force re-evaluation."""
    H = _with_jastrow(load_system(name), jas)
    eng = _engine(H)
    nw = 4
    r_up, r_dn = random_walkers(H, nw, 8)
    n_up, n_dn = r_up.shape[1], r_dn.shape[1]
    elec = list(range(n_up + n_dn)) * 2
    rng = np.random.default_rng(5)
    r_all = np.concatenate([r_up, r_dn], axis=1)
    r_new = r_all[:, elec, :] + rng.normal(scale=0.4, size=(nw, len(elec), 3))
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    dr, jr = eng.move_ratios(r_up, r_dn, Ginv, elec, r_new)
    dr, jr = dr.cpu().numpy(), jr.cpu().numpy()
    wf = H.wavefunction_data
    for w in range(nw):
        for k, e in enumerate(elec):
            up, idx = (True, e) if e < n_up else (False, e - n_up)
            tot = P.wf_ratio_brute_force(wf, r_up[w], r_dn[w], up, idx, r_new[w, k])
            det = P.wf_ratio_brute_force(wf, r_up[w], r_dn[w], up, idx, r_new[w, k], det_only=True)
            np.testing.assert_allclose(dr[w, k], det, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(dr[w, k] * jr[w, k], tot, rtol=1e-9, atol=1e-12)


def test_as_factor(water):
    eng = _engine(water)
    r_up, r_dn = random_walkers(water, 7, 12)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    got = eng.as_reg_fast(G, Ginv).cpu().numpy()
    for w in range(7):
        ref = P.compute_AS_regularization_factor(G[w].cpu().numpy(), Ginv[w].cpu().numpy())
        np.testing.assert_allclose(got[w], ref, rtol=1e-12)


@pytest.mark.parametrize("name,jas,eps,nmpm", [("water_ccecp_ccpvqz", "j2pade", 0.0, 24), ("water_ccecp_ccpvqz", "j1exp_j2exp", 0.05, 16),
                                               ("Li_ae_ccpvdz_cart", "j2pade", 0.0, 20), ("N2_ecp_ccpvtz_cart", "j1pade_j2pade", 0.1, 10),
                                               ("H2_ae_ccpvdz_cart", "none", 0.0, 30),
                                               ("H_ecp_ccpvqz", "j1pade_j2pade", 0.0, 14)])  # fmt: skip  (no down electron: _update_electron_positions_only_up_electron, jqmc_mcmc.py:4536-4722)
def test_mcmc_update_trajectory(name, jas, eps, nmpm):
    """kernel 6 (a28): same keys -> same proposals, bit-exact accept/reject sequence, keys bit-exact,
    positions/G/Ginv to round-off."""
    H = _with_jastrow(load_system(name), jas)
    eng = _engine(H)
    nw = 3
    r_up, r_dn = random_walkers(H, nw, 77, scale=0.6)
    keys = np.array([[0, 4242 + 13 * i] for i in range(nw)], dtype=np.uint32)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    acc, rej, ru, rd, k2, Gi2, G2 = eng.update(r_up, r_dn, keys, nmpm, 2.0, eps, Ginv, G)
    acc, rej, ru, rd, k2, Gi2, G2 = (x.cpu().numpy() for x in (acc, rej, ru, rd, k2, Gi2, G2))
    G, Ginv = G.cpu().numpy(), Ginv.cpu().numpy()
    for w in range(nw):
        a, r_, ru_o, rd_o, key_o, Gi_o, G_o = OD.update_electron_positions(
            H, r_up[w], r_dn[w], (int(keys[w, 0]), int(keys[w, 1])), nmpm, 2.0, eps, Ginv[w], G[w]
        )
        assert (a, r_) == (int(acc[w]), int(rej[w]))
        assert a + r_ == nmpm
        assert tuple(int(x) for x in k2[w]) == tuple(key_o)
        np.testing.assert_allclose(ru[w], ru_o, rtol=0, atol=1e-11)
        np.testing.assert_allclose(rd[w], rd_o, rtol=0, atol=1e-11)
        np.testing.assert_allclose(G2[w], G_o, rtol=1e-9, atol=1e-12 * np.abs(G_o).max())
        np.testing.assert_allclose(Gi2[w], Gi_o, rtol=1e-8, atol=1e-10 * np.abs(Gi_o).max())
        # the running inverse is still the inverse of the geminal at the final positions
        Gfresh = P.compute_geminal_all_elements(H.wavefunction_data.geminal_data, ru[w], rd[w])
        np.testing.assert_allclose(G2[w], Gfresh, rtol=1e-9, atol=1e-12 * np.abs(Gfresh).max())
        np.testing.assert_allclose(Gi2[w] @ Gfresh, np.eye(len(Gfresh)), rtol=0, atol=1e-8)


def test_rng_stream_bit_exact(water):
    """Device RNG vs the NumPy jax.random restatement: key chain after nmpm proposals and RT angles."""
    eng = _engine(water)
    nw = 64
    keys = np.stack([np.arange(nw, dtype=np.uint32) * 7919 + 1, np.arange(nw, dtype=np.uint32) ** 2 + 5], axis=1).astype(np.uint32)
    r_up, r_dn = random_walkers(water, nw, 1)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    out = eng.update(r_up, r_dn, keys, 5, 2.0, 0.0, Ginv, G)
    k2 = out[4].cpu().numpy()
    RT = eng.generate_RTs(keys).cpu().numpy()
    for w in range(nw):
        k = (int(keys[w, 0]), int(keys[w, 1]))
        kk = k
        for _ in range(5 * 6):
            kk, _sub = R.split(kk)
        assert tuple(int(x) for x in k2[w]) == kk
        np.testing.assert_allclose(RT[w], OD.generate_rotation_matrix(k), rtol=0, atol=2e-15)
        np.testing.assert_allclose(RT[w] @ RT[w].T, np.eye(3), atol=1e-14)


def test_full_size_properties(water):
    """BASELINE-size run (4096 walkers): size-independent properties -- counts add up, the running inverse
    stays the inverse, e_L is finite, walkers sharing a key and a configuration stay identical."""
    import torch

    from jqmc_b200 import rng_host
    from jqmc_b200.mcmc import generate_init_electron_configurations

    H = _with_jastrow(water, "j2pade")
    eng = _engine(H)
    nw, nmpm = 4096, 40
    np.random.seed(1)
    r_up, r_dn, _, _ = generate_init_electron_configurations(4, 4, nw, H.coulomb_potential_data.effective_charges, H.structure_data.positions)
    keys = rng_host.split(rng_host.PRNGKey(99), nw)
    # duplicate walker 0 into the last slot (same key, same configuration) -> identical trajectory
    r_up[-1], r_dn[-1], keys[-1] = r_up[0], r_dn[0], keys[0]
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    acc_t = torch.zeros(nw, dtype=torch.int64, device="cuda")
    state = (r_up, r_dn, keys, Ginv, G)
    for _ in range(3):
        acc, rej, ru, rd, k2, Gi2, G2 = eng.update(state[0], state[1], state[2], nmpm, 2.0, 0.0, state[3], state[4])
        assert torch.all(acc + rej == nmpm)
        acc_t += acc
        state = (ru, rd, k2, Gi2, G2)
    ratio = acc_t.double().mean().item() / (3 * nmpm)
    assert 0.2 < ratio < 0.95, ratio
    Gf, Gif = eng.geminal_inv_batched(state[0], state[1])
    err = (torch.bmm(state[3], Gf) - torch.eye(4, device="cuda", dtype=torch.float64)).abs().amax(dim=(1, 2))
    assert err.median().item() < 1e-9 and err.max().item() < 1e-4
    np.testing.assert_allclose(state[4].cpu().numpy(), Gf.cpu().numpy(), rtol=1e-7, atol=1e-12)
    RT = eng.generate_RTs(state[2])
    e_L = eng.e_L_fast(state[0], state[1], RT, state[3])
    assert torch.isfinite(e_L).all()
    assert -30.0 < e_L.mean().item() < -10.0
    assert torch.equal(state[0][0], state[0][-1]) and torch.equal(state[2][0], state[2][-1]) and e_L[0] == e_L[-1]


def test_shape_errors(water):
    eng = _engine(water)
    r_up, r_dn = random_walkers(water, 2, 0)
    with pytest.raises(ValueError):
        eng.geminal_inv_batched(r_up[:, :3], r_dn)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    with pytest.raises(ValueError):
        eng.e_L_fast(r_up, r_dn, np.eye(3)[None], Ginv)  # RT for 1 walker, 2 walkers given
    with pytest.raises(ValueError):
        eng.update(r_up, r_dn, np.zeros((3, 2), dtype=np.uint32), 4, 2.0, 0.0, Ginv, G)


def test_mcmc_driver_runs(water):
    from jqmc_b200.mcmc import MCMC

    H = _with_jastrow(water, "j2pade")
    m = MCMC(H, mcmc_seed=34456, num_walkers=64, num_mcmc_per_measurement=40, Dt=2.0, epsilon_AS=0.0)
    m.run(num_mcmc_steps=30)
    assert m.e_L.shape == (30, 64) and m.w_L.shape == (30, 64)
    E, dE, Var, dVar = m.get_E(num_mcmc_warmup_steps=10, num_mcmc_bin_blocks=5)
    assert -18.5 < E < -15.5 and dE > 0
    # same seed -> same stream
    m2 = MCMC(H, mcmc_seed=34456, num_walkers=64, num_mcmc_per_measurement=40, Dt=2.0, epsilon_AS=0.0)
    m2.run(num_mcmc_steps=5)
    np.testing.assert_array_equal(m2.e_L, m.e_L[:5])


@pytest.mark.parametrize("name,jas", [("water_ccecp_ccpvqz", "j2pade"), ("N2_ecp_ccpvtz_cart", "j1pade_j2pade"), ("Li_ae_ccpvdz_cart", "j2pade"),
                                      ("H2_ecp_ccpvtz", "j1exp_j2exp")])  # fmt: skip
def test_local_energy_fused_equals_staged(name, jas):
    """The fused (one kernel) and staged (kernel chain) local-energy paths agree to round-off."""
    H = _with_jastrow(load_system(name), jas)
    eng = _engine(H)
    nw = 70
    r_up, r_dn = random_walkers(H, nw, 5)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    RT = eng.generate_RTs(np.array([[1, i] for i in range(nw)], dtype=np.uint32))
    eng.set_fused(True)
    a = [x.cpu().numpy() for x in eng.e_L_fast(r_up, r_dn, RT, Ginv, return_parts=True)]
    eng.set_fused(False)
    b = [x.cpu().numpy() for x in eng.e_L_fast(r_up, r_dn, RT, Ginv, return_parts=True)]
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y, rtol=1e-11, atol=1e-11 * np.abs(y).max())


@pytest.mark.parametrize("name,jas,nlm", [("water_ccecp_ccpvqz", "j2pade", "tmove"), ("water_ccecp_ccpvqz", "j1exp_j2exp", "dltmove"),
                                          ("Li_ae_ccpvdz_cart", "j2pade", "tmove"), ("H2_ecp_ccpvtz_cart", "j1pade_j2pade", "tmove"),
                                          ("H2_ae_ccpvdz_cart", "j1exp_j2exp", "tmove"), ("H2_ecp_ccpvtz", "j2pade", "dltmove"),
                                          ("H_ecp_ccpvqz", "j1exp_j2exp", "tmove")])  # fmt: skip
def test_lrdmc_V_elements(name, jas, nlm):
    """a20, a23-a25, a30: V_diag / V_nondiag of the lattice-regularised Hamiltonian."""
    H = _with_jastrow(load_system(name), jas)
    eng = _engine(H)
    nw = 3
    r_up, r_dn = random_walkers(H, nw, 14, scale=0.7)
    RT = eng.generate_RTs(np.array([[9, i] for i in range(nw)], dtype=np.uint32))
    Vd, Vn = eng.V_elements_n(r_up, r_dn, RT, nlm, 0.3)
    Vd, Vn, RT = Vd.cpu().numpy(), Vn.cpu().numpy(), RT.cpu().numpy()
    for w in range(nw):
        d, n = OD.lrdmc_V_elements(H, r_up[w], r_dn[w], RT[w], nlm, 0.3)
        np.testing.assert_allclose(Vd[w], d, rtol=1e-9)
        np.testing.assert_allclose(Vn[w], n, rtol=1e-9)


@pytest.mark.parametrize("name,jas,nlm,mesh", [("water_ccecp_ccpvqz", "j2pade", "tmove", True), ("water_ccecp_ccpvqz", "j2pade", "dltmove", True),
                                               ("Li_ae_ccpvdz_cart", "j1exp_j2exp", "tmove", True), ("H2_ecp_ccpvtz_cart", "j2pade", "tmove", False),
                                               ("H_ecp_ccpvqz", "j1exp_j2exp", "tmove", True)])  # fmt: skip
def test_lrdmc_projection_trajectory(name, jas, nlm, mesh):
    """kernel 6 / a30: same keys -> the same mesh moves are selected (bit-exact positions up to round-off of the
    mesh point), same weights, keys bit-exact."""
    H = _with_jastrow(load_system(name), jas)
    eng = _engine(H)
    nw, nmpm, alat = 3, 6, 0.3
    E_scf = {"water_ccecp_ccpvqz": -17.0, "Li_ae_ccpvdz_cart": -7.4, "H2_ecp_ccpvtz_cart": -1.1, "H_ecp_ccpvqz": -0.45}[name]
    r_up, r_dn = random_walkers(H, nw, 41, scale=0.7)
    keys = np.array([[0, 777 + 5 * i] for i in range(nw)], dtype=np.uint32)
    Ginv = eng.A_inv_n(r_up, r_dn)
    w0 = np.ones(nw)
    out = eng.projection_n(w0, r_up, r_dn, Ginv, keys, E_scf, nmpm, mesh, nlm, alat)
    w, ru, rd, Gi, k2, RT, Vd, Vn = (x.cpu().numpy() for x in out)
    Ginv = Ginv.cpu().numpy()
    for i in range(nw):
        ow, oru, ord_, oGi, okey, oRT, od, on = OD.lrdmc_projection(
            H, 1.0, r_up[i], r_dn[i], Ginv[i], (int(keys[i, 0]), int(keys[i, 1])), E_scf, nmpm, mesh, nlm, alat
        )
        assert tuple(int(x) for x in k2[i]) == tuple(okey)
        np.testing.assert_allclose(ru[i], oru, rtol=0, atol=1e-11)
        np.testing.assert_allclose(rd[i], ord_, rtol=0, atol=1e-11)
        np.testing.assert_allclose(w[i], ow, rtol=1e-8)
        np.testing.assert_allclose(RT[i], oRT, rtol=0, atol=1e-14)
        np.testing.assert_allclose(Vd[i], od, rtol=1e-8)
        np.testing.assert_allclose(Vn[i], on, rtol=1e-8)
        np.testing.assert_allclose(Gi[i], oGi, rtol=1e-7, atol=1e-9 * np.abs(oGi).max())


def _check_projection_t(H, eng, nw, tau, mesh, nlm, alat, seed=41):
    r_up, r_dn = random_walkers(H, nw, seed, scale=0.7)
    keys = np.array([[0, 977 + 5 * i] for i in range(nw)], dtype=np.uint32)
    Ginv = eng.A_inv_n(r_up, r_dn)
    out = eng.projection_t(np.ones(nw), r_up, r_dn, Ginv, keys, tau, mesh, nlm, alat)
    e_L, pc, w, ru, rd, Gi, k2, RT = (x.cpu().numpy() for x in out)
    oe, opc, ow, oru, ord_, oGi, ok2, oRT, n_it = OD.lrdmc_projection_t_loop(H, np.ones(nw), r_up, r_dn, Ginv.cpu().numpy(), keys, tau, mesh, nlm, alat)
    np.testing.assert_array_equal(pc, opc)  # the same number of projections for every walker
    np.testing.assert_array_equal(k2, ok2)  # keys bit-exact, including the no-move iterations of the early finishers
    np.testing.assert_allclose(ru, oru, rtol=0, atol=1e-11)
    np.testing.assert_allclose(rd, ord_, rtol=0, atol=1e-11)
    np.testing.assert_allclose(w, ow, rtol=1e-8)
    np.testing.assert_allclose(e_L, oe, rtol=1e-8)
    np.testing.assert_allclose(RT, oRT, rtol=0, atol=1e-14)
    for i in range(nw):
        np.testing.assert_allclose(Gi[i], oGi[i], rtol=1e-6, atol=1e-8 * np.abs(oGi[i]).max())
    return opc, n_it


@pytest.mark.parametrize("name,jas,nlm,mesh,tau,wpc", [("water_ccecp_ccpvqz", "j2pade", "tmove", True, 0.025, 2), ("water_ccecp_ccpvqz", "j1exp_j2exp", "dltmove", True, 0.02, 0),
                                                       ("Li_ae_ccpvdz_cart", "j1exp_j2exp", "tmove", True, 0.03, 1), ("H2_ecp_ccpvtz_cart", "j2pade", "tmove", False, 0.1, 2),
                                                       ("H_ecp_ccpvqz", "j1exp_j2exp", "tmove", True, 0.2, 3)])  # fmt: skip
def test_lrdmc_projection_t_trajectory(name, jas, nlm, mesh, tau, wpc):
    """(f).2 GFMC_t: the continuous-time projection loop (jqmc/jqmc_gfmc.py:724-1110, 1539-1570) against the oracle's literal
    while_loop: per-walker projection counts, moves, weights, e_L and RT of the LAST iteration, and keys bit-exact.  Several
    CTAs with different iteration counts (wpc walkers per CTA) exercise the tail pass."""
    H = _with_jastrow(load_system(name), jas)
    eng = _engine(H)
    eng.set_walkers_per_cta(wpc)
    pc, n_it = _check_projection_t(H, eng, 5, tau, mesh, nlm, 0.3)
    assert n_it == pc.max() and pc.min() >= 1


def test_gfmc_t_driver_matches_oracle_loop():
    """(f).2: GFMC_t.run on the engine == the same branching steps done with plain oracle calls (same seeds)."""
    from jqmc_b200.gfmc import GFMC_t
    from tests import test_gfmc_host as TH

    H = TH._system()
    hist, ranks = TH._reference_run_t(H, 1)
    g = GFMC_t(H, num_walkers=TH.NW, num_gfmc_collect_steps=1, mcmc_seed=TH.SEED, tau=TH.TAU, alat=TH.ALAT)
    g.run(TH.STEPS)
    np.testing.assert_allclose(g.bare_w_L[:, 0], [h[0] for h in hist], rtol=1e-9)
    np.testing.assert_allclose(g.e_L[:, 0], [h[1] for h in hist][1:], rtol=1e-9)
    np.testing.assert_allclose(g.average_projection_counter, ranks[0]["pc"])
    assert g.num_survived_walkers == sum(h[3] for h in hist)
    np.testing.assert_allclose(g.latest_r_up_carts, ranks[0]["r_up"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(g.latest_r_dn_carts, ranks[0]["r_dn"], rtol=0, atol=1e-11)
    assert [tuple(int(x) for x in k) for k in g.jax_PRNG_key_list] == ranks[0]["keys"]


def test_gfmc_t_driver_water_runs(water):
    """GFMC_t on the BASELINE system: 30 branching steps of tau = 0.05 with 256 walkers; energies in the physical range, the
    projection count matches tau x |V_nondiag|, total weight and time bookkeeping are consistent."""
    from jqmc_b200.gfmc import GFMC_t

    H = _with_jastrow(water, "j2pade")
    g = GFMC_t(H, num_walkers=256, num_gfmc_collect_steps=2, mcmc_seed=11, tau=0.05, alat=0.3)
    g.run(30)
    assert g.mcmc_counter == 28 and g.e_L.shape == (28, 1)
    assert np.all(np.isfinite(g.e_L)) and -19.0 < g.e_L[5:].mean() < -15.5
    apc = g.average_projection_counter
    assert apc.shape == (30,) and np.all(apc > 2.0) and np.all(apc < 40.0)
    E, s, V, sv = g.get_E(num_mcmc_warmup_steps=8, num_mcmc_bin_blocks=5)
    assert -19.0 < E < -15.5 and s > 0
    assert 0.5 < g.num_survived_walkers / (g.num_survived_walkers + g.num_killed_walkers) <= 1.0


@pytest.mark.parametrize("world,nw", [(1, 4096), (2, 1000), (8, 4096), (3, 7)])
def test_lrdmc_branch_indices_bit_exact(water, world, nw):
    """a30 reconfiguration: comb indices over the all-gathered weights are bit-identical to the reference's
    NumPy/MPI arithmetic (np.sum pairwise, np.cumsum, Exscan offsets, searchsorted), survivors counted alike."""
    eng = _engine(water)
    rng = np.random.default_rng(100 + world)
    for trial in range(3):
        w = rng.uniform(0.2, 1.8, size=world * nw) * np.exp(rng.normal(scale=0.3, size=world * nw))
        if trial == 2:
            w[rng.integers(0, world * nw, size=max(1, nw // 10))] *= 40.0  # heavy walkers taking many slots
        zeta = float(rng.random())
        chosen, ns = eng.lrdmc_branch(w, nw, zeta)
        ref, ref_ns = OD.lrdmc_branch_indices(np.split(w, world), zeta)
        np.testing.assert_array_equal(chosen.cpu().numpy(), ref)
        assert int(ns.item()) == ref_ns


def test_lrdmc_collect_and_gather(water):
    eng = _engine(water)
    rng = np.random.default_rng(5)
    nw = 777
    w, Vd, Vn = rng.uniform(0.5, 1.5, nw), rng.normal(-10, 1, nw), rng.normal(-7, 1, nw)
    out = eng.lrdmc_collect(w, Vd, Vn, -17.2).cpu().numpy()
    np.testing.assert_allclose(out, OD.lrdmc_collect(w, Vd, Vn, -17.2), rtol=1e-13)
    src_up, src_dn = rng.normal(size=(3 * nw, 4, 3)), rng.normal(size=(3 * nw, 4, 3))
    idx = rng.integers(0, 3 * nw, size=nw).astype(np.int32)
    du, dd = eng.gather_walkers(idx, src_up, src_dn)
    np.testing.assert_array_equal(du.cpu().numpy(), src_up[idx])
    np.testing.assert_array_equal(dd.cpu().numpy(), src_dn[idx])
    with pytest.raises(ValueError):
        eng.lrdmc_branch(np.ones(10), 3, 0.5)
    with pytest.raises(ValueError):
        eng.lrdmc_branch(np.ones(8), 4, 1.5)


def test_gfmc_n_driver_matches_oracle_loop():
    """a30: GFMC_n.run on the engine == the same branching steps done with plain oracle calls (same seeds):
    stored averages, survivors, final walkers and keys."""
    from jqmc_b200.gfmc import GFMC_n
    from tests import test_gfmc_host as TH

    H = TH._system()
    hist, ranks = TH._reference_run(H, 1)
    g = GFMC_n(H, num_walkers=TH.NW, num_mcmc_per_measurement=TH.NMPM, num_gfmc_collect_steps=1, mcmc_seed=TH.SEED, E_scf=TH.E_SCF,
               alat=TH.ALAT)  # fmt: skip
    g.run(TH.STEPS)
    np.testing.assert_allclose(g.bare_w_L[:, 0], [h[0] for h in hist], rtol=1e-9)
    np.testing.assert_allclose(g.e_L[:, 0], [h[1] for h in hist][1:], rtol=1e-9)
    assert g.num_survived_walkers == sum(h[3] for h in hist)
    np.testing.assert_allclose(g.latest_r_up_carts, ranks[0]["r_up"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(g.latest_r_dn_carts, ranks[0]["r_dn"], rtol=0, atol=1e-11)
    assert [tuple(int(x) for x in k) for k in g.jax_PRNG_key_list] == ranks[0]["keys"]


def test_gfmc_n_driver_water_runs(water):
    """LRDMC on the BASELINE system: 30 branching steps with 64 walkers; energies finite and in the physical range,
    E_scf updated on the fly, get_E works."""
    from jqmc_b200.gfmc import GFMC_n

    H = _with_jastrow(water, "j2pade")
    g = GFMC_n(H, num_walkers=64, num_mcmc_per_measurement=10, num_gfmc_collect_steps=2, mcmc_seed=11, E_scf=-17.0, alat=0.3)
    g.run(30)
    assert g.mcmc_counter == 28 and g.e_L.shape == (28, 1)
    assert np.all(np.isfinite(g.e_L)) and -19.0 < g.e_L[5:].mean() < -15.5
    assert g.E_scf != -17.0
    E, s, V, sv = g.get_E(num_mcmc_warmup_steps=8, num_mcmc_bin_blocks=5)
    assert -19.0 < E < -15.5 and s > 0
    assert 0.5 < g.num_survived_walkers / (g.num_survived_walkers + g.num_killed_walkers) <= 1.0
