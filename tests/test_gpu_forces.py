"""Atomic forces on the GPU: the device-side finite-difference position derivatives of e_L and ln|Psi|
against the oracle -- analytic gradient of ln|Psi| where the oracle has one, the oracle's own central differences for e_L and
for the nuclear derivatives (displaced Hamiltonians) -- then the force products with SWCT and the driver's jackknifed forces."""

import copy

import numpy as np
import pytest

from jqmc_b200.data import Jastrow_data, Jastrow_one_body_data, Jastrow_two_body_data
from oracle import drivers as OD
from oracle import physics as P
from tests.conftest import load_system, random_walkers

pytestmark = pytest.mark.gpu


def _system(name):
    H = copy.deepcopy(load_system(name))
    cp = H.coulomb_potential_data
    core = tuple(cp.z_cores) if cp.ecp_flag else tuple(0.0 for _ in H.structure_data.atomic_numbers)
    H.wavefunction_data.jastrow_data = Jastrow_data(
        jastrow_one_body_data=Jastrow_one_body_data(jastrow_1b_param=1.1, jastrow_1b_type="exp", structure_data=H.structure_data, core_electrons=core),
        jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.8, jastrow_2b_type="pade"),
    )  # fmt: skip
    return H


def _oracle_e_L(H, ru, rd, RT):
    _, Ginv = OD.geminal_inv(H.wavefunction_data.geminal_data, ru, rd)
    return P.compute_local_energy(H, ru, rd, RT, Ginv=Ginv)


@pytest.mark.parametrize("name", ["water_ccecp_ccpvqz", "H2_ae_ccpvdz_cart"])
def test_position_derivatives(name):
    from jqmc_b200.engine import WalkerEngine
    from jqmc_b200.forces import ForceEvaluator, displace_nucleus

    H = _system(name)
    eng = WalkerEngine(H)
    fe = ForceEvaluator(H, eng)
    nw, h = 2, fe.h
    r_up, r_dn = random_walkers(H, nw, 4, scale=0.7)
    RT = eng.generate_RTs(np.array([[1, 5 + i] for i in range(nw)], dtype=np.uint32))
    d = {k: v.cpu().numpy() for k, v in fe(r_up, r_dn, RT).items()}
    RTh = RT.cpu().numpy()
    wf = H.wavefunction_data
    n_at = len(H.structure_data.atomic_numbers)
    for w in range(nw):
        # ln|Psi|: analytic gradient of the oracle (determinant + Jastrow parts)
        gu_d, gd_d, _, _ = P.compute_grads_and_laplacian_ln_Det(wf.geminal_data, r_up[w], r_dn[w])
        gu_j, gd_j, _, _ = P.compute_grads_and_laplacian_Jastrow_part(wf.jastrow_data, r_up[w], r_dn[w])
        np.testing.assert_allclose(d["dln_Psi_dr_up"][w], gu_d + gu_j, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(d["dln_Psi_dr_dn"][w], gd_d + gd_j, rtol=2e-6, atol=1e-7)
        # e_L: the oracle's own central differences with the same step (one electron coordinate per spin, every nucleus)
        for spin, r, key in (("up", r_up, "de_L_dr_up"), ("dn", r_dn, "de_L_dr_dn")):
            for i, c in ((0, 0), (r.shape[1] - 1, 2)):
                rp, rm = r[w].copy(), r[w].copy()
                rp[i, c] += h
                rm[i, c] -= h
                fd = ((_oracle_e_L(H, rp, r_dn[w], RTh[w]) - _oracle_e_L(H, rm, r_dn[w], RTh[w])) if spin == "up" else
                      (_oracle_e_L(H, r_up[w], rp, RTh[w]) - _oracle_e_L(H, r_up[w], rm, RTh[w]))) / (2 * h)  # fmt: skip
                np.testing.assert_allclose(d[key][w, i, c], fd, rtol=1e-5, atol=1e-5)
        for a, c in ((0, 1), (n_at - 1, 0)):
            Hp, Hm = displace_nucleus(H, a, c, h), displace_nucleus(H, a, c, -h)
            fd_e = (_oracle_e_L(Hp, r_up[w], r_dn[w], RTh[w]) - _oracle_e_L(Hm, r_up[w], r_dn[w], RTh[w])) / (2 * h)
            fd_l = (P.evaluate_ln_wavefunction(Hp.wavefunction_data, r_up[w], r_dn[w]) - P.evaluate_ln_wavefunction(Hm.wavefunction_data, r_up[w], r_dn[w])) / (2 * h)
            np.testing.assert_allclose(d["de_L_dR"][w, a, c], fd_e, rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(d["dln_Psi_dR"][w, a, c], fd_l, rtol=1e-5, atol=1e-6)
    # translation invariance: moving every particle together changes nothing -> the derivatives sum to zero
    tot_e = d["de_L_dR"].sum(1) + d["de_L_dr_up"].sum(1) + d["de_L_dr_dn"].sum(1)
    tot_l = d["dln_Psi_dR"].sum(1) + d["dln_Psi_dr_up"].sum(1) + d["dln_Psi_dr_dn"].sum(1)
    np.testing.assert_allclose(tot_e, 0.0, atol=2e-5 * max(1.0, np.abs(d["de_L_dR"]).max()))
    np.testing.assert_allclose(tot_l, 0.0, atol=1e-6 * max(1.0, np.abs(d["dln_Psi_dR"]).max()))


def test_mcmc_forces_h2_symmetric():
    """H2: the two atoms feel opposite forces; the estimate is finite, antisymmetric within its error bars, and survives a
    checkpoint round trip."""
    from jqmc_b200.mcmc import MCMC

    H = _system("H2_ecp_ccpvtz")
    m = MCMC(H, mcmc_seed=7, num_walkers=256, num_mcmc_per_measurement=20, Dt=2.0, epsilon_AS=0.0, comput_position_deriv=True)
    m.run(num_mcmc_steps=30)
    F, dF = m.get_aF(num_mcmc_warmup_steps=5, num_mcmc_bin_blocks=5)
    assert F.shape == (2, 3) and np.all(np.isfinite(F)) and np.all(dF > 0)
    np.testing.assert_allclose(F[0] + F[1], 0.0, atol=6 * np.sqrt(dF[0] ** 2 + dF[1] ** 2).max() + 1e-3)
    axis = np.asarray(H.structure_data.positions)[1] - np.asarray(H.structure_data.positions)[0]
    axis /= np.linalg.norm(axis)
    perp = F[0] - (F[0] @ axis) * axis
    assert np.abs(perp).max() < 6 * dF[0].max() + 1e-3  # no force perpendicular to the bond


def _oracle_e_L_lattice(H, ru, rd, RT, alat, nlm):
    Vd, Vn = OD.lrdmc_V_elements(H, ru, rd, RT, nlm, alat)
    return Vd + Vn


@pytest.mark.parametrize("name,nlm", [("water_ccecp_ccpvqz", "tmove"), ("water_ccecp_ccpvqz", "dltmove"), ("H2_ae_ccpvdz_cart", "tmove")])
def test_lrdmc_position_derivatives(name, nlm):
    """Gradient of the lattice-regularised local energy V_diag + V_nondiag (the LRDMC force term, jqmc_gfmc.py:5630-5667,
    5840-5850) by device-side finite differences against the oracle's own central differences of lrdmc_V_elements."""
    from jqmc_b200.engine import WalkerEngine
    from jqmc_b200.forces import ForceEvaluator, displace_nucleus

    H = _system(name)
    alat = 0.3
    eng = WalkerEngine(H)
    fe = ForceEvaluator(H, eng, lattice=(alat, nlm))
    nw, h = 2, fe.h
    r_up, r_dn = random_walkers(H, nw, 9, scale=0.7)
    RT = eng.generate_RTs(np.array([[2, 11 + i] for i in range(nw)], dtype=np.uint32))
    d = {k: v.cpu().numpy() for k, v in fe(r_up, r_dn, RT).items()}
    RTh = RT.cpu().numpy()
    n_at = len(H.structure_data.atomic_numbers)
    for w in range(nw):
        np.testing.assert_allclose(d["e_L"][w], _oracle_e_L_lattice(H, r_up[w], r_dn[w], RTh[w], alat, nlm), rtol=1e-9)
        for spin, r, key in (("up", r_up, "de_L_dr_up"), ("dn", r_dn, "de_L_dr_dn")):
            for i, c in ((0, 1), (r.shape[1] - 1, 0)):
                rp, rm = r[w].copy(), r[w].copy()
                rp[i, c] += h
                rm[i, c] -= h
                if spin == "up":
                    fd = _oracle_e_L_lattice(H, rp, r_dn[w], RTh[w], alat, nlm) - _oracle_e_L_lattice(H, rm, r_dn[w], RTh[w], alat, nlm)
                else:
                    fd = _oracle_e_L_lattice(H, r_up[w], rp, RTh[w], alat, nlm) - _oracle_e_L_lattice(H, r_up[w], rm, RTh[w], alat, nlm)
                np.testing.assert_allclose(d[key][w, i, c], fd / (2 * h), rtol=1e-5, atol=1e-5)
        for a, c in ((0, 2), (n_at - 1, 1)):
            Hp, Hm = displace_nucleus(H, a, c, h), displace_nucleus(H, a, c, -h)
            fd = (_oracle_e_L_lattice(Hp, r_up[w], r_dn[w], RTh[w], alat, nlm) - _oracle_e_L_lattice(Hm, r_up[w], r_dn[w], RTh[w], alat, nlm)) / (2 * h)
            np.testing.assert_allclose(d["de_L_dR"][w, a, c], fd, rtol=1e-5, atol=1e-5)
    tot_e = d["de_L_dR"].sum(1) + d["de_L_dr_up"].sum(1) + d["de_L_dr_dn"].sum(1)
    np.testing.assert_allclose(tot_e, 0.0, atol=2e-5 * max(1.0, np.abs(d["de_L_dR"]).max()))


@pytest.mark.parametrize("kind", ["n", "t"])
def test_gfmc_forces_h2_symmetric(kind):
    """LRDMC forces on H2 through the drivers (SWCT + Pathak-Wagner switched on): finite, opposite on the two atoms and along the
    bond within the error bars; the force terms do not disturb the chain."""
    from jqmc_b200.gfmc import GFMC_n, GFMC_t

    H = _system("H2_ecp_ccpvtz")
    common = dict(num_walkers=128, num_gfmc_collect_steps=2, mcmc_seed=9, alat=0.3, use_swct=True, epsilon_PW=0.05)
    if kind == "n":
        mk = lambda deriv: GFMC_n(H, num_mcmc_per_measurement=10, E_scf=-1.1, comput_position_deriv=deriv, **common)  # noqa: E731
    else:
        mk = lambda deriv: GFMC_t(H, tau=0.05, comput_position_deriv=deriv, **common)  # noqa: E731
    g, plain = mk(True), mk(False)
    g.run(num_mcmc_steps=24)
    plain.run(num_mcmc_steps=24)
    np.testing.assert_array_equal(g.bare_w_L, plain.bare_w_L)
    np.testing.assert_array_equal(g.e_L, plain.e_L)
    assert g.force_HF.shape == (22, 1, 2, 3)
    np.testing.assert_allclose(g.force_HF.sum(axis=2), 0.0, atol=2e-5 * max(1.0, np.abs(g.force_HF).max()))
    np.testing.assert_allclose(g.force_PP.sum(axis=2), 0.0, atol=2e-6 * max(1.0, np.abs(g.force_PP).max()))
    F, dF = g.get_aF(num_mcmc_warmup_steps=2, num_mcmc_bin_blocks=5)
    assert F.shape == (2, 3) and np.all(np.isfinite(F)) and np.all(dF > 0)
    np.testing.assert_allclose(F[0] + F[1], 0.0, atol=1e-6 + 1e-3 * np.abs(F).max())  # exact up to the finite-difference error
    axis = np.asarray(H.structure_data.positions)[1] - np.asarray(H.structure_data.positions)[0]
    axis /= np.linalg.norm(axis)
    perp = F[0] - (F[0] @ axis) * axis
    assert np.abs(perp).max() < 6 * dF[0].max() + 1e-3
