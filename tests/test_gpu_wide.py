"""GPU parity tests of the general ("wide") kernel family (jqmc_b200/csrc/qe_wide.cu): AO-basis JAGP geminals, three-body
Jastrow (AO- and MO-type orbital sets, spherical and Cartesian), more electrons than the register kernels cover, and the
small MO-basis systems forced onto this path.  Same oracle, same tolerances as tests/test_gpu_parity.py: fp64 1e-10 relative
unless argued otherwise, accept/reject counts, selected moves and keys bit-exact."""

import copy
import zlib
import dataclasses

import numpy as np
import pytest

from jqmc_b200.data import (
    Geminal_data,
    Jastrow_data,
    Jastrow_one_body_data,
    Jastrow_three_body_data,
    Jastrow_two_body_data,
    MOs_data,
    is_cart,
)
from oracle import drivers as OD
from oracle import physics as P
from tests.conftest import load_system, random_walkers

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def sub_basis(aos, lmax):
    """AO subset with l <= lmax (a small orbital set for the three-body Jastrow)."""
    l = np.asarray(aos.angular_momentums)
    keep = np.nonzero(l <= lmax)[0]
    new_index = {int(a): i for i, a in enumerate(keep)}
    oi = np.asarray(aos.orbital_indices)
    pk = np.nonzero(np.isin(oi, keep))[0]
    kw = dict(
        structure_data=aos.structure_data,
        nucleus_index=tuple(int(np.asarray(aos.nucleus_index)[a]) for a in keep),
        num_ao=len(keep),
        num_ao_prim=len(pk),
        angular_momentums=tuple(int(l[a]) for a in keep),
        orbital_indices=tuple(new_index[int(oi[p])] for p in pk),
        exponents=np.asarray(aos.exponents, dtype=np.float64)[pk],
        coefficients=np.asarray(aos.coefficients, dtype=np.float64)[pk],
    )
    if is_cart(aos):
        for f in ("polynominal_order_x", "polynominal_order_y", "polynominal_order_z"):
            kw[f] = tuple(int(np.asarray(getattr(aos, f))[a]) for a in keep)
    else:
        kw["magnetic_quantum_numbers"] = tuple(int(np.asarray(aos.magnetic_quantum_numbers)[a]) for a in keep)
    return type(aos)(**kw)


def _j12(H, j1, j2):
    cp = H.coulomb_potential_data
    core = tuple(cp.z_cores) if cp.ecp_flag else tuple(0 for _ in H.structure_data.atomic_numbers)
    one = None if j1 is None else Jastrow_one_body_data(jastrow_1b_param=0.9, jastrow_1b_type=j1, structure_data=H.structure_data, core_electrons=core)
    two = None if j2 is None else Jastrow_two_body_data(jastrow_2b_param=0.75, jastrow_2b_type=j2)
    return one, two


def _j3(H, kind, seed, lmax=1):
    if kind is None:
        return None
    rng = np.random.default_rng(seed)
    gem = H.wavefunction_data.geminal_data
    aos_full = gem.orb_data_up_spin.aos_data if hasattr(gem.orb_data_up_spin, "aos_data") else gem.orb_data_up_spin
    aos = sub_basis(aos_full, lmax)
    if kind == "ao":
        orb, n = aos, aos.num_ao
    else:
        n = 6
        orb = MOs_data(num_mo=n, aos_data=aos, mo_coefficients=rng.normal(scale=0.5, size=(n, aos.num_ao)))
    M = rng.normal(scale=0.02, size=(n, n))
    M = 0.5 * (M + M.T) + rng.normal(scale=0.004, size=(n, n))  # the reference symmetrises in optimisation; keep it general
    j1v = rng.normal(scale=0.05, size=(n, 1))
    return Jastrow_three_body_data(orb_data=orb, j_matrix=np.hstack([M, j1v]))


def make_case(case):
    """(Hamiltonian, force_wide)"""
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    if case == "water_jsd":
        H = copy.deepcopy(load_system("water_ccecp_ccpvqz"))
        j1, j2 = _j12(H, "exp", "exp")
        H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_one_body_data=j1, jastrow_two_body_data=j2)
        return H, True
    if case == "h_atom":  # one up electron, no down electron
        H = copy.deepcopy(load_system("H_ecp_ccpvqz"))
        j1, j2 = _j12(H, "exp", "pade")
        H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_one_body_data=j1, jastrow_two_body_data=j2)
        return H, True
    if case == "li_ae":
        H = copy.deepcopy(load_system("Li_ae_ccpvdz_cart"))
        j1, j2 = _j12(H, "pade", "pade")
        H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_one_body_data=j1, jastrow_two_body_data=j2)
        return H, True
    if case in ("water_jagp", "water_jagp_j3mo", "n2_jagp_j3ao"):
        H = copy.deepcopy(load_system("N2_ecp_ccpvtz_cart" if case.startswith("n2") else "water_ccecp_ccpvqz"))
        gem = Geminal_data.convert_from_MOs_to_AOs(H.wavefunction_data.geminal_data)
        lam = np.array(gem.lambda_matrix)
        pert = rng.normal(scale=2e-3, size=lam.shape)
        gem = dataclasses.replace(gem, lambda_matrix=lam + pert)
        H.wavefunction_data.geminal_data = gem
        if case == "water_jagp":
            j1, j2 = _j12(H, None, "pade")
            j3 = None
        elif case == "water_jagp_j3mo":
            j1, j2 = _j12(H, "exp", "pade")
            j3 = _j3(load_system("water_ccecp_ccpvqz"), "mo", 3)
        else:
            j1, j2 = _j12(H, None, "exp")
            j3 = _j3(load_system("N2_ecp_ccpvtz_cart"), "ao", 4)
        H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_one_body_data=j1, jastrow_two_body_data=j2, jastrow_three_body_data=j3)
        return H, False
    if case == "water_j3ao":
        H = copy.deepcopy(load_system("water_ccecp_ccpvqz"))
        j1, j2 = _j12(H, None, "pade")
        H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=j2, jastrow_three_body_data=_j3(H, "ao", 5, lmax=2))
        return H, False
    if case == "big":  # 10 up / 9 dn electrons in 12 synthetic MOs over the water AO basis: beyond the register kernels
        H = copy.deepcopy(load_system("water_ccecp_ccpvqz"))
        gem = H.wavefunction_data.geminal_data
        aos = gem.orb_data_up_spin.aos_data
        n_mo, n_up, n_dn = 12, 10, 9
        Cu = rng.normal(scale=0.4, size=(n_mo, aos.num_ao))
        Cd = Cu + rng.normal(scale=0.05, size=Cu.shape)  # unrestricted: different tables per spin
        lam = np.hstack([np.eye(n_mo) + rng.normal(scale=0.05, size=(n_mo, n_mo)), rng.normal(scale=0.5, size=(n_mo, n_up - n_dn))])
        H.wavefunction_data.geminal_data = Geminal_data(
            num_electron_up=n_up, num_electron_dn=n_dn, orb_data_up_spin=MOs_data(num_mo=n_mo, aos_data=aos, mo_coefficients=Cu),
            orb_data_dn_spin=MOs_data(num_mo=n_mo, aos_data=aos, mo_coefficients=Cd), lambda_matrix=lam)  # fmt: skip
        j1, j2 = _j12(H, "exp", "pade")
        H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_one_body_data=j1, jastrow_two_body_data=j2,
                                                        jastrow_three_body_data=_j3(load_system("water_ccecp_ccpvqz"), "ao", 6))  # fmt: skip
        return H, False
    from jqmc_b200 import synthetic as SY

    if case == "benzene":  # BASELINE configs[3] shape: 30 electrons, 258 spherical AOs, 15 MOs, J1+J2+J3
        return SY.benzene_shape(), False
    if case == "benzene_jagp":  # the same with the 258 x 258 AO-basis geminal
        return SY.benzene_shape(jagp=True), False
    if case == "grid48":  # 48 electrons, 360 Cartesian AOs (partial last shell), 24 MOs, J2+J3
        return SY.grid_molecule(12, 4, 30, 24, j3_ao_per_atom=4), False
    if case == "S":  # BASELINE configs[4]: 100 electrons, 1000 Cartesian AOs, 50 MOs
        return SY.grid_molecule(), False
    raise KeyError(case)


CASES = ["water_jsd", "h_atom", "li_ae", "water_jagp", "water_j3ao", "water_jagp_j3mo", "n2_jagp_j3ao", "big"]


def _engine(case):
    from jqmc_b200.engine import WalkerEngine

    H, force = make_case(case)
    eng = WalkerEngine(H)
    if force:
        eng.set_path(True)
    return H, eng


def _walkers(H, nw, seed, scale=0.8):
    if len(H.structure_data.positions) > 3:  # synthetic shapes: electrons spread over the atoms
        from jqmc_b200 import synthetic as SY

        return SY.init_walkers(H, nw, seed, sigma=scale)
    return random_walkers(H, nw, seed, scale)


@pytest.mark.parametrize("case,nw", [("benzene", 2), ("benzene_jagp", 2), ("grid48", 2), ("S", 1)])
def test_wide_synthetic_shapes_local_energy(case, nw):
    """BASELINE configs[3] / [4] shapes (synthetic coefficients): G, ln|Psi|, per-electron kinetic energies, potentials, e_L."""
    H, eng = _engine(case)
    r_up, r_dn = _walkers(H, nw, 3)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    ln, sg = eng.ln_wavefunction(r_up, r_dn)
    keys = np.array([[1, 50 + i] for i in range(nw)], dtype=np.uint32)
    RT = eng.generate_RTs(keys)
    e_L, T, V = eng.e_L_fast(r_up, r_dn, RT, Ginv, return_parts=True)
    G, Ginv, ln, e_L, T, V, RT = (x.cpu().numpy() for x in (G, Ginv, ln, e_L, T, V, RT))
    wf, cp = H.wavefunction_data, H.coulomb_potential_data
    for w in range(nw):
        Gr = P.compute_geminal_all_elements(wf.geminal_data, r_up[w], r_dn[w])
        np.testing.assert_allclose(G[w], Gr, rtol=1e-9, atol=1e-12 * np.abs(Gr).max())
        np.testing.assert_allclose(Ginv[w] @ Gr, np.eye(len(Gr)), rtol=0, atol=1e-13 * np.linalg.cond(Gr))
        np.testing.assert_allclose(ln[w], P.evaluate_ln_wavefunction(wf, r_up[w], r_dn[w]), rtol=1e-10, atol=1e-9)
        Tu, Td = P.compute_kinetic_energy_all_elements(wf, r_up[w], r_dn[w], Ginv[w])
        Tref = np.concatenate([Tu, Td])
        np.testing.assert_allclose(T[w], Tref, rtol=1e-9, atol=1e-9 * np.abs(Tref).max())
        ref = P.compute_local_energy(H, r_up[w], r_dn[w], RT[w], Ginv=Ginv[w])
        np.testing.assert_allclose(e_L[w], ref, rtol=1e-9, atol=1e-9 * np.abs(Tref).max())


@pytest.mark.parametrize("case", ["benzene", "grid48"])
def test_wide_synthetic_shapes_trajectories(case):
    """Metropolis and LRDMC trajectories on the synthetic shapes: decisions / selected moves and keys bit-exact."""
    H, eng = _engine(case)
    nw = 2
    r_up, r_dn = _walkers(H, nw, 9)
    keys = np.array([[0, 31 + 7 * i] for i in range(nw)], dtype=np.uint32)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    acc, rej, ru, rd, k2, Gi2, G2 = (x.cpu().numpy() for x in eng.update(r_up, r_dn, keys, 6, 2.0, 0.0, Ginv, G))
    Gn, Gin = G.cpu().numpy(), Ginv.cpu().numpy()
    for w in range(nw):
        a, r_, ru_o, rd_o, key_o, Gi_o, G_o = OD.update_electron_positions(H, r_up[w], r_dn[w], (0, 31 + 7 * w), 6, 2.0, 0.0, Gin[w], Gn[w])
        assert (a, r_) == (int(acc[w]), int(rej[w])) and tuple(int(x) for x in k2[w]) == tuple(key_o)
        np.testing.assert_allclose(ru[w], ru_o, rtol=0, atol=1e-11)
        np.testing.assert_allclose(rd[w], rd_o, rtol=0, atol=1e-11)
        np.testing.assert_allclose(G2[w], G_o, rtol=1e-7, atol=1e-10 * np.abs(G_o).max())
    out = eng.projection_n(np.ones(nw), r_up, r_dn, Ginv, keys, -40.0, 2, True, "tmove", 0.3)
    w_, ru, rd, Gi, k2, RT, Vd, Vn = (x.cpu().numpy() for x in out)
    for i in range(1 if case == "grid48" else nw):  # (the oracle needs ~15 s per walker on the 48-electron grid)
        ow, oru, ord_, oGi, okey, oRT, od, on = OD.lrdmc_projection(H, 1.0, r_up[i], r_dn[i], Gin[i], (0, 31 + 7 * i), -40.0, 2, True, "tmove", 0.3)
        assert tuple(int(x) for x in k2[i]) == tuple(okey)
        np.testing.assert_allclose(ru[i], oru, rtol=0, atol=1e-11)
        np.testing.assert_allclose(rd[i], ord_, rtol=0, atol=1e-11)
        np.testing.assert_allclose(w_[i], ow, rtol=1e-7)
        np.testing.assert_allclose(Vd[i], od, rtol=1e-7)
        np.testing.assert_allclose(Vn[i], on, rtol=1e-7)


def test_wide_S_size_properties():
    """BASELINE configs[4] (100 electrons / 1000 AOs / 50 MOs) at 256 walkers: counts add up, the running inverse stays the
    inverse of the geminal at the final positions, duplicated walkers stay identical, e_L and V elements are finite."""
    import torch

    from jqmc_b200 import rng_host

    H, eng = _engine("S")
    nw = 256
    r_up, r_dn = _walkers(H, nw, 2)
    keys = rng_host.split(rng_host.PRNGKey(7), nw)
    r_up[-1], r_dn[-1], keys[-1] = r_up[0], r_dn[0], keys[0]
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    acc, rej, ru, rd, k2, Gi2, G2 = eng.update(r_up, r_dn, keys, 30, 2.0, 0.0, Ginv, G)
    assert torch.all(acc + rej == 30) and 0.05 < acc.double().mean().item() / 30 < 0.98
    Gf, Gif = eng.geminal_inv_batched(ru, rd)
    np.testing.assert_allclose(G2.cpu().numpy(), Gf.cpu().numpy(), rtol=1e-6, atol=1e-10 * float(Gf.abs().max()))
    err = (torch.bmm(Gi2, Gf) - torch.eye(50, device="cuda", dtype=torch.float64)).abs().amax(dim=(1, 2))
    assert err.median().item() < 1e-6, err.median().item()
    assert torch.equal(ru[0], ru[-1]) and torch.equal(k2[0], k2[-1])
    RT = eng.generate_RTs(k2)
    e_L = eng.e_L_fast(ru, rd, RT, Gi2)
    Vd, Vn = eng.V_elements_n(ru, rd, RT, "tmove", 0.3, A_inv=Gi2)
    assert torch.isfinite(e_L).all() and torch.isfinite(Vd).all() and torch.isfinite(Vn).all() and e_L[0] == e_L[-1]
    out = eng.projection_n(np.ones(nw), ru, rd, Gi2, k2, float(e_L.mean()) - 50.0, 3, True, "tmove", 0.3)
    assert torch.isfinite(out[0]).all() and torch.equal(out[1][0], out[1][-1])
    Gf2, _ = eng.geminal_inv_batched(out[1], out[2])
    err = (torch.bmm(out[3], Gf2) - torch.eye(50, device="cuda", dtype=torch.float64)).abs().amax(dim=(1, 2))
    assert err.median().item() < 1e-6, err.median().item()


@pytest.mark.parametrize("case", ["water_jagp_j3mo", "water_j3ao", "big"])
def test_wide_orbital_layers(case):
    """Orbital value / gradient / Laplacian of the geminal and J3 orbital sets (tensor-core AO->MO product)."""
    H, eng = _engine(case)
    rng = np.random.default_rng(2)
    Rn = np.asarray(H.structure_data.positions)
    r = Rn[rng.integers(0, len(Rn), 37)] + rng.normal(scale=0.9, size=(37, 3))
    wf = H.wavefunction_data
    for which, orb in (("up", wf.geminal_data.orb_data_up_spin), ("dn", wf.geminal_data.orb_data_dn_spin),
                       ("j3", wf.jastrow_data.jastrow_three_body_data.orb_data)):  # fmt: skip
        ref = np.stack(P.compute_orb_value_grad_lap(orb, r))
        got = eng.eval_orbitals(which, "orb", r).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize("case", CASES)
def test_wide_geminal_inverse_and_ln_wavefunction(case):
    H, eng = _engine(case)
    nw = 5
    r_up, r_dn = _walkers(H, nw, 21)
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
        np.testing.assert_allclose(ln[w], ref, rtol=RTOL, atol=1e-10)
        assert sg[w] == np.sign(np.linalg.det(Gr))


@pytest.mark.parametrize("case", CASES)
def test_wide_local_energy_and_parts(case):
    H, eng = _engine(case)
    nw = 4
    r_up, r_dn = _walkers(H, nw, 33)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    keys = np.array([[3, 100 + i] for i in range(nw)], dtype=np.uint32)
    RT = eng.generate_RTs(keys)
    e_L, T, V = eng.e_L_fast(r_up, r_dn, RT, Ginv, return_parts=True)
    e_L, T, V, RT, Ginv = (x.cpu().numpy() for x in (e_L, T, V, RT, Ginv))
    wf, cp = H.wavefunction_data, H.coulomb_potential_data
    for w in range(nw):
        RTw = OD.generate_rotation_matrix((3, 100 + w))
        Tu, Td = P.compute_kinetic_energy_all_elements(wf, r_up[w], r_dn[w], Ginv[w])
        Tref = np.concatenate([Tu, Td])
        np.testing.assert_allclose(T[w], Tref, rtol=RTOL, atol=1e-10 * np.abs(Tref).max())
        np.testing.assert_allclose(V[w, 0], P.compute_bare_coulomb_potential(cp, r_up[w], r_dn[w]), rtol=RTOL)
        if cp.ecp_flag:
            np.testing.assert_allclose(V[w, 1], P.compute_ecp_local_parts(cp, r_up[w], r_dn[w]), rtol=RTOL, atol=1e-12)
            vnl = P.compute_ecp_non_local_parts_nearest_neighbors(cp, wf, r_up[w], r_dn[w], RTw, NN=1, Nv=6, Ginv=Ginv[w])[3]
            np.testing.assert_allclose(V[w, 2], vnl, rtol=1e-9, atol=1e-10)
        ref = P.compute_local_energy(H, r_up[w], r_dn[w], RTw, Ginv=Ginv[w])
        np.testing.assert_allclose(e_L[w], ref, rtol=RTOL, atol=1e-10 * np.abs(Tref).max())


@pytest.mark.parametrize("case", ["water_jsd", "water_jagp", "water_jagp_j3mo", "n2_jagp_j3ao", "big", "li_ae"])
def test_wide_move_ratios(case):
    """a9 / a17: determinant and Jastrow (J1+J2+J3) ratios of single-electron moves vs brute-force re-evaluation."""
    H, eng = _engine(case)
    nw = 3
    r_up, r_dn = _walkers(H, nw, 8)
    n_up, n_dn = r_up.shape[1], r_dn.shape[1]
    elec = list(range(n_up + n_dn))
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
            np.testing.assert_allclose(dr[w, k], det, rtol=1e-8, atol=1e-11)
            np.testing.assert_allclose(dr[w, k] * jr[w, k], tot, rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("case,eps,nmpm", [("water_jsd", 0.05, 16), ("h_atom", 0.0, 12), ("li_ae", 0.0, 20), ("water_jagp", 0.0, 20), ("water_j3ao", 0.0, 16),
                                           ("water_jagp_j3mo", 0.1, 12), ("n2_jagp_j3ao", 0.0, 12), ("big", 0.0, 24)])  # fmt: skip
def test_wide_mcmc_update_trajectory(case, eps, nmpm):
    """a28: same keys -> bit-exact accept/reject sequence and keys; positions, G, Ginv to round-off."""
    H, eng = _engine(case)
    nw = 3
    r_up, r_dn = _walkers(H, nw, 77, scale=0.6)
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
        np.testing.assert_allclose(G2[w], G_o, rtol=1e-8, atol=1e-11 * np.abs(G_o).max())
        Gfresh = P.compute_geminal_all_elements(H.wavefunction_data.geminal_data, ru[w], rd[w])
        np.testing.assert_allclose(G2[w], Gfresh, rtol=1e-8, atol=1e-11 * np.abs(Gfresh).max())
        np.testing.assert_allclose(Gi2[w] @ Gfresh, np.eye(len(Gfresh)), rtol=0, atol=1e-7)


@pytest.mark.parametrize("case,nlm", [("water_jsd", "tmove"), ("li_ae", "tmove"), ("water_jagp", "dltmove"), ("water_j3ao", "tmove"),
                                      ("water_jagp_j3mo", "dltmove"), ("n2_jagp_j3ao", "tmove"), ("big", "tmove")])  # fmt: skip
def test_wide_lrdmc_V_elements(case, nlm):
    H, eng = _engine(case)
    nw = 3
    r_up, r_dn = _walkers(H, nw, 14, scale=0.7)
    RT = eng.generate_RTs(np.array([[9, i] for i in range(nw)], dtype=np.uint32))
    Vd, Vn = eng.V_elements_n(r_up, r_dn, RT, nlm, 0.3)
    Vd, Vn, RT = Vd.cpu().numpy(), Vn.cpu().numpy(), RT.cpu().numpy()
    for w in range(nw):
        d, n = OD.lrdmc_V_elements(H, r_up[w], r_dn[w], RT[w], nlm, 0.3)
        np.testing.assert_allclose(Vd[w], d, rtol=1e-9)
        np.testing.assert_allclose(Vn[w], n, rtol=1e-9)


@pytest.mark.parametrize("case,nlm,E_scf,nmpm", [("water_jsd", "tmove", -17.0, 6), ("h_atom", "tmove", -0.45, 5), ("li_ae", "tmove", -7.4, 6), ("water_jagp", "tmove", -17.0, 6),
                                                 ("water_j3ao", "dltmove", -17.0, 5), ("water_jagp_j3mo", "tmove", -17.0, 5),
                                                 ("big", "tmove", -60.0, 4)])  # fmt: skip
def test_wide_lrdmc_projection_trajectory(case, nlm, E_scf, nmpm):
    """a30: same keys -> the same mesh moves, weights, keys; running inverse to round-off."""
    H, eng = _engine(case)
    nw, alat = 3, 0.3
    r_up, r_dn = _walkers(H, nw, 41, scale=0.7)
    keys = np.array([[0, 777 + 5 * i] for i in range(nw)], dtype=np.uint32)
    Ginv = eng.A_inv_n(r_up, r_dn)
    out = eng.projection_n(np.ones(nw), r_up, r_dn, Ginv, keys, E_scf, nmpm, True, nlm, alat)
    w, ru, rd, Gi, k2, RT, Vd, Vn = (x.cpu().numpy() for x in out)
    Ginv = Ginv.cpu().numpy()
    for i in range(nw):
        ow, oru, ord_, oGi, okey, oRT, od, on = OD.lrdmc_projection(
            H, 1.0, r_up[i], r_dn[i], Ginv[i], (int(keys[i, 0]), int(keys[i, 1])), E_scf, nmpm, True, nlm, alat
        )
        assert tuple(int(x) for x in k2[i]) == tuple(okey)
        np.testing.assert_allclose(ru[i], oru, rtol=0, atol=1e-11)
        np.testing.assert_allclose(rd[i], ord_, rtol=0, atol=1e-11)
        np.testing.assert_allclose(w[i], ow, rtol=1e-8)
        np.testing.assert_allclose(RT[i], oRT, rtol=0, atol=1e-14)
        np.testing.assert_allclose(Vd[i], od, rtol=1e-8)
        np.testing.assert_allclose(Vn[i], on, rtol=1e-8)
        np.testing.assert_allclose(Gi[i], oGi, rtol=1e-6, atol=1e-8 * np.abs(oGi).max())


def test_wide_walker_slices_give_identical_results():
    """BASELINE configs[4] sweeps to 64k walkers per GPU; the general family then runs a call as consecutive walker slices
    through one workspace.  Slicing must not change anything: every entry, 70 walkers in slices of 32 vs one call, bit for bit."""
    H, eng = _engine("water_jagp_j3mo")
    nw = 70
    r_up, r_dn = _walkers(H, nw, 47, scale=0.7)
    keys = np.array([[0, 31 + 7 * i] for i in range(nw)], dtype=np.uint32)
    RT = eng.generate_RTs(keys)

    def run():
        G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
        e_L = eng.e_L_fast(r_up, r_dn, RT, Ginv)
        ln = eng.ln_wavefunction(r_up, r_dn)
        up = eng.update(r_up, r_dn, keys, 6, 2.0, 0.0, Ginv, G)
        pr = eng.projection_n(np.ones(nw), r_up, r_dn, Ginv, keys, -17.0, 4, True, "tmove", 0.3)
        ve = eng.V_elements_n(r_up, r_dn, RT, "tmove", 0.3)
        out = []
        for x in (G, Ginv, e_L, ln, *up, *pr, *ve):
            out += [y.cpu().numpy() for y in (x if isinstance(x, (tuple, list)) else (x,))]
        return out

    try:
        eng.set_wide_slice(0)
        whole = run()
        eng.set_wide_slice(32)
        sliced = run()
    finally:
        eng.set_wide_slice(0)
    assert len(whole) == len(sliced)
    for a, b in zip(whole, sliced):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("case,nlm,tau", [("water_jsd", "tmove", 0.025), ("li_ae", "tmove", 0.03), ("water_jagp", "dltmove", 0.02), ("water_jagp_j3mo", "tmove", 0.015)])
def test_wide_lrdmc_projection_t_trajectory(case, nlm, tau):
    """(f).2 GFMC_t on the general path: the literal while_loop (every walker runs every iteration, walkers out of time do
    not move) against the oracle: projection counts and keys bit-exact, moves, weights, e_L and RT of the last iteration."""
    H, eng = _engine(case)

    nw, alat = 3, 0.3  # (the oracle's literal while_loop is the test time: ~1 s per projection and walker on the J3 cases)
    r_up, r_dn = _walkers(H, nw, 43, scale=0.7)
    keys = np.array([[0, 977 + 5 * i] for i in range(nw)], dtype=np.uint32)
    Ginv = eng.A_inv_n(r_up, r_dn)
    out = eng.projection_t(np.ones(nw), r_up, r_dn, Ginv, keys, tau, True, nlm, alat)
    e_L, pc, w, ru, rd, Gi, k2, RT = (x.cpu().numpy() for x in out)
    oe, opc, ow, oru, ord_, oGi, ok2, oRT, n_it = OD.lrdmc_projection_t_loop(H, np.ones(nw), r_up, r_dn, Ginv.cpu().numpy(), keys, tau, True, nlm, alat)
    np.testing.assert_array_equal(pc, opc)
    np.testing.assert_array_equal(k2, ok2)
    np.testing.assert_allclose(ru, oru, rtol=0, atol=1e-11)
    np.testing.assert_allclose(rd, ord_, rtol=0, atol=1e-11)
    np.testing.assert_allclose(w, ow, rtol=1e-8)
    np.testing.assert_allclose(e_L, oe, rtol=1e-8)
    np.testing.assert_allclose(RT, oRT, rtol=0, atol=1e-14)
    for i in range(nw):
        np.testing.assert_allclose(Gi[i], oGi[i], rtol=1e-6, atol=1e-8 * np.abs(oGi[i]).max())
    assert n_it == opc.max()


@pytest.mark.parametrize("name", ["w_2b_3b_w_ecp", "w_2b_1b3b_w_ecp", "w_1b_2b_1b3b_ae"])
def test_wide_turborvb_three_body_jastrow_known_answers(name):
    """The GPU path reproduces the TurboRVB known answers of the reference's J3 tests directly
    (tests/test_comparison_with_turborvb_ECP.py:376-416, 518-558; _AE.py:163-271): WF ratio^2 of the golden move through the
    move-ratio entry, kinetic energy and potential through the local-energy entry."""
    from jqmc_b200.engine import WalkerEngine
    from tests.conftest import turbo_j3_case

    H, up, dn, new_up, new_dn, spin, idx, ratio_ref, kin_ref, v_ref = turbo_j3_case(name)
    eng = WalkerEngine(H)
    G0, Ginv0 = eng.geminal_inv_batched(up[None], dn[None])
    e = idx if spin == "up" else len(up) + idx
    new = (new_up if spin == "up" else new_dn)[idx]
    dr, jr = eng.move_ratios(up[None], dn[None], Ginv0, [e], np.array(new)[None, None, :])
    np.testing.assert_almost_equal(((dr * jr) ** 2).item(), ratio_ref, decimal=6)
    G, Ginv = eng.geminal_inv_batched(new_up[None], new_dn[None])
    e_L, T, V = eng.e_L_fast(new_up[None], new_dn[None], np.eye(3)[None], Ginv, return_parts=True)
    np.testing.assert_almost_equal(T.sum().item(), kin_ref, decimal=6)
    np.testing.assert_almost_equal(V[0, :3].sum().item(), v_ref, decimal=5 if H.coulomb_potential_data.ecp_flag else 2)


def test_wide_turborvb_full_metropolis_known_answers():
    """tests/test_comparison_with_turborvb_ECP.py:660-945 through the engine: geminal before / after the golden move, AS
    factors, WF ratio^2 (x AS regularisation, epsilon = 0.3), kinetic energy and potential at the new configuration."""
    from jqmc_b200.engine import WalkerEngine
    from tests.conftest import TURBO_FULL as T, load_turbo_jastrow

    H = copy.deepcopy(load_system("water_ccecp_ccpvqz"))
    H.wavefunction_data.jastrow_data = load_turbo_jastrow("w_2b_1b3b_w_ecp", H.structure_data)
    eng = WalkerEngine(H)
    up, dn = np.array(T["old_up"]), np.array(T["old_dn"])
    new_up = up.copy()
    new_up[2] = T["new_up2"]
    G0, Gi0 = eng.geminal_inv_batched(up[None], dn[None])
    G1, Gi1 = eng.geminal_inv_batched(new_up[None], dn[None])
    np.testing.assert_almost_equal(G0[0].cpu().numpy(), np.array(T["geminal_old_T"]).T, decimal=6)
    np.testing.assert_almost_equal(G1[0].cpu().numpy(), np.array(T["geminal_new_T"]).T, decimal=6)
    R0, R1 = eng.as_reg_fast(G0, Gi0).item(), eng.as_reg_fast(G1, Gi1).item()
    np.testing.assert_almost_equal(R0, T["R_AS_old"], decimal=6)
    np.testing.assert_almost_equal(R1, T["R_AS_new"], decimal=6)
    dr, jr = eng.move_ratios(up[None], dn[None], Gi0, [2], np.array(T["new_up2"])[None, None, :])
    eps = 0.30
    ratio = ((dr * jr) ** 2).item() * ((max(R1, eps) / R1) / (max(R0, eps) / R0)) ** 2
    np.testing.assert_almost_equal(ratio, T["WF_ratio"], decimal=6)
    np.testing.assert_almost_equal(ratio * T["T_ratio"], T["final_ratio"], decimal=6)
    e_L, Tk, V = eng.e_L_fast(new_up[None], dn[None], np.eye(3)[None], Gi1, return_parts=True)
    np.testing.assert_almost_equal(Tk.sum().item(), T["kinc"], decimal=6)
    np.testing.assert_almost_equal(V[0, :3].sum().item(), T["vpot"] + T["vpotoff"], decimal=5)


@pytest.mark.parametrize("case", ["water_jsd", "li_ae", "h_atom", "water_jagp_j3mo", "n2_jagp_j3ao", "water_j3ao", "big", "benzene"])
def test_parameter_derivatives(case):
    """O_k = d ln|Psi| / d{j1, j2, j_matrix, lambda_matrix} per walker (the reference: jax.grad of
    evaluate_ln_wavefunction_fast) against the analytic oracle (itself checked against finite differences on CPU)."""
    H, eng = _engine(case)
    nw = 3
    r_up, r_dn = _walkers(H, nw, 19)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    got = {k: v.cpu().numpy() for k, v in eng.grad_ln_psi_params_fast(r_up, r_dn, Ginv).items()}
    Gi = Ginv.cpu().numpy()
    for w in range(nw):
        ref = P.compute_dln_wf_dparams(H.wavefunction_data, r_up[w], r_dn[w], Ginv=Gi[w])
        for k, v in ref.items():
            if v is None:
                assert k not in got
                continue
            v = np.asarray(v)
            assert got[k][w].shape == v.shape, (k, got[k][w].shape, v.shape)
            np.testing.assert_allclose(got[k][w], v, rtol=1e-9, atol=1e-11 * max(1.0, np.abs(v).max()))


def test_sr_optimisation_moves_a_bad_jastrow_parameter_towards_lower_energy():
    """BASELINE configs[3] capability at test size: VMC with stored O_k, generalised forces and SR steps on the device.  Start
    from a poor two-body parameter (a = 0.9): the natural-gradient steps must move it downhill and the energy must fall."""
    from jqmc_b200.mcmc import MCMC

    H = copy.deepcopy(load_system("water_ccecp_ccpvqz"))
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.9))
    m = MCMC(H, mcmc_seed=5, num_walkers=1024, num_mcmc_per_measurement=16, Dt=2.0, epsilon_AS=0.0, comput_log_WF_param_deriv=True)
    # E(a) of this one-parameter wavefunction falls steeply with a up to a ~ 3 (profiles/r01_sr_landscape.md: E(1.0) = -16.61,
    # E(1.5) = -16.89, E(2.5) = -17.00, f(1.0) = -dE/da = +1.05): every natural-gradient step must increase a
    hist = m.run_optimize(num_mcmc_steps=44, num_opt_steps=5, num_mcmc_warmup_steps=14, delta=0.05, epsilon=1e-3)
    a_new = m.hamiltonian_data.wavefunction_data.jastrow_data.jastrow_two_body_data.jastrow_2b_param
    assert a_new > 1.2, (a_new, hist)
    assert hist[-1][0] < hist[0][0] - 0.1, hist
    m.run(34)
    f, df = m.get_gF(num_mcmc_warmup_steps=4, num_mcmc_bin_blocks=5, blocks=["j2_param"])
    assert f.shape == (1,) and df[0] > 0 and f[0] > 3 * df[0], (f, df)  # still downhill towards larger a
    # with lambda: the flattened O matrix has the reference's block layout
    m2 = MCMC(H, mcmc_seed=5, num_walkers=64, num_mcmc_per_measurement=8, Dt=2.0, epsilon_AS=0.0, comput_log_WF_param_deriv=True)
    m2.run(6)
    O = m2.get_dln_WF(num_mcmc_warmup_steps=1)
    lam = np.shape(H.wavefunction_data.geminal_data.lambda_matrix)
    assert O.shape == (5, 64, 1 + lam[0] * lam[1])
    theta, info = m2.get_sr_direction(1, epsilon=1e-2)
    assert theta.shape == (1 + lam[0] * lam[1],) and np.all(np.isfinite(theta))


def test_wide_equals_register_kernels_at_scale():
    """water JSD + J2, 1000 walkers (not a multiple of any tile): the general path and the register/shared-memory kernels give
    the same e_L, the same Metropolis decisions and the same LRDMC moves; the tensor-core GEMM equals the plain DFMA GEMM."""
    import torch

    from jqmc_b200 import rng_host
    from jqmc_b200.engine import WalkerEngine

    H = copy.deepcopy(load_system("water_ccecp_ccpvqz"))
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=1.0))
    a, b = WalkerEngine(H), WalkerEngine(H)
    b.set_path(True)
    nw = 1000
    r_up, r_dn = random_walkers(H, nw, 3, 0.7)
    keys = rng_host.split(rng_host.PRNGKey(5), nw)
    Ga, Gia = a.geminal_inv_batched(r_up, r_dn)
    Gb, Gib = b.geminal_inv_batched(r_up, r_dn)
    np.testing.assert_allclose(Gb.cpu().numpy(), Ga.cpu().numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(Gib.cpu().numpy(), Gia.cpu().numpy(), rtol=1e-7, atol=1e-9 * float(Gia.abs().max()))
    RT = a.generate_RTs(keys)
    ea = a.e_L_fast(r_up, r_dn, RT, Gia).cpu().numpy()
    eb = b.e_L_fast(r_up, r_dn, RT, Gia).cpu().numpy()
    np.testing.assert_allclose(eb, ea, rtol=1e-9, atol=1e-9)
    b.set_gemm_reference(True)
    eb2 = b.e_L_fast(r_up, r_dn, RT, Gia).cpu().numpy()
    b.set_gemm_reference(False)
    np.testing.assert_allclose(eb2, eb, rtol=1e-11, atol=1e-11)
    oa = a.update(r_up, r_dn, keys, 12, 2.0, 0.0, Gia, Ga)
    ob = b.update(r_up, r_dn, keys, 12, 2.0, 0.0, Gia, Ga)
    # decisions can differ only where |ratio^2 T - u| is at round-off: allow a handful of walkers out of 1000
    same = (oa[0] == ob[0]).cpu().numpy()
    assert same.mean() > 0.995, same.mean()
    sel = torch.from_numpy(same).to(oa[2].device)
    np.testing.assert_allclose(ob[2][sel].cpu().numpy(), oa[2][sel].cpu().numpy(), rtol=0, atol=1e-9)
    assert torch.equal(oa[4], ob[4])
    w0 = np.ones(nw)
    pa = a.projection_n(w0, r_up, r_dn, Gia, keys, -17.0, 5, True, "tmove", 0.3)
    pb = b.projection_n(w0, r_up, r_dn, Gia, keys, -17.0, 5, True, "tmove", 0.3)
    close = (pa[1] - pb[1]).abs().amax(dim=(1, 2)).cpu().numpy() < 1e-9
    assert close.mean() > 0.995, close.mean()
    np.testing.assert_allclose(pb[0].cpu().numpy()[close], pa[0].cpu().numpy()[close], rtol=1e-7)
    np.testing.assert_allclose(pb[6].cpu().numpy()[close], pa[6].cpu().numpy()[close], rtol=1e-7)


def test_wide_drivers_run_jagp_j3():
    """MCMC and GFMC_n drivers on water JAGP + J1J2J3 (BASELINE configs[2] shape at test size): finite energies in the
    physical range, counters consistent."""
    from jqmc_b200.gfmc import GFMC_n
    from jqmc_b200.mcmc import MCMC

    H, _ = make_case("water_jagp_j3mo")
    # the Hartree-Fock geminal in AO form (no random perturbation) and a gentle J3, so that the energies stay physical
    H.wavefunction_data.geminal_data = Geminal_data.convert_from_MOs_to_AOs(load_system("water_ccecp_ccpvqz").wavefunction_data.geminal_data)
    j3 = H.wavefunction_data.jastrow_data.jastrow_three_body_data
    H.wavefunction_data.jastrow_data = Jastrow_data(
        jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=1.0),
        jastrow_three_body_data=dataclasses.replace(j3, j_matrix=0.02 * np.asarray(j3.j_matrix)),
    )
    m = MCMC(H, mcmc_seed=3, num_walkers=32, num_mcmc_per_measurement=16, Dt=2.0, epsilon_AS=0.0)
    m.run(num_mcmc_steps=12)
    assert m.e_L.shape == (12, 32) and np.all(np.isfinite(m.e_L))
    assert -19.5 < m.e_L[4:].mean() < -14.5
    g = GFMC_n(H, num_walkers=32, num_mcmc_per_measurement=6, num_gfmc_collect_steps=2, mcmc_seed=11, E_scf=-17.0, alat=0.3)
    g.run(10)
    assert np.all(np.isfinite(g.e_L)) and -20.0 < g.e_L[2:].mean() < -14.5
