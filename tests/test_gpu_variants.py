"""GPU parity for the corners `qe_create` accepts that the fixture systems do not reach by themselves:

* non-local ECP quadratures Nv in {4, 12, 18} and NN > 1 nearest nuclei (reference tables jqmc/coulomb_potential.py:96-184;
  reference tests tests/test_ecps.py:631-902), on both kernel families;
* multi-channel ccECPs on heavier atoms (Cl2, CuBr, Ti2 cc-pVTZ Cartesian: the reference's own test inputs);
* angular momenta l = 5, 6, spherical and Cartesian (every reference fixture stops at l = 4);
* a singular geminal matrix at (re)initialisation (the reference's thresholded pseudo-inverse stays finite, jqmc_mcmc.py:4258).
"""

import copy

import numpy as np
import pytest

from jqmc_b200.data import Jastrow_data, Jastrow_two_body_data
from oracle import drivers as OD
from oracle import physics as P
from tests.conftest import load_system, random_walkers

pytestmark = pytest.mark.gpu


def _engine(H, **kw):
    from jqmc_b200.engine import WalkerEngine

    return WalkerEngine(H, **kw)


def _j2(H, a=0.9):
    H = copy.deepcopy(H)
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=a, jastrow_2b_type="pade"))
    return H


def _check_ecp(H, eng, nw, Nv, NN, seed, nlm="tmove", alat=0.3, scale=0.8):
    r_up, r_dn = random_walkers(H, nw, seed, scale=scale)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    keys = np.array([[5, 40 + i] for i in range(nw)], dtype=np.uint32)
    RT = eng.generate_RTs(keys)
    e_L, T, V = eng.e_L_fast(r_up, r_dn, RT, Ginv, return_parts=True)
    Vd, Vn = eng.V_elements_n(r_up, r_dn, RT, nlm, alat)
    e_L, V, RT, Ginv, Vd, Vn, G = (x.cpu().numpy() for x in (e_L, V, RT, Ginv, Vd, Vn, G))
    wf, cp = H.wavefunction_data, H.coulomb_potential_data
    for w in range(nw):
        # every mesh ratio goes through G^-1: two correct evaluations differ by O(cond(G) eps) (random 13 x 13 geminal matrices
        # reach cond ~ 1e10; measured: 1.4e-9 relative at cond 6e9), so the bound scales with the condition number
        rt = max(1e-9, 5e-18 * np.linalg.cond(G[w]))
        vnl = P.compute_ecp_non_local_parts_nearest_neighbors(cp, wf, r_up[w], r_dn[w], RT[w], NN=NN, Nv=Nv, Ginv=Ginv[w])[3]
        np.testing.assert_allclose(V[w, 2], vnl, rtol=rt, atol=1e-11)
        np.testing.assert_allclose(V[w, 1], P.compute_ecp_local_parts(cp, r_up[w], r_dn[w]), rtol=1e-10, atol=1e-12)
        ref = P.compute_local_energy(H, r_up[w], r_dn[w], RT[w], Ginv=Ginv[w], NN=NN, Nv=Nv)
        np.testing.assert_allclose(e_L[w], ref, rtol=max(1e-10, rt), atol=1e-9)
        _, Gi = OD.geminal_inv(wf.geminal_data, r_up[w], r_dn[w])
        d, n, _, _ = OD.lrdmc_elements(H, r_up[w], r_dn[w], Gi, RT[w], alat, nlm, NN=NN, Nv=Nv)
        np.testing.assert_allclose(Vd[w], d, rtol=rt)
        np.testing.assert_allclose(Vn[w], n, rtol=rt)


@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("Nv,NN", [(4, 1), (12, 1), (18, 1), (6, 2), (12, 3)])
def test_ecp_quadrature_variants_water(Nv, NN, path):
    """a25: every quadrature table and NN > 1 on water ccECP (one non-local channel on O), both kernel families."""
    H = _j2(load_system("water_ccecp_ccpvqz"))
    eng = _engine(H, Nv=Nv, NN=NN)
    eng.set_path(path)
    _check_ecp(H, eng, 3, Nv, NN, seed=100 + Nv + NN, nlm="dltmove" if Nv == 12 else "tmove")


@pytest.mark.parametrize("name,Nv,NN", [("Cl2_ecp_ccpvtz_cart", 6, 1), ("Cl2_ecp_ccpvtz_cart", 12, 2), ("CuBr_ecp_ccpvtz_cart", 6, 1),
                                        ("CuBr_ecp_ccpvtz_cart", 4, 2), ("Ti2_ecp_ccpvtz_cart", 6, 2)])  # fmt: skip
def test_ecp_multi_channel_heavy_atoms(name, Nv, NN):
    """a25 with several non-local channels per atom (Cl: s, p; Cu / Br / Ti ccECPs): the reference's own heavier inputs."""
    H = _j2(load_system(name))
    eng = _engine(H, Nv=Nv, NN=NN)
    _check_ecp(H, eng, 2, Nv, NN, seed=7, scale=0.6)


def _high_l_system(cart: bool):
    from jqmc_b200 import synthetic as SY
    from jqmc_b200.data import Structure_data

    pos = np.array([[0.0, 0.0, 0.0], [1.6, 0.3, -0.4]])
    st = Structure_data(positions=pos, atomic_numbers=(3, 1), element_symbols=("Li", "H"), atomic_labels=("Li", "H"))
    if cart:
        aos = SY.cart_aos(st, 84)  # complete Cartesian shells s .. i (l = 6): 1 + 3 + 6 + 10 + 15 + 21 + 28
    else:
        aos = SY.sphe_aos(st, [[0, 1, 2, 5, 6, 6], [0, 5, 6]], n_prim=2)
    rng = np.random.default_rng(3)
    return SY._assemble(st, aos, [3, 1], [0, 0], 3, None, rng, j1=False, j2=True)


@pytest.mark.parametrize("cart", [False, True])
def test_high_angular_momentum(cart):
    """a2-a4 with l = 5 and 6 (generated solid harmonics / Cartesian monomials): AO and MO value, gradient, Laplacian, then
    ln|Psi|, local energy and move ratios of a 4-electron all-electron system on both kernel families."""
    H = _high_l_system(cart)
    gem = H.wavefunction_data.geminal_data
    assert max(gem.orb_data_up_spin.aos_data.angular_momentums) == 6
    rng = np.random.default_rng(11)
    Rn = np.asarray(H.structure_data.positions)
    r = Rn[rng.integers(0, 2, 30)] + rng.normal(scale=1.1, size=(30, 3))
    r[0] = Rn[1]
    nw = 3
    r_up, r_dn = random_walkers(H, nw, 5, scale=1.0)
    for path in (0, 1):
        eng = _engine(H)
        eng.set_path(path)
        ref_ao = np.stack(P.compute_AOs_value_grad_lap(gem.orb_data_up_spin.aos_data, r))
        got_ao = eng.eval_orbitals("up", "ao", r).cpu().numpy()
        np.testing.assert_allclose(got_ao, ref_ao, rtol=1e-10, atol=1e-12 * np.abs(ref_ao).max())
        ref_mo = np.stack(P.compute_orb_value_grad_lap(gem.orb_data_up_spin, r))
        got_mo = eng.eval_orbitals("up", "orb", r).cpu().numpy()
        np.testing.assert_allclose(got_mo, ref_mo, rtol=1e-10, atol=1e-12 * np.abs(ref_mo).max())
        G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
        ln, sg = eng.ln_wavefunction(r_up, r_dn)
        e_L = eng.e_L_fast(r_up, r_dn, None, Ginv).cpu().numpy()
        Vd, Vn = (x.cpu().numpy() for x in eng.V_elements_n(r_up, r_dn, np.tile(np.eye(3), (nw, 1, 1)), "tmove", 0.25))
        Ginv_h = Ginv.cpu().numpy()
        for w in range(nw):
            np.testing.assert_allclose(ln[w].item(), P.evaluate_ln_wavefunction(H.wavefunction_data, r_up[w], r_dn[w]), rtol=1e-10, atol=1e-11)
            ref = P.compute_local_energy(H, r_up[w], r_dn[w], np.eye(3), Ginv=Ginv_h[w])
            np.testing.assert_allclose(e_L[w], ref, rtol=1e-9, atol=1e-9)
            d, n = OD.lrdmc_V_elements(H, r_up[w], r_dn[w], np.eye(3), "tmove", 0.25)
            np.testing.assert_allclose(Vd[w], d, rtol=1e-9)
            np.testing.assert_allclose(Vn[w], n, rtol=1e-9)


@pytest.mark.parametrize("path", [0, 1])
def test_singular_geminal_stays_finite(path):
    """Two same-spin electrons on the same point make two rows of G identical; an electron far outside the basis makes a row
    vanish altogether.  The reference's pseudo-inverse (rcond 1e-20) returns finite numbers in both cases; so must the engine:
    no inf / NaN in Ginv, ln|Psi| = -inf or very negative, and the Metropolis kernel keeps running on such a walker."""
    import torch

    H = _j2(load_system("water_ccecp_ccpvqz"))
    eng = _engine(H)
    eng.set_path(path)
    r_up, r_dn = random_walkers(H, 4, 9)
    r_up[1, 2] = r_up[1, 0]  # coincident up electrons: numerically singular
    r_dn[2, 1] = np.array([400.0, -300.0, 250.0])  # every orbital underflows to zero: an exactly vanishing column
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    assert torch.isfinite(Ginv).all() and torch.isfinite(G).all()
    Gh, Gih = G.cpu().numpy(), Ginv.cpu().numpy()
    assert np.all(np.abs(Gh[2][:, 1]) < 1e-300)  # (the underflowing exponentials return ~1e-307, not an exact zero)
    np.testing.assert_array_equal(Gih[2][1, :], 0.0)  # the null direction is projected out, as the pseudo-inverse does
    ln, sg = eng.ln_wavefunction(r_up, r_dn)
    assert ln[2].item() == -np.inf or ln[2].item() < -600.0
    # regular walkers are untouched
    for w in (0, 3):
        np.testing.assert_allclose(Gih[w] @ Gh[w], np.eye(4), atol=1e-9)
    keys = np.array([[0, 77 + i] for i in range(4)], dtype=np.uint32)
    acc, rej, ru, rd, k2, Gi2, G2 = eng.update(r_up, r_dn, keys, 12, 2.0, 0.0, Ginv, G)
    assert torch.all(acc + rej == 12)
    assert torch.isfinite(ru).all() and torch.isfinite(rd).all()
