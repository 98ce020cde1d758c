"""Pin the CPU oracle against the reference's own known-answer numbers (TurboRVB, hard-coded in
jQMC's tests/test_comparison_with_turborvb_ECP.py:75-274) and against itself (fast-update vs
from-scratch, analytic vs finite differences) -- the reference's test patterns B, C and E."""

import numpy as np
import pytest

from jqmc_b200.data import Geminal_data, Jastrow_data, Jastrow_one_body_data, Jastrow_two_body_data
from oracle import physics as P
from tests.conftest import load_system, random_walkers

UP_A = np.array(
    [
        [-1.1345038587576, -0.698914730480577, -0.006290951981744008],
        [-2.07761893946839, 1.30902541938751, -0.05220902114745041],
        [0.276215481293413, 0.422863618938476, 0.27986648725301],
        [-1.60902246286275, 0.499927465264998, 0.70010581636993],
    ]
)
DN_A = np.array(
    [
        [-1.48583455555933, -1.01189391902775, 1.83998639430367],
        [0.635659512640246, 0.398999201990364, -0.745191606127732],
        [-2.00590358216444, 1.90796788491204, -0.195294104680795],
        [-1.12726250654165, -0.739542218156325, -0.04817447678670805],
    ]
)
UP_B = np.array(
    [
        [-1.1345038587576, -0.698914730480577, -0.006290951981744008],
        [-2.30366220171161, 1.47326376760292, 0.126403765463162],
        [0.276215481293413, 0.422863618938476, 0.27986648725301],
        [-2.54518559687882, 0.822753144911055, 0.70010581636993],
    ]
)
DN_B = np.array(
    [
        [-1.42343008909407, -1.13669461924113, 0.525171318204107],
        [1.90701925586575, 0.398999201990364, -0.745191606127732],
        [-2.00590358216444, 1.90796788491204, -0.195294104680795],
        [-1.12726250654165, -0.678049640381367, -0.656537799033216],
    ]
)
NEW_UP2 = [0.276215481293413, -0.270740090536313, 0.27986648725301]


def _check_turborvb(H, old_up, old_dn, ratio_ref, kin_ref, vpot_ref, vpotoff_ref):
    wf = H.wavefunction_data
    new_up = old_up.copy()
    new_up[2] = NEW_UP2
    ratio = (P.evaluate_wavefunction(wf, new_up, old_dn) / P.evaluate_wavefunction(wf, old_up, old_dn)) ** 2
    np.testing.assert_almost_equal(ratio, ratio_ref, decimal=6)
    np.testing.assert_almost_equal(P.compute_kinetic_energy(wf, new_up, old_dn), kin_ref, decimal=6)
    V = P.compute_coulomb_potential(H.coulomb_potential_data, wf, new_up, old_dn, RT=np.eye(3), NN=1, Nv=6)
    np.testing.assert_almost_equal(V, vpot_ref + vpotoff_ref, decimal=5)
    # fast (running inverse) path == brute force path
    Ginv = np.linalg.inv(P.compute_geminal_all_elements(wf.geminal_data, new_up, old_dn))
    Vf = P.compute_coulomb_potential(H.coulomb_potential_data, wf, new_up, old_dn, RT=np.eye(3), Ginv=Ginv)
    np.testing.assert_allclose(Vf, V, rtol=1e-11)
    np.testing.assert_allclose(
        P.compute_kinetic_energy(wf, new_up, old_dn, Ginv), P.compute_kinetic_energy(wf, new_up, old_dn), rtol=1e-10
    )


def test_turborvb_wo_jastrow(water):
    # tests/test_comparison_with_turborvb_ECP.py:105-129
    _check_turborvb(water, UP_A, DN_A, 0.919592366177397, 14.6961809426982, -17.0152290468758, 0.328893830058865)


def test_turborvb_w_2b_jastrow(water):
    # tests/test_comparison_with_turborvb_ECP.py:231-274
    import copy

    H = copy.deepcopy(water)
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.676718854150191))
    _check_turborvb(H, UP_B, DN_B, 0.881124604511419, 11.1237599317225, -27.03387193107, 0.243517439611676)


@pytest.mark.parametrize("name", ["water_ccecp_ccpvqz", "N2_ecp_ccpvtz_cart", "H2_ae_ccpvdz_cart"])
def test_ao_grad_lap_vs_finite_differences(name):
    # pattern B (jQMC tests/test_AOs.py:463-1190)
    H = load_system(name)
    aos = H.wavefunction_data.geminal_data.orb_data_up_spin.aos_data
    rng = np.random.default_rng(3)
    r = np.asarray(H.structure_data.positions)[0] + rng.normal(scale=0.6, size=(3, 3))
    v, gx, gy, gz, lap = P.compute_AOs_value_grad_lap(aos, r)
    np.testing.assert_allclose(v, P.compute_AOs(aos, r), rtol=1e-13, atol=1e-15)
    h = 1e-4
    fd_lap = np.zeros_like(v)
    for c, g in enumerate((gx, gy, gz)):
        e = np.zeros(3)
        e[c] = h
        fp, fm = P.compute_AOs(aos, r + e), P.compute_AOs(aos, r - e)
        np.testing.assert_allclose(g, (fp - fm) / (2 * h), rtol=2e-6, atol=2e-7)
        fd_lap += (fp - 2 * v + fm) / h**2
    np.testing.assert_allclose(lap, fd_lap, rtol=2e-5, atol=2e-5)


def test_jastrow_grad_lap_vs_finite_differences(water):
    jd = Jastrow_data(
        jastrow_one_body_data=Jastrow_one_body_data(
            jastrow_1b_param=0.9, structure_data=water.structure_data, core_electrons=tuple(water.coulomb_potential_data.z_cores)
        ),
        jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.7, jastrow_2b_type="exp"),
    )
    r_up, r_dn = (x[0] for x in random_walkers(water, 1, 5))
    gu, gd, lu, ld = P.compute_grads_and_laplacian_Jastrow_part(jd, r_up, r_dn)
    h = 1e-4
    J0 = P.compute_Jastrow_part(jd, r_up, r_dn)
    for i in range(len(r_up)):
        lap = 0.0
        for c in range(3):
            p, m = r_up.copy(), r_up.copy()
            p[i, c] += h
            m[i, c] -= h
            Jp, Jm = P.compute_Jastrow_part(jd, p, r_dn), P.compute_Jastrow_part(jd, m, r_dn)
            np.testing.assert_allclose(gu[i, c], (Jp - Jm) / (2 * h), rtol=1e-6, atol=1e-8)
            lap += (Jp - 2 * J0 + Jm) / h**2
        np.testing.assert_allclose(lu[i], lap, rtol=1e-5, atol=1e-5)


def test_geminal_mo_equals_ao_representation(water):
    # JSD in the MO basis and its AO-basis (JAGP-form) conversion give the same G (determinant.py:686-722)
    gem = water.wavefunction_data.geminal_data
    gem_ao = Geminal_data.convert_from_MOs_to_AOs(gem)
    gem_ao.sanity_check()
    r_up, r_dn = (x[0] for x in random_walkers(water, 1, 11))
    np.testing.assert_allclose(
        P.compute_geminal_all_elements(gem_ao, r_up, r_dn), P.compute_geminal_all_elements(gem, r_up, r_dn), rtol=1e-10, atol=1e-14
    )


def test_ln_det_grads_fast_vs_scratch_open_shell():
    H = load_system("Li_ae_ccpvdz_cart")  # 2 up, 1 dn: exercises the unpaired block
    gem = H.wavefunction_data.geminal_data
    assert gem.num_electron_up == 2 and gem.num_electron_dn == 1
    r_up, r_dn = (x[0] for x in random_walkers(H, 1, 2))
    G = P.compute_geminal_all_elements(gem, r_up, r_dn)
    a = P.compute_grads_and_laplacian_ln_Det(gem, r_up, r_dn, np.linalg.inv(G))
    b = P.compute_grads_and_laplacian_ln_Det(gem, r_up, r_dn)
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y, rtol=1e-9, atol=1e-12)
    # gradient of ln|det| by finite differences
    h = 1e-5
    for i in range(2):
        for c in range(3):
            p, m = r_up.copy(), r_up.copy()
            p[i, c] += h
            m[i, c] -= h
            fd = (P.compute_ln_det(gem, p, r_dn) - P.compute_ln_det(gem, m, r_dn)) / (2 * h)
            np.testing.assert_allclose(a[0][i, c], fd, rtol=1e-6, atol=1e-7)


def test_ecp_mesh_layout(water):
    r_up, r_dn = (x[0] for x in random_walkers(water, 1, 4))
    mu, md, v, s = P.compute_ecp_non_local_parts_nearest_neighbors(
        water.coulomb_potential_data, water.wavefunction_data, r_up, r_dn, np.eye(3), NN=1, Nv=6
    )
    assert mu.shape == (48, 4, 3) and md.shape == (48, 4, 3) and v.shape == (48,)
    np.testing.assert_allclose(s, v.sum())
    # first 24 configurations move up electrons only
    assert np.all(md[:24] == r_dn) and np.all(mu[24:] == r_up)


@pytest.mark.parametrize("name", ["w_2b_3b_w_ecp", "w_2b_1b3b_w_ecp", "w_1b_2b_1b3b_ae"])
def test_turborvb_three_body_jastrow_known_answers(name):
    """The J3 (and J1) restatement against the TurboRVB numbers of the reference's tests
    (tests/test_comparison_with_turborvb_ECP.py:376-416, 518-558; tests/test_comparison_with_turborvb_AE.py:163-271), with
    the Jastrow factors recovered from the TurboRVB wavefunction files (tools/turbo_jastrow.py).  Reference tolerances."""
    from tests.conftest import turbo_j3_case

    H, up, dn, new_up, new_dn, spin, idx, ratio_ref, kin_ref, v_ref = turbo_j3_case(name)
    wf = H.wavefunction_data
    ratio = (P.evaluate_wavefunction(wf, new_up, new_dn) / P.evaluate_wavefunction(wf, up, dn)) ** 2
    np.testing.assert_almost_equal(ratio, ratio_ref, decimal=6)
    np.testing.assert_almost_equal(P.compute_kinetic_energy(wf, new_up, new_dn), kin_ref, decimal=6)
    V = P.compute_coulomb_potential(H.coulomb_potential_data, wf, new_up, new_dn, RT=np.eye(3), NN=1, Nv=6)
    np.testing.assert_almost_equal(V, v_ref, decimal=5 if H.coulomb_potential_data.ecp_flag else 2)
    # incremental Jastrow ratio == brute force (reference test pattern B for a17)
    new = (new_up if spin == "up" else new_dn)[idx]
    jr = P.jastrow_ratio(wf.jastrow_data, up, dn, spin == "up", idx, new)
    dr = P.wf_ratio_brute_force(wf, up, dn, spin == "up", idx, new, det_only=True)
    np.testing.assert_allclose((jr * dr) ** 2, ratio, rtol=1e-11)


def test_turborvb_full_metropolis_known_answers():
    """tests/test_comparison_with_turborvb_ECP.py:660-945: proposal factors f_a, f_b, T_ratio, geminal matrices before / after
    the move, AS regularisation (S, F, R_AS with epsilon = 0.3), regularised WF ratio^2, final acceptance ratio, kinetic
    energy and potential -- the arithmetic of one Metropolis proposal (a28) and of the AS factor (a12)."""
    import copy

    from oracle import drivers as OD
    from tests.conftest import TURBO_FULL as T, load_turbo_jastrow

    H = copy.deepcopy(load_system("water_ccecp_ccpvqz"))
    H.wavefunction_data.jastrow_data = load_turbo_jastrow("w_2b_1b3b_w_ecp", H.structure_data)
    wf, gem = H.wavefunction_data, H.wavefunction_data.geminal_data
    up, dn = np.array(T["old_up"]), np.array(T["old_dn"])
    new_up = up.copy()
    new_up[2] = T["new_up2"]
    Dt, eps = 2.0, 0.30
    fa, fb = OD._f_l(H, up[2]), OD._f_l(H, new_up[2])
    np.testing.assert_almost_equal(fa, T["fa"], decimal=6)
    np.testing.assert_almost_equal(fb, T["fb"], decimal=6)
    d2 = np.sum((new_up[2] - up[2]) ** 2)
    T_ratio = (fa / fb) * np.exp(-d2 * (1.0 / (2.0 * fb**2 * Dt**2) - 1.0 / (2.0 * fa**2 * Dt**2)))
    np.testing.assert_almost_equal(T_ratio, T["T_ratio"], decimal=6)
    G_old, G_new = P.compute_geminal_all_elements(gem, up, dn), P.compute_geminal_all_elements(gem, new_up, dn)
    np.testing.assert_almost_equal(G_old, np.array(T["geminal_old_T"]).T, decimal=6)
    np.testing.assert_almost_equal(G_new, np.array(T["geminal_new_T"]).T, decimal=6)
    R_old = P.compute_AS_regularization_factor(G_old, np.linalg.inv(G_old))
    R_new = P.compute_AS_regularization_factor(G_new, np.linalg.inv(G_new))
    np.testing.assert_almost_equal(R_old, T["R_AS_old"], decimal=6)
    np.testing.assert_almost_equal(R_new, T["R_AS_new"], decimal=6)
    np.testing.assert_almost_equal(np.sum(np.linalg.inv(G_old) ** 2) / T["F_old"], 1.0, decimal=6)
    ratio = (P.evaluate_wavefunction(wf, new_up, dn) / P.evaluate_wavefunction(wf, up, dn)) ** 2
    ratio *= ((max(R_new, eps) / R_new) / (max(R_old, eps) / R_old)) ** 2
    np.testing.assert_almost_equal(ratio, T["WF_ratio"], decimal=6)
    np.testing.assert_almost_equal(ratio * T_ratio, T["final_ratio"], decimal=6)
    np.testing.assert_almost_equal((R_new / max(R_new, eps)) ** 2, T["reweight"], decimal=6)
    np.testing.assert_almost_equal(P.compute_kinetic_energy(wf, new_up, dn), T["kinc"], decimal=6)
    V = P.compute_coulomb_potential(H.coulomb_potential_data, wf, new_up, dn, RT=np.eye(3), NN=1, Nv=6)
    np.testing.assert_almost_equal(V, T["vpot"] + T["vpotoff"], decimal=5)


def test_parameter_derivatives_match_finite_differences():
    """O_k = d ln|Psi| / d parameter (the reference: jax.grad of evaluate_ln_wavefunction_fast): the analytic restatement
    against central finite differences of the oracle's own ln|Psi| for every block (J1, J2, J3 matrix incl. one-body column,
    lambda incl. the unpaired column) -- reference test pattern C."""
    import copy
    import dataclasses

    from jqmc_b200.data import Jastrow_three_body_data
    from tests.conftest import load_turbo_jastrow

    H = copy.deepcopy(load_system("Li_ae_ccpvdz_cart"))  # 2 up, 1 down: lambda has an unpaired column
    aos = H.wavefunction_data.geminal_data.orb_data_up_spin.aos_data
    rng = np.random.default_rng(3)
    from tests.test_gpu_wide import sub_basis

    j3b = sub_basis(aos, 1)
    n = j3b.num_ao
    jd = Jastrow_data(
        jastrow_one_body_data=Jastrow_one_body_data(jastrow_1b_param=0.8, jastrow_1b_type="exp", structure_data=H.structure_data, core_electrons=(0.0,)),
        jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.7, jastrow_2b_type="pade"),
        jastrow_three_body_data=Jastrow_three_body_data(orb_data=j3b, j_matrix=rng.normal(scale=0.05, size=(n, n + 1))),
    )
    H.wavefunction_data.jastrow_data = jd
    wf = H.wavefunction_data
    r_up, r_dn = random_walkers(H, 1, 4)
    r_up, r_dn = r_up[0], r_dn[0]
    g = P.compute_dln_wf_dparams(wf, r_up, r_dn)

    def fd(setter, h):
        vals = []
        for s in (+1, -1):
            w2 = copy.deepcopy(wf)
            setter(w2, s * h)
            vals.append(P.evaluate_ln_wavefunction(w2, r_up, r_dn))
        return (vals[0] - vals[1]) / (2 * h)

    def set_j1(w, d):
        w.jastrow_data.jastrow_one_body_data = dataclasses.replace(w.jastrow_data.jastrow_one_body_data, jastrow_1b_param=0.8 + d)

    def set_j2(w, d):
        w.jastrow_data.jastrow_two_body_data = dataclasses.replace(w.jastrow_data.jastrow_two_body_data, jastrow_2b_param=0.7 + d)

    np.testing.assert_allclose(g["j1_param"], fd(set_j1, 1e-5), rtol=1e-7)
    np.testing.assert_allclose(g["j2_param"], fd(set_j2, 1e-5), rtol=1e-7)
    for idx in [(0, 0), (2, 1), (1, 3), (n - 1, n), (0, n)]:
        def set_j3(w, d, idx=idx):
            m = np.array(w.jastrow_data.jastrow_three_body_data.j_matrix)
            m[idx] += d
            w.jastrow_data.jastrow_three_body_data = dataclasses.replace(w.jastrow_data.jastrow_three_body_data, j_matrix=m)

        np.testing.assert_allclose(g["j3_matrix"][idx], fd(set_j3, 1e-5), rtol=1e-6, atol=1e-9)
    lam_shape = np.shape(wf.geminal_data.lambda_matrix)
    for idx in [(0, 0), (1, 0), (0, 1), (lam_shape[0] - 1, lam_shape[1] - 1), (1, lam_shape[1] - 1)]:
        def set_lam(w, d, idx=idx):
            m = np.array(w.geminal_data.lambda_matrix)
            m[idx] += d
            w.geminal_data = dataclasses.replace(w.geminal_data, lambda_matrix=m)

        np.testing.assert_allclose(g["lambda_matrix"][idx], fd(set_lam, 1e-6), rtol=1e-6, atol=1e-8)


def test_move_selection_does_not_depend_on_the_summation_order_of_the_normaliser():
    """The reference normalises the move probabilities with ``p_list.sum()`` (jqmc_gfmc.py:5025; order unspecified under XLA); the
    engine sums them sequentially.  Both orders must select the same moves and give the same weights to round-off."""
    import copy

    from jqmc_b200.data import Jastrow_data, Jastrow_two_body_data
    from oracle import drivers as OD
    from tests.conftest import load_system, random_walkers

    H = copy.deepcopy(load_system("H2_ecp_ccpvtz"))
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.8))
    r_up, r_dn = random_walkers(H, 4, 41, scale=0.7)
    for w in range(4):
        _, Ginv = OD.geminal_inv(H.wavefunction_data.geminal_data, r_up[w], r_dn[w])
        ta, tb = [], []
        a = OD.lrdmc_projection(H, 1.0, r_up[w], r_dn[w], Ginv, (0, 777 + w), -1.1, 12, True, "tmove", 0.3, trace=ta, norm_order="reference")
        b = OD.lrdmc_projection(H, 1.0, r_up[w], r_dn[w], Ginv, (0, 777 + w), -1.1, 12, True, "tmove", 0.3, trace=tb, norm_order="sequential")
        assert [t["k"] for t in ta] == [t["k"] for t in tb]
        np.testing.assert_allclose(a[0], b[0], rtol=1e-13)
        np.testing.assert_array_equal(a[1], b[1])
