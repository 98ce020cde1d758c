import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def water():
    from jqmc_b200.trexio_lite import load_golden_system

    return load_golden_system(os.path.join(GOLDEN, "water_ccecp_ccpvqz.npz"))


def load_system(name):
    from jqmc_b200.trexio_lite import load_golden_system

    return load_golden_system(os.path.join(GOLDEN, name + ".npz"))


def random_walkers(H, nw, seed, scale=0.8):
    import numpy as np

    rng = np.random.default_rng(seed)
    R = np.asarray(H.structure_data.positions, dtype=np.float64)
    gem = H.wavefunction_data.geminal_data
    n_up, n_dn = gem.num_electron_up, gem.num_electron_dn
    own_u = rng.integers(0, len(R), size=(nw, n_up))
    own_d = rng.integers(0, len(R), size=(nw, n_dn))
    r_up = R[own_u] + rng.normal(scale=scale, size=(nw, n_up, 3))
    r_dn = R[own_d] + rng.normal(scale=scale, size=(nw, n_dn, 3))
    return r_up, r_dn


def load_turbo_jastrow(suffix, structure):
    """Jastrow factor of the reference's TurboRVB comparison tests (tests/golden/turbo_jastrow_<suffix>.npz, written by
    tools/turbo_jastrow.py from the reference's turborvb_WF_<suffix>.txt)."""
    import numpy as np

    from jqmc_b200.data import AOs_sphe_data, Jastrow_data, Jastrow_one_body_data, Jastrow_three_body_data, Jastrow_two_body_data

    d = np.load(os.path.join(GOLDEN, f"turbo_jastrow_{suffix}.npz"))
    ints = lambda k: tuple(int(x) for x in d[k])  # noqa: E731
    aos = AOs_sphe_data(
        structure_data=structure, nucleus_index=ints("nucleus_index"), num_ao=len(d["nucleus_index"]), num_ao_prim=len(d["exponents"]),
        angular_momentums=ints("angular_momentums"), magnetic_quantum_numbers=ints("magnetic_quantum_numbers"),
        orbital_indices=ints("orbital_indices"), exponents=np.array(d["exponents"]), coefficients=np.array(d["coefficients"]),
    )  # fmt: skip
    j1 = None
    if not np.isnan(d["j1_param"]):
        j1 = Jastrow_one_body_data(jastrow_1b_param=float(d["j1_param"]), jastrow_1b_type="exp", structure_data=structure,
                                   core_electrons=tuple(float(x) for x in d["core_electrons"]))  # fmt: skip
    return Jastrow_data(
        jastrow_one_body_data=j1,
        jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=float(d["j2_param"]), jastrow_2b_type="pade"),
        jastrow_three_body_data=Jastrow_three_body_data(orb_data=aos, j_matrix=np.array(d["j_matrix"])),
    )


# TurboRVB known answers with three-body Jastrow factors, hard-coded in the reference's tests:
#   (system, Jastrow fixture, old up, old dn, moved spin, moved index, new position, WF_ratio^2, kinetic, vpot, vpotoff)
TURBO_J3_CASES = {
    # tests/test_comparison_with_turborvb_ECP.py:376-416
    "w_2b_3b_w_ecp": ("water_ccecp_ccpvqz", "w_2b_3b_w_ecp",
        [[-1.1345038587576, -0.698914730480577, -0.006290951981744008], [-2.30366220171161, 2.32528986358581, -0.20008513679678],
         [0.390190526911041, 0.422863618938476, 1.0981171776173], [-2.4014357356045, 0.623761374394509, 0.70010581636993]],
        [[-1.58454340030273, -1.01943210665261, 0.37014437052153], [1.90701925586575, 0.398999201990364, -0.745191606127732],
         [-2.00590358216444, 2.3178763219103, -0.195294104680795], [-0.103689059569662, -2.18500664943652, -1.56814885512335]],
        "up", 2, [0.390190526911041, -0.270740090536313, 1.0981171776173],
        0.858468162763939, 5.82890200054949, -19.1676316230828, 0.284240877900265),
    # tests/test_comparison_with_turborvb_ECP.py:518-558
    "w_2b_1b3b_w_ecp": ("water_ccecp_ccpvqz", "w_2b_1b3b_w_ecp",
        [[-2.02906771233089, -0.726280132104733, -0.006290951981744008], [-0.332901524462574, 0.626165379953289, -0.60355949374895],
         [-0.197062006804461, -0.396462287261025, 0.207245244485559], [-2.13232697453793, 2.02938760506611, 0.626121128343523]],
        [[-2.27723556201111, -0.226423326809174, 0.525171318204107], [0.635659512640246, -0.128318768826431, -0.479396452798511],
         [-2.00590358216444, 1.90796788491204, -0.195294104680795], [-1.12726250654165, -0.739542218156325, -0.25704043697001]],
        "dn", 0, [-2.27723556201111, 0.7469747620327, 0.525171318204107],
        0.268078593287622, 9.84051921791642, -27.1676371839677, 0.02774284473669801),
    # tests/test_comparison_with_turborvb_AE.py:163-271 (all-electron H2, J1 + J2 + J3; potential to 2 decimals there)
    "w_1b_2b_1b3b_ae": ("H2_ae_ccpvqz", "w_1b_2b_1b3b_ae",
        [[-0.140725692347622, 1.794610704318, 0.541399181483924]], [[1.18814636744078, 0.02606967395580784, -1.62047650291381]],
        "up", 0, [0.985621336113153, 1.794610704318, 0.541399181483924],
        0.539734425254117, 0.06762960720224656, -1.22497631738529, 0.0),
}  # fmt: skip


def turbo_j3_case(name):
    import copy

    import numpy as np

    sysname, jas, up, dn, spin, idx, new, ratio, kin, vpot, vpotoff = TURBO_J3_CASES[name]
    H = copy.deepcopy(load_system(sysname))
    H.wavefunction_data.jastrow_data = load_turbo_jastrow(jas, H.structure_data)
    up, dn = np.array(up), np.array(dn)
    new_up, new_dn = up.copy(), dn.copy()
    (new_up if spin == "up" else new_dn)[idx] = new
    return H, up, dn, new_up, new_dn, spin, idx, ratio, kin, vpot + vpotoff


# The "full" TurboRVB known-answer case of the reference (tests/test_comparison_with_turborvb_ECP.py:660-945): one Metropolis
# proposal on water ccECP with J2 + J1-part-of-J3, Dt = 2, epsilon_AS = 0.3.
TURBO_FULL = dict(
    old_up=[[-1.13450385875760, -0.698914730480577, -6.290951981744008e-003], [-2.25378719009775, 0.693895756460611, -4.612006323250584e-002],
            [-0.753191857352684, 0.314330338959413, 0.456739833308641], [-1.60902246286275, 0.499927465264998, 0.700105816369930]],
    old_dn=[[-1.52590493546481, -1.13601932859996, 0.586518269898014], [0.635659512640246, 0.398999201990364, -0.745191606127732],
            [-2.00590358216444, -0.465069404417879, 0.360171216755478], [-0.302866379660751, -0.890252305196045, 0.345597836490454]],
    new_up2=[-0.753191857352684, -1.183650619518406e-002, 0.456739833308641],
    fa=0.470249568592385, fb=0.437344721251066, T_ratio=1.06518922117014, final_ratio=0.901584512996174,
    WF_ratio=0.846407844801284, kinc=13.8637480286375, vpot=-30.808883190726, vpotoff=0.168465630163985,
    R_AS_old=0.124245223553222, R_AS_new=0.116703654039403, reweight=0.151330476290540,
    F_old=53823.3438428566, S_old=4.833741852724715e-003, F_new=54231.8526090902, S_new=5.669181330133306e-003,
    geminal_old_T=[[0.186887184114679, 7.221020612173907e-003, 2.919097229181558e-002, 5.283664938570871e-002],
                   [2.711743127945612e-002, -1.872067172512451e-002, 5.968147400894821e-002, -1.363982711796792e-002],
                   [0.196787090743597, 9.308561211374290e-002, 5.042625023653007e-003, 0.135256480405370],
                   [0.152575006966341, -3.569426461507245e-002, 9.441528784169728e-002, 1.156481954187638e-002]],
    geminal_new_T=[[0.186887184114679, 7.221020612173907e-03, 7.892154586169434e-02, 5.283664938570871e-02],
                   [2.711743127945612e-02, -1.872067172512451e-02, 6.631227501216763e-02, -1.363982711796792e-02],
                   [0.196787090743597, 9.308561211374290e-02, 4.145830624511430e-02, 0.135256480405370],
                   [0.152575006966341, -3.569426461507245e-02, 0.140717058457383, 1.156481954187638e-02]],
)  # fmt: skip
