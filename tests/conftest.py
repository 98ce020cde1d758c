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
