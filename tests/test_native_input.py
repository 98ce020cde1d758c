"""Native input path: the C++ HDF5 reader inside libjqmc_b200.so parses jQMC's `hamiltonian_data.h5` /
`restart.h5` layout.  CPU part: the library's view of a file (counts and checksums, walker state) equals what Python wrote --
no compute call, no GPU.  GPU part: an engine created by the library from the file gives the same energies as the engine
created from the Python data model."""

import copy

import numpy as np
import pytest

from jqmc_b200 import hdf5_lite as H5
from jqmc_b200.checkpoint import merge_rank_checkpoints, save_dataclass_to_hdf5, save_rank_checkpoint
from jqmc_b200.data import Jastrow_data, Jastrow_one_body_data, Jastrow_two_body_data
from tests.conftest import load_system, load_turbo_jastrow


def _cs(v):
    v = np.asarray(v, dtype=np.float64).ravel()
    return float(np.sum(v * (1 + np.arange(len(v)) % 7)))


def _systems():
    out = {}
    H = copy.deepcopy(load_system("water_ccecp_ccpvqz"))
    cp = H.coulomb_potential_data
    H.wavefunction_data.jastrow_data = Jastrow_data(
        jastrow_one_body_data=Jastrow_one_body_data(jastrow_1b_param=0.9, jastrow_1b_type="pade", structure_data=H.structure_data, core_electrons=tuple(cp.z_cores)),
        jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.7, jastrow_2b_type="exp"),
    )  # fmt: skip
    out["water_j1j2"] = H
    H = copy.deepcopy(load_system("water_ccecp_ccpvqz"))
    H.wavefunction_data.jastrow_data = load_turbo_jastrow("w_2b_3b_w_ecp", H.structure_data)
    out["water_j3"] = H
    out["Li_cart_ae"] = copy.deepcopy(load_system("Li_ae_ccpvdz_cart"))
    return out


@pytest.mark.parametrize("name", ["water_j1j2", "water_j3", "Li_cart_ae"])
def test_library_reads_hamiltonian_tree(tmp_path, name):
    from jqmc_b200.engine import hdf5_summary

    H = _systems()[name]
    p = str(tmp_path / "hamiltonian_data.h5")
    with H5.File(p, "w") as f:
        save_dataclass_to_hdf5(f, H)  # the reference writes the tree at the root of hamiltonian_data.h5 (hamiltonians.py:118-140)
    counts, checks = hdf5_summary(p, "")
    gem, cp, jd = H.wavefunction_data.geminal_data, H.coulomb_potential_data, H.wavefunction_data.jastrow_data
    orb = gem.orb_data_up_spin
    aos = getattr(orb, "aos_data", orb)
    j1, j2, j3 = jd.jastrow_one_body_data, jd.jastrow_two_body_data, jd.jastrow_three_body_data
    j3a = getattr(j3.orb_data, "aos_data", j3.orb_data) if j3 is not None else None
    assert counts[:14] == [
        len(H.structure_data.atomic_numbers), gem.num_electron_up, gem.num_electron_dn, aos.num_ao, aos.num_ao_prim, getattr(orb, "num_mo", 0),
        int(hasattr(aos, "polynominal_order_x")), cp.num_ecps if cp.ecp_flag else 0, int(cp.ecp_flag),
        0 if j1 is None else {"exp": 1, "pade": 2}[j1.jastrow_1b_type], 0 if j2 is None else {"pade": 1, "exp": 2}[j2.jastrow_2b_type],
        int(j3 is not None), j3a.num_ao if j3a is not None else 0, getattr(j3.orb_data, "num_mo", 0) if j3 is not None else 0,
    ]  # fmt: skip
    ref = [_cs(H.structure_data.positions), _cs(cp.effective_charges), _cs(aos.exponents), _cs(aos.coefficients),
           _cs(orb.mo_coefficients) if hasattr(orb, "mo_coefficients") else 0.0, _cs(gem.lambda_matrix),
           _cs(cp.exponents) if cp.ecp_flag else 0.0, _cs(j3.j_matrix) if j3 is not None else 0.0]  # fmt: skip
    np.testing.assert_allclose(checks, ref, rtol=1e-13, atol=0)  # (sequential C++ sum vs NumPy pairwise sum)


def test_library_reads_restart_checkpoint(tmp_path):
    from jqmc_b200.engine import hdf5_read_walkers, hdf5_summary

    H = _systems()["water_j1j2"]
    rng = np.random.default_rng(1)
    tmp = str(tmp_path / "._r{rank}.h5")
    state = []
    for r in range(2):
        up, dn = rng.normal(size=(5, 4, 3)), rng.normal(size=(5, 4, 3))
        keys = rng.integers(0, 2**32, size=(5, 2), dtype=np.uint64).astype(np.uint32)
        state.append((up, dn, keys))
        save_rank_checkpoint(tmp.format(rank=r), driver_type="GFMC_n", driver_config=dict(num_walkers=5), rng_state=dict(jax_PRNG_key_list=keys, mpi_seed=r),
                             walker_state=dict(latest_r_up_carts=up, latest_r_dn_carts=dn), observables=dict(e_L=np.ones((3, 1))))  # fmt: skip
    p = str(tmp_path / "restart.h5")
    merge_rank_checkpoints(p, mpi_size=2, driver_type="GFMC_n", hamiltonian_data=H, tmp_pattern=tmp)
    counts, _ = hdf5_summary(p, "hamiltonian_data")
    assert counts[:3] == [3, 4, 4]
    assert hdf5_summary(p, "")[0] == counts  # the reader finds the tree one level down by itself
    for r in range(2):
        up, dn, keys = hdf5_read_walkers(p, r, 4, 4)
        np.testing.assert_array_equal(up, state[r][0])
        np.testing.assert_array_equal(dn, state[r][1])
        np.testing.assert_array_equal(keys, state[r][2])
    with pytest.raises(ValueError):
        hdf5_read_walkers(p, 7, 4, 4)
    with pytest.raises(ValueError):
        hdf5_summary(str(tmp_path / "missing.h5"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["water_j1j2", "water_j3"])
def test_engine_from_hdf5_matches_python_built_engine(tmp_path, name):
    from jqmc_b200.engine import WalkerEngine
    from tests.conftest import random_walkers

    H = _systems()[name]
    p = str(tmp_path / "hamiltonian_data.h5")
    with H5.File(p, "w") as f:
        save_dataclass_to_hdf5(f, H)
    a, b = WalkerEngine(H), WalkerEngine.from_hdf5(p)
    r_up, r_dn = random_walkers(H, 5, 2)
    out = []
    for eng in (a, b):
        G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
        RT = eng.generate_RTs(np.array([[0, i] for i in range(5)], dtype=np.uint32))
        out.append((eng.e_L_fast(r_up, r_dn, RT, Ginv).cpu().numpy(), eng.ln_wavefunction(r_up, r_dn)[0].cpu().numpy()))
    np.testing.assert_array_equal(out[0][0], out[1][0])  # identical tables -> bit-identical results
    np.testing.assert_array_equal(out[0][1], out[1][1])
