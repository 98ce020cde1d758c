"""Restart checkpoints (jQMC layout, jqmc/_checkpoint.py) written and read by the replacement drivers, without a GPU:
the drivers run on the oracle-backed engine double.  Covers the HDF5 subset writer/reader, the dataclass tree, the per-rank /
merged file layout and -- the point of it -- that save -> load -> run continues a chain bit-identically."""

import copy
import os

import numpy as np
import pytest
import torch

from jqmc_b200 import hdf5_lite as H5
from jqmc_b200.checkpoint import (
    load_checkpoint_meta, load_hamiltonian_from_checkpoint, load_rank_checkpoint, merge_rank_checkpoints, save_checkpoint,
)  # fmt: skip
from jqmc_b200.data import Jastrow_data, Jastrow_one_body_data, Jastrow_two_body_data
from tests.conftest import load_system
from tests.oracle_engine import OracleEngine


def test_hdf5_writer_round_trip(tmp_path):
    p = str(tmp_path / "t.h5")
    rng = np.random.default_rng(0)
    a = rng.normal(size=(3, 4, 5))
    with H5.File(p, "w") as f:
        g = f.create_group("rank_0/driver_config")
        g.attrs["mcmc_seed"] = 34456
        g.attrs["E_scf"] = -17.2
        g.attrs["non_local_move"] = "tmove"
        g.attrs["flag"] = True
        o = f.create_group("rank_0/observables")
        o.create_dataset("e_L", data=a)
        o.create_dataset("keys", data=np.arange(8, dtype=np.uint32).reshape(4, 2))
        o.create_dataset("i64", data=np.array([-3, 2**40], dtype=np.int64))
        o.create_dataset("f32", data=np.array([1.5, -2.25], dtype=np.float32))
        o.create_dataset("names", data=["O", "H", "Hx"])
        o.create_dataset("empty", data=np.empty(0))
        o.create_dataset("scalar", data=2.5)
        many = f.create_group("many")  # more links than one symbol-table node holds (64): several nodes under the B-tree
        for i in range(150):
            many.create_dataset(f"d{i:03d}", data=np.array([i, i + 1]))
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and len(raw) % 8 == 0
    with H5.File(p, "r") as f:
        assert sorted(f.keys()) == ["many", "rank_0"]
        at = f["rank_0"]["driver_config"].attrs
        assert at["mcmc_seed"] == 34456 and at["E_scf"] == -17.2 and at["non_local_move"] == "tmove" and at["flag"] == 1
        ob = f["rank_0/observables"]
        np.testing.assert_array_equal(ob["e_L"][()], a)
        assert ob["keys"][()].dtype == np.uint32 and ob["i64"][()][1] == 2**40 and ob["f32"][()].dtype == np.float32
        assert ob["names"][()] == ["O", "H", "Hx"] and ob["empty"][()].shape == (0,) and ob["scalar"][()] == 2.5
        assert len(list(f["many"].keys())) == 150
        np.testing.assert_array_equal(f["many"]["d149"][()], [149, 150])
        assert "nope" not in f["many"] and "d007" in f["many"]


def _small_system():
    Hm = copy.deepcopy(load_system("H2_ecp_ccpvtz"))
    cp = Hm.coulomb_potential_data
    Hm.wavefunction_data.jastrow_data = Jastrow_data(
        jastrow_one_body_data=Jastrow_one_body_data(jastrow_1b_param=0.9, jastrow_1b_type="exp", structure_data=Hm.structure_data, core_electrons=tuple(cp.z_cores)),
        jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.7, jastrow_2b_type="pade"),
    )  # fmt: skip
    return Hm


def _assert_same_tree(a, b, path="H"):
    import dataclasses

    if dataclasses.is_dataclass(a):
        assert type(a).__name__ == type(b).__name__, path
        for f in dataclasses.fields(a):
            _assert_same_tree(getattr(a, f.name), getattr(b, f.name), f"{path}.{f.name}")
    elif a is None:
        assert b is None, path
    elif isinstance(a, (np.ndarray, list, tuple)):
        assert np.array_equal(np.asarray(a), np.asarray(b)), path
    else:
        assert a == b, path


def test_hamiltonian_tree_round_trip(tmp_path):
    Hm = _small_system()
    p = str(tmp_path / "restart.h5")
    tmp = str(tmp_path / "._r{rank}.h5")
    from jqmc_b200.checkpoint import save_rank_checkpoint

    save_rank_checkpoint(tmp.format(rank=0), driver_type="MCMC", driver_config=dict(mcmc_seed=1, Dt=2.0, flag=True, name="x"),
                         rng_state=dict(jax_PRNG_key_list=np.zeros((2, 2), np.uint32), mpi_seed=1),
                         walker_state=dict(latest_r_up_carts=np.zeros((2, 1, 3))), observables=dict(e_L=np.ones((3, 2)), none=np.empty(0)))  # fmt: skip
    merge_rank_checkpoints(p, mpi_size=1, driver_type="MCMC", hamiltonian_data=Hm, tmp_pattern=tmp)
    assert not os.path.exists(tmp.format(rank=0))
    meta = load_checkpoint_meta(p)
    assert meta["format_version"] == "1.0" and meta["driver_type"] == "MCMC" and meta["mpi_size"] == 1
    H2 = load_hamiltonian_from_checkpoint(p)
    _assert_same_tree(Hm, H2)
    assert isinstance(H2.structure_data.atomic_numbers, tuple) and isinstance(H2.wavefunction_data.geminal_data.lambda_matrix, np.ndarray)
    d = load_rank_checkpoint(p, 0)
    assert d["driver_config"] == dict(mcmc_seed=1, Dt=2.0, flag=1, name="x") and "none" not in d["observables"]
    with H5.File(p, "r") as f:  # the reference's loader imports `_module_name`: it must name jQMC's modules
        assert f["hamiltonian_data"].attrs["_module_name"] == "jqmc.hamiltonians"
        assert f["hamiltonian_data/wavefunction_data/geminal_data"].attrs["_class_name"] == "Geminal_data"
        assert f["hamiltonian_data/wavefunction_data/geminal_data"].attrs["_module_name"] == "jqmc.determinant"


@pytest.mark.parametrize("kind", ["GFMC_n", "GFMC_t", "MCMC"])
def test_save_load_continue_is_identical(tmp_path, kind, monkeypatch):
    """run(a); run(b) == run(a); save; load; run(b): walkers, keys, stored observables, counters."""
    from jqmc_b200.gfmc import GFMC_n, GFMC_t
    from jqmc_b200.mcmc import MCMC

    monkeypatch.chdir(tmp_path)
    Hm = _small_system()
    eng = OracleEngine(Hm)
    if kind == "GFMC_n":
        mk = lambda: GFMC_n(Hm, num_walkers=3, num_mcmc_per_measurement=2, num_gfmc_collect_steps=1, mcmc_seed=11, E_scf=-1.2, alat=0.4, engine=eng)  # noqa: E731
        cls = GFMC_n
    elif kind == "GFMC_t":
        mk = lambda: GFMC_t(Hm, num_walkers=3, num_gfmc_collect_steps=1, mcmc_seed=11, tau=0.05, alat=0.4, engine=eng)  # noqa: E731
        cls = GFMC_t
    else:
        mk = lambda: MCMC(Hm, mcmc_seed=11, num_walkers=3, num_mcmc_per_measurement=4, Dt=2.0, epsilon_AS=0.0, comput_log_WF_param_deriv=True, engine=eng)  # noqa: E731
        cls = MCMC
    full = mk()  # run() refreshes the inverse (and the comb-offset stream) at its start, as the reference does: compare like with like
    full.run(num_mcmc_steps=3)
    full.run(num_mcmc_steps=2)
    part = mk()
    part.run(num_mcmc_steps=3)
    save_checkpoint(part, "restart.h5")
    assert load_checkpoint_meta("restart.h5")["driver_type"] == kind
    rest = cls.load_from_hdf5("restart.h5", engine=eng)
    assert rest.mcmc_counter == part.mcmc_counter
    np.testing.assert_array_equal(rest.latest_r_up_carts, part.latest_r_up_carts)
    np.testing.assert_array_equal(rest.jax_PRNG_key_list, part.jax_PRNG_key_list)
    rest.run(num_mcmc_steps=2)
    assert isinstance(rest.latest_r_up_carts, np.ndarray) and rest.jax_PRNG_key_list.dtype == np.uint32
    np.testing.assert_array_equal(rest.latest_r_up_carts, full.latest_r_up_carts)
    np.testing.assert_array_equal(rest.latest_r_dn_carts, full.latest_r_dn_carts)
    np.testing.assert_array_equal(rest.jax_PRNG_key_list, full.jax_PRNG_key_list)
    np.testing.assert_array_equal(rest.e_L, full.e_L)
    np.testing.assert_array_equal(rest.w_L, full.w_L)
    if kind == "MCMC":
        assert rest.accepted_moves == full.accepted_moves
        for k, v in full.dln_Psi_dc.items():
            np.testing.assert_array_equal(rest.dln_Psi_dc[k], v)
    else:
        assert rest.num_survived_walkers == full.num_survived_walkers
        np.testing.assert_array_equal(rest.bare_w_L, full.bare_w_L)
