"""Host-side data model of the walker engine: plain-Python mirrors of jQMC's L1 dataclasses.

Field names and array conventions follow the reference so that a reference object (a
``flax.struct`` dataclass) and these mirrors are interchangeable for the engine, which only reads
attributes (duck typing):

* ``Structure_data``          jqmc/structure.py:65-114
* ``AOs_sphe_data``           jqmc/atomic_orbital.py:780-929
* ``AOs_cart_data``           jqmc/atomic_orbital.py:87-258
* ``MOs_data``                jqmc/molecular_orbital.py:85-140
* ``Geminal_data``            jqmc/determinant.py:93-120
* ``Jastrow_*_data``          jqmc/jastrow_factor.py:560-600, 1110-1141, 1316-1335, 1864-1893
* ``Coulomb_potential_data``  jqmc/coulomb_potential.py:187-233
* ``Wavefunction_data``       jqmc/wavefunction.py:247-265
* ``Hamiltonian_data``        jqmc/hamiltonians.py:82-116

No JAX, no pytrees: the engine flattens these into device tables once (``jqmc_b200.engine``).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Union

import numpy as np


@dataclass
class Structure_data:
    positions: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    pbc_flag: bool = False
    vec_a: Sequence[float] = ()
    vec_b: Sequence[float] = ()
    vec_c: Sequence[float] = ()
    atomic_numbers: Sequence[int] = ()
    element_symbols: Sequence[str] = ()
    atomic_labels: Sequence[str] = ()

    @property
    def natom(self) -> int:
        return len(self.atomic_numbers)


@dataclass
class AOs_sphe_data:
    structure_data: Structure_data = field(default_factory=Structure_data)
    nucleus_index: Sequence[int] = ()
    num_ao: int = 0
    num_ao_prim: int = 0
    angular_momentums: Sequence[int] = ()
    magnetic_quantum_numbers: Sequence[int] = ()
    orbital_indices: Sequence[int] = ()
    exponents: np.ndarray = field(default_factory=lambda: np.zeros(0))
    coefficients: np.ndarray = field(default_factory=lambda: np.zeros(0))


@dataclass
class AOs_cart_data:
    structure_data: Structure_data = field(default_factory=Structure_data)
    nucleus_index: Sequence[int] = ()
    num_ao: int = 0
    num_ao_prim: int = 0
    angular_momentums: Sequence[int] = ()
    polynominal_order_x: Sequence[int] = ()
    polynominal_order_y: Sequence[int] = ()
    polynominal_order_z: Sequence[int] = ()
    orbital_indices: Sequence[int] = ()
    exponents: np.ndarray = field(default_factory=lambda: np.zeros(0))
    coefficients: np.ndarray = field(default_factory=lambda: np.zeros(0))


AOs_data = Union[AOs_sphe_data, AOs_cart_data]


@dataclass
class MOs_data:
    num_mo: int = 0
    aos_data: AOs_data = field(default_factory=AOs_sphe_data)
    mo_coefficients: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))


def is_mos(orb_data) -> bool:
    return hasattr(orb_data, "mo_coefficients")


def is_cart(aos_data) -> bool:
    return hasattr(aos_data, "polynominal_order_x")


def orb_num(orb_data) -> int:
    return int(orb_data.num_mo) if is_mos(orb_data) else int(orb_data.num_ao)


@dataclass
class Geminal_data:
    num_electron_up: int = 0
    num_electron_dn: int = 0
    orb_data_up_spin: Union[AOs_data, MOs_data] = field(default_factory=AOs_sphe_data)
    orb_data_dn_spin: Union[AOs_data, MOs_data] = field(default_factory=AOs_sphe_data)
    lambda_matrix: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))

    @property
    def orb_num_up(self) -> int:
        return orb_num(self.orb_data_up_spin)

    @property
    def orb_num_dn(self) -> int:
        return orb_num(self.orb_data_dn_spin)

    def sanity_check(self) -> None:
        """Same shape rule as jqmc/determinant.py:122-145 (raises ValueError)."""
        exp = (self.orb_num_up, self.orb_num_dn + self.num_electron_up - self.num_electron_dn)
        if tuple(np.shape(self.lambda_matrix)) != exp:
            raise ValueError(f"lambda_matrix shape {np.shape(self.lambda_matrix)} != {exp}")

    @staticmethod
    def convert_from_MOs_to_AOs(geminal_data: "Geminal_data") -> "Geminal_data":
        """MO-basis geminal -> AO-basis geminal (JSD -> JAGP form), jqmc/determinant.py:686-722.

        lambda_AO = C_up^T  lambda_MO  [C_dn | I_unpaired-block]; the value of G is unchanged.
        """
        up, dn = geminal_data.orb_data_up_spin, geminal_data.orb_data_dn_spin
        if not (is_mos(up) and is_mos(dn)):
            return geminal_data
        lam = np.asarray(geminal_data.lambda_matrix, dtype=np.float64)
        lam_p, lam_u = lam[:, : dn.num_mo], lam[:, dn.num_mo :]
        Cu = np.asarray(up.mo_coefficients, dtype=np.float64)
        Cd = np.asarray(dn.mo_coefficients, dtype=np.float64)
        ao_p = Cu.T @ lam_p @ Cd
        ao_u = Cu.T @ lam_u
        return Geminal_data(
            num_electron_up=geminal_data.num_electron_up,
            num_electron_dn=geminal_data.num_electron_dn,
            orb_data_up_spin=up.aos_data,
            orb_data_dn_spin=dn.aos_data,
            lambda_matrix=np.hstack([ao_p, ao_u]),
        )


@dataclass
class Jastrow_one_body_data:
    jastrow_1b_param: float = 1.0
    jastrow_1b_type: str = "exp"
    structure_data: Structure_data = field(default_factory=Structure_data)
    core_electrons: Sequence[float] = ()


@dataclass
class Jastrow_two_body_data:
    jastrow_2b_param: float = 1.0
    jastrow_2b_type: str = "pade"


@dataclass
class Jastrow_three_body_data:
    orb_data: Union[AOs_data, MOs_data] = field(default_factory=AOs_sphe_data)
    j_matrix: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))

    @property
    def orb_num(self) -> int:
        return orb_num(self.orb_data)


@dataclass
class Jastrow_data:
    jastrow_one_body_data: Optional[Jastrow_one_body_data] = None
    jastrow_two_body_data: Optional[Jastrow_two_body_data] = None
    jastrow_three_body_data: Optional[Jastrow_three_body_data] = None
    jastrow_nn_data: Optional[object] = None  # NN Jastrow: out of scope, rejected by the engine


@dataclass
class Wavefunction_data:
    jastrow_data: Jastrow_data = field(default_factory=Jastrow_data)
    geminal_data: Geminal_data = field(default_factory=Geminal_data)


@dataclass
class Coulomb_potential_data:
    structure_data: Structure_data = field(default_factory=Structure_data)
    ecp_flag: bool = False
    z_cores: Sequence[float] = ()
    max_ang_mom_plus_1: Sequence[int] = ()
    num_ecps: int = 0
    ang_moms: Sequence[int] = ()
    nucleus_index: Sequence[int] = ()
    exponents: Sequence[float] = ()
    coefficients: Sequence[float] = ()
    powers: Sequence[int] = ()

    @property
    def effective_charges(self) -> np.ndarray:
        """jqmc/coulomb_potential.py:312-324."""
        z = np.asarray(self.structure_data.atomic_numbers, dtype=np.float64)
        if self.ecp_flag:
            return z - np.asarray(self.z_cores, dtype=np.float64)
        return z


@dataclass
class Hamiltonian_data:
    structure_data: Structure_data = field(default_factory=Structure_data)
    coulomb_potential_data: Coulomb_potential_data = field(default_factory=Coulomb_potential_data)
    wavefunction_data: Wavefunction_data = field(default_factory=Wavefunction_data)
