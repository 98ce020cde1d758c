"""jqmc_b200 -- B200-native walker engine for jQMC's VMC / LRDMC hot path.

Only what the hot path needs: the host-side mirror of the reference's data model and
batched callables, and the CUDA kernels behind a C ABI (csrc/, include/jqmc_b200.h).
"""

from .data import (  # noqa: F401
    AOs_cart_data,
    AOs_sphe_data,
    Coulomb_potential_data,
    Geminal_data,
    Hamiltonian_data,
    Jastrow_data,
    Jastrow_one_body_data,
    Jastrow_three_body_data,
    Jastrow_two_body_data,
    MOs_data,
    Structure_data,
    Wavefunction_data,
)

__all__ = [
    "AOs_cart_data", "AOs_sphe_data", "Coulomb_potential_data", "Geminal_data", "Hamiltonian_data", "Jastrow_data",
    "Jastrow_one_body_data", "Jastrow_three_body_data", "Jastrow_two_body_data", "MOs_data", "Structure_data",
    "Wavefunction_data",
]  # fmt: skip
