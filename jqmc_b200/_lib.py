"""ctypes binding of libjqmc_b200.so (include/jqmc_b200.h).  Fails loudly when the library is missing:
there is no CPU or XLA fallback anywhere in this package."""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libjqmc_b200.so")

QE_OK, QE_ERR_INVALID, QE_ERR_UNSUPPORTED, QE_ERR_CUDA, QE_ERR_NOMEM = 0, -1, -2, -3, -4

i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


class qe_basis_desc(C.Structure):
    _fields_ = [
        ("cartesian", C.c_int32),
        ("n_ao", C.c_int32),
        ("n_prim", C.c_int32),
        ("nucleus_index", i32p),
        ("angular_momentums", i32p),
        ("magnetic_quantum_numbers", i32p),
        ("polynominal_order_x", i32p),
        ("polynominal_order_y", i32p),
        ("polynominal_order_z", i32p),
        ("orbital_indices", i32p),
        ("exponents", f64p),
        ("coefficients", f64p),
        ("n_mo", C.c_int32),
        ("mo_coefficients", f64p),
    ]


class qe_system_desc(C.Structure):
    _fields_ = [
        ("n_atom", C.c_int32),
        ("positions", f64p),
        ("effective_charges", f64p),
        ("n_up", C.c_int32),
        ("n_dn", C.c_int32),
        ("orb_up", qe_basis_desc),
        ("orb_dn", qe_basis_desc),
        ("lambda_matrix", f64p),
        ("j1_type", C.c_int32),
        ("j1_param", C.c_double),
        ("j1_core_electrons", f64p),
        ("j1_atomic_numbers", f64p),
        ("j2_type", C.c_int32),
        ("j2_param", C.c_double),
        ("j3_flag", C.c_int32),
        ("j3_orb", qe_basis_desc),
        ("j_matrix", f64p),
        ("ecp_flag", C.c_int32),
        ("n_ecp", C.c_int32),
        ("ecp_nucleus_index", i32p),
        ("ecp_ang_moms", i32p),
        ("ecp_exponents", f64p),
        ("ecp_coefficients", f64p),
        ("ecp_powers", i32p),
        ("ecp_max_ang_mom_plus_1", i32p),
        ("Nv", C.c_int32),
        ("NN", C.c_int32),
        ("precision", C.c_int32),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "qe_create": (C.c_int, [C.POINTER(qe_system_desc), C.POINTER(C.c_void_p)]),
    "qe_destroy": (None, [C.c_void_p]),
    "qe_create_from_hdf5": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "qe_hdf5_summary": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_int64), f64p]),
    "qe_hdf5_read_walkers": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, i32p, f64p, f64p, C.POINTER(C.c_uint32)]),
    "qe_last_error": (C.c_char_p, []),
    "qe_version": (C.c_int, []),
    "qe_geminal_init": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p]),
    "qe_mcmc_update": (
        C.c_int,
        [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "qe_rotation": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qe_local_energy": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_void_p]),
    "qe_nearest_nuclei": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_void_p]),
    "qe_local_energy_frozen": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_void_p]),
    "qe_as_factor": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_void_p]),
    "qe_ln_wavefunction": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p]),
    "qe_eval_orbitals": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qe_move_ratios": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, i32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "qe_lrdmc_project": (
        C.c_int,
        [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 4,
    ),
    "qe_lrdmc_project_tau": (
        C.c_int,
        [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_double, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 4,
    ),
    "qe_lrdmc_velements": (
        C.c_int,
        [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_double] + [C.c_void_p] * 3,
    ),
    "qe_lrdmc_velements_frozen": (
        C.c_int,
        [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_double] + [C.c_void_p] * 3,
    ),
    "qe_lrdmc_collect": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_double, C.c_void_p, C.c_void_p]),
    "qe_lrdmc_branch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qe_gather_walkers": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_void_p]),
    "qe_lrdmc_record_len": (C.c_int64, [C.c_void_p, C.c_int]),
    "qe_lrdmc_pack": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_void_p]),
    "qe_lrdmc_reconfigure_packed": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double] + [C.c_void_p] * 5 + [C.c_void_p],
    ),
    "qe_measure_fp64_peak": (C.c_int, [C.c_int, f64p]),
    "qe_launch_count": (C.c_int64, [C.c_void_p]),
    "qe_set_fused": (C.c_int, [C.c_void_p, C.c_int]),
    "qe_set_walkers_per_cta": (C.c_int, [C.c_void_p, C.c_int]),
    "qe_set_walker_warps": (C.c_int, [C.c_void_p, C.c_int]),
    "qe_set_wide_slice": (C.c_int, [C.c_void_p, C.c_int]),
    "qe_dln_wf": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_void_p]),
    "qe_set_path": (C.c_int, [C.c_void_p, C.c_int]),
    "qe_set_gemm_reference": (C.c_int, [C.c_void_p, C.c_int]),
    "qe_phase_clocks": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "qe_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "qe_profile_kernels": (C.c_int, []),
    "qe_profile_name": (C.c_char_p, [C.c_int]),
    "qe_profile_read": (C.c_int, [C.c_void_p, C.c_int, f64p, C.POINTER(C.c_int64)]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  jqmc_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc == QE_OK:
        return
    msg = load().qe_last_error().decode(errors="replace")
    if rc == QE_ERR_INVALID:
        raise ValueError(f"{what}: {msg}")
    if rc == QE_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    if rc == QE_ERR_NOMEM:
        raise MemoryError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")
