"""TREXIO arrays -> engine data model (host logic, NumPy only).

Restates the coefficient arithmetic of ``read_trexio_file`` (jqmc/trexio_wrapper.py:371-424 for
spherical shells, :242-330 for Cartesian shells, :452-523 for the MO/λ block, :525-560 for ECPs)
on a plain ``dict`` of the raw TREXIO datasets/attributes, so that the same input files drive the
engine without ``trexio``/``h5py``.  The ``dict`` is what ``tools/make_golden.py`` stores in
``tests/golden/*.npz`` and what ``tools/mini_hdf5.py`` extracts from a TREXIO HDF5 file.
"""

from __future__ import annotations

import itertools
from math import factorial, pi, sqrt

import numpy as np

from .data import (
    AOs_cart_data,
    AOs_sphe_data,
    Coulomb_potential_data,
    Geminal_data,
    Hamiltonian_data,
    Jastrow_data,
    MOs_data,
    Structure_data,
    Wavefunction_data,
)

_SYMBOLS = (
    "H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
    "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb "
    "Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn"
).split()
_Z = {s: i + 1 for i, s in enumerate(_SYMBOLS)}

TREXIO_KEYS = (
    "nucleus_coord nucleus_label nucleus_charge electron_up_num electron_dn_num ao_cartesian ao_num "
    "ao_normalization basis_shell_num basis_nucleus_index basis_shell_ang_mom basis_shell_factor "
    "basis_shell_index basis_exponent basis_coefficient basis_prim_factor mo_coefficient mo_occupation "
    "mo_spin ecp_num ecp_max_ang_mom_plus_1 ecp_z_core ecp_ang_mom ecp_nucleus_index ecp_exponent "
    "ecp_coefficient ecp_power"
).split()


def read_trexio_arrays(path: str) -> dict:
    """Extract the datasets ``read_trexio_file`` consumes from a TREXIO HDF5 file (build-box tool)."""
    from tools.mini_hdf5 import MiniHDF5  # build/fixture tooling, not needed at run time

    f = MiniHDF5(path)
    out = {}
    for key in TREXIO_KEYS:
        grp = "/" + key.split("_")[0]
        if f.has(f"{grp}/{key}"):
            out[key] = f.read(f"{grp}/{key}")
        else:
            a = f.attrs(grp) if f.has(grp) else {}
            if key in a:
                out[key] = a[key]
    if "mo_spin" not in out:
        out["mo_spin"] = np.zeros(len(out["mo_occupation"]), dtype=np.int64)
    return out


def hamiltonian_from_trexio_arrays(t: dict) -> Hamiltonian_data:
    """Build ``Hamiltonian_data`` (JSD, no Jastrow) from raw TREXIO arrays.  See module docstring."""
    labels = [str(x) for x in np.atleast_1d(t["nucleus_label"])]
    coords = np.asarray(t["nucleus_coord"], dtype=np.float64)
    structure = Structure_data(
        positions=coords,
        atomic_numbers=tuple(_Z[s] for s in labels),
        element_symbols=tuple(labels),
        atomic_labels=tuple(labels),
    )
    n_up, n_dn = int(t["electron_up_num"]), int(t["electron_dn_num"])
    shell_nuc = np.asarray(t["basis_nucleus_index"], dtype=np.int64)
    shell_l = np.asarray(t["basis_shell_ang_mom"], dtype=np.int64)
    shell_fac = np.asarray(t["basis_shell_factor"], dtype=np.float64)
    prim_shell = np.asarray(t["basis_shell_index"], dtype=np.int64)
    prim_exp = np.asarray(t["basis_exponent"], dtype=np.float64)
    prim_coef = np.asarray(t["basis_coefficient"], dtype=np.float64)
    prim_fac = np.asarray(t["basis_prim_factor"], dtype=np.float64)
    ao_norm = np.asarray(t["ao_normalization"], dtype=np.float64)
    cart = bool(int(t["ao_cartesian"]))

    nucleus_index, l_list, m_list, px, py, pz = [], [], [], [], [], []
    orbital_indices, exponents, coefficients = [], [], []
    n_ao = 0
    for s in range(int(t["basis_shell_num"])):
        l = int(shell_l[s])
        prims = np.nonzero(prim_shell == s)[0]
        if cart:
            orders = ["".join(p) for p in itertools.combinations_with_replacement("xyz", l)]
            nx = [o.count("x") for o in orders]
            ny = [o.count("y") for o in orders]
            nz = [o.count("z") for o in orders]
            nfun = len(orders)
            fpart = [
                factorial(a) * factorial(b) * factorial(c) / (factorial(2 * a) * factorial(2 * b) * factorial(2 * c))
                for a, b, c in zip(nx, ny, nz)
            ]
            zpart = [(2.0 * prim_exp[k] / pi) ** 1.5 * (8.0 * prim_exp[k]) ** l for k in prims]
            for p in range(nfun):
                for i, k in enumerate(prims):
                    c = shell_fac[s] * prim_fac[k] / np.sqrt(zpart[i] * fpart[p]) * prim_coef[k]
                    coefficients.append(c * ao_norm[n_ao + p])
                    exponents.append(prim_exp[k])
                    orbital_indices.append(n_ao + p)
            px += nx
            py += ny
            pz += nz
        else:
            mags = [0] + [i * (-1) ** j for i in range(1, l + 1) for j in range(2)]
            nfun = len(mags)
            norms = [
                sqrt(2.0 ** (2 * l + 3) * factorial(l + 1) * (2 * prim_exp[k]) ** (l + 1.5) / (factorial(2 * l + 2) * sqrt(pi)))
                for k in prims
            ]
            base = [
                shell_fac[s] * prim_fac[k] * sqrt(4 * pi) / sqrt(2 * l + 1) / norms[i] * prim_coef[k]
                for i, k in enumerate(prims)
            ]
            for p in range(nfun):
                for i, k in enumerate(prims):
                    coefficients.append(base[i] * ao_norm[n_ao + p])
                    exponents.append(prim_exp[k])
                    orbital_indices.append(n_ao + p)
            m_list += mags
        nucleus_index += [int(shell_nuc[s])] * nfun
        l_list += [l] * nfun
        n_ao += nfun
    if n_ao != int(t["ao_num"]):
        raise ValueError(f"ao_num_count = {n_ao} is inconsistent with the read ao_num = {int(t['ao_num'])}")

    common = dict(
        structure_data=structure,
        nucleus_index=tuple(nucleus_index),
        num_ao=n_ao,
        num_ao_prim=len(exponents),
        angular_momentums=tuple(l_list),
        orbital_indices=tuple(orbital_indices),
        exponents=np.asarray(exponents, dtype=np.float64),
        coefficients=np.asarray(coefficients, dtype=np.float64),
    )
    if cart:
        aos = AOs_cart_data(
            polynominal_order_x=tuple(px), polynominal_order_y=tuple(py), polynominal_order_z=tuple(pz), **common
        )
    else:
        aos = AOs_sphe_data(magnetic_quantum_numbers=tuple(m_list), **common)

    # MOs / lambda (trexio_wrapper.py:452-523)
    mo_c = np.asarray(t["mo_coefficient"], dtype=np.float64)
    mo_occ = np.asarray(t["mo_occupation"], dtype=np.float64)
    mo_spin = np.asarray(t["mo_spin"], dtype=np.int64)
    thr = 1.0e-6
    spin_dep = not np.all(mo_spin == 0)
    if not spin_dep:
        idx = np.nonzero(mo_spin == 0)[0]
        c_all = mo_c[idx]
        keep = np.nonzero(mo_occ[idx] >= thr)[0]
        c_up = c_dn = c_all[keep]
    elif n_up != n_dn:
        iu = np.nonzero(mo_spin == 0)[0]
        idn = np.nonzero(mo_spin == 1)[0]
        ku = np.nonzero(mo_occ[iu] >= thr)[0]
        kd = np.nonzero(mo_occ[idn] >= thr)[0]
        if len(ku) < len(kd):
            raise ValueError("The number of occ. orbitals for up spins should be larger than those of down spins.")
        c_up = mo_c[iu][ku]
        c_dn = mo_c[idn][ku]
    else:
        raise NotImplementedError
    n_mo = c_up.shape[0]
    diff = n_up - n_dn
    lam_p = np.pad(np.eye(n_dn), ((0, n_mo - n_dn), (0, n_mo - n_dn)))
    lam_u = np.pad(np.eye(diff), ((n_dn, n_mo - n_dn - diff), (0, 0)))
    geminal = Geminal_data(
        num_electron_up=n_up,
        num_electron_dn=n_dn,
        orb_data_up_spin=MOs_data(num_mo=n_mo, aos_data=aos, mo_coefficients=c_up),
        orb_data_dn_spin=MOs_data(num_mo=n_mo, aos_data=aos, mo_coefficients=c_dn),
        lambda_matrix=np.hstack([lam_p, lam_u]),
    )

    if "ecp_num" in t and int(t["ecp_num"]) > 0:
        coulomb = Coulomb_potential_data(
            structure_data=structure,
            ecp_flag=True,
            z_cores=tuple(int(x) for x in t["ecp_z_core"]),
            max_ang_mom_plus_1=tuple(int(x) for x in t["ecp_max_ang_mom_plus_1"]),
            num_ecps=int(t["ecp_num"]),
            ang_moms=tuple(int(x) for x in t["ecp_ang_mom"]),
            nucleus_index=tuple(int(x) for x in t["ecp_nucleus_index"]),
            exponents=tuple(float(x) for x in t["ecp_exponent"]),
            coefficients=tuple(float(x) for x in t["ecp_coefficient"]),
            powers=tuple(int(x) + 2 for x in t["ecp_power"]),
        )
    else:
        coulomb = Coulomb_potential_data(structure_data=structure, ecp_flag=False)

    return Hamiltonian_data(
        structure_data=structure,
        coulomb_potential_data=coulomb,
        wavefunction_data=Wavefunction_data(jastrow_data=Jastrow_data(), geminal_data=geminal),
    )


def load_golden_system(npz_path: str) -> Hamiltonian_data:
    """Load a ``tests/golden/*.npz`` TREXIO extract (written by tools/make_golden.py)."""
    with np.load(npz_path, allow_pickle=False) as z:
        t = {k: z[k] for k in z.files}
    for k in ("electron_up_num", "electron_dn_num", "ao_cartesian", "ao_num", "basis_shell_num", "ecp_num"):
        if k in t:
            t[k] = int(t[k])
    return hamiltonian_from_trexio_arrays(t)
