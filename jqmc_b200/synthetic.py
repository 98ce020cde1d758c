"""Synthetic Hamiltonians of a prescribed SHAPE (atoms, electrons, AO / MO counts, ECP channels, Jastrow orbital sets).

BASELINE.json configs 4 and 5 name systems for which the reference ships no input file (benzene ccECP/cc-pVTZ; a
100-electron / 1000-AO JSD molecule).  The reference's own kernel benchmark builds such systems synthetically
(benchmarks/benchmark_mcmc_kernels.py:208-446: atoms on a cubic grid, one primitive per AO, exponents per angular momentum
[5, 3, 1.5, 0.8, 0.4, 0.2], lambda = I + N(0, 0.01), three ECP terms per atom, electrons = atom centres + N(0, 1)); this
module follows that recipe so that timings are comparable in shape.  Coefficients are synthetic: energies mean nothing.
"""

from __future__ import annotations

import numpy as np

from .data import (
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

EXP_BY_L = (5.0, 3.0, 1.5, 0.8, 0.4, 0.2)


def cart_aos(structure, n_ao_per_atom):
    """Cartesian AOs, complete shells s, p, d, ... until n_ao_per_atom functions (the last shell may be partial), one
    primitive each; monomial order of a shell = itertools.combinations_with_replacement('xyz', l)."""
    spec = []
    l = 0
    while len(spec) < n_ao_per_atom:
        for nx in range(l, -1, -1):
            for ny in range(l - nx, -1, -1):
                if len(spec) < n_ao_per_atom:
                    spec.append((l, nx, ny, l - nx - ny, EXP_BY_L[min(l, len(EXP_BY_L) - 1)]))
        l += 1
    n_atom = structure.natom
    rows = [(a,) + s for a in range(n_atom) for s in spec]
    n = len(rows)
    return AOs_cart_data(
        structure_data=structure,
        nucleus_index=tuple(r[0] for r in rows),
        num_ao=n,
        num_ao_prim=n,
        angular_momentums=tuple(r[1] for r in rows),
        polynominal_order_x=tuple(r[2] for r in rows),
        polynominal_order_y=tuple(r[3] for r in rows),
        polynominal_order_z=tuple(r[4] for r in rows),
        orbital_indices=tuple(range(n)),
        exponents=np.array([r[5] for r in rows], dtype=np.float64),
        coefficients=np.ones(n, dtype=np.float64),
    )


def sphe_aos(structure, shells_by_atom, n_prim=3):
    """Spherical contracted AOs.  shells_by_atom[a] = list of l values (one entry per shell); each shell is a contraction of
    n_prim even-tempered primitives; m order inside a shell: 0, +1, -1, ... (TREXIO order, trexio_wrapper.py:375)."""
    nuc, ls, ms, oi, ex, co = [], [], [], [], [], []
    ao = 0
    for a, shells in enumerate(shells_by_atom):
        seen = {}
        for l in shells:
            k = seen.get(l, 0)
            seen[l] = k + 1
            base = EXP_BY_L[min(l, len(EXP_BY_L) - 1)] / (2.2**k)
            z = [base * (2.5**p) for p in range(n_prim)]
            c = [1.0 / (1 + p) for p in range(n_prim)]
            for m in [0] + [s * i for i in range(1, l + 1) for s in (1, -1)]:
                nuc.append(a)
                ls.append(l)
                ms.append(m)
                for zz, cc in zip(z, c):
                    oi.append(ao)
                    ex.append(zz)
                    co.append(cc)
                ao += 1
    return AOs_sphe_data(
        structure_data=structure, nucleus_index=tuple(nuc), num_ao=ao, num_ao_prim=len(oi), angular_momentums=tuple(ls),
        magnetic_quantum_numbers=tuple(ms), orbital_indices=tuple(oi), exponents=np.array(ex), coefficients=np.array(co),
    )  # fmt: skip


def _ecp(structure, z_cores, valence):
    """Three terms per ECP atom: non-local l = 0, l = 1 and the local channel l = 2 (= max_ang_mom_plus_1)."""
    ang, nuc, ex, co, pw, lmax = [], [], [], [], [], []
    for a, zc in enumerate(z_cores):
        if zc <= 0:
            lmax.append(0)
            ang.append(0)
            nuc.append(a)
            ex.append(4.0)
            co.append(0.0)
            pw.append(2)
            continue
        lmax.append(2)
        for l, e, c in ((0, 3.0, -1.0), (1, 2.0, -0.5), (2, 5.0, float(valence[a]))):
            ang.append(l)
            nuc.append(a)
            ex.append(e)
            co.append(c)
            pw.append(2)
    return Coulomb_potential_data(
        structure_data=structure, ecp_flag=True, z_cores=tuple(float(z) for z in z_cores), max_ang_mom_plus_1=tuple(lmax),
        num_ecps=len(ang), ang_moms=tuple(ang), nucleus_index=tuple(nuc), exponents=tuple(ex), coefficients=tuple(co), powers=tuple(pw),
    )  # fmt: skip


def _assemble(structure, aos_det, valence, z_cores, n_mo, j3_aos, rng, j1=True, j2=True):
    n_el = int(sum(valence))
    n_up, n_dn = (n_el + 1) // 2, n_el // 2
    if n_mo:
        q, _ = np.linalg.qr(rng.normal(size=(aos_det.num_ao, n_mo)))
        orb = MOs_data(num_mo=n_mo, aos_data=aos_det, mo_coefficients=np.ascontiguousarray(q.T))
        n_orb = n_mo
    else:
        orb, n_orb = aos_det, aos_det.num_ao
    lam = np.eye(n_orb, n_orb + n_up - n_dn) + rng.normal(0, 0.01, size=(n_orb, n_orb + n_up - n_dn))
    gem = Geminal_data(num_electron_up=n_up, num_electron_dn=n_dn, orb_data_up_spin=orb, orb_data_dn_spin=orb, lambda_matrix=lam)
    cp = _ecp(structure, z_cores, valence) if any(z > 0 for z in z_cores) else Coulomb_potential_data(structure_data=structure, ecp_flag=False)
    j3d = None
    if j3_aos is not None:
        n = j3_aos.num_ao
        M = rng.normal(0, 2e-3, size=(n, n))
        j3d = Jastrow_three_body_data(orb_data=j3_aos, j_matrix=np.hstack([0.5 * (M + M.T), rng.normal(0, 5e-3, size=(n, 1))]))
    jd = Jastrow_data(
        jastrow_one_body_data=Jastrow_one_body_data(jastrow_1b_param=1.0, jastrow_1b_type="exp", structure_data=structure,
                                                    core_electrons=tuple(float(z) for z in z_cores)) if j1 else None,
        jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=1.0, jastrow_2b_type="pade") if j2 else None,
        jastrow_three_body_data=j3d,
    )  # fmt: skip
    return Hamiltonian_data(structure_data=structure, coulomb_potential_data=cp, wavefunction_data=Wavefunction_data(jastrow_data=jd, geminal_data=gem))


def grid_molecule(n_atoms=25, valence=4, n_ao_per_atom=40, n_mo=50, ecp_core=2, spacing=3.0, j3_ao_per_atom=0, j1=False, seed=42):
    """BASELINE config 5: `n_atoms` identical atoms on a cubic grid, Cartesian AOs, optional MO layer (JSD when
    n_mo = number of up electrons).  Defaults: 100 electrons, 1000 AOs, 50 MOs, ECP with two non-local channels."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(n_atoms ** (1.0 / 3.0)))
    pos = np.array([[i, j, k] for i in range(side) for j in range(side) for k in range(side)][:n_atoms], dtype=np.float64) * spacing
    st = Structure_data(positions=pos, atomic_numbers=tuple([valence + ecp_core] * n_atoms), element_symbols=tuple(["X"] * n_atoms),
                        atomic_labels=tuple(f"X{i}" for i in range(n_atoms)))  # fmt: skip
    aos = cart_aos(st, n_ao_per_atom)
    j3 = cart_aos(st, j3_ao_per_atom) if j3_ao_per_atom else None
    return _assemble(st, aos, [valence] * n_atoms, [ecp_core] * n_atoms, n_mo, j3, rng, j1=j1)


def benzene_shape(jagp=False, seed=42):
    """BASELINE config 4 shape: C6H6 geometry, 30 valence electrons (C: He-core ECP), spherical contracted AOs
    C[3s3p2d1f] = 29, H[3s2p1d] = 14 (258 AOs), J1 + J2 + J3 with a smaller J3 orbital set C[2s1p], H[1s] (36 AOs)."""
    rng = np.random.default_rng(seed)
    rc, rh = 2.64, 2.64 + 2.05
    ang = np.arange(6) * np.pi / 3
    pos = np.vstack([np.stack([rc * np.cos(ang), rc * np.sin(ang), 0 * ang], 1), np.stack([rh * np.cos(ang), rh * np.sin(ang), 0 * ang], 1)])
    st = Structure_data(positions=pos, atomic_numbers=tuple([6] * 6 + [1] * 6), element_symbols=tuple(["C"] * 6 + ["H"] * 6),
                        atomic_labels=tuple(["C"] * 6 + ["H"] * 6))  # fmt: skip
    shells_c, shells_h = [0, 0, 0, 1, 1, 1, 2, 2, 3], [0, 0, 0, 1, 1, 2]
    aos = sphe_aos(st, [shells_c] * 6 + [shells_h] * 6)
    j3 = sphe_aos(st, [[0, 0, 1]] * 6 + [[0]] * 6, n_prim=1)
    return _assemble(st, aos, [4] * 6 + [1] * 6, [2] * 6 + [0] * 6, 0 if jagp else 15, j3, rng)


def init_walkers(H, nw, seed=0, sigma=1.0):
    """Electrons = atom centres (round robin) + N(0, sigma) (benchmarks/benchmark_mcmc_kernels.py:440-446)."""
    rng = np.random.default_rng(seed)
    pos = np.asarray(H.structure_data.positions, dtype=np.float64)
    gem = H.wavefunction_data.geminal_data
    n_up, n_dn = gem.num_electron_up, gem.num_electron_dn
    cu = pos[np.arange(n_up) % len(pos)]
    cd = pos[np.arange(n_dn) % len(pos)]
    return cu[None] + rng.normal(0, sigma, size=(nw, n_up, 3)), cd[None] + rng.normal(0, sigma, size=(nw, n_dn, 3))
