"""Build-box tool: recover the Jastrow factors of the reference's TurboRVB comparison tests.

The reference stores them as pickles of its own flax dataclasses (tests/trexio_example_files/jastrow_data_*.pkl), which
cannot be unpickled without jqmc + flax.  The pickles were generated from the TurboRVB wavefunction files next to them
(turborvb_WF_*.txt, "fort.10" format) by tests/trexio_example_files/read_jastrow_factor_from_turbo_wf.py through the
third-party `turbogenius` parser (absent here).  This tool reads the few fort.10 sections that matter with its own parser
and writes tests/golden/turbo_jastrow_<suffix>.npz:

    j2_param, j1_param (nan if none), core_electrons[n_atom], positions[n_atom,3], atomic_numbers[n_atom],
    J3 AO tables (nucleus_index, angular_momentums, magnetic_quantum_numbers, orbital_indices, exponents, coefficients),
    j_matrix[n_ao, n_ao+1]

Conventions taken from the reference script (:95-199): shells are uncontracted normalised Gaussians (TurboRVB orbital types
16 = s, 36 = p, 37 = d; 200 = the constant orbital), m order inside a shell in makefun notation p: (+1, -1, 0),
d: (0, +2, -2, -1, +1); the Jastrow matrix is symmetric; its column against the constant orbital is the one-body vector,
scaled by (N_el - 1).  Whether these conventions are right is decided by the TurboRVB known answers of the reference's tests
(tests/test_oracle_golden.py checks them to the reference's own tolerances).

    python -m tools.turbo_jastrow            # needs /root/reference
"""

import os
import sys

import numpy as np

REF = "/root/reference/tests/trexio_example_files"
SUFFIXES = ["w_2b_3b_w_ecp", "w_2b_1b3b_w_ecp", "w_1b_2b_1b3b_ae"]
MULT_TO_L = {1: 0, 3: 1, 5: 2, 7: 3, 9: 4}


def _sections(path):
    """{header text: list of token lists} for every '#' header of the file."""
    out, cur = {}, None
    for line in open(path):
        if line.lstrip().startswith("#"):
            cur = " ".join(line.replace("#", " ").split())
            out[cur] = []
        elif cur is not None and line.strip():
            out[cur].append(line.split())
    return out


def _find(sec, *words):
    for k, v in sec.items():
        if all(w.lower() in k.lower() for w in words):
            return v
    raise KeyError(words)


def parse(path):
    sec = _sections(path)
    n_up, n_el, n_ion = (int(x) for x in _find(sec, "Nelup")[0])
    jas_type = int(_find(sec, "Jas 2body")[0][0])
    n_jasmat = int(_find(sec, "Det mat", "Jas mat")[0][1])
    ion = np.array([float(x) for row in _find(sec, "Ion coordinates") for x in row]).reshape(n_ion, 5)
    valence, atomic_numbers, positions = ion[:, 0], np.floor(ion[:, 1] + 1e-9), ion[:, 2:5]
    two = _find(sec, "Parameters Jastrow two body")[0]
    n_par = int(two[0])
    pars = [float(x) for x in two[1 : 1 + n_par]]
    if jas_type == -5:
        j2, j1 = pars[0], np.nan
    elif jas_type == -15:
        j2, j1 = pars[0], pars[1]
    else:
        raise NotImplementedError(f"Jastrow type {jas_type}")
    # Jastrow shells: "mult npar type" then "ion par..."
    tok = [t for row in _find(sec, "Parameters atomic Jastrow wf") for t in row]
    nuc, ls, ms, oi, ex, co = [], [], [], [], [], []
    ao, i, const_index = 0, 0, None
    n_orb = 0
    while i < len(tok):
        mult, npar, typ = int(tok[i]), int(tok[i + 1]), int(tok[i + 2])
        atom = int(tok[i + 3]) - 1
        par = [float(x) for x in tok[i + 4 : i + 4 + npar]]
        i += 4 + npar
        if typ == 200:
            const_index = n_orb
            n_orb += 1
            continue
        if typ not in (16, 36, 37) or npar != 1:
            raise NotImplementedError(f"TurboRVB orbital type {typ} with {npar} parameters")
        l = MULT_TO_L[mult]
        m_list = {0: [0], 1: [1, -1, 0], 2: [0, 2, -2, -1, 1]}[l]
        for m in m_list:
            nuc.append(atom), ls.append(l), ms.append(m), oi.append(ao), ex.append(par[0]), co.append(1.0)
            ao += 1
        n_orb += mult
    if const_index is None or const_index != n_orb - 1:
        raise NotImplementedError("the constant Jastrow orbital must be the last one")
    rows = _find(sec, "Nonzero values of jasmat")
    assert len(rows) == n_jasmat, (len(rows), n_jasmat)
    M = np.zeros((ao, ao))
    j1v = np.zeros(ao)
    for r, c, v in rows:
        r, c, v = int(r) - 1, int(c) - 1, float(v.replace("D", "E"))
        if r != const_index and c != const_index:
            M[r, c] = v
            M[c, r] = v
        elif c == const_index and r != const_index:
            j1v[r] = v * (n_el - 1)
    return dict(
        j2_param=j2, j1_param=j1, core_electrons=atomic_numbers - valence, positions=positions, atomic_numbers=atomic_numbers,
        nucleus_index=np.array(nuc), angular_momentums=np.array(ls), magnetic_quantum_numbers=np.array(ms),
        orbital_indices=np.array(oi), exponents=np.array(ex), coefficients=np.array(co), j_matrix=np.column_stack([M, j1v]),
    )  # fmt: skip


def main():
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    for suf in SUFFIXES:
        d = parse(os.path.join(REF, f"turborvb_WF_{suf}.txt"))
        path = os.path.join(out_dir, f"turbo_jastrow_{suf}.npz")
        np.savez_compressed(path, **d)
        print(path, d["j_matrix"].shape, "j2", d["j2_param"], "j1", d["j1_param"], file=sys.stderr)


if __name__ == "__main__":
    main()
