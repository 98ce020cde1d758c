"""Build-box tool: extract the TREXIO inputs of the jQMC reference into tests/golden/*.npz.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python -m tools.make_golden
The .npz files hold the raw TREXIO datasets ``read_trexio_file`` consumes (see
jqmc_b200/trexio_lite.py:TREXIO_KEYS); no reference source code is copied.
"""

import os
import sys

import numpy as np

from jqmc_b200.trexio_lite import read_trexio_arrays

REF = "/root/reference/tests/trexio_example_files"
FILES = [
    "water_ccecp_ccpvqz.h5",
    "H2_ae_ccpvdz_cart.h5",
    "H2_ecp_ccpvtz.h5",
    "H2_ecp_ccpvtz_cart.h5",
    "Li_ae_ccpvdz_cart.h5",
    "N_ae_ccpvdz_cart.h5",
    "N2_ecp_ccpvtz_cart.h5",
    "H_ecp_ccpvqz.h5",
    "H2_ae_ccpvqz.h5",
    # multi-channel ccECPs (several non-local channels, heavier atoms, more electrons than the register kernels hold)
    "Cl2_ecp_ccpvtz_cart.h5",
    "CuBr_ecp_ccpvtz_cart.h5",
    "Ti2_ecp_ccpvtz_cart.h5",
]


def main():
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for fn in FILES:
        t = read_trexio_arrays(os.path.join(REF, fn))
        t["nucleus_label"] = np.array([str(x) for x in t["nucleus_label"]])
        # keep only MOs that can be occupied (+ a few virtuals) to keep fixtures small
        occ = np.asarray(t["mo_occupation"])
        keep = np.nonzero(occ > 1e-6)[0]
        n_keep = min(len(occ), (keep.max() + 1 if len(keep) else 0) + 4)
        spin = np.asarray(t["mo_spin"])
        sel = np.array([i for i in range(len(occ)) if (i % max(1, (spin == 0).sum())) < n_keep], dtype=int)
        if np.all(spin == 0):
            sel = np.arange(n_keep)
        t["mo_coefficient"] = np.asarray(t["mo_coefficient"])[sel]
        t["mo_occupation"] = occ[sel]
        t["mo_spin"] = spin[sel]
        path = os.path.join(out_dir, fn.replace(".h5", ".npz"))
        np.savez_compressed(path, **t)
        print(path, os.path.getsize(path), "bytes", file=sys.stderr)


if __name__ == "__main__":
    main()
