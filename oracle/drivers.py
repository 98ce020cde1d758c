"""NumPy restatement of jQMC's per-walker step kernels (TEST INFRASTRUCTURE -- see oracle/__init__.py).

* ``generate_rotation_matrix``   jqmc/jqmc_mcmc.py:4228-4245
* ``geminal_inv``                jqmc/jqmc_mcmc.py:4248-4261
* ``update_electron_positions``  jqmc/jqmc_mcmc.py:4278-4533 (Metropolis, nmpm single-electron proposals)
* ``lrdmc_projection``           jqmc/jqmc_gfmc.py:4738-5358 (GFMC_n projection, legacy kinetic path)
* ``lrdmc_V_elements``           jqmc/jqmc_gfmc.py:5360-5627

Unlike the reference these evaluate the determinant/Jastrow ratios from scratch (brute force) so
that they do not share the rank-1 algebra of the CUDA kernels; the running inverse is still carried
with Sherman-Morrison because it is part of the interface (returned to the caller).
"""

from __future__ import annotations

import numpy as np

from . import jaxrng as R
from . import physics as P


def rotation_from_angles(alpha, beta, gamma):
    """R = Rz(gamma) Ry(beta) Rx(alpha) as written out at jqmc/jqmc_mcmc.py:4237-4244."""
    ca, sa = np.cos(alpha), np.sin(alpha)
    cb, sb = np.cos(beta), np.sin(beta)
    cg, sg = np.cos(gamma), np.sin(gamma)
    return np.array(
        [
            [cb * cg, cg * sa * sb - ca * sg, sa * sg + ca * cg * sb],
            [cb * sg, ca * cg + sa * sb * sg, ca * sb * sg - cg * sa],
            [-sb, cb * sa, ca * cb],
        ]
    )


def generate_rotation_matrix(key):
    """RT for the MCMC driver: angles from split(key)[1], key NOT advanced; returns R.T."""
    sub = R.split(key)[1]
    a, b, g = R.uniform(sub, 3, -2 * np.pi, 2 * np.pi)
    return rotation_from_angles(a, b, g).T


def geminal_inv(gem, r_up, r_dn):
    G = P.compute_geminal_all_elements(gem, r_up, r_dn)
    return G, P.geminal_inv_svd(G)


def _charges(H):
    return H.coulomb_potential_data.effective_charges


def _f_l(H, r):
    i = P.find_nearest_nucleus_indices(H.structure_data, r, 1)[0]
    Rc = np.asarray(H.structure_data.positions, dtype=np.float64)[i]
    Z = _charges(H)[i]
    d = np.linalg.norm(r - Rc)
    return 1.0 / Z**2 * (1.0 + Z**2 * d) / (1.0 + d)


def update_electron_positions(H, r_up, r_dn, key, nmpm, Dt, epsilon_AS, Ginv, G, trace=None):
    """One walker, ``nmpm`` Metropolis proposals.  Returns (acc, rej, r_up, r_dn, key, Ginv, G).

    ``trace`` (optional list) receives one dict per proposal with the intermediate ratios.
    """
    r_up = np.array(r_up, dtype=np.float64)
    r_dn = np.array(r_dn, dtype=np.float64)
    G = np.array(G, dtype=np.float64)
    Ginv = np.array(Ginv, dtype=np.float64)
    n_up, n_dn = len(r_up), len(r_dn)
    wf = H.wavefunction_data
    gem = wf.geminal_data
    acc = rej = 0
    for _ in range(nmpm):
        key, sub = R.split(key)
        is_up = R.randint(sub, 0, n_up + n_dn) < n_up
        key, sub = R.split(key)
        up_index = R.randint(sub, 0, n_up)
        key, sub = R.split(key)
        dn_index = R.randint(sub, 0, n_dn)
        idx = up_index if is_up else dn_index
        old = (r_up if is_up else r_dn)[idx].copy()

        f_l = _f_l(H, old)
        sigma = f_l * Dt
        key, sub = R.split(key)
        g = R.normal(sub) * sigma
        key, sub = R.split(key)
        axis = R.randint(sub, 0, 3)
        gv = np.zeros(3)
        gv[axis] = g
        new = old + gv
        f_p = _f_l(H, new)
        dn2 = np.linalg.norm(new - old) ** 2
        T_ratio = (f_l / f_p) * np.exp(-dn2 * (1.0 / (2.0 * f_p**2 * Dt**2) - 1.0 / (2.0 * f_l**2 * Dt**2)))

        p_up, p_dn = r_up.copy(), r_dn.copy()
        (p_up if is_up else p_dn)[idx] = new
        J_ratio = np.exp(P.compute_Jastrow_part(wf.jastrow_data, p_up, p_dn) - P.compute_Jastrow_part(wf.jastrow_data, r_up, r_dn))

        # geminal row / column difference, evaluated from scratch
        G_prop_full = P.compute_geminal_all_elements(gem, p_up, p_dn)
        G_old_full = P.compute_geminal_all_elements(gem, r_up, r_dn)
        if is_up:
            v = (G_prop_full[idx, :] - G_old_full[idx, :])[:, None]
            u = np.zeros((n_up, 1))
            u[idx, 0] = 1.0
        else:
            u = (G_prop_full[:, idx] - G_old_full[:, idx])[:, None]
            v = np.zeros((n_up, 1))
            v[idx, 0] = 1.0
        Ainv_u = Ginv @ u
        vT_Ainv = v.T @ Ginv
        Det_ratio = 1.0 + (v.T @ Ainv_u)[0, 0]
        with np.errstate(all="ignore"):
            Ginv_new = Ginv - (Ainv_u @ vT_Ainv) / Det_ratio
            G_new = G.copy()
            if is_up:
                G_new[idx, :] += v[:, 0]
            else:
                G_new[:, idx] += u[:, 0]
            R_p = P.compute_AS_regularization_factor(G_new, Ginv_new)
            R_o = P.compute_AS_regularization_factor(G, Ginv)
            R_AS_ratio = (max(R_p, epsilon_AS) / R_p) / (max(R_o, epsilon_AS) / R_o) if (R_p != 0 and R_o != 0) else np.nan
            R_ratio = (R_AS_ratio * J_ratio * Det_ratio) ** 2.0
            A = min(1.0, R_ratio * T_ratio) if not np.isnan(R_ratio * T_ratio) else np.nan
        key, sub = R.split(key)
        b = R.uniform(sub)
        accepted = bool(b < A)
        if trace is not None:
            trace.append(
                dict(is_up=is_up, idx=idx, axis=axis, g=g, T_ratio=T_ratio, J_ratio=J_ratio, Det_ratio=Det_ratio,
                     R_AS_ratio=R_AS_ratio, b=b, A=A, accepted=accepted)
            )  # fmt: skip
        if accepted:
            acc += 1
            r_up, r_dn, Ginv, G = p_up, p_dn, Ginv_new, G_new
        else:
            rej += 1
    return acc, rej, r_up, r_dn, key, Ginv, G
