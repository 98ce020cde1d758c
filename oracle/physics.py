"""NumPy restatement of jQMC's per-walker physics (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function cites the reference statement it follows (paths relative to the jQMC repository).
The code favours the reference's slow ``_debug`` formulations (closed forms, explicit loops, brute
force determinant ratios) so that it is independent of the CUDA kernels' algebra (factored radial
sums, generated solid-harmonic polynomials, rank-1 updates).  All arithmetic is float64.
"""

from __future__ import annotations

import functools
import itertools
from math import comb, factorial, pi, sqrt

import numpy as np

from jqmc_b200.data import is_cart, is_mos

EPS_SAFE_DISTANCE = 1.0e-12  # jqmc/_setting.py (EPS_safe_distance), used by the Jastrow grads
RCOND_SVD = 1.0e-20  # jqmc/_setting.py:103-106


# --------------------------------------------------------------------------------------
# Atomic orbitals
# --------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _solid_harmonic_tables(l: int, m: int):
    """Coefficient tables of S_lm (pure functions of l, m; cached -- the evaluation below multiplies in the same order as the
    un-tabulated closed form, so the values are bit-identical): [(c, p, q)] for sum c x^p y^q, [(lam, k, n)] for sum lam r^2k z^n."""
    ma = abs(m)
    cs = [1.0, 0.0, -1.0, 0.0]  # cos(k*pi/2)
    sn = [0.0, 1.0, 0.0, -1.0]  # sin(k*pi/2)
    xy = []
    for p in range(ma + 1):
        trig = cs[(ma - p) % 4] if m >= 0 else sn[(ma - p) % 4]
        if trig != 0.0:
            xy.append((comb(ma, p) * trig, p, ma - p))
    zz = []
    for k in range((l - ma) // 2 + 1):
        lam = (
            (-1.0) ** k
            * 2.0 ** (-l)
            * comb(l, k)
            * comb(2 * l - 2 * k, l)
            * factorial(l - 2 * k)
            / factorial(l - 2 * k - ma)
        )
        zz.append((lam, k, l - 2 * k - ma))
    pref = sqrt((2 - int(ma == 0)) * factorial(l - ma) / factorial(l + ma))
    return tuple(xy), tuple(zz), pref


def solid_harmonic(l: int, m: int, d):
    """Real regular solid harmonic S_lm(d), closed form of jqmc/atomic_orbital.py:3019-3109.

    ``d`` has shape (..., 3) and may be complex (used for complex-step derivatives): the formula is
    a polynomial in (x, y, z) once r^(2k) is written as (x^2+y^2+z^2)^k.
    """
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    r2 = x * x + y * y + z * z
    txy, tzz, pref = _solid_harmonic_tables(int(l), int(m))
    # (x,y) part: A_m = Re (x+iy)^|m|, B_m = Im (x+iy)^|m|   (:3069-3083)
    xy = 0.0
    for c, p, q in txy:
        xy = xy + c * x**p * y**q
    # z part (:3086-3103)
    zz = 0.0
    for lam, k, n in tzz:
        zz = zz + lam * r2**k * z**n
    return pref * zz * xy


def _aos_cache(aos):
    """Per-object memo of the primitive lists and the normalised per-AO terms (they depend only on the basis definition and are
    needed at every one of the thousands of AO evaluations of a trajectory test).  The memo lives on the AOs object and is
    valid only while the defining fields are the very same objects (identity check), so a modified copy never sees stale terms."""
    fields = (aos.orbital_indices, aos.exponents, aos.coefficients, aos.angular_momentums, aos.nucleus_index)
    c = getattr(aos, "_oracle_cache", None)
    if c is None or len(c["fields"]) != len(fields) or any(a is not b for a, b in zip(c["fields"], fields)):
        c = {"fields": fields, "prims": None, "terms": {}}
        try:
            object.__setattr__(aos, "_oracle_cache", c)
        except (AttributeError, TypeError):
            pass  # (an object without a __dict__: no memo, same results)
    return c


def _ao_prim_lists(aos):
    c = _aos_cache(aos)
    if c["prims"] is None:
        oi = np.asarray(aos.orbital_indices)
        c["prims"] = [np.nonzero(oi == a)[0] for a in range(aos.num_ao)]
    return c["prims"]


def _norm_sphe(l, Z):
    """jqmc/atomic_orbital.py:2316-2323 (same as :1957-1968)."""
    return np.sqrt(2.0 ** (2 * l + 3) * factorial(l + 1) * (2.0 * Z) ** (l + 1.5) / (factorial(2 * l + 2) * sqrt(pi)))


def _norm_cart(l, nx, ny, nz, Z):
    """jqmc/atomic_orbital.py:2243-2244 (same as :2013-2024)."""
    return np.sqrt(
        (2.0 * Z / pi) ** 1.5
        * (8.0 * Z) ** l
        * factorial(nx)
        * factorial(ny)
        * factorial(nz)
        / (factorial(2 * nx) * factorial(2 * ny) * factorial(2 * nz))
    )


def _ao_terms(aos, a, prims):
    """Per-AO primitive exponents and fully normalised coefficients, and the angular function (memoised per AOs object)."""
    memo = _aos_cache(aos)["terms"]
    hit = memo.get(a)
    if hit is not None and hit[0] is prims:
        return hit[1]
    out = _ao_terms_uncached(aos, a, prims)
    memo[a] = (prims, out)
    return out


def _ao_terms_uncached(aos, a, prims):
    Z = np.asarray(aos.exponents, dtype=np.float64)[prims]
    c = np.asarray(aos.coefficients, dtype=np.float64)[prims]
    l = int(aos.angular_momentums[a])
    if is_cart(aos):
        nx, ny, nz = (int(aos.polynominal_order_x[a]), int(aos.polynominal_order_y[a]), int(aos.polynominal_order_z[a]))
        w = c * _norm_cart(l, nx, ny, nz, Z)

        def ang(d):
            return d[..., 0] ** nx * d[..., 1] ** ny * d[..., 2] ** nz

    else:
        m = int(aos.magnetic_quantum_numbers[a])
        w = c * _norm_sphe(l, Z) * sqrt((2 * l + 1) / (4 * pi))

        def ang(d):
            return solid_harmonic(l, m, d)

    return Z, w, ang


def _ao_eval_complex(aos, a, prims, d):
    """phi_a at displacement d (possibly complex): [sum_p w_p exp(-Z_p r^2)] * ang(d)."""
    Z, w, ang = _ao_terms(aos, a, prims)
    r2 = d[..., 0] ** 2 + d[..., 1] ** 2 + d[..., 2] ** 2
    rad = np.sum(w[:, None] * np.exp(-Z[:, None] * r2[None, :]), axis=0)
    return rad * ang(d)


def compute_AOs(aos, r_carts):
    """AO values, shape (n_ao, N).  Follows _compute_AOs_sphe_debug / _compute_AOs_cart_debug
    (jqmc/atomic_orbital.py:1931-1981, 1984-2033)."""
    r = np.asarray(r_carts, dtype=np.float64).reshape(-1, 3)
    R = np.asarray(aos.structure_data.positions, dtype=np.float64)
    out = np.zeros((aos.num_ao, r.shape[0]))
    for a, prims in enumerate(_ao_prim_lists(aos)):
        d = r - R[int(aos.nucleus_index[a])]
        out[a] = _ao_eval_complex(aos, a, prims, d)
    return out


def compute_AOs_value_grad_lap(aos, r_carts):
    """value, d/dx, d/dy, d/dz, laplacian of every AO; each (n_ao, N).

    Gradient: complex-step derivative of the closed-form AO (exact to rounding, no algebra shared
    with the CUDA kernels).  Laplacian: the reference's analytic identity
    (jqmc/atomic_orbital.py:3575-3588 spherical, :3484-3505 Cartesian)
        lap(phi) = R*lap(A) + A*(4 r^2 R2 - 6 R1) - 4 R1 (r . grad A),
    with R = sum w e^{-Zr^2}, R1 = sum w Z e^{-Zr^2}, R2 = sum w Z^2 e^{-Zr^2}; grad A by complex step,
    lap(A) = 0 for solid harmonics and the closed monomial form for Cartesian AOs.
    """
    r = np.asarray(r_carts, dtype=np.float64).reshape(-1, 3)
    Rn = np.asarray(aos.structure_data.positions, dtype=np.float64)
    n = r.shape[0]
    val = np.zeros((aos.num_ao, n))
    g = np.zeros((3, aos.num_ao, n))
    lap = np.zeros((aos.num_ao, n))
    h = 1.0e-30
    for a, prims in enumerate(_ao_prim_lists(aos)):
        d = r - Rn[int(aos.nucleus_index[a])]
        Z, w, ang = _ao_terms(aos, a, prims)
        r2 = np.sum(d * d, axis=1)
        e = np.exp(-Z[:, None] * r2[None, :])
        R0 = np.sum(w[:, None] * e, axis=0)
        R1 = np.sum((w * Z)[:, None] * e, axis=0)
        R2 = np.sum((w * Z * Z)[:, None] * e, axis=0)
        A = ang(d)
        val[a] = R0 * A
        gA = np.zeros((3, n))
        for c in range(3):
            dc = d.astype(np.complex128)
            dc[:, c] += 1j * h
            g[c, a] = np.imag(_ao_eval_complex(aos, a, prims, dc)) / h
            gA[c] = np.imag(ang(dc)) / h
        if is_cart(aos):
            nx, ny, nz = (int(aos.polynominal_order_x[a]), int(aos.polynominal_order_y[a]), int(aos.polynominal_order_z[a]))
            x, y, z = d[:, 0], d[:, 1], d[:, 2]
            lapA = np.zeros(n)
            if nx >= 2:
                lapA += nx * (nx - 1) * x ** (nx - 2) * y**ny * z**nz
            if ny >= 2:
                lapA += ny * (ny - 1) * x**nx * y ** (ny - 2) * z**nz
            if nz >= 2:
                lapA += nz * (nz - 1) * x**nx * y**ny * z ** (nz - 2)
        else:
            lapA = 0.0
        rdotg = d[:, 0] * gA[0] + d[:, 1] * gA[1] + d[:, 2] * gA[2]
        lap[a] = R0 * lapA + A * (4.0 * r2 * R2 - 6.0 * R1) - 4.0 * R1 * rdotg
    return val, g[0], g[1], g[2], lap


# --------------------------------------------------------------------------------------
# Orbitals (AO or MO layer)      jqmc/molecular_orbital.py:239-301, 375-415
# --------------------------------------------------------------------------------------
def compute_orb(orb, r_carts):
    if is_mos(orb):
        return np.asarray(orb.mo_coefficients, dtype=np.float64) @ compute_AOs(orb.aos_data, r_carts)
    return compute_AOs(orb, r_carts)


def compute_orb_value_grad_lap(orb, r_carts):
    if is_mos(orb):
        C = np.asarray(orb.mo_coefficients, dtype=np.float64)
        return tuple(C @ q for q in compute_AOs_value_grad_lap(orb.aos_data, r_carts))
    return compute_AOs_value_grad_lap(orb, r_carts)


# --------------------------------------------------------------------------------------
# Geminal / determinant            jqmc/determinant.py
# --------------------------------------------------------------------------------------
def _split_lambda(gem):
    lam = np.asarray(gem.lambda_matrix, dtype=np.float64)
    nd = gem.orb_num_dn
    return lam[:, :nd], lam[:, nd:]


def compute_geminal_all_elements(gem, r_up, r_dn):
    """G = Phi_up^T lam_paired Phi_dn || Phi_up^T lam_unpaired  (jqmc/determinant.py:1379-1426)."""
    lp, lu = _split_lambda(gem)
    ou = compute_orb(gem.orb_data_up_spin, r_up)
    od = compute_orb(gem.orb_data_dn_spin, r_dn) if len(r_dn) else np.zeros((gem.orb_num_dn, 0))
    return np.hstack([ou.T @ lp @ od, ou.T @ lu])


def geminal_inv_svd(G):
    """Thresholded-SVD pseudo-inverse (jqmc/jqmc_mcmc.py:4248-4261)."""
    U, s, Vt = np.linalg.svd(G, full_matrices=False)
    s_inv = np.where(s > RCOND_SVD * s[0], 1.0 / s, 0.0)
    return (Vt.T * s_inv[None, :]) @ U.T


def compute_ln_det(gem, r_up, r_dn):
    """ln|det G| (jqmc/determinant.py:1026-1046)."""
    return np.linalg.slogdet(compute_geminal_all_elements(gem, r_up, r_dn))[1]


def compute_det(gem, r_up, r_dn):
    return np.linalg.det(compute_geminal_all_elements(gem, r_up, r_dn))


def compute_AS_regularization_factor(G, Ginv):
    """R_AS = (min(min_i |G_i.|^2, min_j |G_.j|^2) * |Ginv|_F^2)^(-3/8), 0 if the product is <= 0
    (jqmc/determinant.py:1223-1260)."""
    F = np.sum(Ginv**2)
    S = min(np.min(np.sum(G**2, axis=1)), np.min(np.sum(G**2, axis=0)))
    SF = S * F
    return SF ** (-3.0 / 8.0) if SF > 0.0 else 0.0


def compute_grads_and_laplacian_ln_Det(gem, r_up, r_dn, Ginv=None):
    """Per-electron grad ln|det G| and lap ln|det G| (jqmc/determinant.py:2140-2250).
    ``Ginv=None`` inverts G afresh (the non-"fast" variant, :1808-1963)."""
    lp, lu = _split_lambda(gem)
    vu, gxu, gyu, gzu, lu_ = compute_orb_value_grad_lap(gem.orb_data_up_spin, r_up)
    vd, gxd, gyd, gzd, ld_ = compute_orb_value_grad_lap(gem.orb_data_dn_spin, r_dn)
    if Ginv is None:
        Ginv = np.linalg.inv(np.hstack([vu.T @ lp @ vd, vu.T @ lu]))
    n_up, n_dn = len(r_up), len(r_dn)
    grad_up = np.zeros((n_up, 3))
    grad_dn = np.zeros((n_dn, 3))
    for c, (gu, gd) in enumerate(((gxu, gxd), (gyu, gyd), (gzu, gzd))):
        dG_up = np.hstack([gu.T @ lp @ vd, gu.T @ lu])  # d/dr_i of row i
        dG_dn = vu.T @ lp @ gd  # d/dr_j of column j (paired block only)
        grad_up[:, c] = np.einsum("ij,ji->i", dG_up, Ginv)
        grad_dn[:, c] = np.einsum("ij,ji->j", dG_dn, Ginv[:n_dn, :])
    lG_up = np.hstack([lu_.T @ lp @ vd, lu_.T @ lu])
    lG_dn = vu.T @ lp @ ld_
    lap_up = np.einsum("ij,ji->i", lG_up, Ginv) - np.sum(grad_up**2, axis=1)
    lap_dn = np.einsum("ij,ji->j", lG_dn, Ginv[:n_dn, :]) - np.sum(grad_dn**2, axis=1)
    return grad_up, grad_dn, lap_up, lap_dn


# --------------------------------------------------------------------------------------
# Jastrow                           jqmc/jastrow_factor.py
# --------------------------------------------------------------------------------------
def _j1_params(j1):
    z = np.asarray(j1.structure_data.atomic_numbers, dtype=np.float64) - np.asarray(j1.core_electrons, dtype=np.float64)
    return np.asarray(j1.structure_data.positions, dtype=np.float64), z


def compute_Jastrow_one_body(j1, r_up, r_dn):
    """jqmc/jastrow_factor.py:729-777 (_debug)."""
    R, zeff = _j1_params(j1)
    a = float(j1.jastrow_1b_param)
    J = 0.0
    for r in itertools.chain(r_up, r_dn):
        for Rc, Z in zip(R, zeff):
            c = (2.0 * Z) ** 0.25
            d = np.linalg.norm(r - Rc)
            if j1.jastrow_1b_type == "exp":
                f = 1.0 / (2.0 * a) * (1.0 - np.exp(-a * c * d))
            elif j1.jastrow_1b_type == "pade":
                f = d / (2.0 * (1.0 + a * c * d))
            else:
                raise ValueError(f"Unknown jastrow_1b_type: {j1.jastrow_1b_type}")
            J += -((2.0 * Z) ** 0.75) * f
    return J


def compute_Jastrow_two_body(j2, r_up, r_dn):
    """jqmc/jastrow_factor.py:1254-1313 (_debug)."""
    a = float(j2.jastrow_2b_param)

    def f(ri, rj):
        d = np.linalg.norm(ri - rj)
        if j2.jastrow_2b_type == "pade":
            return d / 2.0 / (1.0 + a * d)
        if j2.jastrow_2b_type == "exp":
            return 1.0 / (2.0 * a) * (1.0 - np.exp(-a * d))
        raise ValueError(f"Unknown jastrow_2b_type: {j2.jastrow_2b_type}")

    J = sum(f(u, d) for u, d in itertools.product(r_up, r_dn))
    J += sum(f(a_, b_) for a_, b_ in itertools.combinations(r_up, 2))
    J += sum(f(a_, b_) for a_, b_ in itertools.combinations(r_dn, 2))
    return J


def compute_Jastrow_three_body(j3, r_up, r_dn):
    """jqmc/jastrow_factor.py:1804-1861 (_debug), written with matrix products."""
    jm = np.asarray(j3.j_matrix, dtype=np.float64)
    j1v, M = jm[:, -1], jm[:, :-1]
    xu = compute_orb(j3.orb_data, r_up) if len(r_up) else np.zeros((jm.shape[0], 0))
    xd = compute_orb(j3.orb_data, r_dn) if len(r_dn) else np.zeros((jm.shape[0], 0))
    J = j1v @ xu.sum(axis=1) + j1v @ xd.sum(axis=1)
    Auu = xu.T @ M @ xu
    Add = xd.T @ M @ xd
    J += np.sum(np.triu(Auu, k=1)) + np.sum(np.triu(Add, k=1))
    J += np.sum(xu.T @ M @ xd)
    return J


def compute_Jastrow_part(jd, r_up, r_dn):
    """J = J1 + J2 + J3 (jqmc/jastrow_factor.py:2127-2178)."""
    r_up = np.asarray(r_up, dtype=np.float64).reshape(-1, 3)
    r_dn = np.asarray(r_dn, dtype=np.float64).reshape(-1, 3)
    J = 0.0
    if jd.jastrow_one_body_data is not None:
        J += compute_Jastrow_one_body(jd.jastrow_one_body_data, r_up, r_dn)
    if jd.jastrow_two_body_data is not None:
        J += compute_Jastrow_two_body(jd.jastrow_two_body_data, r_up, r_dn)
    if jd.jastrow_three_body_data is not None:
        J += compute_Jastrow_three_body(jd.jastrow_three_body_data, r_up, r_dn)
    if getattr(jd, "jastrow_nn_data", None) is not None:
        raise NotImplementedError("NN Jastrow is out of scope")
    return J


def compute_grads_and_laplacian_Jastrow_part(jd, r_up, r_dn):
    """Analytic per-electron grad J and lap J (jqmc/jastrow_factor.py:2982-3101;
    J1 :960-1034, J2 :3434-3558, J3 :4076-4156)."""
    r_up = np.asarray(r_up, dtype=np.float64).reshape(-1, 3)
    r_dn = np.asarray(r_dn, dtype=np.float64).reshape(-1, 3)
    n_up, n_dn = len(r_up), len(r_dn)
    r_all = np.vstack([r_up, r_dn])
    grad = np.zeros((n_up + n_dn, 3))
    lap = np.zeros(n_up + n_dn)
    eps = EPS_SAFE_DISTANCE

    j1 = jd.jastrow_one_body_data
    if j1 is not None:
        R, zeff = _j1_params(j1)
        a = float(j1.jastrow_1b_param)
        for i, r in enumerate(r_all):
            for Rc, Z in zip(R, zeff):
                c = (2.0 * Z) ** 0.25
                A = (2.0 * Z) ** 0.75
                diff = r - Rc
                d = max(np.linalg.norm(diff), eps)
                if j1.jastrow_1b_type == "exp":
                    e = np.exp(-a * c * d)
                    fp = -A * (c / 2.0) * e
                    lap[i] += A * (a * c * c / 2.0) * e - A * c * e / d
                else:
                    den = 1.0 + a * c * d
                    fp = -A / (2.0 * den * den)
                    lap[i] += A * a * c / den**3 + 2.0 * fp / d
                grad[i] += fp * diff / d

    j2 = jd.jastrow_two_body_data
    if j2 is not None:
        a = float(j2.jastrow_2b_param)
        for i in range(n_up + n_dn):
            for j in range(n_up + n_dn):
                if i == j:
                    continue
                diff = r_all[i] - r_all[j]
                d = max(np.linalg.norm(diff), eps)
                if j2.jastrow_2b_type == "pade":
                    den = 1.0 + a * d
                    fp = 0.5 / (den * den)
                    lap[i] += -a / den**3 + 2.0 * fp / d
                else:
                    e = np.exp(-a * d)
                    fp = 0.5 * e
                    lap[i] += -(a / 2.0) * e + 2.0 * fp / d
                grad[i] += fp / d * diff

    j3 = jd.jastrow_three_body_data
    if j3 is not None:
        jm = np.asarray(j3.j_matrix, dtype=np.float64)
        j1v, M = jm[:, -1], jm[:, :-1]
        v, gx, gy, gz, lp = compute_orb_value_grad_lap(j3.orb_data, r_all)
        for k in range(n_up + n_dn):
            same = range(0, n_up) if k < n_up else range(n_up, n_up + n_dn)
            other = range(n_up, n_up + n_dn) if k < n_up else range(0, n_up)
            gk = j1v.copy()
            for i in same:
                if i > k:
                    gk += M @ v[:, i]
                elif i < k:
                    gk += M.T @ v[:, i]
            for j in other:
                gk += (M @ v[:, j]) if k < n_up else (M.T @ v[:, j])
            grad[k] += np.array([gk @ gx[:, k], gk @ gy[:, k], gk @ gz[:, k]])
            lap[k] += gk @ lp[:, k]
    return grad[:n_up], grad[n_up:], lap[:n_up], lap[n_up:]


# --------------------------------------------------------------------------------------
# Wavefunction / kinetic energy       jqmc/wavefunction.py
# --------------------------------------------------------------------------------------
def evaluate_ln_wavefunction(wf, r_up, r_dn):
    """ln|Psi| = J + ln|det G| (jqmc/wavefunction.py:677-720)."""
    r_up = np.asarray(r_up, dtype=np.float64).reshape(-1, 3)
    r_dn = np.asarray(r_dn, dtype=np.float64).reshape(-1, 3)
    return compute_Jastrow_part(wf.jastrow_data, r_up, r_dn) + compute_ln_det(wf.geminal_data, r_up, r_dn)


def evaluate_wavefunction(wf, r_up, r_dn):
    """Psi = exp(J) det G."""
    r_up = np.asarray(r_up, dtype=np.float64).reshape(-1, 3)
    r_dn = np.asarray(r_dn, dtype=np.float64).reshape(-1, 3)
    return np.exp(compute_Jastrow_part(wf.jastrow_data, r_up, r_dn)) * compute_det(wf.geminal_data, r_up, r_dn)


def compute_kinetic_energy_all_elements(wf, r_up, r_dn, Ginv=None):
    """T_i = -1/2 (lap ln Psi_i + |grad ln Psi_i|^2)  (jqmc/wavefunction.py:1141-1207, 1270-1297)."""
    r_up = np.asarray(r_up, dtype=np.float64).reshape(-1, 3)
    r_dn = np.asarray(r_dn, dtype=np.float64).reshape(-1, 3)
    gJu, gJd, lJu, lJd = compute_grads_and_laplacian_Jastrow_part(wf.jastrow_data, r_up, r_dn)
    gDu, gDd, lDu, lDd = compute_grads_and_laplacian_ln_Det(wf.geminal_data, r_up, r_dn, Ginv)
    Tu = -0.5 * (lJu + lDu + np.sum((gJu + gDu) ** 2, axis=1))
    Td = -0.5 * (lJd + lDd + np.sum((gJd + gDd) ** 2, axis=1))
    return Tu, Td


def compute_kinetic_energy(wf, r_up, r_dn, Ginv=None):
    Tu, Td = compute_kinetic_energy_all_elements(wf, r_up, r_dn, Ginv)
    return np.sum(Tu) + np.sum(Td)


# --------------------------------------------------------------------------------------
# Derivatives of ln|Psi| with respect to the variational parameters (stochastic reconfiguration O_k)
# --------------------------------------------------------------------------------------
def compute_dln_wf_dparams(wf, r_up, r_dn, Ginv=None):
    """d ln|Psi| / d{jastrow_1b_param, jastrow_2b_param, j_matrix, lambda_matrix} -- what the reference obtains as
    jax.grad(evaluate_ln_wavefunction_fast) (jqmc/jqmc_mcmc.py:4748, 854-876; parameter blocks jqmc/wavefunction.py:515-674),
    written out analytically:  d ln|det G| / d lambda[a,b] = sum_ij Ginv[j,i] dG[i,j]/d lambda[a,b]  with the running
    inverse held fixed, and the explicit parameter derivatives of J1, J2, J3.  Returns a dict (None for absent blocks)."""
    r_up = np.asarray(r_up, dtype=np.float64).reshape(-1, 3)
    r_dn = np.asarray(r_dn, dtype=np.float64).reshape(-1, 3)
    gem, jd = wf.geminal_data, wf.jastrow_data
    n_up, n_dn = len(r_up), len(r_dn)
    out = dict(j1_param=None, j2_param=None, j3_matrix=None, lambda_matrix=None)  # the reference's block names (wavefunction.py:542-624)
    ou = compute_orb(gem.orb_data_up_spin, r_up)
    od = compute_orb(gem.orb_data_dn_spin, r_dn) if n_dn else np.zeros((gem.orb_num_dn, 0))
    if Ginv is None:
        Ginv = np.linalg.inv(compute_geminal_all_elements(gem, r_up, r_dn))
    A = ou @ Ginv.T  # [a, j] = sum_i phi_a(r_i) Ginv[j, i]
    out["lambda_matrix"] = np.hstack([A[:, :n_dn] @ od.T, A[:, n_dn:]])
    j1 = jd.jastrow_one_body_data
    if j1 is not None:
        R, zeff = _j1_params(j1)
        a = float(j1.jastrow_1b_param)
        g = 0.0
        for r in itertools.chain(r_up, r_dn):
            for Rc, Z in zip(R, zeff):
                c, Apre = (2.0 * Z) ** 0.25, (2.0 * Z) ** 0.75
                d = np.linalg.norm(r - Rc)
                if j1.jastrow_1b_type == "exp":
                    e = np.exp(-a * c * d)
                    g += -Apre * (a * c * d * e - (1.0 - e)) / (2.0 * a * a)
                else:
                    g += Apre * c * d * d / (2.0 * (1.0 + a * c * d) ** 2)
        out["j1_param"] = g
    j2 = jd.jastrow_two_body_data
    if j2 is not None:
        a = float(j2.jastrow_2b_param)
        r_all = np.vstack([r_up, r_dn])
        g = 0.0
        for i, j in itertools.combinations(range(len(r_all)), 2):
            d = np.linalg.norm(r_all[i] - r_all[j])
            if j2.jastrow_2b_type == "pade":
                g += -d * d / (2.0 * (1.0 + a * d) ** 2)
            else:
                e = np.exp(-a * d)
                g += (a * d * e - (1.0 - e)) / (2.0 * a * a)
        out["j2_param"] = g
    j3 = jd.jastrow_three_body_data
    if j3 is not None:
        x = compute_orb(j3.orb_data, np.vstack([r_up, r_dn]))  # electrons ordered up then down
        n = x.shape[1]
        dM = np.zeros((x.shape[0], x.shape[0]))
        for i in range(n):
            for j in range(i + 1, n):
                dM += np.outer(x[:, i], x[:, j])
        out["j3_matrix"] = np.hstack([dM, x.sum(axis=1)[:, None]])
    return out


# --------------------------------------------------------------------------------------
# Wavefunction ratios for single-electron moves
# --------------------------------------------------------------------------------------
def wf_ratio_brute_force(wf, r_up, r_dn, spin_up: bool, idx: int, r_new, det_only=False):
    """Psi(r')/Psi(r) by evaluating everything twice (the _debug way,
    jqmc/coulomb_potential.py:956-973)."""
    r_up = np.asarray(r_up, dtype=np.float64)
    r_dn = np.asarray(r_dn, dtype=np.float64)
    nu, nd = r_up.copy(), r_dn.copy()
    (nu if spin_up else nd)[idx] = r_new
    ratio = compute_det(wf.geminal_data, nu, nd) / compute_det(wf.geminal_data, r_up, r_dn)
    if not det_only:
        ratio *= np.exp(compute_Jastrow_part(wf.jastrow_data, nu, nd) - compute_Jastrow_part(wf.jastrow_data, r_up, r_dn))
    return ratio


def det_ratio_fast(gem, r_up, r_dn, Ginv, spin_up: bool, idx: int, r_new):
    """det G'/det G via the matrix-determinant lemma with the running inverse
    (jqmc/determinant.py:1665-1783): row_new . Ginv[:, k]  or  Ginv[k, :] . col_new."""
    lp, lu = _split_lambda(gem)
    r_new = np.asarray(r_new, dtype=np.float64).reshape(1, 3)
    if spin_up:
        ou = compute_orb(gem.orb_data_up_spin, r_new)[:, 0]
        od = compute_orb(gem.orb_data_dn_spin, r_dn) if len(r_dn) else np.zeros((gem.orb_num_dn, 0))
        row = np.concatenate([ou @ lp @ od, ou @ lu])
        return row @ Ginv[:, idx]
    od = compute_orb(gem.orb_data_dn_spin, r_new)[:, 0]
    ou = compute_orb(gem.orb_data_up_spin, r_up)
    col = ou.T @ (lp @ od)
    return Ginv[idx, :] @ col


def jastrow_ratio(jd, r_up, r_dn, spin_up: bool, idx: int, r_new):
    """exp(J(r') - J(r)) (jqmc/jastrow_factor.py:2230-2693 computes the same quantity incrementally)."""
    nu, nd = np.array(r_up, dtype=np.float64), np.array(r_dn, dtype=np.float64)
    (nu if spin_up else nd)[idx] = r_new
    return np.exp(compute_Jastrow_part(jd, nu, nd) - compute_Jastrow_part(jd, r_up, r_dn))


# --------------------------------------------------------------------------------------
# Coulomb / ECP                       jqmc/coulomb_potential.py
# --------------------------------------------------------------------------------------
def compute_bare_coulomb_potential(cp, r_up, r_dn):
    """Sum_{a<b} q_a q_b / r_ab over nuclei+electrons (jqmc/coulomb_potential.py:2224-2249)."""
    R = np.asarray(cp.structure_data.positions, dtype=np.float64)
    q = np.concatenate([cp.effective_charges, -np.ones(len(r_up) + len(r_dn))])
    x = np.vstack([R, np.reshape(r_up, (-1, 3)), np.reshape(r_dn, (-1, 3))])
    V = 0.0
    for a, b in itertools.combinations(range(len(q)), 2):
        V += q[a] * q[b] / np.linalg.norm(x[a] - x[b])
    return V


def compute_bare_coulomb_potential_el_el(r_up, r_dn):
    x = np.vstack([np.reshape(r_up, (-1, 3)), np.reshape(r_dn, (-1, 3))])
    return sum(1.0 / np.linalg.norm(x[a] - x[b]) for a, b in itertools.combinations(range(len(x)), 2))


def compute_bare_coulomb_potential_ion_ion(cp):
    R = np.asarray(cp.structure_data.positions, dtype=np.float64)
    q = cp.effective_charges
    return sum(q[a] * q[b] / np.linalg.norm(R[a] - R[b]) for a, b in itertools.combinations(range(len(q)), 2))


def compute_bare_coulomb_potential_el_ion_element_wise(cp, r_up, r_dn, alat=None):
    """Per-electron -sum_A Z_A / |r_i - R_A|; with ``alat`` the distance is clamped to >= alat
    (jqmc/coulomb_potential.py:2284-2328, discretised variant :2331-2378)."""
    R = np.asarray(cp.structure_data.positions, dtype=np.float64)
    q = cp.effective_charges

    def one(r):
        d = np.linalg.norm(R - r, axis=1)
        if alat is not None:
            d = np.maximum(d, alat)
        return -np.sum(q / d)

    return np.array([one(r) for r in np.reshape(r_up, (-1, 3))]), np.array([one(r) for r in np.reshape(r_dn, (-1, 3))])


def _ecp_terms(cp, i_atom, local: bool):
    out = []
    lmax = int(cp.max_ang_mom_plus_1[i_atom])
    for n, l, z, c, p in zip(cp.nucleus_index, cp.ang_moms, cp.exponents, cp.coefficients, cp.powers):
        if int(n) == i_atom and ((int(l) == lmax) == local):
            out.append((int(l), float(z), float(c), float(p)))
    return out


def compute_ecp_local_parts(cp, r_up, r_dn):
    """jqmc/coulomb_potential.py:601-661."""
    R = np.asarray(cp.structure_data.positions, dtype=np.float64)
    V = 0.0
    for i_atom in range(len(R)):
        terms = _ecp_terms(cp, i_atom, local=True)
        for r in itertools.chain(np.reshape(r_up, (-1, 3)), np.reshape(r_dn, (-1, 3))):
            d = np.linalg.norm(R[i_atom] - r)
            V += d**-2.0 * sum(c * d**p * np.exp(-z * d**2) for (_, z, c, p) in terms)
    return V


def quadrature(Nv: int):
    """Spherical quadrature points/weights (jqmc/coulomb_potential.py:102-184)."""
    if Nv == 4:
        q = 1 / np.sqrt(3)
        return np.full(4, 0.25), np.array([[q, q, q], [q, -q, -q], [-q, q, -q], [-q, -q, q]])
    if Nv == 6:
        g = np.array([[1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0], [0, -1.0, 0], [0, 0, 1.0], [0, 0, -1.0]])
        return np.full(6, 1.0 / 6.0), g
    if Nv == 12:
        t = np.arctan(2)
        sph = [[0.0, 0.0], [np.pi, 0.0]]
        sph += [[t, 2.0 * k * np.pi / 5.0] for k in range(5)]
        sph += [[np.pi - t, (2.0 * k + 1.0) * np.pi / 5.0] for k in range(5)]
        sph = np.array(sph)
        th, ph = sph[:, 0], sph[:, 1]
        return np.full(12, 1.0 / 12.0), np.vstack((np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th))).T
    if Nv == 18:
        p = 1.0 / np.sqrt(2)
        g = [[1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0], [0, -1.0, 0], [0, 0, 1.0], [0, 0, -1.0]]
        g += [[p, p, 0], [p, -p, 0], [-p, p, 0], [-p, -p, 0], [p, 0, p], [p, 0, -p], [-p, 0, p], [-p, 0, -p]]
        g += [[0, p, p], [0, -p, p], [0, p, -p], [0, -p, -p]]
        return np.array([1.0 / 6.0] * 6 + [1.0 / 15.0] * 12), np.array(g, dtype=np.float64)
    raise NotImplementedError


def legendre(l: int, x):
    """P_0..P_6 (jqmc/_function_collections.py:47-65)."""
    from numpy.polynomial import legendre as L

    c = np.zeros(l + 1)
    c[l] = 1.0
    return L.legval(x, c)


def find_nearest_nucleus_indices(structure, r, N):
    """argsort of the distances, first index wins ties (jqmc/structure.py:410-426)."""
    R = np.asarray(structure.positions, dtype=np.float64)
    d = np.sqrt(np.sum((R - np.asarray(r, dtype=np.float64)) ** 2, axis=1))
    return np.argsort(d, kind="stable")[:N]


def compute_ecp_non_local_parts_nearest_neighbors(
    cp, wf, r_up, r_dn, RT=None, NN=1, Nv=6, det_only=False, Ginv=None
):
    """Non-local ECP on the rotated quadrature around the NN nearest nuclei.

    Returns (mesh_r_up, mesh_r_dn, V_nonlocal[points], sum) with the per-point value already summed over
    the angular-momentum channels and ordered (electron, nn, k), up block then down block -- the layout
    of the fast path (jqmc/coulomb_potential.py:1477-1712).  ``Ginv=None`` uses the brute-force
    Psi'/Psi of the _debug twin (:849-1095); otherwise the rank-1 ratio with the given inverse.
    """
    r_up = np.asarray(r_up, dtype=np.float64).reshape(-1, 3)
    r_dn = np.asarray(r_dn, dtype=np.float64).reshape(-1, 3)
    RT = np.eye(3) if RT is None else np.asarray(RT, dtype=np.float64)
    weights, grid = quadrature(Nv)
    grid = grid @ RT
    Rn = np.asarray(cp.structure_data.positions, dtype=np.float64)
    mesh_up, mesh_dn, vals = [], [], []
    for spin_up, rs in ((True, r_up), (False, r_dn)):
        for i, r in enumerate(rs):
            for i_atom in find_nearest_nucleus_indices(cp.structure_data, r, NN):
                rel = Rn[i_atom] - r
                d = np.linalg.norm(rel)
                terms = _ecp_terms(cp, int(i_atom), local=False)
                lmax = int(cp.max_ang_mom_plus_1[i_atom])
                V_l = np.zeros(max(lmax, 1))
                for l, z, c, p in terms:
                    V_l[l] += d**-2.0 * c * d**p * np.exp(-z * d**2)
                for w, g in zip(weights, grid):
                    r_new = r + rel + d * g
                    cos_t = np.dot(-rel / d, g / np.linalg.norm(g))
                    if Ginv is None:
                        ratio = wf_ratio_brute_force(wf, r_up, r_dn, spin_up, i, r_new, det_only)
                    else:
                        ratio = det_ratio_fast(wf.geminal_data, r_up, r_dn, Ginv, spin_up, i, r_new)
                        if not det_only:
                            ratio *= jastrow_ratio(wf.jastrow_data, r_up, r_dn, spin_up, i, r_new)
                    v = sum(V_l[l] * (2 * l + 1) * legendre(l, cos_t) * w * ratio for l in range(lmax))
                    nu, nd = r_up.copy(), r_dn.copy()
                    (nu if spin_up else nd)[i] = r_new
                    mesh_up.append(nu)
                    mesh_dn.append(nd)
                    vals.append(v)
    vals = np.array(vals, dtype=np.float64)
    return np.array(mesh_up), np.array(mesh_dn), vals, float(np.sum(vals))


def compute_coulomb_potential(cp, wf, r_up, r_dn, RT=None, NN=1, Nv=6, Ginv=None):
    """bare + ECP local + ECP non-local (jqmc/coulomb_potential.py:2695-2761)."""
    V = compute_bare_coulomb_potential(cp, r_up, r_dn)
    if cp.ecp_flag:
        V += compute_ecp_local_parts(cp, r_up, r_dn)
        V += compute_ecp_non_local_parts_nearest_neighbors(cp, wf, r_up, r_dn, RT, NN, Nv, Ginv=Ginv)[3]
    return V


# --------------------------------------------------------------------------------------
# Local energy                        jqmc/hamiltonians.py:179-290
# --------------------------------------------------------------------------------------
def compute_local_energy(H, r_up, r_dn, RT=None, Ginv=None, NN=1, Nv=6):
    """e_L = sum_i T_i + V.  ``Ginv=None``: compute_local_energy (fresh inverse, brute-force ratios);
    otherwise compute_local_energy_fast with the supplied running inverse."""
    wf = H.wavefunction_data
    T = compute_kinetic_energy(wf, r_up, r_dn, Ginv)
    V = compute_coulomb_potential(H.coulomb_potential_data, wf, r_up, r_dn, RT, NN, Nv, Ginv)
    return T + V


# ---- space-warp coordinate transformation (jqmc/swct.py:63-150) -----------------------------------------------------------
def swct_omega(structure, r_carts):
    """omega[alpha, i] = kappa_{alpha i} / sum_beta kappa_{beta i}, kappa = 1 / |r_i - R_alpha|^4 (jqmc/swct.py:99-132, debug form)."""
    R = np.asarray(structure.positions, dtype=np.float64)
    r = np.asarray(r_carts, dtype=np.float64)
    om = np.zeros((len(R), len(r)))
    for a in range(len(R)):
        for i in range(len(r)):
            ks = [1.0 / np.linalg.norm(r[i] - R[b]) ** 4 for b in range(len(R))]
            om[a, i] = ks[a] / np.sum(ks)
    return om


def swct_domega(structure, r_carts, h=1.0e-5):
    """sum_i grad_{r_i} omega[alpha, i] -> (n_atom, 3), by central differences of swct_omega (the reference: jacrev, :134-150)."""
    r = np.asarray(r_carts, dtype=np.float64)
    out = np.zeros((len(structure.positions), 3))
    for i in range(len(r)):
        for c in range(3):
            rp, rm = r.copy(), r.copy()
            rp[i, c] += h
            rm[i, c] -= h
            out[:, c] += (swct_omega(structure, rp)[:, i] - swct_omega(structure, rm)[:, i]) / (2 * h)
    return out
