"""NumPy restatement of jQMC's per-walker step kernels (TEST INFRASTRUCTURE -- see oracle/__init__.py).

* ``generate_rotation_matrix``   jqmc/jqmc_mcmc.py:4228-4245
* ``geminal_inv``                jqmc/jqmc_mcmc.py:4248-4261
* ``update_electron_positions``  jqmc/jqmc_mcmc.py:4278-4533 (Metropolis, nmpm single-electron proposals)
* ``lrdmc_projection``           jqmc/jqmc_gfmc.py:4738-5358 (GFMC_n projection, legacy kinetic path)
* ``lrdmc_V_elements``           jqmc/jqmc_gfmc.py:5360-5627
* ``lrdmc_projection_t_step/_loop``  jqmc/jqmc_gfmc.py:724-1110, 1539-1570 (GFMC_t continuous-time projection, legacy path)

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


# --------------------------------------------------------------------------------------
# LRDMC (GFMC_n)                      jqmc/jqmc_gfmc.py:4738-5627
# --------------------------------------------------------------------------------------
_SHIFTS = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)


def discretized_kinetic_elements(wf, r_up, r_dn, Ginv, RT, alat):
    """Off-diagonal LRDMC kinetic elements -Psi'/(2 a^2 Psi) on the 6 N_e mesh (x+,x-,y+,y-,z+,z- per
    electron, up block then down block), shifts rotated as ``shifts @ RT``
    (jqmc/wavefunction.py:1739-1860).  Returns (moves, elements); moves[k] = (spin_up, idx, r_new)."""
    shifts = alat * _SHIFTS @ np.asarray(RT, dtype=np.float64)
    moves, elems = [], []
    for spin_up, rs in ((True, r_up), (False, r_dn)):
        for i, r in enumerate(rs):
            for s in shifts:
                r_new = r + s
                ratio = P.det_ratio_fast(wf.geminal_data, r_up, r_dn, Ginv, spin_up, i, r_new)
                ratio *= P.jastrow_ratio(wf.jastrow_data, r_up, r_dn, spin_up, i, r_new)
                moves.append((spin_up, i, r_new))
                elems.append(-1.0 / (2.0 * alat**2) * ratio)
    return moves, np.array(elems)


def lrdmc_elements(H, r_up, r_dn, Ginv, RT, alat, non_local_move="tmove", NN=1, Nv=6):
    """Diagonal / off-diagonal sums of the lattice-regularised Hamiltonian with importance sampling, the
    move list and the (unnormalised, non-positive) move weights -- the body shared by
    ``_body_step_core`` (jqmc/jqmc_gfmc.py:4807-5031) and ``_compute_V_elements_n`` (:5360-5627)."""
    wf, cp = H.wavefunction_data, H.coulomb_potential_data
    r_up = np.asarray(r_up, dtype=np.float64)
    r_dn = np.asarray(r_dn, dtype=np.float64)
    n_up, n_dn = len(r_up), len(r_dn)
    diag_kin = 3.0 / (2.0 * alat**2) * (n_up + n_dn)
    ke_up, ke_dn = P.compute_kinetic_energy_all_elements(wf, r_up, r_dn, Ginv)
    moves, el = discretized_kinetic_elements(wf, r_up, r_dn, Ginv, RT, alat)
    kin_FN = np.minimum(el, 0.0)
    nondiag_kin = np.sum(kin_FN)
    diag_kin_SP = np.sum(np.maximum(el, 0.0))
    el6 = el.reshape(-1, 6)
    flags = np.any(el6 >= 0, axis=1)
    nd_elem = np.sum(el6 + 1.0 / (4.0 * alat**2), axis=1)
    ei_up, ei_dn = P.compute_bare_coulomb_potential_el_ion_element_wise(cp, r_up, r_dn)
    di_up, di_dn = P.compute_bare_coulomb_potential_el_ion_element_wise(cp, r_up, r_dn, alat=alat)
    e_ion = np.concatenate([ei_up, ei_dn])
    e_ion_disc = np.concatenate([di_up, di_dn])
    ke = np.concatenate([ke_up, ke_dn])
    zv = e_ion + ke - nd_elem
    ei = e_ion if cp.ecp_flag else e_ion_disc
    opt = np.where(flags, np.maximum(zv, ei), zv)
    disc_bare = P.compute_bare_coulomb_potential_el_el(r_up, r_dn) + P.compute_bare_coulomb_potential_ion_ion(cp) + np.sum(opt)
    if cp.ecp_flag:
        local = P.compute_ecp_local_parts(cp, r_up, r_dn)
        det_only = non_local_move == "dltmove"
        mu, md, Vnl, _ = P.compute_ecp_non_local_parts_nearest_neighbors(cp, wf, r_up, r_dn, RT, NN, Nv, det_only=det_only, Ginv=Ginv)
        FN = np.minimum(Vnl, 0.0)
        SP = np.sum(np.maximum(Vnl, 0.0))
        npt = len(Vnl)
        per = NN * Nv
        ecp_moves = []
        for k in range(npt):
            e = k // per
            spin_up = e < n_up
            idx = e if spin_up else e - n_up
            r_new = (mu[k][idx] if spin_up else md[k][idx]).copy()
            ecp_moves.append((spin_up, idx, r_new))
        if det_only:
            jr = np.array([P.jastrow_ratio(wf.jastrow_data, r_up, r_dn, su, i, rn) for (su, i, rn) in ecp_moves])
            FN = FN * jr
        nondiag = nondiag_kin + np.sum(FN)
        diag = diag_kin + disc_bare + local + diag_kin_SP + SP
        p = np.concatenate([kin_FN, FN])
        moves = moves + ecp_moves
    else:
        nondiag = nondiag_kin
        diag = diag_kin + disc_bare + diag_kin_SP
        p = kin_FN
    return diag, nondiag, p, moves


def lrdmc_V_elements(H, r_up, r_dn, RT, non_local_move="tmove", alat=0.3):
    """(V_diag, V_nondiag) at a configuration with a fresh SVD inverse (jqmc/jqmc_gfmc.py:5360-5627)."""
    _, Ginv = geminal_inv(H.wavefunction_data.geminal_data, r_up, r_dn)
    diag, nondiag, _, _ = lrdmc_elements(H, r_up, r_dn, Ginv, RT, alat, non_local_move)
    return diag, nondiag


def lrdmc_projection(H, w, r_up, r_dn, Ginv, key, E_scf, nmpm, random_discretized_mesh, non_local_move, alat, trace=None,
                     norm_order="reference"):
    """``nmpm`` GFMC_n projections of one walker (jqmc/jqmc_gfmc.py:4738-5358).
    Returns (w, r_up, r_dn, Ginv, key, RT, V_diag, V_nondiag).

    norm_order: how the normaliser of the move probabilities is summed.  "reference" = ``p_list.sum()`` as the reference writes
    it (jqmc_gfmc.py:5025; NumPy's pairwise order here, XLA leaves the order unspecified); "sequential" = a running sum in vector
    order, which is what the engine does.  The two differ in the last ulp of the normaliser; tests/test_oracle_golden.py checks
    that they select the same moves on the test seeds."""
    r_up = np.array(r_up, dtype=np.float64)
    r_dn = np.array(r_dn, dtype=np.float64)
    Ginv = np.array(Ginv, dtype=np.float64)
    gem = H.wavefunction_data.geminal_data
    rot_keys, move_keys = [], []
    for _ in range(nmpm):  # _split_step_keys (:5275-5283)
        key, rk = R.split(key)
        key, mk = R.split(key)
        rot_keys.append(rk)
        move_keys.append(mk)
    RT = np.eye(3)
    diag = nondiag = 0.0
    for i in range(nmpm):
        if random_discretized_mesh:
            a, b, g = R.uniform(rot_keys[i], 3, -2 * np.pi, 2 * np.pi)
        else:
            a = b = g = 0.0
        RT = rotation_from_angles(a, b, g).T
        diag, nondiag, p, moves = lrdmc_elements(H, r_up, r_dn, Ginv, RT, alat, non_local_move)
        b_x = 1.0 / (diag - E_scf) * (-nondiag)
        w = w * b_x
        if norm_order == "sequential":
            tot = 0.0
            for x in p:
                tot += x
        else:
            tot = np.asarray(p).sum()
        cdf = np.cumsum(p / tot)
        u = R.uniform(move_keys[i])
        k = min(int(np.searchsorted(cdf, u, side="left")), len(cdf) - 1)
        spin_up, idx, r_new = moves[k]
        # Sherman-Morrison with row/column differences evaluated from scratch (:5083-5141)
        p_up, p_dn = r_up.copy(), r_dn.copy()
        (p_up if spin_up else p_dn)[idx] = r_new
        G_old = P.compute_geminal_all_elements(gem, r_up, r_dn)
        G_new = P.compute_geminal_all_elements(gem, p_up, p_dn)
        n = len(r_up)
        if spin_up:
            v = (G_new[idx, :] - G_old[idx, :])[:, None]
            u_ = np.zeros((n, 1))
            u_[idx, 0] = 1.0
        else:
            u_ = (G_new[:, idx] - G_old[:, idx])[:, None]
            v = np.zeros((n, 1))
            v[idx, 0] = 1.0
        Ainv_u = Ginv @ u_
        vT_Ainv = v.T @ Ginv
        det_ratio = 1.0 + (v.T @ Ainv_u)[0, 0]
        Ginv = Ginv - (Ainv_u @ vT_Ainv) / det_ratio
        if trace is not None:
            trace.append(dict(k=k, u=u, b_x=b_x, diag=diag, nondiag=nondiag, spin_up=spin_up, idx=idx))
        r_up, r_dn = p_up, p_dn
    return w, r_up, r_dn, Ginv, key, RT, diag, nondiag


# --------------------------------------------------------------------------------------
# LRDMC (GFMC_t)                      jqmc/jqmc_gfmc.py:724-1110, 1539-1570
# --------------------------------------------------------------------------------------
def _move_with_sherman_morrison(gem, r_up, r_dn, Ginv, spin_up, idx, r_new):
    """Single-electron move and rank-1 update of the inverse, row/column differences from scratch (:1027-1094)."""
    p_up, p_dn = r_up.copy(), r_dn.copy()
    (p_up if spin_up else p_dn)[idx] = r_new
    G_old = P.compute_geminal_all_elements(gem, r_up, r_dn)
    G_new = P.compute_geminal_all_elements(gem, p_up, p_dn)
    n = len(r_up)
    if spin_up:
        v = (G_new[idx, :] - G_old[idx, :])[:, None]
        u_ = np.zeros((n, 1))
        u_[idx, 0] = 1.0
    else:
        u_ = (G_new[:, idx] - G_old[:, idx])[:, None]
        v = np.zeros((n, 1))
        v[idx, 0] = 1.0
    Ainv_u = Ginv @ u_
    vT_Ainv = v.T @ Ginv
    det_ratio = 1.0 + (v.T @ Ainv_u)[0, 0]
    return p_up, p_dn, Ginv - (Ainv_u @ vT_Ainv) / det_ratio


def lrdmc_projection_t_step(H, pc, tau_left, w, r_up, r_dn, Ginv, key, random_discretized_mesh, non_local_move, alat, trace=None):
    """One ``_projection_t_core`` call for one walker (jqmc/jqmc_gfmc.py:724-1110).
    Returns (e_L, pc, tau_left, w, r_up, r_dn, Ginv, key, RT)."""
    gem = H.wavefunction_data.geminal_data
    if tau_left > 0.0:
        pc = pc + 1
    key, sub = R.split(key)
    if random_discretized_mesh:
        a, b, g = R.uniform(sub, 3, -2 * np.pi, 2 * np.pi)
    else:
        a = b = g = 0.0
    RT = rotation_from_angles(a, b, g).T
    diag, nondiag, p, moves = lrdmc_elements(H, r_up, r_dn, Ginv, RT, alat, non_local_move)
    e_L = diag + nondiag
    key, sub = R.split(key)
    xi = R.uniform(sub)
    tau_update = min(tau_left, np.log(1.0 - xi) / nondiag)
    w = w * np.exp(-tau_update * e_L)
    tau_left = tau_left - tau_update
    key, sub = R.split(key)
    u = R.uniform(sub)
    tot = 0.0
    for x in p:
        tot += x
    cdf = np.cumsum(p / tot)
    k = min(int(np.searchsorted(cdf, u, side="left")), len(cdf) - 1)
    moved = not (tau_left <= 0.0)
    if trace is not None:
        trace.append(dict(k=k, u=u, xi=xi, tau_update=tau_update, e_L=e_L, moved=moved))
    if moved:
        spin_up, idx, r_new = moves[k]
        r_up, r_dn, Ginv = _move_with_sherman_morrison(gem, r_up, r_dn, Ginv, spin_up, idx, r_new)
    return e_L, pc, tau_left, w, r_up, r_dn, Ginv, key, RT


def lrdmc_projection_t_loop(H, w, r_up, r_dn, Ginv, keys, tau, random_discretized_mesh, non_local_move, alat, trace=None):
    """``_run_projection_loop`` over a batch of walkers (jqmc/jqmc_gfmc.py:1539-1570): the vmapped body runs once, then
    again for EVERY walker while any walker has time left -- finished walkers still split their keys and re-evaluate
    e_L with a fresh mesh rotation.  Arrays carry a leading walker axis.
    Returns (e_L, pc, w, r_up, r_dn, Ginv, keys, RT, n_iterations)."""
    nw = len(w)
    st = [
        dict(pc=0, tl=float(tau), w=float(w[i]), ru=np.array(r_up[i], dtype=np.float64), rd=np.array(r_dn[i], dtype=np.float64),
             gi=np.array(Ginv[i], dtype=np.float64), key=(int(keys[i][0]), int(keys[i][1])), e=0.0, RT=np.eye(3))
        for i in range(nw)
    ]  # fmt: skip
    n_it = 0
    while True:
        for i, s in enumerate(st):
            tr = None if trace is None else trace[i]
            s["e"], s["pc"], s["tl"], s["w"], s["ru"], s["rd"], s["gi"], s["key"], s["RT"] = lrdmc_projection_t_step(
                H, s["pc"], s["tl"], s["w"], s["ru"], s["rd"], s["gi"], s["key"], random_discretized_mesh, non_local_move, alat, tr
            )
        n_it += 1
        if not max(s["tl"] for s in st) > 0.0:
            break
    return (
        np.array([s["e"] for s in st]), np.array([s["pc"] for s in st], dtype=np.int32), np.array([s["w"] for s in st]),
        np.array([s["ru"] for s in st]), np.array([s["rd"] for s in st]), np.array([s["gi"] for s in st]),
        np.array([s["key"] for s in st], dtype=np.uint32), np.array([s["RT"] for s in st]), n_it,
    )  # fmt: skip


def lrdmc_collect_t(w, e_L):
    """[nw, sum w, sum w e_L, sum w e_L^2] of one rank (GFMC_t, jqmc/jqmc_gfmc.py:1929-1932)."""
    w, e = np.asarray(w, dtype=np.float64), np.asarray(e_L, dtype=np.float64)
    return np.array([len(w), np.sum(w), np.sum(w * e), np.sum(w * e**2)])


# --------------------------------------------------------------------------------------
# GFMC_n per-step statistics and walker reconfiguration    jqmc/jqmc_gfmc.py:5955-6321
# --------------------------------------------------------------------------------------
def lrdmc_collect(w, V_diag, V_nondiag, E_scf):
    """[nw, sum w, sum w/(Vd-E), sum w/(Vd-E) e_L, sum w/(Vd-E) e_L^2] of one rank (:5971-5976)."""
    w, Vd, Vn = (np.asarray(x, dtype=np.float64) for x in (w, V_diag, V_nondiag))
    e = Vd + Vn
    q = w / (Vd - E_scf)
    return np.array([len(w), np.sum(w), np.sum(q), np.sum(q * e), np.sum(q * e**2)])


def lrdmc_branch_indices(w_per_rank, zeta):
    """Comb reconfiguration indices, rank by rank exactly as the reference computes them with NumPy + MPI
    (:6069-6135): returns (chosen_all[int32, world*nw], n_survived)."""
    w_per_rank = [np.asarray(w, dtype=np.float64) for w in w_per_rank]
    world, nw = len(w_per_rank), len(w_per_rank[0])
    global_weight_sum = 0.0
    for w in w_per_rank:  # allreduce(SUM) of the local np.sum
        global_weight_sum = global_weight_sum + np.sum(w)
    cum, sums = [], []
    for w in w_per_rank:
        p = w / global_weight_sum
        cum.append(np.cumsum(p))
        sums.append(np.sum(p))
    offset = 0.0
    for r in range(world):  # Exscan: rank 0 uses 0.0
        cum[r] = cum[r] + (0.0 if r == 0 else offset)
        offset = offset + sums[r] if r > 0 else sums[0]
    global_cumprob = np.concatenate(cum)
    total = world * nw
    chosen = []
    for r in range(world):
        z_local = (np.arange(r * nw, (r + 1) * nw) + zeta) / total
        chosen.append(np.searchsorted(global_cumprob, z_local).astype(np.int32))
    chosen = np.minimum(np.concatenate(chosen), total - 1).astype(np.int32)
    return chosen, len(np.unique(chosen))
