"""Atomic forces on the walker engine: the position derivatives the reference obtains by automatic
differentiation of its JAX code, here by central finite differences of the engine's own entries on the device, plus the
space-warp coordinate transformation (SWCT) and the Hellmann-Feynman / Pulay force products of ``MCMC.run``.

Reference (all per walker and measurement step, jqmc/jqmc_mcmc.py:749-852):

    de_L/dr_up, de_L/dr_dn, de_L/dR          jax.grad(compute_local_energy)      (:756-781, 4744-4746)
    dln|Psi|/dr_up, dln|Psi|/dr_dn, dln|Psi|/dR   jax.grad(evaluate_ln_wavefunction)  (:786-797)
    omega[alpha, i], sum_i grad_i omega[alpha, i]   SWCT weights 1/|r_i - R_alpha|^4, normalised over atoms (jqmc/swct.py:63-150)
    force_HF = de_L/dR + omega_up . de_L/dr_up + omega_dn . de_L/dr_dn                                    (:826-831)
    force_PP = dlnPsi/dR + omega_up . dlnPsi/dr_up + omega_dn . dlnPsi/dr_dn + 1/2 sum_i grad_i omega       (:833-840)

and ``get_aF`` (jqmc_mcmc.py:1191-1370): F = -<force_HF> - 2 (<e_L force_PP> - <e_L><force_PP>), jackknifed over bins x walkers.

How the derivatives are formed here.  e_L and ln|Psi| are smooth functions of the electron and nuclear coordinates except for
the nearest-nucleus assignment of the non-local ECP, which automatic differentiation treats as a constant; the engine
therefore freezes that assignment at the base point (``qe_nearest_nuclei`` / ``qe_local_energy_frozen``) and takes

    f'(x) = (f(x + h) - f(x - h)) / 2h,   h = 1e-4 bohr   (truncation ~ h^2 f'''/6 ~ 1e-8, round-off ~ 1e-16 |f| / h ~ 1e-11)

for every electron coordinate (two evaluations of inverse + e_L + ln|Psi| each, all walkers at once) and every nuclear
coordinate (a displaced Hamiltonian -> a second engine per displaced geometry, built once).  Water: 2 (24 + 9) = 66 batched
evaluations per measurement step.  For LRDMC (``lattice=(alat, non_local_move)``) the differentiated energy is the lattice-regularised
V_diag + V_nondiag (qe_lrdmc_velements_frozen), whose fixed-node min / max branches automatic differentiation follows piecewise:
a walker within h of such a kink (probability ~ h) gets a one-sided slope mixed in, which the tests bound.  Scope: the register kernel family (the one that takes the frozen assignment); the mesh
rotation RT of the step is shared by all displaced evaluations, as in the reference (RTs is an argument of its gradient).
"""

from __future__ import annotations

import dataclasses

import numpy as np
import torch

from .engine import WalkerEngine


def swct_omega(positions, r):
    """omega[nw, n_atom, n_el] = kappa / sum_atoms kappa, kappa = 1 / |r_i - R_alpha|^4 (jqmc/swct.py:63-103); torch, any device."""
    d = r[:, None, :, :] - positions[None, :, None, :]  # [nw, atom, el, 3]
    kappa = 1.0 / (d * d).sum(-1) ** 2
    return kappa / kappa.sum(dim=1, keepdim=True)


def swct_domega(positions, r):
    """sum_i grad_{r_i} omega[alpha, i] -> [nw, n_atom, 3] (jqmc/swct.py:134-150), analytic:
    grad kappa = -4 d / |d|^6, grad omega_alpha = grad kappa_alpha / S - kappa_alpha sum_beta grad kappa_beta / S^2."""
    d = r[:, None, :, :] - positions[None, :, None, :]
    d2 = (d * d).sum(-1)
    kappa = 1.0 / d2**2
    gk = -4.0 * d / (d2**3)[..., None]  # [nw, atom, el, 3]
    S = kappa.sum(dim=1, keepdim=True)  # [nw, 1, el]
    gS = gk.sum(dim=1, keepdim=True)  # [nw, 1, el, 3]
    g = gk / S[..., None] - kappa[..., None] * gS / (S**2)[..., None]
    return g.sum(dim=2)


class ForceEvaluator:
    """Position derivatives of e_L and ln|Psi| for a batch of walkers (see module docstring)."""

    def __init__(self, hamiltonian_data, engine: WalkerEngine, h: float = 1.0e-4, lattice=None):
        """``lattice = (alat, non_local_move)`` switches the differentiated energy from the VMC local energy
        (compute_local_energy, jqmc_mcmc.py:756-781) to the lattice-regularised one of LRDMC, V_diag + V_nondiag
        (_compute_local_energy_n / _t, jqmc/jqmc_gfmc.py:5630-5667, 1461-1486)."""
        self.H, self.engine, self.h = hamiltonian_data, engine, float(h)
        self.lattice = None if lattice is None else (float(lattice[0]), lattice[1])
        self.n_atom = len(hamiltonian_data.structure_data.atomic_numbers)
        self._displaced = {}
        self.positions = torch.as_tensor(np.asarray(hamiltonian_data.structure_data.positions, dtype=np.float64), device=engine.device)

    def _engine_at(self, atom: int, axis: int, sign: int) -> WalkerEngine:
        key = (atom, axis, sign)
        if key not in self._displaced:
            self._displaced[key] = self.engine.clone_for(displace_nucleus(self.H, atom, axis, sign * self.h))
        return self._displaced[key]

    def _values(self, eng, r_up, r_dn, RT, nn):
        G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
        if self.lattice is None:
            e = eng.e_L_frozen(r_up, r_dn, RT, Ginv, nn)
        else:
            alat, nlm = self.lattice
            if nn is None:  # all-electron: nothing to freeze
                Vd, Vn = eng.V_elements_n(r_up, r_dn, RT, nlm, alat, A_inv=Ginv)
            else:
                Vd, Vn = eng.V_elements_n(r_up, r_dn, RT, nlm, alat, A_inv=Ginv, nn_index=nn)
            e = Vd + Vn
        ln, _ = eng.ln_wavefunction(r_up, r_dn)
        return e, ln

    def __call__(self, r_up, r_dn, RT):
        """dict(de_L_dR [nw,n_atom,3], de_L_dr_up, de_L_dr_dn, dln_Psi_dR, dln_Psi_dr_up, dln_Psi_dr_dn, e_L, ln_Psi)."""
        eng, h = self.engine, self.h
        r_up, r_dn, nw = eng._walkers(r_up, r_dn)
        nn = eng.nearest_nuclei(r_up, r_dn) if eng.ecp_flag else None
        e0, ln0 = self._values(eng, r_up, r_dn, RT, nn)
        out = dict(e_L=e0, ln_Psi=ln0)
        for name, r, other, is_up in (("up", r_up, r_dn, True), ("dn", r_dn, r_up, False)):
            de = torch.zeros_like(r)
            dl = torch.zeros_like(r)
            for i in range(r.shape[1]):
                for c in range(3):
                    vals = []
                    for sg in (1.0, -1.0):
                        rr = r.clone()
                        rr[:, i, c] += sg * h
                        vals.append(self._values(eng, rr, other, RT, nn) if is_up else self._values(eng, other, rr, RT, nn))
                    de[:, i, c] = (vals[0][0] - vals[1][0]) / (2.0 * h)
                    dl[:, i, c] = (vals[0][1] - vals[1][1]) / (2.0 * h)
            out[f"de_L_dr_{name}"], out[f"dln_Psi_dr_{name}"] = de, dl
        deR = torch.zeros((nw, self.n_atom, 3), dtype=torch.float64, device=eng.device)
        dlR = torch.zeros_like(deR)
        for a in range(self.n_atom):
            for c in range(3):
                ep, lp = self._values(self._engine_at(a, c, +1), r_up, r_dn, RT, nn)
                em, lm = self._values(self._engine_at(a, c, -1), r_up, r_dn, RT, nn)
                deR[:, a, c] = (ep - em) / (2.0 * h)
                dlR[:, a, c] = (lp - lm) / (2.0 * h)
        out["de_L_dR"], out["dln_Psi_dR"] = deR, dlR
        return out

    def force_products(self, r_up, r_dn, RT, use_swct: bool = True):
        """(force_HF, force_PP, e_L) per walker, [nw, n_atom, 3] (jqmc_mcmc.py:799-844)."""
        d = self(r_up, r_dn, RT)
        r_up, r_dn, nw = self.engine._walkers(r_up, r_dn)
        if use_swct:
            om_u, om_d = swct_omega(self.positions, r_up), swct_omega(self.positions, r_dn)
            dom = swct_domega(self.positions, r_up) + swct_domega(self.positions, r_dn)
        else:
            om_u = torch.zeros((nw, self.n_atom, r_up.shape[1]), dtype=torch.float64, device=r_up.device)
            om_d = torch.zeros((nw, self.n_atom, r_dn.shape[1]), dtype=torch.float64, device=r_up.device)
            dom = torch.zeros((nw, self.n_atom, 3), dtype=torch.float64, device=r_up.device)
        f_hf = d["de_L_dR"] + torch.einsum("wjk,wkl->wjl", om_u, d["de_L_dr_up"]) + torch.einsum("wjk,wkl->wjl", om_d, d["de_L_dr_dn"])
        f_pp = (d["dln_Psi_dR"] + torch.einsum("wjk,wkl->wjl", om_u, d["dln_Psi_dr_up"])
                + torch.einsum("wjk,wkl->wjl", om_d, d["dln_Psi_dr_dn"]) + 0.5 * dom)  # fmt: skip
        self._last = d
        return f_hf, f_pp, d["e_L"]

    def weighted_force_sums(self, r_up, r_dn, RT, weight, e_L, use_swct: bool, epsilon_PW: float = 0.0):
        """LRDMC per-branching force sums of this rank -> device tensor [3, n_atom, 3]:
        sum_w g_w force_HF_w, sum_w g_w force_PP_w, sum_w g_w e_L_w force_PP_w with g = weight * f_eps, where f_eps is the
        Pathak-Wagner regularisation 7 t^6 - 15 t^4 + 9 t^2 for t = 1 / (|grad ln Psi| epsilon_PW) < 1 and 1 otherwise
        (jqmc/jqmc_gfmc.py:5977-6019 GFMC_n with weight = w / (V_diag - E_scf); :1926-1970 GFMC_t with weight = w)."""
        f_hf, f_pp, _ = self.force_products(r_up, r_dn, RT, use_swct)
        g = weight
        if epsilon_PW > 0.0:
            d = self._last
            gn2 = (d["dln_Psi_dr_up"] ** 2).sum(dim=(1, 2)) + (d["dln_Psi_dr_dn"] ** 2).sum(dim=(1, 2))
            t = 1.0 / torch.sqrt(gn2) / epsilon_PW
            t2 = t * t
            t4 = t2 * t2
            g = g * torch.where(t < 1.0, 7.0 * t4 * t2 - 15.0 * t4 + 9.0 * t2, torch.ones_like(t))
        return torch.stack([torch.einsum("i,ijk->jk", g, f_hf), torch.einsum("i,ijk->jk", g, f_pp),
                            torch.einsum("i,ijk->jk", g * e_L, f_pp)])  # fmt: skip


def displace_nucleus(H, atom: int, axis: int, delta: float):
    """Copy of the Hamiltonian with nucleus `atom` moved by `delta` along `axis`: every Structure_data in the tree (AO centres of
    the geminal and of the J3 orbitals, J1 centres, ECP / Coulomb centres) moves together, as the reference's gradient with
    respect to the `positions` leaves does (jqmc/hamiltonians.py accumulate_position_grad)."""

    def walk(obj):
        if dataclasses.is_dataclass(obj) and not isinstance(obj, type):
            if type(obj).__name__ == "Structure_data":
                pos = np.array(obj.positions, dtype=np.float64, copy=True)
                pos[atom, axis] += delta
                return dataclasses.replace(obj, positions=pos)
            return dataclasses.replace(obj, **{f.name: walk(getattr(obj, f.name)) for f in dataclasses.fields(obj)})
        return obj

    return walk(H)


def jackknife_forces(w_L, e_L, force_HF, force_PP, num_mcmc_bin_blocks: int, device=None):
    """(force_mean, force_std) [n_atom, 3]: F = -<force_HF> - 2 (<e_L force_PP> - <e_L><force_PP>) with the reference's binned
    jackknife over (bins x walkers) samples of all ranks and its two-pass standard deviation (jqmc_mcmc.py:1226-1370)."""
    from .mcmc import _allreduce_sum

    w_L, e_L = np.asarray(w_L), np.asarray(e_L)
    fh, fp = np.asarray(force_HF), np.asarray(force_PP)
    fe = e_L[..., None, None] * fp

    def binned(x):
        s = np.array([np.sum(a, axis=0) for a in np.array_split(x, num_mcmc_bin_blocks, axis=0)])  # [bins, nw, ...]
        return s.reshape((s.shape[0] * s.shape[1],) + s.shape[2:])

    wb, web = binned(w_L), binned(w_L * e_L)
    whf, wpp, wef = binned(w_L[..., None, None] * fh), binned(w_L[..., None, None] * fp), binned(w_L[..., None, None] * fe)
    shp = whf.shape[1:]
    tot = _allreduce_sum(np.concatenate([[wb.sum(), web.sum(), float(wb.size)], whf.sum(0).ravel(), wpp.sum(0).ravel(), wef.sum(0).ravel()]), device)
    n = int(np.prod(shp))
    W, WE, M = tot[0], tot[1], int(round(tot[2]))
    HF, PP, EF = (tot[3 + k * n : 3 + (k + 1) * n].reshape(shp) for k in range(3))
    den = (W - wb)[:, None, None]
    f_hf = -(HF - whf) / den
    f_pl = -2.0 * ((EF - wef) / den - ((WE - web)[:, None, None] / den) * ((PP - wpp) / den))
    f = f_hf + f_pl
    mean = _allreduce_sum(f.sum(0).ravel(), device).reshape(shp) / M
    var = _allreduce_sum(((f - mean) ** 2).sum(0).ravel(), device).reshape(shp) / M
    return mean, np.sqrt((M - 1) * var)
