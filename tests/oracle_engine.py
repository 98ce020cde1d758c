"""Test double of ``jqmc_b200.engine.WalkerEngine`` backed by the CPU oracle (TEST INFRASTRUCTURE).

It lets the host-side drivers (``jqmc_b200.mcmc.MCMC``, ``jqmc_b200.gfmc.GFMC_n``) and their collectives run on a
machine without a GPU (``gloo`` backend), so that the N > 1 orchestration is covered by CPU tests.  It lives under
tests/ on purpose: the product never falls back to it.
"""

import numpy as np
import torch

from oracle import drivers as OD
from oracle import physics as OP


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _key(k):
    return (int(k[0]), int(k[1]))


class OracleEngine:
    device = torch.device("cpu")

    def __init__(self, H):
        self.H = H
        gem = H.wavefunction_data.geminal_data
        self.n_up, self.n_dn = gem.num_electron_up, gem.num_electron_dn

    def _t(self, a, dtype=torch.float64):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a))).to(dtype)

    def geminal_inv_batched(self, r_up, r_dn):
        r_up, r_dn = _np(r_up), _np(r_dn)
        out = [OD.geminal_inv(self.H.wavefunction_data.geminal_data, u, d) for u, d in zip(r_up, r_dn)]
        return self._t([o[0] for o in out]), self._t([o[1] for o in out])

    def A_inv_n(self, r_up, r_dn):
        return self.geminal_inv_batched(r_up, r_dn)[1]

    def update(self, r_up, r_dn, keys, nmpm, Dt, eps, Ginv, G, inplace=False):
        r_up, r_dn, keys, Ginv, G = (_np(x) for x in (r_up, r_dn, keys, Ginv, G))
        res = [
            OD.update_electron_positions(self.H, r_up[w], r_dn[w], _key(keys[w]), nmpm, Dt, eps, Ginv[w], G[w])
            for w in range(len(r_up))
        ]
        acc = torch.tensor([r[0] for r in res], dtype=torch.int32)
        rej = torch.tensor([r[1] for r in res], dtype=torch.int32)
        k2 = torch.from_numpy(np.array([r[4] for r in res], dtype=np.uint32))
        return acc, rej, self._t([r[2] for r in res]), self._t([r[3] for r in res]), k2, self._t([r[5] for r in res]), self._t([r[6] for r in res])

    def grad_ln_psi_params_fast(self, r_up, r_dn, Ginv):
        r_up, r_dn, Ginv = _np(r_up), _np(r_dn), _np(Ginv)
        res = [OP.compute_dln_wf_dparams(self.H.wavefunction_data, r_up[w], r_dn[w], Ginv=Ginv[w]) for w in range(len(r_up))]
        return {k: self._t(np.array([r[k] for r in res])) for k in res[0] if res[0][k] is not None}

    def generate_RTs(self, keys):
        return self._t([OD.generate_rotation_matrix(_key(k)) for k in _np(keys)])

    def e_L_fast(self, r_up, r_dn, RTs, Ginv):
        r_up, r_dn, Ginv = _np(r_up), _np(r_dn), _np(Ginv)
        RTs = _np(RTs) if RTs is not None else [np.eye(3)] * len(r_up)
        return self._t([OP.compute_local_energy(self.H, r_up[w], r_dn[w], RTs[w], Ginv=Ginv[w]) for w in range(len(r_up))])

    def as_reg_fast(self, G, Ginv):
        return self._t([OP.compute_AS_regularization_factor(g, gi) for g, gi in zip(_np(G), _np(Ginv))])

    def projection_n(self, w, r_up, r_dn, A_inv, keys, E_scf, nmpm, mesh, nlm, alat, inplace=False):
        w, r_up, r_dn, A_inv, keys = (_np(x) for x in (w, r_up, r_dn, A_inv, keys))
        res = [
            OD.lrdmc_projection(self.H, w[i], r_up[i], r_dn[i], A_inv[i], _key(keys[i]), E_scf, nmpm, mesh, nlm, alat)
            for i in range(len(w))
        ]
        k2 = torch.from_numpy(np.array([r[4] for r in res], dtype=np.uint32))
        cols = lambda j: self._t([r[j] for r in res])  # noqa: E731
        return cols(0), cols(1), cols(2), cols(3), k2, cols(5), cols(6), cols(7)

    def projection_t(self, w, r_up, r_dn, A_inv, keys, tau, mesh, nlm, alat, inplace=False):
        w, r_up, r_dn, A_inv, keys = (_np(x) for x in (w, r_up, r_dn, A_inv, keys))
        e, pc, w2, ru, rd, gi, k2, RT, _ = OD.lrdmc_projection_t_loop(self.H, w, r_up, r_dn, A_inv, keys, tau, mesh, nlm, alat)
        return self._t(e), torch.from_numpy(pc), self._t(w2), self._t(ru), self._t(rd), self._t(gi), torch.from_numpy(k2), self._t(RT)

    def lrdmc_collect_t(self, w, e_L):
        s = OD.lrdmc_collect_t(_np(w), _np(e_L))
        return self._t([s[0], s[1], s[1], s[2], s[3]])

    # ---- what jqmc_b200.forces.ForceEvaluator needs (no nearest-nucleus freezing on the oracle side: ecp_flag False) --------
    ecp_flag = False

    def clone_for(self, H):
        return OracleEngine(H)

    def _walkers(self, r_up, r_dn):
        r_up, r_dn = self._t(_np(r_up)), self._t(_np(r_dn))
        return r_up, r_dn, r_up.shape[0]

    def ln_wavefunction(self, r_up, r_dn):
        r_up, r_dn = _np(r_up), _np(r_dn)
        ln = [OP.evaluate_ln_wavefunction(self.H.wavefunction_data, u, d) for u, d in zip(r_up, r_dn)]
        return self._t(ln), None

    def V_elements_n(self, r_up, r_dn, RTs, nlm, alat, A_inv=None):
        r_up, r_dn, RTs = _np(r_up), _np(r_dn), _np(RTs)
        res = [OD.lrdmc_V_elements(self.H, r_up[i], r_dn[i], RTs[i], nlm, alat) for i in range(len(r_up))]
        return self._t([r[0] for r in res]), self._t([r[1] for r in res])

    def lrdmc_collect(self, w, Vd, Vn, E_scf):
        return self._t(OD.lrdmc_collect(_np(w), _np(Vd), _np(Vn), E_scf))

    def lrdmc_branch(self, w_all, nw, zeta):
        w_all = _np(w_all)
        chosen, ns = OD.lrdmc_branch_indices(np.split(w_all, len(w_all) // nw), zeta)
        return torch.from_numpy(chosen), torch.tensor([ns], dtype=torch.int32)

    def gather_walkers(self, chosen_local, src_up, src_dn):
        idx = torch.as_tensor(_np(chosen_local).astype(np.int64))
        return src_up[idx].clone(), src_dn[idx].clone()

    # ---- packed reconfiguration exchange (mirrors qe_lrdmc_record_len / qe_lrdmc_pack / qe_lrdmc_reconfigure_packed) ----------
    def lrdmc_record_len(self, nw):
        return 8 + nw * (1 + 3 * self.n_up + 3 * self.n_dn)

    def lrdmc_pack(self, sums5, w, r_up, r_dn, record):
        nw = w.shape[0]
        record[:5] = torch.as_tensor(_np(sums5))
        record[5:8] = 0.0
        record[8 : 8 + nw] = torch.as_tensor(_np(w))
        o = 8 + nw
        record[o : o + nw * 3 * self.n_up] = torch.as_tensor(_np(r_up)).reshape(-1)
        o += nw * 3 * self.n_up
        record[o : o + nw * 3 * self.n_dn] = torch.as_tensor(_np(r_dn)).reshape(-1)
        return record

    def lrdmc_reconfigure_packed(self, records, nw, world, rank, zeta):
        L = self.lrdmc_record_len(nw)
        rec = _np(records).reshape(world, L)
        sums = np.zeros(5)
        for r in range(world):  # rank order, like the device kernel and the reference's MPI reduce
            sums = sums + rec[r, :5]
        chosen, ns = OD.lrdmc_branch_indices([rec[r, 8 : 8 + nw] for r in range(world)], zeta)
        pu, pd = 3 * self.n_up, 3 * self.n_dn
        up = rec[:, 8 + nw : 8 + nw + nw * pu].reshape(world * nw, self.n_up, 3)
        dn = rec[:, 8 + nw + nw * pu :].reshape(world * nw, self.n_dn, 3)
        mine = chosen[rank * nw : (rank + 1) * nw]
        return self._t(sums), self._t(up[mine]), self._t(dn[mine]), torch.tensor([ns], dtype=torch.int32), torch.from_numpy(chosen)
