"""Host-side pieces of the atomic-force path (jqmc_b200/forces.py) on CPU: SWCT weights against the oracle's restatement of
jqmc/swct.py, the analytic sum of their gradients against finite differences, the nuclear displacement of a Hamiltonian tree,
and the force jackknife against a literal transcription of the estimator's definition (jqmc_mcmc.py:1191-1370)."""

import numpy as np
import torch

from jqmc_b200.forces import displace_nucleus, jackknife_forces, swct_domega, swct_omega
from oracle import physics as P
from tests.conftest import load_system


def test_swct_matches_oracle():
    H = load_system("water_ccecp_ccpvqz")
    rng = np.random.default_rng(0)
    R = np.asarray(H.structure_data.positions)
    r = R[rng.integers(0, 3, size=(5, 4))] + rng.normal(scale=0.8, size=(5, 4, 3))
    om = swct_omega(torch.from_numpy(R), torch.from_numpy(r)).numpy()
    dom = swct_domega(torch.from_numpy(R), torch.from_numpy(r)).numpy()
    for w in range(5):
        ref = P.swct_omega(H.structure_data, r[w])
        np.testing.assert_allclose(om[w], ref, rtol=1e-12)
        np.testing.assert_allclose(om[w].sum(axis=0), 1.0, rtol=1e-13)  # weights are normalised over the atoms
        np.testing.assert_allclose(dom[w], P.swct_domega(H.structure_data, r[w]), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(dom.sum(axis=1), 0.0, atol=1e-12)  # sum over atoms of omega is 1 -> its gradient vanishes


def test_displace_nucleus_moves_every_structure_copy():
    import copy

    from jqmc_b200.data import Jastrow_data, Jastrow_one_body_data

    H = copy.deepcopy(load_system("H2_ecp_ccpvtz"))
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_one_body_data=Jastrow_one_body_data(structure_data=H.structure_data, core_electrons=(0.0, 0.0)))
    H2 = displace_nucleus(H, 1, 2, 0.25)
    p0 = np.asarray(H.structure_data.positions)
    for st in (H2.structure_data, H2.coulomb_potential_data.structure_data, H2.wavefunction_data.geminal_data.orb_data_up_spin.aos_data.structure_data,
               H2.wavefunction_data.jastrow_data.jastrow_one_body_data.structure_data):  # fmt: skip
        d = np.asarray(st.positions) - p0
        assert d[1, 2] == 0.25 and np.count_nonzero(d) == 1
    assert np.array_equal(np.asarray(H.structure_data.positions), p0)  # the original is untouched


def test_force_jackknife_matches_definition():
    rng = np.random.default_rng(3)
    M, nw, na, nb = 24, 5, 2, 6
    w = rng.uniform(0.5, 1.5, size=(M, nw))
    e = rng.normal(-1.1, 0.2, size=(M, nw))
    fh = rng.normal(size=(M, nw, na, 3))
    fp = rng.normal(size=(M, nw, na, 3))
    mean, std = jackknife_forces(w, e, fh, fp, nb)
    # literal definition: bins x walkers samples, leave-one-out estimates of -<HF> - 2 (<e PP> - <e><PP>)
    def binned(x):
        return np.concatenate([np.sum(a, axis=0) for a in np.array_split(x, nb, axis=0)], axis=0)

    wb, web = binned(w), binned(w * e)
    whf, wpp, wef = binned(w[..., None, None] * fh), binned(w[..., None, None] * fp), binned(w[..., None, None] * e[..., None, None] * fp)
    est = []
    for j in range(len(wb)):
        W = wb.sum() - wb[j]
        est.append(-(whf.sum(0) - whf[j]) / W - 2 * ((wef.sum(0) - wef[j]) / W - (web.sum() - web[j]) / W * (wpp.sum(0) - wpp[j]) / W))
    est = np.array(est)
    np.testing.assert_allclose(mean, est.mean(0), rtol=1e-12)
    np.testing.assert_allclose(std, np.sqrt((len(est) - 1) * est.var(0)), rtol=1e-10)
