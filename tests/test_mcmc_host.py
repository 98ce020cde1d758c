"""Host-side VMC driver (jqmc_b200.mcmc.MCMC) on CPU with the oracle-backed engine double: stored parameter derivatives,
get_dln_WF layout and the jackknifed generalised forces (reference: jqmc/jqmc_mcmc.py:854-876, 1372-1420, 1516-1692)."""

import numpy as np

from jqmc_b200.data import Jastrow_data, Jastrow_one_body_data, Jastrow_two_body_data
from jqmc_b200.mcmc import MCMC, jackknife_gF
from tests.conftest import load_system
from tests.oracle_engine import OracleEngine


def _system():
    H = load_system("H2_ae_ccpvdz_cart")
    H.wavefunction_data.jastrow_data = Jastrow_data(
        jastrow_one_body_data=Jastrow_one_body_data(jastrow_1b_param=0.9, jastrow_1b_type="exp", structure_data=H.structure_data, core_electrons=(0.0, 0.0)),
        jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.75),
    )
    return H


def test_parameter_derivatives_and_generalised_forces():
    H = _system()
    m = MCMC(H, mcmc_seed=7, num_walkers=3, num_mcmc_per_measurement=2, Dt=2.0, epsilon_AS=0.0, comput_log_WF_param_deriv=True,
             engine=OracleEngine(H))  # fmt: skip
    m.run(8)
    d = m.dln_Psi_dc
    lam_shape = np.shape(H.wavefunction_data.geminal_data.lambda_matrix)
    assert d["j1_param"].shape == (8, 3) and d["j2_param"].shape == (8, 3) and d["lambda_matrix"].shape == (8, 3) + lam_shape
    O = m.get_dln_WF(num_mcmc_warmup_steps=2)
    K = 2 + int(np.prod(lam_shape))
    assert O.shape == (6, 3, K)
    np.testing.assert_array_equal(O[:, :, 0], d["j1_param"][2:])
    np.testing.assert_array_equal(O[:, :, 2:], d["lambda_matrix"][2:].reshape(6, 3, -1))
    f, df = m.get_gF(num_mcmc_warmup_steps=2, num_mcmc_bin_blocks=3)
    assert f.shape == (K,) and df.shape == (K,) and np.all(np.isfinite(f)) and np.all(df >= 0)
    # independent jackknife: leave one (bin, walker) sample out of the weighted averages
    w, e = m.w_L[2:], m.e_L[2:]
    bins = [slice(0, 2), slice(2, 4), slice(4, 6)]
    samples = [(b, wk) for b in bins for wk in range(3)]
    est = []
    for b, wk in samples:
        mask = np.ones_like(w, dtype=bool)
        mask[b, wk] = False
        ww = w * mask
        eo = np.einsum("iw,iwk->k", ww * e, O) / ww.sum()
        est.append(-2.0 * (eo - (ww * e).sum() / ww.sum() * np.einsum("iw,iwk->k", ww, O) / ww.sum()))
    est = np.array(est)
    np.testing.assert_allclose(f, est.mean(axis=0), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(df, np.sqrt((len(est) - 1) * est.var(axis=0)), rtol=1e-9, atol=1e-13)
    # a subset of parameters and of blocks
    f2, _ = m.get_gF(2, 3, chosen_param_index=[0, 2])
    np.testing.assert_allclose(f2, f[[0, 2]], rtol=1e-12)
    f3, _ = m.get_gF(2, 3, blocks=["j2_param"])
    np.testing.assert_allclose(f3, f[[1]], rtol=1e-12)
    g, dg = jackknife_gF(w, e, O, 3)
    np.testing.assert_array_equal(g, f)
