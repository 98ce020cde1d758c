"""GPU parity at the BENCHMARKED launch shape: 4096 walkers of water ccECP/cc-pVQZ JSD + J2 with the default launch
configuration (fused walker kernel: 28 walkers per CTA, 147 CTAs, the last one partly filled; Metropolis kernel: 32 walkers
per CTA, 128 CTAs), compared with the CPU oracle on sampled walkers -- the first and the last walker, both sides of CTA
boundaries of either kernel, the last full CTA and the partly filled one.  The small-batch parity tests
(tests/test_gpu_parity.py) run 1-6 walkers, i.e. one partly filled CTA; this file checks that nothing changes when the grid
is the one bench.py times.  Decisions (accept / reject counts, selected mesh moves through the final positions, PRNG keys)
bit-exact, floating point at the tolerances of the small-batch tests."""

import copy

import numpy as np
import pytest

from jqmc_b200.data import Jastrow_data, Jastrow_two_body_data
from oracle import drivers as OD
from oracle import physics as P
from tests.conftest import load_system, random_walkers

pytestmark = pytest.mark.gpu

NW = 4096
# walker kernel: CTA c owns walkers [28 c, 28 c + 28): 27|28, 55|56 are CTA boundaries, 4059|4060 opens the last full CTA (145),
# 4087|4088 opens the partly filled one (146: 8 walkers).  Metropolis kernel: 32 walkers per CTA (31|32, 4063|4064).
SAMPLE = (0, 1, 27, 28, 31, 32, 55, 56, 2047, 2048, 4059, 4060, 4063, 4064, 4087, 4088, 4094, 4095)


@pytest.fixture(scope="module")
def setup():
    from jqmc_b200 import rng_host
    from jqmc_b200.engine import WalkerEngine

    H = copy.deepcopy(load_system("water_ccecp_ccpvqz"))
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=1.0, jastrow_2b_type="pade"))
    eng = WalkerEngine(H)
    r_up, r_dn = random_walkers(H, NW, 2024, scale=0.7)
    keys = rng_host.split(rng_host.PRNGKey(31337), NW)
    return H, eng, r_up, r_dn, keys


def test_update_full_shape(setup):
    """qe_mcmc_update, nmpm = 40 (the benchmarked call): accept / reject counts and keys bit-exact, state to round-off."""
    H, eng, r_up, r_dn, keys = setup
    nmpm = 40
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    acc, rej, ru, rd, k2, Gi2, G2 = (x.cpu().numpy() for x in eng.update(r_up, r_dn, keys, nmpm, 2.0, 0.0, Ginv, G))
    G, Ginv = G.cpu().numpy(), Ginv.cpu().numpy()
    assert np.all(acc + rej == nmpm)
    for w in SAMPLE:
        a, r_, ru_o, rd_o, key_o, Gi_o, G_o = OD.update_electron_positions(
            H, r_up[w], r_dn[w], (int(keys[w, 0]), int(keys[w, 1])), nmpm, 2.0, 0.0, Ginv[w], G[w]
        )
        assert (a, r_) == (int(acc[w]), int(rej[w])), w
        assert tuple(int(x) for x in k2[w]) == tuple(key_o), w
        np.testing.assert_allclose(ru[w], ru_o, rtol=0, atol=1e-11)
        np.testing.assert_allclose(rd[w], rd_o, rtol=0, atol=1e-11)
        np.testing.assert_allclose(G2[w], G_o, rtol=1e-9, atol=1e-12 * np.abs(G_o).max())
        np.testing.assert_allclose(Gi2[w], Gi_o, rtol=1e-7, atol=1e-9 * np.abs(Gi_o).max())


def test_local_energy_and_V_elements_full_shape(setup):
    """qe_local_energy (fused kernel, mode 2) and qe_lrdmc_velements (mode 1) on all 4096 walkers."""
    H, eng, r_up, r_dn, keys = setup
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    RT = eng.generate_RTs(keys)
    e_L = eng.e_L_fast(r_up, r_dn, RT, Ginv).cpu().numpy()
    Vd, Vn = (x.cpu().numpy() for x in eng.V_elements_n(r_up, r_dn, RT, "tmove", 0.30))
    RT, Ginv = RT.cpu().numpy(), Ginv.cpu().numpy()
    assert np.all(np.isfinite(e_L)) and np.all(np.isfinite(Vd)) and np.all(np.isfinite(Vn))
    for w in SAMPLE:
        RTw = OD.generate_rotation_matrix((int(keys[w, 0]), int(keys[w, 1])))
        np.testing.assert_allclose(RT[w], RTw, rtol=0, atol=1e-14)
        ref = P.compute_local_energy(H, r_up[w], r_dn[w], RTw, Ginv=Ginv[w])
        np.testing.assert_allclose(e_L[w], ref, rtol=1e-10, atol=1e-9)
        d, n = OD.lrdmc_V_elements(H, r_up[w], r_dn[w], RTw, "tmove", 0.30)
        np.testing.assert_allclose(Vd[w], d, rtol=1e-9)
        np.testing.assert_allclose(Vn[w], n, rtol=1e-9)


def test_projection_full_shape(setup):
    """qe_lrdmc_project on all 4096 walkers (the oracle is slow: 4 projections, every second sampled walker)."""
    H, eng, r_up, r_dn, keys = setup
    nmpm, alat, E_scf = 4, 0.30, -17.0
    Ginv = eng.A_inv_n(r_up, r_dn)
    out = eng.projection_n(np.ones(NW), r_up, r_dn, Ginv, keys, E_scf, nmpm, True, "tmove", alat)
    w, ru, rd, Gi, k2, RT, Vd, Vn = (x.cpu().numpy() for x in out)
    Ginv = Ginv.cpu().numpy()
    assert np.all(np.isfinite(w)) and np.all(w > 0)
    for i in SAMPLE[::2] + (4095,):
        ow, oru, ord_, oGi, okey, oRT, od, on = OD.lrdmc_projection(
            H, 1.0, r_up[i], r_dn[i], Ginv[i], (int(keys[i, 0]), int(keys[i, 1])), E_scf, nmpm, True, "tmove", alat
        )
        assert tuple(int(x) for x in k2[i]) == tuple(okey), i
        np.testing.assert_allclose(ru[i], oru, rtol=0, atol=1e-11)  # the same mesh moves were selected
        np.testing.assert_allclose(rd[i], ord_, rtol=0, atol=1e-11)
        np.testing.assert_allclose(w[i], ow, rtol=1e-8)
        np.testing.assert_allclose(RT[i], oRT, rtol=0, atol=1e-14)
        np.testing.assert_allclose(Vd[i], od, rtol=1e-8)
        np.testing.assert_allclose(Vn[i], on, rtol=1e-8)
        np.testing.assert_allclose(Gi[i], oGi, rtol=1e-7, atol=1e-9 * np.abs(oGi).max())
