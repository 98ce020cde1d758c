"""Mixed-precision mode (jqmc/_precision.py:345-374; north_star: "1e-5 in its mixed-precision mode").

`WalkerEngine(H, precision="mixed")` evaluates the reference's low-risk zones in fp32 -- AO values (`ao_eval`) and Jastrow
values / ratios (`jastrow_eval`, `jastrow_ratio`), with r - R formed in fp64 first -- and everything else in fp64.  The checks
compare with the fp64 ORACLE (not with an fp32 restatement): what matters is the distance of the mixed results from the
exact ones, at the reference's fp32 tolerance (`_precision.py` "strict"/float32: atol 1e-5, rtol 1e-3; we hold rtol 1e-5 on
the headline quantities).  Quantities owned by fp64 zones (kinetic energy from cached gradients, potentials, the inverse) must
still agree to fp64 round-off.  Decisions are NOT required to be bit-identical in this mode."""

import copy

import numpy as np
import pytest

from jqmc_b200.data import Jastrow_data, Jastrow_one_body_data, Jastrow_two_body_data
from oracle import drivers as OD
from oracle import physics as P
from tests.conftest import load_system, random_walkers

pytestmark = pytest.mark.gpu


def _system(name="water_ccecp_ccpvqz", j1=False):
    H = copy.deepcopy(load_system(name))
    cp = H.coulomb_potential_data
    j1d = None
    if j1:
        j1d = Jastrow_one_body_data(jastrow_1b_param=0.9, jastrow_1b_type="exp", structure_data=H.structure_data, core_electrons=tuple(cp.z_cores))
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_one_body_data=j1d, jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=0.8, jastrow_2b_type="pade"))
    return H


@pytest.mark.parametrize("j1", [False, True])
def test_mixed_local_energy_and_V_elements(j1):
    from jqmc_b200.engine import WalkerEngine

    H = _system(j1=j1)
    eng = WalkerEngine(H, precision="mixed")
    full = WalkerEngine(H)
    nw = 6
    r_up, r_dn = random_walkers(H, nw, 3)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    RT = eng.generate_RTs(np.array([[2, i] for i in range(nw)], dtype=np.uint32))
    e_m, T_m, V_m = (x.cpu().numpy() for x in eng.e_L_fast(r_up, r_dn, RT, Ginv, return_parts=True))
    e_f, T_f, V_f = (x.cpu().numpy() for x in full.e_L_fast(r_up, r_dn, RT, Ginv, return_parts=True))
    Vd_m, Vn_m = (x.cpu().numpy() for x in eng.V_elements_n(r_up, r_dn, RT, "tmove", 0.3))
    RTh, Gih = RT.cpu().numpy(), Ginv.cpu().numpy()
    differs = False
    for w in range(nw):
        ref = P.compute_local_energy(H, r_up[w], r_dn[w], RTh[w], Ginv=Gih[w])
        np.testing.assert_allclose(e_m[w], ref, rtol=1e-5, atol=1e-5)
        d, n = OD.lrdmc_V_elements(H, r_up[w], r_dn[w], RTh[w], "tmove", 0.3)
        np.testing.assert_allclose(Vd_m[w], d, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(Vn_m[w], n, rtol=1e-5, atol=1e-5)
        differs = differs or e_m[w] != e_f[w]
    # fp64 zones are untouched: kinetic energy (cached fp64 gradients / Laplacians), bare Coulomb, local ECP
    np.testing.assert_allclose(T_m, T_f, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(V_m[:, :2], V_f[:, :2], rtol=1e-13)
    assert differs, "the mixed engine returned bit-identical energies: the fp32 zones did not run"


def test_mixed_metropolis_and_projection_statistics():
    """The fp32 zones perturb ratios by ~1e-6: single decisions may flip, the Markov chain statistics may not."""
    from jqmc_b200.engine import WalkerEngine

    H = _system()
    nw = 512
    r_up, r_dn = random_walkers(H, nw, 17, scale=0.7)
    keys = np.stack([np.zeros(nw, np.uint32), np.arange(nw, dtype=np.uint32) + 99], axis=1)
    out = {}
    for prec in ("full", "mixed"):
        eng = WalkerEngine(H, precision=prec)
        G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
        acc, rej, ru, rd, k2, Gi2, G2 = eng.update(r_up, r_dn, keys, 40, 2.0, 0.0, Ginv, G)
        Gf, _ = eng.geminal_inv_batched(ru, rd)
        w, pu, pd, Gi3, k3, RT, Vd, Vn = eng.projection_n(np.ones(nw), ru, rd, eng.A_inv_n(ru, rd), k2, -17.2, 10, True, "tmove", 0.3)
        out[prec] = dict(acc=acc.cpu().numpy(), ru=ru.cpu().numpy(), keys=k2.cpu().numpy(), w=w.cpu().numpy(), e=(Vd + Vn).cpu().numpy(),
                         inv_err=float((Gi2 @ Gf - __import__("torch").eye(4, device=Gf.device, dtype=Gf.dtype)).abs().max()))  # fmt: skip
    f, m = out["full"], out["mixed"]
    np.testing.assert_array_equal(f["keys"], m["keys"])  # the random stream does not depend on the precision mode
    assert np.mean(f["acc"] == m["acc"]) > 0.9  # almost every walker takes exactly the same decisions
    same = np.all(np.abs(f["ru"] - m["ru"]) < 1e-9, axis=(1, 2))
    assert same.mean() > 0.9
    # accepted rows of G are built from fp32 AO values (relative error ~1e-7, as in the reference: ao_eval feeds the geminal), so
    # the running inverse is the inverse of THAT matrix: against the fp64 geminal it is off by ~1e-7 cond(G), not by round-off
    assert m["inv_err"] < 1e-2
    assert abs(f["acc"].mean() - m["acc"].mean()) < 0.5
    assert np.all(np.isfinite(m["w"])) and np.all(m["w"] > 0)
    assert abs(np.mean(f["e"]) - np.mean(m["e"])) < 0.2


def test_mixed_requires_valid_precision():
    from jqmc_b200.engine import WalkerEngine

    with pytest.raises(ValueError):
        WalkerEngine(_system(), precision="half")
