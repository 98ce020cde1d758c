#!/usr/bin/env python
"""E(a) and the generalised force f(a) = -dE/da of the one-parameter J2 (Pade) water wavefunction: a consistency check of the
sign and scale of MCMC.get_gF against the slope of independently sampled energies (never a bench number)."""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from jqmc_b200.data import Jastrow_data, Jastrow_two_body_data
    from jqmc_b200.mcmc import MCMC

    H0 = bench.make_hamiltonian()
    rows = []
    for a in (0.4, 0.7, 1.0, 1.5, 2.5, 4.0):
        H = copy.deepcopy(H0)
        H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=a))
        m = MCMC(H, mcmc_seed=11, num_walkers=4096, num_mcmc_per_measurement=40, Dt=2.0, epsilon_AS=0.0, comput_log_WF_param_deriv=True)
        m.run(140)
        E, dE, _, _ = m.get_E(40, 10)
        f, df = m.get_gF(40, 10, blocks=["j2_param"])
        rows.append((a, E, dE, f[0], df[0]))
        print(f"a = {a:4.1f}   E = {E:.5f} +- {dE:.5f}   f = -dE/da = {f[0]:+.5f} +- {df[0]:.5f}", flush=True)
    for (a0, E0, d0, f0, _), (a1, E1, d1, f1, _) in zip(rows, rows[1:]):
        print(f"[{a0}, {a1}]: -(E1 - E0)/(a1 - a0) = {-(E1 - E0) / (a1 - a0):+.5f} +- {np.hypot(d0, d1) / (a1 - a0):.5f}   mean f = {(f0 + f1) / 2:+.5f}")


if __name__ == "__main__":
    main()
