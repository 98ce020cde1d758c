#!/usr/bin/env python
"""Short VMC + LRDMC step loop for ncu captures (never a bench number).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --steps 2
    ncu --set full --clock-control none --import-source on -k regex:k_mcmc -s 1 -c 1 -o gpurun_out/prof_mcmc \
        python tools/profile_step.py --steps 2 --vmc-only
"""

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--walkers", type=int, default=4096)
    ap.add_argument("--vmc-only", action="store_true")
    ap.add_argument("--lrdmc-only", action="store_true")
    ap.add_argument("--system", default="water")
    args = ap.parse_args()
    import torch

    import bench
    from jqmc_b200.engine import WalkerEngine

    torch.cuda.set_device(0)
    H = bench.make_hamiltonian(args.system) if args.system != "water" else bench.make_hamiltonian()
    eng = WalkerEngine(H)
    r_up, r_dn, keys = bench.init_walkers(H, args.walkers, bench.SEED)
    dev = eng.device
    r_up, r_dn, keys = (torch.from_numpy(x).to(dev) for x in (r_up, r_dn, keys))
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    if not args.lrdmc_only:
        for _ in range(args.steps):
            acc, rej, r_up, r_dn, keys, Ginv, G = eng.update(r_up, r_dn, keys, bench.NMPM, bench.DT, bench.EPS_AS, Ginv, G, inplace=True)
            RT = eng.generate_RTs(keys)
            e_L = eng.e_L_fast(r_up, r_dn, RT, Ginv)
            eng.as_reg_fast(G, Ginv)
        torch.cuda.synchronize()
        print("VMC e_L mean", float(e_L.mean()))
    if not args.vmc_only:
        w = torch.ones(args.walkers, dtype=torch.float64, device=dev)
        E_scf = -17.0
        for _ in range(args.steps):
            Ginv = eng.A_inv_n(r_up, r_dn)
            w.fill_(1.0)
            w, r_up, r_dn, Ginv, keys, RT, Vd, Vn = eng.projection_n(w, r_up, r_dn, Ginv, keys, E_scf, bench.NMPM, True, "tmove", 0.30, inplace=True)
            Vd, Vn = eng.V_elements_n(r_up, r_dn, RT, "tmove", 0.30)
        torch.cuda.synchronize()
        print("LRDMC e_L mean", float((Vd + Vn).mean()))


if __name__ == "__main__":
    main()
