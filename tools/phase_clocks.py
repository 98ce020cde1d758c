#!/usr/bin/env python
"""Per-phase cycle breakdown of the fused walker kernel (diagnostic, never a bench number).

    python tools/phase_clocks.py [--walkers 4096] [--steps 3] [--out gpurun_out/phases.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch

    import bench
    from jqmc_b200.engine import WalkerEngine

    torch.cuda.set_device(0)
    H = bench.make_hamiltonian()
    eng = WalkerEngine(H)
    r_up, r_dn, keys = bench.init_walkers(H, args.walkers, bench.SEED)
    dev = eng.device
    r_up, r_dn, keys = (torch.from_numpy(x).to(dev) for x in (r_up, r_dn, keys))
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    w = torch.ones(args.walkers, dtype=torch.float64, device=dev)
    res = {}
    for name in ("projection", "V_elements", "e_L"):
        for it in range(args.steps + 1):
            if it == 1:
                eng.phase_clocks(True)
                t0 = torch.cuda.Event(enable_timing=True)
                t1 = torch.cuda.Event(enable_timing=True)
                t0.record()
            if name == "projection":
                Ginv = eng.A_inv_n(r_up, r_dn)
                w.fill_(1.0)
                w, r_up, r_dn, Ginv, keys, RT, Vd, Vn = eng.projection_n(w, r_up, r_dn, Ginv, keys, -17.0, bench.NMPM, True, "tmove", 0.30, inplace=True)
            elif name == "V_elements":
                eng.V_elements_n(r_up, r_dn, RT, "tmove", 0.30)
            else:
                eng.e_L_fast(r_up, r_dn, RT, Ginv)
        t1.record()
        torch.cuda.synchronize()
        clk = eng.phase_clocks(False)
        tot = sum(clk.values()) or 1
        res[name] = {"ms_per_call_incl_clocks": t0.elapsed_time(t1) / args.steps, "share": {k: round(v / tot, 4) for k, v in clk.items() if v}}
        print(name, res[name])
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
