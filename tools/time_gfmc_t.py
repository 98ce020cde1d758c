#!/usr/bin/env python
"""Step time of GFMC_t (lrdmc-tau) on BASELINE configs[1]'s system (never the bench line; bench.py measures VMC + GFMC_n).

    python tools/time_gfmc_t.py --out gpurun_out/gfmc_t.json [--walkers 4096] [--tau 0.10] [--wide]

One step = inverse + the continuous-time projection loop of every walker for the imaginary time tau (CLI defaults of the
reference: tau = 0.10, alat = 0.30, tmove; jqmc/jqmc_miscs.py:165-176) + the per-step sums.  Reported: ms per step (CUDA events on
the launch stream), projections per walker (mean / max), walker-steps/s and projections/s, and the work the reference's
vmapped while_loop would do (every walker runs max-over-walkers iterations) next to what the engine does (every walker runs the
iterations of its own CTA + at most one tail evaluation).
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
    ap.add_argument("--tau", type=float, default=0.10)
    ap.add_argument("--alat", type=float, default=0.30)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--wide", action="store_true", help="force the general kernel family")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import numpy as np
    import torch

    import bench
    from jqmc_b200.engine import WalkerEngine

    torch.cuda.set_device(0)
    H = bench.make_hamiltonian()
    eng = WalkerEngine(H)
    if args.wide:
        eng.set_path(True)
    nw = args.walkers
    r_up, r_dn, keys = bench.init_walkers(H, nw, 5)
    dev = eng.device
    r_up, r_dn, keys = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (r_up, r_dn, keys))
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    for _ in range(3):  # equilibrate with a few Metropolis sweeps so that the walkers are typical configurations
        _, _, r_up, r_dn, keys, Ginv, G = eng.update(r_up, r_dn, keys, 40, 2.0, 0.0, Ginv, G, inplace=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms, pcs = [], []
    e_mean = 0.0
    for it in range(args.steps + 2):
        if it == 2:
            eng.profile(True)
        w = torch.ones(nw, dtype=torch.float64, device=dev)
        ev[0].record()
        Gi = eng.A_inv_n(r_up, r_dn)
        e_L, pc, w, r_up, r_dn, Gi, keys, RT = eng.projection_t(w, r_up, r_dn, Gi, keys, args.tau, True, "tmove", args.alat, inplace=True)
        sums = eng.lrdmc_collect_t(w, e_L)
        ev[1].record()
        torch.cuda.synchronize()
        if it >= 2:
            ms.append(ev[0].elapsed_time(ev[1]))
            pcs.append(pc.cpu().numpy())
            s = sums.cpu().numpy()
            e_mean = s[3] / s[2]
    prof = eng.profile_read()
    eng.profile(False)
    pcs = np.array(pcs)
    t = float(np.mean(ms))
    res = dict(
        system="water ccECP/cc-pVQZ JSD + J2 pade", path="general" if args.wide else "register", walkers=nw, tau=args.tau, alat=args.alat,
        ms_per_step=t, walker_steps_per_s=nw / (t * 1e-3), projections_mean=float(pcs.mean()), projections_max=float(pcs.max(axis=1).mean()),
        projections_per_s=float(pcs.sum(axis=1).mean()) / (t * 1e-3),
        while_loop_iterations_reference=float(pcs.max(axis=1).mean()),  # every walker of the rank, vmapped
        e_L_mixed=float(e_mean),
        kernels={k: dict(ms_per_step=v[0] / args.steps, launches_per_step=v[1] / args.steps) for k, v in prof.items() if v[1]},
    )  # fmt: skip
    print(json.dumps(res), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
