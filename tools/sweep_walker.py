#!/usr/bin/env python
"""Launch-shape sweep of the fused walker kernel (tuning aid, never a bench number): warps per CTA (16 = one CTA per SM,
8 = two, 4 = four) x walkers per CTA, timed with CUDA events on the launch stream for the LRDMC projection (nmpm = 40),
V elements and the VMC local energy of BASELINE configs[1].

    python tools/sweep_walker.py --out gpurun_out/sweep_walker.json
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
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import numpy as np
    import torch

    import bench
    from jqmc_b200.engine import WalkerEngine

    torch.cuda.set_device(0)
    H = bench.make_hamiltonian()
    eng = WalkerEngine(H)
    nw = args.walkers
    r_up, r_dn, keys = bench.init_walkers(H, nw, bench.SEED)
    dev = eng.device
    r_up, r_dn, keys = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (r_up, r_dn, keys))
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    for _ in range(3):
        _, _, r_up, r_dn, keys, Ginv, G = eng.update(r_up, r_dn, keys, 40, 2.0, 0.0, Ginv, G, inplace=True)
    RT = eng.generate_RTs(keys)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2, as bench.py does between steps

    def timed(fn):
        ts = []
        for i in range(args.reps + 1):
            flush.zero_()
            ev[0].record()
            out = fn()
            ev[1].record()
            torch.cuda.synchronize()
            if i:
                ts.append(ev[0].elapsed_time(ev[1]))
        return float(np.mean(ts)), out

    res = []
    ref = None
    for warps, wpcs in ((16, (0, 28, 29, 30, 31, 32)), (8, (0, 14)), (4, (0, 7))):
        for wpc in wpcs:
            eng.set_walker_warps(warps)
            eng.set_walkers_per_cta(wpc)
            try:
                w1 = torch.ones(nw, dtype=torch.float64, device=dev)
                t_p, out = timed(lambda: eng.projection_n(w1, r_up, r_dn, Ginv, keys, -17.2, 40, True, "tmove", 0.30))
                t_v, _ = timed(lambda: eng.V_elements_n(r_up, r_dn, RT, "tmove", 0.30, A_inv=Ginv))
                t_e, _ = timed(lambda: eng.e_L_fast(r_up, r_dn, RT, Ginv))
            except Exception as e:  # launch shape not possible (shared memory)
                print(warps, wpc, "failed:", e, flush=True)
                continue
            eng.profile(True)  # the kernel alone, by the engine's own CUDA events (what bench.py reports)
            for _ in range(args.reps):
                flush.zero_()
                eng.projection_n(w1, r_up, r_dn, Ginv, keys, -17.2, 40, True, "tmove", 0.30)
            torch.cuda.synchronize()
            prof = eng.profile_read()
            eng.profile(False)
            t_k = prof["k_walker(projection)"][0] / max(1, prof["k_walker(projection)"][1])
            wsum = float(out[0].sum())
            if ref is None:
                ref = wsum
            row = dict(warps=warps, wpc=wpc, projection_kernel_ms=t_k, projection_ms=t_p, V_elements_ms=t_v, e_L_ms=t_e, same_result=abs(wsum - ref) <= 1e-9 * abs(ref))
            res.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
