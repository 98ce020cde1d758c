#!/usr/bin/env python
"""Step times of the BASELINE configurations that run on the general kernel family (never the bench line; bench.py measures
configs[1]):  water JAGP (configs[2]), benzene shape JSD+J1J2J3 (configs[3]), 100 e / 1000 AO JSD (configs[4]).

    python tools/time_configs.py --out gpurun_out/configs.json [--cases water_jagp,benzene,S] [--walkers 4096] [--s-walkers 1024]

Per case: ms of one VMC step (nmpm Metropolis proposals + RT + e_L + AS factor) and one LRDMC step (inverse + nmpm projections
+ V elements), CUDA events on the launch stream after warm-up, plus the engine's per-kernel CUDA-event shares (qe_profile).
"""

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(case):
    import dataclasses

    import numpy as np

    import bench
    from jqmc_b200 import synthetic as SY
    from jqmc_b200.data import Geminal_data

    if case == "water_jsd_wide":
        return bench.make_hamiltonian(), True
    if case == "water_jagp":
        H = bench.make_hamiltonian()
        gem = Geminal_data.convert_from_MOs_to_AOs(H.wavefunction_data.geminal_data)
        lam = np.array(gem.lambda_matrix)
        H.wavefunction_data.geminal_data = dataclasses.replace(gem, lambda_matrix=lam + np.random.default_rng(1).normal(scale=1e-3, size=lam.shape))
        return H, False
    if case == "benzene":
        return SY.benzene_shape(), False
    if case == "benzene_jagp":
        return SY.benzene_shape(jagp=True), False
    if case == "S":
        return SY.grid_molecule(), False
    raise KeyError(case)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="water_jsd_wide,water_jagp,benzene,S")
    ap.add_argument("--walkers", type=int, default=4096)
    ap.add_argument("--nmpm", type=int, default=40)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--s-walkers", type=int, default=1024, help="walkers of the 100 e / 1000 AO case (configs[4] sweeps 1k..64k per GPU)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import numpy as np
    import torch

    import bench
    from jqmc_b200 import rng_host
    from jqmc_b200 import synthetic as SY
    from jqmc_b200.engine import WalkerEngine

    torch.cuda.set_device(0)
    res = {}
    for case in args.cases.split(","):
        H, force = build(case)
        eng = WalkerEngine(H)
        if force:
            eng.set_path(True)
        gem = H.wavefunction_data.geminal_data
        nw = args.walkers if case != "S" else args.s_walkers
        if len(H.structure_data.positions) > 3:
            r_up, r_dn = SY.init_walkers(H, nw, 1, sigma=0.8)
            keys = rng_host.split(rng_host.PRNGKey(5), nw)
        else:
            r_up, r_dn, keys = bench.init_walkers(H, nw, 5)
        dev = eng.device
        r_up, r_dn, keys = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (r_up, r_dn, keys))
        G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t_v, t_l = [], []
        E_scf = None
        for it in range(args.steps + 1):
            if it == 1:
                eng.profile(True)
            ev[0].record()
            acc, rej, r_up, r_dn, keys, Ginv, G = eng.update(r_up, r_dn, keys, args.nmpm, 2.0, 0.0, Ginv, G, inplace=True)
            RT = eng.generate_RTs(keys)
            e_L = eng.e_L_fast(r_up, r_dn, RT, Ginv)
            eng.as_reg_fast(G, Ginv)
            ev[1].record()
            if E_scf is None:
                E_scf = float(e_L.mean()) - 0.05 * abs(float(e_L.mean()))
            w = torch.ones(nw, dtype=torch.float64, device=dev)
            ev[2].record()
            Gi = eng.A_inv_n(r_up, r_dn)
            w, ru2, rd2, Gi, k2, RT2, Vd, Vn = eng.projection_n(w, r_up, r_dn, Gi, keys, E_scf, args.nmpm, True, "tmove", 0.30)
            Vd, Vn = eng.V_elements_n(ru2, rd2, RT2, "tmove", 0.30)
            ev[3].record()
            torch.cuda.synchronize()
            if it >= 1:
                t_v.append(ev[0].elapsed_time(ev[1]))
                t_l.append(ev[2].elapsed_time(ev[3]))
        prof = eng.profile_read()
        eng.profile(False)
        tot = sum(v[0] for v in prof.values()) or 1.0
        kern = {k: dict(ms_per_step=v[0] / args.steps, launches_per_step=v[1] / args.steps, share=v[0] / tot) for k, v in prof.items() if v[1]}
        res[case] = dict(
            walkers=nw, n_up=gem.num_electron_up, n_dn=gem.num_electron_dn, n_orb=gem.orb_num_up, nmpm=args.nmpm,
            vmc_ms_per_step=float(np.mean(t_v)), lrdmc_ms_per_step=float(np.mean(t_l)),
            vmc_walker_steps_per_s=nw / (np.mean(t_v) * 1e-3), lrdmc_walker_steps_per_s=nw / (np.mean(t_l) * 1e-3),
            acceptance=float(acc.double().mean()) / args.nmpm, e_L_mean=float(e_L.mean()), e_L_lrdmc=float((Vd + Vn).mean()),
            kernels=kern,
        )  # fmt: skip
        print(case, json.dumps({k: v for k, v in res[case].items() if k != "kernels"}), flush=True)
        for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["share"])[:8]:
            print(f"   {k:24s} {v['ms_per_step']:9.3f} ms/step  {v['launches_per_step']:7.1f} launches  share {v['share']:.3f}", flush=True)
        del eng
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
