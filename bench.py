#!/usr/bin/env python
"""bench.py -- walker-steps/s of the VMC and LRDMC hot path (water ccECP/cc-pVQZ JSD, 4096 walkers per GPU).

A bench "step" is one pass of BOTH drivers' step over every walker of the rank (BASELINE.json configs[1],
"VMC+LRDMC"):
    VMC   (MCMC.run, jqmc_mcmc.py:664-747):   nmpm=40 Metropolis proposals -> rotation draw -> local energy -> AS weight
    LRDMC (GFMC_n.run, jqmc_gfmc.py:5774-6321): nmpm=40 projections -> V_diag/V_nondiag -> weighted sums (+all_reduce)
                                               -> walker reconfiguration (all_gather + comb + gather) -> inverse refresh
`value` counts one walker-step per driver per walker (2 * walkers * steps / time); the per-driver rates are in
the "vmc" and "lrdmc" objects of the same line.

    python bench.py --gpus 1 --steps 20 --warmup 5            # this engine
    python bench.py --impl reference --steps 1 --warmup 0     # CPU oracle port on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # one rank per GPU, walkers sharded (weak scaling)

Prints ONE JSON line on rank 0 (see the keys below).  Nothing here reads /root/reference.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "walker-steps/sec (VMC & LRDMC, water ccECP/cc-pVQZ JSD)"
UNIT = "walker-steps/s"
NW_PER_GPU = 4096
NMPM = 40
DT = 2.0
EPS_AS = 0.0
SEED = 34456
ALAT = 0.30
E_SCF = -17.2
NLM = "tmove"
CPU_PROJ_SAMPLE = 4  # projections actually run per walker by the CPU arm (of NMPM; time scaled by NMPM / CPU_PROJ_SAMPLE)


def make_hamiltonian():
    from jqmc_b200.data import Jastrow_data, Jastrow_two_body_data
    from jqmc_b200.trexio_lite import load_golden_system

    H = load_golden_system(os.path.join(ROOT, "tests", "golden", "water_ccecp_ccpvqz.npz"))
    # J2 Pade a = 1.0 as in the reference's benchmarks/benchmark_local_energy.py:26-48
    H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=1.0, jastrow_2b_type="pade"))
    return H


def init_walkers(H, nw, seed):
    from jqmc_b200 import rng_host
    from jqmc_b200.mcmc import generate_init_electron_configurations

    np.random.seed(seed % (2**32))
    gem = H.wavefunction_data.geminal_data
    r_up, r_dn, _, _ = generate_init_electron_configurations(
        gem.num_electron_up, gem.num_electron_dn, nw, H.coulomb_potential_data.effective_charges, H.structure_data.positions
    )
    keys = rng_host.split(rng_host.PRNGKey(seed), nw)
    return np.ascontiguousarray(r_up), np.ascontiguousarray(r_dn), keys


def algorithmic_flops(H):
    """Irreducible fp64 work per unit, SURVEY.md §8(d) formulas evaluated on this system (exp = 21 flops)."""
    gem = H.wavefunction_data.geminal_data
    aos = gem.orb_data_up_spin.aos_data
    n_ao, n_prim, n_mo = aos.num_ao, aos.num_ao_prim, gem.orb_data_up_spin.num_mo
    n_up, n_dn = gem.num_electron_up, gem.num_electron_dn
    n_e = n_up + n_dn
    n_at = len(H.structure_data.atomic_numbers)
    c_ang, c_val = 60, 15
    pts = n_e * 1 * 6 if H.coulomb_potential_data.ecp_flag else 0
    F_pt = n_prim * 30 + n_ao * c_val + 2 * n_mo * n_ao + 2 * n_up
    F_eL = n_e * (n_prim * 30 + n_ao * c_ang) + 5 * 2 * n_mo * n_ao * n_e + 10 * n_up**3 + pts * F_pt
    F_mh = F_pt + 4 * n_up**2 + 6 * n_up**2 + 20 * n_e + 10 * n_at
    state_bytes = 2 * (24 * n_e + 2 * 8 * n_up**2 + 8) + 32
    # LRDMC: one projection = 6 N_e kinetic + N_e*NN*Nv ECP mesh ratios (F_point each), per-electron kinetic
    # energy from cached derivatives (10 n_mo per electron), then value/grad/lap of the moved electron + Sherman-Morrison
    F_vgl_pt = n_prim * 30 + n_ao * c_ang + 5 * 2 * n_mo * n_ao + 4 * n_up**2
    F_proj = (6 * n_e + pts) * F_pt + n_e * 10 * n_mo + n_e * (20 * n_e + 10 * n_at)
    F_lrdmc_step = NMPM * (F_proj + F_vgl_pt) + (F_proj + n_e * F_vgl_pt) + 2 * (n_e * F_pt + 4 * n_up**3)
    return dict(F_eL=F_eL, F_mh=F_mh, F_point=F_pt, F_step=F_eL + NMPM * F_mh, bytes_step=state_bytes, F_vgl_point=F_vgl_pt,
                F_lrdmc_proj=F_proj, F_lrdmc_step=F_lrdmc_step)


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.f, stderr=subprocess.DEVNULL,
            )  # fmt: skip
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the NumPy oracle port, one walker-step of each driver per task, spread over host processes
# ------------------------------------------------------------------------------------------------
def _cpu_walker_step(args):
    w, r_up, r_dn, key = args
    from oracle import drivers as OD
    from oracle import physics as OP

    H = make_hamiltonian()
    gem = H.wavefunction_data.geminal_data
    G, Ginv = OD.geminal_inv(gem, r_up, r_dn)
    t0 = time.perf_counter()
    _, _, ru, rd, k2, Ginv, G = OD.update_electron_positions(H, r_up, r_dn, key, NMPM, DT, EPS_AS, Ginv, G)
    RT = OD.generate_rotation_matrix(k2)
    e = OP.compute_local_energy(H, ru, rd, RT, Ginv=Ginv)
    OP.compute_AS_regularization_factor(G, Ginv)
    t1 = time.perf_counter()
    # bounded sample: CPU_PROJ_SAMPLE of the NMPM projections are run and their time is scaled (every projection does the
    # same work: 96 mesh ratios + one Sherman-Morrison update); V elements and the inverse refresh are run in full
    wl, ru, rd, _, k3, RT, _, _ = OD.lrdmc_projection(H, 1.0, ru, rd, Ginv, k2, E_SCF, CPU_PROJ_SAMPLE, True, NLM, ALAT)
    t2 = time.perf_counter()
    d, n = OD.lrdmc_V_elements(H, ru, rd, RT, NLM, ALAT)
    OD.geminal_inv(gem, ru, rd)
    t3 = time.perf_counter()
    return t1 - t0, (t2 - t1) * (NMPM / CPU_PROJ_SAMPLE) + (t3 - t2), float(e), float(d + n)


def cpu_sample(n_walkers, procs):
    """Time `n_walkers` oracle walker-steps of each driver on `procs` host processes.
    Returns (combined walker-steps/s, seconds, per-task results)."""
    import multiprocessing as mp

    H = make_hamiltonian()
    r_up, r_dn, keys = init_walkers(H, n_walkers, SEED)
    tasks = [(w, r_up[w], r_dn[w], (int(keys[w, 0]), int(keys[w, 1]))) for w in range(n_walkers)]
    t0 = time.perf_counter()
    if procs > 1:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_walker_step, tasks)
    else:
        res = [_cpu_walker_step(t) for t in tasks]
    wall = time.perf_counter() - t0
    # per-task compute time with the LRDMC projection loop scaled to NMPM; `procs` tasks run concurrently
    dt = float(np.mean([r[0] + r[1] for r in res])) * max(1, (n_walkers + procs - 1) // procs)
    return 2 * n_walkers / dt, dt, res, wall


def _cpu_split(res, procs):
    """Per-driver walker-steps/s of a CPU sample (in-task compute time, `procs` tasks in parallel)."""
    tv = float(np.mean([r[0] for r in res]))
    tl = float(np.mean([r[1] for r in res]))
    return dict(vmc_walker_steps_per_s=procs / tv, lrdmc_walker_steps_per_s=procs / tl)


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    n_w = procs  # one walker-step of each driver per process per bench step
    times, res = [], []
    n_warm, n_steps = min(args.warmup, 1), max(1, min(args.steps, 4))  # bounded: the whole run ends within a few minutes
    for _ in range(n_warm):
        cpu_sample(n_w, procs)
    for _ in range(n_steps):
        v, dt, res, _ = cpu_sample(n_w, procs)
        times.append(dt)
    dt = float(np.mean(times))
    value = 2 * n_w / dt
    sample = (f"{n_w} walkers x [1 VMC step (nmpm={NMPM} + e_L + AS) + 1 LRDMC step ({CPU_PROJ_SAMPLE} of {NMPM} projections run, time scaled "
              f"x{NMPM // CPU_PROJ_SAMPLE}; V elements + inverse in full)] per bench step, {procs} processes, {n_steps} timed step(s) of the "
              f"{args.steps} requested, NumPy restatement (oracle/), not the JAX reference")  # fmt: skip
    line = dict(
        impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=n_steps, warmup=n_warm,
        ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
        config=dict(workload="water ccECP/cc-pVQZ JSD VMC+LRDMC (J2 pade a=1), CPU sample", walkers=n_w, nmpm=NMPM, Dt=DT, epsilon_AS=EPS_AS,
                    alat=ALAT, non_local_move=NLM, E_scf=E_SCF),
        cpu_baseline=dict(value=value, unit=UNIT, cores=procs, kind="port", sample=sample, **_cpu_split(res, procs)),
        e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0,
    )  # fmt: skip
    _emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from jqmc_b200.engine import WalkerEngine, measure_fp64_peak
    from jqmc_b200.gfmc import GFMC_n

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H = make_hamiltonian()
    eng = WalkerEngine(H)
    nw = args.walkers
    r_up_h, r_dn_h, keys_h = init_walkers(H, nw, SEED * (rank + 1))
    r_up = torch.from_numpy(r_up_h).to(dev)
    r_dn = torch.from_numpy(r_dn_h).to(dev)
    keys = torch.from_numpy(keys_h).to(dev)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gf = GFMC_n(H, num_walkers=nw, num_mcmc_per_measurement=NMPM, mcmc_seed=SEED, E_scf=E_SCF, alat=ALAT, non_local_move=NLM, engine=eng)
    zeta_rng = np.random.RandomState(SEED)

    def step_vmc(state):
        r_up, r_dn, keys, G, Ginv = state
        acc, rej, r_up, r_dn, keys, Ginv, G = eng.update(r_up, r_dn, keys, NMPM, DT, EPS_AS, Ginv, G, inplace=True)
        RT = eng.generate_RTs(keys)
        e_L = eng.e_L_fast(r_up, r_dn, RT, Ginv)
        R_AS = eng.as_reg_fast(G, Ginv)
        return (r_up, r_dn, keys, G, Ginv), (e_L, R_AS, acc, rej)

    def step_lrdmc(state):
        r_up, r_dn, keys, A_inv = state
        r_up, r_dn, keys, A_inv, sums, n_surv, _ = gf._step(r_up, r_dn, keys, A_inv, float(zeta_rng.random_sample()), rank, world)
        return (r_up, r_dn, keys, A_inv), (sums, n_surv)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    vstate = (r_up, r_dn, keys, G, Ginv)
    # equilibrate a little so the timed walkers are typical configurations, then warm up
    for _ in range(max(args.warmup, 3)):
        vstate, vobs = step_vmc(vstate)
    lstate = (vstate[0].clone(), vstate[1].clone(), vstate[2].clone(), eng.A_inv_n(vstate[0], vstate[1]))
    for _ in range(max(args.warmup, 3)):
        lstate, lobs = step_lrdmc(lstate)
    torch.cuda.synchronize()

    # ---- device-resident timing: per-step CUDA event pairs, L2 flushed between steps ---------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    mk = lambda: [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]  # noqa: E731
    ev_v, ev_l = mk(), mk()
    l0 = eng.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        ev_v[i][0].record()
        vstate, vobs = step_vmc(vstate)
        ev_v[i][1].record()
        flush.zero_()
        ev_l[i][0].record()
        lstate, lobs = step_lrdmc(lstate)
        ev_l[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count() - l0
    ms_v = float(np.sum([a.elapsed_time(b) for a, b in ev_v]))
    ms_l = float(np.sum([a.elapsed_time(b) for a, b in ev_l]))
    e_mean = float(vobs[0].mean().item())
    acc_ratio = float(vobs[2].double().sum().item() / (nw * NMPM))
    ls = lobs[0].cpu().numpy()
    e_lrdmc = float(ls[3] / ls[2])
    surv = float(lobs[1].item()) / (nw * world)

    # ---- end-to-end: host buffers in, host results out, through the same public calls -----------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    host_v = [pin(t.cpu().numpy()) for t in vstate]
    host_vo = [torch.empty(nw, dtype=torch.float64).pin_memory() for _ in range(2)] + [torch.empty(nw, dtype=torch.int32).pin_memory() for _ in range(2)]
    host_l = [pin(t.cpu().numpy()) for t in lstate[:3]]  # r_up, r_dn, keys (the inverse is rebuilt on the device, as the reference does)
    host_lo = [torch.empty(5, dtype=torch.float64).pin_memory(), torch.empty(1, dtype=torch.int32).pin_memory()]
    nbytes = lambda ts: sum(t.numel() * t.element_size() for t in ts)  # noqa: E731
    h2d = nbytes(host_v) + nbytes(host_l)
    d2h = nbytes(host_v) + nbytes(host_vo) + nbytes(host_l) + nbytes(host_lo)
    n_e2e = max(3, min(args.steps, 20))

    def e2e_step():
        dstate = tuple(t.to(dev, non_blocking=True) for t in host_v)
        dstate, dobs = step_vmc(dstate)
        for h, d in zip(host_v, dstate):
            h.copy_(d, non_blocking=True)
        for h, d in zip(host_vo, dobs):
            h.copy_(d, non_blocking=True)
        lr = tuple(t.to(dev, non_blocking=True) for t in host_l)
        lr = lr + (eng.A_inv_n(lr[0], lr[1]),)
        lr, lo = step_lrdmc(lr)
        for h, d in zip(host_l, lr[:3]):
            h.copy_(d, non_blocking=True)
        for h, d in zip(host_lo, lo):
            h.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    t_e2e = (time.perf_counter() - t0) / n_e2e
    clocks = sampler.stop()

    # ---- per-kernel share and roofline (separate profiled pass; events on the launch stream) ---------
    eng.profile(True)
    n_prof = max(3, min(args.steps, 10))
    for _ in range(n_prof):
        flush.zero_()
        vstate, vobs = step_vmc(vstate)
        flush.zero_()
        lstate, lobs = step_lrdmc(lstate)
    torch.cuda.synchronize()
    prof = eng.profile_read()
    eng.profile(False)
    fp64_peak = measure_fp64_peak(40000) if rank == 0 else 0.0

    # max over ranks
    t = torch.tensor([ms_v, ms_l, t_e2e, t_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_v, ms_l, t_e2e, t_wall = (float(x) for x in t.tolist())

    if rank == 0:
        total_walkers = nw * world
        ms_dev = ms_v + ms_l
        value = 2 * total_walkers * args.steps / (ms_dev * 1e-3)
        e2e_value = 2 * total_walkers / t_e2e
        fl = algorithmic_flops(H)
        tot_ms = sum(v[0] for v in prof.values())
        kern = {k: dict(ms_per_launch=v[0] / v[1], launches_per_step=v[1] / n_prof, share=v[0] / tot_ms) for k, v in prof.items()}
        dom = max(prof.items(), key=lambda kv: kv[1][0])[0]
        per_launch_flops = {
            "k_mcmc": fl["F_mh"] * NMPM * nw,
            "k_walker(e_L)": fl["F_eL"] * nw,
            "k_walker(V_elements)": fl["F_lrdmc_proj"] * nw,
            "k_walker(projection)": (fl["F_lrdmc_proj"] + fl["F_vgl_point"]) * NMPM * nw,
        }
        rl_kernels = {}
        for k, f in per_launch_flops.items():
            if k in kern:
                a = f / (kern[k]["ms_per_launch"] * 1e-3) / 1e12
                rl_kernels[k] = dict(achieved=a, frac=a / fp64_peak if fp64_peak else None, flops_per_launch=f)
        dom_ms = kern[dom]["ms_per_launch"]
        achieved = per_launch_flops.get(dom, fl["F_step"] * nw) / (dom_ms * 1e-3) / 1e12
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:  # per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture, if present
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(dom)
        except Exception:
            pass
        roofline = dict(
            bound="fp64", kernel=dom, achieved=achieved, peak=fp64_peak, unit="TFLOP/s", frac=achieved / fp64_peak if fp64_peak else None,
            traffic=traffic,
            peak_source="DFMA microbenchmark measured live in this run (qe_measure_fp64_peak); MEASURED_PEAKS.json has no fp64 entry",
            note="the path is fp64-ALU bound (SURVEY.md 8d): HBM carries only walker state, see hbm",
            whole_step=dict(achieved=(fl["F_step"] + fl["F_lrdmc_step"]) * nw / (ms_dev / args.steps * 1e-3) / 1e12, unit="TFLOP/s"),
            hbm=dict(achieved=2 * fl["bytes_step"] * nw / (ms_dev / args.steps * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s",
                     peak_source="MEASURED_PEAKS.json" if peaks else "fallback"),
            per_kernel=rl_kernels, kernels=kern,
        )  # fmt: skip
        cpu = None
        if world == 1 and not args.no_cpu:
            procs = max(1, min(os.cpu_count() or 1, 16))
            v, dt, res, wall = cpu_sample(procs, procs)
            cpu = dict(value=v, unit=UNIT, cores=procs, kind="port",
                       sample=(f"{procs} walkers x (1 VMC + 1 LRDMC step; {CPU_PROJ_SAMPLE} of {NMPM} projections run, time scaled) on {procs} "
                               f"processes ({wall:.1f} s wall), NumPy restatement (oracle/), not the JAX reference"),
                       **_cpu_split(res, procs))  # fmt: skip
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
            config=dict(
                workload="water ccECP/cc-pVQZ JSD VMC+LRDMC (J2 pade a=1.0), BASELINE configs[1]", walkers_per_gpu=nw, nmpm=NMPM, Dt=DT,
                epsilon_AS=EPS_AS, alat=ALAT, non_local_move=NLM, E_scf=E_SCF, Nv=6, NN=1,
                parallelism=f"walkers sharded over {world} rank(s); VMC: no data-path collective; LRDMC: all_reduce(5 doubles) + all_gather(w, r) per branching",
                l2="flushed between steps (256 MiB memset outside the per-step CUDA-event pairs)",
                timing="sum of per-step CUDA-event durations on the launch stream, max over ranks",
            ),
            vmc=dict(value=total_walkers * args.steps / (ms_v * 1e-3), unit=UNIT, ms_per_step=ms_v / args.steps),
            lrdmc=dict(value=total_walkers * args.steps / (ms_l * 1e-3), unit=UNIT, ms_per_step=ms_l / args.steps),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=n_e2e,
                     path="WalkerEngine update/generate_RTs/e_L_fast/as_reg_fast + GFMC_n._step on pinned host buffers"),
            gpu_launches=int(launches),
            clocks=clocks, roofline=roofline, cpu_baseline=cpu,
            check=dict(e_L_vmc_mean=e_mean, acceptance=acc_ratio, e_L_lrdmc=e_lrdmc, survived_ratio=surv, wall_s=t_wall),
        )  # fmt: skip
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--walkers", type=int, default=NW_PER_GPU, help="walkers per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 (NCCL prints its version banner there) are sent to
    # stderr, and the result line goes to the saved descriptor
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
