#!/usr/bin/env python
"""bench.py -- walker-steps/s of the VMC and LRDMC hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

Workloads (`--config`, default = BASELINE.json configs[1], the configuration the metric is quoted on):

    water_jsd       configs[1]  water ccECP/cc-pVQZ JSD + J2, VMC + LRDMC, 4096 walkers per GPU            (register kernels)
    water_jagp      configs[2]  water ccECP/cc-pVQZ JAGP (AO-basis geminal 114 x 114), LRDMC with reconfiguration (general family)
    benzene_sr      configs[3]  benzene-shape ccECP/cc-pVTZ JSD + J1J2J3, VMC with O_k collection + one SR solve  (general family)
    synthetic_100e  configs[4]  synthetic 100 e / 1000 AO JSD, VMC + LRDMC, `--walkers` 1k .. 64k per GPU       (general family)

A bench "step" is one pass of the named drivers' step over every walker of the rank:
    VMC   (MCMC.run, jqmc_mcmc.py:664-747):   nmpm=40 Metropolis proposals -> rotation draw -> local energy -> AS weight
    LRDMC (GFMC_n.run, jqmc_gfmc.py:5774-6321): nmpm=40 projections -> V_diag/V_nondiag -> weighted sums -> walker
                                               reconfiguration (one packed all_gather + comb + gather) -> inverse refresh
`value` counts one walker-step per driver per walker; per-driver rates are in the "vmc" / "lrdmc" objects.

    python bench.py [--config water_jsd] --gpus 1 --steps K --warmup W     # this engine
    python bench.py --impl reference --steps K --warmup W                  # CPU arm on the host cores (see run_reference)
    torchrun --nproc-per-node N ... bench.py --gpus N ...                  # one rank per GPU, walkers sharded (weak scaling)

Nothing here reads /root/reference.
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

UNIT = "walker-steps/s"
NW_PER_GPU = 4096
NMPM = 40
DT = 2.0
EPS_AS = 0.0
SEED = 34456
ALAT = 0.30
E_SCF = -17.2
NLM = "tmove"

CONFIGS = {
    "water_jsd": dict(
        baseline=1, metric="walker-steps/sec (VMC & LRDMC, water ccECP/cc-pVQZ JSD)",
        workload="water ccECP/cc-pVQZ JSD VMC+LRDMC (J2 pade a=1.0), BASELINE configs[1]", walkers=NW_PER_GPU, legs=("vmc", "lrdmc"),
        steps=200, warmup=5,
    ),
    "water_jagp": dict(
        baseline=2, metric="walker-steps/sec (LRDMC with walker reconfiguration, water ccECP/cc-pVQZ JAGP)",
        workload="water ccECP/cc-pVQZ JAGP (AO-basis geminal, lambda 114x114) LRDMC with walker reconfiguration (J2 pade a=1.0), BASELINE configs[2]",
        walkers=NW_PER_GPU, legs=("lrdmc",), steps=30, warmup=3,
    ),
    "benzene_sr": dict(
        baseline=3, metric="walker-steps/sec (VMC with O_k collection + SR solve, benzene-shape ccECP/cc-pVTZ JSD+J1J2J3)",
        workload="benzene ccECP/cc-pVTZ SHAPE (12 atoms, 30 e, 258 AOs, 15 MOs, J1+J2+J3 on 36 AOs; synthetic coefficients: the reference ships "
                 "no benzene input) VMC with parameter derivatives + one stochastic-reconfiguration solve, BASELINE configs[3]",
        walkers=1024, legs=("vmc_sr",), steps=20, warmup=3,
    ),
    "synthetic_100e": dict(
        baseline=4, metric="walker-steps/sec (VMC & LRDMC, synthetic 100 e / 1000 AO JSD)",
        workload="synthetic 100-electron / 1000-AO / 50-MO JSD molecule (25 atoms, Cartesian s..g, 2-channel ECP) VMC+LRDMC, BASELINE configs[4]",
        walkers=1024, legs=("vmc", "lrdmc"), steps=3, warmup=3,
    ),
}  # fmt: skip


# ------------------------------------------------------------------------------------------------
# systems and walkers
# ------------------------------------------------------------------------------------------------
def make_hamiltonian(config: str = "water_jsd"):
    import dataclasses

    from jqmc_b200.data import Geminal_data, Jastrow_data, Jastrow_two_body_data
    from jqmc_b200.trexio_lite import load_golden_system

    if config in ("water_jsd", "water_jagp", "water"):
        H = load_golden_system(os.path.join(ROOT, "tests", "golden", "water_ccecp_ccpvqz.npz"))
        # J2 Pade a = 1.0 as in the reference's benchmarks/benchmark_local_energy.py:26-48
        H.wavefunction_data.jastrow_data = Jastrow_data(jastrow_two_body_data=Jastrow_two_body_data(jastrow_2b_param=1.0, jastrow_2b_type="pade"))
        if config == "water_jagp":  # Geminal_data.convert_from_MOs_to_AOs (determinant.py:686-722), SURVEY.md §8(d).3
            gem = Geminal_data.convert_from_MOs_to_AOs(H.wavefunction_data.geminal_data)
            lam = np.array(gem.lambda_matrix)
            H.wavefunction_data.geminal_data = dataclasses.replace(gem, lambda_matrix=lam + np.random.default_rng(1).normal(scale=1e-3, size=lam.shape))
        return H
    from jqmc_b200 import synthetic as SY

    if config == "benzene_sr":
        return SY.benzene_shape()
    if config == "synthetic_100e":
        return SY.grid_molecule()
    raise KeyError(config)


def init_walkers(H, nw, seed):
    from jqmc_b200 import rng_host
    from jqmc_b200.mcmc import generate_init_electron_configurations

    keys = rng_host.split(rng_host.PRNGKey(seed), nw)
    if len(H.structure_data.positions) > 3:  # synthetic systems: electrons = atom centres + N(0, sigma) (benchmark_mcmc_kernels.py:440-446)
        from jqmc_b200 import synthetic as SY

        r_up, r_dn = SY.init_walkers(H, nw, seed % 65536, sigma=0.8)
        return np.ascontiguousarray(r_up), np.ascontiguousarray(r_dn), keys
    np.random.seed(seed % (2**32))
    gem = H.wavefunction_data.geminal_data
    r_up, r_dn, _, _ = generate_init_electron_configurations(
        gem.num_electron_up, gem.num_electron_dn, nw, H.coulomb_potential_data.effective_charges, H.structure_data.positions
    )
    return np.ascontiguousarray(r_up), np.ascontiguousarray(r_dn), keys


def unique_shell_primitives(aos) -> int:
    """Exponentials one AO sweep actually evaluates: the reference stores every shell's primitives once per AO
    (num_ao_prim = 160 for water cc-pVQZ); the engine folds them back into shells (64).  Same grouping as build_basis."""
    oi = np.asarray(aos.orbital_indices)
    ex, co = np.asarray(aos.exponents, dtype=np.float64), np.asarray(aos.coefficients, dtype=np.float64)
    seen = set()
    n = 0
    for a in range(aos.num_ao):
        sel = oi == a
        e, c = ex[sel], co[sel]
        key = (int(aos.nucleus_index[a]), int(aos.angular_momentums[a]), tuple(np.round(e, 12)), tuple(np.round(c / c[np.argmax(np.abs(c))], 9)))
        if key not in seen:
            seen.add(key)
            n += len(e)
    return n


def algorithmic_flops(H, n_prim=None):
    """Irreducible fp64 work per unit, SURVEY.md §8(d) formulas evaluated on this system (exp = 1 transcendental + 20 flops, 30
    flops per primitive with its accumulation).  n_prim: primitives per AO sweep -- default the reference's per-AO replicated
    count (`num_ao_prim`, the §8(d) figure); pass unique_shell_primitives() for the work the engine executes."""
    gem = H.wavefunction_data.geminal_data
    orb = gem.orb_data_up_spin
    aos = getattr(orb, "aos_data", orb)
    n_ao = aos.num_ao
    n_prim = aos.num_ao_prim if n_prim is None else n_prim
    n_up, n_dn = gem.num_electron_up, gem.num_electron_dn
    # width of the per-point contraction: AO -> MO (n_mo columns) for an MO-basis geminal; for an AO-basis (JAGP) geminal the
    # reference contracts the AO row with the pre-contracted M = lambda Phi_dn, i.e. N_up columns (determinant.py:1665-1783)
    n_mo = getattr(orb, "num_mo", 0) or n_up
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
                F_lrdmc_proj=F_proj, F_lrdmc_step=F_lrdmc_step, n_prim=n_prim)  # fmt: skip


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
# CPU arm.  Preferred: the reference itself (jQMC's jitted JAX path on the host cores) when `jax` and `jqmc` import on this
# box; they do not in this image (probe recorded in DESIGN.md §6), so the arm that actually runs is the NumPy
# restatement under oracle/ ("port"): one FULL walker-step of each driver per task -- all nmpm Metropolis proposals, all nmpm
# projections, V elements, inverse refresh; nothing is extrapolated -- on one host process per walker.
# ------------------------------------------------------------------------------------------------
def _cpu_walker_step(args):
    config, r_up, r_dn, key, legs = args
    from oracle import drivers as OD
    from oracle import physics as OP

    H = make_hamiltonian(config)
    gem = H.wavefunction_data.geminal_data
    G, Ginv = OD.geminal_inv(gem, r_up, r_dn)
    t_v = t_l = 0.0
    e = el = float("nan")
    ru, rd, k2 = r_up, r_dn, key
    if "vmc" in legs or "vmc_sr" in legs:
        t0 = time.perf_counter()
        _, _, ru, rd, k2, Ginv, G = OD.update_electron_positions(H, r_up, r_dn, key, NMPM, DT, EPS_AS, Ginv, G)
        RT = OD.generate_rotation_matrix(k2)
        e = float(OP.compute_local_energy(H, ru, rd, RT, Ginv=Ginv))
        OP.compute_AS_regularization_factor(G, Ginv)
        if "vmc_sr" in legs:
            OP.compute_dln_wf_dparams(H.wavefunction_data, ru, rd, Ginv)
        t_v = time.perf_counter() - t0
    if "lrdmc" in legs:
        t0 = time.perf_counter()
        _, ru, rd, _, k3, RT, _, _ = OD.lrdmc_projection(H, 1.0, ru, rd, Ginv, k2, _e_scf(config), NMPM, True, NLM, ALAT)
        d, n = OD.lrdmc_V_elements(H, ru, rd, RT, NLM, ALAT)
        OD.geminal_inv(gem, ru, rd)
        t_l = time.perf_counter() - t0
        el = float(d + n)
    return t_v, t_l, e, el


def _e_scf(config):
    return E_SCF if config.startswith("water") else -1.0e3  # synthetic systems: any value safely below the spectrum


def cpu_sample(config, n_walkers, procs, legs):
    """One full oracle walker-step per driver for `n_walkers` walkers on `procs` host processes.
    Returns (walker-steps/s, seconds of this sample = wall time of the pool, per-task results)."""
    import multiprocessing as mp

    H = make_hamiltonian(config)
    r_up, r_dn, keys = init_walkers(H, n_walkers, SEED)
    tasks = [(config, r_up[w], r_dn[w], (int(keys[w, 0]), int(keys[w, 1])), legs) for w in range(n_walkers)]
    t0 = time.perf_counter()
    if procs > 1:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_walker_step, tasks)
    else:
        res = [_cpu_walker_step(t) for t in tasks]
    wall = time.perf_counter() - t0
    # in-task compute time (excludes interpreter start-up of the pool); `procs` tasks run concurrently
    dt = float(np.mean([r[0] + r[1] for r in res])) * max(1, (n_walkers + procs - 1) // procs)
    n_drivers = sum(1 for leg in legs if leg in ("vmc", "vmc_sr", "lrdmc"))
    return n_drivers * n_walkers / dt, dt, res, wall


def _cpu_split(res, procs):
    tv = float(np.mean([r[0] for r in res]))
    tl = float(np.mean([r[1] for r in res]))
    out = {}
    if tv > 0:
        out["vmc_walker_steps_per_s"] = procs / tv
    if tl > 0:
        out["lrdmc_walker_steps_per_s"] = procs / tl
    return out


def _try_reference_jax(config, n_walkers, steps, warmup):
    """Time the reference itself (jqmc's MCMC.run / GFMC_n.run, JAX on the host cores) when it can be imported on this box.
    Returns None when jax / jqmc are not importable (this image: no jax, no flax, no mpi4py, no wheel for them)."""
    if config != "water_jsd":
        return None
    try:
        os.environ.setdefault("JAX_PLATFORMS", "cpu")
        ref = os.path.join(ROOT, "baseline", "_ref")
        if os.path.isdir(ref) and ref not in sys.path:
            sys.path.insert(0, ref)
        import jax  # noqa: F401

        jax.config.update("jax_enable_x64", True)
        from jqmc.jastrow_factor import Jastrow_data as RJ
        from jqmc.jastrow_factor import Jastrow_two_body_data as RJ2
        from jqmc.jqmc_gfmc import GFMC_n as RGFMC
        from jqmc.jqmc_mcmc import MCMC as RMCMC
    except Exception:
        return None
    try:
        from jqmc.hamiltonians import Hamiltonian_data as RH
        from jqmc.trexio_wrapper import read_trexio_file
        from jqmc.wavefunction import Wavefunction_data as RW

        h5 = os.path.join(ROOT, "baseline", "water_ccecp_ccpvqz.h5")
        if not os.path.exists(h5):
            return None
        st, aos, mos_u, mos_d, gem, cp = read_trexio_file(h5, store_tuple=True)
        jd = RJ(jastrow_one_body_data=None, jastrow_two_body_data=RJ2(jastrow_2b_param=1.0), jastrow_three_body_data=None)
        H = RH(structure_data=st, coulomb_potential_data=cp, wavefunction_data=RW(jastrow_data=jd, geminal_data=gem))
        m = RMCMC(hamiltonian_data=H, mcmc_seed=SEED, num_walkers=n_walkers, num_mcmc_per_measurement=NMPM, Dt=DT, epsilon_AS=EPS_AS)
        m.run(num_mcmc_steps=max(1, warmup))
        t0 = time.perf_counter()
        m.run(num_mcmc_steps=steps)
        t_v = time.perf_counter() - t0
        g = RGFMC(hamiltonian_data=H, num_walkers=n_walkers, num_mcmc_per_measurement=NMPM, mcmc_seed=SEED, E_scf=E_SCF, alat=ALAT, non_local_move=NLM)
        g.run(num_mcmc_steps=max(1, warmup))
        t0 = time.perf_counter()
        g.run(num_mcmc_steps=steps)
        t_l = time.perf_counter() - t0
        return dict(t_vmc=t_v, t_lrdmc=t_l, steps=steps, walkers=n_walkers, jax=jax.__version__)
    except Exception as e:  # an importable but unusable install must not break the arm
        sys.stderr.write(f"reference JAX arm failed ({type(e).__name__}: {e}); falling back to the NumPy port\n")
        return None


def run_reference(args, rank):
    """--impl reference: the CPU implementation of the path on the host cores, same metric / unit / config keys as the GPU
    arm.  Every timed step is a FULL bench step (nothing extrapolated) of a bounded sample: one walker per host process."""
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    legs = cfg["legs"]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    n_w = procs
    jx = _try_reference_jax(args.config, 64, max(1, min(args.steps, 20)), min(args.warmup, 3))
    if jx is not None:
        dt = (jx["t_vmc"] + jx["t_lrdmc"]) / jx["steps"]
        value = 2 * jx["walkers"] / dt
        line = dict(
            impl="reference", metric=cfg["metric"], value=value, unit=UNIT, n_gpus=args.gpus, steps=jx["steps"], warmup=min(args.warmup, 3),
            ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
            config=dict(workload=cfg["workload"] + ", CPU sample", walkers=jx["walkers"], nmpm=NMPM, Dt=DT, epsilon_AS=EPS_AS, alat=ALAT,
                        non_local_move=NLM, E_scf=E_SCF),
            cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="reference", extrapolated=False,
                              sample=f"jQMC MCMC.run + GFMC_n.run, JAX {jx['jax']} on the host cores, {jx['walkers']} walkers, {jx['steps']} steps",
                              vmc_walker_steps_per_s=jx["walkers"] * jx["steps"] / jx["t_vmc"],
                              lrdmc_walker_steps_per_s=jx["walkers"] * jx["steps"] / jx["t_lrdmc"]),
            e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
        )  # fmt: skip
        _emit(line)
        return
    # NumPy port.  A full water step of one walker costs ~30 s of one core (40 brute-force projections), so the number of timed
    # steps is bounded such that the run ends within a few minutes; every step that IS run is run in full.
    n_warm = min(args.warmup, 0 if args.config != "water_jsd" else 1)
    n_steps = max(1, min(args.steps, 3 if args.config == "water_jsd" else 1))
    times, res = [], []
    for _ in range(n_warm):
        cpu_sample(args.config, n_w, procs, legs)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        v, dt, res, _ = cpu_sample(args.config, n_w, procs, legs)
        times.append(dt)
    wall = time.perf_counter() - t0
    dt = float(np.mean(times))
    n_drivers = len(legs)
    value = n_drivers * n_w / dt
    sample = (f"{n_w} walkers (one per host process, {procs} processes) x full bench steps: "
              + " + ".join({"vmc": f"1 VMC step (nmpm={NMPM} proposals + e_L + AS)", "vmc_sr": f"1 VMC step (nmpm={NMPM} + e_L + AS + O_k)",
                            "lrdmc": f"1 LRDMC step (all {NMPM} projections + V elements + inverse)"}[leg] for leg in legs)
              + f"; {n_steps} timed step(s) of the {args.steps} requested ({wall:.0f} s wall); NumPy restatement (oracle/), NOT the JAX reference: "
              "jax / jqmc do not import on this box")  # fmt: skip
    line = dict(
        impl="reference", metric=cfg["metric"], value=value, unit=UNIT, n_gpus=args.gpus, steps=n_steps, warmup=n_warm,
        ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
        config=dict(workload=cfg["workload"] + ", CPU sample", walkers=n_w, nmpm=NMPM, Dt=DT, epsilon_AS=EPS_AS, alat=ALAT, non_local_move=NLM,
                    E_scf=_e_scf(args.config)),
        cpu_baseline=dict(value=value, unit=UNIT, cores=procs, kind="port", extrapolated=False, sample=sample, **_cpu_split(res, procs)),
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

    cfg = CONFIGS[args.config]
    legs = cfg["legs"]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H = make_hamiltonian(args.config)
    eng = WalkerEngine(H, precision=args.precision)
    nw = args.walkers
    e_scf = _e_scf(args.config)
    r_up_h, r_dn_h, keys_h = init_walkers(H, nw, SEED * (rank + 1))
    r_up = torch.from_numpy(r_up_h).to(dev)
    r_dn = torch.from_numpy(r_dn_h).to(dev)
    keys = torch.from_numpy(keys_h).to(dev)
    G, Ginv = eng.geminal_inv_batched(r_up, r_dn)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    has_vmc = "vmc" in legs or "vmc_sr" in legs
    has_lrdmc = "lrdmc" in legs
    with_ok = "vmc_sr" in legs
    gf = None
    if has_lrdmc:
        gf = GFMC_n(H, num_walkers=nw, num_mcmc_per_measurement=NMPM, mcmc_seed=SEED, E_scf=e_scf, alat=ALAT, non_local_move=NLM, engine=eng)
    zeta_rng = np.random.RandomState(SEED)
    ok_store = []  # device-resident O_k samples of the SR leg (never copied to the host)

    def step_vmc(state, keep_ok=False):
        r_up, r_dn, keys, G, Ginv = state
        acc, rej, r_up, r_dn, keys, Ginv, G = eng.update(r_up, r_dn, keys, NMPM, DT, EPS_AS, Ginv, G, inplace=True)
        RT = eng.generate_RTs(keys)
        e_L = eng.e_L_fast(r_up, r_dn, RT, Ginv)
        R_AS = eng.as_reg_fast(G, Ginv)
        if with_ok:
            g = eng.grad_ln_psi_params_fast(r_up, r_dn, Ginv)
            if keep_ok:
                ok_store.append((torch.cat([v.reshape(nw, -1) for v in g.values()], dim=1), e_L))
        return (r_up, r_dn, keys, G, Ginv), (e_L, R_AS, acc, rej)

    def step_lrdmc(state):
        r_up, r_dn, keys, A_inv = state
        r_up, r_dn, keys, A_inv, sums, n_surv, _, _ = gf._step(r_up, r_dn, keys, A_inv, float(zeta_rng.random_sample()), rank, world)
        return (r_up, r_dn, keys, A_inv), (sums, n_surv)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    vstate = (r_up, r_dn, keys, G, Ginv)
    vobs = lobs = None
    n_warm = max(args.warmup, 3)
    # equilibrate a little so the timed walkers are typical configurations, then warm up
    for _ in range(n_warm):
        vstate, vobs = step_vmc(vstate)
    lstate = None
    if has_lrdmc:
        lstate = (vstate[0].clone(), vstate[1].clone(), vstate[2].clone(), eng.A_inv_n(vstate[0], vstate[1]))
        for _ in range(n_warm):
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
        if has_vmc:
            flush.zero_()
            ev_v[i][0].record()
            vstate, vobs = step_vmc(vstate, keep_ok=True)
            ev_v[i][1].record()
        if has_lrdmc:
            flush.zero_()
            ev_l[i][0].record()
            lstate, lobs = step_lrdmc(lstate)
            ev_l[i][1].record()
    sr = None
    if with_ok and ok_store:  # (library warm-up outside the timed solve: cuBLAS / cuSOLVER handles, workspaces)
        from jqmc_b200.sr import sr_natural_gradient as _sr_warm

        _O = torch.stack([o for o, _ in ok_store[:2]])
        _e = torch.stack([e for _, e in ok_store[:2]])
        _sr_warm(torch.ones_like(_e), _e, _O, epsilon=1e-3, use_cg=False, force_dual=False)
        torch.cuda.synchronize()
    if with_ok:  # one stochastic-reconfiguration solve on the samples of the timed steps (all on the device, all ranks)
        from jqmc_b200.sr import sr_natural_gradient

        ev_s = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev_s[0].record()
        O = torch.stack([o for o, _ in ok_store])
        e = torch.stack([e for _, e in ok_store])
        theta, info = sr_natural_gradient(torch.ones_like(e), e, O, epsilon=1e-3, use_cg=O.shape[2] > 2000)
        ev_s[1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count() - l0
    ms_v = float(np.sum([a.elapsed_time(b) for a, b in ev_v])) if has_vmc else 0.0
    ms_l = float(np.sum([a.elapsed_time(b) for a, b in ev_l])) if has_lrdmc else 0.0
    ms_sr = 0.0
    if with_ok:
        ms_sr = float(ev_s[0].elapsed_time(ev_s[1]))
        sr = dict(ms_per_solve=ms_sr, parameters=int(O.shape[2]), samples_per_rank=int(O.shape[0] * O.shape[1]),
                  allreduce_bytes=int(info.get("allreduce_bytes", 0)), theta_norm=float(theta.norm().item()))  # fmt: skip
        del O
        ok_store.clear()
    check = dict(wall_s=t_wall)
    if has_vmc:
        check.update(e_L_vmc_mean=float(vobs[0].mean().item()), acceptance=float(vobs[2].double().sum().item() / (nw * NMPM)))
    if has_lrdmc:
        ls = lobs[0].cpu().numpy()
        check.update(e_L_lrdmc=float(ls[3] / ls[2]), survived_ratio=float(lobs[1].item()) / (nw * world))

    # ---- end-to-end: host buffers in, host results out, through the same public calls -----------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    nbytes = lambda ts: sum(t.numel() * t.element_size() for t in ts)  # noqa: E731
    h2d = d2h = 0
    host_v = host_vo = host_l = host_lo = None
    if has_vmc:
        host_v = [pin(t.cpu().numpy()) for t in vstate]
        host_vo = [torch.empty(nw, dtype=torch.float64).pin_memory() for _ in range(2)] + [torch.empty(nw, dtype=torch.int32).pin_memory() for _ in range(2)]
        h2d += nbytes(host_v)
        d2h += nbytes(host_v) + nbytes(host_vo)
    if has_lrdmc:
        host_l = [pin(t.cpu().numpy()) for t in lstate[:3]]  # r_up, r_dn, keys (the inverse is rebuilt on the device, as the reference does)
        host_lo = [torch.empty(5, dtype=torch.float64).pin_memory(), torch.empty(1, dtype=torch.int32).pin_memory()]
        h2d += nbytes(host_l)
        d2h += nbytes(host_l) + nbytes(host_lo)
    n_e2e = max(3, min(args.steps, 20))

    copy_stream = torch.cuda.Stream(device=dev)  # host<->device copies of one driver overlap the other driver's kernels

    def e2e_step():
        main = torch.cuda.current_stream()
        ev_in_v = ev_in_l = None
        with torch.cuda.stream(copy_stream):  # every input of the step leaves the host at once
            if has_vmc:
                dstate = tuple(t.to(dev, non_blocking=True) for t in host_v)
                ev_in_v = copy_stream.record_event()
            if has_lrdmc:
                lr = tuple(t.to(dev, non_blocking=True) for t in host_l)
                ev_in_l = copy_stream.record_event()
        if has_vmc:
            main.wait_event(ev_in_v)
            dstate, dobs = step_vmc(dstate)
            ev_v = main.record_event()
            with torch.cuda.stream(copy_stream):  # results go back while the LRDMC step runs
                copy_stream.wait_event(ev_v)
                for h, d in zip(host_v, dstate):
                    h.copy_(d, non_blocking=True)
                for h, d in zip(host_vo, dobs):
                    h.copy_(d, non_blocking=True)
        if has_lrdmc:
            main.wait_event(ev_in_l)
            lr = lr + (eng.A_inv_n(lr[0], lr[1]),)
            lr, lo = step_lrdmc(lr)
            for h, d in zip(host_l, lr[:3]):
                h.copy_(d, non_blocking=True)
            for h, d in zip(host_lo, lo):
                h.copy_(d, non_blocking=True)
        copy_stream.synchronize()
        main.synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    t_e2e = (time.perf_counter() - t0) / n_e2e
    if with_ok:  # the SR solve of the timed samples with its result read back to the host, amortised over the steps like `value`
        th_host = torch.empty(O.shape[2], dtype=torch.float64).pin_memory()
        barrier()
        t0 = time.perf_counter()
        th2, _ = sr_natural_gradient(torch.ones_like(e), e, O, epsilon=1e-3, use_cg=O.shape[2] > 2000)
        th_host.copy_(th2, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        barrier()
        t_e2e += (time.perf_counter() - t0) / args.steps
        d2h += th_host.numel() * 8 // args.steps
    clocks = sampler.stop()

    # ---- per-kernel share and roofline (separate profiled pass; events on the launch stream) ---------
    eng.profile(True)
    n_prof = max(2, min(args.steps, 10))
    for _ in range(n_prof):
        if has_vmc:
            flush.zero_()
            vstate, vobs = step_vmc(vstate)
        if has_lrdmc:
            flush.zero_()
            lstate, lobs = step_lrdmc(lstate)
    torch.cuda.synchronize()
    prof = eng.profile_read()
    eng.profile(False)
    fp64_peak = measure_fp64_peak(40000) if rank == 0 else 0.0

    # max over ranks
    t = torch.tensor([ms_v, ms_l, t_e2e, t_wall, ms_sr], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_v, ms_l, t_e2e, t_wall, ms_sr = (float(x) for x in t.tolist())

    if rank == 0:
        total_walkers = nw * world
        n_drivers = int(has_vmc) + int(has_lrdmc)
        ms_dev = ms_v + ms_l + ms_sr
        value = n_drivers * total_walkers * args.steps / (ms_dev * 1e-3)
        e2e_value = n_drivers * total_walkers / t_e2e
        gem = H.wavefunction_data.geminal_data
        aos = getattr(gem.orb_data_up_spin, "aos_data", gem.orb_data_up_spin)
        n_uniq = unique_shell_primitives(aos)
        fl = algorithmic_flops(H)  # SURVEY.md 8(d) convention: the reference's replicated primitive count
        fx = algorithmic_flops(H, n_uniq)  # executed work: unique shell primitives
        tot_ms = sum(v[0] for v in prof.values())
        kern = {k: dict(ms_per_launch=v[0] / v[1], launches_per_step=v[1] / n_prof, share=v[0] / tot_ms) for k, v in prof.items()}
        dom = max(prof.items(), key=lambda kv: kv[1][0])[0]

        def per_launch(f):
            return {
                "k_mcmc": f["F_mh"] * NMPM * nw,
                "k_walker(e_L)": f["F_eL"] * nw,
                "k_walker(V_elements)": f["F_lrdmc_proj"] * nw,
                "k_walker(projection)": (f["F_lrdmc_proj"] + f["F_vgl_point"]) * NMPM * nw,
            }

        pl, px = per_launch(fl), per_launch(fx)
        rl_kernels = {}
        for k in pl:
            if k in kern:
                s = kern[k]["ms_per_launch"] * 1e-3
                rl_kernels[k] = dict(achieved=pl[k] / s / 1e12, frac=pl[k] / s / 1e12 / fp64_peak if fp64_peak else None,
                                     achieved_executed=px[k] / s / 1e12, frac_executed=px[k] / s / 1e12 / fp64_peak if fp64_peak else None,
                                     flops_per_launch=pl[k], flops_per_launch_executed=px[k])  # fmt: skip
        step_flops = (fl["F_step"] if has_vmc else 0) + (fl["F_lrdmc_step"] if has_lrdmc else 0)
        step_flops_x = (fx["F_step"] if has_vmc else 0) + (fx["F_lrdmc_step"] if has_lrdmc else 0)
        step_s = ms_dev / args.steps * 1e-3
        if dom in pl:
            dom_s = kern[dom]["ms_per_launch"] * 1e-3
            achieved, achieved_x = pl[dom] / dom_s / 1e12, px[dom] / dom_s / 1e12
            scope = "dominant kernel, per launch"
        else:  # general family: many kernels per step -- the whole step against the same peak, kernel shares in `kernels`
            achieved, achieved_x = step_flops * nw / step_s / 1e12, step_flops_x * nw / step_s / 1e12
            scope = "whole step (general kernel family: ~16 launches per projection; per-kernel time shares in `kernels`)"
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
            bound="fp64", kernel=dom, scope=scope, achieved=achieved, peak=fp64_peak, unit="TFLOP/s",
            frac=achieved / fp64_peak if fp64_peak else None,
            achieved_executed=achieved_x, frac_executed=achieved_x / fp64_peak if fp64_peak else None,
            convention=(f"`achieved`/`frac`: SURVEY.md 8(d) algorithmic work with the reference's per-AO replicated primitives (n_prim = {fl['n_prim']}, "
                        f"30 flops each); `achieved_executed`/`frac_executed`: the same formulas with the {n_uniq} unique shell primitives the engine "
                        "evaluates -- the hardware-utilisation figure"),
            traffic=traffic,
            peak_source="DFMA microbenchmark measured live in this run (qe_measure_fp64_peak); MEASURED_PEAKS.json has no fp64 entry",
            note=("the path is fp64-ALU bound (SURVEY.md 8d): HBM carries only walker state, see hbm" if args.precision == "full" else
                  "precision 'mixed': the AO values and Jastrow ratios run on the fp32 pipe, so the fp64 peak is NOT the bound of this mode and "
                  "`frac` (reference-formulation flops / fp64 peak) can exceed 1; reported for comparison with the fp64 line only"),
            whole_step=dict(achieved=step_flops * nw / step_s / 1e12, achieved_executed=step_flops_x * nw / step_s / 1e12, unit="TFLOP/s"),
            hbm=dict(achieved=n_drivers * fl["bytes_step"] * nw / step_s / 1e9, peak=hbm_peak, unit="GB/s",
                     peak_source="MEASURED_PEAKS.json" if peaks else "fallback"),
            per_kernel=rl_kernels, kernels=kern,
        )  # fmt: skip
        cpu = None
        if world == 1 and not args.no_cpu and args.config == "synthetic_100e":
            # one brute-force oracle LRDMC step of a 100-electron walker (600 + 600 mesh points x 40 projections) takes hours
            cpu = dict(value=None, unit=UNIT, cores=0, kind="port", extrapolated=False,
                       sample="not run: a full oracle step of this system does not fit a bounded CPU sample; see --config water_jsd")
        elif world == 1 and not args.no_cpu:
            procs = max(1, min(os.cpu_count() or 1, 16))
            v, dt, res, wall = cpu_sample(args.config, procs, procs, legs)
            cpu = dict(value=v, unit=UNIT, cores=procs, kind="port", extrapolated=False,
                       sample=(f"{procs} walkers x one FULL step of each driver ({', '.join(legs)}; all {NMPM} proposals / projections) on {procs} "
                               f"processes ({wall:.1f} s wall), NumPy restatement (oracle/), not the JAX reference"),
                       **_cpu_split(res, procs))  # fmt: skip
        coll = "no data-path collective"
        if has_lrdmc:
            coll = "VMC: no data-path collective; LRDMC: ONE packed all_gather (sums | w | r_up | r_dn) per branching" if has_vmc else \
                   "LRDMC: ONE packed all_gather (sums | w | r_up | r_dn) per branching"
        elif with_ok:
            coll = "VMC: no data-path collective; SR solve: all_reduce of K-vectors (and of the K x K / sample-space matrix)"
        line = dict(
            metric=cfg["metric"], value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=n_warm,
            ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f64" if args.precision == "full" else "mixed (fp32 zones ao_eval, jastrow_eval, jastrow_ratio; fp64 elsewhere)", data="synthetic",
            config=dict(
                workload=cfg["workload"], precision=args.precision, baseline_config_index=cfg["baseline"], walkers_per_gpu=nw, nmpm=NMPM, Dt=DT,
                epsilon_AS=EPS_AS, alat=ALAT, non_local_move=NLM, E_scf=e_scf, Nv=6, NN=1,
                parallelism=f"walkers sharded over {world} rank(s); {coll}",
                l2="flushed between steps (256 MiB memset outside the per-step CUDA-event pairs)",
                timing="sum of per-step CUDA-event durations on the launch stream, max over ranks",
            ),
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=n_e2e,
                     path="WalkerEngine update/generate_RTs/e_L_fast/as_reg_fast (+ grad_ln_psi_params_fast) and GFMC_n._step on pinned host buffers; host<->device copies on a second stream"),
            gpu_launches=int(launches),
            clocks=clocks, roofline=roofline, cpu_baseline=cpu, check=check,
        )  # fmt: skip
        if has_vmc:
            line["vmc"] = dict(value=total_walkers * args.steps / ((ms_v + ms_sr) * 1e-3), unit=UNIT, ms_per_step=ms_v / args.steps)
        if has_lrdmc:
            line["lrdmc"] = dict(value=total_walkers * args.steps / (ms_l * 1e-3), unit=UNIT, ms_per_step=ms_l / args.steps)
        if sr is not None:
            line["sr"] = sr
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
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: per config; 200 for water_jsd)")
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="water_jsd", choices=sorted(CONFIGS))
    ap.add_argument("--walkers", type=int, default=None, help="walkers per GPU (default: per config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--precision", default="full", choices=["full", "mixed"],
                    help="mixed = the reference's mixed-precision mode (fp32 AO values / Jastrow ratios); reported with dtype \"mixed\", never the headline")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.steps is None:
        args.steps = cfg["steps"]
    if args.warmup is None:
        args.warmup = cfg["warmup"]
    if args.walkers is None:
        args.walkers = cfg["walkers"]
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
