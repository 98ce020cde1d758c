"""``MCMC`` driver on the walker engine: the host-side mirror of jQMC's VMC sampler.

Same constructor arguments, step loop, stored observables and ``get_E`` statistics as
``jqmc.jqmc_mcmc.MCMC`` (jqmc/jqmc_mcmc.py:129-258 constructor, :448-1062 ``run``, :1064-1189
``get_E``), with the per-step device work done by ``WalkerEngine`` instead of ``jit(vmap(...))``:

    for each measurement step (jqmc_mcmc.py:664-747)
        update   -> qe_mcmc_update        (nmpm Metropolis proposals per walker)
        RTs      -> qe_rotation
        e_L      -> qe_local_energy
        w_L      -> qe_as_factor, w = (R_AS / max(R_AS, eps))^2

Ranks: one process per GPU; rank r seeds with ``mcmc_seed * (r + 1)`` and owns ``num_walkers``
walkers (jqmc_mcmc.py:191-197).  The only collectives are the ``get_E`` reductions, done with
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) when a process group is initialised.

With ``comput_log_WF_param_deriv=True`` every step also stores O_k = d ln|Psi| / d parameter (qe_dln_wf; blocks ``j1_param``,
``j2_param``, ``j3_matrix``, ``lambda_matrix`` as in jqmc/wavefunction.py:542-624), and ``get_dln_WF`` / ``get_gF`` give the
flattened derivative matrix and the jackknifed generalised forces (jqmc_mcmc.py:1372-1513, 1516-1692).

With ``comput_position_deriv=True`` every step also stores the Hellmann-Feynman and Pulay force products (position
derivatives by finite differences on the device + SWCT, jqmc_b200/forces.py) and ``get_aF`` gives the jackknifed atomic forces.

Out of scope here (SURVEY.md §8f "next"): e_L parameter derivatives (linear method), the adaptive learning rate / SNR filters of
the reference's optimiser.
"""

from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import rng_host
from .engine import WalkerEngine


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def _rank_size():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def should_stop(t_start: float, max_time: float, toml_filename: str, device=None) -> bool:
    """Rank 0 decides (wall-clock limit exceeded, or ``[external_control] stop = true`` in the run's control file) and every
    rank receives that decision, so that all ranks leave the step loop at the same iteration and none is left waiting in a
    collective (jqmc/jqmc_gfmc.py:6386-6414, jqmc/jqmc_mcmc.py:985-1010)."""
    import os

    rank, world = _rank_size()
    stop = False
    if rank == 0:
        stop = max_time < time.perf_counter() - t_start
        if not stop and os.path.isfile(toml_filename):
            import tomllib

            try:
                with open(toml_filename, "rb") as f:
                    stop = bool(tomllib.load(f).get("external_control", {}).get("stop", False))
            except (tomllib.TOMLDecodeError, OSError):
                stop = False
    if world > 1:
        stop = bool(_allreduce_sum([1.0 if stop else 0.0], device)[0] > 0.0)
    return stop


def write_control_file(toml_filename: str) -> None:
    """Rank 0 (re)creates the external control file with ``stop = false`` (jqmc/jqmc_gfmc.py:4684-4698)."""
    rank, _ = _rank_size()
    if rank == 0:
        with open(toml_filename, "w") as f:
            f.write("[external_control]\nstop = false\n")


def _allreduce_sum(values, device=None):
    """Sum a small fp64 vector over ranks (NCCL needs CUDA tensors, gloo CPU tensors)."""
    d = _dist()
    a = np.atleast_1d(np.asarray(values, dtype=np.float64))
    if d is None:
        return a
    dev = device if (device is not None and d.get_backend() == "nccl") else torch.device("cpu")
    t = torch.from_numpy(a.copy()).to(dev)
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return t.cpu().numpy()


def generate_init_electron_configurations(n_up, n_dn, num_walkers, charges, coords):
    """Initial walkers: electrons assigned to nuclei (valence-filling order over a farthest-atom sequence),
    each placed on a uniform shell 0.1-1.0 bohr around its owner.

    TRANSLITERATED FROM the reference's ``_generate_init_electron_configurations`` (jqmc/_jqmc_utility.py:54-218): the
    assignment state machine and -- deliberately -- the order of the ``np.random`` calls (dn offsets, then up offsets;
    distance / theta / phi blocks) are the reference's, because a seeded run has to start from the reference's configurations.
    It is host-side set-up outside the engine's scope: when running under jQMC, call the reference's own function
    and hand its arrays to the drivers; this copy exists so that the package runs where jQMC is not installed.
    Returns (r_up[nw,n_up,3], r_dn[nw,n_dn,3], up_owner, dn_owner)."""
    coords = np.asarray(coords, dtype=np.float64)
    nion = coords.shape[0]
    zeta = np.array([int(round(c)) for c in np.asarray(charges)], dtype=int)
    max_dn = zeta // 2
    # farthest-from-previous atom sequence
    seq = [0]
    free = np.ones(nion, dtype=bool)
    free[0] = False
    for _ in range(1, nion):
        d2 = np.sum((coords[seq[-1]] - coords) ** 2, axis=1)
        nxt = int(np.argmax(np.where(free, d2, -1.0)))
        seq.append(nxt)
        free[nxt] = False
    seq = np.array(seq)
    occ_tot = np.zeros(nion, dtype=int)
    occ_dn = np.zeros(nion, dtype=int)
    occ_up = np.zeros(nion, dtype=int)
    dn_tmpl = np.empty(n_dn, dtype=int)
    j = 0
    for i in range(n_dn):
        while True:
            a = seq[j % nion]
            if np.any(occ_dn < max_dn):
                ok = occ_dn[a] < max_dn[a]
            elif np.any((max_dn == 0) & (occ_tot < zeta)):
                ok = max_dn[a] == 0 and occ_tot[a] < zeta[a]
            else:
                ok = occ_tot[a] < zeta[a]
            j += 1
            if ok:
                dn_tmpl[i] = a
                occ_dn[a] += 1
                occ_tot[a] += 1
                break
    need = zeta - occ_dn
    up_tmpl = np.empty(n_up, dtype=int)
    n_extra = 0
    if n_up <= int(need.sum()):
        p = 0
        for i in range(n_up):
            while True:
                a = seq[p % nion]
                p += 1
                if occ_up[a] < need[a]:
                    up_tmpl[i] = a
                    occ_up[a] += 1
                    break
    else:
        c = 0
        for a in seq:
            for _ in range(int(need[a])):
                up_tmpl[c] = a
                c += 1
        n_extra = n_up - int(need.sum())
    dn_owner = np.broadcast_to(dn_tmpl, (num_walkers, n_dn)).copy() if n_dn else np.empty((num_walkers, 0), dtype=int)
    up_owner = np.empty((num_walkers, n_up), dtype=int)
    if n_extra:
        det = n_up - n_extra
        up_owner[:, :det] = up_tmpl[:det][None, :]
        ridx = np.clip(np.floor(np.random.rand(num_walkers, n_extra) * nion).astype(int), 0, nion - 1)
        up_owner[:, det:] = seq[ridx]
    else:
        up_owner[:] = up_tmpl[None, :]

    def offsets(shape):
        dist = np.random.uniform(0.1, 1.0, size=shape)
        th = np.random.uniform(0.0, np.pi, size=shape)
        ph = np.random.uniform(0.0, 2.0 * np.pi, size=shape)
        st = np.sin(th)
        return np.stack([dist * st * np.cos(ph), dist * st * np.sin(ph), dist * np.cos(th)], axis=-1)

    off_dn = offsets((num_walkers, n_dn)) if n_dn else np.zeros((num_walkers, 0, 3))
    off_up = offsets((num_walkers, n_up)) if n_up else np.zeros((num_walkers, 0, 3))
    return coords[up_owner] + off_up, coords[dn_owner] + off_dn, up_owner, dn_owner


class MCMC:
    """VMC sampler (see module docstring).  Public surface follows jqmc.jqmc_mcmc.MCMC."""

    def __init__(
        self,
        hamiltonian_data=None,
        mcmc_seed: int = 34467,
        num_walkers: int = 40,
        num_mcmc_per_measurement: int = 16,
        Dt: float = 2.0,
        epsilon_AS: float = 1e-1,
        comput_log_WF_param_deriv: bool = False,
        comput_e_L_param_deriv: bool = False,
        comput_position_deriv: bool = False,
        random_discretized_mesh: bool = True,
        use_swct: bool = True,
        engine: WalkerEngine | None = None,
    ) -> None:
        if comput_e_L_param_deriv:
            raise NotImplementedError("e_L parameter derivatives (linear method) are outside the walker engine (SURVEY.md §8f)")
        self.__comput_position_deriv = bool(comput_position_deriv)
        self.__comput_log_WF_param_deriv = bool(comput_log_WF_param_deriv)
        self.hamiltonian_data = hamiltonian_data
        self.__mcmc_seed = mcmc_seed
        self.__num_walkers = int(num_walkers)
        self.__nmpm = int(num_mcmc_per_measurement)
        self.__Dt = float(Dt)
        self.__epsilon_AS = float(epsilon_AS)
        self.__random_discretized_mesh = bool(random_discretized_mesh)
        self.__use_swct = bool(use_swct)
        self.__i_opt = 0
        rank, _ = _rank_size()
        self.__mpi_seed = mcmc_seed * (rank + 1)
        self.engine = engine if engine is not None else WalkerEngine(hamiltonian_data)
        dev = self.engine.device
        keys = rng_host.split(rng_host.PRNGKey(self.__mpi_seed), self.__num_walkers)
        self.__keys = torch.from_numpy(keys).to(dev)
        np.random.seed(self.__mpi_seed % (2**32))
        gem = hamiltonian_data.wavefunction_data.geminal_data
        cp = hamiltonian_data.coulomb_potential_data
        r_up, r_dn, _, _ = generate_init_electron_configurations(
            gem.num_electron_up, gem.num_electron_dn, self.__num_walkers, cp.effective_charges,
            hamiltonian_data.structure_data.positions,
        )  # fmt: skip
        self.__r_up = torch.from_numpy(np.ascontiguousarray(r_up)).to(dev)
        self.__r_dn = torch.from_numpy(np.ascontiguousarray(r_dn)).to(dev)
        self.__forces = None
        if self.__comput_position_deriv:  # atomic forces: finite-difference position derivatives on the device (jqmc_b200/forces.py)
            from .forces import ForceEvaluator

            self.__forces = ForceEvaluator(hamiltonian_data, self.engine)
        self.__init_attributes()

    def __init_attributes(self):
        self.__mcmc_counter = 0
        self.__accepted_moves = 0
        self.__rejected_moves = 0
        self.__stored_e_L = []
        self.__stored_e_L2 = []
        self.__stored_w_L = []
        self.__stored_dln = {}
        self.__stored_force_HF = []
        self.__stored_force_PP = []
        self.__timer = dict(total=0.0, update=0.0, e_L=0.0, misc=0.0)

    # ---- properties (names as in the reference, jqmc_mcmc.py:260-447) --------------------------------
    @property
    def num_walkers(self):
        return self.__num_walkers

    @property
    def mcmc_counter(self):
        return self.__mcmc_counter

    @property
    def e_L(self):
        return np.array(self.__stored_e_L).reshape(-1, self.__num_walkers)

    @property
    def e_L2(self):
        return np.array(self.__stored_e_L2).reshape(-1, self.__num_walkers)

    @property
    def w_L(self):
        return np.array(self.__stored_w_L).reshape(-1, self.__num_walkers)

    @property
    def dln_Psi_dc(self):
        """{block name: array (steps, num_walkers, *block shape)} of d ln|Psi| / d parameter (jqmc_mcmc.py:300-330).  The
        samples are kept on the device while sampling (the SR solve consumes them there); this property copies them out."""
        return {k: self.__dln_block(k).cpu().numpy() for k in self.__stored_dln}

    def __dln_block(self, name):
        """One derivative block as a device tensor (steps, num_walkers, *block shape)."""
        v = self.__stored_dln[name]
        return torch.stack(v) if len(v) else torch.zeros((0, self.__num_walkers), dtype=torch.float64, device=self.engine.device)

    # walker state as NumPy arrays, like the reference's properties (the device tensors stay private)
    @property
    def latest_r_up_carts(self) -> np.ndarray:
        return self.__r_up.detach().cpu().numpy()

    @property
    def latest_r_dn_carts(self) -> np.ndarray:
        return self.__r_dn.detach().cpu().numpy()

    @property
    def jax_PRNG_key_list(self) -> np.ndarray:
        return self.__keys.detach().cpu().numpy()

    # ---- restart checkpoints (jqmc_mcmc.py:3912-4120; layout: jqmc_b200/checkpoint.py) ----------------------------------
    def save_to_hdf5(self, filepath: str) -> None:
        from .checkpoint import save_rank_checkpoint

        cfg = dict(mcmc_seed=int(self.__mcmc_seed), num_walkers=self.__num_walkers, num_mcmc_per_measurement=self.__nmpm, Dt=self.__Dt,
                   epsilon_AS=self.__epsilon_AS, comput_log_WF_param_deriv=self.__comput_log_WF_param_deriv, comput_e_L_param_deriv=False,
                   comput_position_deriv=self.__comput_position_deriv, random_discretized_mesh=self.__random_discretized_mesh, use_swct=self.__use_swct,
                   mcmc_counter=int(self.__mcmc_counter), accepted_moves=int(self.__accepted_moves),
                   rejected_moves=int(self.__rejected_moves), i_opt=int(self.__i_opt))  # fmt: skip
        obs = {"e_L": self.e_L, "e_L2": self.e_L2, "w_L": self.w_L, "param_grads": self.dln_Psi_dc}
        if self.__stored_force_HF:  # (M, 1, ...) per-step arrays in the reference are (M, nw, n_atom, 3) here as well
            obs["force_HF"] = np.array(self.__stored_force_HF)
            obs["force_PP"] = np.array(self.__stored_force_PP)
            obs["E_L_force_PP"] = self.e_L[..., None, None] * np.array(self.__stored_force_PP)
        save_rank_checkpoint(
            filepath, driver_type="MCMC", driver_config=cfg,
            rng_state={"jax_PRNG_key_list": self.jax_PRNG_key_list, "mpi_seed": int(self.__mpi_seed)},
            walker_state={"latest_r_up_carts": self.latest_r_up_carts, "latest_r_dn_carts": self.latest_r_dn_carts},
            observables=obs,
        )  # fmt: skip

    @classmethod
    def load_from_hdf5(cls, filepath: str, rank: int | None = None, engine: WalkerEngine | None = None) -> "MCMC":
        """Restore a sampler from a merged checkpoint (no ``__init__`` call, as in the reference): a following ``run``
        continues the chain exactly."""
        from .checkpoint import check_checkpoint_version, load_hamiltonian_from_checkpoint, load_rank_checkpoint

        if rank is None:
            rank, _ = _rank_size()
        check_checkpoint_version(filepath)
        data = load_rank_checkpoint(filepath, rank)
        cfg, rng, ws, obs = data["driver_config"], data["rng_state"], data["walker_state"], data["observables"]
        if cfg.get("comput_e_L_param_deriv"):
            raise NotImplementedError("e_L parameter derivatives (linear method) are outside the walker engine (SURVEY.md §8f)")
        H = load_hamiltonian_from_checkpoint(filepath)
        obj = cls.__new__(cls)
        obj._MCMC__comput_log_WF_param_deriv = bool(cfg.get("comput_log_WF_param_deriv", False))
        obj._MCMC__comput_position_deriv = bool(cfg.get("comput_position_deriv", False))
        obj.hamiltonian_data = H
        obj._MCMC__mcmc_seed = int(cfg["mcmc_seed"])
        obj._MCMC__num_walkers = int(cfg["num_walkers"])
        obj._MCMC__nmpm = int(cfg["num_mcmc_per_measurement"])
        obj._MCMC__Dt = float(cfg["Dt"])
        obj._MCMC__epsilon_AS = float(cfg["epsilon_AS"])
        obj._MCMC__random_discretized_mesh = bool(cfg.get("random_discretized_mesh", True))
        obj._MCMC__use_swct = bool(cfg.get("use_swct", True))
        obj._MCMC__mpi_seed = int(rng["mpi_seed"])
        obj.engine = engine if engine is not None else WalkerEngine(H)
        dev = obj.engine.device
        obj._MCMC__keys = torch.from_numpy(np.ascontiguousarray(rng["jax_PRNG_key_list"]).astype(np.uint32)).to(dev)
        obj._MCMC__r_up = torch.from_numpy(np.ascontiguousarray(ws["latest_r_up_carts"], dtype=np.float64)).to(dev)
        obj._MCMC__r_dn = torch.from_numpy(np.ascontiguousarray(ws["latest_r_dn_carts"], dtype=np.float64)).to(dev)
        obj._MCMC__forces = None
        if obj._MCMC__comput_position_deriv:
            from .forces import ForceEvaluator

            obj._MCMC__forces = ForceEvaluator(H, obj.engine)
        obj._MCMC__init_attributes()
        if obs.get("force_HF") is not None and np.size(obs["force_HF"]):
            obj._MCMC__stored_force_HF = [x for x in np.asarray(obs["force_HF"])]
            obj._MCMC__stored_force_PP = [x for x in np.asarray(obs["force_PP"])]
        obj._MCMC__mcmc_counter = int(cfg.get("mcmc_counter", 0))
        obj._MCMC__accepted_moves = int(cfg.get("accepted_moves", 0))
        obj._MCMC__rejected_moves = int(cfg.get("rejected_moves", 0))
        obj._MCMC__i_opt = int(cfg.get("i_opt", 0))
        for name, store in (("e_L", "_MCMC__stored_e_L"), ("e_L2", "_MCMC__stored_e_L2"), ("w_L", "_MCMC__stored_w_L")):
            a = obs.get(name)
            setattr(obj, store, [row for row in np.asarray(a)] if a is not None and np.size(a) else [])
        for name, a in (obs.get("param_grads") or {}).items():
            obj._MCMC__stored_dln[name] = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in np.asarray(a)]
        return obj

    @property
    def accepted_moves(self):
        return self.__accepted_moves

    @property
    def rejected_moves(self):
        return self.__rejected_moves

    @property
    def timer(self):
        return dict(self.__timer)

    # ---- sampling ------------------------------------------------------------------------------------
    def run(self, num_mcmc_steps: int = 0, max_time=86400) -> None:
        eng = self.engine
        toml_filename = "external_control_mcmc.toml"  # jqmc_mcmc.py:475
        write_control_file(toml_filename)
        t_start = time.perf_counter()
        G, Ginv = eng.geminal_inv_batched(self.__r_up, self.__r_dn)
        r_up, r_dn, keys = self.__r_up, self.__r_dn, self.__keys
        eps = self.__epsilon_AS
        for _ in range(num_mcmc_steps):
            acc, rej, r_up, r_dn, keys, Ginv, G = eng.update(
                r_up, r_dn, keys, self.__nmpm, self.__Dt, eps, Ginv, G, inplace=True
            )
            if self.__random_discretized_mesh:
                RTs = eng.generate_RTs(keys)
            else:
                RTs = None
            e_L = eng.e_L_fast(r_up, r_dn, RTs, Ginv)
            R_AS = eng.as_reg_fast(G, Ginv)
            if eps > 0:
                w_L = (R_AS / torch.clamp(R_AS, min=eps)) ** 2
            else:  # (R/max(R,0))^2 = 1, NaN when R_AS == 0 (0/0), as in jqmc_mcmc.py:743-747
                w_L = torch.where(R_AS > 0, torch.ones_like(R_AS), torch.full_like(R_AS, float("nan")))
            if self.__forces is not None:  # jqmc_mcmc.py:749-852
                f_hf, f_pp, _ = self.__forces.force_products(r_up, r_dn, RTs, self.__use_swct)
                force_pack = torch.stack([f_hf, f_pp]).cpu().numpy()
            # one device->host read per step (the reference does three: jqmc_mcmc.py:720, 739, 747)
            if self.__comput_log_WF_param_deriv:  # jqmc_mcmc.py:854-876
                for name, g in eng.grad_ln_psi_params_fast(r_up, r_dn, Ginv).items():
                    self.__stored_dln.setdefault(name, []).append(g)  # stays on the device: the SR solve consumes it there
            pack = torch.stack([e_L, w_L, acc.to(torch.float64), rej.to(torch.float64)]).cpu().numpy()
            # rank 0's stop decision (max_time / external stop flag), shared by all ranks; the interrupted step is not
            # counted and its observables are dropped, as in the reference (jqmc_mcmc.py:930-967, 978-985)
            if should_stop(t_start, max_time, toml_filename, eng.device):
                if self.__comput_log_WF_param_deriv:
                    for v in self.__stored_dln.values():
                        v.pop()
                break
            self.__stored_e_L.append(pack[0])
            self.__stored_e_L2.append(pack[0] ** 2)
            self.__stored_w_L.append(pack[1])
            if self.__forces is not None:
                self.__stored_force_HF.append(force_pack[0])
                self.__stored_force_PP.append(force_pack[1])
            self.__accepted_moves += int(pack[2].sum())
            self.__rejected_moves += int(pack[3].sum())
            self.__mcmc_counter += 1
        rank, _ = _rank_size()
        if rank == 0 and os.path.isfile(toml_filename):
            os.remove(toml_filename)
        self.__r_up, self.__r_dn, self.__keys = r_up, r_dn, keys
        self.__timer["total"] += time.perf_counter() - t_start

    def get_E(self, num_mcmc_warmup_steps: int = 50, num_mcmc_bin_blocks: int = 10):
        """(E_mean, E_std, Var_mean, Var_std): binned jackknife over walkers x blocks, reduced over ranks
        exactly as jqmc_mcmc.py:1064-1189 (sum reductions + two-pass variance)."""
        if self.mcmc_counter < num_mcmc_warmup_steps:
            raise ValueError("mcmc_counter should be larger than num_mcmc_warmup_steps")
        if self.mcmc_counter - num_mcmc_warmup_steps < num_mcmc_bin_blocks:
            raise ValueError("(mcmc_counter - num_mcmc_warmup_steps) should be larger than num_mcmc_bin_blocks.")
        e_L = self.e_L[num_mcmc_warmup_steps:]
        e_L2 = self.e_L2[num_mcmc_warmup_steps:]
        w_L = self.w_L[num_mcmc_warmup_steps:]
        return jackknife_E(w_L, e_L, e_L2, num_mcmc_bin_blocks, self.engine.device)


    def get_aF(self, num_mcmc_warmup_steps: int = 50, num_mcmc_bin_blocks: int = 10):
        """(force_mean, force_std) [n_atom, 3] in Hartree / bohr: Hellmann-Feynman + Pulay forces with the reference's jackknife
        (jqmc_mcmc.py:1191-1370).  Needs ``comput_position_deriv=True``."""
        if not self.__stored_force_HF:
            raise ValueError("no force samples stored: construct MCMC with comput_position_deriv=True")
        from .forces import jackknife_forces

        s = slice(num_mcmc_warmup_steps, None)
        return jackknife_forces(self.w_L[s], self.e_L[s], np.array(self.__stored_force_HF)[s], np.array(self.__stored_force_PP)[s],
                                num_mcmc_bin_blocks, self.engine.device)  # fmt: skip

    BLOCK_ORDER = ("j1_param", "j2_param", "j3_matrix", "lambda_matrix")  # jqmc/wavefunction.py:515-674

    def get_dln_WF(self, num_mcmc_warmup_steps: int = 50, chosen_param_index=None, blocks=None, lambda_projectors=None,
                   num_orb_projection=None, symmetrize: bool = True):  # fmt: skip
        """O_matrix (M, num_walkers, K): the stored derivatives after warm-up, blocks concatenated in the reference's order and
        flattened row-major, then -- as ``MCMC.get_dln_WF`` of the reference does (jqmc_mcmc.py:1372-1513) --

        * the lambda block projected when ``lambda_projectors = (L', R', S_up^-1/2, S_dn^-1/2)`` is given (:1424-1459: paired part
          O' = S_up^-1/2 O S_dn^-1/2, then O' - (I - L') O' (I - R'); unpaired columns S_up^-1/2 O), and
        * every block with an internal symmetry symmetrised (:1461-1481): the square part of ``j3_matrix`` when the current
          J3 matrix is symmetric, the paired part of ``lambda_matrix`` when the current lambda is (jqmc/wavefunction.py:170-244).

        ``blocks``: optional list of block names (default: all stored).  The algebra runs on the device."""
        if not self.__stored_dln:
            raise ValueError("no parameter derivatives stored: construct MCMC with comput_log_WF_param_deriv=True")
        O = self._dln_WF_device(num_mcmc_warmup_steps, blocks, lambda_projectors, num_orb_projection, symmetrize).cpu().numpy()
        return O if chosen_param_index is None else O[:, :, chosen_param_index]

    def _block_symmetry(self, name):
        """(symmetric?, number of leading columns that form the square part) of a matrix block of the CURRENT parameters."""
        wf = self.hamiltonian_data.wavefunction_data
        atol, rtol = 1.0e-8, 1.0e-6  # jqmc/_setting.py:74-75
        if name == "j3_matrix":
            j3 = wf.jastrow_data.jastrow_three_body_data
            m = np.asarray(j3.j_matrix)
            sq = m[:, :-1]
            return bool(sq.shape[0] == sq.shape[1] and np.allclose(sq, sq.T, atol=atol)), m.shape[1] - 1
        if name == "lambda_matrix":
            lam = np.asarray(wf.geminal_data.lambda_matrix)
            n = lam.shape[0]
            return bool(np.allclose(lam[:, :n], lam[:, :n].T, atol=atol, rtol=rtol)), n
        return False, 0

    def _dln_WF_device(self, num_mcmc_warmup_steps: int = 0, blocks=None, lambda_projectors=None, num_orb_projection=None,
                       symmetrize: bool = True):  # fmt: skip
        """The same matrix as a device tensor (no host round trip: the path of get_sr_direction / run_optimize)."""
        names = [n for n in self.BLOCK_ORDER if n in self.__stored_dln and (blocks is None or n in blocks)]
        dev = self.engine.device
        parts = []
        for n in names:
            a = self.__dln_block(n)[num_mcmc_warmup_steps:]
            if n == "lambda_matrix" and lambda_projectors is not None and num_orb_projection is not None:
                L, R, Su, Sd = (torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev) for x in lambda_projectors)
                eye = torch.eye(L.shape[0], dtype=torch.float64, device=dev)
                n_pc = R.shape[0]
                paired = Su @ a[..., :n_pc] @ Sd
                paired = paired - (eye - L) @ paired @ (eye - R)
                a = torch.cat([paired, Su @ a[..., n_pc:]], dim=-1)
            if symmetrize and a.dim() == 4:
                sym, n_sq = self._block_symmetry(n)
                if sym:
                    sq = a[..., :n_sq]
                    a = torch.cat([0.5 * (sq + sq.transpose(-1, -2)), a[..., n_sq:]], dim=-1)
            parts.append(a.reshape(a.shape[0], a.shape[1], -1))
        return torch.cat(parts, dim=2)

    def get_gF(self, num_mcmc_warmup_steps: int = 50, num_mcmc_bin_blocks: int = 10, chosen_param_index=None, blocks=None):
        """Generalised forces f_k = -2 (<e_L O_k> - <e_L><O_k>) with jackknife error bars over (bins x walkers) samples of all
        ranks (jqmc_mcmc.py:1516-1692: same binning, same reductions, two-pass standard deviation)."""
        w_L = self.w_L[num_mcmc_warmup_steps:]
        e_L = self.e_L[num_mcmc_warmup_steps:]
        O = self.get_dln_WF(num_mcmc_warmup_steps, chosen_param_index, blocks)
        return jackknife_gF(w_L, e_L, O, num_mcmc_bin_blocks, self.engine.device)

    @staticmethod
    def lambda_projectors(mo_coefficients_up, mo_coefficients_dn, overlap_up, overlap_dn, num_orb_projection: int):
        """(L', R', S_up^-1/2, S_dn^-1/2) of the lambda-subspace projection from the MO coefficients [n_mo, n_ao] and the AO
        overlap matrices (jqmc_mcmc.py:2712-2750): C' = S^1/2 C over the first ``num_orb_projection`` orbitals, L' = C'_up C'_up^T,
        R' = C'_dn C'_dn^T (orthogonal projectors in the S^-1/2-orthogonalised basis)."""
        out = []
        for C, S in ((mo_coefficients_up, overlap_up), (mo_coefficients_dn, overlap_dn)):
            S = 0.5 * (np.asarray(S, dtype=np.float64) + np.asarray(S, dtype=np.float64).T)
            ev, U = np.linalg.eigh(S)
            sq, isq = U @ np.diag(np.sqrt(ev)) @ U.T, U @ np.diag(1.0 / np.sqrt(ev)) @ U.T
            Cp = sq @ np.asarray(C, dtype=np.float64)[:num_orb_projection, :].T
            out.append((Cp @ Cp.T, isq))
        return out[0][0], out[1][0], out[0][1], out[1][1]


    def get_sr_direction(self, num_mcmc_warmup_steps: int = 0, epsilon: float = 1e-3, use_cg: bool = False, blocks=None,
                         cg_max_iter: int = 10000, cg_tol: float = 1e-10):  # fmt: skip
        """Natural-gradient direction theta (flattened over ``blocks`` in BLOCK_ORDER) from the stored samples of all ranks:
        the SR solve of jqmc_mcmc.py:2960-3330 on the device (jqmc_b200.sr).  Returns (theta[K] ndarray, info)."""
        from .sr import sr_natural_gradient

        dev = self.engine.device
        O = self._dln_WF_device(num_mcmc_warmup_steps, blocks)  # (symmetrised like get_dln_WF: f, S and theta inherit the symmetry)
        w = torch.from_numpy(self.w_L[num_mcmc_warmup_steps:]).to(dev)
        e = torch.from_numpy(self.e_L[num_mcmc_warmup_steps:]).to(dev)
        theta, info = sr_natural_gradient(w, e, O, epsilon=epsilon, use_cg=use_cg, cg_max_iter=cg_max_iter, cg_tol=cg_tol)
        return theta.cpu().numpy(), info

    def run_optimize(self, num_mcmc_steps: int = 100, num_opt_steps: int = 1, num_mcmc_warmup_steps: int = 0, delta: float = 1e-2,
                     epsilon: float = 1e-3, use_cg: bool = False, opt_J1_param: bool = True, opt_J2_param: bool = True,
                     opt_J3_param: bool = True, opt_lambda_param: bool = False, num_mcmc_bin_blocks: int = 5):  # fmt: skip
        """Plain SR optimisation loop (the ``sr`` branch of jqmc_mcmc.py:2368-3900 without its adaptive learning rate, SNR
        filters, lambda projection and block symmetrisation): sample, solve, c <- c + delta * theta, rebuild the device
        tables (qe_create) and continue from the current walkers.  Returns [(E, dE, max|f|)] per step."""
        import dataclasses

        if not self.__comput_log_WF_param_deriv:
            raise ValueError("run_optimize needs comput_log_WF_param_deriv=True")
        flags = dict(j1_param=opt_J1_param, j2_param=opt_J2_param, j3_matrix=opt_J3_param, lambda_matrix=opt_lambda_param)
        history = []
        for _ in range(num_opt_steps):
            self.__init_attributes()
            self.run(num_mcmc_steps)
            E, dE, _, _ = self.get_E(num_mcmc_warmup_steps, min(num_mcmc_bin_blocks, self.mcmc_counter - num_mcmc_warmup_steps))
            blocks = [n for n in self.BLOCK_ORDER if n in self.__stored_dln and flags[n]]
            theta, info = self.get_sr_direction(num_mcmc_warmup_steps, epsilon, use_cg, blocks)
            history.append((E, dE, float(info["f"].abs().max())))
            H = self.hamiltonian_data
            wf = H.wavefunction_data
            jd, gem = wf.jastrow_data, wf.geminal_data
            off = 0
            for n in blocks:
                shape = tuple(self.__stored_dln[n][0].shape[1:])
                size = int(np.prod(shape)) if shape else 1
                step = delta * theta[off : off + size].reshape(shape)
                off += size
                if n == "j1_param":
                    j1 = jd.jastrow_one_body_data
                    jd = dataclasses.replace(jd, jastrow_one_body_data=dataclasses.replace(j1, jastrow_1b_param=float(j1.jastrow_1b_param + step)))
                elif n == "j2_param":
                    j2 = jd.jastrow_two_body_data
                    jd = dataclasses.replace(jd, jastrow_two_body_data=dataclasses.replace(j2, jastrow_2b_param=float(j2.jastrow_2b_param + step)))
                elif n == "j3_matrix":
                    j3 = jd.jastrow_three_body_data
                    jd = dataclasses.replace(jd, jastrow_three_body_data=dataclasses.replace(j3, j_matrix=np.asarray(j3.j_matrix) + step))
                else:
                    gem = dataclasses.replace(gem, lambda_matrix=np.asarray(gem.lambda_matrix) + step)
            self.hamiltonian_data = dataclasses.replace(H, wavefunction_data=dataclasses.replace(wf, jastrow_data=jd, geminal_data=gem))
            self.engine = WalkerEngine(self.hamiltonian_data, Nv=self.engine.Nv, NN=self.engine.NN)  # tables change with the parameters
            self.__init_attributes()  # the stored samples belong to the old parameters
        return history


def jackknife_gF(w_L, e_L, O, num_bin_blocks, device=None):
    """(mean[K], std[K]) of -2 (<e_L O> - <e_L><O>) by binned jackknife; sums over ranks via torch.distributed."""

    def binned(x):  # (M, nw, ...) -> (bins*nw, ...)
        s = np.array([np.sum(a, axis=0) for a in np.array_split(x, num_bin_blocks, axis=0)])
        return s.reshape((-1,) + s.shape[2:])

    wb, web = binned(w_L), binned(w_L * e_L)
    wOb, weOb = binned(w_L[:, :, None] * O), binned((w_L * e_L)[:, :, None] * O)
    K = O.shape[2]
    g = _allreduce_sum(np.concatenate([[wb.sum(), web.sum(), wb.size], wOb.sum(axis=0), weOb.sum(axis=0)]), device)
    W, WE, M_total, WO, WEO = g[0], g[1], g[2], g[3 : 3 + K], g[3 + K :]
    den = (W - wb)[:, None]
    f = -2.0 * ((WEO - weOb) / den - ((WE - web) / (W - wb))[:, None] * ((WO - wOb) / den))
    mean = _allreduce_sum(f.sum(axis=0), device) / M_total
    var = _allreduce_sum(np.sum((f - mean) ** 2, axis=0), device) / M_total
    return mean, np.sqrt((M_total - 1) * var)


def jackknife_E(w_L, e_L, e_L2, num_bin_blocks, device=None):
    """Binned jackknife of the weighted energy and variance over (blocks x walkers) samples and all ranks."""

    def binned(x):
        return np.ravel([np.sum(a, axis=0) for a in np.array_split(x, num_bin_blocks, axis=0)])

    wb, web, we2b = binned(w_L), binned(w_L * e_L), binned(w_L * e_L2)
    g = _allreduce_sum([wb.sum(), web.sum(), we2b.sum(), wb.size], device)
    W, WE, WE2, M_total = g[0], g[1], g[2], g[3]
    E_jk = (WE - web) / (W - wb)
    E2_jk = (WE2 - we2b) / (W - wb)
    Var_jk = E2_jk - E_jk**2
    s = _allreduce_sum([E_jk.sum(), Var_jk.sum()], device)
    E_mean, Var_mean = s[0] / M_total, s[1] / M_total
    c = _allreduce_sum([np.sum((E_jk - E_mean) ** 2), np.sum((Var_jk - Var_mean) ** 2)], device)
    E_std = np.sqrt((M_total - 1) * c[0] / M_total)
    Var_std = np.sqrt((M_total - 1) * c[1] / M_total)
    return float(E_mean), float(E_std), float(Var_mean), float(Var_std)
