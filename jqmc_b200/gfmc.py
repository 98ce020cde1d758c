"""``GFMC_n`` (LRDMC, fixed number of projections per branching) and ``GFMC_t`` (LRDMC, fixed imaginary time per
branching) drivers on the walker engine: the host-side mirrors of ``jqmc.jqmc_gfmc.GFMC_n`` (jqmc/jqmc_gfmc.py:4191-6650)
and ``jqmc.jqmc_gfmc.GFMC_t`` (:173-2880).

Same constructor arguments, step loop, stored observables, on-the-fly ``E_scf`` update and ``get_E``
statistics as the reference; the per-step device work is done by ``WalkerEngine``:

    for each branching step (jqmc_gfmc.py:5774-6417)
        w = 1;  projection_n  -> qe_lrdmc_project     (nmpm lattice-regularised projections per walker)
        V_elements_n          -> qe_geminal_init + qe_lrdmc_velements   (e_L = V_diag + V_nondiag)
        weighted sums         -> qe_lrdmc_collect                                             (:5971-6051)
        reconfiguration       -> qe_lrdmc_pack + ONE all_gather + qe_lrdmc_reconfigure_packed (:6059-6318)
        A_inv refresh         -> qe_geminal_init                                             (:6319)
        E_scf update          -> host jackknife of the stored (w, e_L) history               (:6327-6383)

Differences from the reference, all behind the same observable results:

* the reference moves every array to the host and talks mpi4py (reduce -> rank 0, Exscan, Allgather,
  Alltoallv negotiation, Isend/Irecv of the migrating walkers); here the walker state never leaves the
  GPU: ONE ``all_gather`` of a packed record [5 sums | w | r_up | r_dn] per branching over NCCL/NVLink (persistent
  buffers), and every rank evaluates the identical comb redundantly (no negotiation round);
* the comb offset ``zeta`` is rank 0's ``np.random.random()`` after ``np.random.seed(mcmc_seed)`` at the top of
  ``run`` (:4669, :5948-5952); every rank replays that stream locally instead of receiving a broadcast;
* the per-step averages are all-reduced, so every rank stores them (the reference keeps them on rank 0 only).

``GFMC_t`` (jqmc_gfmc.py:646-2391) shares the reconfiguration and the statistics; its step is

    w = 1, tau_left = tau;  projection loop -> qe_lrdmc_project_tau  (until every walker has used up tau)
    weighted sums  sum w, sum w e_L, sum w e_L^2 (no division by V_diag - E_scf, :1929-1932) -> qe_lrdmc_collect
    reconfiguration, A_inv refresh as above; the mean projection count of the rank is stored per step (:2095, 2271)

Restart: ``save_to_hdf5`` / ``load_from_hdf5`` in jQMC's checkpoint layout (jqmc_b200/checkpoint.py).

Atomic forces (``comput_position_deriv=True``; jqmc_gfmc.py:5840-6051 GFMC_n, :1811-1998 GFMC_t): before every branching the
position derivatives of the lattice-regularised local energy and of ln|Psi| are taken on the device by finite differences
(jqmc_b200/forces.py), combined with the space-warp weights (``use_swct``) and the Pathak-Wagner factor (``epsilon_PW``)
into three weighted [n_atom, 3] sums per rank, all-reduced, and stored as (M, 1, n_atom, 3) histories; ``get_aF`` is the
reference's binned jackknife of -<F_HF> - 2 (<e_L F_PP> - <e_L><F_PP>) with the accumulated weights G_L.
"""

from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import rng_host
from .engine import WalkerEngine
from .mcmc import _dist, _rank_size, generate_init_electron_configurations, should_stop, write_control_file

# jqmc/_setting.py:59-61
GFMC_ON_THE_FLY_WARMUP_STEPS = 20
GFMC_ON_THE_FLY_COLLECT_STEPS = 10
GFMC_ON_THE_FLY_BIN_BLOCKS = 10


def compute_G_L(w_L: np.ndarray, num_gfmc_collect_steps: int) -> np.ndarray:
    """G_L[n] = prod_{k=n-c}^{n-1} w_L[k], n = c .. A-1  (jqmc/jqmc_gfmc.py:136-152); w_L has shape (A, x)."""
    w_L = np.asarray(w_L, dtype=np.float64)
    A, c = w_L.shape[0], int(num_gfmc_collect_steps)
    if A <= c:
        return np.zeros((0,) + w_L.shape[1:], dtype=np.float64)
    return np.stack([np.prod(w_L[n - c : n], axis=0) for n in range(c, A)])


def jackknife_E_scf(G_L, G_e_L, num_bin_blocks):
    """On-the-fly E_scf estimate (jqmc/jqmc_gfmc.py:6350-6365): binned jackknife of sum(G e)/sum(G)."""
    G_e_L_binned = np.array([np.sum(x) for x in np.array_split(np.asarray(G_e_L), num_bin_blocks)])
    G_binned = np.array([np.sum(x) for x in np.array_split(np.asarray(G_L), num_bin_blocks)])
    s_e, s_g = np.sum(G_e_L_binned), np.sum(G_binned)
    E_jk = [(s_e - G_e_L_binned[m]) / (s_g - G_binned[m]) for m in range(num_bin_blocks)]
    return float(np.average(E_jk)), float(np.sqrt(num_bin_blocks - 1) * np.std(E_jk))


class _GFMC:
    """What GFMC_n and GFMC_t share: walker initialisation, stored observables, the branching-step loop with its single
    device->host read per print interval, the reconfiguration and ``get_E`` (jqmc/jqmc_gfmc.py:4222-4634, 6500-6700;
    GFMC_t: :208-620, 2393-2587 -- the same code in the reference)."""

    def _setup(self, hamiltonian_data, num_walkers, num_gfmc_collect_steps, mcmc_seed, alat, random_discretized_mesh,
               non_local_move, comput_position_deriv, engine):  # fmt: skip
        self._hamiltonian_data = hamiltonian_data
        self._num_walkers = int(num_walkers)
        self._num_gfmc_collect_steps = int(num_gfmc_collect_steps)
        self._mcmc_seed = int(mcmc_seed)
        self._alat = float(alat)
        self._random_discretized_mesh = bool(random_discretized_mesh)
        self._non_local_move = non_local_move
        self._comput_position_deriv = bool(comput_position_deriv)
        rank, _ = _rank_size()
        self._mpi_seed = self._mcmc_seed * (rank + 1)
        self.engine = engine if engine is not None else WalkerEngine(hamiltonian_data)
        dev = self.engine.device
        keys = rng_host.split(rng_host.PRNGKey(self._mpi_seed), self._num_walkers)
        self._keys = torch.from_numpy(keys).to(dev)
        self._keys_init = np.array(keys, copy=True)  # jax_PRNG_key_list_init of the reference (checkpoint field)
        np.random.seed(self._mpi_seed % (2**32))
        gem = hamiltonian_data.wavefunction_data.geminal_data
        cp = hamiltonian_data.coulomb_potential_data
        r_up, r_dn, _, _ = generate_init_electron_configurations(
            gem.num_electron_up, gem.num_electron_dn, self._num_walkers, cp.effective_charges,
            hamiltonian_data.structure_data.positions,
        )  # fmt: skip
        self._r_up = torch.from_numpy(np.ascontiguousarray(r_up)).to(dev)
        self._r_dn = torch.from_numpy(np.ascontiguousarray(r_dn)).to(dev)
        self._init_forces()
        self._init_attributes()

    def _init_forces(self):
        self._forces = None
        if self._comput_position_deriv:
            from .forces import ForceEvaluator

            self._forces = ForceEvaluator(self._hamiltonian_data, self.engine, lattice=(self._alat, self._non_local_move))

    def _init_attributes(self):
        self._mcmc_counter = 0
        self._num_survived_walkers = 0
        self._num_killed_walkers = 0
        self._stored_w_L = np.zeros((0, 1))
        self._stored_e_L = np.zeros((0, 1))
        self._stored_e_L2 = np.zeros((0, 1))
        self._stored_average_projection_counter = np.zeros((0,))
        n_atoms = len(self._hamiltonian_data.structure_data.atomic_numbers)
        self._stored_force_HF = np.zeros((0, 1, n_atoms, 3))
        self._stored_force_PP = np.zeros((0, 1, n_atoms, 3))
        self._stored_E_L_force_PP = np.zeros((0, 1, n_atoms, 3))
        self._G_L = []
        self._G_e_L = []
        self._timer = dict(total=0.0)

    # ---- properties (jqmc_gfmc.py:4550-4634, 533-620) ------------------------------------------------
    @property
    def hamiltonian_data(self):
        return self._hamiltonian_data

    @property
    def num_gfmc_collect_steps(self):
        return self._num_gfmc_collect_steps

    @num_gfmc_collect_steps.setter
    def num_gfmc_collect_steps(self, n):
        self._num_gfmc_collect_steps = int(n)

    @property
    def mcmc_counter(self) -> int:
        return self._mcmc_counter - self._num_gfmc_collect_steps

    @property
    def num_walkers(self):
        return self._num_walkers

    @property
    def alat(self):
        return self._alat

    @property
    def w_L(self):
        return compute_G_L(self._stored_w_L, self._num_gfmc_collect_steps)

    @property
    def bare_w_L(self):
        return np.asarray(self._stored_w_L)

    @property
    def e_L(self):
        return np.asarray(self._stored_e_L)[self._num_gfmc_collect_steps :]

    @property
    def e_L2(self):
        return np.asarray(self._stored_e_L2)[self._num_gfmc_collect_steps :]

    @property
    def force_HF(self):
        return self._stored_force_HF[self._num_gfmc_collect_steps :]

    @property
    def force_PP(self):
        return self._stored_force_PP[self._num_gfmc_collect_steps :]

    @property
    def E_L_force_PP(self):
        return self._stored_E_L_force_PP[self._num_gfmc_collect_steps :]

    # walker state as NumPy arrays, like the reference's properties (the device tensors stay private: `_r_up`, `_r_dn`, `_keys`)
    @property
    def latest_r_up_carts(self) -> np.ndarray:
        return self._r_up.detach().cpu().numpy()

    @property
    def latest_r_dn_carts(self) -> np.ndarray:
        return self._r_dn.detach().cpu().numpy()

    @property
    def jax_PRNG_key_list(self) -> np.ndarray:
        return self._keys.detach().cpu().numpy()

    @property
    def comput_position_deriv(self) -> bool:
        return self._comput_position_deriv

    # ---- restart checkpoints (jqmc_gfmc.py:377-530 GFMC_t, :4390-4548 GFMC_n; layout: jqmc_b200/checkpoint.py) ------------
    def _driver_config(self) -> dict:
        raise NotImplementedError

    def save_to_hdf5(self, filepath: str) -> None:
        """This rank's state as a temporary per-rank file (merged into restart.h5 by checkpoint.merge_rank_checkpoints)."""
        from .checkpoint import save_rank_checkpoint

        cfg = self._driver_config()
        cfg.update(mcmc_counter=int(self._mcmc_counter), num_survived_walkers=int(self._num_survived_walkers),
                   num_killed_walkers=int(self._num_killed_walkers))  # fmt: skip
        obs = {"e_L": np.asarray(self._stored_e_L), "e_L2": np.asarray(self._stored_e_L2), "w_L": np.asarray(self._stored_w_L)}
        for name, arr in (("force_HF", self._stored_force_HF), ("force_PP", self._stored_force_PP),
                          ("E_L_force_PP", self._stored_E_L_force_PP)):  # fmt: skip
            obs[name] = arr if arr.size > 0 else np.empty(0)
        if isinstance(self, GFMC_t):
            obs["average_projection_counter"] = np.asarray(self._stored_average_projection_counter)
        else:
            obs["G_L"] = np.array(self._G_L) if len(self._G_L) else np.empty(0)
            obs["G_e_L"] = np.array(self._G_e_L) if len(self._G_e_L) else np.empty(0)
        save_rank_checkpoint(
            filepath, driver_type=type(self).__name__, driver_config=cfg,
            rng_state={"jax_PRNG_key_list": self.jax_PRNG_key_list, "jax_PRNG_key_list_init": np.asarray(self._keys_init),
                       "mpi_seed": int(self._mpi_seed)},
            walker_state={"latest_r_up_carts": self.latest_r_up_carts, "latest_r_dn_carts": self.latest_r_dn_carts},
            observables=obs,
        )  # fmt: skip

    @classmethod
    def load_from_hdf5(cls, filepath: str, rank: int | None = None, engine=None):
        """Restore a driver from a merged checkpoint without calling ``__init__`` (same contract as the reference): the
        Hamiltonian comes from the file's root, everything else from this rank's group; a following ``run`` continues the
        chain exactly (walkers, PRNG keys, stored observables, counters)."""
        from .checkpoint import check_checkpoint_version, load_hamiltonian_from_checkpoint, load_rank_checkpoint

        if rank is None:
            rank, _ = _rank_size()
        check_checkpoint_version(filepath)
        data = load_rank_checkpoint(filepath, rank)
        cfg, rng, ws, obs = data["driver_config"], data["rng_state"], data["walker_state"], data["observables"]
        H = load_hamiltonian_from_checkpoint(filepath)
        obj = cls.__new__(cls)
        obj._hamiltonian_data = H
        obj._num_walkers = int(cfg["num_walkers"])
        obj._num_gfmc_collect_steps = int(cfg["num_gfmc_collect_steps"])
        obj._mcmc_seed = int(cfg["mcmc_seed"])
        obj._alat = float(cfg["alat"])
        obj._random_discretized_mesh = bool(cfg.get("random_discretized_mesh", True))
        obj._non_local_move = cfg.get("non_local_move", "tmove")
        obj._comput_position_deriv = bool(cfg.get("comput_position_deriv", False))
        obj._restore_config(cfg)
        obj._mpi_seed = int(rng["mpi_seed"])
        obj.engine = engine if engine is not None else WalkerEngine(H)
        dev = obj.engine.device
        obj._keys = torch.from_numpy(np.ascontiguousarray(rng["jax_PRNG_key_list"]).astype(np.uint32)).to(dev)
        obj._keys_init = np.asarray(rng.get("jax_PRNG_key_list_init", rng["jax_PRNG_key_list"])).astype(np.uint32)
        obj._r_up = torch.from_numpy(np.ascontiguousarray(ws["latest_r_up_carts"], dtype=np.float64)).to(dev)
        obj._r_dn = torch.from_numpy(np.ascontiguousarray(ws["latest_r_dn_carts"], dtype=np.float64)).to(dev)
        obj._init_forces()
        obj._init_attributes()
        obj._mcmc_counter = int(cfg.get("mcmc_counter", 0))
        obj._num_survived_walkers = int(cfg.get("num_survived_walkers", 0))
        obj._num_killed_walkers = int(cfg.get("num_killed_walkers", 0))

        def get(name, default):
            a = obs.get(name)
            return default if a is None or np.size(a) == 0 else np.asarray(a)

        obj._stored_e_L = get("e_L", np.zeros((0, 1)))
        obj._stored_e_L2 = get("e_L2", np.zeros((0, 1)))
        obj._stored_w_L = get("w_L", np.zeros((0, 1)))
        obj._stored_force_HF = get("force_HF", obj._stored_force_HF)
        obj._stored_force_PP = get("force_PP", obj._stored_force_PP)
        obj._stored_E_L_force_PP = get("E_L_force_PP", obj._stored_E_L_force_PP)
        obj._stored_average_projection_counter = get("average_projection_counter", np.zeros((obj._mcmc_counter,)))
        g, ge = get("G_L", None), get("G_e_L", None)
        obj._G_L = [g[i] for i in range(g.shape[0])] if g is not None else []
        obj._G_e_L = [ge[i] for i in range(ge.shape[0])] if ge is not None else []
        return obj

    @property
    def num_survived_walkers(self):
        return self._num_survived_walkers

    @property
    def num_killed_walkers(self):
        return self._num_killed_walkers

    @property
    def timer(self):
        return dict(self._timer)

    # ---- walker reconfiguration on the device (jqmc_gfmc.py:6059-6321 == :2008-2277) -------------------
    def _reconfigure(self, w, r_up, r_dn, sums, zeta, rank, world):
        """One branching: this rank's record [sums | w | r_up | r_dn] goes through ONE all_gather (persistent buffers, NCCL
        over NVLink); every rank then sums the five sums in rank order, evaluates the identical comb and copies its new
        walkers out of the gathered records.  Returns (r_up, r_dn, A_inv, n_survived, sums over all ranks)."""
        eng = self.engine
        nw = self._num_walkers
        L = eng.lrdmc_record_len(nw)
        buf = getattr(self, "_exchange", None)
        if buf is None or buf[0].numel() != L or buf[1].numel() != world * L:
            rec = torch.empty(L, dtype=torch.float64, device=eng.device)
            buf = (rec, torch.empty(world * L, dtype=torch.float64, device=eng.device) if world > 1 else rec)
            self._exchange = buf
        rec, allrec = buf
        eng.lrdmc_pack(sums, w, r_up, r_dn, rec)
        d = _dist()
        if d is not None and world > 1:
            d.all_gather_into_tensor(allrec, rec)
        sums_all, r_up, r_dn, n_surv, _ = eng.lrdmc_reconfigure_packed(allrec, nw, world, rank, zeta)
        A_inv = eng.A_inv_n(r_up, r_dn)
        return r_up, r_dn, A_inv, n_surv, sums_all

    def _step(self, r_up, r_dn, keys, A_inv, zeta, rank, world):  # -> r_up, r_dn, keys, A_inv, sums, n_surv, extra, force sums
        raise NotImplementedError

    def _force_sums(self, r_up, r_dn, RTs, weight, e_L, world):
        """[3, n_atom, 3] weighted force sums over the walkers of all ranks (one small all-reduce), or None."""
        if self._forces is None:
            return None
        fs = self._forces.weighted_force_sums(r_up, r_dn, RTs, weight, e_L, self._use_swct, self._epsilon_PW)
        d = _dist()
        if d is not None and world > 1:
            d.all_reduce(fs)
        return fs

    def _after_interval(self, i, eq_steps, n_bins):
        pass

    def run(self, num_mcmc_steps: int = 50, max_time: int = 86400) -> None:
        rank, world = _rank_size()
        eng = self.engine
        toml_filename = "external_control_gfmc.toml"  # jqmc_gfmc.py:4685
        write_control_file(toml_filename)
        t_start = time.perf_counter()
        zeta_rng = np.random.RandomState(self._mcmc_seed % (2**32))  # rank 0's stream after np.random.seed(mpi_seed), :4669
        A_inv = eng.A_inv_n(self._r_up, self._r_dn)
        r_up, r_dn, keys = self._r_up, self._r_dn, self._keys
        base = self._mcmc_counter
        n_store = base + num_mcmc_steps
        self._stored_e_L = np.concatenate([self._stored_e_L, np.zeros((num_mcmc_steps, 1))])
        self._stored_e_L2 = np.concatenate([self._stored_e_L2, np.zeros((num_mcmc_steps, 1))])
        self._stored_w_L = np.concatenate([self._stored_w_L, np.zeros((num_mcmc_steps, 1))])
        self._stored_average_projection_counter = np.concatenate([self._stored_average_projection_counter, np.zeros(num_mcmc_steps)])
        if self._forces is not None:
            z = np.zeros((num_mcmc_steps,) + self._stored_force_HF.shape[1:])
            self._stored_force_HF = np.concatenate([self._stored_force_HF, z])
            self._stored_force_PP = np.concatenate([self._stored_force_PP, z])
            self._stored_E_L_force_PP = np.concatenate([self._stored_E_L_force_PP, z])
        mcmc_interval = int(np.maximum(num_mcmc_steps / 100, 1))
        eq_steps, n_collect, n_bins = GFMC_ON_THE_FLY_WARMUP_STEPS, GFMC_ON_THE_FLY_COLLECT_STEPS, GFMC_ON_THE_FLY_BIN_BLOCKS
        pending = []  # (step index, device sums, device n_survived, device extra) not yet read back
        done = 0

        def flush():
            # one device->host read for all pending steps (the reference reads every array every step)
            if not pending:
                return
            S = torch.stack([p[1] for p in pending]).cpu().numpy()
            NS = torch.stack([p[2].reshape(()) for p in pending]).cpu().numpy()
            X = torch.stack([p[3].reshape(()) for p in pending]).cpu().numpy() if pending[0][3] is not None else None
            F = torch.stack([p[4] for p in pending]).cpu().numpy() if pending[0][4] is not None else None
            for n, ((i, _, _, _, _), s, ns) in enumerate(zip(pending, S, NS)):
                nw_sum, w_sum, wq, weq, we2q = s
                self._stored_w_L[base + i, 0] = w_sum / nw_sum
                self._stored_e_L[base + i, 0] = weq / wq
                self._stored_e_L2[base + i, 0] = we2q / wq
                if X is not None:
                    self._stored_average_projection_counter[base + i] = X[n]
                if F is not None:  # averaged with the same denominator as e_L (:6031-6035 / :1980-1983)
                    self._stored_force_HF[base + i, 0] = F[n, 0] / wq
                    self._stored_force_PP[base + i, 0] = F[n, 1] / wq
                    self._stored_E_L_force_PP[base + i, 0] = F[n, 2] / wq
                self._num_survived_walkers += int(ns)
                self._num_killed_walkers += int(nw_sum) - int(ns)
                if i >= n_collect:  # :6336-6343
                    G = np.prod(self._stored_w_L[base + i - n_collect : base + i], axis=0)
                    self._G_L.append(G)
                    self._G_e_L.append(G * self._stored_e_L[base + i])
            pending.clear()

        for i in range(num_mcmc_steps):
            zeta = float(zeta_rng.random_sample())
            r_up, r_dn, keys, A_inv, sums, n_surv, extra, fsum = self._step(r_up, r_dn, keys, A_inv, zeta, rank, world)
            pending.append((i, sums, n_surv, extra, fsum))
            if (i + 1) % mcmc_interval == 0 and i > eq_steps:  # :6345-6378
                flush()
                self._after_interval(i, eq_steps, n_bins)
            # stop conditions (max_time, external stop flag): rank 0's decision, received by every rank so that all of them
            # leave at the same step (jqmc_gfmc.py:6386-6414).  Checked once per print interval -- the points where the host
            # reads the device anyway -- instead of every step; the interrupted step is not counted, as in the reference.
            if (i + 1) % mcmc_interval == 0 and should_stop(t_start, max_time, toml_filename, eng.device):
                break
            done += 1
        flush()
        if rank == 0 and os.path.isfile(toml_filename):
            os.remove(toml_filename)
        self._mcmc_counter += done
        ns = self._mcmc_counter
        self._stored_e_L = self._stored_e_L[:ns]
        self._stored_e_L2 = self._stored_e_L2[:ns]
        self._stored_w_L = self._stored_w_L[:ns]
        self._stored_average_projection_counter = self._stored_average_projection_counter[:ns]
        if self._forces is not None:
            self._stored_force_HF = self._stored_force_HF[:ns]
            self._stored_force_PP = self._stored_force_PP[:ns]
            self._stored_E_L_force_PP = self._stored_E_L_force_PP[:ns]
        assert n_store >= ns
        self._r_up, self._r_dn, self._keys = r_up, r_dn, keys
        self._timer["total"] += time.perf_counter() - t_start

    def get_E(self, num_mcmc_warmup_steps: int = 50, num_mcmc_bin_blocks: int = 10):
        """(E_mean, E_std, Var_mean, Var_std): binned jackknife with the accumulated weights G_L
        (jqmc/jqmc_gfmc.py:6500-6700, 2393-2587).  Every rank holds the same (M, 1) history, so no collective is needed."""
        if self.mcmc_counter < num_mcmc_warmup_steps:
            raise ValueError("mcmc_counter should be larger than num_mcmc_warmup_steps")
        if self.mcmc_counter - num_mcmc_warmup_steps < num_mcmc_bin_blocks:
            raise ValueError("(mcmc_counter - num_mcmc_warmup_steps) should be larger than num_mcmc_bin_blocks.")
        e_L = self.e_L[num_mcmc_warmup_steps:]
        e_L2 = self.e_L2[num_mcmc_warmup_steps:]
        w_L = self.w_L[num_mcmc_warmup_steps:]

        def binned(x):
            return np.ravel([np.sum(a, axis=0) for a in np.array_split(x, num_mcmc_bin_blocks, axis=0)])

        wb, web, we2b = binned(w_L), binned(w_L * e_L), binned(w_L * e_L2)
        M = wb.size
        E_jk = (np.sum(web) - web) / (np.sum(wb) - wb)
        E2_jk = (np.sum(we2b) - we2b) / (np.sum(wb) - wb)
        Var_jk = E2_jk - E_jk**2
        E_mean = np.sum(E_jk) / M
        E_std = np.sqrt((M - 1) * np.sum((E_jk - E_mean) ** 2) / M)
        Var_mean = np.sum(Var_jk) / M
        Var_std = np.sqrt((M - 1) * np.sum((Var_jk - Var_mean) ** 2) / M)
        return float(E_mean), float(E_std), float(Var_mean), float(Var_std)

    def get_aF(self, num_mcmc_warmup_steps: int = 50, num_mcmc_bin_blocks: int = 10):
        """(force_mean, force_std) [n_atom, 3]: binned jackknife of -<F_HF> - 2 (<e_L F_PP> - <e_L><F_PP>) over the stored
        histories weighted with G_L (jqmc/jqmc_gfmc.py:6694-6990, 2587-2880; the reference scatters the bins over ranks only to
        share the arithmetic -- every rank holds the same history here and evaluates all bins)."""
        if self._stored_force_HF.shape[0] == 0:
            raise ValueError("no force samples stored: construct the driver with comput_position_deriv=True")
        s = slice(num_mcmc_warmup_steps, None)
        w_L, e_L = self.w_L[s], self.e_L[s]
        f_hf, f_pp, ef_pp = self.force_HF[s], self.force_PP[s], self.E_L_force_PP[s]

        def binned(x):  # (M, 1, ...) -> (bins, ...)
            b = np.array([np.sum(a, axis=0) for a in np.array_split(x, num_mcmc_bin_blocks, axis=0)])
            return b.reshape((b.shape[0] * b.shape[1],) + b.shape[2:])

        wb, web = binned(w_L), binned(w_L * e_L)
        whf, wpp, wef = (binned(w_L[..., None, None] * f) for f in (f_hf, f_pp, ef_pp))
        M = wb.size
        den = (np.sum(wb) - wb)[:, None, None]
        F_hf = -(np.sum(whf, axis=0) - whf) / den
        F_pl = -2.0 * ((np.sum(wef, axis=0) - wef) / den - ((np.sum(web) - web)[:, None, None] / den) * ((np.sum(wpp, axis=0) - wpp) / den))
        F = F_hf + F_pl
        mean = np.sum(F, axis=0) / M
        var = np.sum((F - mean) ** 2, axis=0) / M
        return mean, np.sqrt((M - 1) * var)


class GFMC_n(_GFMC):
    """LRDMC sampler with ``num_mcmc_per_measurement`` projections per branching (see module docstring).  Public surface
    follows jqmc.jqmc_gfmc.GFMC_n."""

    def __init__(
        self,
        hamiltonian_data=None,
        num_walkers: int = 40,
        num_mcmc_per_measurement: int = 16,
        num_gfmc_collect_steps: int = 5,
        mcmc_seed: int = 34467,
        E_scf: float = 0.0,
        alat: float = 0.1,
        random_discretized_mesh: bool = True,
        non_local_move: str = "tmove",
        comput_position_deriv: bool = False,
        epsilon_PW: float = 0.0,
        use_swct: bool = False,
        engine=None,
    ) -> None:
        self._nmpm = int(num_mcmc_per_measurement)
        self._E_scf = float(E_scf)
        self._epsilon_PW, self._use_swct = float(epsilon_PW), bool(use_swct)
        self._setup(hamiltonian_data, num_walkers, num_gfmc_collect_steps, mcmc_seed, alat, random_discretized_mesh, non_local_move,
                    comput_position_deriv, engine)  # fmt: skip

    @property
    def E_scf(self):
        return self._E_scf

    def _driver_config(self) -> dict:  # jqmc_gfmc.py:4400-4416
        return dict(mcmc_seed=self._mcmc_seed, num_walkers=self._num_walkers, num_mcmc_per_measurement=self._nmpm,
                    num_gfmc_collect_steps=self._num_gfmc_collect_steps, E_scf=float(self._E_scf), alat=self._alat,
                    random_discretized_mesh=self._random_discretized_mesh, non_local_move=self._non_local_move,
                    comput_position_deriv=self._comput_position_deriv, epsilon_PW=float(self._epsilon_PW), use_swct=self._use_swct)  # fmt: skip

    def _restore_config(self, cfg) -> None:
        self._nmpm = int(cfg["num_mcmc_per_measurement"])
        self._E_scf = float(cfg["E_scf"])
        self._epsilon_PW = float(cfg.get("epsilon_PW", 0.0))
        self._use_swct = bool(cfg.get("use_swct", False))

    # ---- one branching step on the device (no host synchronisation) ----------------------------------
    def _step(self, r_up, r_dn, keys, A_inv, zeta, rank, world):
        eng = self.engine
        nw = self._num_walkers
        w = torch.ones(nw, dtype=torch.float64, device=eng.device)
        w, r_up, r_dn, A_inv, keys, RTs, _, _ = eng.projection_n(
            w, r_up, r_dn, A_inv, keys, self._E_scf, self._nmpm, self._random_discretized_mesh, self._non_local_move,
            self._alat, inplace=True,
        )  # fmt: skip
        V_diag, V_nondiag = eng.V_elements_n(r_up, r_dn, RTs, self._non_local_move, self._alat)
        sums = eng.lrdmc_collect(w, V_diag, V_nondiag, self._E_scf)
        fsum = self._force_sums(r_up, r_dn, RTs, w / (V_diag - self._E_scf), V_diag + V_nondiag, world)  # :5977-6019
        r_up, r_dn, A_inv, n_surv, sums = self._reconfigure(w, r_up, r_dn, sums, zeta, rank, world)
        return r_up, r_dn, keys, A_inv, sums, n_surv, None, fsum

    def _after_interval(self, i, eq_steps, n_bins):  # on-the-fly E_scf, jqmc_gfmc.py:6345-6378
        n_warm = int(np.minimum(eq_steps, i - eq_steps))
        G_eq, G_e_eq = np.array(self._G_L[n_warm:]), np.array(self._G_e_L[n_warm:])
        if len(G_eq) >= n_bins:
            self._E_scf, _ = jackknife_E_scf(G_eq, G_e_eq, n_bins)


class GFMC_t(_GFMC):
    """LRDMC sampler that propagates every walker for the imaginary time ``tau`` per branching (continuous-time
    projections).  Public surface follows jqmc.jqmc_gfmc.GFMC_t (jqmc/jqmc_gfmc.py:173-2880)."""

    def __init__(
        self,
        hamiltonian_data=None,
        num_walkers: int = 40,
        num_gfmc_collect_steps: int = 5,
        mcmc_seed: int = 34467,
        tau: float = 0.1,
        alat: float = 0.1,
        random_discretized_mesh: bool = True,
        non_local_move: str = "tmove",
        comput_position_deriv: bool = False,
        epsilon_PW: float = 0.0,
        use_swct: bool = False,
        engine=None,
    ) -> None:
        self._tau = float(tau)
        self._epsilon_PW, self._use_swct = float(epsilon_PW), bool(use_swct)
        self._setup(hamiltonian_data, num_walkers, num_gfmc_collect_steps, mcmc_seed, alat, random_discretized_mesh, non_local_move,
                    comput_position_deriv, engine)  # fmt: skip

    @property
    def tau(self):
        return self._tau

    def _driver_config(self) -> dict:  # jqmc_gfmc.py:387-401
        return dict(mcmc_seed=self._mcmc_seed, num_walkers=self._num_walkers, num_gfmc_collect_steps=self._num_gfmc_collect_steps,
                    tau=float(self._tau), alat=self._alat, random_discretized_mesh=self._random_discretized_mesh,
                    non_local_move=self._non_local_move, comput_position_deriv=self._comput_position_deriv,
                    epsilon_PW=float(self._epsilon_PW), use_swct=self._use_swct)  # fmt: skip

    def _restore_config(self, cfg) -> None:
        self._tau = float(cfg["tau"])
        self._epsilon_PW = float(cfg.get("epsilon_PW", 0.0))
        self._use_swct = bool(cfg.get("use_swct", False))

    @property
    def average_projection_counter(self):
        return np.asarray(self._stored_average_projection_counter)

    def _step(self, r_up, r_dn, keys, A_inv, zeta, rank, world):
        eng = self.engine
        nw = self._num_walkers
        w = torch.ones(nw, dtype=torch.float64, device=eng.device)  # weights, time and counter restart every step (:1700-1703)
        e_L, pc, w, r_up, r_dn, A_inv, keys, RTs = eng.projection_t(
            w, r_up, r_dn, A_inv, keys, self._tau, self._random_discretized_mesh, self._non_local_move, self._alat, inplace=True
        )
        sums = eng.lrdmc_collect_t(w, e_L)
        fsum = self._force_sums(r_up, r_dn, RTs, w, e_L, world)  # weights without the V_diag - E_scf division (:1963-1968)
        r_up, r_dn, A_inv, n_surv, sums = self._reconfigure(w, r_up, r_dn, sums, zeta, rank, world)
        return r_up, r_dn, keys, A_inv, sums, n_surv, pc.to(torch.float64).mean(), fsum  # rank-local mean, as the reference (:2095)
