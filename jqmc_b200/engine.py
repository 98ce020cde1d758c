"""Host-side mirror of jQMC's walker-batched callables on top of the C ABI (include/jqmc_b200.h).

``WalkerEngine`` flattens a ``Hamiltonian_data`` (the reference's own object or the mirrors in
``jqmc_b200.data``) into device tables once, and then exposes the seams the reference's drivers call
every step, with the same names, argument order and array layouts:

=============================  ======================================================================
reference callable             engine method
=============================  ======================================================================
``_geminal_inv_batched``       ``geminal_inv_batched(r_up, r_dn) -> (G, Ginv)``  jqmc_mcmc.py:4264
``_jit_vmap_update``           ``update(r_up, r_dn, keys, nmpm, Dt, epsilon_AS, Ginv, G)``  :4728
``_jit_vmap_generate_RTs``     ``generate_RTs(keys)``  :4738
``_jit_vmap_e_L_fast``         ``e_L_fast(r_up, r_dn, RTs, Ginv)``  :4736
``_jit_vmap_as_reg_fast``      ``as_reg_fast(G, Ginv)``  :4739
``evaluate_ln_wavefunction``   ``ln_wavefunction(r_up, r_dn)``  wavefunction.py:677
=============================  ======================================================================

Arrays are fp64 / uint32 CUDA tensors with the walker axis leading.  Inputs may be torch CUDA tensors
(zero copy), any object with ``__dlpack__`` such as a ``jax.Array`` (zero copy via DLPack), or NumPy
arrays (copied host->device).  Outputs are torch CUDA tensors (``jax.dlpack.from_dlpack`` /
``.cpu().numpy()`` on the caller's side).  PyTorch is only used for device memory and streams.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .data import is_cart, is_mos


def _i32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _p_i32(a):
    return a.ctypes.data_as(_lib.i32p)


def _p_f64(a):
    return a.ctypes.data_as(_lib.f64p)


def _basis_desc(orb, keep):
    """qe_basis_desc from AOs_sphe_data / AOs_cart_data / MOs_data (attribute duck typing)."""
    d = _lib.qe_basis_desc()
    aos = orb.aos_data if is_mos(orb) else orb
    d.cartesian = 1 if is_cart(aos) else 0
    d.n_ao = int(aos.num_ao)
    d.n_prim = int(aos.num_ao_prim)
    arrs = dict(
        nucleus_index=_i32(aos.nucleus_index),
        angular_momentums=_i32(aos.angular_momentums),
        orbital_indices=_i32(aos.orbital_indices),
        exponents=_f64(aos.exponents),
        coefficients=_f64(aos.coefficients),
    )
    if len(arrs["exponents"]) != d.n_prim or len(arrs["orbital_indices"]) != d.n_prim:
        raise ValueError("AO tables: exponents/coefficients/orbital_indices must have num_ao_prim entries")
    if len(arrs["nucleus_index"]) != d.n_ao or len(arrs["angular_momentums"]) != d.n_ao:
        raise ValueError("AO tables: nucleus_index/angular_momentums must have num_ao entries")
    if d.cartesian:
        arrs["px"] = _i32(aos.polynominal_order_x)
        arrs["py"] = _i32(aos.polynominal_order_y)
        arrs["pz"] = _i32(aos.polynominal_order_z)
        d.polynominal_order_x, d.polynominal_order_y, d.polynominal_order_z = (_p_i32(arrs[k]) for k in ("px", "py", "pz"))
    else:
        arrs["m"] = _i32(aos.magnetic_quantum_numbers)
        d.magnetic_quantum_numbers = _p_i32(arrs["m"])
    d.nucleus_index = _p_i32(arrs["nucleus_index"])
    d.angular_momentums = _p_i32(arrs["angular_momentums"])
    d.orbital_indices = _p_i32(arrs["orbital_indices"])
    d.exponents = _p_f64(arrs["exponents"])
    d.coefficients = _p_f64(arrs["coefficients"])
    if is_mos(orb):
        c = _f64(orb.mo_coefficients)
        if c.shape != (int(orb.num_mo), d.n_ao):
            raise ValueError(f"mo_coefficients shape {c.shape} != ({orb.num_mo}, {d.n_ao})")
        arrs["C"] = c
        d.n_mo = int(orb.num_mo)
        d.mo_coefficients = _p_f64(c)
    else:
        d.n_mo = 0
    keep.append(arrs)
    return d


class WalkerEngine:
    """Device-resident tables of one Hamiltonian + the batched step kernels (see module docstring)."""

    def __init__(self, hamiltonian_data, Nv: int = 6, NN: int = 1, device=None, precision: str = "full"):
        if not torch.cuda.is_available():
            raise RuntimeError("jqmc_b200.WalkerEngine needs a CUDA device (there is no CPU path)")
        self._lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        H = hamiltonian_data
        wf = H.wavefunction_data
        gem = wf.geminal_data
        jas = wf.jastrow_data
        cp = H.coulomb_potential_data
        st = H.structure_data
        if getattr(st, "pbc_flag", False):
            raise NotImplementedError("periodic systems are not supported (as in the reference: trexio_wrapper.py:112-119)")
        if getattr(jas, "jastrow_nn_data", None) is not None:
            raise NotImplementedError("NN Jastrow is out of scope of the walker engine")
        keep = []
        d = _lib.qe_system_desc()
        pos = _f64(st.positions).reshape(-1, 3)
        z_all = _f64(st.atomic_numbers)
        zeff = z_all - _f64(cp.z_cores) if cp.ecp_flag else z_all
        d.n_atom = pos.shape[0]
        d.positions = _p_f64(pos)
        d.effective_charges = _p_f64(zeff)
        d.n_up, d.n_dn = int(gem.num_electron_up), int(gem.num_electron_dn)
        d.orb_up = _basis_desc(gem.orb_data_up_spin, keep)
        d.orb_dn = _basis_desc(gem.orb_data_dn_spin, keep)
        lam = _f64(gem.lambda_matrix)
        n_orb_up = d.orb_up.n_mo or d.orb_up.n_ao
        n_orb_dn = d.orb_dn.n_mo or d.orb_dn.n_ao
        if lam.shape != (n_orb_up, n_orb_dn + d.n_up - d.n_dn):
            raise ValueError(f"lambda_matrix shape {lam.shape} != ({n_orb_up}, {n_orb_dn + d.n_up - d.n_dn})")
        d.lambda_matrix = _p_f64(lam)
        j1 = getattr(jas, "jastrow_one_body_data", None)
        j2 = getattr(jas, "jastrow_two_body_data", None)
        j3 = getattr(jas, "jastrow_three_body_data", None)
        if j1 is not None:
            d.j1_type = {"exp": 1, "pade": 2}[j1.jastrow_1b_type]
            d.j1_param = float(j1.jastrow_1b_param)
            core = _f64(j1.core_electrons)
            zj = _f64(j1.structure_data.atomic_numbers)
            keep.append((core, zj))
            d.j1_core_electrons, d.j1_atomic_numbers = _p_f64(core), _p_f64(zj)
        if j2 is not None:
            d.j2_type = {"pade": 1, "exp": 2}[j2.jastrow_2b_type]
            d.j2_param = float(j2.jastrow_2b_param)
        if j3 is not None:
            d.j3_flag = 1
            d.j3_orb = _basis_desc(j3.orb_data, keep)
            jm = _f64(j3.j_matrix)
            keep.append(jm)
            d.j_matrix = _p_f64(jm)
        if cp.ecp_flag:
            d.ecp_flag = 1
            d.n_ecp = int(cp.num_ecps)
            e = dict(
                nuc=_i32(cp.nucleus_index), l=_i32(cp.ang_moms), z=_f64(cp.exponents), c=_f64(cp.coefficients),
                p=_i32(cp.powers), lmax=_i32(cp.max_ang_mom_plus_1),
            )  # fmt: skip
            keep.append(e)
            d.ecp_nucleus_index, d.ecp_ang_moms, d.ecp_powers = _p_i32(e["nuc"]), _p_i32(e["l"]), _p_i32(e["p"])
            d.ecp_exponents, d.ecp_coefficients = _p_f64(e["z"]), _p_f64(e["c"])
            d.ecp_max_ang_mom_plus_1 = _p_i32(e["lmax"])
        d.Nv, d.NN = int(Nv), int(NN)
        if precision not in ("full", "mixed"):  # jqmc/_precision.py: the two modes of the reference
            raise ValueError("precision must be 'full' or 'mixed'")
        d.precision = 1 if precision == "mixed" else 0
        self.precision = precision
        keep += [pos, zeff, lam]
        self.n_up, self.n_dn, self.n_atom = d.n_up, d.n_dn, d.n_atom
        self.n_e = d.n_up + d.n_dn
        self.Nv, self.NN = int(Nv), int(NN)
        self.ecp_flag = bool(cp.ecp_flag)
        self.n_orb = n_orb_up
        self._n_ao = int(d.orb_up.n_ao)
        self._has_j1, self._has_j2 = bool(d.j1_type), bool(d.j2_type)
        self._n_ao_j3 = int(d.j3_orb.n_ao) if d.j3_flag else 0
        self._n_orb_j3 = int(d.j3_orb.n_mo or d.j3_orb.n_ao) if d.j3_flag else 0
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.qe_create(C.byref(d), C.byref(h)), "qe_create")
        self._h = h
        del keep

    @classmethod
    def from_hdf5(cls, path: str, group: str = "", Nv: int = 6, NN: int = 1, device=None, precision: str = "full"):
        """Engine built by the library itself from jQMC's ``hamiltonian_data.h5`` (tree at the root) or from a restart
        checkpoint (``group="hamiltonian_data"``): the native input path, no Python data model involved (qe_create_from_hdf5)."""
        if not torch.cuda.is_available():
            raise RuntimeError("jqmc_b200.WalkerEngine needs a CUDA device (there is no CPU path)")
        if precision not in ("full", "mixed"):
            raise ValueError("precision must be 'full' or 'mixed'")
        self = cls.__new__(cls)
        self._lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        counts, checks = hdf5_summary(path, group)
        (self.n_atom, self.n_up, self.n_dn, self._n_ao, _, n_mo, _, _, ecp, j1, j2, j3, n_ao_j3, n_mo_j3) = counts[:14]
        self.n_e = self.n_up + self.n_dn
        self.Nv, self.NN, self.precision = int(Nv), int(NN), precision
        self.ecp_flag = bool(ecp)
        self.n_orb = n_mo or self._n_ao
        self._has_j1, self._has_j2 = bool(j1), bool(j2)
        self._n_ao_j3 = n_ao_j3 if j3 else 0
        self._n_orb_j3 = (n_mo_j3 or n_ao_j3) if j3 else 0
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self._lib.qe_create_from_hdf5(path.encode(), group.encode(), int(Nv), int(NN), 1 if precision == "mixed" else 0, C.byref(h))
        _lib.check(rc, "qe_create_from_hdf5")
        self._h = h
        return self

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and self._lib is not None:  # (at interpreter shutdown the ctypes handle may already be gone)
            self._lib.qe_destroy(h)

    # ---- tensor plumbing -------------------------------------------------------------------------
    def _dev(self, x, dtype=torch.float64):
        if isinstance(x, torch.Tensor):
            t = x
        elif isinstance(x, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(x))
        elif hasattr(x, "__dlpack__"):
            t = torch.from_dlpack(x)
        else:
            t = torch.as_tensor(np.asarray(x))
        if t.dtype != dtype:
            if dtype == torch.uint32 and t.dtype in (torch.int32, torch.int64):
                t = t.to(torch.int64).to(torch.uint32) if t.dtype == torch.int64 else t.view(torch.uint32)
            else:
                t = t.to(dtype)
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        return t.contiguous()

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def clone_for(self, hamiltonian_data):
        """A second engine with this one's settings for another Hamiltonian (the displaced geometries of the force terms)."""
        return WalkerEngine(hamiltonian_data, Nv=self.Nv, NN=self.NN, device=self.device, precision=self.precision)

    def _walkers(self, r_up, r_dn):
        r_up = self._dev(r_up)
        nw = r_up.shape[0]
        if r_up.shape != (nw, self.n_up, 3):
            raise ValueError(f"r_up_carts shape {tuple(r_up.shape)} != (nw, {self.n_up}, 3)")
        r_dn = self._dev(r_dn) if self.n_dn else torch.zeros((nw, 0, 3), dtype=torch.float64, device=self.device)
        if r_dn.shape != (nw, self.n_dn, 3):
            raise ValueError(f"r_dn_carts shape {tuple(r_dn.shape)} != ({nw}, {self.n_dn}, 3)")
        return r_up, r_dn, nw

    def _mat(self, x, nw, name):
        x = self._dev(x)
        if x.shape != (nw, self.n_up, self.n_up):
            raise ValueError(f"{name} shape {tuple(x.shape)} != ({nw}, {self.n_up}, {self.n_up})")
        return x

    def _keys(self, keys, nw):
        keys = self._dev(keys, torch.uint32)
        if keys.shape != (nw, 2):
            raise ValueError(f"keys shape {tuple(keys.shape)} != ({nw}, 2)")
        return keys

    # ---- seams -----------------------------------------------------------------------------------
    def geminal_inv_batched(self, r_up, r_dn):
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        G = torch.empty((nw, self.n_up, self.n_up), dtype=torch.float64, device=self.device)
        Ginv = torch.empty_like(G)
        rc = self._lib.qe_geminal_init(self._h, nw, self._ptr(r_up), self._ptr(r_dn), self._ptr(G), self._ptr(Ginv), self._stream())
        _lib.check(rc, "qe_geminal_init")
        return G, Ginv

    def update(self, r_up, r_dn, keys, num_mcmc_per_measurement, Dt, epsilon_AS, Ginv, G, inplace=False):
        """``nmpm`` Metropolis proposals per walker.  Returns
        ``(accepted[nw], rejected[nw], r_up, r_dn, keys, Ginv, G)`` like the reference.  With
        ``inplace=True`` the caller's CUDA tensors are updated directly (no clone)."""
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        keys = self._keys(keys, nw)
        G, Ginv = self._mat(G, nw, "geminal"), self._mat(Ginv, nw, "geminal_inv")
        if not inplace:
            r_up, r_dn, keys, G, Ginv = (t.clone() for t in (r_up, r_dn, keys, G, Ginv))
        acc = torch.empty(nw, dtype=torch.int32, device=self.device)
        rej = torch.empty(nw, dtype=torch.int32, device=self.device)
        rc = self._lib.qe_mcmc_update(
            self._h, nw, self._ptr(r_up), self._ptr(r_dn), self._ptr(keys), self._ptr(G), self._ptr(Ginv),
            int(num_mcmc_per_measurement), float(Dt), float(epsilon_AS), self._ptr(acc), self._ptr(rej), self._stream(),
        )  # fmt: skip
        _lib.check(rc, "qe_mcmc_update")
        return acc, rej, r_up, r_dn, keys, Ginv, G

    def generate_RTs(self, keys):
        keys = self._dev(keys, torch.uint32)
        nw = keys.shape[0]
        keys = self._keys(keys, nw)
        RT = torch.empty((nw, 3, 3), dtype=torch.float64, device=self.device)
        _lib.check(self._lib.qe_rotation(self._h, nw, self._ptr(keys), self._ptr(RT), self._stream()), "qe_rotation")
        return RT

    def e_L_fast(self, r_up, r_dn, RTs, Ginv, return_parts=False):
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        Ginv = self._mat(Ginv, nw, "geminal_inverse")
        if RTs is None:
            RTs = torch.eye(3, dtype=torch.float64, device=self.device).repeat(nw, 1, 1)
        RTs = self._dev(RTs)
        if RTs.shape != (nw, 3, 3):
            raise ValueError(f"RTs shape {tuple(RTs.shape)} != ({nw}, 3, 3)")
        e_L = torch.empty(nw, dtype=torch.float64, device=self.device)
        T = V = None
        if return_parts:
            T = torch.empty((nw, self.n_e), dtype=torch.float64, device=self.device)
            V = torch.empty((nw, 4), dtype=torch.float64, device=self.device)
        rc = self._lib.qe_local_energy(
            self._h, nw, self._ptr(r_up), self._ptr(r_dn), self._ptr(RTs), self._ptr(Ginv), self._ptr(e_L),
            self._ptr(T), self._ptr(V), self._stream(),
        )  # fmt: skip
        _lib.check(rc, "qe_local_energy")
        return (e_L, T, V) if return_parts else e_L

    def nearest_nuclei(self, r_up, r_dn):
        """nn_index[nw, n_up + n_dn, NN] (int32): the nuclei the non-local ECP of every electron is taken around (qe_nearest_nuclei)."""
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        out = torch.empty((nw, self.n_e, self.NN), dtype=torch.int32, device=self.device)
        rc = self._lib.qe_nearest_nuclei(self._h, nw, self._ptr(r_up), self._ptr(r_dn), self._ptr(out), self._stream())
        _lib.check(rc, "qe_nearest_nuclei")
        return out

    def e_L_frozen(self, r_up, r_dn, RTs, Ginv, nn_index):
        """e_L with the nearest-nucleus assignment of the non-local ECP given by the caller (qe_local_energy_frozen): the
        building block of the finite-difference position derivatives (jqmc_b200/forces.py)."""
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        Ginv = self._mat(Ginv, nw, "A_old_inv")
        if RTs is None:
            RTs = torch.eye(3, dtype=torch.float64, device=self.device).expand(nw, 3, 3).contiguous()
        RTs = self._dev(RTs)
        e_L = torch.empty(nw, dtype=torch.float64, device=self.device)
        rc = self._lib.qe_local_energy_frozen(
            self._h, nw, self._ptr(r_up), self._ptr(r_dn), self._ptr(RTs), self._ptr(Ginv),
            self._ptr(nn_index) if nn_index is not None else None, self._ptr(e_L), self._stream(),
        )  # fmt: skip
        _lib.check(rc, "qe_local_energy_frozen")
        return e_L

    def as_reg_fast(self, G, Ginv):
        G = self._dev(G)
        nw = G.shape[0]
        G, Ginv = self._mat(G, nw, "geminal"), self._mat(Ginv, nw, "geminal_inv")
        out = torch.empty(nw, dtype=torch.float64, device=self.device)
        _lib.check(self._lib.qe_as_factor(self._h, nw, self._ptr(G), self._ptr(Ginv), self._ptr(out), self._stream()), "qe_as_factor")
        return out

    def ln_wavefunction(self, r_up, r_dn):
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        ln = torch.empty(nw, dtype=torch.float64, device=self.device)
        sg = torch.empty(nw, dtype=torch.float64, device=self.device)
        rc = self._lib.qe_ln_wavefunction(self._h, nw, self._ptr(r_up), self._ptr(r_dn), self._ptr(ln), self._ptr(sg), self._stream())
        _lib.check(rc, "qe_ln_wavefunction")
        return ln, sg

    def eval_orbitals(self, which, layer, r):
        """(5, n_orb, n_pts): value, d/dx, d/dy, d/dz, laplacian.  which: 'up'|'dn'|'j3'; layer: 'ao'|'orb'."""
        r = self._dev(r).reshape(-1, 3)
        wi = {"up": 0, "dn": 1, "j3": 2}[which]
        li = {"ao": 0, "orb": 1}[layer]
        n_orb = self._n_out(wi, li)
        out = torch.zeros((5, n_orb, r.shape[0]), dtype=torch.float64, device=self.device)
        rc = self._lib.qe_eval_orbitals(self._h, wi, li, r.shape[0], self._ptr(r), self._ptr(out), self._stream())
        _lib.check(rc, "qe_eval_orbitals")
        return out

    def _n_out(self, wi, li):
        if wi == 2:
            if self._n_ao_j3 == 0:
                raise ValueError("eval_orbitals: this Hamiltonian has no three-body Jastrow orbitals")
            return self._n_ao_j3 if li == 0 else self._n_orb_j3
        return self._n_ao if li == 0 else self.n_orb

    def move_ratios(self, r_up, r_dn, Ginv, elec, r_new, det=True, jas=True):
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        Ginv = self._mat(Ginv, nw, "A_old_inv")
        elec = _i32(elec)
        n_moves = len(elec)
        r_new = self._dev(r_new)
        if r_new.shape != (nw, n_moves, 3):
            raise ValueError(f"r_new shape {tuple(r_new.shape)} != ({nw}, {n_moves}, 3)")
        dr = torch.empty((nw, n_moves), dtype=torch.float64, device=self.device) if det else None
        jr = torch.empty((nw, n_moves), dtype=torch.float64, device=self.device) if jas else None
        rc = self._lib.qe_move_ratios(
            self._h, nw, self._ptr(r_up), self._ptr(r_dn), self._ptr(Ginv), n_moves, _p_i32(elec), self._ptr(r_new),
            self._ptr(dr), self._ptr(jr), self._stream(),
        )  # fmt: skip
        _lib.check(rc, "qe_move_ratios")
        torch.cuda.current_stream(self.device).synchronize()  # elec is a host buffer
        return dr, jr

    def grad_ln_psi_params_fast(self, r_up, r_dn, Ginv):
        """``_jit_vmap_grad_ln_psi_params_fast`` (jqmc/jqmc_mcmc.py:4748): per-walker d ln|Psi| / d parameter for every
        variational block of this Hamiltonian, as a dict keyed like the reference's blocks:
        ``j1_param`` [nw], ``j2_param`` [nw], ``j3_matrix`` [nw, n, n+1], ``lambda_matrix`` [nw, n_orb, n_orb + n_up - n_dn]."""
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        Ginv = self._mat(Ginv, nw, "A_old_inv")
        out = {}
        mk = lambda *shape: torch.empty((nw,) + shape, dtype=torch.float64, device=self.device)  # noqa: E731
        if self._has_j1:
            out["j1_param"] = mk()
        if self._has_j2:
            out["j2_param"] = mk()
        if self._n_orb_j3:
            out["j3_matrix"] = mk(self._n_orb_j3, self._n_orb_j3 + 1)
        out["lambda_matrix"] = mk(self.n_orb, self.n_orb + self.n_up - self.n_dn)
        rc = self._lib.qe_dln_wf(
            self._h, nw, self._ptr(r_up), self._ptr(r_dn), self._ptr(Ginv), self._ptr(out.get("j1_param")), self._ptr(out.get("j2_param")),
            self._ptr(out.get("j3_matrix")), self._ptr(out["lambda_matrix"]), self._stream(),
        )  # fmt: skip
        _lib.check(rc, "qe_dln_wf")
        return out

    def set_fused(self, on: bool):
        """Fused (one kernel, shared-memory resident) vs staged (kernel chain) local energy; same results."""
        _lib.check(self._lib.qe_set_fused(self._h, 1 if on else 0), "qe_set_fused")

    def set_path(self, general: bool):
        """Force the general ("wide") kernel family (True) or let the engine choose (False); see qe_set_path."""
        _lib.check(self._lib.qe_set_path(self._h, 1 if general else 0), "qe_set_path")

    def set_gemm_reference(self, on: bool):
        """Debugging aid: plain DFMA GEMM instead of the fp64 tensor-core kernel in the general path."""
        _lib.check(self._lib.qe_set_gemm_reference(self._h, 1 if on else 0), "qe_set_gemm_reference")

    def set_wide_slice(self, walkers: int):
        """General family: walkers per slice of one call (0 = automatic from the free device memory); process-wide."""
        _lib.check(self._lib.qe_set_wide_slice(self._h, int(walkers)), "qe_set_wide_slice")

    def set_walkers_per_cta(self, wpc: int):
        """Walkers per CTA of the fused walker kernel (0 = automatic)."""
        _lib.check(self._lib.qe_set_walkers_per_cta(self._h, int(wpc)), "qe_set_walkers_per_cta")

    def set_walker_warps(self, warps: int):
        """Warps per CTA of the fused walker kernel (0 = default 16 = one CTA per SM; 8 = two CTAs per SM; 4 = four)."""
        _lib.check(self._lib.qe_set_walker_warps(self._h, int(warps)), "qe_set_walker_warps")

    # ---- LRDMC (GFMC_n) seams: jqmc/jqmc_gfmc.py:4716, 5656-5663 ------------------------------------
    _NLM = {"tmove": 0, "dltmove": 1}

    def A_inv_n(self, r_up, r_dn):
        """``_jit_vmap_A_inv_n``: fresh inverse of the geminal matrix per walker."""
        return self.geminal_inv_batched(r_up, r_dn)[1]

    def projection_n(self, w_L, r_up, r_dn, A_old_inv, keys, E_scf, num_mcmc_per_measurement, random_discretized_mesh,
                     non_local_move, alat, inplace=False):
        """``_jit_vmap_projection_n``: returns (w, r_up, r_dn, A_inv, keys, RT, V_diag, V_nondiag)."""
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        keys = self._keys(keys, nw)
        Ginv = self._mat(A_old_inv, nw, "A_old_inv")
        w = self._dev(w_L)
        if w.shape != (nw,):
            raise ValueError(f"w_L shape {tuple(w.shape)} != ({nw},)")
        if non_local_move not in self._NLM:
            raise NotImplementedError(f"non_local_move = {non_local_move} is not yet implemented.")
        if not inplace:
            w, r_up, r_dn, keys, Ginv = (t.clone() for t in (w, r_up, r_dn, keys, Ginv))
        RT = torch.empty((nw, 3, 3), dtype=torch.float64, device=self.device)
        Vd = torch.empty(nw, dtype=torch.float64, device=self.device)
        Vn = torch.empty(nw, dtype=torch.float64, device=self.device)
        rc = self._lib.qe_lrdmc_project(
            self._h, nw, self._ptr(w), self._ptr(r_up), self._ptr(r_dn), self._ptr(Ginv), self._ptr(keys), float(E_scf),
            int(num_mcmc_per_measurement), 1 if random_discretized_mesh else 0, self._NLM[non_local_move], float(alat),
            self._ptr(RT), self._ptr(Vd), self._ptr(Vn), self._stream(),
        )  # fmt: skip
        _lib.check(rc, "qe_lrdmc_project")
        return w, r_up, r_dn, Ginv, keys, RT, Vd, Vn

    def projection_t(self, w_L, r_up, r_dn, A_old_inv, keys, tau, random_discretized_mesh, non_local_move, alat, inplace=False):
        """``GFMC_t``'s ``_run_projection_loop`` (jqmc/jqmc_gfmc.py:1539-1570; body ``_projection_t_core`` :724-1110) with
        ``projection_counter = 0`` and ``tau_left = tau`` on entry, as the driver sets them every branching step (:1700-1703).
        Returns (e_L, projection_counter, w, r_up, r_dn, A_inv, keys, RT)."""
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        keys = self._keys(keys, nw)
        Ginv = self._mat(A_old_inv, nw, "A_old_inv")
        w = self._dev(w_L)
        if w.shape != (nw,):
            raise ValueError(f"w_L shape {tuple(w.shape)} != ({nw},)")
        if non_local_move not in self._NLM:
            raise NotImplementedError(f"non_local_move = {non_local_move} is not yet implemented.")
        if not inplace:
            w, r_up, r_dn, keys, Ginv = (t.clone() for t in (w, r_up, r_dn, keys, Ginv))
        RT = torch.empty((nw, 3, 3), dtype=torch.float64, device=self.device)
        e_L = torch.empty(nw, dtype=torch.float64, device=self.device)
        pc = torch.zeros(nw, dtype=torch.int32, device=self.device)
        rc = self._lib.qe_lrdmc_project_tau(
            self._h, nw, self._ptr(w), self._ptr(r_up), self._ptr(r_dn), self._ptr(Ginv), self._ptr(keys), float(tau),
            1 if random_discretized_mesh else 0, self._NLM[non_local_move], float(alat), self._ptr(pc), self._ptr(e_L),
            self._ptr(RT), self._stream(),
        )  # fmt: skip
        _lib.check(rc, "qe_lrdmc_project_tau")
        return e_L, pc, w, r_up, r_dn, Ginv, keys, RT

    def V_elements_n(self, r_up, r_dn, RTs, non_local_move, alat, A_inv=None, nn_index=None):
        """``_jit_vmap_V_elements_n``: (V_diag, V_nondiag); the inverse is rebuilt unless ``A_inv`` is given.  With
        ``nn_index`` (``nearest_nuclei`` of a base point) the non-local ECP keeps that nucleus assignment
        (qe_lrdmc_velements_frozen: the function whose finite differences are the LRDMC force terms)."""
        r_up, r_dn, nw = self._walkers(r_up, r_dn)
        if non_local_move not in self._NLM:
            raise NotImplementedError(f"non_local_move = {non_local_move} is not yet implemented.")
        Ginv = self.A_inv_n(r_up, r_dn) if A_inv is None else self._mat(A_inv, nw, "A_inv")
        RTs = self._dev(RTs)
        if RTs.shape != (nw, 3, 3):
            raise ValueError(f"RTs shape {tuple(RTs.shape)} != ({nw}, 3, 3)")
        Vd = torch.empty(nw, dtype=torch.float64, device=self.device)
        Vn = torch.empty(nw, dtype=torch.float64, device=self.device)
        if nn_index is not None:
            nn = nn_index
            if not isinstance(nn, torch.Tensor) or nn.dtype != torch.int32 or tuple(nn.shape) != (nw, self.n_up + self.n_dn, self.NN):
                raise ValueError(f"nn_index must be int32[{nw}, {self.n_up + self.n_dn}, {self.NN}]")
            rc = self._lib.qe_lrdmc_velements_frozen(
                self._h, nw, self._ptr(r_up), self._ptr(r_dn), self._ptr(RTs), self._ptr(Ginv), self._ptr(nn.contiguous()),
                self._NLM[non_local_move], float(alat), self._ptr(Vd), self._ptr(Vn), self._stream(),
            )  # fmt: skip
            _lib.check(rc, "qe_lrdmc_velements_frozen")
            return Vd, Vn
        rc = self._lib.qe_lrdmc_velements(
            self._h, nw, self._ptr(r_up), self._ptr(r_dn), self._ptr(RTs), self._ptr(Ginv), self._NLM[non_local_move],
            float(alat), self._ptr(Vd), self._ptr(Vn), self._stream(),
        )  # fmt: skip
        _lib.check(rc, "qe_lrdmc_velements")
        return Vd, Vn

    # ---- LRDMC per-step statistics and reconfiguration: jqmc/jqmc_gfmc.py:5955-6321 ---------------
    def lrdmc_collect(self, w, V_diag, V_nondiag, E_scf):
        """Device vector [nw, sum w, sum w/(V_diag-E), sum w/(V_diag-E) e_L, sum w/(V_diag-E) e_L^2] of this rank."""
        w = self._dev(w)
        nw = w.shape[0]
        Vd, Vn = self._dev(V_diag), self._dev(V_nondiag)
        if Vd.shape != (nw,) or Vn.shape != (nw,):
            raise ValueError(f"V_diag / V_nondiag shapes {tuple(Vd.shape)} / {tuple(Vn.shape)} != ({nw},)")
        out = torch.empty(5, dtype=torch.float64, device=self.device)
        rc = self._lib.qe_lrdmc_collect(self._h, nw, self._ptr(w), self._ptr(Vd), self._ptr(Vn), float(E_scf), self._ptr(out), self._stream())
        _lib.check(rc, "qe_lrdmc_collect")
        return out

    def lrdmc_collect_t(self, w, e_L):
        """GFMC_t per-step sums (jqmc/jqmc_gfmc.py:1929-1932): device vector [nw, sum w, sum w, sum w e_L, sum w e_L^2]."""
        w, e = self._dev(w), self._dev(e_L)
        nw = w.shape[0]
        if e.shape != (nw,):
            raise ValueError(f"e_L shape {tuple(e.shape)} != ({nw},)")
        out = torch.empty(5, dtype=torch.float64, device=self.device)
        rc = self._lib.qe_lrdmc_collect(self._h, nw, self._ptr(w), None, self._ptr(e), 0.0, self._ptr(out), self._stream())
        _lib.check(rc, "qe_lrdmc_collect")
        return out

    def lrdmc_branch(self, w_all, num_walkers, zeta):
        """Comb selection over the all-gathered weights ``w_all[world*nw]``: returns
        ``(chosen_all int32[world*nw], n_survived int32[1])`` (device), identical on every rank."""
        w_all = self._dev(w_all)
        nw = int(num_walkers)
        if w_all.ndim != 1 or w_all.shape[0] % nw:
            raise ValueError(f"w_all shape {tuple(w_all.shape)} is not (world*{nw},)")
        world = w_all.shape[0] // nw
        chosen = torch.empty(world * nw, dtype=torch.int32, device=self.device)
        nsurv = torch.empty(1, dtype=torch.int32, device=self.device)
        rc = self._lib.qe_lrdmc_branch(self._h, nw, world, self._ptr(w_all), float(zeta), self._ptr(chosen), self._ptr(nsurv), self._stream())
        _lib.check(rc, "qe_lrdmc_branch")
        return chosen, nsurv

    def gather_walkers(self, chosen_local, src_r_up, src_r_dn):
        """New local walkers ``src[chosen_local[i]]`` from the all-gathered coordinates."""
        chosen_local = self._dev(chosen_local, torch.int32)
        nw = chosen_local.shape[0]
        src_up, src_dn = self._dev(src_r_up), self._dev(src_r_dn)
        if src_up.shape[1:] != (self.n_up, 3) or src_dn.shape[1:] != (self.n_dn, 3) or src_up.shape[0] != src_dn.shape[0]:
            raise ValueError(f"gathered coordinate shapes {tuple(src_up.shape)} / {tuple(src_dn.shape)} are inconsistent")
        dst_up = torch.empty((nw, self.n_up, 3), dtype=torch.float64, device=self.device)
        dst_dn = torch.empty((nw, self.n_dn, 3), dtype=torch.float64, device=self.device)
        rc = self._lib.qe_gather_walkers(
            self._h, nw, self._ptr(chosen_local), self._ptr(src_up), self._ptr(src_dn), self._ptr(dst_up), self._ptr(dst_dn), self._stream()
        )
        _lib.check(rc, "qe_gather_walkers")
        return dst_up, dst_dn

    def lrdmc_record_len(self, nw: int) -> int:
        """Doubles per rank record of the packed reconfiguration exchange (qe_lrdmc_record_len)."""
        return int(self._lib.qe_lrdmc_record_len(self._h, int(nw)))

    def lrdmc_pack(self, sums5, w, r_up, r_dn, record):
        """record <- [sums5 | pad | w | r_up | r_dn] (this rank's contribution to the one all_gather of a branching step)."""
        nw = w.shape[0]
        rc = self._lib.qe_lrdmc_pack(self._h, nw, self._ptr(sums5), self._ptr(w), self._ptr(r_up), self._ptr(r_dn), self._ptr(record),
                                     self._stream())  # fmt: skip
        _lib.check(rc, "qe_lrdmc_pack")
        return record

    def lrdmc_reconfigure_packed(self, records, nw: int, world: int, rank: int, zeta: float):
        """Comb + new walkers of this rank from the all-gathered records (rank order): returns (sums5 summed over ranks,
        r_up, r_dn, n_survived, chosen_all)."""
        chosen = torch.empty(world * nw, dtype=torch.int32, device=self.device)
        nsurv = torch.empty(1, dtype=torch.int32, device=self.device)
        sums = torch.empty(5, dtype=torch.float64, device=self.device)
        dst_up = torch.empty((nw, self.n_up, 3), dtype=torch.float64, device=self.device)
        dst_dn = torch.empty((nw, self.n_dn, 3), dtype=torch.float64, device=self.device)
        rc = self._lib.qe_lrdmc_reconfigure_packed(
            self._h, int(nw), int(world), int(rank), self._ptr(records), float(zeta), self._ptr(chosen), self._ptr(nsurv), self._ptr(sums),
            self._ptr(dst_up), self._ptr(dst_dn), self._stream(),
        )  # fmt: skip
        _lib.check(rc, "qe_lrdmc_reconfigure_packed")
        return sums, dst_up, dst_dn, nsurv, chosen

    def launch_count(self) -> int:
        return int(self._lib.qe_launch_count(self._h))

    PHASE_NAMES = ("stage+VGL", "P1 weights/draws", "P2 mesh", "P3a FN split", "P3b sums", "P3d select", "P4 VGL",
                   "P4 reduce", "P4 Sherman-Morrison", "P4 commit", "write-back", "-")  # fmt: skip

    def phase_clocks(self, enable: bool = True):
        """Per-phase cycle sums of the fused walker kernel since the last call (diagnostic); (re)arms or disarms them."""
        out = (C.c_int64 * 12)()
        _lib.check(self._lib.qe_phase_clocks(self._h, 1 if enable else 0, out), "qe_phase_clocks")
        return dict(zip(self.PHASE_NAMES, [int(v) for v in out]))

    def profile(self, enable: bool):
        """Switch per-kernel CUDA-event timing on/off (clears previous records)."""
        _lib.check(self._lib.qe_profile(self._h, 1 if enable else 0), "qe_profile")

    def profile_read(self):
        """{kernel name: (total_ms, launches)} since profile(True)."""
        out = {}
        for i in range(self._lib.qe_profile_kernels()):
            ms, n = C.c_double(), C.c_int64()
            _lib.check(self._lib.qe_profile_read(self._h, i, C.byref(ms), C.byref(n)), "qe_profile_read")
            if n.value:
                out[self._lib.qe_profile_name(i).decode()] = (ms.value, n.value)
        return out


def hdf5_summary(path: str, group: str = ""):
    """(counts[16], checks[8]) of the Hamiltonian tree in an HDF5 file as the LIBRARY's reader sees it (qe_hdf5_summary; no GPU
    needed): n_atom, n_up, n_dn, n_ao, n_prim, n_mo, cartesian, n_ecp, ecp_flag, j1_type, j2_type, j3_flag, n_ao_j3, n_mo_j3."""
    counts = (C.c_int64 * 16)()
    checks = (C.c_double * 8)()
    _lib.check(_lib.load().qe_hdf5_summary(path.encode(), group.encode(), counts, checks), "qe_hdf5_summary")
    return [int(x) for x in counts], [float(x) for x in checks]


def hdf5_read_walkers(path: str, rank: int, n_up: int, n_dn: int):
    """(r_up, r_dn, keys) NumPy arrays of one rank of a restart checkpoint, read by the library (qe_hdf5_read_walkers)."""
    lib = _lib.load()
    nw = C.c_int()
    _lib.check(lib.qe_hdf5_read_walkers(path.encode(), rank, 0, n_up, n_dn, C.byref(nw), None, None, None), "qe_hdf5_read_walkers")
    r_up = np.empty((nw.value, n_up, 3))
    r_dn = np.empty((nw.value, n_dn, 3))
    keys = np.empty((nw.value, 2), dtype=np.uint32)
    rc = lib.qe_hdf5_read_walkers(path.encode(), rank, nw.value, n_up, n_dn, C.byref(nw), r_up.ctypes.data_as(_lib.f64p),
                                  r_dn.ctypes.data_as(_lib.f64p), keys.ctypes.data_as(C.POINTER(C.c_uint32)))  # fmt: skip
    _lib.check(rc, "qe_hdf5_read_walkers")
    return r_up, r_dn, keys


def measure_fp64_peak(iters: int = 20000) -> float:
    """Achieved DFMA TFLOP/s of this GPU (roofline denominator for the fp64 kernels)."""
    out = C.c_double()
    _lib.check(_lib.load().qe_measure_fp64_peak(int(iters), C.byref(out)), "qe_measure_fp64_peak")
    return float(out.value)
