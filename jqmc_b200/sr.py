"""Stochastic-reconfiguration natural gradient on the device (second half).

Mirror of the SR solve inside ``MCMC.run_optimize`` of the reference (jqmc/jqmc_mcmc.py:2960-3330), which runs in NumPy + mpi4py
on the host.  Given the samples of this rank -- weights w_i, local energies e_i and derivatives O_ik = d ln|Psi| / d c_k
(``MCMC.get_dln_WF``; produced on the GPU by ``qe_dln_wf``) -- it forms

    X_ki = sqrt(w_i) (O_ik - <O_k>) / sqrt(sum w),      F_i = -2 sqrt(w_i) (e_i - <e>) / sqrt(sum w)
    S = X X^T  (scale-invariant: X_k <- X_k / sqrt(diag S_k)),      f = X F
    theta = (S + eps I)^-1 f / sqrt(diag S)

with the reference's branches: primal direct solve, primal conjugate gradient (one all-reduce of a K-vector per iteration),
and -- when there are more parameters than samples -- the dual ("push-through") form (X^T X + eps I) y = F, theta = X y.
Sums over samples of all ranks go through ``torch.distributed`` all-reduce / all-gather (NCCL on GPUs, gloo in the CPU
tests); the dense products are plain library GEMMs on whatever device the inputs live on (fp64).  Parameters whose
diag S is below ``min_S_diag_abs`` are frozen (theta = 0), as in the reference (:3080-3100).
"""

from __future__ import annotations

import torch


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


_COLLECTIVE_BYTES = [0]  # payload of the collectives issued since the last sr_natural_gradient call started (reported in info)


def _allreduce(t):
    d = _dist()
    _COLLECTIVE_BYTES[0] += t.numel() * t.element_size()
    if d is not None:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


def _allgather_cat(t):
    """Concatenate a 1-D tensor over ranks (all ranks must pass the same length)."""
    d = _dist()
    if d is None:
        return t
    _COLLECTIVE_BYTES[0] += t.numel() * t.element_size() * d.get_world_size()
    parts = [torch.empty_like(t) for _ in range(d.get_world_size())]
    d.all_gather(parts, t.contiguous())
    return torch.cat(parts)


def conjugate_gradient(b, apply_A, x0, max_iter, tol):
    """Plain CG on A x = b (jqmc/jqmc_mcmc.py: _conjugate_gradient_numpy): returns (x, relative residual, iterations)."""
    x = x0.clone()
    r = b - apply_A(x)
    p = r.clone()
    rs = torch.dot(r, r)
    b_norm = torch.sqrt(torch.dot(b, b)).clamp_min(1e-300)
    it = 0
    for it in range(1, max_iter + 1):
        Ap = apply_A(p)
        alpha = rs / torch.dot(p, Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        rs_new = torch.dot(r, r)
        if torch.sqrt(rs_new) / b_norm < tol:
            rs = rs_new
            break
        p = r + (rs_new / rs) * p
        rs = rs_new
    return x, float(torch.sqrt(rs) / b_norm), it


def sr_natural_gradient(w, e_L, O, epsilon=1e-3, use_cg=False, cg_max_iter=10000, cg_tol=1e-10, min_S_diag_abs=1e-15,
                        force_dual=None):  # fmt: skip
    """theta[K] (see module docstring).  w[n], e_L[n], O[n, K]: this rank's samples (any leading shape is flattened; torch
    tensors on one device, fp64).  Returns (theta, info) with info = dict(f, diag_S, method, cg_iterations, frozen)."""
    _COLLECTIVE_BYTES[0] = 0
    w = w.reshape(-1).to(torch.float64)
    e_L = e_L.reshape(-1).to(torch.float64)
    O = O.reshape(w.numel(), -1).to(torch.float64)
    K, n_local = O.shape[1], w.numel()
    head = _allreduce(torch.cat([w.sum().reshape(1), torch.dot(w, e_L).reshape(1), torch.tensor([float(n_local)], dtype=torch.float64, device=w.device)]))
    W, e_bar, n_total = head[0], head[1] / head[0], int(round(float(head[2])))
    O_bar = _allreduce(w @ O) / W
    sw = torch.sqrt(w)
    X = ((O - O_bar) * sw[:, None] / torch.sqrt(W)).T.contiguous()  # [K, n_local]
    F = -2.0 * sw * (e_L - e_bar) / torch.sqrt(W)
    f = _allreduce(X @ F)  # generalised force <-2 (e_L - E)(O - <O>)>
    diag_S = _allreduce((X * X).sum(dim=1))
    frozen = ~(torch.isfinite(diag_S) & (diag_S > min_S_diag_abs))
    diag_S = torch.where(frozen, torch.full_like(diag_S, min_S_diag_abs), diag_S)
    X = X / torch.sqrt(diag_S)[:, None]
    dual = (K >= n_total) if force_dual is None else bool(force_dual)
    iters = 0
    if not dual:
        XF = _allreduce(X @ F)
        if not use_cg:
            S = _allreduce(X @ X.T)
            S.diagonal().add_(epsilon)
            theta = torch.linalg.solve(S, XF)
            method = "primal-direct"
        else:
            theta, _, iters = conjugate_gradient(XF, lambda v: _allreduce(X @ (X.T @ v)) + epsilon * v, torch.zeros_like(XF), cg_max_iter, cg_tol)
            method = "primal-cg"
    else:
        # push-through identity: (X X^T + eps)^-1 X F = X (X^T X + eps)^-1 F over the samples of ALL ranks
        F_all = _allgather_cat(F)
        d = _dist()
        rank = d.get_rank() if d is not None else 0
        lo = rank * n_local

        def apply_dual(v):  # v over all samples; X^T X v = sum_ranks X_r^T ... needs X of every rank: go through parameter space
            u = _allreduce(X @ v[lo : lo + n_local])  # [K] = X_all v
            return _allgather_cat(X.T @ u) + epsilon * v

        if use_cg:
            y, _, iters = conjugate_gradient(F_all, apply_dual, torch.zeros_like(F_all), cg_max_iter, cg_tol)
            method = "dual-cg"
        else:
            X_all_T = _allgather_cat(X.T.reshape(-1)).reshape(-1, K)  # [n_total, K]
            A = X_all_T @ X_all_T.T
            A.diagonal().add_(epsilon)
            y = torch.linalg.solve(A, F_all)
            method = "dual-direct"
        theta = _allreduce(X @ y[lo : lo + n_local])
    theta = theta / torch.sqrt(diag_S)
    theta = torch.where(frozen, torch.zeros_like(theta), theta)
    return theta, dict(f=f, diag_S=diag_S, method=method, cg_iterations=iters, frozen=int(frozen.sum()), n_total=n_total,
                       allreduce_bytes=_COLLECTIVE_BYTES[0])
