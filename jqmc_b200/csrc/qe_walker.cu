// Exported entries of the fused per-walker kernel (qe_walker_kernel.cuh): GFMC_n projection, V elements, VMC local energy.
// The kernel instantiations themselves live in qe_walker_i_*.cu (one translation unit per orbital padding / basis kind).
#include "qe_walker_kernel.cuh"

size_t lrdmc_draws_bytes(int nw, int nmpm) { return (size_t)nmpm * nw * (2 * 8 + 9 * 8 + 8) + 4 * 256; }
int lrdmc_draws(qe_engine* h, int nw, int nmpm, int random_mesh, uint32_t* keys, WsCarve& c, double** rRT, double** ru, cudaStream_t st) {
  const size_t n_draw = (size_t)nmpm * nw;
  uint2* sub = c.take<uint2>(n_draw * 2);
  *rRT = c.take<double>(n_draw * 9);
  *ru = c.take<double>(n_draw);
  {
    LaunchScope ls_(h, K_KEYCHAIN, st);
    k_lrdmc_keychain<<<nblk(nw, 64), 64, 0, st>>>(nw, nmpm, keys, sub);
  }
  CHECK_LAUNCH();
  {
    LaunchScope ls_(h, K_DRAWS, st);
    k_lrdmc_draws<<<nblk((long long)n_draw, 128), 128, 0, st>>>(nw, nmpm, random_mesh, sub, *rRT, *ru);
  }
  CHECK_LAUNCH();
  return QE_OK;
}

extern "C" int qe_lrdmc_project(qe_engine* h, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys,
                                double E_scf, int nmpm, int random_discretized_mesh, int non_local_move, double alat, double* RT,
                                double* V_diag, double* V_nondiag, void* stream) {
  if (!h || nw <= 0 || nmpm <= 0 || !w || !r_up || !Ginv || !keys || !RT || !V_diag || !V_nondiag || (!r_dn && h->sys.n_dn > 0))
    return fail(QE_ERR_INVALID, "qe_lrdmc_project: bad argument");
  if (!(alat > 0)) return fail(QE_ERR_INVALID, "qe_lrdmc_project: alat must be positive");
  if (non_local_move != 0 && non_local_move != 1) return fail(QE_ERR_INVALID, "qe_lrdmc_project: non_local_move must be 0 (tmove) or 1 (dltmove)");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_wide(h))
    return wide_lrdmc(h, 0, nw, w, r_up, r_dn, Ginv, keys, E_scf, nmpm, random_discretized_mesh, non_local_move, alat, nullptr, RT,
                      V_diag, V_nondiag, st);
  WalkerArgs A{};
  A.nw = nw;
  A.nmpm = nmpm;
  A.mode = 0;
  A.dlt = non_local_move;
  A.alat = alat;
  A.E_scf = E_scf;
  A.w = w;
  A.r_up = r_up;
  A.r_dn = r_dn;
  A.Ginv = Ginv;
  A.RT_out = RT;
  A.V_diag = V_diag;
  A.V_nondiag = V_nondiag;
  A.keys = keys;
  A.random_mesh = random_discretized_mesh;
  return launch_walker<false>(h, A, st, K_LRDMC_PROJ);
}

extern "C" int qe_lrdmc_velements(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                                  int non_local_move, double alat, double* V_diag, double* V_nondiag, void* stream) {
  if (!h || nw <= 0 || !r_up || !Ginv || !V_diag || !V_nondiag || (!r_dn && h->sys.n_dn > 0))
    return fail(QE_ERR_INVALID, "qe_lrdmc_velements: bad argument (Ginv is required: call qe_geminal_init first)");
  if (!(alat > 0)) return fail(QE_ERR_INVALID, "qe_lrdmc_velements: alat must be positive");
  if (use_wide(h))
    return wide_lrdmc(h, 1, nw, nullptr, const_cast<double*>(r_up), const_cast<double*>(r_dn), const_cast<double*>(Ginv), nullptr, 0.0, 1,
                      0, non_local_move, alat, RT, nullptr, V_diag, V_nondiag, (cudaStream_t)stream);
  WalkerArgs A{};
  A.nw = nw;
  A.nmpm = 1;
  A.mode = 1;
  A.dlt = non_local_move;
  A.alat = alat;
  A.r_up = const_cast<double*>(r_up);
  A.r_dn = const_cast<double*>(r_dn);
  A.Ginv = const_cast<double*>(Ginv);
  A.RT_in = RT;
  A.V_diag = V_diag;
  A.V_nondiag = V_nondiag;
  return launch_walker<false>(h, A, (cudaStream_t)stream, K_LRDMC);
}

// V_diag / V_nondiag with the nearest-nucleus assignment of the non-local ECP given by the caller: the lattice-regularised
// local energy whose finite differences are the LRDMC force terms (jqmc/jqmc_gfmc.py:5630-5667, 5840-5850).
extern "C" int qe_lrdmc_velements_frozen(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT,
                                         const double* Ginv, const int32_t* nn_index, int non_local_move, double alat,
                                         double* V_diag, double* V_nondiag, void* stream) {
  if (!h || nw <= 0 || !r_up || !Ginv || !V_diag || !V_nondiag || (!r_dn && h->sys.n_dn > 0))
    return fail(QE_ERR_INVALID, "qe_lrdmc_velements_frozen: bad argument");
  if (!(alat > 0)) return fail(QE_ERR_INVALID, "qe_lrdmc_velements_frozen: alat must be positive");
  if (use_wide(h))
    return fail(QE_ERR_UNSUPPORTED, "qe_lrdmc_velements_frozen: only the fused register kernel takes a fixed nearest-nucleus assignment");
  WalkerArgs A{};
  A.nn_fixed = h->sys.ecp_flag ? nn_index : nullptr;
  A.nw = nw;
  A.nmpm = 1;
  A.mode = 1;
  A.dlt = non_local_move;
  A.alat = alat;
  A.r_up = const_cast<double*>(r_up);
  A.r_dn = const_cast<double*>(r_dn);
  A.Ginv = const_cast<double*>(Ginv);
  A.RT_in = RT;
  A.V_diag = V_diag;
  A.V_nondiag = V_nondiag;
  return launch_walker<false>(h, A, (cudaStream_t)stream, K_LRDMC);
}

// fused local energy (mode 2); returns QE_ERR_UNSUPPORTED when the system does not fit, so that the caller
// (qe_local_energy) can fall back to the staged kernels.
int qe_local_energy_fused(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                          double* e_L, double* T_elem, double* V_parts, cudaStream_t st, const int* nn_fixed) {
  WalkerArgs A{};
  A.nn_fixed = nn_fixed;
  A.nw = nw;
  A.nmpm = 1;
  A.mode = 2;
  A.alat = 1.0;
  A.r_up = const_cast<double*>(r_up);
  A.r_dn = const_cast<double*>(r_dn);
  A.Ginv = const_cast<double*>(Ginv);
  A.RT_in = RT;
  A.e_L = e_L;
  A.T_elem = T_elem;
  A.V_parts = V_parts;
  return launch_walker<false>(h, A, st, K_EL_FUSED);
}

extern "C" int qe_set_walkers_per_cta(qe_engine* h, int wpc) {
  if (!h || wpc < 0 || wpc > 32) return fail(QE_ERR_INVALID, "qe_set_walkers_per_cta: wpc must be 0 (automatic) .. 32");
  h->wpc_override = wpc;
  return QE_OK;
}

extern "C" int qe_set_walker_warps(qe_engine* h, int warps) {
  if (!h || (warps != 0 && warps != 4 && warps != 8 && warps != 16)) return fail(QE_ERR_INVALID, "qe_set_walker_warps: warps must be 0 (default), 4, 8 or 16");
  h->walker_warps = warps;
  return QE_OK;
}

// Diagnostic: per-phase cycle counters of the fused walker kernel (thread 0 of every CTA, summed over CTAs and launches).
// enable != 0 allocates/clears the counters; out12 (host, may be NULL) receives the current sums; enable == 0 switches off.
extern "C" int qe_phase_clocks(qe_engine* h, int enable, int64_t* out12) {
  if (!h) return fail(QE_ERR_INVALID, "qe_phase_clocks: bad argument");
  if (out12 && h->phase_clk) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(out12, h->phase_clk, 12 * sizeof(long long), cudaMemcpyDeviceToHost));
  } else if (out12) {
    for (int i = 0; i < 12; ++i) out12[i] = 0;
  }
  if (enable) {
    if (!h->phase_clk) CUDA_TRY(cudaMalloc(&h->phase_clk, 16 * sizeof(long long)));
    CUDA_TRY(cudaMemset(h->phase_clk, 0, 16 * sizeof(long long)));
  } else if (h->phase_clk) {
    cudaFree(h->phase_clk);
    h->phase_clk = nullptr;
  }
  return QE_OK;
}

// ---- position derivatives (atomic forces) ---------------------------------------------------------------
// The reference differentiates e_L by automatic differentiation (jqmc/jqmc_mcmc.py:749-790), which treats the nearest-nucleus
// assignment of the non-local ECP as a constant.  A finite difference through qe_local_energy would instead jump whenever a
// displaced electron changes its nearest nucleus; these two entries let the caller freeze the assignment of the base point.
__global__ void k_nearest_nuclei(SysDev S, int nw, const double* __restrict__ r_up, const double* __restrict__ r_dn,
                                 int* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = S.n_e * S.NN;
  if (t >= (long long)nw * per) return;
  const int w = (int)(t / per), q = (int)(t % per), e = q / S.NN, nn = q % S.NN;
  const double* p = e < S.n_up ? r_up + ((size_t)w * S.n_up + e) * 3 : r_dn + ((size_t)w * S.n_dn + (e - S.n_up)) * 3;
  out[t] = nearest_atom(S.Rn, S.n_atom, p[0], p[1], p[2], nn, nullptr);
}

extern "C" int qe_nearest_nuclei(qe_engine* h, int nw, const double* r_up, const double* r_dn, int32_t* nn_index, void* stream) {
  if (!h || nw <= 0 || !r_up || !nn_index || (!r_dn && h->sys.n_dn > 0)) return fail(QE_ERR_INVALID, "qe_nearest_nuclei: bad argument");
  if (!h->sys.ecp_flag) return fail(QE_ERR_INVALID, "qe_nearest_nuclei: the Hamiltonian has no ECP");
  const long long n = (long long)nw * h->sys.n_e * h->sys.NN;
  {
    LaunchScope ls_(h, K_ECP_MESH, (cudaStream_t)stream);
    k_nearest_nuclei<<<nblk(n, 128), 128, 0, (cudaStream_t)stream>>>(h->sys, nw, r_up, r_dn, nn_index);
  }
  CHECK_LAUNCH();
  return QE_OK;
}

extern "C" int qe_local_energy_frozen(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT,
                                      const double* Ginv, const int32_t* nn_index, double* e_L, void* stream) {
  if (!h || nw <= 0 || !r_up || !Ginv || !e_L || (!r_dn && h->sys.n_dn > 0)) return fail(QE_ERR_INVALID, "qe_local_energy_frozen: bad argument");
  if (h->sys.ecp_flag && !RT) return fail(QE_ERR_INVALID, "qe_local_energy_frozen: RT required for ECP systems");
  if (use_wide(h) || !h->fused) return fail(QE_ERR_UNSUPPORTED, "qe_local_energy_frozen: only the fused register kernel takes a fixed nearest-nucleus assignment");
  return qe_local_energy_fused(h, nw, r_up, r_dn, RT, Ginv, e_L, nullptr, nullptr, (cudaStream_t)stream, h->sys.ecp_flag ? nn_index : nullptr);
}
