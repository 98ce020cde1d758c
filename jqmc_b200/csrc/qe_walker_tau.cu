// GFMC_t (lrdmc-tau) entry of the fused per-walker kernel: the TAU instantiations of k_walker (qe_walker_kernel.cuh).
#include "qe_walker_kernel.cuh"

// GFMC_t projection: every walker is propagated for the imaginary time tau (jqmc/jqmc_gfmc.py:724-1110, 1539-1570).
// Main pass: a CTA loops until all of its walkers are out of time; every phase of an iteration enumerates only the walkers
// that still have time left (compact list), so a CTA's cost is the SUM of its walkers' projection counts, not
// walkers x the slowest one.  Tail pass: every walker that finished before the slowest walker of the call replays the
// remaining (no-move) iterations of the reference's while_loop -- three key splits each and one last evaluation of e_L
// with that iteration's mesh rotation -- so that keys, e_L and RT equal what the reference returns.
extern "C" int qe_lrdmc_project_tau(qe_engine* h, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys,
                                    double tau, int random_discretized_mesh, int non_local_move, double alat,
                                    int32_t* projection_counter, double* e_L, double* RT, void* stream) {
  if (!h || nw <= 0 || !w || !r_up || !Ginv || !keys || !RT || !e_L || !projection_counter || (!r_dn && h->sys.n_dn > 0))
    return fail(QE_ERR_INVALID, "qe_lrdmc_project_tau: bad argument");
  if (!(alat > 0)) return fail(QE_ERR_INVALID, "qe_lrdmc_project_tau: alat must be positive");
  if (!(tau > 0)) return fail(QE_ERR_INVALID, "qe_lrdmc_project_tau: tau must be positive");
  if (non_local_move != 0 && non_local_move != 1)
    return fail(QE_ERR_INVALID, "qe_lrdmc_project_tau: non_local_move must be 0 (tmove) or 1 (dltmove)");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_wide(h))
    return wide_lrdmc_tau(h, nw, w, r_up, r_dn, Ginv, keys, tau, random_discretized_mesh, non_local_move, alat, projection_counter,
                          e_L, RT, st);
  int rc = ensure_ws(h, 4096);
  if (rc) return rc;
  WsCarve c{(char*)h->ws};
  int* n_max = c.take<int>(4);
  CUDA_TRY(cudaMemsetAsync(n_max, 0, 4 * sizeof(int), st));
  WalkerArgs A{};
  A.nw = nw;
  A.nmpm = 1;
  A.mode = 0;
  A.dlt = non_local_move;
  A.alat = alat;
  A.w = w;
  A.r_up = r_up;
  A.r_dn = r_dn;
  A.Ginv = Ginv;
  A.RT_out = RT;
  A.e_L = e_L;
  A.tau = tau;
  A.random_mesh = random_discretized_mesh;
  A.keys = keys;
  A.pc = projection_counter;
  A.n_max = n_max;
  A.tail = 0;
  rc = launch_walker<true>(h, A, st, K_LRDMC_TAU);
  if (rc) return rc;
  A.tail = 1;
  return launch_walker<true>(h, A, st, K_LRDMC_TAU_TAIL);
}

