/*
 * jqmc_b200.h -- C ABI of the B200-native walker engine for jQMC.
 *
 * The jQMC reference has no FFI: its drivers call module-level, walker-batched JAX callables
 * (SURVEY.md §8b).  Each entry point below replaces one of those callables; the cited file:line is
 * the reference interface it stands in for.  All array arguments are DEVICE pointers (fp64,
 * C-contiguous, walker axis leading, exactly the reference's array layouts) unless the name ends in
 * `_host`; the engine never owns walker state.  Every call is asynchronous on `stream` (a
 * cudaStream_t passed as void*, NULL = default stream) and returns 0 on success or a negative
 * qe_status; qe_last_error() returns the message of the last failure on this thread.
 *
 * No torch / Python types appear here: the library links only against the CUDA runtime.
 */
#ifndef JQMC_B200_H
#define JQMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qe_engine qe_engine;

enum qe_status {
  QE_OK = 0,
  QE_ERR_INVALID = -1,     /* bad argument / inconsistent shapes (Python shim raises ValueError)   */
  QE_ERR_UNSUPPORTED = -2, /* feature outside the engine's scope (e.g. NN Jastrow, l > 6, PBC)     */
  QE_ERR_CUDA = -3,        /* CUDA runtime failure                                                 */
  QE_ERR_NOMEM = -4
};

/* One orbital basis: the AO tables of AOs_sphe_data / AOs_cart_data
 * (jqmc/atomic_orbital.py:780-929, :87-258) plus the optional MO layer of MOs_data
 * (jqmc/molecular_orbital.py:85-140).  Primitives are stored per AO, as in the reference. */
typedef struct {
  int32_t cartesian;              /* 0: spherical (l,m); 1: Cartesian (nx,ny,nz)                    */
  int32_t n_ao;
  int32_t n_prim;                 /* = num_ao_prim                                                  */
  const int32_t* nucleus_index;   /* [n_ao]                                                         */
  const int32_t* angular_momentums; /* [n_ao]                                                       */
  const int32_t* magnetic_quantum_numbers; /* [n_ao]  (spherical)                                   */
  const int32_t* polynominal_order_x;      /* [n_ao]  (Cartesian; reference spelling kept)          */
  const int32_t* polynominal_order_y;
  const int32_t* polynominal_order_z;
  const int32_t* orbital_indices; /* [n_prim] AO index of each primitive                            */
  const double* exponents;        /* [n_prim]                                                       */
  const double* coefficients;     /* [n_prim]                                                       */
  int32_t n_mo;                   /* 0 => the orbitals are the AOs themselves                       */
  const double* mo_coefficients;  /* [n_mo * n_ao] row-major                                        */
} qe_basis_desc;

/* Flat image of Hamiltonian_data (jqmc/hamiltonians.py:82-116): structure, geminal
 * (jqmc/determinant.py:93-120), Jastrow (jqmc/jastrow_factor.py:560-600, 1110-1141, 1316-1335)
 * and Coulomb/ECP tables (jqmc/coulomb_potential.py:187-233).  Host pointers, copied at create. */
typedef struct {
  int32_t n_atom;
  const double* positions;        /* [n_atom*3] bohr                                                */
  const double* effective_charges;/* [n_atom] atomic_numbers - z_cores                              */
  int32_t n_up, n_dn;
  qe_basis_desc orb_up, orb_dn;   /* Geminal_data.orb_data_{up,dn}_spin                             */
  const double* lambda_matrix;    /* [orb_num_up * (orb_num_dn + n_up - n_dn)] row-major            */
  int32_t j1_type;                /* 0 none, 1 'exp', 2 'pade'                                      */
  double j1_param;
  const double* j1_core_electrons;/* [n_atom]                                                       */
  const double* j1_atomic_numbers;/* [n_atom]                                                       */
  int32_t j2_type;                /* 0 none, 1 'pade', 2 'exp'                                      */
  double j2_param;
  int32_t j3_flag;                /* 0 none, 1 analytic three-body                                  */
  qe_basis_desc j3_orb;
  const double* j_matrix;         /* [n_orb_j3 * (n_orb_j3 + 1)] row-major                          */
  int32_t ecp_flag;
  int32_t n_ecp;
  const int32_t* ecp_nucleus_index; /* [n_ecp]                                                      */
  const int32_t* ecp_ang_moms;
  const double* ecp_exponents;
  const double* ecp_coefficients;
  const int32_t* ecp_powers;      /* TREXIO power + 2, as stored by the reference                   */
  const int32_t* ecp_max_ang_mom_plus_1; /* [n_atom]                                                */
  int32_t Nv;                     /* quadrature points: 4, 6, 12 or 18 (jqmc/_setting.py:46)        */
  int32_t NN;                     /* nearest nuclei for the non-local ECP (jqmc/_setting.py:47)     */
  int32_t precision;              /* 0: 'full' (every zone fp64, default); 1: 'mixed' (jqmc/_precision.py:345-374): AO values
                                     (zone ao_eval) and Jastrow values / ratios (jastrow_eval, jastrow_ratio) in fp32 with r - R
                                     formed in fp64 first, everything else -- MO contraction, determinant algebra, gradients
                                     and Laplacians, potentials, assembly -- in fp64.  Honoured by the register kernel family
                                     (bases with l <= 4); the general family evaluates every zone in fp64.              */
} qe_system_desc;

/* Build device tables for one Hamiltonian on the current CUDA device.  Rebuild whenever
 * hamiltonian_data changes (e.g. every optimisation step). */
int qe_create(const qe_system_desc* desc, qe_engine** out);
void qe_destroy(qe_engine* h);

/* Native input path: the same engine built straight from jQMC's HDF5 files -- `hamiltonian_data.h5` (the
 * dataclass tree of jqmc/hamiltonians.py:369-573 at the root: group = "" or "/") or a restart checkpoint (`restart.h5`,
 * group "hamiltonian_data"; jqmc/_checkpoint.py:1-30) -- without the Python stack; the HDF5 subset reader is part of the
 * library (no libhdf5).  Nv / NN / precision as in qe_system_desc.
 * qe_hdf5_summary parses the same tree without touching the device: counts16 = {n_atom, n_up, n_dn, n_ao, n_prim, n_mo,
 * cartesian, n_ecp, ecp_flag, j1_type, j2_type, j3_flag, n_ao_j3, n_mo_j3, 0, 0}, checks8 = position-weighted sums of
 * {positions, effective charges, exponents, coefficients, mo_coefficients, lambda_matrix, ECP exponents, j_matrix}.
 * qe_hdf5_read_walkers reads rank `rank`'s walkers of a restart checkpoint into HOST buffers (r_up_host == NULL: only *nw). */
int qe_create_from_hdf5(const char* path, const char* group, int Nv, int NN, int precision, qe_engine** out);
int qe_hdf5_summary(const char* path, const char* group, int64_t* counts16, double* checks8);
int qe_hdf5_read_walkers(const char* path, int rank, int capacity_walkers, int n_up, int n_dn, int* nw, double* r_up_host,
                         double* r_dn_host, uint32_t* keys_host);
const char* qe_last_error(void);
/* Library/ABI version and whether it was compiled for sm_100a. */
int qe_version(void);

/* _geminal_inv_batched (jqmc/jqmc_mcmc.py:4248-4275) and GFMC_n's _jit_vmap_A_inv_n
 * (jqmc/jqmc_gfmc.py:4706-4717):  r_up[nw,n_up,3], r_dn[nw,n_dn,3] -> G[nw,n_up,n_up], Ginv[...]. */
int qe_geminal_init(qe_engine* h, int nw, const double* r_up, const double* r_dn, double* G, double* Ginv,
                    void* stream);

/* _jit_vmap_update (jqmc/jqmc_mcmc.py:4278-4533, 4728): nmpm single-electron Metropolis proposals
 * per walker, updating r_up, r_dn, keys[nw,2] (uint32, jax.random raw keys), G, Ginv in place;
 * acc/rej[nw] int32 receive the accepted / rejected counts of THIS call. */
int qe_mcmc_update(qe_engine* h, int nw, double* r_up, double* r_dn, uint32_t* keys, double* G, double* Ginv,
                   int nmpm, double Dt, double epsilon_AS, int32_t* acc, int32_t* rej, void* stream);

/* _jit_vmap_generate_RTs (jqmc/jqmc_mcmc.py:4228-4245, 4738): keys[nw,2] -> RT[nw,3,3]; keys not advanced. */
int qe_rotation(qe_engine* h, int nw, const uint32_t* keys, double* RT, void* stream);

/* _jit_vmap_e_L_fast == vmap(compute_local_energy_fast) (jqmc/hamiltonians.py:225-290, jqmc_mcmc.py:4736).
 * Optional per-walker breakdown (any may be NULL): T_elem[nw,n_up+n_dn] per-electron kinetic energies
 * (jqmc/wavefunction.py:1141-1207), V_parts[nw,4] = {bare Coulomb, ECP local, ECP non-local, 0}. */
int qe_local_energy(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT,
                    const double* Ginv, double* e_L, double* T_elem, double* V_parts, void* stream);

/* Position derivatives (atomic forces; the reference: jax.grad of compute_local_energy, jqmc/jqmc_mcmc.py:749-790, 4744-4746).
 * The engine differentiates by central finite differences of qe_local_energy / qe_ln_wavefunction on the device
 * (jqmc_b200/forces.py).  Automatic differentiation treats the nearest-nucleus assignment of the non-local ECP as a constant;
 * to do the same, qe_nearest_nuclei returns that assignment at the base point -- nn_index[nw, n_up + n_dn, NN] (int32) -- and
 * qe_local_energy_frozen evaluates e_L at displaced points with it (register kernel family only; nn_index may be NULL for
 * all-electron systems). */
int qe_nearest_nuclei(qe_engine* h, int nw, const double* r_up, const double* r_dn, int32_t* nn_index, void* stream);
int qe_local_energy_frozen(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                           const int32_t* nn_index, double* e_L, void* stream);
/* The same for the lattice-regularised local energy V_diag + V_nondiag of LRDMC, whose position gradient is the LRDMC force
 * term (grad of _compute_local_energy_n / _compute_local_energy_t, jqmc/jqmc_gfmc.py:5630-5667, 1461-1486). */
int qe_lrdmc_velements_frozen(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                              const int32_t* nn_index, int non_local_move, double alat, double* V_diag, double* V_nondiag,
                              void* stream);

/* _jit_vmap_as_reg_fast (jqmc/determinant.py:1223-1260, jqmc_mcmc.py:4739). */
int qe_as_factor(qe_engine* h, int nw, const double* G, const double* Ginv, double* R_AS, void* stream);

/* vmap(evaluate_ln_wavefunction) (jqmc/wavefunction.py:677-720): ln|Psi| = J + ln|det G|, and the sign of det. */
int qe_ln_wavefunction(qe_engine* h, int nw, const double* r_up, const double* r_dn, double* ln_psi,
                       double* sign, void* stream);

/* Parity/diagnostic entry for kernels 1-2: orbital values, gradients and Laplacians at arbitrary
 * points.  which: 0 = geminal up basis, 1 = geminal dn basis, 2 = J3 basis; layer: 0 = AO layer
 * (compute_AOs_value_grad_lap, jqmc/atomic_orbital.py:3594-3640), 1 = orbital layer
 * (compute_MOs_value_grad_lap, jqmc/molecular_orbital.py:375-415).  r[n_pts,3];
 * out[5, n_orb, n_pts] = value, d/dx, d/dy, d/dz, laplacian (orbital-major like the reference). */
int qe_eval_orbitals(qe_engine* h, int which, int layer, int n_pts, const double* r, double* out, void* stream);

/* Single-electron wavefunction ratios Psi(r')/Psi(r) for a batch of moves per walker
 * (_compute_ratio_determinant_part_split_spin jqmc/determinant.py:1665-1783 times
 *  _compute_ratio_Jastrow_part_split_spin jqmc/jastrow_factor.py:2696-2960).
 * elec[n_moves] = electron index (0..n_up-1 up, n_up.. down) shared by all walkers,
 * r_new[nw,n_moves,3]; det_ratio / jas_ratio [nw,n_moves] (either may be NULL). */
int qe_move_ratios(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* Ginv, int n_moves,
                   const int32_t* elec_host, const double* r_new, double* det_ratio, double* jas_ratio,
                   void* stream);

/* GFMC_n._projection_n (jqmc/jqmc_gfmc.py:4738-5358, 5656): nmpm LRDMC projections per walker.
 * w[nw], r_up, r_dn, Ginv, keys updated in place; RT[nw,3,3], V_diag[nw], V_nondiag[nw] written.
 * non_local_move: 0 = 'tmove', 1 = 'dltmove'. */
int qe_lrdmc_project(qe_engine* h, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys,
                     double E_scf, int nmpm, int random_discretized_mesh, int non_local_move, double alat,
                     double* RT, double* V_diag, double* V_nondiag, void* stream);

/* GFMC_t projection loop (jqmc/jqmc_gfmc.py:724-1110 `_projection_t_core`, driver `_run_projection_loop` :1539-1570):
 * every walker is propagated for the imaginary time tau by continuous-time lattice-regularised projections
 * (time step log(1-xi)/V_nondiag, weight *= exp(-dt e_L), move chosen as in GFMC_n) until its time is used up.
 * w[nw], r_up, r_dn, Ginv, keys updated in place; projection_counter[nw] (int32), e_L[nw], RT[nw,3,3] written.
 * As in the reference's vmapped while_loop, walkers that finish early keep splitting their keys (three splits per
 * iteration) until the slowest walker OF THIS CALL is done, and e_L / RT belong to that last iteration.
 * Register kernel family: asynchronous like every other entry (main pass + tail pass, the projection count stays on the device).
 * General kernel family: NOT asynchronous -- the number of projections is data dependent and that family launches one kernel
 * sequence per projection, so the host reads an "any walker still running" flag back after every projection
 * (cudaStreamSynchronize on `stream`), and all walkers run in one slice. */
int qe_lrdmc_project_tau(qe_engine* h, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys,
                         double tau, int random_discretized_mesh, int non_local_move, double alat,
                         int32_t* projection_counter, double* e_L, double* RT, void* stream);

/* GFMC_n._compute_V_elements_n (jqmc/jqmc_gfmc.py:5360-5627, 5660). */
int qe_lrdmc_velements(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT,
                       const double* Ginv, int non_local_move, double alat, double* V_diag, double* V_nondiag,
                       void* stream);  /* Ginv: inverse at (r_up, r_dn), e.g. from qe_geminal_init as the reference does */

/* Per-step weighted sums of GFMC_n.run (jqmc/jqmc_gfmc.py:5971-5976), e_L = V_diag + V_nondiag:
 * out5 (device) = { nw, sum w, sum w/(V_diag-E_scf), sum w/(V_diag-E_scf) e_L, sum w/(V_diag-E_scf) e_L^2 }.
 * The caller sums out5 over ranks (the reference: mpi reduce, :6016-6020).
 * GFMC_t form (jqmc/jqmc_gfmc.py:1929-1932): V_diag == NULL, V_nondiag holds e_L; out5 = { nw, sum w, sum w, sum w e_L,
 * sum w e_L^2 }. */
int qe_lrdmc_collect(qe_engine* h, int nw, const double* w, const double* V_diag, const double* V_nondiag, double E_scf,
                     double* out5, void* stream);

/* Walker reconfiguration, index part (jqmc/jqmc_gfmc.py:6069-6135).  w_all[world*nw] holds the weights of ALL ranks in
 * rank order (all-gathered by the caller); zeta in [0,1) is rank 0's np.random.random() (:5948-5952).  Every rank
 * evaluates the same comb:  chosen_all[g] = searchsorted(global cumulative probabilities, (g + zeta)/(world*nw)),
 * g = destination slot (rank = g / nw); *n_survived = number of distinct sources.  Summation order = the reference's
 * (np.sum pairwise per rank, np.cumsum sequential, Exscan rank offsets), so the indices agree bit for bit. */
int qe_lrdmc_branch(qe_engine* h, int nw, int world, const double* w_all, double zeta, int32_t* chosen_all,
                    int32_t* n_survived, void* stream);

/* Walker reconfiguration, data part (jqmc/jqmc_gfmc.py:6146-6318): dst walker i <- src walker chosen_local[i], where the
 * src arrays are the all-gathered coordinates [world*nw, n_up|n_dn, 3] and chosen_local = chosen_all + rank*nw. */
int qe_gather_walkers(qe_engine* h, int nw, const int32_t* chosen_local, const double* src_r_up, const double* src_r_dn,
                      double* dst_r_up, double* dst_r_dn, void* stream);

/* Packed exchange of the reconfiguration: ONE all_gather per branching step instead of the reference's reduce + Exscan +
 * Allgather x 2 + Alltoallv + Isend/Irecv rounds (jqmc/jqmc_gfmc.py:6016-6020, 6081-6296).  Every rank packs its record
 *   [5 weighted sums of qe_lrdmc_collect | 3 pad | w[nw] | r_up[nw,n_up,3] | r_dn[nw,n_dn,3]]   (qe_lrdmc_record_len doubles)
 * with qe_lrdmc_pack; the caller all-gathers the records (rank order); qe_lrdmc_reconfigure_packed then sums the five sums
 * over ranks in rank order (sums5 may be NULL), evaluates the comb exactly as qe_lrdmc_branch does (chosen_all[world*nw],
 * *n_survived) and writes this rank's nw new walkers straight from the gathered records. */
int64_t qe_lrdmc_record_len(qe_engine* h, int nw);
int qe_lrdmc_pack(qe_engine* h, int nw, const double* sums5, const double* w, const double* r_up, const double* r_dn,
                  double* record, void* stream);
int qe_lrdmc_reconfigure_packed(qe_engine* h, int nw, int world, int rank, const double* records, double zeta,
                                int32_t* chosen_all, int32_t* n_survived, double* sums5, double* dst_r_up, double* dst_r_dn,
                                void* stream);

/* qe_local_energy runs as ONE fused kernel (shared-memory resident walker state) when the system fits and
 * `on` != 0 (default); otherwise as a chain of staged kernels through a global workspace.  Same results. */
int qe_set_fused(qe_engine* h, int on);

/* _jit_vmap_grad_ln_psi_params_fast == vmap(grad(evaluate_ln_wavefunction_fast)) (jqmc/jqmc_mcmc.py:4748, 854-876): per-walker
 * derivatives of ln|Psi| with respect to the variational parameters, running inverse held fixed -- the O_k of stochastic
 * reconfiguration, in the reference's parameter layout (blocks of jqmc/wavefunction.py:515-674; before get_dln_WF's
 * projection / symmetrisation, jqmc_mcmc.py:1372-1513).  Any output may be NULL:
 *   d_j1[nw], d_j2[nw]                    d/d jastrow_1b_param, d/d jastrow_2b_param
 *   d_j3[nw, n_orb_j3, n_orb_j3 + 1]      d/d j_matrix (last column: one-body vector)
 *   d_lambda[nw, n_orb, n_orb + n_up - n_dn]   d/d lambda_matrix (paired block | unpaired columns) */
int qe_dln_wf(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* Ginv, double* d_j1, double* d_j2,
              double* d_j3, double* d_lambda, void* stream);

/* Kernel family.  0 (default): automatic -- the register/shared-memory kernels when they cover the system (MO-basis
 * geminal with <= 16 orbitals, <= 8 electrons per spin, J1 + J2), otherwise the general path (AO-basis JAGP geminals, any
 * orbital count, <= 112 electrons per spin, three-body Jastrow; contractions on the fp64 tensor cores).  1: always the
 * general path.  Same results to round-off; bit-identical accept/reject and branching decisions are asserted in tests. */
int qe_set_path(qe_engine* h, int path);
/* Debugging aid: nonzero replaces the tensor-core GEMM of the general path by a plain DFMA kernel (this handle only). */
int qe_set_gemm_reference(qe_engine* h, int on);
/* General family: walkers per slice of a call (0 = automatic from the free device memory).  Calls over more walkers run as
 * consecutive slices through one workspace; results do not depend on the slicing.  Per handle, like every other switch. */
int qe_set_wide_slice(qe_engine* h, int walkers);

/* Walkers per CTA of the fused walker kernel: 0 = chosen automatically so that the grid fills the SMs (default),
 * 1..32 = fixed (tuning / tests; results do not depend on it beyond round-off of partial-sum order). */
int qe_set_walkers_per_cta(qe_engine* h, int wpc);
/* Tuning knob: warps per CTA of the fused walker kernel: 0 / 16 = one 16-warp CTA per SM (default), 8 = two CTAs per SM,
 * 4 = four.  Same results. */
int qe_set_walker_warps(qe_engine* h, int warps);

/* Diagnostic: per-phase cycle counters of the fused walker kernel (projection loop phases P0..P4 and write-back), summed
 * over CTAs (thread 0 of each) and launches since they were enabled.  enable != 0 allocates / clears them, out12 (host,
 * may be NULL) receives the sums accumulated so far (synchronises); enable == 0 switches the timing off again. */
int qe_phase_clocks(qe_engine* h, int enable, int64_t* out12);

/* Microbenchmark used by bench.py for the fp64 roofline denominator: runs `iters` dependent-free
 * DFMA per thread on a full grid and returns the achieved TFLOP/s (synchronous). */
int qe_measure_fp64_peak(int iters, double* tflops);

/* Number of kernel launches issued by this engine since creation (bench.py's gpu_launches). */
int64_t qe_launch_count(qe_engine* h);

/* Optional per-kernel device timing for bench.py's roofline line: when enabled, every launch is
 * bracketed by CUDA events on its own stream.  qe_profile(h, on) also clears the records;
 * qe_profile_read() sums the elapsed time and the launch count of kernel `id` (0 .. qe_profile_kernels()-1,
 * named by qe_profile_name) since then (synchronises on the recorded events). */
int qe_profile(qe_engine* h, int enable);
int qe_profile_kernels(void);
const char* qe_profile_name(int id);
int qe_profile_read(qe_engine* h, int id, double* total_ms, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* JQMC_B200_H */
