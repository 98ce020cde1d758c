// Shared declarations of the walker engine translation units (qe_engine.cu, qe_mcmc.cu, qe_walker.cu).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges named like the profile() kernel table

#include "../../include/jqmc_b200.h"
#include "qe_device.cuh"

using namespace qe;
// =================================================================================================
// error handling
// =================================================================================================
int qe_fail(int code, const std::string& msg);  // defined in qe_engine.cu (thread-local last error)
#define fail qe_fail
#define CUDA_TRY(x)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess)                                                                       \
      return fail(QE_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_));                \
  } while (0)

// =================================================================================================
// engine object
// =================================================================================================
struct DevPool {
  std::vector<void*> ptrs;
  template <class T>
  cudaError_t upload(const std::vector<T>& v, const T** out) {
    void* p = nullptr;
    size_t n = std::max<size_t>(v.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) return e;
    ptrs.push_back(p);
    if (!v.empty()) e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    *out = (const T*)p;
    return e;
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
  }
};

#define QE_N_CHUNK 15  // target number of balanced basis chunks (partial sums are added in chunk order)
struct HostBasis {
  BasisDev dev{};
  int n_chunk = 1;   // balanced shell-granular chunks of the basis (deterministic partial sums)
  int off_cseg = 0;  // byte offset of the chunk segment list (int4, same format as the group list) in the blob
  int off_cbeg = 0;  // byte offset of chunk_begin[n_chunk+1] (first segment of each chunk)
  int chunk_warp[32] = {0};  // chunk evaluated by warp w of the Metropolis kernel: warps that share an SM sub-partition (w mod 4)
                             // get chunks of the same angular momentum, i.e. the same code (instruction-cache locality)
  bool present = false;
};

struct SysDev {
  int n_atom, n_up, n_dn, n_e;
  const double* Rn;     // [n_atom*3]
  const double* Zeff;   // [n_atom]
  // geminal lambda in orbital basis: lam_p [nmo_pad][nmo_pad] (zero padded), lam_u [nmo_pad][n_up-n_dn]
  const double* lam_p;
  const double* lam_u;
  int n_unp;
  // Jastrow
  int j1_type;
  double j1_a;
  const double* j1_A;   // (2 Z)^{3/4}
  const double* j1_c;   // (2 Z)^{1/4}
  int j2_type;
  double j2_a;
  // ECP
  int ecp_flag, n_ecp, Nv, NN, ecp_lmax;  // ecp_lmax = global max_ang_mom_plus_1
  const int* ecp_nuc;
  const int* ecp_l;
  const double* ecp_z;
  const double* ecp_c;
  const double* ecp_p;
  const int* ecp_lmax_atom;  // [n_atom] max_ang_mom_plus_1
  const int* ecp_off;        // [n_atom+1] terms sorted by atom
  const double* quad_w;      // [Nv]
  const double* quad_g;      // [Nv*3]
  double v_ion_ion;
};

// Tables of the general ("wide") path (qe_wide.cu): arbitrary orbital counts (AO-basis geminals, many MOs), any number
// of electrons, three-body Jastrow.  All matrices are row-major device arrays; "row" = canonical AO row of the basis
// image (row_ao / row_scale of qe_device.cuh), the per-AO scale is folded into every matrix that touches row space.
struct WideTabs {
  bool present = false;
  int no = 0, n_row = 0, has_mo = 0, restricted = 1;
  const double *CwT_up = nullptr, *CwT_dn = nullptr;  // [no][n_row]   orbital <- AO rows
  const double *Cw_up = nullptr, *Cw_dn = nullptr;    // [n_row][no]   AO-row weights <- orbital weights
  const double *lamP = nullptr, *lamPT = nullptr;     // [no][no]      paired block of lambda and its transpose
  const double* lamU = nullptr;                       // [no][n_unp]
  int j3 = 0, nj = 0, nj_row = 0, j3_mo = 0;
  const double *CjT = nullptr, *Cj = nullptr;         // [nj][nj_row], [nj_row][nj]
  const double *Mj = nullptr, *MjT = nullptr;         // [nj][nj]
  const double* j1v = nullptr;                        // [nj]
};

struct qe_engine {
  DevPool pool;
  WideTabs wt;
  bool narrow_ok = true;  // the register/shared-memory kernels (qe_mcmc.cu, qe_walker.cu) cover this system
  int path = 0;           // 0: automatic (narrow when it covers the system), 1: always the wide path
  HostBasis b_up, b_j3;  // b_up carries both spins' MO coefficient tables (the AO tables are shared: checked at create)
  SysDev sys{};
  int nmo_pad = 4;
  // workspace
  void* ws = nullptr;
  size_t ws_bytes = 0;
  int64_t launches = 0;
  double* gemm_ws = nullptr;    // split-K partial products of the general path's GEMM (grow-only)
  size_t gemm_ws_bytes = 0;
  double* branch_ws = nullptr;  // scratch of qe_lrdmc_branch: rank sums, rank probabilities, cumulative probabilities
  size_t branch_ws_n = 0;
  // optional per-kernel timing (qe_profile): CUDA events recorded on the launch stream around each kernel
  bool profiling = false;
  bool fused = true;  // qe_local_energy uses the fused walker kernel when the system fits
  int wpc_override = 0;  // walkers per CTA of the fused walker kernel (0 = automatic)
  long long* phase_clk = nullptr;  // optional per-phase cycle counters of the fused walker kernel (qe_phase_clocks)
  bool mixed = false;     // precision mode 'mixed' (qe_system_desc.precision = 1): fp32 AO values and Jastrow ratios in the register kernels
  bool gemm_ref = false;  // general family: plain DFMA GEMM instead of the tensor-core kernel (qe_set_gemm_reference)
  int wide_slice = 0;     // general family: walkers per slice of a call, 0 = automatic (qe_set_wide_slice)
  int walker_warps = 0;  // warps per CTA of the fused walker kernel (0 = 16: one CTA per SM; 8: two CTAs per SM; 4: four)
  struct ProfRec { int id; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof;
};

enum KernelId { K_ORB_EL = 0, K_GEMINAL, K_ALGEBRA, K_ECP_MESH, K_REDUCE, K_RATIOS, K_AS, K_ROT, K_KEYCHAIN, K_DRAWS, K_MCMC,
                K_EVAL, K_LRDMC, K_EL_FUSED, K_LRDMC_PROJ, K_COLLECT, K_BRANCH, K_GATHER,
                K_W_AO, K_W_GEMM, K_W_BMM, K_W_INV, K_W_MESH, K_W_ELEC, K_W_SELECT, K_W_COMMIT, K_W_DECIDE, K_W_MISC, K_LRDMC_TAU, K_LRDMC_TAU_TAIL, K_COUNT };
static const char* const KERNEL_NAMES[K_COUNT] = {"k_orb_electrons", "k_geminal", "k_electron_algebra", "k_ecp_mesh", "k_reduce_eL",
                                                  "k_move_ratios", "k_as_factor", "k_rotation", "k_mcmc_keychain", "k_mcmc_draws",
                                                  "k_mcmc", "k_eval_orbitals", "k_walker(V_elements)", "k_walker(e_L)", "k_walker(projection)", "k_lrdmc_collect",
                                                  "k_branch", "k_gather_walkers",
                                                  "kw_ao_store", "kw_dgemm(DMMA)", "kw_bmm", "kw_inverse", "kw_mesh", "kw_electron",
                                                  "kw_lrdmc_select", "kw_lrdmc_commit", "kw_mc_decide", "kw_misc", "k_walker(projection_tau)",
                                                  "k_walker(projection_tau tail)"};
struct LaunchScope {
  qe_engine* h;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int id;
  LaunchScope(qe_engine* h_, int id_, cudaStream_t st_) : h(h_), st(st_), id(id_) {
    h->launches++;
    nvtxRangePushA(KERNEL_NAMES[id]);  // host-side NVTX range per launch (a no-op unless a tool is attached)
    if (h->profiling) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, st);
    }
  }
  ~LaunchScope() {
    nvtxRangePop();
    if (e0) {
      cudaEventRecord(e1, st);
      h->prof.push_back({id, e0, e1});
    }
  }
};

static inline int ensure_ws(qe_engine* h, size_t bytes) {
  if (bytes <= h->ws_bytes) return QE_OK;
  if (h->ws) cudaFree(h->ws);
  h->ws = nullptr;
  h->ws_bytes = 0;
  CUDA_TRY(cudaMalloc(&h->ws, bytes));
  h->ws_bytes = bytes;
  return QE_OK;
}

// =================================================================================================
// small device helpers
// =================================================================================================
// Fast fp64 reciprocal square root / reciprocal for the O(N_e) pair loops evaluated at EVERY mesh point (Jastrow ratios):
// single-precision MUFU seed + two Newton steps (relative error ~1e-16, i.e. fp64 round-off), ~10 DFMA-class instructions
// instead of the ~80 of the IEEE sqrt + divide sequences.  Arguments must lie in the float range (distances, 1 + a d).
__device__ __forceinline__ double qrsqrt(double x) {
  double y = (double)rsqrtf((float)x);
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}
__device__ __forceinline__ double qrcp(double x) {
  double y = (double)__frcp_rn((float)x);
  y = y * fma(-x, y, 2.0);
  y = y * fma(-x, y, 2.0);
  return y;
}
// The same two functions seeded by the fp64 MUFU approximations (rsqrt.approx.ftz.f64 / rcp.approx.ftz.f64: ~2^-22, no
// float conversions) with ONE third-order correction step each: 5 / 3 fp64-pipe instructions, relative error ~1e-16
// (the IEEE sqrt and divide sequences cost ~20 each and carry special-case branches).  x must be a positive normal number.
__device__ __forceinline__ double mrsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);           // 1 - x y^2
  return fma(y * e, fma(0.375, e, 0.5), y);         // y (1 + e/2 + 3 e^2/8)
}
__device__ __forceinline__ double mrcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);                 // 1 - x y
  return fma(y, fma(e, e, e), y);                   // y (1 + e + e^2)
}
// |(dx,dy,dz)| ; distances below 1e-15 bohr are clamped (the reference's Jastrow derivatives clamp at 1e-12)
__device__ __forceinline__ double qdist(double dx, double dy, double dz) {
  const double r2 = fmax(fma(dx, dx, fma(dy, dy, dz * dz)), 1.0e-30);
  return r2 * qrsqrt(r2);
}

// index of the atom with distance-rank `rank` from p (argsort semantics, first index wins ties;
// jqmc/structure.py:410-426)
__device__ __forceinline__ int nearest_atom(const double* __restrict__ Rn, int n_atom, double px, double py, double pz,
                                            int rank, double* dist_out) {
  if (rank == 0) {
    int best = 0;
    double bd = 1e300;
    for (int a = 0; a < n_atom; ++a) {  // squared distances: one sqrt at the end (same order as the norms)
      const double dx = Rn[3 * a] - px, dy = Rn[3 * a + 1] - py, dz = Rn[3 * a + 2] - pz;
      const double d = dx * dx + dy * dy + dz * dz;
      if (d < bd) {
        bd = d;
        best = a;
      }
    }
    if (dist_out) *dist_out = sqrt(bd);
    return best;
  }
  for (int a = 0; a < n_atom; ++a) {
    const double dx = Rn[3 * a] - px, dy = Rn[3 * a + 1] - py, dz = Rn[3 * a + 2] - pz;
    const double da = sqrt(dx * dx + dy * dy + dz * dz);
    int r = 0;
    for (int b = 0; b < n_atom; ++b) {
      const double ex = Rn[3 * b] - px, ey = Rn[3 * b + 1] - py, ez = Rn[3 * b + 2] - pz;
      const double db = sqrt(ex * ex + ey * ey + ez * ez);
      r += (db < da) || (db == da && b < a);
    }
    if (r == rank) {
      if (dist_out) *dist_out = da;
      return a;
    }
  }
  return 0;
}

__device__ __forceinline__ double j1_f(int type, double a, double A, double c, double d) {
  // jqmc/jastrow_factor.py:648-726
  if (type == 1) return -A * (1.0 - qexp(-a * c * d)) / (2.0 * a);
  return -A * d / (2.0 * (1.0 + a * c * d));
}
__device__ __forceinline__ double j2_f(int type, double a, double d) {
  // jqmc/jastrow_factor.py:1180-1251
  if (type == 1) return d / (2.0 * (1.0 + a * d));
  return (1.0 - qexp(-a * d)) / (2.0 * a);
}
// the same functions with the fast reciprocal, `inv2a` = 1 / (2a): used inside the per-mesh-point pair loops
__device__ __forceinline__ double j1_fq(int type, double a, double inv2a, double A, double c, double d) {
  if (type == 1) return -A * (1.0 - qexp(-a * c * d)) * inv2a;
  return -0.5 * A * d * qrcp(fma(a * c, d, 1.0));
}
__device__ __forceinline__ double j2_fq(int type, double a, double inv2a, double d) {
  if (type == 1) return 0.5 * d * qrcp(fma(a, d, 1.0));
  return (1.0 - qexp(-a * d)) * inv2a;
}

// d^n for the small integer powers of the ECP radial terms (stored as doubles; TREXIO power + 2, jqmc/coulomb_potential.py:1562-1568)
__device__ __forceinline__ double ipow(double d, double pw) {
  const int n = (int)pw;
  if ((double)n != pw) return pow(d, pw);
  double r = 1.0;
  const int m = n < 0 ? -n : n;
  for (int i = 0; i < m; ++i) r *= d;
  return n < 0 ? 1.0 / r : r;
}

// Legendre P_l(x), l <= 6 (jqmc/_function_collections.py:47-65)
__device__ __forceinline__ double legendre_l(int l, double x) {
  const double x2 = x * x;
  switch (l) {
    case 0: return 1.0;
    case 1: return x;
    case 2: return 0.5 * (3.0 * x2 - 1.0);
    case 3: return 0.5 * (5.0 * x2 - 3.0) * x;
    case 4: return 0.125 * ((35.0 * x2 - 30.0) * x2 + 3.0);
    case 5: return 0.125 * ((63.0 * x2 - 70.0) * x2 + 15.0) * x;
    default: return 0.0625 * (((231.0 * x2 - 315.0) * x2 + 105.0) * x2 - 5.0);
  }
}


// ECP mesh point (e, nn, k): position, and the channel-summed angular factor  sum_l V_l(d)(2l+1)P_l(cos) * w_k
// (jqmc/coulomb_potential.py:1562-1575, 1607-1645)
__device__ __forceinline__ void ecp_point(const SysDev& S, const double* rt, double x, double y, double z, int nn, int k,
                                          double& px, double& py, double& pz, double& ang_w, bool want_ang) {
  double d;
  const int a = nearest_atom(S.Rn, S.n_atom, x, y, z, nn, &d);
  const double relx = S.Rn[3 * a] - x, rely = S.Rn[3 * a + 1] - y, relz = S.Rn[3 * a + 2] - z;
  d = sqrt(relx * relx + rely * rely + relz * relz);
  const double q0 = S.quad_g[3 * k], q1 = S.quad_g[3 * k + 1], q2 = S.quad_g[3 * k + 2];
  const double gx = q0 * rt[0] + q1 * rt[3] + q2 * rt[6];
  const double gy = q0 * rt[1] + q1 * rt[4] + q2 * rt[7];
  const double gz = q0 * rt[2] + q1 * rt[5] + q2 * rt[8];
  px = x + relx + d * gx;
  py = y + rely + d * gy;
  pz = z + relz + d * gz;
  ang_w = 0.0;
  if (!want_ang) return;
  const double gn = sqrt(gx * gx + gy * gy + gz * gz);
  const double cos_t = (-relx / d) * (gx / gn) + (-rely / d) * (gy / gn) + (-relz / d) * (gz / gn);
  const int lloc = S.ecp_lmax_atom[a];
  double ang = 0.0;
  for (int l = 0; l < lloc; ++l) {
    double vl = 0.0;
    for (int kk = S.ecp_off[a]; kk < S.ecp_off[a + 1]; ++kk)
      if (S.ecp_l[kk] == l) vl += S.ecp_c[kk] * ipow(d, S.ecp_p[kk]) * qexp(-S.ecp_z[kk] * d * d);
    ang = fma(vl / (d * d) * (2 * l + 1), legendre_l(l, cos_t), ang);
  }
  ang_w = ang * S.quad_w[k];
}

// electron position accessors: global AoS arrays r_up[nw][n_up][3], r_dn[nw][n_dn][3]
struct PosGlobal {
  const double* __restrict__ up;
  const double* __restrict__ dn;
  int n_up, n_dn, w;
  __device__ __forceinline__ void get(int e, double& x, double& y, double& z) const {
    const double* p = e < n_up ? up + ((size_t)w * n_up + e) * 3 : dn + ((size_t)w * n_dn + (e - n_up)) * 3;
    x = p[0];
    y = p[1];
    z = p[2];
  }
};

// Part of the Jastrow exponent (J1+J2) that depends on electron e when it sits at (x,y,z):
//   sum_a j1(|x - R_a|) + sum_{j != e} j2(|x - r_j|);   J(r') - J(r) = jastrow_single(r') - jastrow_single(r)
template <class Pos>
__device__ __forceinline__ double jastrow_single(const SysDev& S, const Pos& pos, int e, double x, double y, double z) {
  double J = 0.0;
  if (S.j1_type) {
    for (int a = 0; a < S.n_atom; ++a) {
      const double X = S.Rn[3 * a], Y = S.Rn[3 * a + 1], Z = S.Rn[3 * a + 2];
      const double d = sqrt((x - X) * (x - X) + (y - Y) * (y - Y) + (z - Z) * (z - Z));
      J += j1_f(S.j1_type, S.j1_a, S.j1_A[a], S.j1_c[a], d);
    }
  }
  if (S.j2_type) {
    for (int j = 0; j < S.n_e; ++j) {
      if (j == e) continue;
      double xj, yj, zj;
      pos.get(j, xj, yj, zj);
      const double d = sqrt((x - xj) * (x - xj) + (y - yj) * (y - yj) + (z - zj) * (z - zj));
      J += j2_f(S.j2_type, S.j2_a, d);
    }
  }
  return J;
}

// jastrow_single with the MUFU-seeded reciprocal square root / reciprocal (mesh phase of the fused walker kernel: one call
// per mesh point, N_e + N_at terms each).  Distances are clamped at 1e-150 (a mesh point on top of another particle).
__device__ __forceinline__ double j1_fm(int type, double a, double inv2a, double A, double c, double d) {
  if (type == 1) return -A * (1.0 - qexp(-a * c * d)) * inv2a;
  return -0.5 * A * d * mrcp(fma(a * c, d, 1.0));
}
__device__ __forceinline__ double j2_fm(int type, double a, double inv2a, double d) {
  if (type == 1) return 0.5 * d * mrcp(fma(a, d, 1.0));
  return (1.0 - qexp(-a * d)) * inv2a;
}
template <class Pos>
__device__ __forceinline__ double jastrow_single_m(const SysDev& S, const Pos& pos, int e, double x, double y, double z) {
  double J = 0.0;
  if (S.j1_type) {
    const double inv2a = 1.0 / (2.0 * S.j1_a);
    for (int a = 0; a < S.n_atom; ++a) {
      const double dx = x - S.Rn[3 * a], dy = y - S.Rn[3 * a + 1], dz = z - S.Rn[3 * a + 2];
      const double r2 = fmax(fma(dx, dx, fma(dy, dy, dz * dz)), 1.0e-300);
      J += j1_fm(S.j1_type, S.j1_a, inv2a, S.j1_A[a], S.j1_c[a], r2 * mrsqrt(r2));
    }
  }
  if (S.j2_type) {
    const double inv2a = 1.0 / (2.0 * S.j2_a);
    double J0 = 0.0, J1 = 0.0;  // two accumulators: the terms of consecutive electrons are independent chains
    int j = 0;
    for (; j + 1 < S.n_e; j += 2) {
      double xa, ya, za, xb, yb, zb;
      pos.get(j, xa, ya, za);
      pos.get(j + 1, xb, yb, zb);
      const double ax = x - xa, ay = y - ya, az = z - za, bx = x - xb, by = y - yb, bz = z - zb;
      const double ra = fmax(fma(ax, ax, fma(ay, ay, az * az)), 1.0e-300), rb = fmax(fma(bx, bx, fma(by, by, bz * bz)), 1.0e-300);
      const double fa = j2_fm(S.j2_type, S.j2_a, inv2a, ra * mrsqrt(ra)), fb = j2_fm(S.j2_type, S.j2_a, inv2a, rb * mrsqrt(rb));
      J0 += j == e ? 0.0 : fa;
      J1 += j + 1 == e ? 0.0 : fb;
    }
    if (j < S.n_e && j != e) {
      double xa, ya, za;
      pos.get(j, xa, ya, za);
      const double ax = x - xa, ay = y - ya, az = z - za;
      const double ra = fmax(fma(ax, ax, fma(ay, ay, az * az)), 1.0e-300);
      J0 += j2_fm(S.j2_type, S.j2_a, inv2a, ra * mrsqrt(ra));
    }
    J += J0 + J1;
  }
  return J;
}

// Mixed-precision mode (jqmc/_precision.py:345-374): zones `jastrow_eval` / `jastrow_ratio` in fp32.  The coordinate
// differences are formed in fp64 and then rounded (:61-76); distances, the J1 / J2 functions and their sum run in float.
template <class Pos>
__device__ __forceinline__ float jastrow_single_f32(const SysDev& S, const Pos& pos, int e, double x, double y, double z) {
  float J = 0.0f;
  if (S.j1_type) {
    const float a = (float)S.j1_a, inv2a = 1.0f / (2.0f * a);
    for (int k = 0; k < S.n_atom; ++k) {
      const float dx = (float)(x - S.Rn[3 * k]), dy = (float)(y - S.Rn[3 * k + 1]), dz = (float)(z - S.Rn[3 * k + 2]);
      const float d = sqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1.0e-37f));
      const float A = (float)S.j1_A[k], c = (float)S.j1_c[k];
      J += S.j1_type == 1 ? -A * (1.0f - expf(-a * c * d)) * inv2a : -0.5f * A * d / (1.0f + a * c * d);
    }
  }
  if (S.j2_type) {
    const float a = (float)S.j2_a, inv2a = 1.0f / (2.0f * a);
    for (int j = 0; j < S.n_e; ++j) {
      if (j == e) continue;
      double xj, yj, zj;
      pos.get(j, xj, yj, zj);
      const float dx = (float)(x - xj), dy = (float)(y - yj), dz = (float)(z - zj);
      const float d = sqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1.0e-37f));
      J += S.j2_type == 1 ? 0.5f * d / (1.0f + a * d) : (1.0f - expf(-a * d)) * inv2a;
    }
  }
  return J;
}

// Jastrow (J1+J2) difference J(r') - J(r) for moving electron e from (ox,oy,oz) to (nx,ny,nz)
template <class Pos>
__device__ __forceinline__ double jastrow_delta(const SysDev& S, const Pos& pos, int e, double ox, double oy, double oz,
                                                double nx, double ny, double nz) {
  double dJ = 0.0;
  if (S.j1_type) {
    for (int a = 0; a < S.n_atom; ++a) {
      const double X = S.Rn[3 * a], Y = S.Rn[3 * a + 1], Z = S.Rn[3 * a + 2];
      const double dn_ = sqrt((nx - X) * (nx - X) + (ny - Y) * (ny - Y) + (nz - Z) * (nz - Z));
      const double do_ = sqrt((ox - X) * (ox - X) + (oy - Y) * (oy - Y) + (oz - Z) * (oz - Z));
      const double A = S.j1_A[a], c = S.j1_c[a];
      dJ += j1_f(S.j1_type, S.j1_a, A, c, dn_) - j1_f(S.j1_type, S.j1_a, A, c, do_);
    }
  }
  if (S.j2_type) {
    for (int j = 0; j < S.n_e; ++j) {
      if (j == e) continue;
      double x, y, z;
      pos.get(j, x, y, z);
      const double dn_ = sqrt((nx - x) * (nx - x) + (ny - y) * (ny - y) + (nz - z) * (nz - z));
      const double do_ = sqrt((ox - x) * (ox - x) + (oy - y) * (oy - y) + (oz - z) * (oz - z));
      dJ += j2_f(S.j2_type, S.j2_a, dn_) - j2_f(S.j2_type, S.j2_a, do_);
    }
  }
  return dJ;
}

// The same two functions with the fast reciprocal square root / reciprocal (general path: its mesh kernel is bound by the
// latency of these pair loops when N_e is large; the register kernels are fp64-issue bound and keep the IEEE sequences)
template <class Pos>
__device__ __forceinline__ double jastrow_single_q(const SysDev& S, const Pos& pos, int e, double x, double y, double z) {
  double J = 0.0;
  if (S.j1_type) {
    const double inv2a = 1.0 / (2.0 * S.j1_a);
    for (int a = 0; a < S.n_atom; ++a) {
      const double d = qdist(x - S.Rn[3 * a], y - S.Rn[3 * a + 1], z - S.Rn[3 * a + 2]);
      J += j1_fq(S.j1_type, S.j1_a, inv2a, S.j1_A[a], S.j1_c[a], d);
    }
  }
  if (S.j2_type) {
    const double inv2a = 1.0 / (2.0 * S.j2_a);
    for (int j = 0; j < S.n_e; ++j) {
      if (j == e) continue;
      double xj, yj, zj;
      pos.get(j, xj, yj, zj);
      J += j2_fq(S.j2_type, S.j2_a, inv2a, qdist(x - xj, y - yj, z - zj));
    }
  }
  return J;
}

template <class Pos>
__device__ __forceinline__ double jastrow_delta_q(const SysDev& S, const Pos& pos, int e, double ox, double oy, double oz,
                                                double nx, double ny, double nz) {
  double dJ = 0.0;
  if (S.j1_type) {
    const double inv2a = 1.0 / (2.0 * S.j1_a);
    for (int a = 0; a < S.n_atom; ++a) {
      const double X = S.Rn[3 * a], Y = S.Rn[3 * a + 1], Z = S.Rn[3 * a + 2];
      const double A = S.j1_A[a], c = S.j1_c[a];
      dJ += j1_fq(S.j1_type, S.j1_a, inv2a, A, c, qdist(nx - X, ny - Y, nz - Z)) -
            j1_fq(S.j1_type, S.j1_a, inv2a, A, c, qdist(ox - X, oy - Y, oz - Z));
    }
  }
  if (S.j2_type) {
    const double inv2a = 1.0 / (2.0 * S.j2_a);
    for (int j = 0; j < S.n_e; ++j) {
      if (j == e) continue;
      double x, y, z;
      pos.get(j, x, y, z);
      dJ += j2_fq(S.j2_type, S.j2_a, inv2a, qdist(nx - x, ny - y, nz - z)) - j2_fq(S.j2_type, S.j2_a, inv2a, qdist(ox - x, oy - y, oz - z));
    }
  }
  return dJ;
}


// QE_DEV_MINIMAL (development builds only, never shipped: __graft_entry__.build(minimal=True)): compile the kernels for the
// benchmark shape alone (4 padded orbitals, spherical basis) -- everything else reports QE_ERR_UNSUPPORTED
#ifdef QE_DEV_MINIMAL
#define DISPATCH_NMO_CART(h, CALL)                                                                              \
  do {                                                                                                          \
    if ((h)->b_up.dev.cart != 0 || (h)->nmo_pad != 4) return fail(QE_ERR_UNSUPPORTED, "QE_DEV_MINIMAL build"); \
    CALL(4, false);                                                                                             \
  } while (0)
#define DISPATCH_NMO(h, CALL)                                                            \
  do {                                                                                   \
    if ((h)->nmo_pad != 4) return fail(QE_ERR_UNSUPPORTED, "QE_DEV_MINIMAL build");     \
    CALL(4);                                                                             \
  } while (0)
#else
#define DISPATCH_NMO_CART(h, CALL)                                  \
  do {                                                              \
    const bool cart_ = (h)->b_up.dev.cart != 0;                     \
    switch ((h)->nmo_pad) {                                         \
      case 4: if (cart_) { CALL(4, true); } else { CALL(4, false); } break;   \
      case 8: if (cart_) { CALL(8, true); } else { CALL(8, false); } break;   \
      default: if (cart_) { CALL(16, true); } else { CALL(16, false); } break; \
    }                                                               \
  } while (0)
#define DISPATCH_NMO(h, CALL)       \
  do {                              \
    switch ((h)->nmo_pad) {         \
      case 4: CALL(4); break;       \
      case 8: CALL(8); break;       \
      default: CALL(16); break;     \
    }                               \
  } while (0)
#endif

static inline unsigned nblk(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }
#define CHECK_LAUNCH()                                                                    \
  do {                                                                                    \
    cudaError_t e_ = cudaGetLastError();                                                  \
    if (e_ != cudaSuccess) return fail(QE_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
  } while (0)


// ---- cross-TU entry points -----------------------------------------------------------------------------------------
static inline bool use_wide(const qe_engine* h) { return h->path == 1 || !h->narrow_ok; }
struct WsCarve;
// RNG tables of one qe_mcmc_update call (qe_mcmc.cu) / one qe_lrdmc_project call (qe_walker.cu), carved from `c`
size_t mcmc_draws_bytes(int nw, int nmpm);
int mcmc_draws(qe_engine* h, int nw, int nmpm, uint32_t* keys, WsCarve& c, int** rsel, int** raxis, double** rg, double** rb,
               cudaStream_t st);
size_t lrdmc_draws_bytes(int nw, int nmpm);
int lrdmc_draws(qe_engine* h, int nw, int nmpm, int random_mesh, uint32_t* keys, WsCarve& c, double** rRT, double** ru, cudaStream_t st);
// general path (qe_wide.cu)
int wide_geminal_init(qe_engine* h, int nw, const double* r_up, const double* r_dn, double* G, double* Ginv, double* ln_psi,
                      double* sign, cudaStream_t st);
int wide_local_energy(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                      double* e_L, double* T_elem, double* V_parts, cudaStream_t st);
int wide_move_ratios(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* Ginv, int n_moves,
                     const int32_t* elec_host, const double* r_new, double* det_ratio, double* jas_ratio, cudaStream_t st);
int wide_mcmc_update(qe_engine* h, int nw, double* r_up, double* r_dn, uint32_t* keys, double* G, double* Ginv, int nmpm, double Dt,
                     double epsilon_AS, int32_t* acc, int32_t* rej, cudaStream_t st);
int wide_lrdmc(qe_engine* h, int mode, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys, double E_scf,
               int nmpm, int random_mesh, int non_local_move, double alat, const double* RT_in, double* RT_out, double* V_diag,
               double* V_nondiag, cudaStream_t st);
int wide_lrdmc_tau(qe_engine* h, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys, double tau,
                   int random_mesh, int non_local_move, double alat, int32_t* pc, double* e_L, double* RT_out, cudaStream_t st);
int wide_eval_orbitals(qe_engine* h, int which, int n_pts, const double* r, double* out, cudaStream_t st);
int wide_dln_wf(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* Ginv, double* d_j1, double* d_j2,
                double* d_j3, double* d_lambda, cudaStream_t st);

// workspace carve-up helper
struct WsCarve {
  char* base;
  size_t off = 0;
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = (T*)(base + off);
    off += n * sizeof(T);
    return p;
  }
};

