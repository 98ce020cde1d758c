// Fused per-walker kernel: LRDMC projection (GFMC_n._projection_n, jqmc/jqmc_gfmc.py:4738-5358), its
// observable V_diag / V_nondiag (_compute_V_elements_n, :5360-5627) and the VMC local energy
// (compute_local_energy_fast, jqmc/hamiltonians.py:225-290) share one kernel:
//
//   CTA = 32 walkers (lanes) x NW warps.  Basis / Jastrow / ECP tables are staged in shared memory once per
//   CTA; walker state (positions, running inverse, cached MO value/grad/lap of every electron, ratio weight
//   vectors) lives in shared memory as [item][lane].  Warps take TASKS (a mesh point, an electron, a basis
//   chunk); every lane of a warp executes the same task on its own walker, so control flow is uniform and
//   table reads are warp-broadcast LDS.
//
//   per projection:  P1  ratio weight vectors W[:,e] from the running inverse            (task = electron)
//                    P2  6 N_e kinetic-mesh ratios, N_e*NN*Nv ECP-mesh ratios,
//                        per-electron continuum kinetic energy and potential pieces       (task = point / electron)
//                    P3  warp 0: fixed-node split, diagonal/off-diagonal sums, weight, move selection
//                    P4  value/grad/lap of the moved electron (task = basis chunk), Sherman-Morrison update
#include "qe_common.cuh"

namespace {

struct WalkerArgs {
  int nw, nmpm, mode;  // mode 0: projection, 1: V elements (no move), 2: VMC local energy
  int dlt, ecp_n_chunk;
  double alat, E_scf;
  double* w;
  double* r_up;
  double* r_dn;
  double* Ginv;
  const double* RT_in;  // modes 1, 2: [nw][9]
  double* RT_out;       // mode 0
  double* V_diag;
  double* V_nondiag;
  double* e_L;
  double* T_elem;
  double* V_parts;
  const double* rRT;  // mode 0 draws: [(it*9+c)][nw]
  const double* ru;   //               [it][nw]
  const int* chunk_begin;
  int n_chunk;
};

struct Carve {
  char* p;
  __device__ __forceinline__ Carve(char* base) : p(base) {}
  template <class T>
  __device__ __forceinline__ T* take(size_t n) {
    p = (char*)(((uintptr_t)p + 15) & ~uintptr_t(15));
    T* r = (T*)p;
    p += n * sizeof(T);
    return r;
  }
};

template <class T>
__device__ __forceinline__ const T* stage(Carve& c, const T* src, size_t n, int tid, int nthr) {
  T* dst = c.take<T>(n);
  for (size_t i = tid; i < n; i += nthr) dst[i] = src[i];
  return dst;
}

__device__ __forceinline__ BasisDev stage_basis(const BasisDev& g, Carve& c, int tid, int nthr) {
  BasisDev s = g;
  s.grp_nuc = stage(c, g.grp_nuc, g.n_grp, tid, nthr);
  s.grp_l = stage(c, g.grp_l, g.n_grp, tid, nthr);
  s.grp_sh_begin = stage(c, g.grp_sh_begin, g.n_grp + 1, tid, nthr);
  s.sh_prim_off = stage(c, g.sh_prim_off, g.n_shell + 1, tid, nthr);
  s.sh_slot = stage(c, g.sh_slot, (size_t)g.n_shell * MAXF, tid, nthr);
  s.pr_zc = stage(c, g.pr_zc, g.n_prim, tid, nthr);
  s.Cs = stage(c, g.Cs, (size_t)g.n_ao * g.nmo_pad, tid, nthr);
  return s;
}

__device__ __forceinline__ SysDev stage_sys(const SysDev& g, int nmo, Carve& c, int tid, int nthr) {
  SysDev s = g;
  s.Rn = stage(c, g.Rn, 3 * g.n_atom, tid, nthr);
  s.Zeff = stage(c, g.Zeff, g.n_atom, tid, nthr);
  s.lam_p = stage(c, g.lam_p, (size_t)nmo * nmo, tid, nthr);
  s.lam_u = stage(c, g.lam_u, (size_t)nmo * (g.n_unp > 0 ? g.n_unp : 1), tid, nthr);
  s.j1_A = stage(c, g.j1_A, g.n_atom, tid, nthr);
  s.j1_c = stage(c, g.j1_c, g.n_atom, tid, nthr);
  if (g.ecp_flag) {
    s.ecp_l = stage(c, g.ecp_l, g.n_ecp, tid, nthr);
    s.ecp_z = stage(c, g.ecp_z, g.n_ecp, tid, nthr);
    s.ecp_c = stage(c, g.ecp_c, g.n_ecp, tid, nthr);
    s.ecp_p = stage(c, g.ecp_p, g.n_ecp, tid, nthr);
    s.ecp_lmax_atom = stage(c, g.ecp_lmax_atom, g.n_atom, tid, nthr);
    s.ecp_off = stage(c, g.ecp_off, g.n_atom + 1, tid, nthr);
    s.quad_w = stage(c, g.quad_w, g.Nv, tid, nthr);
    s.quad_g = stage(c, g.quad_g, 3 * g.Nv, tid, nthr);
  }
  return s;
}

size_t table_bytes(const BasisDev& b, const SysDev& s, int nmo) {
  size_t n = 0;
  n += (size_t)(3 * b.n_grp + 1 + b.n_shell + 1) * 4 + (size_t)b.n_shell * MAXF * 2 + (size_t)b.n_prim * 16 + (size_t)b.n_ao * b.nmo_pad * 8;
  n += (size_t)(3 * s.n_atom + 3 * s.n_atom + nmo * nmo + nmo * (s.n_unp > 0 ? s.n_unp : 1)) * 8;
  if (s.ecp_flag) n += (size_t)s.n_ecp * (4 + 24) + (size_t)s.n_atom * 8 + 4 + (size_t)s.Nv * 32;
  return n + 16 * 32;  // alignment slack
}

// electron positions held in shared memory as [e*3+c][lane]
struct PosShared {
  const double* s_r;
  int lane;
  __device__ __forceinline__ void get(int e, double& x, double& y, double& z) const {
    x = s_r[(e * 3 + 0) * 32 + lane];
    y = s_r[(e * 3 + 1) * 32 + lane];
    z = s_r[(e * 3 + 2) * 32 + lane];
  }
};

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
      if (S.ecp_l[kk] == l) vl += S.ecp_c[kk] * pow(d, S.ecp_p[kk]) * exp(-S.ecp_z[kk] * d * d);
    ang = fma(vl / (d * d) * (2 * l + 1), legendre_l(l, cos_t), ang);
  }
  ang_w = ang * S.quad_w[k];
}

template <int NMO, bool CART>
__global__ void __launch_bounds__(512)
k_walker(BasisDev Bu_g, BasisDev Bd_g, SysDev S_g, WalkerArgs P) {
  extern __shared__ __align__(16) char smem_raw[];
  const int lane = threadIdx.x, wid = threadIdx.y, NW = blockDim.y;
  const int tid = wid * 32 + lane, nthr = NW * 32;
  const int w = blockIdx.x * 32 + lane;
  const bool live = w < P.nw;
  const int ww = live ? w : P.nw - 1;

  // ---- stage tables ------------------------------------------------------------------------------
  Carve cv(smem_raw);
  const BasisDev Bu = stage_basis(Bu_g, cv, tid, nthr);
  BasisDev Bd = Bu;
  if (Bd_g.Cs != Bu_g.Cs) Bd.Cs = stage(cv, Bd_g.Cs, (size_t)Bd_g.n_ao * Bd_g.nmo_pad, tid, nthr);
  const SysDev S = stage_sys(S_g, NMO, cv, tid, nthr);
  const int N = S.n_up, Nd = S.n_dn, Ne = S.n_e, NN2 = N * N;
  const int nch = P.n_chunk;
  const int n_kin = P.mode == 2 ? 0 : 6 * Ne;
  const int n_ecp = S.ecp_flag ? Ne * S.NN * S.Nv : 0;
  const int NPT = n_kin + n_ecp;

  double* s_r = cv.take<double>((size_t)Ne * 3 * 32);
  double* s_Gi = cv.take<double>((size_t)NN2 * 32);
  double* s_phi = cv.take<double>((size_t)Ne * 5 * NMO * 32);  // [(e*5+q)*NMO+mo]
  double* s_W = cv.take<double>((size_t)Ne * NMO * 32);
  double* s_p = cv.take<double>((size_t)(NPT > 0 ? NPT : 1) * 32);
  double* s_j = cv.take<double>((size_t)(n_ecp > 0 ? n_ecp : 1) * 32);
  double* s_el = cv.take<double>((size_t)Ne * 5 * 32);  // [e*5 + {ke, ei, eid, loc, ee}]
  double* s_part = cv.take<double>((size_t)2 * nch * NMO * 32);  // double buffer of per-chunk partials, one component at a time
  double* s_stage = cv.take<double>((size_t)5 * NMO * 32);
  double* s_misc = cv.take<double>((size_t)16 * 32);  // 0..8 RT, 9..11 new position, 12 selected electron
#define SR(e, c) s_r[((e) * 3 + (c)) * 32 + lane]
#define SGI(i, j) s_Gi[((i) * N + (j)) * 32 + lane]
#define SPHI(e, q, mo) s_phi[(((e) * 5 + (q)) * NMO + (mo)) * 32 + lane]
#define SW(e, mo) s_W[((e) * NMO + (mo)) * 32 + lane]
#define SEL(e, i) s_el[((e) * 5 + (i)) * 32 + lane]
#define SMISC(i) s_misc[(i) * 32 + lane]

  for (int idx = wid; idx < Ne * 3; idx += NW) {
    const int e = idx / 3, c = idx % 3;
    SR(e, c) = e < N ? P.r_up[((size_t)ww * N + e) * 3 + c] : P.r_dn[((size_t)ww * Nd + (e - N)) * 3 + c];
  }
  for (int idx = wid; idx < NN2; idx += NW) s_Gi[idx * 32 + lane] = P.Ginv[(size_t)ww * NN2 + idx];
  if (P.mode != 0 && wid == 0)
    for (int c = 0; c < 9; ++c) SMISC(c) = P.RT_in ? P.RT_in[(size_t)ww * 9 + c] : (c % 4 == 0 ? 1.0 : 0.0);
  __syncthreads();

  // ---- value/grad/lap of the MOs at every electron (cache): task = electron ----------------------------
  for (int e = wid; e < Ne; e += NW) {
    SinkMO5<NMO> sink;
    sink.init(e < N ? Bu.Cs : Bd.Cs);
    eval_vgl<CART>(Bu, S.Rn, SR(e, 0), SR(e, 1), SR(e, 2), 0, Bu.n_grp, sink);
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
      for (int mo = 0; mo < NMO; ++mo) SPHI(e, q, mo) = sink.acc[q][mo];
  }
  __syncthreads();

  double w_L = (P.mode == 0) ? P.w[ww] : 1.0;
  double diag = 0.0, nondiag = 0.0;
  const double a2 = P.alat * P.alat;
  const int n_it = P.mode == 0 ? P.nmpm : 1;

  for (int it = 0; it < n_it; ++it) {
    // ---- P1: ratio weight vectors ---------------------------------------------------------------------
    if (P.mode == 0 && wid == NW - 1)
      for (int c = 0; c < 9; ++c) SMISC(c) = P.rRT[((size_t)it * 9 + c) * P.nw + ww];
    for (int e = wid; e < Ne; e += NW) {
      double Wv[NMO];
      if (e < N) {
        double y[NMO];
#pragma unroll
        for (int b = 0; b < NMO; ++b) {
          double s = 0;
          for (int j = 0; j < Nd; ++j) s = fma(SPHI(N + j, 0, b), SGI(j, e), s);
          y[b] = s;
        }
#pragma unroll
        for (int a = 0; a < NMO; ++a) {
          double s = 0;
#pragma unroll
          for (int b = 0; b < NMO; ++b) s = fma(S.lam_p[a * NMO + b], y[b], s);
          for (int k = 0; k < S.n_unp; ++k) s = fma(S.lam_u[a * S.n_unp + k], SGI(Nd + k, e), s);
          Wv[a] = s;
        }
      } else {
        const int j = e - N;
        double y[NMO];
#pragma unroll
        for (int a = 0; a < NMO; ++a) {
          double s = 0;
          for (int i = 0; i < N; ++i) s = fma(SPHI(i, 0, a), SGI(j, i), s);
          y[a] = s;
        }
#pragma unroll
        for (int b = 0; b < NMO; ++b) {
          double s = 0;
#pragma unroll
          for (int a = 0; a < NMO; ++a) s = fma(y[a], S.lam_p[a * NMO + b], s);
          Wv[b] = s;
        }
      }
#pragma unroll
      for (int mo = 0; mo < NMO; ++mo) SW(e, mo) = Wv[mo];
    }
    __syncthreads();

    // ---- P2: mesh ratios and per-electron terms ---------------------------------------------------------
    PosShared pos{s_r, lane};
    double rt[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) rt[c] = SMISC(c);
    const int NT = NPT + Ne;
    for (int t = wid; t < NT; t += NW) {
      if (t < NPT) {
        int e;
        double px, py, pz, x, y, z, angw = 0.0;
        const bool kin = t < n_kin;
        if (kin) {
          e = t / 6;
          const int s = t % 6, ax = s >> 1;
          const double sg = (s & 1) ? -P.alat : P.alat;
          pos.get(e, x, y, z);
          px = x + sg * rt[3 * ax];
          py = y + sg * rt[3 * ax + 1];
          pz = z + sg * rt[3 * ax + 2];
        } else {
          const int pt = t - n_kin;
          const int k = pt % S.Nv, nn = (pt / S.Nv) % S.NN;
          e = pt / (S.Nv * S.NN);
          pos.get(e, x, y, z);
          ecp_point(S, rt, x, y, z, nn, k, px, py, pz, angw, true);
        }
        SinkMO<NMO> sink;
        sink.init(e < N ? Bu.Cs : Bd.Cs);
        eval_val<CART>(Bu, S.Rn, px, py, pz, 0, Bu.n_grp, sink);
        double ratio = 0.0;
#pragma unroll
        for (int mo = 0; mo < NMO; ++mo) ratio = fma(sink.acc[mo], SW(e, mo), ratio);
        const double jr = exp(jastrow_delta(S, pos, e, x, y, z, px, py, pz));
        if (kin) {
          s_p[t * 32 + lane] = -1.0 / (2.0 * a2) * (ratio * jr);
        } else {
          s_p[t * 32 + lane] = P.dlt ? angw * ratio : angw * (ratio * jr);
          s_j[(t - n_kin) * 32 + lane] = jr;
        }
      } else {
        // per-electron: continuum kinetic energy, bare / discretised el-ion, ECP local, el-el (pairs j > e)
        const int e = t - NPT;
        double x, y, z;
        pos.get(e, x, y, z);
        double gD[3] = {0, 0, 0}, lD = 0;
#pragma unroll
        for (int mo = 0; mo < NMO; ++mo) {
          const double wv = SW(e, mo);
          gD[0] = fma(SPHI(e, 1, mo), wv, gD[0]);
          gD[1] = fma(SPHI(e, 2, mo), wv, gD[1]);
          gD[2] = fma(SPHI(e, 3, mo), wv, gD[2]);
          lD = fma(SPHI(e, 4, mo), wv, lD);
        }
        lD -= gD[0] * gD[0] + gD[1] * gD[1] + gD[2] * gD[2];
        double gJ[3] = {0, 0, 0}, lJ = 0, ei = 0, eid = 0, loc = 0, ee = 0;
        const double eps = 1.0e-12;
        for (int a = 0; a < S.n_atom; ++a) {
          const double dx = x - S.Rn[3 * a], dy = y - S.Rn[3 * a + 1], dz = z - S.Rn[3 * a + 2];
          const double d = sqrt(dx * dx + dy * dy + dz * dz);
          ei -= S.Zeff[a] / d;
          eid -= S.Zeff[a] / fmax(d, P.alat);
          if (S.j1_type) {
            const double rs = fmax(d, eps);
            const double A = S.j1_A[a], c = S.j1_c[a], aa = S.j1_a;
            double fp;
            if (S.j1_type == 1) {
              const double ex = exp(-aa * c * rs);
              fp = -A * (c * 0.5) * ex;
              lJ += A * (aa * c * c * 0.5) * ex - A * c * ex / rs;
            } else {
              const double den = 1.0 + aa * c * rs;
              fp = -A / (2.0 * den * den);
              lJ += A * aa * c / (den * den * den) + 2.0 * fp / rs;
            }
            const double s = fp / rs;
            gJ[0] = fma(s, dx, gJ[0]);
            gJ[1] = fma(s, dy, gJ[1]);
            gJ[2] = fma(s, dz, gJ[2]);
          }
          if (S.ecp_flag) {
            const int lloc = S.ecp_lmax_atom[a];
            double s = 0.0;
            for (int k = S.ecp_off[a]; k < S.ecp_off[a + 1]; ++k)
              if (S.ecp_l[k] == lloc) s += S.ecp_c[k] * pow(d, S.ecp_p[k]) * exp(-S.ecp_z[k] * d * d);
            loc += s / (d * d);
          }
        }
        for (int j = 0; j < Ne; ++j) {
          if (j == e) continue;
          double x2, y2, z2;
          pos.get(j, x2, y2, z2);
          const double dx = x - x2, dy = y - y2, dz = z - z2;
          const double d = sqrt(dx * dx + dy * dy + dz * dz);
          if (j > e) ee += 1.0 / d;
          if (S.j2_type) {
            const double rs = fmax(d, eps), aa = S.j2_a;
            double fp;
            if (S.j2_type == 1) {
              const double den = 1.0 + aa * rs;
              fp = 0.5 / (den * den);
              lJ += -aa / (den * den * den) + 2.0 * fp / rs;
            } else {
              const double ex = exp(-aa * rs);
              fp = 0.5 * ex;
              lJ += -(aa * 0.5) * ex + 2.0 * fp / rs;
            }
            const double s = fp / rs;
            gJ[0] = fma(s, dx, gJ[0]);
            gJ[1] = fma(s, dy, gJ[1]);
            gJ[2] = fma(s, dz, gJ[2]);
          }
        }
        const double gx = gJ[0] + gD[0], gy = gJ[1] + gD[1], gz = gJ[2] + gD[2];
        SEL(e, 0) = -0.5 * (lJ + lD + gx * gx + gy * gy + gz * gz);
        SEL(e, 1) = ei;
        SEL(e, 2) = eid;
        SEL(e, 3) = loc;
        SEL(e, 4) = ee;
      }
    }
    __syncthreads();

    // ---- P3: assemble (warp 0) ---------------------------------------------------------------------------
    if (wid == 0) {
      if (P.mode == 2) {
        double T = 0, vbare = S.v_ion_ion, vl = 0, vnl = 0;
        for (int e = 0; e < Ne; ++e) {
          T += SEL(e, 0);
          vbare += SEL(e, 1) + SEL(e, 4);
          vl += SEL(e, 3);
          if (P.T_elem && live) P.T_elem[(size_t)w * Ne + e] = SEL(e, 0);
        }
        for (int k = 0; k < n_ecp; ++k) vnl += s_p[k * 32 + lane];
        if (live) {
          P.e_L[w] = T + (vbare + (vl + vnl));
          if (P.V_parts) {
            P.V_parts[(size_t)w * 4 + 0] = vbare;
            P.V_parts[(size_t)w * 4 + 1] = vl;
            P.V_parts[(size_t)w * 4 + 2] = vnl;
            P.V_parts[(size_t)w * 4 + 3] = 0.0;
          }
        }
      } else {
        // fixed-node split and regularised diagonal (jqmc/jqmc_gfmc.py:4829-5053)
        const double diag_kin = 3.0 / (2.0 * a2) * Ne;
        double sum_kinFN = 0, SP_kin = 0, sum_opt_up = 0, sum_opt_dn = 0, ee = 0, loc = 0;
        for (int e = 0; e < Ne; ++e) {
          bool flip = false;
          double nd = 0;
          for (int s = 0; s < 6; ++s) {
            const double v = s_p[(6 * e + s) * 32 + lane];
            flip = flip || (v >= 0.0);
            nd += v + 1.0 / (4.0 * a2);
            const double fn = fmin(v, 0.0);
            sum_kinFN += fn;
            SP_kin += fmax(v, 0.0);
            s_p[(6 * e + s) * 32 + lane] = fn;
          }
          const double zv = SEL(e, 1) + SEL(e, 0) - nd;
          const double eib = S.ecp_flag ? SEL(e, 1) : SEL(e, 2);
          const double opt = flip ? fmax(zv, eib) : zv;
          if (e < N) sum_opt_up += opt; else sum_opt_dn += opt;
          ee += SEL(e, 4);
          loc += SEL(e, 3);
        }
        const double disc_bare = ee + S.v_ion_ion + sum_opt_up + sum_opt_dn;
        double sum_eFN = 0, SP_e = 0;
        for (int k = 0; k < n_ecp; ++k) {
          const double v = s_p[(n_kin + k) * 32 + lane];
          SP_e += fmax(v, 0.0);
          double fn = fmin(v, 0.0);
          if (P.dlt) fn *= s_j[k * 32 + lane];
          sum_eFN += fn;
          s_p[(n_kin + k) * 32 + lane] = fn;
        }
        nondiag = sum_kinFN + sum_eFN;
        diag = S.ecp_flag ? diag_kin + disc_bare + loc + SP_kin + SP_e : diag_kin + disc_bare + SP_kin;
        if (P.mode == 0) {
          const double b_x = 1.0 / (diag - P.E_scf) * (-nondiag);
          w_L *= b_x;
          double tot = 0;
          for (int k = 0; k < NPT; ++k) tot += s_p[k * 32 + lane];
          const double u = P.ru[(size_t)it * P.nw + ww];
          int ksel = NPT - 1;
          double c = 0;
          for (int k = 0; k < NPT; ++k) {
            c += s_p[k * 32 + lane] / tot;
            if (c >= u) {
              ksel = k;
              break;
            }
          }
          int e;
          double x, y, z, px, py, pz, dummy;
          if (ksel < n_kin) {
            e = ksel / 6;
            const int s = ksel % 6, ax = s >> 1;
            const double sg = (s & 1) ? -P.alat : P.alat;
            pos.get(e, x, y, z);
            px = x + sg * rt[3 * ax];
            py = y + sg * rt[3 * ax + 1];
            pz = z + sg * rt[3 * ax + 2];
          } else {
            const int pt = ksel - n_kin;
            const int k = pt % S.Nv, nn = (pt / S.Nv) % S.NN;
            e = pt / (S.Nv * S.NN);
            pos.get(e, x, y, z);
            ecp_point(S, rt, x, y, z, nn, k, px, py, pz, dummy, false);
          }
          SMISC(9) = px;
          SMISC(10) = py;
          SMISC(11) = pz;
          SMISC(12) = (double)e;
        }
      }
    }
    if (P.mode != 0) break;
    __syncthreads();

    // ---- P4: refresh the moved electron, Sherman-Morrison ---------------------------------------------------
    const int es = (int)SMISC(12);
    const double nx = SMISC(9), ny = SMISC(10), nz = SMISC(11);
    {
      SinkMO5<NMO> sink;
      if (wid < nch) {
        sink.init(es < N ? Bu.Cs : Bd.Cs);
        eval_vgl<CART>(Bu, S.Rn, nx, ny, nz, P.chunk_begin[wid], P.chunk_begin[wid + 1], sink);
      }
      // deterministic reduction over chunks, one component at a time through a double buffer
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        double* buf = s_part + (size_t)(q & 1) * nch * NMO * 32;
        if (wid < nch) {
#pragma unroll
          for (int mo = 0; mo < NMO; ++mo) buf[(wid * NMO + mo) * 32 + lane] = sink.acc[q][mo];
        }
        __syncthreads();
        for (int mo = wid; mo < NMO; mo += NW) {
          double s = 0;
          for (int c = 0; c < nch; ++c) s += buf[(c * NMO + mo) * 32 + lane];
          s_stage[(q * NMO + mo) * 32 + lane] = s;
        }
      }
      __syncthreads();
    }
    if (wid == 0) {
      double pn[NMO], po[NMO];
#pragma unroll
      for (int mo = 0; mo < NMO; ++mo) {
        const double s = s_stage[mo * 32 + lane];
        pn[mo] = s - SPHI(es, 0, mo);  // phi_new - phi_old  (row/column DIFFERENCE is linear in it)
        po[mo] = s;
      }
      if (es < N) {
        const int k = es;
        double t[NMO], vvec[16];
#pragma unroll
        for (int b = 0; b < NMO; ++b) {
          double s = 0;
#pragma unroll
          for (int a = 0; a < NMO; ++a) s = fma(pn[a], S.lam_p[a * NMO + b], s);
          t[b] = s;
        }
        double acc = 0;
        for (int j = 0; j < Nd; ++j) {
          double s = 0;
#pragma unroll
          for (int b = 0; b < NMO; ++b) s = fma(t[b], SPHI(N + j, 0, b), s);
          vvec[j] = s;
          acc = fma(s, SGI(j, k), acc);
        }
        for (int q = 0; q < S.n_unp; ++q) {
          double s = 0;
#pragma unroll
          for (int a = 0; a < NMO; ++a) s = fma(pn[a], S.lam_u[a * S.n_unp + q], s);
          vvec[Nd + q] = s;
          acc = fma(s, SGI(Nd + q, k), acc);
        }
        const double invD = 1.0 / (1.0 + acc);
        double col[16], vt[16];
        for (int i = 0; i < N; ++i) col[i] = SGI(i, k);
        for (int jp = 0; jp < N; ++jp) {
          double s = 0;
          for (int j = 0; j < N; ++j) s = fma(vvec[j], SGI(j, jp), s);
          vt[jp] = s;
        }
        for (int i = 0; i < N; ++i)
          for (int jp = 0; jp < N; ++jp) SGI(i, jp) = SGI(i, jp) - (col[i] * vt[jp]) * invD;
      } else {
        const int k = es - N;
        double t[NMO], uvec[16];
#pragma unroll
        for (int a = 0; a < NMO; ++a) {
          double s = 0;
#pragma unroll
          for (int b = 0; b < NMO; ++b) s = fma(S.lam_p[a * NMO + b], pn[b], s);
          t[a] = s;
        }
        for (int i = 0; i < N; ++i) {
          double s = 0;
#pragma unroll
          for (int a = 0; a < NMO; ++a) s = fma(SPHI(i, 0, a), t[a], s);
          uvec[i] = s;
        }
        double au[16], row[16];
        for (int i = 0; i < N; ++i) {
          double s = 0;
          for (int j = 0; j < N; ++j) s = fma(SGI(i, j), uvec[j], s);
          au[i] = s;
        }
        const double invD = 1.0 / (1.0 + au[k]);
        for (int j = 0; j < N; ++j) row[j] = SGI(k, j);
        for (int i = 0; i < N; ++i)
          for (int j = 0; j < N; ++j) SGI(i, j) = SGI(i, j) - (au[i] * row[j]) * invD;
      }
      SR(es, 0) = nx;
      SR(es, 1) = ny;
      SR(es, 2) = nz;
      (void)po;
    }
    __syncthreads();
    for (int item = wid; item < 5 * NMO; item += NW) s_phi[(es * 5 * NMO + item) * 32 + lane] = s_stage[item * 32 + lane];
    __syncthreads();
  }

  // ---- write back -------------------------------------------------------------------------------------------
  if (live && P.mode != 2) {
    if (wid == 0) {
      P.V_diag[w] = diag;
      P.V_nondiag[w] = nondiag;
      if (P.mode == 0) {
        P.w[w] = w_L;
        for (int c = 0; c < 9; ++c) P.RT_out[(size_t)w * 9 + c] = SMISC(c);
      }
    }
    if (P.mode == 0) {
      for (int idx = wid; idx < Ne * 3; idx += NW) {
        const int e = idx / 3, c = idx % 3;
        if (e < N) P.r_up[((size_t)w * N + e) * 3 + c] = SR(e, c);
        else P.r_dn[((size_t)w * Nd + (e - N)) * 3 + c] = SR(e, c);
      }
      for (int idx = wid; idx < NN2; idx += NW) P.Ginv[(size_t)w * NN2 + idx] = s_Gi[idx * 32 + lane];
    }
  }
}

// key chain of the projection loop: two splits per projection (jqmc/jqmc_gfmc.py:5275-5283). thread = walker
__global__ void k_lrdmc_keychain(int nw, int nmpm, uint32_t* __restrict__ keys, uint2* __restrict__ sub) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  Key k{keys[2 * w], keys[2 * w + 1]};
  for (int p = 0; p < nmpm * 2; ++p) {
    Key s;
    rng_split(k, s);
    sub[(size_t)p * nw + w] = make_uint2(s.a, s.b);
  }
  keys[2 * w] = k.a;
  keys[2 * w + 1] = k.b;
}
// rotation matrix (R^T, row-major) and move uniform of every projection: thread = (projection, walker)
__global__ void k_lrdmc_draws(int nw, int nmpm, int random_mesh, const uint2* __restrict__ sub, double* __restrict__ rRT,
                              double* __restrict__ ru) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nmpm * nw) return;
  const int w = (int)(t % nw), p = (int)(t / nw);
  const uint2 rk = sub[((size_t)p * 2) * nw + w], mk = sub[((size_t)p * 2 + 1) * nw + w];
  double al = 0, be = 0, ga = 0;
  if (random_mesh) {
    const double two_pi = 6.283185307179586;
    const Key k{rk.x, rk.y};
    al = rng_uniform_bits(rng_bits64(k, 0u), -two_pi, two_pi);
    be = rng_uniform_bits(rng_bits64(k, 1u), -two_pi, two_pi);
    ga = rng_uniform_bits(rng_bits64(k, 2u), -two_pi, two_pi);
  }
  double sa, ca, sb, cb, sg, cg;
  sincos(al, &sa, &ca);
  sincos(be, &sb, &cb);
  sincos(ga, &sg, &cg);
  const double R[9] = {cb * cg, cg * sa * sb - ca * sg, sa * sg + ca * cg * sb, cb * sg, ca * cg + sa * sb * sg,
                       ca * sb * sg - cg * sa, -sb, cb * sa, ca * cb};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) rRT[((size_t)p * 9 + i * 3 + j) * nw + w] = R[j * 3 + i];
  ru[t] = rng_uniform_bits(rng_bits64(Key{mk.x, mk.y}), 0.0, 1.0);
}

int launch_walker(qe_engine* h, WalkerArgs& A, cudaStream_t st, int kid) {
  const SysDev& S = h->sys;
  const int P = h->nmo_pad;
  if (S.n_up > 16) return fail(QE_ERR_UNSUPPORTED, "more than 16 electrons per spin is not implemented in this build");
  const int nch = h->n_chunk_mc;
  const int NW = 16;
  A.chunk_begin = h->d_chunk_mc;
  A.n_chunk = nch;
  const int Ne = S.n_e;
  const int n_kin = A.mode == 2 ? 0 : 6 * Ne;
  const int n_ecp = S.ecp_flag ? Ne * S.NN * S.Nv : 0;
  size_t per_lane = (size_t)Ne * 3 + (size_t)S.n_up * S.n_up + (size_t)Ne * 5 * P + (size_t)Ne * P + std::max(1, n_kin + n_ecp) +
                    std::max(1, n_ecp) + (size_t)Ne * 5 + (size_t)2 * nch * P + 5 * P + 16;
  size_t smem = table_bytes(h->b_up.dev, S, P) + (h->b_dn.dev.Cs != h->b_up.dev.Cs ? (size_t)h->b_dn.dev.n_ao * P * 8 + 16 : 0) +
                per_lane * 32 * 8 + 16 * 12;
  if (smem > 227 * 1024) return fail(QE_ERR_UNSUPPORTED, "system too large for the fused walker kernel (shared memory)");
  dim3 block(32, NW);
  {
    LaunchScope ls_(h, kid, st);
#define CALL(NMO, CART)                                                                                        \
  do {                                                                                                         \
    CUDA_TRY(cudaFuncSetAttribute(k_walker<NMO, CART>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_walker<NMO, CART><<<nblk(A.nw, 32), block, smem, st>>>(h->b_up.dev, h->b_dn.dev, S, A);                  \
  } while (0)
    DISPATCH_NMO_CART(h, CALL);
#undef CALL
  }
  CHECK_LAUNCH();
  return QE_OK;
}

}  // namespace

extern "C" int qe_lrdmc_project(qe_engine* h, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys,
                                double E_scf, int nmpm, int random_discretized_mesh, int non_local_move, double alat, double* RT,
                                double* V_diag, double* V_nondiag, void* stream) {
  if (!h || nw <= 0 || nmpm <= 0 || !w || !r_up || !Ginv || !keys || !RT || !V_diag || !V_nondiag || (!r_dn && h->sys.n_dn > 0))
    return fail(QE_ERR_INVALID, "qe_lrdmc_project: bad argument");
  if (!(alat > 0)) return fail(QE_ERR_INVALID, "qe_lrdmc_project: alat must be positive");
  if (non_local_move != 0 && non_local_move != 1) return fail(QE_ERR_INVALID, "qe_lrdmc_project: non_local_move must be 0 (tmove) or 1 (dltmove)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_draw = (size_t)nmpm * nw;
  int rc = ensure_ws(h, n_draw * (2 * 8 + 9 * 8 + 8) + 4096);
  if (rc) return rc;
  WsCarve c{(char*)h->ws};
  uint2* sub = c.take<uint2>(n_draw * 2);
  double* rRT = c.take<double>(n_draw * 9);
  double* ru = c.take<double>(n_draw);
  {
    LaunchScope ls_(h, K_KEYCHAIN, st);
    k_lrdmc_keychain<<<nblk(nw, 64), 64, 0, st>>>(nw, nmpm, keys, sub);
  }
  CHECK_LAUNCH();
  {
    LaunchScope ls_(h, K_DRAWS, st);
    k_lrdmc_draws<<<nblk((long long)n_draw, 128), 128, 0, st>>>(nw, nmpm, random_discretized_mesh, sub, rRT, ru);
  }
  CHECK_LAUNCH();
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
  A.rRT = rRT;
  A.ru = ru;
  return launch_walker(h, A, st, K_LRDMC_PROJ);
}

extern "C" int qe_lrdmc_velements(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                                  int non_local_move, double alat, double* V_diag, double* V_nondiag, void* stream) {
  if (!h || nw <= 0 || !r_up || !Ginv || !V_diag || !V_nondiag || (!r_dn && h->sys.n_dn > 0))
    return fail(QE_ERR_INVALID, "qe_lrdmc_velements: bad argument (Ginv is required: call qe_geminal_init first)");
  if (!(alat > 0)) return fail(QE_ERR_INVALID, "qe_lrdmc_velements: alat must be positive");
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
  return launch_walker(h, A, (cudaStream_t)stream, K_LRDMC);
}

// fused local energy (mode 2); returns QE_ERR_UNSUPPORTED when the system does not fit, so that the caller
// (qe_local_energy) can fall back to the staged kernels.
int qe_local_energy_fused(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                          double* e_L, double* T_elem, double* V_parts, cudaStream_t st) {
  WalkerArgs A{};
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
  return launch_walker(h, A, st, K_EL_FUSED);
}
