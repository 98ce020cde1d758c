// Metropolis kernels of the register family (k_mcmc: with AS regularisation; k_mcmc2: warp-parallel decision / update stages) and
// their launch function.  The instantiations live in qe_mcmc_i_*.cu, one translation unit per orbital padding, so that each goes
// through a single-threaded (deterministic) ptxas and the files compile in parallel; qe_mcmc.cu holds the entry points.
#pragma once
#include "qe_common.cuh"

// =================================================================================================
// Metropolis kernel (jqmc/jqmc_mcmc.py:4278-4533).  CTA = 32 walkers (lanes) x (n_chunk + 1) warps.
//   warps 0..n_chunk-1 : AO/MO evaluation of one basis chunk at the proposed position
//   warp  n_chunk      : proposal bookkeeping, T_ratio and Jastrow ratio
//   warp  0            : determinant ratio, Sherman-Morrison update, AS factor, accept/reject
// Walker state lives in shared memory as [item][lane].
// =================================================================================================
struct McmcArgs {
  int nw, nmpm, n_chunk;
  double Dt, eps_AS;
  double* r_up;
  double* r_dn;
  double* G;
  double* Ginv;
  int* acc;
  int* rej;
  const int* rsel;
  const int* raxis;
  const double* rg;
  const double* rb;
  int off_cseg, off_cbeg;
  int chunk_warp[16];  // k_mcmc2: chunk evaluated by warp w (HostBasis::chunk_warp)
  long long* clk;      // optional [12] per-phase cycle counters (qe_phase_clocks), thread 0 of every CTA
};

template <int NMO, bool CART, bool MIXED>
__global__ void __launch_bounds__(512)
k_mcmc(BasisDev B, SysDev S, McmcArgs P) {
  extern __shared__ __align__(16) double sm_all[];
  const int lane = threadIdx.x, wid = threadIdx.y;
  // basis image (shells, primitives, MO coefficients) staged in shared memory: every table read of the AO phase is a
  // warp-broadcast LDS instead of an L1-cached global load on the dependent chain shell -> primitives -> coefficients
  const char* tab = (const char*)sm_all;
  {
    const int4* src = (const int4*)B.g;
    int4* dst = (int4*)sm_all;
    for (int i = wid * 32 + lane; i < B.bytes / 16; i += 32 * (P.n_chunk + 1)) dst[i] = src[i];
  }
  double* sm = sm_all + B.bytes / 8;
  const int w = blockIdx.x * 32 + lane;
  const bool live = w < P.nw;
  const int ww = live ? w : P.nw - 1;  // dead lanes shadow the last walker (no stores)
  const int N = S.n_up, Nd = S.n_dn, Ne = S.n_e, NN2 = N * N;
  const int nch = P.n_chunk;
  const int* cbeg = (const int*)(tab + P.off_cbeg);
  // shared-memory carve-up (all [item][32])
  double* s_r = sm;                       // Ne*3
  double* s_G = s_r + Ne * 3 * 32;        // N*N
  double* s_Gi = s_G + NN2 * 32;          // N*N
  double* s_phi = s_Gi + NN2 * 32;        // Ne*NMO   (orbital values at the electrons: [e][mo])
  double* s_part = s_phi + Ne * NMO * 32; // nch*NMO
  double* s_TJ = s_part + nch * NMO * 32; // 1 + (nch+1): T_ratio, per-warp partial Jastrow exponent differences
  double* s_fl = s_TJ + (nch + 2) * 32;   // Ne + 1: proposal width factor f of every electron at its current position, f'
#define SR(e, c) s_r[((e) * 3 + (c)) * 32 + lane]
#define SG(i, j) s_G[((i) * N + (j)) * 32 + lane]
#define SGI(i, j) s_Gi[((i) * N + (j)) * 32 + lane]
#define SPHI(e, mo) s_phi[((e) * NMO + (mo)) * 32 + lane]
#define SPART(c, mo) s_part[((c) * NMO + (mo)) * 32 + lane]

  // ---- load state -------------------------------------------------------------------------------
  const int tid = wid * 32 + lane, nthr = 32 * (nch + 1);
  for (int idx = wid; idx < Ne * 3; idx += nch + 1) {
    const int e = idx / 3, c = idx % 3;
    SR(e, c) = e < N ? P.r_up[((size_t)ww * N + e) * 3 + c] : P.r_dn[((size_t)ww * Nd + (e - N)) * 3 + c];
  }
  for (int idx = wid; idx < NN2; idx += nch + 1) {
    s_G[idx * 32 + lane] = P.G[(size_t)ww * NN2 + idx];
    s_Gi[idx * 32 + lane] = P.Ginv[(size_t)ww * NN2 + idx];
  }
  (void)tid;
  (void)nthr;
  __syncthreads();
  // f = (1 + Z^2 d) / (Z^2 (1 + d)) with the nearest nucleus (jqmc/jqmc_mcmc.py:4340-4357) for every electron; an accepted
  // move replaces the electron's entry by the f' of the proposal, which is the same function of the same position
  for (int e = wid; e < Ne; e += nch + 1) {
    double dist;
    const int ia = nearest_atom(S.Rn, S.n_atom, SR(e, 0), SR(e, 1), SR(e, 2), 0, &dist);
    const double Zc = S.Zeff[ia];
    s_fl[e * 32 + lane] = 1.0 / (Zc * Zc) * (1.0 + Zc * Zc * dist) / (1.0 + dist);
  }

  // ---- orbital values at every electron (cache for the row/column rebuild) -----------------------
  for (int e = 0; e < Ne; ++e) {
    if (wid < nch) {
      SinkMO<NMO> sink;
      sink.init(tab + (e < N ? B.off_C : B.off_C2));
      eval_val<CART, QE_LMAX>(tab, B, P.off_cseg, SR(e, 0), SR(e, 1), SR(e, 2), cbeg[wid], cbeg[wid + 1], sink);
#pragma unroll
      for (int mo = 0; mo < NMO; ++mo) SPART(wid, mo) = sink.acc[mo];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
      for (int mo = 0; mo < NMO; ++mo) {
        double s = 0;
        for (int c = 0; c < nch; ++c) s += SPART(c, mo);
        SPHI(e, mo) = s;
      }
    }
    __syncthreads();
  }

  int n_acc = 0, n_rej = 0;
  double R_AS_cur = 1.0;
  if (wid == 0 && P.eps_AS > 0.0) {
    double F = 0, Smin = 1e300;
    for (int i = 0; i < NN2; ++i) F = fma(s_Gi[i * 32 + lane], s_Gi[i * 32 + lane], F);
    for (int i = 0; i < N; ++i) {
      double r = 0, c = 0;
      for (int j = 0; j < N; ++j) {
        r = fma(SG(i, j), SG(i, j), r);
        c = fma(SG(j, i), SG(j, i), c);
      }
      Smin = fmin(Smin, fmin(r, c));
    }
    const double SF = Smin * F;
    R_AS_cur = SF > 0.0 ? pow(SF, -0.375) : 0.0;
  }

  // draws of the next proposal are fetched one iteration ahead (their L2 latency hides behind phases B and C)
  int ke_n = 0, axis_n = 0;
  double rg_n = 0.0, rb_n = 0.0;
  if (P.nmpm > 0) {
    ke_n = P.rsel[ww];
    axis_n = P.raxis[ww];
    rg_n = P.rg[ww];
    rb_n = P.rb[ww];
  }
  for (int it = 0; it < P.nmpm; ++it) {
    // ---- phase A: proposal (every thread, redundantly; lane = walker) ----------------------------
    const int ke = ke_n, axis = axis_n;
    const double rg_c = rg_n, rb_c = rb_n;
    if (it + 1 < P.nmpm) {
      const size_t rnext = (size_t)(it + 1) * P.nw + ww;
      ke_n = P.rsel[rnext];
      axis_n = P.raxis[rnext];
      rg_n = P.rg[rnext];
      rb_n = P.rb[rnext];
    }
    const bool up = ke < N;
    const double ox = SR(ke, 0), oy = SR(ke, 1), oz = SR(ke, 2);
    const double f_l = s_fl[ke * 32 + lane];
    const double g = rg_c * (f_l * P.Dt);
    double nx = ox, ny = oy, nz = oz;
    if (axis == 0) nx = ox + g;
    else if (axis == 1) ny = oy + g;
    else nz = oz + g;

    // ---- phase B ---------------------------------------------------------------------------------
    // The J1/J2 terms of J(r') - J(r) (one per nucleus / per other electron) are dealt out over all warps, so that no
    // single warp carries the whole pair loop (it used to bound this phase: profiles/r01_mcmc_v6.md)
    struct PosS {
      const double* s_r;
      int lane;
      __device__ __forceinline__ void get(int e, double& x, double& y, double& z) const {
        x = s_r[(e * 3 + 0) * 32 + lane];
        y = s_r[(e * 3 + 1) * 32 + lane];
        z = s_r[(e * 3 + 2) * 32 + lane];
      }
    } pos{s_r, lane};
    const int n_j1 = S.j1_type ? S.n_atom : 0, n_jt = n_j1 + (S.j2_type ? Ne : 0);
    auto jastrow_terms = [&]() {
      double dJ = 0.0;
      for (int t = wid; t < n_jt; t += nch + 1) {
        if (t < n_j1) {
          const double X = S.Rn[3 * t], Y = S.Rn[3 * t + 1], Z = S.Rn[3 * t + 2];
          const double dn_ = sqrt((nx - X) * (nx - X) + (ny - Y) * (ny - Y) + (nz - Z) * (nz - Z));
          const double do_ = sqrt((ox - X) * (ox - X) + (oy - Y) * (oy - Y) + (oz - Z) * (oz - Z));
          const double A = S.j1_A[t], c = S.j1_c[t];
          if constexpr (MIXED) {  // zone jastrow_ratio in fp32 (differences formed in fp64, then rounded)
            const float af = (float)S.j1_a, Af = (float)A, cf = (float)c, dnf = (float)dn_, dof = (float)do_;
            const float fn = S.j1_type == 1 ? -Af * (1.0f - expf(-af * cf * dnf)) / (2.0f * af) : -0.5f * Af * dnf / (1.0f + af * cf * dnf);
            const float fo = S.j1_type == 1 ? -Af * (1.0f - expf(-af * cf * dof)) / (2.0f * af) : -0.5f * Af * dof / (1.0f + af * cf * dof);
            dJ += (double)(fn - fo);
          } else
          dJ += j1_f(S.j1_type, S.j1_a, A, c, dn_) - j1_f(S.j1_type, S.j1_a, A, c, do_);
        } else {
          const int j = t - n_j1;
          if (j == ke) continue;
          double x, y, z;
          pos.get(j, x, y, z);
          const double dn_ = sqrt((nx - x) * (nx - x) + (ny - y) * (ny - y) + (nz - z) * (nz - z));
          const double do_ = sqrt((ox - x) * (ox - x) + (oy - y) * (oy - y) + (oz - z) * (oz - z));
          if constexpr (MIXED) {
            const float af = (float)S.j2_a, dnf = (float)dn_, dof = (float)do_;
            const float fn = S.j2_type == 1 ? 0.5f * dnf / (1.0f + af * dnf) : (1.0f - expf(-af * dnf)) / (2.0f * af);
            const float fo = S.j2_type == 1 ? 0.5f * dof / (1.0f + af * dof) : (1.0f - expf(-af * dof)) / (2.0f * af);
            dJ += (double)(fn - fo);
          } else
          dJ += j2_f(S.j2_type, S.j2_a, dn_) - j2_f(S.j2_type, S.j2_a, do_);
        }
      }
      s_TJ[(1 + wid) * 32 + lane] = dJ;
    };
    if (wid < nch) {
      // same AO tables for both spins (checked at create); the MO coefficients may differ per lane
      SinkMO<NMO> sink;
      sink.init(tab + (up ? B.off_C : B.off_C2));
      if constexpr (MIXED) eval_val_n_f32<CART, QE_LMAX, 1>(tab, B, P.off_cseg, &nx, &ny, &nz, cbeg[wid], cbeg[wid + 1], sink);  // ao_eval in fp32
      else eval_val<CART, QE_LMAX>(tab, B, P.off_cseg, nx, ny, nz, cbeg[wid], cbeg[wid + 1], sink);
#pragma unroll
      for (int mo = 0; mo < NMO; ++mo) SPART(wid, mo) = sink.acc[mo];
      jastrow_terms();
    } else {
      double dist;
      const int ia = nearest_atom(S.Rn, S.n_atom, nx, ny, nz, 0, &dist);
      const double Zc = S.Zeff[ia];
      const double f_p = 1.0 / (Zc * Zc) * (1.0 + Zc * Zc * dist) / (1.0 + dist);
      const double dd = (nx - ox) * (nx - ox) + (ny - oy) * (ny - oy) + (nz - oz) * (nz - oz);
      const double T_ratio =
          (f_l / f_p) * qexp(-dd * (1.0 / (2.0 * f_p * f_p * P.Dt * P.Dt) - 1.0 / (2.0 * f_l * f_l * P.Dt * P.Dt)));
      s_TJ[lane] = T_ratio;
      s_fl[Ne * 32 + lane] = f_p;
      jastrow_terms();
    }
    __syncthreads();

    // ---- phase C: warp 0 -------------------------------------------------------------------------
    if (wid == 0) {
      // sums over the warps' partials: three interleaved accumulators per quantity (fixed order, short dependent chains --
      // this warp works alone here, so every exposed latency is on the critical path of the proposal)
      double phi[NMO];
      {
        double a0[NMO], a1[NMO], a2[NMO];
#pragma unroll
        for (int mo = 0; mo < NMO; ++mo) a0[mo] = a1[mo] = a2[mo] = 0.0;
        int c = 0;
        for (; c + 2 < nch; c += 3) {
#pragma unroll
          for (int mo = 0; mo < NMO; ++mo) {
            a0[mo] += SPART(c, mo);
            a1[mo] += SPART(c + 1, mo);
            a2[mo] += SPART(c + 2, mo);
          }
        }
        for (; c < nch; ++c) {
#pragma unroll
          for (int mo = 0; mo < NMO; ++mo) a0[mo] += SPART(c, mo);
        }
#pragma unroll
        for (int mo = 0; mo < NMO; ++mo) phi[mo] = (a0[mo] + a1[mo]) + a2[mo];
      }
      double dJ;
      {
        double d0 = 0.0, d1 = 0.0, d2 = 0.0;
        int c = 0;
        for (; c + 2 <= nch; c += 3) {
          d0 += s_TJ[(1 + c) * 32 + lane];
          d1 += s_TJ[(2 + c) * 32 + lane];
          d2 += s_TJ[(3 + c) * 32 + lane];
        }
        for (; c <= nch; ++c) d0 += s_TJ[(1 + c) * 32 + lane];
        dJ = (d0 + d1) + d2;
      }
      // v (row difference) or u (column difference), Det_ratio = 1 + v^T Ginv u.  All loops over electrons run to the
      // compile-time bound NB >= N with a uniform predicate, so that dvec / col / vt stay in registers (runtime bounds put them
      // in local memory and made this single-warp section 40 % of the kernel, profiles/r01_mcmc_v5.md)
      constexpr int NB = NMO < 8 ? NMO : 8;  // N <= 8 (host check) and N <= n_mo <= NMO
      double dvec[NB];
      double Det;
      if (up) {
        const int k = ke;
        double t[NMO];
#pragma unroll
        for (int b = 0; b < NMO; ++b) {
          double s = 0;
#pragma unroll
          for (int a = 0; a < NMO; ++a) s = fma(phi[a], S.lam_p[a * NMO + b], s);
          t[b] = s;
        }
        double acc = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          double s = 0;
          if (j < Nd) {
#pragma unroll
            for (int b = 0; b < NMO; ++b) s = fma(t[b], SPHI(N + j, b), s);
          } else if (j < N) {
            const int q = j - Nd;
#pragma unroll
            for (int a = 0; a < NMO; ++a) s = fma(phi[a], S.lam_u[a * S.n_unp + q], s);
          }
          if (j < N) {
            dvec[j] = s - SG(k, j);
            acc = fma(dvec[j], SGI(j, k), acc);
          } else {
            dvec[j] = 0.0;
          }
        }
        Det = 1.0 + acc;
      } else {
        const int k = ke - N;
        double t[NMO];
#pragma unroll
        for (int a = 0; a < NMO; ++a) {
          double s = 0;
#pragma unroll
          for (int b = 0; b < NMO; ++b) s = fma(S.lam_p[a * NMO + b], phi[b], s);
          t[a] = s;
        }
        double acc = 0;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          if (i < N) {
            double s = 0;
#pragma unroll
            for (int a = 0; a < NMO; ++a) s = fma(SPHI(i, a), t[a], s);
            dvec[i] = s - SG(i, k);
            acc = fma(SGI(k, i), dvec[i], acc);
          } else {
            dvec[i] = 0.0;
          }
        }
        Det = 1.0 + acc;
      }
      const double T_ratio = s_TJ[lane];
      const double J_ratio = qexp(dJ);
      // AS regularisation of the proposed state without materialising it
      double R_AS_ratio = 1.0, R_AS_p = R_AS_cur;
      if (P.eps_AS > 0.0) {
        double F = 0, Smin = 1e300;
        double* s_dv = s_part;  // the partial sums have been consumed: scratch for the dynamically indexed copy of dvec
#pragma unroll
        for (int j = 0; j < NB; ++j) s_dv[j * 32 + lane] = dvec[j];
#define DV(j) s_dv[(j) * 32 + lane]
        if (up) {
          const int k = ke;
          // Ginv' = Ginv - Ginv[:,k] (v^T Ginv) / Det
          for (int jp = 0; jp < N; ++jp) {
            double vt = 0;
            for (int j = 0; j < N; ++j) vt = fma(DV(j), SGI(j, jp), vt);
            vt /= Det;
            for (int i = 0; i < N; ++i) {
              const double x = SGI(i, jp) - SGI(i, k) * vt;
              F = fma(x, x, F);
            }
          }
          for (int i = 0; i < N; ++i) {
            double r = 0, c = 0;
            for (int j = 0; j < N; ++j) {
              const double gij = SG(i, j) + (i == k ? DV(j) : 0.0);
              const double gji = SG(j, i) + (j == k ? DV(i) : 0.0);
              r = fma(gij, gij, r);
              c = fma(gji, gji, c);
            }
            Smin = fmin(Smin, fmin(r, c));
          }
        } else {
          const int k = ke - N;
          // Ginv' = Ginv - (Ginv u) Ginv[k,:] / Det
          for (int i = 0; i < N; ++i) {
            double au = 0;
            for (int j = 0; j < N; ++j) au = fma(SGI(i, j), DV(j), au);
            au /= Det;
            for (int j = 0; j < N; ++j) {
              const double x = SGI(i, j) - au * SGI(k, j);
              F = fma(x, x, F);
            }
          }
          for (int i = 0; i < N; ++i) {
            double r = 0, c = 0;
            for (int j = 0; j < N; ++j) {
              const double gij = SG(i, j) + (j == k ? DV(i) : 0.0);
              const double gji = SG(j, i) + (i == k ? DV(j) : 0.0);
              r = fma(gij, gij, r);
              c = fma(gji, gji, c);
            }
            Smin = fmin(Smin, fmin(r, c));
          }
        }
#undef DV
        const double SF = Smin * F;
        R_AS_p = SF > 0.0 ? pow(SF, -0.375) : 0.0;
        R_AS_ratio = (fmax(R_AS_p, P.eps_AS) / R_AS_p) / (fmax(R_AS_cur, P.eps_AS) / R_AS_cur);
      }
      const double wr = R_AS_ratio * J_ratio * Det;
      const double x = wr * wr * T_ratio;
      const double b = rb_c;
      const bool ok = (x == x) && (b < fmin(1.0, x)) && (Det != 0.0);
      if (ok) {
        ++n_acc;
        R_AS_cur = R_AS_p;
        SR(ke, 0) = nx;
        SR(ke, 1) = ny;
        SR(ke, 2) = nz;
        s_fl[ke * 32 + lane] = s_fl[Ne * 32 + lane];
#pragma unroll
        for (int mo = 0; mo < NMO; ++mo) SPHI(ke, mo) = phi[mo];
        const double invD = 1.0 / Det;
        if (up) {
          const int k = ke;
          double col[NB], vt[NB];
#pragma unroll
          for (int i = 0; i < NB; ++i) col[i] = i < N ? SGI(i, k) : 0.0;
#pragma unroll
          for (int jp = 0; jp < NB; ++jp) {
            double s = 0;
            if (jp < N) {
#pragma unroll
              for (int j = 0; j < NB; ++j)
                if (j < N) s = fma(dvec[j], SGI(j, jp), s);
            }
            vt[jp] = s;
          }
#pragma unroll
          for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int jp = 0; jp < NB; ++jp)
              if (i < N && jp < N) SGI(i, jp) = SGI(i, jp) - (col[i] * vt[jp]) * invD;
#pragma unroll
          for (int j = 0; j < NB; ++j)
            if (j < N) SG(k, j) += dvec[j];
        } else {
          const int k = ke - N;
          double au[NB], row[NB];
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            double s = 0;
            if (i < N) {
#pragma unroll
              for (int j = 0; j < NB; ++j)
                if (j < N) s = fma(SGI(i, j), dvec[j], s);
            }
            au[i] = s;
          }
#pragma unroll
          for (int j = 0; j < NB; ++j) row[j] = j < N ? SGI(k, j) : 0.0;
#pragma unroll
          for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j)
              if (i < N && j < N) SGI(i, j) = SGI(i, j) - (au[i] * row[j]) * invD;
#pragma unroll
          for (int i = 0; i < NB; ++i)
            if (i < N) SG(i, k) += dvec[i];
        }
      } else {
        ++n_rej;
      }
    }
    __syncthreads();
  }

  // ---- write back ---------------------------------------------------------------------------------
  if (live) {
    for (int idx = wid; idx < Ne * 3; idx += nch + 1) {
      const int e = idx / 3, c = idx % 3;
      if (e < N) P.r_up[((size_t)w * N + e) * 3 + c] = SR(e, c);
      else P.r_dn[((size_t)w * Nd + (e - N)) * 3 + c] = SR(e, c);
    }
    for (int idx = wid; idx < NN2; idx += nch + 1) {
      P.G[(size_t)w * NN2 + idx] = s_G[idx * 32 + lane];
      P.Ginv[(size_t)w * NN2 + idx] = s_Gi[idx * 32 + lane];
    }
    if (wid == 0) {
      P.acc[w] = n_acc;
      P.rej[w] = n_rej;
    }
  }
#undef SR
#undef SG
#undef SGI
#undef SPHI
#undef SPART
}


// =================================================================================================
// Metropolis kernel, epsilon_AS = 0 (the CLI default): the work after the AO phase is spread over all warps.
//   The determinant ratio is linear in the orbital values at the proposed point, so every chunk warp contracts ITS partial
//   orbital values to partial row / column differences (with the cached M = lambda Phi_dn, Mt = Phi_up^T lambda) and to a
//   partial ratio; the decision is then a sum over the chunks (one warp), and an accepted move is applied in three short
//   warp-parallel stages (difference vector and new orbital values | v^T Ginv and the cache column | Ginv rows, G row).
//   The previous kernel (kept below for epsilon_AS > 0) did all of it in one warp: 15 of 16 warps waited at the end-of-proposal
//   barrier for ~2.8 us per proposal (profiles/r01_full_v5.md).
// =================================================================================================
template <int NMO, bool CART, int LMAX, bool MIXED>
__global__ void __launch_bounds__(512)
k_mcmc2(BasisDev B, SysDev S, McmcArgs P) {
  extern __shared__ __align__(16) double sm_all[];
  const int lane = threadIdx.x, wid = threadIdx.y, nwarp = P.n_chunk + 1;
  const char* tab = (const char*)sm_all;
  {
    const int4* src = (const int4*)B.g;
    int4* dst = (int4*)sm_all;
    for (int i = wid * 32 + lane; i < B.bytes / 16; i += 32 * nwarp) dst[i] = src[i];
  }
  double* sm = sm_all + B.bytes / 8;
  const int w = blockIdx.x * 32 + lane;
  const bool live = w < P.nw;
  const int ww = live ? w : P.nw - 1;  // dead lanes shadow the last walker (no stores)
  const int N = S.n_up, Nd = S.n_dn, Ne = S.n_e, NN2 = N * N;
  const int nch = P.n_chunk;
  const int* cbeg = (const int*)(tab + P.off_cbeg);
  const int cw = wid < nch ? P.chunk_warp[wid] : 0;  // chunk of this warp (warps of one SM sub-partition sweep the same l)
  const bool clk_on = P.clk != nullptr && lane == 0 && wid == 0;
  long long clk_last = clk_on ? clock64() : 0, clk_acc[6] = {0, 0, 0, 0, 0, 0};
#define PHASE(i)                    \
  if (clk_on) {                     \
    const long long t_ = clock64(); \
    clk_acc[i] += t_ - clk_last;    \
    clk_last = t_;                  \
  }
  double* s_r = sm;                        // Ne*3
  double* s_G = s_r + Ne * 3 * 32;         // N*N
  double* s_Gi = s_G + NN2 * 32;           // N*N
  double* s_phi = s_Gi + NN2 * 32;         // Ne*NMO   orbital values at the electrons [e][mo]
  double* s_M = s_phi + Ne * NMO * 32;     // NMO*N    M[a][j] = sum_b lam_p[a][b] phi_dn_b(r_j) | lam_u[a][j - Nd]
  double* s_Mt = s_M + NMO * N * 32;       // N*NMO    Mt[i][b] = sum_a phi_up_a(r_i) lam_p[a][b]
  double* s_part = s_Mt + N * NMO * 32;    // nch*NMO  partial orbital values of the chunks
  double* s_pdv = s_part + nch * NMO * 32; // nch*N    partial new row (up move) / column (down move) of G
  double* s_pd = s_pdv + nch * N * 32;     // nch      partial determinant ratio
  double* s_TJ = s_pd + nch * 32;          // nwarp+1  T_ratio, per-warp partial Jastrow exponent differences
  double* s_fl = s_TJ + (nwarp + 1) * 32;  // Ne+1     proposal width factor f of every electron, f' of the proposal
  double* s_dv = s_fl + (Ne + 1) * 32;     // N        new - old row / column of G
  double* s_vt = s_dv + N * 32;            // N        v^T Ginv (up) / Ginv u (down)
  double* s_rk = s_vt + N * 32;            // N        old row k of Ginv (down move)
  double* s_phn = s_rk + N * 32;           // NMO      orbital values at the accepted point
  double* s_dec = s_phn + NMO * 32;        // 2        determinant ratio, accepted flag
#define SR(e, c) s_r[((e) * 3 + (c)) * 32 + lane]
#define SG(i, j) s_G[((i) * N + (j)) * 32 + lane]
#define SGI(i, j) s_Gi[((i) * N + (j)) * 32 + lane]
#define SPHI(e, mo) s_phi[((e) * NMO + (mo)) * 32 + lane]
#define SPART(c, mo) s_part[((c) * NMO + (mo)) * 32 + lane]
#define SM_(a, j) s_M[((a) * N + (j)) * 32 + lane]
#define SMT(i, b) s_Mt[((i) * NMO + (b)) * 32 + lane]
#define SPDV(c, j) s_pdv[((c) * N + (j)) * 32 + lane]

  for (int idx = wid; idx < Ne * 3; idx += nwarp) {
    const int e = idx / 3, c = idx % 3;
    SR(e, c) = e < N ? P.r_up[((size_t)ww * N + e) * 3 + c] : P.r_dn[((size_t)ww * Nd + (e - N)) * 3 + c];
  }
  for (int idx = wid; idx < NN2; idx += nwarp) {
    s_G[idx * 32 + lane] = P.G[(size_t)ww * NN2 + idx];
    s_Gi[idx * 32 + lane] = P.Ginv[(size_t)ww * NN2 + idx];
  }
  __syncthreads();
  for (int e = wid; e < Ne; e += nwarp) {  // f = (1 + Z^2 d) / (Z^2 (1 + d)) with the nearest nucleus (jqmc/jqmc_mcmc.py:4340-4357)
    double dist;
    const int ia = nearest_atom(S.Rn, S.n_atom, SR(e, 0), SR(e, 1), SR(e, 2), 0, &dist);
    const double Zc = S.Zeff[ia];
    s_fl[e * 32 + lane] = 1.0 / (Zc * Zc) * (1.0 + Zc * Zc * dist) / (1.0 + dist);
  }
  // orbital values at every electron (fp64 sweep also in the mixed mode: these are the cached geminal operands)
  for (int e = 0; e < Ne; ++e) {
    if (wid < nch) {
      SinkMO<NMO> sink;
      sink.init(tab + (e < N ? B.off_C : B.off_C2));
      eval_val<CART, LMAX>(tab, B, P.off_cseg, SR(e, 0), SR(e, 1), SR(e, 2), cbeg[cw], cbeg[cw + 1], sink);
#pragma unroll
      for (int mo = 0; mo < NMO; ++mo) SPART(cw, mo) = sink.acc[mo];
    }
    __syncthreads();
    for (int mo = wid; mo < NMO; mo += nwarp) {
      double sum = 0;
      for (int c = 0; c < nch; ++c) sum += SPART(c, mo);
      SPHI(e, mo) = sum;
    }
    __syncthreads();
  }
  // M and Mt from the cached orbital values
  for (int t = wid; t < NMO * N; t += nwarp) {
    const int a = t / N, j = t % N;
    double sum = 0;
    if (j < Nd) {
      for (int b = 0; b < NMO; ++b) sum = fma(S.lam_p[a * NMO + b], SPHI(N + j, b), sum);
    } else {
      sum = S.lam_u[a * S.n_unp + (j - Nd)];
    }
    SM_(a, j) = sum;
  }
  for (int t = wid; t < N * NMO; t += nwarp) {
    const int i = t / NMO, b = t % NMO;
    double sum = 0;
    for (int a = 0; a < NMO; ++a) sum = fma(SPHI(i, a), S.lam_p[a * NMO + b], sum);
    SMT(i, b) = sum;
  }
  __syncthreads();
  PHASE(0)

  int n_acc = 0, n_rej = 0;
  int ke_n = 0, axis_n = 0;
  double rg_n = 0.0, rb_n = 0.0;
  if (P.nmpm > 0) {
    ke_n = P.rsel[ww];
    axis_n = P.raxis[ww];
    rg_n = P.rg[ww];
    rb_n = P.rb[ww];
  }
  for (int it = 0; it < P.nmpm; ++it) {
    // ---- phase A: proposal (every thread, redundantly; lane = walker) ----------------------------
    const int ke = ke_n, axis = axis_n;
    const double rg_c = rg_n, rb_c = rb_n;
    if (it + 1 < P.nmpm) {
      const size_t rnext = (size_t)(it + 1) * P.nw + ww;
      ke_n = P.rsel[rnext];
      axis_n = P.raxis[rnext];
      rg_n = P.rg[rnext];
      rb_n = P.rb[rnext];
    }
    const bool up = ke < N;
    const int k = up ? ke : ke - N;
    const double ox = SR(ke, 0), oy = SR(ke, 1), oz = SR(ke, 2);
    const double f_l = s_fl[ke * 32 + lane];
    const double g = rg_c * (f_l * P.Dt);
    double nx = ox, ny = oy, nz = oz;
    if (axis == 0) nx = ox + g;
    else if (axis == 1) ny = oy + g;
    else nz = oz + g;

    // ---- phase B: AO sweep of the chunk, partial row / column of G, partial ratio; Jastrow terms over all warps -----
    const int n_j1 = S.j1_type ? S.n_atom : 0, n_jt = n_j1 + (S.j2_type ? Ne : 0);
    {
      double dJ = 0.0;
#pragma unroll 1
      for (int t = wid; t < n_jt; t += nwarp) {
        double X, Y, Z;
        if (t < n_j1) {
          X = S.Rn[3 * t], Y = S.Rn[3 * t + 1], Z = S.Rn[3 * t + 2];
        } else {
          const int j = t - n_j1;
          if (j == ke) continue;
          X = s_r[(j * 3 + 0) * 32 + lane], Y = s_r[(j * 3 + 1) * 32 + lane], Z = s_r[(j * 3 + 2) * 32 + lane];
        }
        const double dn_ = sqrt((nx - X) * (nx - X) + (ny - Y) * (ny - Y) + (nz - Z) * (nz - Z));
        const double do_ = sqrt((ox - X) * (ox - X) + (oy - Y) * (oy - Y) + (oz - Z) * (oz - Z));
        if constexpr (MIXED) {  // zone jastrow_ratio in fp32 (distances formed in fp64, then rounded)
          const float dnf = (float)dn_, dof = (float)do_;
          if (t < n_j1) {
            const float af = (float)S.j1_a, Af = (float)S.j1_A[t], cf = (float)S.j1_c[t];
            const float fn = S.j1_type == 1 ? -Af * (1.0f - expf(-af * cf * dnf)) / (2.0f * af) : -0.5f * Af * dnf / (1.0f + af * cf * dnf);
            const float fo = S.j1_type == 1 ? -Af * (1.0f - expf(-af * cf * dof)) / (2.0f * af) : -0.5f * Af * dof / (1.0f + af * cf * dof);
            dJ += (double)(fn - fo);
          } else {
            const float af = (float)S.j2_a;
            const float fn = S.j2_type == 1 ? 0.5f * dnf / (1.0f + af * dnf) : (1.0f - expf(-af * dnf)) / (2.0f * af);
            const float fo = S.j2_type == 1 ? 0.5f * dof / (1.0f + af * dof) : (1.0f - expf(-af * dof)) / (2.0f * af);
            dJ += (double)(fn - fo);
          }
        } else if (t < n_j1) {
          const double A = S.j1_A[t], c = S.j1_c[t];
          dJ += j1_f(S.j1_type, S.j1_a, A, c, dn_) - j1_f(S.j1_type, S.j1_a, A, c, do_);
        } else {
          dJ += j2_f(S.j2_type, S.j2_a, dn_) - j2_f(S.j2_type, S.j2_a, do_);
        }
      }
      s_TJ[(1 + wid) * 32 + lane] = dJ;
    }
    if (wid < nch) {
      SinkMO<NMO> sink;
      sink.init(tab + (up ? B.off_C : B.off_C2));
      if constexpr (MIXED) eval_val_n_f32<CART, LMAX, 1>(tab, B, P.off_cseg, &nx, &ny, &nz, cbeg[cw], cbeg[cw + 1], sink);  // ao_eval in fp32
      else eval_val<CART, LMAX>(tab, B, P.off_cseg, nx, ny, nz, cbeg[cw], cbeg[cw + 1], sink);
      double pd = 0.0;
      if (up) {  // new row k of G: G'[k][j] = sum_a phi_a(r') M[a][j]; ratio = sum_j G'[k][j] Ginv[j][k]
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
          double sum = 0;
#pragma unroll
          for (int a = 0; a < NMO; ++a) sum = fma(sink.acc[a], SM_(a, j), sum);
          SPDV(cw, j) = sum;
          pd = fma(sum, SGI(j, k), pd);
        }
      } else {  // new column k of G: G'[i][k] = sum_b Mt[i][b] phi_b(r'); ratio = sum_i Ginv[k][i] G'[i][k]
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
          double sum = 0;
#pragma unroll
          for (int b = 0; b < NMO; ++b) sum = fma(SMT(i, b), sink.acc[b], sum);
          SPDV(cw, i) = sum;
          pd = fma(SGI(k, i), sum, pd);
        }
      }
      s_pd[cw * 32 + lane] = pd;
#pragma unroll
      for (int mo = 0; mo < NMO; ++mo) SPART(cw, mo) = sink.acc[mo];
    } else {
      double dist;
      const int ia = nearest_atom(S.Rn, S.n_atom, nx, ny, nz, 0, &dist);
      const double Zc = S.Zeff[ia];
      const double f_p = 1.0 / (Zc * Zc) * (1.0 + Zc * Zc * dist) / (1.0 + dist);
      const double dd = (nx - ox) * (nx - ox) + (ny - oy) * (ny - oy) + (nz - oz) * (nz - oz);
      s_TJ[lane] = (f_l / f_p) * qexp(-dd * (1.0 / (2.0 * f_p * f_p * P.Dt * P.Dt) - 1.0 / (2.0 * f_l * f_l * P.Dt * P.Dt)));
      s_fl[Ne * 32 + lane] = f_p;
    }
    __syncthreads();
    PHASE(1)

    // ---- phase C: decision (warp 0): sums over the chunks in a fixed order -------------------------
    // (the loops of this and the following stages are kept rolled: a warp running cold straight-line code alone is bound by
    //  instruction fetch, ~30 cycles per 128-byte line of 8 instructions; measured 3.3k cycles for this stage when unrolled)
    int any_ok = 0;
    if (wid == 0) {
      double d0 = 0.0, d1 = 0.0, j0 = 0.0, j1 = 0.0;  // two accumulators each (even / odd terms), fixed order
#pragma unroll 1
      for (int c = 0; c < nwarp; c += 2) {
        if (c < nch) d0 += s_pd[c * 32 + lane];
        if (c + 1 < nch) d1 += s_pd[(c + 1) * 32 + lane];
        j0 += s_TJ[(1 + c) * 32 + lane];
        if (c + 1 < nwarp) j1 += s_TJ[(2 + c) * 32 + lane];
      }
      // Det = 1 + v^T Ginv u with v = new - old row (column): the sum over the chunks is (new row) . Ginv[:,k]; subtracting
      // (old row) . Ginv[:,k] -- 1 up to the drift of the running inverse -- keeps the ratio CONSISTENT with the Sherman-
      // Morrison update below.  Using the chunk sum alone (i.e. assuming Ginv G = 1 exactly) feeds that drift back into
      // every update and it grows exponentially (measured: x100 every 160 proposals).
      double old = 0.0;
#pragma unroll 1
      for (int j = 0; j < N; ++j) old = up ? fma(SG(k, j), SGI(j, k), old) : fma(SGI(k, j), SG(j, k), old);
      const double Det = 1.0 + ((d0 + d1) - old);
      const double wr = qexp(j0 + j1) * Det;
      const double x = wr * wr * s_TJ[lane];
      const bool ok = (x == x) && (rb_c < fmin(1.0, x)) && (Det != 0.0);
      if (ok) ++n_acc; else ++n_rej;
      s_dec[lane] = Det;
      s_dec[32 + lane] = ok ? 1.0 : 0.0;
      any_ok = ok;
    }
    if (!__syncthreads_or(any_ok)) continue;  // every walker of the CTA rejected: nothing to update
    PHASE(2)
    const bool ok = s_dec[32 + lane] != 0.0;

    // ---- U1: difference vector (warps 0..N-1), orbital values at the new point (next NMO warps), position -----------
#pragma unroll 1
    for (int t = wid; t < N + NMO + 1; t += nwarp) {
      if (t < N) {
        double sum = 0;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) sum += SPDV(c, t);
        if (ok) s_dv[t * 32 + lane] = sum - (up ? SG(k, t) : SG(t, k));
      } else if (t < N + NMO) {
        const int mo = t - N;
        double sum = 0;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) sum += SPART(c, mo);
        if (ok) s_phn[mo * 32 + lane] = sum;
      } else if (ok) {
        SR(ke, 0) = nx;
        SR(ke, 1) = ny;
        SR(ke, 2) = nz;
        s_fl[ke * 32 + lane] = s_fl[Ne * 32 + lane];
      }
    }
    __syncthreads();
    PHASE(3)
    // ---- U2a: v^T Ginv (up) / Ginv u and the old row k (down); cache column of the moved electron ---------------------
#pragma unroll 1
    for (int t = wid; t < N + NMO; t += nwarp) {
      if (t < N) {
        double sum = 0;
        if (up) {
#pragma unroll 1
          for (int j = 0; j < N; ++j) sum = fma(s_dv[j * 32 + lane], SGI(j, t), sum);
        } else {
#pragma unroll 1
          for (int j = 0; j < N; ++j) sum = fma(SGI(t, j), s_dv[j * 32 + lane], sum);
          if (ok) s_rk[t * 32 + lane] = SGI(k, t);
        }
        if (ok) s_vt[t * 32 + lane] = sum;
      } else if (ok) {
        const int b = t - N;
        SPHI(ke, b) = s_phn[b * 32 + lane];
        double sum = 0;
        if (up) {  // Mt[k][b] = sum_a phi_new_a lam_p[a][b]
#pragma unroll 1
          for (int a = 0; a < NMO; ++a) sum = fma(s_phn[a * 32 + lane], S.lam_p[a * NMO + b], sum);
          SMT(k, b) = sum;
        } else {  // M[b][k] = sum_c lam_p[b][c] phi_new_c   (k < Nd: a down electron is always a paired column)
#pragma unroll 1
          for (int c = 0; c < NMO; ++c) sum = fma(S.lam_p[b * NMO + c], s_phn[c * 32 + lane], sum);
          SM_(b, k) = sum;
        }
      }
    }
    __syncthreads();
    PHASE(4)
    // ---- U2b: Sherman-Morrison, one warp per row of Ginv; G row / column --------------------------------------------
#pragma unroll 1
    for (int t = wid; t < N + 1; t += nwarp) {
      if (!ok) continue;
      const double invD = 1.0 / s_dec[lane];
      if (t < N) {
        const int i = t;
        if (up) {  // Ginv' = Ginv - Ginv[:,k] (v^T Ginv) / Det
          const double ci = SGI(i, k) * invD;
#pragma unroll 1
          for (int jp = 0; jp < N; ++jp) SGI(i, jp) = SGI(i, jp) - ci * s_vt[jp * 32 + lane];
        } else {  // Ginv' = Ginv - (Ginv u) Ginv[k,:] / Det
          const double ai = s_vt[i * 32 + lane] * invD;
#pragma unroll 1
          for (int j = 0; j < N; ++j) SGI(i, j) = SGI(i, j) - ai * s_rk[j * 32 + lane];
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
          if (up) SG(k, j) += s_dv[j * 32 + lane];
          else SG(j, k) += s_dv[j * 32 + lane];
        }
      }
    }
    __syncthreads();
    PHASE(5)
  }
  if (clk_on)
#pragma unroll 1
    for (int i = 0; i < 6; ++i) atomicAdd((unsigned long long*)&P.clk[i], (unsigned long long)clk_acc[i]);
#undef PHASE

  if (live) {
#pragma unroll 1
    for (int idx = wid; idx < Ne * 3; idx += nwarp) {
      const int e = idx / 3, c = idx % 3;
      if (e < N) P.r_up[((size_t)w * N + e) * 3 + c] = SR(e, c);
      else P.r_dn[((size_t)w * Nd + (e - N)) * 3 + c] = SR(e, c);
    }
#pragma unroll 1
    for (int idx = wid; idx < NN2; idx += nwarp) {
      P.G[(size_t)w * NN2 + idx] = s_G[idx * 32 + lane];
      P.Ginv[(size_t)w * NN2 + idx] = s_Gi[idx * 32 + lane];
    }
    if (wid == 0) {
      P.acc[w] = n_acc;
      P.rej[w] = n_rej;
    }
  }
#undef SR
#undef SG
#undef SGI
#undef SPHI
#undef SPART
#undef SM_
#undef SMT
#undef SPDV
}



// launches the kernel instantiation (NMO_I, CART_I) selected by the engine's orbital padding / basis kind
template <int NMO_I, bool CART_I>
int mcmc_launch_one(qe_engine* h, McmcArgs& A, int nw, double epsilon_AS, cudaStream_t st) {
  const SysDev& S = h->sys;
  const int P = h->nmo_pad, nch = h->b_up.n_chunk;
  static const bool old_env = getenv("QE_MCMC_OLD") && atoi(getenv("QE_MCMC_OLD")) != 0;  // A/B switch (tuning / debugging)
  if (!(epsilon_AS > 0.0) && !old_env) {
    // no AS regularisation (the CLI default): the kernel with the warp-parallel decision / update stages
    const int N_ = S.n_up, Ne_ = S.n_e;
    const size_t smem2 = (size_t)(Ne_ * 3 + 2 * N_ * N_ + Ne_ * P + 2 * P * N_ + nch * P + nch * N_ + nch + (nch + 2) + (Ne_ + 1) + 3 * N_ + P + 2) * 32 * 8 +
                         (size_t)h->b_up.dev.bytes;
    if (smem2 <= 227 * 1024) {
      dim3 block2(32, nch + 1);
      LaunchScope ls_(h, K_MCMC, st);
#define CALL2M(NMO, CART, LMAX, MX)                                                                                        \
  do {                                                                                                                     \
    CUDA_TRY(cudaFuncSetAttribute(k_mcmc2<NMO, CART, LMAX, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
    k_mcmc2<NMO, CART, LMAX, MX><<<nblk(nw, 32), block2, smem2, st>>>(h->b_up.dev, S, A);                                  \
  } while (0)
#define CALL2(NMO, CART)                                                     \
  do {                                                                       \
    if (h->b_up.dev.lmax <= 4) {                                             \
      if (h->mixed) CALL2M(NMO, CART, 4, true);                              \
      else CALL2M(NMO, CART, 4, false);                                      \
    } else {                                                                 \
      CALL2M(NMO, CART, 6, false);                                           \
    }                                                                        \
  } while (0)
      CALL2(NMO_I, CART_I);
#undef CALL2
#undef CALL2M
      CHECK_LAUNCH();
      return QE_OK;
    }
  }
  const size_t smem = (size_t)(S.n_e * 3 + 2 * S.n_up * S.n_up + S.n_e * P + nch * P + 2 + nch + S.n_e + 1) * 32 * 8 + (size_t)h->b_up.dev.bytes;
  if (smem > 227 * 1024) return fail(QE_ERR_UNSUPPORTED, "qe_mcmc_update: basis image does not fit in shared memory");
  dim3 block(32, nch + 1);
  { LaunchScope ls_(h, K_MCMC, st);
#define CALLM(NMO, CART, MX)                                                                                       \
  do {                                                                                                             \
    CUDA_TRY(cudaFuncSetAttribute(k_mcmc<NMO, CART, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_mcmc<NMO, CART, MX><<<nblk(nw, 32), block, smem, st>>>(h->b_up.dev, S, A);                                   \
  } while (0)
#define CALL(NMO, CART)                        \
  do {                                         \
    if (h->mixed) CALLM(NMO, CART, true);      \
    else CALLM(NMO, CART, false);              \
  } while (0)
  CALL(NMO_I, CART_I);
#undef CALL
#undef CALLM
  }
  CHECK_LAUNCH();
  return QE_OK;
}
