// Fused per-walker kernel: LRDMC projection (GFMC_n._projection_n, jqmc/jqmc_gfmc.py:4738-5358), its
// observable V_diag / V_nondiag (_compute_V_elements_n, :5360-5627) and the VMC local energy
// (compute_local_energy_fast, jqmc/hamiltonians.py:225-290) share one kernel.
//
//   CTA = WPC walkers x NWARP warps.  The basis image and the Jastrow / ECP tables are staged in shared memory once per
//   CTA; walker state (positions, running inverse, cached MO value/grad/lap of every electron, ratio weight vectors,
//   mesh elements) lives in shared memory as [item][walker].
//
//   A thread is a (walker, task) pair.  The dominant tasks -- one AO sweep per mesh point -- are the SAME code for every
//   point, so the 32 lanes of a warp hold consecutive (walker, point) pairs: control flow stays uniform and table reads are
//   warp-broadcast LDS.  One 16-warp CTA per SM; WPC is chosen so that the grid is a single wave (4096 walkers -> 28 per
//   CTA, 147 CTAs; choose_wpc below).  Several smaller CTAs per SM were measured slower (instruction-cache thrash between
//   CTAs sitting in different phases).
//
//   per projection:  P0  draws: two Threefry splits, mesh rotation, move uniform (first warp, concurrently with P1)
//                    P1  ratio weight vectors W[:,e] = lambda Phi_dn Ginv[:,e] from the running inverse, and the
//                        Jastrow terms of every electron at its current position             (walker, electron) x 2 task kinds
//                    P2  6 N_e kinetic-mesh ratios + N_e*NN*Nv ECP-mesh ratios, two points per thread (warp rounds handed
//                        out by a shared counter), per-electron continuum kinetic energy and potentials     (walker, point)
//                    P3  fixed-node split and per-electron chunk sums (walker, electron); diagonal / off-diagonal sums,
//                        weight and move selection (walker): whole electrons are skipped by their chunk sums
//                    P4  value/grad/lap of the moved electron (warp = basis chunk), Sherman-Morrison update
//   GFMC_t (template TAU): the same loop until the walkers of the CTA are out of time, every phase enumerating only the
//   walkers that still have time left; see qe_walker_tau.cu.
#pragma once
#include "qe_common.cuh"

// (outside the unnamed namespace: the per-instantiation translation units and the front ends exchange it)
struct WalkerArgs {
  int nw, nmpm, mode;  // mode 0: projection, 1: V elements (no move), 2: VMC local energy
  int dlt, wpc;
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
  int off_cseg, off_cbeg, n_chunk;
  // continuous-time projection (GFMC_t, template TAU): the draws are made in the kernel because the number of
  // projections is not known in advance
  double tau;
  int tail, random_mesh;
  uint32_t* keys;  // [nw][2], advanced in place
  int* pc;         // [nw] projection counter (out in the main pass, in in the tail pass)
  int* n_max;      // [1] max over the walkers of the call (atomicMax in the main pass)
  const int* nn_fixed;  // optional [nw][n_e * NN]: nuclei of the non-local ECP of every electron given by the caller instead of the
                        // nearest-nucleus search (finite-difference derivatives keep the assignment of the base point fixed)
  long long* clk;  // optional [16] per-phase cycle counters (qe_set_phase_clocks; thread 0 of every CTA adds its own)
};

namespace {

// Determinant ratios of TWO mesh points (same spin block): sweep of all AOs at both points, contraction with the MO
// coefficients on the fly, dot product with the ratio weight vectors w0 / w1 (stride ws) of the two (walker, electron) pairs.
template <int NMO, bool CART, int LMAX, bool MIXED>
__device__ __noinline__ double2 mesh_ratio2(const char* __restrict__ tab, const BasisDev B, int coff, double px0, double py0, double pz0,
                                            double px1, double py1, double pz1, const double* __restrict__ w0,
                                            const double* __restrict__ w1, int ws) {
  const double px[2] = {px0, px1}, py[2] = {py0, py1}, pz[2] = {pz0, pz1};
  SinkMOn<NMO, 2> sink;
  sink.init(tab + coff);
  if constexpr (MIXED) eval_val_n_f32<CART, LMAX, 2>(tab, B, B.off_seg, px, py, pz, 0, B.n_grp, sink);  // zone ao_eval in fp32
  else eval_val_n<CART, LMAX, 2>(tab, B, B.off_seg, px, py, pz, 0, B.n_grp, sink);
  double r0 = 0.0, r1 = 0.0;
#pragma unroll
  for (int mo = 0; mo < NMO; ++mo) {
    r0 = fma(sink.acc[0][mo], w0[mo * ws], r0);
    r1 = fma(sink.acc[1][mo], w1[mo * ws], r1);
  }
  return make_double2(r0, r1);
}


// R^T (row-major) of R = Rz(gamma) Ry(beta) Rx(alpha), jqmc/jqmc_mcmc.py:4237-4244
__device__ __forceinline__ void rt_from_angles(double al, double be, double ga, double* RT) {
  double sa, ca, sb, cb, sg, cg;
  sincos(al, &sa, &ca);
  sincos(be, &sb, &cb);
  sincos(ga, &sg, &cg);
  // R rows written transposed (compile-time indices: the matrix stays in registers)
  RT[0] = cb * cg;
  RT[3] = cg * sa * sb - ca * sg;
  RT[6] = sa * sg + ca * cg * sb;
  RT[1] = cb * sg;
  RT[4] = ca * cg + sa * sb * sg;
  RT[7] = ca * sb * sg - cg * sa;
  RT[2] = -sb;
  RT[5] = cb * sa;
  RT[8] = ca * cb;
}

// shared-memory carve-up by byte offsets from the (16-byte aligned) dynamic shared memory base
struct Carve {
  char* base;
  size_t off;
  __device__ Carve(char* b, size_t o) : base(b), off(o) {}
  template <class T>
  __device__ T* take(size_t n) {
    off = (off + 15) & ~size_t(15);
    T* r = (T*)(base + off);
    off += n * sizeof(T);
    return r;
  }
};

template <class T>
__device__ __forceinline__ const T* stage(Carve& c, const T* src, size_t n, int tid, int nthr) {
  T* dst = c.take<T>(n);
  for (size_t i = tid; i < n; i += nthr) dst[i] = src[i];
  return dst;
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

size_t sys_bytes(const SysDev& s, int nmo) {
  size_t n = (size_t)(3 * s.n_atom + 3 * s.n_atom + nmo * nmo + nmo * (s.n_unp > 0 ? s.n_unp : 1)) * 8 + 6 * 16;
  if (s.ecp_flag) n += (size_t)s.n_ecp * (4 + 24) + (size_t)s.n_atom * 8 + 4 + (size_t)s.Nv * 32 + 8 * 16;
  return n;
}

// electron positions held in shared memory as [(e*3+c)][walker]
struct PosShared {
  const double* s_r;
  int wpc, wl;
  __device__ __forceinline__ void get(int e, double& x, double& y, double& z) const {
    x = s_r[(e * 3 + 0) * wpc + wl];
    y = s_r[(e * 3 + 1) * wpc + wl];
    z = s_r[(e * 3 + 2) * wpc + wl];
  }
};

template <int NMO, bool CART, int LMAX, bool TAU, int WS_T, bool MIXED>
__global__ void __launch_bounds__(512, 1)
k_walker(BasisDev B, SysDev S_g, WalkerArgs P) {
  extern __shared__ __align__(16) char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  // GFMC_t tail pass: walkers that ran out of time before the slowest walker of the call replay the no-move iterations the
  // reference's while_loop makes them run (jqmc/jqmc_gfmc.py:1539-1570): 3 key splits each and one last evaluation
  if constexpr (TAU) {
    if (P.tail) {
      const int n_max = *P.n_max;
      int need = 0;
      for (int wl = tid; wl < P.wpc; wl += nthr) need |= (n_max - P.pc[min(blockIdx.x * P.wpc + wl, P.nw - 1)]) > 0;
      if (!__syncthreads_or(need)) return;
    }
  }
  const int lane = tid & 31, wid = tid >> 5, NWARP = nthr >> 5;
  const int WPC = P.wpc;  // live walkers of this CTA
  // walker stride of the [item][walker] shared arrays: the compile-time constant 32 when the padded arrays fit (addresses
  // become immediate offsets: the index arithmetic was a quarter of the kernel's instructions), otherwise WPC (dense)
  const int WS = WS_T ? WS_T : P.wpc;
  const int w0 = blockIdx.x * WPC;
  // optional phase timing: thread 0 accumulates the cycles between consecutive PHASE() marks (placed after barriers)
  long long clk_last = 0;
  long long* clk_acc = nullptr;  // 12 counters in shared memory (carved below), touched by thread 0 only
  const bool clk_on = P.clk != nullptr && tid == 0;
#define PHASE(i)                          \
  if (clk_on) {                           \
    const long long t_ = clock64();       \
    clk_acc[i] += t_ - clk_last;          \
    clk_last = t_;                        \
  }

  // ---- stage tables ------------------------------------------------------------------------------
  {
    const int4* src = (const int4*)B.g;
    int4* dst = (int4*)smem_raw;
    for (int i = tid; i < B.bytes / 16; i += nthr) dst[i] = src[i];
  }
  const char* tab = smem_raw;
  Carve cv(smem_raw, (size_t)B.bytes);
  const SysDev S = stage_sys(S_g, NMO, cv, tid, nthr);
  const int N = S.n_up, Nd = S.n_dn, Ne = S.n_e, NN2 = N * N;
  const int n_kin = P.mode == 2 ? 0 : 6 * Ne;
  const int n_ecp = S.ecp_flag ? Ne * S.NN * S.Nv : 0;
  const int NPT = n_kin + n_ecp;

  double* s_r = cv.take<double>((size_t)Ne * 3 * WS);
  double* s_Gi = cv.take<double>((size_t)NN2 * WS);
  double* s_phi = cv.take<double>((size_t)Ne * 5 * NMO * WS);  // [(e*5+q)*NMO+mo]
  double* s_W = cv.take<double>((size_t)Ne * NMO * WS);
  double* s_el = cv.take<double>((size_t)Ne * 8 * WS);  // [e*8 + {ke|opt, ei, eid, loc, ee, kinFN, kinSP, -}]
  double* s_e2 = cv.take<double>((size_t)Ne * 2 * WS);  // [e*2 + {sum of the fixed-node ECP elements of electron e, of their positive parts}]
  // s_part (partial sums of the chunked value/grad/lap sweeps: start-up and P4) shares its storage with the mesh elements
  // s_p and the ECP Jastrow ratios s_j (P2 -> P3): the move has been selected before P4 overwrites them
  const int n_pj = (NPT > 0 ? NPT : 1) + (n_ecp > 0 ? n_ecp : 1);
  double* s_part = cv.take<double>((size_t)(NWARP * 5 * NMO > n_pj ? NWARP * 5 * NMO : n_pj) * WS);
  double* s_p = s_part;
  double* s_j = s_part + (size_t)(NPT > 0 ? NPT : 1) * WS;
  // non-local ECP data of every (electron, nn-th nearest nucleus), shared by its Nv quadrature points:
  // [0] nucleus index, [1] distance d, [2..4] unit vector nucleus -> electron, [5 + l] V_l(d) (2l+1) / d^2
  const int ECPW = 5 + S.ecp_lmax;
  double* s_ecp = cv.take<double>((size_t)(S.ecp_flag ? Ne * S.NN * ECPW : 1) * WS);
  double* s_stage = cv.take<double>((size_t)5 * NMO * WS);
  double* s_misc = cv.take<double>((size_t)16 * WS);  // 0..8 RT, 9..11 new position, 12 selected electron, 13 total
  int* s_ctr = cv.take<int>(4);
  int* s_act = cv.take<int>((size_t)WPC);  // GFMC_t: compact list of the walkers that still have time left
  uint32_t* s_key = cv.take<uint32_t>((size_t)2 * WPC);  // GFMC_n: running PRNG keys (the draws are made in the kernel)
  clk_acc = cv.take<long long>(12);
  if (clk_on) {
    for (int i = 0; i < 12; ++i) clk_acc[i] = 0;
    clk_last = clock64();
  }
#define SR(e, c) s_r[((e) * 3 + (c)) * WS + wl]
#define SGI(i, j) s_Gi[((i) * N + (j)) * WS + wl]
#define SPHI(e, q, mo) s_phi[(((e) * 5 + (q)) * NMO + (mo)) * WS + wl]
#define SW(e, mo) s_W[((e) * NMO + (mo)) * WS + wl]
#define SEL(e, i) s_el[((e) * 8 + (i)) * WS + wl]
#define SMISC(i) s_misc[(i) * WS + wl]
#define GW(wl_) (min(w0 + (wl_), P.nw - 1)) /* dead walkers of the last CTA shadow the last walker (no stores) */

  for (int idx = tid; idx < Ne * 3 * WPC; idx += nthr) {
    const int wl = idx % WPC, it = idx / WPC, e = it / 3, c = it % 3;
    const int ww = GW(wl);
    s_r[it * WS + wl] = e < N ? P.r_up[((size_t)ww * N + e) * 3 + c] : P.r_dn[((size_t)ww * Nd + (e - N)) * 3 + c];
  }
  for (int idx = tid; idx < NN2 * WPC; idx += nthr) {
    const int wl = idx % WPC, it = idx / WPC;
    s_Gi[it * WS + wl] = P.Ginv[(size_t)GW(wl) * NN2 + it];
  }
  if (P.mode != 0)
    for (int idx = tid; idx < 9 * WPC; idx += nthr) {
      const int wl = idx % WPC, c = idx / WPC;
      s_misc[c * WS + wl] = P.RT_in ? P.RT_in[(size_t)GW(wl) * 9 + c] : (c % 4 == 0 ? 1.0 : 0.0);
    }
  if (tid == 0) s_ctr[0] = 0;
  if (!TAU && P.mode == 0)
    for (int wl = tid; wl < WPC; wl += nthr) {
      s_key[2 * wl] = P.keys[2 * GW(wl)];
      s_key[2 * wl + 1] = P.keys[2 * GW(wl) + 1];
    }
  if constexpr (TAU)
    for (int wl = tid; wl < WPC; wl += nthr) {
      s_act[wl] = wl;
      s_misc[14 * WS + wl] = 1.0;  // has time left
    }
  __syncthreads();

  // ---- value/grad/lap of the MOs at every electron (cache): task = (walker, electron, half of the basis chunks); the two
  //      halves land in the cache and in the scratch of the chunked sweep and are added in a second pass (2 N_e WPC tasks
  //      fill the CTA; one task per (walker, electron) left more than half of the warps idle during the longest sweep) -----
  {
    const int* cbeg = (const int*)(tab + P.off_cbeg);
    const bool split = Ne <= NWARP && P.n_chunk >= 2;  // the scratch holds NWARP x 5 NMO WPC values
    const int c_mid = split ? P.n_chunk / 2 : P.n_chunk;
    for (int s0 = tid; s0 < (split ? 2 : 1) * Ne * WPC; s0 += nthr) {
      const int half = s0 / (Ne * WPC), s = s0 % (Ne * WPC);
      const int wl = s % WPC, e = s / WPC;
      SinkMO5<NMO> sink;
      sink.init(tab + (e < N ? B.off_C : B.off_C2));
      eval_vgl<CART, LMAX>(tab, B, P.off_cseg, SR(e, 0), SR(e, 1), SR(e, 2), half ? cbeg[c_mid] : cbeg[0],
                           half ? cbeg[P.n_chunk] : cbeg[c_mid], sink);
      double* dst = half ? s_part : s_phi;  // same [(e*5+q)*NMO+mo][walker] layout (s_part holds 16 x 5 NMO WPC >= N_e x 5 NMO WPC)
#pragma unroll
      for (int q = 0; q < 5; ++q)
#pragma unroll
        for (int mo = 0; mo < NMO; ++mo) dst[(((e) * 5 + q) * NMO + mo) * WS + wl] = sink.acc[q][mo];
    }
    __syncthreads();
    if (split)
      for (int i = tid; i < Ne * 5 * NMO * WS; i += nthr) s_phi[i] += s_part[i];  // (padding slots included: never read)
  }
  __syncthreads();
  PHASE(0)

  const double a2 = P.alat * P.alat;
  const int n_it = P.mode == 0 ? P.nmpm : 1;
  double w_L = 1.0, diag = 0.0, nondiag = 0.0;  // meaningful in threads tid < WPC (wl = tid)
  if (P.mode == 0 && tid < WPC && !(TAU && P.tail)) w_L = P.w[GW(tid)];
  // GFMC_t per-walker registers (threads tid < WPC)
  Key key{0u, 0u};
  double tau_left = 0.0, xi = 0.0, u_move = 0.0;
  int pc = 0, n_done = 0, n_extra = 0;
  if constexpr (TAU) {
    if (tid < WPC) {
      const int ww = GW(tid);
      key = Key{P.keys[2 * ww], P.keys[2 * ww + 1]};
      if (!P.tail) tau_left = P.tau;
      else {
        n_extra = *P.n_max - P.pc[ww];
        for (int i = 0; i < 3 * (n_extra - 1); ++i) {
          Key sub;
          rng_split(key, sub);
        }
      }
    }
  }
  // walker enumeration of the loop phases: all WPC walkers, or (GFMC_t) the n_act walkers that still have time left
  int n_act = WPC;
#define NACT (TAU ? n_act : WPC)
#define WLOF(c) (TAU ? s_act[c] : (c))
  // mesh-point slot blocks (same spin => same MO coefficient table): kinetic up, kinetic dn, ECP up, ECP dn
  const int n_eu = S.ecp_flag ? N * S.NN * S.Nv : 0, n_ed = S.ecp_flag ? Nd * S.NN * S.Nv : 0;
  const int n_ku = P.mode == 2 ? 0 : 6 * N, n_kd = P.mode == 2 ? 0 : 6 * Nd;

  for (int it = 0; TAU || it < n_it; ++it) {
    if constexpr (TAU) {
      // compact the walkers that still have time left (flag written at the end of the previous iteration)
      if (tid == 0) {
        int n = 0;
        for (int wl = 0; wl < WPC; ++wl)
          if (s_misc[14 * WS + wl] != 0.0) s_act[n++] = wl;
        s_ctr[1] = n;
      }
      __syncthreads();
      n_act = s_ctr[1];
    }
    // (scalars, not arrays: a run-time block index would put arrays into local memory)
    const int bs0 = n_ku * NACT, bs1 = n_kd * NACT, bs2 = n_eu * NACT, bs3 = n_ed * NACT;
    const int bp0 = (bs0 + 1) / 2, bp1 = (bs1 + 1) / 2, bp2 = (bs2 + 1) / 2, bp3 = (bs3 + 1) / 2;
    const int n_rounds_pt = (bp0 + bp1 + bp2 + bp3 + 31) / 32;
    const int n_rounds_el = (Ne * NACT + 31) / 32;
    // ---- P1: ratio weight vectors, task = (walker, electron) -----------------------------------------------------
    if constexpr (TAU) {
      // three splits per projection: rotation angles, time draw, move draw (jqmc/jqmc_gfmc.py:770-776, 1004-1018)
      if (tid < WPC && (P.tail || tau_left > 0.0)) {
        const int wl = tid;
        Key sub;
        rng_split(key, sub);
        double al = 0, be = 0, ga = 0;
        if (P.random_mesh) {
          const double two_pi = 6.283185307179586;
          al = rng_uniform_bits(rng_bits64(sub, 0u), -two_pi, two_pi);
          be = rng_uniform_bits(rng_bits64(sub, 1u), -two_pi, two_pi);
          ga = rng_uniform_bits(rng_bits64(sub, 2u), -two_pi, two_pi);
        }
        double RTl[9];
        rt_from_angles(al, be, ga, RTl);
#pragma unroll
        for (int c = 0; c < 9; ++c) SMISC(c) = RTl[c];
        rng_split(key, sub);
        xi = rng_uniform_bits(rng_bits64(sub), 0.0, 1.0);
        rng_split(key, sub);
        u_move = rng_uniform_bits(rng_bits64(sub), 0.0, 1.0);
      }
    } else if (P.mode == 0) {
      // GFMC_n: two splits per projection, rotation angles and move draw (_split_step_keys, jqmc/jqmc_gfmc.py:5275-5283),
      // made by the first WPC threads while warps 1.. work on P1 (no draw tables in HBM: the launch moves walker state only)
      if (tid < WPC) {
        const int wl = tid;
        Key key{s_key[2 * wl], s_key[2 * wl + 1]}, sub;
        rng_split(key, sub);
        double al = 0, be = 0, ga = 0;
        if (P.random_mesh) {
          const double two_pi = 6.283185307179586;
          al = rng_uniform_bits(rng_bits64(sub, 0u), -two_pi, two_pi);
          be = rng_uniform_bits(rng_bits64(sub, 1u), -two_pi, two_pi);
          ga = rng_uniform_bits(rng_bits64(sub, 2u), -two_pi, two_pi);
        }
        double RTl[9];
        rt_from_angles(al, be, ga, RTl);
#pragma unroll
        for (int c = 0; c < 9; ++c) SMISC(c) = RTl[c];
        rng_split(key, sub);
        SMISC(15) = rng_uniform_bits(rng_bits64(sub), 0.0, 1.0);
        s_key[2 * wl] = key.a;
        s_key[2 * wl + 1] = key.b;
      }
    }
    // (in the projection modes the tasks start at warp 1: the first warp is busy with the draws)
    const int p1_off = (P.mode == 0 && nthr > 32) ? 32 : 0;
    // two task kinds per (walker, electron): the ratio weight vector, and the Jastrow terms at the current position
    const int n_etask = S.ecp_flag ? Ne * S.NN * NACT : 0;
    for (int s0 = tid - p1_off; s0 < 2 * Ne * NACT + n_etask; s0 += nthr - p1_off) {
      if (s0 < 0) continue;
      if (s0 >= 2 * Ne * NACT) {
        // non-local ECP radial factors of (electron, nn): nearest-nucleus search, distance and the channel sums are done
        // once here instead of at each of the Nv quadrature points (jqmc/coulomb_potential.py:1562-1575, 1607-1645)
        const int s = s0 - 2 * Ne * NACT;
        const int wl = WLOF(s % NACT), t = s / NACT, nn = t % S.NN, e = t / S.NN;
        const double x = SR(e, 0), y = SR(e, 1), z = SR(e, 2);
        double d;
        const int a = P.nn_fixed ? P.nn_fixed[((size_t)GW(wl) * Ne + e) * S.NN + nn] : nearest_atom(S.Rn, S.n_atom, x, y, z, nn, &d);
        const double relx = S.Rn[3 * a] - x, rely = S.Rn[3 * a + 1] - y, relz = S.Rn[3 * a + 2] - z;
        d = sqrt(relx * relx + rely * rely + relz * relz);
        double* o = s_ecp + (size_t)t * ECPW * WS + wl;
        o[0] = (double)a;
        o[WS] = d;
        o[2 * WS] = -relx / d;
        o[3 * WS] = -rely / d;
        o[4 * WS] = -relz / d;
        const int lloc = S.ecp_lmax_atom[a];
        for (int l = 0; l < S.ecp_lmax; ++l) {
          double vl = 0.0;
          if (l < lloc)
            for (int kk = S.ecp_off[a]; kk < S.ecp_off[a + 1]; ++kk)
              if (S.ecp_l[kk] == l) vl += S.ecp_c[kk] * ipow(d, S.ecp_p[kk]) * qexp(-S.ecp_z[kk] * d * d);
          o[(5 + l) * WS] = vl / (d * d) * (2 * l + 1);
        }
        continue;
      }
      const bool jtask = s0 >= Ne * NACT;
      const int s = jtask ? s0 - Ne * NACT : s0;
      const int wl = WLOF(s % NACT), e = s / NACT;
      if (jtask) {
        // Jastrow terms of electron e at its current position (shared by all of its mesh points)
        PosShared pos{s_r, WS, wl};
        if constexpr (MIXED) SEL(e, 7) = (double)jastrow_single_f32(S, pos, e, SR(e, 0), SR(e, 1), SR(e, 2));
        else SEL(e, 7) = jastrow_single_m(S, pos, e, SR(e, 0), SR(e, 1), SR(e, 2));
        continue;
      }
      double Wv[NMO];
      if (e < N) {
        double y[NMO];
#pragma unroll
        for (int b = 0; b < NMO; ++b) {
          double sum = 0;
          for (int j = 0; j < Nd; ++j) sum = fma(SPHI(N + j, 0, b), SGI(j, e), sum);
          y[b] = sum;
        }
#pragma unroll
        for (int a = 0; a < NMO; ++a) {
          double sum = 0;
#pragma unroll
          for (int b = 0; b < NMO; ++b) sum = fma(S.lam_p[a * NMO + b], y[b], sum);
          for (int k = 0; k < S.n_unp; ++k) sum = fma(S.lam_u[a * S.n_unp + k], SGI(Nd + k, e), sum);
          Wv[a] = sum;
        }
      } else {
        const int j = e - N;
        double y[NMO];
#pragma unroll
        for (int a = 0; a < NMO; ++a) {
          double sum = 0;
          for (int i = 0; i < N; ++i) sum = fma(SPHI(i, 0, a), SGI(j, i), sum);
          y[a] = sum;
        }
#pragma unroll
        for (int b = 0; b < NMO; ++b) {
          double sum = 0;
#pragma unroll
          for (int a = 0; a < NMO; ++a) sum = fma(y[a], S.lam_p[a * NMO + b], sum);
          Wv[b] = sum;
        }
      }
#pragma unroll
      for (int mo = 0; mo < NMO; ++mo) SW(e, mo) = Wv[mo];
    }
    __syncthreads();
    PHASE(1)

    // ---- P2: mesh ratios (rounds 0 .. n_rounds_pt-1) and per-electron terms (the following n_rounds_el rounds); a warp
    //      takes the next round from a shared counter.  A thread evaluates TWO mesh points of the same spin block at once
    //      (same shell / primitive sequence, coefficient rows loaded once, twice the independent DFMA chains) ------------------
    for (;;) {
      int r = 0;
      if (lane == 0) r = atomicAdd(&s_ctr[0], 1);
      r = __shfl_sync(0xffffffffu, r, 0);
      if (r >= n_rounds_pt + n_rounds_el) break;
      if (r < n_rounds_pt) {
        int pi = r * 32 + lane;  // pair index -> block (kin up, kin dn, ecp up, ecp dn)
        int blk = 0, half = bp0, bstart = 0, bsize = bs0;
        if (pi >= half) { pi -= half; blk = 1; half = bp1; bstart = bs0; bsize = bs1;
          if (pi >= half) { pi -= half; blk = 2; half = bp2; bstart = n_kin * NACT; bsize = bs2;
            if (pi >= half) { pi -= half; blk = 3; half = bp3; bstart = (n_kin + n_eu) * NACT; bsize = bs3;
              if (pi >= half) blk = 4; } } }
        if (blk < 4) {
          const int sA = bstart + pi;
          const bool validB = pi + half < bsize;
          const int sB = validB ? sA + half : sA;
          int sl[2] = {sA, sB};
          double px[2], py[2], pz[2], angw[2], jold[2];
          int el[2], wls[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int wl = WLOF(sl[i] % NACT), t = sl[i] / NACT;
            wls[i] = wl;
            sl[i] = t * WS + wl;  // storage slot
            PosShared pos{s_r, WS, wl};
            double x, y, z;
            angw[i] = 0.0;
            if (blk < 2) {
              const int e = t / 6;
              const int s6 = t % 6, ax = s6 >> 1;
              const double sg = (s6 & 1) ? -P.alat : P.alat;
              pos.get(e, x, y, z);
              px[i] = x + sg * SMISC(3 * ax);
              py[i] = y + sg * SMISC(3 * ax + 1);
              pz[i] = z + sg * SMISC(3 * ax + 2);
              el[i] = e;
            } else {
              const int pt = t - n_kin;
              const int k = pt % S.Nv, en = pt / S.Nv;  // en = e * NN + nn
              el[i] = en / S.NN;
              const double* o = s_ecp + (size_t)en * ECPW * WS + wl;
              const int a = (int)o[0];
              const double d = o[WS];
              // rotated quadrature direction g = q RT (jqmc/coulomb_potential.py:1547), point R_a + d g (:1607-1610)
              const double q0 = S.quad_g[3 * k], q1 = S.quad_g[3 * k + 1], q2 = S.quad_g[3 * k + 2];
              const double gx = q0 * SMISC(0) + q1 * SMISC(3) + q2 * SMISC(6);
              const double gy = q0 * SMISC(1) + q1 * SMISC(4) + q2 * SMISC(7);
              const double gz = q0 * SMISC(2) + q1 * SMISC(5) + q2 * SMISC(8);
              px[i] = fma(d, gx, S.Rn[3 * a]);
              py[i] = fma(d, gy, S.Rn[3 * a + 1]);
              pz[i] = fma(d, gz, S.Rn[3 * a + 2]);
              // cos(theta) between nucleus -> electron and the quadrature direction (:1643-1645), channel sum (:1573-1575)
              const double cos_t = (o[2 * WS] * gx + o[3 * WS] * gy + o[4 * WS] * gz) * mrsqrt(gx * gx + gy * gy + gz * gz);
              double ang = 0.0;
              for (int l = 0; l < S.ecp_lmax; ++l) ang = fma(o[(5 + l) * WS], legendre_l(l, cos_t), ang);
              angw[i] = ang * S.quad_w[k];
            }
            jold[i] = SEL(el[i], 7);
          }
          // AO sweep + MO contraction + dot with the ratio weight vectors: its own function, NOT inlined, so that it gets a
          // register allocation of its own (no spills inside) instead of sharing the kernel's
          const double2 rr = mesh_ratio2<NMO, CART, LMAX, MIXED>(tab, B, (blk & 1) ? B.off_C2 : B.off_C, px[0], py[0], pz[0], px[1], py[1],
                                                                 pz[1], s_W + (el[0] * NMO) * WS + wls[0], s_W + (el[1] * NMO) * WS + wls[1], WS);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int wl = wls[i], e = el[i];
            PosShared pos{s_r, WS, wl};
            const double ratio = i == 0 ? rr.x : rr.y;
            double jr;
            if constexpr (MIXED) jr = (double)expf(jastrow_single_f32(S, pos, e, px[i], py[i], pz[i]) - (float)jold[i]);  // zone jastrow_ratio
            else jr = qexp(jastrow_single_m(S, pos, e, px[i], py[i], pz[i]) - jold[i]);
            if (i == 0 || validB) {
              if (blk < 2) {
                s_p[sl[i]] = -1.0 / (2.0 * a2) * (ratio * jr);
              } else {
                s_p[sl[i]] = P.dlt ? angw[i] * ratio : angw[i] * (ratio * jr);
                s_j[sl[i] - n_kin * WS] = jr;
              }
            }
          }
        }
      } else {
        // per-electron: continuum kinetic energy, bare / discretised el-ion, ECP local, el-el (pairs j > e)
        const int s = (r - n_rounds_pt) * 32 + lane;
        if (s < Ne * NACT) {
          const int wl = WLOF(s % NACT), e = s / NACT;
          PosShared pos{s_r, WS, wl};
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
                const double ex = qexp(-aa * c * rs);
                fp = -A * (c * 0.5) * ex;
                lJ += A * (aa * c * c * 0.5) * ex - A * c * ex / rs;
              } else {
                const double den = 1.0 + aa * c * rs;
                fp = -A / (2.0 * den * den);
                lJ += A * aa * c / (den * den * den) + 2.0 * fp / rs;
              }
              const double sc = fp / rs;
              gJ[0] = fma(sc, dx, gJ[0]);
              gJ[1] = fma(sc, dy, gJ[1]);
              gJ[2] = fma(sc, dz, gJ[2]);
            }
            if (S.ecp_flag) {
              const int lloc = S.ecp_lmax_atom[a];
              double sum = 0.0;
              for (int k = S.ecp_off[a]; k < S.ecp_off[a + 1]; ++k)
                if (S.ecp_l[k] == lloc) sum += S.ecp_c[k] * ipow(d, S.ecp_p[k]) * qexp(-S.ecp_z[k] * d * d);
              loc += sum / (d * d);
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
                const double ex = qexp(-aa * rs);
                fp = 0.5 * ex;
                lJ += -(aa * 0.5) * ex + 2.0 * fp / rs;
              }
              const double sc = fp / rs;
              gJ[0] = fma(sc, dx, gJ[0]);
              gJ[1] = fma(sc, dy, gJ[1]);
              gJ[2] = fma(sc, dz, gJ[2]);
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
    }
    __syncthreads();
    PHASE(2)
    if (tid == 0) s_ctr[0] = 0;  // the next P2 starts after at least one more barrier

    // ---- P3: assemble -----------------------------------------------------------------------------------------------
    if (P.mode == 2) {
      if (tid < WPC) {
        const int wl = tid, w = w0 + wl;
        const bool live = w < P.nw;
        double T = 0, vbare = S.v_ion_ion, vl = 0, vnl = 0;
        for (int e = 0; e < Ne; ++e) {
          T += SEL(e, 0);
          vbare += SEL(e, 1) + SEL(e, 4);
          vl += SEL(e, 3);
          if (P.T_elem && live) P.T_elem[(size_t)w * Ne + e] = SEL(e, 0);
        }
        for (int k = 0; k < n_ecp; ++k) vnl += s_p[k * WS + wl];
        if (live) {
          P.e_L[w] = T + (vbare + (vl + vnl));
          if (P.V_parts) {
            P.V_parts[(size_t)w * 4 + 0] = vbare;
            P.V_parts[(size_t)w * 4 + 1] = vl;
            P.V_parts[(size_t)w * 4 + 2] = vnl;
            P.V_parts[(size_t)w * 4 + 3] = 0.0;
          }
        }
      }
      break;
    }
    // (a) per (walker, electron): fixed-node split of its 6 kinetic elements, regularised el-ion term
    //     (jqmc/jqmc_gfmc.py:4829-4939); the FN values overwrite s_p
    for (int s = tid; s < Ne * NACT; s += nthr) {
      const int wl = WLOF(s % NACT), e = s / NACT;
      bool flip = false;
      double nd = 0, kinFN = 0, kinSP = 0;
#pragma unroll 1  // (the short loops of P3 stay rolled: the phase runs once per projection on cold instruction caches, and every
                  //  instruction line it does not have to fetch counts; A/B on one box: projection 4.240 -> 4.203 ms)
      for (int s6 = 0; s6 < 6; ++s6) {
        const double v = s_p[(6 * e + s6) * WS + wl];
        flip = flip || (v >= 0.0);
        nd += v + 1.0 / (4.0 * a2);
        const double fn = fmin(v, 0.0);
        kinFN += fn;
        kinSP += fmax(v, 0.0);
        s_p[(6 * e + s6) * WS + wl] = fn;
      }
      const double zv = SEL(e, 1) + SEL(e, 0) - nd;
      const double eib = S.ecp_flag ? SEL(e, 1) : SEL(e, 2);
      SEL(e, 0) = flip ? fmax(zv, eib) : zv;  // regularised el-ion term of this electron
      SEL(e, 5) = kinFN;
      SEL(e, 6) = kinSP;
    }
    // (a') per (walker, electron): fixed-node split of its NN*Nv non-local elements (s_j: Jastrow ratio -> positive part) and
    //      their sums.  The 6 kinetic and the NN*Nv ECP elements of an electron are contiguous in the move vector
    //      [kinetic mesh, ECP mesh], so these per-electron sums are also the chunk sums of the move selection below
    if (n_ecp > 0) {
      const int per = S.NN * S.Nv;
      // (these tasks start at the upper half of the CTA when both task lists fit side by side: (a) and (a') then run on
      //  different warps at the same time)
      const int off2 = 2 * Ne * NACT <= nthr ? nthr - Ne * NACT : 0;
      for (int c = tid - off2; c < Ne * NACT; c += nthr) {
        if (c < 0) continue;
        const int wl = WLOF(c % NACT), e = c / NACT;
        double sFN = 0, sSP = 0;
#pragma unroll 1
        for (int k = e * per; k < (e + 1) * per; ++k) {
          const double v = s_p[(n_kin + k) * WS + wl];
          double fn = fmin(v, 0.0);
          if (P.dlt) fn *= s_j[k * WS + wl];
          const double sp = fmax(v, 0.0);
          s_p[(n_kin + k) * WS + wl] = fn;
          s_j[k * WS + wl] = sp;
          sFN += fn;
          sSP += sp;
        }
        s_e2[(e * 2 + 0) * WS + wl] = sFN;
        s_e2[(e * 2 + 1) * WS + wl] = sSP;
      }
    }
    __syncthreads();
    PHASE(3)
    // (b) per walker: sums, weight, normalisation of the move probabilities
    if (tid < WPC && (!TAU || P.tail || tau_left > 0.0)) {
      const int wl = tid;
      const double diag_kin = 3.0 / (2.0 * a2) * Ne;
      double sum_kinFN = 0, SP_kin = 0, sum_opt = 0, ee = 0, loc = 0;
#pragma unroll 1
      for (int e = 0; e < Ne; ++e) {
        sum_kinFN += SEL(e, 5);
        SP_kin += SEL(e, 6);
        sum_opt += SEL(e, 0);
        ee += SEL(e, 4);
        loc += SEL(e, 3);
      }
      const double disc_bare = ee + S.v_ion_ion + sum_opt;
      double sum_eFN = 0, SP_e = 0;
      if (n_ecp > 0)
        for (int e = 0; e < Ne; ++e) {
          sum_eFN += s_e2[(e * 2 + 0) * WS + wl];
          SP_e += s_e2[(e * 2 + 1) * WS + wl];
        }
      nondiag = sum_kinFN + sum_eFN;
      diag = S.ecp_flag ? diag_kin + disc_bare + loc + SP_kin + SP_e : diag_kin + disc_bare + SP_kin;
      if constexpr (TAU) {
        // time spent in this configuration, weight, remaining time (jqmc/jqmc_gfmc.py:1003-1012); the move is suppressed
        // once the time is used up (:1024-1025)
        const double e_L = diag + nondiag;
        if (tau_left > 0.0) ++pc;
        const double tau_update = fmin(tau_left, log(1.0 - xi) / nondiag);
        w_L *= qexp(-tau_update * e_L);
        tau_left -= tau_update;
        SMISC(14) = tau_left <= 0.0 ? 0.0 : 1.0;
        SMISC(13) = nondiag;  // sum of all fixed-node elements = normalisation of the move probabilities
      } else if (P.mode == 0) {
        const double b_x = 1.0 / (diag - P.E_scf) * (-nondiag);
        w_L *= b_x;
        SMISC(13) = nondiag;
      }
    }
    if constexpr (TAU) {
      n_done = it + 1;
      const int mv = (tid < WPC) ? (s_misc[14 * WS + tid] != 0.0) : 0;
      if (!__syncthreads_or(mv)) break;  // every walker of the CTA has used up its time
    } else {
      // (no barrier: (d) below runs on the same threads as (b) and reads, besides its own thread's results, only what was
      //  complete at the previous barrier)
      if (P.mode != 0) break;
    }
    PHASE(4)
    // (d) per walker: first index whose cumulative probability reaches u (searchsorted 'left' on cumsum(p / sum p),
    //     jqmc/jqmc_gfmc.py:5057-5062): skip whole electrons by their chunk sums, then scan element by element; the scan
    //     runs on past the chunk if round-off moved the crossing
    if (tid < WPC && (!TAU || s_misc[14 * WS + tid] != 0.0)) {
      const int wl = tid;
      PosShared pos{s_r, WS, wl};
      const double u = TAU ? u_move : SMISC(15);
      const double tot = SMISC(13);
      const int per = S.ecp_flag ? S.NN * S.Nv : 0, n_ch = n_ecp > 0 ? 2 * Ne : Ne;
      int ksel = NPT - 1, kstart = 0;
      double c = 0;
#pragma unroll 1
      for (int ci = 0; ci < n_ch; ++ci) {
        const double cs = (ci < Ne ? SEL(ci, 5) : s_e2[((ci - Ne) * 2) * WS + wl]) / tot;
        if (c + cs >= u) break;
        c += cs;
        kstart = ci + 1 < Ne ? 6 * (ci + 1) : n_kin + (ci + 1 - Ne) * per;
      }
#pragma unroll 1
      for (int k = kstart; k < NPT; ++k) {
        c += s_p[k * WS + wl] / tot;
        if (c >= u) {
          ksel = k;
          break;
        }
      }
      int e;
      double x, y, z, px, py, pz;
      if (ksel < n_kin) {
        e = ksel / 6;
        const int s6 = ksel % 6, ax = s6 >> 1;
        const double sg = (s6 & 1) ? -P.alat : P.alat;
        pos.get(e, x, y, z);
        px = x + sg * SMISC(3 * ax);
        py = y + sg * SMISC(3 * ax + 1);
        pz = z + sg * SMISC(3 * ax + 2);
      } else {
        const int pt = ksel - n_kin;
        const int k = pt % S.Nv, en = pt / S.Nv;
        e = en / S.NN;
        const double* o = s_ecp + (size_t)en * ECPW * WS + wl;
        const int a = (int)o[0];
        const double d = o[WS];
        const double q0 = S.quad_g[3 * k], q1 = S.quad_g[3 * k + 1], q2 = S.quad_g[3 * k + 2];
        px = fma(d, q0 * SMISC(0) + q1 * SMISC(3) + q2 * SMISC(6), S.Rn[3 * a]);
        py = fma(d, q0 * SMISC(1) + q1 * SMISC(4) + q2 * SMISC(7), S.Rn[3 * a + 1]);
        pz = fma(d, q0 * SMISC(2) + q1 * SMISC(5) + q2 * SMISC(8), S.Rn[3 * a + 2]);
      }
      SMISC(9) = px;
      SMISC(10) = py;
      SMISC(11) = pz;
      SMISC(12) = (double)e;
    }
    __syncthreads();
    PHASE(5)

    // ---- P4: value/grad/lap of the moved electron: warp = basis chunks wid, wid+NWARP, ..., lane = walker --------------
    {
      const int* cbeg = (const int*)(tab + P.off_cbeg);
      for (int wl = lane; wl < WPC; wl += 32) {
        if (TAU && SMISC(14) == 0.0) continue;  // out of time: no move
        const int es = (int)SMISC(12);
        SinkMO5<NMO> sink;
        sink.init(tab + (es < N ? B.off_C : B.off_C2));
        for (int c = wid; c < P.n_chunk; c += NWARP)
          eval_vgl<CART, LMAX>(tab, B, P.off_cseg, SMISC(9), SMISC(10), SMISC(11), cbeg[c], cbeg[c + 1], sink);
#pragma unroll
        for (int q = 0; q < 5; ++q)
#pragma unroll
          for (int mo = 0; mo < NMO; ++mo) s_part[((wid * 5 + q) * NMO + mo) * WS + wl] = sink.acc[q][mo];
      }
    }
    __syncthreads();
    PHASE(6)
    for (int s = tid; s < 5 * NMO * WS; s += nthr) {  // fixed-order sum over the warps' partials
      double sum = 0;
      for (int c = 0; c < NWARP; ++c) sum += s_part[(size_t)c * 5 * NMO * WS + s];
      s_stage[s] = sum;
    }
    __syncthreads();
    PHASE(7)
    // Sherman-Morrison (jqmc/jqmc_gfmc.py:5083-5141): task = (walker, row i); read phase, barrier, write phase
    {  // N * WPC <= blockDim.x is guaranteed by launch_walker: one pass.  All loops over electrons run to the compile-time
       // bound NB >= N with uniform predicates, so that newrow / vvec / uvec stay in registers (no local memory); NB = 4
       // for the 4-orbital instantiation (half the instructions)
      constexpr int NB = NMO <= 4 ? 4 : 8;  // NMO = 4 is dispatched only for n_up <= 4 (qe_create: nmo_pad)
      double newrow[NB];
      const int s = tid;
      bool act = s < N * WPC;
      int wl = 0, i = 0;
      if (act) {
        wl = s % WPC;
        i = s / WPC;
        if (TAU && SMISC(14) == 0.0) act = false;  // walker out of time: no move
      }
      if (act) {
        const int es = (int)SMISC(12);
        double pn[NMO];
#pragma unroll
        for (int mo = 0; mo < NMO; ++mo) pn[mo] = s_stage[mo * WS + wl] - SPHI(es, 0, mo);  // phi_new - phi_old
        if (es < N) {
          const int k = es;
          double t[NMO], vvec[NB];
#pragma unroll
          for (int b = 0; b < NMO; ++b) {
            double sum = 0;
#pragma unroll
            for (int a = 0; a < NMO; ++a) sum = fma(pn[a], S.lam_p[a * NMO + b], sum);
            t[b] = sum;
          }
          double acc = 0;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            double sum = 0;
            if (j < Nd) {
#pragma unroll
              for (int b = 0; b < NMO; ++b) sum = fma(t[b], SPHI(N + j, 0, b), sum);
            } else if (j < N) {
              const int q = j - Nd;
#pragma unroll
              for (int a = 0; a < NMO; ++a) sum = fma(pn[a], S.lam_u[a * S.n_unp + q], sum);
            }
            vvec[j] = sum;
            if (j < N) acc = fma(sum, SGI(j, k), acc);
          }
          const double invD = 1.0 / (1.0 + acc);
          const double coli = SGI(i, k);
#pragma unroll
          for (int jp = 0; jp < NB; ++jp) {
            double vt = 0;
            if (jp < N) {
#pragma unroll
              for (int j = 0; j < NB; ++j)
                if (j < N) vt = fma(vvec[j], SGI(j, jp), vt);
              vt = SGI(i, jp) - (coli * vt) * invD;
            }
            newrow[jp] = vt;
          }
        } else {
          const int k = es - N;
          double t[NMO], uvec[NB];
#pragma unroll
          for (int a = 0; a < NMO; ++a) {
            double sum = 0;
#pragma unroll
            for (int b = 0; b < NMO; ++b) sum = fma(S.lam_p[a * NMO + b], pn[b], sum);
            t[a] = sum;
          }
#pragma unroll
          for (int ii = 0; ii < NB; ++ii) {
            double sum = 0;
            if (ii < N) {
#pragma unroll
              for (int a = 0; a < NMO; ++a) sum = fma(SPHI(ii, 0, a), t[a], sum);
            }
            uvec[ii] = sum;
          }
          double au_i = 0, au_k = 0;
#pragma unroll
          for (int j = 0; j < NB; ++j)
            if (j < N) {
              au_i = fma(SGI(i, j), uvec[j], au_i);
              au_k = fma(SGI(k, j), uvec[j], au_k);
            }
          const double invD = 1.0 / (1.0 + au_k);
#pragma unroll
          for (int j = 0; j < NB; ++j) newrow[j] = j < N ? SGI(i, j) - (au_i * SGI(k, j)) * invD : 0.0;
        }
      }
      __syncthreads();
      if (act) {
#pragma unroll
        for (int j = 0; j < NB; ++j)
          if (j < N) SGI(i, j) = newrow[j];
      }
    }
    __syncthreads();
    PHASE(8)
    for (int s = tid; s < 5 * NMO * WS; s += nthr) {
      const int wl = s % WS, item = s / WS;
      if (wl >= WPC) continue;  // padding slot
      if (TAU && SMISC(14) == 0.0) continue;
      const int es = (int)SMISC(12);
      s_phi[(es * 5 * NMO + item) * WS + wl] = s_stage[s];
    }
    if (tid < WPC && !(TAU && s_misc[14 * WS + tid] == 0.0)) {
      const int wl = tid;
      const int es = (int)SMISC(12);
      SR(es, 0) = SMISC(9);
      SR(es, 1) = SMISC(10);
      SR(es, 2) = SMISC(11);
    }
    __syncthreads();
    PHASE(9)
  }

  // ---- write back -------------------------------------------------------------------------------------------
  if (clk_on) {
    PHASE(10)
    for (int i = 0; i < 12; ++i) atomicAdd((unsigned long long*)&P.clk[i], (unsigned long long)clk_acc[i]);
  }
  if (P.mode == 2) return;
  if constexpr (TAU) {
    if (tid < WPC && w0 + tid < P.nw && (!P.tail || n_extra > 0)) {
      const int wl = tid, w = w0 + tid;
      P.e_L[w] = diag + nondiag;
      for (int c = 0; c < 9; ++c) P.RT_out[(size_t)w * 9 + c] = SMISC(c);
      P.keys[2 * w] = key.a;
      P.keys[2 * w + 1] = key.b;
      if (!P.tail) {
        P.w[w] = w_L;
        P.pc[w] = pc;
      }
    }
    if (P.tail) return;
    if (tid == 0) atomicMax(P.n_max, n_done);
  } else if (tid < WPC && w0 + tid < P.nw) {
    const int wl = tid, w = w0 + tid;
    P.V_diag[w] = diag;
    P.V_nondiag[w] = nondiag;
    if (P.mode == 0) {
      P.w[w] = w_L;
      for (int c = 0; c < 9; ++c) P.RT_out[(size_t)w * 9 + c] = SMISC(c);
      P.keys[2 * w] = s_key[2 * wl];
      P.keys[2 * w + 1] = s_key[2 * wl + 1];
    }
  }
  if (P.mode == 0) {
    for (int idx = tid; idx < Ne * 3 * WPC; idx += nthr) {
      const int wl = idx % WPC, it = idx / WPC, e = it / 3, c = it % 3;
      const int w = w0 + wl;
      if (w >= P.nw) continue;
      if (e < N) P.r_up[((size_t)w * N + e) * 3 + c] = s_r[it * WS + wl];
      else P.r_dn[((size_t)w * Nd + (e - N)) * 3 + c] = s_r[it * WS + wl];
    }
    for (int idx = tid; idx < NN2 * WPC; idx += nthr) {
      const int wl = idx % WPC, it = idx / WPC;
      if (w0 + wl < P.nw) P.Ginv[(size_t)(w0 + wl) * NN2 + it] = s_Gi[it * WS + wl];
    }
  }
#undef PHASE
#undef NACT
#undef WLOF
#undef SR
#undef SGI
#undef SPHI
#undef SW
#undef SEL
#undef SMISC
#undef GW
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

// walkers per CTA.  The kernel is throughput bound (fp64 pipe), so a CTA's time grows linearly with its walkers on top of a
// fixed per-iteration part (serial phases, barriers), with a floor of one full round of thread tasks; the grid runs in
// `waves` of sms*ctas_per_sm CTAs.  Measured on B200, 4096 walkers (profiles/r01_sweep_walker.json): 28 walkers per CTA
// (147 CTAs, one wave) 4.74 ms per 40 projections, 31 (133 CTAs) 5.41 ms, 24 (171 CTAs, two waves) 8.93 ms.
int choose_wpc(int nw, int n_points, int sms, int ctas_per_sm, int wpc_max, int nthr) {
  int best = 1;
  double best_cost = 1e300;
  for (int wpc = wpc_max; wpc >= 1; --wpc) {  // ties: fewer, larger CTAs (table staging amortised)
    const int ctas = (nw + wpc - 1) / wpc;
    const int slots = sms * ctas_per_sm;
    const int waves = (ctas + slots - 1) / slots;
    const double pairs = 0.5 * n_points * wpc;
    const double cost = waves * (std::max(pairs, (double)nthr) / nthr + 0.6);
    if (cost < best_cost) {
      best_cost = cost;
      best = wpc;
    }
  }
  return best;
}

}  // namespace

// One CTA of 16 warps per SM: the warps of an SM then run the same phase of the projection loop at the same time, which
// keeps the instruction working set (one phase, not the whole loop) inside the instruction caches; four independent
// 4-warp CTAs per SM were measured 5x stalled on instruction fetch (profiles/r01_*).
template <bool TAU, int NMO_I, bool CART_I>
int launch_walker_one(qe_engine* h, WalkerArgs& A, cudaStream_t st, int kid) {
  const SysDev& S = h->sys;
  const int P = h->nmo_pad;
  if (S.n_up > 8) return fail(QE_ERR_UNSUPPORTED, "the fused walker kernel holds at most 8 electrons per spin (larger systems run on the general family)");
  const int NWARP = h->walker_warps > 0 ? h->walker_warps : 16;
  const int ctas_per_sm = 16 / NWARP;
  A.off_cseg = h->b_up.off_cseg;
  A.off_cbeg = h->b_up.off_cbeg;
  A.n_chunk = h->b_up.n_chunk;
  A.clk = h->phase_clk;
  const int Ne = S.n_e;
  const int n_kin = A.mode == 2 ? 0 : 6 * Ne;
  const int n_ecp = S.ecp_flag ? Ne * S.NN * S.Nv : 0;
  const size_t per_walker = (size_t)Ne * 3 + (size_t)S.n_up * S.n_up + (size_t)Ne * 5 * P + (size_t)Ne * P +
                            std::max<size_t>((size_t)NWARP * 5 * P, (size_t)std::max(1, n_kin + n_ecp) + std::max(1, n_ecp)) +
                            (S.ecp_flag ? (size_t)Ne * S.NN * (5 + S.ecp_lmax) : 1) + (size_t)Ne * 10 + 5 * P + 16;
  const size_t fixed = (size_t)h->b_up.dev.bytes + sys_bytes(S, P) + 17 * 16 + 64 + 128 + 256 + 112;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t budget = (size_t)227 * 1024 / ctas_per_sm - (ctas_per_sm > 1 ? 1024 : 0);
  if (budget < fixed + per_walker * 8) return fail(QE_ERR_UNSUPPORTED, "system too large for the fused walker kernel (shared memory)");
  // walker stride of the shared arrays: 32 (compile-time: immediate address offsets) when 32 padded slots fit, else dense
  // (instantiated for l <= 4 bases and the GFMC_n / VMC modes only: every instantiation of this kernel costs ~1 min of nvcc)
  // measured on B200 (water, 4096 walkers, profiles/r02_walker_history.md): padded 4.23 ms vs dense 4.37 ms per 40 projections
  static const bool dense_env = getenv("QE_WALKER_DENSE") && atoi(getenv("QE_WALKER_DENSE")) != 0;  // tuning switch
  const bool pad32 = !dense_env && !TAU && h->b_up.dev.lmax <= 4 && fixed + per_walker * 8 * 32 <= budget;
  int wpc_max = pad32 ? 32 : (int)std::min<size_t>(32, (budget - fixed) / (per_walker * 8));
  wpc_max = std::max(1, std::min(wpc_max, NWARP * 32 / S.n_up));  // Sherman-Morrison: one (walker, row) task per thread
  A.wpc = h->wpc_override > 0 ? std::min(h->wpc_override, wpc_max)
                              : choose_wpc(A.nw, std::max(1, n_kin + n_ecp), sms, ctas_per_sm, wpc_max, NWARP * 32);
  const size_t smem = fixed + per_walker * 8 * (pad32 ? 32 : A.wpc);
  {
    LaunchScope ls_(h, kid, st);
#define CALL5(NMO, CART, LMAX, TAU, WS, MX)                                                                                         \
  do {                                                                                                                              \
    CUDA_TRY(cudaFuncSetAttribute(k_walker<NMO, CART, LMAX, TAU, WS, MX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_walker<NMO, CART, LMAX, TAU, WS, MX><<<nblk(A.nw, A.wpc), NWARP * 32, smem, st>>>(h->b_up.dev, S, A);                          \
  } while (0)
#define CALL4(NMO, CART, LMAX, TAU, WS) CALL5(NMO, CART, LMAX, TAU, WS, false)
#define CALL3(NMO, CART, LMAX, TAU)                                        \
  do {                                                                     \
    if constexpr (!(TAU) && (LMAX) == 4) {                                 \
      if (pad32 && h->mixed) CALL5(NMO, CART, LMAX, TAU, 32, true);        \
      else if (pad32) CALL4(NMO, CART, LMAX, TAU, 32);                     \
      else CALL4(NMO, CART, LMAX, TAU, 0);                                 \
    } else {                                                               \
      CALL4(NMO, CART, LMAX, TAU, 0);                                      \
    }                                                                      \
  } while (0)
#define CALL2(NMO, CART, LMAX) CALL3(NMO, CART, LMAX, TAU)
#define CALL(NMO, CART)                                  \
  do {                                                   \
    if (h->b_up.dev.lmax <= 4) CALL2(NMO, CART, 4);      \
    else CALL2(NMO, CART, 6);                            \
  } while (0)
    CALL(NMO_I, CART_I);
#undef CALL
#undef CALL2
#undef CALL3
#undef CALL4
#undef CALL5
  }
  CHECK_LAUNCH();
  return QE_OK;
}

// Front end: the instantiations of the kernel family live in one translation unit per (TAU, orbital padding, Cartesian) --
// jqmc_b200/csrc/qe_walker_i_*.cu -- so that every file goes through a single-threaded (deterministic) ptxas and the files
// compile in parallel.  (`nvcc --split-compile` was measured to produce DIFFERENT machine code from identical sources on
// every run, and builds of this kernel differ by up to 10 % in speed: profiles/r02_walker_history.md.)
#define QE_WALKER_EXTERN(TAU_, NMO_, CART_) \
  extern template int launch_walker_one<TAU_, NMO_, CART_>(qe_engine*, WalkerArgs&, cudaStream_t, int);
QE_WALKER_EXTERN(false, 4, false)
QE_WALKER_EXTERN(false, 4, true)
QE_WALKER_EXTERN(false, 8, false)
QE_WALKER_EXTERN(false, 8, true)
QE_WALKER_EXTERN(false, 16, false)
QE_WALKER_EXTERN(false, 16, true)
QE_WALKER_EXTERN(true, 4, false)
QE_WALKER_EXTERN(true, 4, true)
QE_WALKER_EXTERN(true, 8, false)
QE_WALKER_EXTERN(true, 8, true)
QE_WALKER_EXTERN(true, 16, false)
QE_WALKER_EXTERN(true, 16, true)
#undef QE_WALKER_EXTERN

template <bool TAU>
int launch_walker(qe_engine* h, WalkerArgs& A, cudaStream_t st, int kid) {
  const bool cart = h->b_up.dev.cart != 0;
#ifdef QE_DEV_MINIMAL
  if (cart || h->nmo_pad != 4) return fail(QE_ERR_UNSUPPORTED, "QE_DEV_MINIMAL build");
  return launch_walker_one<TAU, 4, false>(h, A, st, kid);
#else
  switch (h->nmo_pad) {
    case 4: return cart ? launch_walker_one<TAU, 4, true>(h, A, st, kid) : launch_walker_one<TAU, 4, false>(h, A, st, kid);
    case 8: return cart ? launch_walker_one<TAU, 8, true>(h, A, st, kid) : launch_walker_one<TAU, 8, false>(h, A, st, kid);
    default: return cart ? launch_walker_one<TAU, 16, true>(h, A, st, kid) : launch_walker_one<TAU, 16, false>(h, A, st, kid);
  }
#endif
}
