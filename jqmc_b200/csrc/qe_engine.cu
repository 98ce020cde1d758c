// B200-native walker engine for jQMC: host-side table construction, kernels, and the C ABI of
// include/jqmc_b200.h.  Written for sm_100a (fp64 path: the Blackwell tcgen05 tensor cores have no fp64
// mode, so the contractions run as DFMA fused into the AO evaluation).
//
// Thread mapping used by every kernel: LANES ARE WALKERS.  A warp holds 32 consecutive walkers and one
// task (an electron, a quadrature point, a chunk of the AO basis); all lanes therefore execute the same
// shell/primitive sequence, table loads are warp-broadcast and walker-state loads/stores are coalesced
// through [item][walker] workspace layouts.
#include "qe_common.cuh"

static thread_local std::string g_err;
int qe_fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

// =================================================================================================
// host: basis compression (per-AO primitive lists -> shells grouped by (nucleus, l))
// =================================================================================================
static double dfact(int n) {
  double r = 1.0;
  for (int i = 2; i <= n; ++i) r *= i;
  return r;
}

static int cart_index(int l, int nx, int ny, int nz) {
  // position of x^nx y^ny z^nz in itertools.combinations_with_replacement("xyz", l)
  int k = 0;
  for (int ax = l; ax >= 0; --ax)
    for (int ay = l - ax; ay >= 0; --ay) {
      int az = l - ax - ay;
      if (ax == nx && ay == ny && az == nz) return k;
      ++k;
    }
  return -1;
}

struct TmpShell {
  int nuc, l;
  std::vector<double> Z, c;  // c: reference coefficients (unnormalised) of the first AO of the shell
  std::vector<short> slot;
  int first_ao;
};

struct Blob {
  std::vector<char> bytes;
  template <class T>
  int put(const std::vector<T>& v) {
    const size_t off = (bytes.size() + 15) & ~size_t(15);
    bytes.resize(off + std::max<size_t>(v.size(), 1) * sizeof(T), 0);
    if (!v.empty()) std::memcpy(bytes.data() + off, v.data(), v.size() * sizeof(T));
    return (int)off;
  }
  void finish() { bytes.resize((bytes.size() + 15) & ~size_t(15), 0); }
};

// split the shells (in table order) into at most n_target contiguous chunks of roughly equal cost; a chunk is a list of
// segments {nucleus, l, shell_begin, shell_end} (cut at group boundaries).  cseg: segments, cbeg[n_chunk+1]: first segment
static void make_chunks(const std::vector<int4>& grp, const std::vector<double>& sh_cost, const std::vector<double>& grp_overhead,
                        int n_target, std::vector<int4>& cseg, std::vector<int>& cbeg) {
  const int n_sh = (int)sh_cost.size();
  std::vector<int> sh_grp(n_sh, 0);
  for (int g = 0; g < (int)grp.size(); ++g)
    for (int s = grp[g].z; s < grp[g].w; ++s) sh_grp[s] = g;
  n_target = std::max(1, std::min(n_target, n_sh));
  auto count_chunks = [&](double cap, std::vector<int>* cuts) {
    int used = 1;
    double acc = 0;
    int cur_g = -1;
    if (cuts) cuts->assign(1, 0);
    for (int s = 0; s < n_sh; ++s) {
      double c = sh_cost[s] + (sh_grp[s] != cur_g ? grp_overhead[sh_grp[s]] : 0.0);
      if (acc > 0 && acc + c > cap) {
        ++used;
        acc = 0;
        c = sh_cost[s] + grp_overhead[sh_grp[s]];  // a new chunk re-evaluates the angular part of its first group
        if (cuts) cuts->push_back(s);
      }
      acc += c;
      cur_g = sh_grp[s];
    }
    if (cuts) cuts->push_back(n_sh);
    return used;
  };
  double lo = 0, hi = 0;
  for (int s = 0; s < n_sh; ++s) {
    lo = std::max(lo, sh_cost[s] + grp_overhead[sh_grp[s]]);
    hi += sh_cost[s] + grp_overhead[sh_grp[s]];
  }
  for (int it = 0; it < 60; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (count_chunks(mid, nullptr) <= n_target) hi = mid; else lo = mid;
  }
  std::vector<int> cuts;
  count_chunks(hi * (1 + 1e-12), &cuts);
  cseg.clear();
  cbeg.clear();
  for (size_t c = 0; c + 1 < cuts.size(); ++c) {
    cbeg.push_back((int)cseg.size());
    int s = cuts[c];
    while (s < cuts[c + 1]) {
      const int g = sh_grp[s];
      const int e = std::min(cuts[c + 1], grp[g].w);
      cseg.push_back(make_int4(grp[g].x, grp[g].y, s, e));
      s = e;
    }
  }
  cbeg.push_back((int)cseg.size());
}

// d2: second coefficient set sharing the AO tables (down-spin MOs), or nullptr
static int build_basis(const qe_basis_desc& d, const qe_basis_desc* d2, int n_atom, const double* positions, int nmo_pad, int n_chunk_target,
                       DevPool& pool, HostBasis& hb, bool with_C, std::vector<int>& row_ao, std::vector<double>& row_scale) {
  if (d.n_ao <= 0) return fail(QE_ERR_INVALID, "basis: n_ao must be positive");
  const bool cart = d.cartesian != 0;
  std::vector<std::vector<int>> prims(d.n_ao);
  for (int p = 0; p < d.n_prim; ++p) {
    int a = d.orbital_indices[p];
    if (a < 0 || a >= d.n_ao) return fail(QE_ERR_INVALID, "basis: orbital_indices out of range");
    prims[a].push_back(p);
  }
  std::vector<TmpShell> shells;
  std::vector<double> ao_scale(d.n_ao, 1.0);
  for (int a = 0; a < d.n_ao; ++a) {
    const int l = d.angular_momentums[a];
    const int nuc = d.nucleus_index[a];
    if (l < 0 || l > QE_LMAX) return fail(QE_ERR_UNSUPPORTED, "basis: angular momentum > 6 is not supported");
    if (nuc < 0 || nuc >= n_atom) return fail(QE_ERR_INVALID, "basis: nucleus_index out of range");
    int k;
    double extra = 1.0;
    if (cart) {
      const int nx = d.polynominal_order_x[a], ny = d.polynominal_order_y[a], nz = d.polynominal_order_z[a];
      if (nx + ny + nz != l) return fail(QE_ERR_INVALID, "basis: nx+ny+nz != l");
      k = cart_index(l, nx, ny, nz);
      extra = std::sqrt(dfact(nx) * dfact(ny) * dfact(nz) / (dfact(2 * nx) * dfact(2 * ny) * dfact(2 * nz)));
    } else {
      const int m = d.magnetic_quantum_numbers[a];
      if (m < -l || m > l) return fail(QE_ERR_INVALID, "basis: |m| > l");
      k = m + l;
    }
    const auto& pl = prims[a];
    // try to join the last shell
    bool joined = false;
    if (!shells.empty()) {
      TmpShell& s = shells.back();
      if (s.nuc == nuc && s.l == l && s.Z.size() == pl.size() && s.slot[k] < 0 && !pl.empty()) {
        size_t piv = 0;
        for (size_t i = 0; i < pl.size(); ++i)
          if (std::fabs(s.c[i]) > std::fabs(s.c[piv])) piv = i;
        const double ratio = d.coefficients[pl[piv]] / s.c[piv];
        bool ok = std::isfinite(ratio);
        for (size_t i = 0; ok && i < pl.size(); ++i) {
          if (d.exponents[pl[i]] != s.Z[i]) ok = false;
          const double ci = d.coefficients[pl[i]];
          if (std::fabs(ci - ratio * s.c[i]) > 4e-15 * std::fabs(ci)) ok = false;
        }
        if (ok) {
          s.slot[k] = (short)a;
          ao_scale[a] = ratio * extra;
          joined = true;
        }
      }
    }
    if (!joined) {
      TmpShell s;
      s.nuc = nuc;
      s.l = l;
      s.first_ao = a;
      s.slot.assign(MAXF, -1);
      for (int p : pl) {
        s.Z.push_back(d.exponents[p]);
        s.c.push_back(d.coefficients[p]);
      }
      s.slot[k] = (short)a;
      ao_scale[a] = extra;
      shells.push_back(std::move(s));
    }
  }
  std::stable_sort(shells.begin(), shells.end(), [](const TmpShell& x, const TmpShell& y) {
    return x.nuc != y.nuc ? x.nuc < y.nuc : x.l < y.l;
  });
  std::vector<int4> grp, sh;
  std::vector<double2> pr;
  row_ao.clear();
  row_scale.clear();
  std::vector<double> sh_cost, grp_overhead;
  for (size_t s = 0; s < shells.size(); ++s) {
    const TmpShell& t = shells[s];
    const int l = t.l;
    const int nf = cart ? (l + 1) * (l + 2) / 2 : 2 * l + 1;
    if (s == 0 || t.nuc != shells[s - 1].nuc || t.l != shells[s - 1].l) {
      if (!grp.empty()) grp.back().w = (int)s;
      grp.push_back(make_int4(t.nuc, t.l, (int)s, (int)s));
      grp_overhead.push_back(10.0 + 6.0 * l * l);
    }
    const int pb = (int)pr.size();
    for (size_t i = 0; i < t.Z.size(); ++i) {
      const double Z = t.Z[i];
      double N;
      if (cart)  // jqmc/atomic_orbital.py:2243-2244 (Z-dependent part; factorial part lives in the per-AO scale)
        N = std::sqrt(std::pow(2.0 * Z / M_PI, 1.5) * std::pow(8.0 * Z, (double)l));
      else  // jqmc/atomic_orbital.py:2316-2323, times sqrt((2l+1)/4pi) (:2349)
        N = std::sqrt(std::pow(2.0, 2 * l + 3) * dfact(l + 1) * std::pow(2.0 * Z, l + 1.5) / (dfact(2 * l + 2) * std::sqrt(M_PI))) *
            std::sqrt((2 * l + 1) / (4.0 * M_PI));
      pr.push_back(make_double2(-Z, t.c[i] * N));
    }
    sh.push_back(make_int4(pb, (int)pr.size(), (int)row_ao.size(), 0));
    for (int k = 0; k < nf; ++k) {
      const int a = t.slot[k];
      row_ao.push_back(a);
      row_scale.push_back(a >= 0 ? ao_scale[a] : 0.0);
    }
    sh_cost.push_back(22.0 * t.Z.size() + (nmo_pad + 3.0) * nf);
  }
  if (!grp.empty()) grp.back().w = (int)shells.size();
  // pair flags: consecutive uncontracted shells of a group are swept two at a time (eval_seg_val_n)
  for (const int4& g : grp)
    for (int si = g.z; si + 1 < g.w;) {
      if (sh[si].y - sh[si].x == 1 && sh[si + 1].y - sh[si + 1].x == 1) {
        sh[si].w = 1;
        si += 2;
      } else {
        ++si;
      }
    }
  const int n_row = (int)row_ao.size();

  BasisDev& B = hb.dev;
  B = BasisDev{};
  B.n_ao = d.n_ao;
  B.n_mo = d.n_mo;
  B.n_orb = d.n_mo > 0 ? d.n_mo : d.n_ao;
  B.n_grp = (int)grp.size();
  B.n_shell = (int)shells.size();
  B.n_prim = (int)pr.size();
  B.n_row = n_row;
  B.cart = cart ? 1 : 0;
  B.nmo_pad = nmo_pad;
  B.lmax = 0;
  for (const TmpShell& t : shells) B.lmax = std::max(B.lmax, t.l);
  auto c_table = [&](const qe_basis_desc& q) {
    std::vector<double> C((size_t)n_row * nmo_pad, 0.0);
    for (int r = 0; r < n_row; ++r) {
      const int a = row_ao[r];
      if (a < 0) continue;
      for (int mo = 0; mo < q.n_mo; ++mo) C[(size_t)r * nmo_pad + mo] = q.mo_coefficients[(size_t)mo * q.n_ao + a] * ao_scale[a];
    }
    return C;
  };
  std::vector<int4> cseg;
  std::vector<int> cbeg;
  make_chunks(grp, sh_cost, grp_overhead, n_chunk_target, cseg, cbeg);
  hb.n_chunk = (int)cbeg.size() - 1;
  {  // chunks sorted by the angular momentum of their first segment; sorted chunk i -> warp (i / 4) + 4 (i % 4) when 16 warps
    std::vector<int> order(hb.n_chunk);
    for (int c = 0; c < hb.n_chunk; ++c) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cseg[cbeg[a]].y < cseg[cbeg[b]].y; });
    const int nw_ = hb.n_chunk, per = (nw_ + 3) / 4;
    std::vector<int> slot;  // warp ids in the order sub-partition 0, 1, 2, 3
    for (int sp = 0; sp < 4; ++sp)
      for (int w = sp; w < nw_; w += 4) slot.push_back(w);
    (void)per;
    for (int i = 0; i < nw_ && i < 32; ++i) hb.chunk_warp[slot[i]] = order[i];
  }
  Blob blob;
  B.off_seg = blob.put(grp);
  B.off_sh = blob.put(sh);
  B.off_pr = blob.put(pr);
  {
    std::vector<double2> pr2(pr.size());
    for (size_t i = 0; i < pr.size(); ++i) pr2[i] = make_double2(pr[i].x * 46.16624130844683, pr[i].y);  // -Z * 32/ln2
    B.off_pr2 = blob.put(pr2);
    B.off_et = blob.put(std::vector<double>(QE_EXP2_TABLE, QE_EXP2_TABLE + 32));
    std::vector<float2> prf(pr.size());
    for (size_t i = 0; i < pr.size(); ++i) prf[i] = make_float2((float)(pr[i].x * 1.4426950408889634), (float)pr[i].y);  // -Z / ln2
    B.off_prf = blob.put(prf);
  }
  B.off_C = B.off_C2 = 0;
  if (d.n_mo > 0 && with_C) {
    const std::vector<double> Cu = c_table(d);
    B.off_C = blob.put(Cu);
    B.off_C2 = B.off_C;
    if (d2) {
      const std::vector<double> Cd = c_table(*d2);
      if (Cd != Cu) B.off_C2 = blob.put(Cd);  // restricted (same orbitals for both spins): one table
    }
  }
  B.off_rowao = blob.put(row_ao);
  B.off_rowscale = blob.put(row_scale);
  B.off_Rn = blob.put(std::vector<double>(positions, positions + 3 * n_atom));
  hb.off_cseg = blob.put(cseg);
  hb.off_cbeg = blob.put(cbeg);
  blob.finish();
  B.bytes = (int)blob.bytes.size();
  const char* dev = nullptr;
  cudaError_t e = pool.upload(blob.bytes, &dev);
  if (e != cudaSuccess) return fail(QE_ERR_CUDA, std::string("basis upload: ") + cudaGetErrorString(e));
  B.g = dev;
  hb.present = true;
  return QE_OK;
}

static bool same_ao_tables(const qe_basis_desc& x, const qe_basis_desc& y) {
  if (x.cartesian != y.cartesian || x.n_ao != y.n_ao || x.n_prim != y.n_prim) return false;
  for (int a = 0; a < x.n_ao; ++a) {
    if (x.nucleus_index[a] != y.nucleus_index[a] || x.angular_momentums[a] != y.angular_momentums[a]) return false;
    if (x.cartesian) {
      if (x.polynominal_order_x[a] != y.polynominal_order_x[a] || x.polynominal_order_y[a] != y.polynominal_order_y[a] ||
          x.polynominal_order_z[a] != y.polynominal_order_z[a])
        return false;
    } else if (x.magnetic_quantum_numbers[a] != y.magnetic_quantum_numbers[a])
      return false;
  }
  for (int p = 0; p < x.n_prim; ++p)
    if (x.orbital_indices[p] != y.orbital_indices[p] || x.exponents[p] != y.exponents[p] || x.coefficients[p] != y.coefficients[p])
      return false;
  return true;
}

// =================================================================================================
// K1/K2 parity entry: orbital values / VGL at arbitrary points
// =================================================================================================
template <bool CART>
__global__ void k_eval_ao(BasisDev B, int n_pts, const double* __restrict__ r, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pts) return;
  SinkStoreAO sink{out + t, (const int*)(B.g + B.off_rowao), (const double*)(B.g + B.off_rowscale), (long long)B.n_ao * n_pts, n_pts};
  eval_vgl<CART, QE_LMAX>(B.g, B, B.off_seg, r[3 * t], r[3 * t + 1], r[3 * t + 2], 0, B.n_grp, sink);
}
template <int NMO, bool CART>
__global__ void k_eval_mo(BasisDev B, int dn, int n_pts, const double* __restrict__ r, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pts) return;
  SinkMO5<NMO> sink;
  sink.init(B.g + (dn ? B.off_C2 : B.off_C));
  eval_vgl<CART, QE_LMAX>(B.g, B, B.off_seg, r[3 * t], r[3 * t + 1], r[3 * t + 2], 0, B.n_grp, sink);
  for (int q = 0; q < 5; ++q)
#pragma unroll
    for (int mo = 0; mo < NMO; ++mo)
      if (mo < B.n_mo) out[((long long)q * B.n_mo + mo) * n_pts + t] = sink.acc[q][mo];
}

// =================================================================================================
// E1: orbital values (NQ=1) or value/grad/lap (NQ=5) of every MO at every electron of every walker.
// thread = (chunk, electron, walker);  out[chunk][e][q][mo][w]
// =================================================================================================
template <int NMO, bool CART, int NQ>
__global__ void __launch_bounds__(128)
k_orb_electrons(BasisDev B, SysDev S, int nw, const double* __restrict__ r_up, const double* __restrict__ r_dn, int off_cseg,
                int off_cbeg, int n_chunk, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_chunk * S.n_e * nw;
  if (t >= total) return;
  const int w = (int)(t % nw);
  const int e = (int)((t / nw) % S.n_e);
  const int c = (int)(t / ((long long)nw * S.n_e));
  PosGlobal pos{r_up, r_dn, S.n_up, S.n_dn, w};
  double x, y, z;
  pos.get(e, x, y, z);
  const int* cbeg = (const int*)(B.g + off_cbeg);
  const int gb = cbeg[c], ge = cbeg[c + 1];
  const char* Ctab = B.g + (e < S.n_up ? B.off_C : B.off_C2);
  double* o = out + (((size_t)c * S.n_e + e) * NQ * NMO) * nw + w;
  if (NQ == 1) {
    SinkMO<NMO> sink;
    sink.init(Ctab);
    eval_val<CART, QE_LMAX>(B.g, B, off_cseg, x, y, z, gb, ge, sink);
#pragma unroll
    for (int mo = 0; mo < NMO; ++mo) o[(size_t)mo * nw] = sink.acc[mo];
  } else {
    SinkMO5<NMO> sink;
    sink.init(Ctab);
    eval_vgl<CART, QE_LMAX>(B.g, B, off_cseg, x, y, z, gb, ge, sink);
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
      for (int mo = 0; mo < NMO; ++mo) o[((size_t)q * NMO + mo) * nw] = sink.acc[q][mo];
  }
}

// =================================================================================================
// G2: per-walker geminal matrix, inverse (Gauss-Jordan, partial pivoting), ln|det|, Jastrow value.
// thread = walker.  phi[chunk][e][1][mo][w]
// =================================================================================================
// NT = N_up when it is known at compile time (1..8: matrices live in registers, loops fully unrolled), 0 = run-time N <= 16
template <int NMO, int NT>
__global__ void k_geminal(SysDev S, int nw, int n_chunk, const double* __restrict__ phi, const double* __restrict__ r_up,
                          const double* __restrict__ r_dn, double* __restrict__ G_out, double* __restrict__ Ginv_out,
                          double* __restrict__ lnpsi_out, double* __restrict__ sign_out) {
  // block = 32 walkers (x) x 8 (y): the chunk partial sums of the orbital values are reduced cooperatively into shared
  // memory (fixed chunk order), then one thread per walker does the small dense algebra
  extern __shared__ double s_phi_g[];  // [(e*NMO + mo)][32]
  const int lane = threadIdx.x;
  const int wq = blockIdx.x * 32 + lane;
  const int w = wq < nw ? wq : nw - 1;
  constexpr int NMAX = NT ? NT : 16;
  const int N = NT ? NT : S.n_up, Nd = S.n_dn;
  for (int it = threadIdx.y; it < S.n_e * NMO; it += blockDim.y) {
    double s = 0;
    for (int c = 0; c < n_chunk; ++c) s += phi[((size_t)c * S.n_e * NMO + it) * nw + w];
    s_phi_g[it * 32 + lane] = s;
  }
  __syncthreads();
  if (threadIdx.y != 0 || wq >= nw) return;
#define PHIS(e, mo) s_phi_g[((e) * NMO + (mo)) * 32 + lane]
  double G[NMAX * NMAX], I[NMAX * NMAX];
  // G[i][j] = sum_{a,b} PhiU[a][i] lam_p[a][b] PhiD[b][j] ; unpaired columns: sum_a PhiU[a][i] lam_u[a][k]
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    if (i >= N) break;
    double t[NMO];
#pragma unroll
    for (int b = 0; b < NMO; ++b) {
      double s = 0;
#pragma unroll
      for (int a = 0; a < NMO; ++a) s = fma(PHIS(i, a), S.lam_p[a * NMO + b], s);
      t[b] = s;
    }
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
      if (j >= N) break;
      double s = 0;
      if (j < Nd) {
#pragma unroll
        for (int b = 0; b < NMO; ++b) s = fma(t[b], PHIS(N + j, b), s);
      } else {
        const int k = j - Nd;
        for (int a = 0; a < NMO; ++a) s = fma(PHIS(i, a), S.lam_u[a * S.n_unp + k], s);
      }
      G[i * N + j] = s;
    }
  }
#undef PHIS
  if (G_out) {
#pragma unroll
    for (int i = 0; i < NMAX * NMAX; ++i)
      if (i < N * N) G_out[(size_t)w * N * N + i] = G[i];
  }
  // Gauss-Jordan with partial pivoting on a copy.  The reference inverts by a thresholded-SVD pseudo-inverse (rcond 1e-20,
  // jqmc/jqmc_mcmc.py:4258-4260): for a non-singular G both are the inverse; a pivot that vanishes relative to the largest
  // element of G marks a null direction whose row of the result is set to zero instead of dividing by it, so a singular G
  // (two same-spin electrons on top of each other) gives a finite generalised inverse and ln|det| = -inf, never inf / NaN.
  double A[NMAX * NMAX];
  double gmax = 0.0;
#pragma unroll
  for (int i = 0; i < NMAX * NMAX; ++i)
    if (i < N * N) {
      A[i] = G[i];
      gmax = fmax(gmax, fabs(G[i]));
      I[i] = (i / N == i % N) ? 1.0 : 0.0;
    }
  const double tiny = 1.0e-20 * gmax;
  double lndet = 0.0, sgn = 1.0;
#pragma unroll
  for (int c = 0; c < N; ++c) {
    int piv = c;
    double best = fabs(A[c * N + c]);
#pragma unroll
    for (int r = c + 1; r < N; ++r)
      if (fabs(A[r * N + c]) > best) {
        best = fabs(A[r * N + c]);
        piv = r;
      }
    if (piv != c) {
      if (NT) {  // compile-time row indices only (registers): conditional swap against every candidate row
#pragma unroll
        for (int r = c + 1; r < N; ++r)
          if (r == piv) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
              double tmp = A[c * N + j];
              A[c * N + j] = A[r * N + j];
              A[r * N + j] = tmp;
              tmp = I[c * N + j];
              I[c * N + j] = I[r * N + j];
              I[r * N + j] = tmp;
            }
          }
      } else {
        for (int j = 0; j < N; ++j) {
          double tmp = A[c * N + j];
          A[c * N + j] = A[piv * N + j];
          A[piv * N + j] = tmp;
          tmp = I[c * N + j];
          I[c * N + j] = I[piv * N + j];
          I[piv * N + j] = tmp;
        }
      }
      sgn = -sgn;
    }
    const double d = A[c * N + c];
    lndet += log(fabs(d));
    if (d < 0) sgn = -sgn;
    const double inv = fabs(d) > tiny ? 1.0 / d : 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      A[c * N + j] *= inv;
      I[c * N + j] *= inv;
    }
#pragma unroll
    for (int r = 0; r < N; ++r) {
      if (r == c) continue;
      const double f = A[r * N + c];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        A[r * N + j] = fma(-f, A[c * N + j], A[r * N + j]);
        I[r * N + j] = fma(-f, I[c * N + j], I[r * N + j]);
      }
    }
  }
  if (Ginv_out) {
#pragma unroll
    for (int i = 0; i < NMAX * NMAX; ++i)
      if (i < N * N) Ginv_out[(size_t)w * N * N + i] = I[i];
  }
  if (lnpsi_out) {
    // Jastrow value J1 + J2 (jqmc/jastrow_factor.py:2127-2178)
    PosGlobal pos{r_up, r_dn, S.n_up, S.n_dn, w};
    double J = 0.0;
    for (int e = 0; e < S.n_e; ++e) {
      double x, y, z;
      pos.get(e, x, y, z);
      if (S.j1_type)
        for (int a = 0; a < S.n_atom; ++a) {
          const double dx = x - S.Rn[3 * a], dy = y - S.Rn[3 * a + 1], dz = z - S.Rn[3 * a + 2];
          J += j1_f(S.j1_type, S.j1_a, S.j1_A[a], S.j1_c[a], sqrt(dx * dx + dy * dy + dz * dz));
        }
      if (S.j2_type)
        for (int j = e + 1; j < S.n_e; ++j) {
          double x2, y2, z2;
          pos.get(j, x2, y2, z2);
          J += j2_f(S.j2_type, S.j2_a, sqrt((x - x2) * (x - x2) + (y - y2) * (y - y2) + (z - z2) * (z - z2)));
        }
    }
    lnpsi_out[w] = J + lndet;
    if (sign_out) sign_out[w] = sgn;
  }
}

// =================================================================================================
// E2: per-(walker, electron) algebra: sum chunk partials, ratio weight vector W[:,e] = d(det ratio)/d(phi),
//     grad/lap ln|det| (jqmc/determinant.py:2140-2250), Jastrow J1/J2 grad/lap
//     (jqmc/jastrow_factor.py:960-1034, 3434-3558), per-electron kinetic energy
//     (jqmc/wavefunction.py:1198-1207), bare Coulomb / ECP-local pieces (jqmc/coulomb_potential.py:2252-2281,
//     1144-1246).   thread = (electron, walker)
//     phi[chunk][e][NQ][mo][w];  W[e][mo][w];  Te/Vb/Vl[e][w]
// =================================================================================================
template <int NMO, int NQ>
__global__ void k_electron_algebra(SysDev S, int nw, int n_chunk, const double* __restrict__ phi,
                                   const double* __restrict__ r_up, const double* __restrict__ r_dn,
                                   const double* __restrict__ Ginv, double* __restrict__ W, double* __restrict__ Te,
                                   double* __restrict__ Vb, double* __restrict__ Vl) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)S.n_e * nw) return;
  const int w = (int)(t % nw);
  const int e = (int)(t / nw);
  const int N = S.n_up, Nd = S.n_dn;
  const bool up = e < N;
  const double* Gi = Ginv + (size_t)w * N * N;
  auto PHI = [&](int el, int q, int mo) {
    double s = 0;
    for (int c = 0; c < n_chunk; ++c) s += phi[((((size_t)c * S.n_e + el) * NQ + q) * NMO + mo) * nw + w];
    return s;
  };
  // weight vector: det ratio for moving electron e to r' is  sum_mo phi_mo(r') * Wv[mo]
  double Wv[NMO];
  if (up) {
    // Wv[a] = sum_j (lam_p PhiD)[a][j] Ginv[j][e] + sum_k lam_u[a][k] Ginv[Nd+k][e]
    double y[NMO];
    for (int b = 0; b < NMO; ++b) {
      double s = 0;
      for (int j = 0; j < Nd; ++j) s = fma(PHI(N + j, 0, b), Gi[j * N + e], s);
      y[b] = s;
    }
    for (int a = 0; a < NMO; ++a) {
      double s = 0;
      for (int b = 0; b < NMO; ++b) s = fma(S.lam_p[a * NMO + b], y[b], s);
      for (int k = 0; k < S.n_unp; ++k) s = fma(S.lam_u[a * S.n_unp + k], Gi[(Nd + k) * N + e], s);
      Wv[a] = s;
    }
  } else {
    // Wv[b] = sum_i Ginv[j][i] (PhiU^T lam_p)[i][b],  j = e - N
    const int j = e - N;
    double y[NMO];
    for (int a = 0; a < NMO; ++a) {
      double s = 0;
      for (int i = 0; i < N; ++i) s = fma(PHI(i, 0, a), Gi[j * N + i], s);
      y[a] = s;
    }
    for (int b = 0; b < NMO; ++b) {
      double s = 0;
      for (int a = 0; a < NMO; ++a) s = fma(y[a], S.lam_p[a * NMO + b], s);
      Wv[b] = s;
    }
  }
  for (int mo = 0; mo < NMO; ++mo) W[((size_t)e * NMO + mo) * nw + w] = Wv[mo];
  if (NQ == 1) return;

  double gD[3] = {0, 0, 0}, lD = 0;
  for (int mo = 0; mo < NMO; ++mo) {
    gD[0] = fma(PHI(e, 1, mo), Wv[mo], gD[0]);
    gD[1] = fma(PHI(e, 2, mo), Wv[mo], gD[1]);
    gD[2] = fma(PHI(e, 3, mo), Wv[mo], gD[2]);
    lD = fma(PHI(e, 4, mo), Wv[mo], lD);
  }
  lD -= gD[0] * gD[0] + gD[1] * gD[1] + gD[2] * gD[2];

  PosGlobal pos{r_up, r_dn, N, Nd, w};
  double x, y_, z;
  pos.get(e, x, y_, z);
  double gJ[3] = {0, 0, 0}, lJ = 0;
  double v_bare = 0.0, v_loc = 0.0;
  const double eps = 1.0e-12;
  for (int a = 0; a < S.n_atom; ++a) {
    const double dx = x - S.Rn[3 * a], dy = y_ - S.Rn[3 * a + 1], dz = z - S.Rn[3 * a + 2];
    const double d = sqrt(dx * dx + dy * dy + dz * dz);
    v_bare -= S.Zeff[a] / d;
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
      const double s = fp / rs;
      gJ[0] = fma(s, dx, gJ[0]);
      gJ[1] = fma(s, dy, gJ[1]);
      gJ[2] = fma(s, dz, gJ[2]);
    }
    if (S.ecp_flag) {
      const int lloc = S.ecp_lmax_atom[a];
      double s = 0.0;
      for (int k = S.ecp_off[a]; k < S.ecp_off[a + 1]; ++k)
        if (S.ecp_l[k] == lloc) s += S.ecp_c[k] * ipow(d, S.ecp_p[k]) * qexp(-S.ecp_z[k] * d * d);
      v_loc += s / (d * d);
    }
  }
  for (int j = 0; j < S.n_e; ++j) {
    if (j == e) continue;
    double x2, y2, z2;
    pos.get(j, x2, y2, z2);
    const double dx = x - x2, dy = y_ - y2, dz = z - z2;
    const double d = sqrt(dx * dx + dy * dy + dz * dz);
    if (j > e) v_bare += 1.0 / d;
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
      const double s = fp / rs;
      gJ[0] = fma(s, dx, gJ[0]);
      gJ[1] = fma(s, dy, gJ[1]);
      gJ[2] = fma(s, dz, gJ[2]);
    }
  }
  const double gx = gJ[0] + gD[0], gy = gJ[1] + gD[1], gz = gJ[2] + gD[2];
  Te[(size_t)e * nw + w] = -0.5 * (lJ + lD + gx * gx + gy * gy + gz * gz);
  Vb[(size_t)e * nw + w] = v_bare;
  Vl[(size_t)e * nw + w] = v_loc;
}

// =================================================================================================
// E3: non-local ECP on the rotated quadrature (jqmc/coulomb_potential.py:1477-1712).
// thread = (point=(electron, nn, k), walker).   Vnl[pt][w]
// =================================================================================================
template <int NMO, bool CART>
__global__ void __launch_bounds__(128)
k_ecp_mesh(BasisDev B, SysDev S, int nw, const double* __restrict__ r_up, const double* __restrict__ r_dn,
           const double* __restrict__ RT, const double* __restrict__ W, int det_only, double* __restrict__ Vnl,
           double* __restrict__ mesh_xyz) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int npt = S.n_e * S.NN * S.Nv;
  if (t >= (long long)npt * nw) return;
  const int w = (int)(t % nw);
  const int pt = (int)(t / nw);
  const int k = pt % S.Nv;
  const int nn = (pt / S.Nv) % S.NN;
  const int e = pt / (S.Nv * S.NN);
  PosGlobal pos{r_up, r_dn, S.n_up, S.n_dn, w};
  double x, y, z;
  pos.get(e, x, y, z);
  double d;
  const int a = nearest_atom(S.Rn, S.n_atom, x, y, z, nn, &d);
  const double relx = S.Rn[3 * a] - x, rely = S.Rn[3 * a + 1] - y, relz = S.Rn[3 * a + 2] - z;
  d = sqrt(relx * relx + rely * rely + relz * relz);
  // rotated grid point g = grid[k] @ RT
  const double* rt = RT + (size_t)w * 9;
  const double q0 = S.quad_g[3 * k], q1 = S.quad_g[3 * k + 1], q2 = S.quad_g[3 * k + 2];
  const double gx = q0 * rt[0] + q1 * rt[3] + q2 * rt[6];
  const double gy = q0 * rt[1] + q1 * rt[4] + q2 * rt[7];
  const double gz = q0 * rt[2] + q1 * rt[5] + q2 * rt[8];
  const double px = x + relx + d * gx, py = y + rely + d * gy, pz = z + relz + d * gz;
  if (mesh_xyz) {
    double* m = mesh_xyz + ((size_t)pt * 3) * nw + w;
    m[0] = px;
    m[(size_t)nw] = py;
    m[(size_t)2 * nw] = pz;
  }
  const double gn = sqrt(gx * gx + gy * gy + gz * gz);
  const double cos_t = (-relx / d) * (gx / gn) + (-rely / d) * (gy / gn) + (-relz / d) * (gz / gn);
  // radial channels V_l(d) = sum_terms c d^(p-2) exp(-z d^2), non-local terms only
  const int lloc = S.ecp_lmax_atom[a];
  double ang = 0.0;
  for (int l = 0; l < lloc; ++l) {
    double vl = 0.0;
    for (int kk = S.ecp_off[a]; kk < S.ecp_off[a + 1]; ++kk)
      if (S.ecp_l[kk] == l) vl += S.ecp_c[kk] * ipow(d, S.ecp_p[kk]) * qexp(-S.ecp_z[kk] * d * d);
    ang = fma(vl / (d * d) * (2 * l + 1), legendre_l(l, cos_t), ang);
  }
  double val = 0.0;
  if (lloc > 0) {  // uniform per warp only if all lanes agree; divergence here is cheap relative to the AO sweep
    SinkMO<NMO> sink;
    sink.init(B.g + (e < S.n_up ? B.off_C : B.off_C2));
    eval_val<CART, QE_LMAX>(B.g, B, B.off_seg, px, py, pz, 0, B.n_grp, sink);
    double ratio = 0.0;
#pragma unroll
    for (int mo = 0; mo < NMO; ++mo) ratio = fma(sink.acc[mo], W[((size_t)e * NMO + mo) * nw + w], ratio);
    if (!det_only) ratio *= qexp(jastrow_delta(S, pos, e, x, y, z, px, py, pz));
    val = ang * S.quad_w[k] * ratio;
  }
  Vnl[(size_t)pt * nw + w] = val;
}

// E4: e_L = sum_e T_e + V_bare + V_ion_ion + V_ecp_local + sum_pts V_nl   (jqmc/hamiltonians.py:225-290)
__global__ void k_reduce_eL(SysDev S, int nw, const double* __restrict__ Te, const double* __restrict__ Vb,
                            const double* __restrict__ Vl, const double* __restrict__ Vnl, double* __restrict__ e_L,
                            double* __restrict__ T_elem, double* __restrict__ V_parts) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  double T = 0, vb = S.v_ion_ion, vl = 0, vnl = 0;
  for (int e = 0; e < S.n_e; ++e) {
    const double te = Te[(size_t)e * nw + w];
    T += te;
    vb += Vb[(size_t)e * nw + w];
    vl += Vl[(size_t)e * nw + w];
    if (T_elem) T_elem[(size_t)w * S.n_e + e] = te;
  }
  if (S.ecp_flag) {
    const int npt = S.n_e * S.NN * S.Nv;
    for (int p = 0; p < npt; ++p) vnl += Vnl[(size_t)p * nw + w];
  }
  e_L[w] = T + (vb + (vl + vnl));
  if (V_parts) {
    V_parts[(size_t)w * 4 + 0] = vb;
    V_parts[(size_t)w * 4 + 1] = vl;
    V_parts[(size_t)w * 4 + 2] = vnl;
    V_parts[(size_t)w * 4 + 3] = 0.0;
  }
}

// generic single-electron move ratios (parity entry; also the LRDMC kinetic mesh): thread = (move, walker)
template <int NMO, bool CART>
__global__ void __launch_bounds__(128)
k_move_ratios(BasisDev B, SysDev S, int nw, const double* __restrict__ r_up, const double* __restrict__ r_dn,
              const double* __restrict__ W, int n_moves, const int* __restrict__ elec, const double* __restrict__ r_new,
              double* __restrict__ det_ratio, double* __restrict__ jas_ratio) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_moves * nw) return;
  const int w = (int)(t % nw);
  const int mv = (int)(t / nw);
  const int e = elec[mv];
  const double* p = r_new + ((size_t)w * n_moves + mv) * 3;
  const double px = p[0], py = p[1], pz = p[2];
  PosGlobal pos{r_up, r_dn, S.n_up, S.n_dn, w};
  if (det_ratio) {
    SinkMO<NMO> sink;
    sink.init(B.g + (e < S.n_up ? B.off_C : B.off_C2));
    eval_val<CART, QE_LMAX>(B.g, B, B.off_seg, px, py, pz, 0, B.n_grp, sink);
    double ratio = 0.0;
#pragma unroll
    for (int mo = 0; mo < NMO; ++mo) ratio = fma(sink.acc[mo], W[((size_t)e * NMO + mo) * nw + w], ratio);
    det_ratio[(size_t)w * n_moves + mv] = ratio;
  }
  if (jas_ratio) {
    double x, y, z;
    pos.get(e, x, y, z);
    jas_ratio[(size_t)w * n_moves + mv] = qexp(jastrow_delta(S, pos, e, x, y, z, px, py, pz));
  }
}

// =================================================================================================
// AS regularisation factor (jqmc/determinant.py:1223-1260): thread = walker
// =================================================================================================
__global__ void k_as_factor(int N, int nw, const double* __restrict__ G, const double* __restrict__ Ginv,
                            double* __restrict__ R_AS) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  const double* g = G + (size_t)w * N * N;
  const double* gi = Ginv + (size_t)w * N * N;
  double F = 0, S = 1e300;
  for (int i = 0; i < N * N; ++i) F = fma(gi[i], gi[i], F);
  for (int i = 0; i < N; ++i) {
    double r = 0, c = 0;
    for (int j = 0; j < N; ++j) {
      r = fma(g[i * N + j], g[i * N + j], r);
      c = fma(g[j * N + i], g[j * N + i], c);
    }
    S = fmin(S, fmin(r, c));
  }
  const double SF = S * F;
  R_AS[w] = SF > 0.0 ? pow(SF, -0.375) : 0.0;
}

// =================================================================================================
// fp64 peak microbenchmark
// =================================================================================================
__global__ void k_dfma_peak(int iters, double* out) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

int qe_local_energy_fused(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                          double* e_L, double* T_elem, double* V_parts, cudaStream_t st, const int* nn_fixed = nullptr);  // qe_walker.cu

// =================================================================================================
// C ABI
// =================================================================================================
static const double QUAD4[] = {0.5773502691896258, 0.5773502691896258, 0.5773502691896258, 0.5773502691896258, -0.5773502691896258, -0.5773502691896258,
                               -0.5773502691896258, 0.5773502691896258, -0.5773502691896258, -0.5773502691896258, -0.5773502691896258, 0.5773502691896258};

static void quadrature(int Nv, std::vector<double>& w, std::vector<double>& g) {
  // jqmc/coulomb_potential.py:102-184
  w.clear();
  g.clear();
  if (Nv == 4) {
    const double q = 1.0 / std::sqrt(3.0);
    (void)QUAD4;
    const double s[4][3] = {{q, q, q}, {q, -q, -q}, {-q, q, -q}, {-q, -q, q}};
    for (auto& r : s) {
      w.push_back(0.25);
      g.insert(g.end(), r, r + 3);
    }
  } else if (Nv == 6 || Nv == 18) {
    const double s[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (auto& r : s) {
      w.push_back(1.0 / 6.0);
      g.insert(g.end(), r, r + 3);
    }
    if (Nv == 18) {
      for (auto& x : w) x = 1.0 / 6.0;
      const double p = 1.0 / std::sqrt(2.0);
      const double t[12][3] = {{p, p, 0}, {p, -p, 0}, {-p, p, 0}, {-p, -p, 0}, {p, 0, p}, {p, 0, -p},
                               {-p, 0, p}, {-p, 0, -p}, {0, p, p}, {0, -p, p}, {0, p, -p}, {0, -p, -p}};
      for (auto& r : t) {
        w.push_back(1.0 / 15.0);
        g.insert(g.end(), r, r + 3);
      }
    }
  } else if (Nv == 12) {
    const double th = std::atan(2.0);
    std::vector<std::pair<double, double>> sph{{0.0, 0.0}, {M_PI, 0.0}};
    for (int k = 0; k < 5; ++k) sph.push_back({th, 2.0 * k * M_PI / 5.0});
    for (int k = 0; k < 5; ++k) sph.push_back({M_PI - th, (2.0 * k + 1.0) * M_PI / 5.0});
    for (auto& s : sph) {
      w.push_back(1.0 / 12.0);
      g.push_back(std::sin(s.first) * std::cos(s.second));
      g.push_back(std::sin(s.first) * std::sin(s.second));
      g.push_back(std::cos(s.first));
    }
  }
}

extern "C" const char* qe_last_error(void) { return g_err.c_str(); }
extern "C" int qe_version(void) {
  return 100;  // 0.1.0, sm_100a build
}

extern "C" void qe_destroy(qe_engine* h) {
  if (!h) return;
  for (auto& r : h->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  h->pool.release();
  if (h->ws) cudaFree(h->ws);
  if (h->branch_ws) cudaFree(h->branch_ws);
  if (h->gemm_ws) cudaFree(h->gemm_ws);
  if (h->phase_clk) cudaFree(h->phase_clk);
  delete h;
}

// Tables of the general path (qe_wide.cu).  Orbital space = MOs when the geminal / J3 orbitals carry an MO layer, else the
// canonical AO rows of the basis image (per-AO scale folded into lambda / j_matrix; rows a shell does not own are zero).
static int build_wide_tables(const qe_system_desc* d, qe_engine* h, const std::vector<int>& row_ao, const std::vector<double>& row_scale,
                             const std::vector<int>& rowj_ao, const std::vector<double>& rowj_scale) {
  WideTabs& T = h->wt;
  const int n_row = (int)row_ao.size();
  const int n_unp = d->n_up - d->n_dn;
  T.n_row = n_row;
  T.has_mo = d->orb_up.n_mo > 0;
  T.no = T.has_mo ? d->orb_up.n_mo : n_row;
  const int no = T.no;
  const int n_orb_ref = T.has_mo ? d->orb_up.n_mo : d->orb_up.n_ao;  // orbital count of the reference's lambda
  const int lam_cols = n_orb_ref + n_unp;
  cudaError_t e = cudaSuccess;
  auto up = [&](const std::vector<double>& v, const double** out) {
    if (e == cudaSuccess) e = h->pool.upload(v, out);
  };
  std::vector<double> lamP((size_t)no * no, 0.0), lamPT((size_t)no * no, 0.0), lamU((size_t)no * std::max(1, n_unp), 0.0);
  if (T.has_mo) {
    auto tables = [&](const qe_basis_desc& q, std::vector<double>& CT, std::vector<double>& C) {
      CT.assign((size_t)no * n_row, 0.0);
      C.assign((size_t)n_row * no, 0.0);
      for (int r = 0; r < n_row; ++r) {
        const int a = row_ao[r];
        if (a < 0) continue;
        for (int o = 0; o < no; ++o) {
          const double v = q.mo_coefficients[(size_t)o * q.n_ao + a] * row_scale[r];
          CT[(size_t)o * n_row + r] = v;
          C[(size_t)r * no + o] = v;
        }
      }
    };
    std::vector<double> CTu, Cu, CTd, Cd;
    tables(d->orb_up, CTu, Cu);
    tables(d->orb_dn, CTd, Cd);
    T.restricted = CTu == CTd;
    up(CTu, &T.CwT_up);
    up(Cu, &T.Cw_up);
    if (T.restricted) {
      T.CwT_dn = T.CwT_up;
      T.Cw_dn = T.Cw_up;
    } else {
      up(CTd, &T.CwT_dn);
      up(Cd, &T.Cw_dn);
    }
    for (int a = 0; a < no; ++a) {
      for (int b = 0; b < no; ++b) lamP[(size_t)a * no + b] = d->lambda_matrix[(size_t)a * lam_cols + b];
      for (int k = 0; k < n_unp; ++k) lamU[(size_t)a * n_unp + k] = d->lambda_matrix[(size_t)a * lam_cols + n_orb_ref + k];
    }
  } else {
    T.restricted = 1;
    for (int r = 0; r < n_row; ++r) {
      const int a = row_ao[r];
      if (a < 0) continue;
      for (int r2 = 0; r2 < n_row; ++r2) {
        const int b = row_ao[r2];
        if (b >= 0) lamP[(size_t)r * no + r2] = row_scale[r] * d->lambda_matrix[(size_t)a * lam_cols + b] * row_scale[r2];
      }
      for (int k = 0; k < n_unp; ++k) lamU[(size_t)r * n_unp + k] = row_scale[r] * d->lambda_matrix[(size_t)a * lam_cols + n_orb_ref + k];
    }
  }
  for (int a = 0; a < no; ++a)
    for (int b = 0; b < no; ++b) lamPT[(size_t)b * no + a] = lamP[(size_t)a * no + b];
  up(lamP, &T.lamP);
  up(lamPT, &T.lamPT);
  up(lamU, &T.lamU);
  T.j3 = d->j3_flag ? 1 : 0;
  if (T.j3) {
    const int nj_row = (int)rowj_ao.size();
    T.nj_row = nj_row;
    T.j3_mo = d->j3_orb.n_mo > 0;
    T.nj = T.j3_mo ? d->j3_orb.n_mo : nj_row;
    const int nj = T.nj, nj_ref = T.j3_mo ? d->j3_orb.n_mo : d->j3_orb.n_ao;
    std::vector<double> Mj((size_t)nj * nj, 0.0), MjT((size_t)nj * nj, 0.0), j1v(nj, 0.0);
    if (T.j3_mo) {
      std::vector<double> CT((size_t)nj * nj_row, 0.0), C((size_t)nj_row * nj, 0.0);
      for (int r = 0; r < nj_row; ++r) {
        const int a = rowj_ao[r];
        if (a < 0) continue;
        for (int o = 0; o < nj; ++o) {
          const double v = d->j3_orb.mo_coefficients[(size_t)o * d->j3_orb.n_ao + a] * rowj_scale[r];
          CT[(size_t)o * nj_row + r] = v;
          C[(size_t)r * nj + o] = v;
        }
      }
      up(CT, &T.CjT);
      up(C, &T.Cj);
      for (int a = 0; a < nj; ++a) {
        for (int b = 0; b < nj; ++b) Mj[(size_t)a * nj + b] = d->j_matrix[(size_t)a * (nj_ref + 1) + b];
        j1v[a] = d->j_matrix[(size_t)a * (nj_ref + 1) + nj_ref];
      }
    } else {
      for (int r = 0; r < nj_row; ++r) {
        const int a = rowj_ao[r];
        if (a < 0) continue;
        for (int r2 = 0; r2 < nj_row; ++r2) {
          const int b = rowj_ao[r2];
          if (b >= 0) Mj[(size_t)r * nj + r2] = rowj_scale[r] * d->j_matrix[(size_t)a * (nj_ref + 1) + b] * rowj_scale[r2];
        }
        j1v[r] = rowj_scale[r] * d->j_matrix[(size_t)a * (nj_ref + 1) + nj_ref];
      }
    }
    for (int a = 0; a < nj; ++a)
      for (int b = 0; b < nj; ++b) MjT[(size_t)b * nj + a] = Mj[(size_t)a * nj + b];
    up(Mj, &T.Mj);
    up(MjT, &T.MjT);
    up(j1v, &T.j1v);
  }
  if (e != cudaSuccess) return fail(QE_ERR_CUDA, std::string("wide tables upload: ") + cudaGetErrorString(e));
  T.present = true;
  return QE_OK;
}

extern "C" int qe_create(const qe_system_desc* d, qe_engine** out) {
  if (!d || !out) return fail(QE_ERR_INVALID, "qe_create: null argument");
  *out = nullptr;
  if (d->n_atom <= 0 || d->n_up <= 0 || d->n_dn < 0 || d->n_dn > d->n_up)
    return fail(QE_ERR_INVALID, "qe_create: need n_atom > 0 and n_up >= n_dn >= 0, n_up > 0");
  if ((d->orb_up.n_mo > 0) != (d->orb_dn.n_mo > 0)) return fail(QE_ERR_INVALID, "up/dn orbitals must both be MOs or both be AOs");
  if ((d->orb_up.n_mo > 0 ? d->orb_up.n_mo : d->orb_up.n_ao) != (d->orb_dn.n_mo > 0 ? d->orb_dn.n_mo : d->orb_dn.n_ao))
    return fail(QE_ERR_INVALID, "orb_num_up != orb_num_dn");
  if (d->n_up > 112) return fail(QE_ERR_UNSUPPORTED, "more than 112 electrons per spin is not implemented in this build");
  if (d->ecp_flag && !(d->Nv == 4 || d->Nv == 6 || d->Nv == 12 || d->Nv == 18)) return fail(QE_ERR_INVALID, "Nv must be 4, 6, 12 or 18");
  if (d->ecp_flag && (d->NN < 1 || d->NN > d->n_atom)) return fail(QE_ERR_INVALID, "NN must be in 1..n_atom");
  if (!same_ao_tables(d->orb_up, d->orb_dn)) return fail(QE_ERR_UNSUPPORTED, "up/dn orbitals must share the same AO tables");

  qe_engine* h = new qe_engine();
  const int n_mo = d->orb_up.n_mo;
  // the register / shared-memory kernels cover MO-basis geminals with <= 16 orbitals, <= 8 electrons per spin, J1 + J2
  h->narrow_ok = !d->j3_flag && n_mo > 0 && n_mo <= 16 && d->n_up <= 8;
  if (d->precision != 0 && d->precision != 1) {
    qe_destroy(h);
    return fail(QE_ERR_INVALID, "qe_create: precision must be 0 (full) or 1 (mixed)");
  }
  h->mixed = d->precision == 1;
  // orbital padding of the register kernels (template NMO).  A non-singular geminal needs n_mo >= n_up; the padding follows
  // max(n_mo, n_up) so that "NMO = 4 implies n_up <= 4" holds for any input (k_walker sizes its Sherman-Morrison rows by it)
  const int n_pad_src = std::max(n_mo, d->n_up);
  h->nmo_pad = n_pad_src <= 4 ? 4 : (n_pad_src <= 8 ? 8 : 16);
  std::vector<int> row_ao, rowj_ao;
  std::vector<double> row_scale, rowj_scale;
  int rc = build_basis(d->orb_up, &d->orb_dn, d->n_atom, d->positions, h->nmo_pad, QE_N_CHUNK, h->pool, h->b_up, h->narrow_ok, row_ao, row_scale);
  if (rc == QE_OK && d->j3_flag)
    rc = build_basis(d->j3_orb, nullptr, d->n_atom, d->positions, 4, QE_N_CHUNK, h->pool, h->b_j3, false, rowj_ao, rowj_scale);
  if (rc == QE_OK) rc = build_wide_tables(d, h, row_ao, row_scale, rowj_ao, rowj_scale);
  if (rc != QE_OK) {
    qe_destroy(h);
    return rc;
  }
  SysDev& S = h->sys;
  S.n_atom = d->n_atom;
  S.n_up = d->n_up;
  S.n_dn = d->n_dn;
  S.n_e = d->n_up + d->n_dn;
  S.n_unp = d->n_up - d->n_dn;
  std::vector<double> Rn(d->positions, d->positions + 3 * d->n_atom), Z(d->effective_charges, d->effective_charges + d->n_atom);
  const int P = h->nmo_pad;
  const int lam_cols = n_mo + S.n_unp;
  std::vector<double> lam_p((size_t)P * P, 0.0), lam_u((size_t)P * std::max(1, S.n_unp), 0.0);
  for (int a = 0; h->narrow_ok && a < n_mo; ++a) {
    for (int b = 0; b < n_mo; ++b) lam_p[(size_t)a * P + b] = d->lambda_matrix[(size_t)a * lam_cols + b];
    for (int k = 0; k < S.n_unp; ++k) lam_u[(size_t)a * S.n_unp + k] = d->lambda_matrix[(size_t)a * lam_cols + n_mo + k];
  }
  std::vector<double> jA(d->n_atom, 0.0), jc(d->n_atom, 0.0);
  S.j1_type = d->j1_type;
  S.j1_a = d->j1_param;
  if (d->j1_type) {
    for (int a = 0; a < d->n_atom; ++a) {
      const double z = d->j1_atomic_numbers[a] - d->j1_core_electrons[a];
      jA[a] = std::pow(2.0 * z, 0.75);
      jc[a] = std::pow(2.0 * z, 0.25);
    }
  }
  S.j2_type = d->j2_type;
  S.j2_a = d->j2_param;
  S.ecp_flag = d->ecp_flag;
  S.Nv = d->ecp_flag ? d->Nv : 0;
  S.NN = d->ecp_flag ? d->NN : 0;
  std::vector<int> e_nuc, e_l, e_off(d->n_atom + 1, 0), e_lmax(d->n_atom, 0);
  std::vector<double> e_z, e_c, e_p, qw, qg;
  S.ecp_lmax = 0;
  if (d->ecp_flag) {
    for (int a = 0; a < d->n_atom; ++a) {
      e_lmax[a] = d->ecp_max_ang_mom_plus_1[a];
      S.ecp_lmax = std::max(S.ecp_lmax, e_lmax[a]);
      for (int k = 0; k < d->n_ecp; ++k)
        if (d->ecp_nucleus_index[k] == a) {
          e_nuc.push_back(a);
          e_l.push_back(d->ecp_ang_moms[k]);
          e_z.push_back(d->ecp_exponents[k]);
          e_c.push_back(d->ecp_coefficients[k]);
          e_p.push_back((double)d->ecp_powers[k]);
        }
      e_off[a + 1] = (int)e_nuc.size();
    }
    if (S.ecp_lmax > 7) {
      qe_destroy(h);
      return fail(QE_ERR_UNSUPPORTED, "ECP angular momentum > 6");
    }
    quadrature(d->Nv, qw, qg);
  }
  S.n_ecp = (int)e_nuc.size();
  S.v_ion_ion = 0.0;  // jqmc/coulomb_potential.py:2521-2571
  for (int a = 0; a < d->n_atom; ++a)
    for (int b = a + 1; b < d->n_atom; ++b) {
      const double dx = Rn[3 * a] - Rn[3 * b], dy = Rn[3 * a + 1] - Rn[3 * b + 1], dz = Rn[3 * a + 2] - Rn[3 * b + 2];
      S.v_ion_ion += Z[a] * Z[b] / std::sqrt(dx * dx + dy * dy + dz * dz);
    }
  cudaError_t e = cudaSuccess;
  DevPool& pl = h->pool;
  e = pl.upload(Rn, &S.Rn);
  if (e == cudaSuccess) e = pl.upload(Z, &S.Zeff);
  if (e == cudaSuccess) e = pl.upload(lam_p, &S.lam_p);
  if (e == cudaSuccess) e = pl.upload(lam_u, &S.lam_u);
  if (e == cudaSuccess) e = pl.upload(jA, &S.j1_A);
  if (e == cudaSuccess) e = pl.upload(jc, &S.j1_c);
  if (e == cudaSuccess) e = pl.upload(e_nuc, &S.ecp_nuc);
  if (e == cudaSuccess) e = pl.upload(e_l, &S.ecp_l);
  if (e == cudaSuccess) e = pl.upload(e_z, &S.ecp_z);
  if (e == cudaSuccess) e = pl.upload(e_c, &S.ecp_c);
  if (e == cudaSuccess) e = pl.upload(e_p, &S.ecp_p);
  if (e == cudaSuccess) e = pl.upload(e_lmax, &S.ecp_lmax_atom);
  if (e == cudaSuccess) e = pl.upload(e_off, &S.ecp_off);
  if (e == cudaSuccess) e = pl.upload(qw, &S.quad_w);
  if (e == cudaSuccess) e = pl.upload(qg, &S.quad_g);
  if (e != cudaSuccess) {
    qe_destroy(h);
    return fail(QE_ERR_CUDA, std::string("qe_create upload: ") + cudaGetErrorString(e));
  }
  *out = h;
  return QE_OK;
}

extern "C" int64_t qe_launch_count(qe_engine* h) { return h ? h->launches : 0; }
extern "C" int qe_set_path(qe_engine* h, int path) {
  if (!h || (path != 0 && path != 1)) return fail(QE_ERR_INVALID, "qe_set_path: path must be 0 (automatic) or 1 (general path)");
  h->path = path;
  return QE_OK;
}
extern "C" int qe_set_fused(qe_engine* h, int on) {
  if (!h) return fail(QE_ERR_INVALID, "qe_set_fused: null engine");
  h->fused = on != 0;
  return QE_OK;
}

static void prof_clear(qe_engine* h) {
  for (auto& r : h->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  h->prof.clear();
}
extern "C" int qe_profile(qe_engine* h, int enable) {
  if (!h) return fail(QE_ERR_INVALID, "qe_profile: null engine");
  prof_clear(h);
  h->profiling = enable != 0;
  return QE_OK;
}
extern "C" int qe_profile_kernels(void) { return K_COUNT; }
extern "C" const char* qe_profile_name(int id) { return (id >= 0 && id < K_COUNT) ? KERNEL_NAMES[id] : ""; }
extern "C" int qe_profile_read(qe_engine* h, int id, double* total_ms, int64_t* count) {
  if (!h || !total_ms || !count || id < 0 || id >= K_COUNT) return fail(QE_ERR_INVALID, "qe_profile_read: bad argument");
  double tot = 0;
  int64_t n = 0;
  for (auto& r : h->prof)
    if (r.id == id) {
      CUDA_TRY(cudaEventSynchronize(r.e1));
      float ms = 0;
      CUDA_TRY(cudaEventElapsedTime(&ms, r.e0, r.e1));
      tot += ms;
      ++n;
    }
  *total_ms = tot;
  *count = n;
  return QE_OK;
}

extern "C" int qe_eval_orbitals(qe_engine* h, int which, int layer, int n_pts, const double* r, double* out, void* stream) {
  if (!h || !r || !out || n_pts <= 0) return fail(QE_ERR_INVALID, "qe_eval_orbitals: bad argument");
  HostBasis* hb = which == 2 ? &h->b_j3 : &h->b_up;
  if (which < 0 || which > 2 || !hb->present) return fail(QE_ERR_INVALID, "qe_eval_orbitals: basis not present");
  cudaStream_t st = (cudaStream_t)stream;
  const BasisDev& B = hb->dev;
  if (layer != 0 && B.n_mo > 0 && (which == 2 || !h->narrow_ok)) return wide_eval_orbitals(h, which, n_pts, r, out, st);
  { LaunchScope ls_(h, K_EVAL, st);
  if (layer == 0 || B.n_mo == 0) {
    if (B.cart) k_eval_ao<true><<<nblk(n_pts, 128), 128, 0, st>>>(B, n_pts, r, out);
    else k_eval_ao<false><<<nblk(n_pts, 128), 128, 0, st>>>(B, n_pts, r, out);
  } else {
#define CALL(NMO, CART) k_eval_mo<NMO, CART><<<nblk(n_pts, 128), 128, 0, st>>>(B, which == 1, n_pts, r, out)
    switch (B.nmo_pad) {
      case 4: if (B.cart) { CALL(4, true); } else { CALL(4, false); } break;
      case 8: if (B.cart) { CALL(8, true); } else { CALL(8, false); } break;
      default: if (B.cart) { CALL(16, true); } else { CALL(16, false); } break;
    }
#undef CALL
  }
  }
  CHECK_LAUNCH();
  return QE_OK;
}

// k_geminal with N_up as a compile-time constant when it is small (register-resident Gauss-Jordan)
#define GEMINAL_LAUNCH(NMO, NT, ARGS)                                                                                     \
  do {                                                                                                                    \
    const size_t smem_ = (size_t)S.n_e * NMO * 32 * 8;                                                                    \
    CUDA_TRY(cudaFuncSetAttribute(k_geminal<NMO, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));         \
    k_geminal<NMO, NT><<<nblk(nw, 32), dim3(32, 8), smem_, st>>> ARGS;                                                    \
  } while (0)
#define GEMINAL_NT(NMO, ARGS)                          \
  do {                                                 \
    switch (S.n_up) {                                  \
      case 1: GEMINAL_LAUNCH(NMO, 1, ARGS); break;     \
      case 2: GEMINAL_LAUNCH(NMO, 2, ARGS); break;     \
      case 3: GEMINAL_LAUNCH(NMO, 3, ARGS); break;     \
      case 4: GEMINAL_LAUNCH(NMO, 4, ARGS); break;     \
      case 5: GEMINAL_LAUNCH(NMO, 5, ARGS); break;     \
      case 6: GEMINAL_LAUNCH(NMO, 6, ARGS); break;     \
      default: GEMINAL_LAUNCH(NMO, 0, ARGS); break;    \
    }                                                  \
  } while (0)

static size_t ws_need_common(const qe_engine* h, int nw, int n_chunk, int nq) {
  const SysDev& S = h->sys;
  size_t n = 0;
  n += (size_t)n_chunk * S.n_e * nq * h->nmo_pad * nw * 8 + 256;  // phi
  n += (size_t)S.n_e * h->nmo_pad * nw * 8 + 256;                 // W
  n += 3 * ((size_t)S.n_e * nw * 8 + 256);                        // Te, Vb, Vl
  n += (size_t)std::max(1, S.n_e * S.NN * S.Nv) * nw * 8 + 256;   // Vnl
  return n + 4096;
}

extern "C" int qe_geminal_init(qe_engine* h, int nw, const double* r_up, const double* r_dn, double* G, double* Ginv,
                               void* stream) {
  if (!h || nw <= 0 || !r_up || (!r_dn && h->sys.n_dn > 0)) return fail(QE_ERR_INVALID, "qe_geminal_init: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_wide(h)) return wide_geminal_init(h, nw, r_up, r_dn, G, Ginv, nullptr, nullptr, st);
  const SysDev& S = h->sys;
  int rc;
  double* phi;
  rc = ensure_ws(h, ws_need_common(h, nw, h->b_up.n_chunk, 1));
  if (rc) return rc;
  WsCarve c2{(char*)h->ws};
  phi = c2.take<double>((size_t)h->b_up.n_chunk * S.n_e * h->nmo_pad * nw);
  const long long total2 = (long long)h->b_up.n_chunk * S.n_e * nw;
  { LaunchScope ls_(h, K_ORB_EL, st);
#define CALL(NMO, CART)                                                                                              \
  k_orb_electrons<NMO, CART, 1><<<nblk(total2, 128), 128, 0, st>>>(h->b_up.dev, S, nw, r_up, r_dn, h->b_up.off_cseg, \
                                                                   h->b_up.off_cbeg, h->b_up.n_chunk, phi)
  DISPATCH_NMO_CART(h, CALL);
#undef CALL
  }
  CHECK_LAUNCH();
  { LaunchScope ls_(h, K_GEMINAL, st);
#define CALL(NMO) GEMINAL_NT(NMO, (S, nw, h->b_up.n_chunk, phi, r_up, r_dn, G, Ginv, nullptr, nullptr))
  DISPATCH_NMO(h, CALL);
#undef CALL
  }
  CHECK_LAUNCH();
  return QE_OK;
}

extern "C" int qe_ln_wavefunction(qe_engine* h, int nw, const double* r_up, const double* r_dn, double* ln_psi, double* sign,
                                  void* stream) {
  if (!h || nw <= 0 || !r_up || !ln_psi) return fail(QE_ERR_INVALID, "qe_ln_wavefunction: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_wide(h)) return wide_geminal_init(h, nw, r_up, r_dn, nullptr, nullptr, ln_psi, sign, st);
  const SysDev& S = h->sys;
  int rc = ensure_ws(h, ws_need_common(h, nw, h->b_up.n_chunk, 1));
  if (rc) return rc;
  WsCarve c{(char*)h->ws};
  double* phi = c.take<double>((size_t)h->b_up.n_chunk * S.n_e * h->nmo_pad * nw);
  const long long total = (long long)h->b_up.n_chunk * S.n_e * nw;
  { LaunchScope ls_(h, K_ORB_EL, st);
#define CALL(NMO, CART)                                                                                             \
  k_orb_electrons<NMO, CART, 1><<<nblk(total, 128), 128, 0, st>>>(h->b_up.dev, S, nw, r_up, r_dn, h->b_up.off_cseg, \
                                                                   h->b_up.off_cbeg, h->b_up.n_chunk, phi)
  DISPATCH_NMO_CART(h, CALL);
#undef CALL
  }
  CHECK_LAUNCH();
  { LaunchScope ls_(h, K_GEMINAL, st);
#define CALL(NMO) GEMINAL_NT(NMO, (S, nw, h->b_up.n_chunk, phi, r_up, r_dn, nullptr, nullptr, ln_psi, sign))
  DISPATCH_NMO(h, CALL);
#undef CALL
  }
  CHECK_LAUNCH();
  return QE_OK;
}

extern "C" int qe_as_factor(qe_engine* h, int nw, const double* G, const double* Ginv, double* R_AS, void* stream) {
  if (!h || nw <= 0 || !G || !Ginv || !R_AS) return fail(QE_ERR_INVALID, "qe_as_factor: bad argument");
  { LaunchScope ls_(h, K_AS, (cudaStream_t)stream);
  k_as_factor<<<nblk(nw, 128), 128, 0, (cudaStream_t)stream>>>(h->sys.n_up, nw, G, Ginv, R_AS);
  }
  CHECK_LAUNCH();
  return QE_OK;
}

extern "C" int qe_local_energy(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                               double* e_L, double* T_elem, double* V_parts, void* stream) {
  if (!h || nw <= 0 || !r_up || !Ginv || !e_L || (!r_dn && h->sys.n_dn > 0)) return fail(QE_ERR_INVALID, "qe_local_energy: bad argument");
  if (h->sys.ecp_flag && !RT) return fail(QE_ERR_INVALID, "qe_local_energy: RT required for ECP systems");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_wide(h)) return wide_local_energy(h, nw, r_up, r_dn, RT, Ginv, e_L, T_elem, V_parts, st);
  const SysDev& S = h->sys;
  if (h->fused) {
    const int frc = qe_local_energy_fused(h, nw, r_up, r_dn, RT, Ginv, e_L, T_elem, V_parts, st);
    if (frc != QE_ERR_UNSUPPORTED) return frc;
  }
  const int nch = h->b_up.n_chunk, P = h->nmo_pad;
  int rc = ensure_ws(h, ws_need_common(h, nw, nch, 5));
  if (rc) return rc;
  WsCarve c{(char*)h->ws};
  double* phi = c.take<double>((size_t)nch * S.n_e * 5 * P * nw);
  double* W = c.take<double>((size_t)S.n_e * P * nw);
  double* Te = c.take<double>((size_t)S.n_e * nw);
  double* Vb = c.take<double>((size_t)S.n_e * nw);
  double* Vl = c.take<double>((size_t)S.n_e * nw);
  double* Vnl = c.take<double>((size_t)std::max(1, S.n_e * S.NN * S.Nv) * nw);
  const long long t1 = (long long)nch * S.n_e * nw;
  { LaunchScope ls_(h, K_ORB_EL, st);
#define CALL(NMO, CART)                                                                                          \
  k_orb_electrons<NMO, CART, 5><<<nblk(t1, 128), 128, 0, st>>>(h->b_up.dev, S, nw, r_up, r_dn, h->b_up.off_cseg, \
                                                                   h->b_up.off_cbeg, nch, phi)
  DISPATCH_NMO_CART(h, CALL);
#undef CALL
  }
  CHECK_LAUNCH();
  const long long t2 = (long long)S.n_e * nw;
  { LaunchScope ls_(h, K_ALGEBRA, st);
#define CALL(NMO) k_electron_algebra<NMO, 5><<<nblk(t2, 128), 128, 0, st>>>(S, nw, nch, phi, r_up, r_dn, Ginv, W, Te, Vb, Vl)
  DISPATCH_NMO(h, CALL);
#undef CALL
  }
  CHECK_LAUNCH();
  if (S.ecp_flag) {
    const long long t3 = (long long)S.n_e * S.NN * S.Nv * nw;
    { LaunchScope ls_(h, K_ECP_MESH, st);
#define CALL(NMO, CART) \
  k_ecp_mesh<NMO, CART><<<nblk(t3, 128), 128, 0, st>>>(h->b_up.dev, S, nw, r_up, r_dn, RT, W, 0, Vnl, nullptr)
    DISPATCH_NMO_CART(h, CALL);
#undef CALL
    }
    CHECK_LAUNCH();
  }
  { LaunchScope ls_(h, K_REDUCE, st);
  k_reduce_eL<<<nblk(nw, 128), 128, 0, st>>>(S, nw, Te, Vb, Vl, Vnl, e_L, T_elem, V_parts);
  }
  CHECK_LAUNCH();
  return QE_OK;
}

extern "C" int qe_move_ratios(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* Ginv, int n_moves,
                              const int32_t* elec_host, const double* r_new, double* det_ratio, double* jas_ratio, void* stream) {
  if (!h || nw <= 0 || n_moves <= 0 || !r_up || !Ginv || !elec_host || !r_new) return fail(QE_ERR_INVALID, "qe_move_ratios: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const SysDev& S = h->sys;
  for (int i = 0; i < n_moves; ++i)
    if (elec_host[i] < 0 || elec_host[i] >= S.n_e) return fail(QE_ERR_INVALID, "qe_move_ratios: electron index out of range");
  if (use_wide(h)) return wide_move_ratios(h, nw, r_up, r_dn, Ginv, n_moves, elec_host, r_new, det_ratio, jas_ratio, st);
  const int nch = h->b_up.n_chunk, P = h->nmo_pad;
  int rc = ensure_ws(h, ws_need_common(h, nw, nch, 1) + (size_t)n_moves * 4 + 256);
  if (rc) return rc;
  WsCarve c{(char*)h->ws};
  double* phi = c.take<double>((size_t)nch * S.n_e * P * nw);
  double* W = c.take<double>((size_t)S.n_e * P * nw);
  int* elec = c.take<int>(n_moves);
  CUDA_TRY(cudaMemcpyAsync(elec, elec_host, (size_t)n_moves * 4, cudaMemcpyHostToDevice, st));
  const long long t1 = (long long)nch * S.n_e * nw;
  { LaunchScope ls_(h, K_ORB_EL, st);
#define CALL(NMO, CART)                                                                                          \
  k_orb_electrons<NMO, CART, 1><<<nblk(t1, 128), 128, 0, st>>>(h->b_up.dev, S, nw, r_up, r_dn, h->b_up.off_cseg, \
                                                                   h->b_up.off_cbeg, nch, phi)
  DISPATCH_NMO_CART(h, CALL);
#undef CALL
  }
  CHECK_LAUNCH();
  const long long t2 = (long long)S.n_e * nw;
  { LaunchScope ls_(h, K_ALGEBRA, st);
#define CALL(NMO) \
  k_electron_algebra<NMO, 1><<<nblk(t2, 128), 128, 0, st>>>(S, nw, nch, phi, r_up, r_dn, Ginv, W, nullptr, nullptr, nullptr)
  DISPATCH_NMO(h, CALL);
#undef CALL
  }
  CHECK_LAUNCH();
  const long long t3 = (long long)n_moves * nw;
  { LaunchScope ls_(h, K_RATIOS, st);
#define CALL(NMO, CART)                                                                                                   \
  k_move_ratios<NMO, CART><<<nblk(t3, 128), 128, 0, st>>>(h->b_up.dev, S, nw, r_up, r_dn, W, n_moves, elec, \
                                                          r_new, det_ratio, jas_ratio)
  DISPATCH_NMO_CART(h, CALL);
#undef CALL
  }
  CHECK_LAUNCH();
  return QE_OK;
}

extern "C" int qe_dln_wf(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* Ginv, double* d_j1, double* d_j2,
                         double* d_j3, double* d_lambda, void* stream) {
  if (!h || nw <= 0 || !r_up || !Ginv || (!r_dn && h->sys.n_dn > 0)) return fail(QE_ERR_INVALID, "qe_dln_wf: bad argument");
  if (d_j1 && !h->sys.j1_type) return fail(QE_ERR_INVALID, "qe_dln_wf: no one-body Jastrow in this Hamiltonian");
  if (d_j2 && !h->sys.j2_type) return fail(QE_ERR_INVALID, "qe_dln_wf: no two-body Jastrow in this Hamiltonian");
  if (d_j3 && !h->wt.j3) return fail(QE_ERR_INVALID, "qe_dln_wf: no three-body Jastrow in this Hamiltonian");
  return wide_dln_wf(h, nw, r_up, r_dn, Ginv, d_j1, d_j2, d_j3, d_lambda, (cudaStream_t)stream);
}

extern "C" int qe_measure_fp64_peak(int iters, double* tflops) {
  if (!tflops || iters <= 0) return fail(QE_ERR_INVALID, "qe_measure_fp64_peak: bad argument");
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* out = nullptr;
  CUDA_TRY(cudaMalloc(&out, 8));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256;
  k_dfma_peak<<<blocks, threads>>>(iters / 4, out);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaEventRecord(e0));
  k_dfma_peak<<<blocks, threads>>>(iters, out);
  CUDA_TRY(cudaEventRecord(e1));
  CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
  cudaFree(out);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return QE_OK;
}
