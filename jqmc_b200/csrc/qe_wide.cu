// General ("wide") path of the walker engine: any number of orbitals (AO-basis JAGP geminals, large MO sets), any number
// of electrons (<= 112 per spin), one-, two- and three-body Jastrow.  It serves the same C-ABI entry points as the
// register/shared-memory kernels of qe_mcmc.cu / qe_walker.cu (which cover small MO-basis systems) and is selected
// automatically when those do not cover the system (or with qe_set_path(h, 1)).
//
// Layout: every per-walker quantity lives in the engine workspace as [item][walker] (walker innermost), so that lanes are
// walkers, loads/stores coalesce and the contractions against the FIXED matrices of the Hamiltonian (MO coefficients,
// lambda, j_matrix) become ONE dense fp64 GEMM over all walkers:   D[m][(item, w)] = A[m][k] . B[k][(item, w)]
// which runs on the fp64 tensor cores (mma.sync.m8n8k4.f64, kw_dgemm).  Per-walker products (G = Phi^T M, W = M Ginv,
// Sherman-Morrison) are thread-per-output kernels over the same layout.
//
// Algebra (reference: jqmc/determinant.py:1379-1402, 1665-1783, 2140-2250; jqmc/jastrow_factor.py:1744-1801, 2230-2960,
// 4076-4156; jqmc/coulomb_potential.py:1477-1712; jqmc/wavefunction.py:1739-1860; jqmc/jqmc_gfmc.py:4807-5166):
//   Phi[o][e]   orbital values at the electrons          Mfull = [lambda_p Phi_dn | lambda_u]   (no x N_up)
//   MupT = lambda_p^T Phi_up (no x N_up)                  G = Phi_up^T Mfull
//   det ratio of moving up electron k to r':  phi(r') . Worb[:,k],  Worb[:,k] = Mfull Ginv[:,k]
//   det ratio of moving dn electron j to r':  phi(r') . Worb[:,N+j], Worb[:,N+j] = MupT Ginv[j,:]^T
//   in AO-row space: Wrow = Cw Worb, so a mesh ratio costs ONE AO sweep and one multiply-add per AO (no AO->MO product)
//   J3 = j1.sum_i chi_i + sum_{i<j} chi_i^T M chi_j (electrons ordered up then down), hence for electron k
//   J3(r') - J3(r) = (chi(r') - chi(r_k)) . g_k,  g_k = j1 + sum_{i<k} M^T chi_i + sum_{i>k} M chi_i,  grad_k J3 = g_k . grad chi(r_k)
#include "qe_common.cuh"

namespace {

// =================================================================================================
// fp64 tensor-core GEMM   D[m x n] = A[m x k] . B[k x n]   (row-major; n = (item, walker) columns, huge; m, k = orbital counts)
// CTA = 4 warps, tile 32 (m) x 128 (n), k-step 16; warp tile 32 x 32 = 4 x 4 mma.m8n8k4 accumulators.
// Shared tiles are padded by 4 doubles so that both fragment loads are conflict-free per half-warp.
// =================================================================================================
constexpr int GM = 32, GN = 128, GK = 16;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128)
kw_dgemm(int m, long long n, int k, const double* __restrict__ A, int lda, const double* __restrict__ B, long long ldb,
         double* __restrict__ D, long long ldd, long long sB, long long sD, int ksplit, int kchunk, long long sPart) {
  // blockIdx.z = batch * ksplit + ks: split ks covers k in [ks*kchunk, (ks+1)*kchunk) and writes its partial product to
  // D + ks*sPart (ksplit == 1: the product itself); kw_sum_parts adds the partials in a fixed order
  __shared__ double As[GM][GK + 4];
  __shared__ double Bs[GK][GN + 4];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long n0 = (long long)blockIdx.x * GN;
  const int m0 = blockIdx.y * GM;
  const int ks = blockIdx.z % ksplit, bz = blockIdx.z / ksplit;
  B += (long long)bz * sB;
  D += (long long)bz * sD + (long long)ks * sPart;
  const int kbeg = ks * kchunk;
  k = min(k, kbeg + kchunk);
  const int g = lane >> 2, t = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  // software pipeline: the next k-tile travels global -> registers while the tensor cores work on the current shared tiles
  double ra[GM * GK / 128], rb[GK * GN / 128];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < GM * GK / 128; ++u) {
      const int i = tid + u * 128, r = i / GK, c = i % GK;
      ra[u] = (m0 + r < m && k0 + c < k) ? A[(size_t)(m0 + r) * lda + k0 + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < GK * GN / 128; ++u) {
      const int i = tid + u * 128, r = i / GN, c = i % GN;
      rb[u] = (k0 + r < k && n0 + c < n) ? B[(long long)(k0 + r) * ldb + n0 + c] : 0.0;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < GM * GK / 128; ++u) {
      const int i = tid + u * 128;
      As[i / GK][i % GK] = ra[u];
    }
#pragma unroll
    for (int u = 0; u < GK * GN / 128; ++u) {
      const int i = tid + u * 128;
      Bs[i / GN][i % GN] = rb[u];
    }
  };
  if (kbeg < k) {
    fetch(kbeg);
    stash();
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < k; k0 += GK) {
    const bool more = k0 + GK < k;
    if (more) fetch(k0 + GK);
#pragma unroll
    for (int kk = 0; kk < GK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[8 * i + g][kk + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk + t][wid * 32 + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
    if (more) {
      stash();
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + 8 * i + g;
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long col = n0 + wid * 32 + 8 * j + 2 * t;
      double* d = D + (long long)row * ldd + col;
      if (col < n) d[0] = acc[i][j][0];
      if (col + 1 < n) d[1] = acc[i][j][1];
    }
  }
}

// plain DFMA reference of the same product (qe_set_path debugging aid and the GEMM self-check of the tests)
__global__ void kw_dgemm_ref(int m, long long n, int k, const double* __restrict__ A, int lda, const double* __restrict__ B,
                             long long ldb, double* __restrict__ D, long long ldd, long long sB, long long sD) {
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (col >= n) return;
  B += (long long)blockIdx.z * sB;
  D += (long long)blockIdx.z * sD;
  double s = 0.0;
  for (int kk = 0; kk < k; ++kk) s = fma(A[(size_t)row * lda + kk], B[(long long)kk * ldb + col], s);
  D[(long long)row * ldd + col] = s;
}

// D[i] = sum_s part[s][i]  (fixed order: deterministic)
__global__ void kw_sum_parts(long long n, int ksplit, const double* __restrict__ part, double* __restrict__ D, long long rows_n,
                             long long ldd, long long ncol) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double s = 0.0;
  for (int i = 0; i < ksplit; ++i) s += part[(long long)i * n + t];
  // t = (batch*m + row) * ncol + col  ->  D[batch*sD + row*ldd + col] with sD = rows_n (elements per batch in D)
  const long long col = t % ncol, br = t / ncol;
  (void)rows_n;
  D[br * ldd + col] = s;
}

// per-walker product  D[i][j][w] = sum_a X[a][i][w] Y[a][j][w]  with arbitrary (element) strides; thread = (i, j, w)
__global__ void kw_bmm(int ni, int nj, int na, int nw, const double* __restrict__ X, long long sxa, long long sxi,
                       const double* __restrict__ Y, long long sya, long long syj, double* __restrict__ D, long long sdi, long long sdj) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)ni * nj * nw) return;
  const int w = (int)(t % nw);
  const int j = (int)((t / nw) % nj);
  const int i = (int)(t / ((long long)nw * nj));
  const double* x = X + i * sxi + w;
  const double* y = Y + j * syj + w;
  double s = 0.0;
  for (int a = 0; a < na; ++a) s = fma(x[a * sxa], y[a * sya], s);
  D[i * sdi + j * sdj + w] = s;
}

// Ratio weight vectors of both spins in one pass for small determinants (N_up <= 8):
//   Worb[o][e][w]     = sum_j Mfull[o][j][w] Ginv[j][e][w]   (e < N_up)
//   Worb[o][N+j'][w]  = sum_i MupT[o][i][w]  Ginv[j'][i][w]  (j' < N_dn)
// block = 32 walkers (x) x 8 (y): the running inverse of the block's walkers is staged in shared memory once, a thread then
// walks orbitals o = y, y + 8, ... with the 2 N coefficients of the orbital in registers.  Same summation order as kw_bmm.
template <int NB>
__global__ void __launch_bounds__(256)
kw_worb_small(int no, int N, int Nd, int nw, const double* __restrict__ Mfull, const double* __restrict__ MupT,
              const double* __restrict__ Gi, double* __restrict__ Worb) {
  __shared__ double s_g[NB * NB][32];
  const int lane = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
  const int wq = blockIdx.x * 32 + lane;
  const bool live = wq < nw;
  const int w = live ? wq : nw - 1;
  const int Ne = N + Nd;
  for (int i = ty; i < N * N; i += TY) s_g[i][lane] = Gi[(size_t)i * nw + w];
  __syncthreads();
  if (!live) return;
  for (int o = ty + blockIdx.y * TY; o < no; o += TY * gridDim.y) {
    double mf[NB], mu[NB];
#pragma unroll
    for (int a = 0; a < NB; ++a) {
      mf[a] = a < N ? Mfull[((size_t)o * N + a) * nw + w] : 0.0;
      mu[a] = a < N ? MupT[((size_t)o * N + a) * nw + w] : 0.0;
    }
    double* out = Worb + (size_t)o * Ne * nw + w;
#pragma unroll
    for (int e = 0; e < NB; ++e) {
      if (e >= N) break;
      double sacc = 0.0;
#pragma unroll
      for (int a = 0; a < NB; ++a)
        if (a < N) sacc = fma(mf[a], s_g[a * N + e][lane], sacc);
      out[(size_t)e * nw] = sacc;
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (j >= Nd) break;
      double sacc = 0.0;
#pragma unroll
      for (int a = 0; a < NB; ++a)
        if (a < N) sacc = fma(mu[a], s_g[j * N + a][lane], sacc);
      out[(size_t)(N + j) * nw] = sacc;
    }
  }
}

// AoS <-> SoA:  src[w][n_item] -> dst[item][w]  and back
__global__ void kw_to_soa(int nw, int n_item, const double* __restrict__ src, double* __restrict__ dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_item * nw) return;
  const int w = (int)(t % nw), it = (int)(t / nw);
  dst[t] = src[(size_t)w * n_item + it];
}
__global__ void kw_to_aos(int nw, int n_item, const double* __restrict__ src, double* __restrict__ dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_item * nw) return;
  const int w = (int)(t % nw), it = (int)(t / nw);
  dst[(size_t)w * n_item + it] = src[t];
}
// positions: r_up[w][N][3], r_dn[w][Nd][3] <-> rs[(e*3+c)][w]
__global__ void kw_pos_to_soa(int nw, int N, int Nd, const double* __restrict__ r_up, const double* __restrict__ r_dn,
                              double* __restrict__ rs) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ne = N + Nd;
  if (t >= (long long)Ne * 3 * nw) return;
  const int w = (int)(t % nw), it = (int)(t / nw), e = it / 3, c = it % 3;
  rs[t] = e < N ? r_up[((size_t)w * N + e) * 3 + c] : r_dn[((size_t)w * Nd + (e - N)) * 3 + c];
}
__global__ void kw_pos_to_aos(int nw, int N, int Nd, const double* __restrict__ rs, double* __restrict__ r_up, double* __restrict__ r_dn) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ne = N + Nd;
  if (t >= (long long)Ne * 3 * nw) return;
  const int w = (int)(t % nw), it = (int)(t / nw), e = it / 3, c = it % 3;
  if (e < N) r_up[((size_t)w * N + e) * 3 + c] = rs[t];
  else r_dn[((size_t)w * Nd + (e - N)) * 3 + c] = rs[t];
}
// Mfull[o][Nd+q][w] = lambda_u[o][q]
__global__ void kw_fill_unpaired(int no, int N, int Nd, int nw, const double* __restrict__ lamU, double* __restrict__ Mfull) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nu = N - Nd;
  if (t >= (long long)no * nu * nw) return;
  const int w = (int)(t % nw), q = (int)((t / nw) % nu), o = (int)(t / ((long long)nw * nu));
  Mfull[((size_t)o * N + Nd + q) * nw + w] = lamU[(size_t)o * nu + q];
}

struct PosSoA {
  const double* __restrict__ rs;
  int nw, w;
  __device__ __forceinline__ void get(int e, double& x, double& y, double& z) const {
    x = rs[((size_t)e * 3 + 0) * nw + w];
    y = rs[((size_t)e * 3 + 1) * nw + w];
    z = rs[((size_t)e * 3 + 2) * nw + w];
  }
};

// =================================================================================================
// AO sweeps
// =================================================================================================
// raw AO rows (value or value/grad/lap) at points pts[(item*3+c)][w] -> out[q][row][item][w]; thread = (chunk, item, walker)
struct SinkRows {
  double* __restrict__ o;
  size_t sr, sq;
  __device__ __forceinline__ void add(int row, double v) { o[row * sr] = v; }
  __device__ __forceinline__ void add_n(int row, const double* __restrict__ v) { o[row * sr] = v[0]; }
  template <int NF>
  __device__ __forceinline__ void prefetch(int, double*) const {}
  __device__ __forceinline__ void add_nw(int row, const double* __restrict__ v, double) { o[row * sr] = v[0]; }
  __device__ __forceinline__ void add(int row, double v, double gx, double gy, double gz, double lp) {
    double* p = o + row * sr;
    p[0] = v;
    p[sq] = gx;
    p[2 * sq] = gy;
    p[3 * sq] = gz;
    p[4 * sq] = lp;
  }
};
template <bool CART, int NQ>
__global__ void __launch_bounds__(128)
kw_ao_store(BasisDev B, int off_cseg, int off_cbeg, int n_chunk, int n_item, int nw, const double* __restrict__ pts,
            double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_chunk * n_item * nw) return;
  const int w = (int)(t % nw);
  const int item = (int)((t / nw) % n_item);
  const int c = (int)(t / ((long long)nw * n_item));
  const double x = pts[((size_t)item * 3 + 0) * nw + w], y = pts[((size_t)item * 3 + 1) * nw + w], z = pts[((size_t)item * 3 + 2) * nw + w];
  const int* cbeg = (const int*)(B.g + off_cbeg);
  SinkRows sink{out + (size_t)item * nw + w, (size_t)n_item * nw, (size_t)B.n_row * n_item * nw};
  if (NQ == 1) eval_val<CART, QE_LMAX>(B.g, B, off_cseg, x, y, z, cbeg[c], cbeg[c + 1], sink);
  else eval_vgl<CART, QE_LMAX>(B.g, B, off_cseg, x, y, z, cbeg[c], cbeg[c + 1], sink);
}

// dot of the AO rows at NP points with one weight vector W[row] (stride between rows: sr)
template <int NP>
struct SinkDotN {
  const double* __restrict__ W;
  size_t sr;
  double acc[NP];
  __device__ __forceinline__ void init(const double* w_, size_t sr_) {
    W = w_;
    sr = sr_;
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i] = 0.0;
  }
  __device__ __forceinline__ void add_n(int row, const double* __restrict__ v) {
    const double wv = W[row * sr];
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i] = fma(v[i], wv, acc[i]);
  }
  template <int NF>
  __device__ __forceinline__ void prefetch(int row0, double* __restrict__ w) const {
#pragma unroll
    for (int k = 0; k < NF; ++k) w[k] = W[(row0 + k) * sr];
  }
  __device__ __forceinline__ void add_nw(int, const double* __restrict__ v, double wv) {
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i] = fma(v[i], wv, acc[i]);
  }
};

// =================================================================================================
// mesh ratios: LRDMC kinetic mesh (6 N_e points, jqmc/wavefunction.py:1739-1860) and the non-local ECP quadrature
// (N_e NN Nv points, jqmc/coulomb_potential.py:1477-1712).  thread = (pair of points of one electron, walker).
//   p[t][w]: kinetic  -Psi'/(2 a^2 Psi);  ECP  ang * w_k * ratio (tmove: x Jastrow ratio; dltmove: determinant part only)
//   sj[t - n_kin][w]: Jastrow ratio of the ECP points
// =================================================================================================
struct MeshArgs {
  int nw, n_kin, n_ecp, dlt, det_only;
  double alat;
  const double* rs;     // [Ne*3][w]
  const double* RT;     // [9][w]
  const double* Wrow;   // [n_row][Ne][w]
  const double* gJrow;  // [nj_row][Ne][w]
  const double* cJ;     // [Ne][w]  chi(r_e) . g_e
  const double* el;     // [Ne][8][w]: slot 7 = J1+J2 terms of electron e at its current position (kw_electron)
  double* p;
  double* sj;
};
template <bool CART, bool CARTJ, int LMAX, bool SMEM>
__global__ void __launch_bounds__(128)
kw_mesh(BasisDev B, BasisDev BJ, int has_j3, SysDev S, MeshArgs P) {
  // basis images (shell / primitive tables of the determinant and J3 bases) staged in shared memory once per CTA; the CTA
  // then walks its share of the (pair, walker) work items with a grid-stride loop
  extern __shared__ __align__(16) char sm_mesh[];
  const char* tab = B.g;
  const char* tabJ = BJ.g;
  if (SMEM) {
    const int nB = B.bytes / 16, nJ = has_j3 ? BJ.bytes / 16 : 0;
    int4* dst = (int4*)sm_mesh;
    const int4* srcB = (const int4*)B.g;
    const int4* srcJ = (const int4*)BJ.g;
    for (int i = threadIdx.x; i < nB; i += blockDim.x) dst[i] = srcB[i];
    for (int i = threadIdx.x; i < nJ; i += blockDim.x) dst[nB + i] = srcJ[i];
    __syncthreads();
    tab = sm_mesh;
    tabJ = sm_mesh + B.bytes;
  }
  const int n_pairs = (P.n_kin + P.n_ecp) / 2;
  const long long total = (long long)n_pairs * P.nw;
  const int Ne = S.n_e, nw = P.nw;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(t % P.nw);
    const int t0 = 2 * (int)(t / P.nw);
    PosSoA pos{P.rs, nw, w};
    double rt[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) rt[c] = P.RT[(size_t)c * nw + w];
    int e;
    double px[2], py[2], pz[2], angw[2] = {0.0, 0.0};
    double x, y, z;
    const bool kin = t0 < P.n_kin;
    if (kin) {
      e = t0 / 6;
      pos.get(e, x, y, z);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int s6 = (t0 + i) % 6, ax = s6 >> 1;
        const double sg = (s6 & 1) ? -P.alat : P.alat;
        px[i] = x + sg * rt[3 * ax];
        py[i] = y + sg * rt[3 * ax + 1];
        pz[i] = z + sg * rt[3 * ax + 2];
      }
    } else {
      const int pt0 = t0 - P.n_kin;
      e = pt0 / (S.Nv * S.NN);
      pos.get(e, x, y, z);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int pt = pt0 + i;
        const int k = pt % S.Nv, nn = (pt / S.Nv) % S.NN;
        ecp_point(S, rt, x, y, z, nn, k, px[i], py[i], pz[i], angw[i], true);
      }
    }
    SinkDotN<2> sink;
    sink.init(P.Wrow + (size_t)e * nw + w, (size_t)Ne * nw);
    eval_val_n<CART, LMAX, 2>(tab, B, B.off_seg, px, py, pz, 0, B.n_grp, sink);
    double jr[2] = {1.0, 1.0};
    if (!P.det_only) {
      const double jold = P.el[((size_t)e * 8 + 7) * nw + w];
      double d3[2] = {0.0, 0.0};
      if (has_j3) {
        SinkDotN<2> sj3;
        sj3.init(P.gJrow + (size_t)e * nw + w, (size_t)Ne * nw);
        eval_val_n<CARTJ, LMAX, 2>(tabJ, BJ, BJ.off_seg, px, py, pz, 0, BJ.n_grp, sj3);
        const double c0 = P.cJ[(size_t)e * nw + w];
        d3[0] = sj3.acc[0] - c0;
        d3[1] = sj3.acc[1] - c0;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) jr[i] = qexp(jastrow_single_q(S, pos, e, px[i], py[i], pz[i]) - jold + d3[i]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double ratio = sink.acc[i];
      const size_t idx = (size_t)(t0 + i) * nw + w;
      if (kin) {
        P.p[idx] = -1.0 / (2.0 * P.alat * P.alat) * (ratio * jr[i]);
      } else {
        P.p[idx] = P.dlt ? angw[i] * ratio : angw[i] * (ratio * jr[i]);
        if (P.sj) P.sj[idx - (size_t)P.n_kin * nw] = jr[i];
      }
    }
  }
}

// generic single-electron move ratios (parity entry qe_move_ratios): thread = (move, walker)
template <bool CART, bool CARTJ>
__global__ void __launch_bounds__(128)
kw_move_ratios(BasisDev B, BasisDev BJ, int has_j3, SysDev S, int nw, const double* __restrict__ rs, const double* __restrict__ Wrow,
               const double* __restrict__ gJrow, const double* __restrict__ cJ, int n_moves, const int* __restrict__ elec,
               const double* __restrict__ r_new, double* __restrict__ det_ratio, double* __restrict__ jas_ratio) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_moves * nw) return;
  const int w = (int)(t % nw), mv = (int)(t / nw);
  const int e = elec[mv], Ne = S.n_e;
  const double* pn = r_new + ((size_t)w * n_moves + mv) * 3;
  double px[1] = {pn[0]}, py[1] = {pn[1]}, pz[1] = {pn[2]};
  PosSoA pos{rs, nw, w};
  if (det_ratio) {
    SinkDotN<1> sink;
    sink.init(Wrow + (size_t)e * nw + w, (size_t)Ne * nw);
    eval_val_n<CART, QE_LMAX, 1>(B.g, B, B.off_seg, px, py, pz, 0, B.n_grp, sink);
    det_ratio[(size_t)w * n_moves + mv] = sink.acc[0];
  }
  if (jas_ratio) {
    double x, y, z;
    pos.get(e, x, y, z);
    double d = jastrow_single_q(S, pos, e, px[0], py[0], pz[0]) - jastrow_single_q(S, pos, e, x, y, z);
    if (has_j3) {
      SinkDotN<1> sj3;
      sj3.init(gJrow + (size_t)e * nw + w, (size_t)Ne * nw);
      eval_val_n<CARTJ, QE_LMAX, 1>(BJ.g, BJ, BJ.off_seg, px, py, pz, 0, BJ.n_grp, sj3);
      d += sj3.acc[0] - cJ[(size_t)e * nw + w];
    }
    jas_ratio[(size_t)w * n_moves + mv] = qexp(d);
  }
}

// =================================================================================================
// J3 weight vectors  g_e = j1 + sum_{i<e} V[:,i] + sum_{i>e} U[:,i]  (U = M chi, V = M^T chi) and c_e = chi_e . g_e
// =================================================================================================
__global__ void kw_j3_weights(int nj, int Ne, int nw, const double* __restrict__ j1v, const double* __restrict__ U,
                              const double* __restrict__ V, double* __restrict__ gJ) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nj * nw) return;
  const int w = (int)(t % nw), o = (int)(t / nw);
  const double* u = U + (size_t)o * Ne * nw + w;
  const double* v = V + (size_t)o * Ne * nw + w;
  double* g = gJ + (size_t)o * Ne * nw + w;
  double suf = 0.0;
  for (int e = Ne - 1; e >= 0; --e) {  // suffix sums of U into g
    g[(size_t)e * nw] = suf;
    suf += u[(size_t)e * nw];
  }
  double pre = j1v[o];
  for (int e = 0; e < Ne; ++e) {
    g[(size_t)e * nw] += pre;
    pre += v[(size_t)e * nw];
  }
}
// out[e][w] = sum_o X[o][e][w] Y[o][e][w]
__global__ void kw_coldot(int no, int Ne, int nw, const double* __restrict__ X, const double* __restrict__ Y, double* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)Ne * nw) return;
  double s = 0.0;
  for (int o = 0; o < no; ++o) s = fma(X[(size_t)o * Ne * nw + t], Y[(size_t)o * Ne * nw + t], s);
  out[t] = s;
}

// =================================================================================================
// per-(electron, walker) terms: continuum kinetic energy (jqmc/wavefunction.py:1141-1207: grad/lap ln det
// determinant.py:2140-2250, Jastrow jastrow_factor.py:960-1034, 3434-3558, 4076-4156), bare / discretised el-ion,
// ECP local (coulomb_potential.py:1144-1246), el-el (pairs j > e).   el[e][8][w] = {T, ei, ei_disc, loc, ee, -, -, -}
// =================================================================================================
struct ElecArgs {
  int nw, no, nj, has_j3;
  double alat;
  const double* rs;
  const double* Phi;   // [5][no][Ne][w]
  const double* Worb;  // [no][Ne][w]
  const double* Chi;   // [5][nj][Ne][w]
  const double* gJ;    // [nj][Ne][w]
  double* el;
};
__global__ void __launch_bounds__(128)
kw_electron(SysDev S, ElecArgs P) {
  // thread = (electron, walker).  The orbital sums stream 4 derivative planes of Phi plus the weights (~180 MB per launch for
  // water JAGP at 4096 walkers); splitting them over several warps per item was measured slower (62 vs 52 us: the planes are
  // already read at about half the HBM peak and the extra warps only scatter the DRAM pages).
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ne = S.n_e, nw = P.nw;
  if (t >= (long long)Ne * nw) return;
  const int w = (int)(t % nw), e = (int)(t / nw);
  PosSoA pos{P.rs, nw, w};
  double x, y, z;
  pos.get(e, x, y, z);
  const size_t sq = (size_t)P.no * Ne * nw;
  double gD[3] = {0, 0, 0}, lD = 0;
#pragma unroll 4
  for (int o = 0; o < P.no; ++o) {
    const size_t idx = ((size_t)o * Ne + e) * nw + w;
    const double wv = P.Worb[idx];
    gD[0] = fma(P.Phi[sq + idx], wv, gD[0]);
    gD[1] = fma(P.Phi[2 * sq + idx], wv, gD[1]);
    gD[2] = fma(P.Phi[3 * sq + idx], wv, gD[2]);
    lD = fma(P.Phi[4 * sq + idx], wv, lD);
  }
  double gJ[3] = {0, 0, 0}, lJ = 0, ei = 0, eid = 0, loc = 0, ee = 0;
  if (P.has_j3) {
    const size_t sj = (size_t)P.nj * Ne * nw;
    for (int o = 0; o < P.nj; ++o) {
      const size_t idx = ((size_t)o * Ne + e) * nw + w;
      const double g = P.gJ[idx];
      gJ[0] = fma(P.Chi[sj + idx], g, gJ[0]);
      gJ[1] = fma(P.Chi[2 * sj + idx], g, gJ[1]);
      gJ[2] = fma(P.Chi[3 * sj + idx], g, gJ[2]);
      lJ = fma(P.Chi[4 * sj + idx], g, lJ);
    }
  }
  lD -= gD[0] * gD[0] + gD[1] * gD[1] + gD[2] * gD[2];
  const double eps = 1.0e-12;
  for (int a = 0; a < S.n_atom; ++a) {
    const double dx = x - S.Rn[3 * a], dy = y - S.Rn[3 * a + 1], dz = z - S.Rn[3 * a + 2];
    const double d = sqrt(dx * dx + dy * dy + dz * dz);
    ei -= S.Zeff[a] / d;
    eid -= S.Zeff[a] / fmax(d, P.alat);
    if (S.j1_type) {
      const double rs_ = fmax(d, eps);
      const double A = S.j1_A[a], c = S.j1_c[a], aa = S.j1_a;
      double fp;
      if (S.j1_type == 1) {
        const double ex = qexp(-aa * c * rs_);
        fp = -A * (c * 0.5) * ex;
        lJ += A * (aa * c * c * 0.5) * ex - A * c * ex / rs_;
      } else {
        const double den = 1.0 + aa * c * rs_;
        fp = -A / (2.0 * den * den);
        lJ += A * aa * c / (den * den * den) + 2.0 * fp / rs_;
      }
      const double sc = fp / rs_;
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
      const double rs_ = fmax(d, eps), aa = S.j2_a;
      double fp;
      if (S.j2_type == 1) {
        const double den = 1.0 + aa * rs_;
        fp = 0.5 / (den * den);
        lJ += -aa / (den * den * den) + 2.0 * fp / rs_;
      } else {
        const double ex = qexp(-aa * rs_);
        fp = 0.5 * ex;
        lJ += -(aa * 0.5) * ex + 2.0 * fp / rs_;
      }
      const double sc = fp / rs_;
      gJ[0] = fma(sc, dx, gJ[0]);
      gJ[1] = fma(sc, dy, gJ[1]);
      gJ[2] = fma(sc, dz, gJ[2]);
    }
  }
  const double gx = gJ[0] + gD[0], gy = gJ[1] + gD[1], gz = gJ[2] + gD[2];
  double* o = P.el + ((size_t)e * 8) * nw + w;
  o[0] = -0.5 * (lJ + lD + gx * gx + gy * gy + gz * gz);
  o[(size_t)nw] = ei;
  o[(size_t)2 * nw] = eid;
  o[(size_t)3 * nw] = loc;
  o[(size_t)4 * nw] = ee;
  o[(size_t)7 * nw] = jastrow_single_q(S, pos, e, x, y, z);  // shared by all mesh points of this electron
}

// e_L = sum_e T_e + V_bare + V_ion_ion + V_ecp_local + sum_pts V_nl   (jqmc/hamiltonians.py:225-290); thread = walker
__global__ void kw_reduce_eL(SysDev S, int nw, int n_ecp, const double* __restrict__ el, const double* __restrict__ p,
                             double* __restrict__ e_L, double* __restrict__ T_elem, double* __restrict__ V_parts) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  double T = 0, vbare = S.v_ion_ion, vl = 0, vnl = 0;
  for (int e = 0; e < S.n_e; ++e) {
    const double* o = el + ((size_t)e * 8) * nw + w;
    T += o[0];
    vbare += o[(size_t)nw] + o[(size_t)4 * nw];
    vl += o[(size_t)3 * nw];
    if (T_elem) T_elem[(size_t)w * S.n_e + e] = o[0];
  }
  for (int k = 0; k < n_ecp; ++k) vnl += p[(size_t)k * nw + w];
  e_L[w] = T + (vbare + (vl + vnl));
  if (V_parts) {
    V_parts[(size_t)w * 4 + 0] = vbare;
    V_parts[(size_t)w * 4 + 1] = vl;
    V_parts[(size_t)w * 4 + 2] = vnl;
    V_parts[(size_t)w * 4 + 3] = 0.0;
  }
}

// =================================================================================================
// per-walker matrix inverse: one warp per walker, in-place Gauss-Jordan with partial (row) pivoting in shared memory,
// row swaps undone as column swaps at the end.  Gs[i][j][w] -> Gi[i][j][w] (+ AoS copies, ln|det|, sign)
// (reference: thresholded-SVD pseudo-inverse jqmc/jqmc_mcmc.py:4248-4261 -- identical for a non-singular G).  A pivot
// that vanishes relative to the largest element of G (the reference's rcond, 1e-20) marks a null direction: its row and
// column of the result are set to zero instead of dividing by it, so a singular G (e.g. two same-spin electrons on top of
// each other after branching) gives a finite generalised inverse and ln|det| = -inf, never inf / NaN in the running inverse.
// =================================================================================================
__global__ void kw_inverse(int N, int nw, const double* __restrict__ Gs, double* __restrict__ Gi, double* __restrict__ G_aos,
                           double* __restrict__ Ginv_aos, double* __restrict__ lndet_out, double* __restrict__ sign_out) {
  extern __shared__ double sm_inv[];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int w = blockIdx.x * wpb + wl;
  const int ld = N + 1;
  double* A = sm_inv + (size_t)wl * (N * ld + N);
  int* piv = (int*)(A + N * ld);
  if (w >= nw) return;  // whole warp
  double gmax = 0.0;
  for (int idx = lane; idx < N * N; idx += 32) {
    const int i = idx / N, j = idx % N;
    const double v = Gs[(size_t)idx * nw + w];
    A[i * ld + j] = v;
    gmax = fmax(gmax, fabs(v));
    if (G_aos) G_aos[(size_t)w * N * N + idx] = v;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
  const double tiny = 1.0e-20 * gmax;
  __syncwarp();
  double lndet = 0.0, sgn = 1.0;
  for (int c = 0; c < N; ++c) {
    // pivot: first row of maximal |A[r][c]|, r >= c
    double best = -1.0;
    int bi = c;
    for (int r = c + lane; r < N; r += 32) {
      const double v = fabs(A[r * ld + c]);
      if (v > best) {
        best = v;
        bi = r;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    if (lane == 0) piv[c] = bi;
    if (bi != c) {
      for (int j = lane; j < N; j += 32) {
        const double tmp = A[c * ld + j];
        A[c * ld + j] = A[bi * ld + j];
        A[bi * ld + j] = tmp;
      }
      sgn = -sgn;
    }
    __syncwarp();
    const double d = A[c * ld + c];
    lndet += log(fabs(d));
    if (d < 0) sgn = -sgn;
    const double inv = fabs(d) > tiny ? 1.0 / d : 0.0;  // null direction: zero row (and, below, column) of the result
    __syncwarp();
    if (lane == 0) A[c * ld + c] = 1.0;
    __syncwarp();
    for (int j = lane; j < N; j += 32) A[c * ld + j] *= inv;
    __syncwarp();
    for (int r = 0; r < N; ++r) {
      if (r == c) continue;
      const double f = A[r * ld + c];
      __syncwarp();
      if (lane == 0) A[r * ld + c] = 0.0;
      __syncwarp();
      for (int j = lane; j < N; j += 32) A[r * ld + j] = fma(-f, A[c * ld + j], A[r * ld + j]);
      __syncwarp();
    }
  }
  for (int c = N - 1; c >= 0; --c) {
    const int p = piv[c];
    if (p != c)
      for (int i = lane; i < N; i += 32) {
        const double tmp = A[i * ld + c];
        A[i * ld + c] = A[i * ld + p];
        A[i * ld + p] = tmp;
      }
    __syncwarp();
  }
  for (int idx = lane; idx < N * N; idx += 32) {
    const int i = idx / N, j = idx % N;
    const double v = A[i * ld + j];
    if (Gi) Gi[(size_t)idx * nw + w] = v;
    if (Ginv_aos) Ginv_aos[(size_t)w * N * N + idx] = v;
  }
  if (lane == 0) {
    if (lndet_out) lndet_out[w] = lndet;
    if (sign_out) sign_out[w] = sgn;
  }
}

// ln|Psi| = J + ln|det G| (jqmc/wavefunction.py:677-720): adds J1 + J2 + J3 to lndet; thread = walker
__global__ void kw_add_jastrow(SysDev S, int nw, int has_j3, int nj, const double* __restrict__ rs, const double* __restrict__ j1v,
                               const double* __restrict__ Chi, const double* __restrict__ U, double* __restrict__ lnpsi) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  PosSoA pos{rs, nw, w};
  const int Ne = S.n_e;
  double J = 0.0;
  for (int e = 0; e < Ne; ++e) {
    double x, y, z;
    pos.get(e, x, y, z);
    if (S.j1_type)
      for (int a = 0; a < S.n_atom; ++a) {
        const double dx = x - S.Rn[3 * a], dy = y - S.Rn[3 * a + 1], dz = z - S.Rn[3 * a + 2];
        J += j1_f(S.j1_type, S.j1_a, S.j1_A[a], S.j1_c[a], sqrt(dx * dx + dy * dy + dz * dz));
      }
    if (S.j2_type)
      for (int j = e + 1; j < Ne; ++j) {
        double x2, y2, z2;
        pos.get(j, x2, y2, z2);
        J += j2_f(S.j2_type, S.j2_a, sqrt((x - x2) * (x - x2) + (y - y2) * (y - y2) + (z - z2) * (z - z2)));
      }
  }
  if (has_j3) {
    // J3 = sum_o [ j1_o sum_e chi_oe + sum_{i<j} chi_oi U_oj ],  U = M chi
    for (int o = 0; o < nj; ++o) {
      const double* c = Chi + (size_t)o * Ne * nw + w;
      const double* u = U + (size_t)o * Ne * nw + w;
      double suf = 0.0, acc = 0.0, tot = 0.0;
      for (int e = Ne - 1; e >= 0; --e) {
        const double ce = c[(size_t)e * nw];
        acc = fma(ce, suf, acc);
        suf += u[(size_t)e * nw];
        tot += ce;
      }
      J += acc + j1v[o] * tot;
    }
  }
  lnpsi[w] += J;
}

// =================================================================================================
// LRDMC: assembly and move selection (jqmc/jqmc_gfmc.py:4829-5062).  thread = walker; same arithmetic order as the
// fused kernel of qe_walker.cu (sequential sums in the reference's vector order [kinetic mesh, ECP mesh]).
// =================================================================================================
struct SelectArgs {
  int nw, n_kin, n_ecp, dlt, mode, it;  // mode 0: projection, 1: V elements only
  double alat, E_scf;
  const double* rs;
  const double* RT;  // [9][w]
  double* p;
  double* sj;
  double* el;
  double* wL;       // [w] running weights (mode 0)
  const double* ru; // [it][w]
  double* V_diag;
  double* V_nondiag;
  int* es;          // [w] selected electron
  double* pnew;     // [3][w]
  // mode 3 (GFMC_t, jqmc/jqmc_gfmc.py:724-1110): per-walker remaining time, time draw, projection counter, local energy,
  // move mask and the any-walker-still-running flag
  double* tau_left;
  const double* xi;
  int* pc;
  double* e_L;
  int* active;
  int* any_active;
};
__global__ void __launch_bounds__(256)
kw_lrdmc_select(SysDev S, SelectArgs P) {
  // block = 32 walkers (x) x 8 (y).  Each y owns a contiguous block of electrons (their 6 kinetic elements) and a contiguous
  // block of ECP points: fixed-node split and partial sums in parallel, then the move is located by chunk sums of the
  // normalised probabilities (vector order [kinetic mesh, ECP mesh]) and a sequential scan inside the chunk.
  constexpr int TY = 8;
  __shared__ double s_red[7][TY][32];
  __shared__ double s_chunk[2 * TY][32];
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int nw = P.nw, Ne = S.n_e, n_kin = P.n_kin, n_ecp = P.n_ecp, NPT = n_kin + n_ecp;
  const int wq = blockIdx.x * 32 + lane;
  const bool live = wq < nw;
  const int w = live ? wq : nw - 1;  // dead lanes shadow the last walker (no stores)
  const double a2 = P.alat * P.alat;
#define EL(e, i) P.el[((size_t)(e) * 8 + (i)) * nw + w]
#define PP(k) P.p[(size_t)(k) * nw + w]
  const int EB = (Ne + TY - 1) / TY, e0 = min(Ne, ty * EB), e1 = min(Ne, e0 + EB);
  const int PB = (n_ecp + TY - 1) / TY, q0 = min(n_ecp, ty * PB), q1 = min(n_ecp, q0 + PB);
  double sum_kinFN = 0, SP_kin = 0, sum_opt = 0, ee = 0, loc = 0, sum_eFN = 0, SP_e = 0;
  for (int e = e0; e < e1; ++e) {
    double v6[6];
#pragma unroll
    for (int s6 = 0; s6 < 6; ++s6) v6[s6] = PP(6 * e + s6);
    bool flip = false;
    double nd = 0, kinFN = 0, kinSP = 0;
#pragma unroll
    for (int s6 = 0; s6 < 6; ++s6) {
      const double v = v6[s6];
      flip = flip || (v >= 0.0);
      nd += v + 1.0 / (4.0 * a2);
      const double fn = fmin(v, 0.0);
      kinFN += fn;
      kinSP += fmax(v, 0.0);
      if (live) PP(6 * e + s6) = fn;
    }
    const double zv = EL(e, 1) + EL(e, 0) - nd;
    const double eib = S.ecp_flag ? EL(e, 1) : EL(e, 2);
    sum_opt += flip ? fmax(zv, eib) : zv;
    sum_kinFN += kinFN;
    SP_kin += kinSP;
    ee += EL(e, 4);
    loc += EL(e, 3);
  }
  for (int k = q0; k < q1; ++k) {
    const double v = PP(n_kin + k);
    double fn = fmin(v, 0.0);
    if (P.dlt) fn *= P.sj[(size_t)k * nw + w];
    if (live) PP(n_kin + k) = fn;
    sum_eFN += fn;
    SP_e += fmax(v, 0.0);
  }
  s_red[0][ty][lane] = sum_kinFN;
  s_red[1][ty][lane] = SP_kin;
  s_red[2][ty][lane] = sum_opt;
  s_red[3][ty][lane] = ee;
  s_red[4][ty][lane] = loc;
  s_red[5][ty][lane] = sum_eFN;
  s_red[6][ty][lane] = SP_e;
  __syncthreads();
  double tot7[7];
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    double t = 0;
    for (int y = 0; y < TY; ++y) t += s_red[q][y][lane];
    tot7[q] = t;
  }
  const double diag_kin = 3.0 / (2.0 * a2) * Ne;
  const double disc_bare = tot7[3] + S.v_ion_ion + tot7[2];
  const double nondiag = tot7[0] + tot7[5];
  const double diag = S.ecp_flag ? diag_kin + disc_bare + tot7[4] + tot7[1] + tot7[6] : diag_kin + disc_bare + tot7[1];
  if (ty == 0 && live && P.V_diag) {
    P.V_diag[w] = diag;
    P.V_nondiag[w] = nondiag;
  }
  if (P.mode == 3) {
    if (ty == 0 && live) {
      // time spent in this configuration, weight, remaining time; no move once the time is used up (:1003-1025)
      const double e_L = diag + nondiag;
      double tl = P.tau_left[w];
      if (tl > 0.0) P.pc[w] += 1;
      const double tau_update = fmin(tl, log(1.0 - P.xi[w]) / nondiag);
      P.wL[w] *= qexp(-tau_update * e_L);
      tl -= tau_update;
      P.tau_left[w] = tl;
      P.e_L[w] = e_L;
      const int act = tl <= 0.0 ? 0 : 1;
      P.active[w] = act;
      if (act) atomicOr(P.any_active, 1);
    }
  } else {
    if (P.mode != 0) return;
    if (ty == 0 && live) P.wL[w] *= 1.0 / (diag - P.E_scf) * (-nondiag);
  }
  const double tot = nondiag;  // sum of all fixed-node elements = normalisation of the move probabilities
  {
    double ck = 0, ce = 0;
    for (int k = 6 * e0; k < 6 * e1; ++k) ck += PP(k) / tot;
    for (int k = q0; k < q1; ++k) ce += PP(n_kin + k) / tot;
    s_chunk[ty][lane] = ck;
    s_chunk[TY + ty][lane] = ce;
  }
  __syncthreads();
  if (ty != 0) return;
  const double u = P.ru[(size_t)P.it * nw + w];
  // first chunk whose cumulative probability reaches u, then the first element inside it (searchsorted 'left',
  // jqmc/jqmc_gfmc.py:5057-5062); the scan continues past the chunk if round-off moved the crossing
  double c = 0;
  int kstart = 0;
  for (int ci = 0; ci < 2 * TY; ++ci) {
    const double cs = s_chunk[ci][lane];
    const int y = ci < TY ? ci : ci - TY;
    const int kb = ci < TY ? 6 * min(Ne, y * EB) : n_kin + min(n_ecp, y * PB);
    if (c + cs >= u) {
      kstart = kb;
      break;
    }
    c += cs;
    kstart = ci < TY ? 6 * min(Ne, min(Ne, y * EB) + EB) : n_kin + min(n_ecp, min(n_ecp, y * PB) + PB);
  }
  int ksel = NPT - 1;
  for (int k = kstart; k < NPT; ++k) {
    c += PP(k) / tot;
    if (c >= u) {
      ksel = k;
      break;
    }
  }
  PosSoA pos{P.rs, nw, w};
  double rt[9];
#pragma unroll
  for (int cc = 0; cc < 9; ++cc) rt[cc] = P.RT[(size_t)cc * nw + w];
  int e;
  double x, y, z, px, py, pz, dummy;
  if (ksel < n_kin) {
    e = ksel / 6;
    const int s6 = ksel % 6, ax = s6 >> 1;
    const double sg = (s6 & 1) ? -P.alat : P.alat;
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
  if (live) {
    P.es[w] = e;
    P.pnew[w] = px;
    P.pnew[(size_t)nw + w] = py;
    P.pnew[(size_t)2 * nw + w] = pz;
  }
#undef EL
#undef PP
}

// =================================================================================================
// single-electron move: Sherman-Morrison update of the running inverse and of the cached state.
// block = 32 walkers (x) x TY (y); shared: dvec / vt / col  [N][32].
//   MCMC (decide = 1): acceptance test (jqmc/jqmc_mcmc.py:4402-4519) first; the move is committed only if accepted
//   LRDMC (decide = 0): always committed (jqmc/jqmc_gfmc.py:5083-5141)
// =================================================================================================
struct MoveArgs {
  int nw, no, nj, has_j3, NQ, decide, it;
  double eps_AS;
  // state
  double* rs;
  double* Gs;     // [N][N][w] (MCMC only)
  double* Gi;     // [N][N][w]
  double* Phi;    // [NQ][no][Ne][w]
  double* Mfull;  // [no][N][w]
  double* MupT;   // [no][N][w]
  double* Chi;    // [NQ][nj][Ne][w]
  double* U;
  double* V;      // [nj][Ne][w]
  // the move
  const int* es;        // [w]
  const double* pnew;   // [3][w]
  const double* PhiN_up;  // [NQ][no][w]  orbitals at the new point with the up / dn coefficient tables
  const double* PhiN_dn;
  const double* T1;     // lambda_p   PhiN_dn  [no][w]  (new Mfull column if a down electron moves)
  const double* T2;     // lambda_p^T PhiN_up  [no][w]  (new MupT column if an up electron moves)
  const double* ChiN;   // [NQ][nj][w]
  const double* UN;
  const double* VN;     // [nj][w]
  const double* j1v;
  // MCMC
  const double* Tr;     // [w] proposal ratio
  const double* J12;    // [w] J1+J2 exponent difference
  const double* rb;     // [it][w] uniforms
  double* R_AS;         // [w] current AS factor
  int* acc;
  int* rej;
  const int* active;    // [w] optional move mask (GFMC_t: walkers that are out of time do not move)
};
__global__ void __launch_bounds__(512)
kw_move(SysDev S, MoveArgs P) {
  extern __shared__ double sm_mv[];
  const int lane = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
  const int nw = P.nw;
  const int w = blockIdx.x * 32 + lane;
  const bool live = w < nw;
  const int ww = live ? w : nw - 1;
  const int N = S.n_up, Nd = S.n_dn, Ne = S.n_e, no = P.no;
  double* s_d = sm_mv;               // dvec [N][32]
  double* s_t = s_d + N * 32;        // vt / au [N][32]
  double* s_c = s_t + N * 32;        // old column / row [N][32]
  double* s_red = s_c + N * 32;      // [TY][32] partial sums
  double* s_flag = s_red + TY * 32;  // [4][32]: accept flag, Det, ...
  double* s_p1 = s_flag + 4 * 32;    // [max(N, TY)][32] partial orbital sums of phase 1
  const int es = P.es[ww];
  const bool up = es < N;
  const int k = up ? es : es - N;
  const double* PhiN = up ? P.PhiN_up : P.PhiN_dn;
#define GI(i, j) P.Gi[((size_t)(i) * N + (j)) * nw + ww]
#define GS(i, j) P.Gs[((size_t)(i) * N + (j)) * nw + ww]
  // ---- phase 1: row / column difference ----------------------------------------------------------------
  // With few electrons per spin (N < TY) the orbital sum of every element is split into NP = TY / N contiguous parts so
  // that all warps stream the state; the parts are added in a fixed order.
  const int NP = N < TY ? TY / N : 1, OB = (no + NP - 1) / NP;
  for (int q = ty; q < N * NP; q += TY) {
    const int j = q % N, part_i = q / N;
    const int o0 = min(no, part_i * OB), o1 = min(no, o0 + OB);
    double s0 = 0.0, s1 = 0.0;
    const double* Mx = up ? P.Mfull : P.MupT;
    int o = o0;
#pragma unroll 2
    for (; o + 1 < o1; o += 2) {
      s0 = fma(PhiN[(size_t)o * nw + ww] - P.Phi[((size_t)o * Ne + es) * nw + ww], Mx[((size_t)o * N + j) * nw + ww], s0);
      s1 = fma(PhiN[(size_t)(o + 1) * nw + ww] - P.Phi[((size_t)(o + 1) * Ne + es) * nw + ww], Mx[((size_t)(o + 1) * N + j) * nw + ww], s1);
    }
    if (o < o1) s0 = fma(PhiN[(size_t)o * nw + ww] - P.Phi[((size_t)o * Ne + es) * nw + ww], Mx[((size_t)o * N + j) * nw + ww], s0);
    s_p1[q * 32 + lane] = s0 + s1;
  }
  __syncthreads();
  for (int j = ty; j < N; j += TY) {
    double s = 0.0;
    for (int pi = 0; pi < NP; ++pi) s += s_p1[(pi * N + j) * 32 + lane];
    s_d[j * 32 + lane] = s;
    s_c[j * 32 + lane] = up ? GI(j, k) : GI(k, j);  // old column k (up) / old row k (dn)
  }
  // J3 exponent difference, partial over ty
  double part = 0.0;
  if (P.decide && P.has_j3) {
    for (int o = ty; o < P.nj; o += TY) {
      double g = P.j1v[o];
      for (int i = 0; i < es; ++i) g += P.V[((size_t)o * Ne + i) * nw + ww];
      for (int i = es + 1; i < Ne; ++i) g += P.U[((size_t)o * Ne + i) * nw + ww];
      part = fma(P.ChiN[(size_t)o * nw + ww] - P.Chi[((size_t)o * Ne + es) * nw + ww], g, part);
    }
  }
  s_red[ty * 32 + lane] = part;
  __syncthreads();
  // ---- phase 2: vt[jp] = sum_j dvec[j] Ginv[j][jp]  (up)   |   au[i] = sum_j Ginv[i][j] dvec[j]  (dn) ------------------
  for (int j = ty; j < N; j += TY) {
    double s = 0.0;
    if (up) {
      for (int i = 0; i < N; ++i) s = fma(s_d[i * 32 + lane], GI(i, j), s);
    } else {
      for (int i = 0; i < N; ++i) s = fma(GI(j, i), s_d[i * 32 + lane], s);
    }
    s_t[j * 32 + lane] = s;
  }
  __syncthreads();
  // ---- phase 3: decision (ty == 0) ---------------------------------------------------------------------------------
  if (ty == 0) {
    const double Det = 1.0 + s_t[k * 32 + lane];
    bool ok = true;
    if (P.decide) {
      double dJ = P.J12[ww];
      for (int y = 0; y < TY; ++y) dJ += s_red[y * 32 + lane];
      const double J_ratio = qexp(dJ);
      double R_AS_ratio = 1.0, R_AS_p = P.R_AS[ww];
      if (P.eps_AS > 0.0) {
        const double R_AS_cur = R_AS_p;
        // F = |Ginv'|_F^2 with Ginv' = Ginv - col (x) vt / Det (up)  or  Ginv - au (x) row / Det (dn)
        double F = 0, Smin = 1e300;
        for (int i = 0; i < N; ++i)
          for (int j = 0; j < N; ++j) {
            const double xg = up ? GI(i, j) - s_c[i * 32 + lane] * (s_t[j * 32 + lane] / Det)
                                 : GI(i, j) - (s_t[i * 32 + lane] / Det) * s_c[j * 32 + lane];
            F = fma(xg, xg, F);
          }
        for (int i = 0; i < N; ++i) {
          double r = 0, c = 0;
          for (int j = 0; j < N; ++j) {
            double gij, gji;
            if (up) {
              gij = GS(i, j) + (i == k ? s_d[j * 32 + lane] : 0.0);
              gji = GS(j, i) + (j == k ? s_d[i * 32 + lane] : 0.0);
            } else {
              gij = GS(i, j) + (j == k ? s_d[i * 32 + lane] : 0.0);
              gji = GS(j, i) + (i == k ? s_d[j * 32 + lane] : 0.0);
            }
            r = fma(gij, gij, r);
            c = fma(gji, gji, c);
          }
          Smin = fmin(Smin, fmin(r, c));
        }
        const double SF = Smin * F;
        R_AS_p = SF > 0.0 ? pow(SF, -0.375) : 0.0;
        R_AS_ratio = (fmax(R_AS_p, P.eps_AS) / R_AS_p) / (fmax(R_AS_cur, P.eps_AS) / R_AS_cur);
      }
      const double wr = R_AS_ratio * J_ratio * Det;
      const double xx = wr * wr * P.Tr[ww];
      const double b = P.rb[(size_t)P.it * nw + ww];
      ok = (xx == xx) && (b < fmin(1.0, xx)) && (Det != 0.0);
      if (live) {
        if (ok) {
          P.acc[w] += 1;
          P.R_AS[w] = R_AS_p;
        } else {
          P.rej[w] += 1;
        }
      }
    }
    if (P.active && P.active[ww] == 0) ok = false;
    s_flag[lane] = ok ? 1.0 : 0.0;
    s_flag[32 + lane] = 1.0 / Det;
  }
  __syncthreads();
  if (s_flag[lane] == 0.0 || !live) return;  // no barrier below
  const double invD = s_flag[32 + lane];
  // ---- phase 4: commit ------------------------------------------------------------------------------------------------
  for (int i = ty; i < N; i += TY) {
    if (up) {
      const double ci = s_c[i * 32 + lane];
      for (int j = 0; j < N; ++j) GI(i, j) = GI(i, j) - (ci * s_t[j * 32 + lane]) * invD;
      if (P.Gs) GS(k, i) += s_d[i * 32 + lane];
    } else {
      const double ai = s_t[i * 32 + lane];
      for (int j = 0; j < N; ++j) GI(i, j) = GI(i, j) - (ai * s_c[j * 32 + lane]) * invD;
      if (P.Gs) GS(i, k) += s_d[i * 32 + lane];
    }
  }
#pragma unroll 4  // (independent copies: more loads in flight per thread; the kernel is bound by memory latency)
  for (int o = ty; o < no; o += TY) {
    for (int q = 0; q < P.NQ; ++q) P.Phi[(((size_t)q * no + o) * Ne + es) * nw + w] = PhiN[((size_t)q * no + o) * nw + w];
    if (up) P.MupT[((size_t)o * N + k) * nw + w] = P.T2[(size_t)o * nw + w];
    else P.Mfull[((size_t)o * N + k) * nw + w] = P.T1[(size_t)o * nw + w];
  }
  if (P.has_j3)
    for (int o = ty; o < P.nj; o += TY) {
      for (int q = 0; q < P.NQ; ++q) P.Chi[(((size_t)q * P.nj + o) * Ne + es) * nw + w] = P.ChiN[((size_t)q * P.nj + o) * nw + w];
      P.U[((size_t)o * Ne + es) * nw + w] = P.UN[(size_t)o * nw + w];
      P.V[((size_t)o * Ne + es) * nw + w] = P.VN[(size_t)o * nw + w];
    }
  if (ty == 0) {
    P.rs[((size_t)es * 3 + 0) * nw + w] = P.pnew[w];
    P.rs[((size_t)es * 3 + 1) * nw + w] = P.pnew[(size_t)nw + w];
    P.rs[((size_t)es * 3 + 2) * nw + w] = P.pnew[(size_t)2 * nw + w];
  }
#undef GI
#undef GS
}

// Metropolis proposal (jqmc/jqmc_mcmc.py:4340-4401): thread = walker
__global__ void kw_mc_propose(SysDev S, int nw, int it, double Dt, const double* __restrict__ rs, const int* __restrict__ rsel,
                              const int* __restrict__ raxis, const double* __restrict__ rg, int* __restrict__ es,
                              double* __restrict__ pnew, double* __restrict__ Tr, double* __restrict__ J12) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  const size_t ridx = (size_t)it * nw + w;
  const int ke = rsel[ridx], axis = raxis[ridx];
  PosSoA pos{rs, nw, w};
  double ox, oy, oz;
  pos.get(ke, ox, oy, oz);
  double dist;
  int ia = nearest_atom(S.Rn, S.n_atom, ox, oy, oz, 0, &dist);
  double Zc = S.Zeff[ia];
  const double f_l = 1.0 / (Zc * Zc) * (1.0 + Zc * Zc * dist) / (1.0 + dist);
  const double g = rg[ridx] * (f_l * Dt);
  double nx = ox, ny = oy, nz = oz;
  if (axis == 0) nx = ox + g;
  else if (axis == 1) ny = oy + g;
  else nz = oz + g;
  ia = nearest_atom(S.Rn, S.n_atom, nx, ny, nz, 0, &dist);
  Zc = S.Zeff[ia];
  const double f_p = 1.0 / (Zc * Zc) * (1.0 + Zc * Zc * dist) / (1.0 + dist);
  const double dd = (nx - ox) * (nx - ox) + (ny - oy) * (ny - oy) + (nz - oz) * (nz - oz);
  Tr[w] = (f_l / f_p) * qexp(-dd * (1.0 / (2.0 * f_p * f_p * Dt * Dt) - 1.0 / (2.0 * f_l * f_l * Dt * Dt)));
  J12[w] = jastrow_delta_q(S, pos, ke, ox, oy, oz, nx, ny, nz);
  es[w] = ke;
  pnew[w] = nx;
  pnew[(size_t)nw + w] = ny;
  pnew[(size_t)2 * nw + w] = nz;
}

// AS regularisation factor from SoA G, Ginv (jqmc/determinant.py:1223-1260); thread = walker
__global__ void kw_as_factor(int N, int nw, const double* __restrict__ Gs, const double* __restrict__ Gi, double* __restrict__ R_AS) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  double F = 0, Smin = 1e300;
  for (int i = 0; i < N * N; ++i) F = fma(Gi[(size_t)i * nw + w], Gi[(size_t)i * nw + w], F);
  for (int i = 0; i < N; ++i) {
    double r = 0, c = 0;
    for (int j = 0; j < N; ++j) {
      const double a = Gs[((size_t)i * N + j) * nw + w], b = Gs[((size_t)j * N + i) * nw + w];
      r = fma(a, a, r);
      c = fma(b, b, c);
    }
    Smin = fmin(Smin, fmin(r, c));
  }
  const double SF = Smin * F;
  R_AS[w] = SF > 0.0 ? pow(SF, -0.375) : 0.0;
}
__global__ void kw_fill(long long n, double v, double* __restrict__ x) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) x[t] = v;
}
__global__ void kw_fill_i(long long n, int v, int* __restrict__ x) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) x[t] = v;
}
// RT for mode 0 from the pre-drawn table rRT[(it*9+c)][w] -> RT[c][w] is a pointer offset; for modes 1/2: AoS [w][9] -> SoA
__global__ void kw_rt_identity(int nw, double* __restrict__ RT) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 9LL * nw) return;
  const int c = (int)(t / nw);
  RT[t] = (c % 4 == 0) ? 1.0 : 0.0;
}

// =================================================================================================
// parameter derivatives of ln|Psi| (stochastic reconfiguration O_k; the reference: jax.grad of evaluate_ln_wavefunction_fast,
// jqmc/jqmc_mcmc.py:4748, 854-876; parameter blocks jqmc/wavefunction.py:515-674)
// =================================================================================================
// d(J1 + J2)/d(parameter): thread = walker
__global__ void kw_j12_dparam(SysDev S, int nw, const double* __restrict__ rs, double* __restrict__ d_j1, double* __restrict__ d_j2) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  PosSoA pos{rs, nw, w};
  const int Ne = S.n_e;
  double g1 = 0.0, g2 = 0.0;
  for (int e = 0; e < Ne; ++e) {
    double x, y, z;
    pos.get(e, x, y, z);
    if (S.j1_type && d_j1) {
      const double a = S.j1_a;
      for (int n = 0; n < S.n_atom; ++n) {
        const double dx = x - S.Rn[3 * n], dy = y - S.Rn[3 * n + 1], dz = z - S.Rn[3 * n + 2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz), A = S.j1_A[n], c = S.j1_c[n];
        if (S.j1_type == 1) {
          const double ex = qexp(-a * c * d);
          g1 += -A * (a * c * d * ex - (1.0 - ex)) / (2.0 * a * a);
        } else {
          const double den = 1.0 + a * c * d;
          g1 += A * c * d * d / (2.0 * den * den);
        }
      }
    }
    if (S.j2_type && d_j2) {
      const double a = S.j2_a;
      for (int j = e + 1; j < Ne; ++j) {
        double x2, y2, z2;
        pos.get(j, x2, y2, z2);
        const double d = sqrt((x - x2) * (x - x2) + (y - y2) * (y - y2) + (z - z2) * (z - z2));
        if (S.j2_type == 1) {
          const double den = 1.0 + a * d;
          g2 += -d * d / (2.0 * den * den);
        } else {
          const double ex = qexp(-a * d);
          g2 += (a * d * ex - (1.0 - ex)) / (2.0 * a * a);
        }
      }
    }
  }
  if (d_j1) d_j1[w] = g1;
  if (d_j2) d_j2[w] = g2;
}
// exclusive prefix over electrons P[o][e][w] = sum_{i<e} X[o][i][w] and row sums R[o][w] = sum_e X[o][e][w]; thread = (o, w)
__global__ void kw_prefix_excl(int no, int Ne, int nw, const double* __restrict__ X, double* __restrict__ Pf, double* __restrict__ R) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)no * nw) return;
  const int w = (int)(t % nw), o = (int)(t / nw);
  const double* x = X + (size_t)o * Ne * nw + w;
  double* p = Pf + (size_t)o * Ne * nw + w;
  double s = 0.0;
  for (int e = 0; e < Ne; ++e) {
    p[(size_t)e * nw] = s;
    s += x[(size_t)e * nw];
  }
  R[t] = s;
}
// dst[w][m1(r1)][col0 + m2(r2)] = sc1[r1] sc2[r2] src[r1*s1 + r2*s2 + w]   (m = identity, sc = 1 when the map is null;
// rows mapped to -1 are basis-image holes and are skipped): engine row space -> the reference's parameter layout
__global__ void kw_scatter_mat(int nw, int n1, int n2, const int* __restrict__ map1, const double* __restrict__ sc1,
                               const int* __restrict__ map2, const double* __restrict__ sc2, const double* __restrict__ src, long long s1,
                               long long s2, double* __restrict__ dst, int ld, long long stride_w, int col0) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n1 * n2 * nw) return;
  const int w = (int)(t % nw);
  const int r2 = (int)((t / nw) % n2);
  const int r1 = (int)(t / ((long long)nw * n2));
  const int a = map1 ? map1[r1] : r1, b = map2 ? map2[r2] : r2;
  if (a < 0 || b < 0) return;
  double v = src[r1 * s1 + r2 * s2 + w];
  if (sc1) v *= sc1[r1];
  if (sc2) v *= sc2[r2];
  dst[(long long)w * stride_w + (long long)a * ld + col0 + b] = v;
}

// =================================================================================================
// host side
// =================================================================================================
struct WState {
  int nw = 0, NQ = 1;
  double *rs = nullptr, *Gi = nullptr, *Gs = nullptr;
  double *AO = nullptr, *Phi = nullptr, *Mfull = nullptr, *MupT = nullptr, *Worb = nullptr, *Wrow = nullptr;
  double *AOJ = nullptr, *Chi = nullptr, *U = nullptr, *V = nullptr, *gJ = nullptr, *gJrow = nullptr, *cJ = nullptr;
};


int w_gemm(qe_engine* h, cudaStream_t st, int m, long long n, int k, const double* A, int lda, const double* B, long long ldb,
           double* D, long long ldd, int batch = 1, long long sB = 0, long long sD = 0) {
  if (m <= 0 || n <= 0 || k <= 0) return QE_OK;
  if (h->gemm_ref) {  // qe_set_gemm_reference: plain DFMA GEMM instead of the tensor-core kernel
    LaunchScope ls_(h, K_W_GEMM, st);
    dim3 grid(nblk(n, 128), m, batch);
    kw_dgemm_ref<<<grid, 128, 0, st>>>(m, n, k, A, lda, B, ldb, D, ldd, sB, sD);
    CHECK_LAUNCH();
    return QE_OK;
  }
  const long long ctas = (long long)nblk(n, GN) * ((m + GM - 1) / GM) * batch;
  // few output tiles but a long contraction (one new point per walker against ~1000 AO rows): split k over CTAs so that the
  // grid fills the GPU; partial products go to a scratch buffer and are summed in a fixed order.  Needs sD == m*ldd (dense batches).
  int ksplit = 1;
  if (ctas < 148 && k >= 256 && (batch == 1 || sD == (long long)m * ldd)) ksplit = (int)std::min<long long>((k + 63) / 64, std::max<long long>(1, 592 / ctas));
  if (ksplit <= 1) {
    LaunchScope ls_(h, K_W_GEMM, st);
    dim3 grid(nblk(n, GN), (m + GM - 1) / GM, batch);
    kw_dgemm<<<grid, 128, 0, st>>>(m, n, k, A, lda, B, ldb, D, ldd, sB, sD, 1, k, 0);
    CHECK_LAUNCH();
    return QE_OK;
  }
  const int kchunk = (((k + ksplit - 1) / ksplit) + GK - 1) / GK * GK;
  ksplit = (k + kchunk - 1) / kchunk;
  const long long per = (long long)batch * m * n;
  const size_t need = (size_t)per * ksplit * 8;
  if (need > h->gemm_ws_bytes) {
    if (h->gemm_ws) cudaFree(h->gemm_ws);
    h->gemm_ws = nullptr;
    h->gemm_ws_bytes = 0;
    CUDA_TRY(cudaMalloc((void**)&h->gemm_ws, need));
    h->gemm_ws_bytes = need;
  }
  {
    LaunchScope ls_(h, K_W_GEMM, st);
    dim3 grid(nblk(n, GN), (m + GM - 1) / GM, batch * ksplit);
    kw_dgemm<<<grid, 128, 0, st>>>(m, n, k, A, lda, B, ldb, h->gemm_ws, n, sB, (long long)m * n, ksplit, kchunk, per);
    CHECK_LAUNCH();
  }
  {
    LaunchScope ls_(h, K_W_GEMM, st);
    kw_sum_parts<<<nblk(per, 256), 256, 0, st>>>(per, ksplit, h->gemm_ws, D, 0, ldd, n);
    CHECK_LAUNCH();
  }
  return QE_OK;
}
int w_bmm(qe_engine* h, cudaStream_t st, int ni, int nj, int na, int nw, const double* X, long long sxa, long long sxi, const double* Y,
          long long sya, long long syj, double* D, long long sdi, long long sdj) {
  if (ni <= 0 || nj <= 0) return QE_OK;
  LaunchScope ls_(h, K_W_BMM, st);
  kw_bmm<<<nblk((long long)ni * nj * nw, 128), 128, 0, st>>>(ni, nj, na, nw, X, sxa, sxi, Y, sya, syj, D, sdi, sdj);
  CHECK_LAUNCH();
  return QE_OK;
}
int w_ao(qe_engine* h, cudaStream_t st, const HostBasis& hb, int NQ, int n_item, int nw, const double* pts, double* out) {
  LaunchScope ls_(h, K_W_AO, st);
  const long long total = (long long)hb.n_chunk * n_item * nw;
  const BasisDev& B = hb.dev;
#define CALL(CART, NQ_) kw_ao_store<CART, NQ_><<<nblk(total, 128), 128, 0, st>>>(B, hb.off_cseg, hb.off_cbeg, hb.n_chunk, n_item, nw, pts, out)
  if (B.cart) {
    if (NQ == 1) CALL(true, 1); else CALL(true, 5);
  } else {
    if (NQ == 1) CALL(false, 1); else CALL(false, 5);
  }
#undef CALL
  CHECK_LAUNCH();
  return QE_OK;
}
#define TRY(x)            \
  do {                    \
    const int rc_ = (x);  \
    if (rc_) return rc_;  \
  } while (0)
#define MISC(st, ...)                  \
  do {                                 \
    LaunchScope ls_(h, K_W_MISC, st);  \
    __VA_ARGS__;                       \
  } while (0);                         \
  CHECK_LAUNCH()

// orbital layer from raw AO rows: dst[q][o][item][w] = C^T . src[q][row][item][w]; columns [0, n_up_items) use the up table
int w_orbitals(qe_engine* h, cudaStream_t st, int NQ, int n_item, int n_up_items, int nw, const double* AO, double* Phi) {
  const WideTabs& T = h->wt;
  const long long ld = (long long)n_item * nw;
  if (T.restricted || n_up_items == n_item || n_up_items == 0) {
    const double* A = (n_up_items == 0 && !T.restricted) ? T.CwT_dn : T.CwT_up;
    return w_gemm(h, st, T.no, ld, T.n_row, A, T.n_row, AO, ld, Phi, ld, NQ, (long long)T.n_row * ld, (long long)T.no * ld);
  }
  TRY(w_gemm(h, st, T.no, (long long)n_up_items * nw, T.n_row, T.CwT_up, T.n_row, AO, ld, Phi, ld, NQ, (long long)T.n_row * ld,
             (long long)T.no * ld));
  return w_gemm(h, st, T.no, (long long)(n_item - n_up_items) * nw, T.n_row, T.CwT_dn, T.n_row, AO + (size_t)n_up_items * nw, ld,
                Phi + (size_t)n_up_items * nw, ld, NQ, (long long)T.n_row * ld, (long long)T.no * ld);
}

size_t state_bytes(const qe_engine* h, int nw, int NQ, bool with_G) {
  const WideTabs& T = h->wt;
  const SysDev& S = h->sys;
  const size_t N = S.n_up, Ne = S.n_e, W8 = (size_t)nw * 8;
  size_t n = Ne * 3 * W8 + N * N * W8 * (with_G ? 2 : 1);
  n += (size_t)NQ * T.n_row * Ne * W8;                     // AO
  if (T.has_mo) n += (size_t)NQ * T.no * Ne * W8;          // Phi
  n += 2 * (size_t)T.no * N * W8 + (size_t)T.no * Ne * W8;  // Mfull, MupT, Worb
  if (T.has_mo) n += (size_t)T.n_row * Ne * W8;            // Wrow
  if (T.j3) {
    n += (size_t)NQ * T.nj_row * Ne * W8;
    if (T.j3_mo) n += (size_t)NQ * T.nj * Ne * W8 + (size_t)T.nj_row * Ne * W8;
    n += 3 * (size_t)T.nj * Ne * W8 + Ne * W8;
  }
  return n + 32 * 256;
}
void carve_state(const qe_engine* h, WsCarve& c, int nw, int NQ, bool with_G, WState& X) {
  const WideTabs& T = h->wt;
  const SysDev& S = h->sys;
  const size_t N = S.n_up, Ne = S.n_e;
  X.nw = nw;
  X.NQ = NQ;
  X.rs = c.take<double>(Ne * 3 * nw);
  X.Gi = c.take<double>(N * N * nw);
  X.Gs = with_G ? c.take<double>(N * N * nw) : nullptr;
  X.AO = c.take<double>((size_t)NQ * T.n_row * Ne * nw);
  X.Phi = T.has_mo ? c.take<double>((size_t)NQ * T.no * Ne * nw) : X.AO;
  X.Mfull = c.take<double>((size_t)T.no * N * nw);
  X.MupT = c.take<double>((size_t)T.no * N * nw);
  X.Worb = c.take<double>((size_t)T.no * Ne * nw);
  X.Wrow = T.has_mo ? c.take<double>((size_t)T.n_row * Ne * nw) : X.Worb;
  if (T.j3) {
    X.AOJ = c.take<double>((size_t)NQ * T.nj_row * Ne * nw);
    X.Chi = T.j3_mo ? c.take<double>((size_t)NQ * T.nj * Ne * nw) : X.AOJ;
    X.U = c.take<double>((size_t)T.nj * Ne * nw);
    X.V = c.take<double>((size_t)T.nj * Ne * nw);
    X.gJ = c.take<double>((size_t)T.nj * Ne * nw);
    X.gJrow = T.j3_mo ? c.take<double>((size_t)T.nj_row * Ne * nw) : X.gJ;
    X.cJ = c.take<double>(Ne * nw);
  }
}

// positions -> SoA, orbitals at every electron, Mfull, MupT, J3 caches
int build_state(qe_engine* h, cudaStream_t st, WState& X, const double* r_up, const double* r_dn) {
  const WideTabs& T = h->wt;
  const SysDev& S = h->sys;
  const int nw = X.nw, N = S.n_up, Nd = S.n_dn, Ne = S.n_e, NQ = X.NQ;
  const long long ldE = (long long)Ne * nw, ldN = (long long)N * nw;
  MISC(st, kw_pos_to_soa<<<nblk((long long)Ne * 3 * nw, 256), 256, 0, st>>>(nw, N, Nd, r_up, r_dn, X.rs));
  TRY(w_ao(h, st, h->b_up, NQ, Ne, nw, X.rs, X.AO));
  if (T.has_mo) TRY(w_orbitals(h, st, NQ, Ne, N, nw, X.AO, X.Phi));
  // Mfull = [lambda_p Phi_dn | lambda_u],  MupT = lambda_p^T Phi_up
  TRY(w_gemm(h, st, T.no, (long long)Nd * nw, T.no, T.lamP, T.no, X.Phi + (size_t)N * nw, ldE, X.Mfull, ldN));
  if (N > Nd) {
    MISC(st, kw_fill_unpaired<<<nblk((long long)T.no * (N - Nd) * nw, 256), 256, 0, st>>>(T.no, N, Nd, nw, T.lamU, X.Mfull));
  }
  TRY(w_gemm(h, st, T.no, ldN, T.no, T.lamPT, T.no, X.Phi, ldE, X.MupT, ldN));
  if (T.j3) {
    TRY(w_ao(h, st, h->b_j3, NQ, Ne, nw, X.rs, X.AOJ));
    if (T.j3_mo)
      TRY(w_gemm(h, st, T.nj, ldE, T.nj_row, T.CjT, T.nj_row, X.AOJ, ldE, X.Chi, ldE, NQ, (long long)T.nj_row * ldE, (long long)T.nj * ldE));
    TRY(w_gemm(h, st, T.nj, ldE, T.nj, T.Mj, T.nj, X.Chi, ldE, X.U, ldE));
    TRY(w_gemm(h, st, T.nj, ldE, T.nj, T.MjT, T.nj, X.Chi, ldE, X.V, ldE));
  }
  return QE_OK;
}
// G[i][j][w] = sum_o Phi[o][i] Mfull[o][j]
int build_G(qe_engine* h, cudaStream_t st, WState& X, double* Gs) {
  const SysDev& S = h->sys;
  const long long nw = X.nw, N = S.n_up, Ne = S.n_e;
  return w_bmm(h, st, (int)N, (int)N, h->wt.no, (int)nw, X.Phi, Ne * nw, nw, X.Mfull, N * nw, nw, Gs, N * nw, nw);
}
// ratio weight vectors from the running inverse (and the J3 weight vectors)
int build_weights(qe_engine* h, cudaStream_t st, WState& X) {
  const WideTabs& T = h->wt;
  const SysDev& S = h->sys;
  const long long nw = X.nw, N = S.n_up, Nd = S.n_dn, Ne = S.n_e;
  if (N <= 8) {
    LaunchScope ls_(h, K_W_BMM, st);
    const dim3 grid(nblk(nw, 32), 4), block(32, 8);
    if (N <= 4) kw_worb_small<4><<<grid, block, 0, st>>>(T.no, (int)N, (int)Nd, (int)nw, X.Mfull, X.MupT, X.Gi, X.Worb);
    else kw_worb_small<8><<<grid, block, 0, st>>>(T.no, (int)N, (int)Nd, (int)nw, X.Mfull, X.MupT, X.Gi, X.Worb);
    CHECK_LAUNCH();
  } else {
    TRY(w_bmm(h, st, T.no, (int)N, (int)N, (int)nw, X.Mfull, nw, N * nw, X.Gi, N * nw, nw, X.Worb, Ne * nw, nw));
    TRY(w_bmm(h, st, T.no, (int)Nd, (int)N, (int)nw, X.MupT, nw, N * nw, X.Gi, nw, N * nw, X.Worb + N * nw, Ne * nw, nw));
  }
  if (T.has_mo) {
    if (T.restricted) {
      TRY(w_gemm(h, st, T.n_row, Ne * nw, T.no, T.Cw_up, T.no, X.Worb, Ne * nw, X.Wrow, Ne * nw));
    } else {
      TRY(w_gemm(h, st, T.n_row, N * nw, T.no, T.Cw_up, T.no, X.Worb, Ne * nw, X.Wrow, Ne * nw));
      TRY(w_gemm(h, st, T.n_row, Nd * nw, T.no, T.Cw_dn, T.no, X.Worb + N * nw, Ne * nw, X.Wrow + N * nw, Ne * nw));
    }
  }
  if (T.j3) {
    MISC(st, kw_j3_weights<<<nblk((long long)T.nj * nw, 128), 128, 0, st>>>(T.nj, (int)Ne, (int)nw, T.j1v, X.U, X.V, X.gJ));
    MISC(st, kw_coldot<<<nblk(Ne * nw, 128), 128, 0, st>>>(T.nj, (int)Ne, (int)nw, X.Chi, X.gJ, X.cJ));
    if (T.j3_mo) TRY(w_gemm(h, st, T.nj_row, Ne * nw, T.nj, T.Cj, T.nj, X.gJ, Ne * nw, X.gJrow, Ne * nw));
  }
  return QE_OK;
}
int launch_inverse(qe_engine* h, cudaStream_t st, int nw, const double* Gs, double* Gi, double* G_aos, double* Ginv_aos, double* lndet,
                   double* sign) {
  const int N = h->sys.n_up;
  const size_t per_warp = ((size_t)N * (N + 1) + N) * 8;
  int wpb = (int)std::max<size_t>(1, std::min<size_t>(8, (160 * 1024) / per_warp));
  const size_t smem = per_warp * wpb;
  LaunchScope ls_(h, K_W_INV, st);
  CUDA_TRY(cudaFuncSetAttribute(kw_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kw_inverse<<<nblk(nw, wpb), 32 * wpb, smem, st>>>(N, nw, Gs, Gi, G_aos, Ginv_aos, lndet, sign);
  CHECK_LAUNCH();
  return QE_OK;
}
int launch_mesh(qe_engine* h, cudaStream_t st, const MeshArgs& A) {
  const long long total = (long long)((A.n_kin + A.n_ecp) / 2) * A.nw;
  if (total <= 0) return QE_OK;
  LaunchScope ls_(h, K_W_MESH, st);
  const BasisDev& B = h->b_up.dev;
  const BasisDev BJ = h->wt.j3 ? h->b_j3.dev : B;
  const int j3 = h->wt.j3;
  // angular code up to l = LMAX is compiled in (register allocation follows the largest shell type): l <= 2, <= 4, <= 6
  const int lmax = std::max(B.lmax, j3 ? BJ.lmax : 0);
  const size_t smem = (size_t)B.bytes + (j3 ? (size_t)BJ.bytes : 0);
  const bool use_smem = smem <= 48 * 1024;  // 4 resident CTAs per SM keep their images within the 227 KB
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)std::min<long long>(nblk(total, 128), (long long)sms * 16);
#define CALL4(C1, C2, L, SM) kw_mesh<C1, C2, L, SM><<<grid, 128, SM ? smem : 0, st>>>(B, BJ, j3, h->sys, A)
#define CALL3(C1, C2, L)             \
  do {                               \
    if (use_smem) CALL4(C1, C2, L, true); \
    else CALL4(C1, C2, L, false);    \
  } while (0)
#define CALL(C1, C2)                   \
  do {                                 \
    if (lmax <= 2) CALL3(C1, C2, 2);   \
    else if (lmax <= 4) CALL3(C1, C2, 4); \
    else CALL3(C1, C2, 6);             \
  } while (0)
  if (B.cart) {
    if (BJ.cart) CALL(true, true); else CALL(true, false);
  } else {
    if (BJ.cart) CALL(false, true); else CALL(false, false);
  }
#undef CALL
#undef CALL3
#undef CALL4
  CHECK_LAUNCH();
  return QE_OK;
}
int launch_move(qe_engine* h, cudaStream_t st, const MoveArgs& A) {
  const int N = h->sys.n_up;
  const int TY = 16;
  const size_t smem = ((size_t)3 * N + TY + 4 + std::max(N, TY)) * 32 * 8;
  LaunchScope ls_(h, A.decide ? K_W_DECIDE : K_W_COMMIT, st);
  CUDA_TRY(cudaFuncSetAttribute(kw_move, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kw_move<<<nblk(A.nw, 32), dim3(32, TY), smem, st>>>(h->sys, A);
  CHECK_LAUNCH();
  return QE_OK;
}

// orbitals (and J3 orbitals) with derivatives at one new point per walker + the lambda / j_matrix products of the new column
struct NewPoint {
  double *aoN, *PhiN_up, *PhiN_dn, *T1, *T2, *aoJN, *ChiN, *UN, *VN;
};
size_t newpoint_bytes(const qe_engine* h, int nw, int NQ) {
  const WideTabs& T = h->wt;
  size_t n = (size_t)NQ * T.n_row + (size_t)2 * NQ * T.no + 2 * T.no;
  if (T.j3) n += (size_t)NQ * T.nj_row + (size_t)NQ * T.nj + 2 * T.nj;
  return n * nw * 8 + 16 * 256;
}
void carve_newpoint(const qe_engine* h, WsCarve& c, int nw, int NQ, NewPoint& P) {
  const WideTabs& T = h->wt;
  P.aoN = c.take<double>((size_t)NQ * T.n_row * nw);
  if (T.has_mo) {
    P.PhiN_up = c.take<double>((size_t)NQ * T.no * nw);
    P.PhiN_dn = T.restricted ? P.PhiN_up : c.take<double>((size_t)NQ * T.no * nw);
  } else {
    P.PhiN_up = P.PhiN_dn = P.aoN;
  }
  P.T1 = c.take<double>((size_t)T.no * nw);
  P.T2 = c.take<double>((size_t)T.no * nw);
  P.aoJN = P.ChiN = P.UN = P.VN = nullptr;
  if (T.j3) {
    P.aoJN = c.take<double>((size_t)NQ * T.nj_row * nw);
    P.ChiN = T.j3_mo ? c.take<double>((size_t)NQ * T.nj * nw) : P.aoJN;
    P.UN = c.take<double>((size_t)T.nj * nw);
    P.VN = c.take<double>((size_t)T.nj * nw);
  }
}
int eval_newpoint(qe_engine* h, cudaStream_t st, int nw, int NQ, const double* pnew, NewPoint& P) {
  const WideTabs& T = h->wt;
  TRY(w_ao(h, st, h->b_up, NQ, 1, nw, pnew, P.aoN));
  if (T.has_mo) {
    TRY(w_gemm(h, st, T.no, nw, T.n_row, T.CwT_up, T.n_row, P.aoN, nw, P.PhiN_up, nw, NQ, (long long)T.n_row * nw, (long long)T.no * nw));
    if (!T.restricted)
      TRY(w_gemm(h, st, T.no, nw, T.n_row, T.CwT_dn, T.n_row, P.aoN, nw, P.PhiN_dn, nw, NQ, (long long)T.n_row * nw, (long long)T.no * nw));
  }
  TRY(w_gemm(h, st, T.no, nw, T.no, T.lamP, T.no, P.PhiN_dn, nw, P.T1, nw));
  TRY(w_gemm(h, st, T.no, nw, T.no, T.lamPT, T.no, P.PhiN_up, nw, P.T2, nw));
  if (T.j3) {
    TRY(w_ao(h, st, h->b_j3, NQ, 1, nw, pnew, P.aoJN));
    if (T.j3_mo)
      TRY(w_gemm(h, st, T.nj, nw, T.nj_row, T.CjT, T.nj_row, P.aoJN, nw, P.ChiN, nw, NQ, (long long)T.nj_row * nw, (long long)T.nj * nw));
    TRY(w_gemm(h, st, T.nj, nw, T.nj, T.Mj, T.nj, P.ChiN, nw, P.UN, nw));
    TRY(w_gemm(h, st, T.nj, nw, T.nj, T.MjT, T.nj, P.ChiN, nw, P.VN, nw));
  }
  return QE_OK;
}

}  // namespace

extern "C" int qe_set_gemm_reference(qe_engine* h, int on) {
  if (!h) return fail(QE_ERR_INVALID, "qe_set_gemm_reference: bad argument");
  h->gemm_ref = on != 0;
  return QE_OK;
}

// -------------------------------------------------------------------------------------------------
// _geminal_inv_batched / evaluate_ln_wavefunction
// -------------------------------------------------------------------------------------------------
static int wide_geminal_init_1(qe_engine* h, int nw, const double* r_up, const double* r_dn, double* G, double* Ginv, double* ln_psi,
                               double* sign, cudaStream_t st) {
  const SysDev& S = h->sys;
  const WideTabs& T = h->wt;
  TRY(ensure_ws(h, state_bytes(h, nw, 1, true) + 4096));
  WsCarve c{(char*)h->ws};
  WState X;
  carve_state(h, c, nw, 1, true, X);
  TRY(build_state(h, st, X, r_up, r_dn));
  TRY(build_G(h, st, X, X.Gs));
  TRY(launch_inverse(h, st, nw, X.Gs, nullptr, G, Ginv, ln_psi, sign));
  if (ln_psi) {
    MISC(st, kw_add_jastrow<<<nblk(nw, 32), 32, 0, st>>>(S, nw, T.j3, T.nj, X.rs, T.j1v, X.Chi, X.U, ln_psi));
  }
  return QE_OK;
}

// -------------------------------------------------------------------------------------------------
// compute_local_energy_fast
// -------------------------------------------------------------------------------------------------
static int wide_local_energy_1(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                               double* e_L, double* T_elem, double* V_parts, cudaStream_t st) {
  const SysDev& S = h->sys;
  const WideTabs& T = h->wt;
  const int N = S.n_up, Ne = S.n_e;
  const int n_ecp = S.ecp_flag ? Ne * S.NN * S.Nv : 0;
  TRY(ensure_ws(h, state_bytes(h, nw, 5, false) + ((size_t)std::max(1, n_ecp) + (size_t)Ne * 8 + 9) * nw * 8 + 8192));
  WsCarve c{(char*)h->ws};
  WState X;
  carve_state(h, c, nw, 5, false, X);
  double* p = c.take<double>((size_t)std::max(1, n_ecp) * nw);
  double* el = c.take<double>((size_t)Ne * 8 * nw);
  double* RTs = c.take<double>((size_t)9 * nw);
  TRY(build_state(h, st, X, r_up, r_dn));
  MISC(st, kw_to_soa<<<nblk((long long)N * N * nw, 256), 256, 0, st>>>(nw, N * N, Ginv, X.Gi));
  if (RT) {
    MISC(st, kw_to_soa<<<nblk(9LL * nw, 256), 256, 0, st>>>(nw, 9, RT, RTs));
  } else {
    MISC(st, kw_rt_identity<<<nblk(9LL * nw, 256), 256, 0, st>>>(nw, RTs));
  }
  TRY(build_weights(h, st, X));
  ElecArgs E{nw, T.no, T.nj, T.j3, 1.0, X.rs, X.Phi, X.Worb, X.Chi, X.gJ, el};
  {
    LaunchScope ls_(h, K_W_ELEC, st);
    kw_electron<<<nblk((long long)Ne * nw, 128), 128, 0, st>>>(S, E);
  }
  CHECK_LAUNCH();
  if (n_ecp > 0) {
    MeshArgs M{nw, 0, n_ecp, 0, 0, 1.0, X.rs, RTs, X.Wrow, X.gJrow, X.cJ, el, p, nullptr};
    TRY(launch_mesh(h, st, M));
  }
  MISC(st, kw_reduce_eL<<<nblk(nw, 32), 32, 0, st>>>(S, nw, n_ecp, el, p, e_L, T_elem, V_parts));
  return QE_OK;
}

// -------------------------------------------------------------------------------------------------
// qe_move_ratios
// -------------------------------------------------------------------------------------------------
int wide_move_ratios(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* Ginv, int n_moves,
                     const int32_t* elec_host, const double* r_new, double* det_ratio, double* jas_ratio, cudaStream_t st) {
  const SysDev& S = h->sys;
  const WideTabs& T = h->wt;
  const int N = S.n_up;
  TRY(ensure_ws(h, state_bytes(h, nw, 1, false) + (size_t)n_moves * 4 + 8192));
  WsCarve c{(char*)h->ws};
  WState X;
  carve_state(h, c, nw, 1, false, X);
  int* elec = c.take<int>(n_moves);
  CUDA_TRY(cudaMemcpyAsync(elec, elec_host, (size_t)n_moves * 4, cudaMemcpyHostToDevice, st));
  TRY(build_state(h, st, X, r_up, r_dn));
  MISC(st, kw_to_soa<<<nblk((long long)N * N * nw, 256), 256, 0, st>>>(nw, N * N, Ginv, X.Gi));
  TRY(build_weights(h, st, X));
  const BasisDev& B = h->b_up.dev;
  const BasisDev BJ = T.j3 ? h->b_j3.dev : B;
  const long long total = (long long)n_moves * nw;
  {
    LaunchScope ls_(h, K_W_MESH, st);
#define CALL(C1, C2) \
  kw_move_ratios<C1, C2><<<nblk(total, 128), 128, 0, st>>>(B, BJ, T.j3, S, nw, X.rs, X.Wrow, X.gJrow, X.cJ, n_moves, elec, r_new, det_ratio, jas_ratio)
    if (B.cart) {
      if (BJ.cart) CALL(true, true); else CALL(true, false);
    } else {
      if (BJ.cart) CALL(false, true); else CALL(false, false);
    }
#undef CALL
  }
  CHECK_LAUNCH();
  return QE_OK;
}

// orbital layer (MOs) value/grad/lap at arbitrary points: out[5][n_orb][n_pts]
int wide_eval_orbitals(qe_engine* h, int which, int n_pts, const double* r, double* out, cudaStream_t st) {
  const WideTabs& T = h->wt;
  const HostBasis& hb = which == 2 ? h->b_j3 : h->b_up;
  const int n_row = hb.dev.n_row, n_orb = hb.dev.n_mo;
  const double* CT = which == 2 ? T.CjT : (which == 1 ? T.CwT_dn : T.CwT_up);
  // one "walker" per point: pts[c][pt]
  TRY(ensure_ws(h, ((size_t)3 + (size_t)5 * n_row) * n_pts * 8 + 4096));
  WsCarve c{(char*)h->ws};
  double* pts = c.take<double>((size_t)3 * n_pts);
  double* ao = c.take<double>((size_t)5 * n_row * n_pts);
  MISC(st, kw_to_soa<<<nblk(3LL * n_pts, 256), 256, 0, st>>>(n_pts, 3, r, pts));
  TRY(w_ao(h, st, hb, 5, 1, n_pts, pts, ao));
  return w_gemm(h, st, n_orb, n_pts, n_row, CT, n_row, ao, n_pts, out, n_pts, 5, (long long)n_row * n_pts, (long long)n_orb * n_pts);
}

// -------------------------------------------------------------------------------------------------
// _update_electron_positions: nmpm Metropolis proposals
// -------------------------------------------------------------------------------------------------
static int wide_mcmc_update_1(qe_engine* h, int nw, double* r_up, double* r_dn, uint32_t* keys, double* G, double* Ginv, int nmpm,
                              double Dt, double epsilon_AS, int32_t* acc, int32_t* rej, cudaStream_t st) {
  const SysDev& S = h->sys;
  const WideTabs& T = h->wt;
  const int N = S.n_up, Nd = S.n_dn, Ne = S.n_e;
  TRY(ensure_ws(h, mcmc_draws_bytes(nw, nmpm) + state_bytes(h, nw, 1, true) + newpoint_bytes(h, nw, 1) + (size_t)8 * nw * 8 + 16384));
  WsCarve c{(char*)h->ws};
  int *rsel, *raxis;
  double *rg, *rb;
  TRY(mcmc_draws(h, nw, nmpm, keys, c, &rsel, &raxis, &rg, &rb, st));
  WState X;
  carve_state(h, c, nw, 1, true, X);
  NewPoint P;
  carve_newpoint(h, c, nw, 1, P);
  int* es = c.take<int>(nw);
  double* pnew = c.take<double>((size_t)3 * nw);
  double* Tr = c.take<double>(nw);
  double* J12 = c.take<double>(nw);
  double* R_AS = c.take<double>(nw);
  TRY(build_state(h, st, X, r_up, r_dn));
  MISC(st, kw_to_soa<<<nblk((long long)N * N * nw, 256), 256, 0, st>>>(nw, N * N, Ginv, X.Gi));
  MISC(st, kw_to_soa<<<nblk((long long)N * N * nw, 256), 256, 0, st>>>(nw, N * N, G, X.Gs));
  MISC(st, kw_fill_i<<<nblk(nw, 256), 256, 0, st>>>(nw, 0, acc));
  MISC(st, kw_fill_i<<<nblk(nw, 256), 256, 0, st>>>(nw, 0, rej));
  if (epsilon_AS > 0.0) {
    MISC(st, kw_as_factor<<<nblk(nw, 32), 32, 0, st>>>(N, nw, X.Gs, X.Gi, R_AS));
  } else {
    MISC(st, kw_fill<<<nblk(nw, 256), 256, 0, st>>>(nw, 1.0, R_AS));
  }
  for (int it = 0; it < nmpm; ++it) {
    MISC(st, kw_mc_propose<<<nblk(nw, 32), 32, 0, st>>>(S, nw, it, Dt, X.rs, rsel, raxis, rg, es, pnew, Tr, J12));
    TRY(eval_newpoint(h, st, nw, 1, pnew, P));
    MoveArgs A{};
    A.nw = nw; A.no = T.no; A.nj = T.nj; A.has_j3 = T.j3; A.NQ = 1; A.decide = 1; A.it = it;
    A.eps_AS = epsilon_AS;
    A.rs = X.rs; A.Gs = X.Gs; A.Gi = X.Gi; A.Phi = X.Phi; A.Mfull = X.Mfull; A.MupT = X.MupT; A.Chi = X.Chi; A.U = X.U; A.V = X.V;
    A.es = es; A.pnew = pnew; A.PhiN_up = P.PhiN_up; A.PhiN_dn = P.PhiN_dn; A.T1 = P.T1; A.T2 = P.T2;
    A.ChiN = P.ChiN; A.UN = P.UN; A.VN = P.VN; A.j1v = T.j1v;
    A.Tr = Tr; A.J12 = J12; A.rb = rb; A.R_AS = R_AS; A.acc = acc; A.rej = rej;
    TRY(launch_move(h, st, A));
  }
  MISC(st, kw_pos_to_aos<<<nblk((long long)Ne * 3 * nw, 256), 256, 0, st>>>(nw, N, Nd, X.rs, r_up, r_dn));
  MISC(st, kw_to_aos<<<nblk((long long)N * N * nw, 256), 256, 0, st>>>(nw, N * N, X.Gi, Ginv));
  MISC(st, kw_to_aos<<<nblk((long long)N * N * nw, 256), 256, 0, st>>>(nw, N * N, X.Gs, G));
  return QE_OK;
}

// -------------------------------------------------------------------------------------------------
// GFMC_n._projection_n (mode 0) and _compute_V_elements_n (mode 1)
// -------------------------------------------------------------------------------------------------
// GFMC_t draws of one projection: three splits per walker (rotation, time, move), keys advanced in place; thread = walker
__global__ void kw_tau_draws(int nw, int random_mesh, uint32_t* __restrict__ keys, double* __restrict__ RT, double* __restrict__ xi,
                             double* __restrict__ u) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  Key key{keys[2 * w], keys[2 * w + 1]}, sub;
  rng_split(key, sub);
  double al = 0, be = 0, ga = 0;
  if (random_mesh) {
    const double two_pi = 6.283185307179586;
    al = rng_uniform_bits(rng_bits64(sub, 0u), -two_pi, two_pi);
    be = rng_uniform_bits(rng_bits64(sub, 1u), -two_pi, two_pi);
    ga = rng_uniform_bits(rng_bits64(sub, 2u), -two_pi, two_pi);
  }
  double sa, ca, sb, cb, sg, cg;
  sincos(al, &sa, &ca);
  sincos(be, &sb, &cb);
  sincos(ga, &sg, &cg);
  const double R[9] = {cb * cg, cg * sa * sb - ca * sg, sa * sg + ca * cg * sb, cb * sg, ca * cg + sa * sb * sg,
                       ca * sb * sg - cg * sa, -sb, cb * sa, ca * cb};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) RT[(size_t)(i * 3 + j) * nw + w] = R[j * 3 + i];
  rng_split(key, sub);
  xi[w] = rng_uniform_bits(rng_bits64(sub), 0.0, 1.0);
  rng_split(key, sub);
  u[w] = rng_uniform_bits(rng_bits64(sub), 0.0, 1.0);
  keys[2 * w] = key.a;
  keys[2 * w + 1] = key.b;
}

struct TauArgs {
  double tau;
  int32_t* pc;
  double* e_L;
};
static int wide_lrdmc_impl(qe_engine* h, int mode, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys,
                           double E_scf, int nmpm, int random_mesh, int non_local_move, double alat, const double* RT_in,
                           double* RT_out, double* V_diag, double* V_nondiag, const TauArgs* ta, cudaStream_t st);
// -------------------------------------------------------------------------------------------------
// Walker slices.  The cached state of this family is O(5 n_row N_e) doubles per walker (4.9 MB for the 100 e / 1000 AO
// system), so a call over many walkers (BASELINE configs[4] sweeps to 64k per GPU = 320 GB) is run as consecutive slices
// of walkers through the same workspace: walkers are independent (their draws come from their own keys), the caller's
// arrays are walker-major, so a slice is a pointer offset and the results are identical to one big call.
// -------------------------------------------------------------------------------------------------
extern "C" int qe_set_wide_slice(qe_engine* h, int walkers) {
  if (!h || walkers < 0) return fail(QE_ERR_INVALID, "qe_set_wide_slice: walkers must be >= 0");
  h->wide_slice = walkers;  // 0: automatic (workspace budget), > 0: walkers per slice
  return QE_OK;
}
static int wide_slice(qe_engine* h, int nw, int nmpm) {
  if (h->wide_slice > 0) return std::min(nw, h->wide_slice);
  const size_t per_walker = (state_bytes(h, 1024, 5, true) + 2 * newpoint_bytes(h, 1024, 5) + lrdmc_draws_bytes(1024, nmpm) +
                             mcmc_draws_bytes(1024, nmpm)) / 1024 + (size_t)h->sys.n_e * 8 * 130;
  if (per_walker * (size_t)nw <= ((size_t)1 << 30)) return nw;  // small call: no need to ask the driver
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return nw;
  const size_t budget = std::max<size_t>((size_t)1 << 30, (size_t)(0.45 * (double)(free_b + h->ws_bytes)));
  size_t n = budget / std::max<size_t>(per_walker, 1);
  if (n >= (size_t)nw) return nw;
  n = std::max<size_t>(32, n / 32 * 32);
  return (int)std::min<size_t>(n, (size_t)nw);
}
#define OFF(p, stride) ((p) ? (p) + (size_t)w0 * (stride) : nullptr)
int wide_geminal_init(qe_engine* h, int nw, const double* r_up, const double* r_dn, double* G, double* Ginv, double* ln_psi,
                      double* sign, cudaStream_t st) {
  const size_t N = h->sys.n_up, Nd = h->sys.n_dn;
  const int sl = wide_slice(h, nw, 1);
  for (int w0 = 0; w0 < nw; w0 += sl)
    TRY(wide_geminal_init_1(h, std::min(sl, nw - w0), OFF(r_up, N * 3), OFF(r_dn, Nd * 3), OFF(G, N * N), OFF(Ginv, N * N),
                            OFF(ln_psi, 1), OFF(sign, 1), st));
  return QE_OK;
}
int wide_local_energy(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* RT, const double* Ginv,
                      double* e_L, double* T_elem, double* V_parts, cudaStream_t st) {
  const size_t N = h->sys.n_up, Nd = h->sys.n_dn, Ne = h->sys.n_e;
  const int sl = wide_slice(h, nw, 1);
  for (int w0 = 0; w0 < nw; w0 += sl)
    TRY(wide_local_energy_1(h, std::min(sl, nw - w0), OFF(r_up, N * 3), OFF(r_dn, Nd * 3), OFF(RT, 9), OFF(Ginv, N * N), OFF(e_L, 1),
                            OFF(T_elem, Ne), OFF(V_parts, 4), st));
  return QE_OK;
}
int wide_mcmc_update(qe_engine* h, int nw, double* r_up, double* r_dn, uint32_t* keys, double* G, double* Ginv, int nmpm, double Dt,
                     double epsilon_AS, int32_t* acc, int32_t* rej, cudaStream_t st) {
  const size_t N = h->sys.n_up, Nd = h->sys.n_dn;
  const int sl = wide_slice(h, nw, nmpm);
  for (int w0 = 0; w0 < nw; w0 += sl)
    TRY(wide_mcmc_update_1(h, std::min(sl, nw - w0), OFF(r_up, N * 3), OFF(r_dn, Nd * 3), OFF(keys, 2), OFF(G, N * N), OFF(Ginv, N * N),
                           nmpm, Dt, epsilon_AS, OFF(acc, 1), OFF(rej, 1), st));
  return QE_OK;
}
int wide_lrdmc(qe_engine* h, int mode, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys, double E_scf,
               int nmpm, int random_mesh, int non_local_move, double alat, const double* RT_in, double* RT_out, double* V_diag,
               double* V_nondiag, cudaStream_t st) {
  const size_t N = h->sys.n_up, Nd = h->sys.n_dn;
  const int sl = wide_slice(h, nw, nmpm);
  for (int w0 = 0; w0 < nw; w0 += sl)
    TRY(wide_lrdmc_impl(h, mode, std::min(sl, nw - w0), OFF(w, 1), OFF(r_up, N * 3), OFF(r_dn, Nd * 3), OFF(Ginv, N * N), OFF(keys, 2),
                        E_scf, nmpm, random_mesh, non_local_move, alat, OFF(RT_in, 9), OFF(RT_out, 9), OFF(V_diag, 1), OFF(V_nondiag, 1),
                        nullptr, st));
  return QE_OK;
}
#undef OFF
// GFMC_t projection on the general path: the reference's while_loop literally -- every walker runs every iteration (walkers
// that are out of time make no move) until none has time left (jqmc/jqmc_gfmc.py:1539-1570).  The loop condition is read
// back once per iteration (4 bytes; the iteration itself is ~ms of launches on the systems this path serves).
int wide_lrdmc_tau(qe_engine* h, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys, double tau,
                   int random_mesh, int non_local_move, double alat, int32_t* pc, double* e_L, double* RT_out, cudaStream_t st) {
  TauArgs ta{tau, pc, e_L};
  return wide_lrdmc_impl(h, 3, nw, w, r_up, r_dn, Ginv, keys, 0.0, 1, random_mesh, non_local_move, alat, nullptr, RT_out, nullptr,
                         nullptr, &ta, st);
}
static int wide_lrdmc_impl(qe_engine* h, int mode, int nw, double* w, double* r_up, double* r_dn, double* Ginv, uint32_t* keys,
                           double E_scf, int nmpm, int random_mesh, int non_local_move, double alat, const double* RT_in,
                           double* RT_out, double* V_diag, double* V_nondiag, const TauArgs* ta, cudaStream_t st) {
  const SysDev& S = h->sys;
  const WideTabs& T = h->wt;
  const int N = S.n_up, Nd = S.n_dn, Ne = S.n_e;
  const int n_kin = 6 * Ne, n_ecp = S.ecp_flag ? Ne * S.NN * S.Nv : 0, NPT = n_kin + n_ecp;
  size_t need = state_bytes(h, nw, 5, false) + newpoint_bytes(h, nw, 5) +
                ((size_t)NPT + std::max(1, n_ecp) + (size_t)Ne * 8 + 9 + 8 + 6) * nw * 8 + 16384;
  if (mode == 0) need += lrdmc_draws_bytes(nw, nmpm);
  TRY(ensure_ws(h, need));
  WsCarve c{(char*)h->ws};
  double *rRT = nullptr, *ru = nullptr;
  if (mode == 0) TRY(lrdmc_draws(h, nw, nmpm, random_mesh, keys, c, &rRT, &ru, st));
  WState X;
  carve_state(h, c, nw, 5, false, X);
  NewPoint P;
  carve_newpoint(h, c, nw, 5, P);
  double* p = c.take<double>((size_t)NPT * nw);
  double* sj = c.take<double>((size_t)std::max(1, n_ecp) * nw);
  double* el = c.take<double>((size_t)Ne * 8 * nw);
  double* RTs = c.take<double>((size_t)9 * nw);
  int* es = c.take<int>(nw);
  double* pnew = c.take<double>((size_t)3 * nw);
  double *tau_left = nullptr, *xi = nullptr, *u_tau = nullptr;
  int *active = nullptr, *any_active = nullptr;
  if (mode == 3) {
    tau_left = c.take<double>(nw);
    xi = c.take<double>(nw);
    u_tau = c.take<double>(nw);
    active = c.take<int>(nw);
    any_active = c.take<int>(4);
    MISC(st, kw_fill<<<nblk(nw, 256), 256, 0, st>>>(nw, ta->tau, tau_left));
    MISC(st, kw_fill_i<<<nblk(nw, 256), 256, 0, st>>>(nw, 0, ta->pc));
  }
  TRY(build_state(h, st, X, r_up, r_dn));
  MISC(st, kw_to_soa<<<nblk((long long)N * N * nw, 256), 256, 0, st>>>(nw, N * N, Ginv, X.Gi));
  if (mode != 0 && mode != 3) {
    if (RT_in) {
      MISC(st, kw_to_soa<<<nblk(9LL * nw, 256), 256, 0, st>>>(nw, 9, RT_in, RTs));
    } else {
      MISC(st, kw_rt_identity<<<nblk(9LL * nw, 256), 256, 0, st>>>(nw, RTs));
    }
  }
  const int n_it = mode == 0 ? nmpm : 1;
  const double* RTcur = RTs;
  for (int it = 0; mode == 3 || it < n_it; ++it) {
    if (mode == 0) RTcur = rRT + (size_t)it * 9 * nw;  // rRT[(it*9+c)][w]
    if (mode == 3) {
      MISC(st, kw_tau_draws<<<nblk(nw, 128), 128, 0, st>>>(nw, random_mesh, keys, RTs, xi, u_tau));
      CUDA_TRY(cudaMemsetAsync(any_active, 0, sizeof(int), st));
    }
    TRY(build_weights(h, st, X));
    ElecArgs E{nw, T.no, T.nj, T.j3, alat, X.rs, X.Phi, X.Worb, X.Chi, X.gJ, el};
    {
      LaunchScope ls_(h, K_W_ELEC, st);
      kw_electron<<<nblk((long long)Ne * nw, 128), 128, 0, st>>>(S, E);
    }
    CHECK_LAUNCH();
    MeshArgs M{nw, n_kin, n_ecp, non_local_move, 0, alat, X.rs, RTcur, X.Wrow, X.gJrow, X.cJ, el, p, sj};
    TRY(launch_mesh(h, st, M));
    SelectArgs Q{nw, n_kin, n_ecp, non_local_move, mode, mode == 3 ? 0 : it, alat, E_scf, X.rs, RTcur, p, sj, el, w,
                 mode == 3 ? u_tau : ru, V_diag, V_nondiag, es, pnew};
    if (mode == 3) {
      Q.tau_left = tau_left;
      Q.xi = xi;
      Q.pc = ta->pc;
      Q.e_L = ta->e_L;
      Q.active = active;
      Q.any_active = any_active;
    }
    {
      LaunchScope ls_(h, K_W_SELECT, st);
      kw_lrdmc_select<<<nblk(nw, 32), dim3(32, 8), 0, st>>>(S, Q);
    }
    CHECK_LAUNCH();
    if (mode == 3) {
      int any = 0;
      CUDA_TRY(cudaMemcpyAsync(&any, any_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      if (!any) break;
    } else if (mode != 0) break;
    TRY(eval_newpoint(h, st, nw, 5, pnew, P));
    MoveArgs A{};
    A.nw = nw; A.no = T.no; A.nj = T.nj; A.has_j3 = T.j3; A.NQ = 5; A.decide = 0; A.it = it;
    A.rs = X.rs; A.Gs = nullptr; A.Gi = X.Gi; A.Phi = X.Phi; A.Mfull = X.Mfull; A.MupT = X.MupT; A.Chi = X.Chi; A.U = X.U; A.V = X.V;
    A.es = es; A.pnew = pnew; A.PhiN_up = P.PhiN_up; A.PhiN_dn = P.PhiN_dn; A.T1 = P.T1; A.T2 = P.T2;
    A.ChiN = P.ChiN; A.UN = P.UN; A.VN = P.VN; A.j1v = T.j1v;
    A.active = mode == 3 ? active : nullptr;
    TRY(launch_move(h, st, A));
  }
  if (mode == 0 || mode == 3) {
    MISC(st, kw_pos_to_aos<<<nblk((long long)Ne * 3 * nw, 256), 256, 0, st>>>(nw, N, Nd, X.rs, r_up, r_dn));
    MISC(st, kw_to_aos<<<nblk((long long)N * N * nw, 256), 256, 0, st>>>(nw, N * N, X.Gi, Ginv));
    MISC(st, kw_to_aos<<<nblk(9LL * nw, 256), 256, 0, st>>>(nw, 9, RTcur, RT_out));
  }
  return QE_OK;
}

// -------------------------------------------------------------------------------------------------
// vmap(grad(evaluate_ln_wavefunction_fast)) w.r.t. the variational parameters
// -------------------------------------------------------------------------------------------------
int wide_dln_wf(qe_engine* h, int nw, const double* r_up, const double* r_dn, const double* Ginv, double* d_j1, double* d_j2,
                double* d_j3, double* d_lambda, cudaStream_t st) {
  const SysDev& S = h->sys;
  const WideTabs& T = h->wt;
  const long long N = S.n_up, Nd = S.n_dn, Ne = S.n_e, W = nw;
  const int no = T.no, nj = T.nj;
  size_t need = state_bytes(h, nw, 1, false) + ((size_t)no * N + (size_t)no * no) * W * 8 + 8192;
  if (T.j3) need += ((size_t)nj * Ne + nj + (size_t)nj * nj) * W * 8 + 4096;
  TRY(ensure_ws(h, need));
  WsCarve c{(char*)h->ws};
  WState X;
  carve_state(h, c, nw, 1, false, X);
  double* A = c.take<double>((size_t)no * N * W);
  double* DL = c.take<double>((size_t)no * no * W);
  TRY(build_state(h, st, X, r_up, r_dn));
  if (d_j1 || d_j2) {
    MISC(st, kw_j12_dparam<<<nblk(nw, 32), 32, 0, st>>>(S, nw, X.rs, S.j1_type ? d_j1 : nullptr, S.j2_type ? d_j2 : nullptr));
  }
  const BasisDev& B = h->b_up.dev;
  const int* rmap = T.has_mo ? nullptr : (const int*)(B.g + B.off_rowao);
  const double* rsc = T.has_mo ? nullptr : (const double*)(B.g + B.off_rowscale);
  if (d_lambda) {
    const int n_ref = T.has_mo ? no : B.n_ao, ld = n_ref + (int)(N - Nd);
    MISC(st, kw_to_soa<<<nblk(N * N * W, 256), 256, 0, st>>>(nw, (int)(N * N), Ginv, X.Gi));
    // A[a][j] = sum_i Phi_up[a][i] Ginv[j][i];   DL[a][b] = sum_{j < Nd} A[a][j] Phi_dn[b][j]
    TRY(w_bmm(h, st, no, (int)N, (int)N, nw, X.Phi, W, Ne * W, X.Gi, W, N * W, A, N * W, W));
    TRY(w_bmm(h, st, no, no, (int)Nd, nw, A, W, N * W, X.Phi + N * W, W, Ne * W, DL, (long long)no * W, W));
    CUDA_TRY(cudaMemsetAsync(d_lambda, 0, (size_t)nw * n_ref * ld * 8, st));
    MISC(st, kw_scatter_mat<<<nblk((long long)no * no * W, 256), 256, 0, st>>>(nw, no, no, rmap, rsc, rmap, rsc, DL, (long long)no * W, W,
                                                                              d_lambda, ld, (long long)n_ref * ld, 0));
    if (N > Nd) {
      MISC(st, kw_scatter_mat<<<nblk((long long)no * (N - Nd) * W, 256), 256, 0, st>>>(nw, no, (int)(N - Nd), rmap, rsc, nullptr, nullptr,
                                                                                      A + Nd * W, N * W, W, d_lambda, ld,
                                                                                      (long long)n_ref * ld, n_ref));
    }
  }
  if (d_j3 && T.j3) {
    const BasisDev& BJ = h->b_j3.dev;
    const int* jmap = T.j3_mo ? nullptr : (const int*)(BJ.g + BJ.off_rowao);
    const double* jsc = T.j3_mo ? nullptr : (const double*)(BJ.g + BJ.off_rowscale);
    const int n_ref = T.j3_mo ? nj : BJ.n_ao, ld = n_ref + 1;
    double* Pf = c.take<double>((size_t)nj * Ne * W);
    double* Rs = c.take<double>((size_t)nj * W);
    double* DM = c.take<double>((size_t)nj * nj * W);
    MISC(st, kw_prefix_excl<<<nblk((long long)nj * W, 128), 128, 0, st>>>(nj, (int)Ne, nw, X.Chi, Pf, Rs));
    // DM[a][b] = sum_j (sum_{i<j} chi[a][i]) chi[b][j]
    TRY(w_bmm(h, st, nj, nj, (int)Ne, nw, Pf, W, Ne * W, X.Chi, W, Ne * W, DM, (long long)nj * W, W));
    CUDA_TRY(cudaMemsetAsync(d_j3, 0, (size_t)nw * n_ref * ld * 8, st));
    MISC(st, kw_scatter_mat<<<nblk((long long)nj * nj * W, 256), 256, 0, st>>>(nw, nj, nj, jmap, jsc, jmap, jsc, DM, (long long)nj * W, W, d_j3,
                                                                              ld, (long long)n_ref * ld, 0));
    MISC(st, kw_scatter_mat<<<nblk((long long)nj * W, 256), 256, 0, st>>>(nw, nj, 1, jmap, jsc, nullptr, nullptr, Rs, W, 0, d_j3, ld,
                                                                         (long long)n_ref * ld, n_ref));
  }
  return QE_OK;
}
