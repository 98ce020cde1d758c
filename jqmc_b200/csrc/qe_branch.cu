// LRDMC (GFMC_n) per-step statistics and walker reconfiguration on the device
// (jqmc/jqmc_gfmc.py:5955-5976 weighted sums, :6059-6321 branching).
//
// The reference does this part on the host with NumPy + mpi4py: allreduce(sum w) -> local cumulative
// probabilities -> Exscan rank offsets -> Allgather -> searchsorted with the comb z_i = (i + zeta)/N -> negotiated
// point-to-point migration of the chosen walkers.  Here every rank holds the all-gathered weight vector
// (NCCL all_gather of nw doubles per rank, done by the caller) and evaluates the SAME comb redundantly, so the only
// data-path collectives are that all_gather and the one of the walker coordinates; each rank then gathers its
// `nw` destination walkers from the gathered coordinate buffer by index.
//
// Floating-point order follows the reference so that the chosen indices agree bit for bit:
//   S_r   = np.sum(w_r)                 NumPy pairwise summation (8 accumulators, blocks of 128, recursive halving)
//   S     = S_0 + S_1 + ...             rank order
//   p     = w / S ;  c_r = np.cumsum(p_r)   sequential ;  off_r = sum_{q<r} np.sum(p_q)  (pairwise, then rank order)
//   idx_i = searchsorted(c + off, (i + zeta)/N, 'left')
#include "qe_common.cuh"

namespace {

// NumPy's DOUBLE_pairwise_sum for a contiguous array (numpy/core/src/umath/loops_utils.h.src)
__device__ double np_pairwise_sum(const double* __restrict__ a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
      r0 += a[i];
      r1 += a[i + 1];
      r2 += a[i + 2];
      r3 += a[i + 3];
      r4 += a[i + 4];
      r5 += a[i + 5];
      r6 += a[i + 6];
      r7 += a[i + 7];
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// The same sum with the leaves (contiguous runs of <= 128 elements, summed with NumPy's 8 accumulators) computed by different
// threads and combined by thread 0 in NumPy's recursion order: bit-identical to np_pairwise_sum, ~30x shorter critical path.
constexpr int MAX_LEAVES = 1024;
__device__ void np_enum_leaves(int off, int n, int* lo, int* ln, int& cnt) {
  if (n <= 128) {
    lo[cnt] = off;
    ln[cnt] = n;
    ++cnt;
    return;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  np_enum_leaves(off, n2, lo, ln, cnt);
  np_enum_leaves(off + n2, n - n2, lo, ln, cnt);
}
__device__ double np_combine(int n, const double* leaf, int& idx) {
  if (n <= 128) return leaf[idx++];
  int n2 = n / 2;
  n2 -= n2 % 8;
  const double a = np_combine(n2, leaf, idx);
  const double b = np_combine(n - n2, leaf, idx);
  return a + b;
}
// all threads of the block call this; the result is valid in thread 0.  `a` may be global or shared memory.
__device__ double block_np_sum(const double* __restrict__ a, int n, int* s_lo, int* s_ln, double* s_leaf, int* s_cnt) {
  if (n > 128 * MAX_LEAVES / 2) return threadIdx.x == 0 ? np_pairwise_sum(a, n) : 0.0;  // very long vectors: serial
  if (threadIdx.x == 0) {
    int cnt = 0;
    np_enum_leaves(0, n, s_lo, s_ln, cnt);
    *s_cnt = cnt;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < *s_cnt; l += blockDim.x) s_leaf[l] = np_pairwise_sum(a + s_lo[l], s_ln[l]);
  __syncthreads();
  double res = 0.0;
  if (threadIdx.x == 0) {
    int idx = 0;
    res = np_combine(n, s_leaf, idx);
  }
  __syncthreads();
  return res;
}

// block r: S[r] = np.sum(w_all[r*nw : (r+1)*nw])
// (w_stride: distance between the weight vectors of consecutive ranks -- nw for a dense all-gathered vector, the packed
//  record length for qe_lrdmc_reconfigure_packed)
__global__ void k_branch_rank_sums(int nw, size_t w_stride, const double* __restrict__ w_all, double* __restrict__ S) {
  __shared__ int s_lo[MAX_LEAVES], s_ln[MAX_LEAVES], s_cnt;
  __shared__ double s_leaf[MAX_LEAVES];
  const double v = block_np_sum(w_all + (size_t)blockIdx.x * w_stride, nw, s_lo, s_ln, s_leaf, &s_cnt);
  if (threadIdx.x == 0) S[blockIdx.x] = v;
}

// block r: p = w / sum_r S[r];  P[r] = np.sum(p_r);  c_r = cumsum(p_r) (sequential fp64, staged through shared memory)
__global__ void k_branch_cumprob(int nw, int world, size_t w_stride, const double* __restrict__ w_all, const double* __restrict__ S,
                                 double* __restrict__ c, double* __restrict__ P) {
  __shared__ int s_lo[MAX_LEAVES], s_ln[MAX_LEAVES], s_cnt;
  __shared__ double s_leaf[MAX_LEAVES];
  constexpr int CH = 2048;
  __shared__ double s_buf[CH];
  __shared__ double s_acc;
  const int r = blockIdx.x;
  double gsum = 0.0;
  for (int q = 0; q < world; ++q) gsum += S[q];
  double* cr = c + (size_t)r * nw;
  const double* wr = w_all + (size_t)r * w_stride;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) cr[i] = wr[i] / gsum;
  __syncthreads();
  const double pr = block_np_sum(cr, nw, s_lo, s_ln, s_leaf, &s_cnt);
  if (threadIdx.x == 0) {
    P[r] = pr;
    s_acc = 0.0;
  }
  __syncthreads();
  for (int base = 0; base < nw; base += CH) {
    const int m = min(CH, nw - base);
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_buf[i] = cr[base + i];
    __syncthreads();
    if (threadIdx.x == 0) {
      // the sequential fp64 chain of np.cumsum (one dependent add per element, 8 cycles each); loads and stores of 16
      // elements are batched around it so that only the adds sit on the critical path
      double acc = s_acc;
      int i = 0;
      for (; i + 16 <= m; i += 16) {
        double v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = s_buf[i + k];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          acc += v[k];
          v[k] = acc;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) s_buf[i + k] = v[k];
      }
      for (; i < m; ++i) {
        acc += s_buf[i];
        s_buf[i] = acc;
      }
      s_acc = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) cr[base + i] = s_buf[i];
    __syncthreads();
  }
}

// thread g (global destination slot): chosen[g] = searchsorted(c + offset, (g + zeta)/N, 'left'), clamped to N-1
__global__ void k_branch_select(int nw, int world, const double* __restrict__ c, const double* __restrict__ P, double zeta,
                                int* __restrict__ chosen) {
  const int N = nw * world;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ double off[64];
  if (threadIdx.x == 0) {
    // MPI.Exscan(SUM): off_r = P_0 + ... + P_{r-1}; rank 0 uses 0.0 (jqmc/jqmc_gfmc.py:6088-6095)
    double acc = 0.0;
    for (int q = 0; q < world; ++q) {
      off[q] = q == 0 ? 0.0 : acc;
      acc += P[q];
    }
  }
  __syncthreads();
  if (g >= N) return;  // after the barrier: every thread of the last block takes part in it
  const double z = ((double)g + zeta) / (double)N;
  int lo = 0, hi = N;  // first i with cg[i] >= z
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const double v = c[mid] + off[mid / nw];
    if (v < z) lo = mid + 1; else hi = mid;
  }
  chosen[g] = lo < N ? lo : N - 1;
}

// number of distinct chosen sources (chosen is non-decreasing): survivors (jqmc/jqmc_gfmc.py:6134-6135)
__global__ void k_branch_count(int N, const int* __restrict__ chosen, int* __restrict__ n_survived) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  const int flag = (g == 0) || (chosen[g] != chosen[g - 1]);
  const unsigned m = __ballot_sync(__activemask(), flag);
  if ((threadIdx.x & 31) == 0) atomicAdd(n_survived, __popc(m));
}

// dst walker i of this rank <- gathered walker chosen[rank*nw + i]
__global__ void k_gather_walkers(int nw, int per_up, int per_dn, const int* __restrict__ chosen_local,
                                 const double* __restrict__ src_up, const double* __restrict__ src_dn,
                                 double* __restrict__ dst_up, double* __restrict__ dst_dn) {
  const int per = per_up + per_dn;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nw * per) return;
  const int i = (int)(t / per), k = (int)(t % per);
  const size_t s = (size_t)chosen_local[i];
  if (k < per_up) dst_up[(size_t)i * per_up + k] = src_up[s * per_up + k];
  else dst_dn[(size_t)i * per_dn + (k - per_up)] = src_dn[s * per_dn + (k - per_up)];
}

// per-step weighted sums (jqmc/jqmc_gfmc.py:5971-5976): out = {nw, sum w, sum w/(Vd-E), sum w/(Vd-E) e, sum w/(Vd-E) e^2}
// GFMC_t form when Vd == nullptr (jqmc/jqmc_gfmc.py:1929-1932): Vn holds e_L, q = w (no division by V_diag - E_scf)
// one block, fixed-order tree reduction (deterministic)
__global__ void k_lrdmc_collect(int nw, const double* __restrict__ w, const double* __restrict__ Vd,
                                const double* __restrict__ Vn, double E_scf, double* __restrict__ out) {
  __shared__ double sm[4][256];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int i = threadIdx.x; i < nw; i += 256) {
    const double wi = w[i];
    const double e = Vd ? Vd[i] + Vn[i] : Vn[i], q = Vd ? wi / (Vd[i] - E_scf) : wi;
    a0 += wi;
    a1 += q;
    a2 += q * e;
    a3 += q * e * e;
  }
  sm[0][threadIdx.x] = a0;
  sm[1][threadIdx.x] = a1;
  sm[2][threadIdx.x] = a2;
  sm[3][threadIdx.x] = a3;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 4; ++k) sm[k][threadIdx.x] += sm[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = (double)nw;
    for (int k = 0; k < 4; ++k) out[1 + k] = sm[k][0];
  }
}

}  // namespace

extern "C" int qe_lrdmc_collect(qe_engine* h, int nw, const double* w, const double* V_diag, const double* V_nondiag, double E_scf,
                                double* out5, void* stream) {
  if (!h || nw <= 0 || !w || !V_nondiag || !out5) return fail(QE_ERR_INVALID, "qe_lrdmc_collect: bad argument");
  {
    LaunchScope ls_(h, K_COLLECT, (cudaStream_t)stream);
    k_lrdmc_collect<<<1, 256, 0, (cudaStream_t)stream>>>(nw, w, V_diag, V_nondiag, E_scf, out5);
  }
  CHECK_LAUNCH();
  return QE_OK;
}

static int branch_impl(qe_engine* h, int nw, int world, const double* w_all, size_t w_stride, double zeta, int32_t* chosen_all,
                       int32_t* n_survived, void* stream) {
  if (!h || nw <= 0 || world <= 0 || world > 64 || !w_all || !chosen_all || !n_survived)
    return fail(QE_ERR_INVALID, "qe_lrdmc_branch: bad argument (world must be 1..64)");
  if (!(zeta >= 0.0 && zeta < 1.0)) return fail(QE_ERR_INVALID, "qe_lrdmc_branch: zeta must be in [0, 1)");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = nw * world;
  // private scratch (not the shared workspace: the walker buffers of the caller may live there in future)
  if (h->branch_ws_n < (size_t)N + 2 * 64) {
    if (h->branch_ws) cudaFree(h->branch_ws);
    h->branch_ws = nullptr;
    CUDA_TRY(cudaMalloc(&h->branch_ws, ((size_t)N + 2 * 64) * sizeof(double)));
    h->branch_ws_n = (size_t)N + 2 * 64;
  }
  double* S = h->branch_ws;
  double* P = S + 64;
  double* c = P + 64;
  CUDA_TRY(cudaMemsetAsync(n_survived, 0, sizeof(int32_t), st));
  {
    LaunchScope ls_(h, K_BRANCH, st);
    k_branch_rank_sums<<<world, 256, 0, st>>>(nw, w_stride, w_all, S);
    k_branch_cumprob<<<world, 256, 0, st>>>(nw, world, w_stride, w_all, S, c, P);
    k_branch_select<<<nblk(N, 128), 128, 0, st>>>(nw, world, c, P, zeta, chosen_all);
    k_branch_count<<<nblk(N, 128), 128, 0, st>>>(N, chosen_all, n_survived);
    h->launches += 3;
  }
  CHECK_LAUNCH();
  return QE_OK;
}

extern "C" int qe_lrdmc_branch(qe_engine* h, int nw, int world, const double* w_all, double zeta, int32_t* chosen_all,
                               int32_t* n_survived, void* stream) {
  return branch_impl(h, nw, world, w_all, (size_t)nw, zeta, chosen_all, n_survived, stream);
}

// ---- packed exchange: ONE all_gather per branching -----------------------------------------------------------------------
// record of a rank (doubles): [0..4] the five weighted sums of qe_lrdmc_collect, [5..7] padding, then w[nw], r_up[nw*3 n_up],
// r_dn[nw*3 n_dn]
namespace {
__global__ void k_pack_record(int nw, int pu, int pd, const double* __restrict__ sums5, const double* __restrict__ w,
                              const double* __restrict__ r_up, const double* __restrict__ r_dn, double* __restrict__ rec) {
  const long long n = 8 + (long long)nw * (1 + pu + pd);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    double v;
    if (t < 8) v = t < 5 ? sums5[t] : 0.0;
    else if (t < 8 + nw) v = w[t - 8];
    else if (t < 8 + (long long)nw * (1 + pu)) v = r_up[t - 8 - nw];
    else v = r_dn[t - 8 - (long long)nw * (1 + pu)];
    rec[t] = v;
  }
}
// sums over ranks in rank order (the reference: MPI reduce), and this rank's new walkers from the gathered records
__global__ void k_unpack_records(int nw, int world, int pu, int pd, size_t stride, const double* __restrict__ all,
                                 const int* __restrict__ chosen_local, double* __restrict__ sums5, double* __restrict__ dst_up,
                                 double* __restrict__ dst_dn) {
  const int per = pu + pd;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 5 && sums5) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += all[(size_t)r * stride + t];
    sums5[t] = s;
  }
  if (t >= (long long)nw * per) return;
  const int i = (int)(t / per), k = (int)(t % per);
  const int g = chosen_local[i], r = g / nw, j = g % nw;
  const double* rec = all + (size_t)r * stride + 8 + nw;
  if (k < pu) dst_up[(size_t)i * pu + k] = rec[(size_t)j * pu + k];
  else dst_dn[(size_t)i * pd + (k - pu)] = rec[(size_t)nw * pu + (size_t)j * pd + (k - pu)];
}
}  // namespace

extern "C" int64_t qe_lrdmc_record_len(qe_engine* h, int nw) {
  if (!h || nw <= 0) return 0;
  return 8 + (int64_t)nw * (1 + 3 * h->sys.n_up + 3 * h->sys.n_dn);
}

extern "C" int qe_lrdmc_pack(qe_engine* h, int nw, const double* sums5, const double* w, const double* r_up, const double* r_dn,
                             double* record, void* stream) {
  if (!h || nw <= 0 || !sums5 || !w || !r_up || !record || (h->sys.n_dn > 0 && !r_dn))
    return fail(QE_ERR_INVALID, "qe_lrdmc_pack: bad argument");
  const int pu = 3 * h->sys.n_up, pd = 3 * h->sys.n_dn;
  {
    LaunchScope ls_(h, K_GATHER, (cudaStream_t)stream);
    k_pack_record<<<std::min(1024u, nblk(8 + (long long)nw * (1 + pu + pd), 256)), 256, 0, (cudaStream_t)stream>>>(nw, pu, pd, sums5, w, r_up,
                                                                                                                r_dn, record);
  }
  CHECK_LAUNCH();
  return QE_OK;
}

extern "C" int qe_lrdmc_reconfigure_packed(qe_engine* h, int nw, int world, int rank, const double* records, double zeta,
                                           int32_t* chosen_all, int32_t* n_survived, double* sums5, double* dst_r_up,
                                           double* dst_r_dn, void* stream) {
  if (!h || nw <= 0 || world <= 0 || rank < 0 || rank >= world || !records || !chosen_all || !n_survived || !dst_r_up ||
      (h->sys.n_dn > 0 && !dst_r_dn))
    return fail(QE_ERR_INVALID, "qe_lrdmc_reconfigure_packed: bad argument");
  const size_t stride = (size_t)qe_lrdmc_record_len(h, nw);
  int rc = branch_impl(h, nw, world, records + 8, stride, zeta, chosen_all, n_survived, stream);
  if (rc) return rc;
  const int pu = 3 * h->sys.n_up, pd = 3 * h->sys.n_dn;
  {
    LaunchScope ls_(h, K_GATHER, (cudaStream_t)stream);
    k_unpack_records<<<nblk(std::max<long long>(5, (long long)nw * (pu + pd)), 256), 256, 0, (cudaStream_t)stream>>>(
        nw, world, pu, pd, stride, records, chosen_all + (size_t)rank * nw, sums5, dst_r_up, dst_r_dn);
  }
  CHECK_LAUNCH();
  return QE_OK;
}

extern "C" int qe_gather_walkers(qe_engine* h, int nw, const int32_t* chosen_local, const double* src_r_up, const double* src_r_dn,
                                 double* dst_r_up, double* dst_r_dn, void* stream) {
  if (!h || nw <= 0 || !chosen_local || !src_r_up || !dst_r_up || (h->sys.n_dn > 0 && (!src_r_dn || !dst_r_dn)))
    return fail(QE_ERR_INVALID, "qe_gather_walkers: bad argument");
  const int pu = 3 * h->sys.n_up, pd = 3 * h->sys.n_dn;
  {
    LaunchScope ls_(h, K_GATHER, (cudaStream_t)stream);
    k_gather_walkers<<<nblk((long long)nw * (pu + pd), 256), 256, 0, (cudaStream_t)stream>>>(nw, pu, pd, chosen_local, src_r_up,
                                                                                             src_r_dn, dst_r_up, dst_r_dn);
  }
  CHECK_LAUNCH();
  return QE_OK;
}
