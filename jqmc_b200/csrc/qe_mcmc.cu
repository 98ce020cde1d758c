// Metropolis step of the VMC driver: jax.random-compatible draws, rotation matrices and the k_mcmc kernel.
#include "qe_mcmc_kernel.cuh"

// =================================================================================================
// RNG kernels (semantics: oracle/jaxrng.py; call sites jqmc/jqmc_mcmc.py:4322-4367, 4499-4500, 4232-4233)
// =================================================================================================
// rotation matrix RT = R^T from split(key)[1], key not advanced (jqmc/jqmc_mcmc.py:4228-4245)
__device__ __forceinline__ void rotation_RT(double al, double be, double ga, double* __restrict__ o) {
  double sa, ca, sb, cb, sg, cg;
  sincos(al, &sa, &ca);
  sincos(be, &sb, &cb);
  sincos(ga, &sg, &cg);
  // R rows; store transposed
  const double R00 = cb * cg, R01 = cg * sa * sb - ca * sg, R02 = sa * sg + ca * cg * sb;
  const double R10 = cb * sg, R11 = ca * cg + sa * sb * sg, R12 = ca * sb * sg - cg * sa;
  const double R20 = -sb, R21 = cb * sa, R22 = ca * cb;
  o[0] = R00; o[1] = R10; o[2] = R20;
  o[3] = R01; o[4] = R11; o[5] = R21;
  o[6] = R02; o[7] = R12; o[8] = R22;
}
__global__ void k_rotation(int nw, const uint32_t* __restrict__ keys, double* __restrict__ RT) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  Key k{keys[2 * w], keys[2 * w + 1]};
  const Key sub = threefry(k, 0u, 1u);
  const double two_pi = 6.283185307179586;
  const double al = rng_uniform_bits(rng_bits64(sub, 0u), -two_pi, two_pi);
  const double be = rng_uniform_bits(rng_bits64(sub, 1u), -two_pi, two_pi);
  const double ga = rng_uniform_bits(rng_bits64(sub, 2u), -two_pi, two_pi);
  rotation_RT(al, be, ga, RT + (size_t)w * 9);
}

// key chain of the Metropolis loop: 6 splits per proposal.  thread = walker.  sub[(p*6+i)][w]
__global__ void k_mcmc_keychain(int nw, int nmpm, uint32_t* __restrict__ keys, uint2* __restrict__ sub) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  Key k{keys[2 * w], keys[2 * w + 1]};
  for (int p = 0; p < nmpm * 6; ++p) {
    Key s;
    rng_split(k, s);
    sub[(size_t)p * nw + w] = make_uint2(s.a, s.b);
  }
  keys[2 * w] = k.a;
  keys[2 * w + 1] = k.b;
}
// draws of every proposal: thread = (proposal, walker)
__global__ void k_mcmc_draws(int nw, int nmpm, int n_up, int n_dn, const uint2* __restrict__ sub, int* __restrict__ rsel,
                             int* __restrict__ raxis, double* __restrict__ rg, double* __restrict__ rb) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nmpm * nw) return;
  const int w = (int)(t % nw);
  const int p = (int)(t / nw);
  auto K = [&](int i) {
    const uint2 v = sub[((size_t)p * 6 + i) * nw + w];
    return Key{v.x, v.y};
  };
  const bool is_up = rng_randint(K(0), (uint32_t)(n_up + n_dn)) < n_up;
  const int iu = rng_randint(K(1), (uint32_t)n_up);
  const int id = rng_randint(K(2), (uint32_t)n_dn);
  rsel[t] = is_up ? iu : n_up + id;
  rg[t] = rng_normal(K(3));
  raxis[t] = rng_randint(K(4), 3u);
  rb[t] = rng_uniform_bits(rng_bits64(K(5)), 0.0, 1.0);
}

#define QE_MCMC_EXTERN(NMO_, CART_) extern template int mcmc_launch_one<NMO_, CART_>(qe_engine*, McmcArgs&, int, double, cudaStream_t);
QE_MCMC_EXTERN(4, false)
QE_MCMC_EXTERN(4, true)
QE_MCMC_EXTERN(8, false)
QE_MCMC_EXTERN(8, true)
QE_MCMC_EXTERN(16, false)
QE_MCMC_EXTERN(16, true)
#undef QE_MCMC_EXTERN

extern "C" int qe_rotation(qe_engine* h, int nw, const uint32_t* keys, double* RT, void* stream) {
  if (!h || nw <= 0 || !keys || !RT) return fail(QE_ERR_INVALID, "qe_rotation: bad argument");
  { LaunchScope ls_(h, K_ROT, (cudaStream_t)stream);
  k_rotation<<<nblk(nw, 128), 128, 0, (cudaStream_t)stream>>>(nw, keys, RT);
  }
  CHECK_LAUNCH();
  return QE_OK;
}


size_t mcmc_draws_bytes(int nw, int nmpm) {
  const size_t n_draw = (size_t)std::max(1, nmpm) * nw;
  return n_draw * (6 * 8 + 4 + 4 + 8 + 8) + 6 * 256;
}
int mcmc_draws(qe_engine* h, int nw, int nmpm, uint32_t* keys, WsCarve& c, int** rsel, int** raxis, double** rg, double** rb,
               cudaStream_t st) {
  const SysDev& S = h->sys;
  const size_t n_draw = (size_t)std::max(1, nmpm) * nw;
  uint2* sub = c.take<uint2>(n_draw * 6);
  *rsel = c.take<int>(n_draw);
  *raxis = c.take<int>(n_draw);
  *rg = c.take<double>(n_draw);
  *rb = c.take<double>(n_draw);
  if (nmpm > 0) {
    { LaunchScope ls_(h, K_KEYCHAIN, st);
    k_mcmc_keychain<<<nblk(nw, 64), 64, 0, st>>>(nw, nmpm, keys, sub);
    }
    CHECK_LAUNCH();
    { LaunchScope ls_(h, K_DRAWS, st);
    k_mcmc_draws<<<nblk((long long)nmpm * nw, 128), 128, 0, st>>>(nw, nmpm, S.n_up, S.n_dn, sub, *rsel, *raxis, *rg, *rb);
    }
    CHECK_LAUNCH();
  }
  return QE_OK;
}

extern "C" int qe_mcmc_update(qe_engine* h, int nw, double* r_up, double* r_dn, uint32_t* keys, double* G, double* Ginv, int nmpm,
                              double Dt, double epsilon_AS, int32_t* acc, int32_t* rej, void* stream) {
  if (!h || nw <= 0 || nmpm < 0 || !r_up || !keys || !G || !Ginv || !acc || !rej || (!r_dn && h->sys.n_dn > 0))
    return fail(QE_ERR_INVALID, "qe_mcmc_update: bad argument");
  const SysDev& S = h->sys;
  if (use_wide(h)) return wide_mcmc_update(h, nw, r_up, r_dn, keys, G, Ginv, nmpm, Dt, epsilon_AS, acc, rej, (cudaStream_t)stream);
  cudaStream_t st = (cudaStream_t)stream;
  const int P = h->nmo_pad, nch = h->b_up.n_chunk;
  if (S.n_up > std::min(P, 8)) return fail(QE_ERR_UNSUPPORTED, "qe_mcmc_update: register kernel needs n_up <= min(n_mo, 8)");
  int rc = ensure_ws(h, mcmc_draws_bytes(nw, nmpm) + 4096);
  if (rc) return rc;
  WsCarve c{(char*)h->ws};
  int *rsel, *raxis;
  double *rg, *rb;
  rc = mcmc_draws(h, nw, nmpm, keys, c, &rsel, &raxis, &rg, &rb, st);
  if (rc) return rc;
  McmcArgs A{nw, nmpm, nch, Dt, epsilon_AS, r_up, r_dn, G, Ginv, acc, rej, rsel, raxis, rg, rb, h->b_up.off_cseg, h->b_up.off_cbeg, {0}, h->phase_clk};
  for (int i = 0; i < 16; ++i) A.chunk_warp[i] = h->b_up.chunk_warp[i];
  // the kernel instantiations live in qe_mcmc_i_*.cu
  const bool cart = h->b_up.dev.cart != 0;
#ifdef QE_DEV_MINIMAL
  if (cart || P != 4) return fail(QE_ERR_UNSUPPORTED, "QE_DEV_MINIMAL build");
  return mcmc_launch_one<4, false>(h, A, nw, epsilon_AS, st);
#else
  switch (P) {
    case 4: return cart ? mcmc_launch_one<4, true>(h, A, nw, epsilon_AS, st) : mcmc_launch_one<4, false>(h, A, nw, epsilon_AS, st);
    case 8: return cart ? mcmc_launch_one<8, true>(h, A, nw, epsilon_AS, st) : mcmc_launch_one<8, false>(h, A, nw, epsilon_AS, st);
    default: return cart ? mcmc_launch_one<16, true>(h, A, nw, epsilon_AS, st) : mcmc_launch_one<16, false>(h, A, nw, epsilon_AS, st);
  }
#endif
}
