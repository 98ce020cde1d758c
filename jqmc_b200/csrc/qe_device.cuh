// Device-side building blocks of the walker engine: jax.random-compatible counter RNG, the per-thread
// atomic-orbital evaluator, and small fp64 helpers.  All fp64; lanes of a warp are WALKERS and every
// lane walks the same (nucleus, l)-group / shell / primitive sequence, so control flow is uniform and the
// basis tables are warp-broadcast loads.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "angular_gen.cuh"

namespace qe {

constexpr int MAXF = 28;  // max functions per shell: Cartesian l=6 -> 28; spherical l=6 -> 13

// ------------------------------------------------------------------------------------------------
// Basis tables in device memory (built by qe_engine.cu::build_basis from the reference's per-AO tables:
// AOs_sphe_data jqmc/atomic_orbital.py:780-929 / AOs_cart_data :87-258 / MOs_data molecular_orbital.py:85-140)
// ------------------------------------------------------------------------------------------------
struct BasisDev {
  int n_ao, n_mo, n_orb, n_grp, n_shell, cart;
  int nmo_pad;               // MO accumulators per thread (4, 8 or 16); Cs rows are zero-padded to this
  const int* grp_nuc;        // [n_grp]   nucleus of the group
  const int* grp_l;          // [n_grp]   angular momentum of the group
  const int* grp_sh_begin;   // [n_grp+1] shell range of the group
  const int* sh_prim_off;    // [n_shell+1]
  const short* sh_slot;      // [n_shell*MAXF] AO index of canonical function k, or -1
  int n_prim;                // compressed shell primitives
  const double2* pr_zc;      // [n_prim] {exponent, coefficient * N_p * sqrt((2l+1)/4pi)}  (jqmc/atomic_orbital.py:2316-2349)
  const double* ao_scale;    // [n_ao]   per-AO factor (shell-relative coefficient ratio, Cartesian factorial part)
  const double* Cs;          // [n_ao*nmo_pad]  mo_coefficients^T * ao_scale  (MO layer), or nullptr
};

// ------------------------------------------------------------------------------------------------
// Threefry-2x32 and the jax.random draws used by jQMC (semantics: oracle/jaxrng.py)
// ------------------------------------------------------------------------------------------------
struct Key {
  uint32_t a, b;
};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

__device__ __forceinline__ Key threefry(Key k, uint32_t x0, uint32_t x1) {
  const uint32_t ks0 = k.a, ks1 = k.b, ks2 = k.a ^ k.b ^ 0x1BD11BDAu;
  x0 += ks0;
  x1 += ks1;
#define QE_TF_R(r) \
  x0 += x1;        \
  x1 = rotl32(x1, r) ^ x0;
  QE_TF_R(13) QE_TF_R(15) QE_TF_R(26) QE_TF_R(6) x0 += ks1;
  x1 += ks2 + 1u;
  QE_TF_R(17) QE_TF_R(29) QE_TF_R(16) QE_TF_R(24) x0 += ks2;
  x1 += ks0 + 2u;
  QE_TF_R(13) QE_TF_R(15) QE_TF_R(26) QE_TF_R(6) x0 += ks0;
  x1 += ks1 + 3u;
  QE_TF_R(17) QE_TF_R(29) QE_TF_R(16) QE_TF_R(24) x0 += ks1;
  x1 += ks2 + 4u;
  QE_TF_R(13) QE_TF_R(15) QE_TF_R(26) QE_TF_R(6) x0 += ks2;
  x1 += ks0 + 5u;
#undef QE_TF_R
  return Key{x0, x1};
}

// jax.random.split(key): new key = child 0, subkey = child 1
__device__ __forceinline__ void rng_split(Key& key, Key& sub) {
  const Key k = key;
  key = threefry(k, 0u, 0u);
  sub = threefry(k, 0u, 1u);
}
__device__ __forceinline__ uint64_t rng_bits64(Key k, uint32_t i = 0u) {
  const Key o = threefry(k, 0u, i);
  return (uint64_t(o.a) << 32) | uint64_t(o.b);
}
__device__ __forceinline__ double bits_to_unit(uint64_t bits) {
  return __dadd_rn(__longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ull)), -1.0);
}
// jax.random.uniform(key, (), minval, maxval) in fp64 (unfused mul/add, like the NumPy oracle)
__device__ __forceinline__ double rng_uniform_bits(uint64_t bits, double lo, double hi) {
  const double f = bits_to_unit(bits);
  return fmax(lo, __dadd_rn(__dmul_rn(f, __dadd_rn(hi, -lo)), lo));
}
// jax.random.randint(key, (), 0, span) for int64
__device__ __forceinline__ int rng_randint(Key key, uint32_t span_) {
  const Key k1 = threefry(key, 0u, 0u), k2 = threefry(key, 0u, 1u);
  const uint64_t hi = rng_bits64(k1), lo = rng_bits64(k2);
  const uint64_t span = span_ == 0u ? 1ull : uint64_t(span_);
  uint64_t mult = (1ull << 32) % span;
  mult = (mult * mult) % span;
  const uint64_t off = ((hi % span) * mult + (lo % span)) % span;
  return int(off);
}
// XLA's fp64 erf_inv: Giles (2010), evaluated with unfused Horner steps.
__device__ __forceinline__ double erf_inv_giles(double x) {
  double w = -log1p(-__dmul_rn(x, x));
  double p;
#define QE_H(c) p = __dadd_rn(c, __dmul_rn(p, w));
  if (w < 6.25) {
    w = __dadd_rn(w, -3.125);
    p = -3.6444120640178196996e-21;
    QE_H(-1.685059138182016589e-19) QE_H(1.2858480715256400167e-18) QE_H(1.115787767802518096e-17)
    QE_H(-1.333171662854620906e-16) QE_H(2.0972767875968561637e-17) QE_H(6.6376381343583238325e-15)
    QE_H(-4.0545662729752068639e-14) QE_H(-8.1519341976054721522e-14) QE_H(2.6335093153082322977e-12)
    QE_H(-1.2975133253453532498e-11) QE_H(-5.4154120542946279317e-11) QE_H(1.051212273321532285e-09)
    QE_H(-4.1126339803469836976e-09) QE_H(-2.9070369957882005086e-08) QE_H(4.2347877827932403518e-07)
    QE_H(-1.3654692000834678645e-06) QE_H(-1.3882523362786468719e-05) QE_H(0.0001867342080340571352)
    QE_H(-0.00074070253416626697512) QE_H(-0.0060336708714301490533) QE_H(0.24015818242558961693)
    QE_H(1.6536545626831027356)
  } else if (w < 16.0) {
    w = __dadd_rn(sqrt(w), -3.25);
    p = 2.2137376921775787049e-09;
    QE_H(9.0756561938885390979e-08) QE_H(-2.7517406297064545428e-07) QE_H(1.8239629214389227755e-08)
    QE_H(1.5027403968909827627e-06) QE_H(-4.013867526981545969e-06) QE_H(2.9234449089955446044e-06)
    QE_H(1.2475304481671778723e-05) QE_H(-4.7318229009055733981e-05) QE_H(6.8284851459573175448e-05)
    QE_H(2.4031110387097893999e-05) QE_H(-0.0003550375203628474796) QE_H(0.00095328937973738049703)
    QE_H(-0.0016882755560235047313) QE_H(0.0024914420961078508066) QE_H(-0.0037512085075692412107)
    QE_H(0.005370914553590063617) QE_H(1.0052589676941592334) QE_H(3.0838856104922207635)
  } else {
    w = __dadd_rn(sqrt(w), -5.0);
    p = -2.7109920616438573243e-11;
    QE_H(-2.5556418169965252055e-10) QE_H(1.5076572693500548083e-09) QE_H(-3.7894654401267369937e-09)
    QE_H(7.6157012080783393804e-09) QE_H(-1.4960026627149240478e-08) QE_H(2.9147953450901080826e-08)
    QE_H(-6.7711997758452339498e-08) QE_H(2.2900482228026654717e-07) QE_H(-9.9298272942317002539e-07)
    QE_H(4.5260625972231537039e-06) QE_H(-1.9681778105531670567e-05) QE_H(7.5995277030017761139e-05)
    QE_H(-0.00021503011930044477347) QE_H(-0.00013871931833623122026) QE_H(1.0103004648645343977)
    QE_H(4.8499064014085844221)
  }
#undef QE_H
  return __dmul_rn(p, x);
}
// jax.random.normal(key, ()) in fp64
__device__ __forceinline__ double rng_normal(Key k) {
  const double lo = -0.99999999999999988897769753748;  // nextafter(-1, 0)
  const double u = rng_uniform_bits(rng_bits64(k), lo, 1.0);
  return __dmul_rn(1.4142135623730951, erf_inv_giles(u));
}

// ------------------------------------------------------------------------------------------------
// Per-thread AO evaluator.  A sink receives (ao index, value[, gx, gy, gz, lap]) WITHOUT ao_scale.
// ------------------------------------------------------------------------------------------------
template <bool CART, int L>
struct Ang {
  using type = Sph<L>;
};
template <int L>
struct Ang<true, L> {
  using type = Cart<L>;
};

template <class A, class Sink>
__device__ __forceinline__ void eval_group_val(const BasisDev& B, int g, double dx, double dy, double dz, double r2,
                                               Sink& sink) {
  double S[A::NF];
  A::val(dx, dy, dz, S);
  const int sb = B.grp_sh_begin[g], se = B.grp_sh_begin[g + 1];
  for (int s = sb; s < se; ++s) {
    const int pb = B.sh_prim_off[s], pe = B.sh_prim_off[s + 1];
    double R = 0.0;
    int p = pb;
    for (; p + 1 < pe; p += 2) {  // two independent exp chains per trip
      const double2 a = B.pr_zc[p], b = B.pr_zc[p + 1];
      const double ea = exp(-a.x * r2), eb = exp(-b.x * r2);
      R = fma(a.y, ea, R);
      R = fma(b.y, eb, R);
    }
    if (p < pe) {
      const double2 a = B.pr_zc[p];
      R = fma(a.y, exp(-a.x * r2), R);
    }
    const short* slot = B.sh_slot + s * MAXF;
#pragma unroll
    for (int k = 0; k < A::NF; ++k) {
      const int a = slot[k];
      if (a >= 0) sink.add(a, R * S[k]);
    }
  }
}

template <class A, class Sink>
__device__ __forceinline__ void eval_group_vgl(const BasisDev& B, int g, int l, double dx, double dy, double dz, double r2,
                                               Sink& sink) {
  double S[A::NF], Sx[A::NF], Sy[A::NF], Sz[A::NF], Sl[A::NF];
  A::vgl(dx, dy, dz, S, Sx, Sy, Sz, Sl);
  const int sb = B.grp_sh_begin[g], se = B.grp_sh_begin[g + 1];
  for (int s = sb; s < se; ++s) {
    const int pb = B.sh_prim_off[s], pe = B.sh_prim_off[s + 1];
    double R0 = 0.0, R1 = 0.0, R2 = 0.0;
    for (int p = pb; p < pe; ++p) {
      const double2 zc = B.pr_zc[p];
      const double Z = zc.x;
      const double e = zc.y * exp(-Z * r2);
      R0 += e;
      R1 = fma(Z, e, R1);
      R2 = fma(Z * Z, e, R2);
    }
    // phi = R0*A ; grad = R0*gradA - 2 R1 A d ; lap = R0*lapA + A (4 r2 R2 - 6 R1 - 4 l R1)
    // (r.gradA = l*A: A is homogeneous of degree l; jqmc/atomic_orbital.py:3575-3588, :3484-3505)
    const double m2R1 = -2.0 * R1;
    const double lapfac = 4.0 * r2 * R2 - (6.0 + 4.0 * l) * R1;
    const short* slot = B.sh_slot + s * MAXF;
#pragma unroll
    for (int k = 0; k < A::NF; ++k) {
      const int a = slot[k];
      if (a >= 0) {
        const double Ak = S[k];
        const double t = m2R1 * Ak;
        double lp = lapfac * Ak;
        if (!A::HARMONIC) lp = fma(R0, Sl[k], lp);
        sink.add(a, R0 * Ak, fma(R0, Sx[k], t * dx), fma(R0, Sy[k], t * dy), fma(R0, Sz[k], t * dz), lp);
      }
    }
  }
}

// Evaluate groups [gb, ge) of basis B at point (px,py,pz); Rn = nuclear positions [n_atom*3].
template <bool CART, class Sink>
__device__ __forceinline__ void eval_val(const BasisDev& B, const double* __restrict__ Rn, double px, double py, double pz,
                                         int gb, int ge, Sink& sink) {
  for (int g = gb; g < ge; ++g) {
    const int nuc = B.grp_nuc[g];
    const double dx = px - Rn[3 * nuc], dy = py - Rn[3 * nuc + 1], dz = pz - Rn[3 * nuc + 2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    switch (B.grp_l[g]) {
      case 0: eval_group_val<typename Ang<CART, 0>::type>(B, g, dx, dy, dz, r2, sink); break;
      case 1: eval_group_val<typename Ang<CART, 1>::type>(B, g, dx, dy, dz, r2, sink); break;
      case 2: eval_group_val<typename Ang<CART, 2>::type>(B, g, dx, dy, dz, r2, sink); break;
      case 3: eval_group_val<typename Ang<CART, 3>::type>(B, g, dx, dy, dz, r2, sink); break;
      case 4: eval_group_val<typename Ang<CART, 4>::type>(B, g, dx, dy, dz, r2, sink); break;
      case 5: eval_group_val<typename Ang<CART, 5>::type>(B, g, dx, dy, dz, r2, sink); break;
      default: eval_group_val<typename Ang<CART, 6>::type>(B, g, dx, dy, dz, r2, sink); break;
    }
  }
}

template <bool CART, class Sink>
__device__ __forceinline__ void eval_vgl(const BasisDev& B, const double* __restrict__ Rn, double px, double py, double pz,
                                         int gb, int ge, Sink& sink) {
  for (int g = gb; g < ge; ++g) {
    const int nuc = B.grp_nuc[g];
    const double dx = px - Rn[3 * nuc], dy = py - Rn[3 * nuc + 1], dz = pz - Rn[3 * nuc + 2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    const int l = B.grp_l[g];
    switch (l) {
      case 0: eval_group_vgl<typename Ang<CART, 0>::type>(B, g, l, dx, dy, dz, r2, sink); break;
      case 1: eval_group_vgl<typename Ang<CART, 1>::type>(B, g, l, dx, dy, dz, r2, sink); break;
      case 2: eval_group_vgl<typename Ang<CART, 2>::type>(B, g, l, dx, dy, dz, r2, sink); break;
      case 3: eval_group_vgl<typename Ang<CART, 3>::type>(B, g, l, dx, dy, dz, r2, sink); break;
      case 4: eval_group_vgl<typename Ang<CART, 4>::type>(B, g, l, dx, dy, dz, r2, sink); break;
      case 5: eval_group_vgl<typename Ang<CART, 5>::type>(B, g, l, dx, dy, dz, r2, sink); break;
      default: eval_group_vgl<typename Ang<CART, 6>::type>(B, g, l, dx, dy, dz, r2, sink); break;
    }
  }
}

// ---- sinks ---------------------------------------------------------------------------------------
// MO accumulation: acc[mo] += Cs[a][mo] * v   (AO -> MO contraction fused into the AO evaluation,
// jqmc/molecular_orbital.py:239-261; Cs already carries ao_scale)
template <int NMO>
struct SinkMO {
  const double* __restrict__ Cs;
  double acc[NMO];
  __device__ __forceinline__ void init(const double* cs) {
    Cs = cs;
#pragma unroll
    for (int i = 0; i < NMO; ++i) acc[i] = 0.0;
  }
  __device__ __forceinline__ void add(int a, double v) {
    const double* c = Cs + a * NMO;
#pragma unroll
    for (int i = 0; i < NMO; ++i) acc[i] = fma(c[i], v, acc[i]);
  }
};
template <int NMO>
struct SinkMO5 {
  const double* __restrict__ Cs;
  double acc[5][NMO];
  __device__ __forceinline__ void init(const double* cs) {
    Cs = cs;
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
      for (int i = 0; i < NMO; ++i) acc[q][i] = 0.0;
  }
  __device__ __forceinline__ void add(int a, double v, double gx, double gy, double gz, double lp) {
    const double* c = Cs + a * NMO;
#pragma unroll
    for (int i = 0; i < NMO; ++i) {
      const double ci = c[i];
      acc[0][i] = fma(ci, v, acc[0][i]);
      acc[1][i] = fma(ci, gx, acc[1][i]);
      acc[2][i] = fma(ci, gy, acc[2][i]);
      acc[3][i] = fma(ci, gz, acc[3][i]);
      acc[4][i] = fma(ci, lp, acc[4][i]);
    }
  }
};
// Plain AO output (parity entry qe_eval_orbitals, layer 0): out[q][a][pt]
struct SinkStoreAO {
  double* out;
  const double* scale;
  long long stride_q;  // n_ao*n_pts
  int n_pts;
  __device__ __forceinline__ void add(int a, double v) { out[(long long)a * n_pts] = v * scale[a]; }
  __device__ __forceinline__ void add(int a, double v, double gx, double gy, double gz, double lp) {
    const double s = scale[a];
    double* o = out + (long long)a * n_pts;
    o[0] = v * s;
    o[stride_q] = gx * s;
    o[2 * stride_q] = gy * s;
    o[3 * stride_q] = gz * s;
    o[4 * stride_q] = lp * s;
  }
};

}  // namespace qe
