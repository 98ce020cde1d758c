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
// Basis image: ONE contiguous blob per AO basis, built by qe_engine.cu::build_basis from the reference's per-AO
// tables (AOs_sphe_data jqmc/atomic_orbital.py:780-929 / AOs_cart_data :87-258 / MOs_data molecular_orbital.py:85-140)
// and read either from global memory or, after one cooperative copy, from shared memory (`tab` below is the base of
// whichever copy a kernel uses; every table is addressed as tab + byte offset so that the compiler keeps the address
// space of `tab`: LDS for staged tables, LDG otherwise).
//   seg  int4   {nucleus, l, shell_begin, shell_end}       one per (nucleus, l) group; chunk lists reuse the format
//   sh   int4   {prim_begin, prim_end, row0, pair}         row0 = first row of the shell in the row-ordered tables; pair = 1:
//                                                           this shell and the next one (same group) are both uncontracted
//   pr   double2{-exponent, coefficient * N_p * sqrt((2l+1)/4pi)}   (jqmc/atomic_orbital.py:2316-2349; Cartesian :2243-2244)
//   pr2  double2{-exponent * 32/ln2, same coefficient}   value sweeps (qexp_s below);  et double[32] = 2^(j/32)
//   C    double [n_row][nmo_pad]   transposed MO coefficients * per-AO scale, rows in canonical shell order (row0 + k);
//                                  rows of functions a shell does not have are zero
//   row_ao int[n_row], row_scale double[n_row]            AO index (-1: hole) and per-AO scale of each row
//   Rn   double [n_atom*3]
// ------------------------------------------------------------------------------------------------
struct BasisDev {
  int n_ao, n_mo, n_orb, n_grp, n_shell, n_prim, n_row, cart, lmax;
  int nmo_pad;                // MO accumulators per thread (4, 8 or 16); C rows are zero-padded to this
  int off_seg, off_sh, off_pr, off_C, off_C2, off_rowao, off_rowscale, off_Rn;  // byte offsets into the blob (16-aligned)
  int off_pr2, off_et;        // value-sweep primitive table {-exponent * 32/ln2, coefficient}; 2^(j/32) table of qexp_s
  int off_prf;                // float2 {-exponent / ln2, coefficient}: value sweeps of the mixed-precision mode
  int bytes;                  // blob size (multiple of 16)
  const char* g;              // blob in global memory
};

// ------------------------------------------------------------------------------------------------
// exp(x) for -707 <= x < 709 (every exponential on the hot path: Gaussian primitives, Jastrow, ECP radial terms, ratios).
// Branch-free: k = rint(x log2 e) by the magic-number trick, r = x - k ln2 (two-term Cody-Waite with FMA),
// degree-11 near-minimax polynomial (max relative error 1.7e-17 before rounding), scaling by 2^k through the exponent
// field.  Arguments below -707 return a denormal (< 2.3e-308) instead of the exact tiny value: irrelevant at fp64 resolution.
// ------------------------------------------------------------------------------------------------
__constant__ double QE_EXPC[16] = {1.4426950408889634,      6755399441055744.0,      -0.6931471805599453,      -2.3190468138462996e-17,
                                   2.5110037605963777e-08,  2.763263963904103e-07,   2.755724091857897e-06,    2.4801485482328494e-05,
                                   0.00019841269890047113,  0.0013888888952314775,   0.008333333333319601,     0.0416666666664881,
                                   0.1666666666666668,      0.5000000000000019,      1.0,                      -707.0};
__device__ __forceinline__ double qexp(double x) {
  // constants come from the constant bank (DFMA takes a c[][] operand directly; 64-bit immediates would cost two UMOVs each)
  const bool under = x < QE_EXPC[15];  // result below 1e-307: return (a denormal close to) zero
  const double t = fma(x, QE_EXPC[0], QE_EXPC[1]);
  const double k = t - QE_EXPC[1];
  double r = fma(k, QE_EXPC[2], x);
  r = fma(k, QE_EXPC[3], r);
  double p = QE_EXPC[4];
  p = fma(p, r, QE_EXPC[5]);
  p = fma(p, r, QE_EXPC[6]);
  p = fma(p, r, QE_EXPC[7]);
  p = fma(p, r, QE_EXPC[8]);
  p = fma(p, r, QE_EXPC[9]);
  p = fma(p, r, QE_EXPC[10]);
  p = fma(p, r, QE_EXPC[11]);
  p = fma(p, r, QE_EXPC[12]);
  p = fma(p, r, QE_EXPC[13]);
  p = fma(p, r, QE_EXPC[14]);
  p = fma(p, r, QE_EXPC[14]);
  return __hiloint2double(under ? 0 : __double2hiint(p) + (__double2loint(t) << 20), __double2loint(p));
}

// ------------------------------------------------------------------------------------------------
// exp(-Z r2) of the Gaussian primitives in the value sweeps, the dominant arithmetic of the engine: 10 fp64-pipe
// instructions instead of the 19 of `a = -Z r2; qexp(a)`.  With zs = -Z 32/ln2 (table pr2) the argument is reduced in units
// of ln2/32:  k = rint(r2 zs) by the magic-number trick,  u = fma(r2, zs, -k) in [-1/2, 1/2] (one FMA, no separate product),
// exp = 2^(k >> 5) * 2^((k & 31)/32) * 2^(u/32): the middle factor from a 32-entry table (`et`, staged with the basis image
// in shared memory: the lookup runs on the LSU, not on the fp64 pipe), the last one a degree-6 polynomial (truncation
// 3.5e-18), the first one through the exponent field.  Relative error <= 1.2e-16 + |Z r2| 1.1e-16 (rounding of zs; terms
// with a large argument are exponentially small).  Arguments must be <= 0; results below 2^-1020 are flushed to zero.
// ------------------------------------------------------------------------------------------------
__constant__ double QE_EXPS[8] = {6755399441055744.0,     0.02166084939249829,    0.0002345961982022468, 1.693850972437182e-06,
                                  9.172562701824643e-09,  3.973709984549416e-11,  1.4345655584131934e-13, 1.0};
__device__ __forceinline__ double qexp_s(double r2, double zs, const double* __restrict__ et) {
  const double t = fma(r2, zs, QE_EXPS[0]);
  const long long tb = __double_as_longlong(t);
  const int ki = (int)tb;  // rint(r2 zs) <= 0 (two's complement in the low mantissa bits)
  const double kf = t - QE_EXPS[0];
  const double u = fma(r2, zs, -kf);
#ifdef QE_EXP_ESTRIN
  // Estrin: 1 + c1 u + u^2 (c2 + c3 u) + u^4 ((c4 + c5 u) + u^2 c6): dependent depth 4 instead of 6 for two more
  // instructions (same truncation error, rounding differs in the last place).  Selected per translation unit: measured
  // faster in the fused walker kernel only (profiles/r02_walker_history.md, A/B table)
  const double u2 = u * u;
  const double pa = fma(QE_EXPS[1], u, QE_EXPS[7]);
  const double pb = fma(QE_EXPS[3], u, QE_EXPS[2]);
  const double pc = fma(QE_EXPS[5], u, QE_EXPS[4]);
  const double u4 = u2 * u2;
  const double pd = fma(u2, pb, pa);
  const double pe = fma(u2, QE_EXPS[6], pc);
  const double p = fma(u4, pe, pd);
#else
  double p = QE_EXPS[6];
  p = fma(p, u, QE_EXPS[5]);
  p = fma(p, u, QE_EXPS[4]);
  p = fma(p, u, QE_EXPS[3]);
  p = fma(p, u, QE_EXPS[2]);
  p = fma(p, u, QE_EXPS[1]);
  p = fma(p, u, QE_EXPS[7]);
#endif
  const double v = et[ki & 31] * p;
  // valid iff magic - 32640 <= t <= magic (positive doubles order like their bit patterns): k >= -1020 * 32
  const bool under = tb < 0x4337FFFFFFFF8080ll;
  return __hiloint2double(under ? 0 : __double2hiint(v) + ((ki >> 5) << 20), __double2loint(v));
}
// the 2^(j/32) table (copied into every basis image at build time; also used to fill it on the host)
static const double QE_EXP2_TABLE[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924, 1.1387886347566916,
    1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332, 1.2968395546510096,
    1.3252366431597413, 1.3542555469368927, 1.383909881963832, 1.4142135623730951, 1.4451808069770467, 1.4768261459394993,
    1.5091644275934228, 1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429,
    1.718619298122478, 1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103, 1.9152065613971474,
    1.9571441241754002};

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
// Per-thread AO sweep.  A sink receives (row, value[, gx, gy, gz, lap]) WITHOUT the per-AO scale (it is folded into C
// for the MO sinks; AO-layer sinks multiply by row_scale themselves).  `row` = row0 + canonical function index.
// ------------------------------------------------------------------------------------------------
template <bool CART, int L>
struct Ang {
  using type = Sph<L>;
};
template <int L>
struct Ang<true, L> {
  using type = Cart<L>;
};

// radial sums of one shell: R0 = sum c e, and (VGL) R1 = sum Z c e, R2 = sum Z^2 c e, with e = exp(-Z r2)
__device__ __forceinline__ void shell_radial3(const double2* __restrict__ pr, int pb, int pe, double r2, double& R0, double& R1,
                                              double& R2) {
  R0 = 0.0;
  R1 = 0.0;
  R2 = 0.0;
  int p = pb;
  for (; p + 1 < pe; p += 2) {
    const double2 a = pr[p], b = pr[p + 1];
    const double ea = a.y * qexp(a.x * r2), eb = b.y * qexp(b.x * r2);
    R0 += ea;
    R1 = fma(-a.x, ea, R1);
    R2 = fma(a.x * a.x, ea, R2);
    R0 += eb;
    R1 = fma(-b.x, eb, R1);
    R2 = fma(b.x * b.x, eb, R2);
  }
  if (p < pe) {
    const double2 a = pr[p];
    const double ea = a.y * qexp(a.x * r2);
    R0 += ea;
    R1 = fma(-a.x, ea, R1);
    R2 = fma(a.x * a.x, ea, R2);
  }
}

// radial sums R[i] = sum_p c_p exp(-Z_p r2[i]) of one shell at NP points (independent exp chains: 2 primitives x NP points)
// (pr = the value-sweep table pr2: {-exponent * 32/ln2, coefficient}; et = 2^(j/32) table, see qexp_s)
template <int NP>
__device__ __forceinline__ void shell_radial_n(const double2* __restrict__ pr, const double* __restrict__ et, int pb, int pe,
                                               const double* __restrict__ r2, double* __restrict__ R) {
  if (pe - pb == 1) {  // uncontracted shell (most shells of a correlation-consistent basis): straight-line code
    const double2 a = pr[pb];
#pragma unroll
    for (int i = 0; i < NP; ++i) R[i] = a.y * qexp_s(r2[i], a.x, et);
    return;
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) R[i] = 0.0;
  int p = pb;
  for (; p + 1 < pe; p += 2) {
    const double2 a = pr[p], b = pr[p + 1];
    double ea[NP], eb[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      ea[i] = qexp_s(r2[i], a.x, et);
      eb[i] = qexp_s(r2[i], b.x, et);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      R[i] = fma(a.y, ea[i], R[i]);
      R[i] = fma(b.y, eb[i], R[i]);
    }
  }
  if (p < pe) {
    const double2 a = pr[p];
#pragma unroll
    for (int i = 0; i < NP; ++i) R[i] = fma(a.y, qexp_s(r2[i], a.x, et), R[i]);
  }
}

template <class A, int NP, class Sink>
__device__ __forceinline__ void eval_seg_val_n(const int4* __restrict__ sh, const double2* __restrict__ pr,
                                               const double* __restrict__ et, int sb, int se,
                                               const double* __restrict__ dx, const double* __restrict__ dy,
                                               const double* __restrict__ dz, const double* __restrict__ r2, Sink& sink) {
  double S[NP][A::NF];
#pragma unroll
  for (int i = 0; i < NP; ++i) A::val(dx[i], dy[i], dz[i], S[i]);
  for (int s = sb; s < se; ++s) {
    const int4 q = sh[s];
    if (q.w && s + 1 < se) {  // (a chunk boundary may separate the two: then each is swept on its own)
      // two consecutive uncontracted shells (flagged at build time): their exponentials are evaluated together -- 2 NP
      // independent chains in one straight-line block instead of NP -- and contracted one after the other
      const int4 q2 = sh[s + 1];
      double wv[A::NF], wv2[A::NF];
      sink.template prefetch<A::NF>(q.z, wv);
      sink.template prefetch<A::NF>(q2.z, wv2);
      const double2 a = pr[q.x], b = pr[q2.x];
      double Ra[NP], Rb[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        Ra[i] = a.y * qexp_s(r2[i], a.x, et);
        Rb[i] = b.y * qexp_s(r2[i], b.x, et);
      }
#pragma unroll
      for (int k = 0; k < A::NF; ++k) {
        double v[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) v[i] = Ra[i] * S[i][k];
        sink.add_nw(q.z + k, v, wv[k]);
      }
#pragma unroll
      for (int k = 0; k < A::NF; ++k) {
        double v[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) v[i] = Rb[i] * S[i][k];
        sink.add_nw(q2.z + k, v, wv2[k]);
      }
      ++s;
      continue;
    }
    // sinks that contract with a per-walker weight vector issue the loads of this shell's weights BEFORE the exponentials
    // (the loads then overlap the radial part instead of stalling every multiply-add); a no-op for the other sinks
    double wv[A::NF];
    sink.template prefetch<A::NF>(q.z, wv);
    double R[NP];
    shell_radial_n<NP>(pr, et, q.x, q.y, r2, R);
#pragma unroll
    for (int k = 0; k < A::NF; ++k) {
      double v[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) v[i] = R[i] * S[i][k];
      sink.add_nw(q.z + k, v, wv[k]);
    }
  }
}

// ---- mixed-precision mode (jqmc/_precision.py:345-374): the `ao_eval` zone in fp32 -------------------------------------------
// AO VALUES are evaluated in single precision -- r - R is formed in fp64 first and then rounded (jqmc/_precision.py:61-76),
// exponentials through exp2f -- and handed to the sink in fp64: the AO -> MO contraction (zone `mo_eval`) stays in double
// precision, exactly the reference's split.  Gradients / Laplacians (zone `ao_grad_lap`) are never evaluated here.
template <class A, int NP, class Sink>
__device__ __forceinline__ void eval_seg_val_n_f32(const int4* __restrict__ sh, const float2* __restrict__ prf, int sb, int se,
                                                   const float* __restrict__ dx, const float* __restrict__ dy,
                                                   const float* __restrict__ dz, const float* __restrict__ r2, Sink& sink) {
  float S[NP][A::NF];
#pragma unroll
  for (int i = 0; i < NP; ++i) A::template val<float>(dx[i], dy[i], dz[i], S[i]);
  for (int s = sb; s < se; ++s) {
    const int4 q = sh[s];
    double wv[A::NF];
    sink.template prefetch<A::NF>(q.z, wv);
    float R[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) R[i] = 0.0f;
    for (int p = q.x; p < q.y; ++p) {
      const float2 a = prf[p];
#pragma unroll
      for (int i = 0; i < NP; ++i) R[i] = fmaf(a.y, exp2f(a.x * r2[i]), R[i]);
    }
#pragma unroll
    for (int k = 0; k < A::NF; ++k) {
      double v[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) v[i] = (double)(R[i] * S[i][k]);
      sink.add_nw(q.z + k, v, wv[k]);
    }
  }
}
template <bool CART, int LMAX, int NP, class Sink>
__device__ __forceinline__ void eval_val_n_f32(const char* __restrict__ tab, const BasisDev& B, int off_list, const double* __restrict__ px,
                                               const double* __restrict__ py, const double* __restrict__ pz, int gb, int ge,
                                               Sink& sink) {
  const int4* seg = (const int4*)(tab + off_list);
  const int4* sh = (const int4*)(tab + B.off_sh);
  const float2* pr = (const float2*)(tab + B.off_prf);
  const double* Rn = (const double*)(tab + B.off_Rn);
  for (int g = gb; g < ge; ++g) {
    const int4 q = seg[g];
    const double X = Rn[3 * q.x], Y = Rn[3 * q.x + 1], Z = Rn[3 * q.x + 2];
    float dx[NP], dy[NP], dz[NP], r2[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      dx[i] = (float)(px[i] - X);
      dy[i] = (float)(py[i] - Y);
      dz[i] = (float)(pz[i] - Z);
      r2[i] = dx[i] * dx[i] + dy[i] * dy[i] + dz[i] * dz[i];
    }
    switch (q.y) {
      case 0: eval_seg_val_n_f32<typename Ang<CART, 0>::type, NP>(sh, pr, q.z, q.w, dx, dy, dz, r2, sink); break;
      case 1: eval_seg_val_n_f32<typename Ang<CART, 1>::type, NP>(sh, pr, q.z, q.w, dx, dy, dz, r2, sink); break;
      case 2: eval_seg_val_n_f32<typename Ang<CART, 2>::type, NP>(sh, pr, q.z, q.w, dx, dy, dz, r2, sink); break;
      case 3: if (LMAX >= 3) eval_seg_val_n_f32<typename Ang<CART, (LMAX >= 3 ? 3 : 0)>::type, NP>(sh, pr, q.z, q.w, dx, dy, dz, r2, sink); break;
      case 4: if (LMAX >= 4) eval_seg_val_n_f32<typename Ang<CART, (LMAX >= 4 ? 4 : 0)>::type, NP>(sh, pr, q.z, q.w, dx, dy, dz, r2, sink); break;
      case 5: if (LMAX >= 5) eval_seg_val_n_f32<typename Ang<CART, (LMAX >= 5 ? 5 : 0)>::type, NP>(sh, pr, q.z, q.w, dx, dy, dz, r2, sink); break;
      default: if (LMAX >= 6) eval_seg_val_n_f32<typename Ang<CART, (LMAX >= 6 ? 6 : 0)>::type, NP>(sh, pr, q.z, q.w, dx, dy, dz, r2, sink); break;
    }
  }
}

// Evaluate segments [gb, ge) of the list at byte offset off_list (B.off_seg for the whole basis, or a chunk list) at NP
// points per thread (same shell / primitive sequence, tables loaded once).  `tab` = base of the blob copy (shared or global).
template <bool CART, int LMAX, int NP, class Sink>
__device__ __forceinline__ void eval_val_n(const char* __restrict__ tab, const BasisDev& B, int off_list, const double* __restrict__ px,
                                           const double* __restrict__ py, const double* __restrict__ pz, int gb, int ge,
                                           Sink& sink) {
  const int4* seg = (const int4*)(tab + off_list);
  const int4* sh = (const int4*)(tab + B.off_sh);
  const double2* pr = (const double2*)(tab + B.off_pr2);
  const double* et = (const double*)(tab + B.off_et);
  const double* Rn = (const double*)(tab + B.off_Rn);
  for (int g = gb; g < ge; ++g) {
    const int4 q = seg[g];
    const double X = Rn[3 * q.x], Y = Rn[3 * q.x + 1], Z = Rn[3 * q.x + 2];
    double dx[NP], dy[NP], dz[NP], r2[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      dx[i] = px[i] - X;
      dy[i] = py[i] - Y;
      dz[i] = pz[i] - Z;
      r2[i] = dx[i] * dx[i] + dy[i] * dy[i] + dz[i] * dz[i];
    }
    switch (q.y) {
      case 0: eval_seg_val_n<typename Ang<CART, 0>::type, NP>(sh, pr, et, q.z, q.w, dx, dy, dz, r2, sink); break;
      case 1: eval_seg_val_n<typename Ang<CART, 1>::type, NP>(sh, pr, et, q.z, q.w, dx, dy, dz, r2, sink); break;
      case 2: eval_seg_val_n<typename Ang<CART, 2>::type, NP>(sh, pr, et, q.z, q.w, dx, dy, dz, r2, sink); break;
      case 3: if (LMAX >= 3) eval_seg_val_n<typename Ang<CART, (LMAX >= 3 ? 3 : 0)>::type, NP>(sh, pr, et, q.z, q.w, dx, dy, dz, r2, sink); break;
      case 4: if (LMAX >= 4) eval_seg_val_n<typename Ang<CART, (LMAX >= 4 ? 4 : 0)>::type, NP>(sh, pr, et, q.z, q.w, dx, dy, dz, r2, sink); break;
      case 5: if (LMAX >= 5) eval_seg_val_n<typename Ang<CART, (LMAX >= 5 ? 5 : 0)>::type, NP>(sh, pr, et, q.z, q.w, dx, dy, dz, r2, sink); break;
      default: if (LMAX >= 6) eval_seg_val_n<typename Ang<CART, (LMAX >= 6 ? 6 : 0)>::type, NP>(sh, pr, et, q.z, q.w, dx, dy, dz, r2, sink); break;
    }
  }
}
template <bool CART, int LMAX, class Sink>
__device__ __forceinline__ void eval_val(const char* __restrict__ tab, const BasisDev& B, int off_list, double px, double py, double pz,
                                         int gb, int ge, Sink& sink) {
  eval_val_n<CART, LMAX, 1, Sink>(tab, B, off_list, &px, &py, &pz, gb, ge, sink);
}

template <class A, class Sink>
__device__ __forceinline__ void eval_seg_vgl(const int4* __restrict__ sh, const double2* __restrict__ pr, int sb, int se, int l,
                                             double dx, double dy, double dz, double r2, Sink& sink) {
  double S[A::NF], Sx[A::NF], Sy[A::NF], Sz[A::NF], Sl[A::NF];
  A::vgl(dx, dy, dz, S, Sx, Sy, Sz, Sl);
  for (int s = sb; s < se; ++s) {
    const int4 q = sh[s];
    double R0, R1, R2;
    shell_radial3(pr, q.x, q.y, r2, R0, R1, R2);
    // phi = R0*A ; grad = R0*gradA - 2 R1 A d ; lap = R0*lapA + A (4 r2 R2 - 6 R1 - 4 l R1)
    // (r.gradA = l*A: A is homogeneous of degree l; jqmc/atomic_orbital.py:3575-3588, :3484-3505)
    const double m2R1 = -2.0 * R1;
    const double lapfac = 4.0 * r2 * R2 - (6.0 + 4.0 * l) * R1;
#pragma unroll
    for (int k = 0; k < A::NF; ++k) {
      const double Ak = S[k];
      const double t = m2R1 * Ak;
      double lp = lapfac * Ak;
      if (!A::HARMONIC) lp = fma(R0, Sl[k], lp);
      sink.add(q.z + k, R0 * Ak, fma(R0, Sx[k], t * dx), fma(R0, Sy[k], t * dy), fma(R0, Sz[k], t * dz), lp);
    }
  }
}

template <bool CART, int LMAX, class Sink>
__device__ __forceinline__ void eval_vgl(const char* __restrict__ tab, const BasisDev& B, int off_list, double px, double py, double pz,
                                         int gb, int ge, Sink& sink) {
  const int4* seg = (const int4*)(tab + off_list);
  const int4* sh = (const int4*)(tab + B.off_sh);
  const double2* pr = (const double2*)(tab + B.off_pr);
  const double* Rn = (const double*)(tab + B.off_Rn);
  for (int g = gb; g < ge; ++g) {
    const int4 q = seg[g];
    const double dx = px - Rn[3 * q.x], dy = py - Rn[3 * q.x + 1], dz = pz - Rn[3 * q.x + 2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    const int l = q.y;
    switch (l) {
      case 0: eval_seg_vgl<typename Ang<CART, 0>::type>(sh, pr, q.z, q.w, l, dx, dy, dz, r2, sink); break;
      case 1: eval_seg_vgl<typename Ang<CART, 1>::type>(sh, pr, q.z, q.w, l, dx, dy, dz, r2, sink); break;
      case 2: eval_seg_vgl<typename Ang<CART, 2>::type>(sh, pr, q.z, q.w, l, dx, dy, dz, r2, sink); break;
      case 3: if (LMAX >= 3) eval_seg_vgl<typename Ang<CART, (LMAX >= 3 ? 3 : 0)>::type>(sh, pr, q.z, q.w, l, dx, dy, dz, r2, sink); break;
      case 4: if (LMAX >= 4) eval_seg_vgl<typename Ang<CART, (LMAX >= 4 ? 4 : 0)>::type>(sh, pr, q.z, q.w, l, dx, dy, dz, r2, sink); break;
      case 5: if (LMAX >= 5) eval_seg_vgl<typename Ang<CART, (LMAX >= 5 ? 5 : 0)>::type>(sh, pr, q.z, q.w, l, dx, dy, dz, r2, sink); break;
      default: if (LMAX >= 6) eval_seg_vgl<typename Ang<CART, (LMAX >= 6 ? 6 : 0)>::type>(sh, pr, q.z, q.w, l, dx, dy, dz, r2, sink); break;
    }
  }
}

// ---- sinks ---------------------------------------------------------------------------------------
// MO accumulation: acc[mo] += C[row][mo] * v   (AO -> MO contraction fused into the AO evaluation,
// jqmc/molecular_orbital.py:239-261; C already carries the per-AO scale)
template <int NMO>
struct SinkMO {
  const double2* __restrict__ C;
  double acc[NMO];
  __device__ __forceinline__ void init(const char* c_table) {
    C = (const double2*)c_table;
#pragma unroll
    for (int i = 0; i < NMO; ++i) acc[i] = 0.0;
  }
  __device__ __forceinline__ void add(int row, double v) {
    const double2* c = C + row * (NMO / 2);
#pragma unroll
    for (int i = 0; i < NMO / 2; ++i) {
      const double2 ci = c[i];
      acc[2 * i] = fma(ci.x, v, acc[2 * i]);
      acc[2 * i + 1] = fma(ci.y, v, acc[2 * i + 1]);
    }
  }
  __device__ __forceinline__ void add_n(int row, const double* __restrict__ v) { add(row, v[0]); }
  template <int NF>
  __device__ __forceinline__ void prefetch(int, double*) const {}
  __device__ __forceinline__ void add_nw(int row, const double* __restrict__ v, double) { add(row, v[0]); }
};
// NP points per thread sharing ONE coefficient table (the caller pairs points of the same spin): a row is loaded once
template <int NMO, int NP>
struct SinkMOn {
  const double2* __restrict__ C;
  double acc[NP][NMO];
  __device__ __forceinline__ void init(const char* c_table) {
    C = (const double2*)c_table;
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
      for (int i = 0; i < NMO; ++i) acc[p][i] = 0.0;
  }
  __device__ __forceinline__ void add_n(int row, const double* __restrict__ v) {
    const double2* c = C + row * (NMO / 2);
#pragma unroll
    for (int i = 0; i < NMO / 2; ++i) {
      const double2 ci = c[i];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        acc[p][2 * i] = fma(ci.x, v[p], acc[p][2 * i]);
        acc[p][2 * i + 1] = fma(ci.y, v[p], acc[p][2 * i + 1]);
      }
    }
  }
  template <int NF>
  __device__ __forceinline__ void prefetch(int, double*) const {}
  __device__ __forceinline__ void add_nw(int row, const double* __restrict__ v, double) { add_n(row, v); }
};
template <int NMO>
struct SinkMO5 {
  const double2* __restrict__ C;
  double acc[5][NMO];
  __device__ __forceinline__ void init(const char* c_table) {
    C = (const double2*)c_table;
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
      for (int i = 0; i < NMO; ++i) acc[q][i] = 0.0;
  }
  __device__ __forceinline__ void add(int row, double v, double gx, double gy, double gz, double lp) {
    const double2* c = C + row * (NMO / 2);
#pragma unroll
    for (int i = 0; i < NMO / 2; ++i) {
      const double2 ci = c[i];
      acc[0][2 * i] = fma(ci.x, v, acc[0][2 * i]);
      acc[1][2 * i] = fma(ci.x, gx, acc[1][2 * i]);
      acc[2][2 * i] = fma(ci.x, gy, acc[2][2 * i]);
      acc[3][2 * i] = fma(ci.x, gz, acc[3][2 * i]);
      acc[4][2 * i] = fma(ci.x, lp, acc[4][2 * i]);
      acc[0][2 * i + 1] = fma(ci.y, v, acc[0][2 * i + 1]);
      acc[1][2 * i + 1] = fma(ci.y, gx, acc[1][2 * i + 1]);
      acc[2][2 * i + 1] = fma(ci.y, gy, acc[2][2 * i + 1]);
      acc[3][2 * i + 1] = fma(ci.y, gz, acc[3][2 * i + 1]);
      acc[4][2 * i + 1] = fma(ci.y, lp, acc[4][2 * i + 1]);
    }
  }
};
// Plain AO output (parity entry qe_eval_orbitals, layer 0): out[q][a][pt]
struct SinkStoreAO {
  double* out;
  const int* row_ao;
  const double* row_scale;
  long long stride_q;  // n_ao*n_pts
  int n_pts;
  __device__ __forceinline__ void add(int row, double v) {
    const int a = row_ao[row];
    if (a >= 0) out[(long long)a * n_pts] = v * row_scale[row];
  }
  __device__ __forceinline__ void add_n(int row, const double* __restrict__ v) { add(row, v[0]); }
  template <int NF>
  __device__ __forceinline__ void prefetch(int, double*) const {}
  __device__ __forceinline__ void add_nw(int row, const double* __restrict__ v, double) { add(row, v[0]); }
  __device__ __forceinline__ void add(int row, double v, double gx, double gy, double gz, double lp) {
    const int a = row_ao[row];
    if (a < 0) return;
    const double s = row_scale[row];
    double* o = out + (long long)a * n_pts;
    o[0] = v * s;
    o[stride_q] = gx * s;
    o[2 * stride_q] = gy * s;
    o[3 * stride_q] = gz * s;
    o[4 * stride_q] = lp * s;
  }
};

}  // namespace qe
