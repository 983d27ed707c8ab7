// nb_math.cuh - bit-exact arithmetic building blocks of the b200 kernel set.
//
// Everything here must return the same bits as the reference omp3 build (gcc, glibc 2.39,
// -ffp-contract=off): this translation unit is compiled with -fmad=false, so the only fused
// operations are the explicit fma() calls of nb_log (which mirror the FMAs inside glibc's
// own log), and every other + - * / sqrt is one IEEE-754 binary64 round-to-nearest op.
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define NB_HD __host__ __device__ __forceinline__
#define NB_D __device__ __forceinline__
#else
#define NB_HD static inline
#define NB_D static inline
#endif

namespace nb {

// Problem-independent constants, reference neutral_data.h:17-24.
constexpr double kEvToJ = 1.60217646e-19;
constexpr double kAvogadros = 6.02214085774e23;
constexpr double kBarns = 1.0e-28;
constexpr double kParticleMass = 1.674927471213e-27;
constexpr double kMassNo = 1.0e2;
constexpr double kMolarMass = 1.0e-2;
constexpr double kMinEnergyOfInterest = 1.0e0;
constexpr double kOpenBoundCorrection = 1.0e-13;

// ---------------------------------------------------------------------------------------
// Threefry-2x64, 20 rounds: the counter-based generator behind generate_random_numbers
// (reference omp3/neutral.c:632-652; algorithm = Random123/threefry.h:196-286 with the
// 2x64 rotation set of threefry.h:86-93 and the Skein parity constant of :170-171).
// Written from the published algorithm; pinned by the known-answer vectors in
// tests/test_kat.py.
// ---------------------------------------------------------------------------------------
// 64-bit rotate by a compile-time amount. On the device it is spelled as two 32-bit funnel
// shifts (SHF.L.W), or a plain register swap for r == 32: left to itself the compiler builds
// each half from three shift/multiply/logic instructions.
NB_HD uint64_t rotl64(uint64_t v, int r) {
#if defined(__CUDA_ARCH__)
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  if (r >= 32) {
    const unsigned t = lo;
    lo = hi;
    hi = t;
    r -= 32;
  }
  if (r == 0) return ((uint64_t)hi << 32) | lo;
  const unsigned nhi = __funnelshift_l(lo, hi, r);
  const unsigned nlo = __funnelshift_l(hi, lo, r);
  return ((uint64_t)nhi << 32) | nlo;
#else
  return (v << r) | (v >> (64 - r));
#endif
}

#define NB_TF_ROUND(R)            \
  a += b;                         \
  b = rotl64(b, (R)) ^ a;

#define NB_TF_INJECT(K0, K1, S)   \
  a += (K0);                      \
  b += (K1) + (uint64_t)(S);

NB_HD void threefry2x64_20(uint64_t c0, uint64_t c1, uint64_t k0, uint64_t k1,
                           uint64_t& o0, uint64_t& o1) {
  const uint64_t k2 = 0x1BD11BDAA9FC1A22ull ^ k0 ^ k1;
  uint64_t a = c0 + k0;
  uint64_t b = c1 + k1;
  NB_TF_ROUND(16) NB_TF_ROUND(42) NB_TF_ROUND(12) NB_TF_ROUND(31)
  NB_TF_INJECT(k1, k2, 1)
  NB_TF_ROUND(16) NB_TF_ROUND(32) NB_TF_ROUND(24) NB_TF_ROUND(21)
  NB_TF_INJECT(k2, k0, 2)
  NB_TF_ROUND(16) NB_TF_ROUND(42) NB_TF_ROUND(12) NB_TF_ROUND(31)
  NB_TF_INJECT(k0, k1, 3)
  NB_TF_ROUND(16) NB_TF_ROUND(32) NB_TF_ROUND(24) NB_TF_ROUND(21)
  NB_TF_INJECT(k1, k2, 4)
  NB_TF_ROUND(16) NB_TF_ROUND(42) NB_TF_ROUND(12) NB_TF_ROUND(31)
  NB_TF_INJECT(k2, k0, 5)
  o0 = a;
  o1 = b;
}

// Uniform doubles in (0, 1]: u * 2^-64 + 2^-65 (omp3/neutral.c:646-651). The product is
// exact, the sum rounds once.
NB_HD double u64_to_unit(uint64_t u) {
  return (double)u * 0x1p-64 + 0x1p-65;
}

// ctr = {counter, 0}, key = {pkey, master_key}  (omp3/neutral.c:636-641).
NB_HD void random_pair(uint64_t pkey, uint64_t master_key, uint64_t counter, double& r0,
                       double& r1) {
  uint64_t o0, o1;
  threefry2x64_20(counter, 0, pkey, master_key, o0, o1);
  r0 = u64_to_unit(o0);
  r1 = u64_to_unit(o1);
}

// Only the first of the pair is needed for the mean-free-path samples (:129,294).
NB_HD double random_first(uint64_t pkey, uint64_t master_key, uint64_t counter) {
  uint64_t o0, o1;
  threefry2x64_20(counter, 0, pkey, master_key, o0, o1);
  return u64_to_unit(o0);
}

// ---------------------------------------------------------------------------------------
// log(x), bit-identical to glibc 2.39's double-precision log as dispatched on FMA-capable
// x86-64 (__log_fma). `log` is the one operation of the hot path that is not
// IEEE-exact (omp3/neutral.c:130,295), so the kernels replay glibc's exact operation
// sequence - including which multiply-adds glibc's build fused - on
// glibc's own table (glibc_log_table.inc, extracted by tools/gen_glibc_log_table.py).
// Domain: positive normal x (the RNG yields x in [2^-65, 1]); other inputs are not handled.
// Pinned against the host libm in tests/test_log.py (CPU, same source compiled for the
// host) and tests/test_gpu_math.py (device).
// ---------------------------------------------------------------------------------------
struct LogTable {
  double ln2hi, ln2lo;
  double a[5];    // main-path polynomial
  double b[11];   // near-1 polynomial
  double t[256];  // 128 x {invc, logc}
};

NB_HD double bits_to_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d;
  __builtin_memcpy(&d, &u, 8);
  return d;
#endif
}

NB_HD uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u;
  __builtin_memcpy(&u, &d, 8);
  return u;
#endif
}

// The same numbers as compile-time constants: the polynomial coefficients then reach the
// FP64 pipe as constant-bank operands instead of through eighteen loads per call; only the
// {invc, logc} pair, whose index differs from lane to lane, is read from the table in memory.
constexpr double kLogData[2 + 5 + 11 + 256] = {
#include "glibc_log_table.inc"
};
#define NB_LOG_C(name, index) constexpr double name = kLogData[index]

NB_HD double nb_log(double x, const LogTable* __restrict__ L) {
  NB_LOG_C(ln2hi, 0); NB_LOG_C(ln2lo, 1);
  NB_LOG_C(A0, 2); NB_LOG_C(A1, 3); NB_LOG_C(A2, 4); NB_LOG_C(A3, 5); NB_LOG_C(A4, 6);
  NB_LOG_C(B0, 7); NB_LOG_C(B1, 8); NB_LOG_C(B2, 9); NB_LOG_C(B3, 10); NB_LOG_C(B4, 11);
  NB_LOG_C(B5, 12); NB_LOG_C(B6, 13); NB_LOG_C(B7, 14); NB_LOG_C(B8, 15); NB_LOG_C(B9, 16);
  NB_LOG_C(B10, 17);
  const uint64_t ix = double_to_bits(x);
  if (ix - 0x3fee000000000000ull < 0x0003090000000000ull) {
    // 0.9375 <= x < 1.0647: polynomial in r = x - 1 with a split high/low square term.
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = x - 1.0;
    const double r2 = r * r;
    const double r3 = r * r2;
    double p1 = fma(r, B2, B1);
    double p2 = fma(r, B5, B4);
    double p3 = fma(r, B8, B7);
    p1 = fma(r2, B3, p1);
    p2 = fma(r2, B6, p2);
    p3 = fma(r2, B9, p3);
    p3 = fma(r3, B10, p3);
    double q = fma(p3, r3, p2);
    q = fma(q, r3, p1);
    const double t = fma(r, 0x1p27, r);
    const double rhi = fma(-0x1p27, r, t);
    const double rlo = r - rhi;
    const double h2 = rhi * rhi;
    const double hi = fma(h2, B0, r);
    const double lo = fma(h2, B0, r - hi);
    double u = B0 * rlo;
    u = fma(u, r + rhi, lo);
    q = fma(q, r3, u);
    return hi + q;
  }
  const uint64_t tmp = ix - 0x3fe6000000000000ull;
  const int i = (int)((tmp >> 45) & 127);
  const int k = (int)((int64_t)tmp >> 52);
  const double z = bits_to_double(ix - (tmp & 0xfff0000000000000ull));
#if defined(__CUDA_ARCH__)
  const double2 tc = __ldg(reinterpret_cast<const double2*>(L->t) + i);  // one 16-byte load
  const double invc = tc.x, logc = tc.y;
#else
  const double invc = L->t[2 * i];
  const double logc = L->t[2 * i + 1];
#endif
  const double kd = (double)k;
  const double w = fma(kd, ln2hi, logc);
  const double r = fma(z, invc, -1.0);
  const double pa = fma(r, A2, A1);
  const double hi = r + w;
  const double r2 = r * r;
  double lo = (w - hi) + r;
  lo = fma(kd, ln2lo, lo);
  const double r3 = r * r2;
  double pb = fma(r, A4, A3);
  lo = fma(r2, A0, lo);
  pb = fma(pb, r2, pa);
  return fma(r3, pb, lo) + hi;
}

// ---------------------------------------------------------------------------------------
// Derived quantities, each with the reference's operation order.
// ---------------------------------------------------------------------------------------

// number_density = (rho * AVOGADROS / MOLAR_MASS)             omp3/neutral.c:112,289,375
NB_HD double number_density(double rho) { return (rho * kAvogadros) / kMolarMass; }

// macroscopic = number_density * microscopic * BARNS           omp3/neutral.c:113-116
NB_HD double macroscopic(double nd, double micro) { return (nd * micro) * kBarns; }

// speed = sqrt((2 * E * eV_TO_J) / PARTICLE_MASS)              omp3/neutral.c:117,297
NB_HD double speed_of(double e) { return sqrt(((2.0 * e) * kEvToJ) / kParticleMass); }

// Path-length heating estimator                                 omp3/neutral.c:474-495
// heat_response depends on (E, sigma_a/sigma_t) only, so callers may cache it between
// collisions; the value is the same either way.
// `q` is the quotient sigma_a / sigma_t of :482,486 (the same division twice).
NB_HD double heating_response_q(double e, double q) {
  constexpr double c1 =
      (kMassNo * kMassNo + kMassNo + 1) / ((kMassNo + 1) * (kMassNo + 1));
  const double absorb_heat = q * 0.0;
  const double scatter_heat = (1.0 - q) * (e * c1);
  return (e - scatter_heat) - absorb_heat;
}

NB_HD double heating_response(double e, double sig_a, double sig_t) {
  return heating_response_q(e, sig_a / sig_t);
}

NB_HD double deposition(double weight, double path, double sig_t_barns, double response,
                        double nd) {
  return (((weight * path) * sig_t_barns) * response) * nd;
}

}  // namespace nb
