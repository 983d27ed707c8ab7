// nb_sincos.cuh - sin(x) and cos(x), bit-identical to glibc 2.39's double-precision sin/cos as
// dispatched on FMA-capable x86-64 (__sin_fma / __cos_fma, sysdeps/ieee754/dbl-64/s_sin.c).
//
// inject_particles turns a uniform angle theta = 2*pi*r into the direction (cos theta,
// sin theta) with libm (omp3/neutral.c:611-614). These two are - with log - the only
// operations on the path that are not IEEE-exact, so a bank injected on the device is
// bit-identical to the reference's only if the device replays glibc's exact operation
// sequence, including which multiply-adds glibc's build fused. The sequence below was read off
// the disassembly of the container's libm.so.6 (every fma() is a vfmadd/vfnmadd/vfmsub there,
// every other operation a separately rounded one) and runs on glibc's own table
// (glibc_sincos_table.inc, extracted by tools/gen_glibc_sincos_table.py).
//
// Domain: high word of |x| below 0x419921FB, i.e. |x| < ~1.054e8 (the branches below glibc's
// huge-argument reduction); theta lies in (0, 2*pi]. Larger or non-finite arguments return
// NaN - inject never produces them.
// Pinned against the host libm in tests/test_sincos.py (CPU, this source compiled for the
// host) and tests/test_gpu_math.py (device).
#pragma once

#include "nb_math.cuh"

namespace nb {

struct SinCosTable {
  double t[440];  // 110 x {sn, ssn, cs, ccs}
};

namespace sc {
// s_sin.c / usncs.h constants, as found in the binary
constexpr double kBig = 0x1.8p+45;
constexpr double kSn3 = -0x1.5555555555515p-3, kSn5 = 0x1.11110e829872fp-7;
constexpr double kCs2 = 0x1p-1, kCs4 = -0x1.5555555555535p-5, kCs6 = 0x1.6c16bedd9e239p-10;
constexpr double kS1 = -0x1.5555555555555p-3, kS2 = 0x1.1111111110ecep-7,
                 kS3 = -0x1.a01a019db08b8p-13, kS4 = 0x1.71de27b9a7ed9p-19,
                 kS5 = -0x1.addffc2fcdf59p-26;
constexpr double kHp0 = 0x1.921fb54442d18p+0, kHp1 = 0x1.1a62633145c07p-54;
constexpr double kToInt = 0x1.8p+52, kHpInv = 0x1.45f306dc9c883p-1;
constexpr double kMp1 = 0x1.921fb58p+0, kMp2 = -0x1.dde973cp-27;
constexpr double kPp3 = -0x1.cb3b398p-55, kPp4 = -0x1.d747f23e32ed7p-83;

NB_HD double with_sign_of(double mag, double sgn) {
  return bits_to_double((double_to_bits(mag) & 0x7fffffffffffffffull) |
                        (double_to_bits(sgn) & 0x8000000000000000ull));
}
NB_HD double abs_of(double v) { return bits_to_double(double_to_bits(v) & 0x7fffffffffffffffull); }

// TAYLOR_SIN(a*a, a, da)
NB_HD double taylor_sin(double a, double da) {
  const double xx = a * a;
  double p = fma(xx, kS5, kS4);
  p = fma(xx, p, kS3);
  p = fma(xx, p, kS2);
  p = fma(xx, p, kS1);
  const double h = da * 0.5;
  const double t = fma(p, a, -h);
  return a + fma(xx, t, da);
}

// Table index of u = big + |x|: the low word of u, times four.
NB_HD int table_index(double u) { return (int)((unsigned)double_to_bits(u) << 2); }

// do_sin(a, da) for |a| >= 0.126
NB_HD double do_sin(double a, double da, const double* __restrict__ T) {
  if (a <= 0.0) da = -da;
  const double aa = abs_of(a);
  const double u = aa + kBig;
  const double x = aa - (u - kBig);
  const int k = table_index(u);
  const double xx = x * x;
  const double s = x + fma(x * xx, fma(xx, kSn5, kSn3), da);
  const double c = fma(x, da, xx * fma(xx, fma(xx, kCs6, kCs4), kCs2));
  const double sn = T[k], ssn = T[k + 1], cs = T[k + 2], ccs = T[k + 3];
  const double cor = fma(s, cs, fma(-c, sn, fma(s, ccs, ssn)));
  return with_sign_of(sn + cor, a);
}

// do_cos(a, da)
NB_HD double do_cos(double a, double da, const double* __restrict__ T) {
  if (a < 0.0) da = -da;
  const double aa = abs_of(a);
  const double u = aa + kBig;
  const double x = (aa - (u - kBig)) + da;
  const int k = table_index(u);
  const double xx = x * x;
  const double s = fma(x * xx, fma(xx, kSn5, kSn3), x);
  const double c = xx * fma(xx, fma(xx, kCs6, kCs4), kCs2);
  const double sn = T[k], ssn = T[k + 1], cs = T[k + 2], ccs = T[k + 3];
  const double cor = fma(-s, sn, fma(-c, cs, fma(-s, ssn, ccs)));
  return cs + cor;
}

// reduce_sincos: x = n*pi/2 + (a + da), returns n mod 4
NB_HD int reduce(double x, double& a, double& da) {
  const double t = fma(x, kHpInv, kToInt);
  const double xn = t - kToInt;
  const int n = (int)((unsigned)double_to_bits(t) & 3u);
  const double y = fma(-xn, kMp2, fma(-xn, kMp1, x));
  const double t2 = fma(-xn, kPp3, y);
  double db = fma(-kPp3, xn, y - t2);
  const double b = fma(-xn, kPp4, t2);
  db = db + fma(-xn, kPp4, t2 - b);
  a = b;
  da = db;
  return n;
}

// do_sincos(a, da, n)
NB_HD double quadrant(double a, double da, int n, const double* __restrict__ T) {
  double r;
  if (n & 1)
    r = do_cos(a, da, T);
  else
    r = abs_of(a) < 0.126 ? taylor_sin(a, da) : do_sin(a, da, T);
  return (n & 2) ? -r : r;
}
}  // namespace sc

NB_HD double nb_sin(double x, const SinCosTable* __restrict__ tab) {
  const double* T = tab->t;
  const unsigned k = (unsigned)(double_to_bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;
  if (k < 0x3feb6000u)
    return sc::abs_of(x) < 0.126 ? sc::taylor_sin(x, 0.0) : sc::do_sin(x, 0.0, T);
  if (k < 0x400368fdu)
    return sc::with_sign_of(sc::do_cos(sc::kHp0 - sc::abs_of(x), sc::kHp1, T), x);
  if (k < 0x419921fbu) {
    double a, da;
    const int n = sc::reduce(x, a, da);
    return sc::quadrant(a, da, n, T);
  }
  return bits_to_double(0x7ff8000000000000ull);
}

NB_HD double nb_cos(double x, const SinCosTable* __restrict__ tab) {
  const double* T = tab->t;
  const unsigned k = (unsigned)(double_to_bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) return 1.0;
  if (k < 0x3feb6000u) return sc::do_cos(x, 0.0, T);
  if (k < 0x400368fdu) {
    const double y = sc::kHp0 - sc::abs_of(x);
    const double a = y + sc::kHp1;
    const double da = (y - a) + sc::kHp1;
    return sc::abs_of(a) < 0.126 ? sc::taylor_sin(a, da) : sc::do_sin(a, da, T);
  }
  if (k < 0x419921fbu) {
    double a, da;
    const int n = sc::reduce(x, a, da);
    return sc::quadrant(a, da, n + 1, T);
  }
  return bits_to_double(0x7ff8000000000000ull);
}

}  // namespace nb
