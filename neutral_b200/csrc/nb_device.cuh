// nb_device.cuh - device helpers shared by the history kernels.
#pragma once

#include "nb_bank.cuh"

namespace nb {

// --------------------------------------------------------------------------------------
// Cross-section lookup: index `ind` with keys[ind] <= e < keys[ind+1] and the linear
// interpolation of omp3/neutral.c:514-516. The bracketing interval of a strictly increasing
// grid is unique, so bisection finds the reference's `ind`.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ int cs_bracket(const double* __restrict__ keys, int n, double e) {
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (e < __ldg(keys + mid)) hi = mid; else lo = mid;
  }
  return lo;
}

__device__ __forceinline__ double cs_interp(const double* __restrict__ keys,
                                            const double* __restrict__ vals, int ind,
                                            double e) {
  const double k0 = __ldg(keys + ind), k1 = __ldg(keys + ind + 1);
  const double v0 = __ldg(vals + ind), v1 = __ldg(vals + ind + 1);
  return v0 + ((e - k0) / (k1 - k0)) * (v1 - v0);
}

// The two tables share one energy grid, bit for bit: equal lengths (host) and no differing
// grid point found by this timestep's staging kernel (device). The counter is written before
// any kernel that asks and never while one runs.
__device__ __forceinline__ bool same_grid(const StepArgs& a) {
  return a.same_keys && a.totals[kTotGridsDiffer] == 0ull;
}

__device__ __forceinline__ void cs_lookup_pair(const StepArgs& a, double e, double& sig_s,
                                               double& sig_a) {
  const int is = cs_bracket(a.s_keys, a.s_n, e);
  sig_s = cs_interp(a.s_keys, a.s_vals, is, e);
  const int ia = same_grid(a) ? is : cs_bracket(a.a_keys, a.a_n, e);
  sig_a = cs_interp(a.a_keys, a.a_vals, ia, e);
}

// --------------------------------------------------------------------------------------
// The same lookup on the staged tables (CsStage, stage.cu): the bucket index narrows the
// bisection to the few grid points that share the energy's leading bits. The interval found
// is the unique bracketing interval, i.e. the reference's `ind` (clamped to the end
// intervals outside the grid, like the oracle's bisection).
// --------------------------------------------------------------------------------------
__device__ __forceinline__ int cs_bracket_staged(const CsStage& c, double e) {
  const long long d = (long long)(double_to_bits(e) - c.bits0);
  int b = 0;
  if (d > 0) {
    const unsigned long long q = (unsigned long long)d >> c.shift;
    b = q < (unsigned long long)(c.nb - 1) ? (int)q : c.nb - 1;
  }
  int lo = max(__ldg(c.bucket + b) - 1, 0);
  int hi = min(__ldg(c.bucket + b + 1), c.n - 1);
  lo = min(lo, c.n - 2);
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (e < __ldg(&c.kv[mid].x)) hi = mid; else lo = mid;
  }
  return lo;
}

__device__ __forceinline__ void cs_lookup_pair_staged(const StepArgs& a, double e, bool same,
                                                      double& sig_s, double& sig_a) {
  const int is = cs_bracket_staged(a.cs_s, e);
  const double2 s0 = __ldg(a.cs_s.kv + is), s1 = __ldg(a.cs_s.kv + is + 1);
  const double frac = (e - s0.x) / (s1.x - s0.x);
  sig_s = s0.y + frac * (s1.y - s0.y);
  if (same) {
    // identical grids: the interval and the interpolation weight are the same bits
    const double va0 = __ldg(&a.cs_a.kv[is].y), va1 = __ldg(&a.cs_a.kv[is + 1].y);
    sig_a = va0 + frac * (va1 - va0);
  } else {
    const int ia = cs_bracket_staged(a.cs_a, e);
    const double2 a0 = __ldg(a.cs_a.kv + ia), a1 = __ldg(a.cs_a.kv + ia + 1);
    sig_a = a0.y + ((e - a0.x) / (a1.x - a0.x)) * (a1.y - a0.y);
  }
}

// ------------------------------------------------------------------------------------
// Exact division by a divisor whose correctly rounded reciprocal is known.
//   q0 = a*y, two FMA refinements; the last one is Markstein's correction step: with
//   y = RN(1/b) and q1 within one ulp of a/b, RN(q1 + (a - b*q1)*y) == RN(a/b).
// Guarded to operands far from overflow/underflow; anything else takes the IEEE divide.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ bool safe_exponent(double v) {
  const unsigned hi = (unsigned)__double2hiint(v) & 0x7ff00000u;
  return hi - 0x30000000u < 0x20000000u;  // 2^-255 <= |v| < 2^257: quotients stay normal
}

__device__ __forceinline__ double div_by_known_unchecked(double a, double b, double y) {
  const double q0 = a * y;
  const double r0 = fma(-b, q0, a);
  const double q1 = fma(r0, y, q0);
  const double r1 = fma(-b, q1, a);
  return fma(r1, y, q1);
}

__device__ __forceinline__ double div_by_known(double a, double b, double y) {
  if (safe_exponent(a) && safe_exponent(b)) return div_by_known_unchecked(a, b, y);
  return a / b;
}

// Division by a compile-time constant through its (compile-time, correctly rounded)
// reciprocal: the same IEEE quotient, ~5 FP64 operations instead of a full divide.
#define NB_DIV_CONST(a, b) nb::div_by_known((a), (b), 1.0 / (b))

// speed_of (nb_math.cuh; omp3/neutral.c:117,297) with the division by PARTICLE_MASS done that way.
__device__ __forceinline__ double speed_of_fast(double e) {
  return sqrt(NB_DIV_CONST((2.0 * e) * kEvToJ, kParticleMass));
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Event totals: shuffle-reduce per warp, one 64-bit atomic per warp and counter.
__device__ __forceinline__ void flush_totals(unsigned long long* totals, unsigned long long nf,
                                             unsigned long long nc, unsigned long long np,
                                             unsigned long long nz, unsigned long long ndead) {
  nf = warp_sum(nf);
  nc = warp_sum(nc);
  np = warp_sum(np);
  nz = warp_sum(nz);
  ndead = warp_sum(ndead);
  if ((threadIdx.x & 31) == 0) {
    if (nf) atomicAdd(totals + kTotFacets, nf);
    if (nc) atomicAdd(totals + kTotCollisions, nc);
    if (np) atomicAdd(totals + kTotProcessed, np);
    if (nz) atomicAdd(totals + kTotCensus, nz);
    if (ndead) atomicAdd(totals + kTotDeaths, ndead);
  }
}

}  // namespace nb
