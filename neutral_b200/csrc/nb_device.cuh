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

__device__ __forceinline__ void cs_lookup_pair(const StepArgs& a, double e, double& sig_s,
                                               double& sig_a) {
  const int is = cs_bracket(a.s_keys, a.s_n, e);
  sig_s = cs_interp(a.s_keys, a.s_vals, is, e);
  const int ia = a.same_keys ? is : cs_bracket(a.a_keys, a.a_n, e);
  sig_a = cs_interp(a.a_keys, a.a_vals, ia, e);
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
