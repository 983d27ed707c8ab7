// nb_history.cuh - helpers shared by the event-loop kernels (history.cu, collide.cu).
#pragma once

#include "nb_device.cuh"
#include "transport.cuh"

namespace nb {

__device__ __forceinline__ bool is_mixed_tile(double t) {
  return double_to_bits(t) == kMixedTileBits;
}

__device__ __forceinline__ double coarse_value(const StepArgs& a, int cx, int cy) {
  return __ldg(a.tiles.coarse + (cy >> kCoarseShift) * a.tiles.coarse_tx + (cx >> kCoarseShift));
}

__device__ __forceinline__ double fine_value(const StepArgs& a, int cx, int cy) {
  return __ldg(a.tiles.fine + (cy >> kTileShift) * a.tiles.fine_tx + (cx >> kTileShift));
}

// Per-particle flag bits kept in one register.
enum : unsigned {
  kFlagSpeedOk = 1u,      // the speed is inside the proven range of div_by_known
  kFlagCellMfpOk = 2u,    // so is the cell mean free path
  kFlagDivOk = kFlagSpeedOk | kFlagCellMfpOk,
  kFlagInvStale = 64u,    // the two reciprocals have not been recomputed since their divisors
                          // changed: collision chains never need them, the next facet does
  kFlagCoarseMixed = 4u,  // the current coarse tile is not uniform: consult the fine map
  kFlagFineMixed = 8u,    // nor is the current 16x16 tile: densities come from the mesh
  kFlagAnyMixed = kFlagCoarseMixed | kFlagFineMixed,
  kFlagPending = 16u,     // collisions deposited energy that no tally flush has taken yet
  kFlagDead = 32u,
};

// What a facet reads of the quantities that follow from the energy and the cell density
// (omp3/neutral.c:112-117, 135, 481-491).
struct Derived {
  double stb;           // (sigma_s + sigma_a) * BARNS, the deposition's cross section
  double heat;          // heating response of :481-491
  double cell_mfp;      // 1 / Sigma_t, :135
  double cell_mfp_inv;  // correctly rounded 1 / cell_mfp for div_by_known
};

// Recomputes them - and Sigma_s (mean-free-path sampling, :130,295) and p_absorb (:231-232),
// which only collisions read - from the energy and the number density.
__device__ __forceinline__ void derive(const StepArgs& a, double e, double nd, Derived& d,
                                       double& Sig_s, double& p_absorb, unsigned& flags) {
  double sig_s, sig_a;
  cs_lookup_pair_staged(a, e, sig_s, sig_a);
  const double sig_t = sig_s + sig_a;
  d.stb = sig_t * kBarns;
  d.heat = heating_response(e, sig_a, sig_t);
  const double S_s = macroscopic(nd, sig_s);
  const double S_a = macroscopic(nd, sig_a);
  const double S_t = S_s + S_a;
  Sig_s = S_s;
  p_absorb = S_a / S_t;
  d.cell_mfp = 1.0 / S_t;
  flags = (flags & ~kFlagCellMfpOk) | kFlagInvStale;  // cell_mfp_inv is recomputed on demand
}

// Direction of travel along one axis as a cell step: +1, -1, or 0 for a component that is
// exactly zero (omp3/neutral.c:333-366 tests > 0 and < 0 separately).
__device__ __forceinline__ int axis_step(double o) { return (o > 0.0) ? 1 : ((o < 0.0) ? -1 : 0); }

// The same step read off ui = 1 / (o * speed), which the loop already holds: ui has the sign
// of o, and is infinite exactly when o is zero (the speed is positive and finite).
__device__ __forceinline__ int axis_step_from_reciprocal(double ui) {
  const int hi = __double2hiint(ui);
  const int s = (hi >> 31) | 1;
  return ((unsigned)(hi & 0x7fffffff) >= 0x7ff00000u) ? 0 : s;
}

// The edge a particle in cell c is heading for (omp3/neutral.c:442-444, 448-450): the far
// edge when the direction component is >= 0, else the near edge pulled in by
// OPEN_BOUND_CORRECTION.
__device__ __forceinline__ double target_edge(const double* __restrict__ edge, int c, int step) {
  const double v = __ldg(edge + c + (step >= 0 ? 1 : 0));
  return step >= 0 ? v : v - kOpenBoundCorrection;
}

// update_tallies (omp3/neutral.c:408-420). With kPreReduce the lanes of the warp that flush
// into the same cell at the same moment are combined with shuffles first and their leader
// issues one atomic (north_star phase 5). Off by default: measured, it loses - see DESIGN.md 4.
template <bool kPreReduce>
__device__ __forceinline__ void tally_add(double* __restrict__ tally, int cell, double value) {
  if (kPreReduce) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(__activemask(), cell);
    if (peers != (1u << lane)) {
      double sum = 0.0;
      for (unsigned rest = peers; rest; rest &= rest - 1)
        sum += __shfl_sync(peers, value, __ffs(rest) - 1);
      if (lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(tally + cell, sum);
      return;
    }
  }
  atomicAdd(tally + cell, value);
}

}  // namespace nb
