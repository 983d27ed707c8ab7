// nb_history.cuh - helpers of the event-loop kernel (history.cu).
#pragma once

#include "nb_device.cuh"
#include "nb_fastmath.cuh"
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
  // The two reciprocals a facet divides through (1 / speed, 1 / cell mean free path) are NOT
  // both valid: stale, or a divisor outside the proven range of div_by_known. The top bit, so
  // that one OR folds it into the range test of the dividend (history.cu).
  kFlagDivBad = 0x80000000u,
  kFlagInvStale = 64u,    // the two reciprocals have not been recomputed since their divisors
                          // changed: collision chains never need them, the next facet does
  kFlagCoarseMixed = 4u,  // the current coarse tile is not uniform: consult the fine map
  kFlagFineMixed = 8u,    // nor is the current 16x16 tile: densities come from the mesh
  kFlagAnyMixed = kFlagCoarseMixed | kFlagFineMixed,
  kFlagSameGrid = 2u,     // both tables share one energy grid (same_grid(), read once per step)
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

// x / (x + x) is exactly one half for every finite non-zero x (x + x is exact unless it
// overflows, and 0.5 is representable), so when the two microscopic cross sections are the
// same bits - the reference's two tables are the same file, SURVEY.md 2.1 row 6 - the
// quotients sigma_a / sigma_t (:482) and Sigma_a / Sigma_t (:231-232) need no division. Any
// other input takes the divisions; -DNB_HALF_SHORTCUT=0 always does.
#ifndef NB_HALF_SHORTCUT
#define NB_HALF_SHORTCUT 1
#endif
__device__ __forceinline__ bool finite_nonzero(double v) {
  const unsigned hi = (unsigned)__double2hiint(v) & 0x7fffffffu;
  return hi < 0x7ff00000u && (hi | (unsigned)__double2loint(v)) != 0u;
}

// Recomputes them - and Sigma_s (mean-free-path sampling, :130,295) and p_absorb (:231-232),
// which only collisions read - from the energy and the number density.
__device__ __forceinline__ void derive(const StepArgs& a, double e, double nd, Derived& d,
                                       double& Sig_s, double& p_absorb, unsigned& flags) {
  double sig_s, sig_a;
  cs_lookup_pair_staged(a, e, (flags & kFlagSameGrid) != 0u, sig_s, sig_a);
  const double sig_t = sig_s + sig_a;
  d.stb = sig_t * kBarns;
  const double S_s = macroscopic(nd, sig_s);
  const double S_a = macroscopic(nd, sig_a);
  const double S_t = S_s + S_a;
  Sig_s = S_s;
  if (NB_HALF_SHORTCUT && double_to_bits(sig_s) == double_to_bits(sig_a) &&
      finite_nonzero(sig_t) && finite_nonzero(S_t)) {
    d.heat = heating_response_q(e, 0.5);
    p_absorb = 0.5;
  } else {
    d.heat = heating_response_q(e, sig_a / sig_t);
    p_absorb = S_a / S_t;
  }
  d.cell_mfp = 1.0 / S_t;
  flags |= kFlagInvStale | kFlagDivBad;  // cell_mfp_inv is recomputed on demand
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

// ------------------------------------------------------------------------------------
// The elastic-scatter half of collision_event (omp3/neutral.c:262-297) and everything that
// follows from the new energy, as ONE straight-line block for tables that share their energy
// grid: the same expressions in the same order as the plain code in history.cu, with every
// division, reciprocal and square root spelled as its branch-free core (nb_fastmath.cuh), so
// that the three independent chains - direction (two quotients, three roots, the rotation),
// speed, and cross sections (interpolation weight, Sigma, mean free path) - interleave
// instead of running one division after the other.
// Returns false when an operand left the range in which the cores are the IEEE result (or
// the step needs something else the block does not handle); nothing has been committed
// then and the caller redoes the event with the plain operators.
// ------------------------------------------------------------------------------------
// The block writes straight into the particle's state; on `false` that state is garbage and
// the caller restores the direction it parked, and redoes the event from its inputs.
__device__ __forceinline__ bool scatter_fast_same_grid(const StepArgs& a, double e, double a1,
                                                       double nd, double neglog, int cx, int cy,
                                                       double& ox, double& oy, double& v,
                                                       double& uxi, double& uyi, double& ex,
                                                       double& ey, double& mfp, Derived& d,
                                                       double& e_out, double& Sig_s_out,
                                                       double& p_absorb_out) {
  constexpr double kA1sq = (kMassNo + 1.0) * (kMassNo + 1.0);
  // :264-267 - the energy after the scatter
  const double mu = 1.0 - 2.0 * a1;
  const double e_num = e * ((kMassNo * kMassNo + (2.0 * kMassNo) * mu) + 1.0);
  const double e_new = div_by_known_unchecked(e_num, kA1sq, 1.0 / kA1sq);
  bool ok = fm_safe(e) & fm_safe(e_num) & fm_safe(e_new);
  // :285-286 - bracketing interval of the new energy (a short loop of probes), then the loads
  const int is = cs_bracket_staged(a.cs_s, e_new);
  const double2 s0 = __ldg(a.cs_s.kv + is), s1 = __ldg(a.cs_s.kv + is + 1);
  const double va0 = __ldg(&a.cs_a.kv[is].y), va1 = __ldg(&a.cs_a.kv[is + 1].y);
  // :270-279 - direction chain
  const double r_up = div_core(e_new, e);
  const double r_dn = div_core(e, e_new);
  const double ct = 0.5 * ((kMassNo + 1.0) * sqrt_core(r_up) - (kMassNo - 1.0) * sqrt_core(r_dn));
  const double om = 1.0 - ct * ct;
  ok &= fm_safe_positive(om);
  const double st = sqrt_core(om);
  const double nox = ox * ct - oy * st;
  const double noy = ox * st + oy * ct;
  ox = nox;
  oy = noy;
  // :297 - speed chain
  const double v_num = (2.0 * e_new) * kEvToJ;
  const double v_arg = div_by_known_unchecked(v_num, kParticleMass, 1.0 / kParticleMass);
  ok &= fm_safe(v_num) & fm_safe_positive(v_arg);
  v = sqrt_core(v_arg);
  // :285-291, 135, 231-232, 481-491 - cross-section chain. An energy exactly on a grid point
  // gives fa = +0 and the core returns the +0 the division would.
  const double fa = e_new - s0.x, fb = s1.x - s0.x;
  ok &= (fm_safe(fa) | (fa == 0.0)) & fm_safe(fb);
  const double frac = div_core(fa, fb);
  const double sig_s = s0.y + frac * (s1.y - s0.y);
  const double sig_a = va0 + frac * (va1 - va0);
  const double sig_t = sig_s + sig_a;
  const double S_s = macroscopic(nd, sig_s);
  const double S_a = macroscopic(nd, sig_a);
  const double S_t = S_s + S_a;
  double q, pa;
  if (NB_HALF_SHORTCUT && double_to_bits(sig_s) == double_to_bits(sig_a) &&
      finite_nonzero(sig_t) && finite_nonzero(S_t)) {
    q = 0.5;
    pa = 0.5;
  } else {
    ok &= fm_safe(sig_a) & fm_safe(sig_t) & fm_safe(S_a);
    q = div_core(sig_a, sig_t);
    pa = div_core(S_a, S_t);
  }
  d.stb = sig_t * kBarns;
  d.heat = heating_response_q(e_new, q);
  const double wx = nox * v, wy = noy * v;
  ok &= fm_safe(S_t) & fm_safe(S_s) & fm_safe(neglog) & fm_safe(wx) & fm_safe(wy);
  d.cell_mfp = rcp_core(S_t);
  uxi = rcp_core(wx);  // calc_distance_to_facet's reciprocals, :435-436
  uyi = rcp_core(wy);
  mfp = div_core(neglog, S_s);  // :294-295
  ex = target_edge(a.edgex, cx, axis_step(nox));
  ey = target_edge(a.edgey, cy, axis_step(noy));
  e_out = e_new;
  Sig_s_out = S_s;
  p_absorb_out = pa;
  return ok;
}

// update_tallies (omp3/neutral.c:408-420). With kPreReduce the lanes of the warp that flush
// into the same cell at the same moment are combined with shuffles first and their leader
// issues one atomic (north_star phase 5). Off by default: measured, it loses - see DESIGN.md 4.
template <bool kPreReduce>
__device__ __forceinline__ void tally_add(double* __restrict__ tally, int cell, double value) {
#ifdef NB_PROBE_TALLY
  // Measurement builds (never the product; results are wrong): what does the reduction cost
  // the event loop? 1: the value is computed and dropped; 2: a plain store instead of the
  // reduction (same address stream, no read-modify-write at the L2).
  //   NB200_DEFINES=-DNB_PROBE_TALLY=1 NB200_LIB=libneutral_b200.probe1.so python -m neutral_b200.build
  if (NB_PROBE_TALLY == 1) {
    asm volatile("" ::"d"(value), "r"(cell));
  } else {
    tally[cell] = value;
  }
  return;
#endif
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
