// history.cu - the event loop of the b200 kernel set (sm_100a): phase P4 of pipeline.cu.
//
// k_history follows every live particle of the sorted bank from the first event of the
// timestep to its census or death. It restates the loop of handle_particles
// (omp3/neutral.c:134-197) and the three event functions
//   collision_event  omp3/neutral.c:209-300     facet_event   omp3/neutral.c:303-380
//   census_event     omp3/neutral.c:383-405     (+ :408-420, :423-471, :474-495, :498-517)
// with these differences in HOW, none in WHAT is computed:
//
//  * everything that is invariant between two events of one kind lives in registers and is
//    recomputed - with the reference's own expression - only when one of its inputs changes
//    (energy: collisions; density: tile crossings), so the facet iteration is ~25 FP64
//    operations instead of the reference's ~150 including 7 divides;
//  * divisions by the two loop-invariant divisors of a facet (speed, cell mean free path)
//    go through their correctly rounded reciprocals with a Markstein correction - the IEEE
//    quotient, bit for bit (nb_device.cuh: div_by_known);
//  * the cell step/reflect logic (:332-368) is computed on the crossed axis only, without
//    the reference's four-way branch, so a warp does not serialise over directions;
//  * the density of the entered cell comes from the per-step tile map (stage.cu) unless the
//    tile is mixed; cross sections come from the staged tables through the bucket index;
//  * the tally index cy*nx+cx is carried along instead of being rebuilt.
//
// Compiled with -fmad=false: the only fused operations are the explicit fma() of nb_log and
// div_by_known.
#include "nb_device.cuh"
#include "transport.cuh"

namespace nb {

#ifndef NB_HISTORY_MIN_BLOCKS
#define NB_HISTORY_MIN_BLOCKS 6
#endif

__device__ __forceinline__ bool is_mixed_tile(double t) {
  return double_to_bits(t) == kMixedTileBits;
}

__device__ __forceinline__ double tile_value(const StepArgs& a, int cx, int cy) {
  return __ldg(a.tiles.tile_rho + (cy >> kTileShift) * a.tiles.tiles_x + (cx >> kTileShift));
}

// Per-particle flag bits kept in one register.
enum : unsigned {
  kFlagSpeedOk = 1u,    // the speed is inside the proven range of div_by_known
  kFlagCellMfpOk = 2u,  // so is the cell mean free path
  kFlagMixedTile = 4u,  // the current tile is not uniform: densities come from the mesh
  kFlagDead = 8u,
};

// Everything that follows from the energy and the cell density (omp3/neutral.c:112-117,
// 135; 231-232; 481-491), grouped so that the facet loop only carries what it reads.
struct Derived {
  double stb;           // (sigma_s + sigma_a) * BARNS, the deposition's cross section
  double heat;          // heating response of :481-491
  double Sig_s;         // macroscopic scattering cross section (mean-free-path sampling)
  double p_absorb;      // Sigma_a / Sigma_t, :231-232
  double cell_mfp;      // 1 / Sigma_t, :135
  double cell_mfp_inv;  // correctly rounded 1 / cell_mfp for div_by_known
};

__device__ __forceinline__ void derive(const StepArgs& a, double e, double nd, Derived& d,
                                       unsigned& flags) {
  double sig_s, sig_a;
  cs_lookup_pair_staged(a, e, sig_s, sig_a);
  const double sig_t = sig_s + sig_a;
  d.stb = sig_t * kBarns;
  d.heat = heating_response(e, sig_a, sig_t);
  d.Sig_s = macroscopic(nd, sig_s);
  const double Sig_a = macroscopic(nd, sig_a);
  const double Sig_t = d.Sig_s + Sig_a;
  d.p_absorb = Sig_a / Sig_t;
  d.cell_mfp = 1.0 / Sig_t;
  d.cell_mfp_inv = 1.0 / d.cell_mfp;
  flags = safe_exponent(d.cell_mfp) ? (flags | kFlagCellMfpOk) : (flags & ~kFlagCellMfpOk);
}

template <bool kFastDiv>
__global__ void __launch_bounds__(kHistoryThreads, NB_HISTORY_MIN_BLOCKS)
k_history(const StepArgs a, const unsigned* __restrict__ n_live) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned nf = 0, nc = 0, census = 0, processed = 0, died = 0;

  int4 m = make_int4(0, 0, 1, 0);
  if (slot < (int)*n_live) m = a.bank.meta[slot];

  if (!m.z) {
    processed = 1;
    const double2 pos = a.bank.pos[slot];
    const double2 dir = a.bank.dir[slot];
    const double2 ew = a.bank.ew[slot];
    const double2 tm = a.bank.tm[slot];  // k_begin_step left (dt, first path sample) here
    double x = pos.x, y = pos.y, ox = dir.x, oy = dir.y, e = ew.x, w = ew.y;
    double dtc = tm.x, mfp = tm.y;
    int cx = m.x, cy = m.y;
    unsigned counter = 1;  // counter 0 was drawn by k_begin_step (omp3/neutral.c:129)
    unsigned flags = 0;
    double edep = 0.0;

    // ---- the density of the cell (changes on tile crossings only) ...
    double rho = tile_value(a, cx, cy);
    if (is_mixed_tile(rho)) {
      flags |= kFlagMixedTile;
      rho = __ldg(a.density + cy * a.nx + cx);
    }
    double nd = number_density(rho);
    // ---- ... and what follows from the energy (changes on collisions only)
    Derived d;
    derive(a, e, nd, d, flags);
    double v = speed_of(e);
    double v_inv = 1.0 / v;
    if (safe_exponent(v)) flags |= kFlagSpeedOk;
    double uxi = 1.0 / (ox * v);
    double uyi = 1.0 / (oy * v);

    while (dtc > 0.0) {
      // calc_distance_to_facet, omp3/neutral.c:423-471
      const bool xup = ox >= 0.0, yup = oy >= 0.0;
      double ex = __ldg(a.edgex + cx + (xup ? 1 : 0));
      double ey = __ldg(a.edgey + cy + (yup ? 1 : 0));
      if (!xup) ex -= kOpenBoundCorrection;
      if (!yup) ey -= kOpenBoundCorrection;
      const double gx = ex - x;
      const double gy = ey - y;
      const bool x_facet = (gx * uxi) < (gy * uyi);
      const double d_facet = ((x_facet ? gx : gy) * v) * (x_facet ? uxi : uyi);
      const double d_coll = mfp * d.cell_mfp;  // :144-146
      const double d_census = v * dtc;
      const bool collide = d_coll < d_facet && d_coll < d_census;

      if (!collide && d_facet < d_census) {
        // ---- facet_event, :303-380
        nf++;
        double q_mfp, q_dtc;
        if (kFastDiv && (flags & (kFlagSpeedOk | kFlagCellMfpOk)) ==
                            (kFlagSpeedOk | kFlagCellMfpOk) && safe_exponent(d_facet)) {
          q_mfp = div_by_known_unchecked(d_facet, d.cell_mfp, d.cell_mfp_inv);
          q_dtc = div_by_known_unchecked(d_facet, v, v_inv);
        } else {
          q_mfp = d_facet / d.cell_mfp;
          q_dtc = d_facet / v;
        }
        mfp -= q_mfp;
        dtc -= q_dtc;
        edep += deposition(w, d_facet, d.stb, d.heat, nd);
        atomicAdd(a.tally + cy * a.nx + cx, edep * a.inv_ntotal);  // update_tallies, :408-420
        edep = 0.0;
        x += d_facet * ox;
        y += d_facet * oy;
        // :332-368 on the crossed axis: step one cell, or reflect at the mesh boundary
        const double o = x_facet ? ox : oy;
        const int c = x_facet ? cx : cy;
        const int step = (o > 0.0) ? 1 : ((o < 0.0) ? -1 : 0);
        const int cn = c + step;
        if ((unsigned)cn >= (unsigned)(x_facet ? a.nx : a.ny)) {
          if (x_facet) { ox = -ox; uxi = -uxi; } else { oy = -oy; uyi = -uyi; }
        } else {
          if (x_facet) cx = cn; else cy = cn;
          // :372-378 - the macroscopic cross sections follow the density of the new cell.
          // Inside a uniform tile the density cannot change; a new tile is looked up in the
          // tile map; only mixed tiles read the density mesh itself.
          double rho_new = rho;
          if ((cn ^ c) >> kTileShift) {
            rho_new = tile_value(a, cx, cy);
            flags = is_mixed_tile(rho_new) ? (flags | kFlagMixedTile) : (flags & ~kFlagMixedTile);
          }
          if (flags & kFlagMixedTile) rho_new = __ldg(a.density + cy * a.nx + cx);
          if (double_to_bits(rho_new) != double_to_bits(rho)) {
            rho = rho_new;
            nd = number_density(rho);
            derive(a, e, nd, d, flags);
          }
        }
      } else if (collide) {
        // ---- collision_event, :209-300
        nc++;
        const uint64_t pkey = a.pid0 + (uint64_t)(unsigned)m.w;
        edep += deposition(w, d_coll, d.stb, d.heat, nd);
        x += d_coll * ox;
        y += d_coll * oy;
        double a0, a1;
        random_pair(pkey, a.master_key, counter++, a0, a1);
        // :296 - the time to census shrinks by the flight time at the pre-collision speed
        double q_dtc;
        if (kFastDiv && (flags & kFlagSpeedOk) && safe_exponent(d_coll))
          q_dtc = div_by_known_unchecked(d_coll, v, v_inv);
        else
          q_dtc = d_coll / v;
        if (a0 < d.p_absorb) {
          w *= (1.0 - d.p_absorb);
          if (e < kMinEnergyOfInterest) {
            flags |= kFlagDead;
            atomicAdd(a.tally + cy * a.nx + cx, edep * a.inv_ntotal);
            edep = 0.0;
            break;
          }
          // Energy and direction are unchanged: the lookups of :285-291 return the values
          // already held, so nothing that depends on them needs recomputing.
        } else {
          const double mu = 1.0 - 2.0 * a1;
          const double e_new = (e * ((kMassNo * kMassNo + (2.0 * kMassNo) * mu) + 1.0)) /
                               ((kMassNo + 1.0) * (kMassNo + 1.0));
          const double ct = 0.5 * ((kMassNo + 1.0) * sqrt(e_new / e) -
                                   (kMassNo - 1.0) * sqrt(e / e_new));
          const double st = sqrt(1.0 - ct * ct);
          const double nox = ox * ct - oy * st;
          const double noy = ox * st + oy * ct;
          ox = nox;
          oy = noy;
          e = e_new;
          derive(a, e, nd, d, flags);
          v = speed_of(e);
          v_inv = 1.0 / v;
          flags = safe_exponent(v) ? (flags | kFlagSpeedOk) : (flags & ~kFlagSpeedOk);
          uxi = 1.0 / (ox * v);
          uyi = 1.0 / (oy * v);
        }
        mfp = -nb_log(random_first(pkey, a.master_key, counter++), a.logt) / d.Sig_s;
        dtc -= q_dtc;
      } else {
        // ---- census_event, :383-405
        census = 1;
        x += d_census * ox;
        y += d_census * oy;
        mfp -= d_census / d.cell_mfp;
        edep += deposition(w, d_census, d.stb, d.heat, nd);
        atomicAdd(a.tally + cy * a.nx + cx, edep * a.inv_ntotal);
        dtc = 0.0;
        break;
      }
    }

    died = (flags & kFlagDead) ? 1u : 0u;
    a.bank.pos[slot] = make_double2(x, y);
    a.bank.dir[slot] = make_double2(ox, oy);
    a.bank.ew[slot] = make_double2(e, w);
    a.bank.tm[slot] = make_double2(dtc, mfp);
    a.bank.meta[slot] = make_int4(cx, cy, (int)died, m.w);
    if (a.p_facets) a.p_facets[m.w] += nf;
    if (a.p_collisions) a.p_collisions[m.w] += nc;
    if (a.p_census) a.p_census[m.w] += census;
  }
  flush_totals(a.totals, nf, nc, processed, census, died);
}

static inline int blocks_for(int n, int threads) { return (n + threads - 1) / threads; }

int launch_history(const StepArgs& a, const unsigned* n_live, int n_upper, bool fast_div,
                   cudaStream_t st) {
  if (n_upper <= 0) return 0;
  const int blocks = blocks_for(n_upper, kHistoryThreads);
  if (fast_div)
    k_history<true><<<blocks, kHistoryThreads, 0, st>>>(a, n_live);
  else
    k_history<false><<<blocks, kHistoryThreads, 0, st>>>(a, n_live);
  return 1;
}

}  // namespace nb
