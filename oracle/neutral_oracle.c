/*
 * neutral_oracle.c - TEST INFRASTRUCTURE, not product code.
 *
 * A CPU restatement, in plain C, of the particle-history hot path of UoB-HPC/neutral's
 * omp3 kernel set. It exists so that tests/ and bench.py's cpu_baseline leg can check the
 * CUDA path; nothing under neutral_b200/ may call, link or import it.
 *
 * Parity pinning: this port is checked bit-for-bit (every particle field after every
 * timestep, aggregate event counts, tally to 1e-12) against the UNMODIFIED reference
 * omp3/neutral.c compiled from /root/reference (oracle/_ref/libneutral_omp3.so, see
 * oracle/Makefile) by tests/test_oracle.py and tests/test_variants.py, against committed golden vectors
 * generated from that reference build (tests/golden/), and against the Random123
 * known-answer vectors. What it adds over the reference: SoA banks, a global-pid offset
 * (so a shard of the bank replays the same histories) and per-particle event counters.
 *
 * Arithmetic is IEEE binary64, no contraction (build with -ffp-contract=off), expression
 * order exactly as the reference writes it; libm log/sqrt/sin/cos as the reference links.
 * Every function cites the reference lines it restates (paths relative to the reference
 * root).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* neutral_data.h:17-24 */
#define O_EV_TO_J 1.60217646e-19
#define O_AVOGADROS 6.02214085774e23
#define O_BARNS 1.0e-28
#define O_PARTICLE_MASS 1.674927471213e-27
#define O_MASS_NO 1.0e2
#define O_MOLAR_MASS 1.0e-2
#define O_MIN_ENERGY 1.0e0
#define O_OPEN_BOUND 1.0e-13

/* Same member order as the reference's -DSoA Particle (neutral_data.h:48-61). */
typedef struct {
  double* x;
  double* y;
  double* omega_x;
  double* omega_y;
  double* energy;
  double* weight;
  double* dt_to_census;
  double* mfp_to_collision;
  int* cellx;
  int* celly;
  int* dead;
} OracleBank;

/* ---- Threefry-2x64, 20 rounds (Random123/threefry.h:86-93,170-171,196-286) ---------- */

static inline uint64_t rotl64(uint64_t v, unsigned r) {
  return (v << r) | (v >> (64u - r));
}

void oracle_threefry2x64_20(uint64_t c0, uint64_t c1, uint64_t k0, uint64_t k1,
                            uint64_t out[2]) {
  static const unsigned rot[8] = {16, 42, 12, 31, 16, 32, 24, 21};
  const uint64_t ks[3] = {k0, k1, UINT64_C(0x1BD11BDAA9FC1A22) ^ k0 ^ k1};
  uint64_t a = c0 + ks[0];
  uint64_t b = c1 + ks[1];
  for (unsigned round = 0; round < 20; ++round) {
    a += b;
    b = rotl64(b, rot[round & 7u]) ^ a;
    if ((round & 3u) == 3u) {
      /* key injection number s = (round+1)/4 */
      const unsigned s = (round + 1u) >> 2;
      a += ks[s % 3u];
      b += ks[(s + 1u) % 3u] + (uint64_t)s;
    }
  }
  out[0] = a;
  out[1] = b;
}

/* omp3/neutral.c:632-652: ctr = {counter, 0}, key = {pkey, master_key}; uniform in (0,1]. */
void oracle_random_pair(uint64_t pkey, uint64_t master_key, uint64_t counter,
                        double* rn0, double* rn1) {
  uint64_t r[2];
  oracle_threefry2x64_20(counter, 0, pkey, master_key, r);
  const double factor = 1.0 / (18446744073709551615.0 + 1.0);
  const double half_factor = 0.5 * factor;
  *rn0 = (double)r[0] * factor + half_factor;
  *rn1 = (double)r[1] * factor + half_factor;
}

/* ---- cross sections (omp3/neutral.c:498-517) ----------------------------------------- */

/* Index ind with keys[ind] <= energy < keys[ind+1]. The reference walks with a halving
 * stride; the bracketing interval of a strictly increasing table is unique, so a plain
 * bisection returns the same ind. Energies outside the table (never reached by the decks;
 * undefined behaviour in the reference) clamp to the end intervals. */
int oracle_cs_index(const double* keys, int nentries, double energy) {
  int lo = 0, hi = nentries - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (energy < keys[mid]) hi = mid; else lo = mid;
  }
  return lo;
}

double oracle_cs_lookup(const double* keys, const double* values, int nentries,
                        double energy) {
  const int ind = oracle_cs_index(keys, nentries, energy);
  return values[ind] + ((energy - keys[ind]) / (keys[ind + 1] - keys[ind])) *
                           (values[ind + 1] - values[ind]);
}

/* ---- helpers --------------------------------------------------------------------------*/

/* omp3/neutral.c:474-495: path-length heating estimator. */
static inline double energy_deposition(double energy, double weight, double path,
                                       double number_density, double sig_a,
                                       double sig_t) {
  const double absorb_heat = (sig_a / sig_t) * 0.0;
  const double exit_scatter =
      energy * ((O_MASS_NO * O_MASS_NO + O_MASS_NO + 1) /
                ((O_MASS_NO + 1) * (O_MASS_NO + 1)));
  const double scatter_heat = (1.0 - (sig_a / sig_t)) * exit_scatter;
  const double response = (energy - scatter_heat - absorb_heat);
  return weight * path * (sig_t * O_BARNS) * response * number_density;
}

static inline double speed_of(double energy) { /* omp3/neutral.c:117,297 */
  return sqrt((2.0 * energy * O_EV_TO_J) / O_PARTICLE_MASS);
}

/* ---- injection (omp3/neutral.c:560-630; source box from neutral_data.c:65-95) --------- */

/* Particles [pid0, pid0+count) of the global bank go to slots [0, count). */
void oracle_inject(int pid0, int count, int nx, int ny, double left, double bottom,
                   double width, double height, double dt, const double* edgex,
                   const double* edgey, double initial_energy, OracleBank* bank) {
#pragma omp parallel for schedule(static)
  for (int s = 0; s < count; ++s) {
    const uint64_t pid = (uint64_t)(pid0 + s);
    double r0, r1;
    oracle_random_pair(pid, 0, 0, &r0, &r1);
    const double px = left + r0 * width;
    const double py = bottom + r1 * height;
    int cx = 0, cy = 0;
    for (int i = 0; i < nx; ++i) {
      if (px >= edgex[i] && px < edgex[i + 1]) { cx = i; break; }
    }
    for (int j = 0; j < ny; ++j) {
      if (py >= edgey[j] && py < edgey[j + 1]) { cy = j; break; }
    }
    oracle_random_pair(pid, 0, 1, &r0, &r1);
    const double theta = 2.0 * M_PI * r0;
    bank->x[s] = px;
    bank->y[s] = py;
    bank->cellx[s] = cx;
    bank->celly[s] = cy;
    bank->omega_x[s] = cos(theta);
    bank->omega_y[s] = sin(theta);
    bank->energy[s] = initial_energy;
    bank->weight[s] = 1.0;
    bank->dt_to_census[s] = dt;
    bank->mfp_to_collision[s] = 0.0;
    bank->dead[s] = 0;
  }
}

/* ---- one timestep (omp3/neutral.c:43-206 and the event functions it calls) ------------ */

/*
 * Transports slots [0, count) whose global particle ids are pid0 + slot, with pad = 0 and
 * x_off = y_off = 0 (the only configuration the driver produces, main.c:34,101-110).
 * totals[0..2] += facets, collisions, particles processed; per-particle counters (may be
 * NULL) accumulate facets / collisions / census events per slot.
 */
void oracle_transport_step(int nx, int ny, uint64_t master_key, double dt,
                           int ntotal_particles, int pid0, int count,
                           OracleBank* bank, const double* density,
                           const double* edgex, const double* edgey,
                           const double* s_keys, const double* s_values, int s_n,
                           const double* a_keys, const double* a_values, int a_n,
                           double* tally, uint64_t* p_facets,
                           uint64_t* p_collisions, uint64_t* p_census,
                           uint64_t totals[3]) {
  uint64_t tot_f = 0, tot_c = 0, tot_p = 0;
  const double inv_ntotal = 1.0 / (double)ntotal_particles; /* :120 */

#pragma omp parallel for schedule(dynamic, 64) reduction(+ : tot_f, tot_c, tot_p)
  for (int s = 0; s < count; ++s) {
    if (bank->dead[s]) continue; /* :91-93 */
    tot_p++;

    const uint64_t pkey = (uint64_t)(pid0 + s);
    double x = bank->x[s], y = bank->y[s];
    double ox = bank->omega_x[s], oy = bank->omega_y[s];
    double e = bank->energy[s], w = bank->weight[s];
    int cx = bank->cellx[s], cy = bank->celly[s];
    int dead = 0;
    uint64_t nf = 0, nc = 0, nz = 0;

    /* :103-118 */
    double rho = density[(size_t)cy * nx + cx];
    double sig_s = oracle_cs_lookup(s_keys, s_values, s_n, e);
    double sig_a = oracle_cs_lookup(a_keys, a_values, a_n, e);
    double nd = (rho * O_AVOGADROS / O_MOLAR_MASS);
    double Sig_s = nd * sig_s * O_BARNS;
    double Sig_a = nd * sig_a * O_BARNS;
    double v = speed_of(e);
    double edep = 0.0;

    /* :122-131 - `initial` is always 1 (omp3/neutral.c:36) */
    uint64_t counter = 0;
    double r0, r1;
    double dtc = dt;
    oracle_random_pair(pkey, master_key, counter++, &r0, &r1);
    double mfp = -log(r0) / Sig_s;

    while (dtc > 0.0) { /* :134 */
      const double cell_mfp = 1.0 / (Sig_s + Sig_a);

      /* calc_distance_to_facet, :423-471 */
      const double uxi = 1.0 / (ox * v);
      const double uyi = 1.0 / (oy * v);
      const double tx = (ox >= 0.0) ? ((edgex[cx + 1]) - x) * uxi
                                    : ((edgex[cx] - O_OPEN_BOUND) - x) * uxi;
      const double ty = (oy >= 0.0) ? ((edgey[cy + 1]) - y) * uyi
                                    : ((edgey[cy] - O_OPEN_BOUND) - y) * uyi;
      const int x_facet = (tx < ty);
      double d_facet;
      if (x_facet) {
        d_facet = (ox >= 0.0) ? ((edgex[cx + 1]) - x) * v * uxi
                              : ((edgex[cx] - O_OPEN_BOUND) - x) * v * uxi;
      } else {
        d_facet = (oy >= 0.0) ? ((edgey[cy + 1]) - y) * v * uyi
                              : ((edgey[cy] - O_OPEN_BOUND) - y) * v * uyi;
      }

      const double d_coll = mfp * cell_mfp; /* :144-146 */
      const double d_census = v * dtc;

      if (d_coll < d_facet && d_coll < d_census) {
        /* collision_event, :209-300 */
        nc++;
        edep += energy_deposition(e, w, d_coll, nd, sig_a, sig_s + sig_a);
        x += d_coll * ox;
        y += d_coll * oy;
        const double p_absorb = Sig_a / (Sig_s + Sig_a);
        double a0, a1;
        oracle_random_pair(pkey, master_key, counter++, &a0, &a1);
        if (a0 < p_absorb) {
          w *= (1.0 - p_absorb);
          if (e < O_MIN_ENERGY) {
            dead = 1;
#pragma omp atomic update
            tally[(size_t)cy * nx + cx] += edep * inv_ntotal;
            edep = 0.0;
            break;
          }
        } else {
          const double mu = 1.0 - 2.0 * a1;
          const double e_new =
              e * (O_MASS_NO * O_MASS_NO + 2.0 * O_MASS_NO * mu + 1.0) /
              ((O_MASS_NO + 1.0) * (O_MASS_NO + 1.0));
          const double ct = 0.5 * ((O_MASS_NO + 1.0) * sqrt(e_new / e) -
                                   (O_MASS_NO - 1.0) * sqrt(e / e_new));
          const double st = sqrt(1.0 - ct * ct);
          const double nox = (ox * ct - oy * st);
          const double noy = (ox * st + oy * ct);
          ox = nox;
          oy = noy;
          e = e_new;
        }
        sig_s = oracle_cs_lookup(s_keys, s_values, s_n, e);
        sig_a = oracle_cs_lookup(a_keys, a_values, a_n, e);
        nd = (rho * O_AVOGADROS / O_MOLAR_MASS);
        Sig_s = nd * sig_s * O_BARNS;
        Sig_a = nd * sig_a * O_BARNS;
        oracle_random_pair(pkey, master_key, counter++, &r0, &r1);
        mfp = -log(r0) / Sig_s;
        dtc -= d_coll / v;
        v = speed_of(e);
      } else if (d_facet < d_census) {
        /* facet_event, :303-380 */
        nf++;
        mfp -= (d_facet / cell_mfp);
        dtc -= (d_facet / v);
        edep += energy_deposition(e, w, d_facet, nd, sig_a, sig_s + sig_a);
#pragma omp atomic update
        tally[(size_t)cy * nx + cx] += edep * inv_ntotal;
        edep = 0.0;
        x += d_facet * ox;
        y += d_facet * oy;
        if (x_facet) {
          if (ox > 0.0) {
            if (cx >= nx - 1) ox = -ox; else cx++;
          } else if (ox < 0.0) {
            if (cx <= 0) ox = -ox; else cx--;
          }
        } else {
          if (oy > 0.0) {
            if (cy >= ny - 1) oy = -oy; else cy++;
          } else if (oy < 0.0) {
            if (cy <= 0) oy = -oy; else cy--;
          }
        }
        rho = density[(size_t)cy * nx + cx];
        nd = (rho * O_AVOGADROS / O_MOLAR_MASS);
        Sig_s = nd * sig_s * O_BARNS;
        Sig_a = nd * sig_a * O_BARNS;
      } else {
        /* census_event, :383-405 */
        nz++;
        x += d_census * ox;
        y += d_census * oy;
        mfp -= (d_census / cell_mfp);
        edep += energy_deposition(e, w, d_census, nd, sig_a, sig_s + sig_a);
#pragma omp atomic update
        tally[(size_t)cy * nx + cx] += edep * inv_ntotal;
        dtc = 0.0;
        break;
      }
    }

    bank->x[s] = x;
    bank->y[s] = y;
    bank->omega_x[s] = ox;
    bank->omega_y[s] = oy;
    bank->energy[s] = e;
    bank->weight[s] = w;
    bank->dt_to_census[s] = dtc;
    bank->mfp_to_collision[s] = mfp;
    bank->cellx[s] = cx;
    bank->celly[s] = cy;
    bank->dead[s] = dead;
    if (p_facets) p_facets[s] += nf;
    if (p_collisions) p_collisions[s] += nc;
    if (p_census) p_census[s] += nz;
    tot_f += nf;
    tot_c += nc;
  }
  totals[0] += tot_f;
  totals[1] += tot_c;
  totals[2] += tot_p;
}

/* Sum of the tally in index order (validate, omp3/neutral.c:523-527). */
double oracle_tally_sum(const double* tally, size_t n) {
  double s = 0.0;
  for (size_t i = 0; i < n; ++i) s += tally[i];
  return s;
}
