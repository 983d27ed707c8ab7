// pipeline.cu - the phased timestep of the b200 kernel set (sm_100a).
//
// One timestep of solve_transport_2d (reference omp3/neutral.c:19-206) is split into
// event-based phases over the packed-pair SoA bank:
//
//   P1 k_begin_step   per live particle: start-of-step set-up of omp3/neutral.c:103-131
//                     (density, cross sections, first Threefry draw, mean-free-path sample),
//                     classification by NEXT EVENT TYPE (collision / facet / census) and
//                     mesh tile, histogram of the sort keys (warp-aggregated atomics).
//   P2 k_scan_bins    exclusive scan of the histogram -> bin offsets.
//   P3 k_scatter      counting sort of the bank into its double buffer: colliders first,
//                     then streamers, then census-only particles; inside a class by
//                     expected history length (longest first, so that the lanes of a warp
//                     finish together and the long histories start early) and mesh tile;
//                     dead particles compacted to the tail. Warp ballot/match aggregation.
//   P4 k_history      (history.cu) event loop to census/death on event-type-coherent warps.
//
// Bit-exactness: reordering the bank cannot change a history (RNG streams are keyed by the
// particle's origin, omp3/neutral.c:89,632-641) and every cached quantity is recomputed with
// the reference's own expression whenever one of its inputs changes, so the values are the
// ones the reference would recompute each iteration.
#include "nb_device.cuh"
#include "nb_fastmath.cuh"
#include "transport.cuh"

namespace nb {

enum { kClsCollision = 0, kClsFacet = 1, kClsCensus = 2, kNumClasses = 3 };

// Warp-aggregated histogram increment / slot claim: lanes with the same key elect a leader
// (match), the leader adds the group size with one atomic and the members take consecutive
// ranks from a ballot-style prefix count.
__device__ __forceinline__ unsigned claim_in_bin(unsigned* counters, unsigned key,
                                                 bool active) {
  const unsigned lane = threadIdx.x & 31;
  unsigned result = 0;
  const unsigned amask = __ballot_sync(0xffffffffu, active);
  if (active) {
    const unsigned peers = __match_any_sync(amask, key);
    const unsigned leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(counters + key, (unsigned)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    result = base + __popc(peers & ((1u << lane) - 1));
  }
  return result;
}

// ------------------------------------------------------------------------------------
// P1: start-of-step set-up + classification + histogram
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_begin_step(const StepArgs a, const SortArgs s) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = slot < s.n_upper;
  unsigned key = s.nbins - 1;  // dead bin
  if (in_range) {
    const int4 m = a.bank.meta[slot];
    if (!m.z) {
      const uint64_t pkey = a.pid0 + (uint64_t)(unsigned)m.w;
      const double2 pos = a.bank.pos[slot];
      const double2 dir = a.bank.dir[slot];
      const double e = a.bank.ew[slot].x;
      const int cx = m.x, cy = m.y;
      const double rho = __ldg(a.density + (size_t)cy * a.nx + cx);
      double sig_s, sig_a;
      cs_lookup_pair_staged(a, e, same_grid(a), sig_s, sig_a);
      const double nd = number_density(rho);
      const double Sig_s = macroscopic(nd, sig_s);
      const double Sig_a = macroscopic(nd, sig_a);
      const double v = speed_of(e);
      // omp3/neutral.c:127-131: dt_to_census = dt, counter 0 draws the first path sample
      const double mfp = -nb_log(random_first(pkey, a.master_key, 0), a.logt) / Sig_s;
      a.bank.tm[slot] = make_double2(a.dt, mfp);

      // first iteration of the event loop, decision only (:135-150,170)
      const double cell_mfp = 1.0 / (Sig_s + Sig_a);
      const double uxi = 1.0 / (dir.x * v);
      const double uyi = 1.0 / (dir.y * v);
      const double gx = (dir.x >= 0.0) ? (__ldg(a.edgex + cx + 1) - pos.x)
                                       : ((__ldg(a.edgex + cx) - kOpenBoundCorrection) - pos.x);
      const double gy = (dir.y >= 0.0) ? (__ldg(a.edgey + cy + 1) - pos.y)
                                       : ((__ldg(a.edgey + cy) - kOpenBoundCorrection) - pos.y);
      const bool x_facet = (gx * uxi) < (gy * uyi);
      const double d_facet = x_facet ? (gx * v) * uxi : (gy * v) * uyi;
      const double d_coll = mfp * cell_mfp;
      const double d_census = v * a.dt;
      // The class is the kind of work the history is about to do, which is the first event's
      // kind except for a particle of a dense cell that happens to sit next to a cell edge: its
      // first event is a facet, the ~1000 collisions of a collider follow (the path sample
      // carries over into the next cell, :282). Sorted among the streamers it would hold a
      // warp of finished lanes for a collider's lifetime - the split deck's launch ended with
      // 1.2 ms of a dozen such warps (profiles/r02/warp_trace_r2b_split.txt). It is a collider
      // when the sampled collision lies inside this timestep and less than ~8 cells ahead.
      const float cells_per_length =
          fabsf((float)dir.x) * s.inv_dx + fabsf((float)dir.y) * s.inv_dy;
#ifdef NB_CLASS_FIRST_EVENT_ONLY  // A/B build: round 1's rule (the first event's kind)
      const bool collides_soon = false;
#else
      const bool collides_soon = d_coll < d_census && (float)d_coll * cells_per_length < 8.0f;
#endif
      const int cls = ((d_coll < d_facet && d_coll < d_census) || collides_soon) ? kClsCollision
                      : (d_facet < d_census)                                     ? kClsFacet
                                                                                 : kClsCensus;
      const unsigned tile = s.tile_shift >= 0
                                ? (unsigned)((cy >> s.tile_shift) * s.tiles_x + (cx >> s.tile_shift))
                                : 0u;
      // Expected number of events of this history until census, a scheduling hint only
      // (single precision, never feeds back into a particle): streamers cross about
      // (|omega_x|/dx + |omega_y|/dy) * speed * dt facets; a collider scatters its energy
      // down by ~2 % per scatter (omp3/neutral.c:265-267 averaged over mu) and is absorbed
      // as often as it scatters until it falls below MIN_ENERGY_OF_INTEREST (:243).
      unsigned qbin = 0;
      if (s.nq > 1) {
        float est = 1.0f;
        if (cls == kClsFacet)
          est += cells_per_length * (float)v * (float)a.dt;
        else if (cls == kClsCollision)
          est += 101.0f * __logf(fmaxf((float)e, 1.0f));
        const int q = (int)(__log2f(est) * s.q_scale);
        qbin = (unsigned)(s.nq - 1 - min(max(q, 0), s.nq - 1));  // longest histories first
      }
      key = ((unsigned)cls * (unsigned)s.nq + qbin) * (unsigned)s.ntiles + tile;
    }
    s.keys[slot] = key;
  }
  claim_in_bin(s.bin_count, key, in_range);
}

// ------------------------------------------------------------------------------------
// P2: exclusive scan of the histogram -> bin_cursor[b] = first destination slot of bin b;
// *n_live = start of the dead bin. Three small launches (the histogram has up to ~2e5 bins):
// per-chunk sums, scan of the chunk sums by one CTA, per-chunk scan plus chunk offset.
// ------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;  // bins per CTA

// Exclusive scan of one value per thread across the CTA; *total receives the CTA sum.
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* total) {
  __shared__ unsigned warp_sums[32];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (unsigned)o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= (unsigned)o) w += t;
    }
    warp_sums[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const unsigned before = warp ? warp_sums[warp - 1] : 0u;
  *total = warp_sums[(blockDim.x >> 5) - 1];
  return before + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_chunk_sums(SortArgs s, unsigned* chunk_sum) {
  const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  unsigned sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < s.nbins) sum += s.bin_count[base + k];
  unsigned total;
  block_exclusive_scan(sum, &total);
  if (threadIdx.x == 0) chunk_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_chunk_offsets(unsigned* chunk_sum, int nchunks) {
  // nchunks <= 1024 * 8 covers 1.6e7 bins; one value per thread, looped for generality
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nchunks; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const unsigned v = i < nchunks ? chunk_sum[i] : 0u;
    unsigned total;
    const unsigned ex = block_exclusive_scan(v, &total);
    const unsigned c = carry;
    if (i < nchunks) chunk_sum[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_bins(SortArgs s, const unsigned* chunk_sum) {
  const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  unsigned c[kScanItems];
  unsigned sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    c[k] = base + k < s.nbins ? s.bin_count[base + k] : 0u;
    sum += c[k];
  }
  unsigned total;
  unsigned run = chunk_sum[blockIdx.x] + block_exclusive_scan(sum, &total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < s.nbins) {
      s.bin_cursor[base + k] = run;
      if (base + k == s.nbins - 1) s.n_live[0] = run;
      if (base + k == s.nq * s.ntiles) s.n_live[1] = run;  // first bin behind the collision class
    }
    run += c[k];
  }
}

// ------------------------------------------------------------------------------------
// P3: scatter every record of [0, n_upper) to its sorted position in the double buffer.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(const BankView src, const BankView dst,
                                                 const SortArgs s) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = slot < s.n_upper;
  const unsigned key = in_range ? s.keys[slot] : 0u;
  unsigned to = claim_in_bin(s.bin_cursor, key, in_range);
  // Slots behind the live prefix hold particles that died in earlier steps: they keep their
  // place (the double buffer must stay a complete bank for export).
  if (!in_range) to = (unsigned)slot;
  if (slot < s.n) {
    dst.pos[to] = src.pos[slot];
    dst.dir[to] = src.dir[slot];
    dst.ew[to] = src.ew[slot];
    dst.tm[to] = src.tm[slot];
    dst.meta[to] = src.meta[slot];
  }
}

// Self-test hook: q[i] = div_by_known(a[i], b[i], 1/b[i]) next to the IEEE quotient.
__global__ void k_selftest_div(const double* a, const double* b, double* fast, double* ieee,
                               int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double y = 1.0 / b[i];
  fast[i] = div_by_known(a[i], b[i], y);
  ieee[i] = a[i] / b[i];
}

// Self-test hook: the straight-line cores of nb_fastmath.cuh next to the plain operators.
// out = 6 arrays of n: div_core(a,b), a/b, rcp_core(b), 1/b, sqrt_core(|a|), sqrt(|a|).
__global__ void k_selftest_fastmath(const double* a, const double* b, double* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = a[i], y = b[i], ax = fabs(x);
  out[i] = div_core(x, y);
  out[(size_t)n + i] = x / y;
  out[2 * (size_t)n + i] = rcp_core(y);
  out[3 * (size_t)n + i] = 1.0 / y;
  out[4 * (size_t)n + i] = sqrt_core(ax);
  out[5 * (size_t)n + i] = sqrt(ax);
}

// ------------------------------------------------------------------------------------
// Launch wrappers
// ------------------------------------------------------------------------------------
static inline int blocks_for(int n, int threads) { return (n + threads - 1) / threads; }

int launch_sort_phase(const StepArgs& a, const SortArgs& s, const BankView& alt,
                      cudaStream_t st) {
  if (s.n_upper <= 0) return 0;
  cudaMemsetAsync(s.bin_count, 0, sizeof(unsigned) * s.nbins, st);
  k_begin_step<<<blocks_for(s.n_upper, 256), 256, 0, st>>>(a, s);
  const int nchunks = blocks_for(s.nbins, kScanChunk);
  k_scan_chunk_sums<<<nchunks, kScanThreads, 0, st>>>(s, s.chunk_sum);
  k_scan_chunk_offsets<<<1, 1024, 0, st>>>(s.chunk_sum, nchunks);
  k_scan_bins<<<nchunks, kScanThreads, 0, st>>>(s, s.chunk_sum);
  k_scatter<<<blocks_for(s.n, 256), 256, 0, st>>>(a.bank, alt, s);
  return 5;
}

int launch_selftest_fastmath(const double* a, const double* b, double* out, int n,
                             cudaStream_t st) {
  k_selftest_fastmath<<<blocks_for(n, 256), 256, 0, st>>>(a, b, out, n);
  return 1;
}

int launch_selftest_div(const double* a, const double* b, double* fast, double* ieee, int n,
                        cudaStream_t st) {
  k_selftest_div<<<blocks_for(n, 256), 256, 0, st>>>(a, b, fast, ieee, n);
  return 1;
}

}  // namespace nb
