// stage.cu - per-timestep staging of the read-only inputs of the history kernel (sm_100a).
//
// The caller hands solve_transport_2d plain device arrays (neutral_interface.h:11-20):
// two cross-section tables (neutral_data.h:38-43) and the density mesh. Both are re-read
// millions of times per step in a pattern the raw layout serves badly, so every step starts
// with two tiny kernels that restage them:
//
//   k_stage_cs     {key,value} interleaved grid points + a bucket index over the leading
//                  bits of the energy (CsStage). Replaces the ~15 dependent probes of the
//                  reference's search (omp3/neutral.c:506-511) by 1-6 probes of one or two
//                  cache lines. Also checks what the host assumed about the tables.
//   k_stage_tiles  one density value per uniform 16x16-cell tile and, from those
//   k_stage_coarse (k_stage_coarse), per uniform 256x256-cell tile (TileMap): a facet crossing
//                  (omp3/neutral.c:372-378) consults the cache-resident maps instead of a
//                  random 32-byte sector of the 128 MB density mesh.
//
// Restaging every step (a few microseconds) means the caller may change the tables or the
// density between timesteps, exactly as with the reference. Neither structure changes a
// result: the bracketing interval of an energy is unique and a uniform tile's value is the
// cell's own density, bit for bit.
#include "nb_device.cuh"
#include "transport.cuh"

namespace nb {

__device__ __forceinline__ int stage_bucket_id(double key, unsigned long long bits0, int shift,
                                               int nb) {
  const long long d = (long long)(double_to_bits(key) - bits0);
  if (d <= 0) return 0;
  const unsigned long long q = (unsigned long long)d >> shift;
  return q < (unsigned long long)(nb - 1) ? (int)q : nb - 1;
}

// Threads [0, n) interleave the grid points and check them; threads [0, nb] each find one
// entry of the bucket index by bisection (bucket[b] = number of keys with id < b).
__global__ void __launch_bounds__(256) k_stage_cs(const double* __restrict__ keys,
                                                  const double* __restrict__ vals, int n,
                                                  double2* __restrict__ kv,
                                                  int* __restrict__ bucket,
                                                  unsigned long long bits0, int shift, int nb,
                                                  const double* __restrict__ twin_keys,
                                                  unsigned long long* totals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double k = keys[i];
    kv[i] = make_double2(k, vals[i]);
    // the grid must be strictly increasing (the reference's search assumes it,
    // omp3/neutral.c:506-511) ...
    bool fault = i > 0 && !(keys[i - 1] < k);
    if (fault) atomicAdd(totals + kTotFault, 1ull);
    // ... and does the other table (same length) share this energy grid, bit for bit? Decided
    // here, where the data is, every timestep (same_grid() in nb_device.cuh reads the count).
    if (twin_keys && double_to_bits(twin_keys[i]) != double_to_bits(k))
      atomicAdd(totals + kTotGridsDiffer, 1ull);
  }
  if (i <= nb) {
    int lo = 0, hi = n;  // first index whose bucket id is >= i
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (stage_bucket_id(keys[mid], bits0, shift, nb) < i) lo = mid + 1; else hi = mid;
    }
    bucket[i] = lo;
  }
}

// One warp per fine tile; lane l covers row l/2, columns 8*(l%2) .. +7 of the 16x16 tile.
__global__ void __launch_bounds__(256) k_stage_tiles(const double* __restrict__ density, int nx,
                                                     int ny, int tiles_x, int ntiles,
                                                     double* __restrict__ fine) {
  static_assert(kTileShift == 4, "lane mapping below assumes 16x16 tiles");
  const int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (tile >= ntiles) return;
  const int lane = threadIdx.x & 31;
  const int cx0 = (tile % tiles_x) << kTileShift;
  const int cy0 = (tile / tiles_x) << kTileShift;
  const unsigned long long first = double_to_bits(density[(size_t)cy0 * nx + cx0]);
  const int cy = cy0 + (lane >> 1);
  const int cxb = cx0 + ((lane & 1) << 3);
  bool same = true;
  if (cy < ny) {
    const double* row = density + (size_t)cy * nx;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (cxb + c < nx) same = same && (double_to_bits(row[cxb + c]) == first);
  }
  same = __all_sync(0xffffffffu, same) && first != kMixedTileBits;
  if (lane == 0) fine[tile] = bits_to_double(same ? first : kMixedTileBits);
}

// One warp per coarse tile: uniform iff its (up to) 16x16 fine tiles are uniform and equal.
// (One thread per coarse tile, 256 dependent loads each in a single CTA, took 55 us.)
__global__ void __launch_bounds__(256) k_stage_coarse(const double* __restrict__ fine,
                                                      int fine_tx, int fine_ty, int coarse_tx,
                                                      int ncoarse, double* __restrict__ coarse) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= ncoarse) return;
  const int lane = threadIdx.x & 31;
  constexpr int kRatio = 1 << (kCoarseShift - kTileShift);
  const int fx0 = (t % coarse_tx) * kRatio, fy0 = (t / coarse_tx) * kRatio;
  const unsigned long long first = double_to_bits(fine[fy0 * fine_tx + fx0]);
  bool same = first != kMixedTileBits;
  for (int k = lane; k < kRatio * kRatio; k += 32) {
    const int j = k / kRatio, i = k % kRatio;
    if (fy0 + j < fine_ty && fx0 + i < fine_tx)
      same = same && double_to_bits(fine[(fy0 + j) * fine_tx + fx0 + i]) == first;
  }
  same = __all_sync(0xffffffffu, same);
  if (lane == 0) coarse[t] = bits_to_double(same ? first : kMixedTileBits);
}

// Rows 0..3 of StepArgs::edges4 (see nb_bank.cuh).
__global__ void __launch_bounds__(256) k_stage_edges(const double* __restrict__ edgex, int nx,
                                                     const double* __restrict__ edgey, int ny,
                                                     int stride, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= nx) {
    const double e = edgex[i];
    out[i] = e;
    out[2 * (size_t)stride + i] = e - kOpenBoundCorrection;
  }
  if (i <= ny) {
    const double e = edgey[i];
    out[(size_t)stride + i] = e;
    out[3 * (size_t)stride + i] = e - kOpenBoundCorrection;
  }
}

// The same verdict for the direct kernel, which reads the caller's tables unstaged.
__global__ void __launch_bounds__(256) k_compare_grids(const double* __restrict__ a,
                                                       const double* __restrict__ b, int n,
                                                       unsigned long long* totals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && double_to_bits(a[i]) != double_to_bits(b[i]))
    atomicAdd(totals + kTotGridsDiffer, 1ull);
}

static inline int blocks_for(size_t n, int threads) { return (int)((n + threads - 1) / threads); }

int launch_stage_cs(const double* keys, const double* vals, int n, double2* kv, int* bucket,
                    unsigned long long bits0, int shift, int nb, const double* twin_keys,
                    unsigned long long* totals, cudaStream_t st) {
  if (n <= 0) return 0;
  k_stage_cs<<<blocks_for((size_t)(n > nb + 1 ? n : nb + 1), 256), 256, 0, st>>>(keys, vals, n, kv, bucket, bits0, shift, nb,
                                                 twin_keys, totals);
  return 1;
}

int launch_compare_grids(const double* a, const double* b, int n, unsigned long long* totals,
                         cudaStream_t st) {
  if (n <= 0) return 0;
  k_compare_grids<<<blocks_for((size_t)n, 256), 256, 0, st>>>(a, b, n, totals);
  return 1;
}

int launch_stage_edges(const double* edgex, int nx, const double* edgey, int ny, int stride,
                       double* out, cudaStream_t st) {
  const int n = (nx > ny ? nx : ny) + 1;
  k_stage_edges<<<blocks_for((size_t)n, 256), 256, 0, st>>>(edgex, nx, edgey, ny, stride, out);
  return 1;
}

int launch_stage_tiles(const double* density, int nx, int ny, double* fine, double* coarse,
                       TileMap* map, cudaStream_t st) {
  const int fine_tx = ((nx - 1) >> kTileShift) + 1, fine_ty = ((ny - 1) >> kTileShift) + 1;
  const int coarse_tx = ((nx - 1) >> kCoarseShift) + 1, coarse_ty = ((ny - 1) >> kCoarseShift) + 1;
  k_stage_tiles<<<blocks_for((size_t)fine_tx * fine_ty * 32, 256), 256, 0, st>>>(
      density, nx, ny, fine_tx, fine_tx * fine_ty, fine);
  k_stage_coarse<<<blocks_for((size_t)coarse_tx * coarse_ty * 32, 256), 256, 0, st>>>(
      fine, fine_tx, fine_ty, coarse_tx, coarse_tx * coarse_ty, coarse);
  *map = TileMap{fine, coarse, fine_tx, coarse_tx};
  return 2;
}

}  // namespace nb
