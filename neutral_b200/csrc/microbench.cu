// microbench.cu - the ceiling the history kernel actually sits under, measured in place.
//
// Every facet, census and death of a history is one red.global.add.f64 into the tally
// (update_tallies, omp3/neutral.c:408-420) at an address no other lane of the warp shares, so
// the facet-dominated decks are bound by the rate at which the L2 retires FP64 reductions,
// not by HBM bytes. This kernel measures that rate on the GPU the bench runs on:
// each thread issues `iters` reductions into a footprint of `cells` doubles in one of these
// address patterns
//   0  uniformly random cells                       (worst case: every reduction a new sector)
//   1  a mesh walk, +-1 or +-nx per reduction       (the tally pattern of a streaming particle)
//   3  cell += odd stride (mod cells)               (two integer instructions per reduction:
//                                                    the loop cannot be the limit - the PEAK)
//   5  as 3, but the 32 lanes of a warp sit on 32 consecutive cells (is the ceiling per
//      request / sector, or per element?)
//   6  as 3, but lanes pair up on 4 consecutive cells = one 32-byte sector per 4 lanes
// and the launch is timed with CUDA events on the caller's stream.
#include <cuda_runtime.h>
#include <stdint.h>

#include "transport.cuh"

namespace nb {

__global__ void __launch_bounds__(128)
k_red_rate(double* __restrict__ tally, size_t cells, int nx, int iters, int pattern) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  unsigned long long s = 0x9E3779B97F4A7C15ull * (tid + 1);
  if (pattern == 5) s = 0x9E3779B97F4A7C15ull * ((tid >> 5) + 1);
  if (pattern == 6) s = 0x9E3779B97F4A7C15ull * ((tid >> 2) + 1);
  size_t cell = (size_t)(s >> 20) % cells;
  if (pattern == 5) cell = (cell & ~(size_t)31) + lane;
  if (pattern == 6) cell = (cell & ~(size_t)3) + (lane & 3u);
  const double v = 1.0 + 1e-9 * tid;
  const size_t mask = cells - 1;  // cells is a power of two
  for (int i = 0; i < iters; ++i) {
    if (pattern == 1) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      const unsigned r = (unsigned)(s >> 61);
      const long long step = (r & 1) ? ((r & 2) ? 1 : -1) : ((r & 2) ? nx : -nx);
      cell = (size_t)((long long)cell + step) & mask;
    } else if (pattern == 0) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      cell = (size_t)(s >> 20) & mask;
    } else if (pattern == 5) {
      cell = (cell + 0x9E3779A0ull) & mask;  // multiple of 32: the warp stays on one line pair
    } else if (pattern == 6) {
      cell = (cell + 0x9E3779B4ull) & mask;  // multiple of 4: the quad stays on one sector
    } else {
      cell = (cell + 0x9E3779B1ull) & mask;
    }
    atomicAdd(tally + cell, v);
  }
}

// Returns the kernels launched; *seconds receives the timed launch's duration and
// *reductions how many reductions it issued. `scratch` holds `cells` doubles (power of two).
int launch_red_rate(double* scratch, size_t cells, int nx, int iters, int pattern,
                    cudaEvent_t e0, cudaEvent_t e1, double* seconds, double* reductions,
                    cudaStream_t st) {
  const int blocks = 148 * 16, threads = 128;
  k_red_rate<<<blocks, threads, 0, st>>>(scratch, cells, nx, iters / 8 + 1, pattern);  // warm-up
  cudaEventRecord(e0, st);
  k_red_rate<<<blocks, threads, 0, st>>>(scratch, cells, nx, iters, pattern);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, e0, e1);
  *seconds = (double)ms * 1e-3;
  *reductions = (double)blocks * threads * (double)iters;
  return 2;
}

}  // namespace nb
