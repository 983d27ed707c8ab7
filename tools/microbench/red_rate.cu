// red_rate.cu - how many FP64 reductions per second does the B200 L2 absorb?
// Each thread issues `iters` red.global.add.f64 to addresses of one of three patterns inside a
// footprint of `cells` doubles: 0 = uniformly random, 1 = a mesh walk (+-1 or +-nx per step, the
// tally access pattern of a streaming particle), 2 = random with one FP64 multiply-add chain of
// `work` operations between reductions (is the rate hidden behind arithmetic?), 3 = a cheap
// strided sweep (cell = (cell + odd stride) & mask: two integer instructions per reduction, so
// that the loop cannot be the limit), 4 = a straight walk (cell += 1 per step, lanes far
// apart: the tally pattern of particles streaming along x).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/red_rate tools/microbench/red_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void k_red(double* tally, size_t cells, int nx, int iters, int pattern, int work,
                      double* sink) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long s = 0x9E3779B97F4A7C15ull * (tid + 1);
  size_t cell = (s >> 20) % cells;
  double acc = 1.0 + 1e-9 * tid;
  for (int i = 0; i < iters; ++i) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    if (pattern == 1) {
      const unsigned r = (unsigned)(s >> 61);  // 0..7
      const long long step = (r & 1) ? ((r & 2) ? 1 : -1) : ((r & 2) ? nx : -nx);
      long long c = (long long)cell + step;
      if (c < 0) c += cells;
      if ((size_t)c >= cells) c -= cells;
      cell = (size_t)c;
    } else if (pattern == 3) {
      cell = (cell + 0x9E3779B1ull) & (cells - 1);  // cells is a power of two
    } else if (pattern == 4) {
      cell = (cell + 1) & (cells - 1);
    } else {
      cell = (s >> 20) % cells;
    }
    for (int w = 0; w < work; ++w) acc = acc * 1.0000000001 + 1e-12;
    atomicAdd(tally + cell, acc);
  }
  if (acc == 123.456) *sink = acc;
}

int main(int argc, char** argv) {
  const int iters = 2000;
  double *tally, *sink;
  const size_t max_cells = (size_t)64 << 20;  // 512 MB
  cudaMalloc(&tally, max_cells * sizeof(double));
  cudaMalloc(&sink, sizeof(double));
  cudaMemset(tally, 0, max_cells * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = 148 * 16, threads = 128;
  printf("pattern work footprint_MB reds_per_s\n");
  for (int pattern = 0; pattern < 5; ++pattern) {
    for (size_t mb : {1, 16, 64, 128, 256, 512}) {
      const size_t cells = mb * (1 << 20) / 8;
      const int work = pattern == 2 ? 40 : 0;
      k_red<<<blocks, threads>>>(tally, cells, 4000, 100, pattern, work, sink);  // warm-up
      cudaEventRecord(e0);
      k_red<<<blocks, threads>>>(tally, cells, 4000, iters, pattern, work, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("%d %d %zu %.3e\n", pattern, work, mb, (double)blocks * threads * iters / (ms * 1e-3));
    }
  }
  return 0;
}
