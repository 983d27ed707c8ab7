// red_paths.cu - which part of the machine sets the FP64-reduction ceiling of the history kernel,
// and is there a path around it?
//
//   sm-scaling   red.global.add.f64 to spread addresses (pattern 3 of csrc/microbench.cu) from
//                1 CTA of 1024 threads per SM on 37 / 74 / 111 / 148 SMs: a rate that grows with
//                the SM count is an SM-side (LSU) limit, one that saturates is the L2's.
//   tma          the same reductions issued as 16-byte cp.reduce.async.bulk (TMA) operations
//                from shared memory, one per thread and iteration: {value, 0.0} into the aligned
//                cell pair (x + 0.0 == x, so the neighbour cell is unchanged).
//   tma32        32-byte bulk reductions (a whole sector: value + three zeros).
//   pair         red.global.add.f64 with lanes 2k, 2k+1 on the two cells of one 16-byte pair.
//   atom         atomicAdd with the return value consumed (ATOMG instead of REDG).
//
// Build (sm_100a only): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/red_paths tools/microbench/red_paths.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) k_red_spread(double* tally, size_t mask, int iters, int mode,
                                                     double* sink) {
  extern __shared__ double2 pad[];
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  size_t cell = ((size_t)(0x9E3779B97F4A7C15ull * (tid + 1)) >> 20) & mask;
  if (mode == 1) cell = (cell & ~(size_t)1) | (threadIdx.x & 1);  // lane pairs share 16 bytes
  const double v = 1.0 + 1e-9 * tid;
  double acc = 0.0;
  for (int i = 0; i < iters; ++i) {
    cell = (cell + (mode == 1 ? 0x9E3779B2ull : 0x9E3779B1ull)) & mask;
    if (mode == 2) acc += atomicAdd(tally + cell, v);
    else atomicAdd(tally + cell, v);
  }
  if (acc == 123.456) *sink = acc;
}

// One bulk reduction per thread and iteration; `bytes` = 16 or 32.
template <int kBytes>
__global__ void __launch_bounds__(256) k_red_tma(double* tally, size_t mask, int iters) {
  __shared__ __align__(32) double slots[256 * (kBytes / 8)];
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  size_t cell = ((size_t)(0x9E3779B97F4A7C15ull * (tid + 1)) >> 20) & mask;
  const double v = 1.0 + 1e-9 * tid;
  double* mine = slots + threadIdx.x * (kBytes / 8);
  const unsigned smem_addr = (unsigned)__cvta_generic_to_shared(mine);
  constexpr size_t kAlignMask = ~(size_t)(kBytes / 8 - 1);
  for (int i = 0; i < iters; ++i) {
    cell = (cell + 0x9E3779B1ull) & mask;
#pragma unroll
    for (int k = 0; k < kBytes / 8; ++k) mine[k] = ((cell & (kBytes / 8 - 1)) == (size_t)k) ? v : 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    double* dst = tally + (cell & kAlignMask);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_addr), "n"(kBytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static double time_launch(cudaEvent_t e0, cudaEvent_t e1) {
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e-3;
}

int main() {
  const size_t cells = (size_t)2 << 20;  // 16 MiB: L2 resident
  double *tally, *sink;
  cudaMalloc(&tally, cells * sizeof(double));
  cudaMalloc(&sink, sizeof(double));
  cudaMemset(tally, 0, cells * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int big_smem = 200 * 1024;  // one CTA per SM
  cudaFuncSetAttribute(k_red_spread, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
  printf("test sms threads iters reductions_per_s per_sm_per_clk(1.965GHz)\n");
  for (int sms : {37, 74, 111, 148}) {
    const int iters = 4000;
    k_red_spread<<<sms, 1024, big_smem>>>(tally, cells - 1, 200, 0, sink);
    cudaEventRecord(e0);
    k_red_spread<<<sms, 1024, big_smem>>>(tally, cells - 1, iters, 0, sink);
    cudaEventRecord(e1);
    const double s = time_launch(e0, e1);
    const double rate = (double)sms * 1024 * iters / s;
    printf("sm-scaling %d 1024 %d %.3e %.3f\n", sms, iters, rate, rate / sms / 1.965e9);
  }
  for (int mode : {0, 1, 2}) {
    const int iters = 2000, blocks = 148 * 2;
    k_red_spread<<<blocks, 1024, 0>>>(tally, cells - 1, 200, mode, sink);
    cudaEventRecord(e0);
    k_red_spread<<<blocks, 1024, 0>>>(tally, cells - 1, iters, mode, sink);
    cudaEventRecord(e1);
    const double s = time_launch(e0, e1);
    printf("%s 148 %d %d %.3e\n", mode == 0 ? "red" : mode == 1 ? "pair" : "atom", blocks * 1024, iters,
           (double)blocks * 1024 * iters / s);
  }
  {
    const int iters = 500, blocks = 148 * 8;
    k_red_tma<16><<<blocks, 256>>>(tally, cells - 1, 50);
    cudaEventRecord(e0);
    k_red_tma<16><<<blocks, 256>>>(tally, cells - 1, iters);
    cudaEventRecord(e1);
    double s = time_launch(e0, e1);
    printf("tma16 148 %d %d %.3e  (%s)\n", blocks * 256, iters, (double)blocks * 256 * iters / s,
           cudaGetErrorString(cudaGetLastError()));
    k_red_tma<32><<<blocks, 256>>>(tally, cells - 1, 50);
    cudaEventRecord(e0);
    k_red_tma<32><<<blocks, 256>>>(tally, cells - 1, iters);
    cudaEventRecord(e1);
    s = time_launch(e0, e1);
    printf("tma32 148 %d %d %.3e  (%s)\n", blocks * 256, iters, (double)blocks * 256 * iters / s,
           cudaGetErrorString(cudaGetLastError()));
  }
  // sanity: the bulk reductions landed (sum of the tally is finite and positive)
  double* h = (double*)malloc(cells * sizeof(double));
  cudaMemcpy(h, tally, cells * sizeof(double), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (size_t i = 0; i < cells; ++i) sum += h[i];
  printf("tally sum %.6e (%s)\n", sum, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
