// nb_group.cuh - the tally collective of a particle-sharded run, device side.
//
// Every rank (GPU) of a group transports its particle shard into a private per-timestep DELTA
// tally (zero on entry). What the reference keeps in one shared array (the cumulative tally,
// omp3/neutral.c:408-420, main.c never clears it) is OWNED here by slices: rank r owns cells
// [r * chunk, (r + 1) * chunk) of the cumulative tally. After a timestep
//
//     owned_r[i] += sum over ranks d of delta_d[r * chunk + i]            (k_reduce_fold)
//
// is ONE kernel per rank that reads its slice of every peer's delta straight out of the
// peer's memory over NVLink (peer-mapped pointers: cudaDeviceEnablePeerAccess inside one
// process, CUDA IPC between processes) and folds the sum into the slice it owns: a
// reduce-scatter fused with the accumulation, moving (N-1)/N of one tally per GPU and touching
// the owner's HBM once - instead of an all-reduce (twice the NVLink bytes) followed by a fold
// that reads and writes two whole tallies on every GPU. The caller-visible tally is brought up
// to date only when somebody looks at it (k_gather_owned: validate, downloads).
//
// Ranks synchronise on the device, through flags in each other's memory (SyncBlock), never
// through the host: a rank raises `ready` behind its history kernel in stream order, readers
// poll it with acquire loads before they touch the delta, and report `consumed` into the
// owner's block when they are done, which is what the owner waits for before it clears the
// buffer for reuse three timesteps later. All waits are bounded (kWaitTimeoutNs): a peer that
// never arrives raises `fault`, which the host turns into a fatal error, instead of hanging
// the GPU.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace nb {

constexpr int kMaxRanks = 16;
constexpr int kDeltaBuffers = 3;
constexpr unsigned long long kWaitTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

// One per rank, at the start of the rank's slab (peer-mapped on every other rank).
struct SyncBlock {
  unsigned long long ready;                 // last epoch whose delta is complete (owner writes)
  unsigned long long flush_ready;           // last flush epoch whose owned slice is final
  unsigned long long fault;                 // a bounded wait expired (anyone writes)
  unsigned int cta_done;                    // last-CTA election of the owner's reduce kernel
  unsigned int cta_done_flush;              // ... and of its gather kernel
  unsigned long long consumed[kMaxRanks];   // [r]: rank r has read its slice up to this epoch
  unsigned long long flush_consumed[kMaxRanks];
};

// What rank `rank` needs to know about everybody for one kernel (passed by value).
struct GroupView {
  int nranks;
  int rank;
  size_t chunk;                        // cells per owned slice (even; nranks * chunk >= ncells)
  size_t ncells;
  const double* src[kMaxRanks];        // delta buffer of the epoch (reduce) / owned slice (gather)
  SyncBlock* sync[kMaxRanks];
};

// Flags one warp waits for (k_wait_flags) before the kernels behind it on the stream run.
struct WaitList {
  int n;
  const unsigned long long* flag[kMaxRanks];
  unsigned long long* fault;
};
int launch_wait_flags(const WaitList& w, unsigned long long want, cudaStream_t st);
// rank's slice: owned[i] += sum_d delta_d[rank * chunk + i], behind a wait for every rank's
// `ready`; at the end the rank reports `consumed[rank] = epoch` into every rank's block.
// epoch 0 = no flag traffic (the host already ordered everything: single-process flushes).
int launch_reduce_fold(const GroupView& g, double* owned, unsigned long long epoch, int ctas,
                       cudaStream_t st);
// tally[i] += owned_{i / chunk}[i % chunk] for every cell (g.src = owned slices), behind a wait
// for every rank's `flush_ready`; reports `flush_consumed[rank] = epoch` at the end.
int launch_gather_owned(const GroupView& g, double* tally, unsigned long long epoch,
                        cudaStream_t st);
// Raises sync->ready (flush = 0) or sync->flush_ready (flush = 1) to `epoch`, release-ordered
// behind everything in front of it on the stream.
int launch_signal(SyncBlock* mine, int flush, unsigned long long epoch, cudaStream_t st);
// dst[i] += src[i] for i < n (the NCCL flavour's fold of a reduce-scatter result).
int launch_fold_plain(double* dst, const double* src, size_t n, cudaStream_t st);

}  // namespace nb
