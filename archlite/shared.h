/*
 * arch-lite: an original, minimal stand-in for the UoB-HPC `arch` support layer that
 * neutral is normally built inside of (reference README.md:10-17). The real `arch` tree
 * is not part of the reference checkout, so every symbol here is inferred from neutral's
 * call sites. It is support code, not hot path.
 *
 * shared.h: error/exit macro, min/max, transfer directions and the per-kernel-set
 * allocation layer. The allocation layer has two implementations:
 *   archlite/alloc_host.c        - host memory (used by the omp3 reference build)
 *   neutral_b200/csrc/alloc.cu   - device memory on the current GPU (KERNELS=b200)
 *
 * Call sites served: neutral_data.c:54-62,97-105,146-147,168-169; main.c:22,63;
 * omp3/neutral.c:74,510,549,572.
 */
#ifndef ARCHLITE_SHARED_H
#define ARCHLITE_SHARED_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "profiler.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MASTER 0
#define GB (1024.0 * 1024.0 * 1024.0)

/* Directions of copy_buffer: RECV = device -> host, SEND = host -> device. */
enum { RECV = 0, SEND = 1 };

#define TERMINATE(...)                                                         \
  do {                                                                         \
    fprintf(stderr, __VA_ARGS__);                                              \
    fprintf(stderr, " [%s:%d]\n", __FILE__, __LINE__);                         \
    exit(EXIT_FAILURE);                                                        \
  } while (0)

#ifndef __cplusplus
#ifndef min
#define min(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef max
#define max(a, b) (((a) > (b)) ? (a) : (b))
#endif
#endif

/* Zero-filled allocations owned by the kernel set; each returns the bytes allocated. */
size_t allocate_data(double** buf, size_t len);
size_t allocate_float_data(float** buf, size_t len);
size_t allocate_int_data(int** buf, size_t len);
size_t allocate_uint64_data(uint64_t** buf, size_t len);
/* Host staging buffer (always host memory, also in device builds). */
void allocate_host_data(double** buf, size_t len);
void allocate_host_float_data(float** buf, size_t len);
void deallocate_data(double* buf);
void deallocate_host_data(double* buf);

/* Copies len doubles between a kernel-set buffer and a host buffer. */
void copy_buffer(const size_t len, double** src, double** dst, int send);
/* Hands a host buffer over to the kernel set: *dst receives kernel-set memory holding
 * the same len doubles; the host buffer is released. */
void move_host_buffer_to_device(const size_t len, double** src, double** dst);

/* |expected - result| relative to expected, within tol. */
int within_tolerance(const double expected, const double result,
                     const double tolerance);

/* Selects the accelerator for this rank (no-op for host builds). */
void initialise_devices(int rank);

#ifdef __cplusplus
}
#endif

#endif
