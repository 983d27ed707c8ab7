/*
 * arch-lite allocation layer, host flavour: kernel-set memory is plain host memory.
 * Used by the reference omp3 build under oracle/_ref (the device flavour lives in
 * neutral_b200/csrc/alloc.cu). Call sites: neutral_data.c:54-62,97-105,146-147,168-169.
 */
#include "shared.h"

#include <string.h>

static void* zalloc(size_t bytes) {
  void* p = calloc(bytes ? bytes : 1, 1);
  if (!p) {
    TERMINATE("Could not allocate %zu bytes", bytes);
  }
  return p;
}

size_t allocate_data(double** buf, size_t len) {
  *buf = (double*)zalloc(sizeof(double) * len);
  return sizeof(double) * len;
}

size_t allocate_float_data(float** buf, size_t len) {
  *buf = (float*)zalloc(sizeof(float) * len);
  return sizeof(float) * len;
}

size_t allocate_int_data(int** buf, size_t len) {
  *buf = (int*)zalloc(sizeof(int) * len);
  return sizeof(int) * len;
}

size_t allocate_uint64_data(uint64_t** buf, size_t len) {
  *buf = (uint64_t*)zalloc(sizeof(uint64_t) * len);
  return sizeof(uint64_t) * len;
}

void allocate_host_data(double** buf, size_t len) {
  *buf = (double*)zalloc(sizeof(double) * len);
}

void allocate_host_float_data(float** buf, size_t len) {
  *buf = (float*)zalloc(sizeof(float) * len);
}

void deallocate_data(double* buf) { free(buf); }
void deallocate_host_data(double* buf) { free(buf); }

void copy_buffer(const size_t len, double** src, double** dst, int send) {
  (void)send;
  memcpy(*dst, *src, sizeof(double) * len);
}

void move_host_buffer_to_device(const size_t len, double** src, double** dst) {
  (void)len;
  *dst = *src;
  *src = NULL;
}

void initialise_devices(int rank) { (void)rank; }
