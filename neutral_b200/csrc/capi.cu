// capi.cu - the C ABI of libneutral_b200.so (declared in include/neutral_b200.h).
//
// Host-side glue only: argument checking, device memory, stream ordering, counters. The
// arithmetic of the hot path lives in transport.cu. There is deliberately no CPU fallback:
// every compute entry point fails loudly when no CUDA device is usable.
#include "../../include/neutral_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "transport.cuh"
#include "nb_sincos.cuh"

namespace {

using namespace nb;

const SinCosTable kHostSinCosTable = {{
#include "glibc_sincos_table.inc"
}};

const LogTable kHostLogTable = {
#include "glibc_log_table.inc"
};

constexpr uint64_t kBankMagic = 0x6e62323030424e4bull;  // "nb200BNK"

struct Bank {
  BankView cur{};
  BankView alt{};  // double buffer of the per-step sort (allocated on first use)
  bool has_alt = false;
  unsigned* keys = nullptr;
  int n = 0;
  int n_upper = 0;  // slots [n_upper, n) are known to hold dead particles
  uint64_t pid0 = 0;
  SoaView exported{};  // lazily allocated plain SoA view (11 arrays)
  bool has_export = false;
};

// What inject_particles / nb200_bank_create hand out as `Particle*`: the reference's SoA
// struct first (so -DSoA code can read the 11 pointers), our bookkeeping after it.
struct BankHandle {
  nb200_particle_soa view;
  uint64_t magic;
  Bank* impl;
};

constexpr int kCsBuckets = 8192;

// Parameters of the bucket index of one table (CsStage in nb_bank.cuh).
struct CsParams {
  unsigned long long bits0 = 0;
  int shift = 63;
  int nb = 1;  // one bucket = plain bisection over the whole grid (always valid)
};

CsParams cs_params_from_host(const double* keys, int n) {
  CsParams p;
  if (n >= 2 && keys[0] > 0.0 && keys[n - 1] > keys[0]) {
    unsigned long long lo, hi;
    memcpy(&lo, &keys[0], 8);
    memcpy(&hi, &keys[n - 1], 8);
    p.bits0 = lo;
    p.nb = kCsBuckets;
    p.shift = 0;
    while (((hi - lo) >> p.shift) >= (unsigned long long)p.nb) p.shift++;
  }
  return p;
}

struct Context {
  bool ready = false;
  int device = 0;
  cudaStream_t stream = 0;
  LogTable* d_logt = nullptr;
  SinCosTable* d_sct = nullptr;
  int opt_device_inject = 1;
  unsigned long long* d_totals = nullptr;
  unsigned long long* h_totals = nullptr;  // pinned, mapped
  unsigned long long* h_totals_dev = nullptr;  // the same memory as the device addresses it
  cudaEvent_t ev_begin = nullptr, ev_mid = nullptr, ev_end = nullptr;  // phase timing
  // The density tile maps are only read by the event loop: they are staged on a side stream
  // beside the begin-step / sort kernels (fork after the previous history kernel, join
  // before the next one).
  cudaStream_t stage_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_tiles = nullptr;
  int opt_stage_overlap = 1;
  int opt_history_smem_pad = 0;  // occupancy probe: extra dynamic shared memory per CTA
  unsigned* d_bins = nullptr;  // histogram + cursors of the per-step sort
  unsigned* d_n_live = nullptr;
  int bins_capacity = 0;
  int opt_fast_div = 1;
  int opt_tile_shift = 8;
  int opt_length_bins = 512;
  int opt_tally_prereduce = 0;
  int opt_l2_persist = 0;
  int opt_defer_finish = 0;
  Bank* pending_bank = nullptr;  // a step enqueued by solve_transport_2d, not yet finished
  uint64_t pending_launches0 = 0;
  bool l2_limit_set = false;
  size_t l2_setaside = 0;
  // mesh extent, read once per (edgex, edgey) pair for the sort's history-length estimate
  const double* mesh_ex = nullptr;
  const double* mesh_ey = nullptr;
  double mesh_width = 1.0, mesh_height = 1.0;
  struct MeshExtent {
    const double* ex;
    const double* ey;
    int nx, ny;
    double width, height;
  };
  std::vector<MeshExtent> mesh_cache;  // the last few (edgex, edgey) pairs seen
  uint64_t launches = 0;
  uint64_t last_stats[8] = {0};
  int opt_print = 1;
  int opt_pipeline = 1;
  int shard_first = 0;
  int shard_count = -1;
  // cross-section grids seen last: same-grid detection and the bucket-index parameters are
  // derived from a host copy of the keys once per (pointer, size) pair; the staging kernels
  // re-verify them on the device every step (kTotFault)
  const double* cs_s_keys = nullptr;
  const double* cs_a_keys = nullptr;
  int cs_n = 0, cs_a_n = 0;
  int cs_same = 0;
  CsParams cs_s_par, cs_a_par;
  // per-step staging (stage.cu): cross-section tables and the density tile map
  char* d_cs_stage = nullptr;
  size_t cs_stage_bytes = 0;
  double* d_tile_rho = nullptr;
  int tile_capacity = 0;
  double* d_edges4 = nullptr;
  int edges_capacity = 0;
  std::string last_error;
};

Context g;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g.last_error = buf;
}

[[noreturn]] void terminate(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  fprintf(stderr, "neutral_b200: fatal: %s\n", buf);
  fflush(stderr);
  exit(EXIT_FAILURE);
}

#define CU_FATAL(call)                                                              \
  do {                                                                              \
    cudaError_t err__ = (call);                                                     \
    if (err__ != cudaSuccess)                                                       \
      terminate("%s failed: %s [%s:%d]", #call, cudaGetErrorString(err__), __FILE__, \
                __LINE__);                                                          \
  } while (0)

#define CU_TRY(call)                                                                \
  do {                                                                              \
    cudaError_t err__ = (call);                                                     \
    if (err__ != cudaSuccess) {                                                     \
      set_error("%s failed: %s [%s:%d]", #call, cudaGetErrorString(err__), __FILE__, \
                __LINE__);                                                          \
      return -2;                                                                    \
    }                                                                               \
  } while (0)

// Returns 0 when a device is usable and the context is initialised.
int ensure_ready() {
  if (g.ready) return 0;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count <= 0) {
    set_error("no CUDA device available (%s): the b200 kernel set has no CPU fallback",
              err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
    (void)cudaGetLastError();
    return -1;
  }
  CU_TRY(cudaGetDevice(&g.device));
  CU_TRY(cudaMalloc(&g.d_logt, sizeof(LogTable)));
  CU_TRY(cudaMemcpy(g.d_logt, &kHostLogTable, sizeof(LogTable), cudaMemcpyHostToDevice));
  CU_TRY(cudaMalloc(&g.d_sct, sizeof(SinCosTable)));
  CU_TRY(cudaMemcpy(g.d_sct, &kHostSinCosTable, sizeof(SinCosTable), cudaMemcpyHostToDevice));
  CU_TRY(cudaMalloc(&g.d_totals, sizeof(unsigned long long) * kTotCount));
  CU_TRY(cudaHostAlloc(&g.h_totals, sizeof(unsigned long long) * kTotCount, cudaHostAllocMapped));
  CU_TRY(cudaHostGetDevicePointer(&g.h_totals_dev, g.h_totals, 0));
  CU_TRY(cudaEventCreate(&g.ev_begin));
  CU_TRY(cudaEventCreate(&g.ev_mid));
  CU_TRY(cudaEventCreate(&g.ev_end));
  CU_TRY(cudaStreamCreateWithFlags(&g.stage_stream, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&g.ev_tiles, cudaEventDisableTiming));
  CU_TRY(cudaMalloc(&g.d_n_live, sizeof(unsigned)));
  g.ready = true;
  return 0;
}

void require_ready() {
  if (ensure_ready() != 0) terminate("%s", g.last_error.c_str());
}

template <typename T>
size_t device_zalloc(T** buf, size_t len) {
  require_ready();
  const size_t bytes = sizeof(T) * (len ? len : 1);
  CU_FATAL(cudaMalloc((void**)buf, bytes));
  CU_FATAL(cudaMemsetAsync(*buf, 0, bytes, g.stream));
  return sizeof(T) * len;
}

size_t bank_alloc(BankView& b, int n) {
  size_t bytes = 0;
  bytes += device_zalloc(&b.pos, (size_t)n);
  bytes += device_zalloc(&b.dir, (size_t)n);
  bytes += device_zalloc(&b.ew, (size_t)n);
  bytes += device_zalloc(&b.tm, (size_t)n);
  bytes += device_zalloc(&b.meta, (size_t)n);
  return bytes;
}

void bank_release(BankView& b) {
  cudaFree(b.pos);
  cudaFree(b.dir);
  cudaFree(b.ew);
  cudaFree(b.tm);
  cudaFree(b.meta);
  b = BankView{};
}

size_t soa_alloc(SoaView& s, int n) {
  size_t bytes = 0;
  double** d[] = {&s.x, &s.y, &s.omega_x, &s.omega_y, &s.energy, &s.weight,
                  &s.dt_to_census, &s.mfp_to_collision};
  for (double** p : d) bytes += device_zalloc(p, (size_t)n);
  int** i[] = {&s.cellx, &s.celly, &s.dead};
  for (int** p : i) bytes += device_zalloc(p, (size_t)n);
  return bytes;
}

void soa_release(SoaView& s) {
  void* p[] = {s.x, s.y, s.omega_x, s.omega_y, s.energy, s.weight, s.dt_to_census,
               s.mfp_to_collision, s.cellx, s.celly, s.dead};
  for (void* q : p) cudaFree(q);
  s = SoaView{};
}

SoaView as_soa_view(const nb200_particle_soa& p) {
  SoaView s;
  s.x = p.x; s.y = p.y; s.omega_x = p.omega_x; s.omega_y = p.omega_y;
  s.energy = p.energy; s.weight = p.weight; s.dt_to_census = p.dt_to_census;
  s.mfp_to_collision = p.mfp_to_collision;
  s.cellx = p.cellx; s.celly = p.celly; s.dead = p.dead;
  return s;
}

nb200_particle_soa as_public(const SoaView& s) {
  nb200_particle_soa p;
  p.x = s.x; p.y = s.y; p.omega_x = s.omega_x; p.omega_y = s.omega_y;
  p.energy = s.energy; p.weight = s.weight; p.dt_to_census = s.dt_to_census;
  p.mfp_to_collision = s.mfp_to_collision;
  p.cellx = s.cellx; p.celly = s.celly; p.dead = s.dead;
  return p;
}

BankHandle* new_handle(int n, uint64_t pid0, size_t* bytes) {
  BankHandle* h = (BankHandle*)calloc(1, sizeof(BankHandle));
  h->impl = new Bank();
  h->magic = kBankMagic;
  h->impl->n = n;
  h->impl->n_upper = n;
  h->impl->pid0 = pid0;
  const size_t b = bank_alloc(h->impl->cur, n);
  if (bytes) *bytes = b;
  return h;
}

Bank* bank_of(nb200_particle_soa* particles) {
  if (!particles) return nullptr;
  BankHandle* h = reinterpret_cast<BankHandle*>(particles);
  if (h->magic != kBankMagic || !h->impl) return nullptr;
  return h->impl;
}

// Uploads a host SoA bank (count slots) into device bank b, tagging origins from 0.
void upload_host_soa(const SoaView& host, int count, BankView& dst) {
  SoaView staging{};
  soa_alloc(staging, count);
  const double* hd[] = {host.x, host.y, host.omega_x, host.omega_y, host.energy, host.weight,
                        host.dt_to_census, host.mfp_to_collision};
  double* dd[] = {staging.x, staging.y, staging.omega_x, staging.omega_y, staging.energy,
                  staging.weight, staging.dt_to_census, staging.mfp_to_collision};
  for (int k = 0; k < 8; ++k)
    CU_FATAL(cudaMemcpyAsync(dd[k], hd[k], sizeof(double) * count, cudaMemcpyHostToDevice,
                             g.stream));
  const int* hi[] = {host.cellx, host.celly, host.dead};
  int* di[] = {staging.cellx, staging.celly, staging.dead};
  for (int k = 0; k < 3; ++k)
    CU_FATAL(cudaMemcpyAsync(di[k], hi[k], sizeof(int) * count, cudaMemcpyHostToDevice,
                             g.stream));
  g.launches += launch_import_soa(dst, staging, count, 0, g.stream);
  CU_FATAL(cudaStreamSynchronize(g.stream));
  soa_release(staging);
}

// Looks at the two energy grids once per (pointer, size) pair: are they the same grid, and
// which leading bits spread their keys over the bucket index.
// A host that alternates between a few working sets (bench.py's double-buffered e2e pipeline)
// shows a few pairs in turn: the last kTableCacheSlots are remembered, so that no timestep of
// a known pair issues a device-to-host copy - which would queue behind whatever bulk download
// the host has in flight on the same copy engine.
struct InspectedTables {
  const double* s_keys;
  const double* a_keys;
  int s_n, a_n, same;
  CsParams s_par, a_par;
};
constexpr size_t kTableCacheSlots = 8;
std::vector<InspectedTables> g_table_cache;

int inspect_tables(const double* s_keys, int s_n, const double* a_keys, int a_n) {
  if (g.cs_s_keys == s_keys && g.cs_a_keys == a_keys && g.cs_n == s_n && g.cs_a_n == a_n)
    return g.cs_same;
  for (const InspectedTables& t : g_table_cache) {
    if (t.s_keys == s_keys && t.a_keys == a_keys && t.s_n == s_n && t.a_n == a_n) {
      g.cs_s_keys = s_keys;
      g.cs_a_keys = a_keys;
      g.cs_n = s_n;
      g.cs_a_n = a_n;
      g.cs_same = t.same;
      g.cs_s_par = t.s_par;
      g.cs_a_par = t.a_par;
      return g.cs_same;
    }
  }
  std::vector<double> hs(std::max(s_n, 1)), ha(std::max(a_n, 1));
  CU_FATAL(cudaMemcpyAsync(hs.data(), s_keys, sizeof(double) * s_n, cudaMemcpyDeviceToHost,
                           g.stream));
  CU_FATAL(cudaMemcpyAsync(ha.data(), a_keys, sizeof(double) * a_n, cudaMemcpyDeviceToHost,
                           g.stream));
  CU_FATAL(cudaStreamSynchronize(g.stream));
  g.cs_s_keys = s_keys;
  g.cs_a_keys = a_keys;
  g.cs_n = s_n;
  g.cs_a_n = a_n;
  g.cs_same = s_n == a_n && memcmp(hs.data(), ha.data(), sizeof(double) * s_n) == 0;
  g.cs_s_par = cs_params_from_host(hs.data(), s_n);
  g.cs_a_par = cs_params_from_host(ha.data(), a_n);
  if (g_table_cache.size() >= kTableCacheSlots) g_table_cache.erase(g_table_cache.begin());
  g_table_cache.push_back({s_keys, a_keys, s_n, a_n, g.cs_same, g.cs_s_par, g.cs_a_par});
  return g.cs_same;
}

// The step totals reach the host through mapped pinned memory, written by the device itself:
// a cudaMemcpyAsync would wait its turn on the device-to-host copy engine behind any bulk
// download the caller has in flight (measured: 3.4 ms per deck run in bench.py's e2e loop).
__global__ void k_publish_totals(const unsigned long long* __restrict__ totals,
                                 volatile unsigned long long* host_totals) {
  if (threadIdx.x < kTotCount) host_totals[threadIdx.x] = totals[threadIdx.x];
  __threadfence_system();
}

// Restages both tables into the library's own block (stage.cu) and fills the views.
void stage_tables(StepArgs& a) {
  auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t kv_s = align(sizeof(double2) * a.s_n), kv_a = align(sizeof(double2) * a.a_n);
  const size_t bk_s = align(sizeof(int) * (g.cs_s_par.nb + 1));
  const size_t bk_a = align(sizeof(int) * (g.cs_a_par.nb + 1));
  const size_t need = kv_s + kv_a + bk_s + bk_a;
  if (g.cs_stage_bytes < need) {
    cudaFree(g.d_cs_stage);
    CU_FATAL(cudaMalloc(&g.d_cs_stage, need));
    g.cs_stage_bytes = need;
  }
  char* p = g.d_cs_stage;
  double2* d_kv_s = (double2*)p; p += kv_s;
  double2* d_kv_a = (double2*)p; p += kv_a;
  int* d_bk_s = (int*)p; p += bk_s;
  int* d_bk_a = (int*)p;
  g.launches += launch_stage_cs(a.s_keys, a.s_vals, a.s_n, d_kv_s, d_bk_s, g.cs_s_par.bits0,
                                g.cs_s_par.shift, g.cs_s_par.nb,
                                a.same_keys ? a.a_keys : nullptr, a.totals, g.stream);
  g.launches += launch_stage_cs(a.a_keys, a.a_vals, a.a_n, d_kv_a, d_bk_a, g.cs_a_par.bits0,
                                g.cs_a_par.shift, g.cs_a_par.nb, nullptr, a.totals, g.stream);
  a.cs_s = CsStage{d_kv_s, d_bk_s, g.cs_s_par.bits0, g.cs_s_par.shift, g.cs_s_par.nb, a.s_n};
  a.cs_a = CsStage{d_kv_a, d_bk_a, g.cs_a_par.bits0, g.cs_a_par.shift, g.cs_a_par.nb, a.a_n};
}

void stage_tiles(StepArgs& a, cudaStream_t st) {
  const int nfine = (((a.nx - 1) >> kTileShift) + 1) * (((a.ny - 1) >> kTileShift) + 1);
  const int ncoarse = (((a.nx - 1) >> kCoarseShift) + 1) * (((a.ny - 1) >> kCoarseShift) + 1);
  if (g.tile_capacity < nfine + ncoarse) {
    cudaFree(g.d_tile_rho);
    CU_FATAL(cudaMalloc(&g.d_tile_rho, sizeof(double) * (nfine + ncoarse)));
    g.tile_capacity = nfine + ncoarse;
  }
  g.launches += launch_stage_tiles(a.density, a.nx, a.ny, g.d_tile_rho, g.d_tile_rho + nfine,
                                   &a.tiles, st);
  // ... and the target-edge rows (same consumer: the event loop only)
  const int stride = ((std::max(a.nx, a.ny) + 1 + 31) / 32) * 32;
  if (g.edges_capacity < 4 * stride) {
    cudaFree(g.d_edges4);
    CU_FATAL(cudaMalloc(&g.d_edges4, sizeof(double) * 4 * (size_t)stride));
    g.edges_capacity = 4 * stride;
  }
  g.launches += launch_stage_edges(a.edgex, a.nx, a.edgey, a.ny, stride, g.d_edges4, st);
  a.edges4 = g.d_edges4;
  a.edge_stride = stride;
}

void finish_step(uint64_t* facet_events, uint64_t* collision_events);

// The one timestep both flavours share. All pointers are device memory.
void run_step(Bank* bank, int nx, int ny, uint64_t master_key, double dt, int ntotal,
              const double* density, const double* edgex, const double* edgey,
              const double* s_keys, const double* s_vals, int s_n, const double* a_keys,
              const double* a_vals, int a_n, double* tally, uint64_t* r0, uint64_t* r1,
              uint64_t* r2, uint64_t* facet_events, uint64_t* collision_events) {
  if (g.pending_bank)
    terminate("solve_transport_2d: the previous timestep was enqueued with defer_finish=1 and "
              "nb200_solve_finish has not been called");
  const uint64_t launches0 = g.launches;
  StepArgs a{};
  a.nx = nx;
  a.ny = ny;
  a.n = bank->n;
  a.master_key = master_key;
  a.pid0 = bank->pid0;
  a.dt = dt;
  a.inv_ntotal = 1.0 / (double)ntotal;  // omp3/neutral.c:120
  a.density = density;
  a.edgex = edgex;
  a.edgey = edgey;
  a.s_keys = s_keys;
  a.s_vals = s_vals;
  a.s_n = s_n;
  a.a_keys = a_keys;
  a.a_vals = a_vals;
  a.a_n = a_n;
  a.same_keys = inspect_tables(s_keys, s_n, a_keys, a_n);
  a.tally = tally;
  a.p_facets = (unsigned long long*)r0;
  a.p_collisions = (unsigned long long*)r1;
  a.p_census = (unsigned long long*)r2;
  a.totals = g.d_totals;
  a.bank = bank->cur;
  a.logt = g.d_logt;

  CU_FATAL(cudaMemsetAsync(g.d_totals, 0, sizeof(unsigned long long) * kTotCount, g.stream));
  CU_FATAL(cudaEventRecord(g.ev_begin, g.stream));
  if (g.opt_pipeline) {
    if (g.mesh_ex != edgex || g.mesh_ey != edgey) {  // extent of the mesh (a scheduling hint)
      bool known = false;
      for (const auto& m : g.mesh_cache)
        if (m.ex == edgex && m.ey == edgey && m.nx == nx && m.ny == ny) {
          g.mesh_width = m.width;
          g.mesh_height = m.height;
          known = true;
        }
      if (!known) {
        double ex[2], ey[2];
        CU_FATAL(cudaMemcpyAsync(&ex[0], edgex, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
        CU_FATAL(cudaMemcpyAsync(&ex[1], edgex + nx, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
        CU_FATAL(cudaMemcpyAsync(&ey[0], edgey, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
        CU_FATAL(cudaMemcpyAsync(&ey[1], edgey + ny, sizeof(double), cudaMemcpyDeviceToHost, g.stream));
        CU_FATAL(cudaStreamSynchronize(g.stream));
        g.mesh_width = ex[1] > ex[0] ? ex[1] - ex[0] : 1.0;
        g.mesh_height = ey[1] > ey[0] ? ey[1] - ey[0] : 1.0;
        if (g.mesh_cache.size() >= 8) g.mesh_cache.erase(g.mesh_cache.begin());
        g.mesh_cache.push_back({edgex, edgey, nx, ny, g.mesh_width, g.mesh_height});
      }
      g.mesh_ex = edgex;
      g.mesh_ey = edgey;
    }
    if (g.opt_l2_persist && !g.l2_limit_set) {  // set-aside for the persisting window
      size_t want = 4u << 20;
      if (g.opt_l2_persist == 2) {  // experiment: persist (part of) the tally instead
        int max_bytes = 0;
        CU_FATAL(cudaDeviceGetAttribute(&max_bytes, cudaDevAttrMaxPersistingL2CacheSize, g.device));
        want = (size_t)max_bytes;
      }
      CU_FATAL(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
      CU_FATAL(cudaDeviceGetLimit(&g.l2_setaside, cudaLimitPersistingL2CacheSize));
      g.l2_limit_set = true;
    }
    // P0: restage the read-only inputs (cross-section tables, density tile map)
    const bool overlap = g.opt_stage_overlap != 0;
    if (overlap) {
      CU_FATAL(cudaEventRecord(g.ev_fork, g.stream));
      CU_FATAL(cudaStreamWaitEvent(g.stage_stream, g.ev_fork, 0));
      stage_tiles(a, g.stage_stream);
      CU_FATAL(cudaEventRecord(g.ev_tiles, g.stage_stream));
    }
    stage_tables(a);
    if (!overlap) stage_tiles(a, g.stream);
    // P1-P3: begin-step set-up, classification and counting sort into the double buffer
    SortArgs s{};
    s.tile_shift = g.opt_tile_shift;
    s.tiles_x = s.tile_shift >= 0 ? ((nx - 1) >> s.tile_shift) + 1 : 1;
    s.ntiles = s.tile_shift >= 0 ? s.tiles_x * (((ny - 1) >> s.tile_shift) + 1) : 1;
    s.nq = g.opt_length_bins > 1 ? g.opt_length_bins : 1;
    s.q_scale = 24.0f;  // 12 bins across the sqrt(2) spread of facet counts with direction
    s.inv_dx = (float)((double)nx / g.mesh_width);
    s.inv_dy = (float)((double)ny / g.mesh_height);
    s.nbins = 3 * s.nq * s.ntiles + 1;
    s.n_upper = bank->n_upper;
    s.n = bank->n;
    if (!bank->has_alt) {
      bank_alloc(bank->alt, bank->n);
      device_zalloc(&bank->keys, (size_t)bank->n);
      bank->has_alt = true;
    }
    if (g.bins_capacity < s.nbins) {
      cudaFree(g.d_bins);
      // histogram, cursors, and one scan partial per 2048 bins
      CU_FATAL(cudaMalloc(&g.d_bins, sizeof(unsigned) * (2 * (size_t)s.nbins + s.nbins / 2048 + 1)));
      g.bins_capacity = s.nbins;
    }
    s.keys = bank->keys;
    s.bin_count = g.d_bins;
    s.bin_cursor = g.d_bins + s.nbins;
    s.chunk_sum = g.d_bins + 2 * (size_t)s.nbins;
    s.n_live = g.d_n_live;
    g.launches += launch_sort_phase(a, s, bank->alt, g.stream);
    std::swap(bank->cur, bank->alt);
    a.bank = bank->cur;
    if (overlap) CU_FATAL(cudaStreamWaitEvent(g.stream, g.ev_tiles, 0));
    CU_FATAL(cudaEventRecord(g.ev_mid, g.stream));
    // P4: event loop over the sorted live prefix
    g.launches += launch_history(a, g.d_n_live, s.n_upper, g.opt_fast_div != 0,
                                 g.opt_tally_prereduce != 0,
                                 g.opt_l2_persist == 2 ? (const void*)tally
                                 : g.opt_l2_persist   ? (const void*)g.d_cs_stage : nullptr,
                                 g.opt_l2_persist == 2 ? sizeof(double) * (size_t)nx * ny
                                                       : g.cs_stage_bytes,
                                 g.opt_l2_persist == 2 ? g.l2_setaside : 0,
                                 g.opt_history_smem_pad, g.stream);
  } else {
    CU_FATAL(cudaEventRecord(g.ev_mid, g.stream));
    g.launches += launch_history_direct(a, g.stream);
  }
  CU_FATAL(cudaGetLastError());
  CU_FATAL(cudaEventRecord(g.ev_end, g.stream));
  k_publish_totals<<<1, 32, 0, g.stream>>>(g.d_totals, g.h_totals_dev);
  g.launches += 1;
  g.pending_bank = bank;
  g.pending_launches0 = launches0;
  // "defer_finish": the caller overlaps its own host work (launching the collective that
  // combines this step's tally delta, say) with the step and collects the counts later
  if (!g.opt_defer_finish) finish_step(facet_events, collision_events);
}

// Second half of a timestep: waits for the stream, checks the device-side faults, updates
// the live-prefix bound and hands the counts to the caller (omp3/neutral.c:202-205).
void finish_step(uint64_t* facet_events, uint64_t* collision_events) {
  Bank* bank = g.pending_bank;
  if (!bank) return;
  g.pending_bank = nullptr;
  const uint64_t launches0 = g.pending_launches0;
  CU_FATAL(cudaStreamSynchronize(g.stream));
  if (g.h_totals[kTotFault])
    terminate("solve_transport_2d: the cross-section tables are not what they were when first "
              "seen (energy grid not strictly increasing, or the two grids no longer equal)");
  // The sort compacted every particle that was dead at the start of this step behind the
  // live prefix; particles that died DURING the step still sit inside the prefix (the next
  // sort moves them out), so the prefix to visit next step is this step's live count.
  if (g.opt_pipeline) bank->n_upper = (int)g.h_totals[kTotProcessed];

  for (int k = 0; k < kTotCount; ++k) g.last_stats[k] = g.h_totals[k];
  g.last_stats[5] = g.launches - launches0;
  float kernel_ms = 0.0f;
  CU_FATAL(cudaEventElapsedTime(&kernel_ms, g.ev_mid, g.ev_end));
  g.last_stats[6] = (uint64_t)((double)kernel_ms * 1.0e6);  // history kernel, ns on the stream
  CU_FATAL(cudaEventElapsedTime(&kernel_ms, g.ev_begin, g.ev_mid));
  g.last_stats[7] = (uint64_t)((double)kernel_ms * 1.0e6);  // sort phase, ns
  *facet_events += g.h_totals[kTotFacets];          // omp3/neutral.c:202
  *collision_events += g.h_totals[kTotCollisions];  // omp3/neutral.c:203
  if (g.opt_print) {
    printf("Particles  %llu\n", (unsigned long long)g.h_totals[kTotProcessed]);  // :205
  }
}

void check_boundary_args(const char* who, int nx, int ny, int global_nx, int global_ny,
                         int pad, int x_off, int y_off) {
  if (pad != 0 || x_off != 0 || y_off != 0 || nx != global_nx || ny != global_ny) {
    terminate("%s: only the single-rank mesh of main.c:34,42-43 is supported "
              "(pad=%d x_off=%d y_off=%d nx=%d/%d ny=%d/%d)",
              who, pad, x_off, y_off, nx, global_nx, ny, global_ny);
  }
}

// problems/neutral.tests lookup: "<params_filename> result=<value>"
bool find_expected(const char* tests_file, const char* key, double* value) {
  FILE* fp = fopen(tests_file, "r");
  if (!fp) return false;
  char line[4096];
  bool found = false;
  while (fgets(line, sizeof(line), fp)) {
    char name[2048];
    int off = 0;
    if (sscanf(line, "%2047s%n", name, &off) != 1 || name[0] == '#') continue;
    if (strcmp(name, key) != 0) continue;
    const char* eq = strchr(line + off, '=');
    if (!eq) continue;
    *value = strtod(eq + 1, nullptr);
    found = true;
    break;
  }
  fclose(fp);
  return found;
}

__global__ void k_partial_sums(const double* __restrict__ v, size_t n, double* partial) {
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    s += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double warp_part[32];
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? warp_part[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}

double device_sum(const double* v, size_t n) {
  const int blocks = 592, threads = 256;
  double* d_partial = nullptr;
  CU_FATAL(cudaMalloc(&d_partial, sizeof(double) * blocks));
  k_partial_sums<<<blocks, threads, 0, g.stream>>>(v, n, d_partial);
  g.launches++;
  std::vector<double> h(blocks);
  CU_FATAL(cudaMemcpyAsync(h.data(), d_partial, sizeof(double) * blocks,
                           cudaMemcpyDeviceToHost, g.stream));
  CU_FATAL(cudaStreamSynchronize(g.stream));
  cudaFree(d_partial);
  double s = 0.0;
  for (double x : h) s += x;
  return s;
}

}  // namespace

// ======================================================================================
// 1. plugin boundary, device-resident flavour
// ======================================================================================
extern "C" void solve_transport_2d(
    const int nx, const int ny, const int global_nx, const int global_ny,
    const uint64_t master_key, const int pad, const int x_off, const int y_off,
    const double dt, const int ntotal_particles, int* nlocal_particles,
    const int* neighbours, nb200_particle_soa* particles, const double* density,
    const double* edgex, const double* edgey, const double* edgedx, const double* edgedy,
    nb200_cross_section* cs_scatter_table, nb200_cross_section* cs_absorb_table,
    double* energy_deposition_tally, uint64_t* reduce_array0, uint64_t* reduce_array1,
    uint64_t* reduce_array2, uint64_t* facet_events, uint64_t* collision_events) {
  (void)neighbours; (void)edgedx; (void)edgedy;
  if (!(*nlocal_particles)) {  // omp3/neutral.c:30-33
    printf("Out of particles\n");
    return;
  }
  require_ready();
  check_boundary_args("solve_transport_2d", nx, ny, global_nx, global_ny, pad, x_off, y_off);
  Bank* bank = bank_of(particles);
  if (!bank) terminate("solve_transport_2d: `particles` is not a bank created by this "
                       "kernel set's inject_particles / nb200_bank_create");
  run_step(bank, nx, ny, master_key, dt, ntotal_particles, density, edgex, edgey,
           cs_scatter_table->keys, cs_scatter_table->values, cs_scatter_table->nentries,
           cs_absorb_table->keys, cs_absorb_table->values, cs_absorb_table->nentries,
           energy_deposition_tally, reduce_array0, reduce_array1, reduce_array2,
           facet_events, collision_events);
}

extern "C" size_t inject_particles(
    const int nparticles, const int global_nx, const int local_nx, const int local_ny,
    const int pad, const double local_particle_left_off,
    const double local_particle_bottom_off, const double local_particle_width,
    const double local_particle_height, const int x_off, const int y_off, const double dt,
    const double* edgex, const double* edgey, const double initial_energy,
    nb200_particle_soa** particles) {
  require_ready();
  check_boundary_args("inject_particles", local_nx, local_ny, global_nx, local_ny, pad, x_off,
                      y_off);
  const int first = g.shard_count >= 0 ? g.shard_first : 0;
  const int count = g.shard_count >= 0 ? g.shard_count : nparticles;
  if (first < 0 || first + count > nparticles)
    terminate("inject_particles: shard [%d, %d) outside [0, %d)", first, first + count,
              nparticles);

  if (g.opt_device_inject) {
    // The bank is generated where it lives: no host loop, no 80-byte-per-particle upload.
    size_t bytes = 0;
    BankHandle* h = new_handle(count, (uint64_t)first, &bytes);
    InjectArgs ia{edgex, edgey, local_nx, local_ny, local_particle_left_off,
                  local_particle_bottom_off, local_particle_width, local_particle_height, dt,
                  initial_energy};
    g.launches += launch_inject(h->impl->cur, count, (uint64_t)first, ia, g.d_sct, g.stream);
    CU_FATAL(cudaGetLastError());
    CU_FATAL(cudaStreamSynchronize(g.stream));
    *particles = &h->view;
    return bytes;
  }

  // Host flavour ("device_inject" = 0): the same arithmetic with the host's libm.
  // The mesh edges live in kernel-set (device) memory.
  std::vector<double> ex(local_nx + 1), ey(local_ny + 1);
  CU_FATAL(cudaMemcpy(ex.data(), edgex, sizeof(double) * (local_nx + 1),
                      cudaMemcpyDeviceToHost));
  CU_FATAL(cudaMemcpy(ey.data(), edgey, sizeof(double) * (local_ny + 1),
                      cudaMemcpyDeviceToHost));

  const size_t n = (size_t)std::max(count, 1);
  std::vector<double> x(n), y(n), ox(n), oy(n), en(n), wt(n), dtc(n), mfp(n);
  std::vector<int> cx(n), cy(n), dead(n);

  // First cell whose half-open interval holds v; 0 when there is none - what the
  // reference's linear scan returns (omp3/neutral.c:590-603), found by bisection.
  auto locate = [](const std::vector<double>& edge, int ncells, double v) {
    int lo = 0, hi = ncells;
    if (!(v >= edge[0])) return 0;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (v >= edge[mid]) lo = mid; else hi = mid;
    }
    return (v >= edge[lo] && v < edge[lo + 1]) ? lo : 0;
  };

#pragma omp parallel for schedule(static)
  for (int s = 0; s < count; ++s) {
    const uint64_t kk = (uint64_t)(first + s);
    double r0, r1;
    random_pair(kk, 0, 0, r0, r1);  // omp3/neutral.c:581
    x[s] = local_particle_left_off + r0 * local_particle_width;
    y[s] = local_particle_bottom_off + r1 * local_particle_height;
    cx[s] = locate(ex, local_nx, x[s]);
    cy[s] = locate(ey, local_ny, y[s]);
    random_pair(kk, 0, 1, r0, r1);  // omp3/neutral.c:611
    const double theta = 2.0 * M_PI * r0;
    ox[s] = cos(theta);
    oy[s] = sin(theta);
    en[s] = initial_energy;
    wt[s] = 1.0;
    dtc[s] = dt;
    mfp[s] = 0.0;
    dead[s] = 0;
  }

  size_t bytes = 0;
  BankHandle* h = new_handle(count, (uint64_t)first, &bytes);
  SoaView host{x.data(), y.data(), ox.data(), oy.data(), en.data(), wt.data(), dtc.data(),
               mfp.data(), cx.data(), cy.data(), dead.data()};
  upload_host_soa(host, count, h->impl->cur);
  *particles = &h->view;
  return bytes;
}

extern "C" void validate(const int nx, const int ny, const char* params_filename,
                         const int rank, double* energy_tally) {
  require_ready();
  const double total = device_sum(energy_tally, (size_t)nx * ny);
  if (rank != 0) return;
  printf("\nFinal global_energy_tally %.15e\n", total);
  double expected = 0.0;
  if (!find_expected("problems/neutral.tests", params_filename, &expected)) {
    printf("Warning. Test entry was not found, could NOT validate.\n");
    return;
  }
  printf("Expected %.12e, result was %.12e.\n", expected, total);
  const double scale = fabs(expected) > 0.0 ? fabs(expected) : 1.0;
  if (fabs(expected - total) / scale < 1.0e-3) {  // VALIDATE_TOLERANCE, neutral_data.h:27
    printf("PASSED validation.\n");
  } else {
    printf("FAILED validation.\n");
  }
}

// ======================================================================================
// 2. allocation layer
// ======================================================================================
extern "C" size_t allocate_data(double** buf, size_t len) { return device_zalloc(buf, len); }
extern "C" size_t allocate_float_data(float** buf, size_t len) { return device_zalloc(buf, len); }
extern "C" size_t allocate_int_data(int** buf, size_t len) { return device_zalloc(buf, len); }
extern "C" size_t allocate_uint64_data(uint64_t** buf, size_t len) {
  return device_zalloc(buf, len);
}

extern "C" void allocate_host_data(double** buf, size_t len) {
  require_ready();
  CU_FATAL(cudaMallocHost((void**)buf, sizeof(double) * (len ? len : 1)));
  memset(*buf, 0, sizeof(double) * len);
}

extern "C" void allocate_host_float_data(float** buf, size_t len) {
  require_ready();
  CU_FATAL(cudaMallocHost((void**)buf, sizeof(float) * (len ? len : 1)));
  memset(*buf, 0, sizeof(float) * len);
}

extern "C" void deallocate_data(double* buf) { cudaFree(buf); }
extern "C" void deallocate_host_data(double* buf) { cudaFreeHost(buf); }

extern "C" void copy_buffer(const size_t len, double** src, double** dst, int send) {
  require_ready();
  CU_FATAL(cudaMemcpyAsync(*dst, *src, sizeof(double) * len,
                           send ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, g.stream));
  CU_FATAL(cudaStreamSynchronize(g.stream));
}

extern "C" void move_host_buffer_to_device(const size_t len, double** src, double** dst) {
  device_zalloc(dst, len);
  CU_FATAL(cudaMemcpyAsync(*dst, *src, sizeof(double) * len, cudaMemcpyHostToDevice, g.stream));
  CU_FATAL(cudaStreamSynchronize(g.stream));
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, *src) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
    cudaFreeHost(*src);
  } else {
    (void)cudaGetLastError();
    free(*src);
  }
  *src = nullptr;
}

extern "C" void initialise_devices(int rank) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
    terminate("no CUDA device available: the b200 kernel set has no CPU fallback");
  CU_FATAL(cudaSetDevice(rank % count));
  require_ready();
  cudaDeviceProp prop{};
  CU_FATAL(cudaGetDeviceProperties(&prop, rank % count));
  printf("Rank %d using GPU %d: %s (sm_%d%d, %d SMs, %.0f GB)\n", rank, rank % count, prop.name,
         prop.major, prop.minor, prop.multiProcessorCount,
         (double)prop.totalGlobalMem / (1024.0 * 1024.0 * 1024.0));
}

// ======================================================================================
// 3. host-buffer flavour
// ======================================================================================
extern "C" void nb200_solve_transport_2d_host(
    const int nx, const int ny, const int global_nx, const int global_ny,
    const uint64_t master_key, const int pad, const int x_off, const int y_off,
    const double dt, const int ntotal_particles, int* nlocal_particles,
    const int* neighbours, nb200_particle_aos* particles, const double* density,
    const double* edgex, const double* edgey, const double* edgedx, const double* edgedy,
    nb200_cross_section* cs_scatter_table, nb200_cross_section* cs_absorb_table,
    double* energy_deposition_tally, uint64_t* reduce_array0, uint64_t* reduce_array1,
    uint64_t* reduce_array2, uint64_t* facet_events, uint64_t* collision_events) {
  (void)neighbours; (void)edgedx; (void)edgedy;
  static_assert(sizeof(nb200_particle_aos) == 80, "AoS particle must be 80 bytes");
  const int n = *nlocal_particles;
  if (!n) {
    printf("Out of particles\n");
    return;
  }
  require_ready();
  check_boundary_args("nb200_solve_transport_2d_host", nx, ny, global_nx, global_ny, pad,
                      x_off, y_off);
  const size_t ncells = (size_t)nx * ny;
  const int s_n = cs_scatter_table->nentries, a_n = cs_absorb_table->nentries;

  auto upload = [&](const void* host, size_t bytes) {
    void* d = nullptr;
    CU_FATAL(cudaMalloc(&d, bytes ? bytes : 1));
    CU_FATAL(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, g.stream));
    return d;
  };
  double* d_density = (double*)upload(density, sizeof(double) * ncells);
  double* d_edgex = (double*)upload(edgex, sizeof(double) * (nx + 1));
  double* d_edgey = (double*)upload(edgey, sizeof(double) * (ny + 1));
  double* d_sk = (double*)upload(cs_scatter_table->keys, sizeof(double) * s_n);
  double* d_sv = (double*)upload(cs_scatter_table->values, sizeof(double) * s_n);
  double* d_ak = (double*)upload(cs_absorb_table->keys, sizeof(double) * a_n);
  double* d_av = (double*)upload(cs_absorb_table->values, sizeof(double) * a_n);
  double* d_tally = (double*)upload(energy_deposition_tally, sizeof(double) * ncells);
  void* d_aos = upload(particles, sizeof(nb200_particle_aos) * (size_t)n);
  uint64_t* d_r[3] = {nullptr, nullptr, nullptr};
  uint64_t* h_r[3] = {reduce_array0, reduce_array1, reduce_array2};
  for (int k = 0; k < 3; ++k)
    if (h_r[k]) d_r[k] = (uint64_t*)upload(h_r[k], sizeof(uint64_t) * (size_t)n);

  Bank bank;
  bank.n = n;
  bank.n_upper = n;
  bank.pid0 = 0;
  bank_alloc(bank.cur, n);
  g.launches += launch_import_aos(bank.cur, d_aos, n, g.stream);
  g.cs_s_keys = nullptr;  // fresh uploads: never trust the cached table inspection
  run_step(&bank, nx, ny, master_key, dt, ntotal_particles, d_density, d_edgex, d_edgey, d_sk,
           d_sv, s_n, d_ak, d_av, a_n, d_tally, d_r[0], d_r[1], d_r[2], facet_events,
           collision_events);
  finish_step(facet_events, collision_events);  // the host flavour never defers (no-op if done)
  g.cs_s_keys = nullptr;
  g.launches += launch_export_aos(bank.cur, d_aos, n, g.stream);
  CU_FATAL(cudaMemcpyAsync(particles, d_aos, sizeof(nb200_particle_aos) * (size_t)n,
                           cudaMemcpyDeviceToHost, g.stream));
  CU_FATAL(cudaMemcpyAsync(energy_deposition_tally, d_tally, sizeof(double) * ncells,
                           cudaMemcpyDeviceToHost, g.stream));
  for (int k = 0; k < 3; ++k)
    if (h_r[k])
      CU_FATAL(cudaMemcpyAsync(h_r[k], d_r[k], sizeof(uint64_t) * (size_t)n,
                               cudaMemcpyDeviceToHost, g.stream));
  CU_FATAL(cudaStreamSynchronize(g.stream));
  bank_release(bank.cur);
  if (bank.has_alt) {
    bank_release(bank.alt);
    cudaFree(bank.keys);
  }
  void* to_free[] = {d_density, d_edgex, d_edgey, d_sk, d_sv, d_ak, d_av, d_tally, d_aos,
                     d_r[0], d_r[1], d_r[2]};
  for (void* p : to_free) cudaFree(p);
}

// ======================================================================================
// 4. extensions
// ======================================================================================
extern "C" int nb200_abi_version(void) { return NB200_ABI_VERSION; }
extern "C" const char* nb200_last_error(void) { return g.last_error.c_str(); }

extern "C" int nb200_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return count;
}

extern "C" int nb200_set_stream(void* cuda_stream) {
  g.stream = (cudaStream_t)cuda_stream;
  return 0;
}

extern "C" int nb200_set_shard(int first, int count) {
  g.shard_first = first;
  g.shard_count = count;
  return 0;
}

extern "C" int nb200_bank_create(const nb200_particle_soa* host, int count, int pid_first,
                                 nb200_particle_soa** particles) {
  if (ensure_ready() != 0) return -1;
  if (!host || !particles || count < 0) {
    set_error("nb200_bank_create: bad arguments");
    return -3;
  }
  BankHandle* h = new_handle(count, (uint64_t)pid_first, nullptr);
  if (count > 0) upload_host_soa(as_soa_view(*host), count, h->impl->cur);
  *particles = &h->view;
  return 0;
}

extern "C" int nb200_bank_upload(nb200_particle_soa* particles, const nb200_particle_soa* host) {
  if (ensure_ready() != 0) return -1;
  Bank* bank = bank_of(particles);
  if (!bank || !host) {
    set_error("nb200_bank_upload: not a bank handle");
    return -3;
  }
  if (!bank->has_export) {
    soa_alloc(bank->exported, bank->n);
    bank->has_export = true;
    reinterpret_cast<BankHandle*>(particles)->view = as_public(bank->exported);
  }
  // The plain SoA view doubles as the staging area: H2D per field, then one repack kernel.
  const SoaView& s = bank->exported;
  const size_t n = (size_t)bank->n;
  const double* hd[] = {host->x, host->y, host->omega_x, host->omega_y, host->energy,
                        host->weight, host->dt_to_census, host->mfp_to_collision};
  double* dd[] = {s.x, s.y, s.omega_x, s.omega_y, s.energy, s.weight, s.dt_to_census,
                  s.mfp_to_collision};
  for (int k = 0; k < 8; ++k)
    CU_TRY(cudaMemcpyAsync(dd[k], hd[k], sizeof(double) * n, cudaMemcpyHostToDevice, g.stream));
  const int* hi[] = {host->cellx, host->celly, host->dead};
  int* di[] = {s.cellx, s.celly, s.dead};
  for (int k = 0; k < 3; ++k)
    CU_TRY(cudaMemcpyAsync(di[k], hi[k], sizeof(int) * n, cudaMemcpyHostToDevice, g.stream));
  g.launches += launch_import_soa(bank->cur, s, bank->n, 0, g.stream);
  bank->n_upper = bank->n;
  return 0;
}

// The plain device SoA view behind the handle (allocated on first use). Together with
// nb200_bank_import / nb200_bank_export it lets a host move banks with its own asynchronous
// copies: fill the view, then import; export, then read the view.
extern "C" int nb200_bank_view(nb200_particle_soa* particles, nb200_particle_soa* view_out) {
  if (ensure_ready() != 0) return -1;
  Bank* bank = bank_of(particles);
  if (!bank || !view_out) {
    set_error("nb200_bank_view: not a bank handle");
    return -3;
  }
  if (!bank->has_export) {
    soa_alloc(bank->exported, bank->n);
    bank->has_export = true;
    reinterpret_cast<BankHandle*>(particles)->view = as_public(bank->exported);
  }
  *view_out = as_public(bank->exported);
  return 0;
}

// Rebuilds the bank from its plain SoA view (asynchronous on the library's stream).
extern "C" int nb200_bank_import(nb200_particle_soa* particles) {
  if (ensure_ready() != 0) return -1;
  Bank* bank = bank_of(particles);
  if (!bank || !bank->has_export) {
    set_error("nb200_bank_import: not a bank handle, or its view was never requested");
    return -3;
  }
  g.launches += launch_import_soa(bank->cur, bank->exported, bank->n, 0, g.stream);
  bank->n_upper = bank->n;
  return 0;
}

extern "C" int nb200_accumulate(double* dst_device, const double* src_device, size_t n) {
  if (ensure_ready() != 0) return -1;
  g.launches += launch_accumulate(dst_device, src_device, n, g.stream);
  return 0;
}

extern "C" int nb200_accumulate_clear(double* dst_device, double* src_device, size_t n) {
  if (ensure_ready() != 0) return -1;
  g.launches += launch_accumulate_clear(dst_device, src_device, n, g.stream);
  return 0;
}

extern "C" int nb200_accumulate_clear_async(double* dst_device, double* src_device, size_t n,
                                            void* cuda_stream) {
  if (ensure_ready() != 0) return -1;
  g.launches += launch_accumulate_clear(dst_device, src_device, n, (cudaStream_t)cuda_stream);
  return 0;
}

extern "C" int nb200_bank_export(nb200_particle_soa* particles) {
  if (ensure_ready() != 0) return -1;
  Bank* bank = bank_of(particles);
  if (!bank) {
    set_error("nb200_bank_export: not a bank handle");
    return -3;
  }
  if (!bank->has_export) {
    soa_alloc(bank->exported, bank->n);
    bank->has_export = true;
    reinterpret_cast<BankHandle*>(particles)->view = as_public(bank->exported);
  }
  g.launches += launch_export_soa(bank->cur, bank->exported, bank->n, g.stream);
  CU_TRY(cudaStreamSynchronize(g.stream));
  return 0;
}

extern "C" int nb200_bank_download(nb200_particle_soa* particles, nb200_particle_soa* host) {
  const int rc = nb200_bank_export(particles);
  if (rc != 0) return rc;
  Bank* bank = bank_of(particles);
  const SoaView& s = bank->exported;
  const size_t n = (size_t)bank->n;
  const double* dd[] = {s.x, s.y, s.omega_x, s.omega_y, s.energy, s.weight, s.dt_to_census,
                        s.mfp_to_collision};
  double* hd[] = {host->x, host->y, host->omega_x, host->omega_y, host->energy, host->weight,
                  host->dt_to_census, host->mfp_to_collision};
  for (int k = 0; k < 8; ++k)
    CU_TRY(cudaMemcpyAsync(hd[k], dd[k], sizeof(double) * n, cudaMemcpyDeviceToHost, g.stream));
  const int* di[] = {s.cellx, s.celly, s.dead};
  int* hi[] = {host->cellx, host->celly, host->dead};
  for (int k = 0; k < 3; ++k)
    CU_TRY(cudaMemcpyAsync(hi[k], di[k], sizeof(int) * n, cudaMemcpyDeviceToHost, g.stream));
  CU_TRY(cudaStreamSynchronize(g.stream));
  return 0;
}

extern "C" int nb200_bank_copy(nb200_particle_soa* dst, nb200_particle_soa* src) {
  if (ensure_ready() != 0) return -1;
  Bank* d = bank_of(dst);
  Bank* s = bank_of(src);
  if (!d || !s || d->n != s->n) {
    set_error("nb200_bank_copy: handles must be banks of the same size");
    return -3;
  }
  const size_t n = (size_t)s->n;
  CU_TRY(cudaMemcpyAsync(d->cur.pos, s->cur.pos, sizeof(double2) * n, cudaMemcpyDeviceToDevice, g.stream));
  CU_TRY(cudaMemcpyAsync(d->cur.dir, s->cur.dir, sizeof(double2) * n, cudaMemcpyDeviceToDevice, g.stream));
  CU_TRY(cudaMemcpyAsync(d->cur.ew, s->cur.ew, sizeof(double2) * n, cudaMemcpyDeviceToDevice, g.stream));
  CU_TRY(cudaMemcpyAsync(d->cur.tm, s->cur.tm, sizeof(double2) * n, cudaMemcpyDeviceToDevice, g.stream));
  CU_TRY(cudaMemcpyAsync(d->cur.meta, s->cur.meta, sizeof(int4) * n, cudaMemcpyDeviceToDevice, g.stream));
  d->pid0 = s->pid0;
  d->n_upper = s->n_upper;
  return 0;
}

extern "C" int nb200_bank_size(nb200_particle_soa* particles) {
  Bank* bank = bank_of(particles);
  return bank ? bank->n : -3;
}

extern "C" int nb200_bank_free(nb200_particle_soa* particles) {
  Bank* bank = bank_of(particles);
  if (!bank) return -3;
  bank_release(bank->cur);
  if (bank->has_alt) {
    bank_release(bank->alt);
    cudaFree(bank->keys);
  }
  if (bank->has_export) soa_release(bank->exported);
  BankHandle* h = reinterpret_cast<BankHandle*>(particles);
  h->magic = 0;
  delete bank;
  free(h);
  return 0;
}

extern "C" int nb200_memcpy_h2d(void* dst_device, const void* src_host, size_t bytes) {
  if (ensure_ready() != 0) return -1;
  CU_TRY(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, g.stream));
  CU_TRY(cudaStreamSynchronize(g.stream));
  return 0;
}

extern "C" int nb200_memcpy_d2h(void* dst_host, const void* src_device, size_t bytes) {
  if (ensure_ready() != 0) return -1;
  CU_TRY(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, g.stream));
  CU_TRY(cudaStreamSynchronize(g.stream));
  return 0;
}

extern "C" int nb200_memcpy_h2d_async(void* dst_device, const void* src_host, size_t bytes,
                                      void* cuda_stream) {
  if (ensure_ready() != 0) return -1;
  CU_TRY(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice,
                         (cudaStream_t)cuda_stream));
  return 0;
}

extern "C" int nb200_memcpy_d2h_async(void* dst_host, const void* src_device, size_t bytes,
                                      void* cuda_stream) {
  if (ensure_ready() != 0) return -1;
  CU_TRY(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost,
                         (cudaStream_t)cuda_stream));
  return 0;
}

extern "C" int nb200_memset_d(void* dst_device, int value, size_t bytes) {
  if (ensure_ready() != 0) return -1;
  CU_TRY(cudaMemsetAsync(dst_device, value, bytes, g.stream));
  return 0;
}

extern "C" int nb200_synchronize(void) {
  if (ensure_ready() != 0) return -1;
  CU_TRY(cudaStreamSynchronize(g.stream));
  return 0;
}

extern "C" int nb200_set_option(const char* name, int value) {
  int* slot = nullptr;
  if (strcmp(name, "print") == 0) slot = &g.opt_print;
  else if (strcmp(name, "pipeline") == 0) slot = &g.opt_pipeline;
  else if (strcmp(name, "fast_div") == 0) slot = &g.opt_fast_div;
  else if (strcmp(name, "tile_shift") == 0) slot = &g.opt_tile_shift;
  else if (strcmp(name, "length_bins") == 0) slot = &g.opt_length_bins;
  else if (strcmp(name, "tally_prereduce") == 0) slot = &g.opt_tally_prereduce;
  else if (strcmp(name, "l2_persist") == 0) slot = &g.opt_l2_persist;
  else if (strcmp(name, "defer_finish") == 0) slot = &g.opt_defer_finish;
  else if (strcmp(name, "device_inject") == 0) slot = &g.opt_device_inject;
  else if (strcmp(name, "stage_overlap") == 0) slot = &g.opt_stage_overlap;
  else if (strcmp(name, "history_smem_pad") == 0) slot = &g.opt_history_smem_pad;
  if (!slot) {
    set_error("nb200_set_option: unknown option '%s'", name);
    return -3;
  }
  const int prev = *slot;
  *slot = value;
  return prev;
}

extern "C" int nb200_solve_finish(uint64_t* facet_events, uint64_t* collision_events) {
  if (!g.pending_bank) {
    set_error("nb200_solve_finish: no timestep is pending");
    return -3;
  }
  uint64_t f = 0, c = 0;
  finish_step(&f, &c);
  if (facet_events) *facet_events += f;
  if (collision_events) *collision_events += c;
  return 0;
}

extern "C" int nb200_last_step_stats(uint64_t out[8]) {
  for (int k = 0; k < 8; ++k) out[k] = g.last_stats[k];
  return 0;
}

extern "C" uint64_t nb200_kernel_launches(void) { return g.launches; }

extern "C" int nb200_selftest_rng_log(uint64_t pkey0, uint64_t master_key, uint64_t counter,
                                      int n, uint64_t* raw_host, double* unit_host,
                                      double* neglog_host) {
  if (ensure_ready() != 0) return -1;
  uint64_t* d_raw = nullptr;
  double *d_unit = nullptr, *d_nl = nullptr;
  CU_TRY(cudaMalloc(&d_raw, sizeof(uint64_t) * 2 * n));
  CU_TRY(cudaMalloc(&d_unit, sizeof(double) * 2 * n));
  CU_TRY(cudaMalloc(&d_nl, sizeof(double) * 2 * n));
  g.launches += launch_selftest_rng_log(pkey0, master_key, counter, n, g.d_logt, d_raw, d_unit,
                                        d_nl, g.stream);
  CU_TRY(cudaMemcpyAsync(raw_host, d_raw, sizeof(uint64_t) * 2 * n, cudaMemcpyDeviceToHost, g.stream));
  CU_TRY(cudaMemcpyAsync(unit_host, d_unit, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, g.stream));
  CU_TRY(cudaMemcpyAsync(neglog_host, d_nl, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, g.stream));
  CU_TRY(cudaStreamSynchronize(g.stream));
  cudaFree(d_raw);
  cudaFree(d_unit);
  cudaFree(d_nl);
  return 0;
}

extern "C" int nb200_selftest_log(const double* x_host, double* y_host, int n) {
  if (ensure_ready() != 0) return -1;
  double *d_x = nullptr, *d_y = nullptr;
  CU_TRY(cudaMalloc(&d_x, sizeof(double) * n));
  CU_TRY(cudaMalloc(&d_y, sizeof(double) * n));
  CU_TRY(cudaMemcpyAsync(d_x, x_host, sizeof(double) * n, cudaMemcpyHostToDevice, g.stream));
  g.launches += launch_selftest_log(d_x, d_y, n, g.d_logt, g.stream);
  CU_TRY(cudaMemcpyAsync(y_host, d_y, sizeof(double) * n, cudaMemcpyDeviceToHost, g.stream));
  CU_TRY(cudaStreamSynchronize(g.stream));
  cudaFree(d_x);
  cudaFree(d_y);
  return 0;
}

extern "C" int nb200_selftest_cs(const double* keys_host, const double* values_host,
                                 int nentries, const double* energies_host, int n,
                                 int* index_host, double* value_host) {
  if (ensure_ready() != 0) return -1;
  double *d_k = nullptr, *d_v = nullptr, *d_e = nullptr, *d_o = nullptr;
  int* d_i = nullptr;
  CU_TRY(cudaMalloc(&d_k, sizeof(double) * nentries));
  CU_TRY(cudaMalloc(&d_v, sizeof(double) * nentries));
  CU_TRY(cudaMalloc(&d_e, sizeof(double) * n));
  CU_TRY(cudaMalloc(&d_o, sizeof(double) * n));
  CU_TRY(cudaMalloc(&d_i, sizeof(int) * n));
  CU_TRY(cudaMemcpyAsync(d_k, keys_host, sizeof(double) * nentries, cudaMemcpyHostToDevice, g.stream));
  CU_TRY(cudaMemcpyAsync(d_v, values_host, sizeof(double) * nentries, cudaMemcpyHostToDevice, g.stream));
  CU_TRY(cudaMemcpyAsync(d_e, energies_host, sizeof(double) * n, cudaMemcpyHostToDevice, g.stream));
  // stage the table exactly as a timestep does, then run both lookups side by side
  const CsParams par = cs_params_from_host(keys_host, nentries);
  double2* d_kv = nullptr;
  int* d_bk = nullptr;
  CU_TRY(cudaMalloc(&d_kv, sizeof(double2) * nentries));
  CU_TRY(cudaMalloc(&d_bk, sizeof(int) * (par.nb + 1)));
  CU_TRY(cudaMemsetAsync(g.d_totals, 0, sizeof(unsigned long long) * kTotCount, g.stream));
  g.launches += launch_stage_cs(d_k, d_v, nentries, d_kv, d_bk, par.bits0, par.shift, par.nb,
                                nullptr, g.d_totals, g.stream);
  const CsStage staged{d_kv, d_bk, par.bits0, par.shift, par.nb, nentries};
  g.launches += launch_selftest_cs(d_k, d_v, nentries, staged, d_e, n, d_i, d_o, g.stream);
  CU_TRY(cudaMemcpyAsync(index_host, d_i, sizeof(int) * n, cudaMemcpyDeviceToHost, g.stream));
  CU_TRY(cudaMemcpyAsync(value_host, d_o, sizeof(double) * n, cudaMemcpyDeviceToHost, g.stream));
  CU_TRY(cudaStreamSynchronize(g.stream));
  cudaFree(d_k); cudaFree(d_v); cudaFree(d_e); cudaFree(d_o); cudaFree(d_i);
  cudaFree(d_kv); cudaFree(d_bk);
  return 0;
}

extern "C" int nb200_selftest_div(const double* a_host, const double* b_host, int n,
                                  double* fast_host, double* ieee_host) {
  if (ensure_ready() != 0) return -1;
  double* d[4] = {nullptr, nullptr, nullptr, nullptr};
  for (auto& p : d) CU_TRY(cudaMalloc(&p, sizeof(double) * n));
  CU_TRY(cudaMemcpyAsync(d[0], a_host, sizeof(double) * n, cudaMemcpyHostToDevice, g.stream));
  CU_TRY(cudaMemcpyAsync(d[1], b_host, sizeof(double) * n, cudaMemcpyHostToDevice, g.stream));
  g.launches += launch_selftest_div(d[0], d[1], d[2], d[3], n, g.stream);
  CU_TRY(cudaMemcpyAsync(fast_host, d[2], sizeof(double) * n, cudaMemcpyDeviceToHost, g.stream));
  CU_TRY(cudaMemcpyAsync(ieee_host, d[3], sizeof(double) * n, cudaMemcpyDeviceToHost, g.stream));
  CU_TRY(cudaStreamSynchronize(g.stream));
  for (auto& p : d) cudaFree(p);
  return 0;
}

extern "C" int nb200_selftest_fastmath(const double* a_host, const double* b_host, int n,
                                       double* out_host) {
  if (ensure_ready() != 0) return -1;
  double *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
  CU_TRY(cudaMalloc(&d_a, sizeof(double) * n));
  CU_TRY(cudaMalloc(&d_b, sizeof(double) * n));
  CU_TRY(cudaMalloc(&d_o, sizeof(double) * 6 * (size_t)n));
  CU_TRY(cudaMemcpyAsync(d_a, a_host, sizeof(double) * n, cudaMemcpyHostToDevice, g.stream));
  CU_TRY(cudaMemcpyAsync(d_b, b_host, sizeof(double) * n, cudaMemcpyHostToDevice, g.stream));
  g.launches += launch_selftest_fastmath(d_a, d_b, d_o, n, g.stream);
  CU_TRY(cudaMemcpyAsync(out_host, d_o, sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost,
                         g.stream));
  CU_TRY(cudaStreamSynchronize(g.stream));
  cudaFree(d_a);
  cudaFree(d_b);
  cudaFree(d_o);
  return 0;
}

extern "C" void nb200_host_threefry2x64_20(uint64_t c0, uint64_t c1, uint64_t k0, uint64_t k1,
                                           uint64_t out[2]) {
  threefry2x64_20(c0, c1, k0, k1, out[0], out[1]);
}

extern "C" double nb200_host_log(double x) { return nb_log(x, &kHostLogTable); }
extern "C" double nb200_host_sin(double x) { return nb_sin(x, &kHostSinCosTable); }
extern "C" double nb200_host_cos(double x) { return nb_cos(x, &kHostSinCosTable); }

// Batch flavour of the two hooks above (the host build of the device source).
extern "C" void nb200_host_sincos(const double* x, long long n, double* s, double* c) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i) {
    s[i] = nb_sin(x[i], &kHostSinCosTable);
    c[i] = nb_cos(x[i], &kHostSinCosTable);
  }
}

// Compares the host build of nb_sin/nb_cos with this process's libm on n arguments; returns
// the number of arguments where either differs in any bit (first offender in *bad_x).
extern "C" long long nb200_selftest_host_sincos(const double* x, long long n, double* bad_x) {
  long long bad = 0;
  double first_bad = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (long long i = 0; i < n; ++i) {
    const double s = nb_sin(x[i], &kHostSinCosTable), c = nb_cos(x[i], &kHostSinCosTable);
    const double rs = sin(x[i]), rc = cos(x[i]);
    if (memcmp(&s, &rs, 8) != 0 || memcmp(&c, &rc, 8) != 0) {
      if (!bad) {
#pragma omp critical
        first_bad = x[i];
      }
      bad++;
    }
  }
  if (bad_x) *bad_x = first_bad;
  return bad;
}

// sin and cos of n host arguments, evaluated on the device.
extern "C" int nb200_selftest_sincos(const double* x_host, double* s_host, double* c_host,
                                     int n) {
  if (ensure_ready() != 0) return -1;
  double *d_x = nullptr, *d_s = nullptr, *d_c = nullptr;
  CU_TRY(cudaMalloc(&d_x, sizeof(double) * n));
  CU_TRY(cudaMalloc(&d_s, sizeof(double) * n));
  CU_TRY(cudaMalloc(&d_c, sizeof(double) * n));
  CU_TRY(cudaMemcpy(d_x, x_host, sizeof(double) * n, cudaMemcpyHostToDevice));
  g.launches += launch_selftest_sincos(d_x, d_s, d_c, n, g.d_sct, g.stream);
  CU_TRY(cudaStreamSynchronize(g.stream));
  CU_TRY(cudaMemcpy(s_host, d_s, sizeof(double) * n, cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(c_host, d_c, sizeof(double) * n, cudaMemcpyDeviceToHost));
  cudaFree(d_x);
  cudaFree(d_s);
  cudaFree(d_c);
  return 0;
}

