/*
 * arch-lite: model-independent part (params reader, mesh, density painting, comms and
 * profiler stand-ins). Everything that touches kernel-set memory goes through the
 * allocation layer declared in shared.h, so the same object serves the host (omp3) and
 * the device (b200) builds. See the headers for the call sites each piece serves.
 */
#include "comms.h"
#include "mesh.h"
#include "params.h"
#include "profiler.h"
#include "shared.h"
#include "shared_data.h"

#include <ctype.h>
#include <math.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ params ------ */

#define LINE_LEN 4096

/* Copies the first token of `line` into tok; returns a pointer just past it. */
static const char* first_token(const char* line, char* tok, size_t tok_len) {
  while (*line && isspace((unsigned char)*line)) line++;
  size_t n = 0;
  while (*line && !isspace((unsigned char)*line) && *line != '#') {
    if (n + 1 < tok_len) tok[n++] = *line;
    line++;
  }
  tok[n] = '\0';
  return line;
}

/* Finds the first line named param_name; rest receives the text after the name with any
 * trailing comment removed. */
static int find_line(const char* param_name, const char* filename, char* rest,
                     size_t rest_len) {
  FILE* fp = fopen(filename, "r");
  if (!fp) {
    TERMINATE("Could not open the parameter file: %s", filename);
  }
  char line[LINE_LEN];
  char tok[LINE_LEN];
  int found = 0;
  while (fgets(line, sizeof(line), fp)) {
    const char* after = first_token(line, tok, sizeof(tok));
    if (tok[0] == '\0' || strcmp(tok, param_name) != 0) continue;
    strncpy(rest, after, rest_len - 1);
    rest[rest_len - 1] = '\0';
    char* hash = strchr(rest, '#');
    if (hash) *hash = '\0';
    found = 1;
    break;
  }
  fclose(fp);
  return found;
}

double get_double_parameter(const char* param_name, const char* filename) {
  char rest[LINE_LEN];
  if (!find_line(param_name, filename, rest, sizeof(rest))) {
    TERMINATE("Parameter %s was not found in %s", param_name, filename);
  }
  char* end = NULL;
  const double value = strtod(rest, &end);
  if (end == rest) {
    TERMINATE("Parameter %s in %s has no value", param_name, filename);
  }
  return value;
}

int get_int_parameter(const char* param_name, const char* filename) {
  char rest[LINE_LEN];
  if (!find_line(param_name, filename, rest, sizeof(rest))) {
    TERMINATE("Parameter %s was not found in %s", param_name, filename);
  }
  char* end = NULL;
  const long value = strtol(rest, &end, 10);
  if (end == rest) {
    TERMINATE("Parameter %s in %s has no value", param_name, filename);
  }
  return (int)value;
}

int get_key_value_parameter(const char* param_name, const char* filename,
                            char* keys, double* values, int* nkeys) {
  char rest[LINE_LEN];
  *nkeys = 0;
  if (!find_line(param_name, filename, rest, sizeof(rest))) {
    return 0;
  }
  char* save = NULL;
  for (char* tok = strtok_r(rest, " \t\r\n", &save); tok && *nkeys < MAX_KEYS;
       tok = strtok_r(NULL, " \t\r\n", &save)) {
    char* eq = strchr(tok, '=');
    if (!eq) continue;
    *eq = '\0';
    char* key = &keys[(size_t)(*nkeys) * MAX_STR_LEN];
    strncpy(key, tok, MAX_STR_LEN - 1);
    key[MAX_STR_LEN - 1] = '\0';
    values[*nkeys] = strtod(eq + 1, NULL);
    (*nkeys)++;
  }
  return 1;
}

/* ------------------------------------------------------------------- mesh ------- */

void initialise_mesh_2d(Mesh* mesh) {
  const int nxe = mesh->local_nx + 1;
  const int nye = mesh->local_ny + 1;
  double *ex, *ey, *edx, *edy;
  allocate_host_data(&ex, nxe);
  allocate_host_data(&ey, nye);
  allocate_host_data(&edx, nxe);
  allocate_host_data(&edy, nye);

  const double dx = mesh->width / (double)mesh->global_nx;
  const double dy = mesh->height / (double)mesh->global_ny;
  for (int ii = 0; ii < nxe; ++ii) {
    edx[ii] = dx;
    ex[ii] = dx * (double)(mesh->x_off + ii - mesh->pad);
  }
  for (int ii = 0; ii < nye; ++ii) {
    edy[ii] = dy;
    ey[ii] = dy * (double)(mesh->y_off + ii - mesh->pad);
  }

  move_host_buffer_to_device(nxe, &ex, &mesh->edgex);
  move_host_buffer_to_device(nye, &ey, &mesh->edgey);
  move_host_buffer_to_device(nxe, &edx, &mesh->edgedx);
  move_host_buffer_to_device(nye, &edy, &mesh->edgedy);
}

void handle_boundary_2d(const int nx, const int ny, Mesh* mesh, double* arr,
                        const int invert, const int pack) {
  (void)nx; (void)ny; (void)mesh; (void)arr; (void)invert; (void)pack;
}

/* ------------------------------------------------------------ shared data ------- */

void initialise_shared_data_2d(const int local_nx, const int local_ny,
                               const int pad, const double mesh_width,
                               const double mesh_height,
                               const char* problem_def_filename,
                               const double* edgex, const double* edgey,
                               SharedData* shared_data) {
  /* The edges live in kernel-set memory: stage them on the host. */
  double *hex, *hey;
  allocate_host_data(&hex, local_nx + 1);
  allocate_host_data(&hey, local_ny + 1);
  double* dex = (double*)edgex;
  double* dey = (double*)edgey;
  copy_buffer(local_nx + 1, &dex, &hex, RECV);
  copy_buffer(local_ny + 1, &dey, &hey, RECV);

  double* density;
  allocate_host_data(&density, (size_t)local_nx * local_ny);
  for (size_t ii = 0; ii < (size_t)local_nx * local_ny; ++ii) density[ii] = 0.0;

  char* keys = (char*)malloc(sizeof(char) * MAX_KEYS * MAX_STR_LEN);
  double* values = (double*)malloc(sizeof(double) * MAX_KEYS);
  for (int pp = 0;; ++pp) {
    char name[64];
    snprintf(name, sizeof(name), "problem_%d", pp);
    int nkeys = 0;
    if (!get_key_value_parameter(name, problem_def_filename, keys, values,
                                 &nkeys)) {
      break;
    }
    if (nkeys < 5) {
      TERMINATE("Entry %s of %s needs density and xpos ypos width height", name,
                problem_def_filename);
    }
    double rho = values[0];
    for (int kk = 0; kk < nkeys; ++kk) {
      if (strcmp(&keys[(size_t)kk * MAX_STR_LEN], "density") == 0) rho = values[kk];
    }
    const double xpos = values[nkeys - 4] * mesh_width;
    const double ypos = values[nkeys - 3] * mesh_height;
    const double xend = xpos + values[nkeys - 2] * mesh_width;
    const double yend = ypos + values[nkeys - 1] * mesh_height;
    for (int jj = 0; jj < local_ny; ++jj) {
      if (!(hey[jj] >= ypos && hey[jj] < yend)) continue;
      for (int ii = 0; ii < local_nx; ++ii) {
        if (hex[ii] >= xpos && hex[ii] < xend) {
          density[(size_t)jj * local_nx + ii] = rho;
        }
      }
    }
  }
  free(keys);
  free(values);
  (void)pad;

  move_host_buffer_to_device((size_t)local_nx * local_ny, &density,
                             &shared_data->density);
  shared_data->energy = NULL;
  deallocate_host_data(hex);
  deallocate_host_data(hey);
}

/* ------------------------------------------------------------------ comms ------- */

void initialise_mpi(int argc, char** argv, int* rank, int* nranks) {
  (void)argc; (void)argv;
  *rank = MASTER;
  *nranks = 1;
}

void initialise_comms(Mesh* mesh) {
  for (int ii = 0; ii < NNEIGHBOURS; ++ii) mesh->neighbours[ii] = EDGE;
  mesh->x_off = 0;
  mesh->y_off = 0;
  mesh->z_off = 0;
}

void finalise_comms(void) {}
void barrier(void) {}
double reduce_all_sum(double local_val) { return local_val; }
double reduce_all_min(double local_val) { return local_val; }
double reduce_all_max(double local_val) { return local_val; }

/* VisIt dump of one cell-centred field (main.c:129-139,169-200) as a Brick-of-Values pair,
 * which VisIt opens natively: <name>.bov (text header) + <name>.dat (nx*ny doubles, row
 * major). The field may be kernel-set (device) memory or the driver's own host buffer: it is
 * staged through the allocation layer's copy_buffer either way. Single rank, no padding -
 * the only configuration main.c produces (main.c:34,42-43). */
void write_all_ranks_to_visit(const int global_nx, const int global_ny,
                              const int local_nx, const int local_ny,
                              const int pad, const int x_off, const int y_off,
                              const int rank, const int nranks, int* neighbours,
                              double* local_arr, const char* name, const int tt,
                              const double elapsed_sim_time) {
  (void)global_nx; (void)global_ny; (void)pad; (void)x_off; (void)y_off; (void)nranks;
  (void)neighbours; (void)tt;
  if (rank != MASTER) return;
  const size_t n = (size_t)local_nx * (size_t)local_ny;
  double* staged = NULL;
  allocate_host_data(&staged, n);
  copy_buffer(n, &local_arr, &staged, RECV);
  char path[256];
  snprintf(path, sizeof(path), "%s.dat", name);
  FILE* fp = fopen(path, "wb");
  if (!fp || fwrite(staged, sizeof(double), n, fp) != n) {
    TERMINATE("Could not write the VisIt dump %s", path);
  }
  fclose(fp);
  deallocate_host_data(staged);
  snprintf(path, sizeof(path), "%s.bov", name);
  fp = fopen(path, "w");
  if (!fp) {
    TERMINATE("Could not write the VisIt dump %s", path);
  }
  fprintf(fp, "TIME: %.12e\nDATA_FILE: %s.dat\nDATA_SIZE: %d %d 1\nDATA_FORMAT: DOUBLE\n"
              "VARIABLE: %s\nDATA_ENDIAN: LITTLE\nCENTERING: zonal\n"
              "BRICK_ORIGIN: 0. 0. 0.\nBRICK_SIZE: 1. 1. 1.\n",
          elapsed_sim_time, name, local_nx, local_ny, name);
  fclose(fp);
}

int within_tolerance(const double expected, const double result,
                     const double tolerance) {
  const double scale = fabs(expected) > 0.0 ? fabs(expected) : 1.0;
  return fabs(expected - result) / scale < tolerance;
}

/* --------------------------------------------------------------- profiler ------- */

struct Profile compute_profile;

static double now_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1.0e-9 * (double)ts.tv_nsec;
}

void profiler_start(struct Profile* profile) {
  profile->profiler_start = now_seconds();
}

void profiler_end(struct Profile* profile, const char* entry_name) {
  const double elapsed = now_seconds() - profile->profiler_start;
  if (profile != &compute_profile) {
    /* main.c: uninitialised struct, one unterminated char '0'+tt as the name. */
    const unsigned idx = (unsigned char)(entry_name[0] - '1') % PROFILER_MAX_ENTRIES;
    profile->profiler_entries[idx].time = elapsed;
    profile->profiler_entries[idx].calls = 1;
    return;
  }
  for (int ii = 0; ii < profile->profiler_entry_count; ++ii) {
    ProfileEntry* e = &profile->profiler_entries[ii];
    if (strncmp(e->name, entry_name, PROFILER_MAX_NAME - 1) == 0) {
      e->time += elapsed;
      e->calls++;
      return;
    }
  }
  if (profile->profiler_entry_count < PROFILER_MAX_ENTRIES) {
    ProfileEntry* e = &profile->profiler_entries[profile->profiler_entry_count++];
    strncpy(e->name, entry_name, PROFILER_MAX_NAME - 1);
    e->name[PROFILER_MAX_NAME - 1] = '\0';
    e->time = elapsed;
    e->calls = 1;
  }
}

void profiler_print_full_profile(struct Profile* profile) {
  for (int ii = 0; ii < profile->profiler_entry_count; ++ii) {
    const ProfileEntry* e = &profile->profiler_entries[ii];
    printf("%-32s %.6fs  (%d calls)\n", e->name, e->time, e->calls);
  }
}
