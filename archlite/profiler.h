/*
 * arch-lite profiler.h: the START_PROFILING / STOP_PROFILING pair neutral wraps around
 * solve_transport_2d (main.c:82,99,114-116) and inject_particles (omp3/neutral.c:575,627).
 *
 * Two quirks of the callers shape this implementation:
 *  - main.c passes an UNINITIALISED `struct Profile` and a single, unterminated char
 *    '0'+tt as the entry name, then reads profiler_entries[tt-1].time. So for any
 *    profile other than the global compute_profile the entry index is name[0]-'1' and
 *    the time is assigned, never accumulated.
 *  - compute_profile is a zero-initialised global used with real string names; those
 *    entries are looked up by name and accumulated.
 */
#ifndef ARCHLITE_PROFILER_H
#define ARCHLITE_PROFILER_H

#ifdef __cplusplus
extern "C" {
#endif

#define PROFILER_MAX_NAME 64
#define PROFILER_MAX_ENTRIES 256

typedef struct {
  double time;
  int calls;
  char name[PROFILER_MAX_NAME];
} ProfileEntry;

struct Profile {
  ProfileEntry profiler_entries[PROFILER_MAX_ENTRIES];
  int profiler_entry_count;
  double profiler_start;
};

extern struct Profile compute_profile;

void profiler_start(struct Profile* profile);
void profiler_end(struct Profile* profile, const char* entry_name);
void profiler_print_full_profile(struct Profile* profile);

#ifdef ENABLE_PROFILING
#define START_PROFILING(p) profiler_start(p)
#define STOP_PROFILING(p, name) profiler_end(p, name)
#define PRINT_PROFILING_RESULTS(p) profiler_print_full_profile(p)
#else
#define START_PROFILING(p)
#define STOP_PROFILING(p, name)
#define PRINT_PROFILING_RESULTS(p)
#endif

#ifdef __cplusplus
}
#endif

#endif
