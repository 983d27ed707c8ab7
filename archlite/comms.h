/*
 * arch-lite comms.h: single-rank stand-ins for the optional MPI wrappers neutral calls
 * (main.c:62,64,75,112,133,194; omp3/neutral.c:530). neutral hard-sets rank = MASTER and
 * nranks = 1 (main.c:42-43) and implements no halo or particle exchange, so these are
 * identities. Multi-GPU scaling in the b200 kernel set is particle sharding, not ranks.
 */
#ifndef ARCHLITE_COMMS_H
#define ARCHLITE_COMMS_H

#include "mesh.h"
#include "shared.h"

#ifdef __cplusplus
extern "C" {
#endif

void initialise_mpi(int argc, char** argv, int* rank, int* nranks);
void initialise_comms(Mesh* mesh);
void finalise_comms(void);
void barrier(void);
double reduce_all_sum(double local_val);
double reduce_all_min(double local_val);
double reduce_all_max(double local_val);

/* VisIt dump of one cell-centred field: <name>.bov + <name>.dat (Brick of Values). */
void write_all_ranks_to_visit(const int global_nx, const int global_ny,
                              const int local_nx, const int local_ny,
                              const int pad, const int x_off, const int y_off,
                              const int rank, const int nranks, int* neighbours,
                              double* local_arr, const char* name, const int tt,
                              const double elapsed_sim_time);

#ifdef __cplusplus
}
#endif

#endif
