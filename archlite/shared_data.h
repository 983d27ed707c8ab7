/*
 * arch-lite shared_data.h: the cell-centred fields shared between arch mini-apps; neutral
 * only reads `density` (main.c:66-71,105).
 *
 * initialise_shared_data_2d paints density from the deck's problem_0, problem_1, ...
 * lines in order: the last four values of a line are xpos ypos width height as fractions
 * of the mesh, a cell belongs to the box when its lower-left edge lies in
 * [pos, pos + size), and later lines override earlier ones. This rule plus
 * width = height = 1.0 reproduces problems/neutral.tests for scatter / stream / csp
 * (SURVEY.md 4.2).
 */
#ifndef ARCHLITE_SHARED_DATA_H
#define ARCHLITE_SHARED_DATA_H

#include "shared.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double* density; /* kernel-set memory, local_nx * local_ny */
  double* energy;  /* unused by neutral */
} SharedData;

void initialise_shared_data_2d(const int local_nx, const int local_ny,
                               const int pad, const double mesh_width,
                               const double mesh_height,
                               const char* problem_def_filename,
                               const double* edgex, const double* edgey,
                               SharedData* shared_data);

#ifdef __cplusplus
}
#endif

#endif
