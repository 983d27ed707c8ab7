/*
 * arch-lite mesh.h: the structured 2-D mesh neutral's driver fills in (main.c:26-44) and
 * hands to initialise_mesh_2d (main.c:65). Uniform spacing; edge i sits at
 * (width / global_nx) * (x_off + i - pad).
 */
#ifndef ARCHLITE_MESH_H
#define ARCHLITE_MESH_H

#ifdef __cplusplus
extern "C" {
#endif

#define NNEIGHBOURS 6
#define EDGE (-1)

enum { NORTH = 0, EAST, SOUTH, WEST, FRONT, BACK };
enum { NO_INVERT = 0, INVERT_X, INVERT_Y };
enum { NO_PACK = 0, PACK };

typedef struct {
  int global_nx, global_ny, global_nz;
  int local_nx, local_ny, local_nz;
  int pad;
  int x_off, y_off, z_off;
  int niters;
  int rank, nranks, ndims;
  int neighbours[NNEIGHBOURS];

  double width, height, depth;
  double dt, dt_h;
  double sim_end;
  double max_dt;

  /* nx+1 / ny+1 edge coordinates and nx / ny spacings; kernel-set memory. */
  double* edgex;
  double* edgey;
  double* edgedx;
  double* edgedy;
} Mesh;

void initialise_mesh_2d(Mesh* mesh);

/* Halo fill / reflective boundary of a cell-centred field: with pad = 0 on one rank
 * (the only configuration neutral uses, main.c:34,42-43) there is nothing to do. */
void handle_boundary_2d(const int nx, const int ny, Mesh* mesh, double* arr,
                        const int invert, const int pack);

#ifdef __cplusplus
}
#endif

#endif
