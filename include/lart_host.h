/* lart_host.h — C interface of the C++ mini-host.
 *
 * The real host is LaRT's Fortran driver (main.f90, setup.f90, grid_mod_car.f90,
 * observer_rect.f90, output_sum_rect.f90); no Fortran toolchain exists in this
 * image, so this mini-host restates exactly the host-side steps the photon loop
 * needs — namelist subset -> derived parameters -> grid arrays -> observers ->
 * lart_config — and the output normalisation, so that the C ABI in lart_gpu.h
 * can be driven (and tested against the oracle) without Fortran.  It produces
 * HOST arrays only; it contains no transport code and no GPU code.
 */
#ifndef LART_HOST_H
#define LART_HOST_H

#include <stdint.h>

#include "lart_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lart_host_model lart_host_model;

/* scalars the reference prints after grid_create (grid_mod_car.f90:1228-1237) */
typedef struct lart_host_summary {
  double voigt_a, temperature, N_gaspole, N_gashomo, taupole, tauhomo, taupole_dust, tauhomo_dust;
  double Dfreq_ref, vtherm, cross0, atau3, xfreq_min, xfreq_max, dxfreq, dxim, dyim, distance;
  int32_t nx, ny, nz, nxfreq, nobs, nxim, nyim, zonly;
  int64_t nphotons;
  int64_t nclumps;   /* clump medium: N_clumps (0 otherwise) */
} lart_host_summary;

lart_host_model *lart_host_new(void);
void lart_host_free(lart_host_model *m);

/* `par%key = value` exactly as in a LaRT namelist line (setup.f90:27-40);
 * key without the `par%` prefix is accepted too.  Returns non-zero for an
 * unknown key or unparsable value. */
int lart_host_set(lart_host_model *m, const char *key, const char *value);
/* read a whole `&parameters ... /` namelist file (the examples' *.in) */
int lart_host_read_input(lart_host_model *m, const char *path);

/* read_input's derived defaults (setup.f90:42-562) + setup_resonance_line
 * (line_mod.f90:1241-1270) + grid_create (grid_mod_car.f90:11-1238) +
 * observer_create_outside (observer_rect.f90:10-300). */
/* Leaf cells of an octree in the reference's generic AMR format, as generic_amr_read returns them
 * (src/read_generic_amr.f90:52-346: centre, level (root = 0, its children = 1), nH [cm^-3], T [K], bulk velocity
 * [km/s]; box length and lower corner in code units).  Sets par%use_amr_grid; lart_host_setup then builds the tree,
 * the neighbour table and the leaf physics as grid_create_amr does (src/grid_mod_amr.f90:34-526).  The file reader
 * itself stays the reference's.  vx, vy, vz may be NULL (static medium). */
int lart_host_set_amr_leaves(lart_host_model *m, int64_t n, const double *x, const double *y, const double *z, const int32_t *level,
                             const double *nH, const double *T, const double *vx, const double *vy, const double *vz,
                             double boxlen, double origin_x, double origin_y, double origin_z);

int lart_host_setup(lart_host_model *m);

const lart_config *lart_host_config(const lart_host_model *m);
int lart_host_get_summary(const lart_host_model *m, lart_host_summary *out);

/* zero-initialised host tally buffers sized for this model (owned by the model) */
lart_tallies *lart_host_tallies(lart_host_model *m);
int lart_host_zero_tallies(lart_host_model *m);

/* output_normalize_outside (output_sum_rect.f90:151-487): divides the raw sums
 * in place; nscatt_* become per-photon averages. */
int lart_host_normalize(lart_host_model *m);

const char *lart_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
