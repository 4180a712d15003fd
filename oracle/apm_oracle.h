/*
 * apm_oracle.h -- CPU oracle for the LCL hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (libapmlcl.so) never links, loads or calls it.
 *
 * Parity status: PINNED against the reference's own golden vectors for the one case the
 * reference pins (test/fortran/fixtures/coverage_test_*-orig.*; see tests/test_oracle_golden.py).
 * Heterogeneous maps, contact cells, LEFT/RIGHT/BOTTOM outlets and both rate-BC modes are NOT pinned
 * by the reference; for those the oracle is cross-checked against an independent
 * scipy sparse-direct solve of the same system in the tests.  gfortran is absent from the authoring
 * image, so oracle/_ref (the compiled reference) cannot be built; this restatement is the CPU arm.
 */
#ifndef APM_ORACLE_H
#define APM_ORACLE_H

#include "../include/apm_lcl.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Everything the parity tests compare, caller-allocated (NULL = skip). */
typedef struct apm_oracle_out {
    double *width;      /* [nz*nx]                WIDTH (m)                       */
    double *trx;        /* [nz][nx+1]             TRX(1:NZ,0:NX) after FLOW_BOUNDARY */
    double *trz;        /* [nz+1][nx]             TRZ(0:NZ,1:NX) after FLOW_BOUNDARY */
    double *point_map;  /* [nz][nx][4]            AP_POINT_MAP (voxels), 1=BLC 2=BRC 3=TRC 4=TLC */
    double *pressure, *flow_x, *flow_z, *flow_m, *reynolds, *residual;   /* [nz*nx] */
    double *q;          /* [5][nz*nx]             Q(1:5,:)                        */
} apm_oracle_out;

/* Full path with the reference's direct D4 solve.  Returns 0, or 1 on invalid input.
 * *solve_seconds (optional) = CPU time of the D4 solve only, the quantity the reference prints as
 * "REQUIRED TIME FOR SOLUTION" (APERTURE_MAP_FLOW.F:665-669). */
int apm_oracle_solve(const apm_lcl_params *p, const double *aperture, int nx, int nz,
                     apm_oracle_out *out, apm_lcl_scalars *sc, double *solve_seconds);

/* Same, evaluated in __float128 (mirrors the reference's `make test` build, -freal-8-real-16,
 * apmapflow/src/makefile:57-61, which produced the golden files). */
int apm_oracle_solve_quad(const apm_lcl_params *p, const double *aperture, int nx, int nz,
                          apm_oracle_out *out, apm_lcl_scalars *sc, double *solve_seconds);

/* MAP_STATS (MAP_MODULE.F:182-303): rms, stddev, x/z correlation lengths (cells), mean. */
void apm_oracle_map_stats(const double *map, int nx, int nz, double stats[5]);

/* Bytes of UPPER_D4 fill the D4 solve would allocate for an nx*nz interior grid with
 * Dirichlet rows on both flow sides (for sizing the CPU baseline). */
double apm_oracle_d4_fill_bytes(int nx, int nz, int outlet_side);

#ifdef __cplusplus
}
#endif
#endif
