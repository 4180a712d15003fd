/*
 * apm_oracle.c -- CPU restatement of the reference's LCL path.  TEST INFRASTRUCTURE ONLY
 * (see apm_oracle.h for who may use it and for the parity-pinning status).
 *
 * Follows, routine by routine (paths relative to /root/reference/apmapflow/src):
 *   READ_AP_MAP numeric core   APERTURE_MAP_FLOW.F:472-501
 *   CREATE_POINT_MAP           MAP_MODULE.F:91-176
 *   COEF                       APM_MODULE.F:142-271
 *   FLOW_BOUNDARY              APM_MODULE.F:277-376
 *   SOLVER assembly            APERTURE_MAP_FLOW.F:628-688
 *   ADD_DIRICHLET_BOUND_COND   D4_SOLVER_MODULE.F:170-257
 *   ADD_NUEMANNG_BOUND_COND    D4_SOLVER_MODULE.F:263-328
 *   D4_REORDER / D4_SOLVER     D4_SOLVER_MODULE.F:334-617
 *   FRAC_FLOW                  APERTURE_MAP_FLOW.F:729-762
 *   OUTPUT_DATA derived fields APERTURE_MAP_FLOW.F:825-858
 *   MAP_STATS                  MAP_MODULE.F:182-303
 * All reals are REAL(8) (makefile:16 -fdefault-real-8), or __float128 when compiled with
 * -DORACLE_QUAD (the reference's test build, makefile:57-61).  Arrays are kept 1-based like the
 * Fortran so each index expression can be checked against the cited line.
 */
#include "apm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef ORACLE_QUAD
#include <quadmath.h>
typedef __float128 real;
#define R_ATAN atanq
#define R_TAN tanq
#define R_SQRT sqrtq
#define R_FABS fabsq
#define ORACLE_ENTRY apm_oracle_solve_quad
#else
typedef double real;
#define R_ATAN atan
#define R_TAN tan
#define R_SQRT sqrt
#define R_FABS fabs
#define ORACLE_ENTRY apm_oracle_solve
#endif

/* Fortran INT(x)==0 for a REAL x (truncation toward zero; out-of-range -> INT_MIN != 0 on x86). */
static int f_int_is_zero(real x) { return R_FABS(x) < (real)1.0; }

typedef struct {
    int nx, nz;
    real *ap;    /* AP_MAP(IZ,IX)  -> ap[(IZ-1)*nx + IX-1]  (clamped, voxels) */
    real *pm;    /* AP_POINT_MAP(IZ,IX,K) -> pm[((IZ-1)*nx + IX-1)*4 + K-1]   */
    real *width; /* WIDTH(I) 1-based */
    real *trx;   /* TRX(IZ,IX), IX=0..NX -> trx[(IZ-1)*(nx+1) + IX] */
    real *trz;   /* TRZ(IZ,IX), IZ=0..NZ -> trz[IZ*nx + IX-1] */
    real *pr;    /* FRAC_PR(IZ,IX), 0..NZ+1, 0..NX+1 -> pr[IZ*(nx+2) + IX] */
    real *q;     /* Q(K,I) -> q[(K-1)*n + I-1] */
    int *inlet, *outlet, nc;
    int inlet_side;
} model_t;

#define AP(iz, ix) m->ap[(size_t)((iz)-1) * m->nx + (ix)-1]
#define PM(iz, ix, k) m->pm[((size_t)((iz)-1) * m->nx + (ix)-1) * 4 + (k)-1]
#define TRX(iz, ix) m->trx[(size_t)((iz)-1) * (m->nx + 1) + (ix)]
#define TRZ(iz, ix) m->trz[(size_t)(iz) * m->nx + (ix)-1]
#define PR(iz, ix) m->pr[(size_t)(iz) * (m->nx + 2) + (ix)]
#define Q(k, i) m->q[(size_t)((k)-1) * n + (i)-1]

/* MAP_MODULE.F:91-176 */
static void create_point_map(model_t *m)
{
    const int nx = m->nx, nz = m->nz;
    int ix, iz;
    PM(1, 1, 1) = AP(1, 1);
    PM(1, nx, 2) = AP(1, nx);
    PM(nz, nx, 3) = AP(nz, nx);
    PM(nz, 1, 4) = AP(nz, 1);
    for (ix = 1; ix <= nx - 1; ix++)
        for (iz = 1; iz <= nz - 1; iz++) {
            real v = AP(iz, ix);          /* :134 */
            v = v + AP(iz + 1, ix);       /* above  :135 */
            v = v + AP(iz, ix + 1);       /* right  :136 */
            v = v + AP(iz + 1, ix + 1);   /* corner :137 */
            v = v / (real)4.0;
            PM(iz, ix, 3) = v;
            PM(iz + 1, ix + 1, 1) = v;
            PM(iz + 1, ix, 2) = v;
            PM(iz, ix + 1, 4) = v;
        }
    for (iz = 1; iz <= nz - 1; iz++) {   /* left / right map edges :146-158 */
        real v = (AP(iz, 1) + AP(iz + 1, 1)) / (real)2.0;
        PM(iz, 1, 4) = v;
        PM(iz + 1, 1, 1) = v;
        v = (AP(iz, nx) + AP(iz + 1, nx)) / (real)2.0;
        PM(iz, nx, 3) = v;
        PM(iz + 1, nx, 2) = v;
    }
    for (ix = 1; ix <= nx - 1; ix++) {   /* bottom / top map edges :161-173 */
        real v = (AP(1, ix) + AP(1, ix + 1)) / (real)2.0;
        PM(1, ix, 2) = v;
        PM(1, ix + 1, 1) = v;
        v = (AP(nz, ix) + AP(nz, ix + 1)) / (real)2.0;
        PM(nz, ix, 3) = v;
        PM(nz, ix + 1, 4) = v;
    }
}

/* one half-cell tapered-plate term, APM_MODULE.F:200-205 (P side) / :207-211 (F side).
 * a = first factor of W3's numerator, b = second, in the source's order. */
static real tapered_b3(real a, real b, real dhalf)
{
    real theta = R_ATAN(R_FABS(a - b) / dhalf);
    real w3 = (((real)2 * (a * a)) * (b * b)) / (a + b);
    real t = R_TAN(theta);
    real x = t - theta;
    if (f_int_is_zero((real)1e9 * x)) return w3;   /* :205 */
    return ((w3 * (real)3) * x) / ((t * t) * t);   /* :204 left-to-right */
}

/* APM_MODULE.F:142-271 */
static void coef(model_t *m, const apm_lcl_params *p, real diam, real leng)
{
    const int nx = m->nx, nz = m->nz;
    const size_t n = (size_t)nx * nz;
    const real r_fact = p->roughness, low = p->low_mask, vox = p->voxel_size, visc = p->viscosity;
    const real dx = diam / nx, dz = leng / nz;   /* :173-174 (uniform) */
    real *rht = (real *)malloc((n + 1) * sizeof(real));
    real *top = (real *)malloc((n + 1) * sizeof(real));
    int ix, iz;
    size_t i;
    for (i = 0; i < (size_t)nz * (nx + 1); i++) m->trx[i] = 0;
    for (i = 0; i < (size_t)(nz + 1) * nx; i++) m->trz[i] = 0;
    for (iz = 1; iz <= nz; iz++)
        for (ix = 1; ix <= nx; ix++) {
            i = (size_t)(iz - 1) * nx + ix;
            rht[i] = PM(iz, ix, 2) + PM(iz, ix, 3);   /* :182 */
            top[i] = PM(iz, ix, 4) + PM(iz, ix, 3);   /* :183 */
        }
    for (i = 1; i <= n; i++) {
        rht[i] = (real)0.5 * rht[i] - r_fact;         /* :186 */
        top[i] = (real)0.5 * top[i] - r_fact;
        if (rht[i] < low) rht[i] = low;               /* :188 */
        if (top[i] < low) top[i] = low;
        rht[i] = rht[i] * vox;                        /* :190 */
        top[i] = top[i] * vox;
    }
    for (iz = 1; iz <= nz; iz++)
        for (ix = 1; ix <= nx - 1; ix++) {            /* :195-215 */
            real b, dbp, dbf, bp3, bf3;
            i = (size_t)(iz - 1) * nx + ix;
            b = rht[i];
            dbp = dx / 2;
            bp3 = tapered_b3(m->width[i], b, dbp);
            dbf = dx / 2;
            bf3 = tapered_b3(b, m->width[i + 1], dbf);
            TRX(iz, ix) = ((real)1 / (dbp / bp3 + dbp / bf3)) * (dz / 12);   /* :213, DBP twice */
        }
    for (iz = 1; iz <= nz - 1; iz++)
        for (ix = 1; ix <= nx; ix++) {                /* :217-237 */
            real b, dbp, dbf, bp3, bf3;
            i = (size_t)(iz - 1) * nx + ix;
            b = top[i];
            dbp = dz / 2;
            bp3 = tapered_b3(m->width[i], b, dbp);
            dbf = dz / 2;
            bf3 = tapered_b3(b, m->width[i + nx], dbf);
            TRZ(iz, ix) = ((real)1 / (dbp / bp3 + dbp / bf3)) * (dx / 12);   /* :235 */
        }
    for (iz = 1; iz <= nz; iz++) {                    /* :240-250 */
        real bp3;
        i = (size_t)(iz - 1) * nx + nx;
        bp3 = tapered_b3(m->width[i], rht[i], dx / 2);
        TRX(iz, nx) = (dz * bp3) / ((real)12.0 * dx);
    }
    for (ix = 1; ix <= nx; ix++) {                    /* :253-263 */
        real bp3;
        i = (size_t)(nz - 1) * nx + ix;
        bp3 = tapered_b3(m->width[i], top[i], dz / 2);
        TRZ(nz, ix) = (dz * bp3) / ((real)12.0 * dx); /* sic: DZ/DX, :262 */
    }
    for (i = 0; i < (size_t)nz * (nx + 1); i++) m->trx[i] = m->trx[i] / visc;   /* :266-267 */
    for (i = 0; i < (size_t)(nz + 1) * nx; i++) m->trz[i] = m->trz[i] / visc;
    free(rht);
    free(top);
}

/* APM_MODULE.F:277-376 */
static int flow_boundary(model_t *m, int side)
{
    const int nx = m->nx, nz = m->nz;
    int ix, iz;
    if (side == APM_LCL_LEFT || side == APM_LCL_RIGHT) {
        for (iz = 1; iz <= nz; iz++) {
            TRX(iz, 0) = 2 * TRX(iz, 1);
            TRX(iz, nx) = 2 * TRX(iz, nx);
        }
        for (ix = 1; ix <= nx; ix++) {
            TRZ(0, ix) = 0;
            TRZ(nz, ix) = 0;
        }
        m->nc = nz;
        m->inlet_side = (side == APM_LCL_LEFT) ? APM_LCL_RIGHT : APM_LCL_LEFT;
        for (iz = 1; iz <= nz; iz++) {
            int lo = (iz - 1) * nx + 1, hi = (iz - 1) * nx + nx;
            m->inlet[iz] = (side == APM_LCL_LEFT) ? hi : lo;
            m->outlet[iz] = (side == APM_LCL_LEFT) ? lo : hi;
        }
    } else if (side == APM_LCL_TOP || side == APM_LCL_BOTTOM) {
        for (iz = 1; iz <= nz; iz++) {
            TRX(iz, 0) = 0;
            TRX(iz, nx) = 0;
        }
        for (ix = 1; ix <= nx; ix++) {
            TRZ(0, ix) = 2 * TRZ(1, ix);
            TRZ(nz, ix) = 2 * TRZ(nz, ix);
        }
        m->nc = nx;
        m->inlet_side = (side == APM_LCL_TOP) ? APM_LCL_BOTTOM : APM_LCL_TOP;
        for (ix = 1; ix <= nx; ix++) {
            int lo = ix, hi = (nz - 1) * nx + ix;
            m->inlet[ix] = (side == APM_LCL_TOP) ? lo : hi;
            m->outlet[ix] = (side == APM_LCL_TOP) ? hi : lo;
        }
    } else
        return 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Linear system in the solver module's own (growing) grid, D4_SOLVER_MODULE.F:41-57.
 */
typedef struct {
    int onx, onz, snx, snz, xoff, zoff;
    real *cc, *rhs, *tt;      /* cc[1..], rhs[1..], TT(I,ID) -> tt[(ID-1)*(nxz+1) + I] */
    real *ngrow, *ngcol;
    real ngcc, ngrhs;
    size_t cap;
} lsys_t;

#define TT(s, i, id) (s)->tt[(size_t)((id)-1) * ((s)->cap + 1) + (i)]

static void lsys_alloc(lsys_t *s, int nx, int nz)
{
    s->cap = (size_t)nx * nz;
    s->cc = (real *)calloc(s->cap + 1, sizeof(real));
    s->rhs = (real *)calloc(s->cap + 1, sizeof(real));
    s->tt = (real *)calloc(4 * (s->cap + 1), sizeof(real));
    s->ngrow = (real *)calloc(s->cap + 1, sizeof(real));
    s->ngcol = (real *)calloc(s->cap + 1, sizeof(real));
}
static void lsys_free(lsys_t *s)
{
    free(s->cc); free(s->rhs); free(s->tt); free(s->ngrow); free(s->ngcol);
}

/* D4_SOLVER_MODULE.F:170-257 */
static void add_dirichlet(lsys_t *s, int side, real value)
{
    const int onx = s->snx, onz = s->snz;
    lsys_t t = *s;
    int nbc, ib, iz, id;
    int *bc, *mv;
    mv = (int *)malloc(((size_t)onx * onz + 1) * sizeof(int));
    if (side == APM_LCL_BOTTOM) {
        t.snz = onz + 1; t.zoff = s->zoff + 1; nbc = onx;
        bc = (int *)malloc((nbc + 1) * sizeof(int));
        for (ib = 1; ib <= onx; ib++) bc[ib] = ib;
        for (ib = 1; ib <= onx * onz; ib++) mv[ib] = onx + ib;
    } else if (side == APM_LCL_TOP) {
        t.snz = onz + 1; nbc = onx;
        bc = (int *)malloc((nbc + 1) * sizeof(int));
        for (ib = 1; ib <= onx; ib++) bc[ib] = onx * onz + ib;
        for (ib = 1; ib <= onx * onz; ib++) mv[ib] = ib;
    } else if (side == APM_LCL_LEFT) {
        t.snx = onx + 1; t.xoff = s->xoff + 1; nbc = onz;
        bc = (int *)malloc((nbc + 1) * sizeof(int));
        for (iz = 0; iz <= onz - 1; iz++) {
            bc[iz + 1] = iz * t.snx + 1;
            for (ib = 1; ib <= onx; ib++) mv[iz * onx + ib] = iz * t.snx + ib + 1;
        }
    } else { /* RIGHT */
        t.snx = onx + 1; nbc = onz;
        bc = (int *)malloc((nbc + 1) * sizeof(int));
        for (iz = 0; iz <= onz - 1; iz++) {
            bc[iz + 1] = iz * t.snx + t.snx;
            for (ib = 1; ib <= onx; ib++) mv[iz * onx + ib] = iz * t.snx + ib;
        }
    }
    lsys_alloc(&t, t.snx, t.snz);   /* zero filled: NEW_CC, NEW_TT, NEW_RHS = 0 (:236-238) */
    for (ib = 1; ib <= nbc; ib++) {
        t.cc[bc[ib]] = 1;            /* :241 */
        t.rhs[bc[ib]] = value;       /* :245 */
    }
    for (ib = 1; ib <= onx * onz; ib++) {
        t.cc[mv[ib]] = s->cc[ib];
        t.rhs[mv[ib]] = s->rhs[ib];
        for (id = 1; id <= 4; id++) TT(&t, mv[ib], id) = TT(s, ib, id);
    }
    /* ALLOCATE_COEF gives fresh NEUMANNG_ROW/COL; they are (re)zeroed by ADD_NUEMANNG anyway */
    lsys_free(s);
    *s = t;
    free(bc);
    free(mv);
}

/* D4_SOLVER_MODULE.F:263-328 */
static void add_neumann_group(lsys_t *s, int side, real value)
{
    const int nxz = s->snx * s->snz;
    int n, ib, i, id;
    int *bc;
    if (side == APM_LCL_BOTTOM) {
        n = s->snx; id = 3; bc = (int *)malloc((n + 1) * sizeof(int));
        for (ib = 1; ib <= n; ib++) bc[ib] = ib;
    } else if (side == APM_LCL_TOP) {
        n = s->snx; id = 4; bc = (int *)malloc((n + 1) * sizeof(int));
        for (ib = 1; ib <= n; ib++) bc[ib] = (s->snz - 1) * s->snx + ib;
    } else if (side == APM_LCL_LEFT) {
        n = s->snz; id = 1; bc = (int *)malloc((n + 1) * sizeof(int));
        for (ib = 0; ib <= n - 1; ib++) bc[ib + 1] = ib * s->snx + 1;
    } else {
        n = s->snz; id = 2; bc = (int *)malloc((n + 1) * sizeof(int));
        for (ib = 0; ib <= n - 1; ib++) bc[ib + 1] = ib * s->snx + s->snx;
    }
    for (i = 1; i <= nxz; i++) { s->ngrow[i] = 0; s->ngcol[i] = 0; }
    for (i = 1; i <= n; i++) {
        ib = bc[i];
        s->ngcol[ib] = TT(s, ib, id);
        s->ngrow[ib] = TT(s, ib, id);
        TT(s, ib, id) = 0;
    }
    {
        real sum = 0;
        for (i = 1; i <= nxz; i++) sum = sum + s->ngrow[i];
        s->ngcc = -sum;               /* :322 */
    }
    s->ngrhs = value;
    free(bc);
}

/* D4_SOLVER_MODULE.F:334-438 + :444-617.  x (size onx*onz, 1-based) and the group unknown. */
static void d4_solve(lsys_t *s, real *x_out, real *ng_x, double *fill_bytes_only)
{
    const int snx = s->snx, snz = s->snz;
    const int nxz = snx * snz;
    int *ino = (int *)calloc((size_t)nxz + 1, sizeof(int));
    int *inat = (int *)calloc((size_t)nxz + 1, sizeof(int));
    int *link = (int *)calloc(4 * ((size_t)nxz + 1), sizeof(int));
#define LINK(ibr, id) link[(size_t)((id)-1) * (nxz + 1) + (ibr)]
    long long *ibase;
    int ibr = 0, ipass, idiag, i, j, ib, id, jd, num_top, num_total, max_band = 0;
    int ibr_max[3];
    /* reorder: alternating anti-diagonals, :359-388 */
    for (ipass = 1; ipass <= 2; ipass++) {
        if (snx >= snz) {
            for (idiag = ipass + 1; idiag <= snx + snz; idiag += 2)
                for (i = 1; i <= snx; i++) {
                    j = idiag - i;
                    if (j > snz) continue;
                    if (j < 1) break;
                    ib = snx * (j - 1) + i;
                    ino[ib] = ++ibr;
                }
        } else {
            for (idiag = ipass + 1; idiag <= snx + snz; idiag += 2)
                for (j = 1; j <= snz; j++) {
                    i = idiag - j;
                    if (i > snx) continue;
                    if (i < 1) break;
                    ib = snx * (j - 1) + i;
                    ino[ib] = ++ibr;
                }
        }
        ibr_max[ipass] = ibr;
    }
    num_top = ibr_max[1];
    num_total = ibr_max[2];
    /* neighbour pointers, :392-408 */
    ib = 0;
    for (j = 1; j <= snz; j++)
        for (i = 1; i <= snx; i++) {
            ib++;
            ibr = ino[ib];
            if (ibr > 0) {
                inat[ibr] = ib;
                if (i > 1) LINK(ibr, 1) = ino[ib - 1];
                if (i < snx) LINK(ibr, 2) = ino[ib + 1];
                if (j > 1) LINK(ibr, 3) = ino[ib - snx];
                if (j < snz) LINK(ibr, 4) = ino[ib + snx];
            }
        }
    /* profile storage, :412-433 */
    ibase = (long long *)calloc((size_t)num_total + 2, sizeof(long long));
    {
        int ir = 0, il;
        ibase[num_top + 1] = 0;
        for (ibr = num_top + 1; ibr <= num_total; ibr++) {
            il = ibr;
            for (id = 1; id <= 4; id++) {
                int ibru = LINK(ibr, id);
                if (ibru > 0)
                    for (jd = 1; jd <= 4; jd++) {
                        int jbr = LINK(ibru, jd);
                        if (jbr > 0) {
                            if (jbr < il) il = jbr;
                            if (jbr > ir) ir = jbr;
                        }
                    }
            }
            if (ibr - il > max_band) max_band = ibr - il;
            if (ir - ibr > max_band) max_band = ir - ibr;
            ibase[ibr + 1] = ibase[ibr] + ir - ibr;
        }
    }
    if (fill_bytes_only) {
        *fill_bytes_only = (double)ibase[num_total + 1] * sizeof(real);
        free(ino); free(inat); free(link); free(ibase);
        return;
    }
    {
        real *upper = (real *)malloc(((size_t)ibase[num_total + 1] + 1) * sizeof(real));
        real *ngd4col = (real *)calloc((size_t)nxz + 2, sizeof(real));
        real *ngd4row = (real *)calloc((size_t)nxz + 2, sizeof(real));
        real *x = (real *)calloc((size_t)nxz + 2, sizeof(real));
        real *rowbuf = (real *)calloc(2 * (size_t)max_band + 3, sizeof(real));
        real *row = rowbuf + max_band + 1;   /* ROW(-MAX_BAND:MAX_BAND) */
        real factor, temp_rhs, temp_ngcol;
        int icol, jcol, il, jbr, ibru, ibu, ix, iz;
        long long iup;
        /* lower-half elimination, :486-539 */
        for (ibr = num_top + 1; ibr <= num_total; ibr++) {
            ib = inat[ibr];
            temp_rhs = s->rhs[ib];
            temp_ngcol = s->ngcol[ib];
            for (icol = -max_band; icol <= max_band; icol++) row[icol] = 0;
            row[0] = s->cc[ib];
            il = 0;
            for (id = 1; id <= 4; id++) {
                ibru = LINK(ibr, id);
                if (ibru > 0) {
                    ibu = inat[ibru];
                    factor = TT(s, ib, id) / s->cc[ibu];
                    for (jd = 1; jd <= 4; jd++) {
                        jbr = LINK(ibru, jd);
                        if (jbr > 0) {
                            icol = jbr - ibr;
                            if (icol < il) il = icol;
                            row[icol] = row[icol] - factor * TT(s, ibu, jd);
                        }
                    }
                    temp_ngcol = temp_ngcol - factor * s->ngcol[ibu];
                    temp_rhs = temp_rhs - factor * s->rhs[ibu];
                }
            }
            for (icol = il; icol <= -1; icol++) {   /* :517-527 hot loop */
                factor = row[icol];
                jbr = ibr + icol;
                jcol = icol;
                for (iup = ibase[jbr] + 1; iup <= ibase[jbr + 1]; iup++) {
                    jcol++;
                    row[jcol] = row[jcol] - factor * upper[iup];
                }
                temp_ngcol = temp_ngcol - factor * ngd4col[jbr];
                temp_rhs = temp_rhs - factor * x[inat[jbr]];
            }
            factor = (real)1.0 / row[0];
            icol = 0;
            for (iup = ibase[ibr] + 1; iup <= ibase[ibr + 1]; iup++) {
                icol++;
                upper[iup] = factor * row[icol];
            }
            ngd4col[ibr] = factor * temp_ngcol;
            x[ib] = factor * temp_rhs;
        }
        /* group row elimination, :543-569 */
        for (ib = 1; ib <= nxz; ib++) ngd4row[ino[ib]] = s->ngrow[ib];
        for (ibr = 1; ibr <= num_top; ibr++) {
            ib = inat[ibr];
            factor = ngd4row[ibr] / s->cc[ib];
            for (jd = 1; jd <= 4; jd++) {
                jbr = LINK(ibr, jd);
                if (jbr > 0) ngd4row[jbr] = ngd4row[jbr] - factor * TT(s, ib, jd);
            }
            s->ngcc = s->ngcc - factor * s->ngcol[ib];
            s->ngrhs = s->ngrhs - factor * s->rhs[ib];
        }
        for (ibr = num_top + 1; ibr <= num_total; ibr++) {
            factor = ngd4row[ibr];
            jcol = ibr;
            for (iup = ibase[ibr] + 1; iup <= ibase[ibr + 1]; iup++) {
                jcol++;
                ngd4row[jcol] = ngd4row[jcol] - factor * upper[iup];
            }
            s->ngcc = s->ngcc - factor * ngd4col[ibr];
            s->ngrhs = s->ngrhs - factor * x[inat[ibr]];
        }
        *ng_x = s->ngrhs / s->ngcc;
        /* back substitution, :573-597 */
        x[inat[nxz]] = x[inat[nxz]] - ngd4col[nxz] * (*ng_x);
        for (ibr = num_total - 1; ibr >= num_top + 1; ibr--) {
            ib = inat[ibr];
            temp_rhs = x[ib];
            temp_rhs = temp_rhs - ngd4col[ibr] * (*ng_x);
            jbr = ibr;
            for (iup = ibase[ibr] + 1; iup <= ibase[ibr + 1]; iup++) {
                jbr++;
                temp_rhs = temp_rhs - upper[iup] * x[inat[jbr]];
            }
            x[ib] = temp_rhs;
        }
        for (ibr = 1; ibr <= num_top; ibr++) {
            ib = inat[ibr];
            temp_rhs = s->rhs[ib];
            temp_rhs = temp_rhs - s->ngcol[ib] * (*ng_x);
            for (id = 1; id <= 4; id++) {
                jbr = LINK(ibr, id);
                if (jbr > 0) temp_rhs = temp_rhs - TT(s, ib, id) * x[inat[jbr]];
            }
            x[ib] = temp_rhs / s->cc[ib];
        }
        /* strip the BC padding, :600-609 */
        for (iz = 1; iz <= s->onz; iz++)
            for (ix = 1; ix <= s->onx; ix++) {
                ibu = (iz + s->zoff - 1) * snx + ix + s->xoff;
                ib = (iz - 1) * s->onx + ix;
                x_out[ib] = x[ibu];
            }
        free(upper); free(ngd4col); free(ngd4row); free(x); free(rowbuf);
    }
    free(ino); free(inat); free(link); free(ibase);
#undef LINK
}

static double cpu_seconds(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int ORACLE_ENTRY(const apm_lcl_params *p, const double *aperture, int nx, int nz,
                 apm_oracle_out *out, apm_lcl_scalars *sc, double *solve_seconds)
{
    model_t mm, *m = &mm;
    const size_t n = (size_t)nx * nz;
    const int side = p->outlet_side;
    const real high = p->high_mask, low = p->low_mask, vox = p->voxel_size;
    real inlpb = p->inlet_pressure, outpb = p->outlet_pressure;
    const real injrate = p->inlet_rate, outrate_in = p->outlet_rate;
    real raw = 0, avg = 0, minap = 0, maxap = 0, diam, leng, flow = 0, ng_x = 0;
    long long zeros = 0, masked = 0;
    int ix, iz, have_min = 0, presc;
    size_t i;
    lsys_t ls;
    real *x;

    if (nx < 1 || nz < 1 || side < 0 || side > 3 || !(p->viscosity > 0)) return 1;
    memset(m, 0, sizeof(*m));
    m->nx = nx; m->nz = nz;
    m->ap = (real *)malloc(n * sizeof(real));
    m->pm = (real *)calloc(n * 4, sizeof(real));
    m->width = (real *)malloc((n + 1) * sizeof(real));
    m->trx = (real *)malloc((size_t)nz * (nx + 1) * sizeof(real));
    m->trz = (real *)malloc((size_t)(nz + 1) * nx * sizeof(real));
    m->pr = (real *)calloc((size_t)(nz + 2) * (nx + 2), sizeof(real));
    m->q = (real *)calloc(5 * n, sizeof(real));
    m->inlet = (int *)malloc(((size_t)(nx > nz ? nx : nz) + 1) * sizeof(int));
    m->outlet = (int *)malloc(((size_t)(nx > nz ? nx : nz) + 1) * sizeof(int));
    for (i = 0; i < n; i++) m->ap[i] = aperture[i];

    /* APERTURE_MAP_FLOW.F:462-463 */
    diam = nx * vox * (real)p->avg_factor;
    leng = nz * vox * (real)p->avg_factor;

    /* :472-480 -- Fortran array intrinsics walk column-major (IZ fastest) */
    for (ix = 1; ix <= nx; ix++)
        for (iz = 1; iz <= nz; iz++) {
            real a = AP(iz, ix);
            raw = raw + a;
            if (f_int_is_zero(a * (real)1e9)) zeros++;
            if (!(ix == 1 && iz == 1)) { if (a > maxap) maxap = a; } else maxap = a;
            if (a > 0) { if (!have_min || a < minap) { minap = a; have_min = 1; } }
            if (a > high) masked++;
        }
    for (ix = 1; ix <= nx; ix++)
        for (iz = 1; iz <= nz; iz++)
            if (AP(iz, ix) < low) masked++;
    raw = raw / (real)(nx * (long long)nz);
    for (i = 0; i < n; i++) if (m->ap[i] > high) m->ap[i] = high;   /* :478 */
    for (i = 0; i < n; i++) if (m->ap[i] < low) m->ap[i] = low;     /* :479 */
    for (ix = 1; ix <= nx; ix++)
        for (iz = 1; iz <= nz; iz++) avg = avg + AP(iz, ix);
    avg = avg / (real)(nz * (long long)nx) * vox;                   /* :480 */

    create_point_map(m);                                            /* :483 */
    for (iz = 1; iz <= nz; iz++)
        for (ix = 1; ix <= nx; ix++) {
            i = (size_t)(iz - 1) * nx + ix;
            m->width[i] = AP(iz, ix) - (real)p->roughness;          /* :495 */
        }
    for (i = 1; i <= n; i++) if (m->width[i] < low) m->width[i] = low;   /* :498 */
    for (i = 1; i <= n; i++) m->width[i] = m->width[i] * vox;            /* :500 */

    coef(m, p, diam, leng);
    if (flow_boundary(m, side)) return 1;

    /* SOLVER, APERTURE_MAP_FLOW.F:628-659 */
    memset(&ls, 0, sizeof(ls));
    ls.onx = ls.snx = nx; ls.onz = ls.snz = nz;
    lsys_alloc(&ls, nx, nz);
    ls.ngcc = 1; ls.ngrhs = 0;
    for (iz = 1; iz <= nz; iz++)
        for (ix = 1; ix <= nx; ix++) {
            real sum;
            i = (size_t)(iz - 1) * nx + ix;
            TT(&ls, i, 1) = -TRX(iz, ix - 1);
            TT(&ls, i, 2) = -TRX(iz, ix);
            TT(&ls, i, 3) = -TRZ(iz - 1, ix);
            TT(&ls, i, 4) = -TRZ(iz, ix);
            sum = TT(&ls, i, 1);
            sum = sum + TT(&ls, i, 2);
            sum = sum + TT(&ls, i, 3);
            sum = sum + TT(&ls, i, 4);
            ls.cc[i] = -sum;
        }
    if (outpb >= 0) add_dirichlet(&ls, side, outpb);              /* :644-646 */
    if (inlpb >= 0) add_dirichlet(&ls, m->inlet_side, inlpb);     /* :648-650 */
    if (injrate > 0) add_neumann_group(&ls, m->inlet_side, injrate);   /* :653-655 */
    if (outrate_in > 0) add_neumann_group(&ls, side, -outrate_in);     /* :657-659 */

    x = (real *)calloc(n + 1, sizeof(real));
    {
        double t0 = cpu_seconds();
        d4_solve(&ls, x, &ng_x, NULL);
        if (solve_seconds) *solve_seconds = cpu_seconds() - t0;
    }
    for (iz = 1; iz <= nz; iz++)
        for (ix = 1; ix <= nx; ix++) PR(iz, ix) = x[(size_t)(iz - 1) * nx + ix];   /* :672-677 */
    presc = (p->inlet_pressure >= 0) && (p->outlet_pressure >= 0);   /* :137-143 */
    if (!presc) {                                                    /* :680-688 */
        if (injrate > 0) inlpb = ng_x;
        if (outrate_in > 0) outpb = ng_x;
    }
    free(x);
    lsys_free(&ls);

    /* FRAC_FLOW halo pressures :729-749 */
    {
        real left = inlpb, right = inlpb, bot = inlpb, topv = inlpb;
        if (side == APM_LCL_LEFT) left = outpb;
        else if (side == APM_LCL_RIGHT) right = outpb;
        else if (side == APM_LCL_TOP) topv = outpb;
        else bot = outpb;
        /* assignment order in the source: columns 0 and NX+1 first (all rows), then rows 0, NZ+1
           (all columns), so the four corners take the row values */
        for (iz = 0; iz <= nz + 1; iz++) { PR(iz, 0) = left; PR(iz, nx + 1) = right; }
        for (ix = 0; ix <= nx + 1; ix++) { PR(0, ix) = bot; PR(nz + 1, ix) = topv; }
    }
    for (iz = 1; iz <= nz; iz++)
        for (ix = 1; ix <= nx; ix++) {                               /* :753-762 */
            i = (size_t)(iz - 1) * nx + ix;
            Q(1, i) = -TRX(iz, ix - 1) * (PR(iz, ix) - PR(iz, ix - 1));
            Q(2, i) = -TRX(iz, ix) * (PR(iz, ix + 1) - PR(iz, ix));
            Q(3, i) = -TRZ(iz - 1, ix) * (PR(iz, ix) - PR(iz - 1, ix));
            Q(4, i) = -TRZ(iz, ix) * (PR(iz + 1, ix) - PR(iz, ix));
            Q(5, i) = ((Q(1, i) - Q(2, i)) + (Q(3, i) - Q(4, i)));
        }

    /* OUTPUT_DATA :825-858 */
    {
        real maxres = 0, maxres_signed = 0;
        int ic;
        for (i = 1; i <= n; i++) {
            real r = R_FABS(Q(5, i));
            if (i == 1 || r > maxres) { maxres = r; maxres_signed = Q(5, i); }   /* MAXLOC = first max */
        }
        for (ic = 1; ic <= m->nc; ic++) {
            size_t oc = m->outlet[ic];
            if (side == APM_LCL_LEFT) flow = flow + Q(1, oc);
            if (side == APM_LCL_RIGHT) flow = flow + Q(2, oc);
            if (side == APM_LCL_TOP) flow = flow + Q(3, oc);
            if (side == APM_LCL_BOTTOM) flow = flow + Q(4, oc);
        }
        for (i = 1; i <= n; i++) {
            real qx = (side == APM_LCL_LEFT) ? Q(1, i) : Q(2, i);
            real qz = (side == APM_LCL_BOTTOM) ? Q(3, i) : Q(4, i);
            real qm = R_SQRT(qx * qx + qz * qz);
            if (out) {
                if (out->flow_x) out->flow_x[i - 1] = (double)qx;
                if (out->flow_z) out->flow_z[i - 1] = (double)qz;
                if (out->flow_m) out->flow_m[i - 1] = (double)qm;
                if (out->reynolds)
                    out->reynolds[i - 1] = (double)(((real)p->density / (real)p->viscosity) * (qm / m->width[i]));
                if (out->residual) out->residual[i - 1] = (double)(R_FABS(Q(5, i)) / maxres);
                if (out->width) out->width[i - 1] = (double)m->width[i];
            }
        }
        if (sc) {
            memset(sc, 0, sizeof(*sc));
            sc->zeros = zeros; sc->masked = masked;
            sc->raw_aper = (double)raw; sc->avg_aper = (double)avg;
            sc->min_aper = (double)minap; sc->max_aper = (double)maxap;
            sc->inlet_pressure = (double)inlpb; sc->outlet_pressure = (double)outpb;
            sc->flow = (double)flow; sc->outlet_rate = (double)R_FABS(flow);
            sc->max_residual = (double)maxres_signed;
            sc->converged = 1;
        }
    }
    if (out) {
        if (out->pressure)
            for (iz = 1; iz <= nz; iz++)
                for (ix = 1; ix <= nx; ix++) out->pressure[(size_t)(iz - 1) * nx + ix - 1] = (double)PR(iz, ix);
        if (out->trx) for (i = 0; i < (size_t)nz * (nx + 1); i++) out->trx[i] = (double)m->trx[i];
        if (out->trz) for (i = 0; i < (size_t)(nz + 1) * nx; i++) out->trz[i] = (double)m->trz[i];
        if (out->point_map) for (i = 0; i < 4 * n; i++) out->point_map[i] = (double)m->pm[i];
        if (out->q) for (i = 0; i < 5 * n; i++) out->q[i] = (double)m->q[i];
    }
    free(m->ap); free(m->pm); free(m->width); free(m->trx); free(m->trz); free(m->pr); free(m->q);
    free(m->inlet); free(m->outlet);
    return 0;
}

#ifndef ORACLE_QUAD
/* MAP_MODULE.F:182-303.  map is row-major [nz][nx]; Fortran loops are IX outer / IZ inner. */
void apm_oracle_map_stats(const double *map, int nx, int nz, double stats[5])
{
#define DM(iz, ix) map[(size_t)((iz)-1) * nx + (ix)-1]
    double max_len = (double)(nx < nz ? nx : nz) / 2.0;
    double avg = 0, rms = 0, sm = 0, sms = 0, var, dev, corr[2];
    int ix, iz, k, dir;
    for (ix = 1; ix <= nx; ix++)
        for (iz = 1; iz <= nz; iz++) avg += DM(iz, ix);
    avg = avg / ((double)nx * nz);
    for (iz = 1; iz <= nz; iz++)
        for (ix = 1; ix <= nx; ix++) {
            rms = rms + DM(iz, ix) * DM(iz, ix);
            sm = sm + (DM(iz, ix) - avg);
            sms = sms + (DM(iz, ix) - avg) * (DM(iz, ix) - avg);
        }
    sm = (sm * sm) * (1.0 / ((double)nx * nz));
    rms = rms * (1.0 / ((double)nx * nz));
    var = (sms - sm) * (1.0 / ((double)nx * nz - 1));
    rms = sqrt(rms);
    dev = sqrt(var);
    for (dir = 0; dir < 2; dir++) {
        double test1 = 0, test2 = 0, test3 = 0, c = 0;
        for (k = 1; k <= (int)max_len + 1; k++) {
            double cnt = 0, ss = 0, v3;
            for (ix = 1; ix <= nx; ix++)
                for (iz = 1; iz <= nz; iz++) {
                    double d;
                    if (dir == 0) { if (ix + k > nx) continue; d = DM(iz, ix) - DM(iz, ix + k); }
                    else { if (iz + k > nz) continue; d = DM(iz, ix) - DM(iz + k, ix); }
                    cnt += 1;
                    ss += d * d;
                }
            if (cnt == 0) continue;
            v3 = ss / cnt;
            c = (double)k;
            test1 = fabs(v3 / var - 0.90);
            if (v3 / var > 0.90) {
                if (test2 < test1) c = test3;
                break;
            }
            test2 = test1;
            test3 = c;
        }
        corr[dir] = c;
    }
    stats[0] = rms; stats[1] = dev; stats[2] = corr[0]; stats[3] = corr[1]; stats[4] = avg;
#undef DM
}

double apm_oracle_d4_fill_bytes(int nx, int nz, int outlet_side)
{
    lsys_t ls;
    double bytes = 0;
    real ng;
    memset(&ls, 0, sizeof(ls));
    ls.onx = nx; ls.onz = nz;
    if (outlet_side == APM_LCL_TOP || outlet_side == APM_LCL_BOTTOM) { ls.snx = nx; ls.snz = nz + 2; }
    else { ls.snx = nx + 2; ls.snz = nz; }
    d4_solve(&ls, NULL, &ng, &bytes);
    return bytes;
}
#endif
