/*
 * apm_lcl.h -- C ABI of libapmlcl.so, the B200-native local-cubic-law (LCL) flow solve.
 *
 * Drop-in boundary for the hot path of stadelmanma/netl-ap-map-flow.  The reference crosses a
 * PROCESS boundary here: apmapflow/run_model/run_model.py:752-759 Popen()s
 * `apm-lcl-model.exe <file.inp>` (apmapflow/src/APERTURE_MAP_FLOW.F:1-51).  This library is what a
 * ctypes binding replaces that Popen with.  No torch / C++ types cross the boundary: plain
 * pointers, sizes and PODs.  Every pointer that carries map data may be a HOST or a DEVICE pointer
 * (row-major [nz][nx] doubles, row 0 = first line of the map file = bottom of the fracture,
 * APERTURE_MAP_FLOW.F:448-449, cell index I=(IZ-1)*NX+IX :491-497); the library inspects it with
 * cudaPointerGetAttributes and copies only when needed.
 *
 * All functions return 0 on success and 1 on failure (the reference's exit status,
 * APERTURE_MAP_FLOW.F:47-49); apm_lcl_last_error() returns the message (thread local).
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef APM_LCL_H
#define APM_LCL_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APM_LCL_VERSION 100

#if defined(__GNUC__)
#define APM_LCL_API __attribute__((visibility("default")))
#else
#define APM_LCL_API
#endif

/* OUTLET-SIDE, matched by substring in this order (APM_MODULE.F:306-368) */
enum { APM_LCL_LEFT = 0, APM_LCL_RIGHT = 1, APM_LCL_TOP = 2, APM_LCL_BOTTOM = 3 };

/* Model parameters in SI / voxel units exactly as READ_MASTER leaves them
 * (APERTURE_MAP_FLOW.F:88-126 defaults, :226-384 keywords). */
typedef struct apm_lcl_params {
    double inlet_pressure;   /* INLPB  [Pa]; < 0 = not supplied (default -1)            */
    double outlet_pressure;  /* OUTPB  [Pa]; < 0 = not supplied                         */
    double inlet_rate;       /* INJRATE [m^3/s]; <= 0 = not supplied                    */
    double outlet_rate;      /* OUTRATE [m^3/s]; <= 0 = not supplied                    */
    double density;          /* RHO [kg/m^3]                                            */
    double viscosity;        /* AVG_VISC [Pa s], must be > 0                            */
    double voxel_size;       /* VOX_MET [m per voxel]                                   */
    double avg_factor;       /* AVG_FACT, MAP AVERAGING FACTOR                          */
    double roughness;        /* R_FACT [voxels]                                         */
    double high_mask;        /* HIGH_MASK [voxels]                                      */
    double low_mask;         /* LOW_MASK [voxels]                                       */
    int    outlet_side;      /* APM_LCL_LEFT/RIGHT/TOP/BOTTOM                           */
    /* solver controls (new; the reference's direct D4 solve has none) */
    double rel_tol;          /* stop when ||r||2/||b||2 AND ||D^-1/2 r||/||D^-1/2 b|| <= rel_tol; <=0 -> 1e-10 */
    int    max_iter;         /* <=0 -> 500                                              */
    int    flags;            /* APM_LCL_FLAG_* bits, 0 by default                       */
} apm_lcl_params;

/* apm_lcl_bulk_solve: never derive this case from another case's unit solution (see there) */
#define APM_LCL_FLAG_OWN_SOLVE 1

/* Per-cell outputs, SI units, each nz*nx doubles or NULL to skip (host or device pointers).
 * APERTURE_MAP_FLOW.F:825-869 before the output-unit conversion. */
typedef struct apm_lcl_fields {
    double *pressure;   /* PRES  = FRAC_PR(1:NZ,1:NX)                                    */
    double *flow_x;     /* QX_OUT = Q2 (Q1 if outlet LEFT)                               */
    double *flow_z;     /* QZ_OUT = Q4 (Q3 if outlet BOTTOM)                             */
    double *flow_m;     /* QM_OUT = sqrt(QX^2+QZ^2)                                      */
    double *aperture;   /* WIDTH (roughness-adjusted, clamped, metres)                   */
    double *reynolds;   /* REN = rho/mu * QM / WIDTH                                     */
    double *residual;   /* RES = |Q5| / max|Q5|                                          */
} apm_lcl_fields;

typedef struct apm_lcl_scalars {
    /* READ_AP_MAP counters (bit exact) and averages, APERTURE_MAP_FLOW.F:472-480 */
    long long zeros;         /* COUNT(INT(A*1E9)==0)                                    */
    long long masked;        /* COUNT(A>HIGH)+COUNT(A<LOW)                              */
    double raw_aper;         /* mean aperture before masking [voxels]                   */
    double avg_aper;         /* mean aperture after masking [m]                         */
    double min_aper;         /* MINVAL(A, A>0) [voxels]                                 */
    double max_aper;         /* MAXVAL(A) [voxels]                                      */
    /* results */
    double inlet_pressure;   /* INLPB after the solve (rate BC back-fill :680-688) [Pa]  */
    double outlet_pressure;  /* OUTPB after the solve [Pa]                               */
    double flow;             /* signed FLOW = sum over outlet cells (:850-857) [m^3/s]   */
    double outlet_rate;      /* OUTRATE = |FLOW|                                         */
    double max_residual;     /* Q5 at MAXLOC(|Q5|) (signed, :879)                        */
    /* solver diagnostics */
    int    iterations;       /* PCG iterations (sum over solves)                         */
    int    converged;        /* 1 if both stopping tests passed                          */
    double rel_residual;     /* final ||r||2/||b||2 (true residual)                      */
    double rel_residual_scaled; /* final ||D^-1/2 r||2/||D^-1/2 b||2                     */
    double setup_ms;         /* conductances + multigrid hierarchy (device time)         */
    double solve_ms;         /* PCG (device time)                                        */
    double total_ms;         /* aperture-on-device -> outputs-on-device                  */
} apm_lcl_scalars;

typedef struct apm_lcl_solver apm_lcl_solver;   /* opaque, owns all device scratch */

APM_LCL_API int         apm_lcl_version(void);
APM_LCL_API const char *apm_lcl_last_error(void);
/* number of kernels launched by this thread's last compute call (bench.py's gpu_launches) */
APM_LCL_API long long   apm_lcl_last_launch_count(void);

/* Defaults of INITIALIZE_RUN (APERTURE_MAP_FLOW.F:88-126). */
APM_LCL_API void apm_lcl_default_params(apm_lcl_params *p);

/* Persistent solver for maps of one size on one device.  Replaces ALLOCATE_FRAC
 * (APM_MODULE.F:107-123) + INIT_SOLVER (D4_SOLVER_MODULE.F:110-139). */
APM_LCL_API int apm_lcl_solver_create(apm_lcl_solver **out, int device, int nx, int nz);
APM_LCL_API int apm_lcl_solver_destroy(apm_lcl_solver *s);

/* The hot path: READ_AP_MAP numeric core -> CREATE_POINT_MAP -> COEF -> FLOW_BOUNDARY ->
 * SOLVER -> FRAC_FLOW -> OUTPUT_DATA derived fields (APERTURE_MAP_FLOW.F:472-501, 597-693,
 * 700-767, 815-869; APM_MODULE.F:142-376; MAP_MODULE.F:91-176).  The direct D4 solve
 * (D4_SOLVER_MODULE.F:444-617) is replaced by multigrid-preconditioned CG on the same system. */
APM_LCL_API int apm_lcl_solver_solve(apm_lcl_solver *s, const apm_lcl_params *p, const double *aperture,
                         apm_lcl_fields *out, apm_lcl_scalars *sc);

/* One map on several GPUs (new -- the reference is one serial process): the map is cut into row
 * strips, one per rank / process / GPU of ONE node.  Rank k creates its solver with
 * apm_lcl_solver_create_striped, which fills `handle` (apm_lcl_peer_handle_bytes() bytes: the CUDA IPC handle
 * of the rank's device slab); the host gathers the handles of all ranks in rank order (torch.distributed
 * all_gather, MPI, a file -- the library has no rendezvous of its own) and passes the `world` handles to
 * apm_lcl_solver_connect.  From then on apm_lcl_solver_solve is collective: every rank calls it with the same
 * parameters, passes the aperture rows [a_lo, a_hi) of the map (its own rows plus one either side, see
 * apm_lcl_solver_rows) and receives the rows [z_lo, z_hi) of every output field; the scalars are the whole
 * map's and identical on every rank.  Ghost rows and dot products cross the GPUs inside the kernels (peer stores over
 * NVLink); no NCCL call is on the data path.  Before destroying the solvers every rank calls
 * apm_lcl_solver_disconnect, and the host puts a barrier between that and apm_lcl_solver_destroy (a slab must
 * not be freed while a peer still maps it).  Maps too small to be worth striping are solved whole by every
 * rank (then z_lo = 0, z_hi = nz on all of them).  world <= 16. */
APM_LCL_API int apm_lcl_peer_handle_bytes(void);
APM_LCL_API int apm_lcl_solver_create_striped(apm_lcl_solver **out, int device, int nx, int nz, int rank, int world,
                                              void *handle);
APM_LCL_API int apm_lcl_solver_connect(apm_lcl_solver *s, const void *handles);
APM_LCL_API int apm_lcl_solver_disconnect(apm_lcl_solver *s);
APM_LCL_API int apm_lcl_solver_rows(apm_lcl_solver *s, int *z_lo, int *z_hi, int *a_lo, int *a_hi);

/* The same from ONE process: `n_devices` strips, one solver and one host thread per entry of `devices` (a GPU
 * may be named more than once), the strips' memory addressed directly.  Whole-map HOST buffers in and out. */
APM_LCL_API int apm_lcl_solve_striped(const apm_lcl_params *p, const double *aperture, int nx, int nz, int n_devices,
                                      const int *devices, apm_lcl_fields *out, apm_lcl_scalars *sc);

/* One-shot convenience (create + solve + destroy) on `device`. */
APM_LCL_API int apm_lcl_solve(const apm_lcl_params *p, const double *aperture, int nx, int nz, int device,
                  apm_lcl_fields *out, apm_lcl_scalars *sc);

/* Many independent cases of one map size on one device (the tensor-level counterpart of
 * BulkRun, apmapflow/run_model/bulk_run.py:107-128: independent simulations, no exchange between
 * them).  Case k uses params[k], apertures[k] and writes outs[k] / scalars[k]; one solver object is
 * reused, so allocations, the CUDA graph and the multigrid storage are paid once.  Cases that pass the SAME
 * aperture pointer and differ only in inlet_pressure / outlet_pressure (both supplied, no rates: a sweep over
 * OUTLET-PRESS) share one unit solution, p = OUTPB + (INLPB - OUTPB) u: the first is solved, the others are an
 * affine map of its pressure field plus the flux pass and report iterations = 0.  APM_LCL_BULK_SHARE=0 in the
 * environment, or APM_LCL_FLAG_OWN_SOLVE in a case's params.flags, solves every case on its own.  Returns the number
 * of failed cases (0 = all good); a failed case has scalars[k].converged == 0 and the last message in
 * apm_lcl_last_error().  Shard the case list over GPUs by calling this once per device from one
 * thread/process per device (apmapflow_b200.lcl.bulk_solve does that). */
APM_LCL_API int apm_lcl_bulk_solve(int device, int nx, int nz, int n_cases, const apm_lcl_params *params,
                                   const double *const *apertures, apm_lcl_fields *outs, apm_lcl_scalars *scalars);

/* File-level entry point: what `apm-lcl-model.exe file.inp` does (APERTURE_MAP_FLOW.F:1-51).  With the
 * environment variable APM_LCL_DEVICES="0,1,2,3" the map is striped over those GPUs (apm_lcl_solve_striped) and
 * `device` is ignored.  Reads the
 * initialization file (OPEN_IO IO_MODULE.F:51, READ_MASTER APERTURE_MAP_FLOW.F:191), the aperture map
 * (READ_DATA_MAP MAP_MODULE.F:18), runs the hot path on `device` and writes the log, stat (csv+yaml),
 * aperture, flow-x/z/m, pressure and VTK files in the reference's formats.  The transcript the
 * reference prints to stdout is returned in `log` (NUL terminated, truncated at the front if longer
 * than log_cap).  Returns 0, or 1 with the reference's error messages in the transcript. */
APM_LCL_API int apm_lcl_run_file(const char *inp_path, int device, char *log, size_t log_cap);
/* The same with the map striped over `devices` (what run_model(..., devices=[...]) calls). */
APM_LCL_API int apm_lcl_run_file_striped(const char *inp_path, const int *devices, int n_devices, char *log,
                                         size_t log_cap);

/* The statistics the reference prints next to the fields, on the device; `values` / `map` may be
 * host or device pointers.
 * apm_lcl_percentiles: CALC_PERCENTILE (OUTPUT_MODULE.F:53-124): out[j] = the pct[j]-th percentile of the n values in the
 * reference's definition (sorted, duplicates collapsed, first value whose cumulative count exceeds ceil(p/100 n)), found by
 * an exact radix select instead of a sort.
 * apm_lcl_map_stats: MAP_STATS (MAP_MODULE.F:182-303): mean, RMS, standard deviation and the x / z correlation lengths in
 * cells (first lag of the variogram whose semivariance passes 0.9 of the variance); any output pointer may be NULL. */
APM_LCL_API int apm_lcl_percentiles(const double *values, size_t n, const double *pct, int npct, int device, double *out);
APM_LCL_API int apm_lcl_map_stats(const double *map, int nx, int nz, int device, double *avg, double *rms, double *dev,
                                  double *x_corr, double *z_corr);

/* Kernel-level entry points used by the parity tests -------------------------------------- */

/* COEF + FLOW_BOUNDARY: writes WIDTH[nz*nx], TRX[nz][nx+1] (faces 0..NX of each row) and
 * TRZ[nz+1][nx] (faces 0..NZ) in the reference's indexing (APM_MODULE.F:118), any may be NULL. */
APM_LCL_API int apm_lcl_conductance(const apm_lcl_params *p, const double *aperture, int nx, int nz, int device,
                        double *width, double *trx, double *trz, apm_lcl_scalars *sc);

/* Access to the most recent solve's device-resident state (for roofline benches): returns
 * device pointers owned by the solver, valid until the next solve/destroy.  name: "pressure", "width",
 * "tx", "tz", "txl", "tzb"; "r" and "q" (scratch: time_kernel("vcycle") computes q = M^-1 r);
 * diagnostics of the multigrid hierarchy: "nlevels", "gather_level" (first level every rank holds whole;
 * answers in *count), "L<l>.nz", "L<l>.nx", "L<l>.zlo", "L<l>.zhi" (answer in *count), "L<l>.C|S|E|N|NE|NW"
 * (stored 9-point operator of level l >= 1) and "L<l>.hw|he|vs|vn|csw|cse|cnw|cne|th|tv|tc" (interpolation
 * weights level l -> l+1).  A striped solver returns the rows it owns. */
APM_LCL_API int apm_lcl_solver_device_field(apm_lcl_solver *s, const char *name, double **ptr, size_t *count);

/* Times ONE kernel of the solve on the solver's current state, `reps` back-to-back launches on the
 * solver's stream bracketed by CUDA events.  name: "spmv", "update", "direction", "down5", "up5",
 * "smooth9_l1", "residual9_l1", "restrict1", "smooth9_l1_up", "vcycle", "conductance", "flux".  Returns average ms per launch in *ms and the
 * kernel's algorithmic bytes per launch in *bytes. */
APM_LCL_API int apm_lcl_solver_time_kernel(apm_lcl_solver *s, const char *name, int reps, double *ms, double *bytes);

#ifdef __cplusplus
}
#endif
#endif /* APM_LCL_H */
