// common.cuh -- internal declarations shared by the translation units of libapmlcl.so.
// Nothing here crosses the C ABI (include/apm_lcl.h).
#pragma once

#ifndef LCL_CPU_SIM   // tests/sim/cuda_sim.h replaces the runtime when debugging kernels without a GPU
#include <cuda_runtime.h>
#define LCL_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#endif

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/apm_lcl.h"

namespace lcl {

// ---------------------------------------------------------------------------------------------
// error handling / launch accounting (thread local: one solver per host thread is the model)
extern thread_local std::string g_last_error;
extern thread_local long long g_launches;
bool use_pdl();   // api.cu: programmatic dependent launches unless APM_LCL_NO_PDL is set

struct CudaError {
    std::string msg;
};

inline void cuda_check(cudaError_t e, const char *what, const char *file, int line)
{
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %s at %s:%d: %s", what, file, line, cudaGetErrorString(e));
        throw CudaError{buf};
    }
}
#define CUDA_CHECK(x) ::lcl::cuda_check((x), #x, __FILE__, __LINE__)
// every kernel launch goes through this so gpu_launches in bench.py is a count, not a guess
#ifndef LCL_CPU_SIM
#define LAUNCH(kernel, grid, block, smem, stream, ...)                       \
    do {                                                                     \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);          \
        ++::lcl::g_launches;                                                 \
        CUDA_CHECK(cudaGetLastError());                                      \
    } while (0)
// Launch of a kernel of the PCG iteration: allowed to start while its predecessor in the stream drains
// (programmatic dependent launch; the kernel waits in stopped() -> cudaGridDependencySynchronize() before
// it touches memory).  The ~90 launches of an iteration are short dependent kernels on the coarse levels:
// what this hides is their scheduling latency.  APM_LCL_NO_PDL=1 falls back to plain launches.
#define LAUNCH_DEP(kernel, grid_, block_, smem_, stream_, ...)                              \
    do {                                                                                      \
        if (::lcl::use_pdl()) {                                                               \
            cudaLaunchConfig_t cfg_ = {};                                                     \
            cfg_.gridDim = dim3(grid_);                                                       \
            cfg_.blockDim = dim3(block_);                                                     \
            cfg_.dynamicSmemBytes = (smem_);                                                  \
            cfg_.stream = (stream_);                                                          \
            cudaLaunchAttribute attr_[1];                                                     \
            attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                 \
            attr_[0].val.programmaticStreamSerializationAllowed = 1;                          \
            cfg_.attrs = attr_;                                                               \
            cfg_.numAttrs = 1;                                                                \
            CUDA_CHECK(cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__));                       \
            ++::lcl::g_launches;                                                              \
        } else {                                                                              \
            LAUNCH(kernel, grid_, block_, smem_, stream_, __VA_ARGS__);                       \
        }                                                                                     \
    } while (0)
#else
#define LAUNCH(kernel, grid, block, smem, stream, ...)                                        \
    do {                                                                                      \
        ::cusim::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); });     \
        ++::lcl::g_launches;                                                                  \
    } while (0)
#define LAUNCH_DEP LAUNCH
#endif

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Level description.
//   level 0 ("fine"): 5-point operator kept matrix free as face conductances
//        tx[z][x] = right face of cell (z,x)  (x = nx-1 : right boundary face, TRX(IZ,NX))
//        tz[z][x] = top   face of cell (z,x)  (z = nz-1 : top boundary face,   TRZ(NZ,IX))
//        txl[z]   = left boundary face TRX(IZ,0),  tzb[x] = bottom boundary face TRZ(0,IX)
//   level >= 1: Galerkin 9-point operators stored symmetrically in "difference form":
//        couplings E -> (z,x+1), N -> (z+1,x), NE -> (z+1,x+1), NW -> (z+1,x-1) (0 if the neighbour
//        does not exist), absorption S (what is left of the row sum: the Dirichlet boundary terms)
//        and C = S + sum of the eight couplings.  A x is always evaluated as
//        S x_p + sum g (x_p - x_q): a pocket of open cells behind ~1e-14-relative leak conductances
//        keeps its leak information instead of losing it to cancellation (tools/proto_gth.py).
struct Fine5 {
    const double *tx, *tz, *txl, *tzb;
    int nz, nx;
};
struct Coarse9 {
    const double *C, *S, *E, *N, *NE, *NW;
    int nz, nx;
};

// interpolation weights fine(level l) <- coarse(level l+1), one entry per coarse 2x2 block
// (Z,X): hw/he for the h-edge point (2Z,2X+1), vs/vn for the v-edge point (2Z+1,2X),
// csw/cse/cnw/cne for the centre point (2Z+1,2X+1).  th/tv/tc = 1 - (sum of that point's weights),
// computed without subtraction (it is exactly 0 away from Dirichlet boundaries).
struct Weights {
    double *hw, *he, *vs, *vn, *csw, *cse, *cnw, *cne, *th, *tv, *tc;
};

struct Level {
    int nz = 0, nx = 0;
    size_t n = 0;
    // coarse levels only
    double *C = nullptr, *S = nullptr, *E = nullptr, *N = nullptr, *NE = nullptr, *NW = nullptr;
    double *x = nullptr, *b = nullptr, *r = nullptr;
    Weights w{};   // weights to the NEXT coarser level (absent on the coarsest)
    Coarse9 view9() const { return Coarse9{C, S, E, N, NE, NW, nz, nx}; }
};

// PCG scalars live on the device; kernels read alpha/beta from here, the host only polls.
struct PcgState {
    double rz;          // r.z of the current direction
    double rz_new;      // r.z after the latest preconditioner application
    double dq;          // d.A d
    double alpha, beta;
    double rr, rs;      // ||r||^2 and r^T D^-1 r of the latest residual
    double bb, bs;      // same for the right-hand side
    double tol2;        // rel_tol^2
    double dx, dx_prev, dx_prev2; // max |alpha d_i| of the last three updates (pointwise change of the iterate)
    double dx_tol;      // both must be <= dx_tol (absolute, = 2.5e-7 * pressure scale) before stopping
    int res_ok;         // number of iterations at which both residual-norm tests have passed
    int converged;      // set on device when all tests pass
    int failed;         // non-finite / non-positive curvature
    int iter;
};

constexpr int kMaxPartials = 4096;   // per-reduction partial slots (>= any grid we launch)

struct Reducer {
    double *partials = nullptr;      // [kSlots][kMaxPartials]
    unsigned int *counters = nullptr;
};

// ---------------------------------------------------------------------------------------------
// the solver object behind apm_lcl_solver
// The small levels of the hierarchy (<= kTailCells cells each) are cycled by ONE block in one launch
// (k_tail_cycle): by value, so the kernel needs no indirection to find its arrays.
constexpr int kTailMaxLevels = 8, kTailCells = 1024, kTailThreads = 1024;
constexpr size_t kTailMaxSmem = 160 * 1024;   // iterates, residuals and right-hand sides of all tail levels (33 KB for 32x32)
struct TailLevel {
    Coarse9 op;
    Weights w;            // to the next coarser level (unused on the last)
    double *x, *b, *r;
};
struct TailArgs {
    TailLevel lv[kTailMaxLevels];
    int nlev = 0;
    const double *inv = nullptr, *scale = nullptr;   // dense inverse of the last level
    int nu = 2;
};

struct Solver {
    int device = 0, nx = 0, nz = 0;
    size_t n = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // device -> host copies that can run beside the solve
    int num_sms = 148;
    // fine level
    double *ap = nullptr, *width = nullptr, *tx = nullptr, *tz = nullptr, *txl = nullptr, *tzb = nullptr;
    double *x = nullptr, *r = nullptr, *d = nullptr, *q = nullptr, *z = nullptr;
    std::vector<Level> levels;   // levels[0] = fine (only nz,nx,n,w used), then coarse ones
    double *coarse_inv = nullptr;   // dense inverse (Jacobi scaled) of the coarsest operator
    double *coarse_scale = nullptr;
    int coarse_n = 0;
    PcgState *state = nullptr;      // device
    PcgState *state_host = nullptr; // pinned
    double *red = nullptr;          // reduction partials
    unsigned int *red_counters = nullptr;
    double *scal = nullptr;         // misc device scalars (conductance sums, flux reductions)
    double *scal_host = nullptr;    // pinned mirror
    cudaEvent_t ev[6]{};
#ifndef LCL_CPU_SIM
    cudaGraphExec_t iter_graph = nullptr;   // one steady-state PCG iteration (spmv, update, V-cycle, direction)
#endif
    long long iter_graph_launches = 0;
    int nu_fine = 1, nu_coarse = 2, nu_deep = 2;   // sweeps: fine, level 1, levels >= 2
    int tail_top = -1;                             // first level cycled by k_tail_cycle (-1: none)
    TailArgs tail{};
    size_t tail_smem = 0;                          // dynamic shared memory of k_tail_cycle
    int w_lo = 1, w_hi = 0;                        // levels visited twice per cycle (empty range: V-cycle)
    bool have_solution = false;    // x holds the pressure of the last solve
    bool have_hierarchy = false;   // conductances and multigrid operators of the last solve are in place
    bool ap_valid = false;         // ap holds the staged aperture map of the last solve (not flux scratch)
    bool down_configured = false;   // k_down5 opted in to >48 KB dynamic shared memory on this device
    apm_lcl_params last_params{};
    Fine5 view5() const { return Fine5{tx, tz, txl, tzb, nz, nx}; }
};

// conductance.cu -------------------------------------------------------------------------------
// scal layout written by the conductance kernel
enum { SC_RAW_SUM = 0, SC_CLAMP_SUM, SC_MIN_POS, SC_MAX, SC_ZEROS, SC_MASKED, SC_FLOW, SC_MAXRES_ABS,
       SC_MAXRES_SIGNED, SC_COUNT = 16 };
void launch_conductance(Solver &s, const apm_lcl_params &p, const double *ap_dev);
void export_reference_layout(Solver &s, double *trx_dev, double *trz_dev);   // [nz][nx+1], [nz+1][nx]

// mg_setup.cu ----------------------------------------------------------------------------------
void mg_allocate(Solver &s);
void mg_free(Solver &s);
void mg_setup(Solver &s);   // weights + Galerkin operators + coarsest inverse, from s.tx/s.tz

// mg_cycle.cu ----------------------------------------------------------------------------------
// z = M^-1 r on the fine level.  The first kernel also accumulates ||r||^2 and r^T D^-1 r into st
// (convergence test; with init they become the norms of the right-hand side) and the last one r.z
// (-> st->rz, st->beta).  Every kernel is a no-op once st->converged or st->failed is set.
// ztmp is scratch (the pre-smoothed iterate), the result lands in zout.
void mg_vcycle(Solver &s, const double *r, double *ztmp, double *zout, PcgState *st, bool init);

// pcg.cu ---------------------------------------------------------------------------------------
struct BoundaryValues {
    double left, right, bottom, top;   // pressure outside each side (only Dirichlet sides matter)
};
// solves A x = b(bv) into s.x; returns iterations; fills residual norms
int pcg_solve(Solver &s, const BoundaryValues &bv, double rel_tol, int max_iter, apm_lcl_scalars *sc);

// single-kernel launchers for apm_lcl_solver_time_kernel (roofline benches); each returns the
// kernel's algorithmic bytes per launch, or a negative value if the name is not theirs
double bench_launch_pcg(Solver &s, const std::string &name);
double bench_launch_cycle(Solver &s, const std::string &name);

// flux.cu --------------------------------------------------------------------------------------
struct FluxOut {
    double *qx, *qz, *qm, *ren, *res;   // device pointers or nullptr
};
void launch_flux(Solver &s, const apm_lcl_params &p, const double *pres, const BoundaryValues &bv,
                 const FluxOut &o);
void launch_affine(Solver &s, double *x, size_t n, double scale, double shift);   // x = shift + scale*x

}   // namespace lcl
