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

// Loads a kernel into the current device's context.  CUDA loads kernels lazily, at their first launch, and
// that load synchronises with the context: a rank spinning on a peer's flag would then keep the peer --
// another solver of the same process on the same GPU -- from ever launching the kernel that raises the flag.
// Every solver therefore loads all kernels when it is created (also keeps first-launch cost out of timings).
#ifndef LCL_CPU_SIM
template <class K>
inline void preload_one(K kernel)
{
    cudaFuncAttributes attr;
    CUDA_CHECK(cudaFuncGetAttributes(&attr, kernel));
}
#else
template <class K>
inline void preload_one(K) {}
#endif
void preload_conductance();
void preload_flux();
void preload_pcg();
void preload_peer();
void preload_cycle();
void preload_setup();
void preload_stats();

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
//
// Row strips (one map on several GPUs): a level is either striped -- rank k owns the contiguous
// rows [zlo, zhi) and keeps kHalo ghost rows either side, filled by its neighbours through peer memory
// (dist.cuh) -- or replicated on every rank (zlo = 0, zhi = nz).  Every array pointer of a striped level is
// a VIRTUAL base: element (z, x) of the global grid sits at p[z * nx + x] for the rows a rank holds, so no
// kernel translates indices; nz, nx stay the global dimensions (they decide where the physical boundary
// is), zlo/zhi say which rows a kernel writes and reduces over.  One GPU is the case zlo = 0, zhi = nz.
constexpr int kHalo = 6;   // ghost rows kept either side of a strip (the fine down leg reads 4 + 1 face row)
struct Fine5 {
    const double *tx, *tz, *txl, *tzb;
    int nz, nx;
    int zlo, zhi;
};
struct Coarse9 {
    const double *C, *S, *E, *N, *NE, *NW;
    int nz, nx;
    int zlo, zhi;
};
// rows of a striped level a rank may READ: its own plus the ghost rows it holds, inside the grid
__host__ __device__ inline int read_lo(int zlo) { return zlo - kHalo > 0 ? zlo - kHalo : 0; }
__host__ __device__ inline int read_hi(int zhi, int nz) { return zhi + kHalo < nz ? zhi + kHalo : nz; }

// interpolation weights fine(level l) <- coarse(level l+1), one entry per coarse 2x2 block
// (Z,X): hw/he for the h-edge point (2Z,2X+1), vs/vn for the v-edge point (2Z+1,2X),
// csw/cse/cnw/cne for the centre point (2Z+1,2X+1).  th/tv/tc = 1 - (sum of that point's weights),
// computed without subtraction (it is exactly 0 away from Dirichlet boundaries).
struct Weights {
    double *hw, *he, *vs, *vn, *csw, *cse, *cnw, *cne, *th, *tv, *tc;
};

constexpr int kMaxRanks = 16;

struct Level {
    int nz = 0, nx = 0;
    size_t n = 0;              // cells of the whole level
    int zlo = 0, zhi = 0;      // rows this rank owns (0, nz unless the level is striped)
    bool striped = false;      // rows are dealt to the ranks (else every rank holds the whole level)
    int cut[kMaxRanks + 1]{};  // striped: rank k owns rows [cut[k], cut[k+1])
    int czlo = 0, czhi = 0;    // coarse rows (next level's numbering) this rank computes: weights, restriction
    int ccut[kMaxRanks + 1]{}; // the same for every rank (the weight arrays are striped by these)
    // coarse levels only
    double *C = nullptr, *S = nullptr, *E = nullptr, *N = nullptr, *NE = nullptr, *NW = nullptr;
    double *x = nullptr, *b = nullptr, *r = nullptr;
    Weights w{};   // weights to the NEXT coarser level (absent on the coarsest)
    Coarse9 view9() const { return Coarse9{C, S, E, N, NE, NW, nz, nx, zlo, zhi}; }
    size_t own_cells() const { return (size_t)(zhi - zlo) * nx; }
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
    double dx_part;     // this rank's share of the latest one (its strip), combined over ranks by the next down leg
    double dx_tol;      // both must be <= dx_tol (absolute, = 2.5e-7 * pressure scale) before stopping
    int res_ok;         // number of iterations at which both residual-norm tests have passed
    int converged;      // set on device when all tests pass
    int failed;         // non-finite / non-positive curvature
    int iter;
};

constexpr int kMaxPartials = 4096;   // per-reduction partial slots (>= any grid we launch)

// ---------------------------------------------------------------------------------------------
// One map on several GPUs: what the ranks of a striped solve write into each other's memory.  Every rank
// allocates ONE slab with the same layout, exports it (CUDA IPC) and maps the slabs of the others; a
// pointer into the own slab plus peer_off[k] is the same object on rank k.  All remote accesses are
// STORES into the peer's memory followed by a flag store; every wait polls LOCAL memory (peer loads bypass
// L2 and may be served stale from L1).  Counters only grow, every rank runs the same sequence of kernels,
// and a wait refers to a kernel the peer enqueued no later than this one: no cycle, hence no deadlock;
// waits are bounded anyway (dist.cuh) and raise `error` instead of hanging the GPU.
constexpr int kGatherMax = 12;   // doubles per rank in one gather
struct PeerWindow {
    unsigned long long halo_flag[2];              // [0] written by the rank below, [1] by the rank above: exchange count
    unsigned long long gather_flag[kMaxRanks];    // written by rank k: its gather count
    unsigned long long bcast_flag[kMaxRanks];     // written by rank k: its row-broadcast count
    double gather_slot[2][kMaxRanks][kGatherMax]; // rank k's values, double buffered by the parity of the count
};
struct DistLocal {                                // local bookkeeping of the same sequence (never written by peers)
    unsigned long long halo_seq, gather_seq, bcast_seq;
    unsigned int block_counter[4];                // "last block done" of the exchange kernels
    int error;                                    // a wait timed out: 1 halo, 2 gather, 3 broadcast (first one wins)
    int error_peer;                               // ... on the flag of this rank
    unsigned long long error_want, error_have;    // ... which should have reached / had reached this count
    unsigned long long timeout_ns;                // how long a wait may last (APM_LCL_PEER_TIMEOUT_MS, default 20 s)
};
struct Dist {
    int rank = 0, world = 1;
    PeerWindow *win[kMaxRanks]{};     // every rank's window as THIS rank addresses it (win[rank] is local)
    long long peer_off[kMaxRanks]{};  // bytes from an address in the own slab to the same object of rank k
    DistLocal *local = nullptr;
};

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
    // strips (world == 1: one strip holding every row)
    int rank = 0, world = 1;
    int zlo = 0, zhi = 0;           // fine rows this rank owns
    int gather_level = 0;           // first level every rank holds whole (levels below it are striped); 0: none striped
    Dist dist{};
    char *slab = nullptr;           // every device array of the solver lives in this one allocation
    size_t slab_bytes = 0;
    void *peer_slab[kMaxRanks]{};   // the other ranks' slabs as mapped here
    bool peer_ipc[kMaxRanks]{};     // mapped through CUDA IPC (a rank of another process) rather than addressed directly
    bool connected = false;
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
    apm_lcl_scalars last_scalars{};
    Fine5 view5() const { return Fine5{tx, tz, txl, tzb, nz, nx, zlo, zhi}; }
    size_t own_cells() const { return (size_t)(zhi - zlo) * nx; }
    double *own(double *p) const { return p + (size_t)zlo * nx; }   // first owned element of a fine-level array
};

// Bump allocator over the solver's slab.  The layout runs twice: a counting pass (base == nullptr) that
// only measures, then the real one.  Every rank lays out the same sizes in the same order, so an object has
// the same offset in every rank's slab (what Dist::peer_off relies on).
struct Arena {
    char *base = nullptr;
    size_t used = 0;
    void *take_bytes(size_t bytes)
    {
        used = (used + 255) & ~(size_t)255;
        void *p = base + used;   // counting pass: a fake address nobody dereferences
        used += bytes ? bytes : 8;
        return p;
    }
    double *take(size_t count) { return (double *)take_bytes(count * sizeof(double)); }
    // array of `pitch`-wide rows: striped (every rank allocates the tallest strip plus ghost rows and gets the
    // virtual base that makes global row numbers work) or whole
    double *take_rows(int pitch, int nz, bool striped, const int *cut, int rank, int world)
    {
        if (!striped) return take((size_t)nz * pitch);
        int tallest = 0;
        for (int k = 0; k < world; k++) tallest = cut[k + 1] - cut[k] > tallest ? cut[k + 1] - cut[k] : tallest;
        double *p = take((size_t)(tallest + 2 * kHalo) * pitch);
        return p - (long long)(cut[rank] - kHalo) * pitch;
    }
};

// peer.cu --------------------------------------------------------------------------------------
constexpr int kPeerHandleBytes = 128;
void slab_create(Solver &s, size_t bytes, void *handle_out);   // one zeroed allocation (+ its IPC handle if world > 1)
void slab_connect(Solver &s, const void *handles);             // maps the peers' slabs, fills s.dist
void slab_disconnect(Solver &s);
std::string peer_error_report(Solver &s);   // "" unless a wait on a peer timed out
void slab_destroy(Solver &s);
struct HaloSet {
    double *a[12];
    int count = 0;
    void add(double *p) { a[count++] = p; }
};
// Ghost rows of striped arrays (row pitch `pitch`, strips cut[0..world]): every rank stores its lowest `dhi`
// rows into the upper ghost rows of the rank below and its highest `dlo` rows into the lower ghost rows of
// the rank above, then waits for its own ghosts.  A no-op when world == 1.
void halo_exchange(Solver &s, const HaloSet &hs, int pitch, const int *cut, int dlo, int dhi, const PcgState *st);
// Rows [cut[rank], cut[rank+1]) of REPLICATED arrays go to every other rank (an all-gather by stores).
void row_broadcast(Solver &s, const HaloSet &hs, int pitch, const int *cut, const PcgState *st);

// conductance.cu -------------------------------------------------------------------------------
// scal layout written by the conductance kernel
enum { SC_RAW_SUM = 0, SC_CLAMP_SUM, SC_MIN_POS, SC_MAX, SC_ZEROS, SC_MASKED, SC_FLOW, SC_MAXRES_ABS,
       SC_MAXRES_SIGNED, SC_COUNT = 16 };
void launch_conductance(Solver &s, const apm_lcl_params &p, const double *ap_dev);
void export_reference_layout(Solver &s, double *trx_dev, double *trz_dev);   // [nz][nx+1], [nz+1][nx]

// mg_setup.cu ----------------------------------------------------------------------------------
void mg_plan(Solver &s);                  // level sizes, which levels are striped and where the strips are cut
void mg_allocate(Solver &s, Arena &a);    // takes every level's arrays from the arena (called for both passes)
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
