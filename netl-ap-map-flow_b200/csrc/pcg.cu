// pcg.cu -- matrix-free preconditioned conjugate gradients for the LCL pressure system.
//
// The system is the one SOLVER assembles (APERTURE_MAP_FLOW.F:628-650) with the Dirichlet rows of
// ADD_DIRICHLET_BOUND_COND (D4_SOLVER_MODULE.F:170-257) folded in:  sum_f T_f (p_i - p_nbr) = 0 per
// cell, boundary faces contributing T_b p_i on the diagonal and T_b P_b on the right-hand side.
// TT/CC are never materialised: every kernel reads the face conductances tx/tz directly.
// Algorithmic bytes per cell (f64): spmv 32 (d, tx, tz in; q out), update 48 (x, d, r, q in; x, r
// out), direction 24 (z, d in; d out).  All scalars (alpha, beta, norms, convergence flag) stay on
// the device; the host polls one small struct.
#include <cstdlib>

#include "device_utils.cuh"

namespace lcl {

__device__ __forceinline__ bool stopped(const PcgState *st)
{
#ifndef LCL_CPU_SIM
    cudaGridDependencySynchronize();   // LAUNCH_DEP: the predecessor's results are visible from here on
#endif
    return st->converged | st->failed;
}

// r = b (x0 = 0), x = 0
__global__ void __launch_bounds__(kBlock)
k_init_rhs(Fine5 v, BoundaryValues bv, double *__restrict__ r, double *__restrict__ x)
{
    const int nz = v.nz, nx = v.nx;
    const Tiles tiles(v.zlo, v.zhi, nx);
    for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
        int z, xx;
        tiles.cell(t, z, xx);
        if (z >= v.zhi || xx >= nx) continue;
        const size_t i = (size_t)z * nx + xx;
        double b = 0.0;
        if (xx == 0) b += v.txl[z] * bv.left;
        if (xx == nx - 1) b += v.tx[i] * bv.right;
        if (z == 0) b += v.tzb[xx] * bv.bottom;
        if (z == nz - 1) b += v.tz[i] * bv.top;
        r[i] = b;
        x[i] = 0.0;
    }
}

// q = A d and d.q;  alpha = rz / d.q
// VEC (nx even): a thread takes two neighbouring cells and moves d, the faces and q as 128-bit words
// (LDG.E.128 / STG.E.128: half the memory instructions for the same bytes); the arithmetic per cell is the scalar
// kernel's, term by term.
__device__ __forceinline__ void spmv5_finish(PcgState *st, const Dist &dc, const double (&part)[1])
{
    double tot[1] = {part[0]};
    dist_sum<1>(dc, tot);
    st->dq = tot[0];
    if (!(tot[0] > 0.0) || !(st->rz == st->rz)) {
        // curvature lost to rounding: harmless once the residual tests have passed (floating
        // pockets), a failure otherwise
        if (st->res_ok) st->converged = 1; else st->failed = 1;
        st->alpha = 0.0;
    } else {
        st->alpha = st->rz / tot[0];
    }
    st->iter += 1;
}

template <bool VEC>
__global__ void __launch_bounds__(kBlock)
k_spmv5(Fine5 v, const double *__restrict__ d, double *__restrict__ q, PcgState *st, double *partials,
        unsigned int *counter, Dist dc)
{
    if (stopped(st)) return;
    const F5 A{v};
    const int nz = v.nz, nx = v.nx;
    double acc[1] = {0.0};
    if (VEC) {
        const int np = nx >> 1;                       // cell pairs per row
        const Tiles tiles(v.zlo, v.zhi, np);
        for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
            int z, p;
            tiles.cell(t, z, p);
            if (z >= v.zhi || p >= np) continue;
            const int x0 = 2 * p;
            const size_t i = (size_t)z * nx + x0;
            const double2 dc2 = *reinterpret_cast<const double2 *>(d + i);
            const double2 tx2 = *reinterpret_cast<const double2 *>(v.tx + i);
            const double2 tu2 = *reinterpret_cast<const double2 *>(v.tz + i);
            const double2 td2 = z > 0 ? *reinterpret_cast<const double2 *>(v.tz + i - nx) : *reinterpret_cast<const double2 *>(v.tzb + x0);
            const double2 dd2 = z > 0 ? *reinterpret_cast<const double2 *>(d + i - nx) : make_double2(0.0, 0.0);
            const double2 du2 = z + 1 < nz ? *reinterpret_cast<const double2 *>(d + i + nx) : make_double2(0.0, 0.0);
            const double tl = x0 > 0 ? v.tx[i - 1] : v.txl[z];
            const double dl = x0 > 0 ? d[i - 1] : 0.0, dr = x0 + 2 < nx ? d[i + 2] : 0.0;
            // difference form (boundary faces: homogeneous Dirichlet value 0)
            double s0 = tl * (dc2.x - dl);
            s0 += tx2.x * (dc2.x - dc2.y);
            s0 += td2.x * (dc2.x - dd2.x);
            s0 += tu2.x * (dc2.x - du2.x);
            double s1 = tx2.x * (dc2.y - dc2.x);
            s1 += tx2.y * (dc2.y - dr);
            s1 += td2.y * (dc2.y - dd2.y);
            s1 += tu2.y * (dc2.y - du2.y);
            *reinterpret_cast<double2 *>(q + i) = make_double2(s0, s1);
            acc[0] += dc2.x * s0;
            acc[0] += dc2.y * s1;
        }
    } else {
        const Tiles tiles(v.zlo, v.zhi, nx);
        for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
            int z, xx;
            tiles.cell(t, z, xx);
            if (z >= v.zhi || xx >= nx) continue;
            const size_t i = (size_t)z * nx + xx;
            const double tl = A.left(z, xx), tr = A.right(z, xx), td = A.down(z, xx), tu = A.up(z, xx);
            const double di = d[i];
            // difference form (boundary faces: homogeneous Dirichlet value 0)
            double s = tl * (di - (xx > 0 ? d[i - 1] : 0.0));
            s += tr * (di - (xx + 1 < nx ? d[i + 1] : 0.0));
            s += td * (di - (z > 0 ? d[i - nx] : 0.0));
            s += tu * (di - (z + 1 < nz ? d[i + nx] : 0.0));
            q[i] = s;
            acc[0] += di * s;
        }
    }
    grid_sum_finalize<1>(acc, partials, counter, [=](const double(&part)[1]) { spmv5_finish(st, dc, part); });
}

// x += alpha d ; r -= alpha q ; max |alpha d_i| (how far any single pressure moved) rides along
__global__ void __launch_bounds__(256)
k_update(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ d, const double *__restrict__ q,
         size_t n, PcgState *st, double *partials, unsigned int *counter)
{
    if (stopped(st)) return;
    const double alpha = st->alpha;
    double dmax = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double di = d[i];
        x[i] += alpha * di;
        r[i] -= alpha * q[i];
        dmax = fmax(dmax, fabs(di));
    }
    // this strip's largest change; the ranks' values meet in the gather of the down leg (mg_cycle.cu), which
    // is where the stopping rule looks at them
    grid_max_finalize(dmax, partials, counter, [=](double m) {
        st->dx_prev2 = st->dx_prev;
        st->dx_prev = st->dx;
        st->dx_part = fabs(alpha) * m;
    });
}

// d = z + beta d
__global__ void __launch_bounds__(256)
k_direction(double *__restrict__ d, const double *__restrict__ z, size_t n, const PcgState *st, int init)
{
    if (stopped(st)) return;
    const double beta = st->beta;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        d[i] = init ? z[i] : z[i] + beta * d[i];
}

// true residual r = b - A x with both norms (verification after the recurrence says "converged")
__global__ void __launch_bounds__(kBlock)
k_true_residual(Fine5 v, BoundaryValues bv, const double *__restrict__ x, double *__restrict__ r, double *out,
                double *partials, unsigned int *counter, Dist dc)
{
    const F5 A{v};
    const int nz = v.nz, nx = v.nx;
    const Tiles tiles(v.zlo, v.zhi, nx);
    double acc[2] = {0.0, 0.0};
    for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
        int z, xx;
        tiles.cell(t, z, xx);
        if (z >= v.zhi || xx >= nx) continue;
        const size_t i = (size_t)z * nx + xx;
        const double tl = A.left(z, xx), tr = A.right(z, xx), td = A.down(z, xx), tu = A.up(z, xx);
        const double dg = (tl + tr) + (td + tu);
        const double xi = x[i];
        double s = tl * ((xx > 0 ? x[i - 1] : bv.left) - xi);
        s += tr * ((xx + 1 < nx ? x[i + 1] : bv.right) - xi);
        s += td * ((z > 0 ? x[i - nx] : bv.bottom) - xi);
        s += tu * ((z + 1 < nz ? x[i + nx] : bv.top) - xi);
        r[i] = s;
        acc[0] += s * s;
        acc[1] += s * s / dg;
    }
    grid_sum_finalize<2>(acc, partials, counter, [=](const double(&part)[2]) {
        double tot[2] = {part[0], part[1]};
        dist_sum<2>(dc, tot);
        out[0] = tot[0];
        out[1] = tot[1];
    });
}

// d = z + beta d runs one row beyond the strip either side: z arrives there by halo exchange (end of the cycle),
// so d's ghost rows (read by the next spmv) need no exchange of their own
static void launch_direction(Solver &s, const double *z, int init, bool dep)
{
    const int lo = s.levels[0].striped && s.zlo > 0 ? s.zlo - 1 : s.zlo;
    const int hi = s.levels[0].striped && s.zhi < s.nz ? s.zhi + 1 : s.zhi;
    const size_t off = (size_t)lo * s.nx, cnt = (size_t)(hi - lo) * s.nx;
    const int g1 = s.num_sms * 8;
    if (dep) LAUNCH_DEP(k_direction, g1, 256, 0, s.stream, s.d + off, z + off, cnt, (const PcgState *)s.state, init);
    else LAUNCH(k_direction, g1, 256, 0, s.stream, s.d + off, z + off, cnt, (const PcgState *)s.state, init);
}

static void launch_spmv(Solver &s, bool dep)
{
    const Fine5 v5 = s.view5();
    const bool vec = !(s.nx & 1);   // pairs of cells as 128-bit words need rows that start on 16 bytes
    const int g0 = grid_for(s, s.zhi - s.zlo, vec ? s.nx / 2 : s.nx, 8);
    if (vec) {
        if (dep) LAUNCH_DEP(k_spmv5<true>, g0, kBlock, 0, s.stream, v5, (const double *)s.d, s.q, s.state, s.red, s.red_counters, s.dist);
        else LAUNCH(k_spmv5<true>, g0, kBlock, 0, s.stream, v5, (const double *)s.d, s.q, s.state, s.red, s.red_counters, s.dist);
    } else {
        if (dep) LAUNCH_DEP(k_spmv5<false>, g0, kBlock, 0, s.stream, v5, (const double *)s.d, s.q, s.state, s.red, s.red_counters, s.dist);
        else LAUNCH(k_spmv5<false>, g0, kBlock, 0, s.stream, v5, (const double *)s.d, s.q, s.state, s.red, s.red_counters, s.dist);
    }
}

double bench_launch_pcg(Solver &s, const std::string &name)
{
    const Fine5 v5 = s.view5();
    const int g0 = grid_for(s, s.zhi - s.zlo, s.nx, 8), g1 = s.num_sms * 8;
    const double n = (double)s.own_cells();
    if (name == "spmv") {
        launch_spmv(s, false);
        return 32.0 * n;
    }
    if (name == "update") {
        LAUNCH(k_update, g1, 256, 0, s.stream, s.own(s.x), s.own(s.r), (const double *)s.own(s.d), (const double *)s.own(s.q),
               s.own_cells(), s.state, s.red + 4 * kMaxPartials, s.red_counters + 5);
        return 48.0 * n;
    }
    if (name == "direction") {
        launch_direction(s, s.q, 0, false);
        return 24.0 * n;
    }
    return -1.0;
}

static void poll_state(Solver &s)
{
    CUDA_CHECK(cudaMemcpyAsync(s.state_host, s.state, sizeof(PcgState), cudaMemcpyDeviceToHost, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
}

int pcg_solve(Solver &s, const BoundaryValues &bv, double rel_tol, int max_iter, apm_lcl_scalars *sc)
{
    const Fine5 v5 = s.view5();
    const int g0 = grid_for(s, s.zhi - s.zlo, s.nx, 8);
    const int g1 = s.num_sms * 8;
    double *part = s.red;                 // slots 0..15 belong to pcg.cu / conductance / flux
    unsigned int *cnt = s.red_counters;
    const size_t n = s.n, own = s.own_cells();
    // cells per poll: small maps are launch bound, so look at the flag less often (the same on every rank)
    const int check_every = (n >= (size_t)4 << 20 || getenv("APM_LCL_DEBUG")) ? 1 : (n >= (size_t)1 << 18 ? 2 : 4);
    const bool striped = s.levels[0].striped;

    PcgState h{};
    h.tol2 = rel_tol * rel_tol;
    {   // pointwise stagnation test: no pressure moved by more than 2.5e-7 of the pressure scale in the
        // last two updates (global residual norms do not see hydraulically isolated islands)
        const double ps = fmax(fmax(fabs(bv.left), fabs(bv.right)), fmax(fabs(bv.bottom), fabs(bv.top)));
        const char *knob = getenv("APM_LCL_DX_TOL");   // relative; 2.5e-7 keeps weakly connected cells within the 1e-6 parity bar
        h.dx_tol = (knob ? atof(knob) : 2.5e-7) * (ps > 0 ? ps : 1.0);
        h.dx = h.dx_prev = h.dx_prev2 = h.dx_part = 1e300;
    }
    const bool debug = getenv("APM_LCL_DEBUG") != nullptr;
    CUDA_CHECK(cudaMemcpyAsync(s.state, &h, sizeof h, cudaMemcpyHostToDevice, s.stream));
    LAUNCH(k_init_rhs, g0, kBlock, 0, s.stream, v5, bv, s.r, s.x);
    // z = M^-1 r lands in s.q: the cycle is out of place and q (= A d) is dead by then; the next
    // spmv overwrites it after the direction update has consumed it
    mg_vcycle(s, s.r, s.z, s.q, s.state, true);
    launch_direction(s, s.q, 1, false);

    int iters = 0, restarts = 0;
    bool done = false;
    double true_rr = 0.0, true_rs = 0.0, bb = 0.0, bs = 0.0;
    auto peer_error = [&]() {
        if (s.world == 1) return false;
        int e = 0;
        CUDA_CHECK(cudaMemcpyAsync(&e, &s.dist.local->error, sizeof e, cudaMemcpyDeviceToHost, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        return e != 0;
    };
    poll_state(s);
    if (peer_error()) {   // a peer went missing during the set-up or the first cycle
        g_last_error = peer_error_report(s);
        return -1;
    }
    bb = s.state_host->bb; bs = s.state_host->bs;
    if (!(bb > 0.0)) {   // zero right-hand side: x = 0 is the solution
        if (sc) { sc->iterations += 0; sc->converged = 1; sc->rel_residual = 0.0; sc->rel_residual_scaled = 0.0; }
        return 0;
    }
    auto enqueue_iteration = [&]() {
        launch_spmv(s, true);
        LAUNCH_DEP(k_update, g1, 256, 0, s.stream, s.own(s.x), s.own(s.r), (const double *)s.own(s.d),
                   (const double *)s.own(s.q), own, s.state, part + 4 * kMaxPartials, cnt + 5);
        mg_vcycle(s, s.r, s.z, s.q, s.state, false);
        launch_direction(s, s.q, 0, true);
    };
#ifndef LCL_CPU_SIM
    // Every argument of the steady-state iteration is a pointer or size fixed for the solver's lifetime
    // (the scalars live in s.state on the device, the exchange counters in the slab), so the launches are
    // captured once into a CUDA graph and replayed: small maps are launch bound otherwise.
    if (!s.iter_graph && !getenv("APM_LCL_NO_GRAPH")) {
        const long long before = g_launches;
        cudaGraph_t graph = nullptr;
        CUDA_CHECK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        enqueue_iteration();
        CUDA_CHECK(cudaStreamEndCapture(s.stream, &graph));
        CUDA_CHECK(cudaGraphInstantiate(&s.iter_graph, graph, 0));
        CUDA_CHECK(cudaGraphDestroy(graph));
        s.iter_graph_launches = g_launches - before;
        g_launches = before;
    }
#endif
    auto run_iteration = [&]() {
#ifndef LCL_CPU_SIM
        if (s.iter_graph) {
            CUDA_CHECK(cudaGraphLaunch(s.iter_graph, s.stream));
            g_launches += s.iter_graph_launches;
            return;
        }
#endif
        enqueue_iteration();
    };
    while (!done) {
        while (iters < max_iter) {
            for (int k = 0; k < check_every && iters < max_iter; k++, iters++) run_iteration();
            poll_state(s);
            if (debug)
                fprintf(stderr, "[pcg %d/%d] it %3d rr/bb %.3e rs/bs %.3e rz %.3e dq %.3e alpha %.3e beta %.3e dx %.3e conv %d fail %d\n",
                        s.rank, s.world, s.state_host->iter, sqrt(s.state_host->rr / bb), sqrt(s.state_host->rs / bs), s.state_host->rz,
                        s.state_host->dq, s.state_host->alpha, s.state_host->beta, s.state_host->dx,
                        s.state_host->converged, s.state_host->failed);
            if (s.state_host->converged || s.state_host->failed) break;
        }
        const int dev_iters = s.state_host->iter;
        // verify on the true residual (the recurrence can drift under 1e16 contrast)
        if (striped) {
            HaloSet hs;
            hs.add(s.x);
            halo_exchange(s, hs, s.nx, s.levels[0].cut, 1, 1, nullptr);
        }
        LAUNCH(k_true_residual, g0, kBlock, 0, s.stream, v5, bv, (const double *)s.x, s.q, s.scal + 12,
               part + 8 * kMaxPartials, cnt + 2, s.dist);
        CUDA_CHECK(cudaMemcpyAsync(s.scal_host + 12, s.scal + 12, 2 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
        CUDA_CHECK(cudaStreamSynchronize(s.stream));
        if (peer_error()) {
            g_last_error = peer_error_report(s);
            return -1;
        }
        true_rr = s.scal_host[12]; true_rs = s.scal_host[13];
        const bool ok = true_rr <= h.tol2 * bb && true_rs <= h.tol2 * bs;
        iters = dev_iters;
        if (ok || s.state_host->failed || iters >= max_iter || restarts >= 3) {
            done = true;
            if (sc) {
                sc->iterations += iters;
                sc->converged = ok ? 1 : 0;
                sc->rel_residual = sqrt(true_rr / bb);
                sc->rel_residual_scaled = sqrt(true_rs / bs);
            }
            if (s.state_host->failed && !ok) {
                g_last_error = "PCG breakdown (non-positive curvature or non-finite residual)";
                return -1;
            }
        } else {
            // restart from the true residual: r := b - A x, z = M^-1 r, d = z
            restarts++;
            CUDA_CHECK(cudaMemcpyAsync(s.own(s.r), s.own(s.q), own * sizeof(double), cudaMemcpyDeviceToDevice, s.stream));
            PcgState h2 = *s.state_host;
            h2.converged = 0; h2.failed = 0;
            CUDA_CHECK(cudaMemcpyAsync(s.state, &h2, sizeof h2, cudaMemcpyHostToDevice, s.stream));
            CUDA_CHECK(cudaStreamSynchronize(s.stream));
            mg_vcycle(s, s.r, s.z, s.q, s.state, false);
            // after this cycle beta = rz_new/rz_old is meaningless: force a steepest-descent restart
            launch_direction(s, s.q, 1, false);
            poll_state(s);
            if (s.state_host->converged) {
                // the cycle's own test passed on the true residual it was given
                continue;
            }
        }
    }
    return iters;
}

// forces the kernels of this file into the device's context now (see preload_kernels, api.cu)
void preload_pcg()
{
    preload_one(k_init_rhs);
    preload_one(k_spmv5<true>);
    preload_one(k_spmv5<false>);
    preload_one(k_update);
    preload_one(k_direction);
    preload_one(k_true_residual);
}

}   // namespace lcl
