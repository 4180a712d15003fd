// flux.cu -- per-cell fluxes and derived output fields, one fused pass over the pressure map.
//
// Replaces FRAC_FLOW (apmapflow/src/APERTURE_MAP_FLOW.F:729-762: halo pressures, Q1..Q5) and the
// derived fields of OUTPUT_DATA (:825-858: RES, QX/QZ face choice, QM, REN, outlet FLOW sum with the
// reference's face choice LEFT->Q1, RIGHT->Q2, TOP->Q3, BOTTOM->Q4, MAXLOC of |Q5|).
// Algorithmic bytes per cell: read p, tx, tz, width (32 B), write qx, qz, qm, ren, res (40 B);
// RES needs max|Q5| first, so it is normalised by a second 16 B/cell pass.
#include "device_utils.cuh"

namespace lcl {

struct FluxParams {
    double rho_over_mu;
    int side;
};

// What FRAC_FLOW and OUTPUT_DATA do for ONE cell, given its pressure, its four neighbours' and its four faces; shared by the
// scalar and the paired kernel so that both do the same arithmetic term by term.
struct FluxAcc {
    double flow = 0.0;                // outlet FLOW
    double best = -1.0, best_signed = 0.0;
    long long best_idx = -1;          // MAXLOC: first maximum in cell order
};
struct CellOut {
    double qx, qz, qm, ren, a5;
};
__device__ __forceinline__ CellOut flux_cell(const FluxParams &fp, int nz, int nx, int z, int x, size_t i, double pc, double pl,
                                             double pr, double pd, double pu, double tl, double tr, double td, double tu, double w,
                                             FluxAcc &a)
{
    const double q1 = -tl * (pc - pl);    // :756
    const double q2 = -tr * (pr - pc);    // :757
    const double q3 = -td * (pc - pd);    // :758
    const double q4 = -tu * (pu - pc);    // :759
    const double q5 = ((q1 - q2) + (q3 - q4));      // :760
    CellOut c;
    c.qx = fp.side == APM_LCL_LEFT ? q1 : q2;       // :829-835
    c.qz = fp.side == APM_LCL_BOTTOM ? q3 : q4;
    c.qm = sqrt(c.qx * c.qx + c.qz * c.qz);         // :838
    c.ren = fp.rho_over_mu * (c.qm / w);            // :839
    c.a5 = fabs(q5);
    if (c.a5 > a.best || (c.a5 == a.best && (long long)i < a.best_idx)) { a.best = c.a5; a.best_signed = q5; a.best_idx = (long long)i; }
    if (fp.side == APM_LCL_LEFT && x == 0) a.flow += q1;      // :853-856
    if (fp.side == APM_LCL_RIGHT && x == nx - 1) a.flow += q2;
    if (fp.side == APM_LCL_TOP && z == nz - 1) a.flow += q3;
    if (fp.side == APM_LCL_BOTTOM && z == 0) a.flow += q4;
    return c;
}

// VEC (nx even, every array on a 16-byte boundary): a thread takes two neighbouring cells and moves the pressures, faces,
// widths and the five outputs as 128-bit words, like k_spmv5<true>; nine streams of 8-byte accesses are what kept the scalar
// kernel at 0.57 of the copy peak.
template <bool VEC>
__global__ void __launch_bounds__(kBlock)
k_flux(Fine5 v, const double *__restrict__ p, const double *__restrict__ width, BoundaryValues bv, FluxParams fp,
       FluxOut o, double *partials, unsigned int *counter, double *scal, Dist dc)
{
    const F5 A{v};
    const int nz = v.nz, nx = v.nx;
    FluxAcc a;
    if (VEC) {
        const int np = nx >> 1;
        const Tiles tiles(v.zlo, v.zhi, np);   // a strip's own rows; p has a ghost row either side
        for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
            int z, pp;
            tiles.cell(t, z, pp);
            if (z >= v.zhi || pp >= np) continue;
            const int x0 = 2 * pp;
            const size_t i = (size_t)z * nx + x0;
            const double2 pc2 = *reinterpret_cast<const double2 *>(p + i);
            const double2 tx2 = *reinterpret_cast<const double2 *>(v.tx + i);
            const double2 tu2 = *reinterpret_cast<const double2 *>(v.tz + i);
            const double2 td2 = z > 0 ? *reinterpret_cast<const double2 *>(v.tz + i - nx) : *reinterpret_cast<const double2 *>(v.tzb + x0);
            const double2 pd2 = z > 0 ? *reinterpret_cast<const double2 *>(p + i - nx) : make_double2(bv.bottom, bv.bottom);
            const double2 pu2 = z + 1 < nz ? *reinterpret_cast<const double2 *>(p + i + nx) : make_double2(bv.top, bv.top);
            const double2 w2 = *reinterpret_cast<const double2 *>(width + i);
            const double tl = x0 > 0 ? v.tx[i - 1] : v.txl[z];
            const double pl = x0 > 0 ? p[i - 1] : bv.left, pr = x0 + 2 < nx ? p[i + 2] : bv.right;
            const CellOut c0 = flux_cell(fp, nz, nx, z, x0, i, pc2.x, pl, pc2.y, pd2.x, pu2.x, tl, tx2.x, td2.x, tu2.x, w2.x, a);
            const CellOut c1 = flux_cell(fp, nz, nx, z, x0 + 1, i + 1, pc2.y, pc2.x, pr, pd2.y, pu2.y, tx2.x, tx2.y, td2.y, tu2.y, w2.y, a);
            if (o.qx) *reinterpret_cast<double2 *>(o.qx + i) = make_double2(c0.qx, c1.qx);
            if (o.qz) *reinterpret_cast<double2 *>(o.qz + i) = make_double2(c0.qz, c1.qz);
            if (o.qm) *reinterpret_cast<double2 *>(o.qm + i) = make_double2(c0.qm, c1.qm);
            if (o.ren) *reinterpret_cast<double2 *>(o.ren + i) = make_double2(c0.ren, c1.ren);
            if (o.res) *reinterpret_cast<double2 *>(o.res + i) = make_double2(c0.a5, c1.a5);
        }
    } else {
        const Tiles tiles(v.zlo, v.zhi, nx);
        for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
            int z, x;
            tiles.cell(t, z, x);
            if (z >= v.zhi || x >= nx) continue;
            const size_t i = (size_t)z * nx + x;
            const double pc = p[i];
            const double pl = x > 0 ? p[i - 1] : bv.left;
            const double pr = x + 1 < nx ? p[i + 1] : bv.right;
            const double pd = z > 0 ? p[i - nx] : bv.bottom;
            const double pu = z + 1 < nz ? p[i + nx] : bv.top;
            const CellOut c = flux_cell(fp, nz, nx, z, x, i, pc, pl, pr, pd, pu, A.left(z, x), A.right(z, x), A.down(z, x),
                                        A.up(z, x), width[i], a);
            if (o.qx) o.qx[i] = c.qx;
            if (o.qz) o.qz[i] = c.qz;
            if (o.qm) o.qm[i] = c.qm;
            if (o.ren) o.ren[i] = c.ren;
            if (o.res) o.res[i] = c.a5;
        }
    }
    double acc[1] = {a.flow};
    double best = a.best, best_signed = a.best_signed;
    long long best_idx = a.best_idx;
    // block argmax (value, first index)
    __shared__ double sv[32], ss[32];
    __shared__ long long si[32];
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, off);
            const double os = __shfl_xor_sync(0xffffffffu, best_signed, off);
            const long long oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
            if (oi >= 0 && (ov > best || (ov == best && (oi < best_idx || best_idx < 0)))) { best = ov; best_signed = os; best_idx = oi; }
        }
        if (lane == 0) { sv[warp] = best; ss[warp] = best_signed; si[warp] = best_idx; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < (int)(blockDim.x >> 5); k++)
                if (si[k] >= 0 && (sv[k] > best || (sv[k] == best && (si[k] < best_idx || best_idx < 0)))) {
                    best = sv[k]; best_signed = ss[k]; best_idx = si[k];
                }
            partials[1 * kMaxPartials + blockIdx.x] = best;
            partials[2 * kMaxPartials + blockIdx.x] = best_signed;
            partials[3 * kMaxPartials + blockIdx.x] = (double)best_idx;
        }
        __syncthreads();
    }
    grid_sum_finalize<1>(acc, partials, counter, [=](const double(&tot)[1]) {
        double bv_ = -1.0, bs_ = 0.0, bi_ = -1.0;
        for (int b = 0; b < (int)gridDim.x; b++) {
            const double cv = ldcg(&partials[1 * kMaxPartials + b]);
            const double ci = ldcg(&partials[3 * kMaxPartials + b]);
            if (ci >= 0.0 && (cv > bv_ || (cv == bv_ && (ci < bi_ || bi_ < 0.0)))) {
                bv_ = cv; bs_ = ldcg(&partials[2 * kMaxPartials + b]); bi_ = ci;
            }
        }
        // strips: outlet flow summed, MAXLOC over the ranks' candidates (largest value, first cell), in rank order
        const double mine[4] = {tot[0], bv_, bs_, bi_};
        double all[kMaxRanks][4];
        dist_gather<4>(dc, mine, all);
        double flow = 0.0;
        bv_ = -1.0; bs_ = 0.0; bi_ = -1.0;
        for (int r = 0; r < dc.world; r++) {
            flow += all[r][0];
            const double cv = all[r][1], ci = all[r][3];
            if (ci >= 0.0 && (cv > bv_ || (cv == bv_ && (ci < bi_ || bi_ < 0.0)))) { bv_ = cv; bs_ = all[r][2]; bi_ = ci; }
        }
        scal[SC_FLOW] = flow;
        scal[SC_MAXRES_ABS] = bv_;
        scal[SC_MAXRES_SIGNED] = bs_;
    });
}

__global__ void __launch_bounds__(256)
k_normalise(double *__restrict__ res, size_t n, const double *scal)
{
    const double m = scal[SC_MAXRES_ABS];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        res[i] = res[i] / m;   // :826 (0/0 = NaN like the reference when every residual is 0)
}

__global__ void __launch_bounds__(256)
k_affine(double *__restrict__ x, size_t n, double scale, double shift)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = shift + scale * x[i];
}

void launch_flux(Solver &s, const apm_lcl_params &p, const double *pres, const BoundaryValues &bv, const FluxOut &o)
{
    FluxParams fp;
    fp.rho_over_mu = p.density / p.viscosity;
    fp.side = p.outlet_side;
    // pairs of cells as 128-bit words need an even row length and every array on a 16-byte boundary (a caller's own
    // device buffers need not be)
    bool vec = !(s.nx & 1);
    for (const void *q : {(const void *)pres, (const void *)o.qx, (const void *)o.qz, (const void *)o.qm, (const void *)o.ren,
                          (const void *)o.res})
        if ((uintptr_t)q & 15) vec = false;
    const int g0 = grid_for(s, s.zhi - s.zlo, vec ? s.nx / 2 : s.nx, 8);
    if (vec)
        LAUNCH(k_flux<true>, g0, kBlock, 0, s.stream, s.view5(), pres, (const double *)s.width, bv, fp, o,
               s.red + 10 * kMaxPartials, s.red_counters + 3, s.scal, s.dist);
    else
        LAUNCH(k_flux<false>, g0, kBlock, 0, s.stream, s.view5(), pres, (const double *)s.width, bv, fp, o,
               s.red + 10 * kMaxPartials, s.red_counters + 3, s.scal, s.dist);
    if (o.res) LAUNCH(k_normalise, s.num_sms * 8, 256, 0, s.stream, s.own(o.res), s.own_cells(), (const double *)s.scal);
}

void launch_affine(Solver &s, double *x, size_t n, double scale, double shift)
{
    LAUNCH(k_affine, s.num_sms * 8, 256, 0, s.stream, x, n, scale, shift);
}

// forces the kernels of this file into the device's context now (see preload_kernels, api.cu)
void preload_flux()
{
    preload_one(k_flux<true>);
    preload_one(k_flux<false>);
    preload_one(k_normalise);
    preload_one(k_affine);
}

}   // namespace lcl
