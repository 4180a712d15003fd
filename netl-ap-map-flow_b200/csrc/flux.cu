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

__global__ void __launch_bounds__(kBlock)
k_flux(Fine5 v, const double *__restrict__ p, const double *__restrict__ width, BoundaryValues bv, FluxParams fp,
       FluxOut o, double *partials, unsigned int *counter, double *scal, Dist dc)
{
    const F5 A{v};
    const int nz = v.nz, nx = v.nx;
    const Tiles tiles(v.zlo, v.zhi, nx);   // a strip's own rows; p has a ghost row either side
    double acc[1] = {0.0};            // outlet FLOW
    double best = -1.0, best_signed = 0.0;
    long long best_idx = -1;          // MAXLOC: first maximum in cell order
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
        const double q1 = -A.left(z, x) * (pc - pl);    // :756
        const double q2 = -A.right(z, x) * (pr - pc);   // :757
        const double q3 = -A.down(z, x) * (pc - pd);    // :758
        const double q4 = -A.up(z, x) * (pu - pc);      // :759
        const double q5 = ((q1 - q2) + (q3 - q4));      // :760
        const double qx = fp.side == APM_LCL_LEFT ? q1 : q2;     // :829-835
        const double qz = fp.side == APM_LCL_BOTTOM ? q3 : q4;
        const double qm = sqrt(qx * qx + qz * qz);               // :838
        if (o.qx) o.qx[i] = qx;
        if (o.qz) o.qz[i] = qz;
        if (o.qm) o.qm[i] = qm;
        if (o.ren) o.ren[i] = fp.rho_over_mu * (qm / width[i]);  // :839
        const double a5 = fabs(q5);
        if (o.res) o.res[i] = a5;
        if (a5 > best || (a5 == best && (long long)i < best_idx)) { best = a5; best_signed = q5; best_idx = (long long)i; }
        if (fp.side == APM_LCL_LEFT && x == 0) acc[0] += q1;      // :853-856
        if (fp.side == APM_LCL_RIGHT && x == nx - 1) acc[0] += q2;
        if (fp.side == APM_LCL_TOP && z == nz - 1) acc[0] += q3;
        if (fp.side == APM_LCL_BOTTOM && z == 0) acc[0] += q4;
    }
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
    const int g0 = grid_for(s, s.zhi - s.zlo, s.nx, 8);
    LAUNCH(k_flux, g0, kBlock, 0, s.stream, s.view5(), pres, (const double *)s.width, bv, fp, o,
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
    preload_one(k_flux);
    preload_one(k_normalise);
    preload_one(k_affine);
}

}   // namespace lcl
