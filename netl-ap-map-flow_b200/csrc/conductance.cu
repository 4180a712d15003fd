// conductance.cu -- aperture map -> WIDTH and face conductances, one fused pass.
//
// Replaces, for the same inputs and with the same arithmetic order:
//   READ_AP_MAP numeric core   apmapflow/src/APERTURE_MAP_FLOW.F:472-501  (counters, clamps, WIDTH)
//   CREATE_POINT_MAP           apmapflow/src/MAP_MODULE.F:91-176          (corner values, recomputed
//                                                                          in registers, never stored)
//   COEF                       apmapflow/src/APM_MODULE.F:142-271         (tapered-plate half cells)
//   FLOW_BOUNDARY              apmapflow/src/APM_MODULE.F:277-376         (2x / 0 boundary faces)
// Roofline: 8 B read + 24 B written per cell, but 4 atan and half a dozen divisions in FP64 per cell make it
// FP64-pipe bound, not HBM bound.  It runs once per solve.
#include "device_utils.cuh"

namespace lcl {

struct CondParams {
    double high, low, rough, vox, visc, dx, dz;
    // loop constants of the interior faces, so that a face costs three divisions instead of seven: 1 / (d/2) for the
    // taper angle, and (DZ/12) / AVG_VISC / (DX/2) resp. (DX/12) / AVG_VISC / (DZ/2) for the two half cells in series
    // (APM_MODULE.F:213,235,266-267).  The products differ from the reference's chain of divisions in the last ulp or
    // two, like the merged division in tapered_b3; the parity bar on conductances is 1e-9.
    double inv_hx, inv_hz, kx, kz;
    int side;
};

__device__ __forceinline__ double clampa(double a, double high, double low)
{
    if (a > high) a = high;   // APERTURE_MAP_FLOW.F:478
    if (a < low) a = low;     // :479
    return a;
}

// One half-cell Stokes tapered-plate term (APM_MODULE.F:200-205).  a,b in the order the source
// multiplies them; INT(1E9*x)==0 restated as |1e9 x| < 1 (no trap on overflow).
// tan(theta) is the argument of the atan again: the reference evaluates TAN(ATAN(u)), which is u up to
// the last ulp or two of the two libm calls -- the same uncertainty any other libm has, and what the
// 1e-9 parity bar on conductances already allows for -- so the tangent (half of this kernel's FP64
// work) is not recomputed.
__device__ __forceinline__ double tapered_b3(double a, double b, double inv_dhalf)
{
    const double t = fabs(a - b) * inv_dhalf;
    const double theta = atan(t);
    const double x = t - theta;
    const double num = (2.0 * (a * a)) * (b * b);
    if (fabs(1e9 * x) < 1.0) return num / (a + b);
    // W3 * 3 x / t^3 with W3 = num / (a + b): one division instead of two (FP64 divisions are what this
    // kernel spends its time on; the result moves by an ulp, the parity bar on conductances is 1e-9)
    return ((num * 3.0) * x) / ((a + b) * ((t * t) * t));
}

// two half-cell terms in series across a face, times the face's constant k = (width/12) / visc / d:
// k / (d/B1 + d/B2) d = k B1 B2 / (B1 + B2), one division
__device__ __forceinline__ double series(double b1, double b2, double k) { return ((b1 * b2) * k) / (b1 + b2); }

struct ApView {
    const double *a;
    int nz, nx;
    double high, low;
    __device__ __forceinline__ double at(int z, int x) const { return clampa(a[(size_t)z * nx + x], high, low); }
    // corner whose lower-left cell is (cz,cx); cz in [-1,nz-1], cx in [-1,nx-1]  (MAP_MODULE.F:126-173)
    __device__ __forceinline__ double corner(int cz, int cx) const
    {
        const bool zl = cz < 0, zh = cz >= nz - 1, xl = cx < 0, xh = cx >= nx - 1;
        if ((zl || zh) && (xl || xh)) return at(zl ? 0 : nz - 1, xl ? 0 : nx - 1);   // map corners
        if (xl || xh) {                                                             // left/right edge
            const int x = xl ? 0 : nx - 1;
            return (at(cz, x) + at(cz + 1, x)) / 2.0;
        }
        if (zl || zh) {                                                             // bottom/top edge
            const int z = zl ? 0 : nz - 1;
            return (at(z, cx) + at(z, cx + 1)) / 2.0;
        }
        return (((at(cz, cx) + at(cz + 1, cx)) + at(cz, cx + 1)) + at(cz + 1, cx + 1)) / 4.0;   // :134-138
    }
};

__device__ __forceinline__ double width_of(double a_clamped, const CondParams &p)
{
    double w = a_clamped - p.rough;   // APERTURE_MAP_FLOW.F:495
    if (w < p.low) w = p.low;         // :498
    return w * p.vox;                 // :500
}

__global__ void __launch_bounds__(kBlock)
k_conductance(const double *__restrict__ ap, int nz, int nx, int zlo, int zhi, CondParams p, double *__restrict__ width,
              double *__restrict__ tx, double *__restrict__ tz, double *__restrict__ txl,
              double *__restrict__ tzb, double *partials, unsigned int *counter, double *scal, Dist dc)
{
    // a strip [zlo, zhi) reads the aperture rows zlo-1 .. zhi (corner points), global row numbers throughout
    const ApView A{ap, nz, nx, p.high, p.low};
    const Tiles tiles(zlo, zhi, nx);
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // raw sum, clamped sum, -min_pos(as max of neg), max, zeros, masked
    double vmin = 1e300, vmax = -1e300;
    for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
        int z, x;
        tiles.cell(t, z, x);
        if (z >= zhi || x >= nx) continue;
        const size_t i = (size_t)z * nx + x;
        const double raw = ap[i];
        const double a = clampa(raw, p.high, p.low);
        acc[0] += raw;
        acc[1] += a;
        if (raw > 0.0 && raw < vmin) vmin = raw;
        if (raw > vmax) vmax = raw;
        if (fabs(raw * 1e9) < 1.0) acc[4] += 1.0;                    // ZEROS, :473
        acc[5] += (raw > p.high ? 1.0 : 0.0) + (raw < p.low ? 1.0 : 0.0);   // MASK, :476

        const double w = width_of(a, p);
        width[i] = w;
        const double trc = A.corner(z, x), brc = A.corner(z - 1, x), tlc = A.corner(z, x - 1);
        double rht = 0.5 * (brc + trc) - p.rough;   // APM_MODULE.F:182,186
        double top = 0.5 * (tlc + trc) - p.rough;   // :183,187
        if (rht < p.low) rht = p.low;
        if (top < p.low) top = p.low;
        rht *= p.vox;
        top *= p.vox;
        const bool lr_outlet = (p.side == APM_LCL_LEFT || p.side == APM_LCL_RIGHT);

        // right face
        double fx;
        {
            const double bp3 = tapered_b3(w, rht, p.inv_hx);
            if (x < nx - 1) {
                const double wn = width_of(A.at(z, x + 1), p);
                const double bf3 = tapered_b3(rht, wn, p.inv_hx);
                fx = series(bp3, bf3, p.kx);                                   // :213,266
                if (x == 0) txl[z] = lr_outlet ? 2 * fx : 0.0;                 // :309,324 / :339,354
            } else {
                fx = (p.dz * bp3) / (12.0 * p.dx) / p.visc;                    // :249
                if (nx == 1) txl[z] = lr_outlet ? 2 * fx : 0.0;
                fx = lr_outlet ? 2 * fx : 0.0;                                 // :310 / :340
            }
        }
        tx[i] = fx;
        // top face
        double fz;
        {
            const double bp3 = tapered_b3(w, top, p.inv_hz);
            if (z < nz - 1) {
                const double wn = width_of(A.at(z + 1, x), p);
                const double bf3 = tapered_b3(top, wn, p.inv_hz);
                fz = series(bp3, bf3, p.kz);                                   // :235,267
                if (z == 0) tzb[x] = lr_outlet ? 0.0 : 2 * fz;                 // :311 / :341
            } else {
                fz = (p.dz * bp3) / (12.0 * p.dx) / p.visc;                    // :262 (sic DZ/DX)
                if (nz == 1) tzb[x] = lr_outlet ? 0.0 : 2 * fz;
                fz = lr_outlet ? 0.0 : 2 * fz;                                 // :312 / :342
            }
        }
        tz[i] = fz;
    }
    // min / max ride along as separate block reductions through the same partial slots
    __shared__ double smin[32], smax[32];
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int o = 16; o > 0; o >>= 1) {
            vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        if (lane == 0) { smin[warp] = vmin; smax[warp] = vmax; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < (int)(blockDim.x >> 5); k++) { vmin = fmin(vmin, smin[k]); vmax = fmax(vmax, smax[k]); }
            partials[6 * kMaxPartials + blockIdx.x] = vmin;
            partials[7 * kMaxPartials + blockIdx.x] = vmax;
        }
        __syncthreads();
    }
    grid_sum_finalize<6>(acc, partials, counter, [=](const double(&tot)[6]) {
        double mn = 1e300, mx = -1e300;
        for (int b = 0; b < (int)gridDim.x; b++) {
            mn = fmin(mn, ldcg(&partials[6 * kMaxPartials + b]));
            mx = fmax(mx, ldcg(&partials[7 * kMaxPartials + b]));
        }
        // the strips' sums, counters and extrema in rank order: every rank ends up with the same scalars
        const double mine[8] = {tot[0], tot[1], tot[2], tot[3], tot[4], tot[5], mn, mx};
        double all[kMaxRanks][8];
        dist_gather<8>(dc, mine, all);
        double sum[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        mn = 1e300; mx = -1e300;
        for (int r = 0; r < dc.world; r++) {
            for (int k = 0; k < 6; k++) sum[k] += all[r][k];
            mn = fmin(mn, all[r][6]);
            mx = fmax(mx, all[r][7]);
        }
        scal[SC_RAW_SUM] = sum[0];
        scal[SC_CLAMP_SUM] = sum[1];
        scal[SC_MIN_POS] = mn == 1e300 ? 0.0 : mn;
        scal[SC_MAX] = mx;
        scal[SC_ZEROS] = sum[4];
        scal[SC_MASKED] = sum[5];
    });
}

void launch_conductance(Solver &s, const apm_lcl_params &p, const double *ap_dev)
{
    CondParams c;
    c.high = p.high_mask; c.low = p.low_mask; c.rough = p.roughness; c.vox = p.voxel_size; c.visc = p.viscosity;
    // DIAM = NX*VOX*AVG, DX = DIAM/NX (APERTURE_MAP_FLOW.F:462-463, APM_MODULE.F:173-174)
    const double diam = s.nx * p.voxel_size * p.avg_factor, leng = s.nz * p.voxel_size * p.avg_factor;
    c.dx = diam / s.nx; c.dz = leng / s.nz;
    c.inv_hx = 1.0 / (c.dx / 2); c.inv_hz = 1.0 / (c.dz / 2);
    c.kx = (c.dz / 12) / c.visc / (c.dx / 2); c.kz = (c.dx / 12) / c.visc / (c.dz / 2);
    c.side = p.outlet_side;
    const int grid = grid_for(s, s.zhi - s.zlo, s.nx, 8);
    LAUNCH(k_conductance, grid, kBlock, 0, s.stream, ap_dev, s.nz, s.nx, s.zlo, s.zhi, c, s.width, s.tx, s.tz, s.txl, s.tzb,
           s.red, s.red_counters, s.scal, s.dist);
    if (s.levels[0].striped) {   // the faces are read by every fine-level kernel up to 5 rows beyond a strip
        HaloSet hs;
        hs.add(s.tx);
        hs.add(s.tz);
        halo_exchange(s, hs, s.nx, s.levels[0].cut, kHalo, kHalo, nullptr);
        HaloSet h1;
        h1.add(s.txl);
        halo_exchange(s, h1, 1, s.levels[0].cut, kHalo, kHalo, nullptr);
    }
}

// TRX(1:NZ,0:NX) / TRZ(0:NZ,1:NX) in the reference's own layout, for the parity tests (single-GPU solvers)
__global__ void k_export_layout(Fine5 v, double *__restrict__ trx, double *__restrict__ trz)
{
    const F5 A{v};
    const int nz = v.nz, nx = v.nx;
    const long long total = (long long)nz * (nx + 1) + (long long)(nz + 1) * nx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (i < (long long)nz * (nx + 1)) {
            if (!trx) continue;
            const int z = (int)(i / (nx + 1)), f = (int)(i % (nx + 1));
            trx[i] = f == 0 ? v.txl[z] : v.tx[(size_t)z * nx + f - 1];
        } else {
            if (!trz) continue;
            const long long j = i - (long long)nz * (nx + 1);
            const int f = (int)(j / nx), x = (int)(j % nx);
            trz[j] = f == 0 ? v.tzb[x] : v.tz[(size_t)(f - 1) * nx + x];
        }
    }
    (void)A;
}

void export_reference_layout(Solver &s, double *trx_dev, double *trz_dev)
{
    LAUNCH(k_export_layout, s.num_sms * 4, 256, 0, s.stream, s.view5(), trx_dev, trz_dev);
}

// forces the kernels of this file into the device's context now (see preload_kernels, api.cu)
void preload_conductance()
{
    preload_one(k_conductance);
    preload_one(k_export_layout);
}

}   // namespace lcl
