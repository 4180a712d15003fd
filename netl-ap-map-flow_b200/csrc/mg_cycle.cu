// mg_cycle.cu -- one symmetric multigrid V-cycle, z = M^-1 r, the preconditioner of the PCG solve
// that replaces the reference's direct D4 elimination (D4_SOLVER_MODULE.F:444-617).
//
// Fine level: exact red/black Gauss-Seidel on the matrix-free 5-point operator (faces tx/tz); the
// whole down leg (smooth, residual, restriction) is one launch, k_down5, and the whole up leg
// (prolongation, smooth, r.z) another, k_up5.
// Coarse levels: four-colour Gauss-Seidel on the stored symmetric 9-point Galerkin operators, all
// colours and sweeps of a level in one tile kernel (k_smooth9).
// Pre-smoothing runs the colours forward, post-smoothing backward, so M is symmetric (CG needs it;
// tests/test_sim_kernels.py and tests/test_gpu_parity.py check <M a, b> == <a, M b>).
// Algorithmic bytes per launch are stated at each kernel and in DESIGN.md section 5.
#include "device_utils.cuh"

#include <functional>

namespace lcl {

// First statement of every kernel of the iteration: wait for the predecessor (a no-op unless the kernel was
// launched as a programmatic dependent, LAUNCH_DEP), then the device-side stop flag.
__device__ __forceinline__ bool stopped(const PcgState *st)
{
#ifndef LCL_CPU_SIM
    cudaGridDependencySynchronize();
#endif
    return st && (st->converged | st->failed);
}

// -------------------------------------------------------------------------------- coarse levels
__global__ void __launch_bounds__(kBlock)
k_residual9(Coarse9 v, const double *__restrict__ b, const double *__restrict__ x, double *__restrict__ r,
            int z0, int z1, const PcgState *st)
{
    // rows [z0, z1): a strip also computes the row below its own (the restriction of its first coarse row
    // reads it), from the ghost rows of x and of the operator
    if (stopped(st)) return;
    const C9 A{v};
    const int nz = v.nz, nx = v.nx;
    const Tiles tiles(z0, z1, nx);
    for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
        int z, xx;
        tiles.cell(t, z, xx);
        if (z >= z1 || xx >= nx) continue;
        const size_t i = (size_t)z * nx + xx;
        const double xi = x[i];
        double s = b[i] - A.S(z, xx) * xi;
#pragma unroll
        for (int dz = -1; dz <= 1; dz++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                if (dz == 0 && dx == 0) continue;
                const int qz = z + dz, qx = xx + dx;
                if (qz < 0 || qz >= nz || qx < 0 || qx >= nx) continue;
                s -= A.g(z, xx, dz, dx) * (xi - x[(size_t)qz * nx + qx]);
            }
        r[i] = s;
    }
}

// ------------------------------------------------------------------- fused tile smoothers
// value the interpolation P e gives at fine point (z,x)
__device__ __forceinline__ double prolong_value(const Weights &w, const double *__restrict__ e, int nzc, int nxc,
                                                int z, int x)
{
    const int bz = z >> 1, bx = x >> 1;
    const size_t b = (size_t)bz * nxc + bx;
    const bool Xp = bx + 1 < nxc, Zp = bz + 1 < nzc;
    const double e00 = e[b];
    if (!(z & 1)) {
        if (!(x & 1)) return e00;
        return w.hw[b] * e00 + (Xp ? w.he[b] * e[b + 1] : 0.0);
    }
    if (!(x & 1)) return w.vs[b] * e00 + (Zp ? w.vn[b] * e[b + nxc] : 0.0);
    return (w.csw[b] * e00 + (Xp ? w.cse[b] * e[b + 1] : 0.0)) +
           ((Zp ? w.cnw[b] * e[b + nxc] : 0.0) + ((Zp && Xp) ? w.cne[b] * e[b + nxc + 1] : 0.0));
}

constexpr int kSmTZ = 16, kSmTX = 64;   // smoother tile: 1024 points, 256 threads
constexpr int kSmThreads = 256;         // compile-time so the per-tile loops unroll and batch their loads

// Everything one cell of a 9-point level needs for a Gauss-Seidel update, held in registers for the
// whole launch: the 8 couplings (W, E, S, N, SW, SE, NW, NE), the right-hand side and 1/diag.  Cells
// outside the grid get all zeros, so their "update" is 0 and the sweeps need no bounds checks.
struct Cell9 {
    double gw, ge, gs, gn, gsw, gse, gnw, gne, b, inv;
};
// Shared-memory layout of a framed row of x (and of the staged operator rows): the even columns first,
// then the odd ones.  A thread owns a 2x2 block, so the lanes of a warp touch columns of ONE parity at
// stride 2: split by parity they are consecutive doubles -- conflict-free, half the wavefronts of the
// plain row (the sweeps were shared-memory bound: 72 % l1tex, 57 M wavefronts per level-1 launch).
template <int W>
__device__ __forceinline__ int split_pos(int row, int col) { return row * W + (col >> 1) + (col & 1) * (W / 2); }

// One Gauss-Seidel update of a cell from its eight neighbour values, in the fixed order (W E) (S N) (SW SE) (NW NE).
// A thread owns a 2x2 block and keeps its four iterates in registers: three of the eight neighbours of each of
// its cells are its own (register operands), five come from shared memory -- 20 loads per sweep of the block
// instead of 32 (the sweeps are bound by the shared-memory pipe); every new value is published to shared memory
// for the neighbouring threads.
__device__ __forceinline__ double relax9v(const Cell9 &c, double w, double e, double s_, double n, double sw, double se,
                                          double nw, double ne)
{
    double s = c.b;
    s += c.gw * w + c.ge * e;
    s += c.gs * s_ + c.gn * n;
    s += c.gsw * sw + c.gse * se;
    s += c.gnw * nw + c.gne * ne;
    return s * c.inv;
}
// pe0: the block's (row r, even column) cell in the split layout, po0 = pe0 + W/2 its odd-column cell; the row
// above is + W.  West of an even column: the odd plane one to the left (po - 1); east of an odd column: pe + 1.
template <int W>
__device__ __forceinline__ void relax_c00(const Cell9 &c, double *pe0, const double *po0, double &x00, double x01, double x10, double x11)
{
    x00 = relax9v(c, po0[-1], x01, pe0[-W], x10, po0[-W - 1], po0[-W], po0[W - 1], x11);
    pe0[0] = x00;
}
template <int W>
__device__ __forceinline__ void relax_c01(const Cell9 &c, const double *pe0, double *po0, double x00, double &x01, double x10, double x11)
{
    x01 = relax9v(c, x00, pe0[1], po0[-W], x11, pe0[-W], pe0[-W + 1], x10, pe0[W + 1]);
    po0[0] = x01;
}
template <int W>
__device__ __forceinline__ void relax_c10(const Cell9 &c, double *pe0, const double *po0, double x00, double x01, double &x10, double x11)
{
    x10 = relax9v(c, po0[W - 1], x11, x00, pe0[2 * W], po0[-1], x01, po0[2 * W - 1], po0[2 * W]);
    pe0[W] = x10;
}
template <int W>
__device__ __forceinline__ void relax_c11(const Cell9 &c, const double *pe0, double *po0, double x00, double x01, double x10, double &x11)
{
    x11 = relax9v(c, x10, pe0[W + 1], x01, po0[2 * W], x00, pe0[1], pe0[2 * W], pe0[2 * W + 1]);
    po0[W] = x11;
}

// Operator staging of the 9-point smoother: the rows a 16x64 tile needs -- E, N, NE, NW of the tile plus
// the row below and one column either side (rows z0-1..z0+15, columns x0-1..x0+64), b and C of the tile
// itself -- copied asynchronously into shared memory, zero-filled outside the grid.
constexpr int kStRows = kSmTZ + 1, kStCols = kSmTX + 2, kStPlane = kStRows * kStCols, kStTile = kSmTZ * kSmTX;
constexpr int kSmW = kSmTX + 4, kSmH = kSmTZ + 4;   // x with a 2-cell frame: the region stays aligned to coarse blocks
constexpr size_t kSm9Smem = (size_t)(4 * kStPlane + 2 * kStTile + kSmH * kSmW) * sizeof(double);
static_assert(2 * (kSm9Smem + 1024) <= 227 * 1024, "two blocks per SM");

__device__ __forceinline__ void stage_tile9(double *st, const Coarse9 &v, const double *__restrict__ b, int z0, int x0)
{
    // rows a rank may read: its strip plus ghost rows (the whole grid on one GPU); rows it owns: [zlo, zhi)
    const int nx = v.nx, rlo = read_lo(v.zlo), rhi = read_hi(v.zhi, v.nz), zhi = v.zhi;
    double *sE = st, *sN = st + kStPlane, *sNE = st + 2 * kStPlane, *sNW = st + 3 * kStPlane;
    double *sb = st + 4 * kStPlane, *sC = sb + kStTile;
#pragma unroll 2
    for (int k = threadIdx.x; k < kStPlane; k += kSmThreads) {
        const int r = k / kStCols, cc = k % kStCols;
        const int z = z0 - 1 + r, x = x0 - 1 + cc, q = split_pos<kStCols>(r, cc);
        const bool in = z >= rlo && z < rhi && x >= 0 && x < nx;
        const size_t i = in ? (size_t)z * nx + x : 0;
        cp_async8(&sE[q], v.E + i, in);
        cp_async8(&sN[q], v.N + i, in);
        cp_async8(&sNE[q], v.NE + i, in);
        cp_async8(&sNW[q], v.NW + i, in);
    }
#pragma unroll
    for (int k = threadIdx.x; k < kStTile; k += kSmThreads) {
        const int lz = k / kSmTX, lx = k % kSmTX;
        const int z = z0 + lz, x = x0 + lx, q = split_pos<kSmTX>(lz, lx);
        const bool in = z < zhi && x < nx;
        const size_t i = in ? (size_t)z * nx + x : 0;
        cp_async8(&sb[q], b + i, in);
        cp_async8(&sC[q], v.C + i, in);
    }
}
// the staged operator row of tile cell (lz, lx): every coupling a neighbour outside the grid would carry is
// either stored as 0 (E, N, NE, NW where the neighbour is absent) or zero-filled by the staging
__device__ __forceinline__ Cell9 staged_cell9(const double *st, int lz, int lx)
{
    const double *sE = st, *sN = st + kStPlane, *sNE = st + 2 * kStPlane, *sNW = st + 3 * kStPlane;
    const double *sb = st + 4 * kStPlane, *sC = sb + kStTile;
    const int r = lz + 1, cc = lx + 1;
    const int k = split_pos<kStCols>(r, cc), kw = split_pos<kStCols>(r, cc - 1), ke = split_pos<kStCols>(r, cc + 1);
    Cell9 c;
    c.ge = sE[k];
    c.gn = sN[k];
    c.gne = sNE[k];
    c.gnw = sNW[k];
    c.gw = sE[kw];
    c.gs = sN[k - kStCols];
    c.gsw = sNE[kw - kStCols];
    c.gse = sNW[ke - kStCols];
    c.b = sb[split_pos<kSmTX>(lz, lx)];
    const double d = sC[split_pos<kSmTX>(lz, lx)];
    c.inv = d != 0.0 ? 1.0 / d : 0.0;          // cells outside the grid: everything 0, their "update" is 0
    return c;
}

// Second stage, up leg only: what x + P e of the framed tile needs -- x over the framed region, the coarse
// values e of its (H/2+1) x (W/2+1) blocks and the eight weight planes of its H/2 x W/2 blocks.
constexpr int kPrBH = kSmH / 2, kPrBW = kSmW / 2, kPrE = (kPrBH + 1) * (kPrBW + 1), kPrB = kPrBH * kPrBW;
constexpr int kPrStage = kSmH * kSmW + kPrE + 8 * kPrB;
constexpr size_t kSm9SmemUp = kSm9Smem + (size_t)(kPrStage + (kPrStage & 1)) * sizeof(double);
static_assert(2 * (kSm9SmemUp + 1024) <= 227 * 1024, "two blocks per SM");

// rlo, rhi: rows of xin this rank holds; wlo, whi: coarse rows of the weights and of e it holds
__device__ __forceinline__ void stage_prolong9(double *st, const double *__restrict__ xin, const Weights &w,
                                               const double *__restrict__ e, int wlo, int whi, int nxc, int z0, int x0,
                                               int rlo, int rhi, int nx)
{
    double *sXin = st, *sE = st + kSmH * kSmW, *sW = sE + kPrE;
    const int zr = z0 - 2, xr = x0 - 2, bz0 = zr >> 1, bx0 = xr >> 1;   // arithmetic shift: -2 >> 1 == -1
#pragma unroll 2
    for (int k = threadIdx.x; k < kSmH * kSmW; k += kSmThreads) {
        const int z = zr + k / kSmW, x = xr + k % kSmW;
        const bool in = z >= rlo && z < rhi && x >= 0 && x < nx;
        cp_async8(&sXin[k], xin + (in ? (size_t)z * nx + x : 0), in);
    }
    for (int k = threadIdx.x; k < kPrE; k += kSmThreads) {
        const int bz = bz0 + k / (kPrBW + 1), bx = bx0 + k % (kPrBW + 1);
        const bool in = bz >= wlo && bz < whi && bx >= 0 && bx < nxc;
        cp_async8(&sE[k], e + (in ? (size_t)bz * nxc + bx : 0), in);
    }
#pragma unroll
    for (int k = threadIdx.x; k < kPrB; k += kSmThreads) {
        const int bz = bz0 + k / kPrBW, bx = bx0 + k % kPrBW;
        const bool in = bz >= wlo && bz < whi && bx >= 0 && bx < nxc;
        const size_t b = in ? (size_t)bz * nxc + bx : 0;
        cp_async8(&sW[k], w.hw + b, in);
        cp_async8(&sW[kPrB + k], w.he + b, in);
        cp_async8(&sW[2 * kPrB + k], w.vs + b, in);
        cp_async8(&sW[3 * kPrB + k], w.vn + b, in);
        cp_async8(&sW[4 * kPrB + k], w.csw + b, in);
        cp_async8(&sW[5 * kPrB + k], w.cse + b, in);
        cp_async8(&sW[6 * kPrB + k], w.cnw + b, in);
        cp_async8(&sW[7 * kPrB + k], w.cne + b, in);
    }
}
// sx := x + P e over the framed region from the staged operands (one thread per coarse 2x2 block)
// x + P e of the coarse block (bzl, bxl) of the framed region, from the staged operands into sx and into v
__device__ __forceinline__ void prolong_block9(double *sx, const double *st, int nzc, int nxc, int z0, int x0, int nz, int nx,
                                               int bzl, int bxl, double (&v)[4])
{
    constexpr int W = kSmW, BW = kPrBW, EW = kPrBW + 1;
    const double *sXin = st, *sE = st + kSmH * kSmW, *sW = sE + kPrE;
    const int k = bzl * BW + bxl;
    const int bz = ((z0 - 2) >> 1) + bzl, bx = ((x0 - 2) >> 1) + bxl;
    v[0] = v[1] = v[2] = v[3] = 0.0;
    if (bz >= 0 && bx >= 0 && bz < nzc && bx < nxc) {
        const bool xp = 2 * bx + 1 < nx, zp = 2 * bz + 1 < nz;
        const double *ep = sE + bzl * EW + bxl;
        const double e00 = ep[0], e01 = ep[1], e10 = ep[EW], e11 = ep[EW + 1];   // 0 beyond the coarse grid
        const double *xi = sXin + (2 * bzl) * W + 2 * bxl;
        v[0] = xi[0] + e00;
        if (xp) v[1] = xi[1] + (sW[k] * e00 + sW[kPrB + k] * e01);
        if (zp) v[2] = xi[W] + (sW[2 * kPrB + k] * e00 + sW[3 * kPrB + k] * e10);
        if (zp && xp)
            v[3] = xi[W + 1] + ((sW[4 * kPrB + k] * e00 + sW[5 * kPrB + k] * e01) + (sW[6 * kPrB + k] * e10 + sW[7 * kPrB + k] * e11));
    }
    double *p = &sx[(2 * bzl) * W + bxl];               // split rows: even column 2 bxl here, odd one W/2 further
    p[0] = v[0];
    p[W / 2] = v[1];
    p[W] = v[2];
    p[W + W / 2] = v[3];
}
// sx := x + P e over the framed region.  Thread (a, c) builds the block it will smooth -- framed block (a + 1, c + 1) --
// and keeps the four values (own); the 84 blocks of the frame ring are dealt to the first 84 threads.
__device__ __forceinline__ void prolong_staged9(double *sx, const double *st, int nzc, int nxc, int z0, int x0, int nz, int nx,
                                                double (&own)[4])
{
    constexpr int BH = kPrBH, BW = kPrBW;
    static_assert((BH - 2) * (BW - 2) == kSmThreads, "one tile block per thread");
    const int tid = threadIdx.x;
    prolong_block9(sx, st, nzc, nxc, z0, x0, nz, nx, (tid >> 5) + 1, (tid & 31) + 1, own);
    constexpr int kRing = BH * BW - (BH - 2) * (BW - 2);
    if (tid < kRing) {
        int bzl, bxl;
        if (tid < BW) { bzl = 0; bxl = tid; }
        else if (tid < 2 * BW) { bzl = BH - 1; bxl = tid - BW; }
        else { bzl = 1 + ((tid - 2 * BW) >> 1); bxl = ((tid - 2 * BW) & 1) ? BW - 1 : 0; }
        double ring[4];
        prolong_block9(sx, st, nzc, nxc, z0, x0, nz, nx, bzl, bxl, ring);
    }
}

// All four colours and all `nu` sweeps of one 9-point level in ONE launch.  Each block owns a
// 16x64 tile, keeps x (+ frame) in shared memory and runs four-colour Gauss-Seidel on it; halo
// values stay frozen during the launch (hybrid GS: Jacobi between tiles, GS inside -- a regular
// splitting, and the backward variant is its transpose, so the cycle stays symmetric).  A thread
// owns one 2x2 block (one cell of each colour) and holds its operator rows in registers; the sweeps
// touch only registers and shared memory.  The rows reach the registers through a shared-memory stage
// filled by asynchronous copies, and the copies for the block's NEXT tile are issued before the sweeps
// of the current one start: the DRAM latency of a tile hides behind the sweeps and the store of the
// tile before it (two resident blocks per SM could not hide it: 51 % of the copy peak before, see
// profiles/README.md for after).
// ZERO_INIT: x starts from 0 (down leg; nothing is read).  PROLONG: x starts from x + P e (up leg).
template <bool ZERO_INIT, bool PROLONG, bool FORWARD>
__global__ void __launch_bounds__(256, 2)
k_smooth9(Coarse9 v, const double *__restrict__ b, const double *__restrict__ xin, double *__restrict__ xout, int nu,
          Weights w, const double *__restrict__ e, int nzc, int nxc, int wlo, int whi, const PcgState *st)
{
    if (stopped(st)) return;
    const int nz = v.nz, nx = v.nx;
    // a strip walks the tiles of its own rows [zlo, zhi) (cut at multiples of the tile height, mg_plan) and
    // reads the rows [rlo, rhi) it holds: the frame of a tile at a cut is the neighbour's ghost rows
    const int zlo = v.zlo, zhi = v.zhi, rlo = read_lo(zlo), rhi = read_hi(zhi, nz);
    constexpr int W = kSmW, H = kSmH;
    LCL_DYNAMIC_SMEM(double, smem);
    double *stage = smem, *sx = smem + 4 * kStPlane + 2 * kStTile;   // sx at a 16-byte aligned offset
    double *stage2 = sx + kSmH * kSmW;                               // PROLONG only
    const int tiles_x = (nx + kSmTX - 1) / kSmTX, tiles_z = (zhi - zlo + kSmTZ - 1) / kSmTZ;
    const long long ntiles = (long long)tiles_x * tiles_z;
    const int tid = threadIdx.x;
    const int a = tid >> 5, c = tid & 31;
    if (blockIdx.x < ntiles) {
        const int z0 = zlo + (int)(blockIdx.x / tiles_x) * kSmTZ, x0 = (int)(blockIdx.x % tiles_x) * kSmTX;
        stage_tile9(stage, v, b, z0, x0);
        if (PROLONG) stage_prolong9(stage2, xin, w, e, wlo, whi, nxc, z0, x0, rlo, rhi, nx);
    }
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int z0 = zlo + (int)(t / tiles_x) * kSmTZ, x0 = (int)(t % tiles_x) * kSmTX;
        if (!PROLONG) {
            for (int k = tid; k < H * W; k += kSmThreads) {
                const int zf = z0 + k / W - 2, xf = x0 + k % W - 2;
                double val = 0.0;
                if (!ZERO_INIT && zf >= rlo && zf < rhi && xf >= 0 && xf < nx)
                    val = xin[(size_t)zf * nx + xf];     // never the array this launch writes: other tiles read it
                sx[ZERO_INIT ? k : split_pos<W>(k / W, k % W)] = val;
            }
        }
        cp_async_wait_all();
        __syncthreads();
        Cell9 c00 = staged_cell9(stage, 2 * a, 2 * c), c01 = staged_cell9(stage, 2 * a, 2 * c + 1);
        Cell9 c10 = staged_cell9(stage, 2 * a + 1, 2 * c), c11 = staged_cell9(stage, 2 * a + 1, 2 * c + 1);
        // the six couplings INSIDE the block are held by both of their cells (the same staged element read twice):
        // say so, and the compiler drops six loads and six doubles of registers
        c01.gw = c00.ge; c10.gs = c00.gn; c11.gsw = c00.gne;
        c11.gs = c01.gn; c10.gse = c01.gnw; c11.gw = c10.ge;
        double own[4] = {0.0, 0.0, 0.0, 0.0};
        if (PROLONG) prolong_staged9(sx, stage2, nzc, nxc, z0, x0, nz, nx, own);
        __syncthreads();                                 // every operand is in registers or sx: the stages are free again
        if (t + gridDim.x < ntiles) {
            const long long tn = t + gridDim.x;
            const int zn = zlo + (int)(tn / tiles_x) * kSmTZ, xn = (int)(tn % tiles_x) * kSmTX;
            stage_tile9(stage, v, b, zn, xn);
            if (PROLONG) stage_prolong9(stage2, xin, w, e, wlo, whi, nxc, zn, xn, rlo, rhi, nx);
        }
        // the thread's block: framed rows 2a+2, 2a+3, columns 2c+2 (even, half index c+1) and 2c+3 (odd)
        double *pe0 = &sx[(2 * a + 2) * W + c + 1];      // even column
        double *po0 = pe0 + W / 2;                       // odd column; west of the even one: po0 - 1, east of the odd one: pe0 + 1
        // the block's own iterates stay in registers (the up leg has just built them, the zero guess is zero)
        double x00, x01, x10, x11;
        if (PROLONG) { x00 = own[0]; x01 = own[1]; x10 = own[2]; x11 = own[3]; }
        else if (ZERO_INIT) { x00 = x01 = x10 = x11 = 0.0; }
        else { x00 = pe0[0]; x01 = po0[0]; x10 = pe0[W]; x11 = po0[W]; }
        int sweep = 0;
        if (ZERO_INIT && nu > 0) {
            // first forward sweep from x = 0: every value a colour has not reached yet is still 0 (the frame too), and
            // adding 0 * g changes nothing -- the loads of those neighbours are skipped (10 of the 20 of this sweep)
            x00 = c00.b * c00.inv;
            pe0[0] = x00;
            __syncthreads();
            x01 = relax9v(c01, x00, pe0[1], 0.0, 0.0, 0.0, 0.0, 0.0, 0.0);
            po0[0] = x01;
            __syncthreads();
            x10 = relax9v(c10, 0.0, 0.0, x00, pe0[2 * W], po0[-1], x01, po0[2 * W - 1], po0[2 * W]);
            pe0[W] = x10;
            __syncthreads();
            relax_c11<W>(c11, pe0, po0, x00, x01, x10, x11);
            __syncthreads();
            sweep = 1;
        }
        for (; sweep < nu; sweep++) {
            if (FORWARD) {
                relax_c00<W>(c00, pe0, po0, x00, x01, x10, x11);
                __syncthreads();
                relax_c01<W>(c01, pe0, po0, x00, x01, x10, x11);
                __syncthreads();
                relax_c10<W>(c10, pe0, po0, x00, x01, x10, x11);
                __syncthreads();
                relax_c11<W>(c11, pe0, po0, x00, x01, x10, x11);
                __syncthreads();
            } else {
                relax_c11<W>(c11, pe0, po0, x00, x01, x10, x11);
                __syncthreads();
                relax_c10<W>(c10, pe0, po0, x00, x01, x10, x11);
                __syncthreads();
                relax_c01<W>(c01, pe0, po0, x00, x01, x10, x11);
                __syncthreads();
                relax_c00<W>(c00, pe0, po0, x00, x01, x10, x11);
                __syncthreads();
            }
        }
        // the block's four cells go out from the registers (two 16-byte stores); every thread is past the last colour
        // barrier, so shared memory is free for the next tile
        {
            const int zf = z0 + 2 * a, xf = x0 + 2 * c;
            if (xf < nx) {
                double *o = xout + (size_t)zf * nx + xf;
                const bool pair = xf + 1 < nx, vec = pair && !(nx & 1);
                if (zf < zhi) {
                    if (vec) *reinterpret_cast<double2 *>(o) = make_double2(x00, x01);
                    else { o[0] = x00; if (pair) o[1] = x01; }
                }
                if (zf + 1 < zhi) {
                    if (vec) *reinterpret_cast<double2 *>(o + nx) = make_double2(x10, x11);
                    else { o[nx] = x10; if (pair) o[nx + 1] = x11; }
                }
            }
        }
    }
}

// What the down leg's reduction decides: the PCG residual norms and the stopping rule.
__device__ __forceinline__ void down_norms_finalize(PcgState *st, int init, const double (&part)[2], const Dist &dc)
{
    // this strip's sums and its largest pointwise change of the last update (k_update), combined in rank order
    const double mine[3] = {part[0], part[1], st->dx_part};
    double all[kMaxRanks][3];
    dist_gather<3>(dc, mine, all);
    double tot[2] = {0.0, 0.0}, dx = 0.0;
    for (int r = 0; r < dc.world; r++) {
        tot[0] += all[r][0];
        tot[1] += all[r][1];
        dx = fmax(dx, all[r][2]);
    }
    st->dx = dx;
    st->rr = tot[0];
    st->rs = tot[1];
    if (init) {
        st->bb = tot[0];
        st->bs = tot[1];
    } else if (tot[0] <= st->tol2 * st->bb && tot[1] <= st->tol2 * st->bs) {
        // Residual norms are met.  Keep iterating only while every update still shrinks the
        // largest pointwise pressure change: that is how weakly connected cells (which global
        // norms do not see) reach 1e-6.  If the change grows twice in a row (or after 12 such
        // extra iterations) what is left are floating pockets, so stop.  A change of more than
        // 2.5e-3 of the pressure scale while both norms are at rel_tol is a pocket too: its residual is
        // rounding noise, divided by a leak conductance 1e-14 of everything else.
        st->res_ok += 1;
        if ((st->dx <= st->dx_tol && st->dx_prev <= st->dx_tol) || (st->dx > st->dx_prev && st->dx_prev > st->dx_prev2) ||
            st->dx > 1e4 * st->dx_tol || st->res_ok > 12)
            st->converged = 1;
    }
    if (!(tot[0] == tot[0]) || !(tot[1] == tot[1])) st->failed = 1;
}

// Fine level, whole down leg in ONE launch: exact red/black sweep from a zero guess, the residual it
// leaves (non-zero at red cells only) and its restriction to the coarse right-hand side.
// A block stages b and the faces of a 42x58 tile plus frame -- 48 rows of 64 -- in shared memory with
// asynchronous copies (everything in flight at once, zero-filled outside the grid, so the passes need
// no neighbour logic and touch no global memory), then recomputes the frame instead of exchanging it:
//   red  (rows -3..TZ+1)  x = b/diag                    -> needs nothing else
//   black(rows -2..TZ)    x = (b + sum t x_red)/diag    -> exactly what the two-pass version computes
//   residual at red cells (rows -1..TZ-1), in place over b
//   coarse rhs of the 21x29 coarse points the tile owns = r(2Z,2X) + sum w r(2Z+-1,2X+-1)
// A warp walks rows, its lanes the cells of one colour in the row (32 per framed row): all index
// arithmetic is shifts and adds, the row tests are warp-uniform.
// Algorithmic traffic: b, tx, tz in, x out (32 B/cell) + 4 weights in, 1 rhs out per coarse point.
constexpr int kDnTZ = 42, kDnTX = 58, kDnFL = 4, kDnFH = 2, kDnThreads = 512, kDnWarps = kDnThreads / 32;
constexpr int kDnH = kDnTZ + kDnFL + kDnFH, kDnW = kDnTX + kDnFL + kDnFH;
constexpr size_t kDnSmem = 4 * (size_t)kDnH * kDnW * sizeof(double);
static_assert(kDnW == 64 && kDnH % kDnWarps == 0, "a framed row is two lane rounds, the framed tile whole warp rounds");
static_assert(kDnTX / 2 <= 32, "a row of coarse points per warp");
static_assert(2 * (kDnSmem + 2048) <= 228 * 1024, "two blocks per SM");

// 1 / d for the diagonals of the smoothers (d > 0, far from the ends of the exponent range): hardware seed
// (upper 32 bits, ~2^-23) and two Newton steps, 5 instructions without the range checks and the slow path of
// an IEEE division.  The last bit may differ from 1.0 / d; both legs of the cycle use this same function of the
// same sum, so the preconditioner stays symmetric.
__device__ __forceinline__ double rcp_diag(double d)
{
#ifdef LCL_CPU_SIM
    return 1.0 / d;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(fma(-d, r, 1.0), r, r);
    return fma(fma(-d, r, 1.0), r, r);
#endif
}

// Shared-memory layout of a framed row: the 32 even columns first, then the 32 odd ones, so the lanes
// of a warp walking one colour (stride 2 in x) touch consecutive doubles -- no bank conflicts -- and so
// do their left/right neighbours (the other parity) and the rows above and below.
__device__ __forceinline__ int dn_pos(int row, int col) { return row * kDnW + (col >> 1) + ((col & 1) << 5); }

// Stage b, the right face (tx; txl in column -1) and the upper face (tz; tzb in row -1) of every cell
// of a framed H x 64 tile whose frame is F cells on the low sides, asynchronously and zero-filled
// outside the grid.  Lanes 0-15 take the even columns of a 32-column chunk, lanes 16-31 the odd ones:
// one contiguous 256-byte read, two conflict-free 128-byte shared-memory writes.
// INT (interior tile: the whole framed region lies inside the rows this rank holds and inside the grid): no
// predicates, no zero fill, one address per row.
template <int F, int H, bool INT>
__device__ __forceinline__ void stage_tile5(double *sb, double *stx, double *stz, const Fine5 &v,
                                            const double *__restrict__ b, int z0, int x0, int wy, int lane)
{
    if (INT) {
        const int col0 = 2 * (lane & 15) + (lane >> 4);
        size_t i = (size_t)(z0 + wy - F) * v.nx + (x0 - F + col0);
        int k = dn_pos(wy, col0);                      // dn_pos(row, col0 + 32) = k + 16
#pragma unroll
        for (int row = wy; row < H; row += kDnWarps, i += (size_t)kDnWarps * v.nx, k += kDnWarps * kDnW) {
            cp_async8(&sb[k], b + i, true);
            cp_async8(&stx[k], v.tx + i, true);
            cp_async8(&stz[k], v.tz + i, true);
            cp_async8(&sb[k + 16], b + i + 32, true);
            cp_async8(&stx[k + 16], v.tx + i + 32, true);
            cp_async8(&stz[k + 16], v.tz + i + 32, true);
        }
        return;
    }
    // rows this rank holds (strip + ghost rows; the whole grid on one GPU).  Row -1 of the grid is the bottom
    // boundary face tzb; the row below a strip's ghost rows is not held and reads as 0 (it only feeds ghost
    // cells nobody uses)
    const int nx = v.nx, rlo = read_lo(v.zlo), rhi = read_hi(v.zhi, v.nz);
#pragma unroll
    for (int row = wy; row < H; row += kDnWarps) {
        const int z = z0 + row - F;
        const bool zin = z >= rlo && z < rhi, zface = zin || z == -1;
        const size_t zoff = zin ? (size_t)z * nx : 0;
        const double *brow = b + zoff, *txrow = v.tx + zoff, *tzrow = zin ? v.tz + zoff : v.tzb;
        const double *txl = v.txl + (zin ? z : 0);
#pragma unroll
        for (int c = 0; c < kDnW; c += 32) {
            const int col = c + 2 * (lane & 15) + (lane >> 4);
            const int xx = x0 + col - F, k = dn_pos(row, col);
            const bool xin = xx >= 0 && xx < nx;
            const int xs = xin ? xx : 0;
            cp_async8(&sb[k], brow + xs, zin && xin);
            cp_async8(&stx[k], xx >= 0 ? txrow + xs : txl, zin && xx >= -1 && xx < nx);
            cp_async8(&stz[k], tzrow + xs, zface && xin);
        }
    }
}

// One tile of the down leg.  INT: interior tile (every cell of the framed region exists and is held, the whole
// tile is owned, nx even): the passes run without a single bounds test -- 97 % of the tiles of a large map.
template <bool INT>
__device__ __forceinline__ void down5_tile(double *smem, const Fine5 &v, const Weights &w, const double *__restrict__ b,
                                           double *__restrict__ x, double *__restrict__ bc, int nzc, int nxc, int czhi, int z0,
                                           int x0, int rlo, int rhi, double (&acc)[2])
{
    constexpr int W = kDnW, H = kDnH, FL = kDnFL, TZ = kDnTZ, TX = kDnTX;
    double *sb = smem, *stx = smem + H * W, *stz = smem + 2 * H * W, *sx = smem + 3 * H * W;
    const int nz = v.nz, nx = v.nx, zhi = v.zhi;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    stage_tile5<FL, H, INT>(sb, stx, stz, v, b, z0, x0, wy, lane);
    // restriction weights of this thread's coarse points (rows wy, wy + 16 of the tile's 21 x 29) ride along
    constexpr int CR = (TZ / 2 + kDnWarps - 1) / kDnWarps;
    const int X = x0 / 2 + lane;
    double wne[CR], wnw[CR], wse[CR], wsw[CR];
    bool cin[CR];
#pragma unroll
    for (int c = 0; c < CR; c++) {
        const int cr = wy + c * kDnWarps, Z = z0 / 2 + cr;
        cin[c] = cr < TZ / 2 && lane < TX / 2 && (INT || (Z < czhi && X < nxc));
        wne[c] = wnw[c] = wse[c] = wsw[c] = 0.0;
        if (cin[c]) {
            const size_t bk = (size_t)Z * nxc + X;
            const bool xm = INT || X > 0, xp = INT || 2 * X + 1 < nx, zm = INT || Z > 0, zp = INT || 2 * Z + 1 < nz;
            if (zp && xp) wne[c] = w.csw[bk];
            if (zp && xm) wnw[c] = w.cse[bk - 1];
            if (zm && xp) wse[c] = w.cnw[bk - nxc];
            if (zm && xm) wsw[c] = w.cne[bk - nxc - 1];
        }
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- red cells ((lz + lx) even; z0, x0 are even), rows -3..TZ+1, columns -3..TX+1
#pragma unroll
    for (int row = wy; row < TZ + 5; row += kDnWarps) {
        const int lz = row - 3, lx = 2 * lane - 3 + ((lz & 1) ^ 1);
        if (lx > TX + 1) continue;
        const int z = z0 + lz, xx = x0 + lx, par = (lx + FL) & 1;
        const int p = dn_pos(lz + FL, lx + FL), pl = p + (par ? -32 : 31);
        double val = 0.0;
        if (INT || (z >= rlo && z < rhi && xx >= 0 && xx < nx)) {
            const double bi = sb[p];
            const double inv = rcp_diag((stx[pl] + stx[p]) + (stz[p - W] + stz[p]));
            val = bi * inv;
            if (lz >= 0 && lz < TZ && (INT || z < zhi) && lx >= 0 && lx < TX) { acc[0] += bi * bi; acc[1] += bi * bi * inv; }
        }
        sx[p] = val;
    }
    __syncthreads();
    // ---- black cells, rows -2..TZ, columns -2..TX (neighbours outside the grid hold 0)
#pragma unroll
    for (int row = wy; row < TZ + 3; row += kDnWarps) {
        const int lz = row - 2, lx = 2 * lane - 2 + ((lz & 1) ^ 1);
        if (lx > TX) continue;
        const int z = z0 + lz, xx = x0 + lx, par = (lx + FL) & 1;
        const int p = dn_pos(lz + FL, lx + FL), pl = p + (par ? -32 : 31), pr = p + (par ? -31 : 32);
        double val = 0.0;
        if (INT || (z >= rlo && z < rhi && xx >= 0 && xx < nx)) {
            const double tl = stx[pl], tr = stx[p], td = stz[p - W], tu = stz[p];
            const double bi = sb[p];
            double s = bi;
            s += tl * sx[pl];
            s += tr * sx[pr];
            s += td * sx[p - W];
            s += tu * sx[p + W];
            const double inv = rcp_diag((tl + tr) + (td + tu));
            val = s * inv;
            if (lz >= 0 && lz < TZ && (INT || z < zhi) && lx >= 0 && lx < TX) { acc[0] += bi * bi; acc[1] += bi * bi * inv; }
        }
        sx[p] = val;
    }
    __syncthreads();
    // ---- residual at red cells, rows -1..TZ-1, columns -1..TX-1, in place over b (it is 0 at black cells)
#pragma unroll
    for (int row = wy; row < TZ + 1; row += kDnWarps) {
        const int lz = row - 1, lx = 2 * lane - 1 + ((lz & 1) ^ 1);
        if (lx > TX - 1) continue;
        const int z = z0 + lz, xx = x0 + lx, par = (lx + FL) & 1;
        const int p = dn_pos(lz + FL, lx + FL), pl = p + (par ? -32 : 31), pr = p + (par ? -31 : 32);
        if (INT || (z >= rlo && z < rhi && xx >= 0 && xx < nx)) {
            // a red cell holds x = b / diag, so b - diag x vanishes (to the rounding of that one product) and what is
            // left of b - sum t (x - x_nb) is sum t x_nb over the black neighbours: no cancellation, two loads fewer
            double s = stx[pl] * sx[pl];
            s += stx[p] * sx[pr];
            s += stz[p - W] * sx[p - W];
            s += stz[p] * sx[p + W];
            sb[p] = s;
        }
    }
    // ---- x of the tile, out of place (independent of the residual pass: no barrier needed before it)
#pragma unroll
    for (int lz = wy; lz < TZ; lz += kDnWarps) {
        const int z = z0 + lz, xe = x0 + 2 * lane;
        if (lane >= TX / 2 || (!INT && (z >= zhi || xe >= nx))) continue;
        const size_t i = (size_t)z * nx + xe;
        const int p = dn_pos(lz + FL, 2 * lane + FL);      // even column; its right neighbour is at p + 32
        const double v0 = sx[p], v1 = sx[p + 32];
        if (INT || (xe + 1 < nx && !(nx & 1))) {
            *reinterpret_cast<double2 *>(x + i) = make_double2(v0, v1);   // i even when nx even: aligned
        } else {
            x[i] = v0;
            if (xe + 1 < nx) x[i + 1] = v1;
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CR; c++)
        if (cin[c]) {
            const int cr = wy + c * kDnWarps;
            const int p = dn_pos(2 * cr + FL, 2 * lane + FL);  // even column: left neighbour p + 31, right p + 32
            double s = sb[p];
            s += wne[c] * sb[p + W + 32];
            s += wnw[c] * sb[p + W + 31];
            s += wse[c] * sb[p - W + 32];
            s += wsw[c] * sb[p - W + 31];
            bc[(size_t)(z0 / 2 + cr) * nxc + X] = s;
        }
    __syncthreads();
}

__global__ void __launch_bounds__(kDnThreads, 2)
k_down5(Fine5 v, Weights w, const double *__restrict__ b, double *__restrict__ x, double *__restrict__ bc, int nzc,
        int nxc, int czhi, PcgState *st, int init, double *partials, unsigned int *counter, Dist dc)
{
    if (stopped(st)) return;
    constexpr int FL = kDnFL, FH = kDnFH, TZ = kDnTZ, TX = kDnTX;
    LCL_DYNAMIC_SMEM(double, smem);
    // a strip tiles its own rows [zlo, zhi) (zlo even) and owns the coarse rows below czhi; cells exist for it
    // in the rows [rlo, rhi) it holds.  The frame of a tile at a cut is recomputed from the ghost rows of b and
    // of the faces exactly as inside the grid: the sweep is exact red/black, so the result does not depend on
    // where tiles or strips are cut.
    const int nz = v.nz, nx = v.nx, zlo = v.zlo, zhi = v.zhi, rlo = read_lo(zlo), rhi = read_hi(zhi, nz);
    const int tiles_x = (nx + TX - 1) / TX, tiles_z = (zhi - zlo + TZ - 1) / TZ;
    double acc[2] = {0.0, 0.0};
    for (long long t = blockIdx.x; t < (long long)tiles_x * tiles_z; t += gridDim.x) {
        const int z0 = zlo + (int)(t / tiles_x) * TZ, x0 = (int)(t % tiles_x) * TX;
        const bool interior = !(nx & 1) && z0 - FL >= rlo && z0 + TZ + FH <= rhi && z0 + TZ <= zhi && x0 - FL >= 0 &&
                              x0 + TX + FH <= nx;
        if (interior) down5_tile<true>(smem, v, w, b, x, bc, nzc, nxc, czhi, z0, x0, rlo, rhi, acc);
        else down5_tile<false>(smem, v, w, b, x, bc, nzc, nxc, czhi, z0, x0, rlo, rhi, acc);
    }
    grid_sum_finalize<2>(acc, partials, counter, [=](const double(&tot)[2]) { down_norms_finalize(st, init, tot, dc); });
}

// Fine level, whole up leg in ONE launch, staged like k_down5: x += P e, one EXACT black/red sweep and
// b.x (= r.z of PCG).  28x60 tiles with a 2-cell frame (32 rows of 64; taller tiles were measured
// slower here: the x + P e operands are loaded synchronously, one round per 16 coarse rows): the thread that owns a coarse
// 2x2 block builds x + P e for it while the asynchronous copies of b and the faces are in flight.
constexpr int kUpTZ = 28, kUpTX = 60, kUpF = 2;
constexpr int kUpH = kUpTZ + 2 * kUpF;
constexpr size_t kUpSmem = 4 * (size_t)kUpH * kDnW * sizeof(double);
static_assert(kUpH == 2 * kDnWarps && kUpTX + 2 * kUpF == kDnW, "one coarse block per thread; rows as wide as the down leg's");

// One tile of the up leg; INT as in down5_tile (plus: every coarse block of the framed tile, its upper and right
// neighbours included, is held).
template <bool INT>
__device__ __forceinline__ void up5_tile(double *smem, const Fine5 &v, const double *__restrict__ b, const double *__restrict__ xin,
                                         double *__restrict__ xout, const Weights &w, const double *__restrict__ e, int nzc,
                                         int nxc, int blo, int bhi, int z0, int x0, int rlo, int rhi, double (&acc)[2])
{
    constexpr int W = kDnW, H = kUpH, F = kUpF, TZ = kUpTZ, TX = kUpTX;
    double *sb = smem, *stx = smem + H * W, *stz = smem + 2 * H * W, *sx = smem + 3 * H * W;
    const int nz = v.nz, nx = v.nx, zhi = v.zhi;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    stage_tile5<F, H, INT>(sb, stx, stz, v, b, z0, x0, wy, lane);
#pragma unroll
    for (int br = wy; br < H / 2; br += kDnWarps) {
        // x + P e of the coarse block (br, lane) of the framed tile: rows 2br, 2br+1, columns 2lane, 2lane+1
        const int bz = (z0 >> 1) - 1 + br, bx = (x0 >> 1) - 1 + lane;
        double v00 = 0.0, v01 = 0.0, v10 = 0.0, v11 = 0.0;
        if (INT || (bz >= blo && bx >= 0 && bz < bhi && bx < nxc)) {
            const size_t bk = (size_t)bz * nxc + bx;
            const size_t i = (size_t)(2 * bz) * nx + 2 * bx;
            const bool Xp = INT || bx + 1 < nxc, Zp = INT || bz + 1 < nzc;
            const bool xp = INT || 2 * bx + 1 < nx, zp = INT || 2 * bz + 1 < nz;
            const double e00 = e[bk];
            const double e01 = Xp ? e[bk + 1] : 0.0;
            const double e10 = Zp ? e[bk + nxc] : 0.0;
            const double e11 = (Zp && Xp) ? e[bk + nxc + 1] : 0.0;
            v00 = xin[i] + e00;
            if (xp) v01 = xin[i + 1] + (w.hw[bk] * e00 + w.he[bk] * e01);
            if (zp) v10 = xin[i + nx] + (w.vs[bk] * e00 + w.vn[bk] * e10);
            if (zp && xp)
                v11 = xin[i + nx + 1] + ((w.csw[bk] * e00 + w.cse[bk] * e01) + (w.cnw[bk] * e10 + w.cne[bk] * e11));
        }
        const int p = 2 * br * W + lane;               // dn_pos(2 br, 2 lane); the odd column sits 32 further
        sx[p] = v00;
        sx[p + 32] = v01;
        sx[p + W] = v10;
        sx[p + W + 32] = v11;
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- black cells ((lz + lx) odd) of the tile and its first frame ring, rows -1..TZ, columns -1..TX
#pragma unroll
    for (int row = wy; row < TZ + 2; row += kDnWarps) {
        const int lz = row - 1, lx = 2 * lane - 1 + (lz & 1);
        if (lx > TX) continue;
        const int z = z0 + lz, xx = x0 + lx, par = (lx + F) & 1;
        const int p = dn_pos(lz + F, lx + F), pl = p + (par ? -32 : 31), pr = p + (par ? -31 : 32);
        if (INT || (z >= rlo && z < rhi && xx >= 0 && xx < nx)) {
            const double tl = stx[pl], tr = stx[p], td = stz[p - W], tu = stz[p];
            double s = sb[p];
            s += tl * sx[pl];
            s += tr * sx[pr];
            s += td * sx[p - W];
            s += tu * sx[p + W];
            sx[p] = s * rcp_diag((tl + tr) + (td + tu));
        }
    }
    __syncthreads();
    // ---- red cells of the tile: each lane owns a horizontal pair (one red cell to update, one black cell
    // to copy), accumulates b.x over both and writes them out of place
#pragma unroll
    for (int lz = wy; lz < TZ; lz += kDnWarps) {
        const int z = z0 + lz, xe = x0 + 2 * lane;
        if (lane >= TX / 2 || (!INT && (z >= zhi || xe >= nx))) continue;
        const size_t i = (size_t)z * nx + xe;
        const int pe = dn_pos(lz + F, 2 * lane + F), po = pe + 32;
        const int r = lz & 1;                          // 0: the even cell is red, 1: the odd cell is red
        double v0 = sx[pe], v1 = sx[po];
        if (INT || xe + r < nx) {
            const int p = r ? po : pe, pl = r ? pe : pe + 31, pr = r ? po - 31 : po;
            const double tl = stx[pl], tr = stx[p], td = stz[p - W], tu = stz[p];
            double s = sb[p];
            s += tl * sx[pl];
            s += tr * sx[pr];
            s += td * sx[p - W];
            s += tu * sx[p + W];
            const double xn = s * rcp_diag((tl + tr) + (td + tu));
            if (r) v1 = xn; else v0 = xn;
        }
        acc[0] += sb[pe] * v0;
        acc[0] += sb[po] * v1;                         // b and x are both 0 beyond the grid
        if (INT || (xe + 1 < nx && !(nx & 1))) {
            *reinterpret_cast<double2 *>(xout + i) = make_double2(v0, v1);   // i even when nx even: aligned
        } else {
            xout[i] = v0;
            if (xe + 1 < nx) xout[i + 1] = v1;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kDnThreads, 2)
k_up5(Fine5 v, const double *__restrict__ b, const double *__restrict__ xin, double *__restrict__ xout, Weights w,
      const double *__restrict__ e, int nzc, int nxc, int wlo, int whi, PcgState *st, int init, double *partials,
      unsigned int *counter, Dist dc)
{
    if (stopped(st)) return;
    constexpr int F = kUpF, TZ = kUpTZ, TX = kUpTX, H = kUpH;
    LCL_DYNAMIC_SMEM(double, smem);
    // strip: own rows [zlo, zhi), held rows [rlo, rhi); coarse blocks whose weights, coarse values and fine
    // rows this rank holds: [blo, bhi)
    const int nz = v.nz, nx = v.nx, zlo = v.zlo, zhi = v.zhi, rlo = read_lo(zlo), rhi = read_hi(zhi, nz);
    const int blo = wlo > ((rlo + 1) >> 1) ? wlo : ((rlo + 1) >> 1), bhi = whi < ((rhi + 1) >> 1) ? whi : ((rhi + 1) >> 1);
    const int tiles_x = (nx + TX - 1) / TX, tiles_z = (zhi - zlo + TZ - 1) / TZ;
    double acc[2] = {0.0, 0.0};
    for (long long t = blockIdx.x; t < (long long)tiles_x * tiles_z; t += gridDim.x) {
        const int z0 = zlo + (int)(t / tiles_x) * TZ, x0 = (int)(t % tiles_x) * TX;
        // first and last coarse block row / column the framed tile touches (their upper / right neighbours included)
        const int bz0 = (z0 >> 1) - 1, bz1 = bz0 + H / 2 - 1, bx0 = (x0 >> 1) - 1, bx1 = bx0 + 31;
        const bool interior = !(nx & 1) && z0 - F >= rlo && z0 + TZ + F <= rhi && z0 + TZ <= zhi && x0 - F >= 0 &&
                              x0 + TX + F <= nx && bz0 >= blo && bz1 + 1 < bhi && bz1 + 1 < nzc && bx0 >= 0 && bx1 + 1 < nxc &&
                              2 * bz1 + 1 < nz && 2 * bx1 + 1 < nx;
        if (interior) up5_tile<true>(smem, v, b, xin, xout, w, e, nzc, nxc, blo, bhi, z0, x0, rlo, rhi, acc);
        else up5_tile<false>(smem, v, b, xin, xout, w, e, nzc, nxc, blo, bhi, z0, x0, rlo, rhi, acc);
    }
    grid_sum_finalize<2>(acc, partials, counter, [=](const double(&part)[2]) {
        double tot[1] = {part[0]};
        dist_sum<1>(dc, tot);
        st->rz_new = tot[0];
        st->beta = init ? 0.0 : tot[0] / st->rz;
        st->rz = tot[0];
    });
}

// ------------------------------------------------------------------------------------ transfers
// bc = P^T r : thread per coarse point gathers its 3x3 fine neighbourhood
__global__ void __launch_bounds__(kBlock)
k_restrict(Weights w, const double *__restrict__ r, int nz, int nx, double *__restrict__ bc, int nzc, int nxc,
           int czlo, int czhi, const PcgState *st)
{
    if (stopped(st)) return;
    const Tiles tiles(czlo, czhi, nxc);   // the coarse rows this rank owns (all of them on one GPU)
    for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
        int Z, X;
        tiles.cell(t, Z, X);
        if (Z >= czhi || X >= nxc) continue;
        const int z = 2 * Z, x = 2 * X;
        const size_t b = (size_t)Z * nxc + X;
        const size_t i = (size_t)z * nx + x;
        const bool xm = X > 0, xp = x + 1 < nx, zm = Z > 0, zp = z + 1 < nz;
        double s = r[i];
        if (xp) s += w.hw[b] * r[i + 1];
        if (xm) s += w.he[b - 1] * r[i - 1];
        if (zp) s += w.vs[b] * r[i + nx];
        if (zm) s += w.vn[b - nxc] * r[i - nx];
        if (zp && xp) s += w.csw[b] * r[i + nx + 1];
        if (zp && xm) s += w.cse[b - 1] * r[i + nx - 1];
        if (zm && xp) s += w.cnw[b - nxc] * r[i - nx + 1];
        if (zm && xm) s += w.cne[b - nxc - 1] * r[i - nx - 1];
        bc[b] = s;
    }
}

// ------------------------------------------------------------------------------- the tail levels
// Operator row of one cell of a tail level, in registers while the level is being smoothed: couplings
// in the order the sums run (SW, S, SE, W, E, NW, N, NE; 0 where the neighbour is absent), diagonal, and
// the absorption part S of the difference-form residual.
struct TailCell {
    double gsw, gs, gse, gw, ge, gnw, gn, gne, diag, S;
};
__device__ __forceinline__ TailCell tail_load_cell(const Coarse9 &v, int z, int x)
{
    TailCell c{0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0};
    const int nx = v.nx, i = z * nx + x;
    const bool xm = x > 0, xp = x + 1 < nx, zm = z > 0;
    c.ge = v.E[i];                            // E, N, NE, NW are stored 0 where the neighbour is absent
    c.gn = v.N[i];
    c.gne = v.NE[i];
    c.gnw = v.NW[i];
    if (xm) c.gw = v.E[i - 1];
    if (zm) {
        c.gs = v.N[i - nx];
        if (xm) c.gsw = v.NE[i - nx - 1];
        if (xp) c.gse = v.NW[i - nx + 1];
    }
    c.diag = v.C[i];
    c.S = v.S[i];
    return c;
}
// one Gauss-Seidel update of the cell at p (row pitch W, zero frame around the level)
__device__ __forceinline__ void tail_relax(const TailCell &c, double b, double *p, int W)
{
    double s = b;
    s += c.gsw * p[-W - 1];
    s += c.gs * p[-W];
    s += c.gse * p[-W + 1];
    s += c.gw * p[-1];
    s += c.ge * p[1];
    s += c.gnw * p[W - 1];
    s += c.gn * p[W];
    s += c.gne * p[W + 1];
    p[0] = s / c.diag;
}

// All levels of at most kTailCells cells -- 32x32 and below -- in ONE launch by ONE block: the whole
// sub-cycle (pre-smoothing, residual, restriction, ..., dense solve, ..., prolongation, post-smoothing)
// with block barriers where the per-level kernels had launch boundaries.  These levels hold a few
// hundred cells: their cost is pure latency (about 16 launches per visit as separate kernels, several
// visits per cycle with the W-levels above them).  A level is one tile, so the smoother is exact
// four-colour Gauss-Seidel, backward on the way up: the cycle stays symmetric.
// Thread t owns cell t of every level.  The iterates, right-hand sides and residuals of all tail levels
// live in shared memory (iterates and residuals inside a one-cell frame of zeros, so no pass has
// neighbour logic); a thread fetches its operator row, its restriction or interpolation weights and
// its right-hand side with independent loads at the head of each level visit and the colour passes then
// run on registers and shared memory.  One visit of a level costs one L2 round trip instead of one per
// pass (measured at 8192^2: 50 -> see profiles/README.md us per launch).
// rhs in lv[0].b; result in lv[0].r (zero_init == 0: lv[0].r is the initial guess).
__global__ void __launch_bounds__(kTailThreads)
k_tail_cycle(TailArgs a, int zero_init, const PcgState *st)
{
    if (stopped(st)) return;
    const int tid = threadIdx.x, T = blockDim.x;
    LCL_DYNAMIC_SMEM(double, smem);
    // shared-memory map: framed iterate X_l and framed residual R_l per level, then the right-hand sides
    int offX[kTailMaxLevels], offB[kTailMaxLevels];
    int framed = 0, plain = 0;
    for (int l = 0; l < a.nlev; l++) {
        offX[l] = framed;
        offB[l] = plain;
        framed += (a.lv[l].op.nz + 2) * (a.lv[l].op.nx + 2);
        plain += a.lv[l].op.nz * a.lv[l].op.nx;
    }
    double *SX = smem, *SR = smem + framed, *SB = smem + 2 * framed;
    for (int i = tid; i < 2 * framed; i += T) smem[i] = 0.0;
    __syncthreads();

    for (int l = 0; l + 1 < a.nlev; l++) {
        const TailLevel &L = a.lv[l], &Cc = a.lv[l + 1];
        const int nz = L.op.nz, nx = L.op.nx, n = nz * nx, nzc = Cc.op.nz, nxc = Cc.op.nx, W = nx + 2;
        const bool own = tid < n;
        const int z = own ? tid / nx : 0, xx = own ? tid % nx : 0;
        double *X = SX + offX[l], *R = SR + offX[l], *B = SB + offB[l];
        double *p = X + (z + 1) * W + xx + 1;
        TailCell c{};
        double b = 0.0;
        if (own) {
            c = tail_load_cell(L.op, z, xx);
            b = l == 0 ? L.b[tid] : B[tid];
            if (l == 0) B[tid] = b;
            if (l == 0 && !zero_init) *p = L.r[tid];
        }
        // restriction weights of the coarse point this thread owns (0 where the fine neighbour is absent)
        const bool cown = tid < nzc * nxc;
        double whw = 0.0, whe = 0.0, wvs = 0.0, wvn = 0.0, wsw = 0.0, wse = 0.0, wnw = 0.0, wne = 0.0;
        const int Z = cown ? tid / nxc : 0, Xc = cown ? tid % nxc : 0;
        if (cown) {
            const int k = tid;
            const bool xm = Xc > 0, xp = 2 * Xc + 1 < nx, zm = Z > 0, zp = 2 * Z + 1 < nz;
            if (xp) whw = L.w.hw[k];
            if (xm) whe = L.w.he[k - 1];
            if (zp) wvs = L.w.vs[k];
            if (zm) wvn = L.w.vn[k - nxc];
            if (zp && xp) wsw = L.w.csw[k];
            if (zp && xm) wse = L.w.cse[k - 1];
            if (zm && xp) wnw = L.w.cnw[k - nxc];
            if (zm && xm) wne = L.w.cne[k - nxc - 1];
        }
        const int colour = 2 * (z & 1) + (xx & 1);
        __syncthreads();
        for (int sweep = 0; sweep < a.nu; sweep++)
            for (int col = 0; col < 4; col++) {
                if (own && colour == col) tail_relax(c, b, p, W);
                __syncthreads();
            }
        if (own) {   // residual, difference form
            const double xi = *p;
            double s = b - c.S * xi;
            s -= c.gsw * (xi - p[-W - 1]);
            s -= c.gs * (xi - p[-W]);
            s -= c.gse * (xi - p[-W + 1]);
            s -= c.gw * (xi - p[-1]);
            s -= c.ge * (xi - p[1]);
            s -= c.gnw * (xi - p[W - 1]);
            s -= c.gn * (xi - p[W]);
            s -= c.gne * (xi - p[W + 1]);
            R[(z + 1) * W + xx + 1] = s;
        }
        __syncthreads();
        if (cown) {   // restriction, as k_restrict
            const double *r = R + (2 * Z + 1) * W + 2 * Xc + 1;
            double s = r[0];
            s += whw * r[1];
            s += whe * r[-1];
            s += wvs * r[W];
            s += wvn * r[-W];
            s += wsw * r[W + 1];
            s += wse * r[W - 1];
            s += wnw * r[-W + 1];
            s += wne * r[-W - 1];
            SB[offB[l + 1] + tid] = s;
        }
        __syncthreads();
    }
    {   // last level: x = S inv S b
        const int l = a.nlev - 1;
        const TailLevel &L = a.lv[l];
        const int nx = L.op.nx, n = L.op.nz * nx, W = nx + 2;
        double *X = SX + offX[l], *B = SB + offB[l];
        __shared__ double sb[64];
        for (int i = tid; i < n; i += T) sb[i] = a.scale[i] * B[i];
        __syncthreads();
        for (int i = tid; i < n; i += T) {
            double s = 0.0;
            for (int j = 0; j < n; j++) s += a.inv[i * n + j] * sb[j];
            X[(i / nx + 1) * W + i % nx + 1] = a.scale[i] * s;
        }
        __syncthreads();
    }
    for (int l = a.nlev - 2; l >= 0; l--) {
        const TailLevel &L = a.lv[l], &Cc = a.lv[l + 1];
        const int nz = L.op.nz, nx = L.op.nx, n = nz * nx, nxc = Cc.op.nx, W = nx + 2, Wc = nxc + 2;
        const bool own = tid < n;
        const int z = own ? tid / nx : 0, xx = own ? tid % nx : 0;
        double *X = SX + offX[l];
        double *p = X + (z + 1) * W + xx + 1;
        TailCell c{};
        double b = 0.0;
        if (own) {
            c = tail_load_cell(L.op, z, xx);
            b = SB[offB[l] + tid];
            // x += P e for this cell: e sits framed in the coarse level's iterate
            const int k = (z >> 1) * nxc + (xx >> 1);
            const double *e = SX + offX[l + 1] + ((z >> 1) + 1) * Wc + (xx >> 1) + 1;
            const double e00 = e[0], e01 = e[1], e10 = e[Wc], e11 = e[Wc + 1];
            double pe;
            if (!(z & 1)) pe = !(xx & 1) ? e00 : L.w.hw[k] * e00 + L.w.he[k] * e01;
            else if (!(xx & 1)) pe = L.w.vs[k] * e00 + L.w.vn[k] * e10;
            else pe = (L.w.csw[k] * e00 + L.w.cse[k] * e01) + (L.w.cnw[k] * e10 + L.w.cne[k] * e11);
            *p += pe;
        }
        const int colour = 2 * (z & 1) + (xx & 1);
        __syncthreads();
        for (int sweep = 0; sweep < a.nu; sweep++)
            for (int col = 3; col >= 0; col--) {
                if (own && colour == col) tail_relax(c, b, p, W);
                __syncthreads();
            }
        if (l == 0 && own) L.r[tid] = *p;
    }
}

// x = S inv S b on the coarsest level (one CTA)
__global__ void k_coarse_solve(const double *__restrict__ inv, const double *__restrict__ scale,
                               const double *__restrict__ b, double *__restrict__ x, int n, const PcgState *st)
{
    if (stopped(st)) return;
    LCL_DYNAMIC_SMEM(double, sb);
    for (int i = threadIdx.x; i < n; i += blockDim.x) sb[i] = scale[i] * b[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < n; j++) s += inv[(size_t)i * n + j] * sb[j];
        x[i] = scale[i] * s;
    }
}

__global__ void k_copy(const double *__restrict__ a, double *__restrict__ b, size_t n, const PcgState *st)
{
    if (stopped(st)) return;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

// fine-level b.x when the hierarchy has a single (dense) level
__global__ void __launch_bounds__(kBlock)
k_dot_norms(const double *__restrict__ b, const double *__restrict__ x, Fine5 v, PcgState *st, int init,
            double *partials, unsigned int *counter)
{
    if (stopped(st)) return;
    const F5 A{v};
    const size_t n = (size_t)v.nz * v.nx;
    double acc[3] = {0.0, 0.0, 0.0};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(i / v.nx), xx = (int)(i % v.nx);
        const double bi = b[i];
        acc[0] += bi * bi;
        acc[1] += bi * bi / A.diag(z, xx);
        acc[2] += bi * x[i];
    }
    grid_sum_finalize<3>(acc, partials, counter, [=](const double(&tot)[3]) {
        st->dx = st->dx_part;   // a map this small is never striped: the strip's change is the map's
        st->rr = tot[0];
        st->rs = tot[1];
        if (init) { st->bb = tot[0]; st->bs = tot[1]; }
        // the dense "preconditioner" is an exact solve: the next residual test decides
        st->rz_new = tot[2];
        st->beta = init ? 0.0 : tot[2] / st->rz;
        st->rz = tot[2];
        if (!init && tot[0] <= st->tol2 * st->bb && tot[1] <= st->tol2 * st->bs) {
            st->res_ok += 1;
            if ((st->dx <= st->dx_tol && st->dx_prev <= st->dx_tol) || (st->dx > st->dx_prev && st->dx_prev > st->dx_prev2) ||
            st->dx > 1e4 * st->dx_tol || st->res_ok > 12)
                st->converged = 1;
        }
    });
}

// ------------------------------------------------------------------------------------- driver
// persistent grids of the fused fine-level legs: two resident blocks per SM
static int leg_grid(Solver &s, int tz, int tx)
{
    if (!s.down_configured) {   // per device: the attribute lives in the context of the solver's GPU
        CUDA_CHECK(cudaFuncSetAttribute(k_down5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDnSmem));
        CUDA_CHECK(cudaFuncSetAttribute(k_up5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpSmem));
        CUDA_CHECK(cudaFuncSetAttribute(k_tail_cycle, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailMaxSmem));
        CUDA_CHECK(cudaFuncSetAttribute(k_smooth9<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSm9Smem));
        CUDA_CHECK(cudaFuncSetAttribute(k_smooth9<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSm9Smem));
        CUDA_CHECK(cudaFuncSetAttribute(k_smooth9<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSm9SmemUp));
        s.down_configured = true;
    }
    const long long tiles = (long long)div_up(s.nx, tx) * div_up(s.zhi - s.zlo, tz);
    long long cap = (long long)s.num_sms * 2;
    if (cap > kMaxPartials) cap = kMaxPartials;
    return (int)(tiles < cap ? tiles : cap);
}
static int down_grid(Solver &s) { return leg_grid(s, kDnTZ, kDnTX); }
static int up_grid(Solver &s) { return leg_grid(s, kUpTZ, kUpTX); }

// persistent grid of the 9-point smoother over the rows a rank owns: two resident blocks per SM, each pipelining its tiles
static int smooth_grid(const Solver &s, const Level &f)
{
    long long tiles = (long long)div_up(f.nx, kSmTX) * div_up(f.zhi - f.zlo, kSmTZ);
    long long cap = (long long)s.num_sms * 2;
    if (cap > kMaxPartials) cap = kMaxPartials;
    return (int)(tiles < cap ? tiles : cap);
}
// coarse rows of level f's weights (and of the next level's vectors) this rank holds
static int w_lo(const Level &f) { return f.striped ? read_lo(f.czlo) : 0; }
static int w_hi(const Level &f, const Level &c) { return f.striped ? read_hi(f.czhi, c.nz) : c.nz; }
// residual rows: a strip adds the row below its own (restriction of its first coarse row)
static int res_lo(const Level &f) { return f.striped && f.zlo > 0 ? f.zlo - 1 : f.zlo; }

static void broadcast1(Solver &s, double *p, int pitch, const int *cut, const PcgState *st)
{
    HaloSet hs;
    hs.add(p);
    row_broadcast(s, hs, pitch, cut, st);
}
static void exchange1(Solver &s, double *p, int pitch, const int *cut, int dlo, int dhi, const PcgState *st)
{
    HaloSet hs;
    hs.add(p);
    halo_exchange(s, hs, pitch, cut, dlo, dhi, st);
}

double bench_launch_cycle(Solver &s, const std::string &name)
{
    const Fine5 v5 = s.view5();
    double *part = s.red + 16 * kMaxPartials;
    unsigned int *cnt = s.red_counters + 4;
    const double n = (double)s.own_cells();
    if (name == "up5" && s.levels.size() > 1 && s.levels[0].w.hw) {
        Level &f = s.levels[0], &c = s.levels[1];
        LAUNCH(k_up5, up_grid(s), kDnThreads, kUpSmem, s.stream, v5, (const double *)s.r, (const double *)s.z, s.q,
               f.w, (const double *)c.r, c.nz, c.nx, w_lo(f), w_hi(f, c), s.state, 0, part + 4 * kMaxPartials, cnt + 1, s.dist);
        return 40.0 * n + (8.0 + 64.0) * (double)(f.czhi - f.czlo) * c.nx;   // x, b, tx, tz in; x out; coarse e + 8 weights
    }
    if (name == "smooth9_l1" && s.levels.size() > 2) {
        Level &f = s.levels[1];
        LAUNCH((k_smooth9<true, false, true>), smooth_grid(s, f), kSmThreads, kSm9Smem, s.stream, f.view9(), (const double *)f.b,
               (const double *)nullptr, f.x, s.nu_coarse, Weights{}, (const double *)nullptr, 0, 0, 0, 0,
               (const PcgState *)s.state);
        return 56.0 * (double)f.own_cells();   // b, C, E, N, NE, NW in; x out (couplings re-read from L1/L2 per sweep)
    }
    if (name == "down5" && s.levels.size() > 1 && s.levels[0].w.hw) {
        Level &f = s.levels[0], &c = s.levels[1];
        LAUNCH(k_down5, down_grid(s), kDnThreads, kDnSmem, s.stream, v5, f.w, (const double *)s.r, s.z, c.b, c.nz,
               c.nx, f.czhi, s.state, 0, part, cnt, s.dist);
        return 32.0 * n + (8.0 + 32.0) * (double)(f.czhi - f.czlo) * c.nx;   // b, tx, tz in, x out; 4 weights in + coarse rhs out
    }
    if (name == "smooth9_l1_up" && s.levels.size() > 2) {
        Level &f = s.levels[1], &c = s.levels[2];
        LAUNCH((k_smooth9<false, true, false>), smooth_grid(s, f), kSmThreads, kSm9SmemUp, s.stream, f.view9(), (const double *)f.b,
               (const double *)f.x, f.r, s.nu_coarse, f.w, (const double *)c.r, c.nz, c.nx, w_lo(f), w_hi(f, c),
               (const PcgState *)s.state);
        return 64.0 * (double)f.own_cells() + (8.0 + 64.0) * (double)(f.czhi - f.czlo) * c.nx;
    }
    if (name == "residual9_l1" && s.levels.size() > 2) {
        Level &f = s.levels[1];
        LAUNCH(k_residual9, grid_for(s, f.zhi - res_lo(f), f.nx, 8), kBlock, 0, s.stream, f.view9(), (const double *)f.b,
               (const double *)f.x, f.r, res_lo(f), f.zhi, s.state);
        return 64.0 * (double)f.own_cells();   // b, x, S, E, N, NE, NW in; r out
    }
    if (name == "restrict1" && s.levels.size() > 2) {
        Level &f = s.levels[1], &c = s.levels[2];
        LAUNCH(k_restrict, grid_for(s, f.czhi - f.czlo, c.nx, 8), kBlock, 0, s.stream, f.w, (const double *)f.r, f.nz, f.nx,
               c.b, c.nz, c.nx, f.czlo, f.czhi, s.state);
        return 8.0 * (double)f.own_cells() + 72.0 * (double)(f.czhi - f.czlo) * c.nx;
    }
    if (name == "vcycle") {
        mg_vcycle(s, s.r, s.z, s.q, s.state, false);
        return 0.0;
    }
    return -1.0;
}

// One cycle.  On a striped level every kernel walks the rank's own rows; what it reads beyond them are ghost
// rows that the producing step has pushed to the neighbours (halo_exchange, peer stores + a flag, no NCCL):
//   fine level   r (4 rows below / 2 above: the down leg recomputes its frame), the down leg's x (2 / 2), z (1 / 1)
//   9-point level l   x and b after the down smoother (2 / 2: residual of the row below the strip, frame of the up
//                     smoother), the level's result after the up smoother (2 / 2: prolongation one level up, or
//                     the starting guess of a second visit)
// The right-hand side of the first whole level is assembled from all ranks' rows (row_broadcast); from there
// down every rank cycles the same small problem redundantly, and its result needs no exchange on the way up.
void mg_vcycle(Solver &s, const double *r, double *ztmp, double *zout, PcgState *st, bool init_flag)
{
    const int init = init_flag ? 1 : 0;
    const Fine5 v5 = s.view5();
    double *part = s.red + 16 * kMaxPartials;   // slots 16.. belong to the cycle
    unsigned int *cnt = s.red_counters + 4;
    const size_t nlev = s.levels.size();

    if (s.levels[0].w.hw == nullptr) {   // map small enough for the dense solve alone
        Level &c = s.levels[1];
        LAUNCH_DEP(k_coarse_solve, 1, 64, c.n * sizeof(double), s.stream, s.coarse_inv, s.coarse_scale, r, zout, (int)c.n, st);
        LAUNCH_DEP(k_dot_norms, 1, kBlock, 0, s.stream, r, (const double *)zout, v5, st, init, part, cnt);
        return;
    }

    // The fused smoothers are out of place (a tile must not see its neighbours' new values through
    // the halo), so each level ping-pongs: down leg x -> level.x, up leg result -> level.r (the
    // residual it held has been restricted by then).  Fine level: ztmp after the down leg, zout final;
    // zout doubles as the fine residual buffer in between.
    {   // whole fine down leg: smooth, residual, restriction
        Level &f = s.levels[0], &c = s.levels[1];
        if (f.striped) exchange1(s, const_cast<double *>(r), f.nx, f.cut, 4, 2, st);
        LAUNCH_DEP(k_down5, down_grid(s), kDnThreads, kDnSmem, s.stream, v5, f.w, r, ztmp, c.b, c.nz, c.nx, f.czhi, st, init,
               part, cnt, s.dist);
        if (f.striped) {
            exchange1(s, ztmp, f.nx, f.cut, 2, 2, st);
            if (!c.striped) broadcast1(s, c.b, c.nx, f.ccut, st);
        }
    }
    // Levels >= 1: cycle(l) solves level l approximately for the right-hand side in .b, result in .r.
    // Levels in [w_lo, w_hi] are visited twice (second time from the first result): a W-cycle
    // truncated to the few mid levels where the V-cycle loses robustness on contact maps, so the
    // launch-bound tail levels are not revisited exponentially often.  Two steps of a symmetric
    // stationary iteration are again a symmetric operator, so CG still applies.
    const Weights now{};
    std::function<void(size_t, bool)> cycle = [&](size_t l, bool zero_init) {
        Level &f = s.levels[l];
        if ((int)l == s.tail_top) {   // this level and everything below it: one block, one launch
            LAUNCH_DEP(k_tail_cycle, 1, kTailThreads, s.tail_smem, s.stream, s.tail, zero_init ? 1 : 0, (const PcgState *)st);
            return;
        }
        if (l == nlev - 1) {   // coarsest: dense solve
            LAUNCH_DEP(k_coarse_solve, 1, 64, f.n * sizeof(double), s.stream, s.coarse_inv, s.coarse_scale,
                   (const double *)f.b, f.r, (int)f.n, st);
            return;
        }
        Level &c = s.levels[l + 1];
        const int nu = l == 1 ? s.nu_coarse : s.nu_deep;
        if (zero_init)
            LAUNCH_DEP((k_smooth9<true, false, true>), smooth_grid(s, f), kSmThreads, kSm9Smem, s.stream, f.view9(),
                   (const double *)f.b, (const double *)nullptr, f.x, nu, now, (const double *)nullptr, 0, 0, 0, 0,
                   (const PcgState *)st);
        else
            LAUNCH_DEP((k_smooth9<false, false, true>), smooth_grid(s, f), kSmThreads, kSm9Smem, s.stream, f.view9(),
                   (const double *)f.b, (const double *)f.r, f.x, nu, now, (const double *)nullptr, 0, 0, 0, 0,
                   (const PcgState *)st);
        if (f.striped) {   // b rides along: the residual of the row below the strip reads it
            HaloSet hs;
            hs.add(f.x);
            hs.add(f.b);
            halo_exchange(s, hs, f.nx, f.cut, 2, 2, st);
        }
        LAUNCH_DEP(k_residual9, grid_for(s, f.zhi - res_lo(f), f.nx, 8), kBlock, 0, s.stream, f.view9(), (const double *)f.b,
               (const double *)f.x, f.r, res_lo(f), f.zhi, st);
        LAUNCH_DEP(k_restrict, grid_for(s, f.czhi - f.czlo, c.nx, 8), kBlock, 0, s.stream, f.w, (const double *)f.r, f.nz, f.nx,
               c.b, c.nz, c.nx, f.czlo, f.czhi, st);
        if (f.striped && !c.striped) broadcast1(s, c.b, c.nx, f.ccut, st);
        cycle(l + 1, true);
        if ((int)(l + 1) >= s.w_lo && (int)(l + 1) <= s.w_hi && l + 1 < nlev - 1) cycle(l + 1, false);
        LAUNCH_DEP((k_smooth9<false, true, false>), smooth_grid(s, f), kSmThreads, kSm9SmemUp, s.stream, f.view9(),
               (const double *)f.b, (const double *)f.x, f.r, nu, f.w, (const double *)c.r, c.nz, c.nx, w_lo(f), w_hi(f, c),
               (const PcgState *)st);
        if (f.striped) exchange1(s, f.r, f.nx, f.cut, 2, 2, st);
    };
    cycle(1, true);
    if (1 >= s.w_lo && 1 <= s.w_hi && nlev > 2) cycle(1, false);
    {
        Level &f = s.levels[0], &c = s.levels[1];
        LAUNCH_DEP(k_up5, up_grid(s), kDnThreads, kUpSmem, s.stream, v5, r, (const double *)ztmp, zout, f.w,
               (const double *)c.r, c.nz, c.nx, w_lo(f), w_hi(f, c), st, init, part + 4 * kMaxPartials, cnt + 1, s.dist);
        if (f.striped) exchange1(s, zout, f.nx, f.cut, 1, 1, st);
    }
}

// forces the kernels of this file into the device's context now (see preload_kernels, api.cu)
void preload_cycle()
{
    preload_one(k_residual9);
    preload_one((k_smooth9<true, false, true>));
    preload_one((k_smooth9<false, false, true>));
    preload_one((k_smooth9<false, true, false>));
    preload_one(k_down5);
    preload_one(k_up5);
    preload_one(k_restrict);
    preload_one(k_tail_cycle);
    preload_one(k_coarse_solve);
    preload_one(k_copy);
    preload_one(k_dot_norms);
}

}   // namespace lcl
