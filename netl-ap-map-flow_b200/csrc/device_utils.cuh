// device_utils.cuh -- device-side helpers: stencil accessors, deterministic fused reductions.
#pragma once

#include "common.cuh"
#include "dist.cuh"

namespace lcl {

constexpr int kTileX = 32;
constexpr int kTileZ = 8;
constexpr int kBlock = kTileX * kTileZ;   // 256 threads, one cell each per tile

__device__ __forceinline__ double ldcg(const double *p) { return __ldcg(p); }

// 8-byte asynchronous global -> shared copy (LDGSTS: the data never passes through registers, so a
// thread can have a whole tile's worth in flight); !pred zero-fills the destination instead.
__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src, bool pred)
{
#ifdef LCL_CPU_SIM
    *dst_smem = pred ? *src : 0.0;
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int nbytes = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all()
{
#ifndef LCL_CPU_SIM
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// ---------------------------------------------------------------------------------------------
// 5-point fine operator accessors (all couplings are positive conductances)
struct F5 {
    Fine5 v;
    __device__ __forceinline__ double left(int z, int x) const { return x > 0 ? v.tx[(size_t)z * v.nx + x - 1] : v.txl[z]; }
    __device__ __forceinline__ double right(int z, int x) const { return v.tx[(size_t)z * v.nx + x]; }
    __device__ __forceinline__ double down(int z, int x) const { return z > 0 ? v.tz[(size_t)(z - 1) * v.nx + x] : v.tzb[x]; }
    __device__ __forceinline__ double up(int z, int x) const { return v.tz[(size_t)z * v.nx + x]; }
    __device__ __forceinline__ double diag(int z, int x) const
    {
        return (left(z, x) + right(z, x)) + (down(z, x) + up(z, x));
    }
    // absorption: the Dirichlet boundary faces of this cell (0 in the interior)
    __device__ __forceinline__ double S(int z, int x) const
    {
        double s = 0.0;
        if (x == 0) s += v.txl[z];
        if (x == v.nx - 1) s += v.tx[(size_t)z * v.nx + x];
        if (z == 0) s += v.tzb[x];
        if (z == v.nz - 1) s += v.tz[(size_t)z * v.nx + x];
        return s;
    }
    // coupling to an existing neighbour, 0 outside the grid / for the (absent) diagonal neighbours
    __device__ __forceinline__ double g(int z, int x, int dz, int dx) const
    {
        if (dz == 0) {
            if (dx == 1) return x + 1 < v.nx ? right(z, x) : 0.0;
            if (dx == -1) return x > 0 ? left(z, x) : 0.0;
            return 0.0;
        }
        if (dx == 0) {
            if (dz == 1) return z + 1 < v.nz ? up(z, x) : 0.0;
            if (dz == -1) return z > 0 ? down(z, x) : 0.0;
        }
        return 0.0;
    }
    __device__ __forceinline__ int nz() const { return v.nz; }
    __device__ __forceinline__ int nx() const { return v.nx; }
};

// 9-point coarse operator accessors
struct C9 {
    Coarse9 v;
    __device__ __forceinline__ double diag(int z, int x) const { return v.C[(size_t)z * v.nx + x]; }
    __device__ __forceinline__ double S(int z, int x) const { return v.S[(size_t)z * v.nx + x]; }
    __device__ __forceinline__ double g(int z, int x, int dz, int dx) const
    {
        const size_t i = (size_t)z * v.nx + x;
        const int nx = v.nx;
        if (dz == 0) {
            if (dx == 1) return v.E[i];                          // stored 0 when x == nx-1
            if (dx == -1) return x > 0 ? v.E[i - 1] : 0.0;
            return 0.0;
        }
        if (dz == 1) {
            if (dx == 0) return v.N[i];
            if (dx == 1) return v.NE[i];
            return v.NW[i];
        }
        if (z == 0) return 0.0;
        if (dx == 0) return v.N[i - nx];
        if (dx == -1) return x > 0 ? v.NE[i - nx - 1] : 0.0;
        return x + 1 < nx ? v.NW[i - nx + 1] : 0.0;
    }
    __device__ __forceinline__ int nz() const { return v.nz; }
    __device__ __forceinline__ int nx() const { return v.nx; }
};

// weight with which coarse point (Jz,Jx) contributes to fine point (qz,qx); 0 if unrelated
__device__ __forceinline__ double pweight(const Weights &w, int nzc, int nxc, int qz, int qx, int Jz, int Jx)
{
    if (Jz < 0 || Jx < 0 || Jz >= nzc || Jx >= nxc) return 0.0;
    const int dz = qz - 2 * Jz, dx = qx - 2 * Jx;
    if (dz < -1 || dz > 1 || dx < -1 || dx > 1) return 0.0;
    if (dz == 0 && dx == 0) return 1.0;
    const int bz = qz >> 1, bx = qx >> 1;          // block that owns the fine point
    const size_t b = (size_t)bz * nxc + bx;
    if (dz == 0) return dx == 1 ? w.hw[b] : w.he[b];
    if (dx == 0) return dz == 1 ? w.vs[b] : w.vn[b];
    if (dz == 1) return dx == 1 ? w.csw[b] : w.cse[b];
    return dx == 1 ? w.cnw[b] : w.cne[b];
}

// 1 - (sum of the interpolation weights of fine point (qz,qx)); exactly 0 for coarse points
__device__ __forceinline__ double pdeficit(const Weights &w, int nxc, int qz, int qx)
{
    const size_t b = (size_t)(qz >> 1) * nxc + (qx >> 1);
    if (qz & 1) return (qx & 1) ? w.tc[b] : w.tv[b];
    return (qx & 1) ? w.th[b] : 0.0;
}

// The same two lookups with the fine point given RELATIVE to the coarse point (Z,X) that a thread owns:
// fine point (2Z + fz, 2X + fx), coarse point (Z + jz, X + jx).  In fully unrolled loops all four
// offsets are compile-time constants, so the geometry tests fold away and only the loads of the
// structurally non-zero weights remain (the Galerkin kernel went from 3.9 to ... ms with this).
__device__ __forceinline__ double pweight_rel(const Weights &w, int nzc, int nxc, int Z, int X, int fz, int fx, int jz,
                                              int jx)
{
    const int dz = fz - 2 * jz, dx = fx - 2 * jx;
    if (dz < -1 || dz > 1 || dx < -1 || dx > 1) return 0.0;
    const int Jz = Z + jz, Jx = X + jx;
    if (Jz < 0 || Jx < 0 || Jz >= nzc || Jx >= nxc) return 0.0;
    if (dz == 0 && dx == 0) return 1.0;
    // block that owns the fine point: floor(fz / 2), floor(fx / 2) blocks away from (Z, X)
    const int boz = fz >= 0 ? fz / 2 : -((1 - fz) / 2), box = fx >= 0 ? fx / 2 : -((1 - fx) / 2);
    const size_t b = (size_t)(Z + boz) * nxc + (X + box);
    if (dz == 0) return dx == 1 ? w.hw[b] : w.he[b];
    if (dx == 0) return dz == 1 ? w.vs[b] : w.vn[b];
    if (dz == 1) return dx == 1 ? w.csw[b] : w.cse[b];
    return dx == 1 ? w.cnw[b] : w.cne[b];
}
__device__ __forceinline__ double pdeficit_rel(const Weights &w, int nxc, int Z, int X, int fz, int fx)
{
    const int boz = fz >= 0 ? fz / 2 : -((1 - fz) / 2), box = fx >= 0 ? fx / 2 : -((1 - fx) / 2);
    const size_t b = (size_t)(Z + boz) * nxc + (X + box);
    if (fz & 1) return (fx & 1) ? w.tc[b] : w.tv[b];
    return (fx & 1) ? w.th[b] : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Deterministic grid reduction with an in-kernel finalizer ("last block done").  Every block
// contributes NV partial sums; the block that arrives last adds them up in a fixed order and runs
// fin(total[]) on one thread.  gridDim.x must be <= kMaxPartials.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem /* [NV][32] */)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double a = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) smem[k * 32 + warp] = a;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double a = lane < nwarps ? smem[k * 32 + lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            v[k] = a;
        }
    }
    __syncthreads();
}

template <int NV, class Fin>
__device__ __forceinline__ void grid_sum_finalize(double (&v)[NV], double *partials, unsigned int *counter, Fin fin)
{
    __shared__ double smem[NV * 32];
    __shared__ bool is_last;
    block_sum<NV>(v, smem);
    const int nblocks = gridDim.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) partials[k * kMaxPartials + blockIdx.x] = v[k];
        __threadfence();
        const unsigned int t = atomicInc(counter, (unsigned int)(nblocks - 1));
        is_last = (t == (unsigned int)(nblocks - 1));
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) {
            acc[k] = 0.0;
            for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc[k] += ldcg(&partials[k * kMaxPartials + i]);
        }
        block_sum<NV>(acc, smem);
        if (threadIdx.x == 0) fin(acc);
    }
}

// same pattern for a maximum (values >= 0)
template <class Fin>
__device__ __forceinline__ void grid_max_finalize(double v, double *partials, unsigned int *counter, Fin fin)
{
    __shared__ double smem[32];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
    auto block_max = [&](double a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        if (lane == 0) smem[warp] = a;
        __syncthreads();
        a = lane < nwarps ? smem[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        __syncthreads();
        return a;
    };
    v = block_max(v);
    const int nblocks = gridDim.x;
    if (tid == 0) {
        partials[blockIdx.x] = v;
        __threadfence();
        const unsigned int t = atomicInc(counter, (unsigned int)(nblocks - 1));
        is_last = (t == (unsigned int)(nblocks - 1));
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double a = 0.0;
        for (int i = tid; i < nblocks; i += blockDim.x) a = fmax(a, ldcg(&partials[i]));
        a = block_max(a);
        if (tid == 0) fin(a);
    }
}

// tile iteration: 1-D grid, each block walks 32x8 tiles of the rows [zlo, zhi) of a grid nx wide (one cell per
// thread; cells with z >= zhi or x >= nx are the caller's to skip)
struct Tiles {
    int tiles_x, zlo;
    long long ntiles;
    __device__ __forceinline__ Tiles(int zlo_, int zhi, int nx) : zlo(zlo_)
    {
        tiles_x = (nx + kTileX - 1) / kTileX;
        ntiles = (long long)tiles_x * ((zhi - zlo + kTileZ - 1) / kTileZ);
    }
    __device__ __forceinline__ void cell(long long t, int &z, int &x) const
    {
        z = zlo + (int)(t / tiles_x) * kTileZ + (threadIdx.x >> 5);
        x = (int)(t % tiles_x) * kTileX + (threadIdx.x & 31);
    }
};

// grid for a kernel that walks `nz` rows (the rows a rank owns) of a level nx wide
inline int grid_for(const Solver &s, int nz, int nx, int blocks_per_sm = 8)
{
    long long tiles = (long long)div_up(nx, kTileX) * div_up(nz, kTileZ);
    long long cap = (long long)s.num_sms * blocks_per_sm;
    if (cap > kMaxPartials) cap = kMaxPartials;
    return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

}   // namespace lcl
