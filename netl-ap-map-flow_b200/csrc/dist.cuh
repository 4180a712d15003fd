// dist.cuh -- device side of the striped (multi-GPU) solve: flags and small gathers over peer memory.
//
// New work: the reference is a single-process serial program (apmapflow/src/APERTURE_MAP_FLOW.F); SURVEY 8e
// asks for row strips with a halo exchange per operator application and an all-reduce per dot product.
// Both are done by the kernels themselves: the block that finishes a grid reduction stores its rank's
// partial sums straight into every peer's window, raises a flag there and waits for the flags of the
// others (dist_gather); ghost rows are stored into the neighbours' arrays by k_halo_exchange (peer.cu).
// No NCCL call sits on the data path; torch.distributed only carries the IPC handles at start-up.
#pragma once

#include "common.cuh"

#ifdef LCL_CPU_SIM
#include <sched.h>
#endif

namespace lcl {

#ifdef LCL_CPU_SIM
#define LCL_NOINLINE
#else
#define LCL_NOINLINE __noinline__
#endif

#ifdef LCL_CPU_SIM
// the CPU emulation runs each rank in its own process on a POSIX shared-memory slab
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
__device__ __forceinline__ void fence_sys() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
__device__ __forceinline__ double ld_volatile(const double *p) { return *(const volatile double *)p; }
__device__ __forceinline__ void st_volatile(double *p, double v) { *(volatile double *)p = v; }
__device__ __forceinline__ unsigned long long now_ns()
{
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}
__device__ __forceinline__ void spin_pause() { sched_yield(); }
#else
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_sys() { __threadfence_system(); }
__device__ __forceinline__ double ld_volatile(const double *p)
{
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile(double *p, double v)
{
    asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void spin_pause() { __nanosleep(64); }
#endif

// Wait until the local flag *p (raised by rank `peer`) has reached `want`.  Bounded (DistLocal::timeout_ns, 20 s
// unless APM_LCL_PEER_TIMEOUT_MS says otherwise): on time-out the rank records which wait it was and carries on
// with whatever it has, every later wait returns at once, and the host reports the failure.
__device__ __forceinline__ bool wait_flag(const unsigned long long *p, unsigned long long want, const Dist &dc, int kind,
                                          int peer)
{
    if (ld_acquire_sys(p) >= want) return true;
    if (dc.local->error) return false;
    const unsigned long long t0 = now_ns();
    for (;;) {
        for (int k = 0; k < 64; k++) {
            if (ld_acquire_sys(p) >= want) return true;
            spin_pause();
        }
        if (now_ns() - t0 > dc.local->timeout_ns) {
            dc.local->error = kind;
            dc.local->error_peer = peer;
            dc.local->error_want = want;
            dc.local->error_have = ld_acquire_sys(p);
            return false;
        }
    }
}

// All ranks' NV values on every rank, in rank order (so every rank combines them in the same order and
// ends up with bitwise the same result -- the stopping decisions taken from them then agree).  Called by
// ONE thread of ONE block per rank, at the same point of the same kernel sequence on every rank.
// (not inlined: one thread runs it once per launch, and its rank-sized arrays stay out of the callers' frames)
template <int NV>
__device__ LCL_NOINLINE void dist_gather(const Dist &dc, const double (&mine)[NV], double (&all)[kMaxRanks][NV])
{
    static_assert(NV <= kGatherMax, "gather slot too small");
    if (dc.world == 1) {
#pragma unroll
        for (int k = 0; k < NV; k++) all[0][k] = mine[k];
        return;
    }
    const unsigned long long seq = ++dc.local->gather_seq;
    const int par = (int)(seq & 1);
    for (int r = 0; r < dc.world; r++) {
        double *slot = dc.win[r]->gather_slot[par][dc.rank];
#pragma unroll
        for (int k = 0; k < NV; k++) st_volatile(slot + k, mine[k]);
    }
    fence_sys();
    for (int r = 0; r < dc.world; r++) st_release_sys(&dc.win[r]->gather_flag[dc.rank], seq);
    PeerWindow *self = dc.win[dc.rank];
    for (int r = 0; r < dc.world; r++) wait_flag(&self->gather_flag[r], seq, dc, 2, r);
    fence_sys();
    for (int r = 0; r < dc.world; r++)
#pragma unroll
        for (int k = 0; k < NV; k++) all[r][k] = ld_volatile(&self->gather_slot[par][r][k]);
}

// sum over ranks, in rank order
template <int NV>
__device__ LCL_NOINLINE void dist_sum(const Dist &dc, double (&v)[NV])
{
    if (dc.world == 1) return;
    double all[kMaxRanks][NV];
    dist_gather<NV>(dc, v, all);
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double s = 0.0;
        for (int r = 0; r < dc.world; r++) s += all[r][k];
        v[k] = s;
    }
}

}   // namespace lcl
