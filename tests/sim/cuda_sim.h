// cuda_sim.h -- TEST INFRASTRUCTURE ONLY.  A tiny CUDA-on-CPU emulation used to debug the kernel
// sources of libapmlcl in the GPU-less authoring container (tests/sim/build_sim.py compiles
// csrc/*.cu with g++ and this header force-included, producing tests/sim/libapmlcl_sim.so).
//
// It is NOT a CPU fallback: the product binding (apmapflow_b200/_lib.py) only ever loads
// libapmlcl.so built by nvcc, and nothing outside tests/ references the sim library.  The sim runs
// every CUDA thread of a block as a ucontext fiber so __syncthreads / warp shuffles / the
// "last block done" reductions behave as on the device (blocks execute one after another).
#pragma once
#ifndef LCL_CPU_SIM
#define LCL_CPU_SIM 1
#endif

#include <ucontext.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() {}
    dim3(unsigned a) : x(a) {}
    dim3(int a) : x((unsigned)a) {}
};

// Fiber switch.  glibc's swapcontext saves and restores the signal mask with a system call on every switch; a kernel
// with a few barriers per thread spends nearly all of its emulated time there.  On x86-64 the switch is twenty lines of
// assembly instead (sim_globals.cpp: callee-saved registers and the stack pointer, nothing else changes between fibers).
#if defined(__x86_64__)
#define CUSIM_FAST_SWITCH 1
extern "C" void cusim_switch(void **save_sp, void *load_sp);
#endif

namespace cusim {
struct Fiber {
#ifdef CUSIM_FAST_SWITCH
    void *sp = nullptr;
#else
    ucontext_t ctx;
#endif
    char *stack = nullptr;
    bool done = true;
};
inline std::vector<Fiber> &fibers() { static std::vector<Fiber> f; return f; }
#ifdef CUSIM_FAST_SWITCH
inline void *&sched() { static void *sp = nullptr; return sp; }
#else
inline ucontext_t &sched() { static ucontext_t c; return c; }
#endif
inline int &cur() { static int c = 0; return c; }
inline int &nthreads() { static int n = 0; return n; }
inline int &bar_count() { static int n = 0; return n; }
inline int &bar_gen() { static int n = 0; return n; }
inline std::function<void()> *&body() { static std::function<void()> *b = nullptr; return b; }
inline char *&dyn_smem() { static char *p = nullptr; return p; }
struct WarpX {
    double buf[32];
    long long lbuf[32];
    int count = 0, gen = 0;
};
inline std::vector<WarpX> &warps() { static std::vector<WarpX> w; return w; }
constexpr size_t kStack = 256 * 1024;

#ifdef CUSIM_FAST_SWITCH
inline void yield() { cusim_switch(&fibers()[cur()].sp, sched()); }
inline void entry()
{
    (*body())();
    fibers()[cur()].done = true;
    cusim_switch(&fibers()[cur()].sp, sched());
    __builtin_trap();   // a finished fiber is never resumed
}
// a fresh fiber: six zeroed callee-saved registers, then the address `ret` jumps to, then a slot that keeps the stack
// 16-byte aligned at the entry of entry() the way a call would
inline void prepare(Fiber &f)
{
    void **top = (void **)(((uintptr_t)f.stack + kStack) & ~(uintptr_t)15);
    *--top = nullptr;
    *--top = (void *)entry;
    for (int k = 0; k < 6; k++) *--top = nullptr;
    f.sp = (void *)top;
    f.done = false;
}
inline void resume(Fiber &f) { cusim_switch(&sched(), f.sp); }
#else
inline void yield() { swapcontext(&fibers()[cur()].ctx, &sched()); }
inline void entry()
{
    (*body())();
    fibers()[cur()].done = true;
    swapcontext(&fibers()[cur()].ctx, &sched());
}
inline void prepare(Fiber &f)
{
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())entry, 0);
    f.done = false;
}
inline void resume(Fiber &f) { swapcontext(&sched(), &f.ctx); }
#endif
}   // namespace cusim

extern dim3 threadIdx, blockIdx, blockDim, gridDim;

inline void __syncthreads()
{
    const int gen = cusim::bar_gen();
    if (++cusim::bar_count() == cusim::nthreads()) {
        cusim::bar_count() = 0;
        cusim::bar_gen()++;
    } else {
        while (cusim::bar_gen() == gen) cusim::yield();
    }
}
inline void cusim_warp_barrier(int warp, int lanes)
{
    cusim::WarpX &w = cusim::warps()[warp];
    const int gen = w.gen;
    if (++w.count == lanes) {
        w.count = 0;
        w.gen++;
    } else {
        while (w.gen == gen) cusim::yield();
    }
}
inline int cusim_lanes(int warp)
{
    const int rem = cusim::nthreads() - warp * 32;
    return rem < 32 ? rem : 32;
}
inline double __shfl_xor_sync(unsigned, double v, int o)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lanes = cusim_lanes(warp);
    cusim::warps()[warp].buf[lane] = v;
    cusim_warp_barrier(warp, lanes);
    const double r = cusim::warps()[warp].buf[(lane ^ o) < lanes ? (lane ^ o) : lane];
    cusim_warp_barrier(warp, lanes);
    return r;
}
inline long long __shfl_xor_sync(unsigned, long long v, int o)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lanes = cusim_lanes(warp);
    cusim::warps()[warp].lbuf[lane] = v;
    cusim_warp_barrier(warp, lanes);
    const long long r = cusim::warps()[warp].lbuf[(lane ^ o) < lanes ? (lane ^ o) : lane];
    cusim_warp_barrier(warp, lanes);
    return r;
}
struct alignas(16) double2 { double x, y; };
inline double2 make_double2(double a, double b) { return double2{a, b}; }
inline double __ldcg(const double *p) { return *p; }
inline void __threadfence() {}
inline unsigned atomicInc(unsigned *a, unsigned lim)
{
    const unsigned old = *a;
    *a = old >= lim ? 0 : old + 1;
    return old;
}
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::atan; using std::fabs; using std::fmax; using std::fmin; using std::sqrt; using std::tan;

// ------------------------------------------------------------------------------ runtime API shim
typedef int cudaError_t;
typedef void *cudaStream_t;
struct cusimEvent { std::chrono::steady_clock::time_point t; };
typedef cusimEvent *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum { cudaStreamNonBlocking = 1 };
struct cudaPointerAttributes { int type = 0; };
struct cudaDeviceProp { int multiProcessorCount = 4; char name[64] = "cusim"; };
inline const char *cudaGetErrorString(cudaError_t) { return "cusim error"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDeviceCount(int *c) { *c = 1; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return 0; }
enum { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = cudaDeviceProp().multiProcessorCount; return 0; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = 0; return 0; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::malloc(n ? n : 1); std::memset(*p, 0xff, n); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = std::calloc(1, n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, int) { *s = (void *)1; return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cusimEvent(); return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return 0;
}

namespace cusim {
// run one kernel launch: blocks sequentially, the threads of a block as fibers
inline void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> fn)
{
    const int nt = (int)block.x;
    auto &fb = fibers();
    if ((int)fb.size() < nt) fb.resize(nt);
    for (int i = 0; i < nt; i++)
        if (!fb[i].stack) fb[i].stack = (char *)std::malloc(kStack);
    warps().assign((nt + 31) / 32, WarpX());
    std::vector<char> dyn(smem ? smem : 8);
    dyn_smem() = dyn.data();
    body() = &fn;
    nthreads() = nt;
    gridDim = grid;
    blockDim = block;
    for (unsigned b = 0; b < grid.x; b++) {
        bar_count() = 0;
        for (auto &w : warps()) w.count = 0;
        for (int i = 0; i < nt; i++) prepare(fb[i]);
        int remaining = nt;
        while (remaining > 0) {
            for (int i = 0; i < nt; i++) {
                if (fb[i].done) continue;
                cur() = i;
                threadIdx = dim3(i);
                blockIdx = dim3(b);
                resume(fb[i]);
                if (fb[i].done) remaining--;
            }
        }
    }
}
}   // namespace cusim

#define LCL_DYNAMIC_SMEM(type, name) type *name = (type *)::cusim::dyn_smem()
