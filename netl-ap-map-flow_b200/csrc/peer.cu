// peer.cu -- the slab every device array of a solver lives in, its export to the other ranks of a striped
// solve (CUDA IPC), and the two kernels that move rows between ranks by peer stores: k_halo_exchange
// (ghost rows to the two neighbours) and k_row_broadcast (own rows of a replicated array to everyone).
//
// New work (the reference, apmapflow/src/APERTURE_MAP_FLOW.F, is one serial process).  The
// north star names NCCL send/recv for the 1-cell halo; a PCG + multigrid iteration has ~30 such points, each
// a few hundred KB at most, so what matters is latency: a peer store plus a flag costs a few microseconds
// inside the stream, an NCCL call an order of magnitude more, and both ride the same NVLink.
#include <cstdlib>
#include <cstring>
#include <string>

#include "dist.cuh"

#include <unistd.h>
#ifdef LCL_CPU_SIM
#include <fcntl.h>
#include <sys/mman.h>
#endif

namespace lcl {

// ------------------------------------------------------------------------------------------ slab
#ifdef LCL_CPU_SIM
// TEST-ONLY emulation: the slab is POSIX shared memory, the "IPC handle" its name
static std::string g_shm_name[64];
static int g_shm_count = 0;

void slab_create(Solver &s, size_t bytes, void *handle_out)
{
    s.slab_bytes = bytes;
    if (s.world == 1) {
        s.slab = (char *)std::calloc(1, bytes);
        if (!s.slab) throw CudaError{"out of memory"};
        return;
    }
    char name[kPeerHandleBytes];
    snprintf(name, sizeof name, "/apmlcl_sim_%d_%d_%d", (int)getpid(), s.rank, g_shm_count++);
    const int fd = shm_open(name, O_CREAT | O_RDWR | O_EXCL, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) throw CudaError{"shm_open failed"};
    s.slab = (char *)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (s.slab == (char *)MAP_FAILED) throw CudaError{"mmap failed"};
    std::memset(s.slab, 0, bytes);
    g_shm_name[s.rank % 64] = name;
    std::memset(handle_out, 0, kPeerHandleBytes);
    std::memcpy(handle_out, name, strlen(name) + 1);
}

static void *peer_open(Solver &s, const void *handle, int)
{
    const int fd = shm_open((const char *)handle, O_RDWR, 0600);
    if (fd < 0) throw CudaError{"shm_open (peer) failed"};
    void *p = mmap(nullptr, s.slab_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) throw CudaError{"mmap (peer) failed"};
    return p;
}
static void peer_close(Solver &s, void *p, int) { munmap(p, s.slab_bytes); }

void slab_destroy(Solver &s)
{
    if (!s.slab) return;
    if (s.world == 1) std::free(s.slab);
    else {
        munmap(s.slab, s.slab_bytes);
        shm_unlink(g_shm_name[s.rank % 64].c_str());
    }
    s.slab = nullptr;
}
#else
// what a rank publishes: the CUDA IPC handle of its slab for ranks in other processes, and the plain address
// for ranks that live in the same process (several solvers driven by threads of one host program: an IPC
// handle cannot be opened by the process that exported it)
struct PeerHandle {
    cudaIpcMemHandle_t ipc;
    int device;
    int pid;
    void *address;
};
static_assert(sizeof(PeerHandle) <= kPeerHandleBytes, "handle blob too small");

void slab_create(Solver &s, size_t bytes, void *handle_out)
{
    s.slab_bytes = bytes;
    CUDA_CHECK(cudaMalloc((void **)&s.slab, bytes));
    CUDA_CHECK(cudaMemsetAsync(s.slab, 0, bytes, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    if (s.world == 1) return;
    PeerHandle h{};
    CUDA_CHECK(cudaIpcGetMemHandle(&h.ipc, s.slab));
    h.device = s.device;
    h.pid = (int)getpid();
    h.address = s.slab;
    std::memset(handle_out, 0, kPeerHandleBytes);
    std::memcpy(handle_out, &h, sizeof h);
}

static void *peer_open(Solver &s, const void *handle, int r)
{
    PeerHandle h;
    std::memcpy(&h, handle, sizeof h);
    const bool same_process = h.pid == (int)getpid();
    s.peer_ipc[r] = !same_process;
    if (!same_process) {
        // another process (one per GPU under torchrun): map its slab; peer access is enabled on the way
        void *p = nullptr;
        CUDA_CHECK(cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
        return p;
    }
    if (h.device != s.device) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, s.device, h.device) != cudaSuccess || !can) {
            cudaGetLastError();
            throw CudaError{"striped solve needs peer access (NVLink / PCIe P2P) between the GPUs of the ranks"};
        }
        const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(e);
        cudaGetLastError();
    }
    return h.address;
}
static void peer_close(Solver &s, void *p, int r)
{
    if (s.peer_ipc[r]) cudaIpcCloseMemHandle(p);
}

void slab_destroy(Solver &s)
{
    if (s.slab) cudaFree(s.slab);
    s.slab = nullptr;
}
#endif

void slab_connect(Solver &s, const void *handles)
{
    // the window sits at the start of every slab, the local bookkeeping right behind it (layout in api.cu)
    s.dist.rank = s.rank;
    s.dist.world = s.world;
    for (int r = 0; r < s.world; r++) {
        if (r == s.rank) {
            s.peer_slab[r] = s.slab;
        } else {
            if (!handles) throw CudaError{"peer handles are NULL"};
            s.peer_slab[r] = peer_open(s, (const char *)handles + (size_t)r * kPeerHandleBytes, r);
        }
        s.dist.win[r] = (PeerWindow *)s.peer_slab[r];
        s.dist.peer_off[r] = (long long)((char *)s.peer_slab[r] - s.slab);
    }
    unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;   // a peer that is 20 s late is gone
    if (const char *e = getenv("APM_LCL_PEER_TIMEOUT_MS")) timeout_ns = (unsigned long long)atoll(e) * 1000 * 1000;
    CUDA_CHECK(cudaMemcpyAsync(&s.dist.local->timeout_ns, &timeout_ns, sizeof timeout_ns, cudaMemcpyHostToDevice, s.stream));
    CUDA_CHECK(cudaStreamSynchronize(s.stream));
    s.connected = true;
}

// what the host says when a wait timed out: which wait, on whom, and where both sides' counters stand
std::string peer_error_report(Solver &s)
{
    if (s.world == 1) return "";
    DistLocal loc{};
    PeerWindow win{};
    cudaMemcpy(&loc, s.dist.local, sizeof loc, cudaMemcpyDeviceToHost);
    cudaMemcpy(&win, s.dist.win[s.rank], sizeof win, cudaMemcpyDeviceToHost);
    if (!loc.error) return "";
    static const char *kinds[] = {"?", "halo exchange", "gather of partial sums", "row broadcast"};
    char buf[512];
    snprintf(buf, sizeof buf,
             "striped solve: rank %d of %d gave up waiting in a %s for rank %d (flag at %llu, expected %llu); own counts: "
             "halo %llu gather %llu broadcast %llu; flags seen: halo below %llu above %llu",
             s.rank, s.world, kinds[loc.error >= 1 && loc.error <= 3 ? loc.error : 0], loc.error_peer, loc.error_have,
             loc.error_want, loc.halo_seq, loc.gather_seq, loc.bcast_seq, win.halo_flag[0], win.halo_flag[1]);
    std::string msg = buf;
    msg += "; gather flags";
    for (int r = 0; r < s.world; r++) msg += " " + std::to_string(win.gather_flag[r]);
    msg += "; broadcast flags";
    for (int r = 0; r < s.world; r++) msg += " " + std::to_string(win.bcast_flag[r]);
    return msg;
}

void slab_disconnect(Solver &s)
{
    for (int r = 0; r < s.world; r++) {
        if (r != s.rank && s.peer_slab[r]) peer_close(s, s.peer_slab[r], r);
        s.peer_slab[r] = nullptr;
    }
    s.connected = false;
}

// ------------------------------------------------------------------------------------- exchanges
__device__ __forceinline__ bool xchg_stopped(const PcgState *st)
{
#ifndef LCL_CPU_SIM
    cudaGridDependencySynchronize();
#endif
    return st && (st->converged | st->failed);
}

// rows [z0, z1) of every array of the set, from this rank's copy into rank `to`'s copy.  `shift` = elements
// between the two ranks' virtual bases beyond the slab offset (0 for replicated arrays).
__device__ __forceinline__ void push_rows(const Dist &dc, const HaloSet &hs, int to, long long shift, int pitch, int z0, int z1)
{
    const long long per = (long long)(z1 - z0) * pitch, total = per * hs.count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(i / per);
        const long long e = (long long)z0 * pitch + i % per;
        const double *src = hs.a[a] + e;
        double *dst = (double *)((char *)hs.a[a] + dc.peer_off[to]) + shift + e;
        *dst = *src;
    }
}

// "last block done": true in exactly one block per launch, after every block's stores are visible system wide
__device__ __forceinline__ bool last_block(unsigned int *counter)
{
    __shared__ bool last;
    fence_sys();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicInc(counter, gridDim.x - 1);
        last = t == gridDim.x - 1;
    }
    __syncthreads();
    return last;
}

// zlo_me / zlo_dn / zlo_up: first owned row of this rank and of its neighbours (the virtual bases of striped
// arrays differ by (zlo_me - zlo_peer) * pitch elements on top of the slab offset)
__global__ void __launch_bounds__(256)
k_halo_exchange(Dist dc, HaloSet hs, int pitch, int zlo, int zhi, int zlo_dn, int zlo_up, int dlo, int dhi, const PcgState *st)
{
    if (xchg_stopped(st)) return;
    const bool dn = dc.rank > 0, up = dc.rank + 1 < dc.world;
    // my lowest dhi rows are the upper ghosts of the rank below; my highest dlo rows the lower ghosts of the rank above
    if (dn) push_rows(dc, hs, dc.rank - 1, (long long)(zlo - zlo_dn) * pitch, pitch, zlo, zlo + dhi < zhi ? zlo + dhi : zhi);
    if (up) push_rows(dc, hs, dc.rank + 1, (long long)(zlo - zlo_up) * pitch, pitch, zhi - dlo > zlo ? zhi - dlo : zlo, zhi);
    if (!last_block(&dc.local->block_counter[0])) return;
    if (threadIdx.x == 0) {
        const unsigned long long seq = ++dc.local->halo_seq;
        fence_sys();
        if (dn) st_release_sys(&dc.win[dc.rank - 1]->halo_flag[1], seq);   // I am the rank above my lower neighbour
        if (up) st_release_sys(&dc.win[dc.rank + 1]->halo_flag[0], seq);
        PeerWindow *self = dc.win[dc.rank];
        if (dn) wait_flag(&self->halo_flag[0], seq, dc, 1, dc.rank - 1);
        if (up) wait_flag(&self->halo_flag[1], seq, dc, 1, dc.rank + 1);
        fence_sys();
    }
}

// The targets are whole arrays every rank keeps reading until it gets here itself (the right-hand side of the
// first whole level is read all the way through that level's cycle), and ranks that are not neighbours can be
// several exchange points apart.  So the broadcast is two-phase: nobody stores before EVERY rank has arrived
// ("ready", count 2k+1), and nobody leaves before every rank's rows have landed ("done", count 2k+2).
__global__ void __launch_bounds__(256)
k_row_broadcast(Dist dc, HaloSet hs, int pitch, int z0, int z1, const PcgState *st)
{
    if (xchg_stopped(st)) return;
    PeerWindow *self = dc.win[dc.rank];
    const unsigned long long seq = dc.local->bcast_seq;   // advanced by the last block, after every block has read it
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            fence_sys();
            for (int r = 0; r < dc.world; r++)
                if (r != dc.rank) st_release_sys(&dc.win[r]->bcast_flag[dc.rank], seq + 1);
        }
        for (int r = 0; r < dc.world; r++)
            if (r != dc.rank) wait_flag(&self->bcast_flag[r], seq + 1, dc, 3, r);
        fence_sys();
    }
    __syncthreads();
    for (int k = 1; k < dc.world; k++) push_rows(dc, hs, (dc.rank + k) % dc.world, 0, pitch, z0, z1);
    if (!last_block(&dc.local->block_counter[1])) return;
    if (threadIdx.x == 0) {
        dc.local->bcast_seq = seq + 2;
        fence_sys();
        for (int r = 0; r < dc.world; r++)
            if (r != dc.rank) st_release_sys(&dc.win[r]->bcast_flag[dc.rank], seq + 2);
        for (int r = 0; r < dc.world; r++)
            if (r != dc.rank) wait_flag(&self->bcast_flag[r], seq + 2, dc, 3, r);
        fence_sys();
    }
}

static int xchg_grid(const Solver &s, long long elements)
{
    long long g = (elements + 4 * 256 - 1) / (4 * 256);   // a few elements per thread; small exchanges stay one block
    if (g < 1) g = 1;
    if (g > s.num_sms) g = s.num_sms;
    return (int)g;
}

void halo_exchange(Solver &s, const HaloSet &hs, int pitch, const int *cut, int dlo, int dhi, const PcgState *st)
{
    if (s.world == 1 || hs.count == 0) return;
    const int zlo = cut[s.rank], zhi = cut[s.rank + 1];
    const int zlo_dn = s.rank > 0 ? cut[s.rank - 1] : 0, zlo_up = s.rank + 1 < s.world ? cut[s.rank + 1] : 0;
    const long long elements = (long long)(dlo + dhi) * pitch * hs.count;
    LAUNCH_DEP(k_halo_exchange, xchg_grid(s, elements), 256, 0, s.stream, s.dist, hs, pitch, zlo, zhi, zlo_dn, zlo_up, dlo,
               dhi, st);
}

void row_broadcast(Solver &s, const HaloSet &hs, int pitch, const int *cut, const PcgState *st)
{
    if (s.world == 1 || hs.count == 0) return;
    const int z0 = cut[s.rank], z1 = cut[s.rank + 1];
    const long long elements = (long long)(z1 - z0) * pitch * hs.count * (s.world - 1);
    LAUNCH_DEP(k_row_broadcast, xchg_grid(s, elements), 256, 0, s.stream, s.dist, hs, pitch, z0, z1, st);
}

// forces the kernels of this file into the device's context now (see preload_kernels, api.cu)
void preload_peer()
{
    preload_one(k_halo_exchange);
    preload_one(k_row_broadcast);
}

}   // namespace lcl
