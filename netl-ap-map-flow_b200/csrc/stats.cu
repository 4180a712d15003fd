// stats.cu -- the map statistics and percentile tables of the reference on the device.
//
// Replaces, for the file pipeline (host_io.cpp) and for callers that hold fields in HBM:
//   MAP_STATS        apmapflow/src/MAP_MODULE.F:182-303     mean, RMS, standard deviation and the x / z correlation
//                                                           lengths from the lag variograms (first lag whose
//                                                           semivariance passes 0.9 of the variance, :246-266)
//   CALC_PERCENTILE  apmapflow/src/OUTPUT_MODULE.F:53-124   sorts the field, collapses duplicates and walks the
//                                                           cumulative counts: percentile p is the order
//                                                           statistic s[min(ceil(p/100 n), n-1)]
// once the solve takes milliseconds these are the next host hotspot -- the reference's own input
// templates ask for 19 percentiles of six fields, i.e. six sorts of N doubles (8192^2: ~6 s each on one core).
// Here a percentile is found by an exact radix select (eight 8-bit passes over the order-preserving integer image
// of the doubles, all requested percentiles at once), a variogram batch of eight lags is one pass over the map.
// Both are HBM bound: 8 B/value per select pass, 8 B/cell per eight lags.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "device_utils.cuh"

namespace lcl {

namespace {

struct DevBuf {
    void *p = nullptr;
    explicit DevBuf(size_t bytes) { CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 8)); }
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    template <class T> T *as() const { return (T *)p; }
};

// APM_LCL_STATS_TIMING=1: wall time of the stages of the two entry points on stderr
struct Lap {
    bool on = getenv("APM_LCL_STATS_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void operator()(const char *what)
    {
        if (!on) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[apm_lcl stats] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

bool on_device(const void *p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Small per-device scratch (reduction partials, "last block done" counter, results, select histograms), allocated
// once per device and kept for the life of the process: cudaMalloc / cudaFree of a few bytes cost 0.1 - 0.8 s each in
// a process whose allocator (torch's) already holds gigabytes, against 4 ms for the kernels of a whole call.  One
// caller at a time per device (the mutex): these are a few milliseconds.
constexpr int kMaxGroups = 32;
struct Scratch {
    double *part = nullptr, *result = nullptr;
    unsigned int *counter = nullptr;
    unsigned long long *hist = nullptr;
    std::mutex busy;
};
Scratch &scratch_for(int device)
{
    static std::mutex guard;
    static Scratch pool[64];
    std::lock_guard<std::mutex> lock(guard);
    Scratch &sc = pool[device & 63];
    if (!sc.part) {
        CUDA_CHECK(cudaMalloc((void **)&sc.part, (size_t)8 * kMaxPartials * sizeof(double)));
        CUDA_CHECK(cudaMalloc((void **)&sc.result, 8 * sizeof(double)));
        CUDA_CHECK(cudaMalloc((void **)&sc.counter, sizeof(unsigned int)));
        CUDA_CHECK(cudaMalloc((void **)&sc.hist, (size_t)kMaxGroups * 256 * sizeof(unsigned long long)));
        CUDA_CHECK(cudaMemset(sc.counter, 0, sizeof(unsigned int)));
    }
    return sc;
}

}   // namespace

// ----------------------------------------------------------------------------------- moments
// pass 1: sum v; pass 2 (avg known): sum (v - avg), sum (v - avg)^2, sum v^2      (MAP_MODULE.F:224-240)
__global__ void __launch_bounds__(256)
k_stats_moments(const double *__restrict__ v, size_t n, double avg, int pass, double *partials, unsigned int *counter,
                double *out)
{
    double acc[3] = {0.0, 0.0, 0.0};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double a = v[i];
        if (pass == 0) {
            acc[0] += a;
        } else {
            const double d = a - avg;
            acc[0] += d;
            acc[1] += d * d;
            acc[2] += a * a;
        }
    }
    grid_sum_finalize<3>(acc, partials, counter, [=](const double(&tot)[3]) {
        out[0] = tot[0];
        out[1] = tot[1];
        out[2] = tot[2];
    });
}

// sum over all valid pairs of (a(z, x) - a(z, x + k))^2 (dir 0) or (a(z, x) - a(z + k, x))^2 (dir 1) for the lags
// k0 .. k0 + nk - 1 (nk <= 8) in one pass over the map                                 (MAP_MODULE.F:246-254, 271-279)
constexpr int kLagBatch = 8;
__global__ void __launch_bounds__(kBlock)
k_lag_ssd(const double *__restrict__ a, int nz, int nx, int dir, int k0, int nk, double *partials, unsigned int *counter,
          double *out)
{
    const Tiles tiles(0, nz, nx);
    double acc[kLagBatch] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
        int z, x;
        tiles.cell(t, z, x);
        if (z >= nz || x >= nx) continue;
        const size_t i = (size_t)z * nx + x;
        const double ai = a[i];
#pragma unroll
        for (int m = 0; m < kLagBatch; m++) {
            const int k = k0 + m;
            if (m < nk && (dir == 0 ? x + k < nx : z + k < nz)) {
                const double d = ai - a[dir == 0 ? i + k : i + (size_t)k * nx];
                acc[m] += d * d;
            }
        }
    }
    grid_sum_finalize<kLagBatch>(acc, partials, counter, [=](const double(&tot)[kLagBatch]) {
        for (int m = 0; m < kLagBatch; m++) out[m] = tot[m];
    });
}

// ---------------------------------------------------------------------------------- percentiles
// order-preserving image of a double in the unsigned integers
__host__ __device__ inline unsigned long long key_of(double v)
{
    unsigned long long u;
#ifdef __CUDA_ARCH__
    u = (unsigned long long)__double_as_longlong(v);
#else
    std::memcpy(&u, &v, sizeof u);
#endif
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
inline double value_of(unsigned long long k)
{
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
    std::memcpy(&v, &u, sizeof v);
    return v;
}

// One pass of the radix select: for every group g (a key prefix of `done` bits that some requested order statistic
// is known to lie under) the histogram of the next 8 bits over the values that carry that prefix.
struct SelectGroups {
    unsigned long long prefix[kMaxGroups];
    int count;
};
__global__ void __launch_bounds__(256)
k_select_hist(const double *__restrict__ v, size_t n, SelectGroups g, int done, unsigned long long *__restrict__ hist)
{
    LCL_DYNAMIC_SMEM(unsigned int, sh);   // [count][256]
    for (int i = threadIdx.x; i < g.count * 256; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const int shift = 64 - done - 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long key = key_of(v[i]);
        const unsigned long long top = done ? key >> (64 - done) : 0ull;
        const unsigned int digit = (unsigned int)(key >> shift) & 255u;
        for (int j = 0; j < g.count; j++)
            if (top == g.prefix[j]) {
#ifdef LCL_CPU_SIM
                sh[j * 256 + digit] += 1u;
#else
                atomicAdd(&sh[j * 256 + digit], 1u);
#endif
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < g.count * 256; i += blockDim.x)
        if (sh[i]) {
#ifdef LCL_CPU_SIM
            hist[i] += sh[i];
#else
            atomicAdd(&hist[i], (unsigned long long)sh[i]);
#endif
        }
}

void preload_stats()
{
    preload_one(k_stats_moments);
    preload_one(k_lag_ssd);
    preload_one(k_select_hist);
}

}   // namespace lcl

using namespace lcl;

extern "C" {

int apm_lcl_map_stats(const double *map, int nx, int nz, int device, double *avg_out, double *rms, double *dev,
                      double *x_corr, double *z_corr)
{
    try {
        if (!map || nx < 1 || nz < 1) throw CudaError{"apm_lcl_map_stats: bad arguments"};
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            cudaGetLastError();
            throw CudaError{"no CUDA device available: libapmlcl has no CPU fallback"};
        }
        CUDA_CHECK(cudaSetDevice(device));
        Lap lap;
        const size_t n = (size_t)nx * nz;
        const double cells = (double)nx * nz;
        std::unique_ptr<DevBuf> staged;
        const double *a = map;
        if (!on_device(map)) {
            staged.reset(new DevBuf(n * sizeof(double)));
            CUDA_CHECK(cudaMemcpy(staged->p, map, n * sizeof(double), cudaMemcpyHostToDevice));
            a = staged->as<double>();
        }
        Scratch &sc = scratch_for(device);
        std::lock_guard<std::mutex> lock(sc.busy);
        lap("map_stats: staging");
        int sms = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        const int grid = std::max(1, std::min((int)((n + 255) / 256), sms * 8));
        lap("map_stats: attributes");
        double h[8];
        LAUNCH(k_stats_moments, grid, 256, 0, 0, a, n, 0.0, 0, sc.part, sc.counter, sc.result);
        CUDA_CHECK(cudaMemcpy(h, sc.result, 3 * sizeof(double), cudaMemcpyDeviceToHost));
        const double avg = h[0] / cells;                                               // MAP_MODULE.F:222
        LAUNCH(k_stats_moments, grid, 256, 0, 0, a, n, avg, 1, sc.part, sc.counter, sc.result);
        CUDA_CHECK(cudaMemcpy(h, sc.result, 3 * sizeof(double), cudaMemcpyDeviceToHost));
        const double sm = (h[0] * h[0]) * (1.0 / cells);                               // :237
        const double var = (h[1] - sm) * (1.0 / (cells - 1));                          // :239
        if (avg_out) *avg_out = avg;
        if (rms) *rms = std::sqrt(h[2] * (1.0 / cells));                               // :238, 240
        if (dev) *dev = std::sqrt(var);
        // variograms: lags 1 .. INT(min(nx, nz) / 2) + 1, stop at the first lag whose semivariance passes 0.9 var
        const int kmax = (int)((double)std::min(nx, nz) / 2.0) + 1;
        lap("map_stats: moments");
        Solver fake;
        fake.num_sms = sms;
        const int g2 = grid_for(fake, nz, nx, 8);
        for (int dir = 0; dir < 2; dir++) {
            double t1 = 0, t2 = 0, t3 = 0, c = 0;
            bool stop = false;
            for (int k0 = 1; k0 <= kmax && !stop; k0 += kLagBatch) {
                const int nk = std::min(kLagBatch, kmax - k0 + 1);
                LAUNCH(k_lag_ssd, g2, kBlock, 0, 0, a, nz, nx, dir, k0, nk, sc.part, sc.counter, sc.result);
                CUDA_CHECK(cudaMemcpy(h, sc.result, kLagBatch * sizeof(double), cudaMemcpyDeviceToHost));
                for (int m = 0; m < nk; m++) {
                    const int k = k0 + m;
                    const int zmax = dir == 0 ? nz : nz - k, xmax = dir == 0 ? nx - k : nx;
                    const double cnt = zmax > 0 && xmax > 0 ? (double)zmax * (double)xmax : 0.0;
                    if (cnt == 0) continue;                                             // :255
                    const double v3 = h[m] / cnt;
                    c = (double)k;
                    t1 = std::fabs(v3 / var - 0.90);
                    if (v3 / var > 0.90) {                                              // :259-262
                        if (t2 < t1) c = t3;
                        stop = true;
                        break;
                    }
                    t2 = t1;
                    t3 = c;
                }
            }
            if (dir == 0 && x_corr) *x_corr = c;
            if (dir == 1 && z_corr) *z_corr = c;
            lap("map_stats: variogram");
        }
        g_last_error.clear();
        return 0;
    } catch (const CudaError &e) {
        g_last_error = e.msg;
        return 1;
    } catch (...) {
        g_last_error = "unknown error";
        return 1;
    }
}

int apm_lcl_percentiles(const double *values, size_t n, const double *pct, int npct, int device, double *out)
{
    try {
        if (!values || !pct || !out || n == 0 || npct < 1) throw CudaError{"apm_lcl_percentiles: bad arguments"};
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            cudaGetLastError();
            throw CudaError{"no CUDA device available: libapmlcl has no CPU fallback"};
        }
        CUDA_CHECK(cudaSetDevice(device));
        Lap lap;
        std::unique_ptr<DevBuf> staged;
        const double *v = values;
        if (!on_device(values)) {
            staged.reset(new DevBuf(n * sizeof(double)));
            CUDA_CHECK(cudaMemcpy(staged->p, values, n * sizeof(double), cudaMemcpyHostToDevice));
            v = staged->as<double>();
        }
        Scratch &sc = scratch_for(device);
        std::lock_guard<std::mutex> lock(sc.busy);
        lap("percentiles: staging");
        int sms = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        const int grid = std::max(1, std::min((int)((n + 255) / 256), sms * 8));
        lap("percentiles: attributes");
        // percentile p -> 0-based order statistic min(ceil(p / 100 n), n - 1)          (OUTPUT_MODULE.F:113-121)
        std::vector<unsigned long long> rank(npct), prefix(npct, 0ull);
        for (int j = 0; j < npct; j++) {
            const double kk = std::ceil(pct[j] / 100.0 * (double)n);
            rank[j] = kk <= 0 ? 0ull : (kk >= (double)n ? (unsigned long long)(n - 1) : (unsigned long long)kk);
        }
        std::vector<unsigned long long> h((size_t)kMaxGroups * 256);
        for (int done = 0; done < 64; done += 8) {
            // the distinct prefixes still alive (several percentiles often share one); at most kMaxGroups per pass
            for (int first = 0; first < npct;) {
                SelectGroups g{};
                std::vector<int> group_of(npct, -1);
                int j = first;
                for (; j < npct; j++) {
                    int found = -1;
                    for (int q = 0; q < g.count; q++)
                        if (g.prefix[q] == prefix[j]) found = q;
                    if (found < 0) {
                        if (g.count == kMaxGroups) break;
                        g.prefix[g.count] = prefix[j];
                        found = g.count++;
                    }
                    group_of[j] = found;
                }
                CUDA_CHECK(cudaMemset(sc.hist, 0, (size_t)g.count * 256 * sizeof(unsigned long long)));
                LAUNCH(k_select_hist, grid, 256, (size_t)g.count * 256 * sizeof(unsigned int), 0, v, n, g, done, sc.hist);
                CUDA_CHECK(cudaMemcpy(h.data(), sc.hist, (size_t)g.count * 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                for (int q = first; q < j; q++) {
                    const unsigned long long *hq = &h[(size_t)group_of[q] * 256];
                    unsigned long long cum = 0;
                    int d = 0;
                    for (; d < 255; d++) {
                        if (cum + hq[d] > rank[q]) break;
                        cum += hq[d];
                    }
                    rank[q] -= cum;
                    prefix[q] = (prefix[q] << 8) | (unsigned long long)d;
                }
                first = j;
            }
            lap("percentiles: select pass");
        }
        for (int j = 0; j < npct; j++) out[j] = value_of(prefix[j]);
        g_last_error.clear();
        return 0;
    } catch (const CudaError &e) {
        g_last_error = e.msg;
        return 1;
    } catch (...) {
        g_last_error = "unknown error";
        return 1;
    }
}

}   // extern "C"
