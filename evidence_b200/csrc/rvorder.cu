// rvorder.cu — posterior planet-ordering on the device (part of librvlnl.so).
//
// Reference path replaced: evidence/post_processing.py:104-128 -- a pandas `iterrows` loop that,
// for every posterior sample whose planet periods are not non-decreasing, rebuilds an index list
// from np.argsort(periods) and gathers the row through it.  The gather uses the RANK of a column's
// own planet as the source slot (new[i] = old[planets[rank(p_i)][q_i]]), which is the inverse of
// the sorting permutation; reproduced as is (identical results on identical inputs).
//
// One block per TILE of 128 consecutive rows: the tile is one contiguous, 16-byte aligned piece of
// memory, read ONCE with coalesced 16-byte streaming loads into shared memory; one thread per row does
// the "already ordered" test (:113) and, for a row that fails it, permutes the planets' columns in
// place through the ranks; the tile is written back with coalesced 16-byte stores.  16 B of HBM
// traffic per element and ~1 instruction per element (the r1 kernel, one thread per element with its
// own index arithmetic and K period reads, spent 8 and reached 1.1-1.4 TB/s): the bound is HBM.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>

#include "../../include/rvlnl.h"
#include "rvpost.h"

namespace {

struct OrderTab {
    int K, Q;
    int period_col[RVL_FIP_MAX_PLANETS];
    int planet_col[RVL_FIP_MAX_PLANETS][RVL_ORDER_MAX_PARAMS];
};

// numpy's sort order: a before b  <=>  a < b, NaN after everything
__device__ __forceinline__ bool before(double a, double b)
{
    return (a == a) && ((b != b) || a < b);
}

constexpr int kTileRows = 128;  // even: every tile starts on a 16-byte boundary
constexpr int kOrderThreads = 256;

__global__ void __launch_bounds__(kOrderThreads)
order_planets_kernel(const double *in, long long n, int ndim, const OrderTab tab, double *out)
{
    extern __shared__ __align__(16) double tile[];  // [kTileRows][ndim]
    const int tid = threadIdx.x;
    const long long n_tiles = (n + kTileRows - 1) / kTileRows;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long row0 = t * kTileRows;
        const int rows = (int)min((long long)kTileRows, n - row0);
        const long long e0 = row0 * ndim;  // first element of the tile: even, the tile is contiguous
        const int ne = rows * ndim;
        __syncthreads();                   // the previous tile has been written out
        {   // (1) coalesced 16-byte streaming loads: ~1 instruction per element
            const double2 *s2 = reinterpret_cast<const double2 *>(in + e0);
            double2 *d2 = reinterpret_cast<double2 *>(tile);
            for (int i = tid; i < ne / 2; i += blockDim.x) d2[i] = __ldcs(s2 + i);
            if ((ne & 1) && tid == 0) tile[ne - 1] = __ldcs(in + e0 + ne - 1);
        }
        __syncthreads();
        if (tid < rows) {  // (2) one thread per row: order test (:113); permute the planets in place
            double *r = tile + (size_t)tid * ndim;
            double per[RVL_FIP_MAX_PLANETS];
            bool ordered = true;
            for (int j = 0; j < tab.K; ++j) {
                per[j] = r[tab.period_col[j]];
                if (j > 0 && !(per[j - 1] <= per[j])) ordered = false;  // NaN -> not ordered
            }
            if (!ordered) {
                int rank[RVL_FIP_MAX_PLANETS];  // position of planet p in np.argsort(periods): stable, NaN last
                for (int p = 0; p < tab.K; ++p) {
                    int k = 0;
                    for (int j = 0; j < tab.K; ++j)
                        k += (before(per[j], per[p]) || (j < p && !before(per[p], per[j]))) ? 1 : 0;
                    rank[p] = k;
                }
                for (int q = 0; q < tab.Q; ++q) {  // new[planet p][q] = old[planet rank(p)][q]  (:121-124)
                    double v[RVL_FIP_MAX_PLANETS];
                    for (int j = 0; j < tab.K; ++j) v[j] = r[tab.planet_col[j][q]];
                    for (int p = 0; p < tab.K; ++p) r[tab.planet_col[p][q]] = v[rank[p]];
                }
            }
        }
        __syncthreads();
        {   // (3) coalesced 16-byte streaming stores
            const double2 *s2 = reinterpret_cast<const double2 *>(tile);
            double2 *d2 = reinterpret_cast<double2 *>(out + e0);
            for (int i = tid; i < ne / 2; i += blockDim.x) __stcs(d2 + i, s2[i]);
            if ((ne & 1) && tid == 0) __stcs(out + e0 + ne - 1, tile[ne - 1]);
        }
    }
}

thread_local std::string g_order_error;
int ofail(int code, const std::string &msg)
{
    g_order_error = msg;
    return code;
}
#define OCU(call)                                                                               \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return ofail(RVL_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

}  // namespace

extern "C" {

const char *rvl_order_last_error(void) { return g_order_error.c_str(); }

int rvl_order_planets(int32_t device, const double *samples, int64_t n, int32_t ndim,
                      const int32_t *period_cols, const int32_t *planet_cols, int32_t K, int32_t Q,
                      double *out, double *kernel_ms)
{
    if (kernel_ms) *kernel_ms = 0.0;
    if (n < 0 || ndim < 1 || ndim > RVL_MAX_DIM || (n > 0 && (!samples || !out)))
        return ofail(RVL_EINVAL, "bad samples");
    if (K < 1 || K > RVL_FIP_MAX_PLANETS || Q < 1 || Q > RVL_ORDER_MAX_PARAMS || !period_cols ||
        !planet_cols)
        return ofail(RVL_EINVAL, "bad planet tables");
    OrderTab tab{};
    tab.K = K;
    tab.Q = Q;
    int8_t h_planet[RVL_MAX_DIM];
    for (int i = 0; i < RVL_MAX_DIM; ++i) h_planet[i] = -1;
    for (int p = 0; p < K; ++p) {
        if (period_cols[p] < 0 || period_cols[p] >= ndim) return ofail(RVL_EINVAL, "period column out of range");
        tab.period_col[p] = period_cols[p];
        for (int q = 0; q < Q; ++q) {
            const int c = planet_cols[p * Q + q];
            if (c < 0 || c >= ndim) return ofail(RVL_EINVAL, "planet column out of range");
            if (h_planet[c] >= 0) return ofail(RVL_EINVAL, "a column belongs to two planets");
            tab.planet_col[p][q] = c;
            h_planet[c] = (int8_t)p;
        }
    }
    if (n == 0) return RVL_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return ofail(RVL_ENODEV, "no CUDA device: evidence_b200 has no CPU fallback");
    }
    int prev = 0;
    cudaGetDevice(&prev);
    if (device < 0) device = prev;
    if (device >= ndev) return ofail(RVL_EINVAL, "device index out of range");
    OCU(cudaSetDevice(device));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};

    // device buffers and events are kept per device between calls (rvpost.h)
    std::lock_guard<std::mutex> lock(rvpost::g_mutex);
    struct { void *p; } d_in, d_out;
    const size_t nb = (size_t)n * ndim * sizeof(double);
    OCU(rvpost::get(device, 8, nb, &d_in.p));
    OCU(rvpost::get(device, 9, nb, &d_out.p));
    OCU(cudaMemcpyAsync(d_in.p, samples, nb, cudaMemcpyHostToDevice, 0));
    cudaEvent_t e0, e1;
    OCU(rvpost::events(device, &e0, &e1));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const long long n_tiles = (n + kTileRows - 1) / kTileRows;
    const size_t smem = (size_t)kTileRows * ndim * sizeof(double);  // <= 128 KB
    static uint64_t opted = 0;  // one bit per device: the attribute belongs to the device's context
    if (smem > 48 * 1024 && !(device < 64 && (opted >> device & 1))) {
        OCU(cudaFuncSetAttribute(order_planets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024 + 1024));
        if (device < 64) opted |= 1ull << device;
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 2048)));
    const unsigned grid = (unsigned)std::min<long long>(n_tiles, (long long)sms * per_sm);  // resident blocks
    OCU(cudaEventRecord(e0, 0));
    order_planets_kernel<<<grid, kOrderThreads, smem>>>((const double *)d_in.p, n, ndim, tab,
                                                       (double *)d_out.p);
    OCU(cudaEventRecord(e1, 0));
    OCU(cudaGetLastError());
    OCU(cudaMemcpy(out, d_out.p, nb, cudaMemcpyDeviceToHost));
    float ms = 0.f;
    OCU(cudaEventElapsedTime(&ms, e0, e1));
    if (kernel_ms) *kernel_ms = ms;
    return RVL_OK;
}

}  // extern "C"
