// rvorder.cu — posterior planet-ordering on the device (part of librvlnl.so).
//
// Reference path replaced: evidence/post_processing.py:104-128 -- a pandas `iterrows` loop that,
// for every posterior sample whose planet periods are not non-decreasing, rebuilds an index list
// from np.argsort(periods) and gathers the row through it.  The gather uses the RANK of a column's
// own planet as the source slot (new[i] = old[planets[rank(p_i)][q_i]]), which is the inverse of
// the sorting permutation; reproduced as is (identical results on identical inputs).
//
// One warp per row: the row is read ONCE, coalesced, into the warp's slice of shared memory; the K
// periods, the "already ordered" test (:113) and the ranks come from there (broadcast reads), and
// the row is written back coalesced through the gather.  Pure byte movement, 16 B of HBM traffic
// per element and nothing else: the bound is HBM bandwidth.
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

constexpr int kRowsPerBlock = 8;  // warps per block

__global__ void __launch_bounds__(kRowsPerBlock * 32)
order_planets_kernel(const double *in, long long n, int ndim, const OrderTab tab, double *out)
{
    __shared__ double srow[kRowsPerBlock][RVL_MAX_DIM];
    __shared__ int8_t s_planet[RVL_MAX_DIM], s_pos[RVL_MAX_DIM];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // column -> (planet, position within the planet) from the table (K * Q entries)
    for (int c = threadIdx.x; c < ndim; c += blockDim.x) s_planet[c] = -1;
    __syncthreads();
    for (int i = threadIdx.x; i < tab.K * tab.Q; i += blockDim.x) {
        const int p = i / tab.Q, q = i - p * tab.Q;
        s_planet[tab.planet_col[p][q]] = (int8_t)p;
        s_pos[tab.planet_col[p][q]] = (int8_t)q;
    }
    __syncthreads();
    double *r = srow[wib];
    const long long stride = (long long)gridDim.x * kRowsPerBlock;
    for (long long row = (long long)blockIdx.x * kRowsPerBlock + wib; row < n; row += stride) {
        const double *src = in + row * ndim;
        for (int c = lane; c < ndim; c += 32) r[c] = __ldcs(src + c);  // streamed: read once
        __syncwarp();
        double per[RVL_FIP_MAX_PLANETS];
        bool ordered = true;
        for (int j = 0; j < tab.K; ++j) {
            per[j] = r[tab.period_col[j]];
            if (j > 0 && !(per[j - 1] <= per[j])) ordered = false;  // :113 (NaN -> not ordered)
        }
        double *dst = out + row * ndim;
        for (int c = lane; c < ndim; c += 32) {
            int from = c;
            const int p = s_planet[c];
            if (!ordered && p >= 0) {
                int rank = 0;  // position of planet p in np.argsort(periods): stable, NaN last
                for (int j = 0; j < tab.K; ++j)
                    rank += (before(per[j], per[p]) || (j < p && !before(per[p], per[j]))) ? 1 : 0;
                from = tab.planet_col[rank][s_pos[c]];  // :121-124
            }
            __stcs(dst + c, r[from]);
        }
        __syncwarp();
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
struct Buf {
    void *p = nullptr;
    ~Buf() { if (p) cudaFree(p); }
};
struct Ev {  // released on every return path, like Buf
    cudaEvent_t e = nullptr;
    ~Ev() { if (e) cudaEventDestroy(e); }
};

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
    int8_t h_planet[RVL_MAX_DIM], h_pos[RVL_MAX_DIM];
    for (int i = 0; i < RVL_MAX_DIM; ++i) { h_planet[i] = -1; h_pos[i] = 0; }
    for (int p = 0; p < K; ++p) {
        if (period_cols[p] < 0 || period_cols[p] >= ndim) return ofail(RVL_EINVAL, "period column out of range");
        tab.period_col[p] = period_cols[p];
        for (int q = 0; q < Q; ++q) {
            const int c = planet_cols[p * Q + q];
            if (c < 0 || c >= ndim) return ofail(RVL_EINVAL, "planet column out of range");
            if (h_planet[c] >= 0) return ofail(RVL_EINVAL, "a column belongs to two planets");
            tab.planet_col[p][q] = c;
            h_planet[c] = (int8_t)p;
            h_pos[c] = (int8_t)q;
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
    const long long blocks_needed = (n + kRowsPerBlock - 1) / kRowsPerBlock;
    const unsigned grid = (unsigned)std::min<long long>(blocks_needed, (long long)sms * 8);  // resident blocks
    OCU(cudaEventRecord(e0, 0));
    order_planets_kernel<<<grid, kRowsPerBlock * 32>>>((const double *)d_in.p, n, ndim, tab,
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
