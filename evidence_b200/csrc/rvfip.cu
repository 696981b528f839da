// rvfip.cu — FIP-periodogram accumulation on the device (part of librvlnl.so).
//
// Reference path replaced (paths relative to the reference checkout):
//   evidence/fip_criterion.py:303-338   for every posterior sample of every k-planet run: the
//       mean motions 2pi/P of its planets (optionally with the 1-day / 30-day aliases), the
//       frequency-grid bins whose window [nu - w/2, nu + w/2] contains one of them
//       (two np.searchsorted calls), and  fapnu[run, bins] -= p(k|y) * weight  -- a Python loop
//       over samples with a fancy-index update whose duplicate bins subtract once.
//
// Device formulation.  One thread per sample: its <= 5 K mean motions -> [beg, end) bin ranges by
// binary search over the SAME nua / nub arrays the reference searches (so the bin edges are decided
// by the same comparisons), union of the ranges (insertion sort + merge in registers), then
// +w at beg / -w at end of every merged range into a difference array.  The difference array is
// 64-bit FIXED POINT (2^-56): integer atomics are exact and order-independent, so the result is
// bit-reproducible whatever the scheduling; the rounding of one term is <= 2^-57, i.e. ~1e-17
// per sample (sqrt(n) growth), below the reference's own accumulated rounding (~1e-16 per
// subtraction).  A single-block scan turns it into fapnu.  The work is L2-atomic bound; HBM sees
// 8 (k + 1) bytes per sample.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>

#include "../../include/rvlnl.h"
#include "rvpost.h"

namespace {

constexpr int kMaxRanges = 5 * RVL_FIP_MAX_PLANETS;
constexpr double kFix = 72057594037927936.0;  // 2^56

thread_local std::string g_fip_error;

// np.searchsorted(a, v, 'left')  = first i with a[i] >= v ; 'right' = first i with a[i] > v
__device__ __forceinline__ int lower_bound(const double *a, int n, double v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int upper_bound(const double *a, int n, double v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void fip_ranges_kernel(const double *nua, const double *nub, int nfreq,
                                  const double *periods, int k, const double *weights,
                                  long long n, double scale, int with_alias, double fmin, double fmax,
                                  double shift_day, double shift_month, unsigned long long *diff)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int beg[kMaxRanges], end[kMaxRanges];
    int m = 0;
    for (int p = 0; p < k; ++p) {
        const double f0 = __ddiv_rn(6.283185307179586, periods[i * k + p]);  // 2*np.pi/x (:319)
        const int nv = with_alias ? 5 : 1;
        for (int v = 0; v < nv; ++v) {
            double f = f0;
            if (v == 1) f = fabs(__dadd_rn(f0, shift_day));    // :322
            if (v == 2) f = fabs(__dsub_rn(f0, shift_day));    // :323
            if (v == 3) f = fabs(__dadd_rn(f0, shift_month));  // :324
            if (v == 4) f = fabs(__dsub_rn(f0, shift_month));  // :325
            if (with_alias && !(f <= fmax && f >= fmin)) continue;  // :330-331 (NaN drops out too)
            const int b = upper_bound(nub, nfreq, f);  // np.searchsorted(nub, f, 'right') (:333)
            const int e = lower_bound(nua, nfreq, f);  // np.searchsorted(nua, f, 'left')  (:334)
            if (e <= b) continue;                       // range(b, e) is empty
            // insertion sort by beg
            int j = m++;
            while (j > 0 && beg[j - 1] > b) {
                beg[j] = beg[j - 1];
                end[j] = end[j - 1];
                --j;
            }
            beg[j] = b;
            end[j] = e;
        }
    }
    if (m == 0) return;
    // fixed-point weight; the scale carries p(k|y) / sum(weights)  (:313, :338)
    const long long w = llrint(__dmul_rn(__dmul_rn(weights[i], scale), kFix));
    if (w == 0) return;
    // union of the sorted ranges: every bin of the sample is updated once (fancy-index semantics)
    int cb = beg[0], ce = end[0];
    for (int j = 1; j <= m; ++j) {
        if (j < m && beg[j] <= ce) {
            ce = max(ce, end[j]);
            continue;
        }
        atomicAdd(diff + cb, (unsigned long long)w);
        atomicAdd(diff + ce, (unsigned long long)(-w));
        if (j < m) {
            cb = beg[j];
            ce = end[j];
        }
    }
}

// fapnu[j] -= 2^-56 * cumsum(diff)[j]; one block, each thread owns a contiguous run of bins
__global__ void __launch_bounds__(1024) fip_scan_kernel(const unsigned long long *diff, int nfreq,
                                                        double *fapnu)
{
    __shared__ long long part[1024];
    const int T = blockDim.x, t = threadIdx.x;
    const int per = (nfreq + T - 1) / T;
    const int lo = min(nfreq, t * per), hi = min(nfreq, lo + per);
    long long s = 0;
    for (int j = lo; j < hi; ++j) s += (long long)diff[j];
    part[t] = s;
    __syncthreads();
    // exclusive prefix over the T partial sums (T <= 1024: a simple Hillis-Steele scan)
    for (int off = 1; off < T; off <<= 1) {
        const long long v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long acc = part[t] - s;
    for (int j = lo; j < hi; ++j) {
        acc += (long long)diff[j];
        fapnu[j] = __dsub_rn(fapnu[j], __dmul_rn((double)acc, 1.0 / kFix));
    }
}

int fip_fail(int code, const std::string &msg)
{
    g_fip_error = msg;
    return code;
}
#define FCU(call)                                                                               \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fip_fail(RVL_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
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

const char *rvl_fip_last_error(void) { return g_fip_error.c_str(); }

void rvl_post_release(void) { rvpost::release_all(); }

int rvl_fip_accumulate(int32_t device, const double *nua, const double *nub, int32_t nfreq,
                       const double *periods, int32_t k, const double *weights, int64_t n,
                       double pk, int32_t with_alias, double pmin, double pmax, double *fapnu,
                       double *kernel_ms)
{
    if (!nua || !nub || !fapnu || nfreq < 1) return fip_fail(RVL_EINVAL, "bad grid");
    if (k < 1 || k > RVL_FIP_MAX_PLANETS) return fip_fail(RVL_EINVAL, "k out of range");
    if (n < 0 || (n > 0 && (!periods || !weights))) return fip_fail(RVL_EINVAL, "bad samples");
    if (kernel_ms) *kernel_ms = 0.0;
    if (n == 0) return RVL_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fip_fail(RVL_ENODEV, "no CUDA device: evidence_b200 has no CPU fallback");
    }
    int prev = 0;
    cudaGetDevice(&prev);
    if (device < 0) device = prev;
    if (device >= ndev) return fip_fail(RVL_EINVAL, "device index out of range");
    FCU(cudaSetDevice(device));
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};

    // sum of the weights on the host, in numpy's order of summation (pairwise) is not needed to
    // the bit: the normalised weight enters one rounding further down; plain long-double sum
    long double wsum = 0.0L;
    for (int64_t i = 0; i < n; ++i) wsum += weights[i];
    if (!(wsum > 0.0L)) return fip_fail(RVL_EINVAL, "weights do not sum to a positive number");
    const double scale = pk / (double)wsum;

    // device buffers and events are kept per device between calls (rvpost.h)
    std::lock_guard<std::mutex> lock(rvpost::g_mutex);
    struct { void *p; } d_nua, d_nub, d_per, d_w, d_diff, d_fap;
    const size_t gb = (size_t)nfreq * sizeof(double);
    FCU(rvpost::get(device, 0, gb, &d_nua.p));
    FCU(rvpost::get(device, 1, gb, &d_nub.p));
    FCU(rvpost::get(device, 2, gb, &d_fap.p));
    FCU(rvpost::get(device, 3, ((size_t)nfreq + 1) * sizeof(unsigned long long), &d_diff.p));
    FCU(rvpost::get(device, 4, (size_t)n * k * sizeof(double), &d_per.p));
    FCU(rvpost::get(device, 5, (size_t)n * sizeof(double), &d_w.p));
    FCU(cudaMemcpyAsync(d_nua.p, nua, gb, cudaMemcpyHostToDevice, 0));
    FCU(cudaMemcpyAsync(d_nub.p, nub, gb, cudaMemcpyHostToDevice, 0));
    FCU(cudaMemcpyAsync(d_fap.p, fapnu, gb, cudaMemcpyHostToDevice, 0));
    FCU(cudaMemcpyAsync(d_per.p, periods, (size_t)n * k * sizeof(double), cudaMemcpyHostToDevice, 0));
    FCU(cudaMemcpyAsync(d_w.p, weights, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, 0));
    FCU(cudaMemsetAsync(d_diff.p, 0, ((size_t)nfreq + 1) * sizeof(unsigned long long), 0));

    cudaEvent_t e0, e1;
    FCU(rvpost::events(device, &e0, &e1));
    const double two_pi = 6.283185307179586;
    const int tb = 256;
    FCU(cudaEventRecord(e0, 0));
    fip_ranges_kernel<<<(unsigned)((n + tb - 1) / tb), tb>>>(
        (const double *)d_nua.p, (const double *)d_nub.p, nfreq, (const double *)d_per.p, k,
        (const double *)d_w.p, n, scale, with_alias, two_pi / pmax, two_pi / pmin,
        two_pi / 0.99727, two_pi / 30.0, (unsigned long long *)d_diff.p);
    fip_scan_kernel<<<1, 1024>>>((const unsigned long long *)d_diff.p, nfreq, (double *)d_fap.p);
    FCU(cudaEventRecord(e1, 0));
    FCU(cudaGetLastError());
    FCU(cudaMemcpy(fapnu, d_fap.p, gb, cudaMemcpyDeviceToHost));
    float ms = 0.f;
    FCU(cudaEventElapsedTime(&ms, e0, e1));
    if (kernel_ms) *kernel_ms = ms;
    return RVL_OK;
}

}  // extern "C"
