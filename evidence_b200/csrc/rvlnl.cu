// rvlnl.cu — librvlnl.so: B200 (sm_100a) batched Keplerian RV log-likelihood + prior transform
// behind the C-ABI of include/rvlnl.h.  No CPU fallback, no dispatch: sm_100a only.
//
// Reference path replaced (paths relative to the reference checkout):
//   evidence/rvmodel/__init__.py:157-219  RVModel.log_likelihood      -> rv_lnl_kernel
//   evidence/rvmodel/__init__.py:343-463  kep_rv / modelk             -> point_setup + solve_planet
//   evidence/rvmodel/trueanomaly.c:8-41   trueanomaly()               -> solve_planet / trueanomaly_kernel
//   evidence/rvmodel/__init__.py:222-273  drift                       -> epoch term of rv_lnl_kernel
//   evidence/rvmodel/__init__.py:59-80    BaseModel.logL              -> epoch term + slice reduce
//   evidence/ultranest/__init__.py:125-137 prior(hypercube)           -> prior_transform_kernel / point_prepare_kernel
//
// Layout.  Epoch data lives in HBM as columns [ncol][Npad] of doubles (t, vrad, svrad^2, then
// (t-tref)/365.25 when the model has a drift, then the linear-parameter columns) followed by
// Npad instrument ids (uint8).  The epoch axis is cut into S slices of whole 32-epoch chunks;
// block b serves slice b % S and brings that slice of every column into shared memory ONCE with
// 1-D TMA bulk copies (cp.async.bulk + mbarrier), then stays resident: its warps pull parameter
// vectors from a per-slice work counter.  warp <-> one parameter vector, lanes <-> 32 epochs:
// all lanes of a warp share the eccentricity, so Newton iteration counts are nearly uniform and
// the loop exit / sin-cos path choice come from one warp redux per trip.  chi^2 and log-det sums are
// reduced with warp shuffles; with S > 1 a one-warp-per-point prepare pass computes the per-point
// constants once (optionally fused with the prior transform), every slice writes its partial sums
// and a small second kernel adds them in slice order (deterministic).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/rvlnl.h"
#include "rvl_math.h"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kPlanetStride = 8;  // doubles per planet in the per-warp constant block
constexpr int kModelBytes = (int)((sizeof(rvl_model_desc) + 127) / 128 * 128);
// per-planet constants: 0 nmot, 1 M0, 2 ec, 3 A, 4 Bs, 5 Ce, 6 epoch

struct KArgs {
    const rvl_model_desc *model;  // device copy
    const double *cols;           // [ncol][Npad]
    const uint8_t *inst;          // [Npad]
    const double *theta;          // [B][ndim]
    double *lnl;                  // [B]        (written directly when S == 1)
    double *partial;              // [B][S][2]  (S > 1)
    int *flags;                   // [B] 1 = invalid Keplerian (S > 1, or written by the prepare pass)
    const double *consts;         // [B][wstride] per-point constants from point_prepare_kernel, or NULL
    unsigned long long *counters; // 0 newton iters, 1 cap hits, 2 invalid points
    unsigned int *work;           // [S] dynamic work counters
    long long B;
    double cte;  // -0.5 N ln(2 pi)
    int N, Npad, ncol;
    int S, cps;  // slices, chunks per slice
    int wstride; // doubles per warp in the constant block
};

// ---- shared-memory / TMA helpers (sm_90+ PTX) --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D TMA: global -> shared, completion counted in bytes on the mbarrier.  bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ double par_of(const rvl_param &p, const double *row)
{
    return p.slot >= 0 ? row[p.slot] : p.value;
}

// shared-memory loads through a 32-bit shared-window address (one live register per base
// pointer; stops the compiler from re-deriving generic addresses inside the hot loops)
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return (int)v;
}

// huge / non-finite arguments: CUDA libdevice (Payne-Hanek), out of line and returned by value
__device__ __noinline__ double2 sincos_slow(double x)
{
    double s, c;
    sincos(x, &s, &c);
    return make_double2(s, c);
}
// high word of |x| as an integer: orders like |x| itself (positive doubles sort as integers)
__device__ __forceinline__ int abs_hi(double x) { return __double2hiint(x) & 0x7fffffff; }
constexpr int kHiTrigMax = 0x40F86A00;  // 1e5 = 0x40F86A0000000000: |x| < 1e5  <=>  abs_hi < this
constexpr int kHiTiny = 0x3F500000;     // abs_hi(d) <  this  <=>  |d| <  2^-10
constexpr int kHiSmall = 0x3FA00000;    // abs_hi(d) <  this  <=>  |d| <  2^-5
#ifndef RVL_MEDIUM
#define RVL_MEDIUM 1
#endif
constexpr int kHiMedium = 0x3FE80000;   // abs_hi(d) <  this  <=>  |d| <  0.75

template <int U>
__device__ __forceinline__ bool any_big(const double (&E)[U])
{
    int emax = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) emax = max(emax, abs_hi(E[u]));
    return __any_sync(kFull, !(emax < kHiTrigMax));
}

// ---- U epochs per lane x one planet: Kepler solves + RV terms -------------------------------
// VARIANT 0: optimised (reciprocal-multiply Newton step, warp-uniform small-step sin/cos
//            advance).  VARIANT 1: conservative (IEEE division, full sin/cos every step) — kept
//            as the in-product cross-check of the optimisations, selectable with
//            rvl_set_option("variant", 1).
// U independent solves per lane (instruction-level parallelism; control flow, votes and constant
// loads are shared by the U solves).
// Lanes freeze individually (trueanomaly.c:21: per-element stop): a frozen lane's step is forced
// to 0, so its E never moves again and `|d| > tol` keeps it inactive without a separate flag.
template <int VARIANT, int U>
__device__ __forceinline__ void solve_planet(const double (&t)[U], uint32_t pc, double tol,
                                             int itmax, double (&rv)[U], int (&iters)[U],
                                             int &caps)
{
    const double nmot = lds_f64(pc), M0 = lds_f64(pc + 8), ec = lds_f64(pc + 16),
                 epoch = lds_f64(pc + 48);
    double M[U], E[U], s[U], c[U], d[U];
    int last[U];
    bool big = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        M[u] = rvl::mean_anomaly(nmot, t[u], epoch, M0);
        E[u] = M[u];
        big = big || !(abs_hi(M[u]) < kHiTrigMax);
        d[u] = 1e300;  // "no step taken yet": forces the full sin/cos on the first pass
        s[u] = 0.0;
        c[u] = 1.0;
        last[u] = 0;
    }
    // slow: some |M| >= 1e5 (or non-finite), or an eccentricity outside [-0.99, 0.99] (only
    // reachable with a nonsensical direct `ecc`): every sin/cos of this solve goes through
    // libdevice, which is valid for any argument
    const bool slow = __any_sync(kFull, big || !(ec >= -0.99));
    int trip = 0;  // warp-uniform number of Newton steps taken so far
    const int tol_hi = __double2hiint(tol);
#pragma unroll 2
    for (;;) {
        // warp-wide maximum of the high words of |d|: ONE redux decides the sin/cos path and
        // (except in a 1e-6-wide band around tol) the loop exit, all on uniform values
        int hmax = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) hmax = max(hmax, abs_hi(d[u]));
        const int wmax = (int)__reduce_max_sync(kFull, (unsigned)hmax);
        // (1) bring (sin E, cos E) up to date with the step d just taken
        if (VARIANT == 0 && wmax < kHiTiny) {
#pragma unroll
            for (int u = 0; u < U; ++u) rvl::advance_tiny(d[u], s[u], c[u]);
        } else if (VARIANT == 0 && wmax < kHiSmall) {
#pragma unroll
            for (int u = 0; u < U; ++u) rvl::advance_small(d[u], s[u], c[u]);
        } else if (VARIANT == 0 && RVL_MEDIUM && !slow && wmax < kHiMedium) {
#pragma unroll
            for (int u = 0; u < U; ++u) rvl::advance_medium(d[u], s[u], c[u]);
        } else if (slow || (trip > 2 && any_big<U>(E))) {
            // |E| can only leave the fast range after >= 3 Newton steps (|step| <= 100 |f|)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double2 r = sincos_slow(E[u]);
                s[u] = r.x;
                c[u] = r.y;
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) rvl::sincos_fast(E[u], s[u], c[u]);
        }
        // (2) which lanes still iterate (trueanomaly.c:21); the cap (:32-33) bounds the loop
        bool pa[U];
#pragma unroll
        for (int u = 0; u < U; ++u) pa[u] = fabs(d[u]) > tol;
        if (trip >= itmax || wmax < tol_hi) break;  // high word below tol's: every |d| < tol
        if (wmax == tol_hi) {                       // rare tie on the high word: exact vote
            bool any_left = false;
#pragma unroll
            for (int u = 0; u < U; ++u) any_left = any_left || pa[u];
            if (!__any_sync(kFull, any_left)) break;
        }
        ++trip;
        // (3) one Newton step for the active lanes; a frozen lane's step is exactly 0
#pragma unroll
        for (int u = 0; u < U; ++u) {
            double En;
            if (VARIANT == 0) {
                rvl::newton_step(E[u], s[u], c[u], M[u], ec, En);
            } else {
                const double f = rvl::sub(rvl::sub(E[u], rvl::mul(ec, s[u])), M[u]);
                const double fp = rvl::sub(1.0, rvl::mul(ec, c[u]));
                En = rvl::sub(E[u], __ddiv_rn(f, fp));
            }
            En = pa[u] ? En : E[u];
            d[u] = rvl::sub(En, E[u]);  // exact
            E[u] = En;
            last[u] = pa[u] ? trip : last[u];
        }
    }
    const double A = lds_f64(pc + 24), Bs = lds_f64(pc + 32), Ce = lds_f64(pc + 40);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        iters[u] += last[u];
        caps += (fabs(d[u]) > tol) ? 1 : 0;
        rv[u] = rvl::kepler_rv(s[u], c[u], ec, A, Bs, Ce);
    }
}

// ---- per-point setup: theta row -> per-warp constants (modelk :411-457, :181-192) ---------
// lane p < K handles planet p; lane i < n_inst handles instrument i; lane 0 the drift/linpar
// coefficients.  Returns (warp-uniform) whether the point is a valid Keplerian.
__device__ __forceinline__ bool point_setup(const rvl_model_desc &m, const double *row,
                                            double *wc, int lane)
{
    const int K = m.n_planets;
    bool bad = false;
    if (lane < K) {
        const rvl_planet_desc &pl = m.planet[lane];
        double amp = par_of(pl.amp, row);
        if (pl.amp_is_log) amp = exp(amp);
        double per = par_of(pl.period, row);
        if (pl.period_is_log) per = exp(per);
        const double a = par_of(pl.e1, row), b = par_of(pl.e2, row);
        double ecc, omega;
        if (pl.ecc_mode == RVL_ECC_SECOS_SESIN) {
            ecc = rvl::add(rvl::mul(a, a), rvl::mul(b, b));
            omega = atan2(b, a);
            bad = ecc > 1.0;
        } else if (pl.ecc_mode == RVL_ECC_ECOS_ESIN) {
            ecc = sqrt(rvl::add(rvl::mul(a, a), rvl::mul(b, b)));
            omega = atan2(b, a);
            bad = ecc > 1.0;
        } else {
            ecc = a;
            omega = b;
        }
        double M0 = par_of(pl.phase, row);
        if (pl.phase_mode == RVL_PHASE_ML0) M0 = rvl::sub(M0, omega);
        const double ec = ecc > 0.99 ? 0.99 : ecc;  // trueanomaly.c:11-12
        double sw, cw;
        if (abs_hi(omega) < kHiTrigMax) {
            rvl::sincos_fast(omega, sw, cw);
        } else {
            const double2 r = sincos_slow(omega);
            sw = r.x;
            cw = r.y;
        }
        const double root = sqrt(rvl::mul(rvl::sub(1.0, ec), rvl::add(1.0, ec)));
        double *pc = wc + lane * kPlanetStride;
        pc[0] = __ddiv_rn(6.283185307179586, per);  // 2*np.pi/P_day (:459)
        pc[1] = M0;
        pc[2] = ec;
        pc[3] = rvl::mul(amp, cw);
        pc[4] = -rvl::mul(rvl::mul(amp, sw), root);
        pc[5] = rvl::mul(amp, rvl::mul(ecc, cw));
        pc[6] = par_of(pl.epoch, row);
    }
    double *ic = wc + K * kPlanetStride;
    if (lane < m.n_inst) {
        ic[2 * lane] = par_of(m.offset[lane], row);
        double j2 = 0.0;
        if (m.jitter_in_model) {
            const double j = par_of(m.jitter[lane], row);
            j2 = rvl::mul(j, j);
        }
        ic[2 * lane + 1] = j2;
    }
    double *dc = ic + 2 * m.n_inst;
    if (lane < 4) dc[lane] = m.drift_in_model ? par_of(m.drift[lane], row) : 0.0;
    if (lane < m.n_linpar) dc[4 + lane] = par_of(m.linpar[lane], row);
    __syncwarp();
    return !__any_sync(kFull, bad);
}

// ---- the likelihood kernel ------------------------------------------------------------------
// U = epochs per lane in flight (1: 1024 threads/SM at <=64 registers; 2: 512 threads/SM at
// <=128 registers, two chunks of 32 epochs per warp trip).
template <int VARIANT, int U, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) rv_lnl_kernel(const KArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // [0,8) mbarrier | [128, 128+sizeof(model)) model | epoch columns | inst ids | warp consts
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    rvl_model_desc *sm = reinterpret_cast<rvl_model_desc *>(smem_raw + 128);
    const int sl = blockIdx.x % a.S;
    const int Ctot = a.Npad / 32;
    const int c0 = sl * a.cps;
    const int nch = min(a.cps, Ctot - c0);  // chunks in this slice (>= 1 by construction)
    const int ne = nch * 32;
    double *scol = reinterpret_cast<double *>(smem_raw + 128 + kModelBytes);
    uint8_t *sinst = reinterpret_cast<uint8_t *>(scol + (size_t)a.ncol * ne);
    double *wconst = reinterpret_cast<double *>(
        smem_raw + 128 + kModelBytes + (((size_t)a.ncol * ne * 8 + ne + 127) / 128) * 128);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned col_bytes = (unsigned)ne * 8u;
        mbar_expect_tx(bar, col_bytes * (unsigned)a.ncol + (unsigned)ne);
        for (int cidx = 0; cidx < a.ncol; ++cidx)
            tma_load_1d(scol + (size_t)cidx * ne, a.cols + (size_t)cidx * a.Npad + (size_t)c0 * 32,
                        col_bytes, bar);
        tma_load_1d(sinst, a.inst + (size_t)c0 * 32, (unsigned)ne, bar);
    }
    // model description: plain cooperative copy (1.6 KB) while the TMA is in flight
    {
        const int nw = (int)(sizeof(rvl_model_desc) / 4);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.model);
        uint32_t *dst = reinterpret_cast<uint32_t *>(sm);
        for (int i = tid; i < nw; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    mbar_wait(bar, 0);

    const rvl_model_desc &m = *sm;
    const int K = m.n_planets;
    const double tol = m.tol;
    const int itmax = m.itmax;
    const bool has_drift = m.drift_in_model != 0;
    int drift_hi = 0;  // highest drift coefficient that is a parameter or a non-zero constant
    for (int i = 1; i < 4; ++i)
        if (m.drift[i].slot >= 0 || m.drift[i].value != 0.0) drift_hi = i;
    const int nlin = m.n_linpar;
    double *wc = wconst + (size_t)warp * a.wstride;
    // 32-bit shared-window addresses of everything the hot loop reads
    const uint32_t a_wc = smem_u32(wc);
    const uint32_t a_ic = a_wc + (uint32_t)(K * kPlanetStride) * 8u;
    const uint32_t a_dc = a_ic + (uint32_t)(2 * m.n_inst) * 8u;
    const uint32_t a_t = smem_u32(scol) + (uint32_t)lane * 8u;
    const uint32_t colb = (uint32_t)ne * 8u;  // bytes per column
    const uint32_t a_inst = smem_u32(sinst) + (uint32_t)lane;
    const uint32_t lin0 = (uint32_t)(3 + (has_drift ? 1 : 0)) * colb;
    const int e_base = c0 * 32 + lane;  // global epoch index of this lane in chunk 0

    unsigned long long tot_iters = 0, tot_caps = 0, tot_invalid = 0;

    // work queue of this slice: the index of the NEXT item is requested while the current one is
    // computed, so the atomic's round trip (~1 us) is off the critical path
    unsigned idx = 0;
    if (lane == 0) idx = atomicAdd(&a.work[sl], 1u);
    idx = __shfl_sync(kFull, idx, 0);
    while ((long long)idx < a.B) {
        const long long pt = idx;
        const double *row = a.theta + pt * m.ndim;
        unsigned next = 0;
        if (lane == 0) next = atomicAdd(&a.work[sl], 1u);

        __syncwarp();
        bool valid;
        if (a.consts) {  // constants were prepared once per point (S > 1 / fused transform)
            const double *src = a.consts + (size_t)pt * a.wstride;
            for (int i = lane; i < a.wstride; i += 32) wc[i] = __ldg(src + i);
            valid = __ldg(a.flags + pt) == 0;
            __syncwarp();
        } else {
            valid = point_setup(m, row, wc, lane);
        }

        double chi = 0.0, prod = 1.0;
        int esum = 0, iters = 0, caps = 0;
        bool ok = true;
        if (valid) {
            for (int ch = 0; ch < nch; ch += U) {
                // U chunks of 32 epochs; a missing last chunk repeats the previous one, masked
                uint32_t off[U];
                bool live[U];
                double t[U], rvsum[U];
                int it_l[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool have = (ch + u) < nch;
                    const int cu = have ? ch + u : ch;
                    off[u] = (uint32_t)cu * 256u;
                    live[u] = have && (e_base + cu * 32) < a.N;
                    t[u] = lds_f64(a_t + off[u]);
                    rvsum[u] = 0.0;
                    it_l[u] = 0;
                }
                int cap_l = 0;
                for (int p = 0; p < K; ++p) {
                    double v[U];
                    solve_planet<VARIANT, U>(t, a_wc + (uint32_t)(p * kPlanetStride) * 8u, tol,
                                             itmax, v, it_l, cap_l);
#pragma unroll
                    for (int u = 0; u < U; ++u) rvsum[u] = (p == 0) ? v[u] : rvl::add(rvsum[u], v[u]);
                }
                caps += cap_l;  // (padded / repeated lanes included: a cap hit is a cap hit)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t ae = a_t + off[u];
                    const int ii = lds_u8(a_inst + off[u] / 8u);
                    const uint32_t ai = a_ic + (uint32_t)ii * 16u;
                    double rvm = lds_f64(ai);
                    if (K > 0) rvm = rvl::add(rvm, rvsum[u]);
                    if (has_drift) {
                        // lin*tt + quad*tt^2 + cub*tt^3 + quar*tt^4, left to right (:271); a
                        // coefficient that is absent from the model is an exact +0 term: skipped
                        const double tt = lds_f64(ae + 3u * colb);
                        double dr = rvl::mul(lds_f64(a_dc), tt);
                        if (drift_hi > 0) {
                            const double t2 = rvl::mul(tt, tt);
                            dr = rvl::add(dr, rvl::mul(lds_f64(a_dc + 8), t2));
                            if (drift_hi > 1) {
                                dr = rvl::add(dr, rvl::mul(lds_f64(a_dc + 16), rvl::mul(t2, tt)));
                                if (drift_hi > 2)
                                    dr = rvl::add(dr, rvl::mul(lds_f64(a_dc + 24), rvl::mul(t2, t2)));
                            }
                        }
                        rvm = rvl::add(rvm, dr);
                    }
                    for (int l = 0; l < nlin; ++l)
                        rvm = rvl::add(rvm, rvl::mul(lds_f64(a_dc + 32u + (uint32_t)l * 8u),
                                                     lds_f64(ae + lin0 + (uint32_t)l * colb)));
                    const double res = rvl::sub(lds_f64(ae + colb), rvm);
                    const double var = rvl::add(lds_f64(ae + 2u * colb), lds_f64(ai + 8));
                    const double term = rvl::mul(rvl::mul(res, res), rvl::rcp(rvl::add(var, var)));
                    double mant;
                    int ex;
                    const bool okv = rvl::split_pos(var, mant, ex);
                    if (live[u]) {
                        chi = rvl::add(chi, term);
                        prod = rvl::mul(prod, mant);
                        esum += ex;
                        ok = ok && okv;
                        iters += it_l[u];
                    }
                }
                if ((ch & 255) == 254 || (ch & 255) == 255) {  // keep the mantissa product in range
                    double mm;
                    int ee;
                    rvl::split_pos(prod, mm, ee);
                    prod = mm;
                    esum += ee;
                }
            }
        }

        // ---- slice reduction: chi^2 sum, mantissa product, exponent sum ----
        double S1, S2;
        if (valid) {
            const bool all_ok = __all_sync(kFull, ok);
            if (all_ok) {
                {  // every lane's mantissa product back to [1,2): 32 of them multiply to < 2^32
                    double mm;
                    int ee;
                    rvl::split_pos(prod, mm, ee);
                    prod = mm;
                    esum += ee;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    chi = rvl::add(chi, __shfl_xor_sync(kFull, chi, o));
                    prod = rvl::mul(prod, __shfl_xor_sync(kFull, prod, o));
                }
                esum = __reduce_add_sync(kFull, esum);
                // sum ln sqrt(var) = 0.5 (ln prod + esum ln 2)
                const double ld = rvl::fma_((double)esum, rvl::kLn2Hi,
                                           rvl::fma_((double)esum, rvl::kLn2Lo, log(prod)));
                S1 = rvl::mul(0.5, ld);
            } else {
                // a variance that is zero / subnormal / negative / non-finite: plain logs
                double acc = 0.0;
                for (int ch = 0; ch < nch; ++ch) {
                    const uint32_t o8 = (uint32_t)ch * 256u;
                    if ((e_base + ch * 32) < a.N) {
                        const int ii = lds_u8(a_inst + o8 / 8u);
                        const double var = rvl::add(lds_f64(a_t + o8 + 2u * colb),
                                                    lds_f64(a_ic + (uint32_t)ii * 16u + 8u));
                        acc = rvl::add(acc, log(sqrt(var)));
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    acc = rvl::add(acc, __shfl_xor_sync(kFull, acc, o));
                    chi = rvl::add(chi, __shfl_xor_sync(kFull, chi, o));
                }
                S1 = acc;
            }
            S2 = chi;
            tot_iters += (unsigned long long)__reduce_add_sync(kFull, iters);
            tot_caps += (unsigned long long)__reduce_add_sync(kFull, caps);
        } else {
            S1 = 0.0;
            S2 = 0.0;
            if (sl == 0) ++tot_invalid;
        }
        if (lane == 0) {
            if (a.S == 1) {
                // (cte - sum ln sqrt var) - sum r^2/(2 var)   (:80); invalid -> -1e30 (:203)
                a.lnl[pt] = valid ? rvl::sub(rvl::sub(a.cte, S1), S2) : -1e30;
            } else {
                // this slice's partial sums; combine_slices_kernel adds them in slice order
                double *o = a.partial + ((size_t)pt * a.S + sl) * 2;
                o[0] = S1;
                o[1] = S2;
            }
        }
        idx = __shfl_sync(kFull, next, 0);
    }
    if (lane == 0) {
        if (tot_iters) atomicAdd(&a.counters[0], tot_iters);
        if (tot_caps) atomicAdd(&a.counters[1], tot_caps);
        if (tot_invalid) atomicAdd(&a.counters[2], tot_invalid);
    }
}

// ---- once-per-point pass: (optional) unit cube -> theta, then the per-point constants ----------
// One warp per point.  With U != NULL this is the fused prior transform: lane i < ndim evaluates
// ppf_i(u_i), theta is written out (the sampler stores it) and the constants are derived from
// exactly those values.  Used whenever the epoch axis is cut into S > 1 slices (the constants are
// then computed once per point instead of once per slice) and for rvl_transform_loglike.
__device__ double ppf_eval(const rvl_prior_desc &pr, const double *tables, double q);

__global__ void __launch_bounds__(256) point_prepare_kernel(const rvl_model_desc *model,
                                                            const rvl_prior_desc *priors,
                                                            const double *tables, const double *U,
                                                            double *theta, double *consts,
                                                            int *flags,
                                                            unsigned int *work, int n_work,
                                                            long long B, int wstride)
{
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x < n_work) work[threadIdx.x] = 0u;  // per-slice work queues
    // the model description is read many times per point: one cooperative copy to shared memory
    __shared__ __align__(16) unsigned char s_model[sizeof(rvl_model_desc)];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(model);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_model);
        for (int i = threadIdx.x; i < (int)(sizeof(rvl_model_desc) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= B) return;  // whole warps
    const rvl_model_desc &m = *reinterpret_cast<const rvl_model_desc *>(s_model);
    // the row is staged in shared memory with ONE coalesced read per warp: theta / U may live in
    // pinned host memory (zero-copy host-buffer calls), where scattered 8-byte reads are costly
    __shared__ double srow_all[8][RVL_MAX_DIM];
    double *srow = srow_all[threadIdx.x >> 5];
    double *row = theta + pt * m.ndim;
    if (U) {
        for (int i = lane; i < m.ndim; i += 32) {
            const double v = ppf_eval(priors[i], tables, U[pt * m.ndim + i]);
            srow[i] = v;
            row[i] = v;  // theta is an output of the fused call
        }
    } else {
        for (int i = lane; i < m.ndim; i += 32) srow[i] = row[i];
    }
    __syncwarp();
    const bool valid = point_setup(m, srow, consts + (size_t)pt * wstride, lane);
    if (lane == 0) flags[pt] = valid ? 0 : 1;
}

// Peer buffers of the fused all-gather: device pointers into the other ranks' (and our own)
// gathered lnL vectors (NVLink peer / symmetric memory).  Passed by value.
struct PeerOut {
    double *ptr[RVL_MAX_PEERS];
    int n;
    long long offset;  // element offset of this rank's block inside every gathered vector
};

// combine the per-slice partial sums in slice order (deterministic):
// lnL = (cte - sum_s S1) - sum_s S2  (:80); invalid Keplerian -> -1e30 (:203).
// With peers: the result is also stored straight into every rank's gathered vector over
// NVLink -- the all-gather is fused into the producing kernel.
__global__ void combine_slices_kernel(const double *partial, const int *flags, double *lnl,
                                      long long B, int S, double cte, const PeerOut peers)
{
    const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= B) return;
    double v = -1e30;
    if (!flags[pt]) {
        const double *p = partial + (size_t)pt * S * 2;
        double s1 = 0.0, s2 = 0.0;
        for (int s = 0; s < S; ++s) {
            s1 = rvl::add(s1, p[2 * s]);
            s2 = rvl::add(s2, p[2 * s + 1]);
        }
        v = rvl::sub(rvl::sub(cte, s1), s2);
    }
    lnl[pt] = v;
    for (int r = 0; r < peers.n; ++r) peers.ptr[r][peers.offset + pt] = v;
}

// S == 1 (the likelihood kernel wrote lnL itself): push the finished block to the peers
__global__ void scatter_peers_kernel(const double *lnl, long long B, const PeerOut peers)
{
    const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= B) return;
    const double v = lnl[pt];
    for (int r = 0; r < peers.n; ++r) peers.ptr[r][peers.offset + pt] = v;
}

// ---- prior transform: unit cube -> theta (evidence/ultranest/__init__.py:125-137) -----------
__device__ double ppf_eval(const rvl_prior_desc &pr, const double *tables, double q)
{
    const double p0 = pr.p[0], p1 = pr.p[1];
    switch (pr.kind) {
    case RVL_PRIOR_UNIFORM:  // priors.py:41-42
        return rvl::add(p0, rvl::mul(rvl::sub(p1, p0), q));
    case RVL_PRIOR_JEFFREYS:  // :62-63
        return rvl::mul(p0, pow(__ddiv_rn(p1, p0), q));
    case RVL_PRIOR_MODJEFFREYS:  // :82-83
        return rvl::sub(rvl::mul(p0, pow(rvl::add(1.0, __ddiv_rn(p1, p0)), q)), p0);
    case RVL_PRIOR_UNIFORMFREQ:  // :100-101   xmin / (1 - q*(xmax-xmin)/xmax)
        return __ddiv_rn(p0, rvl::sub(1.0, __ddiv_rn(rvl::mul(q, rvl::sub(p1, p0)), p1)));
    case RVL_PRIOR_TRUNCRAYLEIGH: {  // :249-252
        const double s2 = rvl::mul(p0, p0);
        const double A = rvl::sub(1.0, exp(-__ddiv_rn(rvl::mul(p1, p1), rvl::mul(2.0, s2))));
        return sqrt(rvl::mul(rvl::mul(-2.0, s2), log(rvl::sub(1.0, rvl::mul(q, A)))));
    }
    case RVL_PRIOR_NORMAL:  // scipy.stats.norm(loc, scale).ppf
        return rvl::add(rvl::mul(normcdfinv(q), p1), p0);
    case RVL_PRIOR_LOGNORMAL:  // scipy.stats.lognorm(s, loc, scale).ppf
        return rvl::add(rvl::mul(exp(rvl::mul(p0, normcdfinv(q))), pr.p[2]), p1);
    case RVL_PRIOR_TABLE: {
        // interp1d(cdf, x)(q): hi = clip(searchsorted(cdf, q, 'left'), 1, len-1)
        const double *cdf = tables + pr.table_offset;
        const double *x = cdf + pr.table_len;
        int lo = 0, hi = pr.table_len;
        while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if (__ldg(cdf + mid) < q) lo = mid + 1; else hi = mid;
        }
        int k = max(1, min(pr.table_len - 1, lo));
        const double x0 = __ldg(x + k - 1), x1 = __ldg(x + k);
        const double c0 = __ldg(cdf + k - 1), c1 = __ldg(cdf + k);
        if (p1 != 0.0) {  // p1 = 1: slopes dx/dq follow the knots -> cubic Hermite segment
            const double *mk = x + pr.table_len;
            const double m0 = __ldg(mk + k - 1), m1 = __ldg(mk + k);
            if (m0 > 0.0 && m1 > 0.0) {
                const double h = c1 - c0, t = (q - c0) / h, t2 = t * t, t3 = t2 * t;
                const double h00 = 2.0 * t3 - 3.0 * t2 + 1.0, h10 = t3 - 2.0 * t2 + t;
                const double h01 = -2.0 * t3 + 3.0 * t2, h11 = t3 - t2;
                return h00 * x0 + h10 * h * m0 + h01 * x1 + h11 * h * m1;
            }
        }
        const double slope = __ddiv_rn(rvl::sub(x1, x0), rvl::sub(c1, c0));
        const double y = rvl::add(rvl::mul(slope, rvl::sub(q, c0)), x0);
        return p0 != 0.0 ? exp10(y) : y;  // p0 = 1: Log10Normal, 10**interp (priors.py:144)
    }
    default:
        return nan("");
    }
}

__global__ void prior_transform_kernel(const rvl_prior_desc *priors, const double *tables,
                                       const double *U, double *Theta, long long total, int ndim)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int col = (int)(i % ndim);
    Theta[i] = ppf_eval(priors[col], tables, U[i]);
}

// ---- the reference's native FFI on the device (trueanomaly.h:4) --------------------------------
// One warp per 32 elements; a per-warp constant block in shared memory feeds solve_planet with
// nmot = 0, epoch = 0, M0 = M[i]: mean_anomaly gives 0*(t-0) + M[i] = M[i] exactly, and
// A = 1, Bs = 0, Ce = 0 / A = 0, Bs = 1 turn the RV term into cos(nu) / -sin(nu)... simpler and
// exact: replay the same loop on (sin E, cos E) and finish with atan2.
__global__ void trueanomaly_kernel(const double *M, int n, double ecc, double *nu, int itmax,
                                   double tol, int *caphit)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int ii = min(i, n - 1);  // whole warps stay converged; extra lanes duplicate the last
    const double ec = ecc > 0.99 ? 0.99 : ecc;
    const double m = __ldg(M + ii);
    double E = m, s, c, d = 1e300;
    const bool slow = __any_sync(kFull, !(abs_hi(m) < kHiTrigMax) || !(ec >= -0.99));
    if (slow) { const double2 r = sincos_slow(E); s = r.x; c = r.y; }
    else rvl::sincos_fast(E, s, c);
    int trip = 0;
    for (;;) {
        const bool pa = (fabs(d) > tol) && trip < itmax;
        if (!__any_sync(kFull, pa)) break;
        ++trip;
        double En;
        rvl::newton_step(E, s, c, m, ec, En);
        En = pa ? En : E;
        d = rvl::sub(En, E);
        E = En;
        const int h = abs_hi(d);
        if (__all_sync(kFull, h < kHiTiny)) rvl::advance_tiny(d, s, c);
        else if (__all_sync(kFull, h < kHiSmall)) rvl::advance_small(d, s, c);
        else if (slow || (trip > 2 && __any_sync(kFull, !(abs_hi(E) < kHiTrigMax)))) {
            const double2 r = sincos_slow(E); s = r.x; c = r.y;
        } else rvl::sincos_fast(E, s, c);
    }
    const int cap = (fabs(d) > tol) ? 1 : 0;
    if (i < n) {
        // nu = 2 atan(sqrt((1+e)/(1-e)) tan(E/2))  ==  atan2(sqrt(1-e^2) sin E, cos E - e)
        const double root = sqrt(rvl::mul(rvl::sub(1.0, ec), rvl::add(1.0, ec)));
        nu[i] = atan2(rvl::mul(root, s), rvl::sub(c, ec));
        if (cap) atomicExch(caphit, 1);
    }
}

// ---- register-resident DFMA loop: the FP64 roofline denominator -------------------------------
template <int R>
__global__ void __launch_bounds__(1024) dfma_peak_kernel(double *out, int iters, double, double b)
{
    // R independent chains per thread, 32 warps per SM (grid of tools/ubench.cu)
    double x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = threadIdx.x + r;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            // multiplier from the constant bank (uniform-register operand), addend in a
            // register: the operand form that reached the highest rate in tools/ubench.cu
            for (int r = 0; r < R; ++r) x[r] = __fma_rn(x[r], RVL_K(23), b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) s += x[r];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

thread_local std::string g_create_error;

}  // namespace

// =================================================================================================
// handle + C-ABI
// =================================================================================================
struct rvl_handle {
    int device = 0;
    int sm_count = 0, smem_optin = 0, clock_khz = 0;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // host copies of the staged inputs
    int N = 0, Npad = 0, n_inst = 0;
    std::vector<double> h_t, h_rv, h_err;
    std::vector<int32_t> h_inst;
    std::vector<double> h_linpar[RVL_MAX_LINPAR];
    bool have_data = false, have_model = false, have_priors = false, cols_dirty = true;
    rvl_model_desc model{};

    // device
    double *d_cols = nullptr;
    uint8_t *d_inst = nullptr;
    int ncol = 0;
    rvl_model_desc *d_model = nullptr;
    rvl_prior_desc *d_priors = nullptr;
    double *d_tables = nullptr;
    int prior_ndim = 0;

    // per-call scratch (grown on demand)
    long long cap_B = 0, cap_flags = 0;
    size_t cap_partial = 0;
    double *d_theta = nullptr, *d_u = nullptr, *d_lnl = nullptr, *d_partial = nullptr;
    double *d_consts = nullptr;
    size_t cap_consts = 0;
    int *d_flags = nullptr;
    unsigned long long *d_counters = nullptr;  // 3
    unsigned int *d_work = nullptr;            // sm_count

    // options
    int opt_variant = 0, opt_slices = 0, opt_warps = 0, opt_timing = 0, opt_zero_copy = 1, opt_ilp = 2, opt_min_chunks = 8, opt_items_per_warp = 4;

    // bookkeeping
    uint64_t n_points = 0, n_solves = 0, launches = 0;
    double last_ms = 0.0;
    bool timing_pending = false;
};

namespace {

int fail(rvl_t *h, int code, const std::string &msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
#define CU(h, call)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(h, RVL_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// build the epoch columns in HBM: t, vrad, svrad^2, [(t-tref)/365.25], [linpar...], inst ids
int upload_columns(rvl_t *h)
{
    const int N = h->N, Npad = h->Npad;
    const bool drift = h->have_model && h->model.drift_in_model;
    const int nlin = h->have_model ? h->model.n_linpar : 0;
    const int ncol = 3 + (drift ? 1 : 0) + nlin;
    for (int l = 0; l < nlin; ++l)
        if ((int)h->h_linpar[l].size() != N)
            return fail(h, RVL_ESTATE, "linear-parameter column " + std::to_string(l) + " not set");
    std::vector<double> cols((size_t)ncol * Npad);
    std::vector<uint8_t> ids((size_t)Npad);
    for (int j = 0; j < Npad; ++j) {
        const int k = j < N ? j : N - 1;  // padding duplicates the last epoch (masked in-kernel)
        cols[j] = h->h_t[k];
        cols[(size_t)Npad + j] = h->h_rv[k];
        cols[(size_t)2 * Npad + j] = h->h_err[k] * h->h_err[k];  // svrad**2 (:190-192)
        int c = 3;
        if (drift) cols[(size_t)(c++) * Npad + j] = (h->h_t[k] - h->model.tref) / 365.25;  // :270
        for (int l = 0; l < nlin; ++l) cols[(size_t)(c++) * Npad + j] = h->h_linpar[l][k];
        ids[j] = (uint8_t)h->h_inst[k];
    }
    if (h->d_cols) cudaFree(h->d_cols);
    if (h->d_inst) cudaFree(h->d_inst);
    h->d_cols = nullptr; h->d_inst = nullptr;
    CU(h, cudaMalloc(&h->d_cols, cols.size() * sizeof(double)));
    CU(h, cudaMalloc(&h->d_inst, ids.size()));
    CU(h, cudaMemcpy(h->d_cols, cols.data(), cols.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->d_inst, ids.data(), ids.size(), cudaMemcpyHostToDevice));
    h->ncol = ncol;
    h->cols_dirty = false;
    return RVL_OK;
}

// device staging for the host-buffer entry points
int ensure_io(rvl_t *h, long long B)
{
    if (B <= h->cap_B) return RVL_OK;
    const long long cap = std::max<long long>(B, 1024);
    cudaFree(h->d_theta); cudaFree(h->d_u); cudaFree(h->d_lnl);
    h->d_theta = h->d_u = h->d_lnl = nullptr;
    h->cap_B = 0;
    const size_t row = (size_t)std::max(1, std::max(h->model.ndim, h->prior_ndim));
    CU(h, cudaMalloc(&h->d_theta, (size_t)cap * row * sizeof(double)));
    CU(h, cudaMalloc(&h->d_u, (size_t)cap * row * sizeof(double)));
    CU(h, cudaMalloc(&h->d_lnl, (size_t)cap * sizeof(double)));
    h->cap_B = cap;
    return RVL_OK;
}

// per-slice partial sums (S > 1), invalid-point flags and prepared per-point constants
int ensure_partial(rvl_t *h, long long B, int S, bool prepare, int wstride)
{
    if (S <= 1 && !prepare) return RVL_OK;
    if (B > h->cap_flags) {
        cudaFree(h->d_flags);
        h->d_flags = nullptr; h->cap_flags = 0;
        CU(h, cudaMalloc(&h->d_flags, (size_t)B * sizeof(int)));
        h->cap_flags = B;
    }
    const size_t need = (size_t)B * S;
    if (S > 1 && need > h->cap_partial) {
        cudaFree(h->d_partial);
        h->d_partial = nullptr; h->cap_partial = 0;
        CU(h, cudaMalloc(&h->d_partial, need * 2 * sizeof(double)));
        h->cap_partial = need;
    }
    const size_t needc = (size_t)B * wstride;
    if (prepare && needc > h->cap_consts) {
        cudaFree(h->d_consts);
        h->d_consts = nullptr; h->cap_consts = 0;
        CU(h, cudaMalloc(&h->d_consts, needc * sizeof(double)));
        h->cap_consts = needc;
    }
    return RVL_OK;
}

struct Plan {
    int S, cps, W, U, grid, wstride;
    size_t smem;
};

size_t smem_need(int ncol, int ne, int W, int wstride)
{
    return 128 + kModelBytes + (((size_t)ncol * ne * 8 + ne + 127) / 128) * 128 +
           (size_t)W * wstride * 8;
}

int make_plan(rvl_t *h, long long B, Plan &pl)
{
    const rvl_model_desc &m = h->model;
    const int Ctot = h->Npad / 32;
    int wstride = m.n_planets * kPlanetStride + 2 * m.n_inst + 4 + m.n_linpar;
    wstride = (wstride + 1) & ~1;
    const int U = (h->opt_variant == 0 && h->opt_ilp == 2) ? 2 : 1;
    int W = h->opt_warps > 0 ? h->opt_warps : (U == 2 ? 28 : 32);
    W = std::max(1, std::min(32, W));
    // smallest slice count whose slice fits in shared memory
    int S = 1;
    auto fits = [&](int s) {
        const int cps = (Ctot + s - 1) / s;
        return smem_need(h->ncol, cps * 32, W, wstride) <= (size_t)h->smem_optin;
    };
    while (S < Ctot && !fits(S)) ++S;
    if (!fits(S)) return fail(h, RVL_EINVAL, "epoch chunk does not fit in shared memory");
    if (h->opt_slices > 0) {
        S = std::max(S, std::min(h->opt_slices, Ctot));
    } else {
        // small batches: cut epochs finer so that every warp of the chip gets >= ~4 work items,
        // but keep >= 8 chunks per slice so the per-point setup stays amortised
        const long long warps_total = (long long)h->sm_count * W;
        long long want = ((long long)h->opt_items_per_warp * warps_total + B - 1) / std::max<long long>(B, 1);
        const int maxS = std::max(1, Ctot / h->opt_min_chunks);
        S = std::max<long long>(S, std::min<long long>(want, maxS));
    }
    S = std::min(S, h->sm_count);
    int cps = (Ctot + S - 1) / S;
    S = (Ctot + cps - 1) / cps;  // drop empty trailing slices
    if (!fits(S)) return fail(h, RVL_EINVAL, "slice does not fit in shared memory");
    pl.S = S; pl.cps = cps; pl.W = W; pl.U = U; pl.wstride = wstride;
    pl.grid = std::max(1, h->sm_count / S) * S;
    pl.smem = smem_need(h->ncol, cps * 32, W, wstride);
    return RVL_OK;
}

template <int V, int U, int T>
int launch_lnl_v(rvl_t *h, const KArgs &a, const Plan &pl, cudaStream_t st)
{
    CU(h, cudaFuncSetAttribute(rv_lnl_kernel<V, U, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               h->smem_optin));
    rv_lnl_kernel<V, U, T><<<pl.grid, std::min(pl.W * 32, T), pl.smem, st>>>(a);
    CU(h, cudaGetLastError());
    return RVL_OK;
}

// enqueue: [prepare pass], zero work counters, likelihood kernel, [combine].  All pointers are
// device pointers.  dU != NULL: fused prior transform -- theta is WRITTEN to dTheta by the prepare
// pass and the likelihood is evaluated on exactly those values.
int enqueue_loglike(rvl_t *h, const double *dU, double *dTheta, long long B, double *dlnL,
                    cudaStream_t st, bool timed, const PeerOut *peers = nullptr)
{
    if (!h->have_data || !h->have_model) return fail(h, RVL_ESTATE, "set data and model first");
    if (dU && !h->have_priors) return fail(h, RVL_ESTATE, "set priors first");
    if (dU && h->prior_ndim != h->model.ndim) return fail(h, RVL_ESTATE, "priors and model disagree on ndim");
    if (h->cols_dirty) { int rc = upload_columns(h); if (rc) return rc; }
    if (B <= 0) return RVL_OK;
    if (B > 0xfffffff0LL) return fail(h, RVL_EINVAL, "batch too large for one call (max ~4.29e9)");
    Plan pl;
    int rc = make_plan(h, B, pl);
    if (rc) return rc;
    const bool prepare = pl.S > 1 || dU != nullptr;
    rc = ensure_partial(h, B, pl.S, prepare, pl.wstride);
    if (rc) return rc;
    if (prepare) {
        const int wpb = 8;  // warps (= points) per block
        point_prepare_kernel<<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, 0, st>>>(
            h->d_model, h->d_priors, h->d_tables, dU, dTheta, h->d_consts, h->d_flags,
            h->d_work, h->sm_count, B, pl.wstride);
        CU(h, cudaGetLastError());
        ++h->launches;
    } else {
        CU(h, cudaMemsetAsync(h->d_work, 0, sizeof(unsigned) * (size_t)h->sm_count, st));
    }
    KArgs a{};
    a.model = h->d_model; a.cols = h->d_cols; a.inst = h->d_inst; a.theta = dTheta; a.lnl = dlnL;
    a.partial = h->d_partial; a.flags = h->d_flags; a.counters = h->d_counters; a.work = h->d_work;
    a.consts = prepare ? h->d_consts : nullptr;
    a.B = B; a.cte = -0.5 * h->N * log(2 * M_PI); a.N = h->N; a.Npad = h->Npad; a.ncol = h->ncol;
    a.S = pl.S; a.cps = pl.cps; a.wstride = pl.wstride;
    if (timed) CU(h, cudaEventRecord(h->ev0, st));
    if (h->opt_variant == 1) rc = launch_lnl_v<1, 1, 1024>(h, a, pl, st);
    else if (pl.U == 2 && pl.W <= 16) rc = launch_lnl_v<0, 2, 512>(h, a, pl, st);
    else if (pl.U == 2 && pl.W <= 20) rc = launch_lnl_v<0, 2, 640>(h, a, pl, st);
    else if (pl.U == 2 && pl.W <= 24) rc = launch_lnl_v<0, 2, 768>(h, a, pl, st);
    else if (pl.U == 2 && pl.W <= 28) rc = launch_lnl_v<0, 2, 896>(h, a, pl, st);
    else if (pl.U == 2) rc = launch_lnl_v<0, 2, 1024>(h, a, pl, st);
    else rc = launch_lnl_v<0, 1, 1024>(h, a, pl, st);
    if (rc) return rc;
    if (timed) { CU(h, cudaEventRecord(h->ev1, st)); h->timing_pending = true; }
    ++h->launches;
    if (pl.S > 1) {
        const int tb = 256;
        PeerOut none{};
        combine_slices_kernel<<<(unsigned)((B + tb - 1) / tb), tb, 0, st>>>(
            h->d_partial, h->d_flags, dlnL, B, pl.S, a.cte, peers ? *peers : none);
        CU(h, cudaGetLastError());
        ++h->launches;
    } else if (peers && peers->n > 0) {
        const int tb = 256;
        scatter_peers_kernel<<<(unsigned)((B + tb - 1) / tb), tb, 0, st>>>(dlnL, B, *peers);
        CU(h, cudaGetLastError());
        ++h->launches;
    }
    h->n_points += (uint64_t)B;
    h->n_solves += (uint64_t)B * (uint64_t)h->N * (uint64_t)h->model.n_planets;
    return RVL_OK;
}

int enqueue_transform(rvl_t *h, const double *dU, long long B, double *dTheta, cudaStream_t st)
{
    if (!h->have_priors) return fail(h, RVL_ESTATE, "set priors first");
    if (B <= 0) return RVL_OK;
    const long long total = B * h->prior_ndim;
    const int tb = 256;
    prior_transform_kernel<<<(unsigned)((total + tb - 1) / tb), tb, 0, st>>>(
        h->d_priors, h->d_tables, dU, dTheta, total, h->prior_ndim);
    CU(h, cudaGetLastError());
    ++h->launches;
    return RVL_OK;
}

// Device alias of a pinned (page-locked, mapped) host buffer, or NULL for pageable memory.
// The kernels then read theta / write lnL in place over PCIe: no staging copy, no extra launch.
void *pinned_alias(rvl_t *h, const void *p)
{
    if (!h->opt_zero_copy || !p) return nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost)
        return at.devicePointer;
    cudaGetLastError();  // pageable memory: clear the error some driver versions record
    return nullptr;
}

int finish_timing(rvl_t *h)
{
    if (h->timing_pending) {
        float ms = 0.f;
        CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->last_ms = ms;
        h->timing_pending = false;
    }
    return RVL_OK;
}

}  // namespace

extern "C" {

int rvl_abi_version(void) { return RVL_ABI_VERSION; }

int rvl_create(rvl_t **out, int device)
{
    if (!out) return fail(nullptr, RVL_EINVAL, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RVL_ENODEV,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count=0") +
                        " (librvlnl has no CPU fallback)");
    if (device < 0) cudaGetDevice(&device);
    if (device >= ndev) return fail(nullptr, RVL_ENODEV, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return fail(nullptr, RVL_ECUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, RVL_ENODEV,
                    "device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                        "; librvlnl is built for sm_100a only");
    rvl_t *h = new (std::nothrow) rvl_handle();
    if (!h) return fail(nullptr, RVL_ENOMEM, "out of host memory");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    h->clock_khz = prop.clockRate;
    DevGuard g(device);
    auto bail = [&](const char *what, cudaError_t ce) {
        std::string msg = std::string(what) + ": " + cudaGetErrorString(ce);
        rvl_destroy(h);
        return fail(nullptr, RVL_ECUDA, msg);
    };
    cudaError_t ce;
    if ((ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", ce);
    if ((ce = cudaEventCreate(&h->ev0)) != cudaSuccess) return bail("event", ce);
    if ((ce = cudaEventCreate(&h->ev1)) != cudaSuccess) return bail("event", ce);
    if ((ce = cudaMalloc(&h->d_counters, 3 * sizeof(unsigned long long))) != cudaSuccess) return bail("malloc", ce);
    if ((ce = cudaMemset(h->d_counters, 0, 3 * sizeof(unsigned long long))) != cudaSuccess) return bail("memset", ce);
    if ((ce = cudaMalloc(&h->d_work, sizeof(unsigned) * (size_t)h->sm_count)) != cudaSuccess) return bail("malloc", ce);
    if ((ce = cudaMalloc(&h->d_model, sizeof(rvl_model_desc))) != cudaSuccess) return bail("malloc", ce);
    *out = h;
    return RVL_OK;
}

void rvl_destroy(rvl_t *h)
{
    if (!h) return;
    DevGuard g(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_cols); cudaFree(h->d_inst); cudaFree(h->d_model); cudaFree(h->d_priors);
    cudaFree(h->d_tables); cudaFree(h->d_theta); cudaFree(h->d_u); cudaFree(h->d_lnl);
    cudaFree(h->d_partial); cudaFree(h->d_flags); cudaFree(h->d_counters); cudaFree(h->d_work);
    cudaFree(h->d_consts);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char *rvl_last_error(const rvl_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int rvl_set_data(rvl_t *h, const double *t, const double *rv, const double *err,
                 const int32_t *inst, int32_t n, int32_t n_inst)
{
    if (!h) return RVL_EINVAL;
    if (!t || !rv || !err || !inst) return fail(h, RVL_EINVAL, "NULL data pointer");
    if (n <= 0) return fail(h, RVL_EINVAL, "n must be positive");
    if (n_inst <= 0 || n_inst > RVL_MAX_INST)
        return fail(h, RVL_EINVAL, "n_inst must be in [1, " + std::to_string(RVL_MAX_INST) + "]");
    for (int j = 0; j < n; ++j)
        if (inst[j] < 0 || inst[j] >= n_inst) return fail(h, RVL_EINVAL, "instrument id out of range");
    h->N = n; h->Npad = (n + 31) / 32 * 32; h->n_inst = n_inst;
    h->h_t.assign(t, t + n); h->h_rv.assign(rv, rv + n); h->h_err.assign(err, err + n);
    h->h_inst.assign(inst, inst + n);
    for (auto &c : h->h_linpar) c.clear();
    h->have_data = true; h->cols_dirty = true;
    return RVL_OK;
}

int rvl_set_linpar(rvl_t *h, int32_t idx, const double *col, int32_t n)
{
    if (!h) return RVL_EINVAL;
    if (!h->have_data) return fail(h, RVL_ESTATE, "set data first");
    if (idx < 0 || idx >= RVL_MAX_LINPAR || !col || n != h->N)
        return fail(h, RVL_EINVAL, "bad linear-parameter column");
    h->h_linpar[idx].assign(col, col + n);
    h->cols_dirty = true;
    return RVL_OK;
}

static bool param_ok(const rvl_param &p, int ndim) { return p.slot >= -1 && p.slot < ndim; }

int rvl_set_model(rvl_t *h, const rvl_model_desc *d)
{
    if (!h) return RVL_EINVAL;
    if (!d) return fail(h, RVL_EINVAL, "desc is NULL");
    if (d->abi_version != RVL_ABI_VERSION) return fail(h, RVL_EINVAL, "abi_version mismatch");
    if (d->ndim < 0 || d->ndim > RVL_MAX_DIM) return fail(h, RVL_EINVAL, "ndim out of range");
    if (d->n_planets < 0 || d->n_planets > RVL_MAX_PLANETS) return fail(h, RVL_EINVAL, "n_planets out of range");
    if (d->n_inst <= 0 || d->n_inst > RVL_MAX_INST) return fail(h, RVL_EINVAL, "n_inst out of range");
    if (d->n_linpar < 0 || d->n_linpar > RVL_MAX_LINPAR) return fail(h, RVL_EINVAL, "n_linpar out of range");
    if (h->have_data && d->n_inst != h->n_inst) return fail(h, RVL_EINVAL, "n_inst differs from the staged data");
    if (d->itmax < 1) return fail(h, RVL_EINVAL, "itmax must be >= 1");
    if (!(d->tol > 0)) return fail(h, RVL_EINVAL, "tol must be positive");
    for (int p = 0; p < d->n_planets; ++p) {
        const rvl_planet_desc &pl = d->planet[p];
        if (!param_ok(pl.amp, d->ndim) || !param_ok(pl.period, d->ndim) || !param_ok(pl.e1, d->ndim) ||
            !param_ok(pl.e2, d->ndim) || !param_ok(pl.phase, d->ndim) || !param_ok(pl.epoch, d->ndim))
            return fail(h, RVL_EINVAL, "planet parameter slot out of range");
        if (pl.ecc_mode < 0 || pl.ecc_mode > 2 || pl.phase_mode < 0 || pl.phase_mode > 1)
            return fail(h, RVL_EINVAL, "bad parametrisation enum");
    }
    for (int i = 0; i < d->n_inst; ++i)
        if (!param_ok(d->offset[i], d->ndim) || !param_ok(d->jitter[i], d->ndim))
            return fail(h, RVL_EINVAL, "instrument parameter slot out of range");
    for (int i = 0; i < 4; ++i)
        if (!param_ok(d->drift[i], d->ndim)) return fail(h, RVL_EINVAL, "drift slot out of range");
    for (int i = 0; i < d->n_linpar; ++i)
        if (!param_ok(d->linpar[i], d->ndim)) return fail(h, RVL_EINVAL, "linpar slot out of range");
    DevGuard g(h->device);
    h->model = *d;
    CU(h, cudaMemcpy(h->d_model, d, sizeof(*d), cudaMemcpyHostToDevice));
    h->have_model = true; h->cols_dirty = true;
    // scratch rows depend on ndim
    h->cap_B = 0;
    return RVL_OK;
}

int rvl_set_priors(rvl_t *h, const rvl_prior_desc *priors, int32_t ndim, const double *tables,
                   int64_t n_table_doubles)
{
    if (!h) return RVL_EINVAL;
    if (!priors || ndim <= 0 || ndim > RVL_MAX_DIM) return fail(h, RVL_EINVAL, "bad priors / ndim");
    for (int i = 0; i < ndim; ++i) {
        const rvl_prior_desc &p = priors[i];
        if (p.kind < 0 || p.kind > RVL_PRIOR_TABLE) return fail(h, RVL_EINVAL, "unknown prior kind");
        if (p.kind == RVL_PRIOR_TABLE &&
            (p.table_len < 2 || p.table_offset < 0 || !tables ||
             p.table_offset + (p.p[1] != 0.0 ? 3LL : 2LL) * p.table_len > n_table_doubles))
            return fail(h, RVL_EINVAL, "prior table out of bounds");
    }
    DevGuard g(h->device);
    cudaFree(h->d_priors); cudaFree(h->d_tables);
    h->d_priors = nullptr; h->d_tables = nullptr;
    CU(h, cudaMalloc(&h->d_priors, sizeof(rvl_prior_desc) * (size_t)ndim));
    CU(h, cudaMemcpy(h->d_priors, priors, sizeof(rvl_prior_desc) * (size_t)ndim, cudaMemcpyHostToDevice));
    if (n_table_doubles > 0 && tables) {
        CU(h, cudaMalloc(&h->d_tables, sizeof(double) * (size_t)n_table_doubles));
        CU(h, cudaMemcpy(h->d_tables, tables, sizeof(double) * (size_t)n_table_doubles, cudaMemcpyHostToDevice));
    }
    h->prior_ndim = ndim; h->have_priors = true;
    h->cap_B = 0;
    return RVL_OK;
}

int rvl_set_option(rvl_t *h, const char *name, int64_t value)
{
    if (!h || !name) return RVL_EINVAL;
    const std::string n(name);
    if (n == "variant") { if (value < 0 || value > 1) return fail(h, RVL_EINVAL, "variant in {0,1}"); h->opt_variant = (int)value; }
    else if (n == "slices") h->opt_slices = (int)std::max<int64_t>(0, value);
    else if (n == "warps") h->opt_warps = (int)std::max<int64_t>(0, std::min<int64_t>(32, value));
    else if (n == "timing") h->opt_timing = value != 0;
    else if (n == "zero_copy") h->opt_zero_copy = value != 0;
    else if (n == "min_chunks") h->opt_min_chunks = (int)std::max<int64_t>(1, value);
    else if (n == "items_per_warp") h->opt_items_per_warp = (int)std::max<int64_t>(1, value);
    else if (n == "ilp") { if (value != 1 && value != 2) return fail(h, RVL_EINVAL, "ilp in {1,2}"); h->opt_ilp = (int)value; }
    else return fail(h, RVL_EINVAL, "unknown option " + n);
    return RVL_OK;
}

int rvl_loglike_dev(rvl_t *h, const double *dTheta, int64_t B, double *dlnL, void *stream)
{
    if (!h) return RVL_EINVAL;
    if (B < 0 || (B > 0 && (!dTheta || !dlnL))) return fail(h, RVL_EINVAL, "bad arguments");
    DevGuard g(h->device);
    return enqueue_loglike(h, nullptr, const_cast<double *>(dTheta), B, dlnL, (cudaStream_t)stream,
                           h->opt_timing != 0);
}

int rvl_loglike_dev_scatter(rvl_t *h, const double *dTheta, int64_t B, double *dlnL,
                            const uint64_t *peer_ptrs, int32_t n_peers, int64_t offset,
                            void *stream)
{
    if (!h) return RVL_EINVAL;
    if (B < 0 || (B > 0 && (!dTheta || !dlnL))) return fail(h, RVL_EINVAL, "bad arguments");
    if (n_peers < 0 || n_peers > RVL_MAX_PEERS || (n_peers > 0 && !peer_ptrs) || offset < 0)
        return fail(h, RVL_EINVAL, "bad peer list");
    PeerOut po{};
    po.n = n_peers;
    po.offset = offset;
    for (int r = 0; r < n_peers; ++r) po.ptr[r] = reinterpret_cast<double *>(peer_ptrs[r]);
    DevGuard g(h->device);
    return enqueue_loglike(h, nullptr, const_cast<double *>(dTheta), B, dlnL, (cudaStream_t)stream,
                           h->opt_timing != 0, &po);
}

int rvl_transform_dev(rvl_t *h, const double *dU, int64_t B, double *dTheta, void *stream)
{
    if (!h) return RVL_EINVAL;
    if (B < 0 || (B > 0 && (!dU || !dTheta))) return fail(h, RVL_EINVAL, "bad arguments");
    DevGuard g(h->device);
    return enqueue_transform(h, dU, B, dTheta, (cudaStream_t)stream);
}

int rvl_transform_loglike_dev(rvl_t *h, const double *dU, int64_t B, double *dTheta, double *dlnL,
                              void *stream)
{
    if (!h) return RVL_EINVAL;
    if (B < 0 || (B > 0 && (!dU || !dTheta || !dlnL))) return fail(h, RVL_EINVAL, "bad arguments");
    if (h->have_model && h->have_priors && h->prior_ndim != h->model.ndim)
        return fail(h, RVL_ESTATE, "priors and model disagree on ndim");
    DevGuard g(h->device);
    return enqueue_loglike(h, dU, dTheta, B, dlnL, (cudaStream_t)stream, h->opt_timing != 0);
}

int rvl_loglike(rvl_t *h, const double *Theta, int64_t B, double *lnL)
{
    if (!h) return RVL_EINVAL;
    if (B < 0 || (B > 0 && (!Theta || !lnL))) return fail(h, RVL_EINVAL, "bad arguments");
    if (!h->have_data || !h->have_model) return fail(h, RVL_ESTATE, "set data and model first");
    if (B == 0) return RVL_OK;
    DevGuard g(h->device);
    int rc = ensure_io(h, B);
    if (rc) return rc;
    // Pinned caller buffers are used in place (zero-copy) when the once-per-point prepare pass
    // reads theta (one coalesced read per row); otherwise theta is staged with a copy.
    Plan pl;
    rc = make_plan(h, B, pl);
    if (rc) return rc;
    const size_t nb = (size_t)B * h->model.ndim * sizeof(double);
    double *th_dev = pl.S > 1 ? (double *)pinned_alias(h, Theta) : nullptr;
    double *out_dev = (double *)pinned_alias(h, lnL);
    if (!th_dev) {
        th_dev = h->d_theta;
        if (nb) CU(h, cudaMemcpyAsync(h->d_theta, Theta, nb, cudaMemcpyHostToDevice, h->stream));
    }
    rc = enqueue_loglike(h, nullptr, th_dev, B, out_dev ? out_dev : h->d_lnl, h->stream,
                         h->opt_timing != 0);
    if (rc) return rc;
    if (!out_dev)
        CU(h, cudaMemcpyAsync(lnL, h->d_lnl, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return finish_timing(h);
}

int rvl_transform(rvl_t *h, const double *U, int64_t B, double *Theta)
{
    if (!h) return RVL_EINVAL;
    if (B < 0 || (B > 0 && (!U || !Theta))) return fail(h, RVL_EINVAL, "bad arguments");
    if (!h->have_priors) return fail(h, RVL_ESTATE, "set priors first");
    if (B == 0) return RVL_OK;
    DevGuard g(h->device);
    int rc = ensure_io(h, B);
    if (rc) return rc;
    const size_t nb = (size_t)B * h->prior_ndim * sizeof(double);
    CU(h, cudaMemcpyAsync(h->d_u, U, nb, cudaMemcpyHostToDevice, h->stream));
    rc = enqueue_transform(h, h->d_u, B, h->d_theta, h->stream);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(Theta, h->d_theta, nb, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return RVL_OK;
}

int rvl_transform_loglike(rvl_t *h, const double *U, int64_t B, double *Theta, double *lnL)
{
    if (!h) return RVL_EINVAL;
    if (B < 0 || (B > 0 && (!U || !lnL))) return fail(h, RVL_EINVAL, "bad arguments");
    if (!h->have_data || !h->have_model || !h->have_priors)
        return fail(h, RVL_ESTATE, "set data, model and priors first");
    if (h->prior_ndim != h->model.ndim) return fail(h, RVL_ESTATE, "priors and model disagree on ndim");
    if (B == 0) return RVL_OK;
    DevGuard g(h->device);
    int rc = ensure_io(h, B);
    if (rc) return rc;
    const size_t nb = (size_t)B * h->prior_ndim * sizeof(double);
    // pinned caller buffers are read / written in place by the prepare pass and the kernels
    const double *u_dev = (const double *)pinned_alias(h, U);
    double *th_dev = Theta ? (double *)pinned_alias(h, Theta) : nullptr;
    double *out_dev = (double *)pinned_alias(h, lnL);
    if (!u_dev) {
        CU(h, cudaMemcpyAsync(h->d_u, U, nb, cudaMemcpyHostToDevice, h->stream));
        u_dev = h->d_u;
    }
    rc = enqueue_loglike(h, u_dev, th_dev ? th_dev : h->d_theta, B, out_dev ? out_dev : h->d_lnl,
                         h->stream, h->opt_timing != 0);
    if (rc) return rc;
    if (Theta && !th_dev)
        CU(h, cudaMemcpyAsync(Theta, h->d_theta, nb, cudaMemcpyDeviceToHost, h->stream));
    if (!out_dev)
        CU(h, cudaMemcpyAsync(lnL, h->d_lnl, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return finish_timing(h);
}

int rvl_trueanomaly(rvl_t *h, const double *M, int32_t n, double ecc, double *nu,
                    int32_t niterationmax, double tol)
{
    if (!h) return RVL_EINVAL;
    if (n < 0 || (n > 0 && (!M || !nu)) || niterationmax < 1) return fail(h, RVL_EINVAL, "bad arguments");
    if (n == 0) return 0;
    DevGuard g(h->device);
    double *dM = nullptr, *dnu = nullptr;
    int *dcap = nullptr;
    CU(h, cudaMalloc(&dM, sizeof(double) * (size_t)n));
    CU(h, cudaMalloc(&dnu, sizeof(double) * (size_t)n));
    CU(h, cudaMalloc(&dcap, sizeof(int)));
    CU(h, cudaMemsetAsync(dcap, 0, sizeof(int), h->stream));
    CU(h, cudaMemcpyAsync(dM, M, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    const int tb = 128;
    trueanomaly_kernel<<<(n + tb - 1) / tb, tb, 0, h->stream>>>(dM, n, ecc, dnu, niterationmax, tol, dcap);
    ++h->launches;
    int cap = 0;
    cudaError_t e1 = cudaGetLastError();
    cudaMemcpyAsync(nu, dnu, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(&cap, dcap, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e2 = cudaStreamSynchronize(h->stream);
    cudaFree(dM); cudaFree(dnu); cudaFree(dcap);
    if (e1 != cudaSuccess) return fail(h, RVL_ECUDA, cudaGetErrorString(e1));
    if (e2 != cudaSuccess) return fail(h, RVL_ECUDA, cudaGetErrorString(e2));
    return cap ? -1 : 0;  // same 0 / -1 convention as trueanomaly.c:32-40
}

int rvl_counters(rvl_t *h, rvl_counters_t *out)
{
    if (!h || !out) return RVL_EINVAL;
    DevGuard g(h->device);
    unsigned long long c[3] = {0, 0, 0};
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaMemcpy(c, h->d_counters, sizeof c, cudaMemcpyDeviceToHost));
    out->n_points = h->n_points; out->n_solves = h->n_solves;
    out->n_newton_iters = c[0]; out->n_cap_hits = c[1]; out->n_invalid = c[2];
    return RVL_OK;
}

int rvl_reset_counters(rvl_t *h)
{
    if (!h) return RVL_EINVAL;
    DevGuard g(h->device);
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaMemset(h->d_counters, 0, 3 * sizeof(unsigned long long)));
    h->n_points = 0; h->n_solves = 0;
    return RVL_OK;
}

int rvl_last_kernel_ms(rvl_t *h, double *ms)
{
    if (!h || !ms) return RVL_EINVAL;
    if (h->timing_pending) {  // *_dev launches: wait for the kernel's end event
        DevGuard g(h->device);
        CU(h, cudaEventSynchronize(h->ev1));
        int rc = finish_timing(h);
        if (rc) return rc;
    }
    *ms = h->last_ms;
    return RVL_OK;
}

int rvl_launch_count(rvl_t *h, uint64_t *n)
{
    if (!h || !n) return RVL_EINVAL;
    *n = h->launches;
    return RVL_OK;
}

int rvl_fp64_peak(rvl_t *h, double *tflops)
{
    if (!h || !tflops) return RVL_EINVAL;
    DevGuard g(h->device);
    const int blocks = h->sm_count, tb = 1024, iters = 4000;
    double *out = nullptr;
    CU(h, cudaMalloc(&out, sizeof(double) * (size_t)blocks * tb));
    double best = 0.0;
    for (int rep = 0; rep < 9; ++rep) {  // best over chain counts 2, 4, 8 and repetitions
        const int R = rep % 3 == 0 ? 2 : (rep % 3 == 1 ? 4 : 8);
        CU(h, cudaEventRecord(h->ev0, h->stream));
        if (R == 2) dfma_peak_kernel<2><<<blocks, tb, 0, h->stream>>>(out, iters, 0.999999, 1e-9);
        else if (R == 4) dfma_peak_kernel<4><<<blocks, tb, 0, h->stream>>>(out, iters, 0.999999, 1e-9);
        else dfma_peak_kernel<8><<<blocks, tb, 0, h->stream>>>(out, iters, 0.999999, 1e-9);
        CU(h, cudaEventRecord(h->ev1, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        ++h->launches;
        float ms = 0.f;
        CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        const double flops = 2.0 * 8.0 * R * (double)iters * (double)blocks * tb;
        if (rep >= 3) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaFree(out);
    *tflops = best;
    return RVL_OK;
}

int rvl_device_info(rvl_t *h, int32_t *sm_count, int32_t *smem_optin, int32_t *clock_khz)
{
    if (!h) return RVL_EINVAL;
    if (sm_count) *sm_count = h->sm_count;
    if (smem_optin) *smem_optin = h->smem_optin;
    if (clock_khz) *clock_khz = h->clock_khz;
    return RVL_OK;
}

}  // extern "C"
