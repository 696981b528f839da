/*
 * rvlnl_oracle.c — TEST INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * A plain-C, scalar CPU restatement of the `evidence` RV log-likelihood path, used as the
 * checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under evidence_b200/ may call into this file.
 *
 * Parity status: PINNED against outputs of the reference itself (live import of
 * /root/reference/evidence in the build container, oracle/make_golden.py) and against the
 * reference's own C solver compiled from where it lies (oracle/_ref/trueanomaly.so).  The
 * reference's own test-suite holds no golden lnL value for this path, so the
 * pins are the committed fixtures under tests/golden/.
 *
 * Each function cites the reference lines it follows (paths relative to the reference
 * checkout).  Arithmetic is kept in the reference's operation order (no FMA contraction:
 * build with -ffp-contract=off), so it reproduces the reference to the last few ulps.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rvlnl.h"

/* ------------------------------------------------------------------------------------------
 * Kepler solve + true anomaly.   evidence/rvmodel/trueanomaly.c:8-41
 * Newton from E = M (un-reduced), per-element stop once |E - E_prev| <= tol after at least
 * one step; eccentricity clamped to 0.99 inside the solver only; on the iteration cap the
 * whole call returns -1 and leaves the remaining nu untouched.
 * `iters` (optional) accumulates the number of Newton steps taken.
 * ---------------------------------------------------------------------------------------- */
/* The sin / cos of the Newton loop can be swapped at compile time (-D'ORC_SIN(x)=...'): only
 * oracle/experiments/libm_sensitivity.py does that, to measure how far the REFERENCE moves when
 * its libm changes.  Default: the C library's, like the reference. */
#ifdef ORC_PERTURB
/* a libm that is one ulp off on 1/16 of its results (deterministic in the value) */
static double orc_perturb(double y)
{
    union { double d; unsigned long long u; } w;
    w.d = y;
    if (((w.u * 0x9E3779B97F4A7C15ull) >> 60) == 0) w.u ^= 1ull;
    return w.d;
}
#endif
#ifndef ORC_SIN
#define ORC_SIN(x) sin(x)
#endif
#ifndef ORC_COS
#define ORC_COS(x) cos(x)
#endif
int orc_trueanomaly(const double *M, int n, double ecc, double *nu, int itmax, double tol,
                    long long *iters)
{
    const double e = (ecc > 0.99) ? 0.99 : ecc;
    long long total = 0;
    for (int j = 0; j < n; ++j) {
        const double m = M[j];
        double cur = m, prev;
        int k = 0;
        do {
            prev = cur;
            const double f = prev - e * ORC_SIN(prev) - m;
            const double fp = 1 - e * ORC_COS(prev);
            cur = prev - f / fp;
            ++k;
            if (k >= itmax) {
                if (iters) *iters += total + k;
                return -1;
            }
        } while (fabs(cur - prev) > tol);
        total += k;
        nu[j] = 2. * atan(sqrt((1. + e) / (1. - e)) * tan(cur / 2.));
    }
    if (iters) *iters += total;
    return 0;
}

/* numpy's pairwise summation of a contiguous float64 vector (what np.sum does in
 * BaseModel.logL, evidence/rvmodel/__init__.py:80): blocks of <=128 with 8 running
 * accumulators, halved recursively above that. */
static double np_pairwise_sum(const double *a, ptrdiff_t n)
{
    if (n < 8) {
        double r = 0.;
        for (ptrdiff_t i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        ptrdiff_t i;
        for (i = 0; i < 8; ++i) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    ptrdiff_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

typedef struct {
    int n;
    int n_inst;
    const double *t;
    const double *rv;
    const double *err;
    const int32_t *inst;
    const double *linpar[RVL_MAX_LINPAR];
} orc_data;

static inline double par(const rvl_param *p, const double *theta)
{
    return p->slot >= 0 ? theta[p->slot] : p->value;
}

/* One point.  evidence/rvmodel/__init__.py:157-219 (log_likelihood), :343-385 (kep_rv),
 * :388-463 (modelk), :222-273 (drift), :59-80 (logL).  `work` holds 4*n doubles. */
double orc_loglike_one(const rvl_model_desc *m, const orc_data *d, const double *theta,
                       double *work, long long *iters, long long *caphits)
{
    const int n = d->n;
    double *rvm = work, *var = work + n, *ma = work + 2 * n, *nu = work + 3 * n;

    /* :181-192 offsets and jitter-inflated variance */
    for (int j = 0; j < n; ++j) {
        const int i = d->inst[j];
        rvm[j] = 0. + par(&m->offset[i], theta);
        const double s2 = d->err[j] * d->err[j];
        if (m->jitter_in_model) {
            const double jit = par(&m->jitter[i], theta);
            var[j] = s2 + jit * jit;
        } else {
            var[j] = s2;
        }
    }

    /* :195-203, :369-383 planets; rv_planet.sum(axis=0) adds planet rows in order */
    if (m->n_planets > 0) {
        double *acc = (double *)calloc((size_t)n, sizeof(double));
        for (int p = 0; p < m->n_planets; ++p) {
            const rvl_planet_desc *pl = &m->planet[p];
            double K = par(&pl->amp, theta);
            if (pl->amp_is_log) K = exp(K); /* :412-415 */
            double P = par(&pl->period, theta);
            if (pl->period_is_log) P = exp(P); /* :417-420 */
            double ecc, omega;
            const double a = par(&pl->e1, theta), b = par(&pl->e2, theta);
            if (pl->ecc_mode == RVL_ECC_SECOS_SESIN) { /* :425-431 */
                ecc = a * a + b * b;
                omega = atan2(b, a);
                if (ecc > 1) { free(acc); return -1e30; }
            } else if (pl->ecc_mode == RVL_ECC_ECOS_ESIN) { /* :433-439 */
                ecc = sqrt(a * a + b * b);
                omega = atan2(b, a);
                if (ecc > 1) { free(acc); return -1e30; }
            } else { /* :441-447 */
                ecc = a;
                omega = b;
            }
            double M0 = par(&pl->phase, theta); /* :449-454 */
            if (pl->phase_mode == RVL_PHASE_ML0) M0 = M0 - omega;
            const double epoch = par(&pl->epoch, theta);
            const double nmot = 2 * M_PI / P; /* :459 */
            for (int j = 0; j < n; ++j) ma[j] = nmot * (d->t[j] - epoch) + M0;
            memset(nu, 0, (size_t)n * sizeof(double)); /* :488 */
            if (orc_trueanomaly(ma, n, ecc, nu, m->itmax, m->tol, iters) != 0 && caphits)
                ++*caphits; /* :490 return code ignored */
            const double ecw = ecc * cos(omega);
            for (int j = 0; j < n; ++j) { /* :463 */
                const double v = K * (cos(nu[j] + omega) + ecw);
                acc[j] = (p == 0) ? (0. + v) : (acc[j] + v);
            }
        }
        for (int j = 0; j < n; ++j) rvm[j] += acc[j];
        free(acc);
    }

    /* :206-207, :256-271 drift */
    if (m->drift_in_model) {
        const double lin = par(&m->drift[0], theta), quad = par(&m->drift[1], theta);
        const double cub = par(&m->drift[2], theta), quar = par(&m->drift[3], theta);
        for (int j = 0; j < n; ++j) {
            const double tt = (d->t[j] - m->tref) / 365.25;
            rvm[j] += lin * tt + quad * (tt * tt) + cub * pow(tt, 3.) + quar * pow(tt, 4.);
        }
    }

    /* :210-212 linear parameters */
    for (int l = 0; l < m->n_linpar; ++l) {
        const double c = par(&m->linpar[l], theta);
        for (int j = 0; j < n; ++j) rvm[j] += c * d->linpar[l][j];
    }

    /* :215-217, :78-80 */
    for (int j = 0; j < n; ++j) {
        const double r = d->rv[j] - rvm[j];
        ma[j] = log(sqrt(var[j]));
        nu[j] = (r * r) / (2 * var[j]);
    }
    const double cte = -0.5 * n * log(2 * M_PI);
    return cte - np_pairwise_sum(ma, n) - np_pairwise_sum(nu, n);
}

/* Batch driver: rows [0, B) of theta (row stride ndim).  Returns 0. */
int orc_loglike_batch(const rvl_model_desc *m, const double *t, const double *rv,
                      const double *err, const int32_t *inst, int n, int n_inst,
                      const double *const *linpar, const double *theta, long long B,
                      double *out, long long *iters, long long *caphits)
{
    orc_data d;
    memset(&d, 0, sizeof d);
    d.n = n; d.n_inst = n_inst; d.t = t; d.rv = rv; d.err = err; d.inst = inst;
    for (int l = 0; l < m->n_linpar && l < RVL_MAX_LINPAR; ++l) d.linpar[l] = linpar[l];
    double *work = (double *)malloc(sizeof(double) * 4 * (size_t)(n > 0 ? n : 1));
    if (!work) return -1;
    for (long long b = 0; b < B; ++b)
        out[b] = orc_loglike_one(m, &d, theta + b * m->ndim, work, iters, caphits);
    free(work);
    return 0;
}

/* Closed-form inverse CDFs.  evidence/priors.py:41-42, 62-63, 82-83, 100-101, 249-252 */
double orc_ppf_closed(int kind, const double *p, double q)
{
    switch (kind) {
    case RVL_PRIOR_UNIFORM: return p[0] + (p[1] - p[0]) * q;
    case RVL_PRIOR_JEFFREYS: return p[0] * pow(p[1] / p[0], q);
    case RVL_PRIOR_MODJEFFREYS: return p[0] * pow(1 + p[1] / p[0], q) - p[0];
    case RVL_PRIOR_UNIFORMFREQ: return p[0] / (1 - q * (p[1] - p[0]) / p[1]);
    case RVL_PRIOR_TRUNCRAYLEIGH: {
        const double A = 1 - exp(-(p[1] * p[1]) / (2 * (p[0] * p[0])));
        return sqrt(-2 * (p[0] * p[0]) * log(1 - (q * A)));
    }
    default: return NAN;
    }
}

/* Piecewise-linear inverse CDF through (cdf_k, x_k), the arithmetic of scipy's interp1d
 * (kind='linear') as used by evidence/priors.py:124,202,228,287,326,354:
 * slope = (y_hi - y_lo)/(x_hi - x_lo); y = slope*(x_new - x_lo) + y_lo, with
 * hi = clip(searchsorted(x, x_new), 1, len-1), lo = hi - 1. */
double orc_ppf_table(const double *cdf, const double *x, int len, double q)
{
    int lo = 0, hi = len; /* np.searchsorted side='left' */
    while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        if (cdf[mid] < q) lo = mid + 1; else hi = mid;
    }
    int k = lo;
    if (k < 1) k = 1;
    if (k > len - 1) k = len - 1;
    const double slope = (x[k] - x[k - 1]) / (cdf[k] - cdf[k - 1]);
    return slope * (q - cdf[k - 1]) + x[k - 1];
}
