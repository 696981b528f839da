// rvl_math.h — the FP64 arithmetic core of the batched Kepler / RV log-likelihood kernel.
//
// Every function here is plain IEEE-754 double arithmetic written with explicit, never
// contracted operations (mul/add/sub/fma helpers), so that the exact same source compiles for
// sm_100a (the product: rvlnl.cu) and for the host (tests/host_emul, which replays the kernel's
// arithmetic lane by lane on the CPU to validate numerics without a GPU).  The host build is a
// test harness only; the product has no CPU path.
//
// Reference semantics reproduced (paths relative to the reference checkout):
//   evidence/rvmodel/trueanomaly.c:8-41        Newton from E=M, |dE|<=tol stop, e clamp 0.99
//   evidence/rvmodel/__init__.py:459           M = 2*pi/P * (t - epoch) + M0   (no FMA!)
//   evidence/rvmodel/__init__.py:463           rv = K (cos(nu + w) + e cos w)
//
// Why the operation order matters: M and E live near 1e4 rad where one
// ulp is 1.8e-12, and lnL responds to phase errors with a gain of ~1e4.  So the roundings that
// happen ON that 1.8e-12 grid -- forming M, forming E - e sin E, and the Newton update
// E - f/f' -- are reproduced operation for operation.  Everything that is small compared with
// the grid (the quotient f/f', sin/cos values, the rotation to the true anomaly) only has to
// be accurate to an ulp or two, which leaves room to make it cheap:
//   * f/f' is f * rcp(f') with a 2-step Newton reciprocal (no IEEE division),
//   * after a Newton step of size d, (sin E, cos E) is advanced by an angle-addition with a
//     short series in d when every lane of the warp has |d| small (warp-uniform choice),
//   * cos(nu), sin(nu) come from the closed form in (sin E, cos E) instead of atan(tan()).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RVL_HD __host__ __device__ __forceinline__
#else
#define RVL_HD inline
#endif

namespace rvl {

// ---- never-contracted primitives --------------------------------------------------------
RVL_HD double mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;  // host harness is built with -ffp-contract=off
#endif
}
RVL_HD double add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
RVL_HD double sub(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
RVL_HD double fma_(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return ::fma(a, b, c);
#endif
}
RVL_HD int32_t lo32(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    return (int32_t)(uint32_t)u;
#endif
}
RVL_HD int32_t hi32(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    return (int32_t)(uint32_t)(u >> 32);
#endif
}
RVL_HD double from_hilo(int32_t hi, int32_t lo)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}

// ---- reciprocal: hardware seed + ONE cubic refinement (3 DFMA), <= ~1 ulp ------------------
// y0 = MUFU.RCP64H(x) has a relative error e = 1 - x y0 of at most ~2^-20.  1/x = y0 / (1 - e)
// = y0 (1 + e + e^2 + e^3 + ...): y = y0 + y0 (e + e^2) leaves e^3 <= 2^-60, below the rounding
// of the last FMA.  (Two quadratic Newton steps reach the same accuracy with 4 DFMA.)
// Valid for normal, finite x (denominators 1 - e cos E in [0.01, 2], variances).
#ifndef RVL_RCP_CUBIC
#define RVL_RCP_CUBIC 1
#endif
RVL_HD double rcp_seed(double x)
{
    double y;
#if defined(__CUDA_ARCH__)
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // MUFU.RCP64H, ~20 bits
#else
    {  // host harness stand-in with the hardware's seed accuracy (20 bits, truncated)
        float f = 1.0f / (float)x;
        uint32_t u;
        memcpy(&u, &f, 4);
        u &= 0xfffffff0u;
        memcpy(&f, &u, 4);
        y = (double)f;
    }
#endif
    return y;
}
RVL_HD double rcp(double x)
{
    double y = rcp_seed(x);
#if RVL_RCP_CUBIC
    const double e = fma_(-x, y, 1.0);
    const double t = fma_(e, e, e);
    return fma_(y, t, y);
#else
    double e = fma_(-x, y, 1.0);
    y = fma_(y, e, y);
    e = fma_(-x, y, 1.0);
    y = fma_(y, e, y);
    return y;
#endif
}

// ---- constant table -----------------------------------------------------------------------
// On the device the coefficients live in the constant bank: FP64 instructions read them as
// uniform-register / constant operands (LDCU.128 brings two per instruction), instead of the
// two 32-bit immediate moves per coefficient that literals cost.  Same values on the host.
#define RVL_K_TABLE                                                                              \
    {                                                                                            \
        0x1.45f306dc9c883p-1,          /*  0  2/pi                    */                         \
        6755399441055744.0,            /*  1  1.5*2^52 (rint trick)   */                         \
        0x1.921fb54442d18p+0,          /*  2  pi/2 high               */                         \
        0x1.1a62633145c07p-54,         /*  3  pi/2 middle             */                         \
        1.58969099521155010221e-10,    /*  4  S6                      */                         \
        -2.50507602534068634195e-08,   /*  5  S5                      */                         \
        2.75573137070700676789e-06,    /*  6  S4                      */                         \
        -1.98412698298579493134e-04,   /*  7  S3                      */                         \
        8.33333333332248946124e-03,    /*  8  S2                      */                         \
        -1.66666666666666324348e-01,   /*  9  S1                      */                         \
        -1.13596475577881948265e-11,   /* 10  C6                      */                         \
        2.08757232129817482790e-09,    /* 11  C5                      */                         \
        -2.75573143513906633035e-07,   /* 12  C4                      */                         \
        2.48015872894767294178e-05,    /* 13  C3                      */                         \
        -1.38888888888741095749e-03,   /* 14  C2                      */                         \
        4.16666666666666019037e-02,    /* 15  C1                      */                         \
        0.999999,                      /* 16  multiplier of the FP64 peak probe */                         \
        4.21626984126984127e-04,       /* 17  17/40320: d^7 of tan(d/2) */                       \
        4.16666666666666667e-03        /* 18  1/240:    d^5 of tan(d/2) */                       \
    }
// The functions below take the table as their first argument (`kt`): the likelihood kernel
// loads it ONCE per thread into (uniform) registers with loads the compiler may not
// rematerialise, so that the Newton loop holds no constant loads at all -- FP64 instructions of
// sm_100a take register or uniform-register operands only, never a constant-bank address, and
// ptxas otherwise re-loads each path's coefficients on every pass (18 loads per Kepler solve).
constexpr int kKTab = 19;
struct KTab {
    double v[kKTab];
    int vz;  // device: an opaque per-thread zero (see RVL_KV); host: unused
};
static const KTab h_ktab = {RVL_K_TABLE, 0};
#if defined(__CUDACC__)
__constant__ double d_ktab[kKTab] = RVL_K_TABLE;
__device__ __forceinline__ KTab load_ktab()
{
    KTab k;
#pragma unroll
    for (int i = 0; i < kKTab; ++i) k.v[i] = d_ktab[i];
    // zero in every lane (lanemask_lt < 2^31), but a per-lane value for the assembler
    asm volatile("mov.u32 %0, %%lanemask_lt;\n\tshr.u32 %0, %0, 31;" : "=r"(k.vz));
    return k;
}
// the same, pinned: the address carries an opaque warp-uniform zero (the SM id is far below 2^16),
// so the assembler cannot treat the loads as known constants to re-load wherever they are used; they
// become 17 uniform loads at the top of the kernel whose results stay in uniform registers
__device__ __forceinline__ KTab load_ktab_pinned()
{
    KTab k;
    unsigned uz;
    asm volatile("mov.u32 %0, %%smid;\n\tshr.u32 %0, %0, 16;" : "=r"(uz));
    const size_t base = __cvta_generic_to_constant(d_ktab) + (size_t)uz * 8u;
#pragma unroll
    for (int i = 0; i < kKTab; ++i)
        asm volatile("ld.const.f64 %0, [%1];" : "=d"(k.v[i]) : "l"(base + 8u * i));
    // zero in every lane (lanemask_lt < 2^31), but a per-lane value for the assembler
    asm volatile("mov.u32 %0, %%lanemask_lt;\n\tshr.u32 %0, %0, 31;" : "=r"(k.vz));
    return k;
}
#endif
#define RVL_K(i) (kt.v[i])
// A constant that STARTS a Horner chain (or meets a second constant in one instruction) has to sit
// in an ordinary register -- an FP64 instruction takes at most one uniform operand -- and is
// cheaper as one 64-bit constant load on the spot than as two moves out of uniform registers.
// (The index carries an opaque zero that lives in an ordinary register: the load then cannot be
// turned into a uniform load followed by those two moves.)
#if defined(__CUDA_ARCH__)
#define RVL_KV(i) \
    (*reinterpret_cast<const double *>(reinterpret_cast<const char *>(::rvl::d_ktab) + 8 * (i) + kt.vz))
#else
#define RVL_KV(i) (kt.v[i])
#endif

// ---- sin & cos of an un-reduced angle ----------------------------------------------------
// Cody-Waite reduction by pi/2 in two FMA steps (pi/2 to 106 bits: the neglected third term
// is < 1e-28 for |x| < 1e5), then the classic degree-13/14 minimax kernels on [-pi/4, pi/4]
// (fdlibm coefficients).  19 FP64 instructions; ~1 ulp.  Callers route |x| >= kTrigFastMax
// elsewhere (the first reduction step is exact only while |x| 2/pi < 2^17).
constexpr double kTrigFastMax = 100000.0;

RVL_HD void sincos_fast(const KTab &kt, double x, double &s, double &c)
{
    const double q = fma_(x, RVL_K(0), RVL_KV(1));
    const int32_t n = lo32(q);
    const double qf = sub(q, RVL_K(1));
    double r = fma_(-qf, RVL_K(2), x);
    r = fma_(-qf, RVL_K(3), r);
    const double z = mul(r, r);
    double ps = RVL_KV(4);
    ps = fma_(ps, z, RVL_K(5));
    ps = fma_(ps, z, RVL_K(6));
    ps = fma_(ps, z, RVL_K(7));
    ps = fma_(ps, z, RVL_K(8));
    ps = fma_(ps, z, RVL_K(9));
    double pc = RVL_KV(10);
    pc = fma_(pc, z, RVL_K(11));
    pc = fma_(pc, z, RVL_K(12));
    pc = fma_(pc, z, RVL_K(13));
    pc = fma_(pc, z, RVL_K(14));
    pc = fma_(pc, z, RVL_K(15));
    const double sr = fma_(mul(r, z), ps, r);
    const double cr = fma_(z, fma_(z, pc, -0.5), 1.0);
    // quadrant: n mod 4 = 0:(s,c) 1:(c,-s) 2:(-s,-c) 3:(-c,s)   [integer pipe only]
    const bool swp = (n & 1) != 0;
    double ss = swp ? cr : sr;
    double cc = swp ? sr : cr;
    const uint32_t sflip = ((uint32_t)n & 2u) << 30;         // bit 31 set when n&2
    const uint32_t cflip = (((uint32_t)n + 1u) & 2u) << 30;  // bit 31 set when (n+1)&2
    s = from_hilo((int32_t)((uint32_t)hi32(ss) ^ sflip), lo32(ss));
    c = from_hilo((int32_t)((uint32_t)hi32(cc) ^ cflip), lo32(cc));
}

// ---- advance (sin E, cos E) by a small step d: E <- E + d --------------------------------
// s' = s + (c sin d - s (1 - cos d)),  c' = c - (s sin d + c (1 - cos d)).
// Writing the update as "old value + small correction" keeps the added rounding error at half
// an ulp per step however many steps are chained.
// |d| <= 2^-10 : sin d = d - d^3/6 (next term 7e-18), 1-cos d = d^2/2 - d^4/24.   11 instr.
// The rotation (s, c) <- (s + (c sd - s v), c - (s sd + c v)) with sd = sin d, v = 1 - cos d.
// RVL_ROT_FUSED = 0: "old value + small correction" (mul, fma, add per component: 6 instructions,
// one half-ulp rounding of the result); 1: s' = fma(c, sd, fma(-s, v, s)) (4 instructions, two
// roundings of the size of the result).
#ifndef RVL_ROT_FUSED
#define RVL_ROT_FUSED 1
#endif
RVL_HD void rotate(double sd, double v, double &s, double &c)
{
#if RVL_ROT_FUSED
    const double s1 = fma_(c, sd, fma_(-s, v, s));
    const double c1 = fma_(-s, sd, fma_(-c, v, c));
    s = s1;
    c = c1;
#else
    const double ds = fma_(c, sd, -mul(s, v));
    const double dc = fma_(s, sd, mul(c, v));
    s = add(s, ds);
    c = sub(c, dc);
#endif
}

// The rotation with cos d itself, s' = fma(s, cos d, c sin d), c' = fma(c, cos d, -(s sin d)): the
// two products are formed first, then each component is updated IN PLACE -- 4 FP64 instructions
// like the form above, but no result has to wait in a third register for the other one's last
// read (one 64-bit register move per solve and pass on the GPU).  Roundings: cos d (5.5e-17), the
// product (below 1e-17 x 8 |sin d|), the result (5.5e-17): the same size as the form above.
#ifndef RVL_ROT_CD
#define RVL_ROT_CD 1
#endif
RVL_HD void rotate_cd(double sd, double cd, double &s, double &c)
{
    const double p = mul(c, sd);
    const double q = mul(s, sd);
    s = fma_(s, cd, p);
    c = fma_(c, cd, -q);
}

// The same rotation as THREE shears, in place: with t = tan(d/2),
//   [cos d  sin d; -sin d  cos d] = [1 t; 0 1] [1 0; -sin d 1] [1 t; 0 1]
// (1 - t sin d = cos d, t (1 + cos d) = sin d).  3 FP64 instructions instead of 4 and no register
// move (each step overwrites the operand it no longer needs; the 4-FMA form needs the old s for
// c' and the old c for s').  Every step is "old value + small correction": half an ulp of the
// result per step.  Used where tan(d/2) is as short a series as 1 - cos d (|d| <= 2^-5).
#ifndef RVL_ROT_SHEAR
#define RVL_ROT_SHEAR 1
#endif
RVL_HD void rotate_shear(double t, double sd, double &s, double &c)
{
    s = fma_(t, c, s);
    c = fma_(-sd, s, c);
    s = fma_(t, c, s);
}

// (All the short series below use the leading coefficients of the SAME minimax kernels as
// sincos_fast -- S1..S3, C1..C3 differ from -1/6, 1/120, .. by < 4e-16 relative, far below what
// the truncated terms leave -- so that every sin/cos path of the Newton loop draws on one set of
// 16 constants, which the compiler then keeps resident in uniform registers: no constant loads
// inside the loop.)
RVL_HD void advance_tiny(const KTab &kt, double d, double &s, double &c)
{
    const double d2 = mul(d, d);
#if RVL_ROT_SHEAR
    // tan(d/2) = d/2 + d^3/24 (next term d^5/240 < 4e-18); sin d = d - d^3/6 (next 8e-18)
    const double d3 = mul(d, d2);
    const double sd = fma_(d3, RVL_K(9), d);
    const double t = fma_(d3, RVL_K(15), mul(0.5, d));
    rotate_shear(t, sd, s, c);
#else
    const double sd = fma_(mul(d, d2), RVL_K(9), d);
    const double v = mul(d2, fma_(-d2, RVL_K(15), 0.5));
    rotate(sd, v, s, c);
#endif
}
// |d| <= 2^-5 : sin d through d^7 (next 8e-20), 1-cos d through d^8 (next 2e-22).   16 instr.
RVL_HD void advance_small(const KTab &kt, double d, double &s, double &c)
{
    const double d2 = mul(d, d);
    double ps = RVL_KV(7);
    ps = fma_(ps, d2, RVL_K(8));
    ps = fma_(ps, d2, RVL_K(9));
#if RVL_ROT_SHEAR
    // tan(d/2) = d/2 + d^3/24 + d^5/240 + 17 d^7/40320 (next term 31 d^9/725760 < 2e-18)
    const double d3 = mul(d, d2);
    const double sd = fma_(d3, ps, d);
    double pt = RVL_KV(17);
    pt = fma_(pt, d2, RVL_K(18));
    pt = fma_(pt, d2, RVL_K(15));
    const double t = fma_(d3, pt, mul(0.5, d));
    rotate_shear(t, sd, s, c);
#else
    const double sd = fma_(mul(d, d2), ps, d);
    double pc = RVL_KV(13);
    pc = fma_(pc, d2, RVL_K(14));
    pc = fma_(pc, d2, RVL_K(15));
    pc = fma_(pc, d2, -0.5);
    const double v = -mul(d2, pc);
    rotate(sd, v, s, c);
#endif
}
// |d| < 2^-3: sin d through d^9 (next term d^11/11! < 3e-18), 1-cos d through d^10 (next 3e-20);
// the leading minimax coefficients differ from the Taylor ones by less than 1e-18 at this range.
// 15 FP64 instructions.
RVL_HD void advance_mid(const KTab &kt, double d, double &s, double &c)
{
    const double z = mul(d, d);
    double ps = RVL_KV(6);
    ps = fma_(ps, z, RVL_K(7));
    ps = fma_(ps, z, RVL_K(8));
    ps = fma_(ps, z, RVL_K(9));
    double pc = RVL_KV(12);
    pc = fma_(pc, z, RVL_K(13));
    pc = fma_(pc, z, RVL_K(14));
    pc = fma_(pc, z, RVL_K(15));
    const double sd = fma_(mul(d, z), ps, d);
#if RVL_ROT_CD
    rotate_cd(sd, fma_(z, fma_(z, pc, -0.5), 1.0), s, c);
#else
    const double v = -mul(z, fma_(z, pc, -0.5));  // 1 - cos d
    rotate(sd, v, s, c);
#endif
}
// |d| < 0.75 (< pi/4): sin d and 1-cos d from the same minimax kernels as sincos_fast, but with
// no range reduction and no quadrant logic.  21 FP64 instructions, no integer work.
RVL_HD void advance_medium(const KTab &kt, double d, double &s, double &c)
{
    const double z = mul(d, d);
    double ps = RVL_KV(4);
    ps = fma_(ps, z, RVL_K(5));
    ps = fma_(ps, z, RVL_K(6));
    ps = fma_(ps, z, RVL_K(7));
    ps = fma_(ps, z, RVL_K(8));
    ps = fma_(ps, z, RVL_K(9));
    double pc = RVL_KV(10);
    pc = fma_(pc, z, RVL_K(11));
    pc = fma_(pc, z, RVL_K(12));
    pc = fma_(pc, z, RVL_K(13));
    pc = fma_(pc, z, RVL_K(14));
    pc = fma_(pc, z, RVL_K(15));
    const double sd = fma_(mul(d, z), ps, d);
#if RVL_ROT_CD
    rotate_cd(sd, fma_(z, fma_(z, pc, -0.5), 1.0), s, c);
#else
    const double v = -mul(z, fma_(z, pc, -0.5));  // 1 - cos d
    rotate(sd, v, s, c);
#endif
}
// the LAST pass of a solve: every |d| <= tol (1e-4 in the reference): sin d = d - d^3/6 (next term
// 8e-23), 1 - cos d = d^2/2 (next term d^4/24 = 4e-18, below half an ulp of the values it is
// subtracted from).  10 instructions.  Valid for |d| < 2e-4.
RVL_HD void advance_final(const KTab &kt, double d, double &s, double &c)
{
#if RVL_ROT_SHEAR
    advance_tiny(kt, d, s, c);  // (the tiny series holds up to 2^-10; same instruction count)
#else
    const double d2 = mul(d, d);
    const double sd = fma_(mul(d, d2), RVL_K(9), d);
    const double v = mul(0.5, d2);
    rotate(sd, v, s, c);
#endif
}
constexpr int kHiFinal = 0x3F2A36E2;  // high word of 2e-4: abs_hi(d) < this  =>  |d| < 2e-4
constexpr double kTinyStep = 0x1p-10;
constexpr double kSmallStep = 0x1p-5;

// ---- one Newton step of Kepler's equation (trueanomaly.c:25-29) ---------------------------
// Given E with (s, c) = (sin E, cos E):  f = (E - ec s) - M  [two grid roundings, as the
// reference], f' = 1 - ec c, E_new = E - f * rcp(f') [one grid rounding].  Returns E_new - E,
// which is exact (Sterbenz) and is the quantity the reference thresholds against tol (:21).
// RVL_FMA_F = 1: E - ec s is formed with ONE rounding (fma) instead of two (product, then
// difference).  The product's own rounding (<= 1.1e-16 absolute) is below the sin/cos error that
// is already accepted in s, and four orders below the grid of E the reference rounds on.
#ifndef RVL_FMA_F
#define RVL_FMA_F 1
#endif
RVL_HD double newton_step(double E, double s, double c, double M, double ec, double &Enew)
{
    const double r = rcp(fma_(-ec, c, 1.0));
#if RVL_FMA_F
    const double f = sub(fma_(-ec, s, E), M);
#else
    const double f = sub(sub(E, mul(ec, s)), M);
#endif
    Enew = fma_(-f, r, E);
    return sub(Enew, E);
}

// ---- radial velocity of one planet at eccentric anomaly E ---------------------------------
// K (cos(nu + w) + e cos w) with cos nu = (c - ec)/(1 - ec c), sin nu = sqrt(1-ec^2) s/(1 - ec c)
//   = (A (c - ec) + Bs s) / (1 - ec c) + Ce,
// A = K cos w, Bs = -K sin w sqrt(1 - ec^2), Ce = K (e cos w) with the UN-clamped e
// (rvmodel/__init__.py:463).  10 instructions.
RVL_HD double kepler_rv(double s, double c, double ec, double A, double Bs, double Ce)
{
    const double r = rcp(fma_(-ec, c, 1.0));
    const double num = fma_(Bs, s, mul(A, sub(c, ec)));
    return fma_(num, r, Ce);
}

// the same with mAec = -(A ec) prepared once per point: A (c - ec) + Bs s = fma(Bs, s, fma(A, c, mAec)),
// 9 instructions (one rounding fewer; rv only has to be right to a few ulp)
RVL_HD double kepler_rv2(double s, double c, double ec, double A, double Bs, double Ce, double mAec)
{
    const double r = rcp(fma_(-ec, c, 1.0));
    const double num = fma_(Bs, s, fma_(A, c, mAec));
    return fma_(num, r, Ce);
}

// |d| > tol on the bit patterns (positive doubles order like their bits): integer pipe instead of
// the FP64 pipe.  NaN counts as "not converged" (fabs(nan) > tol is false in the reference, where
// the loop then ends; here the warp-uniform exit treats NaN as large either way: see solve_planet).
RVL_HD bool abs_gt(double d, int32_t tol_hi, uint32_t tol_lo)
{
    const int32_t h = hi32(d) & 0x7fffffff;
    return h > tol_hi || (h == tol_hi && (uint32_t)lo32(d) > tol_lo);
}

// mean anomaly, exactly as the reference forms it (rvmodel/__init__.py:459): three roundings
RVL_HD double mean_anomaly(double nmot, double t, double epoch, double M0)
{
    return add(mul(nmot, sub(t, epoch)), M0);
}

// ---- log-determinant bookkeeping: split var into mantissa in [1,2) and exponent ----------
// sum_j ln(var_j) = ln(prod mant_j) + ln2 * sum exp_j.  Returns false for var that is zero,
// subnormal, negative, inf or nan (caller falls back to plain log()).
RVL_HD bool split_pos(double v, double &mant, int32_t &expo)
{
    const int32_t hi = hi32(v);
    const uint32_t be = (uint32_t)hi >> 20;  // sign+biased exponent
    expo = (int32_t)be - 1023;
    mant = from_hilo((hi & 0x000fffff) | 0x3ff00000, lo32(v));
    return (uint32_t)(be - 1u) < 0x7feu;
}
// The same split for the hot loop: `be` is the sign + BIASED exponent field; the caller sums the
// fields (removing 1023 per term at the end) and keeps the unsigned maximum of be - 1, which is
// below 0x7fe exactly when every term was a positive, normal, finite number.
RVL_HD void split_raw(double v, double &mant, uint32_t &be)
{
    const int32_t hi = hi32(v);
    be = (uint32_t)hi >> 20;
    mant = from_hilo((hi & 0x000fffff) | 0x3ff00000, lo32(v));
}
constexpr double kLn2Hi = 0x1.62e42fefa39efp-1;
constexpr double kLn2Lo = 0x1.abc9e3b39803fp-56;

// ---- ln(m) for m in [1, 2): the log-determinant needs ONE logarithm per work item -------------
// fdlibm's kernel: m in [sqrt(1/2), sqrt(2)) after an optional halving (reported in `half`, the
// caller adds it to the exponent sum), f = m - 1, s = f / (2 + f), ln m = f - (f^2/2 - s (f^2/2 + R))
// with R = minimax polynomial in s^2 (Lg1..Lg7, |error| < 2^-58.45).  ~25 FP64 instructions with
// no special cases (the argument is a finite mantissa by construction), against ~80 for the
// general library log().
RVL_HD double log_mantissa(double m, int32_t &half)
{
    half = (m > 0x1.6a09e667f3bcdp+0) ? 1 : 0;  // sqrt(2)
    m = half ? mul(m, 0.5) : m;
    const double f = sub(m, 1.0);
    const double s = mul(f, rcp(add(2.0, f)));
    const double z = mul(s, s);
    const double w = mul(z, z);
    const double t1 = mul(w, fma_(w, fma_(w, 1.531383769920937332e-01, 2.222219843214978396e-01),
                                  3.999999999940941908e-01));
    const double t2 = mul(z, fma_(w, fma_(w, fma_(w, 1.479819860511658591e-01, 1.818357216161805012e-01),
                                          2.857142874366239149e-01), 6.666666666666735130e-01));
    const double R = add(t2, t1);
    const double hfsq = mul(0.5, mul(f, f));
    return sub(f, sub(hfsq, mul(s, add(hfsq, R))));
}

// ---- exp(x), correctly rounded (double-double inside) ---------------------------------------
// Only for the once-per-point decode of the log parametrisations (logperiod, logk1:
// evidence/rvmodel/__init__.py:412-420).  The period is the one hypersensitive input of the path
// (SURVEY.md 0.5): one ulp of P moves lnL by up to ~1e-8, so exp() has to give THE nearest double,
// not a <= 1 ulp neighbour.  x = k ln2 + r (three-part ln2, first product exact), t = r / 256,
// expm1(t) by its Taylor series to t^10 in double-double, eight squarings p <- 2p + p^2, result
// (1 + p) 2^k.  Relative error ~2^-95: the rounding is correct except within 2^-42 of a tie.
// ~400 FP64 operations per call; a likelihood item is >= 10^4.
struct dd {
    double hi, lo;
};
RVL_HD dd dd_two_sum(double a, double b)
{
    const double s = add(a, b), bb = sub(s, a);
    return dd{s, add(sub(a, sub(s, bb)), sub(b, bb))};
}
RVL_HD dd dd_quick(double a, double b)  // |a| >= |b|
{
    const double s = add(a, b);
    return dd{s, sub(b, sub(s, a))};
}
RVL_HD dd dd_add(dd a, dd b)
{
    dd s = dd_two_sum(a.hi, b.hi);
    const dd t = dd_two_sum(a.lo, b.lo);
    s.lo = add(s.lo, t.hi);
    s = dd_quick(s.hi, s.lo);
    s.lo = add(s.lo, t.lo);
    return dd_quick(s.hi, s.lo);
}
RVL_HD dd dd_mul(dd a, dd b)
{
    const double p = mul(a.hi, b.hi);
    double e = fma_(a.hi, b.hi, -p);
    e = add(e, add(mul(a.hi, b.lo), mul(a.lo, b.hi)));
    return dd_quick(p, e);
}
RVL_HD double exp_cr(double x)
{
    if (!(x > -700.0 && x < 700.0)) return ::exp(x);  // overflow / underflow / nan: library
    const double kd = sub(fma_(x, 0x1.71547652b82fep+0, 6755399441055744.0), 6755399441055744.0);
    const int32_t k = (int32_t)kd;
    // r = x - k ln2: k * L1 is exact (L1 = 32 leading bits of ln2), and so is the difference
    const double r0 = sub(x, mul(kd, 0x1.62e42fee00000p-1));
    const double p2 = mul(kd, 0x1.a39ef35793c76p-33);
    const double p2e = fma_(kd, 0x1.a39ef35793c76p-33, -p2);
    dd r = dd_two_sum(r0, -p2);
    r.lo = sub(r.lo, add(p2e, mul(kd, 0x1.cc01f97b57a08p-87)));
    r = dd_quick(r.hi, r.lo);
    const dd t = dd{mul(r.hi, 0x1p-8), mul(r.lo, 0x1p-8)};
    // expm1(t) = t (1 + t (1/2! + t (1/3! + ... + t/11!)))  -> Horner on q = sum_{n>=2} t^(n-2)/n!
    const double fh[10] = {0x1.0000000000000p-1, 0x1.5555555555555p-3, 0x1.5555555555555p-5,
                           0x1.1111111111111p-7, 0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-13,
                           0x1.a01a01a01a01ap-16, 0x1.71de3a556c734p-19, 0x1.27e4fb7789f5cp-22,
                           0x1.ae64567f544e4p-26};
    const double fl[10] = {0.0, 0x1.5555555555555p-57, 0x1.5555555555555p-59, 0x1.1111111111111p-63,
                           -0x1.f49f49f49f49fp-65, 0x1.a01a01a01a01ap-73, 0x1.a01a01a01a01ap-76,
                           -0x1.c154f8ddc6c00p-73, 0x1.cbbc05b4fa99ap-76, -0x1.c062e06d1f209p-80};
    dd q = dd{fh[9], fl[9]};
    for (int n = 8; n >= 0; --n) q = dd_add(dd_mul(q, t), dd{fh[n], fl[n]});
    dd p = dd_add(t, dd_mul(dd_mul(t, t), q));  // t + t^2 q
    for (int i = 0; i < 8; ++i) {               // (1 + p)^2 - 1 = 2p + p^2
        const dd sq = dd_mul(p, p);
        p = dd_add(dd{mul(2.0, p.hi), mul(2.0, p.lo)}, sq);
    }
    const dd one_p = dd_add(dd{1.0, 0.0}, p);
    // scale by 2^k through the exponent field (k in [-1010, 1010]: split in two to stay normal)
    const int32_t k1 = k / 2, k2 = k - k1;
    const double s1 = from_hilo((k1 + 1023) << 20, 0), s2 = from_hilo((k2 + 1023) << 20, 0);
    return mul(mul(one_p.hi, s1), s2);
}

}  // namespace rvl
