// sincos_libm.h -- sin / cos with the bits of the host libm the reference runs on (kepler.cu, trig mode "libm").
//
// The reference turns orbital elements into states with Python's math.sin / math.cos (core/body.py:184-249,
// core/physics.py:43-71), i.e. the C library of the host.  That dependency is not part of /root/reference: on this
// image it is GNU libc 2.39 (Ubuntu GLIBC 2.39-0ubuntu8.5), whose double sin/cos are the IBM Accurate Mathematical
// Library routines (sysdeps/ieee754/dbl-64/s_sin.c, < 0.55 ulp, NOT correctly rounded), in the build the x86-64 ifunc
// selects on every CPU with FMA + AVX2 (`__sin_fma` / `__cos_fma`: the same source compiled with -mfma, i.e. with
// every multiply that feeds an add contracted).  This header restates that published algorithm -- same branch
// points, polynomials, table look-up and contraction pattern -- as a sequence of single IEEE operations, so the
// device produces the bits Python produces on the host:
//   |x| < 2^-26 (sin) / 2^-27 (cos)        x / 1
//   |x| < 0.855469                          table step 1/128 + degree-5/6 corrections (|x| < 0.126: Taylor to x^11)
//   |x| < 2.426265                          via pi/2 - |x| (pi/2 in two doubles)
//   |x| < 105414350                         Cody-Waite reduction, pi/2 in four parts, result as hi + lo
//   beyond, inf, nan                        returns 0: the caller falls back (tolerance path)
// Pinned by tests/test_sincos.py: the host build of this very sequence (gcc -ffp-contract=off, fma() explicit)
// equals math.sin / math.cos bit for bit on millions of arguments per branch; the device build equals the host
// build (tests/test_device.py).  On a host whose libm differs (no FMA, another libc) the routine still returns
// glibc-FMA bits: < 0.55 ulp, and the test against that host's libm skips.
// The table (sincos_libm_tab.h) holds sin/cos of k/128 as double-doubles: mathematical constants, regenerated from
// mpmath by tools/gen_sincos_tab.py.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SL_FN __device__ __forceinline__
#define SL_TAB_DECL static __device__ const
#define SL_ADD(a, b) __dadd_rn((a), (b))
#define SL_SUB(a, b) __dsub_rn((a), (b))
#define SL_MUL(a, b) __dmul_rn((a), (b))
#define SL_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SL_BITS(x) ((uint64_t)__double_as_longlong(x))
#define SL_FROM(u) __longlong_as_double((long long)(u))
#else
#include <math.h>
#include <string.h>
#define SL_FN static inline
#define SL_TAB_DECL static const
#define SL_ADD(a, b) ((a) + (b))
#define SL_SUB(a, b) ((a) - (b))
#define SL_MUL(a, b) ((a) * (b))
#define SL_FMA(a, b, c) fma((a), (b), (c))
static inline uint64_t sl_bits_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double sl_from_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define SL_BITS(x) sl_bits_(x)
#define SL_FROM(u) sl_from_(u)
#endif
#include "sincos_libm_tab.h"

/* Taylor coefficients for |x| < 0.126 and the short polynomials around a table point (s_sin.c / usncs.h) */
#define SL_S1 (-0x1.5555555555555p-3)
#define SL_S2 (0x1.1111111110ecep-7)
#define SL_S3 (-0x1.a01a019db08b8p-13)
#define SL_S4 (0x1.71de27b9a7ed9p-19)
#define SL_S5 (-0x1.addffc2fcdf59p-26)
#define SL_SN3 (-1.66666666666664880952546298448555E-01)
#define SL_SN5 (8.33333214285722277379541354343671E-03)
#define SL_CS2 (4.99999999999999999999950396842453E-01)
#define SL_CS4 (-4.16666666666664434524222570944589E-02)
#define SL_CS6 (1.38888874007937613028114285595617E-03)
#define SL_BIG (0x1.8p45)                    /* ulp 2^-7: big + |x| puts round(128 |x|) in the low word */
#define SL_TOINT (0x1.8p52)
#define SL_HPINV (0x1.45f306dc9c883p-1)      /* 2/pi */
#define SL_HP0 (0x1.921fb54442d18p+0)        /* pi/2 = HP0 + HP1 */
#define SL_HP1 (0x1.1a62633145c07p-54)
#define SL_MP1 (0x1.921fb58000000p+0)        /* pi/2 = MP1 + MP2 + PP3 + PP4, the first three short: k * part is exact */
#define SL_MP2 (-0x1.dde973c000000p-27)
#define SL_PP3 (-0x1.cb3b398000000p-55)
#define SL_PP4 (-0x1.d747f23e32ed7p-83)

SL_FN double sl_abs(double x) { return SL_FROM(SL_BITS(x) & 0x7fffffffffffffffull); }
SL_FN double sl_copysign(double m, double s) {
    return SL_FROM((SL_BITS(m) & 0x7fffffffffffffffull) | (SL_BITS(s) & 0x8000000000000000ull));
}

/* sin(x + dx), |x| < 0.855469 */
SL_FN double sl_do_sin(double x, double dx) {
    const double xold = x;
    const double ax = sl_abs(x);
    if (ax < 0.126) {
        const double xx = SL_MUL(x, x);
        double p = SL_FMA(SL_S5, xx, SL_S4);
        p = SL_FMA(p, xx, SL_S3);
        p = SL_FMA(p, xx, SL_S2);
        p = SL_FMA(p, xx, SL_S1);
        double t = SL_FMA(p, x, -SL_MUL(0.5, dx));
        t = SL_FMA(t, xx, dx);
        return SL_ADD(x, t);
    }
    if (x <= 0) dx = -dx;
    const double u = SL_ADD(SL_BIG, ax);
    x = SL_SUB(ax, SL_SUB(u, SL_BIG));
    const int k = (int)(uint32_t)SL_BITS(u);
    const double xx = SL_MUL(x, x);
    const double s = SL_ADD(x, SL_FMA(SL_MUL(x, xx), SL_FMA(xx, SL_SN5, SL_SN3), dx));
    const double c = SL_FMA(x, dx, SL_MUL(xx, SL_FMA(xx, SL_FMA(xx, SL_CS6, SL_CS4), SL_CS2)));
    const double sn = sl_tab[k][0], ssn = sl_tab[k][1], cs = sl_tab[k][2], ccs = sl_tab[k][3];
    const double cor = SL_FMA(cs, s, SL_FMA(-sn, c, SL_FMA(s, ccs, ssn)));
    return sl_copysign(SL_ADD(sn, cor), xold);
}

/* cos(x + dx), |x| < 0.855469 */
SL_FN double sl_do_cos(double x, double dx) {
    if (x < 0) dx = -dx;
    const double ax = sl_abs(x);
    const double u = SL_ADD(SL_BIG, ax);
    x = SL_ADD(SL_SUB(ax, SL_SUB(u, SL_BIG)), dx);
    const int k = (int)(uint32_t)SL_BITS(u);
    const double xx = SL_MUL(x, x);
    const double s = SL_FMA(SL_MUL(x, xx), SL_FMA(xx, SL_SN5, SL_SN3), x);
    const double c = SL_MUL(xx, SL_FMA(xx, SL_FMA(xx, SL_CS6, SL_CS4), SL_CS2));
    const double sn = sl_tab[k][0], ssn = sl_tab[k][1], cs = sl_tab[k][2], ccs = sl_tab[k][3];
    const double cor = SL_FMA(-sn, s, SL_FMA(-cs, c, SL_FMA(-s, ssn, ccs)));
    return SL_ADD(cs, cor);
}

/* x = n pi/2 + (a + da), 2.426265 <= |x| < 105414350; returns n mod 4 */
SL_FN int sl_reduce(double x, double* a, double* da) {
    const double t = SL_FMA(x, SL_HPINV, SL_TOINT);
    const double xn = SL_SUB(t, SL_TOINT);
    const double y = SL_FMA(-xn, SL_MP2, SL_FMA(-xn, SL_MP1, x));
    const int n = (int)(uint32_t)SL_BITS(t) & 3;
    const double t2 = SL_FMA(-xn, SL_PP3, y);
    double db = SL_FMA(-xn, SL_PP3, SL_SUB(y, t2));
    const double b = SL_FMA(-xn, SL_PP4, t2);
    db = SL_ADD(db, SL_FMA(-xn, SL_PP4, SL_SUB(t2, b)));
    *a = b;
    *da = db;
    return n;
}

SL_FN double sl_quadrant(double a, double da, int n) {
    const double r = (n & 1) ? sl_do_cos(a, da) : sl_do_sin(a, da);
    return (n & 2) ? -r : r;
}

/* returns 0 outside the restated domain (|x| >= 105414350, inf, nan): the caller falls back */
SL_FN int sl_sincos(double x, double* sn, double* cs) {
    const uint32_t k = (uint32_t)(SL_BITS(x) >> 32) & 0x7fffffffu;
    if (k >= 0x419921FBu) return 0;
    if (k < 0x3feb6000u) {
        *sn = k < 0x3e500000u ? x : sl_do_sin(x, 0.0);
        *cs = k < 0x3e400000u ? 1.0 : sl_do_cos(x, 0.0);
    } else if (k < 0x400368fdu) {
        const double y = SL_SUB(SL_HP0, sl_abs(x));
        *sn = sl_copysign(sl_do_cos(y, SL_HP1), x);
        const double a = SL_ADD(y, SL_HP1);
        *cs = sl_do_sin(a, SL_ADD(SL_SUB(y, a), SL_HP1));
    } else {
        double a, da;
        const int n = sl_reduce(x, &a, &da);
        *sn = sl_quadrant(a, da, n);
        *cs = sl_quadrant(a, da, n + 1);
    }
    return 1;
}
