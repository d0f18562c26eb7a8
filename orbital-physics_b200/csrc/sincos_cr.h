// sincos_cr.h -- correctly rounded sin / cos for the initial-condition pipeline (kepler.cu).
//
// The reference turns orbital elements into states with Python's math.sin / math.cos (core/body.py:184-249,
// core/physics.py:43-71), i.e. the host libm.  CUDA's sincos() is accurate to ~1-2 ulp, which made the device
// pipeline a tolerance path (42 % of the states bit-identical).  A libm is not bit-reproducible across hosts --
// glibc's sin / cos are not correctly rounded (0.14 % of random arguments differ from the exact rounding here) -- so
// the only host-independent target is CORRECT ROUNDING, which is what this routine delivers: it then agrees with
// any libm wherever that libm rounds correctly.
//
// Method: Cody-Waite reduction x = k (pi/2) + r with pi/2 in four parts (33 + 33 + 33 + 53 bits: k P_i is exact for
// |k| < 2^20), r kept as a double-double; sin r and cos r by their Taylor series (through r^29 / r^28, |r| <= pi/4:
// truncation < 2^-110) in double-double arithmetic with double-double coefficients; quadrant selection; the high word
// of the normalised result is the correctly rounded value unless the exact result lies within ~2^-100 relative of a
// rounding boundary (probability ~2^-45 per call; no such argument is known for |x| < 2^20 and none is in the tests).
// Every operation is a single IEEE add / multiply / fma, written through the SC_* macros so that the identical
// sequence compiles for the device (explicitly rounded intrinsics: no contraction) and for the host unit test
// (tests/native/sincos_host.c, gcc -ffp-contract=off), which holds it to mpmath on the CPU.
// Domain: |x| < 2^20 (orbital angles are a few radians); beyond that, and for inf / nan, the caller's fallback applies.
#pragma once

#ifdef __CUDACC__
#define SC_FN __device__ __forceinline__
#define SC_ADD(a, b) __dadd_rn((a), (b))
#define SC_SUB(a, b) __dsub_rn((a), (b))
#define SC_MUL(a, b) __dmul_rn((a), (b))
#define SC_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SC_RINT(a) rint(a)
#else
#include <math.h>
#define SC_FN static inline
#define SC_ADD(a, b) ((a) + (b))
#define SC_SUB(a, b) ((a) - (b))
#define SC_MUL(a, b) ((a) * (b))
#define SC_FMA(a, b, c) fma((a), (b), (c))
#define SC_RINT(a) rint(a)
#endif

typedef struct { double hi, lo; } sc_dd;

SC_FN sc_dd sc_two_sum(double a, double b) {
    sc_dd r;
    r.hi = SC_ADD(a, b);
    const double bb = SC_SUB(r.hi, a);
    r.lo = SC_ADD(SC_SUB(a, SC_SUB(r.hi, bb)), SC_SUB(b, bb));
    return r;
}

SC_FN sc_dd sc_fast_two_sum(double a, double b) {      /* |a| >= |b| */
    sc_dd r;
    r.hi = SC_ADD(a, b);
    r.lo = SC_SUB(b, SC_SUB(r.hi, a));
    return r;
}

SC_FN sc_dd sc_add(sc_dd x, sc_dd y) {                 /* accurate double-double sum (error < 3 u^2) */
    sc_dd s = sc_two_sum(x.hi, y.hi);
    const sc_dd t = sc_two_sum(x.lo, y.lo);
    s.lo = SC_ADD(s.lo, t.hi);
    s = sc_fast_two_sum(s.hi, s.lo);
    s.lo = SC_ADD(s.lo, t.lo);
    return sc_fast_two_sum(s.hi, s.lo);
}

SC_FN sc_dd sc_mul(sc_dd x, sc_dd y) {                 /* double-double product (error < 5 u^2) */
    sc_dd p;
    p.hi = SC_MUL(x.hi, y.hi);
    p.lo = SC_FMA(x.hi, y.hi, -p.hi);
    p.lo = SC_FMA(x.hi, y.lo, p.lo);
    p.lo = SC_FMA(x.lo, y.hi, p.lo);
    return sc_fast_two_sum(p.hi, p.lo);
}

/* (-1)^k / (2k+1)!  for k = 1..14  and  (-1)^k / (2k)!  for k = 1..14, as double-doubles */
#define SC_NCOEF 14
#ifdef __CUDACC__
__device__ __constant__
#else
static const
#endif
double sc_sin_c[SC_NCOEF][2] = {
        {-0x1.5555555555555p-3, -0x1.5555555555555p-57},
    {0x1.1111111111111p-7, 0x1.1111111111111p-63},
    {-0x1.a01a01a01a01ap-13, -0x1.a01a01a01a01ap-73},
    {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73},
    {-0x1.ae64567f544e4p-26, 0x1.c062e06d1f209p-80},
    {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87},
    {-0x1.ae7f3e733b81fp-41, -0x1.1d8656b0ee8cbp-97},
    {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
    {-0x1.2f49b46814157p-57, -0x1.2650f61dbdcb4p-112},
    {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120},
    {-0x1.761b41316381ap-75, 0x1.3423c7d91404fp-130},
    {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139},
    {-0x1.d1ab1c2dccea3p-94, -0x1.054d0c78aea14p-149},
    {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157}
};
#ifdef __CUDACC__
__device__ __constant__
#else
static const
#endif
double sc_cos_c[SC_NCOEF][2] = {
        {-0x1.0000000000000p-1, 0x0.0p+0},
    {0x1.5555555555555p-5, 0x1.5555555555555p-59},
    {-0x1.6c16c16c16c17p-10, 0x1.f49f49f49f49fp-65},
    {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76},
    {-0x1.27e4fb7789f5cp-22, -0x1.cbbc05b4fa99ap-76},
    {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83},
    {-0x1.93974a8c07c9dp-37, -0x1.05d6f8a2efd1fp-92},
    {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},
    {-0x1.6827863b97d97p-53, -0x1.eec01221a8b0bp-107},
    {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120},
    {-0x1.0ce396db7f853p-70, 0x1.aebcdbd20331cp-124},
    {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135},
    {-0x1.88e85fc6a4e5ap-89, 0x1.71c37ebd16540p-143},
    {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153}
};

/* sin and cos of the double-double r, |r| <= pi/4 (+ a little) */
SC_FN void sc_sincos_reduced(sc_dd r, sc_dd* s, sc_dd* c) {
    const sc_dd r2 = sc_mul(r, r);
    sc_dd ps, pc;
    ps.hi = sc_sin_c[SC_NCOEF - 1][0]; ps.lo = sc_sin_c[SC_NCOEF - 1][1];
    pc.hi = sc_cos_c[SC_NCOEF - 1][0]; pc.lo = sc_cos_c[SC_NCOEF - 1][1];
    for (int k = SC_NCOEF - 2; k >= 0; --k) {
        sc_dd cs, cc;
        cs.hi = sc_sin_c[k][0]; cs.lo = sc_sin_c[k][1];
        cc.hi = sc_cos_c[k][0]; cc.lo = sc_cos_c[k][1];
        ps = sc_add(sc_mul(ps, r2), cs);
        pc = sc_add(sc_mul(pc, r2), cc);
    }
    /* sin r = r + r (r^2 ps),  cos r = 1 + r^2 pc */
    const sc_dd one = {1.0, 0.0};
    *s = sc_add(r, sc_mul(r, sc_mul(r2, ps)));
    *c = sc_add(one, sc_mul(r2, pc));
}

/* returns 0 if x is outside the supported domain (|x| >= 2^20, inf, nan): the caller falls back */
SC_FN int sc_sincos(double x, double* sn, double* cs) {
    if (!(x > -1048576.0 && x < 1048576.0)) return 0;
    const double kf = SC_RINT(SC_MUL(x, 0x1.45f306dc9c883p-1));
    /* r = x - k pi/2: k 0x1.921fb54400000p+0..0x1.3198a2e000000p-69 are exact products; x - k 0x1.921fb54400000p+0 is exact (Sterbenz), the rest in double-double */
    const double t = SC_FMA(-kf, 0x1.921fb54400000p+0, x);
    sc_dd r = sc_two_sum(t, -SC_MUL(kf, 0x1.0b4611a600000p-34));
    const sc_dd r3 = sc_two_sum(r.hi, -SC_MUL(kf, 0x1.3198a2e000000p-69));
    r.lo = SC_ADD(SC_ADD(r.lo, r3.lo), -SC_MUL(kf, 0x1.b839a252049c1p-104));
    r = sc_fast_two_sum(r3.hi, r.lo);
    sc_dd s, c;
    sc_sincos_reduced(r, &s, &c);
    const int q = (int)((long long)kf & 3);
    const double sv = (q & 1) ? c.hi : s.hi;
    const double cv = (q & 1) ? s.hi : c.hi;
    *sn = (q & 2) ? -sv : sv;
    *cs = ((q + 1) & 2) ? -cv : cv;
    return 1;
}
