/* det_pow.h -- deterministic pow(x, y) for the FLINT-B200 arithmetic contract.
 *
 * Why this exists: the reference evaluates `e**r` with the
 * compiler's libm pow in the step-size controller (src/StepSz.f90:108,153,160).
 * CUDA's pow and glibc's pow do not return the same bits, and on the chaotic
 * problems of BASELINE.json a 1-ulp difference in one step size changes the
 * accepted/rejected counts of the rest of the trajectory.  Bit parity between
 * the CPU oracle and the sm_100a kernels therefore needs ONE pow whose result is
 * a pure function of IEEE-754 basic operations (+ - * / fma) and integer bit
 * manipulation, so that g++ (-ffp-contract=off) and nvcc (--fmad=false) produce
 * identical bits.  This is that function.
 *
 * Accuracy: log2(x) is carried as a double-double (~2^-66 relative), the
 * exponential is 1 + (t + ...) with a degree-15 Taylor tail; the final rounding
 * is the only O(ulp) error.  Measured against mpmath/libm in
 * tests/test_det_pow.py: max error < 1 ulp; the share of results that
 * are not the correctly rounded value is reported there.
 *
 * Domain: x >= 0 (the controller only raises non-negative error norms).  x < 0
 * returns NaN.  Follows C pow() for 0, inf, NaN, y == 0, x == 1.
 */
#ifndef FLINT_B200_DET_POW_H
#define FLINT_B200_DET_POW_H

#ifdef __CUDACC_RTC__              /* NVRTC (rtc.cu): no host headers; the device branches below use intrinsics only */
typedef unsigned long long uint64_t;
typedef long long int64_t;
typedef unsigned int uint32_t;
typedef int int32_t;
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#endif

#if defined(__CUDACC__)
#define FLINT_HD __host__ __device__ __forceinline__
#else
#define FLINT_HD static inline
#endif

FLINT_HD double flint_fma_(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

FLINT_HD uint64_t flint_d2u_(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}

FLINT_HD double flint_u2d_(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}

/* 2^n for any integer n, exact where representable (two-step for subnormals). */
FLINT_HD double flint_scale2_(double v, int n)
{
    if (n > 1023) { v *= 0x1p1023; n -= 1023; if (n > 1023) { v *= 0x1p1023; n -= 1023; if (n > 1023) n = 1023; } }
    else if (n < -1022) { v *= 0x1p-969; n += 969; if (n < -1022) { v *= 0x1p-969; n += 969; if (n < -1022) n = -1022; } }
    return v * flint_u2d_((uint64_t)(n + 1023) << 52);
}

FLINT_HD double flint_det_pow(double x, double y)
{
    /* ---- special cases (C99 pow semantics for the non-negative half-line) ---- */
    if (y == 0.0 || x == 1.0) return 1.0;
    if (x != x || y != y) return x + y;
    if (x < 0.0) return flint_u2d_(0x7ff8000000000000ULL);
    const double inf = flint_u2d_(0x7ff0000000000000ULL);
    if (x == 0.0) return (y > 0.0) ? 0.0 : inf;
    if (x == inf) return (y > 0.0) ? inf : 0.0;
    if (y == inf)  return (x > 1.0) ? inf : 0.0;
    if (y == -inf) return (x > 1.0) ? 0.0 : inf;

    /* ---- x = 2^e * m, m in [sqrt(1/2), sqrt(2)) ---- */
    int e = 0;
    uint64_t ix = flint_d2u_(x);
    if ((ix >> 52) == 0) { x *= 0x1p54; e = -54; ix = flint_d2u_(x); }   /* subnormal */
    e += (int)((ix >> 52) & 0x7ff) - 1023;
    double m = flint_u2d_((ix & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m >= 0x1.6a09e667f3bcdp+0) { m *= 0.5; e += 1; }

    /* ---- log(m) = 2 atanh(s), s = (m-1)/(m+1), as a double-double ---- */
    const double f = m - 1.0;                       /* exact */
    const double p_hi = m + 1.0;
    const double tt = p_hi - m;
    const double p_lo = (m - (p_hi - tt)) + (1.0 - tt);   /* TwoSum: m+1 = p_hi+p_lo */
    const double s_hi = f / p_hi;
    double r = flint_fma_(-s_hi, p_hi, f);
    r = flint_fma_(-s_hi, p_lo, r);
    const double s_lo = r / p_hi;
    const double s2 = s_hi * s_hi;
    /* P(s2) = 2/3 + 2/5 s2 + ... + 2/25 s2^11 */
    double P = 2.0 / 25.0;
    P = flint_fma_(P, s2, 2.0 / 23.0);
    P = flint_fma_(P, s2, 2.0 / 21.0);
    P = flint_fma_(P, s2, 2.0 / 19.0);
    P = flint_fma_(P, s2, 2.0 / 17.0);
    P = flint_fma_(P, s2, 2.0 / 15.0);
    P = flint_fma_(P, s2, 2.0 / 13.0);
    P = flint_fma_(P, s2, 2.0 / 11.0);
    P = flint_fma_(P, s2, 2.0 / 9.0);
    P = flint_fma_(P, s2, 2.0 / 7.0);
    P = flint_fma_(P, s2, 2.0 / 5.0);
    P = flint_fma_(P, s2, 2.0 / 3.0);
    const double a_hi = 2.0 * s_hi;
    const double cor = flint_fma_(s_hi * s2, P, 2.0 * s_lo);
    const double L_hi = a_hi + cor;
    const double L_lo = cor - (L_hi - a_hi);        /* Fast2Sum, |a_hi| >= |cor| */

    /* ---- log2(x) = e + log(m)/ln2 (double-double) ---- */
    const double INVLN2_HI = 0x1.71547652b82fep+0;  /* 1/ln2 rounded      */
    const double INVLN2_LO = 0x1.777d0ffda0d24p-56; /* 1/ln2 - INVLN2_HI  */
    const double q_hi = L_hi * INVLN2_HI;
    const double q_lo = flint_fma_(L_hi, INVLN2_HI, -q_hi) + flint_fma_(L_hi, INVLN2_LO, L_lo * INVLN2_HI);
    const double ed = (double)e;
    const double z_hi = ed + q_hi;
    const double z_lo = ((ed - z_hi) + q_hi) + q_lo; /* Fast2Sum: e == 0 or |e| >= 1 > |q_hi| */

    /* ---- w = y * log2(x) ---- */
    const double w_hi = y * z_hi;
    const double w_lo = flint_fma_(y, z_hi, -w_hi) + y * z_lo;
    if (w_hi >= 1025.0) return inf;
    if (w_hi <= -1080.0) return 0.0;

    /* ---- 2^w = 2^n * exp(r ln2), |r| <= 1/2 ---- */
    const double nd = (double)(long long)(w_hi + (w_hi >= 0.0 ? 0.5 : -0.5));
    const double rr = (w_hi - nd) + w_lo;
    const double LN2_HI = 0x1.62e42fefa39efp-1;
    const double LN2_LO = 0x1.abc9e3b39803fp-56;
    const double t_hi = rr * LN2_HI;
    const double t_lo = flint_fma_(rr, LN2_HI, -t_hi) + rr * LN2_LO;
    /* Q(t) = 1/2! + t/3! + ... + t^13/15! */
    double Q = 1.0 / 1307674368000.0;
    Q = flint_fma_(Q, t_hi, 1.0 / 87178291200.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 6227020800.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 479001600.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 39916800.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 3628800.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 362880.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 40320.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 5040.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 720.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 120.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 24.0);
    Q = flint_fma_(Q, t_hi, 1.0 / 6.0);
    Q = flint_fma_(Q, t_hi, 0.5);
    const double tail = flint_fma_(t_hi * t_hi, Q, t_lo);
    const double v_hi = 1.0 + t_hi;
    const double v_lo = t_hi - (v_hi - 1.0);        /* Fast2Sum, 1 >= |t_hi| */
    const double v = v_hi + (v_lo + tail);
    return flint_scale2_(v, (int)nd);
}

#endif /* FLINT_B200_DET_POW_H */
