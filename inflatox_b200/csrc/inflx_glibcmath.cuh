// inflx_glibcmath.cuh - the libm the REFERENCE runs, restated for the device: bit-identical pow / exp /
// log (and, further down, sin / cos / tanh / atan / tan) to glibc 2.39's x86_64 FMA builds.
//
// Why: the reference's arithmetic is "C compiler + platform libm" (reference
// python/inflatox/compiler.py:299-310 links the generated model with -lm; SURVEY.md 8c "third-party
// arithmetic").  A correctly rounded libm (inflx_crmath.cuh) is *more accurate* than glibc's but is
// not *the reference's result*: glibc's pow is misrounded in ~1e-3 of its calls, and the ill-
// conditioned test models amplify a last-bit difference of one row-level pow into > 1e-10 on a whole
// grid row.  The functions below follow the algorithm glibc 2.39 runs on an x86_64 host with FMA
// (ifunc variants __pow_fma / __exp_fma / __log_fma of sysdeps/ieee754/dbl-64/e_pow.c, e_exp.c,
// e_log.c - Szabolcs Nagy's routines from Arm's optimized-routines) operation by operation,
// INCLUDING which a*b+c the compiler contracted into one FMA in that build (read off the shipped
// binary: the source leaves contraction to the compiler).  glibc is not part of /root/reference; it
// is the un-vendored dependency of the path, pinned here as Ubuntu GLIBC 2.39-0ubuntu8.5 (the libm
// of this image, which is also the libm the oracle links on the GPU box).
//
// tests/test_glibcmath.py compiles this file for the host and demands bit identity with the host's
// libm on >= 10^7 random arguments per function in the CPU suite (10^8+ with INFLX_GLIBC_CHECK_N,
// results in profiles/glibcmath_r2.txt); tests/test_gpu_numerics.py demands device == host build.
//
// Cost: pow ~75, exp ~25, log ~30 FP64 instructions + 2-3 dependent table loads: affordable per
// parameter vector / grid row / grid column (node classes P, R, C), which is where the generator
// uses them (cudagen.LIBM_FUNCTIONS).
//
// Only results are reproduced, not errno / floating-point exception flags (the device has neither).
#pragma once

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define INFLX_GL_FN __device__ __noinline__
#define INFLX_GL_INL __device__ __forceinline__
#define INFLX_GL_TABLE static __device__ const
#define INFLX_GL_ADD(a, b) __dadd_rn((a), (b))
#define INFLX_GL_SUB(a, b) __dsub_rn((a), (b))
#define INFLX_GL_MUL(a, b) __dmul_rn((a), (b))
#define INFLX_GL_FMA(a, b, c) __fma_rn((a), (b), (c))
#define INFLX_GL_DIV(a, b) __ddiv_rn((a), (b))
#define INFLX_GL_BITS(x) ((unsigned long long)__double_as_longlong(x))
#define INFLX_GL_FROM_BITS(b) __longlong_as_double((long long)(b))
#else  // host build for the tests: compile with -ffp-contract=off
#include <math.h>
#include <string.h>
#define INFLX_GL_FN static
#define INFLX_GL_INL static inline
#define INFLX_GL_TABLE static const
#define INFLX_GL_ADD(a, b) ((a) + (b))
#define INFLX_GL_SUB(a, b) ((a) - (b))
#define INFLX_GL_MUL(a, b) ((a) * (b))
#define INFLX_GL_FMA(a, b, c) fma((a), (b), (c))
#define INFLX_GL_DIV(a, b) ((a) / (b))
static inline unsigned long long INFLX_GL_BITS(double x) { unsigned long long b; memcpy(&b, &x, 8); return b; }
static inline double INFLX_GL_FROM_BITS(unsigned long long b) { double x; memcpy(&x, &b, 8); return x; }
#endif

#ifndef INFLX_GLIBC_TABLES_INCLUDED
#define INFLX_GLIBC_TABLES_INCLUDED
#include "inflx_glibc_tables.cuh"
#endif

typedef unsigned long long inflx_gl_u64;
typedef long long inflx_gl_i64;

#define INFLX_GL_NAN INFLX_GL_FROM_BITS(0xfff8000000000000ull) /* x86's default NaN: (x-x)/(x-x) */
#define INFLX_GL_INF INFLX_GL_FROM_BITS(0x7ff0000000000000ull)

// ---- exp ------------------------------------------------------------------------------------------
// e_exp.c specialcase(): 2^(k/N) * (1 + tmp) when the scale factor itself is not representable.
INFLX_GL_INL double inflx_gl_exp_special(double tmp, inflx_gl_u64 sbits, inflx_gl_u64 ki, int pow_abs) {
  if ((ki & 0x80000000ull) == 0) {  // k > 0: the exponent of scale might have overflowed
    sbits -= 1009ull << 52;
    const double scale = INFLX_GL_FROM_BITS(sbits);
    return INFLX_GL_MUL(0x1p1009, INFLX_GL_FMA(scale, tmp, scale));
  }
  sbits += 1022ull << 52;  // k < 0: care in the subnormal range
  const double scale = INFLX_GL_FROM_BITS(sbits);
  const double st = INFLX_GL_MUL(scale, tmp);
  double y = INFLX_GL_ADD(scale, st);
  if ((pow_abs ? fabs(y) : y) < 1.0) {
    // round to the right precision before scaling into the subnormal range (no double rounding)
    double one = 1.0;
    if (pow_abs && y < 0.0) one = -1.0;
    double lo = INFLX_GL_ADD(INFLX_GL_SUB(scale, y), st);
    const double hi = INFLX_GL_ADD(one, y);
    lo = INFLX_GL_ADD(INFLX_GL_ADD(INFLX_GL_SUB(one, hi), y), lo);
    y = INFLX_GL_SUB(INFLX_GL_ADD(lo, hi), one);
    if (y == 0.0) y = pow_abs ? INFLX_GL_FROM_BITS(sbits & 0x8000000000000000ull) : 0.0;
  }
  return INFLX_GL_MUL(0x1p-1022, y);
}

// The shared kernel of exp(x) and pow's exp_inline(x, xtail, sign_bias): exp(x + xtail), the result
// negated through the scale when sign_bias != 0.  `abstop` is already range-checked by the caller.
INFLX_GL_INL double inflx_gl_exp_core(double x, double xtail, int with_tail, inflx_gl_u64 sign_bias,
                                      int special) {
  // x = ln2/N*k + r;  kd = round(x N/ln2) through the 1.5*2^52 shift, contracted with the product
  const double zs = INFLX_GL_FMA(x, INFLX_GL_INVLN2N, INFLX_GL_SHIFT);
  const inflx_gl_u64 ki = INFLX_GL_BITS(zs);
  const double kd = INFLX_GL_SUB(zs, INFLX_GL_SHIFT);
  double r = INFLX_GL_FMA(kd, INFLX_GL_NEGLN2HIN, x);
  r = INFLX_GL_FMA(kd, INFLX_GL_NEGLN2LON, r);
  if (with_tail) r = INFLX_GL_ADD(xtail, r);
  const inflx_gl_u64 idx = 2 * (ki & 127);
  const inflx_gl_u64 top = (ki + sign_bias) << 45;
  const double tail = INFLX_GL_FROM_BITS(inflx_gl_exp_tab[idx]);
  const inflx_gl_u64 sbits = inflx_gl_exp_tab[idx + 1] + top;
  const double r2 = INFLX_GL_MUL(r, r);
  const double p23 = INFLX_GL_FMA(INFLX_GL_EXP_C3, r, INFLX_GL_EXP_C2);
  const double p45 = INFLX_GL_FMA(r, INFLX_GL_EXP_C5, INFLX_GL_EXP_C4);
  double tmp = INFLX_GL_FMA(p23, r2, INFLX_GL_ADD(r, tail));
  tmp = INFLX_GL_FMA(INFLX_GL_MUL(r2, r2), p45, tmp);
  if (special) return inflx_gl_exp_special(tmp, sbits, ki, with_tail);
  const double scale = INFLX_GL_FROM_BITS(sbits);
  return INFLX_GL_FMA(scale, tmp, scale);
}

INFLX_GL_FN double inflx_gl_exp(double x) {
  const inflx_gl_u64 ix = INFLX_GL_BITS(x);
  const unsigned abstop = (unsigned)(ix >> 52) & 0x7ff;
  int special = 0;
  if (abstop - 0x3c9u >= 0x3fu) {           // |x| < 2^-54, |x| >= 512, inf or NaN
    if ((int)(abstop - 0x3c9u) < 0) return INFLX_GL_ADD(1.0, x);
    if (abstop >= 0x409u) {                  // |x| >= 1024
      if (ix == 0xfff0000000000000ull) return 0.0;
      if (abstop >= 0x7ffu) return INFLX_GL_ADD(1.0, x);
      return (ix >> 63) ? 0.0 : INFLX_GL_INF;
    }
    special = 1;
  }
  return inflx_gl_exp_core(x, 0.0, 0, 0, special);
}

// ---- log ------------------------------------------------------------------------------------------
INFLX_GL_FN double inflx_gl_log(double x) {
  inflx_gl_u64 ix = INFLX_GL_BITS(x);
  const unsigned top = (unsigned)(ix >> 48);
  if (ix - 0x3fee000000000000ull < 0x3090000000000ull) {  // 1 - 2^-4 <= x < 1 + 0x1.09p-4
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double* B = inflx_gl_log_poly1;
    const double r = INFLX_GL_SUB(x, 1.0);
    const double r2 = INFLX_GL_MUL(r, r);
    const double r3 = INFLX_GL_MUL(r, r2);
    double q1 = INFLX_GL_FMA(B[2], r, B[1]);
    double q4 = INFLX_GL_FMA(B[5], r, B[4]);
    double q7 = INFLX_GL_FMA(B[8], r, B[7]);
    q1 = INFLX_GL_FMA(B[3], r2, q1);
    q4 = INFLX_GL_FMA(B[6], r2, q4);
    q7 = INFLX_GL_FMA(r2, B[9], q7);
    q7 = INFLX_GL_FMA(B[10], r3, q7);
    q7 = INFLX_GL_FMA(q7, r3, q4);
    q7 = INFLX_GL_FMA(q7, r3, q1);           // y = r3 * q7, folded into the sum below
    const double rw = INFLX_GL_FMA(r, 0x1p27, r);  // r + w, w = r * 2^27
    const double rhi = INFLX_GL_FMA(-0x1p27, r, rw);
    const double rhi2 = INFLX_GL_MUL(rhi, rhi);
    const double rlo = INFLX_GL_SUB(r, rhi);
    const double hi = INFLX_GL_FMA(rhi2, B[0], r);  // B[0] == -0.5
    double lo = INFLX_GL_FMA(rhi2, B[0], INFLX_GL_SUB(r, hi));
    lo = INFLX_GL_FMA(INFLX_GL_MUL(B[0], rlo), INFLX_GL_ADD(r, rhi), lo);
    return INFLX_GL_ADD(hi, INFLX_GL_FMA(q7, r3, lo));
  }
  if (top - 0x0010u >= 0x7ff0u - 0x0010u) {  // x < 2^-1022, inf or NaN
    if (ix * 2 == 0) return -INFLX_GL_INF;
    if (ix == 0x7ff0000000000000ull) return x;
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u)
      return (ix << 1) > 0xffe0000000000000ull ? INFLX_GL_ADD(x, x) : INFLX_GL_NAN;
    ix = INFLX_GL_BITS(INFLX_GL_MUL(x, 0x1p52));  // subnormal: normalise
    ix -= 52ull << 52;
  }
  const inflx_gl_u64 tmp = ix - 0x3fe6000000000000ull;
  const int i = (int)((tmp >> 45) & 127);
  const int k = (int)((inflx_gl_i64)tmp >> 52);
  const double z = INFLX_GL_FROM_BITS(ix - (tmp & (0xfffull << 52)));
  const double invc = inflx_gl_log_tab[2 * i], logc = inflx_gl_log_tab[2 * i + 1];
  const double* A = inflx_gl_log_poly;
  const double kd = (double)k;
  const double w = INFLX_GL_FMA(INFLX_GL_LN2HI, kd, logc);
  const double r = INFLX_GL_FMA(z, invc, -1.0);
  const double p12 = INFLX_GL_FMA(A[2], r, A[1]);
  const double hi = INFLX_GL_ADD(r, w);
  const double r2 = INFLX_GL_MUL(r, r);
  double lo = INFLX_GL_ADD(INFLX_GL_SUB(w, hi), r);
  lo = INFLX_GL_FMA(INFLX_GL_LN2LO, kd, lo);
  const double r3 = INFLX_GL_MUL(r, r2);
  double p = INFLX_GL_FMA(r, A[4], A[3]);
  lo = INFLX_GL_FMA(A[0], r2, lo);
  p = INFLX_GL_FMA(p, r2, p12);
  return INFLX_GL_ADD(INFLX_GL_FMA(r3, p, lo), hi);
}

// ---- pow ------------------------------------------------------------------------------------------
// 0: y is not an integer, 1: odd integer, 2: even integer
INFLX_GL_INL int inflx_gl_checkint(inflx_gl_u64 iy) {
  const int e = (int)(iy >> 52) & 0x7ff;
  if (e < 0x3ff) return 0;
  if (e > 0x3ff + 52) return 2;
  if (iy & ((1ull << (0x3ff + 52 - e)) - 1)) return 0;
  if (iy & (1ull << (0x3ff + 52 - e))) return 1;
  return 2;
}

// log_inline(ix) = hi + lo and its product with y as ehi + elo: the middle of pow, shared by the
// out-of-line routine and the inlined per-point fast path.
INFLX_GL_INL void inflx_gl_pow_log_mul(inflx_gl_u64 ix, double y, double* ehi_out, double* elo_out) {
  // ---- log_inline: log(x) = k ln2 + log(c) + log1p(z/c - 1) as hi + lo ----
  const inflx_gl_u64 tmp = ix - 0x3fe6955500000000ull;
  const int i = (int)((tmp >> 45) & 127);
  const int k = (int)((inflx_gl_i64)tmp >> 52);
  const double z = INFLX_GL_FROM_BITS(ix - (tmp & (0xfffull << 52)));
  const double kd = (double)k;
  const double invc = inflx_gl_pow_log_tab[3 * i], logc = inflx_gl_pow_log_tab[3 * i + 1];
  const double logctail = inflx_gl_pow_log_tab[3 * i + 2];
  const double* A = inflx_gl_pow_poly;
  const double t1 = INFLX_GL_FMA(kd, INFLX_GL_LN2HI, logc);
  const double lo1 = INFLX_GL_FMA(kd, INFLX_GL_LN2LO, logctail);
  const double r = INFLX_GL_FMA(z, invc, -1.0);
  const double ar = INFLX_GL_MUL(r, A[0]);
  const double p12 = INFLX_GL_FMA(A[2], r, A[1]);
  const double p34 = INFLX_GL_FMA(A[4], r, A[3]);
  const double t2 = INFLX_GL_ADD(r, t1);
  const double lo2 = INFLX_GL_ADD(INFLX_GL_SUB(t1, t2), r);
  const double ar2 = INFLX_GL_MUL(r, ar);
  const double ar3 = INFLX_GL_MUL(r, ar2);
  const double lo3 = INFLX_GL_FMA(ar, r, -ar2);
  const double lhi = INFLX_GL_ADD(t2, ar2);
  double p = INFLX_GL_FMA(r, A[6], A[5]);
  p = INFLX_GL_FMA(p, ar2, p34);
  const double lo4 = INFLX_GL_ADD(INFLX_GL_SUB(t2, lhi), ar2);
  p = INFLX_GL_FMA(ar2, p, p12);
  double llo = INFLX_GL_ADD(lo1, lo2);
  llo = INFLX_GL_ADD(llo, lo3);
  llo = INFLX_GL_ADD(llo, lo4);
  llo = INFLX_GL_FMA(ar3, p, llo);
  const double hi = INFLX_GL_ADD(lhi, llo);
  const double lo = INFLX_GL_ADD(INFLX_GL_SUB(lhi, hi), llo);
  // ---- y * log(x) as ehi + elo ----
  *ehi_out = INFLX_GL_MUL(y, hi);
  *elo_out = INFLX_GL_FMA(y, lo, INFLX_GL_FMA(hi, y, -*ehi_out));
}

INFLX_GL_FN double inflx_gl_pow(double x, double y) {
  inflx_gl_u64 sign_bias = 0;
  inflx_gl_u64 ix = INFLX_GL_BITS(x);
  const inflx_gl_u64 iy = INFLX_GL_BITS(y);
  unsigned topx = (unsigned)(ix >> 52);
  const unsigned topy = (unsigned)(iy >> 52);
  if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ffu) - 0x3beu >= 0x43eu - 0x3beu) {
    // x is subnormal, zero, negative, inf or NaN, or |y| is huge, tiny, inf or NaN
    if (2 * iy - 1 >= 2 * 0x7ff0000000000000ull - 1) {  // y is zero, inf or NaN
      if (2 * iy == 0) return 1.0;                       // (signalling NaNs do not occur here)
      if (ix == 0x3ff0000000000000ull) return 1.0;
      if (2 * ix > 2 * 0x7ff0000000000000ull || 2 * iy > 2 * 0x7ff0000000000000ull)
        return INFLX_GL_ADD(x, y);
      if (2 * ix == 2 * 0x3ff0000000000000ull) return 1.0;
      if ((2 * ix < 2 * 0x3ff0000000000000ull) == !(iy >> 63)) return 0.0;
      return INFLX_GL_MUL(y, y);
    }
    if (2 * ix - 1 >= 2 * 0x7ff0000000000000ull - 1) {  // x is zero, inf or NaN
      double x2 = INFLX_GL_MUL(x, x);
      if ((ix >> 63) && inflx_gl_checkint(iy) == 1) x2 = -x2;
      return (iy >> 63) ? INFLX_GL_DIV(1.0, x2) : x2;
    }
    if (ix >> 63) {  // finite x < 0
      const int yint = inflx_gl_checkint(iy);
      if (yint == 0) return INFLX_GL_NAN;
      if (yint == 1) sign_bias = 0x800ull << 7;
      ix &= 0x7fffffffffffffffull;
      topx &= 0x7ffu;
    }
    if ((topy & 0x7ffu) - 0x3beu >= 0x43eu - 0x3beu) {
      if (ix == 0x3ff0000000000000ull) return 1.0;
      if ((topy & 0x7ffu) < 0x3beu)  // |y| < 2^-65: x^y ~ 1 + y log x
        return ix > 0x3ff0000000000000ull ? INFLX_GL_ADD(1.0, y) : INFLX_GL_SUB(1.0, y);
      return (ix > 0x3ff0000000000000ull) == (topy < 0x800u) ? INFLX_GL_INF : 0.0;
    }
    if (topx == 0) {  // subnormal x: normalise so that the exponent becomes negative
      ix = INFLX_GL_BITS(INFLX_GL_MUL(x, 0x1p52));
      ix &= 0x7fffffffffffffffull;
      ix -= 52ull << 52;
    }
  }
  double ehi, elo;
  inflx_gl_pow_log_mul(ix, y, &ehi, &elo);
  // ---- exp_inline(ehi, elo, sign_bias) ----
  const unsigned abstop = (unsigned)(INFLX_GL_BITS(ehi) >> 52) & 0x7ff;
  int special = 0;
  if (abstop - 0x3c9u >= 0x3fu) {
    if ((int)(abstop - 0x3c9u) < 0) {  // tiny: 1 + x
      const double one = INFLX_GL_ADD(1.0, ehi);
      return sign_bias ? -one : one;
    }
    if (abstop >= 0x409u) {
      const double big = (INFLX_GL_BITS(ehi) >> 63) ? 0.0 : INFLX_GL_INF;
      return sign_bias ? -big : big;
    }
    special = 1;
  }
  return inflx_gl_exp_core(ehi, elo, 1, sign_bias, special);
}

// Per grid point (libm flavour "glibc-all"): the main path of pow inlined - x a positive normal,
// |y| neither huge nor tiny, exp(y log x) far from over- and underflow - everything else through
// the out-of-line routine.  Same operations, so the same bits.
INFLX_GL_INL double inflx_gl_pow_m(double x, double y) {
  const inflx_gl_u64 ix = INFLX_GL_BITS(x);
  const unsigned topx = (unsigned)(ix >> 52);
  const unsigned topy = (unsigned)(INFLX_GL_BITS(y) >> 52);
  if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ffu) - 0x3beu >= 0x43eu - 0x3beu)
    return inflx_gl_pow(x, y);
  double ehi, elo;
  inflx_gl_pow_log_mul(ix, y, &ehi, &elo);
  const unsigned abstop = (unsigned)(INFLX_GL_BITS(ehi) >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) return inflx_gl_pow(x, y);
  return inflx_gl_exp_core(ehi, elo, 1, 0, 0);
}

// ---- expm1, tanh ------------------------------------------------------------------------------------
// s_expm1.c (fdlibm) as built for the FMA ifunc variant (__expm1_fma), s_tanh.c (one build, unfused).
#define INFLX_GL_HAVE_TANH 1
INFLX_GL_INL double inflx_gl_add_exponent(double y, int k) {  // SET_HIGH_WORD(y, high + (k << 20))
  const inflx_gl_u64 b = INFLX_GL_BITS(y);
  const unsigned high = (unsigned)(b >> 32) + ((unsigned)k << 20);
  return INFLX_GL_FROM_BITS(((inflx_gl_u64)high << 32) | (b & 0xffffffffull));
}

INFLX_GL_FN double inflx_gl_expm1(double x) {
  const double ln2_hi = 0x1.62e42fee00000p-1, ln2_lo = 0x1.a39ef35793c76p-33;
  const double invln2 = 0x1.71547652b82fep+0;
  const double Q1 = -0x1.11111111110f4p-5, Q2 = 0x1.a01a019fe5585p-10, Q3 = -0x1.4ce199eaadbb7p-14;
  const double Q4 = 0x1.0cfca86e65239p-18, Q5 = -0x1.afdb76e09c32dp-23;
  const inflx_gl_u64 bits = INFLX_GL_BITS(x);
  const unsigned hx = (unsigned)(bits >> 32) & 0x7fffffffu;
  const int neg = (int)(bits >> 63);
  double hi, lo, c = 0.0;
  int k;
  if (hx >= 0x4043687Au) {      // |x| >= 56 ln2
    if (hx >= 0x40862E42u) {    // |x| >= 709.78
      if (hx >= 0x7ff00000u) {
        if (((bits >> 32) & 0xfffff) | (bits & 0xffffffffull)) return INFLX_GL_ADD(x, x);  // NaN
        return neg ? -1.0 : x;
      }
      if (x > 0x1.62e42fefa39efp+9) return INFLX_GL_INF;
    }
    if (neg) return -1.0;        // tiny - one
  }
  if (hx > 0x3fd62e42u) {        // |x| > 0.5 ln2
    if (hx < 0x3FF0A2B2u) {      // and |x| < 1.5 ln2
      if (!neg) { hi = INFLX_GL_SUB(x, ln2_hi); lo = ln2_lo; k = 1; }
      else { hi = INFLX_GL_ADD(x, ln2_hi); lo = -ln2_lo; k = -1; }
    } else {
      k = (int)INFLX_GL_ADD(neg ? -0.5 : 0.5, INFLX_GL_MUL(x, invln2));
      const double t = (double)k;
      hi = INFLX_GL_FMA(-ln2_hi, t, x);
      lo = INFLX_GL_MUL(t, ln2_lo);
    }
    x = INFLX_GL_SUB(hi, lo);
    c = INFLX_GL_SUB(INFLX_GL_SUB(hi, x), lo);
  } else if (hx < 0x3c900000u) {  // |x| < 2^-54
    return x;
  } else {
    k = 0;
  }
  const double hfx = INFLX_GL_MUL(x, 0.5);
  const double hxs = INFLX_GL_MUL(x, hfx);
  const double R2 = INFLX_GL_FMA(Q3, hxs, Q2);
  const double R3 = INFLX_GL_FMA(Q5, hxs, Q4);
  const double h2 = INFLX_GL_MUL(hxs, hxs);
  const double R1 = INFLX_GL_FMA(hxs, Q1, 1.0);
  const double h4 = INFLX_GL_MUL(h2, h2);
  const double r1 = INFLX_GL_FMA(h4, R3, INFLX_GL_FMA(h2, R2, R1));
  const double t = INFLX_GL_FMA(-r1, hfx, 3.0);
  double e = INFLX_GL_MUL(INFLX_GL_DIV(INFLX_GL_SUB(r1, t), INFLX_GL_FMA(-x, t, 6.0)), hxs);
  if (k == 0) return INFLX_GL_SUB(x, INFLX_GL_FMA(e, x, -hxs));
  e = INFLX_GL_FMA(INFLX_GL_SUB(e, c), x, -c);
  e = INFLX_GL_SUB(e, hxs);
  if (k == -1) return INFLX_GL_FMA(INFLX_GL_SUB(x, e), 0.5, -0.5);
  if (k == 1) {
    if (x < -0.25) return INFLX_GL_MUL(INFLX_GL_SUB(e, INFLX_GL_ADD(x, 0.5)), -2.0);
    return INFLX_GL_FMA(2.0, INFLX_GL_SUB(x, e), 1.0);
  }
  if (k <= -2 || k > 56) {  // suffices to return exp(x) - 1
    const double y = INFLX_GL_SUB(1.0, INFLX_GL_SUB(e, x));
    return INFLX_GL_SUB(inflx_gl_add_exponent(y, k), 1.0);
  }
  if (k < 20) {
    const double tk = INFLX_GL_FROM_BITS((inflx_gl_u64)(0x3ff00000u - (0x200000u >> k)) << 32);  // 1 - 2^-k
    return inflx_gl_add_exponent(INFLX_GL_SUB(tk, INFLX_GL_SUB(e, x)), k);
  }
  const double tk = INFLX_GL_FROM_BITS((inflx_gl_u64)((unsigned)(0x3ff - k) << 20) << 32);  // 2^-k
  return inflx_gl_add_exponent(INFLX_GL_ADD(INFLX_GL_SUB(x, INFLX_GL_ADD(e, tk)), 1.0), k);
}

INFLX_GL_FN double inflx_gl_tanh(double x) {
  const inflx_gl_u64 bits = INFLX_GL_BITS(x);
  const unsigned ix = (unsigned)(bits >> 32) & 0x7fffffffu;
  const int neg = (int)(bits >> 63);
  double z;
  if (ix >= 0x7ff00000u)  // inf or NaN
    return neg ? INFLX_GL_SUB(INFLX_GL_DIV(1.0, x), 1.0) : INFLX_GL_ADD(INFLX_GL_DIV(1.0, x), 1.0);
  if (ix < 0x40360000u) {  // |x| < 22
    if ((bits << 1) == 0) return x;
    if (ix < 0x3c800000u) return INFLX_GL_MUL(INFLX_GL_ADD(1.0, x), x);  // |x| < 2^-55
    const double ax = fabs(x);
    if (ix >= 0x3ff00000u) {  // |x| >= 1
      const double t = inflx_gl_expm1(INFLX_GL_ADD(ax, ax));
      z = INFLX_GL_SUB(1.0, INFLX_GL_DIV(2.0, INFLX_GL_ADD(t, 2.0)));
    } else {
      const double t = inflx_gl_expm1(INFLX_GL_MUL(ax, -2.0));
      z = INFLX_GL_DIV(-t, INFLX_GL_ADD(t, 2.0));
    }
  } else {
    z = 1.0;  // one - tiny
  }
  return neg ? -z : z;
}

// ---- sin, cos -----------------------------------------------------------------------------------------
// s_sin.c (IBM accurate mathematical library as simplified in glibc 2.28+) as built for the FMA ifunc
// variants __sin_fma / __cos_fma; branred.c (one build, unfused) for |x| >= 105414350.
#define INFLX_GL_HAVE_SINCOS 1
#define INFLX_GL_SC_BIG 0x1.8000000000000p+45

// sin(x + dx) and cos(x + dx) for |x| < ~0.86 from the 1/128-spaced table plus short series.
INFLX_GL_INL double inflx_gl_do_cos(double x, double dx) {
  const double sn3 = -0x1.5555555555515p-3, sn5 = 0x1.11110e829872fp-7;
  const double cs4 = -0x1.5555555555535p-5, cs6 = 0x1.6c16bedd9e239p-10;
  if (x < 0.0) dx = -dx;
  const double ax = fabs(x);
  const double u = INFLX_GL_ADD(INFLX_GL_SC_BIG, ax);
  x = INFLX_GL_ADD(INFLX_GL_SUB(ax, INFLX_GL_SUB(u, INFLX_GL_SC_BIG)), dx);
  const int k = (int)(unsigned)INFLX_GL_BITS(u) * 4;
  const double xx = INFLX_GL_MUL(x, x);
  const double s = INFLX_GL_FMA(INFLX_GL_MUL(x, xx), INFLX_GL_FMA(sn5, xx, sn3), x);
  const double c = INFLX_GL_MUL(xx, INFLX_GL_FMA(INFLX_GL_FMA(cs6, xx, cs4), xx, 0.5));
  const double sn = inflx_gl_sincostab[k], ssn = inflx_gl_sincostab[k + 1];
  const double cs = inflx_gl_sincostab[k + 2], ccs = inflx_gl_sincostab[k + 3];
  const double cor = INFLX_GL_FMA(-s, sn, INFLX_GL_FMA(-c, cs, INFLX_GL_FMA(-s, ssn, ccs)));
  return INFLX_GL_ADD(cs, cor);
}

INFLX_GL_INL double inflx_gl_do_sin(double x, double dx) {
  const double sn3 = -0x1.5555555555515p-3, sn5 = 0x1.11110e829872fp-7;
  const double cs4 = -0x1.5555555555535p-5, cs6 = 0x1.6c16bedd9e239p-10;
  const double xold = x;
  if (fabs(x) < 0.126) {  // TAYLOR_SIN(x*x, x, dx)
    const double s5 = -0x1.addffc2fcdf59p-26, s4 = 0x1.71de27b9a7ed9p-19, s3 = -0x1.a01a019db08b8p-13;
    const double s2 = 0x1.1111111110ecep-7, s1 = -0x1.5555555555555p-3;
    const double xx = INFLX_GL_MUL(x, x);
    double p = INFLX_GL_FMA(s5, xx, s4);
    p = INFLX_GL_FMA(p, xx, s3);
    p = INFLX_GL_FMA(p, xx, s2);
    p = INFLX_GL_FMA(p, xx, s1);
    const double t = INFLX_GL_FMA(xx, INFLX_GL_FMA(p, x, -INFLX_GL_MUL(dx, 0.5)), dx);
    return INFLX_GL_ADD(x, t);
  }
  if (x <= 0.0) dx = -dx;
  const double ax = fabs(x);
  const double u = INFLX_GL_ADD(INFLX_GL_SC_BIG, ax);
  x = INFLX_GL_SUB(ax, INFLX_GL_SUB(u, INFLX_GL_SC_BIG));
  const int k = (int)(unsigned)INFLX_GL_BITS(u) * 4;
  const double xx = INFLX_GL_MUL(x, x);
  const double s = INFLX_GL_ADD(x, INFLX_GL_FMA(INFLX_GL_MUL(x, xx), INFLX_GL_FMA(sn5, xx, sn3), dx));
  const double c = INFLX_GL_FMA(x, dx, INFLX_GL_MUL(xx, INFLX_GL_FMA(INFLX_GL_FMA(cs6, xx, cs4), xx, 0.5)));
  const double sn = inflx_gl_sincostab[k], ssn = inflx_gl_sincostab[k + 1];
  const double cs = inflx_gl_sincostab[k + 2], ccs = inflx_gl_sincostab[k + 3];
  const double cor = INFLX_GL_FMA(s, cs, INFLX_GL_FMA(-c, sn, INFLX_GL_FMA(s, ccs, ssn)));
  const double r = INFLX_GL_ADD(sn, cor);
  return INFLX_GL_FROM_BITS((INFLX_GL_BITS(r) & 0x7fffffffffffffffull) |
                            (INFLX_GL_BITS(xold) & 0x8000000000000000ull));
}

// x -> (a + da, quadrant) for 2.426 < |x| < 105414350: three-step Cody-Waite with pi/2 to 136 bits
INFLX_GL_INL int inflx_gl_reduce_sincos(double x, double* a, double* da) {
  const double hpinv = 0x1.45f306dc9c883p-1, toint = 0x1.8000000000000p+52;
  const double mp1 = 0x1.921fb58000000p+0, mp2 = -0x1.dde973c000000p-27;
  const double pp3 = -0x1.cb3b398000000p-55, pp4 = -0x1.d747f23e32ed7p-83;
  const double t = INFLX_GL_FMA(x, hpinv, toint);
  const double xn = INFLX_GL_SUB(t, toint);
  const int n = (int)(unsigned)INFLX_GL_BITS(t) & 3;
  const double y = INFLX_GL_FMA(-mp2, xn, INFLX_GL_FMA(-mp1, xn, x));
  const double t2 = INFLX_GL_FMA(-xn, pp3, y);
  double db = INFLX_GL_FMA(-pp3, xn, INFLX_GL_SUB(y, t2));
  const double b = INFLX_GL_FMA(-xn, pp4, t2);
  db = INFLX_GL_ADD(db, INFLX_GL_FMA(-xn, pp4, INFLX_GL_SUB(t2, b)));
  *a = b;
  *da = db;
  return n;
}

// branred.c: x -> (a + aa, quadrant) for huge |x| with 2/pi from a table of 24-bit digits
INFLX_GL_INL void inflx_gl_branred_half(double x, double* b_out, double* bb_out, double* sum_out) {
  const double big = 0x1.8000000000000p+52, big1 = 0x1.8000000000000p+54, tm24 = 0x1p-24;
  double r[6];
  int k = (int)((INFLX_GL_BITS(x) >> 52) & 2047);
  k = (k - 450) / 24;
  if (k < 0) k = 0;
  double gor = INFLX_GL_FROM_BITS((inflx_gl_u64)(0x63f00000u - ((unsigned)(k * 24) << 20)) << 32);  // 2^576 / 2^(24k)
  for (int i = 0; i < 6; i++) {
    r[i] = INFLX_GL_MUL(INFLX_GL_MUL(x, inflx_gl_toverp[k + i]), gor);
    gor = INFLX_GL_MUL(gor, tm24);
  }
  double sum = 0.0;
  for (int i = 0; i < 3; i++) {
    const double s = INFLX_GL_SUB(INFLX_GL_ADD(r[i], big), big);
    sum = INFLX_GL_ADD(sum, s);
    r[i] = INFLX_GL_SUB(r[i], s);
  }
  double t = 0.0;
  for (int i = 0; i < 6; i++) t = INFLX_GL_ADD(t, r[5 - i]);
  double bb = INFLX_GL_SUB(r[0], t);
  for (int i = 1; i < 6; i++) bb = INFLX_GL_ADD(bb, r[i]);
  double s = INFLX_GL_SUB(INFLX_GL_ADD(t, big), big);
  sum = INFLX_GL_ADD(sum, s);
  t = INFLX_GL_SUB(t, s);
  const double b = INFLX_GL_ADD(t, bb);
  bb = INFLX_GL_ADD(INFLX_GL_SUB(t, b), bb);
  s = INFLX_GL_SUB(INFLX_GL_ADD(sum, big1), big1);
  sum = INFLX_GL_SUB(sum, s);
  *b_out = b;
  *bb_out = bb;
  *sum_out = sum;
}

INFLX_GL_FN int inflx_gl_branred(double x, double* a, double* aa) {
  const double split = 0x1.0000002000000p+27, hp0 = 0x1.921fb54442d18p+0, hp1 = 0x1.1a62633145c07p-54;
  const double mp1 = 0x1.921fb58000000p+0, mp2 = -0x1.dde9740000000p-27;
  x = INFLX_GL_MUL(x, 0x1p-600);
  double t = INFLX_GL_MUL(x, split);
  const double x1 = INFLX_GL_SUB(t, INFLX_GL_SUB(t, x));
  const double x2 = INFLX_GL_SUB(x, x1);
  double b1, bb1, sum1, b2, bb2, sum2;
  inflx_gl_branred_half(x1, &b1, &bb1, &sum1);
  inflx_gl_branred_half(x2, &b2, &bb2, &sum2);
  double sum = INFLX_GL_ADD(sum1, sum2);
  double b = INFLX_GL_ADD(b1, b2);
  double bb = (fabs(b1) > fabs(b2)) ? INFLX_GL_ADD(INFLX_GL_SUB(b1, b), b2)
                                    : INFLX_GL_ADD(INFLX_GL_SUB(b2, b), b1);
  if (b > 0.5) {
    b = INFLX_GL_SUB(b, 1.0);
    sum = INFLX_GL_ADD(sum, 1.0);
  } else if (b < -0.5) {
    b = INFLX_GL_ADD(b, 1.0);
    sum = INFLX_GL_SUB(sum, 1.0);
  }
  double s = INFLX_GL_ADD(b, INFLX_GL_ADD(INFLX_GL_ADD(bb, bb1), bb2));
  t = INFLX_GL_ADD(INFLX_GL_ADD(INFLX_GL_SUB(b, s), bb), INFLX_GL_ADD(bb1, bb2));
  b = INFLX_GL_MUL(s, split);
  const double t1 = INFLX_GL_SUB(b, INFLX_GL_SUB(b, s));
  const double t2 = INFLX_GL_SUB(s, t1);
  b = INFLX_GL_MUL(s, hp0);
  bb = INFLX_GL_ADD(
      INFLX_GL_ADD(INFLX_GL_ADD(INFLX_GL_SUB(INFLX_GL_MUL(t1, mp1), b), INFLX_GL_MUL(t1, mp2)),
                   INFLX_GL_MUL(t2, mp1)),
      INFLX_GL_ADD(INFLX_GL_ADD(INFLX_GL_MUL(t2, mp2), INFLX_GL_MUL(s, hp1)), INFLX_GL_MUL(t, hp0)));
  s = INFLX_GL_ADD(b, bb);
  t = INFLX_GL_ADD(INFLX_GL_SUB(b, s), bb);
  *a = s;
  *aa = t;
  return ((int)sum) & 3;
}

INFLX_GL_INL double inflx_gl_do_sincos(double a, double da, int n) {
  const double r = (n & 1) ? inflx_gl_do_cos(a, da) : inflx_gl_do_sin(a, da);
  return (n & 2) ? -r : r;
}

INFLX_GL_FN double inflx_gl_sin(double x) {
  const double hp0 = 0x1.921fb54442d18p+0, hp1 = 0x1.1a62633145c07p-54;
  const unsigned k = (unsigned)(INFLX_GL_BITS(x) >> 32) & 0x7fffffffu;
  double a, da;
  if (k < 0x3e500000u) return x;                       // |x| < 2^-26
  if (k < 0x3feb6000u) return inflx_gl_do_sin(x, 0.0);  // |x| < 0.855469
  if (k < 0x400368fdu) {                               // |x| < 2.426265
    const double r = inflx_gl_do_cos(INFLX_GL_SUB(hp0, fabs(x)), hp1);
    return INFLX_GL_FROM_BITS((INFLX_GL_BITS(r) & 0x7fffffffffffffffull) |
                              (INFLX_GL_BITS(x) & 0x8000000000000000ull));
  }
  if (k < 0x419921FBu) {                               // |x| < 105414350
    const int n = inflx_gl_reduce_sincos(x, &a, &da);
    return inflx_gl_do_sincos(a, da, n);
  }
  if (k < 0x7ff00000u) {
    const int n = inflx_gl_branred(x, &a, &da);
    return inflx_gl_do_sincos(a, da, n);
  }
  return INFLX_GL_DIV(x, x);                           // inf, NaN
}

INFLX_GL_FN double inflx_gl_cos(double x) {
  const double hp0 = 0x1.921fb54442d18p+0, hp1 = 0x1.1a62633145c07p-54;
  const unsigned k = (unsigned)(INFLX_GL_BITS(x) >> 32) & 0x7fffffffu;
  double a, da;
  if (k < 0x3e400000u) return 1.0;                     // |x| < 2^-27
  if (k < 0x3feb6000u) return inflx_gl_do_cos(x, 0.0);
  if (k < 0x400368fdu) {
    const double y = INFLX_GL_SUB(hp0, fabs(x));
    a = INFLX_GL_ADD(y, hp1);
    da = INFLX_GL_ADD(INFLX_GL_SUB(y, a), hp1);
    return inflx_gl_do_sin(a, da);
  }
  if (k < 0x419921FBu) {
    const int n = inflx_gl_reduce_sincos(x, &a, &da);
    return inflx_gl_do_sincos(a, da, n + 1);
  }
  if (k < 0x7ff00000u) {
    const int n = inflx_gl_branred(x, &a, &da);
    return inflx_gl_do_sincos(a, da, n + 1);
  }
  return INFLX_GL_DIV(x, x);
}

// ---- atan, tan --------------------------------------------------------------------------------------
// The reference's epilogue (src/anguelova.rs:130-134) takes delta = atan(|V_wv / V_vv|) and
// eta = omega * tan(delta) - 3 through Rust's f64::atan / f64::tan, i.e. the platform libm.  glibc
// 2.39's atan and tan are IBM's "accurate mathematical library" routines (sysdeps/ieee754/dbl-64/
// s_atan.c, s_tan.c) with the multi-precision fall-backs removed: table look-up (uatan.tbl: 241 x 7,
// utan.tbl: 186 x 4; extracted into inflx_glibc_tables.cuh) + short polynomials + double-double
// corrections.  Restated from the FMA ifunc variants of the shipped binary (__atan_fma, __tan_fma),
// contraction included.  The grid kernels call them in libm flavour "glibc-all" only: the default
// epilogue uses inflx_atan_tan (inflx_device.cuh; <= 1 ulp, half the instructions, no gathers).
#define INFLX_GL_HAVE_ATAN_TAN 1
// Inlined into the epilogue: measured 4-6 % cheaper than out-of-line calls (angular, doc 16384^2:
// +12 % over the default epilogue instead of +16...18 %; -DINFLX_GL_CALL_ATAN_TAN restores calls).
#ifdef INFLX_GL_CALL_ATAN_TAN
#define INFLX_GL_EPI INFLX_GL_FN
#else
#define INFLX_GL_EPI INFLX_GL_INL
#endif

INFLX_GL_INL double inflx_gl_copysign(double mag, double sgn) {
  return INFLX_GL_FROM_BITS((INFLX_GL_BITS(mag) & 0x7fffffffffffffffull) |
                            (INFLX_GL_BITS(sgn) & 0x8000000000000000ull));
}

INFLX_GL_EPI double inflx_gl_atan(double x) {
  const double d3 = -0x1.5555555555555p-2, d5 = 0x1.99999999997fdp-3, d7 = -0x1.24924923f7603p-3;
  const double d9 = 0x1.c71c6e5129a3bp-4, d11 = -0x1.7458022b13c25p-4, d13 = 0x1.375f08b31cbcep-4;
  const double hpi = 0x1.921fb54442d18p+0, hpi1 = 0x1.1a62633145c07p-54, two52 = 0x1p+52;
  const inflx_gl_u64 bits = INFLX_GL_BITS(x);
  if (((bits >> 52) & 0x7ff) == 0x7ff && (bits & 0x000fffffffffffffull)) return INFLX_GL_ADD(x, x);
  const double u = fabs(x);
  if (u < 1.0) {
    if (u < 0.0625) {
      if (u < 0x1.bb67ap-27) return x;
      const double v = INFLX_GL_MUL(x, x);
      double t = INFLX_GL_FMA(v, d13, d11);
      t = INFLX_GL_FMA(v, t, d9);
      t = INFLX_GL_FMA(v, t, d7);
      t = INFLX_GL_FMA(v, t, d5);
      t = INFLX_GL_FMA(v, t, d3);
      return INFLX_GL_FMA(INFLX_GL_MUL(x, v), t, x);
    }
    const int i = (int)INFLX_GL_SUB(INFLX_GL_FMA(u, 256.0, two52), two52) - 16;
    const double* c = inflx_gl_atan_cij + 7 * i;
    const double z = INFLX_GL_SUB(u, c[0]);
    double yy = INFLX_GL_FMA(z, c[6], c[5]);
    yy = INFLX_GL_FMA(z, yy, c[4]);
    yy = INFLX_GL_FMA(z, yy, c[3]);
    yy = INFLX_GL_FMA(z, yy, c[2]);
    return inflx_gl_copysign(INFLX_GL_FMA(yy, z, c[1]), x);
  }
  if (u < 16.0) {  // atan(u) = pi/2 - atan(1/u), 1/u to double-double
    const double w = INFLX_GL_DIV(1.0, u);
    const double t1 = INFLX_GL_MUL(w, u);
    const double t2 = INFLX_GL_FMA(u, w, -t1);
    const double r = INFLX_GL_SUB(INFLX_GL_SUB(1.0, t1), t2);
    const int i = (int)INFLX_GL_SUB(INFLX_GL_FMA(w, 256.0, two52), two52) - 16;
    const double* c = inflx_gl_atan_cij + 7 * i;
    const double z = INFLX_GL_FMA(r, w, INFLX_GL_SUB(w, c[0]));
    double yy = INFLX_GL_FMA(z, c[6], c[5]);
    yy = INFLX_GL_FMA(z, yy, c[4]);
    yy = INFLX_GL_FMA(z, yy, c[3]);
    yy = INFLX_GL_FMA(z, yy, c[2]);
    yy = INFLX_GL_FMA(-yy, z, hpi1);
    return inflx_gl_copysign(INFLX_GL_ADD(INFLX_GL_SUB(hpi, c[1]), yy), x);
  }
  if (u < 0x1.49ff2p+52) {
    const double w = INFLX_GL_DIV(1.0, u);
    const double t1 = INFLX_GL_MUL(w, u);
    const double t3 = INFLX_GL_SUB(hpi, w);
    const double v = INFLX_GL_MUL(w, w);
    double p = INFLX_GL_FMA(v, d13, d11);
    p = INFLX_GL_FMA(v, p, d9);
    p = INFLX_GL_FMA(v, p, d7);
    p = INFLX_GL_FMA(v, p, d5);
    p = INFLX_GL_FMA(v, p, d3);
    const double cor = INFLX_GL_ADD(INFLX_GL_SUB(INFLX_GL_SUB(hpi, t3), w), hpi1);
    const double t2 = INFLX_GL_FMA(u, w, -t1);
    const double r = INFLX_GL_SUB(INFLX_GL_SUB(1.0, t1), t2);
    const double s = INFLX_GL_FMA(-r, w, cor);
    const double yy = INFLX_GL_FMA(-INFLX_GL_MUL(w, v), p, s);
    return inflx_gl_copysign(INFLX_GL_ADD(t3, yy), x);
  }
  return x > 0.0 ? hpi : -hpi;
}

// tan for |x| <= 25 (the epilogue's argument is an atan: 0 <= x <= pi/2, or NaN).  Larger finite
// arguments - s_tan.c's 25 < |x| <= 1e8 and __branred ranges, unreachable from the path - are NOT
// restated and return NaN.
INFLX_GL_EPI double inflx_gl_tan(double x) {
  const double d3 = 0x1.5555555555555p-2, d5 = 0x1.11111111107c6p-3, d7 = 0x1.ba1ba1cdb8745p-5;
  const double d9 = 0x1.664ed49cfc666p-6, d11 = 0x1.2385a3cf2e4eap-7;
  const double e0 = 0x1.5555555554dbdp-2, e1 = 0x1.11112e0a6b45fp-3, mfftnhf = -15.5;
  const inflx_gl_u64 bits = INFLX_GL_BITS(x);
  if (((bits >> 52) & 0x7ff) == 0x7ff) return INFLX_GL_SUB(x, x);  // inf, NaN
  const double w = (x < 0.0) ? -x : x;
  if (w <= 0x1.b096cp-27) return x;
  if (w <= 0x1.f212dp-5) {
    const double x2 = INFLX_GL_MUL(x, x);
    double t = INFLX_GL_FMA(x2, d11, d9);
    t = INFLX_GL_FMA(x2, t, d7);
    t = INFLX_GL_FMA(x2, t, d5);
    t = INFLX_GL_FMA(x2, t, d3);
    return INFLX_GL_FMA(INFLX_GL_MUL(x, x2), t, x);
  }
  if (w <= 0x1.92f1ap-1) {
    const int i = (int)INFLX_GL_FMA(w, 256.0, mfftnhf);
    const double* g = inflx_gl_tan_xfg + 4 * i;
    const double z = INFLX_GL_SUB(w, g[0]);
    const double z2 = INFLX_GL_MUL(z, z);
    const double pz = INFLX_GL_FMA(INFLX_GL_MUL(z, z2), INFLX_GL_FMA(z2, e1, e0), z);
    const double fi = g[1], gi = g[2];
    const double t2 = INFLX_GL_DIV(INFLX_GL_MUL(INFLX_GL_ADD(fi, gi), pz), INFLX_GL_SUB(gi, pz));
    return INFLX_GL_MUL(INFLX_GL_ADD(t2, fi), (x < 0.0) ? -1.0 : 1.0);
  }
  if (!(w <= 25.0)) return INFLX_GL_NAN;  // not restated (see above)
  // 0.787 < |x| <= 25: x = n pi/2 + (a + da)
  const double hpinv = 0x1.45f306dc9c883p-1, toint = 0x1.8p+52;
  const double mp1 = 0x1.921fb58p+0, mp2 = -0x1.dde973cp-27, mp3 = -0x1.cb3b399d747f2p-55;
  const double t = INFLX_GL_FMA(x, hpinv, toint);
  const double xn = INFLX_GL_SUB(t, toint);
  const int n = (int)(INFLX_GL_BITS(t) & 1);
  const double t1 = INFLX_GL_FMA(-xn, mp2, INFLX_GL_FMA(-xn, mp1, x));
  const double a = INFLX_GL_FMA(-xn, mp3, t1);
  const double da = INFLX_GL_FMA(-xn, mp3, INFLX_GL_SUB(t1, a));
  double ya = a, yya = da, sy = 1.0;
  if (0.0 > a) {
    ya = -a;
    yya = -da;
    sy = -1.0;
  }
  if (!(0x1.f212dp-5 < ya)) {  // |a| <= 0.0608: polynomial, and for odd n -cot through a dd division
    const double a2 = INFLX_GL_MUL(a, a);
    double p = INFLX_GL_FMA(a2, d11, d9);
    p = INFLX_GL_FMA(a2, p, d7);
    p = INFLX_GL_FMA(a2, p, d5);
    p = INFLX_GL_FMA(a2, p, d3);
    const double t2 = INFLX_GL_FMA(INFLX_GL_MUL(a, a2), p, da);
    const double b = INFLX_GL_ADD(a, t2);
    if (!n) return b;
    const double db = (fabs(a) > fabs(t2)) ? INFLX_GL_ADD(INFLX_GL_SUB(a, b), t2)
                                           : INFLX_GL_ADD(INFLX_GL_SUB(t2, b), a);
    const double c = INFLX_GL_DIV(1.0, b);
    const double ph = INFLX_GL_MUL(c, b);
    const double pl = INFLX_GL_FMA(c, b, -ph);
    double r = INFLX_GL_ADD(INFLX_GL_SUB(INFLX_GL_SUB(1.0, ph), pl), 0.0);
    r = INFLX_GL_FMA(-db, c, r);
    const double cc = INFLX_GL_DIV(r, b);
    const double zh = INFLX_GL_ADD(c, cc);
    const double zl = INFLX_GL_ADD(INFLX_GL_SUB(c, zh), cc);
    return -INFLX_GL_ADD(zl, zh);
  }
  const int i = (int)INFLX_GL_FMA(ya, 256.0, mfftnhf);
  const double* g = inflx_gl_tan_xfg + 4 * i;
  const double z = INFLX_GL_ADD(INFLX_GL_SUB(ya, g[0]), yya);
  const double z2 = INFLX_GL_MUL(z, z);
  const double pz = INFLX_GL_FMA(INFLX_GL_MUL(z, z2), INFLX_GL_FMA(z2, e1, e0), z);
  const double fi = g[1], gi = g[2];
  const double s = INFLX_GL_MUL(INFLX_GL_ADD(fi, gi), pz);
  if (n) return INFLX_GL_MUL(INFLX_GL_SUB(gi, INFLX_GL_DIV(s, INFLX_GL_ADD(pz, fi))), -sy);
  return INFLX_GL_MUL(INFLX_GL_ADD(INFLX_GL_DIV(s, INFLX_GL_SUB(gi, pz)), fi), sy);
}
