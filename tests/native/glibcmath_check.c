/* Host-side check of inflatox_b200/csrc/inflx_glibcmath.cuh against the HOST'S libm (TEST CODE).
 * The claim under test is bit identity: every result must be the one `pow` / `exp` / `log` / ... of
 * the libm the oracle links returns (NaNs compare equal to NaNs).  Prints one line of counters;
 * tests/test_glibcmath.py parses it.  Build: gcc -O2 -march=native -ffp-contract=off -fopenmp. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include INFLX_GLIBCMATH_HEADER

typedef struct { uint64_t s; } rng_t;
static inline uint64_t next(rng_t* r) { /* splitmix64 */
  uint64_t z = (r->s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static inline double unif(rng_t* r) { return (double)(next(r) >> 11) * 0x1p-53; }
static inline double anybits(rng_t* r) { return INFLX_GL_FROM_BITS(next(r)); }
static inline int same(double a, double b) {
  return INFLX_GL_BITS(a) == INFLX_GL_BITS(b) || (a != a && b != b);
}

/* argument families: `f` cycles through them */
static double arg_x(rng_t* r, int f) {
  switch (f) {
    case 0: return 0.4 + 0.2 * unif(r);                         /* EGNO rows */
    case 1: return exp(40 * (unif(r) - 0.5));
    case 2: return 1.0 + (unif(r) - 0.5) * ldexp(1.0, -(int)(50 * unif(r)));
    case 3: return ldexp(1 + unif(r), (int)(2100 * (unif(r) - 0.5)));
    case 4: return anybits(r);                                  /* every class incl. NaN, inf, subnormal */
    case 5: return -exp(10 * (unif(r) - 0.5));
    default: return 36 * unif(r);
  }
}
static double arg_y(rng_t* r, int f) {
  switch (f) {
    case 0: return -3.0 * (0.05 + 2 * unif(r));
    case 1: return 40 * (unif(r) - 0.5);
    case 2: return ldexp(unif(r) - 0.5, (int)(60 * unif(r)));
    case 3: return 3 * (unif(r) - 0.5);
    case 4: return anybits(r);
    case 5: return (double)(long)(40 * (unif(r) - 0.5));        /* integer exponents (negative x) */
    default: return (double)(long)(1 + 8 * unif(r)) * (unif(r) < 0.3 ? 0.5 : 1.0);
  }
}

int main(int argc, char** argv) {
  long n = argc > 1 ? atol(argv[1]) : 1000000;
  long bad_pow = 0, bad_exp = 0, bad_log = 0, bad_sin = 0, bad_cos = 0, bad_tanh = 0, bad_expm1 = 0, bad_special = 0, bad_atan = 0, bad_tan = 0;
  double worst[9][2] = {{0}};
#pragma omp parallel for reduction(+ : bad_pow, bad_exp, bad_log, bad_sin, bad_cos, bad_tanh, bad_expm1, bad_atan, bad_tan) schedule(static)
  for (long i = 0; i < n; i++) {
    rng_t r = {0x1234567ull + 0x9e3779b97f4a7c15ull * (uint64_t)i};
    const int f = (int)(i % 7);
    const double x = arg_x(&r, f), y = arg_y(&r, f);
    if (!same(inflx_gl_pow(x, y), pow(x, y))) { bad_pow++; worst[0][0] = x; worst[0][1] = y; }
    if (!same(inflx_gl_pow_m(x, y), pow(x, y))) { bad_pow++; worst[0][0] = x; worst[0][1] = y; }  /* per-point variant */
    double xe;
    switch (i % 5) {
      case 0: xe = 1500 * (unif(&r) - 0.5); break;
      case 1: xe = ldexp(unif(&r) - 0.5, -(int)(70 * unif(&r))); break;
      case 2: xe = -745.2 + 40 * (unif(&r) - 0.5); break;       /* subnormal results */
      case 3: xe = 709.7 + 2 * (unif(&r) - 0.5); break;         /* overflow threshold */
      default: xe = anybits(&r);
    }
    if (!same(inflx_gl_exp(xe), exp(xe))) { bad_exp++; worst[1][0] = xe; }
    const double xl = (i % 6 == 5) ? ldexp(unif(&r), -1040) : fabs(x);
    if (!same(inflx_gl_log(xl), log(xl))) { bad_log++; worst[2][0] = xl; }
    if (!same(inflx_gl_log(x), log(x))) { bad_log++; worst[2][0] = x; }
#ifdef INFLX_GL_HAVE_SINCOS
    double xs;
    switch (i % 6) {
      case 0: xs = 4 * M_PI * unif(&r); break;                                  /* d5: [0, 4 pi] */
      case 1: xs = 2000 * (unif(&r) - 0.5); break;
      case 2: xs = ldexp(unif(&r) - 0.5, -(int)(40 * unif(&r))); break;
      case 3: xs = (double)(long)(1000 * unif(&r)) * M_PI_2 * (1 + (unif(&r) - 0.5) * 1e-9); break;
      case 4: xs = 2e8 * (unif(&r) - 0.5); break;
      default: xs = anybits(&r);
    }
    if (!same(inflx_gl_sin(xs), sin(xs))) { bad_sin++; worst[3][0] = xs; }
    if (!same(inflx_gl_cos(xs), cos(xs))) { bad_cos++; worst[4][0] = xs; }
#endif
#ifdef INFLX_GL_HAVE_TANH
    double xt;
    switch (i % 4) {
      case 0: xt = 60 * (unif(&r) - 0.5); break;
      case 1: xt = ldexp(unif(&r) - 0.5, -(int)(60 * unif(&r))); break;
      case 2: xt = (2 * unif(&r) - 1) / (0.05 + 1.95 * unif(&r)); break;        /* C5: x / L */
      default: xt = anybits(&r);
    }
    if (!same(inflx_gl_tanh(xt), tanh(xt))) { bad_tanh++; worst[5][0] = xt; }
    if (!same(inflx_gl_expm1(xt), expm1(xt))) { bad_expm1++; worst[6][0] = xt; }
    if (!same(inflx_gl_expm1(xe), expm1(xe))) { bad_expm1++; worst[6][0] = xe; }
#endif
#ifdef INFLX_GL_HAVE_ATAN_TAN
    double xa;
    switch (i % 6) {
      case 0: xa = exp(60 * (unif(&r) - 0.5)); break;                           /* every branch of atan */
      case 1: xa = 20 * (unif(&r) - 0.5); break;
      case 2: xa = ldexp(unif(&r) - 0.5, -(int)(40 * unif(&r))); break;
      case 3: xa = (unif(&r) - 0.5) * 2.2; break;                               /* around the 1/16 and 1 seams */
      case 4: xa = ldexp(1 + unif(&r), (int)(120 * (unif(&r) - 0.5))); break;
      default: xa = anybits(&r);
    }
    if (!same(inflx_gl_atan(xa), atan(xa))) { bad_atan++; worst[7][0] = xa; }
    double xq;
    switch (i % 5) {
      case 0: xq = M_PI_2 * unif(&r); break;                                    /* the epilogue's range */
      case 1: xq = atan(exp(60 * (unif(&r) - 0.5))); break;                     /* tan(atan(y)), any y */
      case 2: xq = 50 * (unif(&r) - 0.5); break;                                /* |x| <= 25 */
      case 3: xq = ldexp(unif(&r) - 0.5, -(int)(40 * unif(&r))); break;
      default: xq = (double)(long)(16 * unif(&r)) * M_PI_2 * (1 + (unif(&r) - 0.5) * ldexp(1.0, -(int)(50 * unif(&r))));
    }
    if (fabs(xq) <= 25.0 && !same(inflx_gl_tan(xq), tan(xq))) { bad_tan++; worst[8][0] = xq; }
#endif
  }
  /* hand-picked irregular arguments: signs of zeros and infinities included */
  const double sp[][2] = {{0, 2}, {-1, 2}, {-8, 1.0 / 3}, {INFINITY, -1}, {NAN, 1}, {2, NAN}, {2, 0},
                          {1, 5}, {4, 0.5}, {2, 1023.5}, {2, -1074}, {1e-310, 2}, {10, 308.5},
                          {0.5, 1e10}, {-0.0, -1}, {3, INFINITY}, {-0.0, 3}, {-0.0, -3}, {0.0, -2},
                          {-INFINITY, 3}, {-INFINITY, -3}, {-INFINITY, 2.5}, {-2, 3}, {-2, 4},
                          {-2, 1075}, {-2, -1075}, {-2, 0.5}, {1, NAN}, {NAN, 0}, {-1, INFINITY},
                          {0.5, -INFINITY}, {2, -INFINITY}, {2, 1e-300}, {0.5, 1e-300}, {2, 1024},
                          {2, -1022}, {2, -1023}, {2, -1050.5}, {-3, 1e300}, {1e-320, -1},
                          {-1e-320, 3}, {1.0000000001, 1e13}, {0.9999999999, 1e13}};
  for (unsigned i = 0; i < sizeof sp / sizeof sp[0]; i++)
    bad_special += !same(inflx_gl_pow(sp[i][0], sp[i][1]), pow(sp[i][0], sp[i][1]));
  const double s1[] = {0.0, -0.0, -1.0, INFINITY, -INFINITY, NAN, 1e-310, 1.0, 710.0, -750.0, 1e300,
                       -1e300, 709.782712893384, 709.782712893385, -745.13321910194, -745.2, -708.4,
                       0.9375, 1.0647, 1e-20, -1e-20, 5e-324, 1.5, 0x1p-1022};
  for (unsigned i = 0; i < sizeof s1 / sizeof s1[0]; i++) {
    bad_special += !same(inflx_gl_log(s1[i]), log(s1[i]));
    bad_special += !same(inflx_gl_exp(s1[i]), exp(s1[i]));
#ifdef INFLX_GL_HAVE_SINCOS
    bad_special += !same(inflx_gl_sin(s1[i]), sin(s1[i]));
    bad_special += !same(inflx_gl_cos(s1[i]), cos(s1[i]));
#endif
#ifdef INFLX_GL_HAVE_TANH
    bad_special += !same(inflx_gl_tanh(s1[i]), tanh(s1[i]));
    bad_special += !same(inflx_gl_expm1(s1[i]), expm1(s1[i]));
#endif
#ifdef INFLX_GL_HAVE_ATAN_TAN
    bad_special += !same(inflx_gl_atan(s1[i]), atan(s1[i]));
    bad_special += !same(inflx_gl_atan(-s1[i]), atan(-s1[i]));
    if (!(fabs(s1[i]) > 25.0)) bad_special += !same(inflx_gl_tan(s1[i]), tan(s1[i]));
#endif
  }
#ifdef INFLX_GL_HAVE_ATAN_TAN
  const double s2[] = {0.0625, 1.0, 16.0, 0x1.49ff2p+52, 0x1.bb67ap-27, 0x1.b096cp-27, 0x1.f212dp-5,
                       0x1.92f1ap-1, 25.0, M_PI_2, M_PI_4, 0x1.921fb54442d18p+0, 0x1.921fb54442d19p+0,
                       0x1.921fb54442d17p+0, 3 * M_PI_2, 0.99999999999999989, 15.999999999999998};
  for (unsigned i = 0; i < sizeof s2 / sizeof s2[0]; i++)
    for (int sg = -1; sg <= 1; sg += 2)
      for (int d = -2; d <= 2; d++) {
        double v = s2[i];
        for (int k = 0; k < (d < 0 ? -d : d); k++) v = nextafter(v, d < 0 ? 0.0 : INFINITY);
        v *= sg;
        bad_special += !same(inflx_gl_atan(v), atan(v));
        if (fabs(v) <= 25.0) bad_special += !same(inflx_gl_tan(v), tan(v));
      }
#endif
  printf("n=%ld bad_pow=%ld bad_exp=%ld bad_log=%ld bad_sin=%ld bad_cos=%ld bad_tanh=%ld bad_expm1=%ld bad_atan=%ld bad_tan=%ld bad_special=%ld\n",
         n, bad_pow, bad_exp, bad_log, bad_sin, bad_cos, bad_tanh, bad_expm1, bad_atan, bad_tan, bad_special);
  if (bad_pow) printf("pow example: x=%a y=%a ours=%a libm=%a\n", worst[0][0], worst[0][1],
                      inflx_gl_pow(worst[0][0], worst[0][1]), pow(worst[0][0], worst[0][1]));
  if (bad_exp) printf("exp example: x=%a ours=%a libm=%a\n", worst[1][0], inflx_gl_exp(worst[1][0]), exp(worst[1][0]));
  if (bad_log) printf("log example: x=%a ours=%a libm=%a\n", worst[2][0], inflx_gl_log(worst[2][0]), log(worst[2][0]));
#ifdef INFLX_GL_HAVE_SINCOS
  if (bad_sin) printf("sin example: x=%a ours=%a libm=%a\n", worst[3][0], inflx_gl_sin(worst[3][0]), sin(worst[3][0]));
  if (bad_cos) printf("cos example: x=%a ours=%a libm=%a\n", worst[4][0], inflx_gl_cos(worst[4][0]), cos(worst[4][0]));
#endif
#ifdef INFLX_GL_HAVE_TANH
  if (bad_expm1) printf("expm1 example: x=%a ours=%a libm=%a\n", worst[6][0], inflx_gl_expm1(worst[6][0]), expm1(worst[6][0]));
  if (bad_tanh) printf("tanh example: x=%a ours=%a libm=%a\n", worst[5][0], inflx_gl_tanh(worst[5][0]), tanh(worst[5][0]));
#endif
#ifdef INFLX_GL_HAVE_ATAN_TAN
  if (bad_atan) printf("atan example: x=%a ours=%a libm=%a\n", worst[7][0], inflx_gl_atan(worst[7][0]), atan(worst[7][0]));
  if (bad_tan) printf("tan example: x=%a ours=%a libm=%a\n", worst[8][0], inflx_gl_tan(worst[8][0]), tan(worst[8][0]));
#endif
  return 0;
}
