/* Host-side check of inflatox_b200/csrc/inflx_crmath.cuh against libquadmath (TEST CODE).
 * Prints one line of counters; tests/test_crmath.py parses it. */
#include <quadmath.h>
#include <stdio.h>
#include <stdlib.h>
#include INFLX_CRMATH_HEADER

static double rnd(void) { return (double)rand() / RAND_MAX; }

int main(int argc, char** argv) {
  long n = argc > 1 ? atol(argv[1]) : 200000;
  srand(12345);
  long bad_pow = 0, bad_log = 0, bad_exp = 0, bad_sin = 0, bad_cos = 0, glibc_pow = 0, special = 0;
  for (long i = 0; i < n; i++) {
    double x, y, t;
    switch (i % 4) {
      case 0: x = 0.4 + 0.2 * rnd(); y = -3.0 * (0.1 + rnd()); break; /* EGNO: pow(x, -3 alpha) */
      case 1: x = exp(20 * (rnd() - 0.5)); y = 40 * (rnd() - 0.5); break;
      case 2: x = 1.0 + (rnd() - 0.5) * 1e-3; y = 1e3 * (rnd() - 0.5); break;
      default: x = ldexp(1 + rnd(), (int)(200 * (rnd() - 0.5))); y = 3 * (rnd() - 0.5);
    }
    t = (double)powq((__float128)x, (__float128)y);
    bad_pow += inflx_cr_pow(x, y) != t;
    glibc_pow += pow(x, y) != t;
    bad_log += inflx_cr_log(x) != (double)logq((__float128)x);
    double xe = (i % 4 == 2) ? (rnd() - 0.5) * 1e-2 : 1400 * (rnd() - 0.5);
    if (fabs(xe) < 700) bad_exp += inflx_cr_exp(xe) != (double)expq((__float128)xe);
    double xs;
    switch (i % 5) {
      case 0: xs = 4 * M_PI * rnd(); break;                                   /* d5: [0, 4 pi] */
      case 1: xs = 2000 * (rnd() - 0.5); break;
      case 2: xs = ldexp(rnd() - 0.5, -(int)(40 * rnd())); break;
      case 3: xs = (double)(long)(1000 * rnd()) * M_PI_2 * (1 + (rnd() - 0.5) * 1e-9); break;
      default: xs = 1e6 * (rnd() - 0.5);
    }
    bad_sin += inflx_cr_sin(xs) != (double)sinq((__float128)xs);
    bad_cos += inflx_cr_cos(xs) != (double)cosq((__float128)xs);
  }
  /* irregular arguments take libm: identical results, signs and NaNs included */
  const double sp[][2] = {{0, 2}, {-1, 2}, {-8, 1.0 / 3}, {INFINITY, -1}, {NAN, 1}, {2, NAN}, {2, 0},
                          {1, 5}, {4, 0.5}, {2, 1023.5}, {2, -1074}, {1e-310, 2}, {10, 308.5},
                          {0.5, 1e10}, {-0.0, -1}, {3, INFINITY}};
  for (unsigned i = 0; i < sizeof sp / sizeof sp[0]; i++) {
    double a = inflx_cr_pow(sp[i][0], sp[i][1]), g = pow(sp[i][0], sp[i][1]);
    special += !((a == g && signbit(a) == signbit(g)) || (a != a && g != g));
  }
  const double s1[] = {0.0, -0.0, -1.0, INFINITY, -INFINITY, NAN, 1e-310, 1.0, 710.0, -750.0, 1e300};
  for (unsigned i = 0; i < sizeof s1 / sizeof s1[0]; i++) {
    double a[4] = {inflx_cr_log(s1[i]), inflx_cr_exp(s1[i]), inflx_cr_sin(s1[i]), inflx_cr_cos(s1[i])};
    double g[4] = {log(s1[i]), exp(s1[i]), sin(s1[i]), cos(s1[i])};
    for (int k = 0; k < 4; k++)
      special += !((a[k] == g[k] && signbit(a[k]) == signbit(g[k])) || (a[k] != a[k] && g[k] != g[k]));
  }
  printf("n=%ld bad_pow=%ld bad_log=%ld bad_exp=%ld bad_sin=%ld bad_cos=%ld glibc_pow=%ld special=%ld\n",
         n, bad_pow, bad_log, bad_exp, bad_sin, bad_cos, glibc_pow, special);
  return 0;
}
