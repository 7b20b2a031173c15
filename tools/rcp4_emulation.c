/* CPU emulation of the speculative division of csrc/inflx_device.cuh with the default (5-step) and
 * the experimental 4-step reciprocal refinement (INFLX_EXPERIMENT_RCP4), for a MODELLED seed:
 * y0 = (1/b)(1 + d), d uniform in +-2^-k, kept to the high 32 bits of the double with the low word
 * set to 1 (what inflx_mufu_rcp64h builds from MUFU.RCP64H).  Everything after the seed is fma /
 * multiply, i.e. exactly reproducible on the host.  Counts quotients that differ from a / b.
 *
 *   gcc -O2 -march=native -ffp-contract=off -fopenmp tools/rcp4_emulation.c -o /tmp/rcp4 -lm
 *   /tmp/rcp4 <log2 pairs> <k>
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

static inline uint64_t rng(uint64_t* s) {  /* xorshift64* */
  uint64_t x = *s;
  x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
  *s = x;
  return x * 0x2545F4914F6CDD1Dull;
}
static inline double u01(uint64_t* s) { return (double)(rng(s) >> 11) * 0x1p-53; }
static inline double seed_of(double b, double d) {
  double y = (1.0 / b) * (1.0 + d);
  uint64_t u;
  memcpy(&u, &y, 8);
  u = (u & 0xffffffff00000000ull) | 1ull;
  memcpy(&y, &u, 8);
  return y;
}
static inline double quot(double a, double b, double y) {
  const double q0 = a * y;
  const double r = fma(q0, -b, a);
  return fma(y, r, q0);
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 28, k = argc > 2 ? atoi(argv[2]) : 20;
  const uint64_t n = 1ull << lg;
  uint64_t bad5 = 0, bad4 = 0, rcp_mis5 = 0, rcp_mis4 = 0;
#pragma omp parallel reduction(+ : bad5, bad4, rcp_mis5, rcp_mis4)
  {
    uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(1 + omp_get_thread_num());
#pragma omp for schedule(static)
    for (uint64_t i = 0; i < n; i++) {
      const double a = ldexp(1.0 + u01(&s), (int)(rng(&s) % 40) - 20);
      const double b = ldexp(1.0 + u01(&s), (int)(rng(&s) % 40) - 20);
      const double d = ldexp(2.0 * u01(&s) - 1.0, -k);
      const double y0 = seed_of(b, d);
      /* default: e, e + e^2, y1, e2, y2 */
      double e = fma(y0, -b, 1.0);
      const double ec = fma(e, e, e);
      double y1 = fma(y0, ec, y0);
      double e2 = fma(y1, -b, 1.0);
      const double y5 = fma(y1, e2, y1);
      /* experiment: without the cubic step */
      y1 = fma(y0, e, y0);
      e2 = fma(y1, -b, 1.0);
      const double y4 = fma(y1, e2, y1);
      const double q = a / b, yt = 1.0 / b;
      bad5 += quot(a, b, y5) != q;
      bad4 += quot(a, b, y4) != q;
      rcp_mis5 += y5 != yt;
      rcp_mis4 += y4 != yt;
    }
  }
  printf("pairs 2^%d, seed error <= 2^-%d: reciprocal != 1/b: 5-step %llu, 4-step %llu; "
         "quotient != a/b: 5-step %llu, 4-step %llu\n",
         lg, k, (unsigned long long)rcp_mis5, (unsigned long long)rcp_mis4,
         (unsigned long long)bad5, (unsigned long long)bad4);
  return 0;
}
