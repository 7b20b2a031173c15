/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * CPU restatement of inflatox's grid-evaluation hot path, used ONLY as the parity checker
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).  Nothing
 * under inflatox_b200/ may include, link, dlopen or execute this file.
 *
 * What it restates (all citations relative to /root/reference):
 *   - src/dylib.rs:67-161, 163-183   open the model artefact, ABI check, symbol names v<row><col>
 *   - src/hesse_bindings.rs:195-232  Hesse2D: fns[0]=v00, fns[2]=v10, fns[3]=v11
 *   - src/anguelova.rs:84-94         convert_ranges (spacing = (stop-start)/N, endpoint excluded)
 *   - src/anguelova.rs:99-171        mod ops (operation order kept verbatim)
 *   - src/anguelova.rs:219-251,508-540  flat index -> (row, col) -> field-space point
 *   - src/anguelova.rs:633-977       on-trajectory variants
 *   - src/hesse_bindings.rs:68-85, 150-192  potential_array / hesse_array index maps
 *
 * The model artefact is the reference's OWN generated C (tests/golden/c/<model>.c.gz, produced by
 * running the unmodified reference compiler.py, see tests/golden/make_golden.py) built with the
 * reference's flag set (compiler.py:299-310) by oracle/build.py.  Rust's f64 arithmetic never
 * contracts a*b+c into an fma, so this file must be compiled with -ffp-contract=off and without
 * any -ffast-math style flag; `powi(2)` is x*x, `.recip()` is 1.0/x, `.abs()` is fabs,
 * atan/tan/sqrt are the platform libm's, exactly as rustc lowers them on x86-64 linux.
 *
 * Parity pinning: tests/test_oracle.py checks this file against every known answer the
 * reference's own tests hold for the path (tests/test_doc.py:50,51,58).  The Rust crate itself
 * cannot be built in this image (no rustc/cargo), so grids are pinned by construction (verbatim
 * operation order) plus those known answers - see DESIGN.md "Oracle".
 *
 * The same source is compiled a second time with -DINFLX_QUAD (REAL = __float128, libquadmath)
 * against a quad-precision transliteration of the same generated C: that build is the "truth"
 * used to measure conditioning, never a parity target by itself.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#ifdef INFLX_QUAD
#include <quadmath.h>
typedef __float128 REAL;
#define R_FABS fabsq
#define R_ATAN atanq
#define R_TAN tanq
#define R_SQRT sqrtq
#define SYM(name) name##_quad
#else
typedef double REAL;
#define R_FABS fabs
#define R_ATAN atan
#define R_TAN tan
#define R_SQRT sqrt
#define SYM(name) name
#endif

typedef REAL (*scalar_fn)(const REAL *x, const REAL *args);
typedef void (*vector_fn)(const REAL *x, const REAL *args, REAL *out);
typedef REAL (*inner_fn)(const REAL *x, const REAL *args, const REAL *v1, const REAL *v2);

typedef struct {
  void *dl;
  uint32_t dim, n_par;
  uint16_t version[3];
  char name[128];
  scalar_fn V, grad2;
  scalar_fn hesse[4]; /* row-major: v00 v01 v10 v11 (dylib.rs:163-183) */
  vector_fn basis[2]; /* "v", "w1" */
  inner_fn inner;
} oracle_model;

#define MAX_PAR 64

/* ---------------------------------------------------------------------------------------- */
/* dylib.rs:67-161                                                                          */
/* ---------------------------------------------------------------------------------------- */
int SYM(oracle_open)(const char *path, oracle_model **out) {
  void *dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!dl) {
    fprintf(stderr, "oracle_open: %s\n", dlerror());
    return 1;
  }
  oracle_model *m = calloc(1, sizeof *m);
  m->dl = dl;
  const uint16_t *ver = dlsym(dl, "VERSION");
  const uint32_t *dim = dlsym(dl, "DIM");
  const uint32_t *npar = dlsym(dl, "N_PARAMETERS");
  char *const *name = dlsym(dl, "MODEL_NAME");
  if (!ver || !dim || !npar || !name) return 2;
  /* inflatox_version.rs:48-53: only major and minor take part in the comparison; ABI 5.0 */
  if (ver[0] != 5 || ver[1] != 0) return 3;
  memcpy(m->version, ver, sizeof m->version);
  m->dim = *dim;
  m->n_par = *npar;
  if (m->n_par > MAX_PAR) return 4;
  snprintf(m->name, sizeof m->name, "%s", *name);
  m->V = (scalar_fn)dlsym(dl, "V");
  m->grad2 = (scalar_fn)dlsym(dl, "grad_norm_squared");
  if (!m->V || !m->grad2) return 2;
  if (m->dim == 2) {
    const char *names[4] = {"v00", "v01", "v10", "v11"};
    for (int i = 0; i < 4; ++i) {
      m->hesse[i] = (scalar_fn)dlsym(dl, names[i]);
      if (!m->hesse[i]) return 2;
    }
    m->basis[0] = (vector_fn)dlsym(dl, "v");
    m->basis[1] = (vector_fn)dlsym(dl, "w1");
    m->inner = (inner_fn)dlsym(dl, "inner_prod");
  }
  *out = m;
  return 0;
}

void SYM(oracle_close)(oracle_model *m) {
  if (!m) return;
  dlclose(m->dl);
  free(m);
}

uint32_t SYM(oracle_n_fields)(const oracle_model *m) { return m->dim; }
uint32_t SYM(oracle_n_params)(const oracle_model *m) { return m->n_par; }
const char *SYM(oracle_name)(const oracle_model *m) { return m->name; }

static void load_params(const oracle_model *m, const double *p, REAL *pr) {
  for (uint32_t i = 0; i < m->n_par; ++i) pr[i] = (REAL)p[i];
}

/* ---------------------------------------------------------------------------------------- */
/* anguelova.rs:99-171  mod ops                                                             */
/* ---------------------------------------------------------------------------------------- */
static inline void op_complete_analysis(const oracle_model *m, const REAL *x, const REAL *p,
                                        REAL *val) {
  /* anguelova.rs:110 - evaluation order V, v11, v10, v00 (pure functions, order immaterial) */
  const REAL v = m->V(x, p), v11 = m->hesse[3](x, p), v10 = m->hesse[2](x, p),
             v00 = m->hesse[0](x, p);
  /* :112-117 */
  const REAL lhs = v11 / v;
  const REAL q1 = v00 / v10, q2 = v10 / v00;
  const REAL rhs = (REAL)3. + (REAL)3. * (q1 * q1) + (v00 / v) * (q2 * q2);
  const REAL consistency = R_FABS(lhs - rhs) / (R_FABS(lhs) + R_FABS(rhs));
  /* :119 (no factor 1/2 here, unlike epsilon_v_only) */
  const REAL epsilon_v = m->grad2(x, p) / (v * v);
  /* :121-122 */
  const REAL vtt = (v00 * (v10 * v10) + v11 * (v00 * v00) - (REAL)2. * v00 * (v10 * v10)) /
                   (v00 * v00 + v10 * v10);
  /* :124 */
  const REAL q3 = v00 / v10;
  const REAL vt2 = epsilon_v * ((REAL)1. / ((REAL)1. + q3 * q3));
  /* :126 */
  const REAL epsilon_h =
      (REAL)3. * (epsilon_v - vt2) * ((REAL)1. / (epsilon_v + R_FABS(vtt) / v - vt2));
  /* :128 */
  const REAL delta = R_ATAN(R_FABS(v10 / v00));
  /* :130 */
  const REAL omega = R_SQRT((vtt / v) * ((REAL)3. - epsilon_h));
  /* :132 */
  const REAL eta_parallel = omega * R_TAN(delta) - (REAL)3.;
  /* :134 */
  val[0] = consistency;
  val[1] = epsilon_v;
  val[2] = epsilon_h;
  val[3] = eta_parallel;
  val[4] = delta;
  val[5] = omega;
}

static inline REAL op_epsilon_v_only(const oracle_model *m, const REAL *x, const REAL *p) {
  /* :139  0.5 * g2 / V^2, left to right */
  const REAL v = m->V(x, p);
  return (REAL)0.5 * m->grad2(x, p) / (v * v);
}

static inline REAL op_consistency_rapidturn_only(const oracle_model *m, const REAL *x,
                                                 const REAL *p) {
  /* :149-153 */
  const REAL v = m->V(x, p), v11 = m->hesse[3](x, p), v10 = m->hesse[2](x, p),
             v00 = m->hesse[0](x, p);
  const REAL lhs = v11 / v;
  const REAL q = v10 / v00;
  const REAL rhs = (REAL)3. * (q * q);
  return R_FABS(R_FABS(lhs) - R_FABS(rhs)) / (R_FABS(lhs) + R_FABS(rhs));
}

static inline REAL op_consistency_only(const oracle_model *m, const REAL *x, const REAL *p) {
  /* :158-162 */
  const REAL v = m->V(x, p), v11 = m->hesse[3](x, p), v10 = m->hesse[2](x, p),
             v00 = m->hesse[0](x, p);
  const REAL lhs = v11 / v - (REAL)3.;
  const REAL q1 = v00 / v10, q2 = v10 / v00;
  const REAL rhs = (REAL)3. * (q1 * q1) + (v00 / v) * (q2 * q2);
  return R_FABS(R_FABS(lhs) - R_FABS(rhs)) / (R_FABS(lhs) + R_FABS(rhs));
}

static inline uint8_t op_flag_quantum_diff(const oracle_model *m, const REAL *x, const REAL *p,
                                           double accuracy) {
  /* :166-170 - basis function "v" (hesse_bindings.rs:42-43), signed compare, all components */
  REAL out[2] = {0, 0};
  m->basis[0](x, p, out);
  return (out[0] <= (REAL)accuracy) && (out[1] <= (REAL)accuracy);
}

/* ---------------------------------------------------------------------------------------- */
/* grid drivers: anguelova.rs:84-94 (ranges) and :219-251 / :508-540 (index map)            */
/* Rows [row_begin,row_end) of the N0 x N1 grid are evaluated; `out` points at row           */
/* `row_begin` (so a shard writes a contiguous slice).  Coordinates always use the GLOBAL    */
/* row index.  The coordinate arithmetic is done in double in both builds: the grid points   */
/* are inputs.                                                                               */
/* ---------------------------------------------------------------------------------------- */
typedef struct {
  double dx0, dx1, of0, of1;
} ranges;

static ranges convert_ranges(const double *ss, uint64_t n0, uint64_t n1) {
  ranges r;
  r.dx0 = (ss[1] - ss[0]) / (double)n0; /* start_stop[0] = [x0_start, x0_stop] */
  r.dx1 = (ss[3] - ss[2]) / (double)n1;
  r.of0 = ss[0];
  r.of1 = ss[2];
  return r;
}

static inline void grid_point(const ranges *r, uint64_t idx, uint64_t n1, REAL *x) {
  const double i0 = (double)(idx / n1), i1 = (double)(idx % n1);
  const double a = i0 * r->dx0, b = i1 * r->dx1; /* mul then add, never fused */
  x[0] = (REAL)(a + r->of0);
  x[1] = (REAL)(b + r->of1);
}

#define GRID_LOOP(BODY)                                                                  \
  REAL pr[MAX_PAR];                                                                      \
  load_params(m, p, pr);                                                                 \
  const ranges r = convert_ranges(start_stop, n0, n1);                                   \
  const int64_t first = (int64_t)(row_begin * n1), last = (int64_t)(row_end * n1);       \
  _Pragma("omp parallel for schedule(dynamic, 4096) num_threads(threads)") for (         \
      int64_t idx = first; idx < last; ++idx) {                                          \
    REAL x[2];                                                                           \
    grid_point(&r, (uint64_t)idx, n1, x);                                                \
    const int64_t o = idx - first;                                                       \
    BODY                                                                                 \
  }

static int fix_threads(int threads) {
  if (threads > 0) return threads;
  const char *e = getenv("OMP_NUM_THREADS");
  if (e && atoi(e) > 0) return atoi(e);
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

void SYM(oracle_complete_analysis)(const oracle_model *m, const double *p, double *out,
                                   uint64_t n0, uint64_t n1, const double *start_stop,
                                   uint64_t row_begin, uint64_t row_end, int threads) {
  threads = fix_threads(threads);
  GRID_LOOP({
    REAL val[6];
    op_complete_analysis(m, x, pr, val);
    for (int k = 0; k < 6; ++k) out[o * 6 + k] = (double)val[k];
  })
}

void SYM(oracle_consistency_only)(const oracle_model *m, const double *p, double *out,
                                  uint64_t n0, uint64_t n1, const double *start_stop,
                                  uint64_t row_begin, uint64_t row_end, int threads) {
  threads = fix_threads(threads);
  GRID_LOOP({ out[o] = (double)op_consistency_only(m, x, pr); })
}

void SYM(oracle_consistency_rapidturn_only)(const oracle_model *m, const double *p, double *out,
                                            uint64_t n0, uint64_t n1, const double *start_stop,
                                            uint64_t row_begin, uint64_t row_end, int threads) {
  threads = fix_threads(threads);
  GRID_LOOP({ out[o] = (double)op_consistency_rapidturn_only(m, x, pr); })
}

void SYM(oracle_epsilon_v_only)(const oracle_model *m, const double *p, double *out, uint64_t n0,
                                uint64_t n1, const double *start_stop, uint64_t row_begin,
                                uint64_t row_end, int threads) {
  threads = fix_threads(threads);
  GRID_LOOP({ out[o] = (double)op_epsilon_v_only(m, x, pr); })
}

void SYM(oracle_flag_quantum_dif)(const oracle_model *m, const double *p, uint8_t *out,
                                  uint64_t n0, uint64_t n1, const double *start_stop,
                                  uint64_t row_begin, uint64_t row_end, double accuracy,
                                  int threads) {
  threads = fix_threads(threads);
  GRID_LOOP({ out[o] = op_flag_quantum_diff(m, x, pr, accuracy); })
}

/* the five model functions themselves on a grid: [V, v00, v10, v11, g2] per point.  Not a
 * reference entry point; lets tests localise a mismatch to the model functions or the ops. */
void SYM(oracle_model_functions)(const oracle_model *m, const double *p, double *out, uint64_t n0,
                                 uint64_t n1, const double *start_stop, uint64_t row_begin,
                                 uint64_t row_end, int threads) {
  threads = fix_threads(threads);
  GRID_LOOP({
    out[o * 5 + 0] = (double)m->V(x, pr);
    out[o * 5 + 1] = (double)m->hesse[0](x, pr);
    out[o * 5 + 2] = (double)m->hesse[2](x, pr);
    out[o * 5 + 3] = (double)m->hesse[3](x, pr);
    out[o * 5 + 4] = (double)m->grad2(x, pr);
  })
}

/* ---------------------------------------------------------------------------------------- */
/* on-trajectory variants: anguelova.rs:633-977 (x is (n,2) row-major, read as given)       */
/* ---------------------------------------------------------------------------------------- */
#define TRAJ_LOOP(BODY)                                                                  \
  REAL pr[MAX_PAR];                                                                      \
  load_params(m, p, pr);                                                                 \
  _Pragma("omp parallel for schedule(static) num_threads(threads)") for (int64_t o = 0;  \
                                                                         o < (int64_t)n; \
                                                                         ++o) {          \
    REAL x[2] = {(REAL)xs[2 * o], (REAL)xs[2 * o + 1]};                                  \
    BODY                                                                                 \
  }

void SYM(oracle_complete_analysis_on_trajectory)(const oracle_model *m, const double *p,
                                                 const double *xs, double *out, uint64_t n,
                                                 int threads) {
  threads = fix_threads(threads);
  TRAJ_LOOP({
    REAL val[6];
    op_complete_analysis(m, x, pr, val);
    for (int k = 0; k < 6; ++k) out[o * 6 + k] = (double)val[k];
  })
}

void SYM(oracle_consistency_only_on_trajectory)(const oracle_model *m, const double *p,
                                                const double *xs, double *out, uint64_t n,
                                                int threads) {
  threads = fix_threads(threads);
  TRAJ_LOOP({ out[o] = (double)op_consistency_only(m, x, pr); })
}

void SYM(oracle_consistency_rapidturn_only_on_trajectory)(const oracle_model *m, const double *p,
                                                          const double *xs, double *out,
                                                          uint64_t n, int threads) {
  threads = fix_threads(threads);
  TRAJ_LOOP({ out[o] = (double)op_consistency_rapidturn_only(m, x, pr); })
}

void SYM(oracle_epsilon_v_only_on_trajectory)(const oracle_model *m, const double *p,
                                              const double *xs, double *out, uint64_t n,
                                              int threads) {
  threads = fix_threads(threads);
  TRAJ_LOOP({ out[o] = (double)op_epsilon_v_only(m, x, pr); })
}

/* ---------------------------------------------------------------------------------------- */
/* scalar entry points: lib.rs:309-339 (potential), :384-419 + hesse_bindings.rs:132-136     */
/* ---------------------------------------------------------------------------------------- */
double SYM(oracle_potential)(const oracle_model *m, const double *x, const double *p) {
  REAL pr[MAX_PAR], xr[2] = {(REAL)x[0], (REAL)x[1]};
  load_params(m, p, pr);
  return (double)m->V(xr, pr);
}

void SYM(oracle_hesse)(const oracle_model *m, const double *x, const double *p, double *out4) {
  REAL pr[MAX_PAR], xr[2] = {(REAL)x[0], (REAL)x[1]};
  load_params(m, p, pr);
  for (int i = 0; i < 4; ++i) out4[i] = (double)m->hesse[i](xr, pr);
}

double SYM(oracle_grad_norm_squared)(const oracle_model *m, const double *x, const double *p) {
  REAL pr[MAX_PAR], xr[2] = {(REAL)x[0], (REAL)x[1]};
  load_params(m, p, pr);
  return (double)m->grad2(xr, pr);
}

/* basis vector `which` (0 = v, 1 = w1) and the metric inner product (lib.rs:142-203) */
void SYM(oracle_basis)(const oracle_model *m, int which, const double *x, const double *p,
                       double *out2) {
  REAL pr[MAX_PAR], xr[2] = {(REAL)x[0], (REAL)x[1]}, o[2] = {0, 0};
  load_params(m, p, pr);
  m->basis[which](xr, pr, o);
  out2[0] = (double)o[0];
  out2[1] = (double)o[1];
}

double SYM(oracle_inner_prod)(const oracle_model *m, const double *x, const double *p,
                              const double *v1, const double *v2) {
  REAL pr[MAX_PAR], xr[2] = {(REAL)x[0], (REAL)x[1]};
  REAL a[2] = {(REAL)v1[0], (REAL)v1[1]}, b[2] = {(REAL)v2[0], (REAL)v2[1]};
  load_params(m, p, pr);
  return (double)m->inner(xr, pr, a, b);
}

/* hesse_bindings.rs:68-85: serial N-d fill, x_k = idx_k * spacing_k + start_k (2-field case) */
void SYM(oracle_potential_array)(const oracle_model *m, const double *p, double *out, uint64_t n0,
                                 uint64_t n1, const double *start_stop) {
  REAL pr[MAX_PAR];
  load_params(m, p, pr);
  const ranges r = convert_ranges(start_stop, n0, n1);
  for (uint64_t idx = 0; idx < n0 * n1; ++idx) {
    REAL x[2];
    grid_point(&r, idx, n1, x);
    out[idx] = (double)m->V(x, pr);
  }
}

/* hesse_bindings.rs:150-192: output shape (2,2,N0,N1), component-major */
void SYM(oracle_hesse_array)(const oracle_model *m, const double *p, double *out, uint64_t n0,
                             uint64_t n1, const double *start_stop) {
  REAL pr[MAX_PAR];
  load_params(m, p, pr);
  const ranges r = convert_ranges(start_stop, n0, n1);
  for (int c = 0; c < 4; ++c)
    for (uint64_t idx = 0; idx < n0 * n1; ++idx) {
      REAL x[2];
      grid_point(&r, idx, n1, x);
      out[(uint64_t)c * n0 * n1 + idx] = (double)m->hesse[c](x, pr);
    }
}
