/* inflx_b200.h - C ABI of the B200-native inflatox grid-evaluation back-end.
 *
 * This is the drop-in boundary for the reference's Rust extension module `inflatox.libinflx_rs`
 * (reference pyproject.toml:51-54, src/lib.rs:68-92) on the grid-evaluation hot path: every entry
 * point below names the reference interface it replaces.  Plain pointers and sizes only; all
 * array arguments are HOST memory unless stated otherwise, C-contiguous, fp64 (the reference
 * panics on non-contiguous input, src/anguelova.rs:189-191, 210-212; callers of this ABI pass
 * contiguous buffers).  Outputs are caller-allocated and filled in place, as in the reference
 * (PyReadwriteArray*, src/anguelova.rs:458-465).
 *
 * Every function returns an inflx_status; on failure inflx_last_error() holds the message the
 * reference would have formatted for the corresponding LibInflxRsErr (src/err.rs:40-61).  All
 * calls block until the result is in `out`.  A handle may be used from several host threads.
 *
 * There is no CPU fall-back anywhere behind this interface: without a CUDA driver and a GPU the
 * compute entry points fail with INFLX_ERR_CUDA.
 */
#ifndef INFLX_B200_H
#define INFLX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors LibInflxRsErr (reference src/err.rs:28-38) + the two failure classes a GPU adds. */
typedef enum {
  INFLX_OK = 0,
  INFLX_ERR_IO = 1,             /* Io             -> IOError     */
  INFLX_ERR_MISSING_SYMBOL = 2, /* MissingSymbol  -> SystemError */
  INFLX_ERR_VERSION = 3,        /* Version        -> SystemError */
  INFLX_ERR_THREADS = 4,        /* Rayon          -> SystemError */
  INFLX_ERR_SHAPE = 5,          /* Shape          -> Exception   */
  INFLX_ERR_FIELD_DIM = 6,      /* FieldDim       -> Exception   */
  INFLX_ERR_BASIS_NORM = 7,     /* BasisNorm      -> Exception   */
  INFLX_ERR_BASIS_OTH = 8,      /* BasisOth       -> Exception   */
  INFLX_ERR_CUDA = 9,           /* driver / launch failure       -> SystemError */
  INFLX_ERR_NVRTC = 10          /* kernel compilation failure    -> Exception   */
} inflx_status;

/* Operations of `mod ops` (reference src/anguelova.rs:99-171) and the array fills of
 * src/hesse_bindings.rs:68-85, 150-192. */
typedef enum {
  INFLX_OP_COMPLETE_ANALYSIS = 0,          /* 6 f64 / point */
  INFLX_OP_CONSISTENCY_ONLY = 1,           /* 1 f64 / point */
  INFLX_OP_CONSISTENCY_RAPIDTURN_ONLY = 2, /* 1 f64 / point */
  INFLX_OP_EPSILON_V_ONLY = 3,             /* 1 f64 / point */
  INFLX_OP_FLAG_QUANTUM_DIF = 4,           /* 1 byte (bool) / point */
  INFLX_OP_POTENTIAL = 5,                  /* 1 f64 / point */
  INFLX_OP_HESSE = 6,                      /* 4 f64 / point: v00 v01 v10 v11 */
  INFLX_OP_BASIS = 7                       /* points only: v0 v1 w0 w1 <v,v> <v,w> <w,w> */
} inflx_op;

/* An opened model artefact; replaces InflatoxPyDyLib / InflatoxDylib (reference src/lib.rs:104-106,
 * src/dylib.rs:32-61). */
typedef struct inflx_lib inflx_lib;

/* Thread-local message of the last failed call on this thread. */
const char *inflx_last_error(void);

/* ---- kernel compilation (used by inflatox_b200.Compiler; replaces the `zig cc` subprocess of
 * reference python/inflatox/compiler.py:568-598).  Compiles CUDA C++ `source` with NVRTC into a
 * cubin; `*cubin` and `*log` are malloc'ed and released with inflx_free.  Needs no GPU. */
inflx_status inflx_nvrtc_compile(const char *source, const char *name, const char *const *options,
                                 int n_options, void **cubin, size_t *cubin_size, char **log);
void inflx_free(void *ptr);
/* NVRTC version (part of the cubin cache key of inflatox_b200.Compiler); INFLX_ERR_NVRTC and
 * 0.0 when NVRTC cannot be loaded. */
inflx_status inflx_nvrtc_version(int *major, int *minor);

/* ---- artefact handle ------------------------------------------------------------------------ */
/* open_inflx_dylib(lib_path, check_basis) (reference src/lib.rs:108-115; loader src/dylib.rs:67-161,
 * ABI check src/inflatox_version.rs:48-53, random basis validation src/lib.rs:142-203). */
inflx_status inflx_open(const char *lib_path, int check_basis, inflx_lib **out);
void inflx_close(inflx_lib *lib);
uint32_t inflx_n_fields(const inflx_lib *lib);     /* InflatoxDylib::n_fields, dylib.rs */
uint32_t inflx_n_parameters(const inflx_lib *lib); /* InflatoxDylib::n_pars */
const char *inflx_model_name(const inflx_lib *lib);
void inflx_abi_version(const inflx_lib *lib, uint16_t out[3]);

/* Devices the handle shards grid rows / parameter vectors over.  Default: the
 * ordinals in $INFLATOX_DEVICES ("0,1,.." or "all"), else $LOCAL_RANK when set (one process per
 * GPU under torchrun), else all visible devices. */
inflx_status inflx_set_devices(inflx_lib *lib, const int *ordinals, int n);
int inflx_get_devices(const inflx_lib *lib, int *ordinals, int capacity);

/* ---- grid operations: the four grid pyfunctions + flag_quantum_dif_py ------------------------ */
/* complete_analysis(lib, p, out, start_stop, progress, threads)  (reference src/anguelova.rs:458-550)
 *   p[p_len]; out[n0][n1][n_last] with n_last == 6; start_stop[ss_rows][ss_cols] must be (2,2):
 *   [[x0_start,x0_stop],[x1_start,x1_stop]].  `threads` is accepted for signature compatibility
 *   (0 = all); the work is spread over the handle's devices instead of a rayon pool. */
inflx_status inflx_complete_analysis(inflx_lib *lib, const double *p, size_t p_len, double *out,
                                     size_t n0, size_t n1, size_t n_last,
                                     const double *start_stop, size_t ss_rows, size_t ss_cols,
                                     int progress, size_t threads);
/* consistency_only (src/anguelova.rs:176-261), consistency_rapidturn_only (:267-353),
 * epsilon_v_only (:359-447): out[n0][n1]. */
inflx_status inflx_consistency_only(inflx_lib *lib, const double *p, size_t p_len, double *out,
                                    size_t n0, size_t n1, const double *start_stop,
                                    size_t ss_rows, size_t ss_cols, int progress, size_t threads);
inflx_status inflx_consistency_rapidturn_only(inflx_lib *lib, const double *p, size_t p_len,
                                              double *out, size_t n0, size_t n1,
                                              const double *start_stop, size_t ss_rows,
                                              size_t ss_cols, int progress, size_t threads);
inflx_status inflx_epsilon_v_only(inflx_lib *lib, const double *p, size_t p_len, double *out,
                                  size_t n0, size_t n1, const double *start_stop, size_t ss_rows,
                                  size_t ss_cols, int progress, size_t threads);
/* flag_quantum_dif_py(lib, p, x, start_stop, progress, accuracy) (src/anguelova.rs:569-626):
 * x[n0][n1] of C bool (1 byte). */
inflx_status inflx_flag_quantum_dif(inflx_lib *lib, const double *p, size_t p_len, uint8_t *x,
                                    size_t n0, size_t n1, const double *start_stop,
                                    size_t ss_rows, size_t ss_cols, int progress,
                                    double accuracy);

/* ---- on-trajectory operations (src/anguelova.rs:633-977): x[n][2], out[n][6] or out[n] ------- */
inflx_status inflx_complete_analysis_on_trajectory(inflx_lib *lib, const double *p, size_t p_len,
                                                   const double *x, size_t n, size_t x_cols,
                                                   double *out, size_t out_rows, size_t out_cols,
                                                   int progress, size_t threads);
inflx_status inflx_consistency_only_on_trajectory(inflx_lib *lib, const double *p, size_t p_len,
                                                  const double *x, size_t n, size_t x_cols,
                                                  double *out, size_t out_len, int progress,
                                                  size_t threads);
inflx_status inflx_consistency_rapidturn_only_on_trajectory(inflx_lib *lib, const double *p,
                                                            size_t p_len, const double *x,
                                                            size_t n, size_t x_cols, double *out,
                                                            size_t out_len, int progress,
                                                            size_t threads);
inflx_status inflx_epsilon_v_only_on_trajectory(inflx_lib *lib, const double *p, size_t p_len,
                                                const double *x, size_t n, size_t x_cols,
                                                double *out, size_t out_len, int progress,
                                                size_t threads);

/* ---- InflatoxPyDyLib methods (src/lib.rs:205-463) -------------------------------------------- */
/* potential(x, p) -> f64 (lib.rs:309-339); hesse(x, p) -> (2,2) row-major (lib.rs:384-419). */
inflx_status inflx_potential(inflx_lib *lib, const double *x, size_t x_len, const double *p,
                             size_t p_len, double *value);
inflx_status inflx_hesse(inflx_lib *lib, const double *x, size_t x_len, const double *p,
                         size_t p_len, double *out4);
/* potential_array(x_out, p, start_stop) (lib.rs:341-382): x_out[n0][n1]. */
inflx_status inflx_potential_array(inflx_lib *lib, double *x_out, size_t n0, size_t n1,
                                   const double *p, size_t p_len, const double *start_stop,
                                   size_t ss_rows, size_t ss_cols);
/* hesse_array(nx, p, start_stop) (lib.rs:421-462): out[2][2][n0][n1], caller allocated here. */
inflx_status inflx_hesse_array(inflx_lib *lib, double *out, size_t n0, size_t n1, const double *p,
                               size_t p_len, const double *start_stop, size_t ss_rows,
                               size_t ss_cols);
/* validate_basis_on_domain(num_points, p, start_stop, accuracy) (lib.rs:207-307). */
inflx_status inflx_validate_basis_on_domain(inflx_lib *lib, const uint32_t *num_points,
                                            size_t n_axes, const double *p, size_t p_len,
                                            const double *start_stop, size_t ss_rows,
                                            size_t ss_cols, double accuracy);

/* ---- extended grid interface: row shards, fused parameter sweep, device-resident output ------ */
typedef struct {
  int op;                 /* inflx_op */
  const double *params;   /* host, [n_vectors][n_parameters] */
  uint64_t n_vectors;     /* parameter-sweep axis, fused into the launch grid (>= 1) */
  uint64_t n0, n1;        /* FULL grid shape; coordinates always use global indices */
  double start_stop[4];   /* x0_start, x0_stop, x1_start, x1_stop */
  uint64_t row_begin;     /* rows [row_begin, row_end) are evaluated */
  uint64_t row_end;
  double aux;             /* accuracy of flag_quantum_dif, else unused */
  void *out;              /* [n_vectors][row_end-row_begin][n1][k]; hesse, component-major:
                             host output [n_vectors][4][rows][n1], device-resident output
                             [4][n_vectors][rows][n1] */
  int out_is_device;      /* 1: `out` is device memory on `device` and stays there (no copy);
                             a sweep (n_vectors > 1) at most 65535 * (rows per CTA) rows per call */
  int device;             /* ordinal for out_is_device / single-device host calls; -1: shard
                             over the handle's devices */
  void *stream;           /* CUstream to launch on when out_is_device (NULL: internal stream).
                             With a stream the call returns without synchronising; later calls on
                             the same device (any stream) are ordered after it by an event, and
                             inflx_close waits for it */
} inflx_grid_request;

typedef struct {
  double kernel_ms;       /* device time of the grid kernels of this call (CUDA events, max over
                             devices) */
  double grid_ms;         /* out_is_device only: first grid-kernel launch -> last one done, i.e. the
                             dominant kernel without the parameter/row pre-passes */
  double total_ms;        /* host wall time of the call */
  uint64_t launches;      /* kernels launched by this call */
  uint64_t d2h_bytes;     /* bytes copied device -> host */
  uint64_t h2d_bytes;     /* bytes copied host -> device */
  int n_devices;
} inflx_grid_report;

inflx_status inflx_grid_eval(inflx_lib *lib, const inflx_grid_request *req,
                             inflx_grid_report *report /* may be NULL */);

/* The static sharding rule (the reference has a single shared-memory pool,
 * src/anguelova.rs:219-251): shard `index` of `count` owns rows [out[0], out[1]) and parameter
 * vectors [out[2], out[3]).  A sweep with at least `count` vectors is cut into blocks of vectors,
 * anything else into contiguous row blocks.  inflx_grid_eval applies it over a handle's devices,
 * inflatox_b200.sharding over the ranks of a one-process-per-GPU job.  Needs no GPU. */
inflx_status inflx_shard_of(uint64_t n_rows, uint64_t n_vectors, uint64_t index, uint64_t count,
                            uint64_t out[4]);

/* Point list evaluation behind the on-trajectory and scalar entry points. */
inflx_status inflx_points_eval(inflx_lib *lib, int op, const double *p, const double *xs,
                               uint64_t n, double aux, double *out);

/* ---- pinned host memory (north_star (2): "pinned host buffers"): outputs allocated here are
 * written by DMA straight from the device; any other host pointer goes through a pinned staging
 * ring + parallel memcpy.  Blocks >= 64 MiB are huge-page mappings pre-faulted in parallel and
 * registered with the driver (2-3x faster than cuMemHostAlloc, which remains the fall-back). */
inflx_status inflx_host_alloc(size_t bytes, void **ptr);
inflx_status inflx_host_free(void *ptr);
/* The same for a block that `n_devices` devices will fill with equal row shards: slice d of the
 * block is first-touched on the NUMA node device `devices[d]` is attached to, so every GPU's DMA
 * writes stay on its own socket (inflx_host_alloc = this with the process's default devices;
 * INFLATOX_NUMA=0 disables the placement). */
inflx_status inflx_host_alloc_on(size_t bytes, const int *devices, int n_devices, void **ptr);
int inflx_device_numa_node(int device);  /* -1: unknown */

/* ---- measurement support: sustained FP64 FMA rate (2 flop per DFMA) of one device, from a
 * register-resident DFMA micro-kernel; the roofline denominator MEASURED_PEAKS.json lacks. */
inflx_status inflx_measure_fp64_peak(int device, int repeats, double *tflops_best,
                                     double *tflops_median);

/* ---- introspection --------------------------------------------------------------------------- */
int inflx_device_count(void);            /* -1 when no CUDA driver is present */
uint64_t inflx_kernel_launches(void);    /* process-wide count of kernels launched by this library */
const char *inflx_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* INFLX_B200_H */
