// TEST INFRASTRUCTURE - compiles the *generated* CUDA translation unit (device header + model
// kernels) for the HOST, so that the generator's output can be executed and compared with the
// oracle on a machine without a GPU (tests/test_host_emulation.py).
//
// One emulated thread per CTA (INFLX_BLOCK = 1): the cooperative shared-memory staging loop then
// copies the CTA's whole row block and __syncthreads() is a no-op.  blockIdx / threadIdx are
// thread-local variables set by the launcher in emulate.cpp; __constant__ memory is a plain array.
// The hardware seeds become IEEE values (INFLX_HOST_EMULATION in inflx_device.cuh), everything
// else is the same source text.  Build with -ffp-contract=off: `a * b + c` must stay two roundings.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#define INFLX_HOST_EMULATION 1
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __constant__ static
#define __shared__ static thread_local
#define __launch_bounds__(...)

struct inflx_emu_dim3 {
  unsigned x, y, z;
};
static thread_local inflx_emu_dim3 blockIdx, threadIdx, blockDim, gridDim;
static inline void __syncthreads() {}
static inline void __syncwarp() {}

struct alignas(16) double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) {
  double2 r;
  r.x = x;
  r.y = y;
  return r;
}
template <class T>
static inline T __ldg(const T* p) {
  return *p;
}

static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline int __double2hiint(double x) {
  uint64_t u;
  memcpy(&u, &x, 8);
  return (int)(u >> 32);
}
static inline int __double2loint(double x) {
  uint64_t u;
  memcpy(&u, &x, 8);
  return (int)(u & 0xffffffffu);
}
static inline double __hiloint2double(int hi, int lo) {
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double x;
  memcpy(&x, &u, 8);
  return x;
}
static inline long long __double_as_longlong(double x) {
  long long u;
  memcpy(&u, &x, 8);
  return u;
}
static inline double __longlong_as_double(long long u) {
  double x;
  memcpy(&x, &u, 8);
  return x;
}
static inline float __int_as_float(int i) {
  float f;
  memcpy(&f, &i, 4);
  return f;
}
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned __activemask() { return 1u; }
static inline int __any_sync(unsigned, int pred) { return pred; }
