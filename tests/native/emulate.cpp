// TEST INFRASTRUCTURE - host launcher for the generated kernels (see cuda_host_shim.h).
// Compiled together with one generated translation unit:
//   g++ -O1 -ffp-contract=off -fopenmp -shared -fPIC -DINFLX_BLOCK=1 -DINFLX_RPT=<r>
//       -DINFLX_GENERATED_CU='"<model>_<group>.cu"' tests/native/emulate.cpp -o emu.so
// and driven through ctypes.  Mirrors the launch sequence of csrc/inflx_engine.cpp run_shard() for
// one chunk: one vector -> inflx_prologue (parameters by value; P-frontier -> __constant__ bank,
// row frontier), unless `fused` is 0; a sweep (or fused == 0) -> inflx_params -> bank, inflx_rows;
// inflx_cols (when the unit has one); then the grid kernel over (column tiles) x (row tiles) x
// (vectors), with the engine's two tile heights (n_big tiles of rpt rows, the rest rpt_tail).
#include "cuda_host_shim.h"
#include INFLX_GENERATED_CU

#include <vector>

#ifndef INFLX_EMU_KERNEL  // the grid kernel pair to launch: -DINFLX_EMU_KERNEL=inflx_grid_<op> ...
#define INFLX_EMU_KERNEL inflx_grid_complete_analysis
#define INFLX_EMU_KERNEL_SWEEP inflx_grid_complete_analysis_sweep
#endif

extern "C" int emu_info(int what) {
  switch (what) {
    case 0: return INFLX_NPF;
    case 1: return INFLX_NRF;
    case 2: return INFLX_NCF;
    case 3: return INFLX_NP;
    default: return -1;
  }
}

// p: [n_vectors][INFLX_NP]; out: [n_vectors][n0][n1][per_point] (the engine's device layout);
// start_stop: x0_start, x0_stop, x1_start, x1_stop; rows [row_begin, row_end) of the n0-row grid.
extern "C" int emu_grid(const double* p, unsigned n_vectors, double* out, unsigned long long n0,
                        unsigned n1, const double* start_stop, unsigned long long row_begin,
                        unsigned long long row_end, unsigned rpt, double aux, int fused,
                        unsigned n_big, unsigned rpt_tail) {
  const double dx0 = (start_stop[1] - start_stop[0]) / (double)n0;  // inflx_engine.cpp run_shard
  const double dx1 = (start_stop[3] - start_stop[2]) / (double)n1;
  const double of0 = start_stop[0], of1 = start_stop[2];
  const unsigned n_rows = (unsigned)(row_end - row_begin);
  if ((unsigned long long)n_vectors * (INFLX_NPF ? INFLX_NPF : 1) > INFLX_PC_CAP) return 1;
  std::vector<double> rc((size_t)n_vectors * n_rows * (INFLX_NRF ? INFLX_NRF : 1) + 2);
  std::vector<double> cc((size_t)n_vectors * n1 * (INFLX_NCF ? INFLX_NCF : 1) + 2);
  const bool use_prologue = fused && n_vectors == 1;
  if (use_prologue) {
    // (1+2) the engine's single-vector path: every thread of ceil(n_rows / 128) CTAs
    inflx_pvec pv = {};
    for (int k = 0; k < INFLX_NP; ++k) pv.v[k] = p[k];
    blockDim = {128, 1, 1};
    const unsigned blocks = n_rows ? (n_rows + 127) / 128 : 1;
    for (unsigned i = 0; i < blocks * 128; ++i) {
      blockIdx = {i / 128, 0, 0};
      threadIdx = {i % 128, 0, 0};
      inflx_prologue(pv, inflx_pc, rc.data(), of0, dx0, row_begin, n_rows);
    }
  } else {
    // (1) parameter block, straight into the constant bank
    blockDim = {64, 1, 1};
    for (unsigned s = 0; s < n_vectors; ++s) {
      blockIdx = {s / 64, 0, 0};
      threadIdx = {s % 64, 0, 0};
      inflx_params(p, inflx_pc, n_vectors);
    }
  }
  // (2) pre-passes
  blockDim = {128, 1, 1};
  for (unsigned s = 0; s < n_vectors; ++s) {
    for (unsigned i = 0; i < n_rows && !use_prologue; ++i) {
      blockIdx = {i / 128, s, 0};
      threadIdx = {i % 128, 0, 0};
      inflx_rows(rc.data(), of0, dx0, row_begin, n_rows);
    }
#if INFLX_NCF > 0
    for (unsigned c = 0; c < n1; ++c) {
      blockIdx = {c / 128, s, 0};
      threadIdx = {c % 128, 0, 0};
      inflx_cols(cc.data(), of1, dx1, n1);
    }
#endif
  }
  // (3) grid kernel, one emulated thread per CTA (INFLX_BLOCK == 1)
  // two tile heights as in run_shard: n_big tiles of rpt rows, then tiles of rpt_tail rows
  if (rpt_tail == 0 || rpt_tail > rpt) rpt_tail = rpt;
  if ((unsigned long long)n_big * rpt > n_rows) n_big = n_rows / rpt;
  const unsigned row_tiles = n_big + (n_rows - n_big * rpt + rpt_tail - 1) / rpt_tail;
  const unsigned long long comp_stride = (unsigned long long)n_vectors * n_rows * n1;
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
  for (unsigned rt = 0; rt < row_tiles; ++rt)
    for (unsigned c = 0; c < n1; ++c)
      for (unsigned s = 0; s < n_vectors; ++s) {
        blockDim = {1, 1, 1};
        threadIdx = {0, 0, 0};
        blockIdx = {c, rt, s};
        if (n_vectors > 1)
          INFLX_EMU_KERNEL_SWEEP(out, rc.data(), of1, dx1, n1, n_rows, comp_stride, aux, rpt,
                                 cc.data(), n_big, rpt_tail);
        else
          INFLX_EMU_KERNEL(out, rc.data(), of1, dx1, n1, n_rows, comp_stride, aux, rpt, cc.data(),
                           n_big, rpt_tail);
      }
  return 0;
}
