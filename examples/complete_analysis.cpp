// Host-side use of the C ABI (include/inflx_b200.h) from C++ - what a non-Python embedder, or the
// reference's Rust crate through its FFI stub (part B), does:
//
//   g++ -std=c++17 -Iinclude examples/complete_analysis.cpp -o /tmp/complete_analysis
//       -Linflatox_b200 -linflx_b200 -Wl,-rpath,$PWD/inflatox_b200        (one command line)
//   /tmp/complete_analysis <artefact.bin> N0 N1 x0_start x0_stop x1_start x1_stop p0 [p1 ...]
//
// The artefact is what `inflatox_b200.Compiler(model).compile()` writes
// (`CompilationArtifact.shared_object_path`).  Prints the engine's own view of the model, runs
// complete_analysis into a page-locked buffer (direct DMA) and reports min / max / NaN count per
// output plane.  Every failure is the engine's message - there is no CPU fall-back to hide it.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "inflx_b200.h"

static int die(const char* what, inflx_status st) {
  std::fprintf(stderr, "%s failed (status %d): %s\n", what, (int)st, inflx_last_error());
  return 1;
}

int main(int argc, char** argv) {
  if (argc < 9) {
    std::fprintf(stderr,
                 "usage: %s artefact.bin N0 N1 x0_start x0_stop x1_start x1_stop p0 [p1 ...]\n",
                 argv[0]);
    return 2;
  }
  const size_t n0 = std::strtoull(argv[2], nullptr, 10), n1 = std::strtoull(argv[3], nullptr, 10);
  const double start_stop[4] = {std::atof(argv[4]), std::atof(argv[5]), std::atof(argv[6]),
                                std::atof(argv[7])};
  std::vector<double> p;
  for (int i = 8; i < argc; ++i) p.push_back(std::atof(argv[i]));

  inflx_lib* lib = nullptr;
  inflx_status st = inflx_open(argv[1], /*check_basis=*/0, &lib);
  if (st != INFLX_OK) return die("inflx_open", st);
  uint16_t abi[3];
  inflx_abi_version(lib, abi);
  std::printf("model \"%s\": %u fields, %u parameters, artefact ABI %u.%u.%u\n",
              inflx_model_name(lib), inflx_n_fields(lib), inflx_n_parameters(lib), abi[0], abi[1],
              abi[2]);

  void* buf = nullptr;
  const size_t bytes = n0 * n1 * 6 * sizeof(double);
  st = inflx_host_alloc(bytes, &buf);  // page-locked: the GPUs write it by DMA
  if (st != INFLX_OK) {
    inflx_close(lib);
    return die("inflx_host_alloc", st);
  }
  double* out = static_cast<double*>(buf);
  st = inflx_complete_analysis(lib, p.data(), p.size(), out, n0, n1, 6, start_stop, 2, 2,
                               /*progress=*/0, /*threads=*/0);
  int rc = 0;
  if (st != INFLX_OK) {
    rc = die("inflx_complete_analysis", st);
  } else {
    static const char* names[6] = {"consistency", "eps_V", "eps_H", "eta", "delta", "omega"};
    for (int k = 0; k < 6; ++k) {
      double lo = INFINITY, hi = -INFINITY;
      size_t nans = 0;
      for (size_t i = 0; i < n0 * n1; ++i) {
        const double v = out[i * 6 + k];
        if (std::isnan(v)) {
          ++nans;
        } else {
          lo = std::fmin(lo, v);
          hi = std::fmax(hi, v);
        }
      }
      std::printf("%-12s min % .6e  max % .6e  NaN %zu\n", names[k], lo, hi, nans);
    }
  }
  inflx_host_free(buf);
  inflx_close(lib);
  return rc;
}
