// inflx_cuda_dl.h - lazily bound CUDA driver API + NVRTC.
//
// The library is linked against neither libcuda nor libnvrtc: both are dlopen'ed at first use so
// that the C ABI loads (and `inflx_open` can read an artefact's metadata) on a machine without a
// GPU, and so that a missing driver is reported as INFLX_ERR_CUDA instead of a loader failure.
// This file plays the role of reference src/dylib.rs' libloading layer, pointed at the driver
// instead of at a generated C dylib.
#pragma once
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <mutex>
#include <string>
#include <vector>

#define INFLX_STR(x) #x
#define INFLX_XSTR(x) INFLX_STR(x)

#define INFLX_CU_FUNCS(X)                                                                       \
  X(cuInit) X(cuDriverGetVersion) X(cuDeviceGetCount) X(cuDeviceGet) X(cuDeviceGetName)         \
  X(cuDeviceGetAttribute) X(cuDevicePrimaryCtxRetain) X(cuDevicePrimaryCtxRelease)              \
  X(cuCtxSetCurrent) X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction)               \
  X(cuModuleGetGlobal) X(cuMemAlloc) X(cuMemFree) X(cuMemcpyHtoDAsync) X(cuMemcpyDtoHAsync)     \
  X(cuMemcpyDtoDAsync) X(cuMemHostAlloc) X(cuMemFreeHost) X(cuStreamCreate)                     \
  X(cuStreamSynchronize) X(cuStreamDestroy) X(cuStreamWaitEvent) X(cuEventCreate)               \
  X(cuEventRecord) X(cuEventSynchronize) X(cuEventElapsedTime) X(cuEventDestroy)                \
  X(cuLaunchKernel) X(cuGetErrorString) X(cuGetErrorName) X(cuPointerGetAttribute)              \
  X(cuMemGetInfo) X(cuMemHostRegister) X(cuMemHostUnregister) X(cuMemHostGetDevicePointer)    \
  X(cuCtxGetCurrent) X(cuDeviceGetPCIBusId) X(cuOccupancyMaxActiveBlocksPerMultiprocessor)

#define INFLX_NVRTC_FUNCS(X)                                                                    \
  X(nvrtcCreateProgram) X(nvrtcDestroyProgram) X(nvrtcCompileProgram) X(nvrtcGetProgramLogSize) \
  X(nvrtcGetProgramLog) X(nvrtcGetCUBINSize) X(nvrtcGetCUBIN) X(nvrtcGetErrorString)            \
  X(nvrtcVersion)

namespace inflx {

struct CudaDriver {
#define X(name) decltype(&name) p_##name = nullptr;
  INFLX_CU_FUNCS(X)
#undef X
  void* handle = nullptr;
  bool ok = false;
  std::string error;

  static CudaDriver& get() {
    static CudaDriver d;
    static std::once_flag once;
    std::call_once(once, [] { d.load(); });
    return d;
  }

 private:
  void load() {
    const char* names[] = {"libcuda.so.1", "libcuda.so"};
    for (const char* n : names) {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      error = std::string("no CUDA driver: ") + dlerror();
      return;
    }
#define X(name)                                                                  \
  p_##name = reinterpret_cast<decltype(&name)>(dlsym(handle, INFLX_XSTR(name))); \
  if (!p_##name) {                                                               \
    error = std::string("CUDA driver lacks ") + INFLX_XSTR(name);                \
    return;                                                                      \
  }
    INFLX_CU_FUNCS(X)
#undef X
    CUresult r = p_cuInit(0);
    if (r != CUDA_SUCCESS) {
      const char* s = nullptr;
      p_cuGetErrorString(r, &s);
      error = std::string("cuInit failed: ") + (s ? s : "unknown error");
      return;
    }
    ok = true;
  }
};

struct Nvrtc {
#define X(name) decltype(&name) p_##name = nullptr;
  INFLX_NVRTC_FUNCS(X)
#undef X
  void* handle = nullptr;
  bool ok = false;
  std::string error;

  static Nvrtc& get() {
    static Nvrtc d;
    static std::once_flag once;
    std::call_once(once, [] { d.load(); });
    return d;
  }

 private:
  void load() {
    // The CUDA toolkit's NVRTC first, BY PATH: a bare soname resolves to whatever libnvrtc.so.12 is
    // already mapped into the process, and `import torch` maps its own bundled (older) one - the
    // cubin cache would then be keyed, and the kernels compiled, by a different compiler depending
    // on import order (seen on the GPU box: bench.py imports torch, the tests do not).
    std::vector<std::string> names;
    if (const char* home = getenv("CUDA_HOME")) {
      names.push_back(std::string(home) + "/lib64/libnvrtc.so.12");
      names.push_back(std::string(home) + "/lib64/libnvrtc.so");
    }
    for (const char* n : {"/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so",
                          "libnvrtc.so.12", "libnvrtc.so", "libnvrtc.so.13"})
      names.push_back(n);
    std::string errs;
    for (const std::string& n : names) {
      handle = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
      if (handle) break;
      errs += std::string(dlerror()) + "; ";
    }
    if (!handle) {
      error = "NVRTC not found: " + errs;
      return;
    }
#define X(name)                                                                  \
  p_##name = reinterpret_cast<decltype(&name)>(dlsym(handle, INFLX_XSTR(name))); \
  if (!p_##name) {                                                               \
    error = std::string("NVRTC lacks ") + INFLX_XSTR(name);                      \
    return;                                                                      \
  }
    INFLX_NVRTC_FUNCS(X)
#undef X
    ok = true;
  }
};

}  // namespace inflx
