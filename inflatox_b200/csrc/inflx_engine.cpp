// inflx_engine.cpp - host side of the B200 grid-evaluation back-end behind include/inflx_b200.h.
//
// Replaces, for the hot path, the reference's Rust extension: src/dylib.rs (artefact loader ->
// CUDA module loader), src/hesse_bindings.rs (vanishes: model functions are inlined into the
// kernels), the rayon grid drivers of src/anguelova.rs:173-550 (-> launch + row/parameter
// sharding over the GPUs of one box + pipelined device->host copies) and the checks/messages of
// src/lib.rs:117-463 and src/err.rs.  Pure driver-API code; no CPU evaluation path exists here.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <thread>
#include <vector>

#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/inflx_b200.h"
#include "inflx_cuda_dl.h"

namespace inflx {

// ---------------------------------------------------------------------------------------------
// errors and messages
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

static inflx_status fail(inflx_status code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

static std::string fmt(const char* f, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}

static std::string vec_dbg(const std::vector<size_t>& v) {  // Rust's {:?} of a Vec<usize>
  std::string s = "[";
  for (size_t i = 0; i < v.size(); ++i) s += (i ? ", " : "") + std::to_string(v[i]);
  return s + "]";
}

// LibInflxRsErr::Shape (reference src/err.rs:52)
static inflx_status shape_error(const std::vector<size_t>& expected, const std::vector<size_t>& got,
                                const std::string& msg) {
  return fail(INFLX_ERR_SHAPE, "Expected array with shape " + vec_dbg(expected) +
                                   ", received array with shape " + vec_dbg(got) +
                                   ". Context: " + msg);
}

static bool quiet() {
  const char* e = getenv("INFLATOX_QUIET");
  return e && *e && *e != '0';
}
static void info(const std::string& msg) {  // reference src/lib.rs:53-56 BADGE_INFO, to stderr
  if (!quiet()) fprintf(stderr, "\033[1;35m[Inflatox Info]\033[0m\n%s\n", msg.c_str());
}
static void warn(const std::string& msg) {  // BADGE_WARN
  fprintf(stderr, "\033[1;33m[Inflatox Warning]\033[0m\n%s\n", msg.c_str());
}

#define CU_TRY(call)                                                                       \
  do {                                                                                     \
    CUresult _r = (call);                                                                  \
    if (_r != CUDA_SUCCESS) {                                                              \
      const char *_n = nullptr, *_s = nullptr;                                             \
      cu.p_cuGetErrorName(_r, &_n);                                                        \
      cu.p_cuGetErrorString(_r, &_s);                                                      \
      return fail(INFLX_ERR_CUDA, fmt("CUDA driver call failed: %s -> %s (%s) at %s:%d",   \
                                      #call, _n ? _n : "?", _s ? _s : "?", __FILE__,       \
                                      __LINE__));                                          \
    }                                                                                      \
  } while (0)

// ---------------------------------------------------------------------------------------------
// artefact container (written by inflatox_b200/compiler.py)
// ---------------------------------------------------------------------------------------------
#pragma pack(push, 1)
struct FileHeader {
  char magic[8];  // "INFLXB2\0"
  uint32_t container_version;
  uint16_t abi[3];
  uint16_t reserved;
  uint32_t dim, n_params, n_groups, rpt, block, flags;
  char model_name[128];
  uint64_t json_offset, json_size;
  uint32_t pad;
};
struct GroupEntry {
  char name[4];
  uint32_t ncf;  // column-frontier values per column (0: the column block runs inside the grid kernel)
  uint32_t npf, nrf;
  uint64_t cubin_offset, cubin_size;
};
#pragma pack(pop)
static_assert(sizeof(FileHeader) == 192, "container header layout");
static_assert(sizeof(GroupEntry) == 32, "container group layout");

static const uint16_t kAbi[3] = {5, 0, 0};  // V_INFLX_ABI, reference src/lib.rs:50
static const uint32_t kContainerVersion = 2;  // inflatox_b200/version.py __container_version__
static const uint32_t kPcCapacity = 7680;   // doubles, must match cudagen.PC_CAPACITY

struct OpInfo {
  const char* name;
  const char* group;
  uint32_t out_bytes;  // per point
};
static const OpInfo kOps[] = {
    {"complete_analysis", "cmp", 48}, {"consistency_only", "con", 8},
    {"consistency_rapidturn_only", "con", 8}, {"epsilon_v_only", "eps", 8},
    {"flag_quantum_dif", "bas", 1}, {"potential", "pot", 8},
    {"hesse", "hes", 32}, {"basis", "bas", 56},
};

// ---------------------------------------------------------------------------------------------
// devices
// ---------------------------------------------------------------------------------------------
struct DevBuf {
  CUdeviceptr ptr = 0;
  size_t cap = 0;
};
struct PinBuf {
  void* ptr = nullptr;
  size_t cap = 0;
};

struct DeviceState {
  int ordinal = -1;
  CUdevice dev = 0;
  CUcontext ctx = nullptr;
  CUstream compute = nullptr, copy = nullptr;
  CUevent ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  CUevent ev_t0 = nullptr, ev_t1 = nullptr, ev_g0 = nullptr, ev_g1 = nullptr;
  std::mutex mu;  // one grid call at a time per device (scratch buffers are shared)
  DevBuf d_p, d_pc, d_rc, d_cc, d_xs, d_out[2];
  PinBuf stage[2], stage_in[2];
  // A call that ran on a CALLER's stream returns without synchronising; its kernels may still be
  // reading the shared scratch (d_p, d_pc, d_rc, d_cc) and the module's __constant__ bank.
  // ev_user marks its end: the next user of the scratch waits for it on its own stream.
  CUevent ev_user = nullptr;
  bool user_pending = false;
  // scalar entry points: mapped page-locked scratch (inputs then outputs), one launch per call
  void* h_scalar = nullptr;
  CUdeviceptr d_scalar = 0;
  std::string name;
  int sm_count = 148;
};

static inflx_status ensure_dev(CudaDriver& cu, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return INFLX_OK;
  if (b.ptr) cu.p_cuMemFree(b.ptr);
  b.ptr = 0;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CU_TRY(cu.p_cuMemAlloc(&b.ptr, want));
  b.cap = want;
  return INFLX_OK;
}
static inflx_status ensure_pin(CudaDriver& cu, PinBuf& b, size_t bytes) {
  if (bytes <= b.cap) return INFLX_OK;
  if (b.ptr) cu.p_cuMemFreeHost(b.ptr);
  b.ptr = nullptr;
  b.cap = 0;
  CU_TRY(cu.p_cuMemHostAlloc(&b.ptr, bytes, CU_MEMHOSTALLOC_PORTABLE));
  b.cap = bytes;
  return INFLX_OK;
}

static std::mutex g_host_mu;
static std::map<void*, size_t> g_host_maps;  // inflx_host_alloc blocks that are registered mmaps

static std::mutex g_dev_mu;
static std::map<int, std::unique_ptr<DeviceState>> g_devices;

static inflx_status get_device(int ordinal, DeviceState** out) {
  CudaDriver& cu = CudaDriver::get();
  if (!cu.ok) return fail(INFLX_ERR_CUDA, cu.error);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  auto it = g_devices.find(ordinal);
  if (it != g_devices.end()) {
    *out = it->second.get();
    return INFLX_OK;
  }
  int count = 0;
  CU_TRY(cu.p_cuDeviceGetCount(&count));
  if (ordinal < 0 || ordinal >= count)
    return fail(INFLX_ERR_CUDA, fmt("CUDA device %d requested, %d visible", ordinal, count));
  auto d = std::make_unique<DeviceState>();
  d->ordinal = ordinal;
  CU_TRY(cu.p_cuDeviceGet(&d->dev, ordinal));
  char nm[128] = {0};
  cu.p_cuDeviceGetName(nm, sizeof nm, d->dev);
  d->name = nm;
  cu.p_cuDeviceGetAttribute(&d->sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, d->dev);
  if (d->sm_count <= 0) d->sm_count = 148;
  CU_TRY(cu.p_cuDevicePrimaryCtxRetain(&d->ctx, d->dev));
  CU_TRY(cu.p_cuCtxSetCurrent(d->ctx));
  CU_TRY(cu.p_cuStreamCreate(&d->compute, CU_STREAM_NON_BLOCKING));
  CU_TRY(cu.p_cuStreamCreate(&d->copy, CU_STREAM_NON_BLOCKING));
  for (int i = 0; i < 2; ++i) {
    CU_TRY(cu.p_cuEventCreate(&d->ev_done[i], CU_EVENT_DISABLE_TIMING));
    CU_TRY(cu.p_cuEventCreate(&d->ev_copied[i], CU_EVENT_DISABLE_TIMING));
  }
  CU_TRY(cu.p_cuEventCreate(&d->ev_t0, CU_EVENT_DEFAULT));
  CU_TRY(cu.p_cuEventCreate(&d->ev_t1, CU_EVENT_DEFAULT));
  CU_TRY(cu.p_cuEventCreate(&d->ev_g0, CU_EVENT_DEFAULT));
  CU_TRY(cu.p_cuEventCreate(&d->ev_g1, CU_EVENT_DEFAULT));
  CU_TRY(cu.p_cuEventCreate(&d->ev_user, CU_EVENT_DISABLE_TIMING));
  *out = d.get();
  g_devices[ordinal] = std::move(d);
  return INFLX_OK;
}

// Order the stream `cs` (about to touch the device's shared scratch / a module's __constant__ bank)
// after a previous call that was left running on a caller's stream.  dev->mu must be held.
static inflx_status order_after_user_stream(CudaDriver& cu, DeviceState* dev, CUstream cs) {
  if (!dev->user_pending) return INFLX_OK;
  CU_TRY(cu.p_cuStreamWaitEvent(cs, dev->ev_user, 0));
  return INFLX_OK;
}
static const size_t kScalarScratch = 32u << 10;  // bytes: first half inputs, second half outputs
static const uint64_t kScalarMaxPoints = 64;

// ---------------------------------------------------------------------------------------------
// staged (pageable destination) copy-out: pinned staging buffer -> caller's array, in parallel
// and with non-temporal stores (the destination is written once and not read back here, so
// streaming stores save the read-for-ownership traffic of a plain memcpy)
// ---------------------------------------------------------------------------------------------
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void stream_copy_avx2(char* d, const char* s, size_t n) {
  size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
  if (head > n) head = n;
  memcpy(d, s, head);
  d += head;
  s += head;
  n -= head;
  const size_t blocks = n / 128;
  for (size_t i = 0; i < blocks; ++i, s += 128, d += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 64));
    const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 96), e);
  }
  _mm_sfence();
  memcpy(d, s, n - blocks * 128);
}
#endif

static void copy_out(char* d, const char* s, size_t n) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && n >= 4096) {
    stream_copy_avx2(d, s, n);
    return;
  }
#endif
  memcpy(d, s, n);
}

static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
  static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  unsigned n = (unsigned)std::min<size_t>(std::min(hw, 16u), bytes / (2u << 20));
  if (n <= 1) {
    copy_out((char*)dst, (const char*)src, bytes);
    return;
  }
  std::vector<std::thread> th;
  size_t per = ((bytes / n) + 4095) & ~size_t(4095);
  for (unsigned i = 0; i < n; ++i) {
    size_t b = (size_t)i * per;
    if (b >= bytes) break;
    size_t e = std::min(bytes, b + per);
    th.emplace_back([=] { copy_out((char*)dst + b, (const char*)src + b, e - b); });
  }
  for (auto& t : th) t.join();
}

}  // namespace inflx

using namespace inflx;

// ---------------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------------
struct GroupModule {
  CUmodule mod = nullptr;
  CUdeviceptr pc_sym = 0;
  CUfunction params = nullptr, rows = nullptr, cols = nullptr, prologue = nullptr;
  bool tail_tiles = false;  // grid kernels take (n_big, rpt_tail): two tile heights per launch
  uint32_t tail_div = 0;    // the generator's advice: tail tiles of rpt / tail_div rows (0: none)
  std::map<std::string, CUfunction> fns;
  std::map<CUfunction, int> resident;  // CTAs per SM of a grid kernel (occupancy query, cached)
};

struct inflx_lib {
  std::string path;
  std::vector<uint8_t> file;
  FileHeader hdr;
  std::vector<GroupEntry> groups;
  std::vector<int> devices;
  std::mutex mu;
  std::map<std::pair<int, std::string>, std::unique_ptr<GroupModule>> modules;  // (device, group)

  const GroupEntry* group(const char* name) const {
    for (auto& g : groups)
      if (strncmp(g.name, name, sizeof g.name) == 0) return &g;
    return nullptr;
  }
};

static inflx_status load_module(inflx_lib* lib, DeviceState* dev, const char* group,
                                GroupModule** out) {
  CudaDriver& cu = CudaDriver::get();
  std::lock_guard<std::mutex> lk(lib->mu);
  auto key = std::make_pair(dev->ordinal, std::string(group));
  auto it = lib->modules.find(key);
  if (it != lib->modules.end()) {
    *out = it->second.get();
    return INFLX_OK;
  }
  const GroupEntry* g = lib->group(group);
  if (!g)  // LibInflxRsErr::MissingSymbol (reference src/err.rs:44-50)
    return fail(INFLX_ERR_MISSING_SYMBOL,
                fmt("Could not find symbol \"%s\" in %s", group, lib->path.c_str()));
  auto gm = std::make_unique<GroupModule>();
  CU_TRY(cu.p_cuCtxSetCurrent(dev->ctx));
  CU_TRY(cu.p_cuModuleLoadData(&gm->mod, lib->file.data() + g->cubin_offset));
  size_t bytes = 0;
  CU_TRY(cu.p_cuModuleGetGlobal(&gm->pc_sym, &bytes, gm->mod, "inflx_pc"));
  CU_TRY(cu.p_cuModuleGetFunction(&gm->params, gm->mod, "inflx_params"));
  CU_TRY(cu.p_cuModuleGetFunction(&gm->rows, gm->mod, "inflx_rows"));
  if (g->ncf) CU_TRY(cu.p_cuModuleGetFunction(&gm->cols, gm->mod, "inflx_cols"));
  if (cu.p_cuModuleGetFunction(&gm->prologue, gm->mod, "inflx_prologue") != CUDA_SUCCESS)
    gm->prologue = nullptr;  // artefact of an older generator: the three-step prologue is used
  {
    CUdeviceptr marker = 0;
    size_t mbytes = 0;
    gm->tail_tiles =
        cu.p_cuModuleGetGlobal(&marker, &mbytes, gm->mod, "inflx_has_tail_tiles") == CUDA_SUCCESS;
    if (gm->tail_tiles && mbytes == sizeof(uint32_t)) {
      CU_TRY(cu.p_cuMemcpyDtoHAsync(&gm->tail_div, marker, sizeof(uint32_t), dev->compute));
      CU_TRY(cu.p_cuStreamSynchronize(dev->compute));
    }
  }
  *out = gm.get();
  lib->modules[key] = std::move(gm);
  return INFLX_OK;
}

static inflx_status get_fn(inflx_lib* lib, GroupModule* gm, const std::string& name,
                           CUfunction* out) {
  CudaDriver& cu = CudaDriver::get();
  std::lock_guard<std::mutex> lk(lib->mu);
  auto it = gm->fns.find(name);
  if (it != gm->fns.end()) {
    *out = it->second;
    return INFLX_OK;
  }
  CUfunction f = nullptr;
  CUresult r = cu.p_cuModuleGetFunction(&f, gm->mod, name.c_str());
  if (r != CUDA_SUCCESS)
    return fail(INFLX_ERR_MISSING_SYMBOL,
                fmt("Could not find symbol \"%s\" in %s", name.c_str(), lib->path.c_str()));
  gm->fns[name] = f;
  *out = f;
  return INFLX_OK;
}

static inflx_status launch(CudaDriver& cu, CUfunction f, unsigned gx, unsigned gy, unsigned gz,
                           unsigned bx, CUstream st, void** args) {
  CU_TRY(cu.p_cuLaunchKernel(f, gx, gy, gz, bx, 1, 1, 0, st, args, nullptr));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return INFLX_OK;
}

static std::vector<int> default_devices() {
  std::vector<int> out;
  CudaDriver& cu = CudaDriver::get();
  int count = 0;
  if (cu.ok) cu.p_cuDeviceGetCount(&count);
  const char* e = getenv("INFLATOX_DEVICES");
  if (e && *e && strcmp(e, "all") != 0) {
    std::stringstream ss(e);
    std::string tok;
    while (std::getline(ss, tok, ',')) out.push_back(atoi(tok.c_str()));
    return out;
  }
  const char* lr = getenv("LOCAL_RANK");
  if ((!e || !*e) && lr && *lr) {
    out.push_back(count > 0 ? atoi(lr) % count : 0);
    return out;
  }
  for (int i = 0; i < count; ++i) out.push_back(i);
  if (out.empty()) out.push_back(0);
  return out;
}

// ---------------------------------------------------------------------------------------------
// one shard (rows [rb,re) x vectors [sb,se)) on one device
// ---------------------------------------------------------------------------------------------
struct Shard {
  uint64_t rb, re, sb, se;
  int device;
};

struct ShardResult {
  inflx_status status = INFLX_OK;
  std::string error;
  double kernel_ms = 0, grid_ms = 0;
  uint64_t launches = 0, d2h = 0, h2d = 0;
};

static bool is_pinned_host(CudaDriver& cu, const void* p) {
  unsigned int mt = 0;
  CUresult r = cu.p_cuPointerGetAttribute(&mt, CU_POINTER_ATTRIBUTE_MEMORY_TYPE, (CUdeviceptr)p);
  return r == CUDA_SUCCESS && mt == CU_MEMORYTYPE_HOST;
}

static inflx_status run_shard(inflx_lib* lib, const inflx_grid_request& rq, const Shard& sh,
                              ShardResult& res) {
  CudaDriver& cu = CudaDriver::get();
  DeviceState* dev = nullptr;
  inflx_status st = get_device(sh.device, &dev);
  if (st) return st;
  std::lock_guard<std::mutex> dev_lock(dev->mu);
  CU_TRY(cu.p_cuCtxSetCurrent(dev->ctx));

  const OpInfo& op = kOps[rq.op];
  GroupModule* gm = nullptr;
  if ((st = load_module(lib, dev, op.group, &gm))) return st;
  const GroupEntry* ge = lib->group(op.group);
  const uint32_t P = lib->hdr.n_params, NPF = ge->npf, NRF = ge->nrf, NCF = ge->ncf;
  const uint32_t RPT = lib->hdr.rpt, BLOCK = lib->hdr.block;
  const uint64_t S = sh.se - sh.sb, rows_total = sh.re - sh.rb, n1 = rq.n1;
  if (S == 0 || rows_total == 0 || n1 == 0) return INFLX_OK;
  const bool sweep = rq.n_vectors > 1;
  CUfunction grid_fn = nullptr;
  if ((st = get_fn(lib, gm, std::string("inflx_grid_") + op.name + (sweep ? "_sweep" : ""),
                   &grid_fn)))
    return st;
  const bool to_device = rq.out_is_device != 0;
  CUstream cs = (to_device && rq.stream) ? (CUstream)rq.stream : dev->compute;
  if ((st = order_after_user_stream(cu, dev, cs))) return st;

  // (1) parameters -> P-frontier values of every vector of the shard.  One vector: the fused
  // prologue kernel of the first row chunk does it (parameters by value in its argument buffer,
  // no H2D copy, no separate inflx_params / inflx_rows launch); a sweep: H2D + inflx_params.
  const bool fused = S == 1 && !sweep && gm->prologue != nullptr && getenv("INFLATOX_NO_FUSED_PROLOGUE") == nullptr;
  std::vector<double> pv(std::max<uint32_t>(P, 1), 0.0);
  if (P > 0) memcpy(pv.data(), rq.params + sh.sb * P, P * 8);
  if (NPF > 0 && (st = ensure_dev(cu, dev->d_pc, S * NPF * 8))) return st;
  if (!fused) {
    if (P > 0) {
      if ((st = ensure_dev(cu, dev->d_p, S * P * 8))) return st;
      CU_TRY(cu.p_cuMemcpyHtoDAsync(dev->d_p.ptr, rq.params + sh.sb * P, S * P * 8, cs));
      res.h2d += S * P * 8;
    }
    if (NPF > 0) {
      uint32_t nv = (uint32_t)S;
      void* args[] = {&dev->d_p.ptr, &dev->d_pc.ptr, &nv};
      if ((st = launch(cu, gm->params, (unsigned)((S + 63) / 64), 1, 1, 64, cs, args))) return st;
      res.launches++;
    }
  } else {
    res.h2d += P * 8;  // the bytes travel in the launch's argument buffer
  }

  // (2) chunking: vectors first (bounded by the constant bank), then rows
  const double ranges[4] = {rq.start_stop[0], rq.start_stop[1], rq.start_stop[2], rq.start_stop[3]};
  double dx0 = (ranges[1] - ranges[0]) / (double)rq.n0;  // reference src/anguelova.rs:84-94
  double dx1 = (ranges[3] - ranges[2]) / (double)rq.n1;
  double of0 = ranges[0], of1 = ranges[2];
  const uint64_t opb = op.out_bytes;
  const uint64_t bytes_per_vec = rows_total * n1 * opb;
  uint64_t target = 64ull << 20;
  if (const char* e = getenv("INFLATOX_CHUNK_MB")) target = std::max(1, atoi(e)) * (1ull << 20);
  uint64_t s_cap = NPF ? kPcCapacity / NPF : 65535;
  s_cap = std::min<uint64_t>(std::max<uint64_t>(s_cap, 1), 65535);
  uint64_t s_chunk, rows_chunk;
  if (to_device) {
    rows_chunk = rows_total;
    s_chunk = std::min<uint64_t>(S, s_cap);
    // bound the row-frontier scratch (<= 2 GiB)
    if (NRF) s_chunk = std::max<uint64_t>(1, std::min<uint64_t>(s_chunk, (2ull << 30) / (rows_total * NRF * 8 + 1)));
  } else if (bytes_per_vec <= target) {
    rows_chunk = rows_total;
    s_chunk = std::min<uint64_t>(std::min<uint64_t>(S, s_cap), std::max<uint64_t>(1, target / bytes_per_vec));
  } else {
    s_chunk = 1;
    rows_chunk = std::max<uint64_t>(RPT, target / (n1 * opb));
    rows_chunk = std::min<uint64_t>(rows_chunk - rows_chunk % RPT, rows_total);
  }
  // gridDim.y holds 65535 row tiles: taller shards take several launches.  A device-resident
  // SWEEP cannot be cut by rows (its layout [vector][row][col] is indexed with the launch's row
  // count), a single grid can: the launches write consecutive row blocks of the caller's buffer.
  if (to_device && S > 1 && rows_total > 65535ull * RPT)
    return fail(INFLX_ERR_SHAPE, fmt("a device-resident sweep is limited to %llu rows per call "
                                     "(%llu requested): split the request by rows",
                                     (unsigned long long)(65535ull * RPT),
                                     (unsigned long long)rows_total));
  rows_chunk = std::min<uint64_t>(rows_chunk, 65535ull * RPT);

  const bool direct = !to_device && is_pinned_host(cu, rq.out);
  const uint64_t chunk_bytes = s_chunk * rows_chunk * n1 * opb;
  if (!to_device) {
    for (int i = 0; i < 2; ++i) {
      if ((st = ensure_dev(cu, dev->d_out[i], chunk_bytes))) return st;
      if (!direct && (st = ensure_pin(cu, dev->stage[i], chunk_bytes))) return st;
    }
  }
  if (NRF && (st = ensure_dev(cu, dev->d_rc, s_chunk * rows_chunk * NRF * 8))) return st;
  if (NCF && (st = ensure_dev(cu, dev->d_cc, s_chunk * NCF * n1 * 8))) return st;

  // host layout of the request: [n_vectors][rows of the REQUEST][n1][k]; hesse: [n_vectors][4][rows][n1]
  const uint64_t req_rows = rq.row_end - rq.row_begin;
  const bool hesse = rq.op == INFLX_OP_HESSE;
  struct Pending {
    bool active = false;
    uint64_t s0 = 0, sc = 0, r0 = 0, rc = 0;
  } pending[2];

  auto host_ptr = [&](uint64_t s_global, uint64_t comp, uint64_t row_global) -> char* {
    // address of (vector s, component comp (hesse only), row) in rq.out
    uint64_t row_local = row_global - rq.row_begin;
    if (hesse)
      return (char*)rq.out + (((s_global * 4 + comp) * req_rows + row_local) * n1) * 8;
    return (char*)rq.out + ((s_global * req_rows + row_local) * n1) * opb;
  };
  // device chunk layout: [sc][rc][n1][k]; hesse: [4][sc][rc][n1]
  auto issue_copies = [&](int slot, const Pending& pd) -> inflx_status {
    char* hbase = direct ? nullptr : (char*)dev->stage[slot].ptr;
    if (hesse) {
      for (uint64_t c = 0; c < 4; ++c)
        for (uint64_t s = 0; s < pd.sc; ++s) {
          uint64_t off = ((c * pd.sc + s) * pd.rc) * n1 * 8, bytes = pd.rc * n1 * 8;
          void* dst = direct ? (void*)host_ptr(sh.sb + pd.s0 + s, c, pd.r0) : (void*)(hbase + off);
          CU_TRY(cu.p_cuMemcpyDtoHAsync(dst, dev->d_out[slot].ptr + off, bytes, dev->copy));
          res.d2h += bytes;
        }
    } else if (direct && pd.rc != req_rows) {
      for (uint64_t s = 0; s < pd.sc; ++s) {
        uint64_t off = s * pd.rc * n1 * opb, bytes = pd.rc * n1 * opb;
        CU_TRY(cu.p_cuMemcpyDtoHAsync(host_ptr(sh.sb + pd.s0 + s, 0, pd.r0),
                                      dev->d_out[slot].ptr + off, bytes, dev->copy));
        res.d2h += bytes;
      }
    } else {
      uint64_t bytes = pd.sc * pd.rc * n1 * opb;
      void* dst = direct ? (void*)host_ptr(sh.sb + pd.s0, 0, pd.r0) : (void*)hbase;
      CU_TRY(cu.p_cuMemcpyDtoHAsync(dst, dev->d_out[slot].ptr, bytes, dev->copy));
      res.d2h += bytes;
    }
    return INFLX_OK;
  };
  auto finish = [&](int slot) -> inflx_status {  // wait for the copy of `slot`, unstage if needed
    Pending& pd = pending[slot];
    if (!pd.active) return INFLX_OK;
    CU_TRY(cu.p_cuEventSynchronize(dev->ev_copied[slot]));
    if (!direct) {
      const char* hbase = (const char*)dev->stage[slot].ptr;
      if (hesse) {
        for (uint64_t c = 0; c < 4; ++c)
          for (uint64_t s = 0; s < pd.sc; ++s)
            parallel_memcpy(host_ptr(sh.sb + pd.s0 + s, c, pd.r0),
                            hbase + ((c * pd.sc + s) * pd.rc) * n1 * 8, pd.rc * n1 * 8);
      } else if (pd.rc != req_rows) {
        for (uint64_t s = 0; s < pd.sc; ++s)
          parallel_memcpy(host_ptr(sh.sb + pd.s0 + s, 0, pd.r0), hbase + s * pd.rc * n1 * opb,
                          pd.rc * n1 * opb);
      } else {
        parallel_memcpy(host_ptr(sh.sb + pd.s0, 0, pd.r0), hbase, pd.sc * pd.rc * n1 * opb);
      }
    }
    pd.active = false;
    return INFLX_OK;
  };

  CU_TRY(cu.p_cuEventRecord(dev->ev_t0, cs));
  uint64_t k = 0;
  for (uint64_t s0 = 0; s0 < S; s0 += s_chunk) {
    const uint64_t sc = std::min(s_chunk, S - s0);
    // P-frontier values -> the module's __constant__ bank, then the column pre-pass (which reads
    // the bank); with the fused prologue both follow the first row chunk's prologue launch
    auto bank_and_columns = [&]() -> inflx_status {
      if (NPF)
        CU_TRY(cu.p_cuMemcpyDtoDAsync(gm->pc_sym, dev->d_pc.ptr + s0 * NPF * 8, sc * NPF * 8, cs));
      if (NCF) {  // column pre-pass: the column block holds an out-of-line libm call
        uint32_t n1c = (uint32_t)n1;
        void* args[] = {&dev->d_cc.ptr, &of1, &dx1, &n1c};
        inflx_status s2 = launch(cu, gm->cols, (unsigned)((n1 + 127) / 128), (unsigned)sc, 1, 128, cs, args);
        if (s2) return s2;
        res.launches++;
      }
      return INFLX_OK;
    };
    if (!fused && (st = bank_and_columns())) return st;
    bool first_chunk = true;
    for (uint64_t r0 = sh.rb; r0 < sh.re; r0 += rows_chunk, ++k) {
      const uint64_t rc = std::min(rows_chunk, sh.re - r0);
      const int slot = (int)(k & 1);
      if (!to_device) {
        if ((st = finish(slot))) return st;  // buffer `slot` must be drained before reuse
      }
      uint32_t n_rows = (uint32_t)rc, n1u = (uint32_t)n1;
      if (fused && first_chunk) {
        uint64_t rbeg = r0;
        void* args[] = {pv.data(), &dev->d_pc.ptr, &dev->d_rc.ptr, &of0, &dx0, &rbeg, &n_rows};
        const unsigned blocks = (unsigned)std::max<uint64_t>(1, (rc + 127) / 128);
        if ((st = launch(cu, gm->prologue, blocks, 1, 1, 128, cs, args))) return st;
        res.launches++;
        if ((st = bank_and_columns())) return st;
      } else if (NRF) {
        uint64_t rbeg = r0;
        void* args[] = {&dev->d_rc.ptr, &of0, &dx0, &rbeg, &n_rows};
        if ((st = launch(cu, gm->rows, (unsigned)((rc + 127) / 128), (unsigned)sc, 1, 128, cs, args)))
          return st;
        res.launches++;
      }
      first_chunk = false;
      CUdeviceptr outp;
      uint64_t comp_stride = sc * rc * n1;
      if (to_device) {
        // user buffer layout [S][rows_total][n1][k] (hesse: [4] outermost over the whole request)
        outp = (CUdeviceptr)rq.out + (hesse ? s0 * rows_total * n1 * 8 : s0 * rows_total * n1 * opb);
        outp += (r0 - sh.rb) * n1 * (hesse ? 8 : opb);  // row block of a tall single grid (S == 1)
        if (hesse) comp_stride = S * rows_total * n1;
      } else {
        outp = dev->d_out[slot].ptr;
      }
      double aux = rq.aux;
      if (k == 0) CU_TRY(cu.p_cuEventRecord(dev->ev_g0, cs));
      // rows per CTA (tools/tune.py sweeps, profiles/tune_rpt_r1.log): the artefact's maximum (16)
      // when that still leaves ~16 waves of CTAs - the column block and the smem staging then
      // amortise over more rows (16384^2 grids: hyper +5.5 %, angular +1.6 %, d5 +1 %, EGNO +-0
      // against 8 rows); else the largest of 8/4/2 that leaves ~4 waves (4096^2: 8 rows are 3 %
      // faster than 16; 1000^2-2048^2: 2-4 rows are 5-30 % faster than 8 on all but the cheapest
      // model, whose 20 us launch does not care).
      const uint64_t col_tiles = (n1 + BLOCK - 1) / BLOCK;
      uint32_t rpt = RPT;
      if (const char* e = getenv("INFLATOX_RPT")) {
        rpt = (uint32_t)std::min<long>(std::max<long>(atol(e), 1), RPT);
      } else {
        const uint64_t wave = (uint64_t)dev->sm_count * 8;
        auto ctas = [&](uint32_t r) { return col_tiles * ((rc + r - 1) / r) * sc; };
        if (rpt > 8 && ctas(rpt) < 16 * wave) rpt = 8;
        while (rpt > 2 && ctas(rpt) < 4 * wave) rpt /= 2;
      }
      while (rpt < RPT && (rc + rpt - 1) / rpt > 65535) rpt *= 2;  // gridDim.y limit
      // Two tile heights (tools/shard_probe.py, profiles/shard_probe_r2.txt): CTAs are dispatched
      // in blockIdx order and all of a wave take the same time, so a launch of uniform tiles ends
      // with the machine draining for about one full tile's duration - 1.5-3 % of a 2048-row
      // shard (one of 8 GPUs on C3), whatever the tile height, because shorter tiles pay the
      // column block more often.  The last ~1.5 waves' worth of rows therefore go out as tiles of
      // a quarter of the height: the drain shrinks with them, the extra column blocks touch a few
      // per cent of the rows.  Single grids only: in a sweep (blockIdx.z = vector) only the last
      // vector's tiles are the launch's tail.
      uint32_t rpt_tail = rpt, n_big = (uint32_t)(rc / rpt);
      if (gm->tail_tiles && sc == 1 && rpt >= 4) {
        uint32_t want_tail = gm->tail_div ? std::max<uint32_t>(2, rpt / gm->tail_div) : 0;
        if (const char* e = getenv("INFLATOX_RPT_TAIL"))
          want_tail = (uint32_t)std::min<long>(std::max<long>(atol(e), 0), rpt);
        if (want_tail && want_tail < rpt) {
          int per_sm = 0;
          {
            std::lock_guard<std::mutex> lk(lib->mu);
            auto it = gm->resident.find(grid_fn);
            if (it != gm->resident.end()) {
              per_sm = it->second;
            } else {
              if (cu.p_cuOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, grid_fn, (int)BLOCK,
                                                                   0) != CUDA_SUCCESS ||
                  per_sm < 1)
                per_sm = 4;
              gm->resident[grid_fn] = per_sm;
            }
          }
          const uint64_t slots = (uint64_t)dev->sm_count * per_sm;
          const uint64_t tail_tiles = (3 * slots / 2 + col_tiles - 1) / col_tiles;  // of `rpt` rows
          const uint64_t big = rc / rpt;
          const uint64_t nb = big > tail_tiles ? big - tail_tiles : 0;
          const uint64_t y = nb + (rc - nb * rpt + want_tail - 1) / want_tail;
          if (y <= 65535) {
            n_big = (uint32_t)nb;
            rpt_tail = want_tail;
          }
        }
      }
      const uint64_t grid_y = n_big + (rc - (uint64_t)n_big * rpt + rpt_tail - 1) / rpt_tail;
      void* args[] = {&outp, &dev->d_rc.ptr, &of1, &dx1, &n1u, &n_rows, &comp_stride, &aux, &rpt,
                      &dev->d_cc.ptr, &n_big, &rpt_tail};
      if ((st = launch(cu, grid_fn, (unsigned)col_tiles, (unsigned)grid_y, (unsigned)sc, BLOCK, cs,
                       args)))
        return st;
      res.launches++;
      if (to_device) CU_TRY(cu.p_cuEventRecord(dev->ev_g1, cs));
      if (!to_device) {
        CU_TRY(cu.p_cuEventRecord(dev->ev_done[slot], cs));
        CU_TRY(cu.p_cuStreamWaitEvent(dev->copy, dev->ev_done[slot], 0));
        pending[slot].active = true;
        pending[slot].s0 = s0;
        pending[slot].sc = sc;
        pending[slot].r0 = r0;
        pending[slot].rc = rc;
        if ((st = issue_copies(slot, pending[slot]))) return st;
        CU_TRY(cu.p_cuEventRecord(dev->ev_copied[slot], dev->copy));
      }
    }
  }
  CU_TRY(cu.p_cuEventRecord(dev->ev_t1, cs));
  if (!to_device) {
    if ((st = finish((int)(k & 1)))) return st;
    if ((st = finish((int)((k & 1) ^ 1)))) return st;
  }
  if (to_device && rq.stream) {
    // asynchronous return: whoever uses this device's scratch next is ordered after this point
    CU_TRY(cu.p_cuEventRecord(dev->ev_user, cs));
    dev->user_pending = true;
  } else {
    CU_TRY(cu.p_cuStreamSynchronize(cs));
    dev->user_pending = false;  // cs waited for ev_user above, so that work has finished too
    float ms = 0;
    CU_TRY(cu.p_cuEventElapsedTime(&ms, dev->ev_t0, dev->ev_t1));
    res.kernel_ms = ms;
    if (to_device) {
      CU_TRY(cu.p_cuEventElapsedTime(&ms, dev->ev_g0, dev->ev_g1));
      res.grid_ms = ms;
    }
  }
  return INFLX_OK;
}

// ---------------------------------------------------------------------------------------------
// exported API
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* inflx_last_error(void) { return g_last_error.c_str(); }

void inflx_free(void* p) { free(p); }

uint64_t inflx_kernel_launches(void) { return g_launches.load(); }

const char* inflx_build_info(void) {
  return "inflatox_b200 engine; CUDA driver API + NVRTC, dlopen'ed; kernels: sm_100a";
}

int inflx_device_count(void) {
  CudaDriver& cu = CudaDriver::get();
  if (!cu.ok) {
    g_last_error = cu.error;
    return -1;
  }
  int n = 0;
  if (cu.p_cuDeviceGetCount(&n) != CUDA_SUCCESS) return -1;
  return n;
}

inflx_status inflx_nvrtc_compile(const char* source, const char* name, const char* const* options,
                                 int n_options, void** cubin, size_t* cubin_size, char** log) {
  Nvrtc& rt = Nvrtc::get();
  if (cubin) *cubin = nullptr;
  if (log) *log = nullptr;
  if (!rt.ok) return fail(INFLX_ERR_NVRTC, rt.error);
  nvrtcProgram prog = nullptr;
  nvrtcResult r = rt.p_nvrtcCreateProgram(&prog, source, name, 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS)
    return fail(INFLX_ERR_NVRTC, std::string("nvrtcCreateProgram: ") + rt.p_nvrtcGetErrorString(r));
  r = rt.p_nvrtcCompileProgram(prog, n_options, options);
  size_t log_size = 0;
  rt.p_nvrtcGetProgramLogSize(prog, &log_size);
  std::string logs(log_size, '\0');
  if (log_size) rt.p_nvrtcGetProgramLog(prog, &logs[0]);
  if (log) {
    *log = (char*)malloc(log_size + 1);
    memcpy(*log, logs.c_str(), log_size);
    (*log)[log_size] = 0;
  }
  if (r != NVRTC_SUCCESS) {
    rt.p_nvrtcDestroyProgram(&prog);
    return fail(INFLX_ERR_NVRTC, std::string("NVRTC compilation of ") + name + " failed (" +
                                     rt.p_nvrtcGetErrorString(r) + "):\n" + logs);
  }
  size_t sz = 0;
  r = rt.p_nvrtcGetCUBINSize(prog, &sz);
  if (r != NVRTC_SUCCESS || sz == 0) {
    rt.p_nvrtcDestroyProgram(&prog);
    return fail(INFLX_ERR_NVRTC, "NVRTC produced no cubin (was a real --gpu-architecture=sm_XX given?)");
  }
  *cubin = malloc(sz);
  rt.p_nvrtcGetCUBIN(prog, (char*)*cubin);
  *cubin_size = sz;
  rt.p_nvrtcDestroyProgram(&prog);
  return INFLX_OK;
}

inflx_status inflx_nvrtc_version(int* major, int* minor) {
  *major = *minor = 0;
  Nvrtc& rt = Nvrtc::get();
  if (!rt.ok) return fail(INFLX_ERR_NVRTC, rt.error);
  if (rt.p_nvrtcVersion(major, minor) != NVRTC_SUCCESS)
    return fail(INFLX_ERR_NVRTC, "nvrtcVersion failed");
  return INFLX_OK;
}

// ---- artefact -------------------------------------------------------------------------------
static inflx_status validate_basis_at_random(inflx_lib* lib);
void inflx_close(inflx_lib* lib);

inflx_status inflx_open(const char* lib_path, int check_basis, inflx_lib** out) {
  *out = nullptr;
  FILE* f = fopen(lib_path, "rb");
  if (!f)  // LibInflxRsErr::Io (reference src/err.rs:43)
    return fail(INFLX_ERR_IO, fmt("Could not load Inflatox Compilation Artefact (path: %s). "
                                  "Error: \"%s\"", lib_path, strerror(errno)));
  auto lib = std::make_unique<inflx_lib>();
  lib->path = lib_path;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  lib->file.resize(sz > 0 ? sz : 0);
  size_t got = sz > 0 ? fread(lib->file.data(), 1, sz, f) : 0;
  fclose(f);
  if (got != (size_t)sz || (size_t)sz < sizeof(FileHeader) ||
      memcmp(lib->file.data(), "INFLXB2\0", 8) != 0)
    return fail(INFLX_ERR_IO, fmt("Could not load Inflatox Compilation Artefact (path: %s). "
                                  "Error: \"not an inflatox_b200 CUDA artefact\"", lib_path));
  memcpy(&lib->hdr, lib->file.data(), sizeof(FileHeader));
  const FileHeader& h = lib->hdr;
  // reference src/inflatox_version.rs:48-53: major and minor must match
  if (h.abi[0] != kAbi[0] || h.abi[1] != kAbi[1])
    return fail(INFLX_ERR_VERSION,
                fmt("Cannot load Inflatox Compilation Artefact compiled for Inflatox ABI v%u.%u.%u "
                    "using current Inflatox installation (v%u.%u.%u)", h.abi[0], h.abi[1], h.abi[2],
                    kAbi[0], kAbi[1], kAbi[2]));
  if (h.container_version != kContainerVersion)
    return fail(INFLX_ERR_VERSION,
                fmt("Cannot load Inflatox Compilation Artefact with container layout v%u using "
                    "current Inflatox installation (container layout v%u): compile the model again",
                    h.container_version, kContainerVersion));
  if (sizeof(FileHeader) + (size_t)h.n_groups * sizeof(GroupEntry) > lib->file.size())
    return fail(INFLX_ERR_IO, fmt("Could not load Inflatox Compilation Artefact (path: %s). "
                                  "Error: \"truncated artefact\"", lib_path));
  if (h.rpt == 0 || h.rpt > 1024 || h.block == 0 || h.block > 1024 || (h.block & 31))
    return fail(INFLX_ERR_IO, fmt("Could not load Inflatox Compilation Artefact (path: %s). "
                                  "Error: \"corrupt header (rows per CTA %u, CTA width %u)\"",
                                  lib_path, h.rpt, h.block));
  lib->groups.resize(h.n_groups);
  memcpy(lib->groups.data(), lib->file.data() + sizeof(FileHeader), h.n_groups * sizeof(GroupEntry));
  for (auto& g : lib->groups)
    if (g.cubin_offset + g.cubin_size > lib->file.size())
      return fail(INFLX_ERR_IO, fmt("Could not load Inflatox Compilation Artefact (path: %s). "
                                    "Error: \"truncated artefact\"", lib_path));
  // the symbols the reference loader insists on (src/dylib.rs:101-131): V and grad_norm_squared
  for (const char* need : {"pot", "eps"})
    if (!lib->group(need))
      return fail(INFLX_ERR_MISSING_SYMBOL,
                  fmt("Could not find symbol \"%s\" in %s", need, lib_path));
  lib->devices = default_devices();
  if (check_basis) {
    inflx_status st = validate_basis_at_random(lib.get());
    if (st) {
      const std::string why = g_last_error;
      inflx_close(lib.release());  // unloads the modules the validation loaded
      return fail(st, why);
    }
  }
  *out = lib.release();
  return INFLX_OK;
}

void inflx_close(inflx_lib* lib) {
  if (!lib) return;
  CudaDriver& cu = CudaDriver::get();
  if (cu.ok) {
    for (auto& kv : lib->modules) {
      DeviceState* dev = nullptr;
      if (get_device(kv.first.first, &dev) == INFLX_OK) {
        std::lock_guard<std::mutex> lk(dev->mu);
        cu.p_cuCtxSetCurrent(dev->ctx);
        if (dev->user_pending) {  // a kernel of this module may still run on a caller's stream
          cu.p_cuEventSynchronize(dev->ev_user);
          dev->user_pending = false;
        }
        cu.p_cuStreamSynchronize(dev->compute);
        cu.p_cuModuleUnload(kv.second->mod);
      }
    }
  }
  delete lib;
}

uint32_t inflx_n_fields(const inflx_lib* lib) { return lib->hdr.dim; }
uint32_t inflx_n_parameters(const inflx_lib* lib) { return lib->hdr.n_params; }
const char* inflx_model_name(const inflx_lib* lib) { return lib->hdr.model_name; }
void inflx_abi_version(const inflx_lib* lib, uint16_t out[3]) { memcpy(out, lib->hdr.abi, 6); }

inflx_status inflx_set_devices(inflx_lib* lib, const int* ordinals, int n) {
  if (n <= 0) {
    lib->devices = default_devices();
    return INFLX_OK;
  }
  lib->devices.assign(ordinals, ordinals + n);
  return INFLX_OK;
}
int inflx_get_devices(const inflx_lib* lib, int* ordinals, int capacity) {
  int n = (int)lib->devices.size();
  for (int i = 0; i < n && i < capacity; ++i) ordinals[i] = lib->devices[i];
  return n;
}

// ---- static sharding: the ONE implementation of the rule ------------------------
inflx_status inflx_shard_of(uint64_t n_rows, uint64_t n_vectors, uint64_t index, uint64_t count,
                            uint64_t out[4]) {
  if (count == 0 || index >= count)
    return fail(INFLX_ERR_SHAPE, fmt("shard %llu outside a partition into %llu",
                                     (unsigned long long)index, (unsigned long long)count));
  if (n_vectors >= count && count > 1 && n_vectors > 1) {  // a sweep: blocks of parameter vectors
    out[0] = 0;
    out[1] = n_rows;
    out[2] = n_vectors * index / count;
    out[3] = n_vectors * (index + 1) / count;
  } else {  // one grid (or fewer vectors than shards): contiguous row blocks
    out[0] = n_rows * index / count;
    out[1] = n_rows * (index + 1) / count;
    out[2] = 0;
    out[3] = n_vectors;
  }
  return INFLX_OK;
}

// ---- generic grid evaluation -----------------------------------------------------------------
inflx_status inflx_grid_eval(inflx_lib* lib, const inflx_grid_request* rq,
                             inflx_grid_report* report) {
  auto t0 = std::chrono::steady_clock::now();
  if (rq->op < 0 || rq->op > INFLX_OP_HESSE)
    return fail(INFLX_ERR_SHAPE, "unknown grid operation");
  if (lib->hdr.dim != 2)  // Hesse2D::new asserts n_fields == 2 (reference src/hesse_bindings.rs:203)
    return shape_error({2}, {lib->hdr.dim},
                       "the Anguelova & Lazaroiu consistency condition requires a 2-field model.");
  if (rq->row_end > rq->n0 || rq->row_begin > rq->row_end)
    return fail(INFLX_ERR_SHAPE, "row range outside the grid");
  if (rq->n_vectors < 1) return fail(INFLX_ERR_SHAPE, "n_vectors must be >= 1");
  if (rq->n1 > 0xffffffffull || rq->n0 > 0xffffffffull)
    return fail(INFLX_ERR_SHAPE, "grid axes are limited to 2^32-1 points");
  std::vector<int> devs;
  if (rq->device >= 0)
    devs.push_back(rq->device);
  else
    devs = lib->devices;
  if (rq->out_is_device && devs.size() != 1) devs.resize(1);
  if (devs.empty()) return fail(INFLX_ERR_CUDA, "no CUDA device selected");

  // shards: parameter vectors block-distributed when there are enough, else row blocks - the one
  // rule (inflx_shard_of) that inflatox_b200.sharding applies across the ranks of a torchrun job
  std::vector<Shard> shards;
  const uint64_t nd = devs.size(), S = rq->n_vectors, R = rq->row_end - rq->row_begin;
  for (uint64_t d = 0; d < nd; ++d) {
    uint64_t o[4];
    inflx_shard_of(R, S, d, nd, o);
    if (o[1] > o[0] && o[3] > o[2])
      shards.push_back({rq->row_begin + o[0], rq->row_begin + o[1], o[2], o[3], devs[d]});
  }
  std::vector<ShardResult> results(shards.size());
  auto work = [&](size_t i) {
    g_last_error.clear();
    results[i].status = run_shard(lib, *rq, shards[i], results[i]);
    if (results[i].status) results[i].error = g_last_error;
  };
  if (shards.size() <= 1) {
    if (!shards.empty()) work(0);
  } else {
    std::vector<std::thread> th;
    for (size_t i = 0; i < shards.size(); ++i) th.emplace_back(work, i);
    for (auto& t : th) t.join();
  }
  inflx_grid_report rep = {};
  rep.n_devices = (int)shards.size();
  for (auto& r : results) {
    if (r.status) return fail(r.status, r.error);
    rep.kernel_ms = std::max(rep.kernel_ms, r.kernel_ms);
    rep.grid_ms = std::max(rep.grid_ms, r.grid_ms);
    rep.launches += r.launches;
    rep.d2h_bytes += r.d2h;
    rep.h2d_bytes += r.h2d;
  }
  rep.total_ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (report) *report = rep;
  return INFLX_OK;
}

inflx_status inflx_points_eval(inflx_lib* lib, int opi, const double* p, const double* xs,
                               uint64_t n, double aux, double* out) {
  if (opi < 0 || opi > INFLX_OP_BASIS) return fail(INFLX_ERR_SHAPE, "unknown point operation");
  if (opi == INFLX_OP_FLAG_QUANTUM_DIF)
    return fail(INFLX_ERR_SHAPE, "flag_quantum_dif has no point-list form");
  if (lib->hdr.dim != 2)
    return shape_error({2}, {lib->hdr.dim}, "point evaluation requires a 2-field model.");
  if (n == 0) return INFLX_OK;
  CudaDriver& cu = CudaDriver::get();
  if (lib->devices.empty()) return fail(INFLX_ERR_CUDA, "no CUDA device selected");
  DeviceState* dev = nullptr;
  inflx_status st = get_device(lib->devices[0], &dev);
  if (st) return st;
  std::lock_guard<std::mutex> lk(dev->mu);
  CU_TRY(cu.p_cuCtxSetCurrent(dev->ctx));
  const OpInfo& op = kOps[opi];
  GroupModule* gm = nullptr;
  if ((st = load_module(lib, dev, op.group, &gm))) return st;
  const GroupEntry* ge = lib->group(op.group);
  CUfunction fn = nullptr;
  if ((st = get_fn(lib, gm, std::string("inflx_points_") + op.name, &fn))) return st;
  const uint32_t P = lib->hdr.n_params, NPF = ge->npf;
  CUstream cs = dev->compute;
  if ((st = order_after_user_stream(cu, dev, cs))) return st;

  // scalar fast path (calc_V / calc_H, reference src/lib.rs:309-339, 384-419): one launch; inputs
  // and result live in mapped page-locked memory, so no copy is enqueued at all
  const uint64_t opb_s = op.out_bytes, xs_at = (P + 1) & ~1u;
  if ((opi == INFLX_OP_POTENTIAL || opi == INFLX_OP_HESSE) && n <= kScalarMaxPoints &&
      (xs_at + 2 * n) * 8 <= kScalarScratch / 2 && n * opb_s <= kScalarScratch / 2) {
    CUfunction sfn = nullptr;
    if (get_fn(lib, gm, std::string("inflx_scalar_") + op.name, &sfn) == INFLX_OK) {
      if (!dev->h_scalar) {
        CU_TRY(cu.p_cuMemHostAlloc(&dev->h_scalar, kScalarScratch,
                                   CU_MEMHOSTALLOC_PORTABLE | CU_MEMHOSTALLOC_DEVICEMAP));
        CU_TRY(cu.p_cuMemHostGetDevicePointer(&dev->d_scalar, dev->h_scalar, 0));
      }
      double* in = static_cast<double*>(dev->h_scalar);
      double* res = in + kScalarScratch / 16;
      if (P) memcpy(in, p, P * 8);
      memcpy(in + xs_at, xs, n * 16);
      CUdeviceptr d_in = dev->d_scalar, d_res = dev->d_scalar + kScalarScratch / 2;
      void* args[] = {&d_res, &d_in, &n};
      if ((st = launch(cu, sfn, (unsigned)((n + 63) / 64), 1, 1, 64, cs, args))) return st;
      CU_TRY(cu.p_cuStreamSynchronize(cs));
      dev->user_pending = false;
      memcpy(out, res, n * opb_s);
      return INFLX_OK;
    }
    g_last_error.clear();  // artefact without scalar kernels: the general path below
  }
  if (P) {
    if ((st = ensure_dev(cu, dev->d_p, P * 8))) return st;
    CU_TRY(cu.p_cuMemcpyHtoDAsync(dev->d_p.ptr, p, P * 8, cs));
  }
  if (NPF) {
    if ((st = ensure_dev(cu, dev->d_pc, NPF * 8))) return st;
    uint32_t one = 1;
    void* args[] = {&dev->d_p.ptr, &dev->d_pc.ptr, &one};
    if ((st = launch(cu, gm->params, 1, 1, 1, 64, cs, args))) return st;
    CU_TRY(cu.p_cuMemcpyDtoDAsync(gm->pc_sym, dev->d_pc.ptr, NPF * 8, cs));
  }
  const uint64_t opb = op.out_bytes;
  if (n <= (1ull << 16)) {  // a trajectory of the reference tests' size: one launch, plain copies
    if ((st = ensure_dev(cu, dev->d_xs, n * 16))) return st;
    if ((st = ensure_dev(cu, dev->d_out[0], n * opb))) return st;
    CU_TRY(cu.p_cuMemcpyHtoDAsync(dev->d_xs.ptr, xs, n * 16, cs));
    void* args[] = {&dev->d_out[0].ptr, &dev->d_xs.ptr, &n, &aux};
    if ((st = launch(cu, fn, (unsigned)((n + 127) / 128), 1, 1, 128, cs, args))) return st;
    CU_TRY(cu.p_cuMemcpyDtoHAsync(out, dev->d_out[0].ptr, n * opb, cs));
    CU_TRY(cu.p_cuStreamSynchronize(cs));
  } else {
    // long point lists: 1 Mi-point chunks through page-locked staging, two slots - the upload and
    // kernel of chunk k+1 run while chunk k's result travels back on the copy stream and chunk
    // k-1's is unstaged into the caller's (pageable) array
    const uint64_t chunk = 1ull << 20;
    if ((st = ensure_dev(cu, dev->d_xs, chunk * 16))) return st;
    for (int i = 0; i < 2; ++i) {
      if ((st = ensure_dev(cu, dev->d_out[i], chunk * opb))) return st;
      if ((st = ensure_pin(cu, dev->stage[i], chunk * opb))) return st;
      if ((st = ensure_pin(cu, dev->stage_in[i], chunk * 16))) return st;
    }
    struct {
      bool active = false;
      uint64_t b = 0, m = 0;
    } pend[2];
    auto drain = [&](int slot) -> inflx_status {
      if (!pend[slot].active) return INFLX_OK;
      CU_TRY(cu.p_cuEventSynchronize(dev->ev_copied[slot]));
      parallel_memcpy((char*)out + pend[slot].b * opb, dev->stage[slot].ptr, pend[slot].m * opb);
      pend[slot].active = false;
      return INFLX_OK;
    };
    uint64_t k = 0;
    for (uint64_t b = 0; b < n; b += chunk, ++k) {
      const int slot = (int)(k & 1);
      uint64_t m = std::min(chunk, n - b);
      if ((st = drain(slot))) return st;  // also: stage_in[slot] is free again (its kernel ran)
      parallel_memcpy(dev->stage_in[slot].ptr, xs + 2 * b, m * 16);
      CU_TRY(cu.p_cuMemcpyHtoDAsync(dev->d_xs.ptr, dev->stage_in[slot].ptr, m * 16, cs));
      void* args[] = {&dev->d_out[slot].ptr, &dev->d_xs.ptr, &m, &aux};
      if ((st = launch(cu, fn, (unsigned)((m + 127) / 128), 1, 1, 128, cs, args))) return st;
      CU_TRY(cu.p_cuEventRecord(dev->ev_done[slot], cs));
      CU_TRY(cu.p_cuStreamWaitEvent(dev->copy, dev->ev_done[slot], 0));
      CU_TRY(cu.p_cuMemcpyDtoHAsync(dev->stage[slot].ptr, dev->d_out[slot].ptr, m * opb, dev->copy));
      CU_TRY(cu.p_cuEventRecord(dev->ev_copied[slot], dev->copy));
      pend[slot].active = true;
      pend[slot].b = b;
      pend[slot].m = m;
    }
    if ((st = drain((int)(k & 1)))) return st;
    if ((st = drain((int)((k & 1) ^ 1)))) return st;
    CU_TRY(cu.p_cuStreamSynchronize(cs));
  }
  dev->user_pending = false;
  return INFLX_OK;
}

// ---- reference-shaped entry points ------------------------------------------------------------
static std::string human_duration(double s) {  // indicatif::HumanDuration, roughly
  if (s < 1.0) return "0 seconds";
  if (s < 60.0) return fmt("%d second%s", (int)s, (int)s == 1 ? "" : "s");
  if (s < 3600.0) return fmt("%d minute%s", (int)(s / 60), (int)(s / 60) == 1 ? "" : "s");
  return fmt("%d hour%s", (int)(s / 3600), (int)(s / 3600) == 1 ? "" : "s");
}

// validiate_p (reference src/anguelova.rs:70-80; the reference reports `expected: [2]` there)
static inflx_status check_params(const inflx_lib* lib, size_t p_len) {
  if (p_len != lib->hdr.n_params)
    return shape_error({2}, {p_len}, fmt("model \"%s\" has %u paramters", lib->hdr.model_name,
                                         lib->hdr.n_params));
  return INFLX_OK;
}
// convert_start_stop (reference src/lib.rs:117-139)
static inflx_status check_start_stop(size_t rows, size_t cols, size_t n_fields) {
  if (rows != 2 || cols != n_fields)
    return shape_error({2, n_fields}, {rows, cols},
                       "start_stop array should have 2 rows and as many columns as there are fields");
  return INFLX_OK;
}

static inflx_status grid_entry(inflx_lib* lib, int op, const char* hello, const double* p,
                               size_t p_len, void* out, size_t n0, size_t n1,
                               const double* start_stop, size_t ss_rows, size_t ss_cols,
                               double aux) {
  inflx_status st;
  if (lib->hdr.dim != 2)
    return shape_error({2}, {lib->hdr.dim},
                       "the Anguelova & Lazaroiu consistency condition requires a 2-field model.");
  if ((st = check_params(lib, p_len))) return st;
  if ((st = check_start_stop(ss_rows, ss_cols, 2))) return st;
  inflx_grid_request rq = {};
  rq.op = op;
  rq.params = p;
  rq.n_vectors = 1;
  rq.n0 = n0;
  rq.n1 = n1;
  memcpy(rq.start_stop, start_stop, sizeof rq.start_stop);
  rq.row_begin = 0;
  rq.row_end = n0;
  rq.aux = aux;
  rq.out = out;
  rq.device = -1;
  info(fmt("%s on %zu CUDA device(s).", hello, lib->devices.size()));
  inflx_grid_report rep;
  if ((st = inflx_grid_eval(lib, &rq, &rep))) return st;
  info(fmt("Calculation finished. Took %s.", human_duration(rep.total_ms / 1e3).c_str()));
  return INFLX_OK;
}

inflx_status inflx_complete_analysis(inflx_lib* lib, const double* p, size_t p_len, double* out,
                                     size_t n0, size_t n1, size_t n_last, const double* start_stop,
                                     size_t ss_rows, size_t ss_cols, int, size_t) {
  if (n_last != 6)  // reference src/anguelova.rs:480-486
    return shape_error({n0, n1, 6}, {n0, n1, n_last},
                       "Output array should be 3D. Last axis must have lenght 6");
  return grid_entry(lib, INFLX_OP_COMPLETE_ANALYSIS, "Calculating full analysis", p, p_len, out,
                    n0, n1, start_stop, ss_rows, ss_cols, 0.0);
}
inflx_status inflx_consistency_only(inflx_lib* lib, const double* p, size_t p_len, double* out,
                                    size_t n0, size_t n1, const double* start_stop, size_t ss_rows,
                                    size_t ss_cols, int, size_t) {
  return grid_entry(lib, INFLX_OP_CONSISTENCY_ONLY, "Calculating consistency condition ONLY", p,
                    p_len, out, n0, n1, start_stop, ss_rows, ss_cols, 0.0);
}
inflx_status inflx_consistency_rapidturn_only(inflx_lib* lib, const double* p, size_t p_len,
                                              double* out, size_t n0, size_t n1,
                                              const double* start_stop, size_t ss_rows,
                                              size_t ss_cols, int, size_t) {
  return grid_entry(lib, INFLX_OP_CONSISTENCY_RAPIDTURN_ONLY,
                    "Calculating consistency condition with rapid-turn approximation", p, p_len,
                    out, n0, n1, start_stop, ss_rows, ss_cols, 0.0);
}
inflx_status inflx_epsilon_v_only(inflx_lib* lib, const double* p, size_t p_len, double* out,
                                  size_t n0, size_t n1, const double* start_stop, size_t ss_rows,
                                  size_t ss_cols, int, size_t) {
  return grid_entry(lib, INFLX_OP_EPSILON_V_ONLY, "Calculating potential slow-roll parameter ε_V", p,
                    p_len, out, n0, n1, start_stop, ss_rows, ss_cols, 0.0);
}
inflx_status inflx_flag_quantum_dif(inflx_lib* lib, const double* p, size_t p_len, uint8_t* x,
                                    size_t n0, size_t n1, const double* start_stop, size_t ss_rows,
                                    size_t ss_cols, int, double accuracy) {
  return grid_entry(lib, INFLX_OP_FLAG_QUANTUM_DIF, "Calculating zeros of the potential gradient",
                    p, p_len, x, n0, n1, start_stop, ss_rows, ss_cols, accuracy);
}

static inflx_status traj_entry(inflx_lib* lib, int op, const char* hello, const double* p,
                               size_t p_len, const double* x, size_t n, size_t x_cols, double* out,
                               size_t out_rows) {
  inflx_status st;
  if (lib->hdr.dim != 2)
    return shape_error({2}, {lib->hdr.dim},
                       "the Anguelova & Lazaroiu consistency condition requires a 2-field model.");
  if ((st = check_params(lib, p_len))) return st;
  if (x_cols != 2)
    return shape_error({n, 2}, {n, x_cols}, "field-space array should have shape (n, 2)");
  if (out_rows != n)  // reference src/anguelova.rs:672-679
    return shape_error({n}, {out_rows},
                       "First axis of output array and field-space array should have the same length");
  info(fmt("%s on 1 CUDA device.", hello));
  auto t0 = std::chrono::steady_clock::now();
  if ((st = inflx_points_eval(lib, op, p, x, n, 0.0, out))) return st;
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  info(fmt("Calculation finished. Took %s.", human_duration(s).c_str()));
  return INFLX_OK;
}

inflx_status inflx_complete_analysis_on_trajectory(inflx_lib* lib, const double* p, size_t p_len,
                                                   const double* x, size_t n, size_t x_cols,
                                                   double* out, size_t out_rows, size_t out_cols,
                                                   int, size_t) {
  if (out_cols != 6)  // reference src/anguelova.rs:665-671
    return shape_error({out_rows, 6}, {out_rows, out_cols},
                       "Output array should be 2D. Last axis must have lenght 6");
  return traj_entry(lib, INFLX_OP_COMPLETE_ANALYSIS, "Calculating full analysis on trajectory", p,
                    p_len, x, n, x_cols, out, out_rows);
}
inflx_status inflx_consistency_only_on_trajectory(inflx_lib* lib, const double* p, size_t p_len,
                                                  const double* x, size_t n, size_t x_cols,
                                                  double* out, size_t out_len, int, size_t) {
  return traj_entry(lib, INFLX_OP_CONSISTENCY_ONLY,
                    "Calculating consistency condition ONLY on trajectory", p, p_len, x, n, x_cols,
                    out, out_len);
}
inflx_status inflx_consistency_rapidturn_only_on_trajectory(inflx_lib* lib, const double* p,
                                                            size_t p_len, const double* x, size_t n,
                                                            size_t x_cols, double* out,
                                                            size_t out_len, int, size_t) {
  return traj_entry(lib, INFLX_OP_CONSISTENCY_RAPIDTURN_ONLY,
                    "Calculating consistency condition with rapid-turn approximation on trajectory",
                    p, p_len, x, n, x_cols, out, out_len);
}
inflx_status inflx_epsilon_v_only_on_trajectory(inflx_lib* lib, const double* p, size_t p_len,
                                                const double* x, size_t n, size_t x_cols,
                                                double* out, size_t out_len, int, size_t) {
  return traj_entry(lib, INFLX_OP_EPSILON_V_ONLY,
                    "Calculating potential slow-roll parameter ε_V on trajectory", p, p_len, x, n,
                    x_cols, out, out_len);
}

// ---- InflatoxPyDyLib methods -------------------------------------------------------------------
static inflx_status check_xp(const inflx_lib* lib, size_t x_len, size_t p_len) {
  if (x_len != lib->hdr.dim)  // reference src/lib.rs:316-323
    return shape_error({lib->hdr.dim}, {x_len},
                       "expected a 1D array with as many elements as there are field-space coordinates");
  if (p_len != lib->hdr.n_params)  // :328-334
    return shape_error({lib->hdr.n_params}, {p_len},
                       "expected a 1D array with as many elements as there are model parameters");
  return INFLX_OK;
}

inflx_status inflx_potential(inflx_lib* lib, const double* x, size_t x_len, const double* p,
                             size_t p_len, double* value) {
  inflx_status st;
  if ((st = check_xp(lib, x_len, p_len))) return st;
  return inflx_points_eval(lib, INFLX_OP_POTENTIAL, p, x, 1, 0.0, value);
}

inflx_status inflx_hesse(inflx_lib* lib, const double* x, size_t x_len, const double* p,
                         size_t p_len, double* out4) {
  inflx_status st;
  if ((st = check_xp(lib, x_len, p_len))) return st;
  return inflx_points_eval(lib, INFLX_OP_HESSE, p, x, 1, 0.0, out4);
}

static inflx_status array_entry(inflx_lib* lib, int op, double* out, size_t n0, size_t n1,
                                const double* p, size_t p_len, const double* start_stop,
                                size_t ss_rows, size_t ss_cols) {
  inflx_status st;
  if ((st = check_start_stop(ss_rows, ss_cols, lib->hdr.dim))) return st;
  if (p_len != lib->hdr.n_params)
    return shape_error({lib->hdr.n_params}, {p_len},
                       "expected a 1D array with as many elements as there are model parameters");
  inflx_grid_request rq = {};
  rq.op = op;
  rq.params = p;
  rq.n_vectors = 1;
  rq.n0 = n0;
  rq.n1 = n1;
  memcpy(rq.start_stop, start_stop, sizeof rq.start_stop);
  rq.row_begin = 0;
  rq.row_end = n0;
  rq.out = out;
  rq.device = -1;
  return inflx_grid_eval(lib, &rq, nullptr);
}

inflx_status inflx_potential_array(inflx_lib* lib, double* x_out, size_t n0, size_t n1,
                                   const double* p, size_t p_len, const double* start_stop,
                                   size_t ss_rows, size_t ss_cols) {
  return array_entry(lib, INFLX_OP_POTENTIAL, x_out, n0, n1, p, p_len, start_stop, ss_rows, ss_cols);
}
inflx_status inflx_hesse_array(inflx_lib* lib, double* out, size_t n0, size_t n1, const double* p,
                               size_t p_len, const double* start_stop, size_t ss_rows,
                               size_t ss_cols) {
  return array_entry(lib, INFLX_OP_HESSE, out, n0, n1, p, p_len, start_stop, ss_rows, ss_cols);
}

// ---- basis validation (reference src/lib.rs:142-307) ------------------------------------------
static bool is_normal(double v) { return std::isnormal(v); }

// checks one evaluated point; returns non-zero status on a hard failure, sets *nan on soft failure
static inflx_status check_basis_point(const double* r /*7*/, const double* x, double accuracy,
                                      bool* nan) {
  const double ip[3] = {r[4], r[5], r[6]};
  const double* vec[2] = {r, r + 2};
  const int pairs[3][2] = {{0, 0}, {0, 1}, {1, 1}};
  for (int k = 0; k < 3; ++k) {
    int i = pairs[k][0], j = pairs[k][1];
    double v = ip[k];
    if (i == j) {
      if (!is_normal(v)) {
        warn(fmt("Norm of basisvector %d is %g at field-space point [%.3f, %.3f]. v%d=[%.3f, %.3f]\n"
                 "Are we outside the model's domain?", i, v, x[0], x[1], i, vec[i][0], vec[i][1]));
        *nan = true;
      } else if (std::fabs(v - 1.) >= accuracy) {
        return fail(INFLX_ERR_BASIS_NORM,
                    fmt("Expected basis vector %d to be normalised everywhere in the models domain. "
                        "Instead, found norm %g at [%.3f, %.3f].", i, v, x[0], x[1]));
      }
    } else {
      if (!is_normal(v) && v != 0.0) {
        warn(fmt("w%d•w%d = %g at field-space point [%.3f, %.3f].\nv%d=[%.3f, %.3f]\nv%d=[%.3f, %.3f]\n"
                 "Are we outside the model's domain?", i, j, v, x[0], x[1], i, vec[i][0], vec[i][1],
                 j, vec[j][0], vec[j][1]));
        *nan = true;
      } else if (std::fabs(v) >= accuracy) {
        return fail(INFLX_ERR_BASIS_OTH,
                    fmt("Expected basis vectors w%d and w%d to be orthogonal everywhere in the model's "
                        "domain. Instead, found inner product %g at [%.3f, %.3f].", i, j, v, x[0], x[1]));
      }
    }
  }
  return INFLX_OK;
}

static std::string params_dbg(const std::vector<double>& p) {
  std::string s = "[";
  for (size_t i = 0; i < p.size(); ++i) s += (i ? ", " : "") + fmt("%.3f", p[i]);
  return s + "]";
}

static inflx_status validate_basis_at_random(inflx_lib* lib) {
  if (lib->hdr.dim != 2 || !lib->group("bas")) return INFLX_OK;
  const double accuracy = 1e-3;
  const int num_points = 100;
  std::mt19937_64 rng(std::random_device{}());
  std::uniform_real_distribution<double> u(0.0, 1.0);
  std::vector<double> p(lib->hdr.n_params), xs(2 * num_points), out(7 * num_points);
  for (auto& v : p) v = 10. * (-1. + 2. * u(rng));
  for (auto& v : xs) v = -1. + 2. * u(rng);
  inflx_status st = inflx_points_eval(lib, INFLX_OP_BASIS, p.data(), xs.data(), num_points, 0.0, out.data());
  if (st) return st;
  int failed = 0;
  for (int k = 0; k < num_points; ++k) {
    bool nan = false;
    if ((st = check_basis_point(&out[7 * k], &xs[2 * k], accuracy, &nan))) return st;
    failed += nan;
  }
  if (failed)
    warn(fmt("Inflatox was unable to verify basis orthonormality at %d out of %d tested points.\n"
             "This could be indicative of a defective model.\nUsed parameter values: p=%s",
             failed, num_points, params_dbg(p).c_str()));
  return INFLX_OK;
}

inflx_status inflx_validate_basis_on_domain(inflx_lib* lib, const uint32_t* num_points,
                                            size_t n_axes, const double* p, size_t p_len,
                                            const double* start_stop, size_t ss_rows,
                                            size_t ss_cols, double accuracy) {
  info("Validating basis orthonormality on specified domain. This may take a while...");
  inflx_status st;
  if (n_axes != lib->hdr.dim)
    return shape_error({}, {n_axes},
                       "expected an array with with the same number of axes as there are field-space coordinates");
  if ((st = check_start_stop(ss_rows, ss_cols, lib->hdr.dim))) return st;
  if (p_len != lib->hdr.n_params)
    return shape_error({lib->hdr.n_params}, {p_len},
                       "expected a 1D array with as many elements as there are model parameters");
  // reference src/lib.rs:250-258: along each axis the points are start-point with that axis set
  // to `stop + spacing*idx` (sic) - reproduced as written.
  std::vector<double> xs;
  for (size_t axis = 0; axis < 2; ++axis) {
    double start = start_stop[axis * 2], stop = start_stop[axis * 2 + 1];
    double spacing = (stop - start) / (double)num_points[axis];
    for (uint32_t idx = 0; idx < num_points[axis]; ++idx) {
      double pt[2] = {start_stop[0], start_stop[2]};
      pt[axis] = stop + spacing * (double)idx;
      xs.push_back(pt[0]);
      xs.push_back(pt[1]);
    }
  }
  const uint64_t n = xs.size() / 2;
  std::vector<double> out(7 * n);
  if ((st = inflx_points_eval(lib, INFLX_OP_BASIS, p, xs.data(), n, 0.0, out.data()))) return st;
  int failed = 0;
  for (uint64_t k = 0; k < n; ++k) {
    bool nan = false;
    if ((st = check_basis_point(&out[7 * k], &xs[2 * k], accuracy, &nan))) return st;
    failed += nan;
  }
  if (failed) {
    std::vector<double> pv(p, p + p_len);
    warn(fmt("Inflatox was unable to verify basis orthonormality at %d out of %g tested points.\n"
             "This could be indicative of a defective model.\nUsed parameter values: p=%s", failed,
             (double)num_points[0] * (double)num_points[1], params_dbg(pv).c_str()));
  }
  return INFLX_OK;
}

// ---- roofline denominator: sustained FP64 FMA rate of one device -------------------------------
// MEASURED_PEAKS.json carries no fp64 entry, so the bench measures it with this micro-kernel:
// 8 independent DFMA chains per thread, enough resident warps to saturate the FP64 pipe.
static const char* kPeakSrc = R"(
extern "C" __global__ void inflx_dfma_peak(double* out, double a, double b, int iters) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4.,
         x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;
}
)";

inflx_status inflx_measure_fp64_peak(int device, int repeats, double* tflops_best,
                                     double* tflops_median) {
  CudaDriver& cu = CudaDriver::get();
  DeviceState* dev = nullptr;
  inflx_status st = get_device(device, &dev);
  if (st) return st;
  void* cubin = nullptr;
  size_t size = 0;
  char* log = nullptr;
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17"};
  st = inflx_nvrtc_compile(kPeakSrc, "inflx_peak.cu", opts, 2, &cubin, &size, &log);
  free(log);
  if (st) return st;
  std::lock_guard<std::mutex> lk(dev->mu);
  CU_TRY(cu.p_cuCtxSetCurrent(dev->ctx));
  CUmodule mod = nullptr;
  CUresult r = cu.p_cuModuleLoadData(&mod, cubin);
  free(cubin);
  if (r != CUDA_SUCCESS) return fail(INFLX_ERR_CUDA, "cuModuleLoadData(peak kernel) failed");
  CUfunction fn = nullptr;
  CU_TRY(cu.p_cuModuleGetFunction(&fn, mod, "inflx_dfma_peak"));
  if ((st = order_after_user_stream(cu, dev, dev->compute))) return st;
  if ((st = ensure_dev(cu, dev->d_p, 64))) return st;
  int sms = 0;
  cu.p_cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev->dev);
  const unsigned blocks = (unsigned)sms * 8, threads = 256;
  int iters = 1 << 15;
  double a = 1.0000001, b = 1e-9;
  void* args[] = {&dev->d_p.ptr, &a, &b, &iters};
  std::vector<double> tf;
  for (int k = 0; k < repeats + 2; ++k) {
    CU_TRY(cu.p_cuEventRecord(dev->ev_t0, dev->compute));
    if ((st = launch(cu, fn, blocks, 1, 1, threads, dev->compute, args))) return st;
    CU_TRY(cu.p_cuEventRecord(dev->ev_t1, dev->compute));
    CU_TRY(cu.p_cuStreamSynchronize(dev->compute));
    float ms = 0;
    CU_TRY(cu.p_cuEventElapsedTime(&ms, dev->ev_t0, dev->ev_t1));
    if (k >= 2) tf.push_back(2.0 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12);
  }
  cu.p_cuModuleUnload(mod);
  std::sort(tf.begin(), tf.end());
  *tflops_best = tf.back();
  *tflops_median = tf[tf.size() / 2];
  return INFLX_OK;
}

// ---- pinned host memory ------------------------------------------------------------------------
// Large blocks: anonymous mmap + MADV_HUGEPAGE, pre-faulted by a few threads, then registered with
// the driver.  Measured on the round-1 box for 12 GiB (tools/pin_probe.py): cuMemHostAlloc 5.7 s
// (and it holds the driver's lock for all of it, stalling every concurrent launch/copy);
// populate 1.3-1.4 s (no driver involvement) + cuMemHostRegister 1.0-2.3 s; same 55 GB/s D2H rate.
// NUMA placement.  The D2H copies are DMA writes into this memory; on a two-socket box a GPU whose
// destination pages live on the other socket pushes its whole output over the inter-socket link
// (round 1, 8 GPUs: 92 GB/s aggregate with everything on one node).  A block is therefore cut into
// one slice per device - the same equal split the row sharding uses - and each slice is
// first-touched by threads pinned to the CPUs of the NUMA node its GPU hangs off
// (/sys/bus/pci/devices/<bus id>/numa_node).  INFLATOX_NUMA=0 turns the placement off.
static int device_numa_node(CudaDriver& cu, int ordinal) {
  CUdevice dev;
  if (cu.p_cuDeviceGet(&dev, ordinal) != CUDA_SUCCESS) return -1;
  char bus[64] = {0};
  if (cu.p_cuDeviceGetPCIBusId(bus, sizeof bus, dev) != CUDA_SUCCESS) return -1;
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  FILE* f = fopen((std::string("/sys/bus/pci/devices/") + bus + "/numa_node").c_str(), "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}
static bool node_cpu_set(int node, cpu_set_t* set) {
  CPU_ZERO(set);
  if (node < 0) return false;
  FILE* f = fopen(fmt("/sys/devices/system/node/node%d/cpulist", node).c_str(), "r");
  if (!f) return false;
  char buf[4096] = {0};
  const bool ok = fgets(buf, sizeof buf, f) != nullptr;
  fclose(f);
  if (!ok) return false;
  int n = 0;
  for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int a = 0, b = 0;
    const int got = sscanf(tok, "%d-%d", &a, &b);
    if (got == 1) b = a;
    if (got >= 1)
      for (int c = a; c <= b && c < CPU_SETSIZE; ++c) {
        CPU_SET(c, set);
        ++n;
      }
  }
  return n > 0;
}

inflx_status inflx_host_alloc_on(size_t bytes, const int* devices, int n_devices, void** ptr) {
  *ptr = nullptr;
  CudaDriver& cu = CudaDriver::get();
  if (!cu.ok) return fail(INFLX_ERR_CUDA, cu.error);
  std::vector<int> devs;
  if (devices && n_devices > 0)
    devs.assign(devices, devices + n_devices);
  else
    devs = default_devices();
  if (devs.empty()) devs.push_back(0);
  DeviceState* dev = nullptr;  // a context must be current for cuMemHostAlloc / cuMemHostRegister
  inflx_status st = get_device(devs[0], &dev);
  if (st) return st;
  CU_TRY(cu.p_cuCtxSetCurrent(dev->ctx));
  const char* mode = getenv("INFLATOX_HOST_ALLOC");
  const bool want_mmap = bytes >= (size_t(64) << 20) && !(mode && !strcmp(mode, "cuda"));
  if (want_mmap) {
    const size_t huge = size_t(2) << 20, len = (bytes + huge - 1) & ~(huge - 1);
    void* m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m != MAP_FAILED) {
      const bool dbg = getenv("INFLATOX_DEBUG_POOL") != nullptr;
      const auto t0 = std::chrono::steady_clock::now();
      madvise(m, len, MADV_HUGEPAGE);
      const char* numa_env = getenv("INFLATOX_NUMA");
      const bool numa = !(numa_env && !strcmp(numa_env, "0"));
      const size_t nd = devs.size();
      const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
      const unsigned per_slice = std::max<unsigned>(1, std::min(4u, hw) / (unsigned)nd);
      std::vector<std::thread> th;
      std::string placed;
      for (size_t d = 0; d < nd; ++d) {
        // slice d = the bytes device d's row shard lands in (equal split, 2 MiB granularity)
        const size_t s_lo = (len / huge * d / nd) * huge, s_hi = (len / huge * (d + 1) / nd) * huge;
        const int node = numa ? device_numa_node(cu, devs[d]) : -1;
        if (dbg) placed += fmt(" dev%d->node%d", devs[d], node);
        const size_t per = (((s_hi - s_lo) / per_slice) + huge - 1) & ~(huge - 1);
        for (unsigned k = 0; k < per_slice; ++k) {
          const size_t lo = std::min(s_hi, s_lo + k * per), hi = std::min(s_hi, lo + per);
          if (lo < hi)
            th.emplace_back([=] {
              cpu_set_t set;
              if (node_cpu_set(node, &set))  // best effort: a cpuset that excludes the node stays put
                pthread_setaffinity_np(pthread_self(), sizeof set, &set);
              char* q = static_cast<char*>(m);
              if (madvise(q + lo, hi - lo, 23 /* MADV_POPULATE_WRITE */) != 0)
                for (size_t o = lo; o < hi; o += 4096) q[o] = 0;
            });
        }
      }
      for (auto& t : th) t.join();
      const auto t1 = std::chrono::steady_clock::now();
      const CUresult reg = cu.p_cuMemHostRegister(m, len, CU_MEMHOSTREGISTER_PORTABLE);
      if (dbg)
        fprintf(stderr, "[inflx_host_alloc] %zu MiB:%s populate %.0f ms, register %.0f ms (rc %d)\n",
                len >> 20, placed.c_str(),
                std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1)
                    .count(),
                (int)reg);
      if (reg == CUDA_SUCCESS) {
        std::lock_guard<std::mutex> lk(g_host_mu);
        g_host_maps[m] = len;
        *ptr = m;
        return INFLX_OK;
      }
      munmap(m, len);  // fall through to the driver's own allocator
    }
  }
  CU_TRY(cu.p_cuMemHostAlloc(ptr, bytes ? bytes : 1, CU_MEMHOSTALLOC_PORTABLE));
  return INFLX_OK;
}
inflx_status inflx_host_alloc(size_t bytes, void** ptr) {
  return inflx_host_alloc_on(bytes, nullptr, 0, ptr);
}
// NUMA node a device's PCIe link is attached to, -1 when the platform does not say.
int inflx_device_numa_node(int device) {
  CudaDriver& cu = CudaDriver::get();
  return cu.ok ? device_numa_node(cu, device) : -1;
}
inflx_status inflx_host_free(void* ptr) {
  if (!ptr) return INFLX_OK;
  CudaDriver& cu = CudaDriver::get();
  if (!cu.ok) return fail(INFLX_ERR_CUDA, cu.error);
  size_t len = 0;
  {
    std::lock_guard<std::mutex> lk(g_host_mu);
    auto it = g_host_maps.find(ptr);
    if (it != g_host_maps.end()) {
      len = it->second;
      g_host_maps.erase(it);
    }
  }
  // may run on any thread (a numpy view's finaliser): make a context current like inflx_host_alloc
  CUcontext cur = nullptr;
  if (cu.p_cuCtxGetCurrent(&cur) != CUDA_SUCCESS || !cur) {
    DeviceState* dev = nullptr;
    std::vector<int> d = default_devices();
    inflx_status st = get_device(d.empty() ? 0 : d[0], &dev);
    if (st) return st;
    CU_TRY(cu.p_cuCtxSetCurrent(dev->ctx));
  }
  if (len) {
    CUresult r = cu.p_cuMemHostUnregister(ptr);
    if (r != CUDA_SUCCESS) {
      // the driver still holds the registration: unmapping the pages under it would leave a
      // dangling DMA mapping, so the block is kept (leaked) and the error reported
      std::lock_guard<std::mutex> lk(g_host_mu);
      g_host_maps[ptr] = len;
      CU_TRY(r);
    }
    munmap(ptr, len);
    return INFLX_OK;
  }
  CU_TRY(cu.p_cuMemFreeHost(ptr));
  return INFLX_OK;
}

}  // extern "C"
