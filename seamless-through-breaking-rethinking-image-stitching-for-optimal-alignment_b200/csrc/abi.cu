// abi.cu — library-wide state of libstitchb200: error string, device check,
// launch counter.  No kernels here.
#include <atomic>
#include <mutex>
#include <string.h>

#include "common.cuh"

namespace sb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<int> g_tune[SB_TUNE_COUNT];   // 0 = "use the built-in default"
int tune_get(int key, int dflt) {
  if (key < 0 || key >= SB_TUNE_COUNT) return dflt;
  const int v = g_tune[key].load(std::memory_order_relaxed);
  return v ? v : dflt;
}

// Word written (system scope) just before a protocol-timeout trap. It lives in
// mapped pinned host memory so the host can still read it after the context died.
// Portable: every device of the process (one host thread per GPU under nn.DataParallel) sees the
// same mapped word; under unified addressing its device pointer is the same on all of them.
static unsigned int* g_dbg_host = nullptr;
static std::atomic<unsigned int*> g_dbg{nullptr};
static std::mutex g_dbg_mutex;
unsigned int* debug_word_device() {
  unsigned int* d = g_dbg.load(std::memory_order_acquire);
  if (d) return d;
  std::lock_guard<std::mutex> lock(g_dbg_mutex);
  d = g_dbg.load(std::memory_order_relaxed);
  if (d) return d;
  unsigned int* h = nullptr;
  cudaError_t e = cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e == cudaSuccess) {
    *h = 0;
    e = cudaHostGetDevicePointer(&d, h, 0);
    if (e != cudaSuccess) cudaFreeHost(h);
  }
  if (e != cudaSuccess) {
    set_error("debug word allocation failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  g_dbg_host = h;
  g_dbg.store(d, std::memory_order_release);
  return d;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_device() {
  // Cached per device ordinal; a process drives one GPU (one process per GPU).
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = SB_OK;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s (no CUDA device? this library has no CPU fallback)",
              cudaGetErrorString(e));
    return SB_ECUDA;
  }
  if (dev == cached_dev) {
    if (cached_rc != SB_OK) set_error("device %d is not sm_100 (B200)", dev);
    return cached_rc;
  }
  int major = 0, minor = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
    return SB_ECUDA;
  }
  cached_dev = dev;
  cached_rc = (major == 10) ? SB_OK : SB_EARCH;
  if (cached_rc != SB_OK)
    set_error("device %d is sm_%d%d; libstitchb200 only runs on sm_100 (B200)", dev, major, minor);
  return cached_rc;
}

}  // namespace sb

extern "C" {

int sb_version(void) { return SB_VERSION; }
const char* sb_last_error(void) { return sb::g_err; }
int sb_device_check(void) { return sb::check_device(); }
long long sb_launch_count(void) { return sb::g_launches.load(std::memory_order_relaxed); }
unsigned int sb_debug_word(void) { return sb::g_dbg_host ? *sb::g_dbg_host : 0u; }
int sb_tune(int key, int value) {
  if (key < 0 || key >= SB_TUNE_COUNT) {
    sb::set_error("sb_tune: unknown key %d", key);
    return SB_EINVAL;
  }
  sb::g_tune[key].store(value, std::memory_order_relaxed);
  return SB_OK;
}
// Pinned host staging buffers for the host-resident data path (pipeline.StreamedHotPath).
// write_combined != 0: cudaHostAllocWriteCombined — not snooped by the CPU caches, faster for the
// device to read over PCIe; meant for buffers the CPU only ever fills sequentially.
void* sb_host_alloc(size_t bytes, int write_combined) {
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0));
  if (e != cudaSuccess) {
    sb::set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}
int sb_host_free(void* p) {
  if (!p) return SB_OK;
  cudaError_t e = cudaFreeHost(p);
  if (e != cudaSuccess) {
    sb::set_error("cudaFreeHost failed: %s", cudaGetErrorString(e));
    return SB_ECUDA;
  }
  return SB_OK;
}
void sb_reset_launch_count(void) { sb::g_launches.store(0, std::memory_order_relaxed); }
// Releases what the library itself holds (the 64-byte mapped debug word) and resets the tuning knobs and
// the launch counter.  Outputs, workspaces and sb_host_alloc blocks belong to the caller.  The library can be
// used again afterwards (state is re-created on demand).
int sb_shutdown(void) {
  for (int k = 0; k < SB_TUNE_COUNT; ++k) sb::g_tune[k].store(0, std::memory_order_relaxed);
  sb::g_launches.store(0, std::memory_order_relaxed);
  std::lock_guard<std::mutex> lock(sb::g_dbg_mutex);
  if (sb::g_dbg_host) {
    cudaError_t e = cudaFreeHost(sb::g_dbg_host);
    sb::g_dbg_host = nullptr;
    sb::g_dbg.store(nullptr, std::memory_order_release);
    if (e != cudaSuccess) {
      sb::set_error("sb_shutdown: cudaFreeHost failed: %s", cudaGetErrorString(e));
      return SB_ECUDA;
    }
  }
  return SB_OK;
}

}  // extern "C"
