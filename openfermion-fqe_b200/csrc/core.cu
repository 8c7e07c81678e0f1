// Error reporting, device probing and the launch counter of libfqe_b200.so.
#include "fqeb_common.cuh"

#include <string.h>

namespace fqeb {

static thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libfqe_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return FQEB_ERR_NODEVICE;
  }
  return FQEB_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Small host -> device uploads (operator tables) go through an internal non-blocking stream
// and a stream-ordered allocation, so that preparing the next operator neither waits for nor
// stalls the sigma build that is still running on the caller's stream (a plain cudaMemcpy
// synchronises with the legacy default stream, cudaMalloc / cudaFree with the whole device).
static cudaStream_t upload_stream() {
  static cudaStream_t streams[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!streams[dev] &&
      cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess)
    streams[dev] = nullptr;
  return streams[dev];
}

int upload_alloc(void **d_ptr, const void *h_src, size_t bytes) {
  cudaStream_t st = upload_stream();
  FQEB_REQUIRE(st != nullptr, "cannot create the upload stream");
  *d_ptr = nullptr;
  FQEB_CUDA(cudaMallocAsync(d_ptr, bytes, st));
  if (h_src) FQEB_CUDA(cudaMemcpyAsync(*d_ptr, h_src, bytes, cudaMemcpyHostToDevice, st));
  return FQEB_OK;
}

int upload_finish() {
  cudaStream_t st = upload_stream();
  FQEB_REQUIRE(st != nullptr, "cannot create the upload stream");
  FQEB_CUDA(cudaStreamSynchronize(st));
  return FQEB_OK;
}

}  // namespace fqeb

extern "C" const char *fqeb_last_error(void) { return fqeb::g_err; }
extern "C" int fqeb_version(void) { return 100; }
extern "C" int fqeb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
extern "C" int fqeb_set_device(int device) {
  int rc = fqeb::require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_CUDA(cudaSetDevice(device));
  return FQEB_OK;
}
extern "C" uint64_t fqeb_launch_count(void) {
  return fqeb::g_launches.load(std::memory_order_relaxed);
}
