// Error plumbing, device check and version for the srgd_b200 C-ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace srgd {

static thread_local char g_err[512] = "";
thread_local long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail_cuda(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return SRGD_E_CUDA;
}

static int g_sm_count = 0;

int check_device() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = SRGD_E_DEVICE;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no usable CUDA device (%s); srgd_b200 has no CPU fallback", cudaGetErrorString(e));
    return SRGD_E_DEVICE;
  }
  if (dev == cached_dev) {
    if (cached_rc != SRGD_OK) set_error("device %d is not sm_100 (B200); srgd_b200 has no fallback path", dev);
    return cached_rc;
  }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceProperties");
  cached_dev = dev;
  if (p.major != 10) {
    set_error("device %d (%s, sm_%d%d) is not sm_100 (B200); srgd_b200 has no fallback path", dev, p.name,
              p.major, p.minor);
    cached_rc = SRGD_E_DEVICE;
  } else {
    cached_rc = SRGD_OK;
    g_sm_count = p.multiProcessorCount;
  }
  return cached_rc;
}

int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

}  // namespace srgd

extern "C" {

int srgd_version(void) { return SRGD_B200_VERSION; }

const char* srgd_last_error(void) { return srgd::g_err; }

int srgd_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    srgd::set_error("no usable CUDA device %d (%s); srgd_b200 has no CPU fallback", device,
                    e == cudaSuccess ? "device index out of range" : cudaGetErrorString(e));
    return SRGD_E_DEVICE;
  }
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(device);
  int rc = srgd::check_device();
  cudaSetDevice(cur);
  return rc;
}

}  // extern "C"
