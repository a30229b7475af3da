// Error plumbing, device check and version for the srgd_b200 C-ABI.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

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
int g_batch_invariant = 0;

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

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("SRGD_PDL");
    cached = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return cached != 0;
}

// ---- profiling ---------------------------------------------------------------------------------
bool g_prof_on = false;
struct ProfRec {
  int kind;
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static std::vector<ProfRec> g_prof_recs;
struct ProfDone {
  int kind;
  double flops, bytes, ms;
};
static std::vector<ProfDone> g_prof_done;
static double g_prof_ms[SRGD_PK_COUNT], g_prof_flops[SRGD_PK_COUNT], g_prof_bytes[SRGD_PK_COUNT];
static int g_prof_n[SRGD_PK_COUNT];

void prof_open(int kind, double flops, double bytes, cudaStream_t st) {
  ProfRec r{kind, flops, bytes, nullptr, nullptr};
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  g_prof_recs.push_back(r);
}
void prof_close(cudaStream_t st) {
  if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().e1, st);
}

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

long long srgd_launch_count(void) { return (long long)srgd::g_launches; }

int srgd_set_batch_invariant(int on) {
  const int prev = srgd::g_batch_invariant;
  srgd::g_batch_invariant = on ? 1 : 0;
  return prev;
}

int srgd_profile_begin(void) {
  for (auto& r : srgd::g_prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  srgd::g_prof_recs.clear();
  srgd::g_prof_done.clear();
  for (int k = 0; k < SRGD_PK_COUNT; ++k) {
    srgd::g_prof_ms[k] = srgd::g_prof_flops[k] = srgd::g_prof_bytes[k] = 0.0;
    srgd::g_prof_n[k] = 0;
  }
  srgd::g_prof_on = true;
  return SRGD_OK;
}

int srgd_profile_end(void) {
  srgd::g_prof_on = false;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return srgd::fail_cuda(e, "cudaDeviceSynchronize");
  for (auto& r : srgd::g_prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess && r.kind >= 0 && r.kind < SRGD_PK_COUNT) {
      srgd::g_prof_ms[r.kind] += ms;
      srgd::g_prof_flops[r.kind] += r.flops;
      srgd::g_prof_bytes[r.kind] += r.bytes;
      srgd::g_prof_n[r.kind] += 1;
      srgd::g_prof_done.push_back({r.kind, r.flops, r.bytes, (double)ms});
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  srgd::g_prof_recs.clear();
  return SRGD_OK;
}

int srgd_profile_get(int kind, double* ms, double* flops, double* bytes, int* launches) {
  if (kind < 0 || kind >= SRGD_PK_COUNT) {
    srgd::set_error("profile_get: bad kind %d", kind);
    return SRGD_E_ARG;
  }
  if (ms) *ms = srgd::g_prof_ms[kind];
  if (flops) *flops = srgd::g_prof_flops[kind];
  if (bytes) *bytes = srgd::g_prof_bytes[kind];
  if (launches) *launches = srgd::g_prof_n[kind];
  return SRGD_OK;
}

int srgd_profile_record_count(void) { return (int)srgd::g_prof_done.size(); }

int srgd_profile_record(int index, int* kind, double* ms, double* flops, double* bytes) {
  if (index < 0 || index >= (int)srgd::g_prof_done.size()) {
    srgd::set_error("profile_record: index %d out of range", index);
    return SRGD_E_ARG;
  }
  const srgd::ProfDone& d = srgd::g_prof_done[index];
  if (kind) *kind = d.kind;
  if (ms) *ms = d.ms;
  if (flops) *flops = d.flops;
  if (bytes) *bytes = d.bytes;
  return SRGD_OK;
}

}  // extern "C"
