// Library-level entry points: version, error string, device checks.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace fosvos {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FOSVOS_ERR_LAUNCH;
  }
  return FOSVOS_OK;
}

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return dev < 0 ? 0 : (dev > 63 ? 63 : dev);
}

// SM count of the CURRENT device (a process may drive several GPUs: cached per device)
int num_sms() {
  static int cached[64] = {0};
  const int slot = device_slot();
  if (cached[slot] > 0) return cached[slot];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, slot) != cudaSuccess || n <= 0) { cudaGetLastError(); return 148; }
  cached[slot] = n;
  return n;
}

}  // namespace fosvos

extern "C" {

int fosvos_abi_version(void) { return 1; }

const char* fosvos_last_error(void) { return fosvos::g_err; }

int fosvos_device_check(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    fosvos::set_error("no CUDA device (%s); fosvos_b200 has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    return FOSVOS_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    fosvos::set_error("device %d out of range [0,%d)", device, count);
    return FOSVOS_ERR_BAD_ARG;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (major != 10) {
    fosvos::set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
    return FOSVOS_ERR_ARCH;
  }
  return FOSVOS_OK;
}

int fosvos_num_sms(int device) {
  int rc = fosvos_device_check(device);
  if (rc != FOSVOS_OK) return rc;
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  return n;
}

}  // extern "C"
