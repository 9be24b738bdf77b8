// Library-level C-ABI plumbing: error text, version, device probing.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";
long long g_dpf_launches = 0;

void dpf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

DPF_API const char* dpf_last_error(void) { return g_err; }

DPF_API int dpf_version(void) { return 100; }

// Returns 0 when the current device can run this library (compute capability 10.x).
DPF_API int dpf_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    dpf_set_error("dpf_device_check: %s", cudaGetErrorString(e));
    return (int)e;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    dpf_set_error("dpf_device_check: device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
    return DPF_ERR_UNSUPPORTED;
  }
  return DPF_OK;
}

// Number of kernels this library has launched since load (bench.py's gpu_launches evidence).
DPF_API int dpf_launch_count(long long* count) {
  if (!count) { dpf_set_error("dpf_launch_count: null pointer"); return DPF_ERR_NULL_PTR; }
  *count = g_dpf_launches;
  return DPF_OK;
}
