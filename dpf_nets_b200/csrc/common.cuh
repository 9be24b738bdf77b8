// Shared helpers for the dpf-nets B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DPF_API extern "C" __attribute__((visibility("default")))

// Error codes returned by every C-ABI entry point (0 = ok, >0 = cudaError_t, <0 = argument error).
#define DPF_OK 0
#define DPF_ERR_BAD_ARG (-1)
#define DPF_ERR_NULL_PTR (-2)
#define DPF_ERR_UNSUPPORTED (-3)
#define DPF_ERR_ALIGN (-4)
#define DPF_ERR_BARRIER (-5)   /* a grid barrier of an earlier decoder pass timed out; that pass was aborted */

void dpf_set_error(const char* fmt, ...);

extern long long g_dpf_launches;   // kernels launched by this library (api.cu)

static inline int dpf_check_launch(const char* what) {
  ++g_dpf_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    dpf_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return DPF_OK;
}

#define DPF_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      dpf_set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int dpf_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Consecutive per-layer kernels of one pass are launched with programmatic stream serialization:
// a kernel's CTAs may start (barrier init, TMEM allocation, TMA weight loads) while the previous
// kernel drains; pdl_wait() blocks until the previous grid has completed and its writes are
// visible, and MUST precede the first read of anything that grid produced.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

extern int g_dpf_pdl;   // decoder.cu: dpf_set_option(2, v)

template <typename... KArgs, typename... Args>
static inline void dpf_launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_dpf_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
