// Hardware probes used by bench.py to MEASURE the ceilings its roofline fractions are quoted against
// (instead of assuming them): special-function-unit (MUFU ex2) throughput for the EMD kernels, packed fp32x2 FMA
// throughput for the Chamfer kernels.  Not on any product path.
#include "common.cuh"

namespace {

// 8 independent ex2.approx chains per thread; `iters` rounds -> 8 * iters MUFU results per thread
__global__ void __launch_bounds__(256)
mufu_probe_kernel(int iters, float seed, float* __restrict__ out) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + 1e-3f * (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float r;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v[i]));
      v[i] = r - 1.0f;            // keeps the argument small; the FADD shares no pipe with the MUFU
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 123.456f) out[0] = s;   // never true: keeps the chains alive
}

// 8 independent packed FFMA2 chains per thread -> 16 fp32 FMA lanes per round per thread
__global__ void __launch_bounds__(256)
ffma2_probe_kernel(int iters, float seed, float* __restrict__ out) {
  unsigned long long v[8], a, b;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(1.0f - 1e-7f * seed));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(1e-9f * seed));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float x = seed + (float)(threadIdx.x + i);
    asm("mov.b64 %0, {%1, %1};" : "=l"(v[i]) : "f"(x));
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v[i]) : "l"(v[i]), "l"(a), "l"(b));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
    s += lo + hi;
  }
  if (s == 123.456f) out[0] = s;
}

}  // namespace

// Launches `ctas` CTAs of 256 threads running `iters` rounds of 8 independent operations per thread:
// which = 0: ex2.approx (MUFU) -> 8 results per round per thread; which = 1: fma.rn.f32x2 -> 16 fp32 FMA lanes per round
// per thread.  The caller times the launch with CUDA events; *ops_per_launch = operations (MUFU results / FMA lanes) it executes.
DPF_API int dpf_throughput_probe(int which, int ctas, int iters, float* scratch, long long* ops_per_launch, void* stream) {
  DPF_REQUIRE(which == 0 || which == 1, DPF_ERR_BAD_ARG, "dpf_throughput_probe: which must be 0 (ex2) or 1 (fma.f32x2)");
  DPF_REQUIRE(ctas > 0 && iters > 0 && scratch && ops_per_launch, DPF_ERR_BAD_ARG, "dpf_throughput_probe: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (which == 0) mufu_probe_kernel<<<ctas, 256, 0, s>>>(iters, 0.25f, scratch);
  else ffma2_probe_kernel<<<ctas, 256, 0, s>>>(iters, 0.25f, scratch);
  *ops_per_launch = (long long)ctas * 256 * iters * (which == 0 ? 8 : 16);
  return dpf_check_launch("throughput_probe_kernel");
}
