// Fused AMSGrad step with the reference's exact (non-standard) update, lib/networks/optimizers.py:53-74:
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; vmax = max(vmax, v)
//   denom = sqrt(vmax or v) / bc2 + eps ;  p -= wd * p + lr * (m / bc1) / denom
// One elementwise launch per tensor (the decoder is ONE arena tensor); HBM-bound: 20 B read + 16 B
// written per element with AMSGrad.
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 float* __restrict__ vmax, long long n, float lr, float b1, float b2, float eps, float wd, float bc1,
                 float bc2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float vm = vi;
    if (vmax) {
      vm = fmaxf(vmax[i], vi);
      vmax[i] = vm;
    }
    const float denom = sqrtf(vm) / bc2 + eps;
    const float pi = p[i];
    p[i] = pi - (pi * wd + lr * ((mi / bc1) / denom));
  }
}
}  // namespace

// vmax may be NULL (amsgrad off).  bc1 = 1 - beta1^t, bc2 = sqrt(1 - beta2^t).
DPF_API int dpf_adam_step(float* p, const float* g, float* m, float* v, float* vmax, long long n, float lr, float b1,
                          float b2, float eps, float wd, float bc1, float bc2, void* stream) {
  DPF_REQUIRE(n >= 0, DPF_ERR_BAD_ARG, "dpf_adam_step: negative size");
  if (n == 0) return DPF_OK;
  DPF_REQUIRE(p && g && m && v, DPF_ERR_NULL_PTR, "dpf_adam_step: null pointer");
  const int grid = (int)min((long long)dpf_num_sms() * 8, (n + 255) / 256);
  adam_step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, vmax, n, lr, b1, b2, eps, wd, bc1, bc2);
  return dpf_check_launch("adam_step_kernel");
}
