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
// Multi-tensor form: up to ADAM_CHUNK tensors per launch, pointers passed by value in the kernel
// arguments; block -> (tensor, 4096-element slab) through a prefix table.
constexpr int ADAM_CHUNK = 48;
constexpr int ADAM_SLAB = 4096;
struct AdamTensors {
  float* p[ADAM_CHUNK];
  const float* g[ADAM_CHUNK];
  float* m[ADAM_CHUNK];
  float* v[ADAM_CHUNK];
  float* vmax[ADAM_CHUNK];
  int numel[ADAM_CHUNK];
  int block_start[ADAM_CHUNK + 1];
  int count;
};

// hyper (nullable): device array {lr, b1, b2, eps, wd, bc1, bc2} that overrides the by-value scalars, so that a
// CUDA graph holding this launch can be replayed with a new learning rate / step count (dpf_adam_step_multi_dev)
__global__ void __launch_bounds__(256)
adam_step_multi_kernel(const __grid_constant__ AdamTensors t, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2,
                       const float* __restrict__ hyper) {
  if (hyper) {
    lr = hyper[0]; b1 = hyper[1]; b2 = hyper[2]; eps = hyper[3]; wd = hyper[4]; bc1 = hyper[5]; bc2 = hyper[6];
  }
  int k = 0;
  while (k + 1 < t.count && (int)blockIdx.x >= t.block_start[k + 1]) ++k;
  const int base = ((int)blockIdx.x - t.block_start[k]) * ADAM_SLAB;
  const int end = min(t.numel[k], base + ADAM_SLAB);
  float* __restrict__ p = t.p[k];
  const float* __restrict__ g = t.g[k];
  float* __restrict__ m = t.m[k];
  float* __restrict__ v = t.v[k];
  float* __restrict__ vmax = t.vmax[k];
  auto upd = [&](float gi, float& mi, float& vi, float& vm, float& pi) {
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    vm = vmax ? fmaxf(vm, vi) : vi;
    const float denom = sqrtf(vm) / bc2 + eps;
    pi = pi - (pi * wd + lr * ((mi / bc1) / denom));
  };
  // 16-byte accesses when the five arrays allow it (the slab start is a multiple of 4 elements): the update is pure
  // streaming, 9 arrays' worth of traffic per element
  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)vmax) & 15) == 0);
  int i = base + 4 * threadIdx.x;
  if (vec) {
    for (; i + 4 <= end; i += 4 * 256) {
      const float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i), p4 = *reinterpret_cast<const float4*>(p + i);
      float4 x4 = vmax ? *reinterpret_cast<const float4*>(vmax + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      upd(g4.x, m4.x, v4.x, x4.x, p4.x);
      upd(g4.y, m4.y, v4.y, x4.y, p4.y);
      upd(g4.z, m4.z, v4.z, x4.z, p4.z);
      upd(g4.w, m4.w, v4.w, x4.w, p4.w);
      *reinterpret_cast<float4*>(m + i) = m4;
      *reinterpret_cast<float4*>(v + i) = v4;
      if (vmax) *reinterpret_cast<float4*>(vmax + i) = x4;
      *reinterpret_cast<float4*>(p + i) = p4;
    }
    // the (at most 3) elements after the last full group of four: handled by the thread whose group straddles `end`
    if (i < end) {
      for (int e = i; e < end; ++e) {
        float mi = m[e], vi = v[e], vm = vmax ? vmax[e] : 0.f, pi = p[e];
        upd(g[e], mi, vi, vm, pi);
        m[e] = mi; v[e] = vi; if (vmax) vmax[e] = vm; p[e] = pi;
      }
    }
  } else {
    for (int e = base + threadIdx.x; e < end; e += 256) {
      float mi = m[e], vi = v[e], vm = vmax ? vmax[e] : 0.f, pi = p[e];
      upd(g[e], mi, vi, vm, pi);
      m[e] = mi; v[e] = vi; if (vmax) vmax[e] = vm; p[e] = pi;
    }
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

// The same update for n tensors in ceil(n / 48) launches (tensors < 2^31 elements; larger ones go through
// dpf_adam_step).  p/g/m/v/vmax: host arrays of n device pointers (vmax may be NULL = amsgrad off);
// numel: host array of n sizes.  All tensors share the hyper-parameters and the step count.
static int adam_multi_impl(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                           float* const* vmax, const long long* numel, float lr, float b1, float b2, float eps,
                           float wd, float bc1, float bc2, const float* hyper, void* stream) {
  DPF_REQUIRE(n >= 0, DPF_ERR_BAD_ARG, "dpf_adam_step_multi: negative count");
  if (n == 0) return DPF_OK;
  DPF_REQUIRE(p && g && m && v && numel, DPF_ERR_NULL_PTR, "dpf_adam_step_multi: null table");
  for (int i0 = 0; i0 < n;) {
    AdamTensors t{};
    int blocks = 0, c = 0;
    for (; i0 < n && c < ADAM_CHUNK; ++i0) {
      if (numel[i0] == 0) continue;
      DPF_REQUIRE(numel[i0] > 0 && numel[i0] < (1LL << 31) - ADAM_SLAB, DPF_ERR_BAD_ARG, "dpf_adam_step_multi: tensor %d has %lld elements", i0, numel[i0]);
      DPF_REQUIRE(p[i0] && g[i0] && m[i0] && v[i0], DPF_ERR_NULL_PTR, "dpf_adam_step_multi: null pointer in tensor %d", i0);
      t.p[c] = p[i0]; t.g[c] = g[i0]; t.m[c] = m[i0]; t.v[c] = v[i0];
      t.vmax[c] = vmax ? vmax[i0] : nullptr;
      t.numel[c] = (int)numel[i0];
      t.block_start[c] = blocks;
      blocks += (int)((numel[i0] + ADAM_SLAB - 1) / ADAM_SLAB);
      ++c;
    }
    if (c == 0) break;
    t.block_start[c] = blocks;
    t.count = c;
    adam_step_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(t, lr, b1, b2, eps, wd, bc1, bc2, hyper);
    int rc = dpf_check_launch("adam_step_multi_kernel");
    if (rc) return rc;
  }
  return DPF_OK;
}

DPF_API int dpf_adam_step_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                                float* const* vmax, const long long* numel, float lr, float b1, float b2, float eps,
                                float wd, float bc1, float bc2, void* stream) {
  return adam_multi_impl(n, p, g, m, v, vmax, numel, lr, b1, b2, eps, wd, bc1, bc2, nullptr, stream);
}

// Graph-replayable form: the hyper-parameters {lr, b1, b2, eps, wd, bc1, bc2} are read from DEVICE memory at run time.
DPF_API int dpf_adam_step_multi_dev(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                                    float* const* vmax, const long long* numel, const float* hyper_dev, void* stream) {
  DPF_REQUIRE(hyper_dev, DPF_ERR_NULL_PTR, "dpf_adam_step_multi_dev: null hyper-parameter pointer");
  return adam_multi_impl(n, p, g, m, v, vmax, numel, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 1.f, hyper_dev, stream);
}
