// Latent-side (per-shape) building blocks of the DPF-Nets models, fused: the shape-latent flows and feature heads work on
// (B, F) matrices with B = 32..64 rows, where the reference's module chains (lib/networks/flows.py:163-213 RealNVPFlow,
// lib/networks/encoders.py:31-83 FeatureEncoder) turn into ~20 forward and ~45 backward kernels per coupling layer, each
// a few microseconds of launch-bound work.  Two fused pairs replace the elementwise / reduction part:
//
//   bn_swish        y = swish(BatchNorm1d(x))           nn.BatchNorm1d over the batch dimension (batch statistics in
//                                                       training, running statistics otherwise) + Swish (layers.py:5-10)
//   latent_affine   logvar = log(eps + exp(raw_lv)), scatter of (mu, logvar) to the warped positions,
//                   g_out = exp(+-logvar/2) g (+-) mu   (flows.py:196-211)
//
// One thread owns one column (feature) and walks over the B rows: every reduction is a per-thread loop, accesses are
// coalesced across the threads of a warp.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

__global__ void __launch_bounds__(128)
bn_swish_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ rm, float* __restrict__ rv, int B, int F, float eps, float momentum, int training,
                    float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_istd) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= F) return;
  float mean, var;
  if (training) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += x[(size_t)b * F + c];
    mean = s / (float)B;
    float q = 0.f;
    for (int b = 0; b < B; ++b) {
      const float d = x[(size_t)b * F + c] - mean;
      q = fmaf(d, d, q);
    }
    var = q / (float)B;
    if (rm) {
      rm[c] = (1.f - momentum) * rm[c] + momentum * mean;
      rv[c] = (1.f - momentum) * rv[c] + momentum * var * ((float)B / (float)max(B - 1, 1));
    }
  } else {
    mean = rm[c];
    var = rv[c];
  }
  const float istd = 1.f / sqrtf(var + eps);
  const float g = gamma[c], bt = beta[c];
  for (int b = 0; b < B; ++b) {
    const float z = fmaf((x[(size_t)b * F + c] - mean) * istd, g, bt);
    y[(size_t)b * F + c] = z * sigmoidf_(z);
  }
  save_mean[c] = mean;
  save_istd[c] = istd;
}

__global__ void __launch_bounds__(128)
bn_swish_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ save_mean, const float* __restrict__ save_istd,
                    int B, int F, int training, float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= F) return;
  const float mean = save_mean[c], istd = save_istd[c], g = gamma[c], bt = beta[c];
  float dg = 0.f, db = 0.f;
  for (int b = 0; b < B; ++b) {
    const float xh = (x[(size_t)b * F + c] - mean) * istd;
    const float z = fmaf(xh, g, bt);
    const float s = sigmoidf_(z);
    const float dz = dy[(size_t)b * F + c] * (s + z * s * (1.f - s));
    dg = fmaf(dz, xh, dg);
    db += dz;
  }
  dgamma[c] = dg;
  dbeta[c] = db;
  // d xhat = dz * gamma;  training: dx = istd (d xhat - mean_b(d xhat) - xhat mean_b(d xhat xhat))
  const float m1 = training ? g * db / (float)B : 0.f;
  const float m2 = training ? g * dg / (float)B : 0.f;
  for (int b = 0; b < B; ++b) {
    const float xh = (x[(size_t)b * F + c] - mean) * istd;
    const float z = fmaf(xh, g, bt);
    const float s = sigmoidf_(z);
    const float dz = dy[(size_t)b * F + c] * (s + z * s * (1.f - s));
    dx[(size_t)b * F + c] = istd * (dz * g - m1 - xh * m2);
  }
}

// pos[j] = index of latent position j in the layer's warp list, or -1 for a kept position
__global__ void __launch_bounds__(256)
latent_affine_fwd_kernel(const float* __restrict__ g, const float* __restrict__ raw_mu, const float* __restrict__ raw_lv,
                         const int* __restrict__ pos, int B, int G, int W, float eps, int inverse,
                         float* __restrict__ g_out, float* __restrict__ mu, float* __restrict__ lv) {
  const long long total = (long long)B * G;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int b = (int)(e / G), j = (int)(e - (long long)b * G);
    const int p = pos[j];
    float m = 0.f, l = 0.f;
    if (p >= 0) {
      m = raw_mu[(size_t)b * W + p];
      l = logf(eps + expf(raw_lv[(size_t)b * W + p]));
    }
    const float gv = g[e];
    mu[e] = m;
    lv[e] = l;
    g_out[e] = inverse ? expf(-0.5f * l) * (gv - m) : fmaf(expf(0.5f * l), gv, m);
  }
}

// cotangents dgo (of g_out), dmu_f, dlv_f (of the full-width mu / logvar outputs; nullable) -> dg (B,G), draw_mu, draw_lv (B,W)
__global__ void __launch_bounds__(256)
latent_affine_bwd_kernel(const float* __restrict__ dgo, const float* __restrict__ dmu_f, const float* __restrict__ dlv_f,
                         const float* __restrict__ g, const float* __restrict__ raw_mu, const float* __restrict__ raw_lv,
                         const int* __restrict__ pos, int B, int G, int W, float eps, int inverse,
                         float* __restrict__ dg, float* __restrict__ draw_mu, float* __restrict__ draw_lv) {
  const long long total = (long long)B * G;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int b = (int)(e / G), j = (int)(e - (long long)b * G);
    const int p = pos[j];
    const float d = dgo ? dgo[e] : 0.f;
    if (p < 0) {
      dg[e] = d;                       // kept position: mu = logvar = 0, g_out = g
      continue;
    }
    const float raw = raw_lv[(size_t)b * W + p];
    const float ex = expf(raw);
    const float l = logf(eps + ex);
    const float m = raw_mu[(size_t)b * W + p];
    const float gv = g[e];
    float dm = dmu_f ? dmu_f[e] : 0.f, dl = dlv_f ? dlv_f[e] : 0.f, dgv;
    if (inverse) {
      const float sc = expf(-0.5f * l);
      dgv = d * sc;
      dm -= d * sc;
      dl -= 0.5f * d * sc * (gv - m);
    } else {
      const float sc = expf(0.5f * l);
      dgv = d * sc;
      dm += d;
      dl += 0.5f * d * sc * gv;
    }
    dg[e] = dgv;
    draw_mu[(size_t)b * W + p] = dm;
    draw_lv[(size_t)b * W + p] = dl * ex / (eps + ex);
  }
}

}  // namespace

// y = swish(BatchNorm1d(x)) for x (B,F) fp32; training != 0: batch statistics (biased variance) and, when rm / rv are given,
// the running-statistics update (momentum, unbiased variance); otherwise rm / rv are the statistics used.
// save_mean / save_istd (F) are what the backward needs.
DPF_API int dpf_bn_swish_forward(const float* x, const float* gamma, const float* beta, float* rm, float* rv, int B, int F,
                                 float eps, float momentum, int training, float* y, float* save_mean, float* save_istd,
                                 void* stream) {
  DPF_REQUIRE(x && gamma && beta && y && save_mean && save_istd, DPF_ERR_NULL_PTR, "dpf_bn_swish_forward: null pointer");
  DPF_REQUIRE(training || (rm && rv), DPF_ERR_NULL_PTR, "dpf_bn_swish_forward: eval mode needs running statistics");
  DPF_REQUIRE(B > 0 && F > 0, DPF_ERR_BAD_ARG, "dpf_bn_swish_forward: bad sizes");
  bn_swish_fwd_kernel<<<(F + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, gamma, beta, rm, rv, B, F, eps, momentum, training, y,
                                                                        save_mean, save_istd);
  return dpf_check_launch("bn_swish_fwd_kernel");
}

DPF_API int dpf_bn_swish_backward(const float* dy, const float* x, const float* gamma, const float* beta, const float* save_mean,
                                  const float* save_istd, int B, int F, int training, float* dx, float* dgamma, float* dbeta,
                                  void* stream) {
  DPF_REQUIRE(dy && x && gamma && beta && save_mean && save_istd && dx && dgamma && dbeta, DPF_ERR_NULL_PTR,
              "dpf_bn_swish_backward: null pointer");
  DPF_REQUIRE(B > 0 && F > 0, DPF_ERR_BAD_ARG, "dpf_bn_swish_backward: bad sizes");
  bn_swish_bwd_kernel<<<(F + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dy, x, gamma, beta, save_mean, save_istd, B, F, training, dx,
                                                                        dgamma, dbeta);
  return dpf_check_launch("bn_swish_bwd_kernel");
}

// RealNVPFlow's transform (flows.py:196-211): g (B,G), raw_mu / raw_lv (B,W) = the two nets' outputs for the W warped
// positions, pos (G) int32 = index into the warp list or -1; inverse != 0: g_out = exp(-logvar/2) (g - mu), else
// g_out = exp(logvar/2) g + mu; mu / lv (B,G) are the full-width outputs (zero at kept positions).
DPF_API int dpf_latent_affine_forward(const float* g, const float* raw_mu, const float* raw_lv, const int* pos, int B, int G, int W,
                                      float eps, int inverse, float* g_out, float* mu, float* lv, void* stream) {
  DPF_REQUIRE(g && raw_mu && raw_lv && pos && g_out && mu && lv, DPF_ERR_NULL_PTR, "dpf_latent_affine_forward: null pointer");
  DPF_REQUIRE(B > 0 && G > 0 && W > 0 && W <= G, DPF_ERR_BAD_ARG, "dpf_latent_affine_forward: bad sizes");
  const int grid = (int)min((long long)dpf_num_sms() * 4, ((long long)B * G + 255) / 256);
  latent_affine_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, raw_mu, raw_lv, pos, B, G, W, eps, inverse, g_out, mu, lv);
  return dpf_check_launch("latent_affine_fwd_kernel");
}

// dgo / dmu_f / dlv_f: cotangents of g_out / mu / lv (each nullable = zero) -> dg (B,G), draw_mu, draw_lv (B,W) (every
// element written).
DPF_API int dpf_latent_affine_backward(const float* dgo, const float* dmu_f, const float* dlv_f, const float* g, const float* raw_mu,
                                       const float* raw_lv, const int* pos, int B, int G, int W, float eps, int inverse, float* dg,
                                       float* draw_mu, float* draw_lv, void* stream) {
  DPF_REQUIRE(g && raw_mu && raw_lv && pos && dg && draw_mu && draw_lv, DPF_ERR_NULL_PTR, "dpf_latent_affine_backward: null pointer");
  DPF_REQUIRE(B > 0 && G > 0 && W > 0 && W <= G, DPF_ERR_BAD_ARG, "dpf_latent_affine_backward: bad sizes");
  const int grid = (int)min((long long)dpf_num_sms() * 4, ((long long)B * G + 255) / 256);
  latent_affine_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dgo, dmu_f, dlv_f, g, raw_mu, raw_lv, pos, B, G, W, eps, inverse, dg,
                                                                  draw_mu, draw_lv);
  return dpf_check_launch("latent_affine_bwd_kernel");
}
