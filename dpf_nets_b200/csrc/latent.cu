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
// bn_swish: a CTA covers 32 columns x 8 row groups, every thread keeps its rows in registers (one coalesced read of the
// matrix, independent loads), the batch reductions are small shared-memory trees.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

// Thread layout: a CTA covers 32 columns (threadIdx.x) with 8 row groups (threadIdx.y); a thread keeps its rows
// ty, ty + 8, ... (up to BN_MAXR of them) in registers, so the matrix is read once with independent coalesced loads and
// the two batch reductions go through a small shared-memory tree.
constexpr int BN_TY = 8;       // row groups; the kernels are instantiated for 4 / 8 / 32 rows per thread (B <= 32 / 64 / 256)

__device__ __forceinline__ float col_sum(float v, float (*red)[33]) {      // sum over the 8 row groups of a column
  __syncthreads();
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < BN_TY; ++i) s += red[i][threadIdx.x];
  return s;
}

template <int BN_MAXR>
__global__ void __launch_bounds__(32 * BN_TY)
bn_swish_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ rm, float* __restrict__ rv, int B, int F, float eps, float momentum, int training,
                    float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_istd) {
  __shared__ float red[BN_TY][33];
  const int c = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
  const bool ok = c < F;
  float v[BN_MAXR];
#pragma unroll
  for (int r = 0; r < BN_MAXR; ++r) {
    const int b = ty + r * BN_TY;
    v[r] = (ok && b < B) ? x[(size_t)b * F + c] : 0.f;
  }
  float mean, var;
  if (training) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < BN_MAXR; ++r) s += v[r];
    mean = col_sum(s, red) / (float)B;
    float q = 0.f;
#pragma unroll
    for (int r = 0; r < BN_MAXR; ++r) {
      const float d = (ty + r * BN_TY < B) ? v[r] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    var = col_sum(q, red) / (float)B;
    if (rm && ok && ty == 0) {
      rm[c] = (1.f - momentum) * rm[c] + momentum * mean;
      rv[c] = (1.f - momentum) * rv[c] + momentum * var * ((float)B / (float)max(B - 1, 1));
    }
  } else {
    mean = ok ? rm[c] : 0.f;
    var = ok ? rv[c] : 1.f;
  }
  if (!ok) return;
  const float istd = 1.f / sqrtf(var + eps);
  const float g = gamma[c], bt = beta[c];
#pragma unroll
  for (int r = 0; r < BN_MAXR; ++r) {
    const int b = ty + r * BN_TY;
    if (b < B) {
      const float z = fmaf((v[r] - mean) * istd, g, bt);
      y[(size_t)b * F + c] = z * sigmoidf_(z);
    }
  }
  if (ty == 0) {
    save_mean[c] = mean;
    save_istd[c] = istd;
  }
}

template <int BN_MAXR>
__global__ void __launch_bounds__(32 * BN_TY)
bn_swish_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ save_mean, const float* __restrict__ save_istd,
                    int B, int F, int training, float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[BN_TY][33];
  const int c = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
  const bool ok = c < F;
  const float mean = ok ? save_mean[c] : 0.f, istd = ok ? save_istd[c] : 0.f, g = ok ? gamma[c] : 0.f, bt = ok ? beta[c] : 0.f;
  float xh[BN_MAXR], dz[BN_MAXR];
  float dg = 0.f, db = 0.f;
#pragma unroll
  for (int r = 0; r < BN_MAXR; ++r) {
    const int b = ty + r * BN_TY;
    const bool in = ok && b < B;
    xh[r] = in ? (x[(size_t)b * F + c] - mean) * istd : 0.f;
    const float z = fmaf(xh[r], g, bt);
    const float s = sigmoidf_(z);
    dz[r] = in ? dy[(size_t)b * F + c] * (s + z * s * (1.f - s)) : 0.f;
    dg = fmaf(dz[r], xh[r], dg);
    db += dz[r];
  }
  dg = col_sum(dg, red);
  db = col_sum(db, red);
  if (!ok) return;
  if (ty == 0) {
    dgamma[c] = dg;
    dbeta[c] = db;
  }
  // d xhat = dz * gamma;  training: dx = istd (d xhat - mean_b(d xhat) - xhat mean_b(d xhat xhat))
  const float m1 = training ? g * db / (float)B : 0.f;
  const float m2 = training ? g * dg / (float)B : 0.f;
#pragma unroll
  for (int r = 0; r < BN_MAXR; ++r) {
    const int b = ty + r * BN_TY;
    if (b < B) dx[(size_t)b * F + c] = istd * (dz[r] * g - m1 - xh[r] * m2);
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
  DPF_REQUIRE(B <= BN_TY * 32, DPF_ERR_UNSUPPORTED, "dpf_bn_swish_forward: B <= %d (got %d)", BN_TY * 32, B);
  const dim3 grid((F + 31) / 32), block(32, BN_TY);
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= BN_TY * 4) bn_swish_fwd_kernel<4><<<grid, block, 0, st>>>(x, gamma, beta, rm, rv, B, F, eps, momentum, training, y, save_mean, save_istd);
  else if (B <= BN_TY * 8) bn_swish_fwd_kernel<8><<<grid, block, 0, st>>>(x, gamma, beta, rm, rv, B, F, eps, momentum, training, y, save_mean, save_istd);
  else bn_swish_fwd_kernel<32><<<grid, block, 0, st>>>(x, gamma, beta, rm, rv, B, F, eps, momentum, training, y, save_mean, save_istd);
  return dpf_check_launch("bn_swish_fwd_kernel");
}

DPF_API int dpf_bn_swish_backward(const float* dy, const float* x, const float* gamma, const float* beta, const float* save_mean,
                                  const float* save_istd, int B, int F, int training, float* dx, float* dgamma, float* dbeta,
                                  void* stream) {
  DPF_REQUIRE(dy && x && gamma && beta && save_mean && save_istd && dx && dgamma && dbeta, DPF_ERR_NULL_PTR,
              "dpf_bn_swish_backward: null pointer");
  DPF_REQUIRE(B > 0 && F > 0, DPF_ERR_BAD_ARG, "dpf_bn_swish_backward: bad sizes");
  DPF_REQUIRE(B <= BN_TY * 32, DPF_ERR_UNSUPPORTED, "dpf_bn_swish_backward: B <= %d (got %d)", BN_TY * 32, B);
  const dim3 grid((F + 31) / 32), block(32, BN_TY);
  cudaStream_t st = (cudaStream_t)stream;
  if (B <= BN_TY * 4) bn_swish_bwd_kernel<4><<<grid, block, 0, st>>>(dy, x, gamma, beta, save_mean, save_istd, B, F, training, dx, dgamma, dbeta);
  else if (B <= BN_TY * 8) bn_swish_bwd_kernel<8><<<grid, block, 0, st>>>(dy, x, gamma, beta, save_mean, save_istd, B, F, training, dx, dgamma, dbeta);
  else bn_swish_bwd_kernel<32><<<grid, block, 0, st>>>(dy, x, gamma, beta, save_mean, save_istd, B, F, training, dx, dgamma, dbeta);
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
