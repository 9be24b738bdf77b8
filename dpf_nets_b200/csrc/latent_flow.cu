// One coupling layer of the shape-latent flows (RealNVPFlow, reference lib/networks/flows.py:163-213) as ONE kernel forward
// and ONE kernel backward.  Per layer and branch (mu, logvar):
//     kept = g[:, keep]  ->  hpre = kept Wa^T  ->  BatchNorm1d (batch statistics)  ->  Swish  ->  raw = s Wb^T + bb
//     logvar = log(eps + exp(raw_lv)),  g_out[warp] = exp(+-logvar / 2) g[warp] (+-) mu,  g_out[keep] = g[keep]
// The library path is 4 GEMMs of 32 x 64 x 128 (SIMT sgemm, ~5 us of pure latency each), 2 BatchNorm + Swish kernels, the
// transform and a handful of index ops per layer forward, and three times that backward: ~100 launches of a few us for a
// MFLOP of work, 14 layers per training step.  Here the batch (B <= 64 shapes) lives in one CTA: LANE = BATCH ROW for every
// product with a weight matrix (the weight elements are warp-uniform 16-byte loads straight from global memory / L2: each
// element is read by exactly one warp, once), THREAD = COLUMN for the reductions over the batch (BatchNorm statistics, the
// weight gradients).  fp32 CUDA-core arithmetic throughout: same accuracy class as the library path.
#include "common.cuh"

namespace {

constexpr int LF_T = 512;            // 16 warps
constexpr int LF_CH = 8;             // columns per register-blocked chunk

struct LfBranch {
  const float* Wa;                   // (H, Kk)
  const float* gamma;                // (H)
  const float* beta;                 // (H)
  float* rm;                         // (H) running mean, nullable
  float* rv;                         // (H) running var, nullable
  const float* Wb;                   // (Wn, H)
  const float* bb;                   // (Wn)
};
struct LfGrads {
  float* dWa; float* dgamma; float* dbeta; float* dWb; float* dbb;
};

struct LfFwdArgs {
  const float* g; const int* pos; const int* keep_idx;
  LfBranch br[2];
  int B, D, H, Kk, Wn;
  float bn_eps, momentum, eps;
  int training, inverse;
  float* g_out; float* mu; float* lv;
  float* hpre;                       // [2][B][H]   saved for the backward
  float* stat;                       // [2][2][H]   {mean, istd}
  float* raw;                        // [2][B][Wn]
};

struct LfBwdArgs {
  const float* dgo; const float* dmu_f; const float* dlv_f;      // cotangents of g_out, mu, logvar (B,D); nullable
  const float* g; const int* pos; const int* keep_idx;
  LfBranch br[2];
  int B, D, H, Kk, Wn;
  float eps;
  int training, inverse;
  const float* hpre; const float* stat; const float* raw;
  float* dg;
  LfGrads gr[2];
};

__device__ __forceinline__ float sigmoid_(float z) { return 1.f / (1.f + expf(-z)); }

// C[b][c] = sum_k A[b][k] W[c][k] for the rows b = lane + 32 r of a shared-memory matrix A (leading dimension lda, 16-byte
// aligned rows) and 8 weight rows c (global memory, K contiguous floats each, 16-byte aligned): warp-uniform weight loads
template <int RB>
__device__ __forceinline__ void rows_times_wrows(const float* __restrict__ A, int lda, int K, const float* (&w)[LF_CH], int lane,
                                                 float (&acc)[RB][LF_CH]) {
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int j = 0; j < LF_CH; ++j) acc[r][j] = 0.f;
  for (int k = 0; k < K; k += 4) {
    float4 av[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) av[r] = *reinterpret_cast<const float4*>(A + (size_t)(lane + 32 * r) * lda + k);
#pragma unroll
    for (int j = 0; j < LF_CH; ++j) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w[j] + k));
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        acc[r][j] = fmaf(av[r].x, wv.x, acc[r][j]);
        acc[r][j] = fmaf(av[r].y, wv.y, acc[r][j]);
        acc[r][j] = fmaf(av[r].z, wv.z, acc[r][j]);
        acc[r][j] = fmaf(av[r].w, wv.w, acc[r][j]);
      }
    }
  }
}

template <int RB>
__global__ void __launch_bounds__(LF_T, 1)
latent_flow_fwd_kernel(const LfFwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int B = a.B, D = a.D, H = a.H, Kk = a.Kk, Wn = a.Wn;
  constexpr int BR = 32 * RB;
  const int ldk = Kk + 4, ldh = H + 4, ldw = Wn + 4;
  float* kept = sm;                          // [BR][ldk]
  float* h = kept + BR * ldk;                // [2][BR][ldh]
  float* raw = h + 2 * BR * ldh;             // [2][BR][ldw]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < BR * Kk; i += LF_T) {
    const int b = i / Kk, k = i - b * Kk;
    kept[b * ldk + k] = b < B ? a.g[(size_t)b * D + a.keep_idx[k]] : 0.f;
  }
  __syncthreads();
  // ---- hpre = kept Wa^T, both branches: 2 H columns in chunks of 8 over the warps ----
  for (int c0 = warp * LF_CH; c0 < 2 * H; c0 += (LF_T / 32) * LF_CH) {
    const int br = c0 / H, j0 = c0 - br * H;
    const float* Wa = br ? a.br[1].Wa : a.br[0].Wa;
    const float* w[LF_CH];
#pragma unroll
    for (int j = 0; j < LF_CH; ++j) w[j] = Wa + (size_t)(j0 + j) * Kk;
    float acc[RB][LF_CH];
    rows_times_wrows<RB>(kept, ldk, Kk, w, lane, acc);
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int b = lane + 32 * r;
      float* hr = h + ((size_t)br * BR + b) * ldh + j0;
      *reinterpret_cast<float4*>(hr) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      *reinterpret_cast<float4*>(hr + 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
      if (b < B) {
        float* gp = a.hpre + ((size_t)br * B + b) * H + j0;
        *reinterpret_cast<float4*>(gp) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        *reinterpret_cast<float4*>(gp + 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
      }
    }
  }
  __syncthreads();
  // ---- BatchNorm1d + Swish, one thread per (branch, feature) column ----
  for (int c = tid; c < 2 * H; c += LF_T) {
    const int br = c / H, j = c - br * H;
    const LfBranch& P = br ? a.br[1] : a.br[0];
    float* col = h + (size_t)br * BR * ldh + j;
    float mean, var;
    if (a.training) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += col[(size_t)b * ldh];
      mean = s / (float)B;
      float q = 0.f;
      for (int b = 0; b < B; ++b) {
        const float d = col[(size_t)b * ldh] - mean;
        q = fmaf(d, d, q);
      }
      var = q / (float)B;
      if (P.rm) {
        P.rm[j] = (1.f - a.momentum) * P.rm[j] + a.momentum * mean;
        P.rv[j] = (1.f - a.momentum) * P.rv[j] + a.momentum * var * ((float)B / (float)max(B - 1, 1));
      }
    } else {
      mean = P.rm[j];
      var = P.rv[j];
    }
    const float istd = 1.f / sqrtf(var + a.bn_eps);
    const float ga = P.gamma[j], be = P.beta[j];
    a.stat[(size_t)(br * 2 + 0) * H + j] = mean;
    a.stat[(size_t)(br * 2 + 1) * H + j] = istd;
    for (int b = 0; b < B; ++b) {
      const float z = fmaf((col[(size_t)b * ldh] - mean) * istd, ga, be);
      col[(size_t)b * ldh] = z * sigmoid_(z);
    }
    for (int b = B; b < BR; ++b) col[(size_t)b * ldh] = 0.f;
  }
  __syncthreads();
  // ---- raw = s Wb^T + bb, both branches: 2 Wn columns ----
  for (int c0 = warp * LF_CH; c0 < 2 * Wn; c0 += (LF_T / 32) * LF_CH) {
    const int br = c0 / Wn, j0 = c0 - br * Wn;
    const LfBranch& P = br ? a.br[1] : a.br[0];
    const float* w[LF_CH];
#pragma unroll
    for (int j = 0; j < LF_CH; ++j) w[j] = P.Wb + (size_t)(j0 + j) * H;
    float acc[RB][LF_CH];
    rows_times_wrows<RB>(h + (size_t)br * BR * ldh, ldh, H, w, lane, acc);
    float bias[LF_CH];
#pragma unroll
    for (int j = 0; j < LF_CH; ++j) bias[j] = P.bb[j0 + j];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int b = lane + 32 * r;
      float* rr = raw + ((size_t)br * BR + b) * ldw + j0;
#pragma unroll
      for (int j = 0; j < LF_CH; ++j) rr[j] = acc[r][j] + bias[j];
      if (b < B) {
        float* gp = a.raw + ((size_t)br * B + b) * Wn + j0;
#pragma unroll
        for (int j = 0; j < LF_CH; ++j) gp[j] = acc[r][j] + bias[j];
      }
    }
  }
  __syncthreads();
  // ---- the transform (same arithmetic as latent_affine_fwd_kernel) ----
  for (int e = tid; e < B * D; e += LF_T) {
    const int b = e / D, j = e - b * D;
    const int p = a.pos[j];
    float m = 0.f, l = 0.f;
    if (p >= 0) {
      m = raw[((size_t)0 * BR + b) * ldw + p];
      l = logf(a.eps + expf(raw[((size_t)1 * BR + b) * ldw + p]));
    }
    const float gv = a.g[e];
    a.mu[e] = m;
    a.lv[e] = l;
    a.g_out[e] = a.inverse ? expf(-0.5f * l) * (gv - m) : fmaf(expf(0.5f * l), gv, m);
  }
}

template <int RB>
__global__ void __launch_bounds__(LF_T, 1)
latent_flow_bwd_kernel(const LfBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int B = a.B, D = a.D, H = a.H, Kk = a.Kk, Wn = a.Wn;
  constexpr int BR = 32 * RB;
  const int ldk = Kk + 4, ldh = H + 4, ldw = Wn + 4;
  float* kept = sm;                          // [BR][ldk]
  float* S = kept + BR * ldk;                // [2][BR][ldh]   s = swish(y)
  float* DY = S + 2 * BR * ldh;              // [2][BR][ldh]   swish'(y) -> d y -> d hpre
  float* DR = DY + 2 * BR * ldh;             // [2][BR][ldw]   d raw
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- 0. kept rows, recomputed activations ----
  for (int i = tid; i < BR * Kk; i += LF_T) {
    const int b = i / Kk, k = i - b * Kk;
    kept[b * ldk + k] = b < B ? a.g[(size_t)b * D + a.keep_idx[k]] : 0.f;
  }
  for (int i = tid; i < 2 * BR * H; i += LF_T) {
    const int j = i % H, b = (i / H) % BR, br = i / (H * BR);
    float s = 0.f, ds = 0.f;
    if (b < B) {
      const LfBranch& P = br ? a.br[1] : a.br[0];
      const float mean = a.stat[(size_t)(br * 2 + 0) * H + j], istd = a.stat[(size_t)(br * 2 + 1) * H + j];
      const float z = fmaf((a.hpre[((size_t)br * B + b) * H + j] - mean) * istd, P.gamma[j], P.beta[j]);
      const float sg = sigmoid_(z);
      s = z * sg;
      ds = sg + z * sg * (1.f - sg);
    }
    S[((size_t)br * BR + b) * ldh + j] = s;
    DY[((size_t)br * BR + b) * ldh + j] = ds;
  }
  for (int i = tid; i < 2 * BR * ldw; i += LF_T) DR[i] = 0.f;
  __syncthreads();
  // ---- 1. the transform backward (same arithmetic as latent_affine_bwd_kernel): d raw, dg of the warped positions ----
  for (int e = tid; e < B * D; e += LF_T) {
    const int b = e / D, j = e - b * D;
    const int p = a.pos[j];
    const float d = a.dgo ? a.dgo[e] : 0.f;
    if (p < 0) continue;                 // kept positions: written in step 6
    const float rw = a.raw[((size_t)1 * B + b) * Wn + p];
    const float ex = expf(rw);
    const float l = logf(a.eps + ex);
    const float m = a.raw[((size_t)0 * B + b) * Wn + p];
    const float gv = a.g[e];
    float dm = a.dmu_f ? a.dmu_f[e] : 0.f, dl = a.dlv_f ? a.dlv_f[e] : 0.f, dgv;
    if (a.inverse) {
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
    a.dg[e] = dgv;
    DR[((size_t)0 * BR + b) * ldw + p] = dm;
    DR[((size_t)1 * BR + b) * ldw + p] = dl * ex / (a.eps + ex);
  }
  __syncthreads();
  // ---- 2. dbb = sum_b d raw;  dWb[w][j] = sum_b d raw[b][w] s[b][j]   (thread = (branch, j) column x 32 values of w) ----
  for (int c = tid; c < 2 * Wn; c += LF_T) {
    const int br = c / Wn, w = c - br * Wn;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += DR[((size_t)br * BR + b) * ldw + w];
    (br ? a.gr[1].dbb : a.gr[0].dbb)[w] = s;
  }
  for (int item = tid; item < 2 * H * ((Wn + 31) / 32); item += LF_T) {
    const int c = item % (2 * H), wblk = item / (2 * H);
    const int br = c / H, j = c - br * H, w0 = wblk * 32;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    const float* sc = S + (size_t)br * BR * ldh + j;
    const float* dr = DR + (size_t)br * BR * ldw + w0;
    for (int b = 0; b < B; ++b) {
      const float sv = sc[(size_t)b * ldh];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 d4 = *reinterpret_cast<const float4*>(dr + (size_t)b * ldw + 4 * i);
        acc[4 * i + 0] = fmaf(d4.x, sv, acc[4 * i + 0]);
        acc[4 * i + 1] = fmaf(d4.y, sv, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(d4.z, sv, acc[4 * i + 2]);
        acc[4 * i + 3] = fmaf(d4.w, sv, acc[4 * i + 3]);
      }
    }
    float* dWb = br ? a.gr[1].dWb : a.gr[0].dWb;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (w0 + i < Wn) dWb[(size_t)(w0 + i) * H + j] = acc[i];
  }
  // ---- 3. d y = swish'(y) * (d raw Wb): lanes = rows, 8 columns j per chunk, K = Wn ----
  for (int c0 = warp * LF_CH; c0 < 2 * H; c0 += (LF_T / 32) * LF_CH) {
    const int br = c0 / H, j0 = c0 - br * H;
    const float* Wb = br ? a.br[1].Wb : a.br[0].Wb;
    float acc[RB][LF_CH];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int j = 0; j < LF_CH; ++j) acc[r][j] = 0.f;
    const float* dr = DR + (size_t)br * BR * ldw;
    for (int w = 0; w < Wn; w += 4) {
      float4 dv[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) dv[r] = *reinterpret_cast<const float4*>(dr + (size_t)(lane + 32 * r) * ldw + w);
#pragma unroll
      for (int ww = 0; ww < 4; ++ww) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(Wb + (size_t)(w + ww) * H + j0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(Wb + (size_t)(w + ww) * H + j0 + 4));
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const float d = ww == 0 ? dv[r].x : ww == 1 ? dv[r].y : ww == 2 ? dv[r].z : dv[r].w;
          acc[r][0] = fmaf(d, w0.x, acc[r][0]); acc[r][1] = fmaf(d, w0.y, acc[r][1]);
          acc[r][2] = fmaf(d, w0.z, acc[r][2]); acc[r][3] = fmaf(d, w0.w, acc[r][3]);
          acc[r][4] = fmaf(d, w1.x, acc[r][4]); acc[r][5] = fmaf(d, w1.y, acc[r][5]);
          acc[r][6] = fmaf(d, w1.z, acc[r][6]); acc[r][7] = fmaf(d, w1.w, acc[r][7]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float* dy = DY + ((size_t)br * BR + lane + 32 * r) * ldh + j0;
#pragma unroll
      for (int j = 0; j < LF_CH; ++j) dy[j] *= acc[r][j];
    }
  }
  __syncthreads();
  // ---- 4. BatchNorm backward per (branch, feature) column: dgamma, dbeta, d hpre (in place) ----
  for (int c = tid; c < 2 * H; c += LF_T) {
    const int br = c / H, j = c - br * H;
    const LfBranch& P = br ? a.br[1] : a.br[0];
    const float mean = a.stat[(size_t)(br * 2 + 0) * H + j], istd = a.stat[(size_t)(br * 2 + 1) * H + j];
    const float ga = P.gamma[j];
    float* col = DY + (size_t)br * BR * ldh + j;
    const float* hp = a.hpre + (size_t)br * B * H + j;
    float dgam = 0.f, dbet = 0.f;
    for (int b = 0; b < B; ++b) {
      const float xh = (hp[(size_t)b * H] - mean) * istd;
      const float dz = col[(size_t)b * ldh];
      dgam = fmaf(dz, xh, dgam);
      dbet += dz;
    }
    (br ? a.gr[1].dgamma : a.gr[0].dgamma)[j] = dgam;
    (br ? a.gr[1].dbeta : a.gr[0].dbeta)[j] = dbet;
    const float m1 = a.training ? ga * dbet / (float)B : 0.f;
    const float m2 = a.training ? ga * dgam / (float)B : 0.f;
    for (int b = 0; b < B; ++b) {
      const float xh = (hp[(size_t)b * H] - mean) * istd;
      col[(size_t)b * ldh] = istd * (col[(size_t)b * ldh] * ga - m1 - xh * m2);
    }
  }
  __syncthreads();
  // ---- 5. dWa[j][k] = sum_b d hpre[b][j] kept[b][k]   (thread = (branch, j) column x 32 values of k) ----
  for (int item = tid; item < 2 * H * ((Kk + 31) / 32); item += LF_T) {
    const int c = item % (2 * H), kblk = item / (2 * H);
    const int br = c / H, j = c - br * H, k0 = kblk * 32;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    const float* dc = DY + (size_t)br * BR * ldh + j;
    for (int b = 0; b < B; ++b) {
      const float dv = dc[(size_t)b * ldh];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 k4 = *reinterpret_cast<const float4*>(kept + (size_t)b * ldk + k0 + 4 * i);
        acc[4 * i + 0] = fmaf(k4.x, dv, acc[4 * i + 0]);
        acc[4 * i + 1] = fmaf(k4.y, dv, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(k4.z, dv, acc[4 * i + 2]);
        acc[4 * i + 3] = fmaf(k4.w, dv, acc[4 * i + 3]);
      }
    }
    float* dWa = (br ? a.gr[1].dWa : a.gr[0].dWa) + (size_t)j * Kk + k0;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (k0 + i < Kk) dWa[i] = acc[i];
  }
  // ---- 6. d kept = d hpre Wa (both branches): lanes = rows, 4 columns k per warp item; dg[keep] = dgo[keep] + d kept ----
  for (int k0 = warp * 4; k0 < Kk; k0 += (LF_T / 32) * 4) {
    float acc[RB][4];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
    for (int br = 0; br < 2; ++br) {
      const float* Wa = br ? a.br[1].Wa : a.br[0].Wa;
      const float* dh = DY + (size_t)br * BR * ldh;
      for (int j = 0; j < H; j += 4) {
        float4 dv[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) dv[r] = *reinterpret_cast<const float4*>(dh + (size_t)(lane + 32 * r) * ldh + j);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(Wa + (size_t)(j + jj) * Kk + k0));
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const float d = jj == 0 ? dv[r].x : jj == 1 ? dv[r].y : jj == 2 ? dv[r].z : dv[r].w;
            acc[r][0] = fmaf(d, wv.x, acc[r][0]);
            acc[r][1] = fmaf(d, wv.y, acc[r][1]);
            acc[r][2] = fmaf(d, wv.z, acc[r][2]);
            acc[r][3] = fmaf(d, wv.w, acc[r][3]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int b = lane + 32 * r;
      if (b < B) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const size_t e = (size_t)b * D + a.keep_idx[k0 + i];
          a.dg[e] = (a.dgo ? a.dgo[e] : 0.f) + acc[r][i];
        }
      }
    }
  }
}

size_t lf_fwd_smem(int RB, int H, int Kk, int Wn) { return sizeof(float) * (size_t)(32 * RB) * ((Kk + 4) + 2 * (H + 4) + 2 * (Wn + 4)); }
size_t lf_bwd_smem(int RB, int H, int Kk, int Wn) { return sizeof(float) * (size_t)(32 * RB) * ((Kk + 4) + 4 * (H + 4) + 2 * (Wn + 4)); }

int lf_check_dims(const char* what, int B, int D, int H, int Kk, int Wn) {
  DPF_REQUIRE(B > 0 && D > 0 && H > 0 && Kk > 0 && Wn > 0 && Kk <= D && Wn <= D, DPF_ERR_BAD_ARG, "%s: bad sizes", what);
  DPF_REQUIRE(B <= 64 && H % 8 == 0 && Kk % 32 == 0 && Wn % 32 == 0, DPF_ERR_UNSUPPORTED,
              "%s: needs B <= 64, H %% 8 == 0, kept / warped widths multiples of 32 (got B=%d H=%d kept=%d warped=%d)", what, B, H, Kk, Wn);
  return DPF_OK;
}

bool lf_aligned(const LfBranch& b) { return (((uintptr_t)b.Wa | (uintptr_t)b.Wb) & 15) == 0; }

}  // namespace

// One RealNVPFlow layer forward.  g (B,D); pos (D,) int32: index in the warp list or -1; keep_idx (Kk,) int32: the kept
// positions in order; per branch b in {0: mu, 1: logvar}: Wa[b] (H,Kk), gamma[b], beta[b] (H), rm[b], rv[b] (H) running
// statistics (updated in place in training mode when not null; read in eval mode), Wb[b] (Wn,H), bb[b] (Wn).
// Outputs g_out, mu, lv (B,D) and, for the backward, hpre (2,B,H), stat (2,2,H) {mean, istd}, raw (2,B,Wn).
DPF_API int dpf_latent_flow_forward(const float* g, const int* pos, const int* keep_idx, const float* const* Wa, const float* const* gamma,
                                    const float* const* beta, float* const* rm, float* const* rv, const float* const* Wb,
                                    const float* const* bb, int B, int D, int H, int Kk, int Wn, float bn_eps, float momentum, int training,
                                    float eps, int inverse, float* g_out, float* mu, float* lv, float* hpre, float* stat, float* raw,
                                    void* stream) {
  DPF_REQUIRE(g && pos && keep_idx && Wa && gamma && beta && rm && rv && Wb && bb && g_out && mu && lv && hpre && stat && raw, DPF_ERR_NULL_PTR,
              "dpf_latent_flow_forward: null pointer");
  int rc = lf_check_dims("dpf_latent_flow_forward", B, D, H, Kk, Wn);
  if (rc) return rc;
  LfFwdArgs a{};
  a.g = g; a.pos = pos; a.keep_idx = keep_idx;
  for (int b = 0; b < 2; ++b) {
    a.br[b] = LfBranch{Wa[b], gamma[b], beta[b], rm[b], rv[b], Wb[b], bb[b]};
    DPF_REQUIRE(Wa[b] && gamma[b] && beta[b] && Wb[b] && bb[b], DPF_ERR_NULL_PTR, "dpf_latent_flow_forward: null branch pointer");
    DPF_REQUIRE(training || (rm[b] && rv[b]), DPF_ERR_NULL_PTR, "dpf_latent_flow_forward: eval mode needs running statistics");
    DPF_REQUIRE(lf_aligned(a.br[b]), DPF_ERR_ALIGN, "dpf_latent_flow_forward: weight matrices must be 16-byte aligned");
  }
  DPF_REQUIRE((((uintptr_t)hpre | (uintptr_t)raw) & 15) == 0, DPF_ERR_ALIGN, "dpf_latent_flow_forward: hpre / raw alignment");
  a.B = B; a.D = D; a.H = H; a.Kk = Kk; a.Wn = Wn;
  a.bn_eps = bn_eps; a.momentum = momentum; a.eps = eps; a.training = training; a.inverse = inverse;
  a.g_out = g_out; a.mu = mu; a.lv = lv; a.hpre = hpre; a.stat = stat; a.raw = raw;
  const int RB = B <= 32 ? 1 : 2;
  const size_t smem = lf_fwd_smem(RB, H, Kk, Wn);
  DPF_REQUIRE(smem <= 227 * 1024, DPF_ERR_UNSUPPORTED, "dpf_latent_flow_forward: layer too wide for one CTA (%zu bytes of shared memory)", smem);
  cudaStream_t s = (cudaStream_t)stream;
  if (RB == 1) {
    cudaFuncSetAttribute(latent_flow_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    latent_flow_fwd_kernel<1><<<1, LF_T, smem, s>>>(a);
  } else {
    cudaFuncSetAttribute(latent_flow_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    latent_flow_fwd_kernel<2><<<1, LF_T, smem, s>>>(a);
  }
  return dpf_check_launch("latent_flow_fwd_kernel");
}

// One RealNVPFlow layer backward: cotangents dgo, dmu_f, dlv_f (B,D; each nullable = zero) of g_out, mu, logvar -> dg (B,D) and,
// per branch, dWa (H,Kk), dgamma, dbeta (H), dWb (Wn,H), dbb (Wn).  hpre / stat / raw: the forward's saved buffers.
DPF_API int dpf_latent_flow_backward(const float* dgo, const float* dmu_f, const float* dlv_f, const float* g, const int* pos,
                                     const int* keep_idx, const float* const* Wa, const float* const* gamma, const float* const* beta,
                                     const float* const* Wb, int B, int D, int H, int Kk, int Wn, int training, float eps, int inverse,
                                     const float* hpre, const float* stat, const float* raw, float* dg, float* const* dWa,
                                     float* const* dgamma, float* const* dbeta, float* const* dWb, float* const* dbb, void* stream) {
  DPF_REQUIRE(g && pos && keep_idx && Wa && gamma && beta && Wb && hpre && stat && raw && dg && dWa && dgamma && dbeta && dWb && dbb,
              DPF_ERR_NULL_PTR, "dpf_latent_flow_backward: null pointer");
  int rc = lf_check_dims("dpf_latent_flow_backward", B, D, H, Kk, Wn);
  if (rc) return rc;
  LfBwdArgs a{};
  a.dgo = dgo; a.dmu_f = dmu_f; a.dlv_f = dlv_f; a.g = g; a.pos = pos; a.keep_idx = keep_idx;
  for (int b = 0; b < 2; ++b) {
    DPF_REQUIRE(Wa[b] && gamma[b] && beta[b] && Wb[b] && dWa[b] && dgamma[b] && dbeta[b] && dWb[b] && dbb[b], DPF_ERR_NULL_PTR,
                "dpf_latent_flow_backward: null branch pointer");
    a.br[b] = LfBranch{Wa[b], gamma[b], beta[b], nullptr, nullptr, Wb[b], nullptr};
    a.gr[b] = LfGrads{dWa[b], dgamma[b], dbeta[b], dWb[b], dbb[b]};
    DPF_REQUIRE(lf_aligned(a.br[b]), DPF_ERR_ALIGN, "dpf_latent_flow_backward: weight matrices must be 16-byte aligned");
  }
  a.B = B; a.D = D; a.H = H; a.Kk = Kk; a.Wn = Wn; a.eps = eps; a.training = training; a.inverse = inverse;
  a.hpre = hpre; a.stat = stat; a.raw = raw; a.dg = dg;
  const int RB = B <= 32 ? 1 : 2;
  const size_t smem = lf_bwd_smem(RB, H, Kk, Wn);
  DPF_REQUIRE(smem <= 227 * 1024, DPF_ERR_UNSUPPORTED, "dpf_latent_flow_backward: layer too wide for one CTA (%zu bytes of shared memory)", smem);
  cudaStream_t s = (cudaStream_t)stream;
  if (RB == 1) {
    cudaFuncSetAttribute(latent_flow_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    latent_flow_bwd_kernel<1><<<1, LF_T, smem, s>>>(a);
  } else {
    cudaFuncSetAttribute(latent_flow_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    latent_flow_bwd_kernel<2><<<1, LF_T, smem, s>>>(a);
  }
  return dpf_check_launch("latent_flow_bwd_kernel");
}
