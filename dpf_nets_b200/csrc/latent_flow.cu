// One coupling layer of the shape-latent flows (RealNVPFlow, reference lib/networks/flows.py:163-213) as ONE kernel forward
// and ONE kernel backward.  Per layer and branch (mu, logvar):
//     kept = g[:, keep]  ->  hpre = kept Wa^T  ->  BatchNorm1d (batch statistics)  ->  Swish  ->  raw = s Wb^T + bb
//     logvar = log(eps + exp(raw_lv)),  g_out[warp] = exp(+-logvar / 2) g[warp] (+-) mu,  g_out[keep] = g[keep]
// The library path is 4 GEMMs of 32 x 64 x 128 (SIMT sgemm, ~5 us of pure latency each), 2 BatchNorm + Swish kernels, the
// transform and a handful of index ops per layer forward, and three times that backward: ~100 launches of a few us for a
// MFLOP of work, 14 layers per training step.
//
// Here a layer is one THREAD-BLOCK CLUSTER of 8 CTAs (a first version with one CTA per layer measured 27 / 70 us forward /
// backward: one SM's instruction issue is the bound).  The hidden width is split over the CTAs: every CTA owns H / 8 hidden
// columns per branch - its slices of Wa, of the BatchNorm statistics (the whole batch, B <= 64 shapes, lives in every CTA,
// so the statistics need no exchange) and of Wb's columns - and the two places where all hidden columns meet go through
// DISTRIBUTED SHARED MEMORY: forward, every CTA pushes its slice of s = swish(bn(hpre)) into all eight CTAs' copies before
// the second product (each CTA then forms Wn / 8 output columns per branch and transforms those positions); backward, the
// partial sums of d kept = d hpre Wa over the hidden slices are reduced by reading the eight CTAs' buffers.
// LANE = BATCH ROW for the products with weight matrices (weight elements are warp-uniform shared-memory reads) and for the
// BatchNorm reductions (warp shuffles); fp32 CUDA-core arithmetic throughout: same accuracy class as the library path.
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int LC_NC = 8;             // CTAs per cluster
constexpr int LC_T = 256;            // threads per CTA (8 warps)
constexpr int LC_CH = 4;             // columns per register-blocked chunk

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

// phase stamps of the most recent launches (thread 0 of cluster rank 0; eight stores per kernel): dpf_latent_flow_stamps()
__device__ unsigned long long g_lf_stamps[2][12];
__device__ __forceinline__ void lf_stamp(int which, int i, int rk, int tid) {
  if (rk == 0 && tid == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_lf_stamps[which][i] = t;
  }
}

// acc[r][j] = sum_k A[lane + 32 r][k] W[j][k]: A in shared memory (leading dimension lda), 4 weight rows in shared memory
// (warp-uniform 16-byte reads), K a multiple of 4
template <int RB>
__device__ __forceinline__ void rows_times_wrows(const float* __restrict__ A, int lda, int K, const float* w0, int ldw, int lane,
                                                 float (&acc)[RB][LC_CH]) {
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int j = 0; j < LC_CH; ++j) acc[r][j] = 0.f;
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float4 av[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) av[r] = *reinterpret_cast<const float4*>(A + (size_t)(lane + 32 * r) * lda + k);
#pragma unroll
    for (int j = 0; j < LC_CH; ++j) {
      const float4 wv = *reinterpret_cast<const float4*>(w0 + (size_t)j * ldw + k);
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

__device__ __forceinline__ void copy4(float* dst, const float* __restrict__ src, int n4, int tid) {
  for (int i = tid; i < n4; i += LC_T) reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
}

template <int RB>
__global__ void __cluster_dims__(LC_NC, 1, 1) __launch_bounds__(LC_T, 1)
latent_flow_fwd_kernel(const LfFwdArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rk = (int)cluster.block_rank();
  extern __shared__ __align__(16) float sm[];
  const int B = a.B, D = a.D, H = a.H, Kk = a.Kk, Wn = a.Wn;
  constexpr int BR = 32 * RB;
  const int JS = H / LC_NC, WS = Wn / LC_NC;        // hidden / output columns per branch owned by this CTA
  const int ldk = Kk + 4, ldh = H + 4;
  float* kept = sm;                          // [BR][ldk]
  float* sW1 = kept + BR * ldk;              // [2][JS][Kk]   rows rk*JS.. of Wa
  float* sW2 = sW1 + 2 * JS * Kk;            // [2][WS][H]    rows rk*WS.. of Wb
  float* hl = sW2 + 2 * WS * H;              // [2][BR][JS]   own hidden slice: hpre -> s
  float* s_all = hl + 2 * BR * JS;           // [2][BR][ldh]  all hidden columns (pushed by the eight CTAs)
  float* rawl = s_all + 2 * BR * ldh;        // [2][BR][WS]
  float* gS = rawl + 2 * BR * WS;            // [B][D]        the layer input (coalesced copy: every later access is shared memory)
  int* kidx = reinterpret_cast<int*>(gS + (size_t)B * D);      // [Kk] kept positions
  int* wd = kidx + Kk;                       // [Wn] warped positions (inverse of pos)
  float* colp = reinterpret_cast<float*>(wd + Wn);             // [2 JS][4] {gamma, beta, running mean, running var} of the own columns
  float* biasl = colp + 8 * JS;              // [2 WS] bias of the own output columns
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  lf_stamp(0, 0, rk, tid);
  pdl_launch_dependents();      // the next layer's kernel may start its own prologue
  cluster.barrier_arrive();     // "this CTA has started": waited for only right before the first remote write
  // everything that does not depend on the previous layer's kernel first (programmatic dependent launch: this part
  // overlaps that kernel's tail); all global reads of the prologue are independent of each other (one round trip)
  for (int k = tid; k < Kk; k += LC_T) kidx[k] = a.keep_idx[k];
  for (int d = tid; d < D; d += LC_T) {
    const int p = a.pos[d];
    if (p >= 0) wd[p] = d;
  }
  for (int c = tid; c < 2 * JS; c += LC_T) {
    const int br = c / JS, j = rk * JS + (c - br * JS);
    const LfBranch& P = br ? a.br[1] : a.br[0];
    const bool stats = !a.training || P.rm;
    colp[4 * c + 0] = P.gamma[j];
    colp[4 * c + 1] = P.beta[j];
    colp[4 * c + 2] = stats ? P.rm[j] : 0.f;
    colp[4 * c + 3] = stats ? P.rv[j] : 1.f;
  }
  for (int c = tid; c < 2 * WS; c += LC_T) {
    const int br = c / WS;
    biasl[c] = (br ? a.br[1].bb : a.br[0].bb)[rk * WS + (c - br * WS)];
  }
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    copy4(sW1 + br * JS * Kk, (br ? a.br[1].Wa : a.br[0].Wa) + (size_t)rk * JS * Kk, JS * Kk / 4, tid);
    copy4(sW2 + br * WS * H, (br ? a.br[1].Wb : a.br[0].Wb) + (size_t)rk * WS * H, WS * H / 4, tid);
  }
  lf_stamp(0, 1, rk, tid);
  pdl_wait();                   // the layer input (and every buffer this kernel writes) belongs to us from here on
  lf_stamp(0, 2, rk, tid);
  copy4(gS, a.g, B * D / 4, tid);
  __syncthreads();
  for (int i = tid; i < BR * Kk; i += LC_T) {
    const int b = i / Kk, k = i - b * Kk;
    kept[b * ldk + k] = b < B ? gS[(size_t)b * D + kidx[k]] : 0.f;
  }
  __syncthreads();
  lf_stamp(0, 3, rk, tid);
  // ---- hpre = kept Wa^T on the own 2 JS columns ----
  for (int c0 = warp * LC_CH; c0 < 2 * JS; c0 += (LC_T / 32) * LC_CH) {
    const int br = c0 / JS, jl = c0 - br * JS;
    float acc[RB][LC_CH];
    rows_times_wrows<RB>(kept, ldk, Kk, sW1 + (size_t)(br * JS + jl) * Kk, Kk, lane, acc);
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int b = lane + 32 * r;
      const float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      *reinterpret_cast<float4*>(hl + ((size_t)br * BR + b) * JS + jl) = v;
      if (b < B) *reinterpret_cast<float4*>(a.hpre + ((size_t)br * B + b) * H + rk * JS + jl) = v;
    }
  }
  __syncthreads();
  lf_stamp(0, 4, rk, tid);
  // ---- BatchNorm1d + Swish: four columns per warp item (independent shuffle chains in flight together), lane = batch row ----
  for (int c0 = warp * 4; c0 < 2 * JS; c0 += (LC_T / 32) * 4) {
    const int br = c0 / JS, jl = c0 - br * JS, j = rk * JS + jl;
    const LfBranch& P = br ? a.br[1] : a.br[0];
    float4 v[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r)
      v[r] = (lane + 32 * r < B) ? *reinterpret_cast<const float4*>(hl + ((size_t)br * BR + lane + 32 * r) * JS + jl) : make_float4(0.f, 0.f, 0.f, 0.f);
    float mean[4], var[4];
    if (a.training) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < RB; ++r) { s[0] += v[r].x; s[1] += v[r].y; s[2] += v[r].z; s[3] += v[r].w; }
#pragma unroll
      for (int q = 0; q < 4; ++q) mean[q] = warp_sum(s[q]) / (float)B;
      float q2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (lane + 32 * r < B) {
          const float d0 = v[r].x - mean[0], d1 = v[r].y - mean[1], d2 = v[r].z - mean[2], d3 = v[r].w - mean[3];
          q2[0] = fmaf(d0, d0, q2[0]); q2[1] = fmaf(d1, d1, q2[1]); q2[2] = fmaf(d2, d2, q2[2]); q2[3] = fmaf(d3, d3, q2[3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) var[q] = warp_sum(q2[q]) / (float)B;
      if (P.rm && lane < 4) {
        const float mq = lane == 0 ? mean[0] : lane == 1 ? mean[1] : lane == 2 ? mean[2] : mean[3];
        const float vq = lane == 0 ? var[0] : lane == 1 ? var[1] : lane == 2 ? var[2] : var[3];
        P.rm[j + lane] = (1.f - a.momentum) * colp[4 * (c0 + lane) + 2] + a.momentum * mq;
        P.rv[j + lane] = (1.f - a.momentum) * colp[4 * (c0 + lane) + 3] + a.momentum * vq * ((float)B / (float)max(B - 1, 1));
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) { mean[q] = colp[4 * (c0 + q) + 2]; var[q] = colp[4 * (c0 + q) + 3]; }
    }
    float istd[4], ga[4], be[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      istd[q] = 1.f / sqrtf(var[q] + a.bn_eps);
      ga[q] = colp[4 * (c0 + q) + 0];
      be[q] = colp[4 * (c0 + q) + 1];
    }
    if (lane < 4) {
      const float mq = lane == 0 ? mean[0] : lane == 1 ? mean[1] : lane == 2 ? mean[2] : mean[3];
      const float iq = lane == 0 ? istd[0] : lane == 1 ? istd[1] : lane == 2 ? istd[2] : istd[3];
      a.stat[(size_t)(br * 2 + 0) * H + j + lane] = mq;
      a.stat[(size_t)(br * 2 + 1) * H + j + lane] = iq;
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const bool in = lane + 32 * r < B;
      const float z0 = fmaf((v[r].x - mean[0]) * istd[0], ga[0], be[0]), z1 = fmaf((v[r].y - mean[1]) * istd[1], ga[1], be[1]);
      const float z2 = fmaf((v[r].z - mean[2]) * istd[2], ga[2], be[2]), z3 = fmaf((v[r].w - mean[3]) * istd[3], ga[3], be[3]);
      *reinterpret_cast<float4*>(hl + ((size_t)br * BR + lane + 32 * r) * JS + jl) =
          in ? make_float4(z0 * sigmoid_(z0), z1 * sigmoid_(z1), z2 * sigmoid_(z2), z3 * sigmoid_(z3)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  lf_stamp(0, 5, rk, tid);
  // ---- push the own slice of s into every CTA's copy (distributed shared memory, 16-byte stores) ----
  cluster.barrier_wait();       // every CTA of the cluster has started: its shared memory may be written remotely
  {
    const int q4 = JS / 4, n4 = 2 * BR * q4;
    for (int i = tid; i < n4; i += LC_T) {
      const int q = i % q4, row = i / q4;                       // row = br * BR + b
      const float4 v = *reinterpret_cast<const float4*>(hl + (size_t)row * JS + 4 * q);
#pragma unroll
      for (int t = 0; t < LC_NC; ++t) {
        float* dst = cluster.map_shared_rank(s_all, (unsigned)((rk + t) % LC_NC));
        *reinterpret_cast<float4*>(dst + (size_t)row * ldh + rk * JS + 4 * q) = v;
      }
    }
  }
  cluster.sync();
  lf_stamp(0, 6, rk, tid);
  // ---- raw = s Wb^T + bb on the own 2 WS output columns ----
  for (int c0 = warp * LC_CH; c0 < 2 * WS; c0 += (LC_T / 32) * LC_CH) {
    const int br = c0 / WS, wl = c0 - br * WS;
    float acc[RB][LC_CH];
    rows_times_wrows<RB>(s_all + (size_t)br * BR * ldh, ldh, H, sW2 + (size_t)(br * WS + wl) * H, H, lane, acc);
    const float4 bias = *reinterpret_cast<const float4*>(biasl + br * WS + wl);
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int b = lane + 32 * r;
      const float4 v = make_float4(acc[r][0] + bias.x, acc[r][1] + bias.y, acc[r][2] + bias.z, acc[r][3] + bias.w);
      *reinterpret_cast<float4*>(rawl + ((size_t)br * BR + b) * WS + wl) = v;
      if (b < B) *reinterpret_cast<float4*>(a.raw + ((size_t)br * B + b) * Wn + rk * WS + wl) = v;
    }
  }
  __syncthreads();
  lf_stamp(0, 7, rk, tid);
  // ---- the transform (same arithmetic as latent_affine_fwd_kernel) on the warped positions whose raw columns this CTA
  //      formed; the kept positions are dealt over the CTAs by kept-column range ----
  for (int i = tid; i < B * WS; i += LC_T) {
    const int wl = i % WS, b = i / WS;
    const size_t e = (size_t)b * D + wd[rk * WS + wl];
    const float m = rawl[((size_t)0 * BR + b) * WS + wl];
    const float l = logf(a.eps + expf(rawl[((size_t)1 * BR + b) * WS + wl]));
    const float gv = gS[e];
    a.mu[e] = m;
    a.lv[e] = l;
    a.g_out[e] = a.inverse ? expf(-0.5f * l) * (gv - m) : fmaf(expf(0.5f * l), gv, m);
  }
  const int KSf = Kk / LC_NC;
  for (int i = tid; i < B * KSf; i += LC_T) {
    const int kl = i % KSf, b = i / KSf;
    const size_t e = (size_t)b * D + kidx[rk * KSf + kl];
    a.mu[e] = 0.f;
    a.lv[e] = 0.f;
    a.g_out[e] = gS[e];
  }
  lf_stamp(0, 8, rk, tid);
}

// GS: the layer input g is staged in shared memory (when it fits; wide latents read it from global memory instead)
template <int RB, bool GS>
__global__ void __cluster_dims__(LC_NC, 1, 1) __launch_bounds__(LC_T, 1)
latent_flow_bwd_kernel(const LfBwdArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rk = (int)cluster.block_rank();
  extern __shared__ __align__(16) float sm[];
  const int B = a.B, D = a.D, H = a.H, Kk = a.Kk, Wn = a.Wn;
  constexpr int BR = 32 * RB;
  const int JS = H / LC_NC, KS = Kk / LC_NC;
  const int ldk = Kk + 4, ldw = Wn + 4;
  float* kept = sm;                          // [BR][ldk]
  float* DR = kept + BR * ldk;               // [2][BR][ldw]  d raw, all output columns (formed by every CTA)
  float* Sl = DR + 2 * BR * ldw;             // [2][BR][JS]   s = swish(y) of the own hidden columns
  float* XH = Sl + 2 * BR * JS;              // [2][BR][JS]   xhat
  float* DY = XH + 2 * BR * JS;              // [2][BR][JS]   swish'(y) -> d y -> d hpre
  float* sWbT = DY + 2 * BR * JS;            // [2][Wn][JS]   columns rk*JS.. of Wb
  float* sWa = sWbT + 2 * Wn * JS;           // [2][JS][Kk]   rows rk*JS.. of Wa
  float* dkp = sWa + 2 * JS * Kk;            // [BR][ldk]     this CTA's partial d kept
  int* kidx = reinterpret_cast<int*>(dkp + BR * ldk);          // [Kk] kept positions
  int* wd = kidx + Kk;                       // [Wn] warped positions (inverse of pos)
  float* colp = reinterpret_cast<float*>(wd + Wn);             // [2 JS][4] {gamma, beta, mean, istd} of the own columns
  float* gS = colp + 8 * JS;                 // GS: [B][D] the layer input
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- 0. operands: all of them were produced by the forward pass, none by the previous backward kernel, so this whole step
  //         overlaps that kernel's tail (programmatic dependent launch); independent global reads first ----
  lf_stamp(1, 0, rk, tid);
  pdl_launch_dependents();
  cluster.barrier_arrive();     // "this CTA has started": waited for right before the first remote write
  if (GS) copy4(gS, a.g, B * D / 4, tid);
  for (int c = tid; c < 2 * JS; c += LC_T) {
    const int br = c / JS, j = rk * JS + (c - br * JS);
    const LfBranch& P = br ? a.br[1] : a.br[0];
    colp[4 * c + 0] = P.gamma[j];
    colp[4 * c + 1] = P.beta[j];
    colp[4 * c + 2] = a.stat[(size_t)(br * 2 + 0) * H + j];
    colp[4 * c + 3] = a.stat[(size_t)(br * 2 + 1) * H + j];
  }
  for (int k = tid; k < Kk; k += LC_T) kidx[k] = a.keep_idx[k];
  for (int d = tid; d < D; d += LC_T) {
    const int p = a.pos[d];
    if (p >= 0) wd[p] = d;
  }
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    copy4(sWa + br * JS * Kk, (br ? a.br[1].Wa : a.br[0].Wa) + (size_t)rk * JS * Kk, JS * Kk / 4, tid);
    const float* Wb = br ? a.br[1].Wb : a.br[0].Wb;
    const int q4 = JS / 4;
    for (int i = tid; i < Wn * q4; i += LC_T) {
      const int q = i % q4, w = i / q4;
      *reinterpret_cast<float4*>(sWbT + ((size_t)br * Wn + w) * JS + 4 * q) =
          __ldg(reinterpret_cast<const float4*>(Wb + (size_t)w * H + rk * JS + 4 * q));
    }
  }
  __syncthreads();
#pragma unroll 4
  for (int i = tid; i < 2 * BR * JS; i += LC_T) {
    const int jl = i % JS, b = (i / JS) % BR, br = i / (JS * BR), j = rk * JS + jl;
    float s = 0.f, ds = 0.f, xh = 0.f;
    if (b < B) {
      const float* cp = colp + 4 * (br * JS + jl);
      xh = (a.hpre[((size_t)br * B + b) * H + j] - cp[2]) * cp[3];
      const float z = fmaf(xh, cp[0], cp[1]);
      const float sg = sigmoid_(z);
      s = z * sg;
      ds = sg + z * sg * (1.f - sg);
    }
    Sl[i] = s;
    XH[i] = xh;
    DY[i] = ds;
  }
  for (int i = tid; i < 2 * (BR - B) * ldw; i += LC_T) {        // rows beyond the batch are zero operands
    const int br = i / ((BR - B) * ldw), rest = i - br * (BR - B) * ldw;
    DR[((size_t)br * BR + B) * ldw + rest] = 0.f;
  }
#pragma unroll 4
  for (int i = tid; i < BR * Kk; i += LC_T) {
    const int b = i / Kk, k = i - b * Kk;
    kept[b * ldk + k] = b < B ? (GS ? gS[(size_t)b * D + kidx[k]] : a.g[(size_t)b * D + kidx[k]]) : 0.f;
  }
  // ---- 1. the transform backward (same arithmetic as latent_affine_bwd_kernel) on the own Wn / 8 warped columns; the two
  //         d raw values of an item are pushed into every CTA's copy (distributed shared memory); dg of those positions ----
  lf_stamp(1, 1, rk, tid);
  pdl_wait();          // the cotangents come from the previous backward kernels; every global write of this kernel follows
  lf_stamp(1, 2, rk, tid);
  cluster.barrier_wait();       // every CTA of the cluster has started: remote writes may begin
  {
    const int WSb = Wn / LC_NC;
    for (int i = tid; i < B * WSb; i += LC_T) {
      const int p = rk * WSb + i % WSb, b = i / WSb;
      const size_t e = (size_t)b * D + wd[p];
      const float d = a.dgo ? a.dgo[e] : 0.f;
      const float rw = a.raw[((size_t)1 * B + b) * Wn + p];
      const float m = a.raw[((size_t)0 * B + b) * Wn + p];
      float dm = a.dmu_f ? a.dmu_f[e] : 0.f, dl = a.dlv_f ? a.dlv_f[e] : 0.f, dgv;
      const float ex = expf(rw);
      const float l = logf(a.eps + ex);
      const float gv = GS ? gS[e] : a.g[e];
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
      const float dlr = dl * ex / (a.eps + ex);
#pragma unroll
      for (int t = 0; t < LC_NC; ++t) {
        float* dst = cluster.map_shared_rank(DR, (unsigned)((rk + t) % LC_NC));
        dst[((size_t)0 * BR + b) * ldw + p] = dm;
        dst[((size_t)1 * BR + b) * ldw + p] = dlr;
      }
    }
  }
  cluster.sync();
  lf_stamp(1, 3, rk, tid);
  // ---- 2. dbb (its columns dealt over the CTAs);  dWb[w][own j] = sum_b d raw[b][w] s[b][j] ----
  for (int c = rk * (2 * Wn / LC_NC) + tid; c < (rk + 1) * (2 * Wn / LC_NC); c += LC_T) {
    const int br = c / Wn, w = c - br * Wn;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += DR[((size_t)br * BR + b) * ldw + w];
    (br ? a.gr[1].dbb : a.gr[0].dbb)[w] = s;
  }
  {
    const int q4 = JS / 4;
    for (int i = tid; i < 2 * Wn * q4; i += LC_T) {
      const int q = i % q4, w = (i / q4) % Wn, br = i / (q4 * Wn);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int b = 0; b < B; ++b) {
        const float d = DR[((size_t)br * BR + b) * ldw + w];
        const float4 s4 = *reinterpret_cast<const float4*>(Sl + ((size_t)br * BR + b) * JS + 4 * q);
        acc.x = fmaf(d, s4.x, acc.x); acc.y = fmaf(d, s4.y, acc.y); acc.z = fmaf(d, s4.z, acc.z); acc.w = fmaf(d, s4.w, acc.w);
      }
      *reinterpret_cast<float4*>((br ? a.gr[1].dWb : a.gr[0].dWb) + (size_t)w * H + rk * JS + 4 * q) = acc;
    }
  }
  lf_stamp(1, 4, rk, tid);
  // ---- 3. d y = swish'(y) * (d raw Wb) on the own columns: lanes = rows, K = Wn ----
  for (int c0 = warp * LC_CH; c0 < 2 * JS; c0 += (LC_T / 32) * LC_CH) {
    const int br = c0 / JS, jl = c0 - br * JS;
    float acc[RB][LC_CH];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int j = 0; j < LC_CH; ++j) acc[r][j] = 0.f;
    const float* dr = DR + (size_t)br * BR * ldw;
    const float* wt = sWbT + (size_t)br * Wn * JS + jl;
    for (int w = 0; w < Wn; w += 4) {
      float4 dv[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) dv[r] = *reinterpret_cast<const float4*>(dr + (size_t)(lane + 32 * r) * ldw + w);
#pragma unroll
      for (int ww = 0; ww < 4; ++ww) {
        const float4 wv = *reinterpret_cast<const float4*>(wt + (size_t)(w + ww) * JS);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const float d = ww == 0 ? dv[r].x : ww == 1 ? dv[r].y : ww == 2 ? dv[r].z : dv[r].w;
          acc[r][0] = fmaf(d, wv.x, acc[r][0]); acc[r][1] = fmaf(d, wv.y, acc[r][1]);
          acc[r][2] = fmaf(d, wv.z, acc[r][2]); acc[r][3] = fmaf(d, wv.w, acc[r][3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float4* dy = reinterpret_cast<float4*>(DY + ((size_t)br * BR + lane + 32 * r) * JS + jl);
      float4 v = *dy;
      v.x *= acc[r][0]; v.y *= acc[r][1]; v.z *= acc[r][2]; v.w *= acc[r][3];
      *dy = v;
    }
  }
  __syncthreads();
  lf_stamp(1, 5, rk, tid);
  // ---- 4. BatchNorm backward: four columns per warp item, lane = batch row ----
  for (int c0 = warp * 4; c0 < 2 * JS; c0 += (LC_T / 32) * 4) {
    const int br = c0 / JS, jl = c0 - br * JS, j = rk * JS + jl;
    float4 dz[RB], xh[RB];
    float dgam[4] = {0.f, 0.f, 0.f, 0.f}, dbet[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const size_t o = ((size_t)br * BR + lane + 32 * r) * JS + jl;
      dz[r] = *reinterpret_cast<const float4*>(DY + o);            // zero beyond the batch (swish' was stored as 0 there)
      xh[r] = *reinterpret_cast<const float4*>(XH + o);
      dgam[0] = fmaf(dz[r].x, xh[r].x, dgam[0]); dgam[1] = fmaf(dz[r].y, xh[r].y, dgam[1]);
      dgam[2] = fmaf(dz[r].z, xh[r].z, dgam[2]); dgam[3] = fmaf(dz[r].w, xh[r].w, dgam[3]);
      dbet[0] += dz[r].x; dbet[1] += dz[r].y; dbet[2] += dz[r].z; dbet[3] += dz[r].w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      dgam[q] = warp_sum(dgam[q]);
      dbet[q] = warp_sum(dbet[q]);
    }
    if (lane < 4) {
      const float gq = lane == 0 ? dgam[0] : lane == 1 ? dgam[1] : lane == 2 ? dgam[2] : dgam[3];
      const float bq = lane == 0 ? dbet[0] : lane == 1 ? dbet[1] : lane == 2 ? dbet[2] : dbet[3];
      (br ? a.gr[1].dgamma : a.gr[0].dgamma)[j + lane] = gq;
      (br ? a.gr[1].dbeta : a.gr[0].dbeta)[j + lane] = bq;
    }
    float istd[4], ga[4], m1[4], m2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      istd[q] = colp[4 * (c0 + q) + 3];
      ga[q] = colp[4 * (c0 + q) + 0];
      m1[q] = a.training ? ga[q] * dbet[q] / (float)B : 0.f;
      m2[q] = a.training ? ga[q] * dgam[q] / (float)B : 0.f;
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const bool in = lane + 32 * r < B;
      *reinterpret_cast<float4*>(DY + ((size_t)br * BR + lane + 32 * r) * JS + jl) =
          in ? make_float4(istd[0] * (dz[r].x * ga[0] - m1[0] - xh[r].x * m2[0]), istd[1] * (dz[r].y * ga[1] - m1[1] - xh[r].y * m2[1]),
                           istd[2] * (dz[r].z * ga[2] - m1[2] - xh[r].z * m2[2]), istd[3] * (dz[r].w * ga[3] - m1[3] - xh[r].w * m2[3]))
             : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  lf_stamp(1, 6, rk, tid);
  // ---- 5. dWa[own j][k] = sum_b d hpre[b][j] kept[b][k] ----
  {
    const int q4 = Kk / 4;
    for (int i = tid; i < 2 * JS * q4; i += LC_T) {
      const int q = i % q4, jl = (i / q4) % JS, br = i / (q4 * JS);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int b = 0; b < B; ++b) {
        const float dv = DY[((size_t)br * BR + b) * JS + jl];
        const float4 k4 = *reinterpret_cast<const float4*>(kept + (size_t)b * ldk + 4 * q);
        acc.x = fmaf(dv, k4.x, acc.x); acc.y = fmaf(dv, k4.y, acc.y); acc.z = fmaf(dv, k4.z, acc.z); acc.w = fmaf(dv, k4.w, acc.w);
      }
      *reinterpret_cast<float4*>((br ? a.gr[1].dWa : a.gr[0].dWa) + (size_t)(rk * JS + jl) * Kk + 4 * q) = acc;
    }
  }
  lf_stamp(1, 7, rk, tid);
  // ---- 6. partial d kept over the own hidden columns of both branches: lanes = rows, 4 columns k per warp item ----
  for (int k0 = warp * 4; k0 < Kk; k0 += (LC_T / 32) * 4) {
    float acc[RB][4];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
    for (int br = 0; br < 2; ++br) {
      for (int jl = 0; jl < JS; jl += 4) {
        float4 dv[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) dv[r] = *reinterpret_cast<const float4*>(DY + ((size_t)br * BR + lane + 32 * r) * JS + jl);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 wv = *reinterpret_cast<const float4*>(sWa + (size_t)(br * JS + jl + jj) * Kk + k0);
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const float d = jj == 0 ? dv[r].x : jj == 1 ? dv[r].y : jj == 2 ? dv[r].z : dv[r].w;
            acc[r][0] = fmaf(d, wv.x, acc[r][0]); acc[r][1] = fmaf(d, wv.y, acc[r][1]);
            acc[r][2] = fmaf(d, wv.z, acc[r][2]); acc[r][3] = fmaf(d, wv.w, acc[r][3]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r)
      *reinterpret_cast<float4*>(dkp + (size_t)(lane + 32 * r) * ldk + k0) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  lf_stamp(1, 8, rk, tid);
  cluster.sync();
  lf_stamp(1, 9, rk, tid);
  // ---- 7. reduce the eight partials on the own KS kept columns (distributed shared memory reads): dg[keep] = dgo[keep] + d kept ----
  for (int i = tid; i < B * KS; i += LC_T) {
    const int kl = i % KS, b = i / KS, k = rk * KS + kl;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < LC_NC; ++t) s += cluster.map_shared_rank(dkp, (unsigned)t)[(size_t)b * ldk + k];
    const size_t e = (size_t)b * D + kidx[k];
    a.dg[e] = (a.dgo ? a.dgo[e] : 0.f) + s;
  }
  lf_stamp(1, 10, rk, tid);
  cluster.sync();          // no CTA may exit while its partial is still being read
  lf_stamp(1, 11, rk, tid);
}

size_t lf_fwd_smem(int RB, int B, int D, int H, int Kk, int Wn) {
  const size_t BR = 32 * RB, JS = H / LC_NC, WS = Wn / LC_NC;
  return sizeof(float) * (BR * (Kk + 4) + 2 * JS * Kk + 2 * WS * H + 2 * BR * JS + 2 * BR * (H + 4) + 2 * BR * WS + (size_t)B * D + Kk + Wn + 8 * JS + 2 * WS);
}
size_t lf_bwd_smem(int RB, int B, int D, int H, int Kk, int Wn, bool gs) {
  const size_t BR = 32 * RB, JS = H / LC_NC;
  return sizeof(float) * (2 * BR * (Kk + 4) + 2 * BR * (Wn + 4) + 6 * BR * JS + 2 * (size_t)Wn * JS + 2 * JS * Kk + Kk + Wn + 8 * JS + (gs ? (size_t)B * D : 0));
}
constexpr size_t LF_SMEM_MAX = 227 * 1024;

template <typename K, typename A>
int lf_launch(K kernel, const A& a, size_t smem, cudaStream_t s, const char* what) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dpf_launch_pdl(kernel, LC_NC, LC_T, smem, s, a);      // cluster dims come from the kernel's __cluster_dims__
  return dpf_check_launch(what);
}

int lf_check_dims(const char* what, int B, int D, int H, int Kk, int Wn) {
  DPF_REQUIRE(B > 0 && D > 0 && H > 0 && Kk > 0 && Wn > 0 && Kk <= D && Wn <= D, DPF_ERR_BAD_ARG, "%s: bad sizes", what);
  DPF_REQUIRE(B <= 64 && H % 32 == 0 && Kk % 32 == 0 && Wn % 32 == 0 && D == Kk + Wn, DPF_ERR_UNSUPPORTED,
              "%s: needs B <= 64 and hidden / kept / warped widths multiples of 32 (got B=%d H=%d kept=%d warped=%d)", what, B, H, Kk, Wn);
  return DPF_OK;
}

bool lf_aligned(const LfBranch& b) { return (((uintptr_t)b.Wa | (uintptr_t)b.Wb | (uintptr_t)b.bb) & 15) == 0; }

}  // namespace

// development aid: globaltimer stamps (ns) of the phases of the most recent forward (row 0) and backward (row 1) launch
DPF_API int dpf_latent_flow_stamps(unsigned long long* out24) {
  DPF_REQUIRE(out24, DPF_ERR_NULL_PTR, "dpf_latent_flow_stamps: null pointer");
  cudaError_t e = cudaMemcpyFromSymbol(out24, g_lf_stamps, sizeof(unsigned long long) * 24);
  return e == cudaSuccess ? DPF_OK : (int)e;
}

// DPF_OK when dpf_latent_flow_forward AND _backward handle a layer of these sizes, DPF_ERR_UNSUPPORTED otherwise (the caller
// then keeps the block kernels + library GEMMs)
DPF_API int dpf_latent_flow_supported(int B, int D, int H, int Kk, int Wn) {
  int rc = lf_check_dims("dpf_latent_flow_supported", B, D, H, Kk, Wn);
  if (rc) return rc;
  const int RB = B <= 32 ? 1 : 2;
  DPF_REQUIRE(lf_fwd_smem(RB, B, D, H, Kk, Wn) <= LF_SMEM_MAX && lf_bwd_smem(RB, B, D, H, Kk, Wn, false) <= LF_SMEM_MAX, DPF_ERR_UNSUPPORTED,
              "dpf_latent_flow_supported: layer too wide for the shared memory of a CTA (B=%d D=%d H=%d)", B, D, H);
  return DPF_OK;
}

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
  DPF_REQUIRE((((uintptr_t)hpre | (uintptr_t)raw | (uintptr_t)g) & 15) == 0, DPF_ERR_ALIGN, "dpf_latent_flow_forward: g / hpre / raw alignment");
  a.B = B; a.D = D; a.H = H; a.Kk = Kk; a.Wn = Wn;
  a.bn_eps = bn_eps; a.momentum = momentum; a.eps = eps; a.training = training; a.inverse = inverse;
  a.g_out = g_out; a.mu = mu; a.lv = lv; a.hpre = hpre; a.stat = stat; a.raw = raw;
  const int RB = B <= 32 ? 1 : 2;
  const size_t smem = lf_fwd_smem(RB, B, D, H, Kk, Wn);
  DPF_REQUIRE(smem <= LF_SMEM_MAX, DPF_ERR_UNSUPPORTED, "dpf_latent_flow_forward: layer too wide (%zu bytes of shared memory per CTA)", smem);
  cudaStream_t s = (cudaStream_t)stream;
  return RB == 1 ? lf_launch(latent_flow_fwd_kernel<1>, a, smem, s, "latent_flow_fwd_kernel")
                 : lf_launch(latent_flow_fwd_kernel<2>, a, smem, s, "latent_flow_fwd_kernel");
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
    DPF_REQUIRE(lf_aligned(a.br[b]) && (((uintptr_t)dWa[b] | (uintptr_t)dWb[b]) & 15) == 0, DPF_ERR_ALIGN, "dpf_latent_flow_backward: weight matrices must be 16-byte aligned");
  }
  a.B = B; a.D = D; a.H = H; a.Kk = Kk; a.Wn = Wn; a.eps = eps; a.training = training; a.inverse = inverse;
  DPF_REQUIRE(((uintptr_t)g & 15) == 0, DPF_ERR_ALIGN, "dpf_latent_flow_backward: g alignment");
  a.hpre = hpre; a.stat = stat; a.raw = raw; a.dg = dg;
  const int RB = B <= 32 ? 1 : 2;
  const bool gs = lf_bwd_smem(RB, B, D, H, Kk, Wn, true) <= LF_SMEM_MAX;
  const size_t smem = lf_bwd_smem(RB, B, D, H, Kk, Wn, gs);
  DPF_REQUIRE(smem <= LF_SMEM_MAX, DPF_ERR_UNSUPPORTED, "dpf_latent_flow_backward: layer too wide (%zu bytes of shared memory per CTA)", smem);
  cudaStream_t s = (cudaStream_t)stream;
  if (RB == 1) return gs ? lf_launch(latent_flow_bwd_kernel<1, true>, a, smem, s, "latent_flow_bwd_kernel")
                         : lf_launch(latent_flow_bwd_kernel<1, false>, a, smem, s, "latent_flow_bwd_kernel");
  return gs ? lf_launch(latent_flow_bwd_kernel<2, true>, a, smem, s, "latent_flow_bwd_kernel")
            : lf_launch(latent_flow_bwd_kernel<2, false>, a, smem, s, "latent_flow_bwd_kernel");
}
