// Train-mode PointNet cloud encoder, the narrow layers (init_sd 3 -> 64, sd0 64 -> 128, sd1 128 -> 256) forward AND backward.
// Reference: PointNetCloudEncoder.features.{init_sd, sd0, sd1}{, _bn, _relu} in .train() (lib/networks/encoders.py:9-28) and
// torch.autograd through them; the library path is one bmm + a batch-norm + a clamp per layer forward and twice that backward
// (15 passes over (B, C, N) activations, fp32 SIMT GEMMs).
//
// What is stored per layer is only the PRE-BatchNorm output Z_l = W_l A_{l-1}; the activation A_l = relu(sc_l Z_l + sh_l) is
// recomputed by whoever loads Z_l (sc, sh = the folded batch statistics).  Layer 0 is never stored at all: its batch
// statistics follow analytically from the first and second moments of the input cloud (mean_c = W0[c] . E[x],
// var_c = W0[c]^T Cov(x) W0[c]), so A_0 is recomputed from the three coordinates wherever it is an operand.
//
// Two tcgen05 kernels, both with ONE CHANNEL PER TMEM LANE (a reduction over points is a per-thread register reduction, a
// store is 128 contiguous bytes per thread) and operand tiles [channels x 64 points] built by the CTA from fp32 global data
// as bf16 hi | lo images (three UMMA chains hi*hi + lo*hi + hi*lo: fp32-class accuracy):
//   pl_gemm_kernel    Out[M x pts] = Wimg[M x K] f(In[K x pts])      forward layers (f = BN + ReLU of the previous layer, output
//                     statistics in the epilogue) and the backward dgrads (f = BatchNorm/ReLU backward of (dA, Z), Wimg = W^T)
//   pl_wgrad_kernel   D[MP x NQ] += sum_pts P[MP x pts] Q[NQ x pts]^T   weight gradients (P = BN/ReLU backward of (dA, Z),
//                     Q = recomputed activations) and the Gram matrix of the last layer's analytic backward (P = Q)
// The [channels x 64 points] tile with 128-byte swizzled rows is at the same time a K-major operand over the points (wgrad) and
// an MN-major operand over the channels (forward / dgrad): one shared-memory image serves both descriptors.
// Plus two streaming reductions (BatchNorm-backward sums; layer 0's backward sums).
#include "common.cuh"
#include "umma.cuh"
#include "pointnet_tiles.cuh"
#include <math_constants.h>

namespace {

constexpr int PL_T = pnt::T;                    // threads per CTA
constexpr int PL_NT = pnt::NT;                  // points per tile
constexpr uint32_t PL_KB = 128 * 128;           // 16 KB: [128 rows x 64 bf16] block of a K-major image
constexpr int PL_TAB = pnt::TAB;                // floats per channel in a loader table (pointnet_tiles.cuh)
constexpr uint64_t PL_DESC_K = umma::make_desc_template(16, 1024, umma::LAYOUT_SW128);        // K-major, 8-row groups 1024 B apart
constexpr uint64_t PL_DESC_MN = umma::make_desc_template(1024, 1024, umma::LAYOUT_SW128);     // MN-major, one 64-wide block

using pnt::LD_X3; using pnt::LD_AFFINE; using pnt::LD_BNBWD;

// ---------------------------------------------------------------------------------------------------------------
// weight images: matrix [R x K] (element (r, k) at W[r * rs + k * cs], so that W and W^T share the code) ->
// [ceil(R / 128) chunks][hi, lo][K / 64 blocks][128 x 64] swizzled bf16, rows beyond R zero
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pl_pack_kernel(const float* __restrict__ W, int R, int K, long long rs, long long cs, unsigned char* __restrict__ img) {
  const int qpr = K / 8;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int rpad = ((R + 127) / 128) * 128;
  if (i >= rpad * qpr) return;
  const int r = i / qpr, q = i % qpr;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = r < R ? W[(long long)r * rs + (long long)(q * 8 + 2 * e) * cs] : 0.f;
    const float b = r < R ? W[(long long)r * rs + (long long)(q * 8 + 2 * e + 1) * cs] : 0.f;
    hi[e] = umma::pack_bf16(a, b);
    lo[e] = umma::pack_bf16(a - __uint_as_float(hi[e] << 16), b - __uint_as_float(hi[e] & 0xffff0000u));
  }
  const uint32_t half = (uint32_t)(K / 64) * PL_KB;
  unsigned char* base = img + (size_t)(r >> 7) * 2 * half + (size_t)(q >> 3) * PL_KB + umma::sw128_offset(r & 127, q & 7);
  *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(base + half) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// CTAs along the point tiles: the kernels hold an SM's shared memory (nearly) alone, so one CTA per SM and output chunk in
// total, each with a contiguous range of the B * ceil(N / 64) tiles (at least 2 per CTA: the weight image load and the
// weight gradients' atomic flush are amortised)
int pl_ctas(int B, int N, int chunks) {
  const int total = B * ((N + PL_NT - 1) / PL_NT);
  return max(1, min((total + 1) / 2, max(1, dpf_num_sms() / chunks)));
}

// ---------------------------------------------------------------------------------------------------------------
// Out[M x pts] = Wimg[M x K] f(In[K x pts]) per tile of 64 points; CTA = (shape b, group of tiles) x (chunk of 128 output rows)
// ---------------------------------------------------------------------------------------------------------------
struct PlGemmArgs {
  const float* in0;             // LD_X3: x (B,3,N); LD_AFFINE: Z (B,K,N); LD_BNBWD: dA (B,K,N)
  const float* in1;             // LD_BNBWD: Z (B,K,N)
  const float* tab;             // (K, PL_TAB)
  const unsigned char* wimg;    // pl_pack_kernel image of the [Mout x K] matrix
  float* out;                   // (B, Mout, N) or null
  const float* row_off;         // (Mout,) added to every element of a row, or null
  float* stat;                  // (gridDim.x, Mout, 3) {count, mean, sum of squared deviations} or null
  int B, N, Mout;
};

template <int K, int LOADER, bool STATS>
__global__ void __launch_bounds__(PL_T, 1)
pl_gemm_kernel(const PlGemmArgs a) {
  constexpr int NST = K == 256 ? 1 : 2;           // operand stages (K = 256: the 128 KB weight image leaves room for one)
  constexpr int LAG = NST - 1;                    // the epilogue of tile i - LAG runs under the UMMAs of tile i
  constexpr uint32_t WHALF = (K / 64) * PL_KB, BHALF = K * 128;
  extern __shared__ unsigned char smraw[];
  unsigned char* sm = smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u);
  unsigned char* sA = sm;                                       // weight chunk hi | lo
  unsigned char* sB = sm + 2 * WHALF;                           // NST x (tile hi | lo)
  float (*red)[128][4] = reinterpret_cast<float (*)[128][4]>(sB + NST * 2 * BHALF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(red) + 2 * 128 * 4 * sizeof(float));   // [0] weights, [1..2] TMEM buffers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int lane_c = tid & 127, part = tid >> 7, quarter = warp & 3;
  const int chunk = blockIdx.y;
  // the CTA's contiguous range of the B * n_tiles point tiles (tiles of all shapes in one flat index: balanced to one tile)
  const int n_tiles = (a.N + PL_NT - 1) / PL_NT;
  const int total = a.B * n_tiles;
  const int t0 = (int)((long long)total * blockIdx.x / gridDim.x);
  const int nt = (int)((long long)total * (blockIdx.x + 1) / gridDim.x) - t0;
  if (tid == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::mbar_init(&bars[2], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 128);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
  if (tid == 0) {
    umma::mbar_expect_tx(&bars[0], 2 * WHALF);
    umma::bulk_g2s(sA, a.wimg + (size_t)chunk * 2 * WHALF, 2 * WHALF, &bars[0]);
  }
  typename pnt::LoaderFor<K, LOADER>::type ld;
  ld.init(a.in0, a.in1, a.tab, a.N, tid);
  constexpr uint32_t IDESC = umma::make_idesc_bf16(128, PL_NT, 0, 1);

  const int row = chunk * 128 + lane_c;
  const bool row_ok = row < a.Mout;
  const float roff = (a.row_off && row_ok) ? a.row_off[row] : 0.f;
  float sum = 0.f, sq = 0.f, shift = 0.f;
  int cnt = 0;
  uint32_t par[2] = {0u, 0u};
  const bool vec_ok = (a.N % 4) == 0;

  if (nt > 0) ld.load(t0);
  for (int i = 0; i < nt + LAG; ++i) {
    if (i < nt) {
      unsigned char* st = sB + (size_t)(i % NST) * 2 * BHALF;
      ld.template convert<false, false>(st, st + BHALF, t0 + i, nullptr);     // tail columns are never read back
      umma::fence_async_smem();
      if (i + 1 < nt) ld.load(t0 + i + 1);       // in flight under the UMMAs and the epilogue
    }
    umma::fence_before_sync();
    __syncthreads();
    if (i < nt && tid == 0) {
      if (i == 0) umma::mbar_wait(&bars[0], 0);
      umma::fence_after_sync();
      const uint32_t d = tmem + (uint32_t)(i & 1) * PL_NT;
      const uint32_t b0 = umma::smem_u32(sB + (size_t)(i % NST) * 2 * BHALF);
#pragma unroll
      for (int chain = 0; chain < 3; ++chain) {           // hi*hi, lo*hi, hi*lo
        const uint32_t a_base = umma::smem_u32(sA) + (chain == 1 ? WHALF : 0);
        const uint32_t b_base = b0 + (chain == 2 ? BHALF : 0);
#pragma unroll
        for (int kb = 0; kb < K / 64; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma::mma_bf16(d, umma::desc_at(PL_DESC_K, a_base + kb * PL_KB + 32 * k),
                           umma::desc_at(PL_DESC_MN, b_base + (kb * 4 + k) * 2048), IDESC, (chain | kb | k) > 0);
      }
      umma::mma_commit(&bars[1 + (i & 1)]);
    }
    const int e = i - LAG;
    if (e >= 0) {
      umma::mbar_wait(&bars[1 + (e & 1)], par[e & 1]);
      par[e & 1] ^= 1u;
      umma::fence_after_sync();
      float v[32];
      umma::tmem_ld32(tmem + lane_off + (uint32_t)(e & 1) * PL_NT + part * 32, v);
      const int b = (t0 + e) / n_tiles;
      const int nbase = (t0 + e - b * n_tiles) * PL_NT + part * 32;
      if (STATS) {
        if (nbase + 32 <= a.N) {      // all 32 columns valid: no per-element bookkeeping
          if (cnt == 0) shift = v[0];
          cnt += 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float dx = v[j] - shift;
            sum += dx;
            sq = fmaf(dx, dx, sq);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nbase + j < a.N) {
              const float x = v[j];
              if (cnt == 0) shift = x;
              ++cnt;
              const float dx = x - shift;
              sum += dx;
              sq = fmaf(dx, dx, sq);
            }
          }
        }
      }
      if (a.out && row_ok) {
        float* dst = a.out + ((size_t)b * a.Mout + row) * a.N + nbase;
        if (vec_ok && nbase + 32 <= a.N) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j] + roff, v[4 * j + 1] + roff, v[4 * j + 2] + roff, v[4 * j + 3] + roff);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nbase + j < a.N) dst[j] = v[j] + roff;
        }
      }
    }
  }
  if (STATS) {
    // this thread: count, mean = shift + sum / cnt, M2 = sq - sum^2 / cnt; the two point halves merge pairwise (Chan et al.)
    const float fc = (float)cnt;
    const float mean_t = cnt ? shift + sum / fc : 0.f;
    const float m2_t = cnt ? fmaxf(sq - sum * sum / fc, 0.f) : 0.f;
    red[part][lane_c][0] = fc; red[part][lane_c][1] = mean_t; red[part][lane_c][2] = m2_t;
    __syncthreads();
    if (part == 0 && row_ok) {
      const float n2 = red[1][lane_c][0], mean2 = red[1][lane_c][1], m22 = red[1][lane_c][2];
      const float ntot = fc + n2;
      const float delta = mean2 - mean_t;
      float* o = a.stat + ((size_t)blockIdx.x * a.Mout + row) * 3;
      o[0] = ntot;
      o[1] = ntot > 0.f ? mean_t + delta * (n2 / ntot) : 0.f;
      o[2] = ntot > 0.f ? m2_t + m22 + delta * delta * (fc * n2 / ntot) : 0.f;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

template <int K>
constexpr size_t pl_gemm_smem() {
  return 1024 + 2 * (size_t)(K / 64) * PL_KB + (K == 256 ? 1 : 2) * 2 * (size_t)K * 128 + 2 * 128 * 4 * sizeof(float) + 64;
}

template <int K, int LOADER, bool STATS>
int pl_launch_gemm(const PlGemmArgs& a, cudaStream_t s) {
  constexpr size_t smem = pl_gemm_smem<K>();
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pl_gemm_kernel<K, LOADER, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  pl_gemm_kernel<K, LOADER, STATS><<<dim3(pl_ctas(a.B, a.N, (a.Mout + 127) / 128), (a.Mout + 127) / 128), PL_T, smem, s>>>(a);
  return dpf_check_launch("pl_gemm_kernel");
}

// ---------------------------------------------------------------------------------------------------------------
// D[MP x NQ] += sum over the CTA's points of P[MP x pts] Q[NQ x pts]^T ; accumulated in TMEM over the CTA's tiles, then
// written as this CTA's partial sum (MP x NQ row-major fp32); a small kernel adds the partials to `out`.  GRAM: Q is P
// (one image, MP == NQ).
// ---------------------------------------------------------------------------------------------------------------
struct PlWgradArgs {
  const float* p_in0; const float* p_in1; const float* p_tab;
  const float* q_in0; const float* q_tab;
  float* partial;               // (gridDim.x, MP, NQ): this CTA's sum over its tiles (summed by pl_reduce_partials_kernel)
  int B, N;
};

template <int MP, int PLOADER, int NQ, int QLOADER, bool GRAM>
__global__ void __launch_bounds__(PL_T, 1)
pl_wgrad_kernel(const PlWgradArgs a) {
  constexpr uint32_t PHALF = MP * 128, QHALF = GRAM ? 0 : NQ * 128;
  constexpr uint32_t STAGE = 2 * PHALF + 2 * QHALF;
  constexpr uint32_t COLS = (MP / 128) * NQ;                   // fp32 accumulator columns
  constexpr uint32_t TCOLS = COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
  extern __shared__ unsigned char smraw[];
  unsigned char* sm = smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * STAGE);      // [0..1] stage free, [2] all done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int lane_c = tid & 127, part = tid >> 7, quarter = warp & 3;
  const int n_tiles = (a.N + PL_NT - 1) / PL_NT;
  const int total = a.B * n_tiles;
  const int t0 = (int)((long long)total * blockIdx.x / gridDim.x);
  const int nt = (int)((long long)total * (blockIdx.x + 1) / gridDim.x) - t0;
  if (tid == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::mbar_init(&bars[2], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, TCOLS);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;

  typename pnt::LoaderFor<MP, PLOADER>::type lp;
  lp.init(a.p_in0, a.p_in1, a.p_tab, a.N, tid);
  typename pnt::LoaderFor<GRAM ? MP : NQ, GRAM ? PLOADER : QLOADER>::type lq;
  if (!GRAM) lq.init(a.q_in0, nullptr, a.q_tab, a.N, tid);
  constexpr uint32_t IDESC = umma::make_idesc_bf16(128, NQ, 0, 0);
  uint32_t par[2] = {0u, 0u};

  if (nt > 0) {
    lp.load(t0);
    if (!GRAM) lq.load(t0);
  }
  for (int i = 0; i < nt; ++i) {
    unsigned char* st = sm + (size_t)(i & 1) * STAGE;
    if (i >= 2) {                                   // the UMMAs of tile i - 2 read this stage
      umma::mbar_wait(&bars[i & 1], par[i & 1]);
      par[i & 1] ^= 1u;
    }
    lp.template convert<true, false>(st, st + PHALF, t0 + i, nullptr);          // points beyond N must not enter the sums
    if (!GRAM) lq.template convert<true, false>(st + 2 * PHALF, st + 2 * PHALF + QHALF, t0 + i, nullptr);
    umma::fence_async_smem();
    if (i + 1 < nt) {
      lp.load(t0 + i + 1);
      if (!GRAM) lq.load(t0 + i + 1);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t p0 = umma::smem_u32(st), q0 = GRAM ? p0 : umma::smem_u32(st + 2 * PHALF);
      constexpr uint32_t QH = GRAM ? PHALF : QHALF;
#pragma unroll
      for (int mc = 0; mc < MP / 128; ++mc) {
#pragma unroll
        for (int chain = 0; chain < 3; ++chain) {         // hi*hi, lo*hi, hi*lo
          const uint32_t pa = p0 + (chain == 1 ? PHALF : 0) + mc * PL_KB;
          const uint32_t qa = q0 + (chain == 2 ? QH : 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma::mma_bf16(tmem + mc * NQ, umma::desc_at(PL_DESC_K, pa + 32 * k), umma::desc_at(PL_DESC_K, qa + 32 * k), IDESC,
                           (i | chain | k) > 0);
        }
      }
      umma::mma_commit(i + 1 < nt ? &bars[i & 1] : &bars[2]);
    }
  }
  if (nt > 0) {
    umma::mbar_wait(&bars[2], 0);
    umma::fence_after_sync();
  }
  // lane = P row; this part stores half of the columns, 32 at a time (128 contiguous bytes per thread).  Per-CTA partials
  // + a reduction kernel instead of float atomics: 148 CTAs x 32 768 .. 65 536 atomics on the same addresses were half
  // of this kernel's time (ncu: lg_throttle on the REDs).
#pragma unroll 1
  for (int mc = 0; mc < MP / 128; ++mc) {
#pragma unroll 1
    for (int c0 = part * 32; c0 < NQ; c0 += 64) {
      float v[32];
      if (nt > 0) {
        umma::tmem_ld32(tmem + lane_off + mc * NQ + c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(a.partial + ((size_t)blockIdx.x * MP + mc * 128 + lane_c) * NQ + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, TCOLS);
}

// out[j] += sum over the CTAs' partials; the CTA range is split over blockIdx.y (4 atomics per address instead of 148)
__global__ void __launch_bounds__(128)
pl_reduce_partials_kernel(const float* __restrict__ partial, int n_ctas, int n4, float* __restrict__ out) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= n4) return;
  const int c0 = (int)((long long)n_ctas * blockIdx.y / gridDim.y), c1 = (int)((long long)n_ctas * (blockIdx.y + 1) / gridDim.y);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int c = c0; c < c1; ++c) {
    const float4 v = reinterpret_cast<const float4*>(partial)[(size_t)c * n4 + j];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  atomicAdd(out + 4 * j + 0, acc.x);
  atomicAdd(out + 4 * j + 1, acc.y);
  atomicAdd(out + 4 * j + 2, acc.z);
  atomicAdd(out + 4 * j + 3, acc.w);
}

template <int MP, int PLOADER, int NQ, int QLOADER, bool GRAM>
int pl_launch_wgrad(const PlWgradArgs& a, float* out, cudaStream_t s) {
  constexpr size_t smem = 1024 + 2 * (2 * (size_t)MP * 128 + (GRAM ? 0 : 2 * (size_t)NQ * 128)) + 64;
  static_assert(smem <= 232448, "wgrad stages exceed the shared memory of an SM");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pl_wgrad_kernel<MP, PLOADER, NQ, QLOADER, GRAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const int ctas = pl_ctas(a.B, a.N, 1);
  pl_wgrad_kernel<MP, PLOADER, NQ, QLOADER, GRAM><<<ctas, PL_T, smem, s>>>(a);
  int rc = dpf_check_launch("pl_wgrad_kernel");
  if (rc) return rc;
  const int n4 = MP * NQ / 4;
  pl_reduce_partials_kernel<<<dim3((n4 + 127) / 128, 4), 128, 0, s>>>(a.partial, ctas, n4, out);
  return dpf_check_launch("pl_reduce_partials_kernel");
}

// ---------------------------------------------------------------------------------------------------------------
// streaming reductions.  One block per (shape, channel) row of N points.
//   BatchNorm + ReLU backward sums:  s[c] = {sum dA [y > 0], sum dA [y > 0] (z - mu)},  y = sc z + sh
//   layer 0 backward sums:           s[c] = {sum dA m, sum dA m x0, sum dA m x1, sum dA m x2},  m = [a . x + c > 0]
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_add_double(double* dst, const float* v, int nv, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = 0; j < nv; ++j) {
    const double w = warp_sum_d((double)v[j]);
    if (lane == 0) scratch[warp * 4 + j] = w;
  }
  __syncthreads();
  if (threadIdx.x < nv) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w * 4 + threadIdx.x];
    atomicAdd(dst + threadIdx.x, t);
  }
}

__global__ void __launch_bounds__(128)
pl_bn_bwd_sums_kernel(const float* __restrict__ dA, const float* __restrict__ Z, const float* __restrict__ tab, int C, int N,
                      double* __restrict__ sums) {
  __shared__ double scratch[4 * 4];
  const int c = blockIdx.x % C;
  const size_t base = (size_t)blockIdx.x * N;
  const float sc = tab[c * PL_TAB + 0], sh = tab[c * PL_TAB + 1], mu = tab[c * PL_TAB + 5];
  float acc[2] = {0.f, 0.f};
  for (int n = threadIdx.x; n < N; n += 128) {
    const float z = Z[base + n], d = dA[base + n];
    const float dm = fmaf(sc, z, sh) > 0.f ? d : 0.f;
    acc[0] += dm;
    acc[1] = fmaf(dm, z - mu, acc[1]);
  }
  block_add_double(sums + (size_t)c * 2, acc, 2, scratch);
}

__global__ void __launch_bounds__(128)
pl_layer0_bwd_sums_kernel(const float* __restrict__ dA, const float* __restrict__ x, const float* __restrict__ tab, int C, int N,
                          double* __restrict__ sums) {
  __shared__ double scratch[4 * 4];
  const int c = blockIdx.x % C, b = blockIdx.x / C;
  const size_t base = (size_t)blockIdx.x * N;
  const float* xb = x + (size_t)b * 3 * N;
  const float a0 = tab[c * PL_TAB + 0], a1 = tab[c * PL_TAB + 1], a2 = tab[c * PL_TAB + 2], cc = tab[c * PL_TAB + 3];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n = threadIdx.x; n < N; n += 128) {
    const float x0 = xb[n], x1 = xb[N + n], x2 = xb[2 * (size_t)N + n];
    const float dm = fmaf(a0, x0, fmaf(a1, x1, fmaf(a2, x2, cc))) > 0.f ? dA[base + n] : 0.f;
    acc[0] += dm;
    acc[1] = fmaf(dm, x0, acc[1]);
    acc[2] = fmaf(dm, x1, acc[2]);
    acc[3] = fmaf(dm, x2, acc[3]);
  }
  block_add_double(sums + (size_t)c * 4, acc, 4, scratch);
}


// ---------------------------------------------------------------------------------------------------------------
// per-channel finalisation kernels (one thread per channel, double arithmetic): what would otherwise be ~150 tiny
// library launches per training step between the big kernels
// ---------------------------------------------------------------------------------------------------------------
// moments of the cloud: sums[0..2] = sum x_i, sums[3..8] = sum x_i x_j (00 01 02 11 12 22) over all B * N points
__global__ void __launch_bounds__(256)
pl_input_moments_kernel(const float* __restrict__ x, int B, int N, double* __restrict__ sums) {
  __shared__ double scratch[8 * 9];
  double acc[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) acc[j] = 0.0;
  const long long total = (long long)B * N;
  for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < total; p += (long long)gridDim.x * 256) {
    const int b = (int)(p / N), n = (int)(p - (long long)b * N);
    const float* xb = x + (size_t)b * 3 * N + n;
    const double x0 = xb[0], x1 = xb[N], x2 = xb[2 * (size_t)N];
    acc[0] += x0; acc[1] += x1; acc[2] += x2;
    acc[3] += x0 * x0; acc[4] += x0 * x1; acc[5] += x0 * x2; acc[6] += x1 * x1; acc[7] += x1 * x2; acc[8] += x2 * x2;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const double w = warp_sum_d(acc[j]);
    if (lane == 0) scratch[warp * 9 + j] = w;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += scratch[w * 9 + threadIdx.x];
    atomicAdd(sums + threadIdx.x, t);
  }
}

__device__ __forceinline__ void load_moments(const double* sums, double M, double xm[3], double cov[3][3]) {
  for (int i = 0; i < 3; ++i) xm[i] = sums[i] / M;
  const double s2[3][3] = {{sums[3], sums[4], sums[5]}, {sums[4], sums[6], sums[7]}, {sums[5], sums[7], sums[8]}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cov[i][j] = s2[i][j] / M - xm[i] * xm[j];      // double: no cancellation issue at |x| ~ 1
}

// layer 0: analytic batch statistics mean_c = W0[c] . E[x], var_c = W0[c]^T Cov(x) W0[c] -> loader table {sc W0[c,:], beta - sc mean},
// stats (C, 3) fp32 {mean, biased var, istd}
__global__ void pl_layer0_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ W0, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, int C, double M, float eps, float* __restrict__ tab,
                                          float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double xm[3], cov[3][3];
  load_moments(sums, M, xm, cov);
  const double w[3] = {W0[c * 3], W0[c * 3 + 1], W0[c * 3 + 2]};
  double mean = 0.0, var = 0.0;
  for (int i = 0; i < 3; ++i) {
    mean += w[i] * xm[i];
    for (int j = 0; j < 3; ++j) var += w[i] * cov[i][j] * w[j];
  }
  var = fmax(var, 0.0);
  const double istd = rsqrt(var + (double)eps);
  const double sc = (double)gamma[c] * istd;
  float* t = tab + (size_t)c * PL_TAB;
  t[0] = (float)(sc * w[0]); t[1] = (float)(sc * w[1]); t[2] = (float)(sc * w[2]); t[3] = (float)((double)beta[c] - sc * mean);
  t[4] = 0.f; t[5] = 0.f; t[6] = 0.f; t[7] = 0.f;
  stats[c * 3] = (float)mean; stats[c * 3 + 1] = (float)var; stats[c * 3 + 2] = (float)istd;
}

// merge of the per-CTA {count, mean, M2} triples (Chan et al.) -> loader table {sc, sh, 0, 0, 0, mean}, stats (C, 3) {mean, biased var, istd}.
// stride3 = 3: stat (G, C, 3) of pl_gemm_kernel; stride3 = 2: stat (G, C, 2) {mean, M2} of the pool kernel with `count` points each.
__global__ void __launch_bounds__(128)
pl_stats_finalize_kernel(const float* __restrict__ stat, int G, int C, int stride3, float count, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, float* __restrict__ tab, float* __restrict__ stats) {
  // one WARP per channel, the G groups strided over its lanes; two passes of plain sums (no serial dependency):
  //   n = sum n_g, mean = sum n_g m_g / n, M2 = sum [M2_g + n_g (m_g - mean)^2]
  const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double n = 0.0, nm = 0.0;
  for (int g = lane; g < G; g += 32) {
    const float* e = stat + ((size_t)g * C + c) * stride3;
    const double ng = stride3 == 3 ? (double)e[0] : (double)count;
    n += ng;
    nm += ng * (double)e[stride3 - 2];
  }
  n = warp_sum_d(n);
  nm = warp_sum_d(nm);
  const double mean = n > 0.0 ? nm / n : 0.0;
  double m2 = 0.0;
  for (int g = lane; g < G; g += 32) {
    const float* e = stat + ((size_t)g * C + c) * stride3;
    const double ng = stride3 == 3 ? (double)e[0] : (double)count;
    const double d = (double)e[stride3 - 2] - mean;
    m2 += (double)e[stride3 - 1] + ng * d * d;
  }
  m2 = warp_sum_d(m2);
  if (lane != 0) return;
  const double var = n > 0.0 ? fmax(m2 / n, 0.0) : 0.0;
  const double istd = rsqrt(var + (double)eps);
  if (tab) {
    const double sc = (double)gamma[c] * istd;
    float* t = tab + (size_t)c * PL_TAB;
    t[0] = (float)sc; t[1] = (float)((double)beta[c] - sc * mean); t[2] = 0.f; t[3] = 0.f; t[4] = 0.f; t[5] = (float)mean; t[6] = 0.f; t[7] = 0.f;
  }
  stats[c * 3] = (float)mean; stats[c * 3 + 1] = (float)var; stats[c * 3 + 2] = (float)istd;
}

// BatchNorm backward of a layer from the two sums: dbeta = s1, dgamma = s2 istd; loader-2 table {sc, sh, g, g m1, g m2 istd, mu}
__global__ void pl_bwd_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ tab_fwd, const float* __restrict__ gamma,
                                       const float* __restrict__ stats, int C, double M, float* __restrict__ tab_bwd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double istd = stats[c * 3 + 2];
  const double s1 = sums[c * 2], dg = sums[c * 2 + 1] * istd;
  const double g = (double)gamma[c] * istd;
  const float* tf = tab_fwd + (size_t)c * PL_TAB;
  float* t = tab_bwd + (size_t)c * PL_TAB;
  t[0] = tf[0]; t[1] = tf[1]; t[2] = (float)g; t[3] = (float)(g * s1 / M); t[4] = (float)(g * dg / M * istd); t[5] = tf[5]; t[6] = 0.f; t[7] = 0.f;
  dgamma[c] = (float)dg;
  dbeta[c] = (float)s1;
}

// layer 0 backward from its four sums per channel {S, T_0, T_1, T_2} and the input moments
__global__ void pl_layer0_bwd_finalize_kernel(const double* __restrict__ sums4, const double* __restrict__ moments, const float* __restrict__ W0,
                                              const float* __restrict__ gamma, const float* __restrict__ stats, int C, double M,
                                              float* __restrict__ dW0, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double xm[3], cov[3][3];
  load_moments(moments, M, xm, cov);
  const double istd = stats[c * 3 + 2];
  const double w[3] = {W0[c * 3], W0[c * 3 + 1], W0[c * 3 + 2]};
  const double S = sums4[c * 4];
  double Tc[3], dg = 0.0;
  for (int j = 0; j < 3; ++j) {
    Tc[j] = sums4[c * 4 + 1 + j] - S * xm[j];          // sum dy (x_j - mean_j)
    dg += w[j] * Tc[j];
  }
  dg *= istd;
  const double gi = (double)gamma[c] * istd;
  for (int j = 0; j < 3; ++j) {
    double wc = 0.0;
    for (int i = 0; i < 3; ++i) wc += w[i] * cov[i][j];
    dW0[c * 3 + j] = (float)(gi * (Tc[j] - dg * istd * wc));      // m2 istd M (W0 cov)_j with m2 = dg / M
  }
  dgamma[c] = (float)dg;
  dbeta[c] = (float)S;
}


// ---------------------------------------------------------------------------------------------------------------
// the SPARSE part of the pooled last layer's backward (ops/pointnet_pool.py): the max-pool routes the cotangent of
// (shape b, channel c) to ONE point n* = idx[b][c].  Per (b, c) and input channel k:
//     a = relu(sc_k Z[b][k][n*] + sh_k)          the layer input at the selected point (recomputed)
//     T[c][k]       += coef[b][c] a              (weight-gradient term; summed over the shapes)
//     dA[b][k][n*]  += coef[b][c] W[c][k]        (input-gradient term, on top of the dense part already in dA)
// instead of gather -> affine -> relu -> einsum and outer product -> scatter_add over three (B,256,512) temporaries.
// Block = (shape, 8 channels), thread = input channel k; several channels of a shape often select the same point: atomics.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PS_CG = 8;
__global__ void __launch_bounds__(256)
pl_pool_sparse_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ tab, const long long* __restrict__ idx,
                          const float* __restrict__ coef, const float* __restrict__ W, int N, int C, float* __restrict__ T,
                          float* __restrict__ dA) {
  const int k = threadIdx.x, b = blockIdx.y, c0 = blockIdx.x * PS_CG;
  const float sc = tab[k * PL_TAB], sh = tab[k * PL_TAB + 1];
  const float* zrow = Z + ((size_t)b * 256 + k) * N;
  float* drow = dA + ((size_t)b * 256 + k) * N;
  float zv[PS_CG], cf[PS_CG];
  int np[PS_CG];
#pragma unroll
  for (int j = 0; j < PS_CG; ++j) {           // all gathers of the block's channels in flight together
    np[j] = (int)idx[(size_t)b * C + c0 + j];
    cf[j] = coef[(size_t)b * C + c0 + j];
    zv[j] = zrow[np[j]];
  }
#pragma unroll
  for (int j = 0; j < PS_CG; ++j) {
    if (cf[j] != 0.f) {
      const float a = fmaxf(fmaf(sc, zv[j], sh), 0.f);
      if (a != 0.f) atomicAdd(T + (size_t)(c0 + j) * 256 + k, cf[j] * a);
      atomicAdd(drow + np[j], cf[j] * W[(size_t)(c0 + j) * 256 + k]);
    }
  }
}

}  // namespace

// ===============================================================================================================
// C ABI
// ===============================================================================================================
// bytes of the packed image of an [R x K] matrix (K in {64, 128, 256})
DPF_API int dpf_pointnet_layer_image_bytes(int R, int K, long long* bytes) {
  DPF_REQUIRE(bytes, DPF_ERR_NULL_PTR, "dpf_pointnet_layer_image_bytes: null out pointer");
  DPF_REQUIRE(R > 0 && (K == 64 || K == 128 || K == 256), DPF_ERR_BAD_ARG, "dpf_pointnet_layer_image_bytes: R=%d K=%d", R, K);
  *bytes = (long long)((R + 127) / 128) * 2 * (K / 64) * PL_KB;
  return DPF_OK;
}

// W: matrix [R x K], element (r, k) at W[r * row_stride + k * col_stride] (so W^T needs no copy) -> image (256-byte aligned)
DPF_API int dpf_pointnet_layer_pack(const float* W, int R, int K, long long row_stride, long long col_stride, void* image, void* stream) {
  DPF_REQUIRE(W && image, DPF_ERR_NULL_PTR, "dpf_pointnet_layer_pack: null pointer");
  DPF_REQUIRE(R > 0 && (K == 64 || K == 128 || K == 256), DPF_ERR_BAD_ARG, "dpf_pointnet_layer_pack: R=%d K=%d", R, K);
  DPF_REQUIRE(((uintptr_t)image & 255) == 0, DPF_ERR_ALIGN, "dpf_pointnet_layer_pack: image alignment");
  const int n = ((R + 127) / 128) * 128 * (K / 8);
  pl_pack_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, R, K, row_stride, col_stride, (unsigned char*)image);
  return dpf_check_launch("pl_pack_kernel");
}

// rows of the stat buffer of dpf_pointnet_layer_gemm: one per CTA along the point tiles
DPF_API int dpf_pointnet_layer_groups(int B, int N, int Mout, int* groups) {
  DPF_REQUIRE(groups, DPF_ERR_NULL_PTR, "dpf_pointnet_layer_groups: null out pointer");
  DPF_REQUIRE(B > 0 && N > 0 && Mout > 0, DPF_ERR_BAD_ARG, "dpf_pointnet_layer_groups: B=%d N=%d Mout=%d", B, N, Mout);
  *groups = pl_ctas(B, N, (Mout + 127) / 128);
  return DPF_OK;
}

// Out (B, Mout, N) = image[Mout x K] f(in) with the loader `loader` (0 layer 0 from x (B,3,N), K = 64; 1 relu(sc z + sh) - sub of
// in0 (B,K,N); 2 BatchNorm + ReLU backward of (in0 = dA, in1 = Z), both (B,K,N)); tab (K, 8) fp32 per-channel loader constants;
// out nullable; row_off (Mout,) nullable; stat (dpf_pointnet_layer_groups, Mout, 3) {count, mean, M2} nullable (forward layers).
DPF_API int dpf_pointnet_layer_gemm(int loader, int K, const float* in0, const float* in1, const float* tab, const void* image,
                                    int B, int N, int Mout, float* out, const float* row_off, float* stat, void* stream) {
  DPF_REQUIRE(in0 && tab && image && (out || stat), DPF_ERR_NULL_PTR, "dpf_pointnet_layer_gemm: null pointer");
  DPF_REQUIRE(loader != LD_BNBWD || in1, DPF_ERR_NULL_PTR, "dpf_pointnet_layer_gemm: the backward loader needs Z");
  DPF_REQUIRE(B > 0 && N > 0 && Mout > 0, DPF_ERR_BAD_ARG, "dpf_pointnet_layer_gemm: bad sizes B=%d N=%d Mout=%d", B, N, Mout);
  DPF_REQUIRE(((uintptr_t)image & 255) == 0 && ((uintptr_t)in0 & 15) == 0 && ((uintptr_t)in1 & 15) == 0 && ((uintptr_t)out & 15) == 0,
              DPF_ERR_ALIGN, "dpf_pointnet_layer_gemm: alignment");
  PlGemmArgs a{in0, in1, tab, (const unsigned char*)image, out, row_off, stat, B, N, Mout};
  cudaStream_t s = (cudaStream_t)stream;
  const bool st = stat != nullptr;
  if (loader == LD_X3 && K == 64) return st ? pl_launch_gemm<64, LD_X3, true>(a, s) : pl_launch_gemm<64, LD_X3, false>(a, s);
  if (loader == LD_AFFINE && K == 128) return st ? pl_launch_gemm<128, LD_AFFINE, true>(a, s) : pl_launch_gemm<128, LD_AFFINE, false>(a, s);
  if (loader == LD_AFFINE && K == 256 && !st) return pl_launch_gemm<256, LD_AFFINE, false>(a, s);
  if (loader == LD_BNBWD && K == 128 && !st) return pl_launch_gemm<128, LD_BNBWD, false>(a, s);
  if (loader == LD_BNBWD && K == 256 && !st) return pl_launch_gemm<256, LD_BNBWD, false>(a, s);
  DPF_REQUIRE(false, DPF_ERR_UNSUPPORTED, "dpf_pointnet_layer_gemm: no kernel for loader %d, K = %d, stats %d", loader, K, (int)st);
  return DPF_ERR_UNSUPPORTED;
}

// bytes of the scratch buffer of dpf_pointnet_layer_wgrad (per-CTA partial sums)
DPF_API int dpf_pointnet_layer_wgrad_scratch_bytes(int MP, int NQ, int B, int N, long long* bytes) {
  DPF_REQUIRE(bytes, DPF_ERR_NULL_PTR, "dpf_pointnet_layer_wgrad_scratch_bytes: null out pointer");
  DPF_REQUIRE(MP > 0 && NQ > 0 && B > 0 && N > 0, DPF_ERR_BAD_ARG, "dpf_pointnet_layer_wgrad_scratch_bytes: bad sizes");
  *bytes = (long long)pl_ctas(B, N, 1) * MP * NQ * (long long)sizeof(float);
  return DPF_OK;
}

// out (MP, NQ) += sum over all points of P Q^T; P = BatchNorm + ReLU backward of (dA, Z) (B,MP,N) with p_tab, or (gram != 0)
// the centred activations relu(sc z + sh) - sub of Z (B,MP,N) with Q = P; Q = layer 0 from x (NQ = 64, q_loader 0) or
// relu(sc z + sh) - sub of q_in (B,NQ,N) (q_loader 1).  out is ACCUMULATED into (zero it first); scratch:
// dpf_pointnet_layer_wgrad_scratch_bytes() bytes, 16-byte aligned.
DPF_API int dpf_pointnet_layer_wgrad(int MP, int NQ, int gram, int q_loader, const float* p_in0, const float* p_in1, const float* p_tab,
                                     const float* q_in, const float* q_tab, int B, int N, void* scratch, float* out, void* stream) {
  DPF_REQUIRE(p_in0 && p_tab && out && scratch && (gram || (p_in1 && q_in && q_tab)), DPF_ERR_NULL_PTR, "dpf_pointnet_layer_wgrad: null pointer");
  DPF_REQUIRE(((uintptr_t)scratch & 15) == 0 && ((uintptr_t)out & 15) == 0, DPF_ERR_ALIGN, "dpf_pointnet_layer_wgrad: scratch / out alignment");
  DPF_REQUIRE(B > 0 && N > 0, DPF_ERR_BAD_ARG, "dpf_pointnet_layer_wgrad: bad sizes B=%d N=%d", B, N);
  DPF_REQUIRE(((uintptr_t)p_in0 & 15) == 0 && ((uintptr_t)p_in1 & 15) == 0 && ((uintptr_t)q_in & 15) == 0, DPF_ERR_ALIGN,
              "dpf_pointnet_layer_wgrad: alignment");
  PlWgradArgs a{p_in0, p_in1, p_tab, q_in, q_tab, (float*)scratch, B, N};
  cudaStream_t s = (cudaStream_t)stream;
  if (gram && MP == 256 && NQ == 256) return pl_launch_wgrad<256, LD_AFFINE, 256, LD_AFFINE, true>(a, out, s);
  if (!gram && MP == 256 && NQ == 128 && q_loader == LD_AFFINE) return pl_launch_wgrad<256, LD_BNBWD, 128, LD_AFFINE, false>(a, out, s);
  if (!gram && MP == 128 && NQ == 64 && q_loader == LD_X3) return pl_launch_wgrad<128, LD_BNBWD, 64, LD_X3, false>(a, out, s);
  DPF_REQUIRE(false, DPF_ERR_UNSUPPORTED, "dpf_pointnet_layer_wgrad: no kernel for MP = %d, NQ = %d, gram %d, q_loader %d", MP, NQ, gram, q_loader);
  return DPF_ERR_UNSUPPORTED;
}

// sums (C, 2) double += {sum dA [y > 0], sum dA [y > 0] (z - mu)} over all points, y = sc z + sh (tab as for loader 2)
DPF_API int dpf_pointnet_bn_bwd_sums(const float* dA, const float* Z, const float* tab, int B, int C, int N, double* sums, void* stream) {
  DPF_REQUIRE(dA && Z && tab && sums, DPF_ERR_NULL_PTR, "dpf_pointnet_bn_bwd_sums: null pointer");
  DPF_REQUIRE(B > 0 && C > 0 && N > 0, DPF_ERR_BAD_ARG, "dpf_pointnet_bn_bwd_sums: bad sizes");
  pl_bn_bwd_sums_kernel<<<B * C, 128, 0, (cudaStream_t)stream>>>(dA, Z, tab, C, N, sums);
  return dpf_check_launch("pl_bn_bwd_sums_kernel");
}

// sums (C, 4) double += {sum dA m, sum dA m x0, sum dA m x1, sum dA m x2}, m = [a . x + c > 0] (tab as for loader 0)
DPF_API int dpf_pointnet_layer0_bwd_sums(const float* dA, const float* x, const float* tab, int B, int C, int N, double* sums, void* stream) {
  DPF_REQUIRE(dA && x && tab && sums, DPF_ERR_NULL_PTR, "dpf_pointnet_layer0_bwd_sums: null pointer");
  DPF_REQUIRE(B > 0 && C > 0 && N > 0, DPF_ERR_BAD_ARG, "dpf_pointnet_layer0_bwd_sums: bad sizes");
  pl_layer0_bwd_sums_kernel<<<B * C, 128, 0, (cudaStream_t)stream>>>(dA, x, tab, C, N, sums);
  return dpf_check_launch("pl_layer0_bwd_sums_kernel");
}

// ---- per-channel finalisation (tiny kernels; everything stays on the device) -------------------------------------
// moments (9,) double += {sum x_i, sum x_i x_j (00 01 02 11 12 22)} over the B * N points of x (B,3,N): zero it first
DPF_API int dpf_pointnet_input_moments(const float* x, int B, int N, double* moments, void* stream) {
  DPF_REQUIRE(x && moments, DPF_ERR_NULL_PTR, "dpf_pointnet_input_moments: null pointer");
  DPF_REQUIRE(B > 0 && N > 0, DPF_ERR_BAD_ARG, "dpf_pointnet_input_moments: bad sizes");
  const long long total = (long long)B * N;
  const int blocks = (int)((total + 256 * 4 - 1) / (256 * 4) < 2 * dpf_num_sms() ? (total + 256 * 4 - 1) / (256 * 4) : 2 * dpf_num_sms());
  pl_input_moments_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, B, N, moments);
  return dpf_check_launch("pl_input_moments_kernel");
}

// layer 0 (W0 (C,3)): tab (C,8) loader-0 table, stats (C,3) {mean, biased var, istd} from the input moments
DPF_API int dpf_pointnet_layer0_finalize(const double* moments, const float* W0, const float* gamma, const float* beta, int C, int B, int N,
                                         float eps, float* tab, float* stats, void* stream) {
  DPF_REQUIRE(moments && W0 && gamma && beta && tab && stats, DPF_ERR_NULL_PTR, "dpf_pointnet_layer0_finalize: null pointer");
  pl_layer0_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(moments, W0, gamma, beta, C, (double)B * N, eps, tab, stats);
  return dpf_check_launch("pl_layer0_finalize_kernel");
}

// stat: (G, C, 3) {count, mean, M2} of dpf_pointnet_layer_gemm (width 3) or (G, C, 2) {mean, M2} with `count` points each of the
// pool kernel (width 2) -> tab (C,8) loader-1 table {sc, sh, 0, 0, 0, mean} (nullable), stats (C,3) {mean, biased var, istd}
DPF_API int dpf_pointnet_stats_finalize(const float* stat, int G, int C, int width, float count, const float* gamma, const float* beta, float eps,
                                        float* tab, float* stats, void* stream) {
  DPF_REQUIRE(stat && stats && (!tab || (gamma && beta)), DPF_ERR_NULL_PTR, "dpf_pointnet_stats_finalize: null pointer");
  DPF_REQUIRE(G > 0 && C > 0 && (width == 2 || width == 3), DPF_ERR_BAD_ARG, "dpf_pointnet_stats_finalize: bad sizes");
  pl_stats_finalize_kernel<<<(C + 3) / 4, 128, 0, (cudaStream_t)stream>>>(stat, G, C, width, count, gamma, beta, eps, tab, stats);
  return dpf_check_launch("pl_stats_finalize_kernel");
}

// sums (C,2) of dpf_pointnet_bn_bwd_sums -> dgamma, dbeta (C,) and the loader-2 table tab_bwd (C,8); tab_fwd: the layer's
// loader-1 table, stats: its {mean, var, istd}
DPF_API int dpf_pointnet_bwd_finalize(const double* sums, const float* tab_fwd, const float* gamma, const float* stats, int C, int B, int N,
                                      float* tab_bwd, float* dgamma, float* dbeta, void* stream) {
  DPF_REQUIRE(sums && tab_fwd && gamma && stats && tab_bwd && dgamma && dbeta, DPF_ERR_NULL_PTR, "dpf_pointnet_bwd_finalize: null pointer");
  pl_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, tab_fwd, gamma, stats, C, (double)B * N, tab_bwd, dgamma, dbeta);
  return dpf_check_launch("pl_bwd_finalize_kernel");
}

// layer 0 backward: sums4 (C,4) of dpf_pointnet_layer0_bwd_sums + the input moments -> dW0 (C,3), dgamma, dbeta (C,)
DPF_API int dpf_pointnet_layer0_bwd_finalize(const double* sums4, const double* moments, const float* W0, const float* gamma, const float* stats,
                                             int C, int B, int N, float* dW0, float* dgamma, float* dbeta, void* stream) {
  DPF_REQUIRE(sums4 && moments && W0 && gamma && stats && dW0 && dgamma && dbeta, DPF_ERR_NULL_PTR, "dpf_pointnet_layer0_bwd_finalize: null pointer");
  pl_layer0_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums4, moments, W0, gamma, stats, C, (double)B * N, dW0, dgamma, dbeta);
  return dpf_check_launch("pl_layer0_bwd_finalize_kernel");
}

// The sparse part of the pooled last layer's backward: Z (B,256,N) the layer's pre-BatchNorm input with its loader-1 table tab
// (256,8), idx (B,C) int64 the selected point per (shape, channel), coef (B,C) the cotangent there, W (C,256).
// T (C,256) += sum_b coef[b][c] a[b][:][n*] (zero it first), dA (B,256,N) += coef[b][c] W[c][:] at point n* (float atomics).
DPF_API int dpf_pointnet_pool_sparse_backward(const float* Z, const float* tab, const long long* idx, const float* coef, const float* W,
                                              int B, int N, int C, float* T, float* dA, void* stream) {
  DPF_REQUIRE(Z && tab && idx && coef && W && T && dA, DPF_ERR_NULL_PTR, "dpf_pointnet_pool_sparse_backward: null pointer");
  DPF_REQUIRE(B > 0 && N > 0 && C > 0 && C % PS_CG == 0 && B <= 65535, DPF_ERR_BAD_ARG, "dpf_pointnet_pool_sparse_backward: bad sizes B=%d N=%d C=%d", B, N, C);
  pl_pool_sparse_bwd_kernel<<<dim3(C / PS_CG, B), 256, 0, (cudaStream_t)stream>>>(Z, tab, idx, coef, W, N, C, T, dA);
  return dpf_check_launch("pl_pool_sparse_bwd_kernel");
}
