// Fused eval-mode decoder: ALL coupling layers of the stack in ONE launch (sampling, mode 'direct',
// and eval-mode 'inverse').  Reference: LocalCondRNVPDecoder.forward in .eval() - decoders.py:54-72
// over CondRealNVPFlow3D.forward, flows.py:95-117 (SURVEY.md section 8 b2 "decoder_fwd_fused_eval").
//
// In eval mode every BatchNorm uses running statistics, so no layer needs a reduction over points:
// a tile of 128 points runs through the whole stack without leaving the SM.  The three coordinates of
// a point stay in registers from layer to layer; only the per-layer list outputs (P, MU, LV - the
// reference API returns all of them) are written to HBM.
//
//   eval_tables_kernel   per layer : folded BN_a {A00, A01, c0}, b2, channel roles   (EvalLayerTab)
//                        per (layer, shape): epilogue table {S, T, W2_0, W2_1} with
//                        S = film_s / sqrt(rv_b + eps), T = film_t - S * rm_b          (BN_b x FiLM fold)
//   decoder_eval_tc_kernel  256 threads = two threads per point (row = tid & 127 = TMEM lane,
//                        part = tid >> 7 = 32 of the 64 conditioner channels).  Per (tile, layer) item:
//                        TMA bulk loads of the layer's W1 image(s) + tables (mbarrier), h1 -> swizzled
//                        bf16 tile -> tcgen05.mma into TMEM, fused epilogue, exchange of the two
//                        partial last-SharedDot sums through shared memory, affine transform.
//                        Tables are double-buffered; the weight buffer is refilled as soon as the
//                        layer's second UMMA chain has completed (overlaps the epilogue).
//                        Shared memory: 46 KB (bf16) -> 4 CTAs/SM (4 x 128 TMEM columns).
#include "coupling.cuh"
#include "umma.cuh"
#include "tc_tiles.cuh"

namespace {

constexpr int NT = 256;

struct EvalLayerTab {              // DPF_EVAL_LTAB_BYTES, 16-byte multiple (bulk-copied)
  float4 PA[2][F / 2][2];          // folded eval BN_a in CHANNEL-PAIR layout (a = channel 2j, b = 2j+1) for the packed fp32x2 math:
                                   //   {A00_a, A00_b, c0_a, c0_b}, {A01_a, A01_b, -, -}
  float b2[2][2];
  int k, w, keep0, keep1, warp0, warp1;
  int pad[6];
};
static_assert(sizeof(EvalLayerTab) == DPF_EVAL_LTAB_BYTES, "EvalLayerTab size");
constexpr uint32_t EPI_BYTES = 2 * F * sizeof(float4);

__global__ void __launch_bounds__(128)
eval_tables_kernel(const float* __restrict__ arena, const float* __restrict__ stats, const LayerMeta* __restrict__ meta,
                   const float* __restrict__ film, unsigned char* __restrict__ ltab_out, float4* __restrict__ epi_out, int B, int G) {
  const int l = blockIdx.x, b = blockIdx.y;
  const int br = threadIdx.x >> 6, c = threadIdx.x & 63;
  const LayerMeta m = meta[l];
  const int k = (int)m.k, w = (int)m.w;
  const BranchLayout lay = branch_layout(k, w, G);
  const float* prm = arena + m.param_off + (size_t)br * lay.size;
  const float* st = stats + m.stat_off + (size_t)br * ST_COUNT * F;
  if (b == 0) {
    EvalLayerTab* lt = reinterpret_cast<EvalLayerTab*>(ltab_out + (size_t)l * sizeof(EvalLayerTab));
    const float w0 = prm[lay.W0 + c * k + 0];
    const float w1 = (k == 2) ? prm[lay.W0 + c * k + 1] : 0.f;
    const float istd = 1.f / sqrtf(st[ST_BNA_RV * F + c] + DPF_BN_EPS);
    const float gi = prm[lay.bnA_w + c] * istd;
    float* pa = reinterpret_cast<float*>(&lt->PA[br][c >> 1][0]);
    const int ln = c & 1;
    pa[ln] = gi * w0;
    pa[2 + ln] = prm[lay.bnA_b + c] - gi * st[ST_BNA_RM * F + c];
    pa[4 + ln] = gi * w1;
    pa[6 + ln] = 0.f;
    if (c < 2) lt->b2[br][c] = (c < w) ? prm[lay.b2 + c] : 0.f;
    if (threadIdx.x == 0) {
      lt->k = k; lt->w = w;
      lt->keep0 = (int)m.keep0; lt->keep1 = (int)m.keep1; lt->warp0 = (int)m.warp0; lt->warp1 = (int)m.warp1;
    }
  }
  const float ib = 1.f / sqrtf(st[ST_BNB_RV * F + c] + DPF_BN_EPS);
  const float mb = st[ST_BNB_RM * F + c];
  const float sc = film[(((size_t)l * 4 + br * 2 + 0) * B + b) * F + c];
  const float sh = film[(((size_t)l * 4 + br * 2 + 1) * B + b) * F + c];
  const float S = sc * ib;
  // channel-pair layout: [l][b][br][F/2] x {S_a, S_b, T_a, T_b}, {W20_a, W20_b, W21_a, W21_b}
  float* pe = reinterpret_cast<float*>(epi_out + ((((size_t)l * B + b) * 2 + br) * (F / 2) + (c >> 1)) * 2);
  const int ln = c & 1;
  pe[ln] = S;
  pe[2 + ln] = fmaf(-S, mb, sh);
  pe[4 + ln] = prm[lay.W2 + c];
  pe[6 + ln] = (w == 2) ? prm[lay.W2 + F + c] : 0.f;
}

struct EvalArgs {
  const float* p;                  // (B,3,N) input of the first processed layer
  float* P; float* MU; float* LV;  // (L,B,3,N) list outputs, indexed by layer (MU nullable: not written)
  float* SLV;                      // nullable (B,3,N): sum over all layers of logvar, accumulated in registers
  const unsigned char* ltab;       // [L] EvalLayerTab
  const float4* epi;               // [L][B][2][F]
  const unsigned short* wimg;      // [L][2][N_IMG] weight images (pack_w1_kernel)
  int L, B, N, tiles_per_b, n_tiles;
  float eps;
};

template <bool SPLIT>
struct EvalSmem {
  unsigned char W[(SPLIT ? 4 : 2) * IMG_W];     // [br][W1 hi (, W1 lo)] of the layer in flight
  unsigned char H[(SPLIT ? 2 : 1) * IMG_H];     // h1 tile (hi (, lo)) of the branch in flight
  struct Tab { EvalLayerTab lt; float4 epi[2][F / 2][2]; } tab[2];
  float4 obuf[2][DPF_TILE];                     // partial last-SharedDot sums of the two parts
  uint64_t bar_mma, bar_w;
  uint32_t tmem_base;
};

// 32 channels (16 pairs) of one branch: h3 = relu(S * acc + T), partial sums of the last SharedDot; packed fp32x2 math
// (FFMA2: two IEEE fp32 lanes per instruction), even / odd channels accumulate in the two lanes of o0_2 / o1_2
template <bool W2ND>
__device__ __forceinline__ void eval_epilogue32(const float4 (*__restrict__ epi)[2], uint32_t taddr, f32x2& o0_2, f32x2& o1_2) {
  uint32_t ra[16], rb[16];
  umma::tmem_ld16_issue(taddr, ra);
  umma::tmem_ld16_issue(taddr + 16, rb);
  umma::tmem_ld_wait16(ra);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 E0 = epi[j][0], E1 = epi[j][1];
    const f32x2 a2 = f2_fma(f2_pack(E0.x, E0.y), f2_pack(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1])), f2_pack(E0.z, E0.w));
    float ha, hb;
    f2_unpack(a2, ha, hb);
    const f32x2 h2 = f2_pack(fmaxf(ha, 0.f), fmaxf(hb, 0.f));
    o0_2 = f2_fma(f2_pack(E1.x, E1.y), h2, o0_2);
    if (W2ND) o1_2 = f2_fma(f2_pack(E1.z, E1.w), h2, o1_2);
  }
  umma::tmem_ld_wait16(rb);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 E0 = epi[8 + j][0], E1 = epi[8 + j][1];
    const f32x2 a2 = f2_fma(f2_pack(E0.x, E0.y), f2_pack(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1])), f2_pack(E0.z, E0.w));
    float ha, hb;
    f2_unpack(a2, ha, hb);
    const f32x2 h2 = f2_pack(fmaxf(ha, 0.f), fmaxf(hb, 0.f));
    o0_2 = f2_fma(f2_pack(E1.x, E1.y), h2, o0_2);
    if (W2ND) o1_2 = f2_fma(f2_pack(E1.z, E1.w), h2, o1_2);
  }
}

template <int MODE, bool SPLIT>
__global__ void __launch_bounds__(NT, SPLIT ? 2 : 4)
decoder_eval_tc_kernel(const EvalArgs a) {
  extern __shared__ unsigned char smraw[];
  using Smem = EvalSmem<SPLIT>;
  Smem& s = *reinterpret_cast<Smem*>(smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, part = tid >> 7, quarter = warp & 3;
  constexpr uint32_t W_BR = (SPLIT ? 2 : 1) * IMG_W;                 // staged bytes per branch
  constexpr uint32_t ITEM_TX = 2 * W_BR + (uint32_t)sizeof(EvalLayerTab) + EPI_BYTES;

  if (tid == 0) {
    umma::mbar_init(&s.bar_mma, 1);
    umma::mbar_init(&s.bar_w, 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(&s.tmem_base, 128);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);

  const int per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(a.n_tiles, t0 + per);

  // loads of one (tile, step) item: weight image(s) of both branches + layer table + epilogue table
  auto issue_item_loads = [&](int tile, int q, uint32_t stage) {
    const int l = MODE == 0 ? q : a.L - 1 - q;
    const int b = tile / a.tiles_per_b;
    const unsigned char* wl = reinterpret_cast<const unsigned char*>(a.wimg) + (size_t)l * 2 * N_IMG * IMG_W;
    umma::mbar_expect_tx(&s.bar_w, ITEM_TX);
    umma::bulk_g2s(s.W, wl, W_BR, &s.bar_w);
    umma::bulk_g2s(s.W + W_BR, wl + N_IMG * IMG_W, W_BR, &s.bar_w);
    umma::bulk_g2s(&s.tab[stage].lt, a.ltab + (size_t)l * sizeof(EvalLayerTab), (uint32_t)sizeof(EvalLayerTab), &s.bar_w);
    umma::bulk_g2s(&s.tab[stage].epi[0][0], a.epi + ((size_t)l * a.B + b) * 2 * F, EPI_BYTES, &s.bar_w);
  };
  if (tid == 0 && t0 < t1) issue_item_loads(t0, 0, 0);

  const float sig1 = sqrtf(a.eps + 1.0f);
  uint32_t it = 0, mma_phase = 0;
  for (int tile = t0; tile < t1; ++tile) {
    const int b = tile / a.tiles_per_b;
    const int n = (tile - b * a.tiles_per_b) * DPF_TILE + row;
    const bool valid = n < a.N;
    const size_t pbase = (size_t)b * 3 * a.N + n;
    float xin[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) xin[ch] = valid ? a.p[pbase + (size_t)ch * a.N] : 0.f;
    float slv[3] = {0.f, 0.f, 0.f};

    for (int q = 0; q < a.L; ++q, ++it) {
      const int l = MODE == 0 ? q : a.L - 1 - q;
      const uint32_t stage = it & 1u;
      umma::mbar_wait(&s.bar_w, it & 1u);
      const typename Smem::Tab& tb = s.tab[stage];
      const int k = tb.lt.k;
      const float xk0 = pick3(xin, tb.lt.keep0);
      const float xk1 = (k == 2) ? pick3(xin, tb.lt.keep1) : 0.f;
      const f32x2 xk0_2 = f2_pack(xk0, xk0), xk1_2 = f2_pack(xk1, xk1);

#pragma unroll
      for (int br = 0; br < 2; ++br) {
        // h1 = relu(A0 . x_keep + c0) of this part's 32 channels -> swizzled bf16 tile; packed fp32x2 math over channel pairs
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int ch8 = part * 4 + qq;
          uint32_t w[4], wl[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 P0 = tb.lt.PA[br][ch8 * 4 + i][0];
            f32x2 v2 = f2_fma(f2_pack(P0.x, P0.y), xk0_2, f2_pack(P0.z, P0.w));
            if (k == 2) {                   // uniform per layer
              const float4 P1 = tb.lt.PA[br][ch8 * 4 + i][1];
              v2 = f2_fma(f2_pack(P1.x, P1.y), xk1_2, v2);
            }
            float va, vb;
            f2_unpack(v2, va, vb);
            if (SPLIT) {
              va = fmaxf(va, 0.f);
              vb = fmaxf(vb, 0.f);
              w[i] = umma::pack_bf16(va, vb);
              wl[i] = umma::pack_bf16(va - __uint_as_float(w[i] << 16), vb - __uint_as_float(w[i] & 0xffff0000u));
            } else {
              w[i] = umma::pack_bf16_relu(va, vb);     // ReLU inside the conversion
            }
          }
          const uint32_t off = umma::sw128_offset(row, ch8);
          *reinterpret_cast<uint4*>(s.H + off) = make_uint4(w[0], w[1], w[2], w[3]);
          if (SPLIT) *reinterpret_cast<uint4*>(s.H + IMG_H + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
        umma::fence_async_smem();
        umma::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
          umma::fence_after_sync();
          issue_gemm1<SPLIT>(tmem + br * F, s.H, s.H + IMG_H, s.W + br * W_BR, s.W + br * W_BR + IMG_W);
          umma::mma_commit(&s.bar_mma);
        }
        umma::mbar_wait(&s.bar_mma, mma_phase);
        mma_phase ^= 1u;
        umma::fence_after_sync();
      }
      // both chains are complete: the weight buffer and the other table stage are free -> prefetch the next item
      if (tid == 0) {
        if (q + 1 < a.L) issue_item_loads(tile, q + 1, stage ^ 1u);
        else if (tile + 1 < t1) issue_item_loads(tile + 1, 0, stage ^ 1u);
      }

      const int wn = tb.lt.w;
      float o[2][2];
#pragma unroll
      for (int br = 0; br < 2; ++br) {
        f32x2 o0_2 = f2_pack(part == 0 ? tb.lt.b2[br][0] : 0.f, 0.f), o1_2 = f2_pack(part == 0 ? tb.lt.b2[br][1] : 0.f, 0.f);
        if (wn == 2) eval_epilogue32<true>(&tb.epi[br][part * 16], lane_addr + br * F + part * 32, o0_2, o1_2);
        else eval_epilogue32<false>(&tb.epi[br][part * 16], lane_addr + br * F + part * 32, o0_2, o1_2);
        float lo, hi;
        f2_unpack(o0_2, lo, hi);
        o[br][0] = lo + hi;
        f2_unpack(o1_2, lo, hi);
        o[br][1] = lo + hi;
      }
      s.obuf[part][row] = make_float4(o[0][0], o[0][1], o[1][0], o[1][1]);
      umma::fence_before_sync();
      __syncthreads();
      const float4 u0 = s.obuf[0][row], u1 = s.obuf[1][row];
      const float omu[2] = {u0.x + u1.x, u0.y + u1.y};
      const float olv[2] = {u0.z + u1.z, u0.w + u1.w};

      // affine transform (both threads of the point keep the new coordinates; part 0 writes P and MU, part 1 LV)
      const int warp0 = tb.lt.warp0, warp1 = tb.lt.warp1;
      float yv[3], muv[3] = {0.f, 0.f, 0.f}, lvv[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) yv[ch] = (MODE == 0) ? sig1 * xin[ch] : xin[ch] / sig1;
#pragma unroll
      for (int wi = 0; wi < 2; ++wi) {
        if (wi < wn) {
          const int ch = wi == 0 ? warp0 : warp1;
          const float lg = softsign(olv[wi]);
          const float sig = sqrtf(a.eps + expf(lg));
          const float m = omu[wi];
          const float xv = pick3(xin, ch);
          const float r = (MODE == 0) ? fmaf(sig, xv, m) : (xv - m) / sig;
#pragma unroll
          for (int c3 = 0; c3 < 3; ++c3)
            if (c3 == ch) { yv[c3] = r; muv[c3] = m; lvv[c3] = lg; }
        }
      }
      if (valid) {
        const size_t obase = (size_t)l * a.B * 3 * a.N + pbase;
        if (part == 0) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            a.P[obase + (size_t)ch * a.N] = yv[ch];
            if (a.MU) a.MU[obase + (size_t)ch * a.N] = muv[ch];
          }
        } else {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) a.LV[obase + (size_t)ch * a.N] = lvv[ch];
        }
      }
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        xin[ch] = yv[ch];
        slv[ch] += lvv[ch];
      }
    }
    if (a.SLV && valid && part == 1) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) a.SLV[pbase + (size_t)ch * a.N] = slv[ch];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

template <int MODE, bool SPLIT>
int launch_eval_t(const EvalArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(EvalSmem<SPLIT>) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(decoder_eval_tc_kernel<MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const int slots = dpf_num_sms() * (SPLIT ? 2 : 4);
  const int grid = a.n_tiles < slots ? a.n_tiles : slots;
  decoder_eval_tc_kernel<MODE, SPLIT><<<grid, NT, smem, st>>>(a);
  return dpf_check_launch("decoder_eval_tc_kernel");
}

}  // namespace

// Whole eval-mode stack in two launches (tables + fused decoder).  film = workspace FiLM table
// [L][4][B][F] (film_forward_kernel), wimg = packed W1 images (pack_w1_kernel).
int launch_decoder_eval_tc(const float* arena, const float* stats, const LayerMeta* meta_dev, const float* film,
                           const unsigned short* wimg, unsigned char* ltab, float* epi, const float* p, float* P, float* MU,
                           float* LV, float* SLV, int L, int G, int B, int N, int mode, int split, float eps, cudaStream_t s) {
  eval_tables_kernel<<<dim3(L, B), 128, 0, s>>>(arena, stats, meta_dev, film, ltab, reinterpret_cast<float4*>(epi), B, G);
  int rc = dpf_check_launch("eval_tables_kernel");
  if (rc) return rc;
  EvalArgs a{};
  a.p = p; a.P = P; a.MU = MU; a.LV = LV; a.SLV = SLV;
  a.ltab = ltab; a.epi = reinterpret_cast<const float4*>(epi); a.wimg = wimg;
  a.L = L; a.B = B; a.N = N;
  a.tiles_per_b = (N + DPF_TILE - 1) / DPF_TILE;
  a.n_tiles = a.tiles_per_b * B;
  a.eps = eps;
  if (mode == 0) return split ? launch_eval_t<0, true>(a, s) : launch_eval_t<0, false>(a, s);
  return split ? launch_eval_t<1, true>(a, s) : launch_eval_t<1, false>(a, s);
}
