// Tensor-core (tcgen05 / TMEM) kernels of the conditional coupling stack - the BF16 path.
//
// Tile = 128 consecutive points of one shape = the 128 TMEM lanes of one CTA (thread t owns point
// t and TMEM lane t).  Per tile and branch the 64x64 SharedDot is a UMMA chain
//     D[128 x 64] (TMEM, fp32) = H1[128 x 64] (smem, bf16, K-major SW128) x W1^T (smem, bf16)
// whose A operand the CUDA cores produce in place (first SharedDot + folded BN_a + ReLU -> bf16 ->
// swizzled st.shared), and whose accumulator row each thread reads back with tcgen05.ld for the
// fused epilogue (BN_b x FiLM fold, ReLU, last SharedDot, softsign, exp, sqrt, affine transform,
// log-det output) - no activation ever goes to HBM.  Weight images (pre-swizzled bf16, packed per
// step by pack_w1_kernel) are staged with TMA bulk copies (cp.async.bulk + mbarrier).
//
// SPLIT (precision "bf16x3"): both operands are carried as bf16 hi + bf16 lo and the product is
// accumulated from three chains hi*hi + lo*hi + hi*lo (fp32-class accuracy, ~2^-17 per operand).
// Train-mode BatchNorm divides by small batch deviations and amplifies single-bf16 rounding beyond
// the 2e-2 budget (measured: tests/test_decoder_gpu.py), so training defaults to SPLIT; plain
// "bf16" stays available (sampling / eval: ~1e-3).
//
// Backward pass 2 chains three UMMAs per tile: the forward recompute, dgrad
//     DH1[128 x 64] = DH2[128 x 64] x W1          (K-major operands, W1^T image)
// and wgrad accumulated in TMEM across all tiles of the CTA
//     DW1[c, j] += sum_p DH2[p, c] * H1[p, j]      (the SAME smem tiles read as MN-major operands).
// Per-channel reductions over points (BN statistics, FiLM / BN / W2 gradients) go through a
// shared-memory transpose (colreduce) instead of warp shuffles.
#include "coupling.cuh"
#include "umma.cuh"
#include "tc_tiles.cuh"

// Phase timing for development (build with -DDPF_STAMPS, tools/stamp_probe.py): thread 0 of every CTA
// records clock64() at phase boundaries of the last launch of each kernel class.  Compiled out of
// the product library.
#ifdef DPF_STAMPS
__device__ unsigned long long* g_stamps = nullptr;      // [3 classes][512 CTAs][16]
#define DPF_STAMP(cls, i)                                                                                     \
  do {                                                                                                        \
    if (threadIdx.x == 0 && g_stamps && blockIdx.x < 512) g_stamps[((cls) * 512 + blockIdx.x) * 16 + (i)] = clock64(); \
  } while (0)
#define DPF_STAMP_NS(cls, i)                                                                                  \
  do {                                                                                                        \
    if (threadIdx.x == 0 && g_stamps && blockIdx.x < 512) {                                                   \
      unsigned long long t_;                                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                  \
      g_stamps[((cls) * 512 + blockIdx.x) * 16 + (i)] = t_;                                                   \
    }                                                                                                         \
  } while (0)
#else
#define DPF_STAMP(cls, i) do { } while (0)
#define DPF_STAMP_NS(cls, i) do { } while (0)
#endif
extern int g_dpf_p2_two_tiles;   // decoder.cu: dpf_set_option(3, v)
#ifdef DPF_EXP_NOATOMICS
#define DPF_GATOMIC(stmt) do { } while (0)
#else
#define DPF_GATOMIC(stmt) stmt
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// Weight packing: fp32 W1 (arena) -> bf16 128B-swizzled images per (layer, branch)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_w1_kernel(const float* __restrict__ arena, const LayerMeta* __restrict__ meta, int G, unsigned short* __restrict__ out) {
  const int l = blockIdx.x >> 1, br = blockIdx.x & 1;
  const LayerMeta m = meta[l];
  const BranchLayout lay = branch_layout((int)m.k, (int)m.w, G);
  const float* W1 = arena + m.param_off + (size_t)br * lay.size + lay.W1;
  unsigned char* img = reinterpret_cast<unsigned char*>(out) + (size_t)(l * 2 + br) * N_IMG * IMG_W;
  for (int e = threadIdx.x; e < N_IMG * F * 8; e += 256) {
    const int t = e / (F * 8), r = (e / 8) % F, q = e & 7;   // image t (0: W1 hi, 1: W1 lo, 2: W1^T hi), row r, chunk q
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c0 = q * 8 + 2 * i;
      float lo = t == 2 ? W1[c0 * F + r] : W1[r * F + c0];
      float hi = t == 2 ? W1[(c0 + 1) * F + r] : W1[r * F + c0 + 1];
      if (t == 1) {   // residual of the bf16 rounding: W1 = hi + lo to ~2^-17
        lo -= __bfloat162float(__float2bfloat16_rn(lo));
        hi -= __bfloat162float(__float2bfloat16_rn(hi));
      }
      w[i] = umma::pack_bf16(lo, hi);
    }
    *reinterpret_cast<uint4*>(img + t * IMG_W + umma::sw128_offset(r, q)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// Shared building blocks
// ---------------------------------------------------------------------------------------------
struct TcCommon {                       // small per-CTA tables (after the 1024-aligned image area)
  float4 A0[2][F];                      // folded BN_a: {A00, A01, c0, -}
  float4 epi[2][F];                     // per tile: {S, T, W2_0, W2_1},  a = S*acc + T
  // the same two tables in CHANNEL-PAIR layout (a = channel 2j, b = 2j+1) for the packed fp32x2 math (FFMA2):
  float4 A0p[2][F / 2][2];              //   {A00_a, A00_b, c0_a, c0_b}, {A01_a, A01_b, -, -}
  float4 epip[2][F / 2][2];             //   {S_a, S_b, T_a, T_b}, {W20_a, W20_b, W21_a, W21_b}
  float mb[2][F], ib[2][F], sraw[2][F];
  float W2[2][2][F];
  float b2[2][2];
  double pend[20];
  uint64_t bar_mma, bar_load, bar_aux;
  uint32_t tmem_base;
};

// (all table helpers index by row = threadIdx.x & 127: in the 256-thread kernels both threads of a
// point compute identical entries; side effects are restricted to threadIdx.x < 128)
__device__ __forceinline__ void tc_prologue_tables(const CouplingArgs& a, const BranchLayout& lay, TcCommon& s, bool writer, bool need_bnb) {
  const int tid = threadIdx.x & 127, br = tid >> 6, c = tid & 63;
  writer = writer && threadIdx.x < 128;
  float A00, A01, c0;
  fold_bn_a(a, lay, br, c, writer, A00, A01, c0, nullptr, nullptr);
  s.A0[br][c] = make_float4(A00, A01, c0, 0.f);
  {
    float* pa = reinterpret_cast<float*>(&s.A0p[br][c >> 1][0]);
    const int ln = c & 1;
    pa[ln] = A00; pa[2 + ln] = c0; pa[4 + ln] = A01; pa[6 + ln] = 0.f;
  }
  if (need_bnb) {
    float mean, istd;
    bn_b_stats(a, br, c, writer, mean, istd);
    s.mb[br][c] = mean;
    s.ib[br][c] = istd;
    const float* prm = a.prm + (size_t)br * lay.size;
    s.W2[br][0][c] = prm[lay.W2 + c];
    s.W2[br][1][c] = (a.w == 2) ? prm[lay.W2 + F + c] : 0.f;
    if (c < 2) s.b2[br][c] = (c < a.w) ? prm[lay.b2 + c] : 0.f;
  }
}

// backward kernels: the same tables loaded from the per-pass table (bwd_tables_kernel)
__device__ __forceinline__ void tc_prologue_tables_ltab(const float* __restrict__ ltab, TcCommon& s) {
  const int tid = threadIdx.x & 127, br = tid >> 6, c = tid & 63;
  const float4 t0 = reinterpret_cast<const float4*>(ltab)[tid * (DPF_LTAB_ROW / 4) + 0];
  const float4 t1 = reinterpret_cast<const float4*>(ltab)[tid * (DPF_LTAB_ROW / 4) + 1];
  s.A0[br][c] = make_float4(t0.x, t0.y, t0.z, 0.f);
  {
    float* pa = reinterpret_cast<float*>(&s.A0p[br][c >> 1][0]);
    const int ln = c & 1;
    pa[ln] = t0.x; pa[2 + ln] = t0.z; pa[4 + ln] = t0.y; pa[6 + ln] = 0.f;
  }
  s.mb[br][c] = t0.w;
  s.ib[br][c] = t1.x;
  s.W2[br][0][c] = t1.y;
  s.W2[br][1][c] = t1.z;
  if (c < 2) s.b2[br][c] = t1.w;
}

__device__ __forceinline__ void tc_tile_film(const CouplingArgs& a, TcCommon& s, int b) {
  const int tid = threadIdx.x & 127, br = tid >> 6, c = tid & 63;
  const float sc = a.film[((size_t)(br * 2 + 0) * a.B + b) * F + c];
  const float sh = a.film[((size_t)(br * 2 + 1) * a.B + b) * F + c];
  const float S = sc * s.ib[br][c];
  s.sraw[br][c] = sc;
  const float T = fmaf(-S, s.mb[br][c], sh);
  s.epi[br][c] = make_float4(S, T, s.W2[br][0][c], s.W2[br][1][c]);
  float* pe = reinterpret_cast<float*>(&s.epip[br][c >> 1][0]);
  const int ln = c & 1;
  pe[ln] = S; pe[2 + ln] = T; pe[4 + ln] = s.W2[br][0][c]; pe[6 + ln] = s.W2[br][1][c];
}

// h1 = relu(A0 . x_keep + c0) of the 8 channels of 16-byte chunk q of one branch (packed fp32x2 math over the channel
// pairs of the pair-layout table) -> bf16 hi (and lo residual) words
template <int K, bool SPLIT>
__device__ __forceinline__ void h1_chunk8(const float4 (*__restrict__ A0p)[2], int q, f32x2 xk0_2, f32x2 xk1_2, uint32_t w[4], uint32_t wl[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 P0 = A0p[q * 4 + i][0];
    f32x2 v2 = f2_fma(f2_pack(P0.x, P0.y), xk0_2, f2_pack(P0.z, P0.w));
    if (K == 2) {
      const float4 P1 = A0p[q * 4 + i][1];
      v2 = f2_fma(f2_pack(P1.x, P1.y), xk1_2, v2);
    }
    float va, vb;
    f2_unpack(v2, va, vb);
    va = fmaxf(va, 0.f);
    vb = fmaxf(vb, 0.f);
    w[i] = umma::pack_bf16(va, vb);
    if (SPLIT) wl[i] = umma::pack_bf16(va - __uint_as_float(w[i] << 16), vb - __uint_as_float(w[i] & 0xffff0000u));
  }
}

// CTA setup: barriers and TMEM allocation (the weight images are TMA-bulk-loaded by the caller).
__device__ __forceinline__ uint32_t tc_setup(TcCommon& s, uint32_t tmem_cols) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    umma::mbar_init(&s.bar_mma, 1);
    umma::mbar_init(&s.bar_load, 1);
    umma::mbar_init(&s.bar_aux, 1);
    umma::mbar_fence_init();
  }
  if ((tid >> 5) == 0) umma::tmem_alloc(&s.tmem_base, tmem_cols);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  return s.tmem_base;
}

// TMA bulk loads of this layer's weight images.  ALL = {hi, lo, T} of both branches in one copy
// (backward pass 2); otherwise only {hi, lo} of each branch, packed [br][hi, lo] (32 KB).
template <bool ALL>
__device__ __forceinline__ void tc_load_weights(TcCommon& s, unsigned char* W, const unsigned short* wimg) {
  if (threadIdx.x == 0) {
    if (ALL) {
      umma::mbar_expect_tx(&s.bar_load, 2 * N_IMG * IMG_W);
      umma::bulk_g2s(W, wimg, 2 * N_IMG * IMG_W, &s.bar_load);
    } else {
      umma::mbar_expect_tx(&s.bar_load, 4 * IMG_W);
      umma::bulk_g2s(W, wimg, 2 * IMG_W, &s.bar_load);
      umma::bulk_g2s(W + 2 * IMG_W, reinterpret_cast<const unsigned char*>(wimg) + N_IMG * IMG_W, 2 * IMG_W, &s.bar_load);
    }
  }
}
// image t (0 hi, 1 lo, 2 T) of branch br inside the staged block
template <bool ALL>
__device__ __forceinline__ const unsigned char* wimg_at(const unsigned char* W, int br, int t) { return W + (br * (ALL ? N_IMG : 2) + t) * IMG_W; }

__device__ __forceinline__ void tile_range(int n_tiles, int& t0, int& t1) {   // contiguous tiles per CTA
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  t0 = blockIdx.x * per;
  t1 = min(n_tiles, t0 + per);
}

// Tiles whose SharedDot accumulators one CTA of the merged train-mode forward keeps resident in TMEM
// across the grid barrier (RES * 128 columns; two CTAs per SM use all 512).
constexpr int RES = 2;

// Software grid barrier.  Every CTA of the launch must be resident for it to complete: the launcher sizes the grid to
// what fits (2 CTAs per SM by shared memory, registers and TMEM) and VERIFIES that co-residency once per process with
// a probe launch of the same footprint (tc_verify_coresidency).  Should a CTA nevertheless wait longer than ~2 s
// (another tenant on the SMs), the barrier does not continue with partial sums: it raises the host-visible failure
// flag and traps, so the pass ends with a CUDA error and every later decoder call reports the timeout.
__device__ int* g_barrier_fail_dev = nullptr;      // device alias of a pinned, mapped host int (tc_barrier_setup)

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v < expected) {
        __nanosleep(64);
        if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz
          atomicExch(counter + 1, 1u);
          if (g_barrier_fail_dev) {
            *reinterpret_cast<volatile int*>(g_barrier_fail_dev) = 1;
            __threadfence_system();
          }
          __trap();
        }
      }
    } while (v < expected);
  }
  __syncthreads();
}

// Probe with the resource footprint of the cooperative kernels (threads, dynamic shared memory, TMEM columns,
// <= 128 registers): all CTAs meet at a barrier with a short timeout; counter[1] != 0 afterwards = they were NOT all
// resident at the same time.
__global__ void __launch_bounds__(576)
coresidency_probe_kernel(unsigned int* counter, int tmem_cols) {
  extern __shared__ unsigned char smraw[];
  __shared__ uint32_t tmem_base;
  if (threadIdx.x == 0) smraw[0] = 0;
  if ((threadIdx.x >> 5) == 0) umma::tmem_alloc(&tmem_base, tmem_cols);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v < gridDim.x) {
        __nanosleep(64);
        if (clock64() - t0 > 100000000LL) {   // ~50 ms
          atomicExch(counter + 1, 1u);
          break;
        }
      }
    } while (v < gridDim.x);
  }
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) umma::tmem_dealloc(tmem_base, tmem_cols);
}

// =============================================================================================
// Backward helpers (same math as coupling_bwd.cu)
// =============================================================================================
__device__ Pending tc_compute_pending(const BwdArgs& a, bool writer, double* red) {
  Pending P{0.f, 0.f, 0.f, 0.f, 0.f};
  if (!a.has_pending) return P;
  const int tid = threadIdx.x & 127, br = tid >> 6, c = tid & 63;
  writer = writer && threadIdx.x < 128;
  const BranchLayout nlay = branch_layout(a.nk, a.nw, a.f.G);
  const double M = (double)a.f.B * (double)a.f.N;
  BnA bn;
  if (a.n_ltab) {   // bn_a_of() of that layer, precomputed by bwd_tables_kernel
    const float4 t2 = reinterpret_cast<const float4*>(a.n_ltab)[tid * (DPF_LTAB_ROW / 4) + 2];
    const float4 t3 = reinterpret_cast<const float4*>(a.n_ltab)[tid * (DPF_LTAB_ROW / 4) + 3];
    bn.w0 = t2.x; bn.w1 = t2.y; bn.gamma = t2.z; bn.mean = t2.w; bn.istd = t3.x;
  } else {
    bn = bn_a_of(a.nprm, a.nstat, nlay, a.n_mom, M, a.nk, a.nkeep0, a.nkeep1, a.f.training, br, c);
  }
  const double dbeta = a.n_bna_sums[(br * F + c) * 4 + 0];
  const double E0 = a.n_bna_sums[(br * F + c) * 4 + 1];
  const double E1 = a.n_bna_sums[(br * F + c) * 4 + 2];
  const double istd = bn.istd, mean = bn.mean, w0 = bn.w0, w1 = bn.w1, gam = bn.gamma;
  const double dgamma = istd * (w0 * E0 + w1 * E1 - mean * dbeta);
  const double rM = 1.0 / M;
  const double n1 = a.f.training ? gam * dbeta * rM : 0.0;
  const double n2 = a.f.training ? gam * dgamma * rM : 0.0;
  const double i2n2 = istd * istd * n2;
  double v[5] = {w0 * istd * n1 - i2n2 * mean * w0, w1 * istd * n1 - i2n2 * mean * w1, i2n2 * w0 * w0, i2n2 * w0 * w1, i2n2 * w1 * w1};
  if (writer) {
    float* d = a.ndprm + (size_t)br * nlay.size;
    d[nlay.bnA_b + c] = (float)dbeta;
    d[nlay.bnA_w + c] = (float)dgamma;
    double S1[2] = {0.0, 0.0}, S2[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    if (a.f.training) {
      S1[0] = a.n_mom[a.nkeep0];
      S2[0][0] = a.n_mom[mom2_index(a.nkeep0, a.nkeep0)];
      if (a.nk == 2) {
        S1[1] = a.n_mom[a.nkeep1];
        S2[0][1] = S2[1][0] = a.n_mom[mom2_index(a.nkeep0, a.nkeep1)];
        S2[1][1] = a.n_mom[mom2_index(a.nkeep1, a.nkeep1)];
      }
    }
    const double E[2] = {E0, E1};
    for (int j = 0; j < a.nk; ++j) {
      const double sx = istd * (w0 * S2[0][j] + w1 * S2[1][j] - mean * S1[j]);
      d[nlay.W0 + c * a.nk + j] = (float)(istd * (gam * E[j] - n1 * S1[j] - n2 * sx));
    }
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const double sm = warp_sum_d(v[i]);
    if (lane == 0) red[i * 4 + warp] = sm;
  }
  __syncthreads();
  P.c0 = (float)(red[0] + red[1] + red[2] + red[3]);
  P.c1 = (float)(red[4] + red[5] + red[6] + red[7]);
  P.q00 = (float)(red[8] + red[9] + red[10] + red[11]);
  P.q01 = (float)(red[12] + red[13] + red[14] + red[15]);
  P.q11 = (float)(red[16] + red[17] + red[18] + red[19]);
  __syncthreads();
  return P;
}

struct TcPoint { float x[3], dy[3], sig[2], do_mu[2], do_lv[2]; };
struct TcRaw { float x[3], y[3], lv[3], dy[3], dmu[3], dlv[3]; };

// Global loads of one point's backward inputs.  (Issuing them one tile ahead, or prefetching them
// into L1, was measured and does not shorten the step: profiles/README.md.)
__device__ __forceinline__ TcRaw tc_load_raw(const BwdArgs& a, int b, int n, bool valid) {
  TcRaw r;
  const int N = a.f.N;
  const size_t base = (size_t)b * 3 * N + n;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const size_t o = base + (size_t)ch * N;
    r.x[ch] = valid ? a.f.x[o] : 0.f;
    r.y[ch] = valid ? a.yv[o] : 0.f;
    r.lv[ch] = valid ? a.lvv[o] : 0.f;
    float d = 0.f;
    if (valid && a.dy_chain) d += a.dy_chain[o];
    if (valid && a.dP) d += a.dP[o];
    r.dy[ch] = d;
    r.dmu[ch] = (valid && a.dMU) ? a.dMU[o] : 0.f;
    r.dlv[ch] = (valid && a.dLV) ? a.dLV[o] : 0.f;
  }
  return r;
}
template <int MODE>
__device__ __forceinline__ TcPoint tc_finish_point(const BwdArgs& a, const Pending& P, const TcRaw& r, bool valid) {
  TcPoint g;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    g.x[ch] = r.x[ch];
    g.dy[ch] = r.dy[ch];
  }
  if (a.has_pending && valid) {
    const float y0 = pick3(r.y, a.nkeep0);
    const float y1 = a.nk == 2 ? pick3(r.y, a.nkeep1) : 0.f;
    const float corr0 = P.c0 + P.q00 * y0 + P.q01 * y1;
    const float corr1 = P.c1 + P.q01 * y0 + P.q11 * y1;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      if (ch == a.nkeep0) g.dy[ch] -= corr0;
      if (a.nk == 2 && ch == a.nkeep1) g.dy[ch] -= corr1;
    }
  }
#pragma unroll
  for (int wi = 0; wi < 2; ++wi) {
    g.sig[wi] = 1.f; g.do_mu[wi] = 0.f; g.do_lv[wi] = 0.f;
    if (wi < a.f.w) {
      const int ch = wi == 0 ? a.f.warp0 : a.f.warp1;
      const float l = pick3(r.lv, ch), dyv = pick3(g.dy, ch);
      const float e = expf(l);
      const float s2 = a.f.eps + e;
      const float sg = sqrtf(s2);
      g.sig[wi] = sg;
      float dm, dl;
      if (MODE == 1) {
        dm = -dyv / sg;
        dl = -dyv * pick3(r.y, ch) * e / (2.f * s2);
      } else {
        dm = dyv;
        dl = dyv * pick3(g.x, ch) * e / (2.f * sg);
      }
      dm += pick3(r.dmu, ch);
      dl += pick3(r.dlv, ch);
      const float om = 1.f - fabsf(l);
      g.do_mu[wi] = valid ? dm : 0.f;
      g.do_lv[wi] = valid ? dl * om * om : 0.f;
    }
  }
  return g;
}

// =============================================================================================
// Backward pass 2, two threads per point (256 threads per CTA): thread (row = tid & 127, part =
// tid >> 7) owns channels [32*part, 32*part+32) of BOTH branches of its point - same TMEM lane, half
// of the columns.  Twice the warps per SM for the same shared memory: the latency-bound phases
// (TMEM loads, column reductions, tile writes) interleave across 8 warps instead of 4.
// =============================================================================================
constexpr int NT2 = 256;

// column sums of ONE quantity per part: scratch = float[2 parts][128][33]; returns the partial over
// rows [32*q, 32*q+32) (q = warp & 3) of column (lane) of this thread's part
__device__ __forceinline__ float colreduce32_part(float* scratch, const float va[32], int row, int part, int lane, int quarter) {
  float* sc = scratch + part * (DPF_TILE * 33);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; ++i) sc[row * 33 + i] = va[i];
  __syncthreads();
  float sa = 0.f;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) sa += sc[(quarter * 32 + r) * 33 + lane];
  return sa;
}

struct TcP2Smem2 {
  unsigned char W[2 * N_IMG * IMG_W];   // [br][W1 hi, W1 lo, W1^T hi]
  unsigned char H[2 * IMG_H];           // h1 hi tiles
  unsigned char D[2 * IMG_H];           // h1 lo tiles, then dh2pre tiles, then dz tiles
  unsigned char X[2 * IMG_H];           // per-point weights {1, xk0 hi, xk0 lo, xk1 hi, xk1 lo, 0...} (block 1 = zeros); 1024-aligned
  TcCommon c;
  float m1[2][F], m2[2][F];
  float bnrow[5][2 * F];                // rows 0..4 of the BN_a-sum accumulator at kernel end
  float t1buf[DPF_TILE][2];
};

template <int K, int MODE, bool SPLIT>
__global__ void __launch_bounds__(NT2)
coupling_bwd_p2_tc2_kernel(const BwdArgs a, const unsigned short* __restrict__ wimg) {
  extern __shared__ unsigned char smraw[];
  TcP2Smem2& s = *reinterpret_cast<TcP2Smem2*>(smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, part = tid >> 7, quarter = warp & 3;
  const BranchLayout lay = branch_layout(a.f.k, a.f.w, a.f.G);
  int t0, t1;
  tile_range(a.f.n_tiles, t0, t1);
  DPF_STAMP(2, 0);
  DPF_STAMP_NS(2, 14);
  pdl_launch_dependents();
  const uint32_t tmem = tc_setup(s.c, 512);
  DPF_STAMP(2, 1);
  tc_load_weights<true>(s.c, s.W, wimg);
  pdl_wait();          // pass 1's FiLM sums are read from here on
  tc_prologue_tables(a.f, lay, s.c, false, true);
  DPF_STAMP(2, 2);
  if (tid < 128) {
    const int br = tid >> 6, c = tid & 63;
    float m1 = 0.f, m2 = 0.f;
    if (a.f.training) {
      double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
      const float* fs = a.f.film + (size_t)(br * 2 + 0) * a.f.B * F + c;
      const float* dsr = a.dfilm + (size_t)(br * 2 + 0) * a.f.B * F + c;
      const float* dtr = a.dfilm + (size_t)(br * 2 + 1) * a.f.B * F + c;
      int b = 0;
      for (; b + 4 <= a.f.B; b += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double sc = fs[(size_t)(b + u) * F];
          s1[u] += sc * (double)dtr[(size_t)(b + u) * F];
          s2[u] += sc * (double)dsr[(size_t)(b + u) * F];
        }
      }
      for (; b < a.f.B; ++b) {
        const double sc = fs[(size_t)b * F];
        s1[0] += sc * (double)dtr[(size_t)b * F];
        s2[0] += sc * (double)dsr[(size_t)b * F];
      }
      const double M = (double)a.f.B * (double)a.f.N;
      m1 = (float)(((s1[0] + s1[1]) + (s1[2] + s1[3])) / M);
      m2 = (float)(((s2[0] + s2[1]) + (s2[2] + s2[3])) / M);
    }
    s.m1[br][c] = m1;
    s.m2[br][c] = m2;
  }
  for (int i = tid; i < (int)(2 * IMG_H / 16); i += NT2) reinterpret_cast<uint4*>(s.X)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const Pending P = tc_compute_pending(a, false, s.c.pend);
  DPF_STAMP(2, 3);
  const float sig1 = sqrtf(a.f.eps + 1.0f);
  umma::mbar_wait(&s.c.bar_load, 0);
  __syncthreads();
  DPF_STAMP(2, 4);

  uint32_t phase = 0, phase_aux = 0;
  const uint32_t T_FWD = tmem, T_DG = tmem + 128, T_WG = tmem + 256, T_BN = tmem + 384;
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
  for (int tile = t0; tile < t1; ++tile) {
    const int b = tile / a.f.tiles_per_b;
    const int n = (tile - b * a.f.tiles_per_b) * DPF_TILE + row;
    const bool valid = n < a.f.N;
    tc_tile_film(a.f, s.c, b);
    const TcPoint g = tc_finish_point<MODE>(a, P, tc_load_raw(a, b, n, valid), valid);
    if (tile == t0) DPF_STAMP(2, 5);
    const float xk0 = pick3(g.x, a.f.keep0);
    const float xk1 = (K == 2) ? pick3(g.x, a.f.keep1) : 0.f;
    if (tile > t0) {   // the previous tile's BN-sum UMMAs still read the D / X tiles
      umma::mbar_wait(&s.c.bar_aux, phase_aux);
      phase_aux ^= 1;
    }
    // h1 (hi / lo) chunks of this part: 4 x 16 B per branch
#pragma unroll
    for (int br = 0; br < 2; ++br) {
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const int q = part * 4 + qq;
        uint32_t w[4], wl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 Aa = s.c.A0[br][q * 8 + 2 * i], Ab = s.c.A0[br][q * 8 + 2 * i + 1];
          float va = fmaf(Aa.x, xk0, Aa.z), vb = fmaf(Ab.x, xk0, Ab.z);
          if (K == 2) { va = fmaf(Aa.y, xk1, va); vb = fmaf(Ab.y, xk1, vb); }
          va = fmaxf(va, 0.f);
          vb = fmaxf(vb, 0.f);
          w[i] = umma::pack_bf16(va, vb);
          if (SPLIT) wl[i] = umma::pack_bf16(va - __uint_as_float(w[i] << 16), vb - __uint_as_float(w[i] & 0xffff0000u));
        }
        const uint32_t off = umma::sw128_offset(row, q);
        *reinterpret_cast<uint4*>(s.H + br * IMG_H + off) = make_uint4(w[0], w[1], w[2], w[3]);
        if (SPLIT) *reinterpret_cast<uint4*>(s.D + br * IMG_H + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
      }
    }
    if (tile == t0) DPF_STAMP(2, 6);
    umma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      issue_gemm1<SPLIT>(T_FWD, s.H, s.D, wimg_at<true>(s.W, 0, 0), wimg_at<true>(s.W, 0, 1));
      issue_gemm1<SPLIT>(T_FWD + F, s.H + IMG_H, s.D + IMG_H, wimg_at<true>(s.W, 1, 0), wimg_at<true>(s.W, 1, 1));
      umma::mma_commit(&s.c.bar_mma);
    }
    umma::mbar_wait(&s.c.bar_mma, phase);
    phase ^= 1;
    umma::fence_after_sync();
    if (tile == t0) DPF_STAMP(2, 7);
    // ---- epilogue A: dh2pre (bf16) of this part's 32 channels of each branch -> D tiles ----
#pragma unroll 1
    for (int br = 0; br < 2; ++br) {
      const float d0 = br == 0 ? g.do_mu[0] : g.do_lv[0];
      const float d1 = br == 0 ? g.do_mu[1] : g.do_lv[1];
      float v[32];
      umma::tmem_ld32(T_FWD + lane_off + br * F + part * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int c = part * 32 + i;
        const float4 e = s.c.epi[br][c];
        const float h2n = (v[i] - s.c.mb[br][c]) * s.c.ib[br][c];
        const float av = fmaf(e.x, v[i], e.y);
        const float da = av > 0.f ? fmaf(e.z, d0, e.w * d1) : 0.f;
        const float dh = s.c.ib[br][c] * (da * s.c.sraw[br][c] - s.m1[br][c] - h2n * s.m2[br][c]);
        v[i] = valid ? dh : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 pk = make_uint4(umma::pack_bf16(v[8 * q + 0], v[8 * q + 1]), umma::pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                    umma::pack_bf16(v[8 * q + 4], v[8 * q + 5]), umma::pack_bf16(v[8 * q + 6], v[8 * q + 7]));
        *reinterpret_cast<uint4*>(s.D + br * IMG_H + umma::sw128_offset(row, part * 4 + q)) = pk;
      }
    }
    if (tile == t0) DPF_STAMP(2, 8);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int br = 0; br < 2; ++br)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma::mma_bf16(T_DG + br * F, umma::desc_at(DESC_K, umma::smem_u32(s.D + br * IMG_H) + 32 * k),
                         umma::desc_at(DESC_K, umma::smem_u32(wimg_at<true>(s.W, br, 2)) + 32 * k), IDESC_GEMM, k > 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma::mma_bf16(T_WG, umma::desc_at(DESC_MN, umma::smem_u32(s.D) + 2048 * k), umma::desc_at(DESC_MN, umma::smem_u32(s.H) + 2048 * k),
                       IDESC_WGRAD, (tile > t0 || k > 0) ? 1u : 0u);
      umma::mma_commit(&s.c.bar_mma);
    }
    umma::mbar_wait(&s.c.bar_mma, phase);
    phase ^= 1;
    umma::fence_after_sync();
    if (tile == t0) DPF_STAMP(2, 9);
    // ---- epilogue B: dz, T1, BN_a sums over this part's channels ----
    // The per-channel sums over points  dbeta = sum dz,  E = sum dz * x_keep  run on the tensor cores:
    // dz (bf16) overwrites the D tiles (their UMMAs are complete) and is multiplied by the per-point
    // weight tile X = {1, xk0 hi, xk0 lo, xk1 hi, xk1 lo} - the same MN-major operand pair as wgrad,
    // accumulated in TMEM across all tiles of the CTA.
    float T1_0 = 0.f, T1_1 = 0.f;
#pragma unroll 1
    for (int br = 0; br < 2; ++br) {
      float v[32];
      umma::tmem_ld32(T_DG + lane_off + br * F + part * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float4 A = s.c.A0[br][part * 32 + i];
        float z = fmaf(A.x, xk0, A.z);
        if (K == 2) z = fmaf(A.y, xk1, z);
        const float dz = (z > 0.f && valid) ? v[i] : 0.f;
        T1_0 = fmaf(A.x, dz, T1_0);
        if (K == 2) T1_1 = fmaf(A.y, dz, T1_1);
        v[i] = dz;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 pk = make_uint4(umma::pack_bf16(v[8 * q + 0], v[8 * q + 1]), umma::pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                    umma::pack_bf16(v[8 * q + 4], v[8 * q + 5]), umma::pack_bf16(v[8 * q + 6], v[8 * q + 7]));
        *reinterpret_cast<uint4*>(s.D + br * IMG_H + umma::sw128_offset(row, part * 4 + q)) = pk;
      }
    }
    if (part == 0) {
      const float one = valid ? 1.f : 0.f;
      const uint32_t w0 = umma::pack_bf16(one, xk0);                                   // {1, xk0 hi}
      const float xk0_lo = xk0 - __uint_as_float(w0 & 0xffff0000u);
      const uint32_t w1 = umma::pack_bf16(xk0_lo, xk1);                                // {xk0 lo, xk1 hi}
      const float xk1_lo = xk1 - __uint_as_float(w1 & 0xffff0000u);
      const uint32_t w2 = umma::pack_bf16(xk1_lo, 0.f);                                // {xk1 lo, 0}
      *reinterpret_cast<uint4*>(s.X + umma::sw128_offset(row, 0)) = make_uint4(w0, w1, w2, 0u);
    } else {
      s.t1buf[row][0] = T1_0;
      s.t1buf[row][1] = T1_1;
    }
    if (tile == t0) DPF_STAMP(2, 10);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma::mma_bf16(T_BN, umma::desc_at(DESC_MN, umma::smem_u32(s.X) + 2048 * k), umma::desc_at(DESC_MN, umma::smem_u32(s.D) + 2048 * k),
                       IDESC_WGRAD, (tile > t0 || k > 0) ? 1u : 0u);
      umma::mma_commit(&s.c.bar_aux);
    }
    if (part == 0 && valid) {
      T1_0 += s.t1buf[row][0];
      T1_1 += s.t1buf[row][1];
      float dx[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) dx[ch] = (MODE == 1) ? g.dy[ch] / sig1 : g.dy[ch] * sig1;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        if (ch == a.f.keep0) dx[ch] += T1_0;
        if (K == 2 && ch == a.f.keep1) dx[ch] += T1_1;
        if (ch == a.f.warp0) dx[ch] = (MODE == 1) ? g.dy[ch] / g.sig[0] : g.dy[ch] * g.sig[0];
        if (K == 1 && ch == a.f.warp1) dx[ch] = (MODE == 1) ? g.dy[ch] / g.sig[1] : g.dy[ch] * g.sig[1];
      }
      const size_t base = (size_t)b * 3 * a.f.N + n;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) a.dx_out[base + (size_t)ch * a.f.N] = dx[ch];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tile == t0) DPF_STAMP(2, 11);
  }
  // ---- CTA epilogue ----
  DPF_STAMP(2, 12);
  if (t1 > t0) {
    umma::mbar_wait(&s.c.bar_aux, phase_aux);
    umma::fence_after_sync();
    if (quarter == 0) {   // accumulator rows 0..4 live in TMEM lanes 0..4: warps 0 (columns 0..63) and 4 (64..127)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
        umma::tmem_ld32(T_BN + part * F + half * 32, v);
        if (lane < 5) {
#pragma unroll
          for (int i = 0; i < 32; ++i) s.bnrow[lane][part * F + half * 32 + i] = v[i];
        }
      }
    }
  }
  __syncthreads();
  if (tid < 128 && t1 > t0) {
    DPF_GATOMIC(atomicAdd(&a.bna_sums[tid * 4 + 0], (double)s.bnrow[0][tid]));
    DPF_GATOMIC(atomicAdd(&a.bna_sums[tid * 4 + 1], (double)s.bnrow[1][tid] + (double)s.bnrow[2][tid]));
    DPF_GATOMIC(atomicAdd(&a.bna_sums[tid * 4 + 2], (double)s.bnrow[3][tid] + (double)s.bnrow[4][tid]));
  }
  {
    // accumulator row = branch*64 + channel; this thread stores columns [part*32, +32) of its branch's block
    const int br = row >> 6, c = row & 63;
    float4* d = reinterpret_cast<float4*>(a.dw1_partial + ((size_t)blockIdx.x * 2 + br) * (F * F) + c * F + part * 32);
    umma::fence_after_sync();
    float v[32];
    if (t1 > t0) {
      umma::tmem_ld32(T_WG + lane_off + br * F + part * 32, v);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
  DPF_STAMP(2, 13);
  DPF_STAMP_NS(2, 15);
}

constexpr uint32_t IDESC_RED = umma::make_idesc_bf16(128, 16, 1, 1);

// =============================================================================================
// Backward pass 2, TWO tiles in flight per SM: a 544-thread CTA = two 256-thread halves, each running
// the two-threads-per-point tile pipeline above on its own tile (own H / D / X tiles, own 128 TMEM
// columns for the recompute / dgrad accumulators), plus ONE MMA-issuer warp that serves both halves:
// a half writes its operand tiles, arrives on its request mbarrier, the issuer's elected lane issues
// that stage's UMMAs and commits to the half's completion mbarrier.  One issuing thread keeps every
// accumulation into the shared wgrad / BN_a-sum accumulators in program order.  While one half waits
// on a UMMA round trip or its global loads, the other half's CUDA-core phase runs.
//   TMEM: [0,128) half 0 recompute -> dgrad, [128,256) half 1, [256,384) wgrad, [384,400) BN_a sums.
// The BN_a sums use the operand roles of pass 1's reductions: A = dz tile (M = 128 channels of both
// branches, MN-major), B = per-point weights {1, xk0 hi, xk0 lo, xk1 hi, xk1 lo} (N = 16), so lane m
// = channel, columns 0..4 = the five sums.
// The deferred correction P and the BN_b batch terms m1, m2 are not recomputed: pass 1 hands them over
// (pend_store, m12_rep).
// =============================================================================================
constexpr int NT4 = 512 + 32;

struct TcP2Half {
  unsigned char H[2 * IMG_H];           // h1 hi tiles [br]
  unsigned char D[2 * IMG_H];           // h1 lo tiles, then dh2pre tiles, then dz tiles
  unsigned char X[IMG_H];               // per-point weights, chunk 0 of every row (the rest stays zero)
};
struct TcP2Smem4 {
  unsigned char W[2 * N_IMG * IMG_W];   // [br][W1 hi, W1 lo, W1^T hi]
  TcP2Half h[2];
  // tables in CHANNEL-PAIR layout (a = channel 2j, b = channel 2j+1) for the packed fp32x2 epilogue math
  float4 PA[2][F / 2][2];               // [br][j]: {A00_a, A00_b, c0_a, c0_b}, {A01_a, A01_b, -, -}   (folded BN_a)
  float4 PE[2][2][F / 2][3];            // [half][br][j]: {S_a, S_b, T_a, T_b}, {W20_a, W20_b, W21_a, W21_b}, {-c3_a, -c3_b, -c2_a, -c2_b}
                                        //   a = S*acc + T ;  dh2pre = S*da - c2 - c3*acc
  float4 lt0[2][F];                     // per channel: {bnB istd, bnB mean, c2, c3}
  float2 lt1[2][F];                     // per channel: {W2_0, W2_1}
  float pend[8];
  float t1buf[2][DPF_TILE][2];
  uint64_t bar_load, bar_req[2], bar_done[2];
  uint32_t tmem_base;
};

template <int K, int MODE, bool SPLIT>
__global__ void __launch_bounds__(NT4, 1)   // 17 warps are allocated as 20: 96 registers per thread
coupling_bwd_p2_tc4_kernel(const BwdArgs a, const unsigned short* __restrict__ wimg) {
  extern __shared__ unsigned char smraw[];
  TcP2Smem4& s = *reinterpret_cast<TcP2Smem4*>(smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = tid >= 512;
  const int half = (tid >> 8) & 1, row = tid & 127, part = (tid >> 7) & 1, quarter = warp & 3;
  const BranchLayout lay = branch_layout(a.f.k, a.f.w, a.f.G);
  const int n_workers = 2 * gridDim.x;
  const int worker = 2 * blockIdx.x + half;
  auto iters_of = [&](int w) { return w < a.f.n_tiles ? (a.f.n_tiles - w + n_workers - 1) / n_workers : 0; };
  DPF_STAMP(2, 0);
  DPF_STAMP_NS(2, 14);
  pdl_launch_dependents();
  if (tid == 0) {
    umma::mbar_init(&s.bar_load, 1);
    umma::mbar_init(&s.bar_req[0], 256);
    umma::mbar_init(&s.bar_req[1], 256);
    umma::mbar_init(&s.bar_done[0], 1);
    umma::mbar_init(&s.bar_done[1], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(&s.tmem_base, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  if (tid == 0) {
    umma::mbar_expect_tx(&s.bar_load, 2 * N_IMG * IMG_W);
    umma::bulk_g2s(s.W, wimg, 2 * N_IMG * IMG_W, &s.bar_load);
  }
  if (!issuer) {       // independent of the previous kernels: zero the per-point weight tiles
    for (int i = tid; i < (int)(IMG_H / 16); i += 512) {
      reinterpret_cast<uint4*>(s.h[0].X)[i] = make_uint4(0u, 0u, 0u, 0u);
      reinterpret_cast<uint4*>(s.h[1].X)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  DPF_STAMP(2, 1);
  pdl_wait();          // pass 1's sums / pending correction are read from here on
  DPF_STAMP(2, 2);
  DPF_STAMP_NS(2, 13);
  const uint32_t T_WG = tmem + 256, T_BN = tmem + 384;

  if (issuer) {
    // ------------------------------ MMA issuer warp ------------------------------
    if (lane == 0) {
      umma::mbar_wait(&s.bar_load, 0);
      int n_it[2] = {iters_of(2 * (int)blockIdx.x), iters_of(2 * (int)blockIdx.x + 1)};
      int it[2] = {0, 0}, stage[2] = {0, 0};
      uint32_t ph[2] = {0u, 0u};
      uint32_t wg_acc = 0u, bn_acc = 0u;
      int remaining = 3 * (n_it[0] + n_it[1]);
      while (remaining > 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (it[h] < n_it[h] && umma::mbar_test(&s.bar_req[h], ph[h])) {
            ph[h] ^= 1u;
            umma::fence_after_sync();
            TcP2Half& hb = s.h[h];
            const uint32_t T_F = tmem + h * 128;
            if (stage[h] == 0) {          // forward recompute, both branches
              issue_gemm1<SPLIT>(T_F, hb.H, hb.D, wimg_at<true>(s.W, 0, 0), wimg_at<true>(s.W, 0, 1));
              issue_gemm1<SPLIT>(T_F + F, hb.H + IMG_H, hb.D + IMG_H, wimg_at<true>(s.W, 1, 0), wimg_at<true>(s.W, 1, 1));
            } else if (stage[h] == 1) {   // dgrad (overwrites the recompute accumulators) + wgrad
#pragma unroll
              for (int br = 0; br < 2; ++br)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma::mma_bf16(T_F + br * F, umma::desc_at(DESC_K, umma::smem_u32(hb.D + br * IMG_H) + 32 * k),
                                 umma::desc_at(DESC_K, umma::smem_u32(wimg_at<true>(s.W, br, 2)) + 32 * k), IDESC_GEMM, k > 0);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                umma::mma_bf16(T_WG, umma::desc_at(DESC_MN, umma::smem_u32(hb.D) + 2048 * k), umma::desc_at(DESC_MN, umma::smem_u32(hb.H) + 2048 * k),
                               IDESC_WGRAD, wg_acc);
                wg_acc = 1u;
              }
            } else {                      // BN_a sums
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                umma::mma_bf16(T_BN, umma::desc_at(DESC_MN, umma::smem_u32(hb.D) + 2048 * k), umma::desc_at(DESC_MN, umma::smem_u32(hb.X) + 2048 * k),
                               IDESC_RED, bn_acc);
                bn_acc = 1u;
              }
            }
            umma::mma_commit(&s.bar_done[h]);
            if (++stage[h] == 3) { stage[h] = 0; ++it[h]; }
            --remaining;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ the two tile pipelines ------------------------------
    TcP2Half& hb = s.h[half];
    const int n_it = iters_of(worker);
    // first tile's global loads are issued before the table work
    TcRaw raw0;
    {
      const int b0 = worker / a.f.tiles_per_b;
      const int n0 = (worker - b0 * a.f.tiles_per_b) * DPF_TILE + row;
      raw0 = tc_load_raw(a, b0, n0, n_it > 0 && n0 < a.f.N);
    }
    // per-channel constants.  BN_b backward of a channel:
    //   dh2pre = ib*(da*s - m1 - h2n*m2),  h2n = (acc - mb)*ib   =>   dh2pre = (ib*s)*da - c2 - c3*acc
    if (tid < 128) {
      const float4 t0 = reinterpret_cast<const float4*>(a.ltab)[tid * (DPF_LTAB_ROW / 4) + 0];
      const float4 t1 = reinterpret_cast<const float4*>(a.ltab)[tid * (DPF_LTAB_ROW / 4) + 1];
      float c2 = 0.f, c3 = 0.f;
      if (a.f.training) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int r = 0; r < DPF_M12_REP; ++r) {
          const double2 v = *reinterpret_cast<const double2*>(a.m12_rep + (size_t)r * (2 * F * 2) + (size_t)tid * 2);
          s1 += v.x;
          s2 += v.y;
        }
        const float rM = 1.f / ((float)a.f.B * (float)a.f.N);
        const float m1 = (float)s1 * rM, m2 = (float)s2 * rM;
        c3 = t1.x * t1.x * m2;
        c2 = t1.x * m1 - c3 * t0.w;
      }
      {
        float* pa = reinterpret_cast<float*>(&s.PA[tid >> 6][(tid & 63) >> 1][0]);
        const int ln = tid & 1;
        pa[ln] = t0.x;        // A00
        pa[2 + ln] = t0.z;    // c0
        pa[4 + ln] = t0.y;    // A01
        pa[6 + ln] = 0.f;
      }
      s.lt0[tid >> 6][tid & 63] = make_float4(t1.x, t0.w, c2, c3);
      s.lt1[tid >> 6][tid & 63] = make_float2(t1.y, t1.z);
    } else if (tid < 128 + 5) {
      s.pend[tid - 128] = a.has_pending ? a.pend_store[tid - 128] : 0.f;
    }
    const float sig1 = sqrtf(a.f.eps + 1.0f);
    umma::named_bar_sync(3, 512);        // tables + zeroed X tiles visible to both halves
    DPF_STAMP(2, 3);

    uint32_t ph = 0;
    const uint32_t T_F = tmem + half * 128;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint64_t* req = &s.bar_req[half];
    uint64_t* done = &s.bar_done[half];
    for (int it = 0; it < n_it; ++it) {
      const int tile = worker + it * n_workers;
      const int b = tile / a.f.tiles_per_b;
      const int n = (tile - b * a.f.tiles_per_b) * DPF_TILE + row;
      const bool valid = n < a.f.N;
      if (part == 0) {   // FiLM fold of this tile's shape (read again only after the next request / completion round trip)
        const int br = row >> 6, c = row & 63;
        const float sc = a.f.film[((size_t)(br * 2 + 0) * a.f.B + b) * F + c];
        const float sh = a.f.film[((size_t)(br * 2 + 1) * a.f.B + b) * F + c];
        const float4 l0 = s.lt0[br][c];
        const float2 l1 = s.lt1[br][c];
        const float S = sc * l0.x;
        float* pe = reinterpret_cast<float*>(&s.PE[half][br][c >> 1][0]);
        const int ln = c & 1;
        pe[ln] = S;
        pe[2 + ln] = fmaf(-S, l0.y, sh);
        pe[4 + ln] = l1.x;
        pe[6 + ln] = l1.y;
        pe[8 + ln] = -l0.w;     // -c3
        pe[10 + ln] = -l0.z;    // -c2
      }
      const Pending P{s.pend[0], s.pend[1], s.pend[2], s.pend[3], s.pend[4]};
      const TcPoint g = tc_finish_point<MODE>(a, P, it == 0 ? raw0 : tc_load_raw(a, b, n, valid), valid);
      const float xk0 = pick3(g.x, a.f.keep0);
      const float xk1 = (K == 2) ? pick3(g.x, a.f.keep1) : 0.f;
      const f32x2 xk0_2 = f2_pack(xk0, xk0), xk1_2 = f2_pack(xk1, xk1);
      if (it == 0) DPF_STAMP(2, 4);
      // ---- stage 0: h1 hi -> H, lo -> D (the previous tile's BN_a-sum UMMAs still read D / X: wait for them) ----
      if (it > 0) {
        umma::mbar_wait(done, ph);
        ph ^= 1;
      }
#pragma unroll
      for (int br = 0; br < 2; ++br) {
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int q = part * 4 + qq;
          uint32_t w[4], wl[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 P0 = s.PA[br][q * 4 + i][0];
            f32x2 v2 = f2_fma(f2_pack(P0.x, P0.y), xk0_2, f2_pack(P0.z, P0.w));
            if (K == 2) {
              const float4 P1 = s.PA[br][q * 4 + i][1];
              v2 = f2_fma(f2_pack(P1.x, P1.y), xk1_2, v2);
            }
            float va, vb;
            f2_unpack(v2, va, vb);
            va = fmaxf(va, 0.f);
            vb = fmaxf(vb, 0.f);
            w[i] = umma::pack_bf16(va, vb);
            if (SPLIT) wl[i] = umma::pack_bf16(va - __uint_as_float(w[i] << 16), vb - __uint_as_float(w[i] & 0xffff0000u));
          }
          const uint32_t off = umma::sw128_offset(row, q);
          *reinterpret_cast<uint4*>(hb.H + br * IMG_H + off) = make_uint4(w[0], w[1], w[2], w[3]);
          if (SPLIT) *reinterpret_cast<uint4*>(hb.D + br * IMG_H + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
      }
      umma::fence_async_smem();
      if (it == 0) DPF_STAMP(2, 5);
      umma::mbar_arrive(req);
      umma::mbar_wait(done, ph);
      ph ^= 1;
      umma::fence_after_sync();
      if (it == 0) DPF_STAMP(2, 6);
      // ---- stage 1: epilogue A: dh2pre (bf16) of this part's 32 channels of each branch -> D tiles ----
#pragma unroll 1
      for (int br = 0; br < 2; ++br) {
        const float d0 = br == 0 ? g.do_mu[0] : g.do_lv[0];
        const float d1 = br == 0 ? g.do_mu[1] : g.do_lv[1];
        const f32x2 d0_2 = f2_pack(d0, d0), d1_2 = f2_pack(d1, d1);
#pragma unroll 1
        for (int hc = 0; hc < 2; ++hc) {     // 16 columns at a time (register pressure: 96 per thread)
          uint32_t r[16];
          umma::tmem_ld16_issue(T_F + lane_off + br * F + part * 32 + hc * 16, r);
          umma::tmem_ld_wait16(r);
          uint32_t wv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {      // channel pair (2j, 2j+1) of this chunk, packed fp32x2 math
            const float4* pe = s.PE[half][br][part * 16 + hc * 8 + j];
            const float4 E0 = pe[0], E1 = pe[1], E2 = pe[2];
            const f32x2 S2 = f2_pack(E0.x, E0.y);
            const f32x2 v2 = f2_pack(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            const f32x2 av2 = f2_fma(S2, v2, f2_pack(E0.z, E0.w));
            const f32x2 g2 = f2_fma(f2_pack(E1.x, E1.y), d0_2, f2_mul(f2_pack(E1.z, E1.w), d1_2));
            float av0, av1, g0, g1;
            f2_unpack(av2, av0, av1);
            f2_unpack(g2, g0, g1);
            const f32x2 da2 = f2_pack(av0 > 0.f ? g0 : 0.f, av1 > 0.f ? g1 : 0.f);
            const f32x2 dh2 = f2_fma(S2, da2, f2_fma(f2_pack(E2.x, E2.y), v2, f2_pack(E2.z, E2.w)));
            float h0, h1;
            f2_unpack(dh2, h0, h1);
            wv[j] = valid ? umma::pack_bf16(h0, h1) : 0u;
          }
#pragma unroll
          for (int q = 0; q < 2; ++q)
            *reinterpret_cast<uint4*>(hb.D + br * IMG_H + umma::sw128_offset(row, part * 4 + hc * 2 + q)) =
                make_uint4(wv[4 * q], wv[4 * q + 1], wv[4 * q + 2], wv[4 * q + 3]);
        }
      }
      umma::fence_async_smem();
      umma::fence_before_sync();
      if (it == 0) DPF_STAMP(2, 7);
      umma::mbar_arrive(req);
      umma::mbar_wait(done, ph);
      ph ^= 1;
      umma::fence_after_sync();
      if (it == 0) DPF_STAMP(2, 8);
      // ---- stage 2: epilogue B: dz (bf16) -> D tiles, T1 = A0^T dz, per-point weight row -> X ----
      f32x2 T1_0_2 = f2_pack(0.f, 0.f), T1_1_2 = f2_pack(0.f, 0.f);     // even / odd channel partial sums
#pragma unroll 1
      for (int br = 0; br < 2; ++br) {
#pragma unroll 1
        for (int hc = 0; hc < 2; ++hc) {
          uint32_t r[16];
          umma::tmem_ld16_issue(T_F + lane_off + br * F + part * 32 + hc * 16, r);
          umma::tmem_ld_wait16(r);
          uint32_t wv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4* pa = s.PA[br][part * 16 + hc * 8 + j];
            const float4 P0 = pa[0];
            const f32x2 A00_2 = f2_pack(P0.x, P0.y);
            f32x2 z2 = f2_fma(A00_2, xk0_2, f2_pack(P0.z, P0.w));
            f32x2 A01_2 = 0;
            if (K == 2) {
              const float4 P1 = pa[1];
              A01_2 = f2_pack(P1.x, P1.y);
              z2 = f2_fma(A01_2, xk1_2, z2);
            }
            float z0, z1;
            f2_unpack(z2, z0, z1);
            const float dz0 = (z0 > 0.f && valid) ? __uint_as_float(r[2 * j]) : 0.f;
            const float dz1 = (z1 > 0.f && valid) ? __uint_as_float(r[2 * j + 1]) : 0.f;
            const f32x2 dz2 = f2_pack(dz0, dz1);
            T1_0_2 = f2_fma(A00_2, dz2, T1_0_2);
            if (K == 2) T1_1_2 = f2_fma(A01_2, dz2, T1_1_2);
            wv[j] = umma::pack_bf16(dz0, dz1);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q)
            *reinterpret_cast<uint4*>(hb.D + br * IMG_H + umma::sw128_offset(row, part * 4 + hc * 2 + q)) =
                make_uint4(wv[4 * q], wv[4 * q + 1], wv[4 * q + 2], wv[4 * q + 3]);
        }
      }
      float T1_0, T1_1;
      {
        float lo, hi;
        f2_unpack(T1_0_2, lo, hi);
        T1_0 = lo + hi;
        f2_unpack(T1_1_2, lo, hi);
        T1_1 = lo + hi;
      }
      if (part == 0) {
        const float one = valid ? 1.f : 0.f;
        const uint32_t w0 = umma::pack_bf16(one, xk0);                                   // {1, xk0 hi}
        const float xk0_lo = xk0 - __uint_as_float(w0 & 0xffff0000u);
        const uint32_t w1 = umma::pack_bf16(xk0_lo, xk1);                                // {xk0 lo, xk1 hi}
        const float xk1_lo = xk1 - __uint_as_float(w1 & 0xffff0000u);
        const uint32_t w2 = umma::pack_bf16(xk1_lo, 0.f);                                // {xk1 lo, 0}
        *reinterpret_cast<uint4*>(hb.X + umma::sw128_offset(row, 0)) = make_uint4(w0, w1, w2, 0u);
      } else {
        s.t1buf[half][row][0] = T1_0;
        s.t1buf[half][row][1] = T1_1;
      }
      umma::fence_async_smem();
      umma::fence_before_sync();
      if (it == 0) DPF_STAMP(2, 9);
      umma::mbar_arrive(req);
      umma::named_bar_sync(1 + half, 256);   // t1buf hand-over inside the half
      if (part == 0 && valid) {
        T1_0 += s.t1buf[half][row][0];
        T1_1 += s.t1buf[half][row][1];
        float dx[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) dx[ch] = (MODE == 1) ? g.dy[ch] / sig1 : g.dy[ch] * sig1;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          if (ch == a.f.keep0) dx[ch] += T1_0;
          if (K == 2 && ch == a.f.keep1) dx[ch] += T1_1;
          if (ch == a.f.warp0) dx[ch] = (MODE == 1) ? g.dy[ch] / g.sig[0] : g.dy[ch] * g.sig[0];
          if (K == 1 && ch == a.f.warp1) dx[ch] = (MODE == 1) ? g.dy[ch] / g.sig[1] : g.dy[ch] * g.sig[1];
        }
        const size_t base = (size_t)b * 3 * a.f.N + n;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) a.dx_out[base + (size_t)ch * a.f.N] = dx[ch];
      }
    }
    if (n_it > 0) DPF_STAMP(2, 10);
    if (n_it > 0) {     // the last tile's BN_a-sum UMMAs
      umma::mbar_wait(done, ph);
      ph ^= 1;
    }
    umma::fence_before_sync();
  }
  // ---- CTA epilogue (every UMMA of both halves has completed) ----
  __syncthreads();
  DPF_STAMP(2, 11);
  umma::fence_after_sync();
  const bool had_work = 2 * (int)blockIdx.x < a.f.n_tiles;
  if (tid < 128 && had_work) {     // lane = channel m = br*64 + c; columns {dbeta, E0 hi, E0 lo, E1 hi, E1 lo}
    uint32_t r0[4], r1[4];
    umma::tmem_ld4(T_BN + ((uint32_t)(quarter * 32) << 16), r0);
    umma::tmem_ld4(T_BN + ((uint32_t)(quarter * 32) << 16) + 4, r1);
    umma::tmem_ld_wait4(r0);
    umma::tmem_ld_wait4(r1);
    DPF_GATOMIC(atomicAdd(&a.bna_sums[tid * 4 + 0], (double)__uint_as_float(r0[0])));
    DPF_GATOMIC(atomicAdd(&a.bna_sums[tid * 4 + 1], (double)__uint_as_float(r0[1]) + (double)__uint_as_float(r0[2])));
    DPF_GATOMIC(atomicAdd(&a.bna_sums[tid * 4 + 2], (double)__uint_as_float(r0[3]) + (double)__uint_as_float(r1[0])));
  }
  if (!issuer) {
    // accumulator row = branch*64 + channel; this thread stores 16 columns of its branch's 64x64 block
    const int br = row >> 6, c = row & 63, cg = half * 2 + part;
    float4* d = reinterpret_cast<float4*>(a.dw1_partial + ((size_t)blockIdx.x * 2 + br) * (F * F) + c * F + cg * 16);
    uint32_t r[16];
    if (had_work) {
      umma::tmem_ld16_issue(T_WG + ((uint32_t)(quarter * 32) << 16) + br * F + cg * 16, r);
      umma::tmem_ld_wait16(r);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = 0u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      d[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
  DPF_STAMP(2, 12);
  DPF_STAMP_NS(2, 15);
}

// =============================================================================================
// MERGED backward launch of one layer: pass 1 -> software grid barrier -> pass 2 in ONE kernel (one CTA per SM,
// the two-pipelines + MMA-issuer-warp organisation of coupling_bwd_p2_tc4_kernel for BOTH phases).  What it removes
// per layer: one launch transition (the pass-2 CTAs own a whole SM, so PDL cannot overlap pass 1 <-> pass 2), one
// set of per-CTA prologues (barrier init, TMEM allocation, 48 KB weight stage, tables), the pend / m12 hand-over
// through global memory.  Phase 1 (the per-shape FiLM / last-SharedDot reductions on the tensor cores, see the
// pass-1 kernel below) reuses phase 2's buffers: h3 hi | lo tiles in H[0] | H[1], the 0/1 mask tile in D[1] (its
// ignored second M block is X), the per-point weight rows in X, its accumulators in the wgrad TMEM columns.
//   TMEM: [0,128) / [128,256) the halves' recompute -> dgrad accumulators; phase 1: [256,320) / [320,384) the halves'
//   reduction accumulators {br: h3 sums [0,16), mask sums [16,32)}; phase 2: [256,384) wgrad, [384,400) BN_a sums.
// Tiles: phase 1 gives every pipeline a CONTIGUOUS run of tiles (mostly one shape: one accumulator flush per run),
// phase 2 the strided assignment of the tc4 kernel.
// =============================================================================================
struct TcBwdMSmem {
  unsigned char W[2 * N_IMG * IMG_W];   // [br][W1 hi, W1 lo, W1^T hi]
  TcP2Half h[2];
  float4 PA[2][F / 2][2];               // folded BN_a, channel-pair layout (see TcP2Smem4)
  float4 PE[2][2][F / 2][3];            // [half][br][j]: {S_a, S_b, T_a, T_b}, {W20_a, W20_b, W21_a, W21_b}, phase 2: {-c3_a, -c3_b, -c2_a, -c2_b};
                                        //   phase 1: slot [2] = {s_raw_a, s_raw_b, shift_a, shift_b} of the tile's shape
  float4 lt0[2][F];                     // per channel: {bnB istd, bnB mean, c2, c3}
  float2 lt1[2][F];                     // per channel: {W2_0, W2_1}
  float t1buf[2][DPF_TILE][2];          // phase 2: T1 hand-over inside a half; phase 1: lo-part sums of the flush [half][br][w][64]
  double pend_red[20];
  float b2fin[2][2];
  uint64_t bar_load, bar_req[2], bar_done[2];
  uint32_t tmem_base;
};

static_assert(sizeof(TcBwdMSmem) + 1024 <= 232448, "merged backward: shared memory exceeds the 227 KB per-CTA limit");

template <int K, int MODE, bool SPLIT>
__global__ void __launch_bounds__(NT4, 1)
coupling_bwd_merged_kernel(const BwdArgs a, const unsigned short* __restrict__ wimg, unsigned int* __restrict__ barrier_counter) {
  extern __shared__ unsigned char smraw[];
  TcBwdMSmem& s = *reinterpret_cast<TcBwdMSmem*>(smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = tid >= 512;
  const int half = (tid >> 8) & 1, row = tid & 127, part = (tid >> 7) & 1, quarter = warp & 3;
  const BranchLayout lay = branch_layout(a.f.k, a.f.w, a.f.G);
  const int n_workers = 2 * gridDim.x;
  const int worker = 2 * blockIdx.x + half;
  // phase 1: contiguous tiles [p1_lo(w), p1_lo(w+1)); phase 2: tiles w, w + n_workers, ...
  const int p1_per = (a.f.n_tiles + n_workers - 1) / n_workers;
  auto p1_lo = [&](int w) { return min(a.f.n_tiles, w * p1_per); };
  auto iters_of = [&](int w) { return w < a.f.n_tiles ? (a.f.n_tiles - w + n_workers - 1) / n_workers : 0; };
  pdl_launch_dependents();
  if (tid == 0) {
    umma::mbar_init(&s.bar_load, 1);
    umma::mbar_init(&s.bar_req[0], 256);
    umma::mbar_init(&s.bar_req[1], 256);
    umma::mbar_init(&s.bar_done[0], 1);
    umma::mbar_init(&s.bar_done[1], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(&s.tmem_base, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  if (tid == 0) {
    umma::mbar_expect_tx(&s.bar_load, 2 * N_IMG * IMG_W);
    umma::bulk_g2s(s.W, wimg, 2 * N_IMG * IMG_W, &s.bar_load);
  }
  if (!issuer) {       // independent of the previous kernels: zero the per-point weight tiles
    for (int i = tid; i < (int)(IMG_H / 16); i += 512) {
      reinterpret_cast<uint4*>(s.h[0].X)[i] = make_uint4(0u, 0u, 0u, 0u);
      reinterpret_cast<uint4*>(s.h[1].X)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (tid < 4) (&s.b2fin[0][0])[tid] = 0.f;
  pdl_wait();          // the previous backward kernel's sums / gradients are read from here on
  // static tables of this layer (bwd_tables_kernel) + the deferred BN_a correction owed by the previous step
  if (tid < 128) {
    const float4 t0 = reinterpret_cast<const float4*>(a.ltab)[tid * (DPF_LTAB_ROW / 4) + 0];
    const float4 t1 = reinterpret_cast<const float4*>(a.ltab)[tid * (DPF_LTAB_ROW / 4) + 1];
    float* pa = reinterpret_cast<float*>(&s.PA[tid >> 6][(tid & 63) >> 1][0]);
    const int ln = tid & 1;
    pa[ln] = t0.x;        // A00
    pa[2 + ln] = t0.z;    // c0
    pa[4 + ln] = t0.y;    // A01
    pa[6 + ln] = 0.f;
    s.lt0[tid >> 6][tid & 63] = make_float4(t1.x, t0.w, 0.f, 0.f);
    s.lt1[tid >> 6][tid & 63] = make_float2(t1.y, t1.z);
  }
  const Pending P = tc_compute_pending(a, blockIdx.x == 0, s.pend_red);     // all threads (contains __syncthreads)
  const float sig1 = sqrtf(a.f.eps + 1.0f);
  const uint32_t T_WG = tmem + 256, T_BN = tmem + 384;
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;

  // =========================================== phase 1 ===========================================
  if (issuer) {
    if (lane == 0) {
      umma::mbar_wait(&s.bar_load, 0);
      int t_lo[2] = {p1_lo(2 * (int)blockIdx.x), p1_lo(2 * (int)blockIdx.x + 1)};
      int n_it[2] = {p1_lo(2 * (int)blockIdx.x + 1) - t_lo[0], p1_lo(2 * (int)blockIdx.x + 2) - t_lo[1]};
      int it[2] = {0, 0}, stage[2] = {0, 0}, cur_b[2] = {-1, -1};
      uint32_t ph[2] = {0u, 0u}, acc[2] = {0u, 0u};
      int remaining = 3 * (n_it[0] + n_it[1]);
      while (remaining > 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (it[h] < n_it[h] && umma::mbar_test(&s.bar_req[h], ph[h])) {
            ph[h] ^= 1u;
            umma::fence_after_sync();
            TcP2Half& hb = s.h[h];
            const uint32_t T_F = tmem + h * 128, T_R = tmem + 256 + h * 64;
            if (stage[h] == 0) {          // forward recompute, both branches
              const int b = (t_lo[h] + it[h]) / a.f.tiles_per_b;
              acc[h] = (b == cur_b[h]) ? 1u : 0u;       // a new shape starts a new accumulation run
              cur_b[h] = b;
              issue_gemm1<SPLIT>(T_F, hb.H, hb.D, wimg_at<true>(s.W, 0, 0), wimg_at<true>(s.W, 0, 1));
              issue_gemm1<SPLIT>(T_F + F, hb.H + IMG_H, hb.D + IMG_H, wimg_at<true>(s.W, 1, 0), wimg_at<true>(s.W, 1, 1));
            } else {                      // reductions of branch stage - 1: h3 (hi | lo) and mask tiles against the weight rows
              const int br = stage[h] - 1;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma::mma_bf16(T_R + br * 32, umma::desc_at(DESC_MN, umma::smem_u32(hb.H) + 2048 * k),
                               umma::desc_at(DESC_MN, umma::smem_u32(hb.X) + 2048 * k), IDESC_RED, (acc[h] || k > 0) ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma::mma_bf16(T_R + br * 32 + 16, umma::desc_at(DESC_MN, umma::smem_u32(hb.D + IMG_H) + 2048 * k),
                               umma::desc_at(DESC_MN, umma::smem_u32(hb.X) + 2048 * k), IDESC_RED, (acc[h] || k > 0) ? 1u : 0u);
            }
            umma::mma_commit(&s.bar_done[h]);
            if (++stage[h] == 3) { stage[h] = 0; ++it[h]; }
            --remaining;
          }
        }
      }
    }
    __syncwarp();
  } else {
    TcP2Half& hb = s.h[half];
    const int t_lo = p1_lo(worker), n_it = p1_lo(worker + 1) - t_lo;
    const uint32_t T_F = tmem + half * 128, T_R = tmem + 256 + half * 64;
    uint64_t* req = &s.bar_req[half];
    uint64_t* done = &s.bar_done[half];
    uint32_t ph = 0;
    float dW2acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};      // part 0, row < 64: channel row of branch [br], output [w]
    float b2acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    int cur_b = -1;
    umma::named_bar_sync(3, 512);        // tables + zeroed X tiles visible to both halves

    // sums of the finished run (all tiles of shape b) -> FiLM gradients, BN_b batch terms, dW2 (registers)
    auto flush_run = [&](int b) {
      umma::fence_after_sync();
      if (part == 0) {
        if (row >= F) {                    // lanes 64..127 hold the h3-lo part of channel row - 64
#pragma unroll
          for (int br = 0; br < 2; ++br) {
            uint32_t d1[4];
            umma::tmem_ld4(T_R + lane_off + br * 32 + br * 4, d1);
            umma::tmem_ld_wait4(d1);
            s.t1buf[half][(br * 2 + 0) * F / 2 + (row - F) / 2][(row - F) & 1] = __uint_as_float(d1[0]) + __uint_as_float(d1[1]);
            s.t1buf[half][(br * 2 + 1) * F / 2 + (row - F) / 2][(row - F) & 1] = __uint_as_float(d1[2]) + __uint_as_float(d1[3]);
          }
        }
      }
      umma::named_bar_sync(1 + half, 256);
      if (part == 0 && row < F) {
        const int c = row;
#pragma unroll
        for (int br = 0; br < 2; ++br) {
          uint32_t d1[4], d2[4];
          umma::tmem_ld4(T_R + lane_off + br * 32 + br * 4, d1);
          umma::tmem_ld4(T_R + lane_off + br * 32 + 16 + br * 4, d2);
          umma::tmem_ld_wait4(d1);
          umma::tmem_ld_wait4(d2);
          const float S0 = __uint_as_float(d1[0]) + __uint_as_float(d1[1]) + s.t1buf[half][(br * 2 + 0) * F / 2 + c / 2][c & 1];
          const float S1 = __uint_as_float(d1[2]) + __uint_as_float(d1[3]) + s.t1buf[half][(br * 2 + 1) * F / 2 + c / 2][c & 1];
          const float M0 = __uint_as_float(d2[0]) + __uint_as_float(d2[1]);
          const float M1 = __uint_as_float(d2[2]) + __uint_as_float(d2[3]);
          const float2 w2 = s.lt1[br][c];
          const float* pe2 = reinterpret_cast<const float*>(&s.PE[half][br][c >> 1][2]);
          const float sraw = pe2[c & 1], shift = pe2[2 + (c & 1)];
          const float dt = fmaf(w2.x, M0, w2.y * M1);
          const float ds = (fmaf(w2.x, S0, w2.y * S1) - shift * dt) / sraw;
          atomicAdd(&a.dfilm[((size_t)(br * 2 + 0) * a.f.B + b) * F + c], ds);
          atomicAdd(&a.dfilm[((size_t)(br * 2 + 1) * a.f.B + b) * F + c], dt);
          if (a.f.training) {              // BN_b batch terms for phase 2: m1 = sum_b s*dt / M, m2 = sum_b s*ds / M
            double* rep = a.m12_rep + (size_t)(worker & (DPF_M12_REP - 1)) * (2 * F * 2) + (size_t)(br * F + c) * 2;
            DPF_GATOMIC(atomicAdd(rep + 0, (double)sraw * (double)dt));
            DPF_GATOMIC(atomicAdd(rep + 1, (double)sraw * (double)ds));
          }
          dW2acc[br][0] += S0;
          dW2acc[br][1] += S1;
        }
      }
      umma::fence_before_sync();
      umma::named_bar_sync(1 + half, 256);   // t1buf / accumulators are free again
    };

    for (int it = 0; it < n_it; ++it) {
      const int tile = t_lo + it;
      const int b = tile / a.f.tiles_per_b;
      const int n = (tile - b * a.f.tiles_per_b) * DPF_TILE + row;
      const bool valid = n < a.f.N;
      const TcRaw raw = tc_load_raw(a, b, n, valid);
      if (it > 0) {      // the previous tile's branch-1 reductions read H / D[1] / X and feed the accumulators
        umma::mbar_wait(done, ph);
        ph ^= 1;
      }
      if (b != cur_b) {
        if (cur_b >= 0) flush_run(cur_b);
        cur_b = b;
      }
      if (part == 0) {   // FiLM fold of this tile's shape
        const int br = row >> 6, c = row & 63;
        const float sc = a.f.film[((size_t)(br * 2 + 0) * a.f.B + b) * F + c];
        const float sh = a.f.film[((size_t)(br * 2 + 1) * a.f.B + b) * F + c];
        const float4 l0 = s.lt0[br][c];
        const float S = sc * l0.x;
        float* pe = reinterpret_cast<float*>(&s.PE[half][br][c >> 1][0]);
        const int ln = c & 1;
        pe[ln] = S;
        pe[2 + ln] = fmaf(-S, l0.y, sh);
        pe[8 + ln] = sc;
        pe[10 + ln] = sh;
      }
      const TcPoint g = tc_finish_point<MODE>(a, P, raw, valid);
      const float xk0 = pick3(g.x, a.f.keep0);
      const float xk1 = (K == 2) ? pick3(g.x, a.f.keep1) : 0.f;
      const f32x2 xk0_2 = f2_pack(xk0, xk0), xk1_2 = f2_pack(xk1, xk1);
      // ---- stage 0: h1 hi -> H, lo -> D (both branches), per-point weight row -> X ----
#pragma unroll
      for (int br = 0; br < 2; ++br) {
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int q = part * 4 + qq;
          uint32_t w[4], wl[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 P0 = s.PA[br][q * 4 + i][0];
            f32x2 v2 = f2_fma(f2_pack(P0.x, P0.y), xk0_2, f2_pack(P0.z, P0.w));
            if (K == 2) {
              const float4 P1 = s.PA[br][q * 4 + i][1];
              v2 = f2_fma(f2_pack(P1.x, P1.y), xk1_2, v2);
            }
            float va, vb;
            f2_unpack(v2, va, vb);
            va = fmaxf(va, 0.f);
            vb = fmaxf(vb, 0.f);
            w[i] = umma::pack_bf16(va, vb);
            if (SPLIT) wl[i] = umma::pack_bf16(va - __uint_as_float(w[i] << 16), vb - __uint_as_float(w[i] & 0xffff0000u));
          }
          const uint32_t off = umma::sw128_offset(row, q);
          *reinterpret_cast<uint4*>(hb.H + br * IMG_H + off) = make_uint4(w[0], w[1], w[2], w[3]);
          if (SPLIT) *reinterpret_cast<uint4*>(hb.D + br * IMG_H + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
      }
      if (part == 0) {   // weights of both branches: {mu0 hi, mu0 lo, mu1 hi, mu1 lo, lv0 hi, lv0 lo, lv1 hi, lv1 lo}
        uint32_t w[4];
        const float dv[4] = {g.do_mu[0], g.do_mu[1], g.do_lv[0], g.do_lv[1]};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t hi = umma::pack_bf16(dv[i], 0.f) & 0xffffu;
          const float lo = dv[i] - __uint_as_float(hi << 16);
          w[i] = umma::pack_bf16(0.f, lo) | hi;
        }
        *reinterpret_cast<uint4*>(hb.X + umma::sw128_offset(row, 0)) = make_uint4(w[0], w[1], w[2], w[3]);
        b2acc[0][0] += g.do_mu[0]; b2acc[0][1] += g.do_mu[1];
        b2acc[1][0] += g.do_lv[0]; b2acc[1][1] += g.do_lv[1];
      }
      umma::fence_async_smem();
      umma::mbar_arrive(req);
      umma::mbar_wait(done, ph);
      ph ^= 1;
      umma::fence_after_sync();
      // ---- stages 1, 2: h3 (hi | lo) and mask of one branch -> H[0] | H[1], D[1]; reductions on the tensor cores ----
#pragma unroll 1
      for (int br = 0; br < 2; ++br) {
        if (br == 1) {   // branch 0's reductions still read the tiles
          umma::mbar_wait(done, ph);
          ph ^= 1;
          umma::fence_after_sync();
        }
#pragma unroll 1
        for (int hc = 0; hc < 2; ++hc) {
          uint32_t r[16];
          umma::tmem_ld16_issue(T_F + lane_off + br * F + part * 32 + hc * 16, r);
          umma::tmem_ld_wait16(r);
          uint32_t wh[8], wl[8], wm[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 E0 = s.PE[half][br][part * 16 + hc * 8 + j][0];
            const f32x2 a2 = f2_fma(f2_pack(E0.x, E0.y), f2_pack(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), f2_pack(E0.z, E0.w));
            float ha, hbv;
            f2_unpack(a2, ha, hbv);
            ha = fmaxf(ha, 0.f);
            hbv = fmaxf(hbv, 0.f);
            wh[j] = umma::pack_bf16(ha, hbv);
            wl[j] = umma::pack_bf16(ha - __uint_as_float(wh[j] << 16), hbv - __uint_as_float(wh[j] & 0xffff0000u));
            wm[j] = (ha > 0.f ? 0x3f80u : 0u) | (hbv > 0.f ? 0x3f800000u : 0u);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t off = umma::sw128_offset(row, part * 4 + hc * 2 + q);
            *reinterpret_cast<uint4*>(hb.H + off) = make_uint4(wh[4 * q], wh[4 * q + 1], wh[4 * q + 2], wh[4 * q + 3]);
            *reinterpret_cast<uint4*>(hb.H + IMG_H + off) = make_uint4(wl[4 * q], wl[4 * q + 1], wl[4 * q + 2], wl[4 * q + 3]);
            *reinterpret_cast<uint4*>(hb.D + IMG_H + off) = make_uint4(wm[4 * q], wm[4 * q + 1], wm[4 * q + 2], wm[4 * q + 3]);
          }
        }
        umma::fence_async_smem();
        umma::fence_before_sync();
        umma::mbar_arrive(req);
      }
    }
    if (n_it > 0) {
      umma::mbar_wait(done, ph);
      ph ^= 1;
      flush_run(cur_b);
    }
    // last-SharedDot gradients of this pipeline
    if (part == 0) {
#pragma unroll
      for (int br = 0; br < 2; ++br)
#pragma unroll
        for (int wi = 0; wi < 2; ++wi) {
          const float v = warp_sum(b2acc[br][wi]);
          if (lane == 0 && n_it > 0) atomicAdd(&s.b2fin[br][wi], v);
        }
      if (row < F && n_it > 0) {
#pragma unroll
        for (int br = 0; br < 2; ++br) {
          float* d = a.dprm + (size_t)br * lay.size;
          DPF_GATOMIC(atomicAdd(&d[lay.W2 + row], dW2acc[br][0]));
          if (a.f.w == 2) DPF_GATOMIC(atomicAdd(&d[lay.W2 + F + row], dW2acc[br][1]));
        }
      }
    }
    umma::fence_before_sync();
  }
  __syncthreads();
  if (tid < 4) {
    const int br = tid >> 1, c = tid & 1;
    if (c < a.f.w) DPF_GATOMIC(atomicAdd(&a.dprm[(size_t)br * lay.size + lay.b2 + c], s.b2fin[br][c]));
  }
  // every CTA's FiLM / BN_b sums are complete and visible after this barrier
  grid_barrier(barrier_counter, gridDim.x);
  umma::fence_after_sync();

  // =========================================== phase 2 ===========================================
  // (the body of coupling_bwd_p2_tc4_kernel; the mbarrier phases continue from phase 1)
  if (issuer) {
    if (lane == 0) {
      int n_it[2] = {iters_of(2 * (int)blockIdx.x), iters_of(2 * (int)blockIdx.x + 1)};
      int p1n[2] = {p1_lo(2 * (int)blockIdx.x + 1) - p1_lo(2 * (int)blockIdx.x), p1_lo(2 * (int)blockIdx.x + 2) - p1_lo(2 * (int)blockIdx.x + 1)};
      int it[2] = {0, 0}, stage[2] = {0, 0};
      uint32_t ph[2] = {(uint32_t)(3 * p1n[0]) & 1u, (uint32_t)(3 * p1n[1]) & 1u};
      uint32_t wg_acc = 0u, bn_acc = 0u;
      int remaining = 3 * (n_it[0] + n_it[1]);
      while (remaining > 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (it[h] < n_it[h] && umma::mbar_test(&s.bar_req[h], ph[h])) {
            ph[h] ^= 1u;
            umma::fence_after_sync();
            TcP2Half& hb = s.h[h];
            const uint32_t T_F = tmem + h * 128;
            if (stage[h] == 0) {          // forward recompute, both branches
              issue_gemm1<SPLIT>(T_F, hb.H, hb.D, wimg_at<true>(s.W, 0, 0), wimg_at<true>(s.W, 0, 1));
              issue_gemm1<SPLIT>(T_F + F, hb.H + IMG_H, hb.D + IMG_H, wimg_at<true>(s.W, 1, 0), wimg_at<true>(s.W, 1, 1));
            } else if (stage[h] == 1) {   // dgrad (overwrites the recompute accumulators) + wgrad
#pragma unroll
              for (int br = 0; br < 2; ++br)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma::mma_bf16(T_F + br * F, umma::desc_at(DESC_K, umma::smem_u32(hb.D + br * IMG_H) + 32 * k),
                                 umma::desc_at(DESC_K, umma::smem_u32(wimg_at<true>(s.W, br, 2)) + 32 * k), IDESC_GEMM, k > 0);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                umma::mma_bf16(T_WG, umma::desc_at(DESC_MN, umma::smem_u32(hb.D) + 2048 * k), umma::desc_at(DESC_MN, umma::smem_u32(hb.H) + 2048 * k),
                               IDESC_WGRAD, wg_acc);
                wg_acc = 1u;
              }
            } else {                      // BN_a sums
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                umma::mma_bf16(T_BN, umma::desc_at(DESC_MN, umma::smem_u32(hb.D) + 2048 * k), umma::desc_at(DESC_MN, umma::smem_u32(hb.X) + 2048 * k),
                               IDESC_RED, bn_acc);
                bn_acc = 1u;
              }
            }
            umma::mma_commit(&s.bar_done[h]);
            if (++stage[h] == 3) { stage[h] = 0; ++it[h]; }
            --remaining;
          }
        }
      }
    }
    __syncwarp();
  } else {
    TcP2Half& hb = s.h[half];
    const int n_it = iters_of(worker);
    const int p1n = p1_lo(worker + 1) - p1_lo(worker);
    TcRaw raw0;
    {
      const int b0 = worker / a.f.tiles_per_b;
      const int n0 = (worker - b0 * a.f.tiles_per_b) * DPF_TILE + row;
      raw0 = tc_load_raw(a, b0, n0, n_it > 0 && n0 < a.f.N);
    }
    // BN_b backward of a channel: dh2pre = ib*(da*s - m1 - h2n*m2), h2n = (acc - mb)*ib  =>  dh2pre = (ib*s)*da - c2 - c3*acc
    if (tid < 128) {
      float c2 = 0.f, c3 = 0.f;
      const float4 l0 = s.lt0[tid >> 6][tid & 63];
      if (a.f.training) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int r = 0; r < DPF_M12_REP; ++r) {
          const double2 v = __ldcg(reinterpret_cast<const double2*>(a.m12_rep + (size_t)r * (2 * F * 2) + (size_t)tid * 2));
          s1 += v.x;
          s2 += v.y;
        }
        const float rM = 1.f / ((float)a.f.B * (float)a.f.N);
        const float m1 = (float)s1 * rM, m2 = (float)s2 * rM;
        c3 = l0.x * l0.x * m2;
        c2 = l0.x * m1 - c3 * l0.y;
      }
      s.lt0[tid >> 6][tid & 63] = make_float4(l0.x, l0.y, c2, c3);
    }
    // the X tiles hold phase 1's weight rows in chunk 0 of every row: phase 2 overwrites chunk 0 of every row of its tiles
    // (rows of invalid points get {0, ...}), nothing else of X was touched
    umma::named_bar_sync(3, 512);

    uint32_t ph = (uint32_t)(3 * p1n) & 1u;
    const uint32_t T_F = tmem + half * 128;
    uint64_t* req = &s.bar_req[half];
    uint64_t* done = &s.bar_done[half];
    for (int it = 0; it < n_it; ++it) {
      const int tile = worker + it * n_workers;
      const int b = tile / a.f.tiles_per_b;
      const int n = (tile - b * a.f.tiles_per_b) * DPF_TILE + row;
      const bool valid = n < a.f.N;
      if (part == 0) {   // FiLM fold of this tile's shape (read again only after the next request / completion round trip)
        const int br = row >> 6, c = row & 63;
        const float sc = a.f.film[((size_t)(br * 2 + 0) * a.f.B + b) * F + c];
        const float sh = a.f.film[((size_t)(br * 2 + 1) * a.f.B + b) * F + c];
        const float4 l0 = s.lt0[br][c];
        const float2 l1 = s.lt1[br][c];
        const float S = sc * l0.x;
        float* pe = reinterpret_cast<float*>(&s.PE[half][br][c >> 1][0]);
        const int ln = c & 1;
        pe[ln] = S;
        pe[2 + ln] = fmaf(-S, l0.y, sh);
        pe[4 + ln] = l1.x;
        pe[6 + ln] = l1.y;
        pe[8 + ln] = -l0.w;     // -c3
        pe[10 + ln] = -l0.z;    // -c2
      }
      const TcPoint g = tc_finish_point<MODE>(a, P, it == 0 ? raw0 : tc_load_raw(a, b, n, valid), valid);
      const float xk0 = pick3(g.x, a.f.keep0);
      const float xk1 = (K == 2) ? pick3(g.x, a.f.keep1) : 0.f;
      const f32x2 xk0_2 = f2_pack(xk0, xk0), xk1_2 = f2_pack(xk1, xk1);
      // ---- stage 0: h1 hi -> H, lo -> D (the previous tile's BN_a-sum UMMAs still read D / X: wait for them) ----
      if (it > 0) {
        umma::mbar_wait(done, ph);
        ph ^= 1;
      }
#pragma unroll
      for (int br = 0; br < 2; ++br) {
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int q = part * 4 + qq;
          uint32_t w[4], wl[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 P0 = s.PA[br][q * 4 + i][0];
            f32x2 v2 = f2_fma(f2_pack(P0.x, P0.y), xk0_2, f2_pack(P0.z, P0.w));
            if (K == 2) {
              const float4 P1 = s.PA[br][q * 4 + i][1];
              v2 = f2_fma(f2_pack(P1.x, P1.y), xk1_2, v2);
            }
            float va, vb;
            f2_unpack(v2, va, vb);
            va = fmaxf(va, 0.f);
            vb = fmaxf(vb, 0.f);
            w[i] = umma::pack_bf16(va, vb);
            if (SPLIT) wl[i] = umma::pack_bf16(va - __uint_as_float(w[i] << 16), vb - __uint_as_float(w[i] & 0xffff0000u));
          }
          const uint32_t off = umma::sw128_offset(row, q);
          *reinterpret_cast<uint4*>(hb.H + br * IMG_H + off) = make_uint4(w[0], w[1], w[2], w[3]);
          if (SPLIT) *reinterpret_cast<uint4*>(hb.D + br * IMG_H + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
      }
      umma::fence_async_smem();
      umma::mbar_arrive(req);
      umma::mbar_wait(done, ph);
      ph ^= 1;
      umma::fence_after_sync();
      // ---- stage 1: epilogue A: dh2pre (bf16) of this part's 32 channels of each branch -> D tiles ----
#pragma unroll 1
      for (int br = 0; br < 2; ++br) {
        const float d0 = br == 0 ? g.do_mu[0] : g.do_lv[0];
        const float d1 = br == 0 ? g.do_mu[1] : g.do_lv[1];
        const f32x2 d0_2 = f2_pack(d0, d0), d1_2 = f2_pack(d1, d1);
#pragma unroll 1
        for (int hc = 0; hc < 2; ++hc) {
          uint32_t r[16];
          umma::tmem_ld16_issue(T_F + lane_off + br * F + part * 32 + hc * 16, r);
          umma::tmem_ld_wait16(r);
          uint32_t wv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4* pe = s.PE[half][br][part * 16 + hc * 8 + j];
            const float4 E0 = pe[0], E1 = pe[1], E2 = pe[2];
            const f32x2 S2 = f2_pack(E0.x, E0.y);
            const f32x2 v2 = f2_pack(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            const f32x2 av2 = f2_fma(S2, v2, f2_pack(E0.z, E0.w));
            const f32x2 g2 = f2_fma(f2_pack(E1.x, E1.y), d0_2, f2_mul(f2_pack(E1.z, E1.w), d1_2));
            float av0, av1, g0, g1;
            f2_unpack(av2, av0, av1);
            f2_unpack(g2, g0, g1);
            const f32x2 da2 = f2_pack(av0 > 0.f ? g0 : 0.f, av1 > 0.f ? g1 : 0.f);
            const f32x2 dh2 = f2_fma(S2, da2, f2_fma(f2_pack(E2.x, E2.y), v2, f2_pack(E2.z, E2.w)));
            float h0, h1;
            f2_unpack(dh2, h0, h1);
            wv[j] = valid ? umma::pack_bf16(h0, h1) : 0u;
          }
#pragma unroll
          for (int q = 0; q < 2; ++q)
            *reinterpret_cast<uint4*>(hb.D + br * IMG_H + umma::sw128_offset(row, part * 4 + hc * 2 + q)) =
                make_uint4(wv[4 * q], wv[4 * q + 1], wv[4 * q + 2], wv[4 * q + 3]);
        }
      }
      umma::fence_async_smem();
      umma::fence_before_sync();
      umma::mbar_arrive(req);
      umma::mbar_wait(done, ph);
      ph ^= 1;
      umma::fence_after_sync();
      // ---- stage 2: epilogue B: dz (bf16) -> D tiles, T1 = A0^T dz, per-point weight row -> X ----
      f32x2 T1_0_2 = f2_pack(0.f, 0.f), T1_1_2 = f2_pack(0.f, 0.f);     // even / odd channel partial sums
#pragma unroll 1
      for (int br = 0; br < 2; ++br) {
#pragma unroll 1
        for (int hc = 0; hc < 2; ++hc) {
          uint32_t r[16];
          umma::tmem_ld16_issue(T_F + lane_off + br * F + part * 32 + hc * 16, r);
          umma::tmem_ld_wait16(r);
          uint32_t wv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4* pa = s.PA[br][part * 16 + hc * 8 + j];
            const float4 P0 = pa[0];
            const f32x2 A00_2 = f2_pack(P0.x, P0.y);
            f32x2 z2 = f2_fma(A00_2, xk0_2, f2_pack(P0.z, P0.w));
            f32x2 A01_2 = 0;
            if (K == 2) {
              const float4 P1 = pa[1];
              A01_2 = f2_pack(P1.x, P1.y);
              z2 = f2_fma(A01_2, xk1_2, z2);
            }
            float z0, z1;
            f2_unpack(z2, z0, z1);
            const float dz0 = (z0 > 0.f && valid) ? __uint_as_float(r[2 * j]) : 0.f;
            const float dz1 = (z1 > 0.f && valid) ? __uint_as_float(r[2 * j + 1]) : 0.f;
            const f32x2 dz2 = f2_pack(dz0, dz1);
            T1_0_2 = f2_fma(A00_2, dz2, T1_0_2);
            if (K == 2) T1_1_2 = f2_fma(A01_2, dz2, T1_1_2);
            wv[j] = umma::pack_bf16(dz0, dz1);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q)
            *reinterpret_cast<uint4*>(hb.D + br * IMG_H + umma::sw128_offset(row, part * 4 + hc * 2 + q)) =
                make_uint4(wv[4 * q], wv[4 * q + 1], wv[4 * q + 2], wv[4 * q + 3]);
        }
      }
      float T1_0, T1_1;
      {
        float lo, hi;
        f2_unpack(T1_0_2, lo, hi);
        T1_0 = lo + hi;
        f2_unpack(T1_1_2, lo, hi);
        T1_1 = lo + hi;
      }
      if (part == 0) {
        const float one = valid ? 1.f : 0.f;
        const uint32_t w0 = umma::pack_bf16(one, xk0);                                   // {1, xk0 hi}
        const float xk0_lo = xk0 - __uint_as_float(w0 & 0xffff0000u);
        const uint32_t w1 = umma::pack_bf16(xk0_lo, xk1);                                // {xk0 lo, xk1 hi}
        const float xk1_lo = xk1 - __uint_as_float(w1 & 0xffff0000u);
        const uint32_t w2 = umma::pack_bf16(xk1_lo, 0.f);                                // {xk1 lo, 0}
        *reinterpret_cast<uint4*>(hb.X + umma::sw128_offset(row, 0)) = make_uint4(w0, w1, w2, 0u);
      } else {
        s.t1buf[half][row][0] = T1_0;
        s.t1buf[half][row][1] = T1_1;
      }
      umma::fence_async_smem();
      umma::fence_before_sync();
      umma::mbar_arrive(req);
      umma::named_bar_sync(1 + half, 256);   // t1buf hand-over inside the half
      if (part == 0 && valid) {
        T1_0 += s.t1buf[half][row][0];
        T1_1 += s.t1buf[half][row][1];
        float dx[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) dx[ch] = (MODE == 1) ? g.dy[ch] / sig1 : g.dy[ch] * sig1;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          if (ch == a.f.keep0) dx[ch] += T1_0;
          if (K == 2 && ch == a.f.keep1) dx[ch] += T1_1;
          if (ch == a.f.warp0) dx[ch] = (MODE == 1) ? g.dy[ch] / g.sig[0] : g.dy[ch] * g.sig[0];
          if (K == 1 && ch == a.f.warp1) dx[ch] = (MODE == 1) ? g.dy[ch] / g.sig[1] : g.dy[ch] * g.sig[1];
        }
        const size_t base = (size_t)b * 3 * a.f.N + n;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) a.dx_out[base + (size_t)ch * a.f.N] = dx[ch];
      }
    }
    if (n_it > 0) {     // the last tile's BN_a-sum UMMAs
      umma::mbar_wait(done, ph);
      ph ^= 1;
    }
    umma::fence_before_sync();
  }
  // ---- CTA epilogue (every UMMA of both halves has completed) ----
  __syncthreads();
  umma::fence_after_sync();
  const bool had_work = 2 * (int)blockIdx.x < a.f.n_tiles;
  if (tid < 128 && had_work) {     // lane = channel m = br*64 + c; columns {dbeta, E0 hi, E0 lo, E1 hi, E1 lo}
    uint32_t r0[4], r1[4];
    umma::tmem_ld4(T_BN + ((uint32_t)(quarter * 32) << 16), r0);
    umma::tmem_ld4(T_BN + ((uint32_t)(quarter * 32) << 16) + 4, r1);
    umma::tmem_ld_wait4(r0);
    umma::tmem_ld_wait4(r1);
    DPF_GATOMIC(atomicAdd(&a.bna_sums[tid * 4 + 0], (double)__uint_as_float(r0[0])));
    DPF_GATOMIC(atomicAdd(&a.bna_sums[tid * 4 + 1], (double)__uint_as_float(r0[1]) + (double)__uint_as_float(r0[2])));
    DPF_GATOMIC(atomicAdd(&a.bna_sums[tid * 4 + 2], (double)__uint_as_float(r0[3]) + (double)__uint_as_float(r1[0])));
  }
  if (!issuer) {
    const int br = row >> 6, c = row & 63, cg = half * 2 + part;
    float4* d = reinterpret_cast<float4*>(a.dw1_partial + ((size_t)blockIdx.x * 2 + br) * (F * F) + c * F + cg * 16);
    uint32_t r[16];
    if (had_work) {
      umma::tmem_ld16_issue(T_WG + ((uint32_t)(quarter * 32) << 16) + br * F + cg * 16, r);
      umma::tmem_ld_wait16(r);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = 0u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      d[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// =============================================================================================
// Backward pass 1, two threads per point, reductions over points on the tensor cores.
//
// With mask m = [a > 0], a = s*h2n + t (FiLM scale s, shift t), h3 = m*a, da = m*(W2_0 d0 + W2_1 d1):
//     S_w[c] = sum_p d_w[p] h3[p,c]      M_w[c] = sum_p d_w[p] m[p,c]          (w = 0, 1)
//     dW2_w = S_w     dt = W2_0 M_0 + W2_1 M_1     ds = sum_p da*h2n = (W2_0 S_0 + W2_1 S_1 - t*dt) / s
// so the four column sums per branch are all pass 1 needs.  They are UMMAs with K = the 128 points:
//     D[channel, j] += sum_p A[p, channel] * X[p, j]
// A = the h3 (bf16 hi | lo) tiles / the 0-1 mask tile read as MN-major M = 128 operands, X = the
// per-point weights {d0 hi, d0 lo, d1 hi, d1 lo} of both branches (MN-major, N = 16), accumulated in
// TMEM over the consecutive tiles of one shape and read back one channel per lane.
// =============================================================================================
struct TcP1Smem2 {
  unsigned char W[4 * IMG_W];           // [br][W1 hi, W1 lo]
  unsigned char H[2 * IMG_H];           // h1 hi | lo, then h3 hi | lo of the branch in flight
  unsigned char Mk[IMG_H];              // mask tile; the M = 128 operand's second block is X (ignored lanes)
  unsigned char X[IMG_H];               // per-point weights, chunk 0 of every row (chunk 1 stays zero)
  TcCommon c;
  float shift[2][F];                    // FiLM shift of the current shape
  float sbuf[2][2][DPF_TILE];           // S_w partials: lanes 0..63 = h3 hi part, 64..127 = h3 lo part
  float mbuf[2][2][F];
  float b2fin[2][2];
};

template <int K, int MODE, bool SPLIT>
__global__ void __launch_bounds__(NT2, 2)
coupling_bwd_p1_tc2_kernel(const BwdArgs a, const unsigned short* __restrict__ wimg) {
  extern __shared__ unsigned char smraw[];
  TcP1Smem2& s = *reinterpret_cast<TcP1Smem2*>(smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, part = tid >> 7, quarter = warp & 3;
  const BranchLayout lay = branch_layout(a.f.k, a.f.w, a.f.G);
  int t0, t1;
  tile_range(a.f.n_tiles, t0, t1);
  DPF_STAMP(1, 0);
  DPF_STAMP_NS(1, 14);
  pdl_launch_dependents();
  const uint32_t tmem = tc_setup(s.c, 256);
  DPF_STAMP(1, 1);
  tc_load_weights<false>(s.c, s.W, wimg);
  pdl_wait();          // the previous backward kernel's sums / gradients are read from here on
  if (a.ltab) tc_prologue_tables_ltab(a.ltab, s.c);
  else tc_prologue_tables(a.f, lay, s.c, false, true);
  DPF_STAMP(1, 2);
  for (int i = tid; i < (int)(IMG_H / 16); i += NT2) reinterpret_cast<uint4*>(s.X)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 4) (&s.b2fin[0][0])[tid] = 0.f;
  __syncthreads();
  // the first tile's global loads are issued before the (latency-bound) pending computation
  TcRaw raw0;
  {
    const int b0 = t0 / a.f.tiles_per_b;
    const int n0 = (t0 - b0 * a.f.tiles_per_b) * DPF_TILE + row;
    raw0 = tc_load_raw(a, b0, n0, t0 < t1 && n0 < a.f.N);
  }
  // The deferred correction P needs the previous backward kernel's BN_a sums (a latency-bound double-precision
  // chain); the first tile's h1 tiles and forward UMMA need only x and the tables, so P is computed while that
  // UMMA is in flight (inside the tile loop, first tile / first branch).
  Pending P{0.f, 0.f, 0.f, 0.f, 0.f};
  DPF_STAMP(1, 3);
  float b2acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float dW2acc[2] = {0.f, 0.f};        // threads < 128: (branch, channel) = (tid >> 6, tid & 63)
  umma::mbar_wait(&s.c.bar_load, 0);
  __syncthreads();
  DPF_STAMP(1, 4);

  uint32_t phase = 0, phase_aux = 0;
  const uint32_t T_FWD = tmem, T_RED = tmem + 128;   // T_RED + br*32: [0,16) h3 sums, [16,32) mask sums
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
  int cur_b = -1;
  bool run_start = true;

  // sums of the finished run of tiles (all of shape b) -> FiLM gradients (global) and dW2 (registers)
  auto flush_run = [&](int b) {
    if (part == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int br = 0; br < 2; ++br) {
        uint32_t d1[4], d2[4];
        umma::tmem_ld4(T_RED + lane_off + br * 32 + br * 4, d1);
        umma::tmem_ld4(T_RED + lane_off + br * 32 + 16 + br * 4, d2);
        umma::tmem_ld_wait4(d1);
        umma::tmem_ld_wait4(d2);
        s.sbuf[br][0][row] = __uint_as_float(d1[0]) + __uint_as_float(d1[1]);
        s.sbuf[br][1][row] = __uint_as_float(d1[2]) + __uint_as_float(d1[3]);
        if (row < F) {
          s.mbuf[br][0][row] = __uint_as_float(d2[0]) + __uint_as_float(d2[1]);
          s.mbuf[br][1][row] = __uint_as_float(d2[2]) + __uint_as_float(d2[3]);
        }
      }
      umma::fence_before_sync();
    }
    __syncthreads();
    if (tid < 2 * F) {
      const int br = tid >> 6, c = tid & 63;
      const float S0 = s.sbuf[br][0][c] + s.sbuf[br][0][F + c], S1 = s.sbuf[br][1][c] + s.sbuf[br][1][F + c];
      const float M0 = s.mbuf[br][0][c], M1 = s.mbuf[br][1][c];
      const float W20 = s.c.W2[br][0][c], W21 = s.c.W2[br][1][c];
      const float dt = fmaf(W20, M0, W21 * M1);
      const float ds = (fmaf(W20, S0, W21 * S1) - s.shift[br][c] * dt) / s.c.sraw[br][c];
      atomicAdd(&a.dfilm[((size_t)(br * 2 + 0) * a.f.B + b) * F + c], ds);
      atomicAdd(&a.dfilm[((size_t)(br * 2 + 1) * a.f.B + b) * F + c], dt);
      if (a.m12_rep && a.f.training) {   // BN_b batch terms for pass 2: m1 = sum_b s*dt / M, m2 = sum_b s*ds / M
        double* rep = a.m12_rep + (size_t)(blockIdx.x & (DPF_M12_REP - 1)) * (2 * F * 2) + (size_t)tid * 2;
        const double sc = (double)s.c.sraw[br][c];
        DPF_GATOMIC(atomicAdd(rep + 0, sc * (double)dt));
        DPF_GATOMIC(atomicAdd(rep + 1, sc * (double)ds));
      }
      dW2acc[0] += S0;
      dW2acc[1] += S1;
    }
    __syncthreads();
  };

  for (int tile = t0; tile < t1; ++tile) {
    const int b = tile / a.f.tiles_per_b;
    const int n = (tile - b * a.f.tiles_per_b) * DPF_TILE + row;
    const bool valid = n < a.f.N;
    if (tile > t0) {   // the previous tile's reduction UMMAs read H / Mk / X and feed the accumulators
      umma::mbar_wait(&s.c.bar_aux, phase_aux);
      phase_aux ^= 1;
    }
    if (b != cur_b) {
      if (cur_b >= 0) flush_run(cur_b);
      cur_b = b;
      run_start = true;
      tc_tile_film(a.f, s.c, b);
      if (tid < 2 * F) s.shift[tid >> 6][tid & 63] = a.f.film[((size_t)((tid >> 6) * 2 + 1) * a.f.B + b) * F + (tid & 63)];
    }
    const TcRaw raw = tile == t0 ? raw0 : tc_load_raw(a, b, n, valid);
    if (tile == t0) DPF_STAMP(1, 5);
    const float xk0 = pick3(raw.x, a.f.keep0);
    const float xk1 = (K == 2) ? pick3(raw.x, a.f.keep1) : 0.f;
    const f32x2 xk0_2 = f2_pack(xk0, xk0), xk1_2 = f2_pack(xk1, xk1);
    TcPoint g;
#pragma unroll 1
    for (int br = 0; br < 2; ++br) {
      if (br == 1) {   // branch 0's reductions still read the h3 tiles in H
        umma::mbar_wait(&s.c.bar_aux, phase_aux);
        phase_aux ^= 1;
      }
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const int q = part * 4 + qq;
        uint32_t w[4], wl[4];
        h1_chunk8<K, SPLIT>(s.c.A0p[br], q, xk0_2, xk1_2, w, wl);
        const uint32_t off = umma::sw128_offset(row, q);
        *reinterpret_cast<uint4*>(s.H + off) = make_uint4(w[0], w[1], w[2], w[3]);
        if (SPLIT) *reinterpret_cast<uint4*>(s.H + IMG_H + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
      }
      umma::fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        umma::fence_after_sync();
        issue_gemm1<SPLIT>(T_FWD + br * F, s.H, s.H + IMG_H, wimg_at<false>(s.W, br, 0), wimg_at<false>(s.W, br, 1));
        umma::mma_commit(&s.c.bar_mma);
      }
      if (br == 0) {   // while the forward UMMA runs: pending correction (once), this point's cotangents, weight row
        if (tile == t0) {
          P = tc_compute_pending(a, blockIdx.x == 0, s.c.pend);
          if (blockIdx.x == 0 && tid == 0 && a.pend_store) {   // pass 2 of this layer reads it instead of recomputing
            a.pend_store[0] = P.c0; a.pend_store[1] = P.c1; a.pend_store[2] = P.q00; a.pend_store[3] = P.q01; a.pend_store[4] = P.q11;
          }
        }
        g = tc_finish_point<MODE>(a, P, raw, valid);
        if (part == 0) {   // weights of both branches: {mu0 hi, mu0 lo, mu1 hi, mu1 lo, lv0 hi, lv0 lo, lv1 hi, lv1 lo}
          uint32_t w[4];
          const float dv[4] = {g.do_mu[0], g.do_mu[1], g.do_lv[0], g.do_lv[1]};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t hi = umma::pack_bf16(dv[i], 0.f) & 0xffffu;
            const float lo = dv[i] - __uint_as_float(hi << 16);
            w[i] = umma::pack_bf16(0.f, lo) | hi;
          }
          *reinterpret_cast<uint4*>(s.X + umma::sw128_offset(row, 0)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      umma::mbar_wait(&s.c.bar_mma, phase);
      phase ^= 1;
      umma::fence_after_sync();
      if (tile == t0 && br == 0) DPF_STAMP(1, 6);
      // h3 (hi | lo) and mask of this part's 32 channels -> tiles
      {
        float v[32];
        umma::tmem_ld32(T_FWD + lane_off + br * F + part * 32, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t wh[4], wl[4], wm[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 E0 = s.c.epip[br][part * 16 + q * 4 + i][0];
            const f32x2 a2 = f2_fma(f2_pack(E0.x, E0.y), f2_pack(v[q * 8 + 2 * i], v[q * 8 + 2 * i + 1]), f2_pack(E0.z, E0.w));
            float ha, hb;
            f2_unpack(a2, ha, hb);
            ha = fmaxf(ha, 0.f);
            hb = fmaxf(hb, 0.f);
            wh[i] = umma::pack_bf16(ha, hb);
            wl[i] = umma::pack_bf16(ha - __uint_as_float(wh[i] << 16), hb - __uint_as_float(wh[i] & 0xffff0000u));
            wm[i] = (ha > 0.f ? 0x3f80u : 0u) | (hb > 0.f ? 0x3f800000u : 0u);
          }
          const uint32_t off = umma::sw128_offset(row, part * 4 + q);
          *reinterpret_cast<uint4*>(s.H + off) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
          *reinterpret_cast<uint4*>(s.H + IMG_H + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
          *reinterpret_cast<uint4*>(s.Mk + off) = make_uint4(wm[0], wm[1], wm[2], wm[3]);
        }
      }
      umma::fence_async_smem();
      umma::fence_before_sync();
      __syncthreads();
      if (tile == t0 && br == 0) DPF_STAMP(1, 7);
      if (tid == 0) {
        umma::fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma::mma_bf16(T_RED + br * 32, umma::desc_at(DESC_MN, umma::smem_u32(s.H) + 2048 * k),
                         umma::desc_at(DESC_MN, umma::smem_u32(s.X) + 2048 * k), IDESC_RED, (!run_start || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma::mma_bf16(T_RED + br * 32 + 16, umma::desc_at(DESC_MN, umma::smem_u32(s.Mk) + 2048 * k),
                         umma::desc_at(DESC_MN, umma::smem_u32(s.X) + 2048 * k), IDESC_RED, (!run_start || k > 0) ? 1u : 0u);
        umma::mma_commit(&s.c.bar_aux);
      }
    }
    run_start = false;
    if (tile == t0) DPF_STAMP(1, 8);
    b2acc[0][0] += g.do_mu[0]; b2acc[0][1] += g.do_mu[1];
    b2acc[1][0] += g.do_lv[0]; b2acc[1][1] += g.do_lv[1];
  }
  if (t1 > t0) {
    umma::mbar_wait(&s.c.bar_aux, phase_aux);
    DPF_STAMP(1, 9);
    flush_run(cur_b);
  }
  DPF_STAMP(1, 10);
  if (part == 0) {
#pragma unroll
    for (int br = 0; br < 2; ++br)
#pragma unroll
      for (int wi = 0; wi < 2; ++wi) {
        const float v = warp_sum(b2acc[br][wi]);
        if (lane == 0) atomicAdd(&s.b2fin[br][wi], v);
      }
  }
  __syncthreads();
  if (tid < 2 * F) {
    const int br = tid >> 6, c = tid & 63;
    float* d = a.dprm + (size_t)br * lay.size;
    DPF_GATOMIC(atomicAdd(&d[lay.W2 + c], dW2acc[0]));
    if (a.f.w == 2) DPF_GATOMIC(atomicAdd(&d[lay.W2 + F + c], dW2acc[1]));
    if (c < a.f.w) DPF_GATOMIC(atomicAdd(&d[lay.b2 + c], s.b2fin[br][c]));
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
  DPF_STAMP(1, 11);
  DPF_STAMP_NS(1, 15);
}

// =============================================================================================
// Forward kernels, two threads per point (256 threads per CTA, same shared memory as the 128-thread
// form => 16 warps per SM): statistics pass, apply pass, and the merged train-mode launch.
// =============================================================================================
struct TcFwdSmem2 {
  unsigned char W[4 * IMG_W];           // [br][W1 hi, W1 lo]
  unsigned char H[2 * IMG_H];           // hi, lo tile of the branch in flight
  TcCommon c;
  float scratch[2 * DPF_TILE * 33];     // column-sum scratch, one plane per part
  float fin[2][2 * F];
  float obuf[DPF_TILE][4];              // part 1 -> part 0 hand-over of the partial last-SharedDot sums
  double mom[9][4];
};

// sum and sum of squares over the 128 rows for this part's 32 columns
__device__ __forceinline__ void colreduce32_sq_part(float* scratch, const float va[32], int row, int part, int lane, int quarter,
                                                    float& rs, float& rq) {
  float* sc = scratch + part * (DPF_TILE * 33);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; ++i) sc[row * 33 + i] = va[i];
  __syncthreads();
  float sa = 0.f, sb = 0.f;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    const float v = sc[(quarter * 32 + r) * 33 + lane];
    sa += v;
    sb = fmaf(v, v, sb);
  }
  rs = sa;
  rq = sb;
}

// h1 chunks of this part for one branch -> tiles, then UMMA chain into TMEM columns [tcol, tcol+64)
template <int K, bool SPLIT>
__device__ __forceinline__ void fwd2_gemm(TcFwdSmem2& s, int br, uint32_t tcol, float xk0, float xk1, int row, int part, uint32_t& phase) {
  const f32x2 xk0_2 = f2_pack(xk0, xk0), xk1_2 = f2_pack(xk1, xk1);
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) {
    const int q = part * 4 + qq;
    uint32_t w[4], wl[4];
    h1_chunk8<K, SPLIT>(s.c.A0p[br], q, xk0_2, xk1_2, w, wl);
    const uint32_t off = umma::sw128_offset(row, q);
    *reinterpret_cast<uint4*>(s.H + off) = make_uint4(w[0], w[1], w[2], w[3]);
    if (SPLIT) *reinterpret_cast<uint4*>(s.H + IMG_H + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
  }
  umma::fence_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    umma::fence_after_sync();
    issue_gemm1<SPLIT>(tcol, s.H, s.H + IMG_H, wimg_at<false>(s.W, br, 0), wimg_at<false>(s.W, br, 1));
    umma::mma_commit(&s.c.bar_mma);
  }
  umma::mbar_wait(&s.c.bar_mma, phase);
  phase ^= 1;
  umma::fence_after_sync();
}

// partial last-SharedDot sums of this part's 32 channels of one branch from TMEM columns [tcol + 32*part, +32):
// packed fp32x2 math over channel pairs (even / odd channels accumulate in the two lanes, summed at the end)
__device__ __forceinline__ void fwd2_epilogue(const TcFwdSmem2& s, int br, uint32_t taddr, int part, float& o0, float& o1) {
  float v[32];
  umma::tmem_ld32(taddr + part * 32, v);
  f32x2 o0_2 = f2_pack(o0, 0.f), o1_2 = f2_pack(o1, 0.f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 E0 = s.c.epip[br][part * 16 + j][0], E1 = s.c.epip[br][part * 16 + j][1];
    const f32x2 a2 = f2_fma(f2_pack(E0.x, E0.y), f2_pack(v[2 * j], v[2 * j + 1]), f2_pack(E0.z, E0.w));
    float ha, hb;
    f2_unpack(a2, ha, hb);
    const f32x2 h2 = f2_pack(fmaxf(ha, 0.f), fmaxf(hb, 0.f));
    o0_2 = f2_fma(f2_pack(E1.x, E1.y), h2, o0_2);
    o1_2 = f2_fma(f2_pack(E1.z, E1.w), h2, o1_2);
  }
  float lo, hi;
  f2_unpack(o0_2, lo, hi);
  o0 = lo + hi;
  f2_unpack(o1_2, lo, hi);
  o1 = lo + hi;
}

// transform + outputs + moment accumulation of one point (part 0 threads)
template <int MODE>
__device__ __forceinline__ void fwd2_finish(const CouplingArgs& a, const float xin[3], const float o[2][2], int b, int n, bool valid,
                                            float macc[9]) {
  float yv[3], muv[3] = {0.f, 0.f, 0.f}, lvv[3] = {0.f, 0.f, 0.f};
  const float sig1 = sqrtf(a.eps + 1.0f);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) yv[ch] = (MODE == 0) ? sig1 * xin[ch] : xin[ch] / sig1;
#pragma unroll
  for (int wi = 0; wi < 2; ++wi) {
    if (wi < a.w) {
      const int ch = wi == 0 ? a.warp0 : a.warp1;
      const float l = softsign(o[1][wi]);
      const float sig = sqrtf(a.eps + expf(l));
      const float m = o[0][wi];
      const float xv = pick3(xin, ch);
      const float r = (MODE == 0) ? fmaf(sig, xv, m) : (xv - m) / sig;
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (q == ch) { yv[q] = r; muv[q] = m; lvv[q] = l; }
    }
  }
  if (valid) {
    store_point_outputs(a, (size_t)b * 3 * a.N + n, yv, muv, lvv);
    macc[0] += yv[0]; macc[1] += yv[1]; macc[2] += yv[2];
    macc[3] = fmaf(yv[0], yv[0], macc[3]); macc[4] = fmaf(yv[0], yv[1], macc[4]); macc[5] = fmaf(yv[0], yv[2], macc[5]);
    macc[6] = fmaf(yv[1], yv[1], macc[6]); macc[7] = fmaf(yv[1], yv[2], macc[7]); macc[8] = fmaf(yv[2], yv[2], macc[8]);
  }
}

// apply epilogue of one tile from TMEM columns [tbase, tbase+128): both branches, pair hand-over, finish
template <int MODE>
__device__ __forceinline__ void fwd2_apply_tile(const CouplingArgs& a, TcFwdSmem2& s, uint32_t tbase_lane, const float xin[3], int b,
                                                int n, bool valid, int row, int part, float macc[9]) {
  float o[2][2];
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    o[br][0] = part == 0 ? s.c.b2[br][0] : 0.f;
    o[br][1] = part == 0 ? s.c.b2[br][1] : 0.f;
    fwd2_epilogue(s, br, tbase_lane + br * F, part, o[br][0], o[br][1]);
  }
  if (part == 1) *reinterpret_cast<float4*>(s.obuf[row]) = make_float4(o[0][0], o[0][1], o[1][0], o[1][1]);
  __syncthreads();
  if (part == 0) {
    const float4 t = *reinterpret_cast<const float4*>(s.obuf[row]);
    o[0][0] += t.x; o[0][1] += t.y; o[1][0] += t.z; o[1][1] += t.w;
    fwd2_finish<MODE>(a, xin, o, b, n, valid, macc);
  }
}

__device__ __forceinline__ void fwd2_flush_stats(const CouplingArgs& a, TcFwdSmem2& s, const float sacc[2][2], int tid, int part, int lane) {
  float* fin = &s.fin[0][0];
  for (int i = tid; i < 4 * F; i += NT2) fin[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    atomicAdd(&s.fin[0][br * F + part * 32 + lane], sacc[br][0]);
    atomicAdd(&s.fin[1][br * F + part * 32 + lane], sacc[br][1]);
  }
  __syncthreads();
  if (tid < 128) {
    // merged forward: DPF_BNB_REP replicas, so that ~300 CTAs do not serialise on 256 addresses right before the grid barrier
    double* dst = a.bnb_rep ? a.bnb_rep + (size_t)(blockIdx.x & (DPF_BNB_REP - 1)) * (2 * F * 2) : a.bnb_sums;
    DPF_GATOMIC(atomicAdd(&dst[tid * 2 + 0], (double)s.fin[0][tid]));   // tid == br*F + c
    DPF_GATOMIC(atomicAdd(&dst[tid * 2 + 1], (double)s.fin[1][tid]));
  }
}

__device__ __forceinline__ void fwd2_flush_moments(const CouplingArgs& a, TcFwdSmem2& s, const float macc[9], int tid) {
  // only part-0 threads (warps 0..3) carry moments
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const double v = warp_sum_d((double)macc[i]);
    if (lane == 0 && warp < 4) s.mom[i][warp] = v;
  }
  __syncthreads();
  if (tid < 9) DPF_GATOMIC(atomicAdd(a.mom_out + tid, s.mom[tid][0] + s.mom[tid][1] + s.mom[tid][2] + s.mom[tid][3]));
}

template <int K, int MODE, bool STATS, bool SPLIT>
__global__ void __launch_bounds__(NT2, 2)
coupling_fwd_tc2_kernel(const CouplingArgs a, const unsigned short* __restrict__ wimg) {
  extern __shared__ unsigned char smraw[];
  TcFwdSmem2& s = *reinterpret_cast<TcFwdSmem2*>(smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, part = tid >> 7, quarter = warp & 3;
  const BranchLayout lay = branch_layout(a.k, a.w, a.G);
  const bool writer = (blockIdx.x == 0) && a.update_stats && !STATS;
  const uint32_t tmem = tc_setup(s.c, 128);
  tc_load_weights<false>(s.c, s.W, wimg);
  tc_prologue_tables(a, lay, s.c, writer, !STATS);
  float sacc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float macc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) macc[i] = 0.f;
  umma::mbar_wait(&s.c.bar_load, 0);
  __syncthreads();

  int t0, t1;
  tile_range(a.n_tiles, t0, t1);
  uint32_t phase = 0;
  const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
  for (int tile = t0; tile < t1; ++tile) {
    const int b = tile / a.tiles_per_b;
    const int n = (tile - b * a.tiles_per_b) * DPF_TILE + row;
    const bool valid = n < a.N;
    if (!STATS) tc_tile_film(a, s.c, b);
    const float* px = a.x + (size_t)b * 3 * a.N + n;
    float xin[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) xin[ch] = valid ? px[(size_t)ch * a.N] : 0.f;
    const float xk0 = pick3(xin, a.keep0);
    const float xk1 = (K == 2) ? pick3(xin, a.keep1) : 0.f;
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      fwd2_gemm<K, SPLIT>(s, br, tmem + br * F, xk0, xk1, row, part, phase);
      if (STATS) {
        float v[32];
        umma::tmem_ld32(lane_addr + br * F + part * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = valid ? v[i] : 0.f;
        float rs, rq;
        colreduce32_sq_part(s.scratch, v, row, part, lane, quarter, rs, rq);
        sacc[br][0] += rs;
        sacc[br][1] += rq;
        umma::fence_before_sync();
        __syncthreads();
      } else if (br == 0) {
        // branch 1's tiles overwrite H only after this sync; branch 0's accumulator stays in its TMEM columns
        umma::fence_before_sync();
        __syncthreads();
      }
    }
    if (!STATS) {
      fwd2_apply_tile<MODE>(a, s, lane_addr, xin, b, n, valid, row, part, macc);
      umma::fence_before_sync();
      __syncthreads();     // TMEM columns, tiles, epi table and obuf are free again
    }
  }
  if (STATS) {
    fwd2_flush_stats(a, s, sacc, tid, part, lane);
  } else if (a.mom_out) {
    fwd2_flush_moments(a, s, macc, tid);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

template <int K, int MODE, bool SPLIT>
__global__ void __launch_bounds__(NT2, 2)
coupling_fwd_train_tc2_kernel(const CouplingArgs a, const unsigned short* __restrict__ wimg, unsigned int* __restrict__ barrier_counter) {
  extern __shared__ unsigned char smraw[];
  TcFwdSmem2& s = *reinterpret_cast<TcFwdSmem2*>(smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, part = tid >> 7, quarter = warp & 3;
  const BranchLayout lay = branch_layout(a.k, a.w, a.G);
  const bool writer = (blockIdx.x == 0) && a.update_stats && tid < 128;
  DPF_STAMP(0, 0);
  DPF_STAMP_NS(0, 14);
  pdl_launch_dependents();
  const uint32_t tmem = tc_setup(s.c, RES * 128);
  DPF_STAMP(0, 1);
  tc_load_weights<false>(s.c, s.W, wimg);
  pdl_wait();          // everything above is independent of the previous layer's kernel
  tc_prologue_tables(a, lay, s.c, writer, false);
  DPF_STAMP(0, 2);
  if (tid < 128) {
    const int br = tid >> 6, c = tid & 63;
    const float* prm = a.prm + (size_t)br * lay.size;
    s.c.W2[br][0][c] = prm[lay.W2 + c];
    s.c.W2[br][1][c] = (a.w == 2) ? prm[lay.W2 + F + c] : 0.f;
    if (c < 2) s.c.b2[br][c] = (c < a.w) ? prm[lay.b2 + c] : 0.f;
  }
  float sacc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  umma::mbar_wait(&s.c.bar_load, 0);
  __syncthreads();
  DPF_STAMP(0, 3);

  const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
  uint32_t phase = 0;
  float xin[RES][3];
  // ---------------- phase 1: h1 -> UMMA -> per-channel sum / sum of squares ----------------
#pragma unroll
  for (int ts = 0; ts < RES; ++ts) {
    const int tile = blockIdx.x + ts * gridDim.x;
    if (tile < a.n_tiles) {
      const int b = tile / a.tiles_per_b;
      const int n = (tile - b * a.tiles_per_b) * DPF_TILE + row;
      const bool valid = n < a.N;
      const float* px = a.x + (size_t)b * 3 * a.N + n;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) xin[ts][ch] = valid ? px[(size_t)ch * a.N] : 0.f;
      const float xk0 = pick3(xin[ts], a.keep0);
      const float xk1 = (K == 2) ? pick3(xin[ts], a.keep1) : 0.f;
#pragma unroll
      for (int br = 0; br < 2; ++br) {
        fwd2_gemm<K, SPLIT>(s, br, tmem + ts * 128 + br * F, xk0, xk1, row, part, phase);
        float v[32];
        umma::tmem_ld32(lane_addr + ts * 128 + br * F + part * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = valid ? v[i] : 0.f;
        float rs, rq;
        colreduce32_sq_part(s.scratch, v, row, part, lane, quarter, rs, rq);
        sacc[br][0] += rs;
        sacc[br][1] += rq;
        umma::fence_before_sync();
        __syncthreads();
      }
    }
  }
  DPF_STAMP(0, 4);
  fwd2_flush_stats(a, s, sacc, tid, part, lane);
  DPF_STAMP(0, 5);
  grid_barrier(barrier_counter, gridDim.x);
  DPF_STAMP(0, 6);
  // ---------------- phase 2: BN_b x FiLM fold, epilogue from the resident accumulators ----------------
  if (tid < 128) {
    const int br = tid >> 6, c = tid & 63;
    const double M = (double)a.B * (double)a.N;
    double sm = 0.0, sq = 0.0;
#pragma unroll
    for (int r = 0; r < DPF_BNB_REP; ++r) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(a.bnb_rep + (size_t)r * (2 * F * 2) + (br * F + c) * 2));
      sm += v.x;
      sq += v.y;
    }
    if (blockIdx.x == 0) {   // canonical copy for the backward kernels (bn_b_stats)
      a.bnb_sums[(br * F + c) * 2 + 0] = sm;
      a.bnb_sums[(br * F + c) * 2 + 1] = sq;
    }
    const double dm = sm / M;
    const double dv = fmax(sq / M - dm * dm, 0.0);
    s.c.mb[br][c] = (float)dm;
    s.c.ib[br][c] = 1.f / sqrtf((float)dv + DPF_BN_EPS);
    if (writer) {
      float* st = a.stat + (size_t)br * ST_COUNT * F;
      st[ST_BNB_RM * F + c] = (1.f - DPF_BN_MOM) * st[ST_BNB_RM * F + c] + DPF_BN_MOM * (float)dm;
      st[ST_BNB_RV * F + c] = (1.f - DPF_BN_MOM) * st[ST_BNB_RV * F + c] + DPF_BN_MOM * (float)(dv * (M / fmax(M - 1.0, 1.0)));
    }
  }
  __syncthreads();
  float macc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) macc[i] = 0.f;
  DPF_STAMP(0, 7);
  umma::fence_after_sync();
#pragma unroll
  for (int ts = 0; ts < RES; ++ts) {
    const int tile = blockIdx.x + ts * gridDim.x;
    if (tile < a.n_tiles) {
      const int b = tile / a.tiles_per_b;
      const int n = (tile - b * a.tiles_per_b) * DPF_TILE + row;
      const bool valid = n < a.N;
      __syncthreads();
      tc_tile_film(a, s.c, b);
      __syncthreads();
      fwd2_apply_tile<MODE>(a, s, lane_addr + ts * 128, xin[ts], b, n, valid, row, part, macc);
    }
  }
  DPF_STAMP(0, 8);
  if (a.mom_out) fwd2_flush_moments(a, s, macc, tid);
  DPF_STAMP(0, 9);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, RES * 128);
  DPF_STAMP(0, 10);
  DPF_STAMP_NS(0, 15);
}

template <typename T>
inline size_t smem_for() { return sizeof(T) + 1024; }

template <int K, int MODE, bool STATS, bool SPLIT>
int launch_fwd_tc_t(const CouplingArgs& a, const unsigned short* wimg, int grid, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(coupling_fwd_tc2_kernel<K, MODE, STATS, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for<TcFwdSmem2>());
    attr = true;
  }
  coupling_fwd_tc2_kernel<K, MODE, STATS, SPLIT><<<grid, NT2, smem_for<TcFwdSmem2>(), st>>>(a, wimg);
  return dpf_check_launch("coupling_fwd_tc2_kernel");
}

template <int K, int MODE, bool SPLIT>
int launch_bwd_tc_t(const BwdArgs& a, const unsigned short* wimg, int pass, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(coupling_bwd_p1_tc2_kernel<K, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for<TcP1Smem2>());
    cudaFuncSetAttribute(coupling_bwd_p2_tc2_kernel<K, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for<TcP2Smem2>());
    attr = true;
  }
  if (pass == 1) {
    const int grid = min(a.f.n_tiles, dpf_num_sms() * 2);
    dpf_launch_pdl(coupling_bwd_p1_tc2_kernel<K, MODE, SPLIT>, grid, NT2, smem_for<TcP1Smem2>(), st, a, wimg);
    return dpf_check_launch("coupling_bwd_p1_tc2_kernel");
  }
  const int grid = min(a.f.n_tiles, dpf_num_sms());
  if (g_dpf_p2_two_tiles && a.pend_store && a.m12_rep && a.ltab) {   // two tiles in flight per SM (dpf_set_option(3, 0) = one)
    static bool attr4 = false;
    if (!attr4) {
      cudaFuncSetAttribute(coupling_bwd_p2_tc4_kernel<K, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for<TcP2Smem4>());
      attr4 = true;
    }
    dpf_launch_pdl(coupling_bwd_p2_tc4_kernel<K, MODE, SPLIT>, grid, NT4, smem_for<TcP2Smem4>(), st, a, wimg);
    return dpf_check_launch("coupling_bwd_p2_tc4_kernel");
  }
  dpf_launch_pdl(coupling_bwd_p2_tc2_kernel<K, MODE, SPLIT>, grid, NT2, smem_for<TcP2Smem2>(), st, a, wimg);
  return dpf_check_launch("coupling_bwd_p2_tc2_kernel");
}

static int g_coop_occupancy = -1;
static int* g_barrier_fail_host = nullptr;     // pinned + mapped; written by a timed-out grid barrier
static int g_coresident_ok[2] = {-1, -1};     // per kernel footprint (0 merged forward, 1 merged backward): -1 not probed, 0 failed, 1 verified

static int tc_barrier_setup() {
  if (g_barrier_fail_host) return DPF_OK;
  int* h = nullptr;
  cudaError_t e = cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped);
  if (e != cudaSuccess) { dpf_set_error("grid barrier flag: %s", cudaGetErrorString(e)); return (int)e; }
  *h = 0;
  int* d = nullptr;
  e = cudaHostGetDevicePointer(&d, h, 0);
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_barrier_fail_dev, &d, sizeof(d));
  if (e != cudaSuccess) { cudaFreeHost(h); dpf_set_error("grid barrier flag: %s", cudaGetErrorString(e)); return (int)e; }
  g_barrier_fail_host = h;
  return DPF_OK;
}

// One synchronous probe launch per process and footprint: `grid` CTAs of `threads` threads with `smem` dynamic bytes
// and `tmem_cols` TMEM columns each must all be resident at once (the probe uses fewer registers than any kernel it
// stands for, whose own __launch_bounds__ guarantee the register fit).  Cannot run during stream capture (-1).
static int tc_verify_coresidency(int which, int grid, int threads, size_t smem, int tmem_cols, cudaStream_t st) {
  if (g_coresident_ok[which] >= 0) return g_coresident_ok[which];
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  if (cap != cudaStreamCaptureStatusNone) return -1;
  if (tc_barrier_setup() != DPF_OK) return -1;
  unsigned int* counter = nullptr;
  if (cudaMalloc(&counter, 2 * sizeof(unsigned int)) != cudaSuccess) return -1;
  cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned int), st);
  cudaFuncSetAttribute(coresidency_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(coresidency_probe_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  coresidency_probe_kernel<<<grid, threads, smem, st>>>(counter, tmem_cols);
  ++g_dpf_launches;
  unsigned int host[2] = {0u, 1u};
  cudaError_t e = cudaMemcpyAsync(host, counter, sizeof(host), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(counter);
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  g_coresident_ok[which] = (host[0] == (unsigned int)grid && host[1] == 0u) ? 1 : 0;
  return g_coresident_ok[which];
}

// merged backward launch of one layer (pass 1 -> grid barrier -> pass 2); DPF_ERR_UNSUPPORTED = use the two launches
template <int K, int MODE, bool SPLIT>
int launch_bwd_merged_t(const BwdArgs& a, const unsigned short* wimg, unsigned int* counter, cudaStream_t st) {
  auto kern = coupling_bwd_merged_kernel<K, MODE, SPLIT>;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for<TcBwdMSmem>());
    attr = true;
  }
  const int ok = tc_verify_coresidency(1, dpf_num_sms(), NT4, smem_for<TcBwdMSmem>(), 512, st);
  if (ok != 1) {
    dpf_set_error("merged backward not used: co-residency of %d CTAs %s", dpf_num_sms(), ok == 0 ? "could not be established" : "not verified yet (stream capture)");
    return DPF_ERR_UNSUPPORTED;
  }
  const int grid = min(a.f.n_tiles, dpf_num_sms());      // = the CTA count dpf_decoder_backward sized the wgrad partials for
  dpf_launch_pdl(kern, grid, NT4, smem_for<TcBwdMSmem>(), st, a, wimg, counter);
  return dpf_check_launch("coupling_bwd_merged_kernel");
}

template <int K, int MODE, bool SPLIT>
int launch_fwd_train_t(const CouplingArgs& a, const unsigned short* wimg, unsigned int* counter, cudaStream_t st) {
  static int max_grid = -1;
  auto kern = coupling_fwd_train_tc2_kernel<K, MODE, SPLIT>;
  if (max_grid < 0) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for<TcFwdSmem2>());
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT2, smem_for<TcFwdSmem2>());
    g_coop_occupancy = per_sm;
    // shared memory allows 2 CTAs per SM (2 x ~92 KB) and so do the 512 TMEM columns (RES*128 each);
    // the cooperative launch itself validates co-residency and we fall back if it refuses
    per_sm = 512 / (RES * 128);
    max_grid = per_sm * dpf_num_sms();
  }
  if (max_grid <= 0 || a.n_tiles > RES * max_grid) {
    dpf_set_error("merged forward not used: n_tiles=%d max_grid=%d smem=%zu", a.n_tiles, max_grid, smem_for<TcFwdSmem2>());
    return DPF_ERR_UNSUPPORTED;
  }
  // Not cudaLaunchCooperativeKernel: the runtime's co-residency check assumes ONE CTA per SM for any
  // kernel that allocates TMEM (occupancy query returns 1 at every shared-memory size), although two
  // CTAs with 256 columns each do share an SM.  Instead the full-size grid's co-residency is verified
  // once per process with a probe launch of the same footprint; until / unless that succeeds the caller
  // uses the two-launch form (no grid barrier).
  const int ok = tc_verify_coresidency(0, max_grid, NT2, smem_for<TcFwdSmem2>(), RES * 128, st);
  if (ok != 1) {
    dpf_set_error("merged forward not used: co-residency of %d CTAs %s", max_grid, ok == 0 ? "could not be established" : "not verified yet (stream capture)");
    return DPF_ERR_UNSUPPORTED;
  }
  const int grid = min(a.n_tiles, max_grid);
  dpf_launch_pdl(kern, grid, NT2, smem_for<TcFwdSmem2>(), st, a, wimg, counter);
  return dpf_check_launch("coupling_fwd_train_tc_kernel");
}

template <int K, bool SPLIT>
int launch_fwd_tc_k(const CouplingArgs& a, const unsigned short* wimg, int mode, bool stats_pass, int grid, cudaStream_t s) {
  if (mode == 0) return stats_pass ? launch_fwd_tc_t<K, 0, true, SPLIT>(a, wimg, grid, s) : launch_fwd_tc_t<K, 0, false, SPLIT>(a, wimg, grid, s);
  return stats_pass ? launch_fwd_tc_t<K, 1, true, SPLIT>(a, wimg, grid, s) : launch_fwd_tc_t<K, 1, false, SPLIT>(a, wimg, grid, s);
}

}  // namespace

// Folded per-(branch, channel) tables of every layer for the backward kernels, once per pass:
// {A00, A01, c0} (BN_a folded into the first SharedDot), BN_b mean / istd, last SharedDot weights and bias.
__global__ void __launch_bounds__(2 * DPF_F)
bwd_tables_kernel(const float* __restrict__ arena, float* stats, const LayerMeta* __restrict__ meta, const double* __restrict__ moments,
                  double* bnb_sums, float* __restrict__ ltab, int L, int G, int B, int N, int mode, int training) {
  const int l = blockIdx.x, q = mode == 0 ? l : L - 1 - l;
  const LayerMeta m = meta[l];
  CouplingArgs a{};
  a.prm = arena + m.param_off;
  a.stat = stats + m.stat_off;
  a.mom_in = moments + (size_t)q * 16;
  a.bnb_sums = bnb_sums + (size_t)l * 2 * DPF_F * 2;
  a.B = B; a.N = N; a.G = G;
  a.k = (int)m.k; a.w = (int)m.w; a.keep0 = (int)m.keep0; a.keep1 = (int)m.keep1;
  a.training = training;
  const BranchLayout lay = branch_layout(a.k, a.w, G);
  const int br = threadIdx.x >> 6, c = threadIdx.x & 63;
  float A00, A01, c0, mean, istd, meanA, istdA;
  fold_bn_a(a, lay, br, c, false, A00, A01, c0, &meanA, &istdA);
  bn_b_stats(a, br, c, false, mean, istd);
  const float* prm = a.prm + (size_t)br * lay.size;
  const float w20 = prm[lay.W2 + c];
  const float w21 = (a.w == 2) ? prm[lay.W2 + DPF_F + c] : 0.f;
  const float b2 = (c < a.w) ? prm[lay.b2 + c] : 0.f;
  float4* d = reinterpret_cast<float4*>(ltab + ((size_t)l * 2 * DPF_F + threadIdx.x) * DPF_LTAB_ROW);
  d[0] = make_float4(A00, A01, c0, mean);
  d[1] = make_float4(istd, w20, w21, b2);
  d[2] = make_float4(prm[lay.W0 + c * a.k], (a.k == 2) ? prm[lay.W0 + c * a.k + 1] : 0.f, prm[lay.bnA_w + c], meanA);   // = bn_a_of()
  d[3] = make_float4(istdA, 0.f, 0.f, 0.f);
}

int launch_bwd_tables(const float* arena, float* stats, const LayerMeta* meta_dev, const double* moments, const double* bnb_sums,
                      float* ltab, int L, int G, int B, int N, int mode, int training, cudaStream_t s) {
  bwd_tables_kernel<<<L, 2 * DPF_F, 0, s>>>(arena, stats, meta_dev, moments, const_cast<double*>(bnb_sums), ltab, L, G, B, N, mode, training);
  return dpf_check_launch("bwd_tables_kernel");
}

// dW1 of every layer = sum over the pass-2 CTAs' partials (deterministic, no global atomics).
// grid = (L*2, 16): blockIdx.x = layer*2 + branch, each thread owns one of the 64x64 entries.
__global__ void __launch_bounds__(256)
dw1_reduce_kernel(const float* __restrict__ partial, int n_cta, float* __restrict__ darena, const LayerMeta* __restrict__ meta, int G) {
  const int l = blockIdx.x >> 1, br = blockIdx.x & 1;
  const int e = blockIdx.y * 256 + threadIdx.x;
  const LayerMeta m = meta[l];
  const BranchLayout lay = branch_layout((int)m.k, (int)m.w, G);
  const float* p = partial + ((size_t)l * n_cta * 2 + br) * (DPF_F * DPF_F) + e;
  float s0 = 0.f, s1 = 0.f;
  int c = 0;
  for (; c + 2 <= n_cta; c += 2) {
    s0 += p[(size_t)c * 2 * DPF_F * DPF_F];
    s1 += p[(size_t)(c + 1) * 2 * DPF_F * DPF_F];
  }
  if (c < n_cta) s0 += p[(size_t)c * 2 * DPF_F * DPF_F];
  darena[m.param_off + (size_t)br * lay.size + lay.W1 + e] = s0 + s1;
}

int launch_dw1_reduce(const float* partial, int n_cta, float* darena, const LayerMeta* meta_dev, int L, int G, cudaStream_t s) {
  dw1_reduce_kernel<<<dim3(L * 2, 16), 256, 0, s>>>(partial, n_cta, darena, meta_dev, G);
  return dpf_check_launch("dw1_reduce_kernel");
}

int tc_bwd_p2_max_ctas() { return dpf_num_sms(); }

// 1 when a grid barrier of this process has timed out (the pass that hit it trapped; outputs are invalid)
int tc_barrier_failed() { return g_barrier_fail_host && *reinterpret_cast<volatile int*>(g_barrier_fail_host) != 0; }
int tc_coresidency_state() { return g_coresident_ok[0] < g_coresident_ok[1] ? g_coresident_ok[0] : g_coresident_ok[1]; }

size_t tc_weight_image_elems_per_layer() { return (size_t)2 * N_IMG * F * F; }

int launch_pack_w1(const float* arena, const LayerMeta* meta_dev, int L, int G, unsigned short* out, cudaStream_t s) {
  pack_w1_kernel<<<L * 2, 256, 0, s>>>(arena, meta_dev, G, out);
  return dpf_check_launch("pack_w1_kernel");
}

// wimg: this layer's images [br][W1 hi, W1 lo, W1^T hi][4096 bf16]; split != 0 = bf16x3
int launch_coupling_fwd_tc(const CouplingArgs& a, const unsigned short* wimg, int mode, bool stats_pass, int split, cudaStream_t s) {
  const int grid = min(a.n_tiles, dpf_num_sms() * 2);
  if (a.k == 2) return split ? launch_fwd_tc_k<2, true>(a, wimg, mode, stats_pass, grid, s) : launch_fwd_tc_k<2, false>(a, wimg, mode, stats_pass, grid, s);
  return split ? launch_fwd_tc_k<1, true>(a, wimg, mode, stats_pass, grid, s) : launch_fwd_tc_k<1, false>(a, wimg, mode, stats_pass, grid, s);
}

// merged statistics + apply launch of one train-mode layer (DPF_ERR_UNSUPPORTED = does not fit)
int launch_coupling_fwd_train_tc(const CouplingArgs& a, const unsigned short* wimg, int mode, int split, unsigned int* counter,
                                 cudaStream_t s) {
  if (a.k == 2) {
    if (split) return mode == 0 ? launch_fwd_train_t<2, 0, true>(a, wimg, counter, s) : launch_fwd_train_t<2, 1, true>(a, wimg, counter, s);
    return mode == 0 ? launch_fwd_train_t<2, 0, false>(a, wimg, counter, s) : launch_fwd_train_t<2, 1, false>(a, wimg, counter, s);
  }
  if (split) return mode == 0 ? launch_fwd_train_t<1, 0, true>(a, wimg, counter, s) : launch_fwd_train_t<1, 1, true>(a, wimg, counter, s);
  return mode == 0 ? launch_fwd_train_t<1, 0, false>(a, wimg, counter, s) : launch_fwd_train_t<1, 1, false>(a, wimg, counter, s);
}

// occupancy (CTAs per SM) the runtime computes for the merged / the plain apply kernel at `smem` dynamic bytes
int tc_debug_occupancy(int which, int smem) {
  int n = -1;
  if (which == 0) {
    auto k = coupling_fwd_train_tc2_kernel<1, 1, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, NT2, smem);
  } else {
    auto k = coupling_fwd_tc2_kernel<1, 1, false, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, NT2, smem);
  }
  return n;
}

int launch_coupling_bwd_tc(const BwdArgs& a, const unsigned short* wimg, int mode, int pass, int split, cudaStream_t s) {
  if (a.f.k == 2) {
    if (split) return mode == 0 ? launch_bwd_tc_t<2, 0, true>(a, wimg, pass, s) : launch_bwd_tc_t<2, 1, true>(a, wimg, pass, s);
    return mode == 0 ? launch_bwd_tc_t<2, 0, false>(a, wimg, pass, s) : launch_bwd_tc_t<2, 1, false>(a, wimg, pass, s);
  }
  if (split) return mode == 0 ? launch_bwd_tc_t<1, 0, true>(a, wimg, pass, s) : launch_bwd_tc_t<1, 1, true>(a, wimg, pass, s);
  return mode == 0 ? launch_bwd_tc_t<1, 0, false>(a, wimg, pass, s) : launch_bwd_tc_t<1, 1, false>(a, wimg, pass, s);
}

int launch_coupling_bwd_merged_tc(const BwdArgs& a, const unsigned short* wimg, int mode, int split, unsigned int* counter, cudaStream_t s) {
  if (!(a.ltab && a.m12_rep)) return DPF_ERR_UNSUPPORTED;
  if (a.f.k == 2) {
    if (split) return mode == 0 ? launch_bwd_merged_t<2, 0, true>(a, wimg, counter, s) : launch_bwd_merged_t<2, 1, true>(a, wimg, counter, s);
    return mode == 0 ? launch_bwd_merged_t<2, 0, false>(a, wimg, counter, s) : launch_bwd_merged_t<2, 1, false>(a, wimg, counter, s);
  }
  if (split) return mode == 0 ? launch_bwd_merged_t<1, 0, true>(a, wimg, counter, s) : launch_bwd_merged_t<1, 1, true>(a, wimg, counter, s);
  return mode == 0 ? launch_bwd_merged_t<1, 0, false>(a, wimg, counter, s) : launch_bwd_merged_t<1, 1, false>(a, wimg, counter, s);
}

#ifdef DPF_STAMPS
extern "C" __attribute__((visibility("default"))) int dpf_debug_stamps(unsigned long long* buf) {
  return (int)cudaMemcpyToSymbol(g_stamps, &buf, sizeof(buf));
}
#endif
