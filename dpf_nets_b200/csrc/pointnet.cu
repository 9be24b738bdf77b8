// Fused eval-mode PointNet cloud encoder + max-pool: x (B,3,N) -> global feature (B,512) in ONE kernel.
// Reference: PointNetCloudEncoder.forward (lib/networks/encoders.py:9-28: 4 x [SharedDot(no bias) ->
// BatchNorm1d -> ReLU], widths 3 -> 64 -> 128 -> 256 -> 512) followed by torch.max over the points
// (lib/networks/models.py:130-131).  SURVEY.md section 8 row a8 / b2 "pointnet_fwd" (inference half;
// the train-mode path - batch statistics, backward - still runs on the library GEMMs).
//
// In eval mode every BatchNorm is an affine map of running statistics and folds into the preceding
// SharedDot: a_l = relu(W'_l a_{l-1} + b'_l).  A tile of 128 points runs through all four layers
// without leaving the SM:
//   layer 1 (K = 3)         CUDA cores                      -> bf16 tile A1 [128 x 64]  (smem, SW128)
//   layer 2  z2 = A1 W2'^T  tcgen05 M=128 N=128 K=64        -> TMEM -> relu -> A2 [128 x 128]
//   layer 3  z3 = A2 W3'^T  tcgen05 M=128 N=256 K=128       -> TMEM -> relu -> A3 [128 x 256]
//   layer 4  z4^T = W4' A3^T, four chunks of 128 channels:  tcgen05 M=128 (channels) N=128 (points)
//            K=256 -> TMEM with ONE CHANNEL PER LANE, so the max over the points of the tile is a
//            per-thread register reduction (no cross-thread traffic); relu and the bias commute with max.
// Weights are pre-folded / pre-swizzled bf16 images (pointnet_pack_kernel) staged by TMA bulk copies;
// W4 (256 KB) streams through two 64 KB buffers.  One CTA per (shape, slice of its tiles); slices are
// combined with an atomic max on the non-negative outputs.
#include "coupling.cuh"
#include "umma.cuh"

namespace {

constexpr int PN_T = 256;                       // threads per CTA: row = tid & 127, part = tid >> 7
constexpr int C1 = 64, C2 = 128, C3 = 256, C4 = 512;
constexpr uint32_t KB_TILE = 128 * 128;         // 16 KB: [128 rows x 64 bf16] SW128 K-block
constexpr uint64_t PN_DESC = umma::make_desc_template(16, 1024, umma::LAYOUT_SW128);
constexpr uint32_t IDESC_N128 = umma::make_idesc_bf16(128, 128, 0, 0);
constexpr uint32_t IDESC_N256 = umma::make_idesc_bf16(128, 256, 0, 0);

// workspace layout (bytes)
constexpr size_t WS_TAB1 = 0;                                   // float4[64] {w0', w1', w2', b'}
constexpr size_t WS_B2 = WS_TAB1 + C1 * 16;                     // float[128]
constexpr size_t WS_B3 = WS_B2 + C2 * 4;                        // float[256]
constexpr size_t WS_B4 = WS_B3 + C3 * 4;                        // float[512]
constexpr size_t WS_W2 = 8192;                                  // [128 x 64] bf16 image (16 KB); tables end at 4608
constexpr size_t WS_W3 = WS_W2 + (size_t)C2 * C1 * 2;           // 2 K-blocks x [256 x 64] (64 KB)
constexpr size_t WS_W4 = WS_W3 + (size_t)C3 * C2 * 2;           // 4 chunks x 4 K-blocks x [128 x 64] (256 KB)
constexpr size_t WS_BYTES = WS_W4 + (size_t)C4 * C3 * 2;
static_assert(WS_B4 + C4 * 4 <= WS_W2, "pointnet workspace tables overlap the weight images");

struct PnParams {
  const float* W[4];        // SharedDot weights (out, in) row-major
  const float* gamma[4];
  const float* beta[4];
  const float* rm[4];
  const float* rv[4];
  float bn_eps;
};

__device__ __forceinline__ float pn_scale(const PnParams& p, int l, int c) { return p.gamma[l][c] / sqrtf(p.rv[l][c] + p.bn_eps); }

// folded tables + swizzled bf16 weight images; one thread per 16-byte chunk (8 input channels of one output row)
__global__ void __launch_bounds__(256)
pointnet_pack_kernel(const PnParams p, unsigned char* __restrict__ ws) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < C1) {
    const float g = pn_scale(p, 0, i);
    reinterpret_cast<float4*>(ws + WS_TAB1)[i] =
        make_float4(g * p.W[0][i * 3 + 0], g * p.W[0][i * 3 + 1], g * p.W[0][i * 3 + 2], p.beta[0][i] - g * p.rm[0][i]);
  }
  if (i < C2) reinterpret_cast<float*>(ws + WS_B2)[i] = p.beta[1][i] - pn_scale(p, 1, i) * p.rm[1][i];
  if (i < C3) reinterpret_cast<float*>(ws + WS_B3)[i] = p.beta[2][i] - pn_scale(p, 2, i) * p.rm[2][i];
  if (i < C4) reinterpret_cast<float*>(ws + WS_B4)[i] = p.beta[3][i] - pn_scale(p, 3, i) * p.rm[3][i];
  int l, r, q, kin;
  unsigned char* dst;
  if (i < C2 * (C1 / 8)) {                               // W2: rows 128, 8 chunks
    l = 1; kin = C1; r = i / 8; q = i % 8;
    dst = ws + WS_W2 + umma::sw128_offset(r, q);
  } else if (i < C2 * (C1 / 8) + C3 * (C2 / 8)) {        // W3: rows 256, 16 chunks = 2 K-blocks
    const int j = i - C2 * (C1 / 8);
    l = 2; kin = C2; r = j / 16; q = j % 16;
    dst = ws + WS_W3 + (size_t)(q >> 3) * (C3 * 128) + umma::sw128_offset(r, q & 7);
  } else if (i < C2 * (C1 / 8) + C3 * (C2 / 8) + C4 * (C3 / 8)) {   // W4: rows 512 (4 chunks of 128), 32 chunks = 4 K-blocks
    const int j = i - C2 * (C1 / 8) - C3 * (C2 / 8);
    l = 3; kin = C3; r = j / 32; q = j % 32;
    dst = ws + WS_W4 + (size_t)(r >> 7) * (4 * KB_TILE) + (size_t)(q >> 3) * KB_TILE + umma::sw128_offset(r & 127, q & 7);
  } else {
    return;
  }
  const float g = pn_scale(p, l, r);
  const float* src = p.W[l] + (size_t)r * kin + q * 8;
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) w[e] = umma::pack_bf16(g * src[2 * e], g * src[2 * e + 1]);
  *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

struct PnSmem {
  unsigned char buf0[4 * KB_TILE];      // A1 (16 KB) + A2 (2 K-blocks), then W4 chunks 0 / 2
  unsigned char buf1[4 * KB_TILE];      // W3 (2 x 32 KB), then W4 chunks 1 / 3
  unsigned char buf2[4 * KB_TILE];      // A3 (4 K-blocks)
  unsigned char w2[KB_TILE];            // resident
  float4 tab1[C1];
  float b2[C2], b3[C3], b4[C4];
  float red[2][C4];
  uint64_t bar_mma, bar_w2, bar_w3, bar_c[2];
  uint32_t tmem_base;
};

// relu(acc + bias) of 32 accumulator columns -> four 16-byte chunks of row `row` of a K-block
__device__ __forceinline__ void pn_store32(unsigned char* kblock, int row, int chunk0, const float v[32], const float* bias) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      w[e] = umma::pack_bf16(fmaxf(v[q * 8 + 2 * e] + bias[q * 8 + 2 * e], 0.f), fmaxf(v[q * 8 + 2 * e + 1] + bias[q * 8 + 2 * e + 1], 0.f));
    *reinterpret_cast<uint4*>(kblock + umma::sw128_offset(row, chunk0 + q)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void __launch_bounds__(PN_T, 1)
pointnet_eval_kernel(const float* __restrict__ x, const unsigned char* __restrict__ ws, float* __restrict__ out, int B, int N) {
  extern __shared__ unsigned char smraw[];
  PnSmem& s = *reinterpret_cast<PnSmem*>(smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, part = tid >> 7, quarter = warp & 3;
  const int b = blockIdx.y;
  const int tiles_per_b = (N + DPF_TILE - 1) / DPF_TILE;
  const int per = (tiles_per_b + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per, t_end = min(tiles_per_b, t_begin + per);
  if (t_begin >= t_end) return;

  if (tid == 0) {
    umma::mbar_init(&s.bar_mma, 1);
    umma::mbar_init(&s.bar_w2, 1);
    umma::mbar_init(&s.bar_w3, 1);
    umma::mbar_init(&s.bar_c[0], 1);
    umma::mbar_init(&s.bar_c[1], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(&s.tmem_base, 512);
  for (int i = tid; i < C1; i += PN_T) s.tab1[i] = reinterpret_cast<const float4*>(ws + WS_TAB1)[i];
  for (int i = tid; i < C2; i += PN_T) s.b2[i] = reinterpret_cast<const float*>(ws + WS_B2)[i];
  for (int i = tid; i < C3; i += PN_T) s.b3[i] = reinterpret_cast<const float*>(ws + WS_B3)[i];
  for (int i = tid; i < C4; i += PN_T) s.b4[i] = reinterpret_cast<const float*>(ws + WS_B4)[i];
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  const uint32_t T_Z2 = tmem, T_Z3 = tmem + 128, T_Z4[2] = {tmem, tmem + 384};
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
  if (tid == 0) {
    umma::mbar_expect_tx(&s.bar_w2, KB_TILE);
    umma::bulk_g2s(s.w2, ws + WS_W2, KB_TILE, &s.bar_w2);
  }
  uint32_t ph_mma = 0, ph_w3 = 0, ph_c[2] = {0, 0};
  float runmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  unsigned char* const A1 = s.buf0;
  unsigned char* const A2 = s.buf0 + KB_TILE;
  umma::mbar_wait(&s.bar_w2, 0);

  for (int tile = t_begin; tile < t_end; ++tile) {
    const int n = tile * DPF_TILE + row;
    const bool valid = n < N;
    const int nvalid = min(DPF_TILE, N - tile * DPF_TILE);
    if (tid == 0) {   // W3 -> buf1 (free: the previous tile's chunk UMMAs have completed)
      umma::mbar_expect_tx(&s.bar_w3, 4 * KB_TILE);
      umma::bulk_g2s(s.buf1, ws + WS_W3, 4 * KB_TILE, &s.bar_w3);
    }
    // ---- layer 1 on the CUDA cores: this part's 32 channels of its point -> A1 ----
    {
      const size_t base = (size_t)b * 3 * N + n;
      const float x0 = valid ? x[base] : 0.f, x1 = valid ? x[base + N] : 0.f, x2 = valid ? x[base + 2 * (size_t)N] : 0.f;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const int q = part * 4 + qq;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 ta = s.tab1[q * 8 + 2 * e], tb = s.tab1[q * 8 + 2 * e + 1];
          const float va = fmaxf(fmaf(ta.z, x2, fmaf(ta.y, x1, fmaf(ta.x, x0, ta.w))), 0.f);
          const float vb = fmaxf(fmaf(tb.z, x2, fmaf(tb.y, x1, fmaf(tb.x, x0, tb.w))), 0.f);
          w[e] = umma::pack_bf16(va, vb);
        }
        *reinterpret_cast<uint4*>(A1 + umma::sw128_offset(row, q)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    // ---- layer 2: z2[128 x 128] = A1 W2'^T ----
    if (tid == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma::mma_bf16(T_Z2, umma::desc_at(PN_DESC, umma::smem_u32(A1) + 32 * k), umma::desc_at(PN_DESC, umma::smem_u32(s.w2) + 32 * k),
                       IDESC_N128, k > 0);
      umma::mma_commit(&s.bar_mma);
    }
    umma::mbar_wait(&s.bar_mma, ph_mma);
    ph_mma ^= 1;
    umma::fence_after_sync();
#pragma unroll
    for (int j = 0; j < 2; ++j) {   // this part's 64 channels = K-block `part` of A2
      float v[32];
      umma::tmem_ld32(T_Z2 + lane_off + part * 64 + j * 32, v);
      pn_store32(A2 + part * KB_TILE, row, j * 4, v, &s.b2[part * 64 + j * 32]);
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    // ---- layer 3: z3[128 x 256] = A2 W3'^T ----
    umma::mbar_wait(&s.bar_w3, ph_w3);
    ph_w3 ^= 1;
    if (tid == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma::mma_bf16(T_Z3, umma::desc_at(PN_DESC, umma::smem_u32(A2) + kb * KB_TILE + 32 * k),
                         umma::desc_at(PN_DESC, umma::smem_u32(s.buf1) + kb * 2 * KB_TILE + 32 * k), IDESC_N256, (kb | k) > 0);
      umma::mma_commit(&s.bar_mma);
    }
    umma::mbar_wait(&s.bar_mma, ph_mma);
    ph_mma ^= 1;
    umma::fence_after_sync();
    if (tid == 0) {   // A1 / A2 / W3 are consumed: stream the first two W4 chunks
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        umma::mbar_expect_tx(&s.bar_c[c], 4 * KB_TILE);
        umma::bulk_g2s(c == 0 ? s.buf0 : s.buf1, ws + WS_W4 + (size_t)c * 4 * KB_TILE, 4 * KB_TILE, &s.bar_c[c]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // this part's 128 channels = K-blocks 2*part, 2*part+1 of A3
      float v[32];
      umma::tmem_ld32(T_Z3 + lane_off + part * 128 + j * 32, v);
      pn_store32(s.buf2 + (part * 2 + (j >> 1)) * KB_TILE, row, (j & 1) * 4, v, &s.b3[part * 128 + j * 32]);
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    // ---- layer 4, channel-major: z4^T[128 channels x 128 points] per chunk = W4'(chunk) A3^T ----
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int bsel = c & 1;
      unsigned char* wbuf = bsel == 0 ? s.buf0 : s.buf1;
      umma::mbar_wait(&s.bar_c[bsel], ph_c[bsel]);
      ph_c[bsel] ^= 1;
      if (tid == 0) {
        umma::fence_after_sync();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma::mma_bf16(T_Z4[bsel], umma::desc_at(PN_DESC, umma::smem_u32(wbuf) + kb * KB_TILE + 32 * k),
                           umma::desc_at(PN_DESC, umma::smem_u32(s.buf2) + kb * KB_TILE + 32 * k), IDESC_N128, (kb | k) > 0);
        umma::mma_commit(&s.bar_mma);
      }
      umma::mbar_wait(&s.bar_mma, ph_mma);
      ph_mma ^= 1;
      umma::fence_after_sync();
      if (tid == 0 && c + 2 < 4) {   // this chunk's weights are consumed: refill the buffer
        umma::mbar_expect_tx(&s.bar_c[bsel], 4 * KB_TILE);
        umma::bulk_g2s(wbuf, ws + WS_W4 + (size_t)(c + 2) * 4 * KB_TILE, 4 * KB_TILE, &s.bar_c[bsel]);
      }
      float m = runmax[c];
#pragma unroll
      for (int j = 0; j < 2; ++j) {   // lane = channel; this part's 64 points
        float v[32];
        umma::tmem_ld32(T_Z4[bsel] + lane_off + part * 64 + j * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) m = fmaxf(m, (part * 64 + j * 32 + i < nvalid) ? v[i] : -INFINITY);
      }
      runmax[c] = m;
      umma::fence_before_sync();
      __syncthreads();   // every thread has read this accumulator buffer before chunk c+2 (or the next tile) overwrites it
    }
  }
  // ---- max over the CTA's tiles: combine the two point halves, bias + relu, atomic max across CTAs ----
#pragma unroll
  for (int c = 0; c < 4; ++c) s.red[part][c * 128 + row] = runmax[c];
  __syncthreads();
  for (int ch = tid; ch < C4; ch += PN_T) {
    const float v = fmaxf(fmaxf(s.red[0][ch], s.red[1][ch]) + s.b4[ch], 0.f) + 0.f;   // (+0.f: never -0.0, whose bits would win the uint max)
    atomicMax(reinterpret_cast<unsigned int*>(out) + (size_t)b * C4 + ch, __float_as_uint(v));   // v >= 0: uint order == float order
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

DPF_API int dpf_pointnet_workspace_bytes(long long* bytes) {
  DPF_REQUIRE(bytes, DPF_ERR_NULL_PTR, "dpf_pointnet_workspace_bytes: null out pointer");
  *bytes = (long long)WS_BYTES;
  return DPF_OK;
}

// weights[l] (out, in) fp32 row-major for the four SharedDots (64x3, 128x64, 256x128, 512x256);
// bn[l] = {gamma, beta, running_mean, running_var} device pointers (4 per layer, 16 in total).
DPF_API int dpf_pointnet_eval_forward(const float* x, int B, int N, const float* const* weights, const float* const* bn,
                                      float bn_eps, void* workspace, float* out, void* stream) {
  DPF_REQUIRE(x && weights && bn && workspace && out, DPF_ERR_NULL_PTR, "dpf_pointnet_eval_forward: null pointer");
  DPF_REQUIRE(B > 0 && N > 0 && B <= 65535, DPF_ERR_BAD_ARG, "dpf_pointnet_eval_forward: bad sizes B=%d N=%d", B, N);
  DPF_REQUIRE(((uintptr_t)workspace & 255) == 0, DPF_ERR_ALIGN, "dpf_pointnet_eval_forward: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  PnParams p{};
  for (int l = 0; l < 4; ++l) {
    DPF_REQUIRE(weights[l] && bn[4 * l] && bn[4 * l + 1] && bn[4 * l + 2] && bn[4 * l + 3], DPF_ERR_NULL_PTR,
                "dpf_pointnet_eval_forward: null parameter pointer (layer %d)", l);
    p.W[l] = weights[l];
    p.gamma[l] = bn[4 * l]; p.beta[l] = bn[4 * l + 1]; p.rm[l] = bn[4 * l + 2]; p.rv[l] = bn[4 * l + 3];
  }
  p.bn_eps = bn_eps;
  const int chunks = C2 * (C1 / 8) + C3 * (C2 / 8) + C4 * (C3 / 8);
  pointnet_pack_kernel<<<(chunks + 255) / 256, 256, 0, s>>>(p, (unsigned char*)workspace);
  int rc = dpf_check_launch("pointnet_pack_kernel");
  if (rc) return rc;
  cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * C4, s);
  const size_t smem = sizeof(PnSmem) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pointnet_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const int tiles_per_b = (N + DPF_TILE - 1) / DPF_TILE;
  int S = dpf_num_sms() / B;                      // slices per shape so that B * S <= #SMs (one wave)
  S = S < 1 ? 1 : (S > tiles_per_b ? tiles_per_b : S);
  pointnet_eval_kernel<<<dim3(S, B), PN_T, smem, s>>>(x, (const unsigned char*)workspace, out, B, N);
  return dpf_check_launch("pointnet_eval_kernel");
}
