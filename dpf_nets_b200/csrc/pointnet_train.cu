// Train-mode PointNet cloud encoder, last layer + max-pool: everything the batch-statistics BatchNorm and the max over the
// points need from  h = W h2  (W: 512 x 256 SharedDot weight, h2: (B,256,N) post-ReLU activations of the layer before)
// WITHOUT materialising h (134 MB at 32 x 2048; the library path writes it, normalises it, clamps it and reduces it in
// five full passes).  Reference: PointNetCloudEncoder.features.{sd2, sd2_bn, sd2_relu} in .train() (lib/networks/
// encoders.py:9-28) followed by torch.max over the points (lib/networks/models.py:130-131).
//
// BatchNorm + ReLU are monotone per channel, so  max_n relu(gamma (h - mu)/sigma + beta)  is a function of max_n h (gamma >= 0)
// or min_n h (gamma < 0): one pass over the points yields, per (shape, channel), the mean and the sum of squared deviations of
// h (merged into the batch statistics by the caller) and max / min of h with their point indices (the max-pool's
// selection for the backward).
//
// One CTA = one shape x one chunk of 128 output channels.  z^T[128 channels x 64 points] = W(chunk) h2(tile) on the tensor
// cores with ONE CHANNEL PER TMEM LANE, so every reduction over points is a per-thread register reduction:
//   A = W chunk [128 x 256] K-major SW128 (bf16 hi + lo images, TMA bulk copy, resident for the CTA's lifetime)
//   B = h2 tile [256 x 64 points] MN-major SW128 (points contiguous - the tensor's own layout; fp32 -> bf16 hi + lo by the CTA)
//   three UMMA chains hi*hi + lo*hi + hi*lo (fp32-class accuracy: batch statistics divide by small deviations).
#include "coupling.cuh"
#include "umma.cuh"
#include "pointnet_tiles.cuh"
#include <math_constants.h>

namespace {

constexpr int PT_T = 512;                       // threads: lane = tid & 127 = channel of the chunk, part = tid >> 7 = quarter of the tile's points
constexpr int PT_CIN = 256, PT_COUT = 512, PT_NT = 64;     // input channels, output channels, points per tile
constexpr int PT_PARTS = PT_T / 128, PT_PC = PT_NT / PT_PARTS;      // 4 parts x 16 columns
constexpr uint32_t PT_KB = 128 * 128;           // 16 KB: [128 rows x 64 bf16] K-block of the weight chunk
constexpr uint32_t PT_WIMG = 4 * PT_KB;         // 64 KB: one chunk image (4 K-blocks)
constexpr uint32_t PT_BIMG = PT_CIN * 128;      // 32 KB: [256 K-rows x 64 points] bf16
constexpr uint64_t PT_DESC_A = umma::make_desc_template(16, 1024, umma::LAYOUT_SW128);
constexpr uint64_t PT_DESC_B = umma::make_desc_template(PT_BIMG, 1024, umma::LAYOUT_SW128);     // MN-major, a single 64-wide block
constexpr uint32_t PT_IDESC = umma::make_idesc_bf16(128, PT_NT, 0, 1);

// W (512 x 256 fp32 row-major) -> [4 chunks][hi, lo][4 K-blocks][128 x 64] swizzled bf16 images; one thread per 16-byte chunk
__global__ void __launch_bounds__(256)
pool_pack_kernel(const float* __restrict__ W, unsigned char* __restrict__ img) {
  const int i = blockIdx.x * 256 + threadIdx.x;            // (row r, 16-byte chunk q of the row's 256 inputs)
  if (i >= PT_COUT * (PT_CIN / 8)) return;
  const int r = i / 32, q = i % 32;
  const float* src = W + (size_t)r * PT_CIN + q * 8;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = src[2 * e], b = src[2 * e + 1];
    hi[e] = umma::pack_bf16(a, b);
    lo[e] = umma::pack_bf16(a - __uint_as_float(hi[e] << 16), b - __uint_as_float(hi[e] & 0xffff0000u));
  }
  unsigned char* base = img + (size_t)(r >> 7) * 2 * PT_WIMG + (size_t)(q >> 3) * PT_KB + umma::sw128_offset(r & 127, q & 7);
  *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(base + PT_WIMG) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

struct PtSmem {
  unsigned char A[2 * PT_WIMG];         // W chunk hi | lo
  unsigned char Bt[2 * PT_BIMG];        // h2 tile hi | lo
  float red[PT_PARTS][128][8];          // part -> {mean, M2, max, min, argmax, argmin, count} hand-over
  uint64_t bar_w, bar_mma;
  uint32_t tmem_base;
};

template <bool IN_ACT>
__global__ void __launch_bounds__(PT_T, 1)
pool_forward_kernel(const float* __restrict__ h2, const float* __restrict__ tab, const unsigned char* __restrict__ wimg, int B, int N,
                    float* __restrict__ stat, float* __restrict__ vmax, float* __restrict__ vmin,
                    int* __restrict__ imax, int* __restrict__ imin, float* __restrict__ asum) {
  extern __shared__ unsigned char smraw[];
  PtSmem& s = *reinterpret_cast<PtSmem*>(smraw + ((1024u - (umma::smem_u32(smraw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5;
  const int lane_c = tid & 127, part = tid >> 7, quarter = warp & 3;
  const int b = blockIdx.x, chunk = blockIdx.y;
  if (tid == 0) {
    umma::mbar_init(&s.bar_w, 1);
    umma::mbar_init(&s.bar_mma, 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(&s.tmem_base, 64);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
  if (tid == 0) {
    umma::mbar_expect_tx(&s.bar_w, 2 * PT_WIMG);
    umma::bulk_g2s(s.A, wimg + (size_t)chunk * 2 * PT_WIMG, 2 * PT_WIMG, &s.bar_w);
  }
  // statistics as SHIFTED sums  sum (x - r), sum (x - r)^2  with r = the channel's first value in this thread: the variance
  // then comes out of numbers of the size of the deviations, not of E[x^2] - mean^2 (batch statistics divide by it)
  float sum = 0.f, sq = 0.f, mx = -CUDART_INF_F, mn = CUDART_INF_F, shift = 0.f;
  int amx = 0, amn = 0, cnt = 0;
  const int n_tiles = (N + PT_NT - 1) / PT_NT;
  uint32_t ph = 0;
  // thread k owns input channel k: its 64 consecutive points of a tile are 16 independent 16-byte loads, all issued before
  // the first use and - for the NEXT tile - before this tile's UMMA wait and epilogue, so the global-load latency (the
  // kernel's top stall in ncu) overlaps the tensor-core work
  // tab != null: h2 is the layer's PRE-BatchNorm input Z and the operand is relu(sc z + sh) (8 floats per channel, {sc, sh, 0, ..});
  // columns beyond N then hold relu(sh) instead of 0, which is harmless: the epilogue never reads them.
  // Coalesced row loads + neighbour swap (pointnet_tiles.cuh); the NEXT tile's loads are issued before this tile's UMMA wait
  // and epilogue, so the global-load latency overlaps the tensor-core work.
  const bool want_asum = asum != nullptr && chunk == 0;
  pnt::TileLoader<PT_CIN, IN_ACT ? pnt::LD_AFFINE : pnt::LD_RAW, PT_T / 32> ld;
  ld.init(h2, nullptr, tab, N, tid);
  constexpr int NSLOT = PT_CIN / (4 * (PT_T / 32));       // (pair, lane parity) slots per thread
  float act_sum[NSLOT];             // per-slot sums of the operand over the valid points (the analytic backward's S)
#pragma unroll
  for (int i = 0; i < NSLOT; ++i) act_sum[i] = 0.f;
  const int tile0 = b * n_tiles;    // flat tile index of the loaders
  ld.load(tile0);
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int n0 = tile * PT_NT;
    // ---- h2 tile -> bf16 hi | lo, MN-major (8 chunks of 16 bytes per K-row) ----
    if (want_asum) ld.template convert<false, true>(s.Bt, s.Bt + PT_BIMG, tile0 + tile, act_sum);
    else ld.template convert<false, false>(s.Bt, s.Bt + PT_BIMG, tile0 + tile, nullptr);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      if (tile == 0) umma::mbar_wait(&s.bar_w, 0);
      umma::fence_after_sync();
      // K = 256 = 4 K-blocks (A) x 4 steps of 16; the B tile advances 16 K-rows = 2048 bytes per step
#pragma unroll
      for (int chain = 0; chain < 3; ++chain) {         // hi*hi, lo*hi, hi*lo
        const uint32_t a_base = umma::smem_u32(s.A) + (chain == 1 ? PT_WIMG : 0);
        const uint32_t b_base = umma::smem_u32(s.Bt) + (chain == 2 ? PT_BIMG : 0);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma::mma_bf16(tmem, umma::desc_at(PT_DESC_A, a_base + kb * PT_KB + 32 * k),
                           umma::desc_at(PT_DESC_B, b_base + (kb * 4 + k) * 2048), PT_IDESC, (chain | kb | k) > 0);
      }
      umma::mma_commit(&s.bar_mma);
    }
    if (tile + 1 < n_tiles) {                         // in flight during the UMMA chain and the epilogue
      ld.load(tile0 + tile + 1);
    }
    umma::mbar_wait(&s.bar_mma, ph);
    ph ^= 1;
    umma::fence_after_sync();
    // ---- lane = channel: statistics and max / min over this part's 16 points, in registers ----
    {
      uint32_t r[PT_PC];
      umma::tmem_ld16_issue(tmem + lane_off + part * PT_PC, r);
      umma::tmem_ld_wait16(r);
      const int nbase = n0 + part * PT_PC;
      if (nbase + PT_PC <= N) {       // all columns valid (every tile but a ragged last one): no per-element bookkeeping
        if (cnt == 0) shift = __uint_as_float(r[0]);
        cnt += PT_PC;
#pragma unroll
        for (int i = 0; i < PT_PC; ++i) {
          const float x = __uint_as_float(r[i]);
          const float dx = x - shift;
          sum += dx;
          sq = fmaf(dx, dx, sq);
          if (x > mx) { mx = x; amx = nbase + i; }
          if (x < mn) { mn = x; amn = nbase + i; }
        }
      } else {
#pragma unroll
        for (int i = 0; i < PT_PC; ++i) {
          if (nbase + i < N) {
            const float x = __uint_as_float(r[i]);
            if (cnt == 0) shift = x;
            ++cnt;
            const float dx = x - shift;
            sum += dx;
            sq = fmaf(dx, dx, sq);
            if (x > mx) { mx = x; amx = nbase + i; }
            if (x < mn) { mn = x; amn = nbase + i; }
          }
        }
      }
    }
    umma::fence_before_sync();
    __syncthreads();          // the accumulator and the tile buffer are free again
  }
  if (want_asum) {      // a slot's row is shared by the 8 lanes of the same parity in the same half-warp
#pragma unroll
    for (int i = 0; i < NSLOT; ++i) {
      float v = act_sum[i];
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      if ((tid & 14) == 0) asum[(size_t)b * PT_CIN + ld.slot_row(i)] = v;
    }
  }
  // ---- combine the four point quarters ----
  // per thread: count, mean = shift + sum / cnt, M2 = sq - sum^2 / cnt (sum of squared deviations from that mean)
  const float fc = (float)cnt;
  const float mean_t = cnt ? shift + sum / fc : 0.f;
  const float m2_t = cnt ? fmaxf(sq - sum * sum / fc, 0.f) : 0.f;
  s.red[part][lane_c][0] = mean_t; s.red[part][lane_c][1] = m2_t; s.red[part][lane_c][2] = mx; s.red[part][lane_c][3] = mn;
  s.red[part][lane_c][4] = __int_as_float(amx); s.red[part][lane_c][5] = __int_as_float(amn);
  s.red[part][lane_c][6] = fc;
  __syncthreads();
  if (part == 0) {
    // parts hold increasing point indices within every tile: on ties the lower index wins
    const int c = chunk * 128 + lane_c;
    float m1 = mx, m0 = mn;
    int a1 = amx, a0 = amn;
    float n1 = fc, mean = mean_t, m2 = m2_t;
#pragma unroll
    for (int p = 1; p < PT_PARTS; ++p) {
      const float* o = s.red[p][lane_c];
      const int oamx = __float_as_int(o[4]), oamn = __float_as_int(o[5]);
      if (o[2] > m1 || (o[2] == m1 && oamx < a1)) { m1 = o[2]; a1 = oamx; }
      if (o[3] < m0 || (o[3] == m0 && oamn < a0)) { m0 = o[3]; a0 = oamn; }
      // pairwise merge of the (count, mean, M2) triples (Chan et al.)
      const float n2 = o[6], nt = n1 + n2;
      if (nt > 0.f) {
        const float delta = o[0] - mean;
        mean += delta * (n2 / nt);
        m2 += o[1] + delta * delta * (n1 * n2 / nt);
        n1 = nt;
      }
    }
    vmax[(size_t)b * PT_COUT + c] = m1;
    vmin[(size_t)b * PT_COUT + c] = m0;
    imax[(size_t)b * PT_COUT + c] = a1;
    imin[(size_t)b * PT_COUT + c] = a0;
    stat[((size_t)b * PT_COUT + c) * 2 + 0] = n1 > 0.f ? mean : 0.f;
    stat[((size_t)b * PT_COUT + c) * 2 + 1] = n1 > 0.f ? m2 : 0.f;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

}  // namespace

DPF_API int dpf_pointnet_pool_workspace_bytes(long long* bytes) {
  DPF_REQUIRE(bytes, DPF_ERR_NULL_PTR, "dpf_pointnet_pool_workspace_bytes: null out pointer");
  *bytes = (long long)4 * 2 * PT_WIMG;
  return DPF_OK;
}

// h2 (B,256,N) fp32, W (512,256) fp32 ->  stat (B,512,2) fp32 {mean_n h, sum_n (h - mean)^2} per (shape, channel) (the caller
// merges the B equal-sized groups into the batch statistics), vmax / vmin (B,512) fp32 = max / min over the points of
// h = W h2, imax / imin (B,512) int32 = their point indices (lowest index on exact ties).
// workspace: dpf_pointnet_pool_workspace_bytes() bytes, 256-byte aligned (weight images).
// in_tab (256, 8) fp32 nullable: per input channel {sc, sh, ...}: the operand is relu(sc h2 + sh) (h2 = the pre-BatchNorm
// output of the layer before, its BatchNorm + ReLU applied on load)
// asum (B, 256) fp32 nullable: per (shape, input channel) sum of the operand over the points.
DPF_API int dpf_pointnet_pool_forward_ex(const float* h2, const float* in_tab, const float* W, int B, int N, void* workspace, float* stat,
                                         float* vmax, float* vmin, int* imax, int* imin, float* asum, void* stream) {
  DPF_REQUIRE(h2 && W && workspace && stat && vmax && vmin && imax && imin, DPF_ERR_NULL_PTR, "dpf_pointnet_pool_forward: null pointer");
  DPF_REQUIRE(B > 0 && N > 0 && B <= 65535, DPF_ERR_BAD_ARG, "dpf_pointnet_pool_forward: bad sizes B=%d N=%d", B, N);
  DPF_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)h2 & 15) == 0, DPF_ERR_ALIGN, "dpf_pointnet_pool_forward: workspace / h2 alignment");
  cudaStream_t s = (cudaStream_t)stream;
  pool_pack_kernel<<<(PT_COUT * (PT_CIN / 8) + 255) / 256, 256, 0, s>>>(W, (unsigned char*)workspace);
  int rc = dpf_check_launch("pool_pack_kernel");
  if (rc) return rc;
  const size_t smem = sizeof(PtSmem) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pool_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(pool_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  if (in_tab) pool_forward_kernel<true><<<dim3(B, 4), PT_T, smem, s>>>(h2, in_tab, (const unsigned char*)workspace, B, N, stat, vmax, vmin, imax, imin, asum);
  else pool_forward_kernel<false><<<dim3(B, 4), PT_T, smem, s>>>(h2, in_tab, (const unsigned char*)workspace, B, N, stat, vmax, vmin, imax, imin, asum);
  return dpf_check_launch("pool_forward_kernel");
}

DPF_API int dpf_pointnet_pool_forward(const float* h2, const float* W, int B, int N, void* workspace, float* stat,
                                      float* vmax, float* vmin, int* imax, int* imin, void* stream) {
  return dpf_pointnet_pool_forward_ex(h2, nullptr, W, B, N, workspace, stat, vmax, vmin, imax, imin, nullptr, stream);
}
