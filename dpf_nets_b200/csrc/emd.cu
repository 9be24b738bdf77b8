// Approximate earth-mover distance (soft auction over 9 temperature levels) for sm_100a.
//
// Reference: approxmatchkernel / matchcostkernel / matchcostgrad{1,2}kernel
// (lib/metrics/pytorch_structural_losses/src/approxmatch.cu:3-326, semantics in SURVEY.md App. D).
//
// Two forms:
//  * FFI-compatible: dpf_approxmatch materialises match (b, m, n) like the reference, dpf_matchcost /
//    dpf_matchcost_grad consume it.  One CTA per batch element, per-thread sums in ascending index
//    order exactly like the reference's sequential inner loops (same rounding), 64-bit indexing (the
//    reference overflows int for b > 511 at 2048^2), scratch in shared memory instead of `temp`.
//  * Fused all-pairs cost: dpf_pairwise_emd never writes `match` (16.8 MB per cloud pair at 2048^2, RMW
//    9x in the reference): cost = sum over levels of w * |x1 - x2| is accumulated while the level's
//    weights are formed.  One CTA per cloud pair, MUFU(ex2)-bound, no HBM traffic beyond the points.
#include "common.cuh"

namespace {

constexpr int EMD_THREADS = 512;
constexpr int EMD_TILE = 1024;   // points staged per shared-memory tile (float4: x, y, z, weight)

// exp(level * d) as ONE multiply + MUFU.EX2: `level2` = level * log2(e) is folded on the caller's side (the reference's
// __expf(level * d) is ex2.approx((level * d) * log2e): one more multiply, argument rounded once more - the two differ
// by <= 1 ulp of an argument of magnitude < ~100 wherever the weight is not negligible, far inside the 1e-4 gate)
__device__ __forceinline__ float emd_w(float level2, float x1, float y1, float z1, float x2, float y2, float z2) {
  const float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
  float w;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(d * level2));
  return w;
}

// Packed fp32x2 evaluation (FADD2 / FMUL2 / FFMA2 of sm_100a, per-lane IEEE rounding): one register-resident point
// against TWO staged points per instruction.  The staged tile holds point pairs as
//   buf[2i] = {x_a, x_b, y_a, y_b},  buf[2i+1] = {z_a, z_b, w_a, w_b}      (a = point 2i, b = point 2i+1 of the tile)
// so that every packed operand is an aligned register pair of one LDS.128.  An odd tail is padded with a zero-weight
// point.  The sums over staged points stay sequential in ascending index (a, then b) like the reference's loops.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 emd_pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void emd_unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 emd_sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 emd_mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 emd_fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float emd_ex2(float x) {
  float w;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(x));
  return w;
}
// squared distances of the staged pair (A, B) to the point (x1, y1, z1), and exp(level * d) of both
__device__ __forceinline__ void emd_pair(float4 A, float4 B, float x1, float y1, float z1, float level2, float& d_a, float& d_b,
                                         float& e_a, float& e_b) {
  const f32x2 dx = emd_sub2(emd_pack2(A.x, A.y), emd_pack2(x1, x1));
  const f32x2 dy = emd_sub2(emd_pack2(A.z, A.w), emd_pack2(y1, y1));
  const f32x2 dz = emd_sub2(emd_pack2(B.x, B.y), emd_pack2(z1, z1));
  const f32x2 d = emd_fma2(dz, dz, emd_fma2(dy, dy, emd_mul2(dx, dx)));
  float a0, a1;
  emd_unpack2(emd_mul2(d, emd_pack2(level2, level2)), a0, a1);
  emd_unpack2(d, d_a, d_b);
  e_a = emd_ex2(a0);
  e_b = emd_ex2(a1);
}
// stage `cnt` points (xyz of src + weight) in the pair layout, zero-weight padding up to an even count
__device__ __forceinline__ void emd_stage(float4* buf, const float* __restrict__ src, const float* wsrc, int cnt) {
  float* fb = reinterpret_cast<float*>(buf);
  const int padded = (cnt + 1) & ~1;
  for (int l = threadIdx.x; l < padded; l += EMD_THREADS) {
    const bool ok = l < cnt;
    const int base = (l >> 1) * 8 + (l & 1);
    fb[base + 0] = ok ? src[l * 3 + 0] : 0.f;
    fb[base + 2] = ok ? src[l * 3 + 1] : 0.f;
    fb[base + 4] = ok ? src[l * 3 + 2] : 0.f;
    fb[base + 6] = ok ? wsrc[l] : 0.f;
  }
}

// One soft-assignment run for the cloud pair (p1[n], p2[m]).  remainL/ratioL [n], remainR/ratioR [m]
// live in shared memory.  MATCH: accumulate into match (m, n) like the reference.  COST: accumulate
// sum w * dist into the per-thread `cost`.
template <bool MATCH, bool COST>
__device__ void approxmatch_pair(int n, int m, const float* __restrict__ p1, const float* __restrict__ p2,
                                 float* __restrict__ match, float* remainL, float* remainR, float* ratioL,
                                 float* ratioR, float4* buf, float& cost) {
  const int tid = threadIdx.x;
  float multiL, multiR;
  if (n >= m) { multiL = 1; multiR = n / m; }
  else        { multiL = m / n; multiR = 1; }
  if (MATCH)
    for (size_t j = tid; j < (size_t)n * m; j += EMD_THREADS) match[j] = 0;
  for (int j = tid; j < n; j += EMD_THREADS) remainL[j] = multiL;
  for (int j = tid; j < m; j += EMD_THREADS) remainR[j] = multiR;
  __syncthreads();
  for (int j = 7; j > -2; j--) {
    const float level = -powf(4.0f, j) * 1.4426950408889634f;   // level * log2(e), see emd_w
    // (1) ratioL[k] = remainL[k] / (1e-9 + sum_l exp(level d_kl) remainR[l])
    for (int k0 = 0; k0 < n; k0 += EMD_THREADS) {
      const int k = k0 + tid;
      float x1 = 0, y1 = 0, z1 = 0;
      if (k < n) { x1 = p1[k * 3 + 0]; y1 = p1[k * 3 + 1]; z1 = p1[k * 3 + 2]; }
      float suml = 1e-9f;
      for (int l0 = 0; l0 < m; l0 += EMD_TILE) {
        const int lend = min(m, l0 + EMD_TILE) - l0;
        emd_stage(buf, p2 + (size_t)l0 * 3, remainR + l0, lend);
        __syncthreads();
        for (int pi = 0; pi < (lend + 1) / 2; pi++) {
          const float4 A = buf[2 * pi], Bq = buf[2 * pi + 1];
          float da, db, ea, eb;
          emd_pair(A, Bq, x1, y1, z1, level, da, db, ea, eb);
          suml += ea * Bq.z;
          suml += eb * Bq.w;
        }
        __syncthreads();
      }
      if (k < n) ratioL[k] = remainL[k] / suml;
    }
    __syncthreads();
    // (2) per l: consumption and new remainR
    for (int l0 = 0; l0 < m; l0 += EMD_THREADS) {
      const int l = l0 + tid;
      float x2 = 0, y2 = 0, z2 = 0;
      if (l < m) { x2 = p2[l * 3 + 0]; y2 = p2[l * 3 + 1]; z2 = p2[l * 3 + 2]; }
      float sumr = 0;
      for (int k0 = 0; k0 < n; k0 += EMD_TILE) {
        const int kend = min(n, k0 + EMD_TILE) - k0;
        emd_stage(buf, p1 + (size_t)k0 * 3, ratioL + k0, kend);
        __syncthreads();
        for (int pi = 0; pi < (kend + 1) / 2; pi++) {
          const float4 A = buf[2 * pi], Bq = buf[2 * pi + 1];
          float da, db, ea, eb;
          emd_pair(A, Bq, x2, y2, z2, level, da, db, ea, eb);
          sumr += ea * Bq.z;
          sumr += eb * Bq.w;
        }
        __syncthreads();
      }
      if (l < m) {
        sumr *= remainR[l];
        const float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
        ratioR[l] = consumption * remainR[l];
        remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
      }
    }
    __syncthreads();
    // (3) weights of this level: match += w, remainL -= sum_l w (and the fused cost)
    for (int k0 = 0; k0 < n; k0 += EMD_THREADS) {
      const int k = k0 + tid;
      float x1 = 0, y1 = 0, z1 = 0;
      if (k < n) { x1 = p1[k * 3 + 0]; y1 = p1[k * 3 + 1]; z1 = p1[k * 3 + 2]; }
      float suml = 0;
      const float rl = (k < n) ? ratioL[k] : 0.f;
      for (int l0 = 0; l0 < m; l0 += EMD_TILE) {
        const int lend = min(m, l0 + EMD_TILE) - l0;
        emd_stage(buf, p2 + (size_t)l0 * 3, ratioR + l0, lend);
        __syncthreads();
        if (k < n) {
          for (int pi = 0; pi < (lend + 1) / 2; pi++) {
            const float4 A = buf[2 * pi], Bq = buf[2 * pi + 1];
            float da, db, ea, eb;
            emd_pair(A, Bq, x1, y1, z1, level, da, db, ea, eb);
            const float wa = ea * rl * Bq.z, wb = eb * rl * Bq.w;
            if (MATCH) {
              match[(size_t)(l0 + 2 * pi) * n + k] += wa;
              if (2 * pi + 1 < lend) match[(size_t)(l0 + 2 * pi + 1) * n + k] += wb;
            }
            if (COST) {   // |x1 - x2| = d * rsqrt(d) without the IEEE sqrt sequence
              cost = fmaf(wa, da * rsqrtf(fmaxf(da, 1e-30f)), cost);
              cost = fmaf(wb, db * rsqrtf(fmaxf(db, 1e-30f)), cost);
            }
            suml += wa;
            suml += wb;
          }
        }
        __syncthreads();
      }
      if (k < n) remainL[k] = fmaxf(0.0f, remainL[k] - suml);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(EMD_THREADS)
approxmatch_kernel(int b, int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                   float* __restrict__ match) {
  extern __shared__ float4 emd_sm[];
  float4* buf = emd_sm;
  float* remainL = reinterpret_cast<float*>(buf + EMD_TILE);
  float* remainR = remainL + n;
  float* ratioL = remainR + m;
  float* ratioR = ratioL + n;
  float cost = 0.f;
  for (int i = blockIdx.x; i < b; i += gridDim.x)
    approxmatch_pair<true, false>(n, m, xyz1 + (size_t)i * n * 3, xyz2 + (size_t)i * m * 3, match + (size_t)i * n * m,
                                  remainL, remainR, ratioL, ratioR, buf, cost);
}

// out[i, j] = approx EMD cost between A_i and B_j (un-normalised, like MatchCost's output)
__global__ void __launch_bounds__(EMD_THREADS)
pairwise_emd_kernel(int S2, int n, int m, const float* __restrict__ A, const float* __restrict__ B,
                    float* __restrict__ out, int row_start, int row_step, long long n_pairs) {
  extern __shared__ float4 emd_sm[];
  __shared__ float red[EMD_THREADS / 32];
  float4* buf = emd_sm;
  float* remainL = reinterpret_cast<float*>(buf + EMD_TILE);
  float* remainR = remainL + n;
  float* ratioL = remainR + m;
  float* ratioR = ratioL + n;
  for (long long e = blockIdx.x; e < n_pairs; e += gridDim.x) {
    const int i = row_start + (int)(e / S2) * row_step, j = (int)(e % S2);
    float cost = 0.f;
    approxmatch_pair<false, true>(n, m, A + (size_t)i * n * 3, B + (size_t)j * m * 3, nullptr, remainL, remainR, ratioL,
                                  ratioR, buf, cost);
    cost = warp_sum(cost);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cost;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < EMD_THREADS / 32; ++w) t += red[w];
      out[(size_t)i * S2 + j] = t;
    }
    __syncthreads();
  }
}

// out[i] = sum_l sum_k match[i,l,k] |x1_k - x2_l|   (approxmatch.cu:184-224)
__global__ void __launch_bounds__(256)
matchcost_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                 const float* __restrict__ match, float* __restrict__ out, int lchunks) {
  __shared__ float red[8];
  const int i = blockIdx.x / lchunks, lc = blockIdx.x - i * lchunks;
  const int l0 = (int)((long long)m * lc / lchunks), l1 = (int)((long long)m * (lc + 1) / lchunks);
  const float* p1 = xyz1 + (size_t)i * n * 3;
  const float* p2 = xyz2 + (size_t)i * m * 3;
  const float* mt = match + (size_t)i * n * m;
  float sub = 0.f;
  for (int l = l0; l < l1; ++l) {
    const float x2 = p2[l * 3 + 0], y2 = p2[l * 3 + 1], z2 = p2[l * 3 + 2];
    for (int k = threadIdx.x; k < n; k += 256) {
      const float dx = x2 - p1[k * 3 + 0], dy = y2 - p1[k * 3 + 1], dz = z2 - p1[k * 3 + 2];
      sub = fmaf(mt[(size_t)l * n + k], sqrtf(dx * dx + dy * dy + dz * dz), sub);
    }
  }
  sub = warp_sum(sub);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sub;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out + i, t);
  }
}

// grad1[i,k] = sum_l match[l,k] (x1_k - x2_l)/max(|.|,1e-10);  grad2[i,l] = sum_k match[l,k] (x2_l - x1_k)/...
// (approxmatch.cu:229-291)
__global__ void __launch_bounds__(256)
matchcost_grad1_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                       const float* __restrict__ match, float* __restrict__ grad1, int kchunks) {
  const int i = blockIdx.x / kchunks, kc = blockIdx.x - i * kchunks;
  const int k = kc * 256 + threadIdx.x;
  if (k >= n) return;
  const float* p1 = xyz1 + (size_t)i * n * 3;
  const float* p2 = xyz2 + (size_t)i * m * 3;
  const float* mt = match + (size_t)i * n * m;
  const float x1 = p1[k * 3 + 0], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
  float dx = 0, dy = 0, dz = 0;
  for (int l = 0; l < m; ++l) {
    const float x2 = p2[l * 3 + 0], y2 = p2[l * 3 + 1], z2 = p2[l * 3 + 2];
    const float d = mt[(size_t)l * n + k] * rsqrtf(fmaxf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2), 1e-20f));
    dx += (x1 - x2) * d; dy += (y1 - y2) * d; dz += (z1 - z2) * d;
  }
  float* g = grad1 + ((size_t)i * n + k) * 3;
  g[0] = dx; g[1] = dy; g[2] = dz;
}

__global__ void __launch_bounds__(256)
matchcost_grad2_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                       const float* __restrict__ match, float* __restrict__ grad2) {
  __shared__ float red[3][8];
  const int i = blockIdx.x / m, l = blockIdx.x - i * m;
  const float* p1 = xyz1 + (size_t)i * n * 3;
  const float* p2 = xyz2 + ((size_t)i * m + l) * 3;
  const float* mt = match + (size_t)i * n * m + (size_t)l * n;
  const float x2 = p2[0], y2 = p2[1], z2 = p2[2];
  float sx = 0, sy = 0, sz = 0;
  for (int k = threadIdx.x; k < n; k += 256) {
    const float x1 = x2 - p1[k * 3 + 0], y1 = y2 - p1[k * 3 + 1], z1 = z2 - p1[k * 3 + 2];
    const float d = mt[k] * rsqrtf(fmaxf(x1 * x1 + y1 * y1 + z1 * z1, 1e-20f));
    sx += x1 * d; sy += y1 * d; sz += z1 * d;
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sx; red[1][threadIdx.x >> 5] = sy; red[2][threadIdx.x >> 5] = sz; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    grad2[((size_t)i * m + l) * 3 + threadIdx.x] = t;
  }
}

size_t emd_smem_bytes(int n, int m) { return sizeof(float4) * EMD_TILE + sizeof(float) * 2 * ((size_t)n + m); }

}  // namespace

// Replaces approxmatch() (approxmatch.cuh:6, approxmatch.cu:299-307; shim structural_loss.cpp:22-37).
// match (b, m, n) is overwritten.  `temp` of the reference signature is not needed (scratch lives in
// shared memory) and may be NULL.
DPF_API int dpf_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, float* temp,
                            void* stream) {
  (void)temp;
  DPF_REQUIRE(b >= 0 && n > 0 && m > 0, DPF_ERR_BAD_ARG, "dpf_approxmatch: bad sizes");
  if (b == 0) return DPF_OK;
  DPF_REQUIRE(xyz1 && xyz2 && match, DPF_ERR_NULL_PTR, "dpf_approxmatch: null pointer");
  const size_t smem = emd_smem_bytes(n, m);
  DPF_REQUIRE(smem <= 200 * 1024, DPF_ERR_UNSUPPORTED, "dpf_approxmatch: n + m = %d exceeds the shared-memory scratch (max ~23000)", n + m);
  cudaFuncSetAttribute(approxmatch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  approxmatch_kernel<<<b, EMD_THREADS, smem, (cudaStream_t)stream>>>(b, n, m, xyz1, xyz2, match);
  return dpf_check_launch("approxmatch_kernel");
}

// Replaces matchcost() (approxmatch.cuh:7, approxmatch.cu:309-316; shim structural_loss.cpp:39-52).
DPF_API int dpf_matchcost(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* out,
                          void* stream) {
  DPF_REQUIRE(b >= 0 && n > 0 && m > 0, DPF_ERR_BAD_ARG, "dpf_matchcost: bad sizes");
  if (b == 0) return DPF_OK;
  DPF_REQUIRE(xyz1 && xyz2 && match && out, DPF_ERR_NULL_PTR, "dpf_matchcost: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(float) * b, s);
  int lchunks = 1;
  while ((long long)b * lchunks < 4LL * dpf_num_sms() && lchunks * 2 <= m) lchunks *= 2;
  matchcost_kernel<<<b * lchunks, 256, 0, s>>>(n, m, xyz1, xyz2, match, out, lchunks);
  return dpf_check_launch("matchcost_kernel");
}

// Replaces matchcostgrad() (approxmatch.cuh:8, approxmatch.cu:318-326; shim structural_loss.cpp:54-69).
DPF_API int dpf_matchcost_grad(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match,
                               float* grad1, float* grad2, void* stream) {
  DPF_REQUIRE(b >= 0 && n > 0 && m > 0, DPF_ERR_BAD_ARG, "dpf_matchcost_grad: bad sizes");
  if (b == 0) return DPF_OK;
  DPF_REQUIRE(xyz1 && xyz2 && match && grad1 && grad2, DPF_ERR_NULL_PTR, "dpf_matchcost_grad: null pointer");
  DPF_REQUIRE((long long)b * m < 2147483647LL, DPF_ERR_BAD_ARG, "dpf_matchcost_grad: b*m too large");
  cudaStream_t s = (cudaStream_t)stream;
  const int kchunks = (n + 255) / 256;
  matchcost_grad1_kernel<<<b * kchunks, 256, 0, s>>>(n, m, xyz1, xyz2, match, grad1, kchunks);
  int rc = dpf_check_launch("matchcost_grad1_kernel");
  if (rc) return rc;
  matchcost_grad2_kernel<<<b * m, 256, 0, s>>>(n, m, xyz1, xyz2, match, grad2);
  return dpf_check_launch("matchcost_grad2_kernel");
}

// Fused EMD half of _pairwise_EMD_CD_ (lib/metrics/evaluation_metrics.py:85-121) without the dense
// match: out[i*S2+j] = MatchCost(A_i, B_j) for rows i = row_start + t*row_step, t < n_rows
// (divide by n for emd_approx, evaluation_metrics.py:26-31).
DPF_API int dpf_pairwise_emd(int S1, int S2, int n, int m, const float* A, const float* B, float* out, int row_start,
                             int row_step, int n_rows, void* stream) {
  DPF_REQUIRE(S1 >= 0 && S2 >= 0 && n > 0 && m > 0, DPF_ERR_BAD_ARG, "dpf_pairwise_emd: bad sizes");
  DPF_REQUIRE(row_start >= 0 && row_step >= 1 && n_rows >= 0, DPF_ERR_BAD_ARG, "dpf_pairwise_emd: bad row range");
  if (n_rows == 0 || S2 == 0) return DPF_OK;
  DPF_REQUIRE(row_start + (long long)(n_rows - 1) * row_step < S1, DPF_ERR_BAD_ARG, "dpf_pairwise_emd: rows exceed S1");
  DPF_REQUIRE(A && B && out, DPF_ERR_NULL_PTR, "dpf_pairwise_emd: null pointer");
  const size_t smem = emd_smem_bytes(n, m);
  DPF_REQUIRE(smem <= 200 * 1024, DPF_ERR_UNSUPPORTED, "dpf_pairwise_emd: n + m = %d exceeds the shared-memory scratch", n + m);
  cudaFuncSetAttribute(pairwise_emd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const long long n_pairs = (long long)n_rows * S2;
  const int per_sm = (int)max((size_t)1, min((size_t)4, (size_t)(220 * 1024) / (smem + 1024)));
  const int grid = (int)min(n_pairs, (long long)dpf_num_sms() * per_sm);
  pairwise_emd_kernel<<<grid, EMD_THREADS, smem, (cudaStream_t)stream>>>(S2, n, m, A, B, out, row_start, row_step, n_pairs);
  return dpf_check_launch("pairwise_emd_kernel");
}
