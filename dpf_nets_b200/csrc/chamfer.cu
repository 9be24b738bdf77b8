// Chamfer nearest-neighbour kernels for sm_100a.
//
// Replaces the reference's NmDistanceKernel / NmDistanceGradKernel
// (lib/metrics/pytorch_structural_losses/src/nndistance.cu:2-154) and the Python
// all-pairs loop around them (lib/networks/utils.py:90-117).
//
// Arithmetic contract (bit-exact with the reference for finite inputs):
//   d(q,t) = fma(dz,dz, fma(dx,dx, dy*dy)),  dx = t.x - q.x (fp32, round-to-nearest) - the
//   contraction nvcc emits for the reference's `x2*x2+y2*y2+z2*z2` (checked in its SASS)
//   nearest = strict '<' scan in ascending target index  => lowest index wins exact ties.
//
// Layout: clouds are (batch, points, 3) fp32 contiguous.  Targets are staged in shared
// memory as float4 so that one broadcast LDS.128 feeds R register-resident queries; the
// kernels are FP32-issue-bound, not HBM-bound; pairs of queries are evaluated with packed fp32x2 instructions
// (FADD2 / FMUL2 / FFMA2 + FMNMX: ~4 issue slots per point pair instead of 7).
#include "common.cuh"
#include <math_constants.h>

extern int g_fused_pairwise;

namespace {

constexpr int kTile = 2048;  // targets per shared-memory tile (32 KB as float4)

__device__ __forceinline__ float sqdist(float4 t, float qx, float qy, float qz) {
  const float dx = __fsub_rn(t.x, qx);
  const float dy = __fsub_rn(t.y, qy);
  const float dz = __fsub_rn(t.z, qz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Packed fp32x2 form (sm_100a FADD2 / FMUL2 / FFMA2: two IEEE fp32 lanes per instruction, each rounded exactly like
// the scalar instruction, so the arithmetic contract above is unchanged): one target against TWO register-resident
// queries in 6 issue slots instead of 12.  The target coordinate enters as the scalar-broadcast operand.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// d(t, q0), d(t, q1) for the packed query pair (qx2, qy2, qz2)
__device__ __forceinline__ void sqdist2(float4 t, f32x2 qx2, f32x2 qy2, f32x2 qz2, float& d0, float& d1) {
  const f32x2 dx = sub2(pack2(t.x, t.x), qx2);
  const f32x2 dy = sub2(pack2(t.y, t.y), qy2);
  const f32x2 dz = sub2(pack2(t.z, t.z), qz2);
  unpack2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), d0, d1);
}

// Cooperative AoS(xyz) -> float4 tile load; reads are fully coalesced over the flat float array.
template <int THREADS>
__device__ __forceinline__ void load_tile(float4* tile, const float* __restrict__ src, int cnt) {
  float* t = reinterpret_cast<float*>(tile);
  const int total = cnt * 3;
  for (int e = threadIdx.x; e < total; e += THREADS) {
    const int p = e / 3;
    t[p * 4 + (e - p * 3)] = __ldg(src + e);
  }
}

// ---------------------------------------------------------------------------------------------
// One direction of the NN search with indices (the FFI-compatible path).
// grid.x = batch * chunks, chunk = THREADS*R queries.
// ---------------------------------------------------------------------------------------------
template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS)
nn_search_kernel(int nq, const float* __restrict__ Q, int nt, const float* __restrict__ T,
                 float* __restrict__ dist, int* __restrict__ idx, int chunks) {
  __shared__ float4 tile[kTile];
  const int bi = blockIdx.x / chunks;
  const int chunk = blockIdx.x - bi * chunks;
  const float* q = Q + (size_t)bi * nq * 3;
  const float* t = T + (size_t)bi * nt * 3;

  float qx[R], qy[R], qz[R], best[R];
  int besti[R];
  int qi[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    qi[r] = chunk * (THREADS * R) + r * THREADS + threadIdx.x;
    const bool ok = qi[r] < nq;
    const int s = ok ? qi[r] : 0;
    qx[r] = ok ? __ldg(q + (size_t)s * 3 + 0) : 0.f;
    qy[r] = ok ? __ldg(q + (size_t)s * 3 + 1) : 0.f;
    qz[r] = ok ? __ldg(q + (size_t)s * 3 + 2) : 0.f;
    best[r] = CUDART_INF_F;
    besti[r] = 0;
  }
  for (int t0 = 0; t0 < nt; t0 += kTile) {
    const int cnt = min(kTile, nt - t0);
    __syncthreads();
    load_tile<THREADS>(tile, t + (size_t)t0 * 3, cnt);
    __syncthreads();
    if (R % 2 == 0) {      // packed pairs of queries
      f32x2 qx2[(R + 1) / 2], qy2[(R + 1) / 2], qz2[(R + 1) / 2];
#pragma unroll
      for (int r = 0; r + 1 < R; r += 2) {
        qx2[r / 2] = pack2(qx[r], qx[r + 1]);
        qy2[r / 2] = pack2(qy[r], qy[r + 1]);
        qz2[r / 2] = pack2(qz[r], qz[r + 1]);
      }
#pragma unroll 4
      for (int k = 0; k < cnt; ++k) {
        const float4 p = tile[k];
#pragma unroll
        for (int r = 0; r + 1 < R; r += 2) {
          float d0, d1;
          sqdist2(p, qx2[r / 2], qy2[r / 2], qz2[r / 2], d0, d1);
          if (d0 < best[r]) {
            best[r] = d0;
            besti[r] = t0 + k;
          }
          if (d1 < best[r + 1]) {
            best[r + 1] = d1;
            besti[r + 1] = t0 + k;
          }
        }
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < cnt; ++k) {
        const float4 p = tile[k];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float d = sqdist(p, qx[r], qy[r], qz[r]);
          if (d < best[r]) {
            best[r] = d;
            besti[r] = t0 + k;
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (qi[r] < nq) {
      dist[(size_t)bi * nq + qi[r]] = best[r];
      idx[(size_t)bi * nq + qi[r]] = besti[r];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused all-pairs Chamfer matrix: out[i, j] = mean_k min_l d(A_i[k], B_j[l]) + mean_l min_k d(...)
// One CTA owns row cloud i and a block of JB column clouds; no expand copy, no index traffic,
// one launch per matrix.  With symmetric != 0 only j >= i is computed (A == B).
// ---------------------------------------------------------------------------------------------
template <int R, int THREADS>
__device__ __forceinline__ float one_direction_sum(float4* tile, const float* __restrict__ q, int nq,
                                                   const float* __restrict__ t, int nt) {
  float total = 0.f;
  for (int q0 = 0; q0 < nq; q0 += THREADS * R) {
    float qx[R], qy[R], qz[R], best[R];
    bool ok[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int s = q0 + r * THREADS + threadIdx.x;
      ok[r] = s < nq;
      const int ss = ok[r] ? s : 0;
      qx[r] = __ldg(q + (size_t)ss * 3 + 0);
      qy[r] = __ldg(q + (size_t)ss * 3 + 1);
      qz[r] = __ldg(q + (size_t)ss * 3 + 2);
      best[r] = CUDART_INF_F;
    }
    for (int t0 = 0; t0 < nt; t0 += kTile) {
      const int cnt = min(kTile, nt - t0);
      __syncthreads();
      load_tile<THREADS>(tile, t + (size_t)t0 * 3, cnt);
      __syncthreads();
      if (R % 2 == 0) {    // packed pairs of queries
        f32x2 qx2[(R + 1) / 2], qy2[(R + 1) / 2], qz2[(R + 1) / 2];
#pragma unroll
        for (int r = 0; r + 1 < R; r += 2) {
          qx2[r / 2] = pack2(qx[r], qx[r + 1]);
          qy2[r / 2] = pack2(qy[r], qy[r + 1]);
          qz2[r / 2] = pack2(qz[r], qz[r + 1]);
        }
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
          const float4 p = tile[k];
#pragma unroll
          for (int r = 0; r + 1 < R; r += 2) {
            float d0, d1;
            sqdist2(p, qx2[r / 2], qy2[r / 2], qz2[r / 2], d0, d1);
            best[r] = fminf(best[r], d0);
            best[r + 1] = fminf(best[r + 1], d1);
          }
        }
      } else {
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
          const float4 p = tile[k];
#pragma unroll
          for (int r = 0; r < R; ++r) best[r] = fminf(best[r], sqdist(p, qx[r], qy[r], qz[r]));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) total += ok[r] ? best[r] : 0.f;
  }
  return total;
}

template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS)
pairwise_cd_kernel(int S2, int n, int m, const float* __restrict__ A, const float* __restrict__ B,
                   float* __restrict__ out, int row_start, int row_step, int jblocks, int JB,
                   int symmetric) {
  __shared__ float4 tile[kTile];
  __shared__ float red[2][THREADS / 32];
  const int rt = blockIdx.x / jblocks;
  const int jb = blockIdx.x - rt * jblocks;
  const int i = row_start + rt * row_step;
  int j0 = jb * JB;
  const int j1 = min(S2, j0 + JB);
  if (symmetric) j0 = max(j0, i);
  const float* a = A + (size_t)i * n * 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = j0; j < j1; ++j) {
    const float* b = B + (size_t)j * m * 3;
    float s1 = one_direction_sum<R, THREADS>(tile, a, n, b, m);
    float s2 = one_direction_sum<R, THREADS>(tile, b, m, a, n);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    __syncthreads();
    if (lane == 0) {
      red[0][warp] = s1;
      red[1][warp] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) {
        t1 += red[0][w];
        t2 += red[1][w];
      }
      out[(size_t)i * S2 + j] = t1 / (float)n + t2 / (float)m;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused all-pairs Chamfer matrix, BOTH directions from ONE distance evaluation per point pair
// (the reference launches NmDistanceKernel twice, nndistance.cu:125-128, i.e. 2 n m evaluations per cloud pair).
//
// A thread keeps R query points of the row cloud A_i in registers (two per packed fp32x2 lane pair) and streams
// the targets of B_j through shared memory, two per iteration.  For every distance d(q_r, t_k):
//   * row direction  : best[r] = min(best[r], d(q_r,t_k), d(q_r,t_k+1))            one 3-input FMNMX3 per 2 evaluations
//   * column direction: c_k = min_r d(q_r, t_k) in registers (FMNMX3 tree), then ONE warp-wide CREDUX.MIN on the fp32
//     bit pattern (distances are >= +0, so unsigned order == float order) and a one-lane store into the warp's
//     row of a [warps][targets] shared-memory table; the 8 warp rows are min-combined once per cloud pair.
// Minima are exact, so per-point distances stay bit-identical with the reference kernel; the two sums run over
// shared-memory arrays of per-point minima with ONE canonical summation order, which keeps out[i,j] == out[j,i]
// bit-exact for A == B.  Requires n <= THREADS * R (one register-resident pass over the row cloud); larger row clouds
// use the two-direction kernel above.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }   // -> FMNMX3

__device__ __forceinline__ float warp_min_nonneg(float v) {
  unsigned r;
  asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(__float_as_uint(v)));
  return __uint_as_float(r);
}

// canonical sum of arr[0..len): strided per-thread partials in ascending order, xor-tree per warp, warps in order
template <int THREADS>
__device__ __forceinline__ float canonical_sum(const float* __restrict__ arr, int len, float* red) {
  float s = 0.f;
  for (int e = threadIdx.x; e < len; e += THREADS) s += arr[e];
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) t += red[w];
  return t;
}

template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
pairwise_cd_fused_kernel(int S2, int n, int m, const float* __restrict__ A, const float* __restrict__ B,
                         float* __restrict__ out, int row_start, int row_step, int jblocks, int JB, int symmetric) {
  static_assert(R % 2 == 0, "queries are processed in packed pairs");
  constexpr int WARPS = THREADS / 32;
  extern __shared__ float4 dyn_smem[];
  float4* tile = dyn_smem;                                        // [kTile] targets of the chunk
  float* colw = reinterpret_cast<float*>(tile + kTile);           // [WARPS][kTile] per-warp column minima
  float* rowm = colw + WARPS * kTile;                             // [THREADS * R] row minima of the pair
  __shared__ float red[WARPS];
  const int rt = blockIdx.x / jblocks;
  const int jb = blockIdx.x - rt * jblocks;
  const int i = row_start + rt * row_step;
  int j0 = jb * JB;
  const int j1 = min(S2, j0 + JB);
  if (symmetric) j0 = max(j0, i);
  const float* a = A + (size_t)i * n * 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // the row cloud's queries: registers for the whole CTA lifetime; slots beyond n sit at +inf (d = +inf: never a minimum)
  f32x2 qx2[R / 2], qy2[R / 2], qz2[R / 2];
#pragma unroll
  for (int r = 0; r < R; r += 2) {
    float c[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int sidx = (r + h) * THREADS + threadIdx.x;
      const bool ok = sidx < n;
#pragma unroll
      for (int d = 0; d < 3; ++d) c[h][d] = ok ? __ldg(a + (size_t)sidx * 3 + d) : CUDART_INF_F;
    }
    qx2[r / 2] = pack2(c[0][0], c[1][0]);
    qy2[r / 2] = pack2(c[0][1], c[1][1]);
    qz2[r / 2] = pack2(c[0][2], c[1][2]);
  }

  for (int j = j0; j < j1; ++j) {
    const float* b = B + (size_t)j * m * 3;
    float best[R];
#pragma unroll
    for (int r = 0; r < R; ++r) best[r] = CUDART_INF_F;
    float s2 = 0.f;
    for (int t0 = 0; t0 < m; t0 += kTile) {
      const int cnt = min(kTile, m - t0);
      __syncthreads();                                            // previous chunk / pair fully consumed
      load_tile<THREADS>(tile, b + (size_t)t0 * 3, cnt);
      if ((cnt & 1) && threadIdx.x == 0) tile[cnt] = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);   // pad to a pair
      __syncthreads();
      float* cw = colw + warp * kTile;
#pragma unroll 2
      for (int k = 0; k < cnt; k += 2) {
        const float4 p0 = tile[k], p1 = tile[k + 1];
        float d0[R], d1[R];
#pragma unroll
        for (int r = 0; r < R; r += 2) {
          sqdist2(p0, qx2[r / 2], qy2[r / 2], qz2[r / 2], d0[r], d0[r + 1]);
          sqdist2(p1, qx2[r / 2], qy2[r / 2], qz2[r / 2], d1[r], d1[r + 1]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) best[r] = min3f(best[r], d0[r], d1[r]);
        float c0 = d0[0], c1 = d1[0];
#pragma unroll
        for (int r = 1; r + 1 < R; r += 2) {
          c0 = min3f(c0, d0[r], d0[r + 1]);
          c1 = min3f(c1, d1[r], d1[r + 1]);
        }
        c0 = fminf(c0, d0[R - 1]);
        c1 = fminf(c1, d1[R - 1]);
        c0 = warp_min_nonneg(c0);
        c1 = warp_min_nonneg(c1);
        if (lane == 0) *reinterpret_cast<float2*>(cw + k) = make_float2(c0, c1);
      }
      __syncthreads();
      // combine the warps' rows of this chunk (in place into row 0), then the canonical sum over the chunk
      for (int e = threadIdx.x; e < cnt; e += THREADS) {
        float v = colw[e];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) v = fminf(v, colw[w * kTile + e]);
        colw[e] = v;
      }
      __syncthreads();
      s2 += canonical_sum<THREADS>(colw, cnt, red);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) rowm[r * THREADS + threadIdx.x] = best[r];
    __syncthreads();
    const float s1 = canonical_sum<THREADS>(rowm, n, red);
    if (threadIdx.x == 0) out[(size_t)i * S2 + j] = s1 / (float)n + s2 / (float)m;
  }
}

__global__ void symmetrize_upper_kernel(float* __restrict__ M, int S) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j < S && j > i) M[(size_t)j * S + i] = M[(size_t)i * S + j];
}

// Gradient of the squared NN distances (reference: nndistance.cu:129-154).
__global__ void nn_grad_kernel(int b, int n, const float* __restrict__ xyz1, int m,
                               const float* __restrict__ xyz2, const float* __restrict__ grad_dist1,
                               const int* __restrict__ idx1, float* __restrict__ grad_xyz1,
                               float* __restrict__ grad_xyz2) {
  const size_t total = (size_t)b * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e / n;
    const int j2 = idx1[e];
    const float* p1 = xyz1 + e * 3;
    const float* p2 = xyz2 + (i * m + j2) * 3;
    const float g = grad_dist1[e] * 2.f;
    const float gx = g * (p1[0] - p2[0]), gy = g * (p1[1] - p2[1]), gz = g * (p1[2] - p2[2]);
    atomicAdd(grad_xyz1 + e * 3 + 0, gx);
    atomicAdd(grad_xyz1 + e * 3 + 1, gy);
    atomicAdd(grad_xyz1 + e * 3 + 2, gz);
    atomicAdd(grad_xyz2 + (i * m + j2) * 3 + 0, -gx);
    atomicAdd(grad_xyz2 + (i * m + j2) * 3 + 1, -gy);
    atomicAdd(grad_xyz2 + (i * m + j2) * 3 + 2, -gz);
  }
}

template <int R, int THREADS>
int launch_nn(int b, int nq, const float* Q, int nt, const float* T, float* dist, int* idx,
              cudaStream_t s) {
  const int chunks = (nq + THREADS * R - 1) / (THREADS * R);
  nn_search_kernel<R, THREADS><<<b * chunks, THREADS, 0, s>>>(nq, Q, nt, T, dist, idx, chunks);
  return dpf_check_launch("nn_search_kernel");
}

int nn_one_direction(int b, int nq, const float* Q, int nt, const float* T, float* dist, int* idx,
                     cudaStream_t s) {
  if (b == 0 || nq == 0) return DPF_OK;
  // Pick the register blocking so that the grid covers the chip at least ~2x.
  const long long target = 2LL * dpf_num_sms();
  if ((long long)b * ((nq + 1023) / 1024) >= target) return launch_nn<8, 128>(b, nq, Q, nt, T, dist, idx, s);
  if ((long long)b * ((nq + 511) / 512) >= target) return launch_nn<4, 128>(b, nq, Q, nt, T, dist, idx, s);
  if ((long long)b * ((nq + 255) / 256) >= target) return launch_nn<2, 128>(b, nq, Q, nt, T, dist, idx, s);
  return launch_nn<1, 128>(b, nq, Q, nt, T, dist, idx, s);
}

}  // namespace

// dpf_set_option(4, v): 1 (default) = the one-evaluation all-pairs kernel, 0 = the two-direction kernel (tests compare both)
int g_fused_pairwise = 1;

// Replaces nndistance() (nndistance.cuh:1, nndistance.cu:125-128).
DPF_API int dpf_nndistance(int b, int n, const float* xyz, int m, const float* xyz2, float* result,
                           int* result_i, float* result2, int* result2_i, void* stream) {
  DPF_REQUIRE(b >= 0 && n >= 0 && m >= 0, DPF_ERR_BAD_ARG, "dpf_nndistance: negative size");
  if (b == 0) return DPF_OK;
  DPF_REQUIRE((n == 0 || (xyz && result && result_i)) && (m == 0 || (xyz2 && result2 && result2_i)),
              DPF_ERR_NULL_PTR, "dpf_nndistance: null pointer");
  DPF_REQUIRE(n == 0 || m > 0, DPF_ERR_BAD_ARG, "dpf_nndistance: empty target cloud (m == 0)");
  DPF_REQUIRE(m == 0 || n > 0, DPF_ERR_BAD_ARG, "dpf_nndistance: empty target cloud (n == 0)");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = nn_one_direction(b, n, xyz, m, xyz2, result, result_i, s);
  if (rc) return rc;
  return nn_one_direction(b, m, xyz2, n, xyz, result2, result2_i, s);
}

// Replaces nndistancegrad() (nndistance.cuh:2, nndistance.cu:149-154); the zero-fill runs on the
// caller's stream (the reference memsets on the null stream).
DPF_API int dpf_nndistance_grad(int b, int n, const float* xyz1, int m, const float* xyz2,
                                const float* grad_dist1, const int* idx1, const float* grad_dist2,
                                const int* idx2, float* grad_xyz1, float* grad_xyz2, void* stream) {
  DPF_REQUIRE(b >= 0 && n >= 0 && m >= 0, DPF_ERR_BAD_ARG, "dpf_nndistance_grad: negative size");
  if (b == 0 || (n == 0 && m == 0)) return DPF_OK;
  DPF_REQUIRE(xyz1 && xyz2 && grad_dist1 && idx1 && grad_dist2 && idx2 && grad_xyz1 && grad_xyz2,
              DPF_ERR_NULL_PTR, "dpf_nndistance_grad: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(grad_xyz1, 0, (size_t)b * n * 3 * sizeof(float), s);
  cudaMemsetAsync(grad_xyz2, 0, (size_t)b * m * 3 * sizeof(float), s);
  const int threads = 256;
  auto blocks = [&](size_t total) {
    return (int)min((size_t)dpf_num_sms() * 8, (total + threads - 1) / threads);
  };
  if (n > 0) nn_grad_kernel<<<blocks((size_t)b * n), threads, 0, s>>>(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_xyz1, grad_xyz2);
  if (m > 0) nn_grad_kernel<<<blocks((size_t)b * m), threads, 0, s>>>(b, m, xyz2, n, xyz1, grad_dist2, idx2, grad_xyz2, grad_xyz1);
  return dpf_check_launch("nn_grad_kernel");
}

// Fused replacement of pairwise_CD (lib/networks/utils.py:90-117): writes rows
// row_start + t*row_step (t < n_rows) of the (S1 x S2) matrix; with symmetric != 0 (A == B)
// only the upper triangle j >= i of those rows is written (mirror with dpf_symmetrize_upper).
DPF_API int dpf_pairwise_cd(int S1, int S2, int n, int m, const float* A, const float* B, float* out,
                            int row_start, int row_step, int n_rows, int symmetric, void* stream) {
  DPF_REQUIRE(S1 >= 0 && S2 >= 0 && n > 0 && m > 0, DPF_ERR_BAD_ARG, "dpf_pairwise_cd: bad sizes");
  DPF_REQUIRE(row_start >= 0 && row_step >= 1 && n_rows >= 0, DPF_ERR_BAD_ARG, "dpf_pairwise_cd: bad row range");
  if (n_rows == 0 || S2 == 0) return DPF_OK;
  DPF_REQUIRE(row_start + (long long)(n_rows - 1) * row_step < S1, DPF_ERR_BAD_ARG, "dpf_pairwise_cd: rows exceed S1");
  DPF_REQUIRE(A && B && out, DPF_ERR_NULL_PTR, "dpf_pairwise_cd: null pointer");
  DPF_REQUIRE(!symmetric || (S1 == S2 && n == m), DPF_ERR_BAD_ARG, "dpf_pairwise_cd: symmetric needs S1==S2, n==m");
  // Column blocking: enough CTAs for >= 4 waves, each CTA amortising its row cloud over JB columns.
  int JB = 8;
  while (JB > 1 && (long long)n_rows * ((S2 + JB - 1) / JB) < 8LL * dpf_num_sms()) JB >>= 1;
  const int jblocks = (S2 + JB - 1) / JB;
  const long long grid = (long long)n_rows * jblocks;
  DPF_REQUIRE(grid < 2147483647LL, DPF_ERR_BAD_ARG, "dpf_pairwise_cd: grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 2048 && g_fused_pairwise) {        // one distance evaluation per point pair for both directions
    constexpr int T = 256;
    auto launch = [&](auto kern, int R) {
      const size_t smem = sizeof(float4) * kTile + sizeof(float) * ((size_t)(T / 32) * kTile + (size_t)T * R);
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<(int)grid, T, smem, st>>>(S2, n, m, A, B, out, row_start, row_step, jblocks, JB, symmetric);
    };
    if (n <= 512) launch(pairwise_cd_fused_kernel<2, T>, 2);
    else if (n <= 1024) launch(pairwise_cd_fused_kernel<4, T>, 4);
    else launch(pairwise_cd_fused_kernel<8, T>, 8);
    return dpf_check_launch("pairwise_cd_fused_kernel");
  }
  pairwise_cd_kernel<8, 256><<<(int)grid, 256, 0, st>>>(S2, n, m, A, B, out, row_start, row_step, jblocks, JB, symmetric);
  return dpf_check_launch("pairwise_cd_kernel");
}

DPF_API int dpf_symmetrize_upper(float* M, int S, void* stream) {
  DPF_REQUIRE(S >= 0, DPF_ERR_BAD_ARG, "dpf_symmetrize_upper: bad size");
  if (S == 0) return DPF_OK;
  DPF_REQUIRE(M, DPF_ERR_NULL_PTR, "dpf_symmetrize_upper: null pointer");
  dim3 grid((S + 255) / 256, S);
  symmetrize_upper_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, S);
  return dpf_check_launch("symmetrize_upper_kernel");
}
