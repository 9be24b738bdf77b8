// GPU-side reductions behind the generation scores: everything the reference computes on the
// all-pairs matrices with torch ops + host syncs (lib/networks/utils.py:120-144: COV = unique argmins,
// MMD = mean of column minima, KNN = leave-one-out 1-NN accuracy on the (S1+S2)^2 block matrix) in ONE
// launch over the three device-resident Chamfer matrices, and the voxel-occupancy histogram of JSD
// (utils.py:45-79) as a GPU histogram.  No matrix ever goes back to the host; the caller reads 4 numbers.
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int SC_THREADS = 128;

struct MinIdx { float v; int i; };

__device__ __forceinline__ MinIdx better(MinIdx a, MinIdx b) {   // lowest value, then lowest index
  return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

__device__ __forceinline__ MinIdx block_min(MinIdx m, MinIdx* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MinIdx t;
    t.v = __shfl_xor_sync(0xffffffffu, m.v, o);
    t.i = __shfl_xor_sync(0xffffffffu, m.i, o);
    m = better(m, t);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = m;
  __syncthreads();
  MinIdx r = red[0];
#pragma unroll
  for (int w = 1; w < SC_THREADS / 32; ++w) r = better(r, red[w]);
  return r;
}

// scan `cnt` values base[k * stride], skipping index `skip` (-1 = none)
__device__ __forceinline__ MinIdx scan_min(const float* __restrict__ base, int cnt, size_t stride, int skip) {
  MinIdx m{CUDART_INF_F, 0x7fffffff};
  for (int k = threadIdx.x; k < cnt; k += SC_THREADS) {
    if (k == skip) continue;
    const float v = base[(size_t)k * stride];
    if (v < m.v) { m.v = v; m.i = k; }     // ascending k per thread: strict '<' keeps the lowest index
  }
  return m;
}

// One CTA per item of the concatenated set [S1 generated | S2 reference].
//   gen i : nearest neighbour over gg[i, j != i] (label gen) and gt[i, :] (label ref); COV marks argmin_j gt[i, j]
//   ref j : nearest neighbour over gt[:, j] (label gen) and tt[j, k != j] (label ref); MMD takes min_i gt[i, j]
// correct[item] = 1 when the 1-NN carries the item's own label (ties: the lower index of the concatenated order,
// i.e. the generated block first).
__global__ void __launch_bounds__(SC_THREADS)
cd_scores_kernel(const float* __restrict__ gg, const float* __restrict__ gt, const float* __restrict__ tt, int S1, int S2,
                 float* __restrict__ correct, float* __restrict__ col_min, int* __restrict__ cov_flag) {
  __shared__ MinIdx red[SC_THREADS / 32];
  const int item = blockIdx.x;
  if (item < S1) {
    const int i = item;
    const MinIdx a = block_min(scan_min(gg + (size_t)i * S1, S1, 1, i), red);
    const MinIdx b = block_min(scan_min(gt + (size_t)i * S2, S2, 1, -1), red);
    if (threadIdx.x == 0) {
      correct[item] = (a.v <= b.v && a.i != 0x7fffffff) ? 1.f : 0.f;   // same-label neighbour at least as close (gen block comes first)
      if (b.i != 0x7fffffff) cov_flag[b.i] = 1;
    }
  } else {
    const int j = item - S1;
    const MinIdx a = block_min(scan_min(gt + j, S1, (size_t)S2, -1), red);
    const MinIdx b = block_min(scan_min(tt + (size_t)j * S2, S2, 1, j), red);
    if (threadIdx.x == 0) {
      correct[item] = (b.v < a.v || a.i == 0x7fffffff) ? 1.f : 0.f;    // ties go to the generated block (lower index)
      col_min[j] = a.v;
    }
  }
}

// out[0] = COV (fraction of references matched), out[1] = MMD (mean of column minima), out[2] = 1-NN accuracy
__global__ void __launch_bounds__(256)
cd_scores_final_kernel(const float* __restrict__ correct, const float* __restrict__ col_min, const int* __restrict__ cov_flag,
                       int S1, int S2, float* __restrict__ out) {
  __shared__ double red[3][8];
  double c = 0.0, m = 0.0, f = 0.0;
  for (int k = threadIdx.x; k < S1 + S2; k += 256) c += (double)correct[k];
  for (int k = threadIdx.x; k < S2; k += 256) {
    m += (double)col_min[k];
    f += (double)cov_flag[k];
  }
  c = warp_sum_d(c); m = warp_sum_d(m); f = warp_sum_d(f);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = c; red[1][threadIdx.x >> 5] = m; red[2][threadIdx.x >> 5] = f; }
  __syncthreads();
  if (threadIdx.x == 0) {
    c = m = f = 0.0;
    for (int w = 0; w < 8; ++w) { c += red[0][w]; m += red[1][w]; f += red[2][w]; }
    out[0] = (float)(f / (double)S2);
    out[1] = (float)(m / (double)S2);
    out[2] = (float)(c / (double)(S1 + S2));
  }
}

// Voxel-occupancy counts: a point lands in cell (i,j,k) when edges[i] <= x < edges[i+1] (double comparison of the
// fp32 coordinate against the caller's double edges, exactly the reference's numpy test) for all three
// coordinates; anything outside (or NaN) is dropped.
__global__ void __launch_bounds__(256)
voxel_hist_kernel(const float* __restrict__ pts, long long n_points, int res, const double* __restrict__ edges,
                  unsigned long long* __restrict__ hist) {
  extern __shared__ double e[];
  for (int i = threadIdx.x; i <= res; i += 256) e[i] = edges[i];
  __syncthreads();
  const double lo = e[0], inv = (double)res / (e[res] - e[0]);
  for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < n_points; p += (long long)gridDim.x * 256) {
    int cell[3];
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double x = (double)pts[p * 3 + c];
      int i = (int)floor((x - lo) * inv);
      if (!(x >= lo) || !(x < e[res])) { ok = false; i = 0; }    // also rejects NaN
      i = max(0, min(res - 1, i));
      while (ok && i > 0 && x < e[i]) --i;                           // the closed form can be one cell off at an edge
      while (ok && i < res - 1 && x >= e[i + 1]) ++i;
      cell[c] = i;
    }
    if (ok) atomicAdd(&hist[((size_t)cell[0] * res + cell[1]) * res + cell[2]], 1ULL);
  }
}

}  // namespace

// Generation scores from the three all-pairs matrices (device, fp32, row-major; gg and tt must be full symmetric
// matrices): out[0] = COV(gt) (utils.py:120-121), out[1] = MMD(gt) (:124-125), out[2] = KNN(gg, gt, tt, 1) (:128-144).
// scratch: (S1 + S2) floats + S2 floats + S2 ints = dpf_cd_scores_scratch_bytes.
DPF_API int dpf_cd_scores_scratch_bytes(int S1, int S2, long long* bytes) {
  DPF_REQUIRE(bytes && S1 >= 0 && S2 >= 0, DPF_ERR_BAD_ARG, "dpf_cd_scores_scratch_bytes: bad arguments");
  *bytes = (long long)sizeof(float) * ((long long)S1 + 3LL * S2) + 256;
  return DPF_OK;
}

DPF_API int dpf_cd_scores(const float* gg, const float* gt, const float* tt, int S1, int S2, void* scratch, float* out,
                          void* stream) {
  DPF_REQUIRE(S1 > 0 && S2 > 0, DPF_ERR_BAD_ARG, "dpf_cd_scores: both sets must be non-empty");
  DPF_REQUIRE(gg && gt && tt && scratch && out, DPF_ERR_NULL_PTR, "dpf_cd_scores: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  float* correct = (float*)scratch;
  float* col_min = correct + S1 + S2;
  int* cov_flag = (int*)(col_min + S2);
  cudaMemsetAsync(cov_flag, 0, sizeof(int) * (size_t)S2, s);
  cd_scores_kernel<<<S1 + S2, SC_THREADS, 0, s>>>(gg, gt, tt, S1, S2, correct, col_min, cov_flag);
  int rc = dpf_check_launch("cd_scores_kernel");
  if (rc) return rc;
  cd_scores_final_kernel<<<1, 256, 0, s>>>(correct, col_min, cov_flag, S1, S2, out);
  return dpf_check_launch("cd_scores_final_kernel");
}

// Voxel-occupancy histogram of get_voxel_occ_dist (utils.py:45-79): pts (n_points,3) fp32, edges (res+1) doubles on the
// device, hist (res^3) uint64 counts (zeroed here).
DPF_API int dpf_voxel_hist(const float* pts, long long n_points, int res, const double* edges, unsigned long long* hist,
                           void* stream) {
  DPF_REQUIRE(n_points >= 0 && res > 0 && res <= 1024, DPF_ERR_BAD_ARG, "dpf_voxel_hist: bad sizes");
  DPF_REQUIRE(edges && hist && (pts || n_points == 0), DPF_ERR_NULL_PTR, "dpf_voxel_hist: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * (size_t)res * res * res, s);
  if (n_points == 0) return DPF_OK;
  const long long want = (n_points + 255) / 256;
  const int grid = (int)(want < 8LL * dpf_num_sms() ? want : 8LL * dpf_num_sms());
  voxel_hist_kernel<<<grid, 256, sizeof(double) * (res + 1), s>>>(pts, n_points, res, edges, hist);
  return dpf_check_launch("voxel_hist_kernel");
}
