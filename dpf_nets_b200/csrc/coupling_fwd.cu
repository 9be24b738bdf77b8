// Forward kernels of the conditional coupling stack, fp32 CUDA-core ("exact") path, plus the
// FiLM pre-compute and the input-moment reduction shared with the tensor-core path.
//
// One thread owns one point (tile = 128 consecutive points of one shape): the thread carries the
// 64 hidden activations in registers, the 64x64 weights are broadcast from shared memory, and the
// FiLM / ReLU / last SharedDot / softsign / exp / sqrt / affine transform run in the same thread
// without ever writing an activation to HBM.  Train-mode BatchNorm:
//   BN_a (after the first SharedDot) is folded analytically from the mean / covariance of the kept
//        coordinates (the first SharedDot is linear), accumulated by the previous step's epilogue;
//   BN_b (after the 64x64 SharedDot) needs one statistics pass (STATS=true) before the apply pass.
#include "coupling.cuh"

namespace {

constexpr int F = DPF_F;

// ------------------------------------------------------------------------------------------
// FiLM nets of ALL layers in one launch: grid = (L*4, b-chunks); net = (branch, kind w|b).
//   out = Lin1( Swish( BN_batch( Lin0 g ) ) ) ; w-nets store s = eps + exp(out).
// flows.py:33-45,68-80,100-106.
// ------------------------------------------------------------------------------------------
constexpr int FILM_THREADS = 256;
constexpr int FILM_BCHUNK = 256;

__global__ void __launch_bounds__(FILM_THREADS)
film_forward_kernel(const float* __restrict__ arena, float* __restrict__ stats, const LayerMeta* __restrict__ meta,
                    const float* __restrict__ g, float* __restrict__ film, int B, int G, int training,
                    int update_stats, float eps, int cap /* rows the shared-memory tiles are sized for */) {
  extern __shared__ float sm[];
  const int l = blockIdx.x >> 2, net = blockIdx.x & 3, br = net >> 1, kind = net & 1;
  const int b0 = blockIdx.y * FILM_BCHUNK;
  const int nb = min(FILM_BCHUNK, B - b0);
  const LayerMeta m = meta[l];
  const BranchLayout lay = branch_layout((int)m.k, (int)m.w, G);
  const float* prm = arena + m.param_off + (size_t)br * lay.size;
  float* st = stats + m.stat_off + (size_t)br * ST_COUNT * F;
  const float* W0 = prm + (kind ? lay.fb0_W : lay.fw0_W);
  const float* bnw = prm + (kind ? lay.fb0_bnw : lay.fw0_bnw);
  const float* bnb = prm + (kind ? lay.fb0_bnb : lay.fw0_bnb);
  const float* W1 = prm + (kind ? lay.fb1_W : lay.fw1_W);
  const float* b1 = prm + (kind ? lay.fb1_b : lay.fw1_b);
  float* rm = st + (kind ? ST_FB_RM : ST_FW_RM) * F;
  float* rv = st + (kind ? ST_FB_RV : ST_FW_RV) * F;

  float* u = sm;                         // [nb][F]
  float* Ws = u + (size_t)cap * F;       // [32][F+1]  (later reused as W1s [F][F+1])
  float* gs = Ws + F * (F + 1);          // transposed g chunk [32][4][8] (at least 1024 floats)
  float* mean_s = gs + ((size_t)cap * 33 > 1024 ? (size_t)cap * 33 : 1024);
  float* istd_s = mean_s + F;
  const int tid = threadIdx.x;
  const int c = tid & 63, bq = tid >> 6;

  // u[b][c] = sum_i g[b][i] W0[c][i]: register tile of 8 shapes per thread (shapes bq + 4 j): one weight load and two
  // 16-byte broadcast loads of the transposed g chunk feed 8 FMAs (the one-shape-per-thread form was shared-memory-load bound)
  const int nj = (nb + 3) >> 2;               // shapes per bq group
  for (int jb = 0; jb < nj; jb += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int i0 = 0; i0 < G; i0 += 32) {
      const int ni = min(32, G - i0);
      __syncthreads();
      for (int e = tid; e < F * 32; e += FILM_THREADS) {
        const int cc = e >> 5, i = e & 31;
        Ws[i * (F + 1) + cc] = (i < ni) ? __ldg(W0 + (size_t)cc * G + i0 + i) : 0.f;
      }
      // gs (reused as gT): [i][q][8] = g[b0 + q + 4 (jb + j)][i0 + i], zero-padded
      for (int e = tid; e < 32 * 32; e += FILM_THREADS) {
        const int i = e & 31, qj = e >> 5, q = qj >> 3, j = qj & 7;
        const int b = q + 4 * (jb + j);
        gs[i * 32 + qj] = (i < ni && b < nb) ? __ldg(g + (size_t)(b0 + b) * G + i0 + i) : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        const float w = Ws[i * (F + 1) + c];
        const float4 ga = *reinterpret_cast<const float4*>(gs + i * 32 + bq * 8);
        const float4 gb = *reinterpret_cast<const float4*>(gs + i * 32 + bq * 8 + 4);
        acc[0] = fmaf(ga.x, w, acc[0]); acc[1] = fmaf(ga.y, w, acc[1]); acc[2] = fmaf(ga.z, w, acc[2]); acc[3] = fmaf(ga.w, w, acc[3]);
        acc[4] = fmaf(gb.x, w, acc[4]); acc[5] = fmaf(gb.y, w, acc[5]); acc[6] = fmaf(gb.z, w, acc[6]); acc[7] = fmaf(gb.w, w, acc[7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = bq + 4 * (jb + j);
      if (b < nb) u[b * F + c] = acc[j];
    }
  }
  __syncthreads();
  if (tid < F) {
    float mean, var;
    if (training) {
      float s = 0.f;
      for (int b = 0; b < nb; ++b) s += u[b * F + tid];
      mean = s / (float)nb;
      float q = 0.f;
      for (int b = 0; b < nb; ++b) {
        const float d = u[b * F + tid] - mean;
        q = fmaf(d, d, q);
      }
      var = q / (float)nb;
      if (update_stats) {
        rm[tid] = (1.f - DPF_BN_MOM) * rm[tid] + DPF_BN_MOM * mean;
        rv[tid] = (1.f - DPF_BN_MOM) * rv[tid] + DPF_BN_MOM * var * ((float)nb / (float)max(nb - 1, 1));
      }
    } else {
      mean = rm[tid];
      var = rv[tid];
    }
    mean_s[tid] = mean;
    istd_s[tid] = 1.f / sqrtf(var + DPF_BN_EPS);
  }
  __syncthreads();
  {
    const float gam = __ldg(bnw + c), bet = __ldg(bnb + c), mean = mean_s[c], istd = istd_s[c];
    for (int b = bq; b < nb; b += FILM_THREADS / 64) {
      const float yv = fmaf((u[b * F + c] - mean) * istd, gam, bet);
      u[b * F + c] = yv * (1.f / (1.f + expf(-yv)));
    }
  }
  for (int e = tid; e < F * F; e += FILM_THREADS) {
    const int co = e >> 6, ci = e & 63;
    Ws[co * (F + 1) + ci] = __ldg(W1 + e);
  }
  __syncthreads();
  {
    const float bias = __ldg(b1 + c);
    float* out = film + ((size_t)(l * 4 + net) * B + b0) * F;
    for (int b = bq; b < nb; b += FILM_THREADS / 64) {
      float acc = bias;
#pragma unroll 16
      for (int ci = 0; ci < F; ++ci) acc = fmaf(Ws[c * (F + 1) + ci], u[b * F + ci], acc);
      out[(size_t)b * F + c] = kind ? acc : (eps + expf(acc));
    }
  }
}

// ------------------------------------------------------------------------------------------
// sum x_c and sum x_c x_c' over all B*N points (double accumulation) -> mom[0..8].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
moments_kernel(const float* __restrict__ x, int B, int N, double* __restrict__ mom) {
  double a[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) a[i] = 0.0;
  const long long total = (long long)B * N;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / N;
    const int n = (int)(e - b * N);
    const float* px = x + (size_t)b * 3 * N + n;
    const double x0 = px[0], x1 = px[N], x2 = px[2 * (size_t)N];
    a[0] += x0; a[1] += x1; a[2] += x2;
    a[3] += x0 * x0; a[4] += x0 * x1; a[5] += x0 * x2;
    a[6] += x1 * x1; a[7] += x1 * x2; a[8] += x2 * x2;
  }
  __shared__ double red[9][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const double v = warp_sum_d(a[i]);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    atomicAdd(mom + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------------------------------
// Per-layer forward, fp32 path.  MODE 0 = 'direct' (sampling), 1 = 'inverse' (training NLL).
// STATS=true: only accumulates sum / sumsq of h2pre per channel (BN_b batch statistics).
// ------------------------------------------------------------------------------------------
struct FwdSmem {
  float W1[2][F * F];       // [branch][c][j]
  float4 A0[2][F];          // {A00, A01, c0, -}
  float S[2][F], T[2][F];   // per-tile FiLM x BN_b fold: a = S*h2pre + T
  float mb[2][F], ib[2][F];
  float W2[2][2][F];
  float b2[2][2];
  float red[2][F][2];       // STATS: per-CTA sum / sumsq
  double mom[9][4];
};

template <int K, int MODE, bool STATS>
__global__ void __launch_bounds__(DPF_TILE)
coupling_fwd_fp32_kernel(const CouplingArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  FwdSmem& s = *reinterpret_cast<FwdSmem*>(smraw);
  const int tid = threadIdx.x;
  const BranchLayout lay = branch_layout(a.k, a.w, a.G);
  const bool writer = (blockIdx.x == 0) && a.update_stats && !STATS;

  for (int e = tid; e < 2 * F * F; e += DPF_TILE) {
    const int br = e / (F * F), r = e - br * F * F;
    s.W1[br][r] = a.prm[(size_t)br * lay.size + lay.W1 + r];
  }
  {
    const int br = tid >> 6, c = tid & 63;
    float A00, A01, c0;
    fold_bn_a(a, lay, br, c, writer, A00, A01, c0, nullptr, nullptr);
    s.A0[br][c] = make_float4(A00, A01, c0, 0.f);
    if (!STATS) {
      float mean, istd;
      bn_b_stats(a, br, c, writer, mean, istd);
      s.mb[br][c] = mean;
      s.ib[br][c] = istd;
      const float* prm = a.prm + (size_t)br * lay.size;
      s.W2[br][0][c] = prm[lay.W2 + c];
      s.W2[br][1][c] = (a.w == 2) ? prm[lay.W2 + F + c] : 0.f;
      if (c < 2) s.b2[br][c] = (c < a.w) ? prm[lay.b2 + c] : 0.f;
    } else {
      s.red[br][c][0] = 0.f;
      s.red[br][c][1] = 0.f;
    }
  }
  float macc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) macc[i] = 0.f;
  __syncthreads();

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_b;
    const int n = (tile - b * a.tiles_per_b) * DPF_TILE + tid;
    const bool valid = n < a.N;
    if (!STATS) {
      __syncthreads();  // previous tile done with S/T
      const int br = tid >> 6, c = tid & 63;
      const float sc = a.film[((size_t)(br * 2 + 0) * a.B + b) * F + c];
      const float sh = a.film[((size_t)(br * 2 + 1) * a.B + b) * F + c];
      const float S = sc * s.ib[br][c];
      s.S[br][c] = S;
      s.T[br][c] = fmaf(-S, s.mb[br][c], sh);
      __syncthreads();
    }
    const float* px = a.x + (size_t)b * 3 * a.N + n;
    float xin[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) xin[ch] = valid ? px[(size_t)ch * a.N] : 0.f;
    const float xk0 = xin[a.keep0];
    const float xk1 = (K == 2) ? xin[a.keep1] : 0.f;
    float o[2][2];
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      float h1[F];
#pragma unroll
      for (int c = 0; c < F; ++c) {
        const float4 A = s.A0[br][c];
        float v = fmaf(A.x, xk0, A.z);
        if (K == 2) v = fmaf(A.y, xk1, v);
        h1[c] = fmaxf(v, 0.f);
      }
      float o0 = STATS ? 0.f : s.b2[br][0], o1 = STATS ? 0.f : s.b2[br][1];
#pragma unroll 2
      for (int c = 0; c < F; ++c) {
        const float4* wrow = reinterpret_cast<const float4*>(&s.W1[br][c * F]);
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int j = 0; j < F / 4; j += 2) {
          const float4 wa = wrow[j], wb = wrow[j + 1];
          acc0 = fmaf(wa.x, h1[4 * j + 0], acc0);
          acc0 = fmaf(wa.y, h1[4 * j + 1], acc0);
          acc0 = fmaf(wa.z, h1[4 * j + 2], acc0);
          acc0 = fmaf(wa.w, h1[4 * j + 3], acc0);
          acc1 = fmaf(wb.x, h1[4 * j + 4], acc1);
          acc1 = fmaf(wb.y, h1[4 * j + 5], acc1);
          acc1 = fmaf(wb.z, h1[4 * j + 6], acc1);
          acc1 = fmaf(wb.w, h1[4 * j + 7], acc1);
        }
        const float acc = acc0 + acc1;
        if (STATS) {
          const float v = valid ? acc : 0.f;
          const float s1 = warp_sum(v), s2 = warp_sum(v * v);
          if ((tid & 31) == 0) {
            atomicAdd(&s.red[br][c][0], s1);
            atomicAdd(&s.red[br][c][1], s2);
          }
        } else {
          const float h3 = fmaxf(fmaf(s.S[br][c], acc, s.T[br][c]), 0.f);
          o0 = fmaf(s.W2[br][0][c], h3, o0);
          o1 = fmaf(s.W2[br][1][c], h3, o1);
        }
      }
      o[br][0] = o0;
      o[br][1] = o1;
    }
    if (!STATS) {
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
          const float xv = xin[ch];
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
  }
  if (STATS) {
    __syncthreads();
    const int br = tid >> 6, c = tid & 63;
    atomicAdd(&a.bnb_sums[(br * F + c) * 2 + 0], (double)s.red[br][c][0]);
    atomicAdd(&a.bnb_sums[(br * F + c) * 2 + 1], (double)s.red[br][c][1]);
  } else if (a.mom_out) {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double v = warp_sum_d((double)macc[i]);
      if (lane == 0) s.mom[i][warp] = v;
    }
    __syncthreads();
    if (tid < 9) atomicAdd(a.mom_out + tid, s.mom[tid][0] + s.mom[tid][1] + s.mom[tid][2] + s.mom[tid][3]);
  }
}

}  // namespace

// ---- launchers used by decoder.cu ----------------------------------------------------------
int launch_film_forward(const float* arena, float* stats, const LayerMeta* meta_dev, const float* g, float* film,
                        int L, int B, int G, int training, int update_stats, float eps, cudaStream_t s) {
  // shared memory sized for the actual batch: 30 KB at B = 32 -> all L*4 CTAs resident in one wave
  // (at the FILM_BCHUNK cap, 116 KB, it is one CTA per SM and 252 CTAs take two waves)
  const int cap = B < FILM_BCHUNK ? B : FILM_BCHUNK;
  const size_t smem = sizeof(float) * ((size_t)cap * F + F * (F + 1) + ((size_t)cap * 33 > 1024 ? (size_t)cap * 33 : 1024) + 2 * F);
  static bool attr = false;
  if (!attr) {
    const size_t smem_max = sizeof(float) * ((size_t)FILM_BCHUNK * F + F * (F + 1) + (size_t)FILM_BCHUNK * 33 + 2 * F);
    cudaFuncSetAttribute(film_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    attr = true;
  }
  dim3 grid(L * 4, (B + FILM_BCHUNK - 1) / FILM_BCHUNK);
  film_forward_kernel<<<grid, FILM_THREADS, smem, s>>>(arena, stats, meta_dev, g, film, B, G, training, update_stats, eps, cap);
  return dpf_check_launch("film_forward_kernel");
}

int launch_moments(const float* x, int B, int N, double* mom, cudaStream_t s) {
  const long long total = (long long)B * N;
  const int grid = (int)min((long long)dpf_num_sms() * 2, (total + 255) / 256);
  moments_kernel<<<grid, 256, 0, s>>>(x, B, N, mom);
  return dpf_check_launch("moments_kernel");
}

template <int K, int MODE, bool STATS>
static int launch_fwd_t(const CouplingArgs& a, int grid, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(coupling_fwd_fp32_kernel<K, MODE, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(FwdSmem));
    attr = true;
  }
  coupling_fwd_fp32_kernel<K, MODE, STATS><<<grid, DPF_TILE, sizeof(FwdSmem), s>>>(a);
  return dpf_check_launch("coupling_fwd_fp32_kernel");
}

int launch_coupling_fwd_fp32(const CouplingArgs& a, int mode, bool stats_pass, cudaStream_t s) {
  const int grid = min(a.n_tiles, dpf_num_sms() * 3);
  if (a.k == 2) {
    if (mode == 0) return stats_pass ? launch_fwd_t<2, 0, true>(a, grid, s) : launch_fwd_t<2, 0, false>(a, grid, s);
    return stats_pass ? launch_fwd_t<2, 1, true>(a, grid, s) : launch_fwd_t<2, 1, false>(a, grid, s);
  }
  if (mode == 0) return stats_pass ? launch_fwd_t<1, 0, true>(a, grid, s) : launch_fwd_t<1, 0, false>(a, grid, s);
  return stats_pass ? launch_fwd_t<1, 1, true>(a, grid, s) : launch_fwd_t<1, 1, false>(a, grid, s);
}
