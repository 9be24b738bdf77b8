// Backward kernels of the conditional coupling stack, fp32 CUDA-core path.
//
// What autograd does for the reference restated as two fused passes per
// layer + one launch for all FiLM nets (algebra validated on CPU by
// oracle/flow_oracle.py::coupling_backward_two_pass against the reference's autograd):
//
//   pass 1  recompute h1 -> W1.h1 -> FiLM; from (x, y, logvar, dy) form do_mu, do_logvar;
//           reduce per (shape, channel) dt = sum_n da, ds = sum_n da*h2n   (FiLM grads, and the
//           BatchNorm_b batch terms m1, m2 follow from them), dW2, db2.
//   pass 2  recompute again, apply the BN_b backward with m1/m2, dgrad (rank-1 updates fused into
//           the same sweep over W1 rows), ReLU mask, per-point T1 = A0^T dz, reductions
//           dbeta = sum dz, E = sum dz x_keep^T, wgrad dW1 = sum dh2pre h1^T (smem-tiled),
//           writes the stored input gradient.
//   BN_a's batch terms never touch the points again: they collapse to an affine correction
//           dx[keep] -= cvec + Q x[keep] that the NEXT backward step applies while loading its dy
//           ("pending"), and to closed forms for dgamma, dbeta, dW0 (finalize, in pass-1 prologue).
#include "coupling.cuh"

namespace {

constexpr int F = DPF_F;
constexpr int PITCH = DPF_TILE + 4;  // smem tile pitch (floats): float4-aligned, conflict-light

__device__ __forceinline__ float pick3(const float v[3], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }

// Deferred BN_a correction owed by the layer processed before us in backward; also (writer CTA)
// finalizes that layer's dgamma / dbeta / dW0.  All 128 threads participate: thread = (branch, channel).
__device__ Pending compute_pending(const BwdArgs& a, bool writer, double* red /* [5][4] smem */) {
  Pending P{0.f, 0.f, 0.f, 0.f, 0.f};
  if (!a.has_pending) return P;
  const int tid = threadIdx.x, br = tid >> 6, c = tid & 63;
  const BranchLayout nlay = branch_layout(a.nk, a.nw, a.f.G);
  const double M = (double)a.f.B * (double)a.f.N;
  const BnA bn = bn_a_of(a.nprm, a.nstat, nlay, a.n_mom, M, a.nk, a.nkeep0, a.nkeep1, a.f.training, br, c);
  const double dbeta = a.n_bna_sums[(br * F + c) * 4 + 0];
  const double E0 = a.n_bna_sums[(br * F + c) * 4 + 1];
  const double E1 = a.n_bna_sums[(br * F + c) * 4 + 2];
  const double istd = bn.istd, mean = bn.mean, w0 = bn.w0, w1 = bn.w1, gam = bn.gamma;
  const double dgamma = istd * (w0 * E0 + w1 * E1 - mean * dbeta);
  const double n1 = a.f.training ? gam * dbeta / M : 0.0;
  const double n2 = a.f.training ? gam * dgamma / M : 0.0;
  double v[5];
  const double i2n2 = istd * istd * n2;
  v[0] = w0 * istd * n1 - i2n2 * mean * w0;   // cvec_j = u_j - r_j
  v[1] = w1 * istd * n1 - i2n2 * mean * w1;
  v[2] = i2n2 * w0 * w0;
  v[3] = i2n2 * w0 * w1;
  v[4] = i2n2 * w1 * w1;
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
    const double s = warp_sum_d(v[i]);
    if (lane == 0) red[i * 4 + warp] = s;
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

// Per-point loads shared by both passes: x, y, logvar, dy (chain + external + pending) and the
// conditioner output cotangents do_mu[wi], do_lv[wi].
struct PointGrad {
  float x[3], dy[3], sig[2], do_mu[2], do_lv[2];
};

template <int MODE>
__device__ __forceinline__ PointGrad load_point(const BwdArgs& a, const Pending& P, int b, int n, bool valid) {
  PointGrad g;
  const int N = a.f.N;
  const size_t base = (size_t)b * 3 * N + n;
  float y[3], lv[3], dmu[3], dlv[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const size_t o = base + (size_t)ch * N;
    g.x[ch] = valid ? a.f.x[o] : 0.f;
    y[ch] = valid ? a.yv[o] : 0.f;
    lv[ch] = valid ? a.lvv[o] : 0.f;
    float d = 0.f;
    if (valid && a.dy_chain) d += a.dy_chain[o];
    if (valid && a.dP) d += a.dP[o];
    g.dy[ch] = d;
    dmu[ch] = (valid && a.dMU) ? a.dMU[o] : 0.f;
    dlv[ch] = (valid && a.dLV) ? a.dLV[o] : 0.f;
  }
  if (a.has_pending && valid) {
    const float y0 = pick3(y, a.nkeep0);
    const float y1 = a.nk == 2 ? pick3(y, a.nkeep1) : 0.f;
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
      const float l = pick3(lv, ch), dyv = pick3(g.dy, ch);
      const float e = expf(l);
      const float s2 = a.f.eps + e;
      const float sg = sqrtf(s2);
      g.sig[wi] = sg;
      float dm, dl;
      if (MODE == 1) {  // inverse: y = (x - mu) / sigma
        dm = -dyv / sg;
        dl = -dyv * pick3(y, ch) * e / (2.f * s2);
      } else {          // direct: y = sigma x + mu
        dm = dyv;
        dl = dyv * pick3(g.x, ch) * e / (2.f * sg);
      }
      dm += pick3(dmu, ch);
      dl += pick3(dlv, ch);
      const float om = 1.f - fabsf(l);            // softsign'(o) = (1 - |softsign(o)|)^2
      g.do_mu[wi] = valid ? dm : 0.f;
      g.do_lv[wi] = valid ? dl * om * om : 0.f;
    }
  }
  return g;
}

struct BwdSmemCommon {
  float W1[2][F * F];
  float4 A0[2][F];
  float S[2][F], T[2][F], sraw[2][F];
  float mb[2][F], ib[2][F];
  float W2[2][2][F];
  double pend[20];
};

__device__ __forceinline__ void bwd_prologue(const BwdArgs& a, const BranchLayout& lay, BwdSmemCommon& s) {
  const int tid = threadIdx.x;
  for (int e = tid; e < 2 * F * F; e += DPF_TILE) {
    const int br = e / (F * F), r = e - br * F * F;
    s.W1[br][r] = a.f.prm[(size_t)br * lay.size + lay.W1 + r];
  }
  const int br = tid >> 6, c = tid & 63;
  float A00, A01, c0, mean, istd;
  fold_bn_a(a.f, lay, br, c, false, A00, A01, c0, nullptr, nullptr);
  s.A0[br][c] = make_float4(A00, A01, c0, 0.f);
  bn_b_stats(a.f, br, c, false, mean, istd);
  s.mb[br][c] = mean;
  s.ib[br][c] = istd;
  const float* prm = a.f.prm + (size_t)br * lay.size;
  s.W2[br][0][c] = prm[lay.W2 + c];
  s.W2[br][1][c] = (a.f.w == 2) ? prm[lay.W2 + F + c] : 0.f;
}

__device__ __forceinline__ void tile_film(const BwdArgs& a, BwdSmemCommon& s, int b) {
  const int tid = threadIdx.x, br = tid >> 6, c = tid & 63;
  const float sc = a.f.film[((size_t)(br * 2 + 0) * a.f.B + b) * F + c];
  const float sh = a.f.film[((size_t)(br * 2 + 1) * a.f.B + b) * F + c];
  const float S = sc * s.ib[br][c];
  s.sraw[br][c] = sc;
  s.S[br][c] = S;
  s.T[br][c] = fmaf(-S, s.mb[br][c], sh);
}

template <int K>
__device__ __forceinline__ void compute_h1(float h1[F], const float4* A0, float xk0, float xk1) {
#pragma unroll
  for (int c = 0; c < F; ++c) {
    const float4 A = A0[c];
    float v = fmaf(A.x, xk0, A.z);
    if (K == 2) v = fmaf(A.y, xk1, v);
    h1[c] = fmaxf(v, 0.f);
  }
}

__device__ __forceinline__ float dot_row(const float* wrow_f, const float h1[F]) {
  const float4* wrow = reinterpret_cast<const float4*>(wrow_f);
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
  return acc0 + acc1;
}

// ------------------------------------------------------------------------------------------
// Pass 1
// ------------------------------------------------------------------------------------------
struct P1Smem {
  BwdSmemCommon c;
  float tile_red[2][F][2];   // per tile: dt, ds
  float w2acc[2][2][F];      // per CTA: dW2
  float b2acc[2][2];
};

template <int K, int MODE>
__global__ void __launch_bounds__(DPF_TILE)
coupling_bwd_p1_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  P1Smem& s = *reinterpret_cast<P1Smem*>(smraw);
  const int tid = threadIdx.x, lane = tid & 31;
  const BranchLayout lay = branch_layout(a.f.k, a.f.w, a.f.G);
  bwd_prologue(a, lay, s.c);
  {
    const int br = tid >> 6, c = tid & 63;
    s.w2acc[br][0][c] = 0.f;
    s.w2acc[br][1][c] = 0.f;
    if (c < 2) s.b2acc[br][c] = 0.f;
  }
  __syncthreads();
  const Pending P = compute_pending(a, blockIdx.x == 0, s.c.pend);

  for (int tile = blockIdx.x; tile < a.f.n_tiles; tile += gridDim.x) {
    const int b = tile / a.f.tiles_per_b;
    const int n = (tile - b * a.f.tiles_per_b) * DPF_TILE + tid;
    const bool valid = n < a.f.N;
    __syncthreads();
    tile_film(a, s.c, b);
    {
      const int br = tid >> 6, c = tid & 63;
      s.tile_red[br][c][0] = 0.f;
      s.tile_red[br][c][1] = 0.f;
    }
    __syncthreads();
    const PointGrad g = load_point<MODE>(a, P, b, n, valid);
    const float xk0 = pick3(g.x, a.f.keep0);
    const float xk1 = (K == 2) ? pick3(g.x, a.f.keep1) : 0.f;
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      float h1[F];
      compute_h1<K>(h1, s.c.A0[br], xk0, xk1);
      const float d0 = br == 0 ? g.do_mu[0] : g.do_lv[0];
      const float d1 = br == 0 ? g.do_mu[1] : g.do_lv[1];
#pragma unroll 2
      for (int c = 0; c < F; ++c) {
        const float acc = dot_row(&s.c.W1[br][c * F], h1);
        const float h2n = (acc - s.c.mb[br][c]) * s.c.ib[br][c];
        const float av = fmaf(s.c.S[br][c], acc, s.c.T[br][c]);
        const float h3 = fmaxf(av, 0.f);
        const float da = av > 0.f ? fmaf(s.c.W2[br][0][c], d0, s.c.W2[br][1][c] * d1) : 0.f;
        const float r0 = warp_sum(da), r1 = warp_sum(da * h2n);
        const float r2 = warp_sum(d0 * h3);
        float r3 = 0.f;
        if (K == 1) r3 = warp_sum(d1 * h3);   // w == 2
        if (lane == 0) {
          atomicAdd(&s.tile_red[br][c][0], r0);
          atomicAdd(&s.tile_red[br][c][1], r1);
          atomicAdd(&s.w2acc[br][0][c], r2);
          if (K == 1) atomicAdd(&s.w2acc[br][1][c], r3);
        }
      }
      const float q0 = warp_sum(d0), q1 = warp_sum(d1);
      if (lane == 0) {
        atomicAdd(&s.b2acc[br][0], q0);
        atomicAdd(&s.b2acc[br][1], q1);
      }
    }
    __syncthreads();
    {
      const int br = tid >> 6, c = tid & 63;
      atomicAdd(&a.dfilm[((size_t)(br * 2 + 0) * a.f.B + b) * F + c], s.tile_red[br][c][1]);  // ds_raw
      atomicAdd(&a.dfilm[((size_t)(br * 2 + 1) * a.f.B + b) * F + c], s.tile_red[br][c][0]);  // dt
    }
  }
  __syncthreads();
  {
    const int br = tid >> 6, c = tid & 63;
    float* d = a.dprm + (size_t)br * lay.size;
    atomicAdd(&d[lay.W2 + c], s.w2acc[br][0][c]);
    if (a.f.w == 2) atomicAdd(&d[lay.W2 + F + c], s.w2acc[br][1][c]);
    if (c < a.f.w) atomicAdd(&d[lay.b2 + c], s.b2acc[br][c]);
  }
}

// ------------------------------------------------------------------------------------------
// Pass 2
// ------------------------------------------------------------------------------------------
struct P2Smem {
  BwdSmemCommon c;
  float m1[2][F], m2[2][F];
  float D[F * PITCH];        // dh2pre tile  [c][point]
  float H[F * PITCH];        // h1 tile      [j][point]
  float w1acc[2][F * F];     // per CTA: dW1
  float bna[2][F][3];        // per CTA: dbeta, E0, E1
};

template <int K, int MODE>
__global__ void __launch_bounds__(DPF_TILE)
coupling_bwd_p2_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  P2Smem& s = *reinterpret_cast<P2Smem*>(smraw);
  const int tid = threadIdx.x, lane = tid & 31;
  const BranchLayout lay = branch_layout(a.f.k, a.f.w, a.f.G);
  bwd_prologue(a, lay, s.c);
  {
    const int br = tid >> 6, c = tid & 63;
    float m1 = 0.f, m2 = 0.f;
    if (a.f.training) {
      double s1 = 0.0, s2 = 0.0;
      for (int b = 0; b < a.f.B; ++b) {
        const double sc = a.f.film[((size_t)(br * 2 + 0) * a.f.B + b) * F + c];
        s1 += sc * (double)a.dfilm[((size_t)(br * 2 + 1) * a.f.B + b) * F + c];   // s * dt
        s2 += sc * (double)a.dfilm[((size_t)(br * 2 + 0) * a.f.B + b) * F + c];   // s * ds_raw
      }
      const double M = (double)a.f.B * (double)a.f.N;
      m1 = (float)(s1 / M);
      m2 = (float)(s2 / M);
    }
    s.m1[br][c] = m1;
    s.m2[br][c] = m2;
    s.bna[br][c][0] = s.bna[br][c][1] = s.bna[br][c][2] = 0.f;
  }
  for (int e = tid; e < 2 * F * F; e += DPF_TILE) (&s.w1acc[0][0])[e] = 0.f;
  __syncthreads();
  const Pending P = compute_pending(a, false, s.c.pend);
  const float sig1 = sqrtf(a.f.eps + 1.0f);

  for (int tile = blockIdx.x; tile < a.f.n_tiles; tile += gridDim.x) {
    const int b = tile / a.f.tiles_per_b;
    const int n = (tile - b * a.f.tiles_per_b) * DPF_TILE + tid;
    const bool valid = n < a.f.N;
    __syncthreads();
    tile_film(a, s.c, b);
    __syncthreads();
    const PointGrad g = load_point<MODE>(a, P, b, n, valid);
    const float xk0 = pick3(g.x, a.f.keep0);
    const float xk1 = (K == 2) ? pick3(g.x, a.f.keep1) : 0.f;
    float T1_0 = 0.f, T1_1 = 0.f;
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      float h1[F], dh1[F];
      compute_h1<K>(h1, s.c.A0[br], xk0, xk1);
#pragma unroll
      for (int j = 0; j < F; ++j) {
        s.H[j * PITCH + tid] = h1[j];
        dh1[j] = 0.f;
      }
      const float d0 = br == 0 ? g.do_mu[0] : g.do_lv[0];
      const float d1 = br == 0 ? g.do_mu[1] : g.do_lv[1];
#pragma unroll 1
      for (int c = 0; c < F; ++c) {
        const float4* wrow = reinterpret_cast<const float4*>(&s.c.W1[br][c * F]);
        const float acc = dot_row(&s.c.W1[br][c * F], h1);
        const float h2n = (acc - s.c.mb[br][c]) * s.c.ib[br][c];
        const float av = fmaf(s.c.S[br][c], acc, s.c.T[br][c]);
        const float da = av > 0.f ? fmaf(s.c.W2[br][0][c], d0, s.c.W2[br][1][c] * d1) : 0.f;
        float dh2pre = s.c.ib[br][c] * (da * s.c.sraw[br][c] - s.m1[br][c] - h2n * s.m2[br][c]);
        dh2pre = valid ? dh2pre : 0.f;
        s.D[c * PITCH + tid] = dh2pre;
#pragma unroll
        for (int j = 0; j < F / 4; ++j) {   // dgrad: rank-1 update with the same W1 row
          const float4 wv = wrow[j];
          dh1[4 * j + 0] = fmaf(wv.x, dh2pre, dh1[4 * j + 0]);
          dh1[4 * j + 1] = fmaf(wv.y, dh2pre, dh1[4 * j + 1]);
          dh1[4 * j + 2] = fmaf(wv.z, dh2pre, dh1[4 * j + 2]);
          dh1[4 * j + 3] = fmaf(wv.w, dh2pre, dh1[4 * j + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < F; ++j) {
        const float dz = h1[j] > 0.f ? dh1[j] : 0.f;
        const float4 A = s.c.A0[br][j];
        T1_0 = fmaf(A.x, dz, T1_0);
        if (K == 2) T1_1 = fmaf(A.y, dz, T1_1);
        const float r0 = warp_sum(dz), r1 = warp_sum(dz * xk0);
        float r2 = 0.f;
        if (K == 2) r2 = warp_sum(dz * xk1);
        if (lane == 0) {
          atomicAdd(&s.bna[br][j][0], r0);
          atomicAdd(&s.bna[br][j][1], r1);
          if (K == 2) atomicAdd(&s.bna[br][j][2], r2);
        }
      }
      __syncthreads();
      {  // wgrad: dW1[c][j] += sum_p D[c][p] * H[j][p]; thread owns c = cb + 8*ci, j = jb + 16*ji
        const int cb = tid >> 4, jb = tid & 15;
        float acc[8][4];
#pragma unroll
        for (int ci = 0; ci < 8; ++ci)
#pragma unroll
          for (int ji = 0; ji < 4; ++ji) acc[ci][ji] = 0.f;
#pragma unroll 2
        for (int p4 = 0; p4 < DPF_TILE / 4; ++p4) {
          float4 dv[8], hv[4];
#pragma unroll
          for (int ci = 0; ci < 8; ++ci) dv[ci] = *reinterpret_cast<const float4*>(&s.D[(cb + 8 * ci) * PITCH + 4 * p4]);
#pragma unroll
          for (int ji = 0; ji < 4; ++ji) hv[ji] = *reinterpret_cast<const float4*>(&s.H[(jb + 16 * ji) * PITCH + 4 * p4]);
#pragma unroll
          for (int ci = 0; ci < 8; ++ci)
#pragma unroll
            for (int ji = 0; ji < 4; ++ji) {
              acc[ci][ji] = fmaf(dv[ci].x, hv[ji].x, acc[ci][ji]);
              acc[ci][ji] = fmaf(dv[ci].y, hv[ji].y, acc[ci][ji]);
              acc[ci][ji] = fmaf(dv[ci].z, hv[ji].z, acc[ci][ji]);
              acc[ci][ji] = fmaf(dv[ci].w, hv[ji].w, acc[ci][ji]);
            }
        }
#pragma unroll
        for (int ci = 0; ci < 8; ++ci)
#pragma unroll
          for (int ji = 0; ji < 4; ++ji) s.w1acc[br][(cb + 8 * ci) * F + jb + 16 * ji] += acc[ci][ji];
      }
      __syncthreads();
    }
    if (valid) {
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
  __syncthreads();
  {
    float* d = a.dprm;
    for (int e = tid; e < 2 * F * F; e += DPF_TILE) {
      const int br = e / (F * F), r = e - br * F * F;
      atomicAdd(&d[(size_t)br * lay.size + lay.W1 + r], s.w1acc[br][r]);
    }
    const int br = tid >> 6, c = tid & 63;
    atomicAdd(&a.bna_sums[(br * F + c) * 4 + 0], (double)s.bna[br][c][0]);
    atomicAdd(&a.bna_sums[(br * F + c) * 4 + 1], (double)s.bna[br][c][1]);
    atomicAdd(&a.bna_sums[(br * F + c) * 4 + 2], (double)s.bna[br][c][2]);
  }
}

// ------------------------------------------------------------------------------------------
// After the last backward step: finalize that layer's BN_a grads and (optionally) resolve the
// pending correction on the gradient w.r.t. the stack input.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DPF_TILE)
coupling_bwd_final_kernel(const BwdArgs a, const float* __restrict__ p_in, const float* __restrict__ dx_stored,
                          float* __restrict__ dp) {
  __shared__ double red[20];
  const Pending P = compute_pending(a, blockIdx.x == 0, red);
  if (!dp) return;
  const long long total = (long long)a.f.B * a.f.N;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / a.f.N;
    const int n = (int)(e - b * a.f.N);
    const size_t base = (size_t)b * 3 * a.f.N + n;
    float x[3], d[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      x[ch] = p_in[base + (size_t)ch * a.f.N];
      d[ch] = dx_stored[base + (size_t)ch * a.f.N];
    }
    const float y0 = pick3(x, a.nkeep0), y1 = a.nk == 2 ? pick3(x, a.nkeep1) : 0.f;
    const float corr0 = P.c0 + P.q00 * y0 + P.q01 * y1, corr1 = P.c1 + P.q01 * y0 + P.q11 * y1;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      if (ch == a.nkeep0) d[ch] -= corr0;
      if (a.nk == 2 && ch == a.nkeep1) d[ch] -= corr1;
      dp[base + (size_t)ch * a.f.N] = d[ch];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward of ALL FiLM nets in one launch: grid = L*4, one CTA per (layer, net).
// Consumes dfilm (ds_raw / dt), produces the nets' parameter grads and accumulates dg.
// ------------------------------------------------------------------------------------------
constexpr int FB_THREADS = 256;
constexpr int FB_MAXB = 128;
constexpr int GS_LD = 36;               // leading dimension of the g chunk rows (32 + padding, multiple of 4 floats)

__global__ void __launch_bounds__(FB_THREADS)
film_backward_kernel(const float* __restrict__ arena, const float* __restrict__ stats, float* __restrict__ darena,
                     const LayerMeta* __restrict__ meta, const float* __restrict__ g, const float* __restrict__ film,
                     const float* __restrict__ dfilm, float* __restrict__ dg, int B, int G, int training, float eps) {
  extern __shared__ float sm[];        // tiles sized for B rows (launch_film_backward)
  const int l = blockIdx.x >> 2, net = blockIdx.x & 3, br = net >> 1, kind = net & 1;
  const LayerMeta m = meta[l];
  const BranchLayout lay = branch_layout((int)m.k, (int)m.w, G);
  const float* prm = arena + m.param_off + (size_t)br * lay.size;
  float* dprm = darena + m.param_off + (size_t)br * lay.size;
  const float* st = stats + m.stat_off + (size_t)br * ST_COUNT * F;
  const int oW0 = kind ? lay.fb0_W : lay.fw0_W, obnw = kind ? lay.fb0_bnw : lay.fw0_bnw;
  const int obnb = kind ? lay.fb0_bnb : lay.fw0_bnb, oW1 = kind ? lay.fb1_W : lay.fw1_W, ob1 = kind ? lay.fb1_b : lay.fw1_b;
  const float* rm = st + (kind ? ST_FB_RM : ST_FW_RM) * F;
  const float* rv = st + (kind ? ST_FB_RV : ST_FW_RV) * F;

  float* XH = sm;                       // [B][F]  u, then xhat
  float* DO = XH + (size_t)B * F;       // [B][F]  cotangent of the net output
  float* DU = DO + (size_t)B * F;       // [B][F]  v (swish output), then dxhat, then du
  float* Ws = DU + (size_t)B * F;       // [F][F+1] / [32][F+1]
  float* gs = Ws + F * (F + 1);         // [B][GS_LD] g chunk, rows 16-byte aligned
  float* gT = gs + (size_t)B * GS_LD;   // [32][4][8] transposed g chunk (register-tiled recompute)
  float* vec = gT + 1024;               // mean, istd, m1, m2, dgam[4][F], dbet[4][F]
  float* mean_s = vec, *istd_s = vec + F, *m1_s = vec + 2 * F, *m2_s = vec + 3 * F;
  float* dgam_s = vec + 4 * F, *dbet_s = vec + 8 * F;
  const int tid = threadIdx.x, c = tid & 63, bq = tid >> 6;

  // ---- recompute u = g W0^T: register tile of 8 shapes per thread (shapes bq + 4 j), like film_forward_kernel ----
  {
    const int nj = (B + 3) >> 2;
    for (int jb = 0; jb < nj; jb += 8) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int i0 = 0; i0 < G; i0 += 32) {
        const int ni = min(32, G - i0);
        __syncthreads();
        for (int e = tid; e < F * 32; e += FB_THREADS) {
          const int cc = e >> 5, i = e & 31;
          Ws[i * (F + 1) + cc] = (i < ni) ? prm[oW0 + (size_t)cc * G + i0 + i] : 0.f;
        }
        for (int e = tid; e < 32 * 32; e += FB_THREADS) {        // gT[i][q][8] = g[q + 4 (jb + j)][i0 + i]
          const int i = e & 31, qj = e >> 5, q = qj >> 3, j = qj & 7;
          const int b = q + 4 * (jb + j);
          gT[i * 32 + qj] = (i < ni && b < B) ? g[(size_t)b * G + i0 + i] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
          const float w = Ws[i * (F + 1) + c];
          const float4 ga = *reinterpret_cast<const float4*>(gT + i * 32 + bq * 8);
          const float4 gb = *reinterpret_cast<const float4*>(gT + i * 32 + bq * 8 + 4);
          acc[0] = fmaf(ga.x, w, acc[0]); acc[1] = fmaf(ga.y, w, acc[1]); acc[2] = fmaf(ga.z, w, acc[2]); acc[3] = fmaf(ga.w, w, acc[3]);
          acc[4] = fmaf(gb.x, w, acc[4]); acc[5] = fmaf(gb.y, w, acc[5]); acc[6] = fmaf(gb.z, w, acc[6]); acc[7] = fmaf(gb.w, w, acc[7]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int b = bq + 4 * (jb + j);
        if (b < B) XH[b * F + c] = acc[j];
      }
    }
  }
  __syncthreads();
  if (tid < F) {
    float mean, var;
    if (training) {
      float sacc = 0.f;
      for (int b = 0; b < B; ++b) sacc += XH[b * F + tid];
      mean = sacc / (float)B;
      float q = 0.f;
      for (int b = 0; b < B; ++b) {
        const float d = XH[b * F + tid] - mean;
        q = fmaf(d, d, q);
      }
      var = q / (float)B;
    } else {
      mean = rm[tid];
      var = rv[tid];
    }
    mean_s[tid] = mean;
    istd_s[tid] = 1.f / sqrtf(var + DPF_BN_EPS);
  }
  __syncthreads();
  const float gam = prm[obnw + c], bet = prm[obnb + c];
  for (int b = bq; b < B; b += 4) {
    const float xh = (XH[b * F + c] - mean_s[c]) * istd_s[c];
    XH[b * F + c] = xh;
    const float yv = fmaf(xh, gam, bet);
    DU[b * F + c] = yv / (1.f + expf(-yv));                       // v = swish(y)
    const float f = film[((size_t)(l * 4 + net) * B + b) * F + c];
    const float d = dfilm[((size_t)(l * 4 + net) * B + b) * F + c];
    DO[b * F + c] = kind ? d : d * (f - eps);                     // s = eps + exp(wraw) -> dwraw = ds (s - eps)
  }
  for (int e = tid; e < F * F; e += FB_THREADS) Ws[(e >> 6) * (F + 1) + (e & 63)] = prm[oW1 + e];
  __syncthreads();
  // ---- dW1[c'][cc] = sum_b DO[b][c'] v[b][cc] ; db1 ----
  {
    const int co = tid & 63, cq = tid >> 6;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    float accb = 0.f;
    for (int b = 0; b < B; ++b) {
      const float d = DO[b * F + co];
      accb += d;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(d, DU[b * F + cq * 16 + i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) dprm[oW1 + co * F + cq * 16 + i] = acc[i];
    if (cq == 0) dprm[ob1 + co] = accb;
  }
  __syncthreads();
  // ---- dv = DO W1 ; dy = dv * swish'(y) ; dgamma, dbeta ; dxhat -> DU ----
  {
    float dgacc = 0.f, dbacc = 0.f;
    for (int b = bq; b < B; b += 4) {
      float dv = 0.f;
#pragma unroll 16
      for (int co = 0; co < F; ++co) dv = fmaf(DO[b * F + co], Ws[co * (F + 1) + c], dv);
      const float xh = XH[b * F + c];
      const float yv = fmaf(xh, gam, bet);
      const float sg = 1.f / (1.f + expf(-yv));
      const float dy = dv * (sg + yv * sg * (1.f - sg));
      dgacc = fmaf(dy, xh, dgacc);
      dbacc += dy;
      DU[b * F + c] = dy * gam;
    }
    dgam_s[bq * F + c] = dgacc;
    dbet_s[bq * F + c] = dbacc;
  }
  __syncthreads();
  if (tid < F) {
    dprm[obnw + tid] = dgam_s[tid] + dgam_s[F + tid] + dgam_s[2 * F + tid] + dgam_s[3 * F + tid];
    dprm[obnb + tid] = dbet_s[tid] + dbet_s[F + tid] + dbet_s[2 * F + tid] + dbet_s[3 * F + tid];
    float m1 = 0.f, m2 = 0.f;
    if (training) {
      for (int b = 0; b < B; ++b) {
        m1 += DU[b * F + tid];
        m2 = fmaf(DU[b * F + tid], XH[b * F + tid], m2);
      }
      m1 /= (float)B;
      m2 /= (float)B;
    }
    m1_s[tid] = m1;
    m2_s[tid] = m2;
  }
  __syncthreads();
  for (int b = bq; b < B; b += 4)
    DU[b * F + c] = istd_s[c] * (DU[b * F + c] - m1_s[c] - XH[b * F + c] * m2_s[c]);
  // ---- dW0[c][i] = sum_b du[b][c] g[b][i] ; dg[b][i] += sum_c du[b][c] W0[c][i] ----
  for (int i0 = 0; i0 < G; i0 += 32) {
    const int ni = min(32, G - i0);
    __syncthreads();
    for (int e = tid; e < F * 32; e += FB_THREADS) {
      const int cc = e >> 5, i = e & 31;
      Ws[i * (F + 1) + cc] = (i < ni) ? prm[oW0 + (size_t)cc * G + i0 + i] : 0.f;
    }
    for (int e = tid; e < B * 32; e += FB_THREADS) {
      const int b = e >> 5, i = e & 31;
      gs[b * GS_LD + i] = (i < ni) ? g[(size_t)b * G + i0 + i] : 0.f;
    }
    __syncthreads();
    {
      const int iq = tid >> 6;   // 8 consecutive i per thread: two 16-byte broadcast loads of g feed 8 FMAs
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int b = 0; b < B; ++b) {
        const float d = DU[b * F + c];
        const float4 ga = *reinterpret_cast<const float4*>(gs + b * GS_LD + iq * 8);
        const float4 gb = *reinterpret_cast<const float4*>(gs + b * GS_LD + iq * 8 + 4);
        acc[0] = fmaf(d, ga.x, acc[0]); acc[1] = fmaf(d, ga.y, acc[1]); acc[2] = fmaf(d, ga.z, acc[2]); acc[3] = fmaf(d, ga.w, acc[3]);
        acc[4] = fmaf(d, gb.x, acc[4]); acc[5] = fmaf(d, gb.y, acc[5]); acc[6] = fmaf(d, gb.z, acc[6]); acc[7] = fmaf(d, gb.w, acc[7]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (iq * 8 + i < ni) dprm[oW0 + (size_t)c * G + i0 + iq * 8 + i] = acc[i];
    }
    {
      // dg[b][i] += sum_c du[b][c] W0[c][i]: 4 shapes (b = bb + 8 j) per thread, du rows read as 16-byte broadcasts
      const int i = tid & 31, bb = tid >> 5;
      for (int b0 = bb; b0 < B; b0 += 32) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int cc = 0; cc < F; cc += 4) {
          const float w0 = Ws[i * (F + 1) + cc], w1 = Ws[i * (F + 1) + cc + 1], w2 = Ws[i * (F + 1) + cc + 2], w3 = Ws[i * (F + 1) + cc + 3];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int b = min(b0 + 8 * j, B - 1);
            const float4 d = *reinterpret_cast<const float4*>(DU + b * F + cc);
            acc[j] = fmaf(d.x, w0, fmaf(d.y, w1, fmaf(d.z, w2, fmaf(d.w, w3, acc[j]))));
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int b = b0 + 8 * j;
          if (b < B && i < ni) atomicAdd(&dg[(size_t)b * G + i0 + i], acc[j]);
        }
      }
    }
  }
}

template <int K, int MODE>
int launch_bwd_t(const BwdArgs& a, int grid, int pass, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(coupling_bwd_p1_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P1Smem));
    cudaFuncSetAttribute(coupling_bwd_p2_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2Smem));
    attr = true;
  }
  if (pass == 1) {
    coupling_bwd_p1_kernel<K, MODE><<<grid, DPF_TILE, sizeof(P1Smem), st>>>(a);
    return dpf_check_launch("coupling_bwd_p1_kernel");
  }
  coupling_bwd_p2_kernel<K, MODE><<<grid, DPF_TILE, sizeof(P2Smem), st>>>(a);
  return dpf_check_launch("coupling_bwd_p2_kernel");
}

}  // namespace

int launch_coupling_bwd_fp32(const BwdArgs& a, int mode, int pass, cudaStream_t s) {
  const int grid = min(a.f.n_tiles, dpf_num_sms());
  if (a.f.k == 2) return mode == 0 ? launch_bwd_t<2, 0>(a, grid, pass, s) : launch_bwd_t<2, 1>(a, grid, pass, s);
  return mode == 0 ? launch_bwd_t<1, 0>(a, grid, pass, s) : launch_bwd_t<1, 1>(a, grid, pass, s);
}

int launch_coupling_bwd_final(const BwdArgs& a, const float* p_in, const float* dx_stored, float* dp, cudaStream_t s) {
  const long long total = (long long)a.f.B * a.f.N;
  const int grid = dp ? (int)min((long long)dpf_num_sms() * 4, (total + DPF_TILE - 1) / DPF_TILE) : 1;
  coupling_bwd_final_kernel<<<grid, DPF_TILE, 0, s>>>(a, p_in, dx_stored, dp);
  return dpf_check_launch("coupling_bwd_final_kernel");
}

int launch_film_backward(const float* arena, const float* stats, float* darena, const LayerMeta* meta_dev,
                         const float* g, const float* film, const float* dfilm, float* dg, int L, int B, int G,
                         int training, float eps, cudaStream_t s) {
  DPF_REQUIRE(B <= FB_MAXB, DPF_ERR_UNSUPPORTED, "film backward supports B <= %d (got %d)", FB_MAXB, B);
  // shared memory sized for the actual batch: 48 KB at B = 32 -> 4 CTAs per SM, all L*4 CTAs in one wave
  // (at the FB_MAXB cap, 137 KB, it is one CTA per SM and 252 CTAs take two waves)
  const size_t smem = sizeof(float) * ((size_t)3 * B * F + F * (F + 1) + (size_t)B * GS_LD + 1024 + 12 * F);
  static bool attr = false;
  if (!attr) {
    const size_t smem_max = sizeof(float) * ((size_t)3 * FB_MAXB * F + F * (F + 1) + (size_t)FB_MAXB * GS_LD + 1024 + 12 * F);
    cudaFuncSetAttribute(film_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    attr = true;
  }
  film_backward_kernel<<<L * 4, FB_THREADS, smem, s>>>(arena, stats, darena, meta_dev, g, film, dfilm, dg, B, G, training, eps);
  return dpf_check_launch("film_backward_kernel");
}
