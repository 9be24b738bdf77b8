// Shared definitions of the point-flow decoder kernels (conditional affine coupling stack).
//
// Reference semantics: CondRealNVPFlow3D / ...Triple / LocalCondRNVPDecoder
// (lib/networks/flows.py:10-160, lib/networks/decoders.py:41-72); arithmetic in SURVEY.md App. A.
//
// Parameter arena (fp32): per coupling layer two branches [mu, logvar], each laid out as
//   W0[F,k] bnA_w[F] bnA_b[F] W1[F,F]
//   fw0_W[F,G] fw0_bnw[F] fw0_bnb[F] fw1_W[F,F] fw1_b[F]       (FiLM scale net  T_*_0_cond_w)
//   fb0_W[F,G] fb0_bnw[F] fb0_bnb[F] fb1_W[F,F] fb1_b[F]       (FiLM shift net  T_*_0_cond_b)
//   W2[w,F] b2[w]
// The gradient arena has the same layout.  Stats arena per branch: 8 vectors of F:
//   bnA_rm bnA_rv bnB_rm bnB_rv fw_rm fw_rv fb_rm fb_rv.
// dpf_nets_b200/lib/networks/_arena.py mirrors this layout on the host.
#pragma once
#include "common.cuh"

#define DPF_F 64          // conditioner width (p_decoder_n_features); the kernels are specialised on it
#define DPF_BN_EPS 1e-5f
#define DPF_BN_MOM 0.1f
#define DPF_TILE 128      // points per tile = threads per CTA (one TMEM lane / one thread per point)
#define DPF_BNB_REP 16     // replicas of the BN_b sum accumulators in the merged train-mode forward
#define DPF_LTAB_ROW 16     // floats per (branch, channel) row of the backward tables
#define DPF_M12_REP 8      // replicas of the BN_b backward mean accumulators (pass 1 -> pass 2)
#define DPF_EVAL_LTAB_BYTES 2112   // sizeof(EvalLayerTab), coupling_tc.cu

struct LayerMeta {        // 8 x int64 per layer, identical on host (numpy int64) and device
  long long param_off, stat_off, k, w, keep0, keep1, warp0, warp1;
};

struct BranchLayout {
  int W0, bnA_w, bnA_b, W1, fw0_W, fw0_bnw, fw0_bnb, fw1_W, fw1_b, fb0_W, fb0_bnw, fb0_bnb, fb1_W, fb1_b, W2, b2, size;
};

__host__ __device__ inline BranchLayout branch_layout(int k, int w, int G) {
  const int F = DPF_F;
  BranchLayout o;
  int x = 0;
  o.W0 = x;      x += F * k;
  o.bnA_w = x;   x += F;
  o.bnA_b = x;   x += F;
  o.W1 = x;      x += F * F;
  o.fw0_W = x;   x += F * G;
  o.fw0_bnw = x; x += F;
  o.fw0_bnb = x; x += F;
  o.fw1_W = x;   x += F * F;
  o.fw1_b = x;   x += F;
  o.fb0_W = x;   x += F * G;
  o.fb0_bnw = x; x += F;
  o.fb0_bnb = x; x += F;
  o.fb1_W = x;   x += F * F;
  o.fb1_b = x;   x += F;
  o.W2 = x;      x += w * F;
  o.b2 = x;      x += w;
  o.size = x;
  return o;
}

// stats vectors (index * F) inside a branch's stats block of 8*F floats
enum { ST_BNA_RM = 0, ST_BNA_RV, ST_BNB_RM, ST_BNB_RV, ST_FW_RM, ST_FW_RV, ST_FB_RM, ST_FB_RV, ST_COUNT };

// Workspace carved by decoder.cu; all sub-buffers 256-byte aligned.
struct DecoderWorkspace {
  float* film;        // [L][4][B][F]   s_mu, t_mu, s_lv, t_lv  (s already = eps + exp(.))
  double* moments;    // [L+1][16]      sum x_c (3), sum x_c x_c' (xx,xy,xz,yy,yz,zz) of each step's input
  double* bnb_sums;   // [L][2][F][2]   sum / sum of squares of h2pre (train)
  double* bnb_rep;    // [L][DPF_BNB_REP][2][F][2] replicated accumulators of the merged forward (cuts same-address atomics)
  float* dfilm;       // [L][4][B][F]   backward: ds_raw_mu, dt_mu, ds_raw_lv, dt_lv
  double* bna_sums;   // [L][2][F][4]   backward: dbeta, E0, E1, pad
  float* dx[2];       // [B][3][N]      ping-pong stored input gradients
  unsigned short* w1_bf16;  // [L][2][3][F*F] bf16 images {W1 hi, W1^T hi, W1 lo} in UMMA smem layout (tensor path)
  unsigned char* eval_ltab; // [L] EvalLayerTab (fused eval decoder)
  float* eval_epi;          // [L][B][2][F] float4 {S, T, W2_0, W2_1} (fused eval decoder)
  float* ltab;              // [L][2*F][DPF_LTAB_ROW] backward: per-(branch, channel) {A00, A01, c0, bnB mean | bnB istd, W2_0, W2_1, b2 |
                            //   W0_0, W0_1, bnA gamma, bnA mean | bnA istd, -, -, -} (bwd_tables_kernel)
  float* pend;              // [L][8]  backward: deferred BN_a correction {c0, c1, q00, q01, q11} pass 1 hands to pass 2
  double* m12_rep;          // [L][DPF_M12_REP][2][F][2] backward: replicated sum_b s*dt, sum_b s*ds (BN_b batch terms m1, m2)
  unsigned int* barriers;   // [L][32] grid-barrier counters of the merged train-mode forward (one 128-B line per layer)
  unsigned int* barriers_bwd;   // [L][32] the same for the merged backward launch
  size_t bytes;
};

__host__ inline size_t dpf_align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ inline DecoderWorkspace carve_workspace(void* base, int L, int G, int B, int N) {
  DecoderWorkspace w;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += dpf_align256(bytes); return r; };
  w.film = (float*)take(sizeof(float) * (size_t)L * 4 * B * DPF_F);
  w.moments = (double*)take(sizeof(double) * (size_t)(L + 1) * 16);
  w.bnb_sums = (double*)take(sizeof(double) * (size_t)L * 2 * DPF_F * 2);
  w.bnb_rep = (double*)take(sizeof(double) * (size_t)L * DPF_BNB_REP * 2 * DPF_F * 2);
  w.dfilm = (float*)take(sizeof(float) * (size_t)L * 4 * B * DPF_F);
  w.bna_sums = (double*)take(sizeof(double) * (size_t)L * 2 * DPF_F * 4);
  w.dx[0] = (float*)take(sizeof(float) * (size_t)B * 3 * N);
  w.dx[1] = (float*)take(sizeof(float) * (size_t)B * 3 * N);
  w.w1_bf16 = (unsigned short*)take(sizeof(unsigned short) * (size_t)L * 6 * DPF_F * DPF_F);
  w.barriers = (unsigned int*)take(sizeof(unsigned int) * (size_t)L * 32);
  w.barriers_bwd = (unsigned int*)take(sizeof(unsigned int) * (size_t)L * 32);
  w.ltab = (float*)take(sizeof(float) * (size_t)L * 2 * DPF_F * DPF_LTAB_ROW);
  w.pend = (float*)take(sizeof(float) * (size_t)L * 8);
  w.m12_rep = (double*)take(sizeof(double) * (size_t)L * DPF_M12_REP * 2 * DPF_F * 2);
  w.eval_ltab = (unsigned char*)take((size_t)L * DPF_EVAL_LTAB_BYTES);
  w.eval_epi = (float*)take(sizeof(float) * (size_t)L * B * 2 * DPF_F * 4);
  w.bytes = off;
  (void)G;
  return w;
}

// Per-launch arguments of the per-layer kernels (passed by value).
struct CouplingArgs {
  const float* x;        // layer input (B,3,N)
  float* y;              // outputs (B,3,N): transformed points, mu (nullable: not written), logvar
  float* mu;
  float* lv;
  float* slv;            // nullable: running sum over the processed layers of logvar (B,3,N) - the per-point log-det the
  int slv_init;          //   flow NLL consumes (losses.py:12-13); slv_init != 0 = first processed layer (store, not add)
  const float* prm;      // this layer's parameters (arena + param_off)
  float* stat;           // this layer's stats block (2 branches x 8 x F)
  const float* film;     // [4][B][F] of this layer
  const double* mom_in;  // [16] moments of x (train)
  double* mom_out;       // [16] accumulators for moments of y (train, next step) or null
  double* bnb_sums;      // [2][F][2]
  double* bnb_rep;       // [DPF_BNB_REP][2][F][2] or null (merged forward only)
  int B, N, G;
  int k, w, keep0, keep1, warp0, warp1;
  int training, update_stats;
  float eps;
  int tiles_per_b, n_tiles;
};

__device__ __forceinline__ float softsign(float o) { return o / (1.f + fabsf(o)); }

// outputs of one point of one layer: y, logvar always; mu and the running log-det sum when requested
__device__ __forceinline__ void store_point_outputs(const CouplingArgs& a, size_t base, const float yv[3], const float muv[3],
                                                    const float lvv[3]) {
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const size_t o = base + (size_t)ch * a.N;
    a.y[o] = yv[ch];
    if (a.mu) a.mu[o] = muv[ch];
    a.lv[o] = lvv[ch];
    if (a.slv) {
      if (a.slv_init) a.slv[o] = lvv[ch];
      else if (ch == a.warp0 || ch == a.warp1) a.slv[o] += lvv[ch];     // kept channels contribute exactly 0
    }
  }
}

// second-moment index of (i,j), i<=j, in the 9-vector layout [s0 s1 s2 xx xy xz yy yz zz]
__device__ __forceinline__ int mom2_index(int i, int j) {
  if (i > j) { const int t = i; i = j; j = t; }
  return 3 + (i == 0 ? j : (i == 1 ? 2 + j : 5));
}

// Folded BN_a of one (branch, channel): h1 = relu(A0.x_keep + c0).  Shared by all coupling kernels.
__device__ inline void fold_bn_a(const CouplingArgs& a, const BranchLayout& lay, int br, int c, bool write_stats,
                          float& A00, float& A01, float& c0, float* mean_out, float* istd_out) {
  const float* prm = a.prm + (size_t)br * lay.size;
  float* st = a.stat + (size_t)br * ST_COUNT * DPF_F;
  const float w0 = prm[lay.W0 + c * a.k + 0];
  const float w1 = (a.k == 2) ? prm[lay.W0 + c * a.k + 1] : 0.f;
  float mean, var;
  if (a.training) {
    const double M = (double)a.B * (double)a.N;
    const double rM = 1.0 / M;        // one double-precision division instead of five (this runs in every CTA's prologue)
    const double m0 = a.mom_in[a.keep0] * rM;
    const double v00 = a.mom_in[mom2_index(a.keep0, a.keep0)] * rM - m0 * m0;
    double dm = (double)w0 * m0, dv = (double)w0 * w0 * v00;
    if (a.k == 2) {
      const double m1 = a.mom_in[a.keep1] * rM;
      const double v11 = a.mom_in[mom2_index(a.keep1, a.keep1)] * rM - m1 * m1;
      const double v01 = a.mom_in[mom2_index(a.keep0, a.keep1)] * rM - m0 * m1;
      dm += (double)w1 * m1;
      dv += (double)w1 * w1 * v11 + 2.0 * (double)w0 * w1 * v01;
    }
    mean = (float)dm;
    var = (float)fmax(dv, 0.0);
    if (write_stats) {
      st[ST_BNA_RM * DPF_F + c] = (1.f - DPF_BN_MOM) * st[ST_BNA_RM * DPF_F + c] + DPF_BN_MOM * mean;
      st[ST_BNA_RV * DPF_F + c] = (1.f - DPF_BN_MOM) * st[ST_BNA_RV * DPF_F + c] + DPF_BN_MOM * (float)(fmax(dv, 0.0) * (M / fmax(M - 1.0, 1.0)));
    }
  } else {
    mean = st[ST_BNA_RM * DPF_F + c];
    var = st[ST_BNA_RV * DPF_F + c];
  }
  const float istd = 1.f / sqrtf(var + DPF_BN_EPS);
  const float gi = prm[lay.bnA_w + c] * istd;
  A00 = gi * w0;
  A01 = gi * w1;
  c0 = prm[lay.bnA_b + c] - gi * mean;
  if (mean_out) *mean_out = mean;
  if (istd_out) *istd_out = istd;
}

// BN_b statistics of one (branch, channel) from the stats pass (train) or the running buffers (eval).
__device__ inline void bn_b_stats(const CouplingArgs& a, int br, int c, bool write_stats, float& mean, float& istd) {
  float* st = a.stat + (size_t)br * ST_COUNT * DPF_F;
  float var;
  if (a.training) {
    const double M = (double)a.B * (double)a.N;
    const double s = a.bnb_sums[(br * DPF_F + c) * 2 + 0], q = a.bnb_sums[(br * DPF_F + c) * 2 + 1];
    const double dm = s / M;
    const double dv = fmax(q / M - dm * dm, 0.0);
    mean = (float)dm;
    var = (float)dv;
    if (write_stats) {
      st[ST_BNB_RM * DPF_F + c] = (1.f - DPF_BN_MOM) * st[ST_BNB_RM * DPF_F + c] + DPF_BN_MOM * mean;
      st[ST_BNB_RV * DPF_F + c] = (1.f - DPF_BN_MOM) * st[ST_BNB_RV * DPF_F + c] + DPF_BN_MOM * (float)(dv * (M / fmax(M - 1.0, 1.0)));
    }
  } else {
    mean = st[ST_BNB_RM * DPF_F + c];
    var = st[ST_BNB_RV * DPF_F + c];
  }
  istd = 1.f / sqrtf(var + DPF_BN_EPS);
}


// ---- backward -------------------------------------------------------------------------------
struct BnA { float w0, w1, mean, istd, gamma; };

// BN_a quantities of (branch, channel) of an arbitrary layer (used for the deferred correction).
__device__ inline BnA bn_a_of(const float* prm_layer, const float* stat_layer, const BranchLayout& lay,
                              const double* mom, double M, int k, int keep0, int keep1, int training, int br, int c) {
  const float* prm = prm_layer + (size_t)br * lay.size;
  const float* st = stat_layer + (size_t)br * ST_COUNT * DPF_F;
  BnA o;
  o.w0 = prm[lay.W0 + c * k + 0];
  o.w1 = (k == 2) ? prm[lay.W0 + c * k + 1] : 0.f;
  o.gamma = prm[lay.bnA_w + c];
  float var;
  if (training) {
    const double rM = 1.0 / M;        // same arithmetic as fold_bn_a (identical mean / istd)
    const double m0 = mom[keep0] * rM;
    const double v00 = mom[mom2_index(keep0, keep0)] * rM - m0 * m0;
    double dm = (double)o.w0 * m0, dv = (double)o.w0 * o.w0 * v00;
    if (k == 2) {
      const double m1 = mom[keep1] * rM;
      const double v11 = mom[mom2_index(keep1, keep1)] * rM - m1 * m1;
      const double v01 = mom[mom2_index(keep0, keep1)] * rM - m0 * m1;
      dm += (double)o.w1 * m1;
      dv += (double)o.w1 * o.w1 * v11 + 2.0 * (double)o.w0 * o.w1 * v01;
    }
    o.mean = (float)dm;
    var = (float)fmax(dv, 0.0);
  } else {
    o.mean = st[ST_BNA_RM * DPF_F + c];
    var = st[ST_BNA_RV * DPF_F + c];
  }
  o.istd = 1.f / sqrtf(var + DPF_BN_EPS);
  return o;
}

// Arguments of the two backward passes of one layer (step q of the processing order).
struct BwdArgs {
  CouplingArgs f;          // forward view of this layer (x, prm, stat, film, mom_in, bnb_sums, sizes ...)
  const float* yv;         // this layer's outputs P_out[l], LV[l]
  const float* lvv;
  const float* dy_chain;   // stored gradient w.r.t. y from the previous backward step (null at the first)
  const float* dP;         // external cotangents of this layer's list entries (nullable)
  const float* dMU;
  const float* dLV;
  float* dx_out;           // stored gradient w.r.t. x
  float* dfilm;            // [4][B][F]   ds_raw_mu, dt_mu, ds_raw_lv, dt_lv (accumulated)
  float* dprm;             // this layer's slice of the gradient arena
  double* bna_sums;        // [2][F][4]   dbeta, E0, E1 (accumulated in pass 2)
  const float* ltab;       // tensor path: [2*F][DPF_LTAB_ROW] this layer's folded tables (bwd_tables_kernel), nullable
  const float* n_ltab;     // the same table of the layer that owes the pending correction (nullable)
  float* pend_store;       // tensor path: [8] pass 1 (CTA 0) stores this step's Pending, pass 2 loads it (nullable)
  double* m12_rep;         // tensor path: [DPF_M12_REP][2][F][2] pass 1 accumulates, pass 2 sums (nullable)
  float* dw1_partial;      // tensor path: per-CTA wgrad partials [grid][2][F*F] of this layer (reduced once per pass)
  // deferred BN_a correction owed by the layer processed just before in backward (its input == our y)
  int has_pending;
  const float* nprm;
  const float* nstat;
  float* ndprm;
  const double* n_bna_sums;
  const double* n_mom;
  int nk, nw, nkeep0, nkeep1;
};

struct Pending { float c0, c1, q00, q01, q11; };
