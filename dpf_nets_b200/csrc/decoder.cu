// Host-side orchestration of the coupling stack behind the C ABI: one call runs all L layers
// (FiLM pre-compute, per-layer statistics + apply launches) on the caller's stream, so the Python
// side pays one FFI call per decoder pass and the whole pass can be captured in a CUDA graph.
#include "coupling.cuh"
#include <vector>

// ---- optional per-kernel-class timing (CUDA events on the launching stream; bench.py roofline) ----
namespace {
enum { CAT_FILM_FWD = 0, CAT_MOMENTS, CAT_FWD_STATS, CAT_FWD_APPLY, CAT_BWD_P1, CAT_BWD_P2, CAT_BWD_FINAL, CAT_FILM_BWD, CAT_COUNT };
struct ProfRec { int cat; cudaEvent_t a, b; };
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  std::vector<ProfRec> recs;
  cudaEvent_t get() {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
  }
} g_prof;
struct ProfScope {
  int cat; cudaStream_t s; cudaEvent_t a;
  ProfScope(int c, cudaStream_t st) : cat(c), s(st), a(nullptr) { if (g_prof.on) { a = g_prof.get(); cudaEventRecord(a, s); } }
  ~ProfScope() { if (g_prof.on) { cudaEvent_t b = g_prof.get(); cudaEventRecord(b, s); g_prof.recs.push_back({cat, a, b}); } }
};
}  // namespace

// Enables (and clears) / disables per-kernel-class event timing of the decoder entry points.
DPF_API int dpf_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.used = 0;
  g_prof.recs.clear();
  return DPF_OK;
}

// Sums the recorded durations: ms[c], counts[c] for the n (<= 8) kernel classes
// {film_fwd, moments, fwd_stats, fwd_apply, bwd_p1, bwd_p2, bwd_final, film_bwd}.
DPF_API int dpf_profile_collect(double* ms, long long* counts, int n) {
  DPF_REQUIRE(ms && counts && n > 0, DPF_ERR_BAD_ARG, "dpf_profile_collect: bad arguments");
  for (int i = 0; i < n; ++i) { ms[i] = 0.0; counts[i] = 0; }
  for (const ProfRec& r : g_prof.recs) {
    cudaEventSynchronize(r.b);
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    if (r.cat < n) { ms[r.cat] += t; counts[r.cat] += 1; }
  }
  g_prof.used = 0;
  g_prof.recs.clear();
  return DPF_OK;
}

// launchers (coupling_fwd.cu / coupling_bwd.cu / coupling_tc.cu)
int launch_film_forward(const float* arena, float* stats, const LayerMeta* meta_dev, const float* g, float* film,
                        int L, int B, int G, int training, int update_stats, float eps, cudaStream_t s);
int launch_moments(const float* x, int B, int N, double* mom, cudaStream_t s);
int launch_coupling_fwd_fp32(const CouplingArgs& a, int mode, bool stats_pass, cudaStream_t s);
int launch_pack_w1(const float* arena, const LayerMeta* meta_dev, int L, int G, unsigned short* out, cudaStream_t s);
int launch_coupling_fwd_tc(const CouplingArgs& a, const unsigned short* wimg, int mode, bool stats_pass, int split, cudaStream_t s);
size_t tc_weight_image_elems_per_layer();
int launch_coupling_fwd_train_tc(const CouplingArgs& a, const unsigned short* wimg, int mode, int split, unsigned int* counter,
                                 cudaStream_t s);

// The merged (cooperative, TMEM-resident) train-mode forward is on by default; dpf_set_option(0, 0)
// forces the two-launch form (tests compare both).
static bool g_merged_forward = true;
// Option 1: the fused all-layer eval decoder (coupling_eval.cu) is on by default for the tensor
// precisions; dpf_set_option(1, 0) forces one launch per layer (tests compare both).
static bool g_fused_eval = true;
int launch_decoder_eval_tc(const float* arena, const float* stats, const LayerMeta* meta_dev, const float* film,
                           const unsigned short* wimg, unsigned char* ltab, float* epi, const float* p, float* P, float* MU,
                           float* LV, float* SLV, int L, int G, int B, int N, int mode, int split, float eps, cudaStream_t s);
// Option 2: programmatic dependent launch of the per-layer tensor-path kernels (common.cuh).
int g_dpf_pdl = 1;
// Option 3: backward pass 2 with two tiles in flight per SM (544-thread CTAs, MMA issuer warp); 0 = the
// one-tile-per-SM form (tests compare both).
int g_dpf_p2_two_tiles = 1;
extern int g_fused_pairwise;   // chamfer.cu
// Option 5: backward of a layer as ONE launch (pass 1 -> grid barrier -> pass 2).  Measured r02 (32 x 2048, bf16x3, 20 timed
// steps, A/B/A/B on one B200): 4.813 ms per step merged vs 4.765 ms with two PDL-chained launches - programmatic dependent
// launch already hides the launch boundary, and the in-kernel grid barrier costs what the boundary did.  Default = two
// launches; the merged form stays available (dpf_set_option(5, 1)) and is compared against it in the tests.
static bool g_merged_backward = false;
int launch_coupling_bwd_merged_tc(const BwdArgs& a, const unsigned short* wimg, int mode, int split, unsigned int* counter, cudaStream_t s);
DPF_API int dpf_set_option(int option, int value) {
  DPF_REQUIRE(option >= 0 && option <= 5, DPF_ERR_BAD_ARG, "dpf_set_option: unknown option %d", option);
  if (option == 5) { g_merged_backward = value != 0; return DPF_OK; }
  if (option == 0) g_merged_forward = value != 0;
  else if (option == 1) g_fused_eval = value != 0;
  else if (option == 2) g_dpf_pdl = value != 0;
  else if (option == 3) g_dpf_p2_two_tiles = value != 0;
  else g_fused_pairwise = value != 0;
  return DPF_OK;
}

int tc_barrier_failed();
int tc_coresidency_state();

static int validate_common(const long long* meta_host, int L, int G, int B, int N, int mode, int precision) {
  DPF_REQUIRE(!tc_barrier_failed(), DPF_ERR_BARRIER,
              "decoder: a grid barrier of an earlier pass timed out (its CTAs were not co-resident: another tenant on the GPU?); "
              "that pass was aborted and its outputs are invalid");
  DPF_REQUIRE(meta_host, DPF_ERR_NULL_PTR, "decoder: layer table is null");
  DPF_REQUIRE(L > 0 && G > 0 && B > 0 && N > 0, DPF_ERR_BAD_ARG, "decoder: L, G, B, N must be positive");
  DPF_REQUIRE(mode == 0 || mode == 1, DPF_ERR_BAD_ARG, "decoder: mode must be 0 (direct) or 1 (inverse)");
  DPF_REQUIRE(precision >= 0 && precision <= 2, DPF_ERR_BAD_ARG, "decoder: precision must be 0 (fp32), 1 (bf16) or 2 (bf16x3)");
  for (int l = 0; l < L; ++l) {
    const LayerMeta& m = reinterpret_cast<const LayerMeta*>(meta_host)[l];
    DPF_REQUIRE((m.k == 1 && m.w == 2) || (m.k == 2 && m.w == 1), DPF_ERR_BAD_ARG, "decoder: layer %d has k=%lld w=%lld", l, m.k, m.w);
  }
  return DPF_OK;
}

DPF_API int dpf_decoder_workspace_bytes(int L, int G, int B, int N, long long* bytes) {
  DPF_REQUIRE(bytes, DPF_ERR_NULL_PTR, "dpf_decoder_workspace_bytes: null out pointer");
  DPF_REQUIRE(L > 0 && G > 0 && B > 0 && N > 0, DPF_ERR_BAD_ARG, "dpf_decoder_workspace_bytes: bad sizes");
  *bytes = (long long)carve_workspace(nullptr, L, G, B, N).bytes;
  return DPF_OK;
}

static CouplingArgs make_args(const LayerMeta& m, const float* arena, float* stats, const DecoderWorkspace& ws,
                              int l, int q, int G, int B, int N, int training, int update_stats, float eps) {
  CouplingArgs a{};
  a.prm = arena + m.param_off;
  a.stat = stats + m.stat_off;
  a.film = ws.film + (size_t)l * 4 * B * DPF_F;
  a.mom_in = ws.moments + (size_t)q * 16;
  a.mom_out = training ? ws.moments + (size_t)(q + 1) * 16 : nullptr;
  a.bnb_sums = ws.bnb_sums + (size_t)l * 2 * DPF_F * 2;
  a.bnb_rep = nullptr;
  a.B = B; a.N = N; a.G = G;
  a.k = (int)m.k; a.w = (int)m.w;
  a.keep0 = (int)m.keep0; a.keep1 = (int)m.keep1; a.warp0 = (int)m.warp0; a.warp1 = (int)m.warp1;
  a.training = training; a.update_stats = update_stats; a.eps = eps;
  a.tiles_per_b = (N + DPF_TILE - 1) / DPF_TILE;
  a.n_tiles = a.tiles_per_b * B;
  return a;
}

// Forward of the whole stack: LocalCondRNVPDecoder.forward (decoders.py:54-72) over
// CondRealNVPFlow3D.forward (flows.py:95-117).  mode 0 = 'direct' (layer 0 first), 1 = 'inverse'
// (layer L-1 first); outputs are indexed by LAYER, like the reference's lists.
// MU may be NULL (the per-layer mu outputs are not written: nothing on the training path reads them, losses.py:7-15,
// and the backward does not need them); SLV (B,3,N), nullable, receives sum_l LV[l] - the per-point log-det sum
// of PointFlowNLL, accumulated in the kernels' epilogues instead of a separate reduction over the 63 planes.
DPF_API int dpf_decoder_forward_ex(const long long* meta_host, const long long* meta_dev, const float* arena,
                                   float* stats, const float* p, const float* g, float* P_out, float* MU, float* LV,
                                   float* SLV, void* workspace, int L, int G, int B, int N, int mode, int training,
                                   int update_stats, int precision, float eps, void* stream) {
  int rc = validate_common(meta_host, L, G, B, N, mode, precision);
  if (rc) return rc;
  DPF_REQUIRE(meta_dev && arena && stats && p && g && P_out && LV && workspace, DPF_ERR_NULL_PTR,
              "dpf_decoder_forward: null pointer");
  DPF_REQUIRE(!training || B <= 256, DPF_ERR_UNSUPPORTED, "dpf_decoder_forward: train-mode FiLM BatchNorm supports B <= 256 (got %d)", B);
  DPF_REQUIRE(!training || ((long long)B * N > 1 && B > 1), DPF_ERR_BAD_ARG, "dpf_decoder_forward: BatchNorm in training mode needs more than 1 value per channel");
  cudaStream_t s = (cudaStream_t)stream;
  const LayerMeta* meta = reinterpret_cast<const LayerMeta*>(meta_host);
  DecoderWorkspace ws = carve_workspace(workspace, L, G, B, N);
  const size_t plane = (size_t)B * 3 * N;

  {
    ProfScope ps(CAT_FILM_FWD, s);
    rc = launch_film_forward(arena, stats, reinterpret_cast<const LayerMeta*>(meta_dev), g, ws.film, L, B, G, training,
                             update_stats, eps, s);
  }
  if (rc) return rc;
  if (precision >= 1) {
    rc = launch_pack_w1(arena, reinterpret_cast<const LayerMeta*>(meta_dev), L, G, ws.w1_bf16, s);
    if (rc) return rc;
  }
  if (training) {
    cudaMemsetAsync(ws.moments, 0, sizeof(double) * (size_t)(L + 1) * 16, s);
    cudaMemsetAsync(ws.bnb_sums, 0, sizeof(double) * (size_t)L * 2 * DPF_F * 2, s);
    if (precision >= 1 && g_merged_forward) cudaMemsetAsync(ws.bnb_rep, 0, sizeof(double) * (size_t)L * DPF_BNB_REP * 2 * DPF_F * 2, s);
    ProfScope ps(CAT_MOMENTS, s);
    rc = launch_moments(p, B, N, ws.moments, s);
    if (rc) return rc;
  }
  if (!training && precision >= 1 && g_fused_eval) {
    // eval mode has no cross-point reduction: the whole stack is one launch (+ one for its tables)
    ProfScope ps(CAT_FWD_APPLY, s);
    return launch_decoder_eval_tc(arena, stats, reinterpret_cast<const LayerMeta*>(meta_dev), ws.film, ws.w1_bf16, ws.eval_ltab,
                                  ws.eval_epi, p, P_out, MU, LV, SLV, L, G, B, N, mode, precision == 2, eps, s);
  }
  const float* x = p;
  bool merged_ok = g_merged_forward;
  if (training && precision >= 1) cudaMemsetAsync(ws.barriers, 0, sizeof(unsigned int) * (size_t)L * 32, s);
  for (int q = 0; q < L; ++q) {
    const int l = mode == 0 ? q : L - 1 - q;
    CouplingArgs a = make_args(meta[l], arena, stats, ws, l, q, G, B, N, training, update_stats, eps);
    a.x = x;
    a.y = P_out + (size_t)l * plane;
    a.mu = MU ? MU + (size_t)l * plane : nullptr;
    a.lv = LV + (size_t)l * plane;
    a.slv = SLV;
    a.slv_init = q == 0;
    const unsigned short* wimg = ws.w1_bf16 + (size_t)l * tc_weight_image_elems_per_layer();
    if (training && precision >= 1 && merged_ok) {
      // one cooperative launch: statistics -> grid barrier -> apply from TMEM-resident accumulators
      ProfScope ps(CAT_FWD_APPLY, s);
      a.bnb_rep = ws.bnb_rep + (size_t)l * DPF_BNB_REP * 2 * DPF_F * 2;
      rc = launch_coupling_fwd_train_tc(a, wimg, mode, precision == 2, ws.barriers + (size_t)l * 32, s);
      a.bnb_rep = nullptr;
      if (rc == DPF_OK) { x = a.y; continue; }
      if (rc != DPF_ERR_UNSUPPORTED) return rc;
      merged_ok = false;                        // does not fit: two-launch form for this and later layers
    }
    if (training) {
      ProfScope ps(CAT_FWD_STATS, s);
      rc = precision >= 1 ? launch_coupling_fwd_tc(a, wimg, mode, true, precision == 2, s) : launch_coupling_fwd_fp32(a, mode, true, s);
      if (rc) return rc;
    }
    {
      ProfScope ps(CAT_FWD_APPLY, s);
      rc = precision >= 1 ? launch_coupling_fwd_tc(a, wimg, mode, false, precision == 2, s) : launch_coupling_fwd_fp32(a, mode, false, s);
    }
    if (rc) return rc;
    x = a.y;
  }
  return DPF_OK;
}

DPF_API int dpf_decoder_forward(const long long* meta_host, const long long* meta_dev, const float* arena,
                                float* stats, const float* p, const float* g, float* P_out, float* MU, float* LV,
                                void* workspace, int L, int G, int B, int N, int mode, int training,
                                int update_stats, int precision, float eps, void* stream) {
  DPF_REQUIRE(MU, DPF_ERR_NULL_PTR, "dpf_decoder_forward: null pointer (MU); dpf_decoder_forward_ex accepts MU == NULL");
  return dpf_decoder_forward_ex(meta_host, meta_dev, arena, stats, p, g, P_out, MU, LV, nullptr, workspace, L, G, B, N, mode,
                                training, update_stats, precision, eps, stream);
}

// Synchronous health check of a forward workspace: *flag = number of layers whose grid barrier gave up
// (must be 0; a non-zero value means the merged forward's CTAs were not co-resident and the outputs
// of that pass are invalid).
DPF_API int dpf_decoder_status(const void* workspace, int L, int G, int B, int N, int* flag) {
  DPF_REQUIRE(workspace && flag && L > 0, DPF_ERR_BAD_ARG, "dpf_decoder_status: bad arguments");
  DecoderWorkspace ws = carve_workspace(const_cast<void*>(workspace), L, G, B, N);
  std::vector<unsigned int> host((size_t)L * 32);
  cudaError_t e = cudaMemcpy(host.data(), ws.barriers, host.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    dpf_set_error("dpf_decoder_status: %s", cudaGetErrorString(e));
    return (int)e;
  }
  int bad = tc_barrier_failed() ? 1 : 0;
  for (int l = 0; l < L; ++l) bad += host[(size_t)l * 32 + 1] != 0;
  *flag = bad;
  return DPF_OK;
}

// Non-blocking: *failed = 1 when a grid barrier of this process has timed out (reads a pinned host flag, no
// device synchronisation); *coresident = 1 verified by the probe launch, 0 probe failed (the merged forward is
// then not used), -1 not probed yet.
DPF_API int dpf_decoder_barrier_state(int* failed, int* coresident) {
  DPF_REQUIRE(failed && coresident, DPF_ERR_NULL_PTR, "dpf_decoder_barrier_state: null pointer");
  *failed = tc_barrier_failed();
  *coresident = tc_coresidency_state();
  return DPF_OK;
}

int launch_coupling_bwd_fp32(const BwdArgs& a, int mode, int pass, cudaStream_t s);
int launch_coupling_bwd_tc(const BwdArgs& a, const unsigned short* wimg, int mode, int pass, int split, cudaStream_t s);
int launch_dw1_reduce(const float* partial, int n_cta, float* darena, const LayerMeta* meta_dev, int L, int G, cudaStream_t s);
int tc_bwd_p2_max_ctas();
int launch_bwd_tables(const float* arena, float* stats, const LayerMeta* meta_dev, const double* moments, const double* bnb_sums,
                      float* ltab, int L, int G, int B, int N, int mode, int training, cudaStream_t s);

// Scratch of the tensor-core backward: per-CTA wgrad partials [L][ctas][2][64*64] fp32.
DPF_API int dpf_decoder_backward_scratch_bytes(int L, int B, int N, long long* bytes) {
  DPF_REQUIRE(bytes && L > 0 && B > 0 && N > 0, DPF_ERR_BAD_ARG, "dpf_decoder_backward_scratch_bytes: bad arguments");
  const long long tiles = (long long)B * ((N + DPF_TILE - 1) / DPF_TILE);
  const long long ctas = tiles < tc_bwd_p2_max_ctas() ? tiles : tc_bwd_p2_max_ctas();
  *bytes = (long long)sizeof(float) * L * ctas * 2 * DPF_F * DPF_F;
  return DPF_OK;
}
int launch_coupling_bwd_final(const BwdArgs& a, const float* p_in, const float* dx_stored, float* dp, cudaStream_t s);
int launch_film_backward(const float* arena, const float* stats, float* darena, const LayerMeta* meta_dev,
                         const float* g, const float* film, const float* dfilm, float* dg, int L, int B, int G,
                         int training, float eps, cudaStream_t s);

// side stream + fork / join events for the independent kernels at the tail of the backward (one set per device)
namespace {
struct TailStreams { cudaStream_t side; cudaEvent_t fork, join; bool ok; };
TailStreams* tail_streams() {
  static TailStreams per_dev[16] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  TailStreams& t = per_dev[dev];
  if (!t.ok) {
    if (cudaStreamCreateWithFlags(&t.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&t.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&t.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    t.ok = true;
  }
  return &t;
}
}  // namespace

// Backward of dpf_decoder_forward (what torch.autograd does for the reference modules; formulas in
// SURVEY.md Appendix F).  `workspace` must be the one the forward call used (FiLM outputs, input
// moments and BN_b sums live there).  dP / dMU / dLV are the cotangents of the stacked outputs
// (nullable; *_stride = elements between consecutive layers, 0 = the same (B,3,N) block for every
// layer).  darena (n_params) and dg (B,G) are overwritten; dp (B,3,N) is optional.
DPF_API int dpf_decoder_backward(const long long* meta_host, const long long* meta_dev, const float* arena,
                                 float* stats, const float* p, const float* g, const float* P_out, const float* LV,
                                 const float* dP, long long dP_stride, const float* dMU, long long dMU_stride,
                                 const float* dLV, long long dLV_stride, float* darena, long long n_params,
                                 float* dg, float* dp, void* workspace, void* bwd_scratch, int L, int G, int B, int N,
                                 int mode, int training, int precision, float eps, void* stream) {
  int rc = validate_common(meta_host, L, G, B, N, mode, precision);
  if (rc) return rc;
  DPF_REQUIRE(meta_dev && arena && stats && p && g && P_out && LV && darena && dg && workspace, DPF_ERR_NULL_PTR,
              "dpf_decoder_backward: null pointer");
  DPF_REQUIRE(precision == 0 || bwd_scratch, DPF_ERR_NULL_PTR, "dpf_decoder_backward: the tensor path needs bwd_scratch (dpf_decoder_backward_scratch_bytes)");
  cudaStream_t s = (cudaStream_t)stream;
  const LayerMeta* meta = reinterpret_cast<const LayerMeta*>(meta_host);
  DecoderWorkspace ws = carve_workspace(workspace, L, G, B, N);
  const size_t plane = (size_t)B * 3 * N;
  const long long n_tiles_all = (long long)B * ((N + DPF_TILE - 1) / DPF_TILE);
  const int p2_ctas = (int)(n_tiles_all < tc_bwd_p2_max_ctas() ? n_tiles_all : tc_bwd_p2_max_ctas());
  cudaMemsetAsync(darena, 0, sizeof(float) * (size_t)n_params, s);
  cudaMemsetAsync(dg, 0, sizeof(float) * (size_t)B * G, s);
  cudaMemsetAsync(ws.dfilm, 0, sizeof(float) * (size_t)L * 4 * B * DPF_F, s);
  cudaMemsetAsync(ws.bna_sums, 0, sizeof(double) * (size_t)L * 2 * DPF_F * 4, s);
  if (precision >= 1) cudaMemsetAsync(ws.m12_rep, 0, sizeof(double) * (size_t)L * DPF_M12_REP * 2 * DPF_F * 2, s);
  if (precision >= 1) cudaMemsetAsync(ws.barriers_bwd, 0, sizeof(unsigned int) * (size_t)L * 32, s);
  bool merged_bwd = precision >= 1 && g_merged_backward;

  if (precision >= 1) {   // folded BN_a / BN_b tables of every layer, once per pass (the per-layer kernels only load them)
    rc = launch_bwd_tables(arena, stats, reinterpret_cast<const LayerMeta*>(meta_dev), ws.moments, ws.bnb_sums, ws.ltab, L, G, B, N,
                           mode, training, s);
    if (rc) return rc;
  }
  auto layer_of = [&](int q) { return mode == 0 ? q : L - 1 - q; };
  auto set_pending = [&](BwdArgs& a, int qn) {   // correction owed by the layer of step qn
    const int ln = layer_of(qn);
    const LayerMeta& mn = meta[ln];
    a.has_pending = 1;
    a.nprm = arena + mn.param_off;
    a.nstat = stats + mn.stat_off;
    a.ndprm = darena + mn.param_off;
    a.n_bna_sums = ws.bna_sums + (size_t)ln * 2 * DPF_F * 4;
    a.n_mom = ws.moments + (size_t)qn * 16;
    a.n_ltab = precision >= 1 ? ws.ltab + (size_t)ln * 2 * DPF_F * DPF_LTAB_ROW : nullptr;
    a.nk = (int)mn.k; a.nw = (int)mn.w; a.nkeep0 = (int)mn.keep0; a.nkeep1 = (int)mn.keep1;
  };
  for (int q = L - 1; q >= 0; --q) {
    const int l = layer_of(q);
    BwdArgs a{};
    a.f = make_args(meta[l], arena, stats, ws, l, q, G, B, N, training, 0, eps);
    a.f.x = q == 0 ? p : P_out + (size_t)layer_of(q - 1) * plane;
    a.yv = P_out + (size_t)l * plane;
    a.lvv = LV + (size_t)l * plane;
    a.dy_chain = q == L - 1 ? nullptr : ws.dx[(q + 1) & 1];
    a.dP = dP ? (dP_stride < 0 ? (l == 0 ? dP : nullptr) : dP + (size_t)l * dP_stride) : nullptr;
    a.dMU = dMU ? dMU + (size_t)l * dMU_stride : nullptr;
    a.dLV = dLV ? dLV + (size_t)l * dLV_stride : nullptr;
    a.dx_out = ws.dx[q & 1];
    a.dfilm = ws.dfilm + (size_t)l * 4 * B * DPF_F;
    a.dprm = darena + meta[l].param_off;
    a.bna_sums = ws.bna_sums + (size_t)l * 2 * DPF_F * 4;
    a.ltab = precision >= 1 ? ws.ltab + (size_t)l * 2 * DPF_F * DPF_LTAB_ROW : nullptr;
    a.pend_store = precision >= 1 ? ws.pend + (size_t)l * 8 : nullptr;
    a.m12_rep = precision >= 1 ? ws.m12_rep + (size_t)l * DPF_M12_REP * 2 * DPF_F * 2 : nullptr;
    a.dw1_partial = precision >= 1 ? (float*)bwd_scratch + (size_t)l * p2_ctas * 2 * DPF_F * DPF_F : nullptr;
    if (q < L - 1) set_pending(a, q + 1);
    const unsigned short* wimg = ws.w1_bf16 + (size_t)l * tc_weight_image_elems_per_layer();
    if (merged_bwd) {
      ProfScope ps(CAT_BWD_P2, s);
      rc = launch_coupling_bwd_merged_tc(a, wimg, mode, precision == 2, ws.barriers_bwd + (size_t)l * 32, s);
      if (rc == DPF_OK) continue;
      if (rc != DPF_ERR_UNSUPPORTED) return rc;
      merged_bwd = false;                       // not available (stream capture before the probe ran): two launches
    }
    {
      ProfScope ps(CAT_BWD_P1, s);
      rc = precision >= 1 ? launch_coupling_bwd_tc(a, wimg, mode, 1, precision == 2, s) : launch_coupling_bwd_fp32(a, mode, 1, s);
    }
    if (rc) return rc;
    {
      ProfScope ps(CAT_BWD_P2, s);
      rc = precision >= 1 ? launch_coupling_bwd_tc(a, wimg, mode, 2, precision == 2, s) : launch_coupling_bwd_fp32(a, mode, 2, s);
    }
    if (rc) return rc;
  }
  {
    BwdArgs a{};
    a.f.B = B; a.f.N = N; a.f.G = G; a.f.training = training; a.f.eps = eps;
    set_pending(a, 0);
    // The tail of the pass is three independent kernels - the backward of all FiLM nets (252 CTAs, ~89 us, latency-bound),
    // the last layer's finalisation and the wgrad reduction over the per-CTA partials (2016 CTAs streaming 305 MB from
    // HBM, ~52 us).  The FiLM kernel is launched FIRST on `s` so that its few CTAs are placed before the reduction's
    // fill the chip from a side stream forked from / joined back into `s` (capturable).
    TailStreams* ts = precision >= 1 ? tail_streams() : nullptr;
    if (ts) {
      cudaEventRecord(ts->fork, s);
      cudaStreamWaitEvent(ts->side, ts->fork, 0);
    }
    {
      ProfScope ps(CAT_FILM_BWD, s);
      rc = launch_film_backward(arena, stats, darena, reinterpret_cast<const LayerMeta*>(meta_dev), g, ws.film,
                                ws.dfilm, dg, L, B, G, training, eps, s);
    }
    cudaStream_t s2 = ts ? ts->side : s;
    if (rc == DPF_OK) {
      ProfScope ps(CAT_BWD_FINAL, s);
      rc = launch_coupling_bwd_final(a, p, ws.dx[0], dp, s2);
      if (rc == DPF_OK && precision >= 1)
        rc = launch_dw1_reduce((const float*)bwd_scratch, p2_ctas, darena, reinterpret_cast<const LayerMeta*>(meta_dev), L, G, s2);
    }
    if (ts) cudaEventRecord(ts->join, ts->side);
    if (ts) cudaStreamWaitEvent(s, ts->join, 0);
    return rc;
  }
}
