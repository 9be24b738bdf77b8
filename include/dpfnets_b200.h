/*
 * dpfnets_b200.h — C ABI of libdpfnets_b200.so (sm_100a).
 *
 * Drop-in boundary for the data-parallel hot path of Regenerator/dpf-nets.  Every entry point
 * takes plain device pointers and sizes, runs asynchronously on the given CUDA stream
 * (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream) on the CURRENT
 * device, and returns 0 on success, a positive cudaError_t, or a negative DPF_ERR_* code.
 * dpf_last_error() returns the message of the last failure on the calling thread.
 * No torch types appear in any signature.  Citations are relative to the reference tree.
 */
#ifndef DPFNETS_B200_H
#define DPFNETS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define DPF_OK 0
#define DPF_ERR_BAD_ARG (-1)
#define DPF_ERR_NULL_PTR (-2)
#define DPF_ERR_UNSUPPORTED (-3)
#define DPF_ERR_ALIGN (-4)
#define DPF_ERR_BARRIER (-5)   /* a grid barrier of an earlier decoder pass timed out; that pass was aborted */

const char* dpf_last_error(void);
int dpf_version(void);
int dpf_device_check(void);
/* Kernels launched by this library since load (evidence for bench.py's gpu_launches). */
int dpf_launch_count(long long* count);
/* Per-kernel-class CUDA-event timing of the decoder entry points (off by default).
 * classes: 0 film_fwd, 1 moments, 2 fwd_stats, 3 fwd_apply, 4 bwd_p1, 5 bwd_p2, 6 bwd_final, 7 film_bwd */
/* option 0: merged train-mode forward (1 = on, default; 0 = two launches per layer).
 * option 1: fused all-layer eval-mode decoder, one launch for the whole stack (1 = on, default;
 *           0 = one launch per layer).  Both exist so that tests can compare the two forms.
 * option 2: programmatic dependent launch of the per-layer kernels (1 = on, default).
 * option 3: backward pass 2 with two tiles in flight per SM (1 = on, default; 0 = one tile per SM).
 * option 4: all-pairs Chamfer with both directions from one distance evaluation per point pair (1 = on, default;
 *           0 = one pass per direction like the reference's two launches).
 * option 5: backward of a coupling layer as ONE launch, pass 1 -> grid barrier -> pass 2 (1 = on; 0 = two
 *           PDL-chained launches per layer, default: measured 1 % faster, see csrc/decoder.cu). */
int dpf_set_option(int option, int value);
int dpf_profile_enable(int on);
int dpf_profile_collect(double* ms, long long* counts, int n);

/* ---- Chamfer / nearest-neighbour distance --------------------------------------------------
 * dpf_nndistance replaces nndistance() — lib/metrics/pytorch_structural_losses/src/nndistance.cuh:1
 * (kernel nndistance.cu:2-128; shim structural_loss.cpp:80-99; pybind bind.cpp:13).
 * xyz (b,n,3), xyz2 (b,m,3) fp32 -> result (b,n) squared NN distance of every xyz point in xyz2,
 * result_i (b,n) int32 argmin (lowest index on exact ties), result2/result2_i (b,m) the reverse. */
int dpf_nndistance(int b, int n, const float* xyz, int m, const float* xyz2, float* result,
                   int* result_i, float* result2, int* result2_i, void* stream);

/* Replaces nndistancegrad() — nndistance.cuh:2, nndistance.cu:129-154, structural_loss.cpp:101-124. */
int dpf_nndistance_grad(int b, int n, const float* xyz1, int m, const float* xyz2,
                        const float* grad_dist1, const int* idx1, const float* grad_dist2,
                        const int* idx2, float* grad_xyz1, float* grad_xyz2, void* stream);

/* Fused replacement of the Python all-pairs loop pairwise_CD — lib/networks/utils.py:90-117
 * (and _pairwise_EMD_CD_'s CD half, lib/metrics/evaluation_metrics.py:85-121):
 * out[i*S2+j] = mean_k min_l |A_i[k]-B_j[l]|^2 + mean_l min_k |A_i[k]-B_j[l]|^2
 * for rows i = row_start + t*row_step, t < n_rows (row sharding across GPUs).
 * A (S1,n,3), B (S2,m,3), out (S1,S2) fp32.  symmetric != 0 (A == B): only j >= i is written. */
int dpf_pairwise_cd(int S1, int S2, int n, int m, const float* A, const float* B, float* out,
                    int row_start, int row_step, int n_rows, int symmetric, void* stream);

/* Mirrors the upper triangle of an (S,S) matrix into the lower one (after dpf_pairwise_cd symmetric). */
int dpf_symmetrize_upper(float* M, int S, void* stream);

/* Generation scores on the device-resident all-pairs matrices, one launch + a 3-number result instead of the
 * reference's torch ops with host syncs — lib/networks/utils.py:120-121 (COV), :124-125 (MMD), :128-144 (KNN, k=1):
 * gg (S1,S1), gt (S1,S2), tt (S2,S2) fp32 row-major, gg / tt full symmetric matrices.
 * out[0] = COV(gt) = #distinct argmin_j gt[i,j] / S2;  out[1] = MMD(gt) = mean_j min_i gt[i,j];
 * out[2] = leave-one-out 1-NN accuracy on [[gg, gt], [gt^T, tt]] (exact ties: lowest index).
 * scratch: dpf_cd_scores_scratch_bytes() bytes of device memory. */
int dpf_cd_scores_scratch_bytes(int S1, int S2, long long* bytes);
int dpf_cd_scores(const float* gg, const float* gt, const float* tt, int S1, int S2, void* scratch, float* out,
                  void* stream);

/* Voxel-occupancy counts of get_voxel_occ_dist — lib/networks/utils.py:45-79 (the histogram behind JSD, :82-87):
 * pts (n_points,3) fp32; edges: res+1 DEVICE doubles (the reference's -0.5 + arange(res+1)/res); a point is
 * counted in cell (i,j,k) when edges[i] <= x < edges[i+1] holds for each coordinate, otherwise dropped.
 * hist: res^3 uint64 counts, overwritten. */
int dpf_voxel_hist(const float* pts, long long n_points, int res, const double* edges, unsigned long long* hist,
                   void* stream);

/* ---- Approximate EMD (soft auction, 9 levels) -----------------------------------------------
 * dpf_approxmatch replaces approxmatch() — src/approxmatch.cuh:6 (kernel approxmatch.cu:3-182, shim
 * structural_loss.cpp:22-37): xyz1 (b,n,3), xyz2 (b,m,3) -> match (b,m,n).  `temp` (b,2(n+m)) of the
 * reference signature is accepted and unused (may be NULL).
 * dpf_matchcost replaces matchcost() — approxmatch.cuh:7 (approxmatch.cu:184-224): out (b).
 * dpf_matchcost_grad replaces matchcostgrad() — approxmatch.cuh:8 (approxmatch.cu:229-291). */
int dpf_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, float* temp,
                    void* stream);
int dpf_matchcost(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* out,
                  void* stream);
int dpf_matchcost_grad(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match,
                       float* grad1, float* grad2, void* stream);
/* Fused EMD half of _pairwise_EMD_CD_ (lib/metrics/evaluation_metrics.py:85-121): out[i*S2+j] =
 * MatchCost(A_i, B_j) without materialising the dense match; rows row_start + t*row_step, t < n_rows. */
int dpf_pairwise_emd(int S1, int S2, int n, int m, const float* A, const float* B, float* out, int row_start,
                     int row_step, int n_rows, void* stream);

/* ---- Point decoder: stack of conditional affine-coupling layers ------------------------------
 * The reference has no native interface here; the replaceable unit is the nn.Module
 * (LocalCondRNVPDecoder.forward, lib/networks/decoders.py:54-72, over CondRealNVPFlow3D.forward,
 * lib/networks/flows.py:95-117).  These entry points are what that module's forward/backward bind.
 *
 * layer table: L rows of 8 int64 {param_off, stat_off, k, w, keep0, keep1, warp0, warp1}
 *   (meta_host = host copy, meta_dev = device copy); arena / stats layouts: csrc/coupling.cuh.
 * p (B,3,N), g (B,G) fp32; P_out, MU, LV (L,B,3,N) fp32 indexed by layer like the reference lists.
 * mode 0 = 'direct' (sampling), 1 = 'inverse' (training NLL); training != 0 = BatchNorm batch
 * statistics (+ running-stat update when update_stats != 0); precision 0 = fp32 CUDA cores,
 * 1 = bf16 tcgen05 tensor cores with fp32 accumulation. */
int dpf_decoder_workspace_bytes(int L, int G, int B, int N, long long* bytes);
int dpf_decoder_forward(const long long* meta_host, const long long* meta_dev, const float* arena,
                        float* stats, const float* p, const float* g, float* P_out, float* MU, float* LV,
                        void* workspace, int L, int G, int B, int N, int mode, int training,
                        int update_stats, int precision, float eps, void* stream);

/* dpf_decoder_forward with the outputs the flow NLL actually consumes (PointFlowNLL, lib/networks/losses.py:7-15:
 * samples[0], sum_l logvars[l]): MU may be NULL (per-layer mu not written); SLV (B,3,N), nullable, receives
 * sum_l LV[l] accumulated in the kernels' epilogues (the reference adds 63 full tensors with Python sum()). */
int dpf_decoder_forward_ex(const long long* meta_host, const long long* meta_dev, const float* arena,
                           float* stats, const float* p, const float* g, float* P_out, float* MU, float* LV,
                           float* SLV, void* workspace, int L, int G, int B, int N, int mode, int training,
                           int update_stats, int precision, float eps, void* stream);

/* Backward of dpf_decoder_forward = what torch.autograd derives for flows.py:95-117 (the reference
 * has no hand-written backward; formulas: SURVEY.md Appendix F).  `workspace` must be the buffer
 * the forward call used.  dP/dMU/dLV: cotangents of the stacked outputs (nullable), *_stride =
 * elements between layers (0 = one (B,3,N) block shared by all layers; dP_stride < 0 = dP is ONE (B,3,N) block
 * that belongs to layer 0 only - the cotangent of samples[0], all the flow NLL produces).  darena (n_params, arena
 * layout) and dg (B,G) are overwritten; dp (B,3,N) is optional (NULL = not needed). */
int dpf_decoder_backward(const long long* meta_host, const long long* meta_dev, const float* arena,
                         float* stats, const float* p, const float* g, const float* P_out, const float* LV,
                         const float* dP, long long dP_stride, const float* dMU, long long dMU_stride,
                         const float* dLV, long long dLV_stride, float* darena, long long n_params,
                         float* dg, float* dp, void* workspace, void* bwd_scratch, int L, int G, int B, int N,
                         int mode, int training, int precision, float eps, void* stream);
/* Synchronous health check of a forward workspace: *flag = layers whose merged-forward grid barrier
 * timed out (must be 0). */
int dpf_decoder_status(const void* workspace, int L, int G, int B, int N, int* flag);
/* Non-blocking (reads a pinned host flag): *failed = 1 when a grid barrier of this process has timed out - the
 * kernel that hit it trapped, and every later dpf_decoder_forward / dpf_decoder_backward returns DPF_ERR_BARRIER;
 * *coresident = 1 the merged forward's full grid was verified co-resident by a probe launch, 0 the probe failed
 * (the two-launch form is used instead), -1 not probed yet. */
int dpf_decoder_barrier_state(int* failed, int* coresident);
/* bwd_scratch: per-CTA wgrad partials of the tensor path (may be NULL for precision 0). */
int dpf_decoder_backward_scratch_bytes(int L, int B, int N, long long* bytes);

/* ---- PointNet cloud encoder, eval mode ------------------------------------------------------
 * Replaces PointNetCloudEncoder.forward in .eval() (lib/networks/encoders.py:9-28: SharedDot -> BatchNorm1d
 * -> ReLU x 4, widths 3 -> 64 -> 128 -> 256 -> 512) together with the max over points the models apply to
 * it (lib/networks/models.py:130-131): x (B,3,N) fp32 -> out (B,512) fp32, one fused kernel (BatchNorm
 * running statistics folded into the weights, bf16 tensor cores with fp32 accumulation).
 * weights: HOST array of 4 device pointers, SharedDot weights (out,in) row-major;
 * bn: HOST array of 16 device pointers, {weight, bias, running_mean, running_var} per layer.
 * workspace: dpf_pointnet_workspace_bytes() bytes, 256-byte aligned (folded tables + weight images). */
int dpf_pointnet_workspace_bytes(long long* bytes);
int dpf_pointnet_eval_forward(const float* x, int B, int N, const float* const* weights, const float* const* bn,
                              float bn_eps, void* workspace, float* out, void* stream);

/* ---- PointNet cloud encoder, train mode: last layer + max-pool ------------------------------
 * Replaces features.{sd2, sd2_bn, sd2_relu} in .train() (lib/networks/encoders.py:9-28: SharedDot 256 -> 512, BatchNorm1d
 * with batch statistics, ReLU) + the max over the points (lib/networks/models.py:130-131) WITHOUT materialising the
 * (B,512,N) activation: BatchNorm + ReLU are monotone per channel, so the pooled output is a function of max_n / min_n of
 * h = W h2 and of the batch statistics.  h2 (B,256,N) fp32 (post-ReLU output of the layer before), W (512,256) fp32 ->
 * stat (B,512,2) fp32 {mean_n h, sum_n (h - mean)^2} per (shape, channel), vmax / vmin (B,512) fp32, imax / imin (B,512)
 * int32 (lowest index on ties).
 * bf16 tensor cores with split operands (hi*hi + lo*hi + hi*lo, fp32 accumulation). */
int dpf_pointnet_pool_workspace_bytes(long long* bytes);
int dpf_pointnet_pool_forward(const float* h2, const float* W, int B, int N, void* workspace, float* stat,
                              float* vmax, float* vmin, int* imax, int* imin, void* stream);
/* in_tab (256, 8) fp32, nullable: per input channel {sc, sh, ...}; the operand is relu(sc h2 + sh), i.e. h2 is the
 * PRE-BatchNorm output of the layer before and its BatchNorm + ReLU are applied while the tile is loaded.
 * asum (B, 256) fp32, nullable: per (shape, input channel) sum of the operand over the points (the analytic backward's S). */
int dpf_pointnet_pool_forward_ex(const float* h2, const float* in_tab, const float* W, int B, int N, void* workspace, float* stat,
                                 float* vmax, float* vmin, int* imax, int* imin, float* asum, void* stream);

/* ---- PointNet cloud encoder, train mode: the narrow layers, forward and backward -------------
 * Replaces features.{init_sd, sd0, sd1}{, _bn, _relu} in .train() (lib/networks/encoders.py:9-28: SharedDot 3 -> 64 -> 128
 * -> 256, BatchNorm1d with batch statistics, ReLU) and torch.autograd through them.  Per layer only the pre-BatchNorm
 * output Z = W A_prev is stored; activations are recomputed on load from Z and the folded statistics (layer 0: from the
 * three input coordinates, its statistics follow analytically from the input moments).  Loader tables: 8 floats per
 * channel, loader 0 {a0, a1, a2, c, sub}: relu(a . x + c) - sub; loader 1 {sc, sh, sub}: relu(sc z + sh) - sub;
 * loader 2 {sc, sh, g, gm1, k2, mu}: (sc z + sh > 0 ? g dA : 0) - gm1 - k2 (z - mu)  (BatchNorm + ReLU backward).
 * bf16 tensor cores with split operands (hi*hi + lo*hi + hi*lo, fp32 accumulation). */
int dpf_pointnet_layer_image_bytes(int R, int K, long long* bytes);
/* W: matrix [R x K], element (r, k) at W[r * row_stride + k * col_stride] (W^T needs no copy); K in {64, 128, 256} */
int dpf_pointnet_layer_pack(const float* W, int R, int K, long long row_stride, long long col_stride, void* image, void* stream);
int dpf_pointnet_layer_groups(int B, int N, int Mout, int* groups);   /* rows of the stat buffer below */
/* out (B, Mout, N) = image[Mout x K] f(in); out nullable; row_off (Mout,) nullable; stat (groups, Mout, 3)
 * {count, mean, sum of squared deviations} per CTA work item, nullable (forward layers: loader 0 with K = 64, loader 1 with K = 128) */
int dpf_pointnet_layer_gemm(int loader, int K, const float* in0, const float* in1, const float* tab, const void* image,
                            int B, int N, int Mout, float* out, const float* row_off, float* stat, void* stream);
int dpf_pointnet_layer_wgrad_scratch_bytes(int MP, int NQ, int B, int N, long long* bytes);
/* out (MP, NQ) += sum over all points of P Q^T (accumulated: zero it first; scratch: per-CTA partial sums, 16-byte aligned).  gram = 0: P = loader 2 of (p_in0 = dA,
 * p_in1 = Z) (B,MP,N), Q = loader q_loader of q_in ((B,NQ,N), or x (B,3,N) with q_loader 0 and NQ = 64); gram = 1: P = Q =
 * loader 1 of p_in0 (B,256,N).  (MP, NQ, q_loader) in {(256,128,1), (128,64,0)} or gram with 256 x 256. */
int dpf_pointnet_layer_wgrad(int MP, int NQ, int gram, int q_loader, const float* p_in0, const float* p_in1, const float* p_tab,
                             const float* q_in, const float* q_tab, int B, int N, void* scratch, float* out, void* stream);
/* sums (C, 2) double += {sum dA [y > 0], sum dA [y > 0] (z - mu)} over all points, y = sc z + sh (table of loader 2) */
int dpf_pointnet_bn_bwd_sums(const float* dA, const float* Z, const float* tab, int B, int C, int N, double* sums, void* stream);
/* sums (C, 4) double += {sum dA m, sum dA m x0, sum dA m x1, sum dA m x2}, m = [a . x + c > 0] (table of loader 0) */
int dpf_pointnet_layer0_bwd_sums(const float* dA, const float* x, const float* tab, int B, int C, int N, double* sums, void* stream);

/* per-channel finalisation kernels (everything between the big kernels stays on the device):
 * input moments (9,) double {sum x_i, sum x_i x_j (00 01 02 11 12 22)} (accumulated: zero first); layer 0's analytic statistics
 * and loader-0 table; merge of the per-CTA statistics into {mean, biased var, istd} + loader-1 table; BatchNorm-backward sums
 * -> dgamma, dbeta, loader-2 table; layer 0's backward from its four sums per channel. */
int dpf_pointnet_input_moments(const float* x, int B, int N, double* moments, void* stream);
int dpf_pointnet_layer0_finalize(const double* moments, const float* W0, const float* gamma, const float* beta, int C, int B, int N,
                                 float eps, float* tab, float* stats, void* stream);
int dpf_pointnet_stats_finalize(const float* stat, int G, int C, int width, float count, const float* gamma, const float* beta, float eps,
                                float* tab, float* stats, void* stream);
int dpf_pointnet_bwd_finalize(const double* sums, const float* tab_fwd, const float* gamma, const float* stats, int C, int B, int N,
                              float* tab_bwd, float* dgamma, float* dbeta, void* stream);
int dpf_pointnet_layer0_bwd_finalize(const double* sums4, const double* moments, const float* W0, const float* gamma, const float* stats,
                                     int C, int B, int N, float* dW0, float* dgamma, float* dbeta, void* stream);
/* the sparse part of the pooled last layer's backward (the max-pool routes each (shape, channel) cotangent to one point):
 * T (C,256) += sum_b coef[b][c] relu(sc Z[b][:][n*] + sh) (zero it first), dA (B,256,N) += coef[b][c] W[c][:] at n* = idx[b][c] */
int dpf_pointnet_pool_sparse_backward(const float* Z, const float* tab, const long long* idx, const float* coef, const float* W,
                                      int B, int N, int C, float* T, float* dA, void* stream);


/* ---- Latent-side fused blocks -------------------------------------------
 * The shape-latent flows and feature heads work on (B,F) matrices with B = 32..64 rows; their module chains
 * (RealNVPFlow lib/networks/flows.py:163-213, FeatureEncoder lib/networks/encoders.py:31-83) are launch-bound.
 * dpf_bn_swish_*: y = swish(BatchNorm1d(x)) over the batch dimension (training: batch statistics, biased variance, running
 *   statistics updated with `momentum` and the unbiased variance when rm / rv are given; eval: rm / rv are used) and its
 *   backward (dx, dgamma, dbeta); save_mean / save_istd (F) carry the statistics from forward to backward.
 * dpf_latent_affine_*: flows.py:196-211 - logvar = log(eps + exp(raw_lv)), mu / logvar scattered to the warped positions
 *   (pos (G) int32: index into the warp list, -1 = kept), g_out = exp(-logvar/2)(g - mu) (inverse != 0) or
 *   exp(logvar/2) g + mu; backward from the cotangents of g_out / mu / logvar (each nullable). */
int dpf_bn_swish_forward(const float* x, const float* gamma, const float* beta, float* rm, float* rv, int B, int F,
                         float eps, float momentum, int training, float* y, float* save_mean, float* save_istd, void* stream);
int dpf_bn_swish_backward(const float* dy, const float* x, const float* gamma, const float* beta, const float* save_mean,
                          const float* save_istd, int B, int F, int training, float* dx, float* dgamma, float* dbeta, void* stream);
int dpf_latent_affine_forward(const float* g, const float* raw_mu, const float* raw_lv, const int* pos, int B, int G, int W,
                              float eps, int inverse, float* g_out, float* mu, float* lv, void* stream);
int dpf_latent_affine_backward(const float* dgo, const float* dmu_f, const float* dlv_f, const float* g, const float* raw_mu,
                               const float* raw_lv, const int* pos, int B, int G, int W, float eps, int inverse, float* dg,
                               float* draw_mu, float* draw_lv, void* stream);

/* ---- shape-latent flows: one RealNVPFlow coupling layer per kernel ----------------------------
 * Replaces RealNVPFlow.forward (lib/networks/flows.py:163-213: per branch Linear -> BatchNorm1d -> Swish -> Linear, then
 * logvar = log(eps + exp(.)), g_out = exp(+-logvar/2) g (+-) mu) and torch.autograd through it.  g (B,D), B <= 64; pos (D,)
 * int32: index in the warp list or -1; keep_idx (Kk,) int32; per branch b in {0: mu, 1: logvar}: Wa[b] (H,Kk), gamma[b],
 * beta[b] (H), rm[b], rv[b] (H; running statistics, updated in place in training mode when not null, read in eval mode),
 * Wb[b] (Wn,H), bb[b] (Wn).  H, Kk, Wn multiples of 32, else DPF_ERR_UNSUPPORTED.  The forward saves hpre (2,B,H), stat
 * (2,2,H) {mean, istd} and raw (2,B,Wn) for the backward.  One thread-block cluster of 8 CTAs per layer (hidden columns
 * split over the CTAs, exchange through distributed shared memory); fp32 CUDA-core arithmetic. */
/* development aid: globaltimer stamps (ns) of the phases of the most recent forward (row 0) / backward (row 1) launch, 2 x 12 */
int dpf_latent_flow_stamps(unsigned long long* out24);
/* DPF_OK when both calls below handle a layer of these sizes, DPF_ERR_UNSUPPORTED otherwise */
int dpf_latent_flow_supported(int B, int D, int H, int Kk, int Wn);
int dpf_latent_flow_forward(const float* g, const int* pos, const int* keep_idx, const float* const* Wa, const float* const* gamma,
                            const float* const* beta, float* const* rm, float* const* rv, const float* const* Wb,
                            const float* const* bb, int B, int D, int H, int Kk, int Wn, float bn_eps, float momentum, int training,
                            float eps, int inverse, float* g_out, float* mu, float* lv, float* hpre, float* stat, float* raw,
                            void* stream);
/* cotangents dgo, dmu_f, dlv_f (B,D; each nullable = zero) -> dg (B,D) and per branch dWa (H,Kk), dgamma, dbeta (H), dWb (Wn,H), dbb (Wn) */
int dpf_latent_flow_backward(const float* dgo, const float* dmu_f, const float* dlv_f, const float* g, const int* pos,
                             const int* keep_idx, const float* const* Wa, const float* const* gamma, const float* const* beta,
                             const float* const* Wb, int B, int D, int H, int Kk, int Wn, int training, float eps, int inverse,
                             const float* hpre, const float* stat, const float* raw, float* dg, float* const* dWa,
                             float* const* dgamma, float* const* dbeta, float* const* dWb, float* const* dbb, void* stream);


/* Fused AMSGrad step with the reference's exact update (lib/networks/optimizers.py:53-74):
 * denom = sqrt(max_exp_avg_sq or exp_avg_sq)/bc2 + eps; p -= wd*p + lr*(exp_avg/bc1)/denom.
 * vmax may be NULL (amsgrad off); bc1 = 1-beta1^t, bc2 = sqrt(1-beta2^t). */
int dpf_adam_step(float* p, const float* g, float* m, float* v, float* vmax, long long n, float lr, float b1,
                  float b2, float eps, float wd, float bc1, float bc2, void* stream);

/* The same update for n tensors in ceil(n/48) launches instead of the reference's per-parameter
 * Python loop (optimizers.py:20-74; ~170 parameter tensors in the generation model).  p/g/m/v/vmax:
 * HOST arrays of n device pointers (vmax NULL = amsgrad off), numel: host array of n sizes. */
int dpf_adam_step_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                        float* const* vmax, const long long* numel, float lr, float b1, float b2, float eps,
                        float wd, float bc1, float bc2, void* stream);

/* dpf_adam_step_multi with the hyper-parameters {lr, b1, b2, eps, wd, bc1, bc2} read from DEVICE memory when the
 * kernel runs, so that a CUDA graph holding the launch can be replayed with a new learning rate and step
 * count (the reference recomputes them on the host every step, optimizers.py:53-66, 89-97). */
int dpf_adam_step_multi_dev(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                            float* const* vmax, const long long* numel, const float* hyper_dev, void* stream);

/* tcgen05 self-test (tests/test_umma_gpu.py): D[128,ncols] = sum_k A_k B_k from raw shared-memory
 * operand images and descriptor fields; validates the UMMA layouts the coupling kernels rely on. */
int dpf_umma_selftest(const void* a_img, int a_bytes, const void* b_img, int b_bytes,
                      unsigned long long a_templ, unsigned long long b_templ, unsigned int idesc,
                      int num_k, int a_kstep, int b_kstep, int ncols, int use_bulk, float* d_out,
                      void* stream);

/* Measured pipe ceilings for bench.py's roofline fractions: `ctas` CTAs x 256 threads x `iters` rounds of 8 independent
 * operations per thread; which = 0: ex2.approx (MUFU; the EMD kernels' bound), 1: fma.rn.f32x2 (packed fp32 FMA; the
 * Chamfer kernels' bound).  *ops_per_launch = MUFU results / fp32 FMA lanes executed; time the call with CUDA events. */
int dpf_throughput_probe(int which, int ctas, int iters, float* scratch, long long* ops_per_launch, void* stream);

/* Debug probe: CTAs/SM the runtime reports for the merged (0) / plain (1) forward kernel at `smem` bytes. */
int dpf_debug_occupancy(int which, int smem, int* out);

#ifdef __cplusplus
}
#endif
#endif /* DPFNETS_B200_H */
