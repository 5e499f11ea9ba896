/*
 * pcrl.h -- C ABI of libpcrl.so: hand-written sm_100a kernels for pyrl's PointNet SAC/DrQ update path.
 *
 * The reference (lz1oceani/pointcloud_rl, `pyrl`) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md section 8b); its extension point is the Python registry.  This header is therefore the
 * boundary the host-side mirror (`pointcloud_rl_b200/`) binds with ctypes.  Every entry point names
 * the reference code whose arithmetic it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no host syncs, no
 *     allocation inside, CUDA-graph capturable
 *   - return value: 0 = ok, otherwise a PCRL_E* code; pcrl_last_error() gives a message
 *   - row-major fp32 everywhere unless stated otherwise
 *   - R = clouds (batch rows after augmentation), N = points per cloud, NP = N rounded up to a
 *     multiple of 128 (padding rows replicate the cloud's point 0, so they tie with it and the
 *     smallest-index rule keeps them out of the argmax), CP = channels padded to 8 or 16
 */
#ifndef PCRL_H_
#define PCRL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCRL_OK 0
#define PCRL_EINVAL 1   /* bad argument / unsupported shape */
#define PCRL_ECUDA 2    /* a CUDA runtime call or launch failed */
#define PCRL_EUNSUPPORTED 3

#define PCRL_ABI_VERSION 1

/* augmentation kinds for pcrl_stage_points */
#define PCRL_AUG_NONE 0
#define PCRL_AUG_JITTER 1 /* RandomJitterPoints: xyz += U(lo,hi) per coordinate (pcd_aug.py:307-322) */
#define PCRL_AUG_ROTZ 2   /* GlobalRotScaleTrans, rot only: one z-angle ~U(lo,hi) per cloud (pcd_aug.py:178-215) */
#define PCRL_AUG_SHIFT 3  /* GlobalRotScaleTrans, translation only (pn_shift.py): one t ~U(lo,hi)^3 per cloud
                             (pcd_aug.py:192-197, apply_rot_trans :84-123).  Axes: bits 8..10 of aug_kind select the
                             shifted axes (x,y,z), 0 = all three -- dm_control/pn_shift.py uses [0.04, 0, 0.04] */
#define PCRL_AUG_SHIFT_AXES(mask) (PCRL_AUG_SHIFT | ((mask) << 8))
#define PCRL_AUG_DOWNSAMPLE 4 /* RandomDownSample (pn_dropout.py; pcd_aug.py:240-257): one random subset of the points
                                 per call, shared by all clouds.  `noise` is REQUIRED: an int32 source map [N] (kept point
                                 -> itself, dropped point -> a kept one), from pcrl_downsample_map or the caller */

int pcrl_abi_version(void);
const char* pcrl_last_error(void);
/* number of SMs of the current device (grid sizing on the host side) */
int pcrl_sm_count(void);

/* Per-device context.  Everything the library keeps between calls -- the SM count and the internal side stream +
 * fork/join events pcrl_pointnet_bwd runs its weight-gradient GEMMs on -- lives in one context per CUDA device; there is
 * no other process-global state.  Calls use the context of the CURRENT device, creating it on first use; a caller that
 * wants to own the lifetime creates it up front and destroys it when done (after synchronising the device; captured CUDA
 * graphs that contain pcrl_pointnet_bwd reference the context's stream and must be destroyed first).
 * pcrl_create returns an opaque handle (0 on failure). */
int64_t pcrl_create(int device);
int pcrl_destroy(int64_t handle);

/* tf32 = 1 requests whose operands were not TMA-addressable (16-byte base, row pitch % 4 floats) run on the exact FFMA
 * kernel instead.  pcrl_tf32_fallbacks() counts them (process-wide); pcrl_set_strict_tf32(1) turns such a call into
 * PCRL_EUNSUPPORTED so a fast-mode shape can never lose the tensor path silently. */
int64_t pcrl_tf32_fallbacks(void);
int pcrl_set_strict_tf32(int on);

/* ---------------------------------------------------------------------------------------------
 * (1) Staging + augmentation.  Replaces GDict.repeat / repeat_interleave (drq.py:58-63,
 * array_ops.py:106-121), RandomJitterPoints / GlobalRotScaleTrans (pcd_aug.py) and
 * PointCloudBase.preprocess (pointnet.py:48-63) with one pass: every source cloud is read once and
 * `repeat` augmented copies are written, interleaved ([b0a0, b0a1, b1a0, ...]).
 *
 * inputs (channel-major, as the replay buffer stores them):
 *   xyz  f32 [B,3,N];  rgb u8 or f32 [B,3,N] (may be NULL; rgb_is_u8 selects /255);
 *   pos  u8 [B,n_pos,N] (may be NULL);  seg u8/bool [B,n_seg,N] (may be NULL)
 * randomness: if `noise` != NULL it is used verbatim (parity mode): JITTER -> f32 [B*repeat,3,N],
 *   ROTZ -> f32 [B*repeat] angles, SHIFT -> f32 [B*repeat,3] translations.  Otherwise Philox4x32-10 keyed by
 *   (seed, *counter_dev + stream_id).
 * outputs:
 *   xf   f32  [B*repeat, NP, CP]  point-major rows (xyz | rgb/255 | pos | seg | 0 pad); may be NULL when xh is given
 *        (the no-grad target branch of the tensor-core path never reads it)
 *   xh   bf16 tile images for the tcgen05 path, [B*repeat*NP/128][128x16] (may be NULL), holding
 *        hi parts of all channels, a constant-1 channel (bias) and the lo parts of xyz
 * ------------------------------------------------------------------------------------------- */
int pcrl_stage_points(const float* xyz, const void* rgb, int rgb_is_u8, const uint8_t* pos, int n_pos,
                      const uint8_t* seg, int n_seg, int B, int N, int repeat, int aug_kind, float aug_lo,
                      float aug_hi, const float* noise, uint64_t seed, const uint64_t* counter_dev,
                      uint32_t stream_id, float* xf, void* xh, int CP, void* stream);

/* RandomDownSample's random subset drawn on the device (Philox keyed by seed, *counter_dev, stream_id): n_drop =
 * int(N*drop_ratio) if fixed_ratio else uniform in [0, int(N*drop_ratio)), keep the N - n_drop points with the smallest
 * random keys (pcd_aug.py:244-251, array_ops.py:659-673).  src_map: int32 [N] for pcrl_stage_points.  N <= 4096. */
int pcrl_downsample_map(int N, float drop_ratio, int fixed_ratio, uint64_t seed, const uint64_t* counter_dev,
                        uint32_t stream_id, int32_t* src_map, void* stream);

/* Device-resident replay sampling (replaces ReplayMemory.sample's numpy `take` per key + the per-leaf H2D copies,
 * replay_buffer.py:297-322, dict_array.py:308-318): for every leaf l of a transition, dst_l[b] = src_l[idx[b]].
 * src_ptrs / dst_ptrs / row_bytes are DEVICE arrays of n_leaves entries (leaf base addresses and bytes per row),
 * idx is a device array of B ring positions (drawn on the host by the sampler, sampling_strategy.py:26-31). */
int pcrl_gather_transitions(const uint64_t* src_ptrs, const uint64_t* dst_ptrs, const int64_t* row_bytes, int n_leaves,
                            const int64_t* idx, int B, void* stream);

/* Host-side helper of the batch upload (GDict.to_torch, dict_array.py:308-318): memcpy between two HOST buffers on
 * `threads` threads of a small persistent pool (caller included).  Used to stage the pageable arrays `memory.sample`
 * returns into the pinned buffer the H2D copy reads; both pointers are host pointers. */
int pcrl_host_memcpy_mt(void* dst_host, const void* src_host, int64_t nbytes, int threads);

/* The whole batch upload in one call (GDict.to_torch(device=...), dict_array.py:308-318, for every leaf of
 * `memory.sample(B)`): leaf i = sizes[i] bytes from HOST pointer srcs_host[i] (NULL: the caller already wrote it) is
 * staged at pinned_host + offsets[i] with the pool memcpy above, and its host->device copy to landing_dev + offsets[i]
 * is enqueued on `stream` immediately, so the DMA of a leaf overlaps the staging of the next one.  pinned_host must be
 * page-locked; srcs_host / offsets / sizes are HOST arrays of n_leaves entries. */
int pcrl_upload_leaves(void* pinned_host, void* landing_dev, const void* const* srcs_host, const int64_t* offsets,
                       const int64_t* sizes, int n_leaves, int threads, void* stream);

/* ColorJitterPoints (pyrl/utils/augmentations/pcd_aug.py:269-303 -> torchvision ColorJitter on [B',3,1,N] uint8):
 * rgb u8 [B,3,N] -> out u8 [B,3,N].  One parameter set per CALL is shared by all clouds (so the num_aug copies of a
 * sample are identical: run it on the B source clouds and let pcrl_stage_points repeat them).
 * params_dev: NULL = draw from Philox (seed, *counter_dev, stream_id); else 8 device floats
 * [order0..3 (a permutation of 0 brightness, 1 contrast, 2 saturation, 3 hue), brightness, contrast, saturation, hue
 * factors] = the reference's draws (parity mode).  brightness / contrast / saturation / hue: the config magnitudes
 * (factor ranges [max(0,1-x), 1+x]; hue [-h, h], h <= 0.5).  Bit-exact vs torchvision 0.26 for uint8 input. */
int pcrl_color_jitter_points(const uint8_t* rgb, int B, int N, const float* params_dev, float brightness, float contrast,
                             float saturation, float hue, uint64_t seed, const uint64_t* counter_dev, uint32_t stream_id,
                             uint8_t* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (2) PointNet per-point shared MLP + max-pool.  Replaces ConvMLP (mlp.py:15-94; block_utils.py),
 * LN1d (nn_layer.py:192-225) and feature.max(-1) (pointnet.py:151):
 *   h0 = relu(W0 x + b0);  h1 = relu(LN(W1 h0; g1,be1,eps));  h2 = relu(LN(W2 h1; g2,be2,eps))
 *   pooled[r,c] = max_n h2[r,n,c], argmax[r,c] = smallest n attaining it.
 * weights: w0 [c1,C] (row stride C), b0 [c1], w1 [c2,c1], g1,be1 [c2], w2 [c3,c2], g2,be2 [c3].
 *
 * pcrl_pointnet_fwd_f32: exact-fp32 CUDA-core path (the parity path; argmax bit-exact up to fp32
 *   summation order).  workspace: pcrl_pointnet_fwd_f32_workspace(...) bytes.
 * pcrl_pointnet_fwd_bf16: fused tcgen05/TMEM path (bf16 operands, fp32 accumulate and LN statistics),
 *   reads `xh` tile images and pre-packed weight images (pcrl_pointnet_pack_weights).
 * outputs: pooled f32 [R,c3], argmax i32 [R,c3] (may be NULL on the bf16 path).
 * ------------------------------------------------------------------------------------------- */
/* bytes needed to process `clouds` clouds at a time (the call loops over chunks if given less than R) */
int64_t pcrl_pointnet_fwd_f32_workspace(int clouds, int NP, int c1, int c2, int c3);
int pcrl_pointnet_fwd_f32(const float* xf, int R, int N, int NP, int CP, int C, const float* w0, const float* b0,
                          const float* w1, const float* g1, const float* be1, const float* w2, const float* g2,
                          const float* be2, int c1, int c2, int c3, float ln_eps, float* pooled, int32_t* argmax,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* Reference-precision tier on the tensor cores: same contract as pcrl_pointnet_fwd_f32, layers 1 and 2 on the TF32 tcgen05
 * GEMM (10-bit mantissa operands, fp32 accumulate), everything else exact fp32 (layer 0, LayerNorm, max / argmax with
 * untruncated 64-bit keys).  Tolerance class: 1e-3 relative; argmax equal up to near-ties. */
int64_t pcrl_pointnet_fwd_tf32_workspace(int clouds, int NP, int c1, int c2, int c3);
int pcrl_pointnet_fwd_tf32(const float* xf, int R, int N, int NP, int CP, int C, const float* w0, const float* b0,
                           const float* w1, const float* g1, const float* be1, const float* w2, const float* g2,
                           const float* be2, int c1, int c2, int c3, float ln_eps, float* pooled, int32_t* argmax,
                           void* workspace, int64_t workspace_bytes, void* stream);

int64_t pcrl_pointnet_wpack_bytes(int c1, int c2, int c3);
int pcrl_pointnet_pack_weights(const float* w0, const float* b0, const float* w1, const float* g1, const float* be1,
                               const float* w2, const float* g2, const float* be2, int C, int c1, int c2, int c3,
                               int rgb_u8 /* xh holds raw 0..255 rgb: fold 1/255 into w0 */, void* wpack,
                               void* stream);
/* Same, restricted to one of the two images inside `wpack`: PCRL_WPACK_FWD = what pcrl_pointnet_fwd_bf16 reads (on the
 * critical path of every encode), PCRL_WPACK_BWD = what pcrl_pointnet_bwd's recompute reads (needed only ~0.4 ms later,
 * so the caller can pack it on a side stream, and not at all before an encode that has no backward). */
#define PCRL_WPACK_FWD 1
#define PCRL_WPACK_BWD 2
int pcrl_pointnet_pack_weights_part(const float* w0, const float* b0, const float* w1, const float* g1, const float* be1,
                                    const float* w2, const float* g2, const float* be2, int C, int c1, int c2, int c3,
                                    int rgb_u8, int which /* PCRL_WPACK_FWD | PCRL_WPACK_BWD */, void* wpack, void* stream);
int pcrl_pointnet_fwd_bf16(const void* xh, int R, int N, int NP, const void* wpack, int c1, int c2, int c3,
                           float ln_eps, uint64_t* pool_keys /* [R,c3] scratch: zero before the first call, every call leaves it zeroed */, float* pooled, int32_t* argmax,
                           void* stream);
/* Same, but output cloud r reads the staged tiles of source cloud r * src_cloud_stride: DrQ's actor step encodes the
 * first of the num_aug staged copies of every sample (drq.py:115) without gathering them first. */
int pcrl_pointnet_fwd_bf16_strided(const void* xh, int R, int src_cloud_stride, int N, int NP, const void* wpack, int c1,
                                   int c2, int c3, float ln_eps, uint64_t* pool_keys, float* pooled, int32_t* argmax,
                                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * (3) Sparse backward of (2) through the saved argmax (autograd of pointnet.py:151 + ConvMLP).
 * Only points that won the max for at least one channel with a non-zero pooled value carry
 * gradient; they are compacted, their forward is recomputed and the dense backward runs on the
 * compacted set.  Gradients are ACCUMULATED into dw0.. (caller zeroes them).
 *   capacity = max active points the workspace is sized for (<= R*c3).
 * ------------------------------------------------------------------------------------------- */
int64_t pcrl_pointnet_bwd_workspace(int R, int NP, int c1, int c2, int c3, int CP);
int pcrl_pointnet_bwd(const float* xf, int R, int N, int NP, int CP, int C, const float* pooled,
                      const int32_t* argmax, const float* dpooled, const float* w0, const float* b0, const float* w1,
                      const float* g1, const float* be1, const float* w2, const float* g2, const float* be2, int c1,
                      int c2, int c3, float ln_eps, float* dw0, float* db0, float* dw1, float* dg1, float* dbe1,
                      float* dw2, float* dg2, float* dbe2, void* workspace, int64_t workspace_bytes, int tf32,
                      const void* xh /* optional: bf16 tile images of the same staged points */,
                      const void* wpack /* optional: packed weights; with xh the recompute runs on the fused tcgen05 kernel */,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * (4) Dense layers.  y = act(x W^T + b): nn.Linear (+ReLU) of LinearMLP (mlp.py:85-100) and
 * PointNet.final_mlp[0] (pointnet.py:110).  x [M,K] with row stride ldx, w [Nout,K], y [M,Nout] row
 * stride ldy.  relu: 0/1.
 * pcrl_linear_bwd: given dy [M,Nout] (already masked by the caller's activation) computes
 *   dw += dy^T x, db += colsum(dy) (skipped when dw / db are NULL), and (if dx != NULL) dx = dy W
 *   (dx row stride lddx).  relu_mask (optional, [M,K] row stride ld_mask): the post-ReLU input x itself --
 *   dx is zeroed where relu_mask <= 0, i.e. the previous layer's ReLU backward fused into this GEMM's epilogue.
 * ------------------------------------------------------------------------------------------- */
int pcrl_linear_fwd(const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int M, int K,
                    int Nout, int relu, int tf32, void* stream);
int pcrl_linear_bwd(const float* x, int ldx, const float* w, const float* dy, int lddy, float* dw, float* db,
                    float* dx, int lddx, const float* relu_mask, int ld_mask, int M, int K, int Nout, int tf32,
                    void* stream);
/* tf32 != 0: run the GEMMs on the tcgen05 TF32 tensor-core kernel (TMA-fed, fp32 inputs consumed as TF32)
 * whenever the operands meet TMA's alignment rules (16-byte base, row pitch % 4 == 0), else and for
 * tf32 == 0 the exact-fp32 FFMA kernel runs.
 *
 * The raw tensor-core GEMM (tests / microbenchmarks): C[i][j] (op)= sum_l A(i,l) B(l,j) (+bias[j]) (ReLU);
 * a_mn == 0: A[i*lda+l], a_mn == 1: A[l*lda+i]; b_mn == 0: B[j*ldb+l], b_mn == 1: B[l*ldb+j];
 * mode 0 store, 1 accumulate, 2 atomic accumulate (required for split_k > 1). */
int pcrl_gemm_tf32(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, const float* bias, float* C,
                   int ldc, int M, int N, int K, int relu, int mode, int split_k, void* stream);
/* dy *= (y > 0) in place: ReLU backward on the saved post-activation */
int pcrl_relu_bwd(float* dy, const float* y, int64_t n, void* stream);
/* out[m,j] = a[m,j] + b[m,j] for j < width (summing the two Q heads' input gradients) */
int pcrl_add_cols(const float* a, int lda, const float* b, int ldb, float* out, int ldo, int M, int width,
                  void* stream);

/* LayerNorm over the last dim of [M,D] (nn.LayerNorm, pointnet.py:110, eps 1e-5): y = (x-mu)*rstd*g + b.
 * fwd saves xhat [M,D] and rstd [M] for bwd.  bwd: dg += sum dy*xhat, db += sum dy, dx (may alias dy). */
int pcrl_layernorm_fwd(const float* x, const float* g, const float* b, float* y, int ldy, float* xhat, float* rstd,
                       int M, int D, float eps, void* stream);
int pcrl_layernorm_bwd(const float* dy, int lddy, const float* xhat, const float* rstd, const float* g, float* dg,
                       float* db, float* dx, int M, int D, void* stream);

/* copy columns: dst[m, dst_off + j] = src[m * src_row_step, j]  for j < width (Visuomotor's torch.cat of
 * feature | robot state | action, visuomotor.py:130-144; src_row_step = repeat_interleave / first-aug slicing) */
int pcrl_copy_cols(const float* src, int lds, int src_row_div, int src_row_mul, float* dst, int ldd, int dst_off,
                   int M, int width, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (5) Policy head.  TanhGaussianHead 'max-entropy' (gaussian.py:23-50,83-87; distributions.py:45-127):
 *   mu,ls = chunk(out,2); std = exp(clamp(ls,lo,hi)); u = mu + std*eps; a = tanh(u)*scale+bias;
 *   neglogp = -sum_j [ logN(u;mu,std) - log(scale*(1-tanh(u)^2)+1e-6) ].
 * eps: injected [M,A] or NULL -> Philox normal.  eps_out [M,A] always written (needed by bwd).
 * bwd: given da [M,A] and the scalar g_nlp = dL/dneglogp (same for every row) -> dout [M,2A].
 * ------------------------------------------------------------------------------------------- */
int pcrl_tanh_gaussian_fwd(const float* out, int M, int A, float ls_lo, float ls_hi, float scale, float bias,
                           const float* eps, uint64_t seed, const uint64_t* counter_dev, uint32_t stream_id,
                           float* action, int ld_action, float* neglogp, float* eps_out, void* stream);
int pcrl_tanh_gaussian_bwd(const float* out, const float* eps, const float* daction, int ld_daction, float g_nlp,
                           int M, int A, float ls_lo, float ls_hi, float scale, float* dout, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (6) Targets and losses (sac.py:125-137,182-195; drq.py:75-90).  `scalars` is a device array of
 * PCRL_NUM_SCALARS floats that receives everything update_parameters() logs; one D2H copy per update.
 *   alpha_dev: device float holding the cached python float `self.alpha`.
 * ------------------------------------------------------------------------------------------- */
#define PCRL_S_CRITIC_LOSS 0
#define PCRL_S_MAX_ABS_ERR 1
#define PCRL_S_Q 2
#define PCRL_S_Q_TARGET 3
#define PCRL_S_CRITIC_GRAD_SQ 4
#define PCRL_S_ACTOR_LOSS 5
#define PCRL_S_ALPHA_LOSS 6
#define PCRL_S_ENTROPY 7
#define PCRL_S_ACTOR_GRAD_SQ 8
#define PCRL_S_ALPHA 9
#define PCRL_S_ALPHA_GRAD 10
#define PCRL_NUM_SCALARS 16

/* y[r] = r*reward_scale + (1-done)*gamma*(min(q0,q1) + alpha*neglogp); then mean over each group of
 * `group` consecutive rows, broadcast back (DrQ, drq.py:84-87; group=1 for SAC).  qt [R,2] row-major. */
int pcrl_td_target(const float* qt, const float* neglogp, const float* rewards, const uint8_t* dones, int B,
                   int group, float gamma, float reward_scale, int ignore_dones, const float* alpha_dev, float* y,
                   void* stream);
/* critic loss = mse(q, y)*2 over q [R,2]; writes dq [R,2] and scalars 0..3 */
int pcrl_critic_loss(const float* q, const float* y, int R, float* dq, float* scalars, void* stream);
/* actor loss = -(mean(min_h q) + alpha*mean(neglogp)); alpha loss = exp(log_alpha)*(entropy - target_entropy).
 * writes dq [M,2] (-1/M routed to the min head), scalars 5..7, *g_nlp_out... (g_nlp = -alpha/M is computed by
 * the caller from alpha_dev inside pcrl_tanh_gaussian_bwd_dev), and dlog_alpha[0]. */
int pcrl_actor_loss(const float* q, const float* neglogp, int M, const float* alpha_dev, const float* log_alpha,
                    float target_entropy, float* dq, float* dlog_alpha, float* scalars, void* stream);
/* as pcrl_tanh_gaussian_bwd but g_nlp = -(*alpha_dev)/M read on the device */
int pcrl_tanh_gaussian_bwd_dev(const float* out, const float* eps, const float* daction, int ld_daction,
                               const float* alpha_dev, int M, int A, float ls_lo, float ls_hi, float scale,
                               float* dout, void* stream);
/* alpha_dev[0] = exp(log_alpha[0]); scalars[PCRL_S_ALPHA] = alpha_dev[0]  (sac.py:195) */
int pcrl_refresh_alpha(const float* log_alpha, float* alpha_dev, float* scalars, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (7) Fused multi-tensor Adam + grad-norm + Polyak over flat buffers.  Replaces torch.optim.Adam.step
 * built by build_optimizer (optimizer_utils.py:31-64), ExtendedModuleBase.grad_norm
 * (module_utils.py:40-45) and soft_update (ops.py:60-90).
 *   step_dev: device int32 step counter, incremented by the kernel (bias correction reads it).
 *   grad_scale: multiplies every gradient first (1/world_size after an all-reduce(sum)).
 *   gradsq_out: device float, receives sum(g^2) of the scaled gradients (sqrt on the host).
 *   polyak: if target != NULL, target[i] = target[i]*(1-tau) + p_new[i]*tau for i in [poly_begin, poly_end)
 *           of THIS buffer, target indexed from 0.
 * ------------------------------------------------------------------------------------------- */
int pcrl_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float grad_scale, int32_t* step_dev, float* gradsq_out, float* target,
                   int64_t poly_begin, int64_t poly_end, float tau, void* stream);
/* One optimizer group updated in several calls (data-parallel runs: the part of the group whose gradient all-reduce has
 * landed steps while the rest is still in flight).  Exactly one call per update passes bump_step = 1 (it must come first:
 * the bias corrections read the bumped count) and zero_gradsq = 1; every part adds its share to *gradsq_out. */
int pcrl_adam_step_part(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                        float eps, float grad_scale, int32_t* step_dev, int bump_step, float* gradsq_out, int zero_gradsq,
                        float* target, int64_t poly_begin, int64_t poly_end, float tau, void* stream);
int pcrl_polyak(float* target, const float* source, int64_t n, float tau, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Gradient all-reduce over NVLink peer memory (replaces DistributedDataParallel's NCCL buckets, module_utils.py:322-349,
 * for the ranks of one node).  In-place SUM of elements [off, off + n) of a SYMMETRIC fp32 buffer: every rank holds a
 * buffer of the same layout and bufs_dev[r] (a DEVICE array of `world` addresses) is rank r's buffer as mapped into THIS
 * process (torch.distributed._symmetric_memory / cudaIpcOpenMemHandle); flags_dev[r] likewise addresses rank r's flag
 * block (pcrl_p2p_flag_bytes() bytes, zero-initialised once, symmetric too).  state_dev: pcrl_p2p_state_bytes() bytes
 * of LOCAL zero-initialised device memory (per-channel call counters).  channel < PCRL_P2P_CHANNELS: reductions that can
 * be in flight at the same time (different streams) must use different channels; every rank must issue the same
 * sequence of calls per channel.  One kernel, stream-ordered, CUDA-graph-capturable; the result is bit-identical on
 * all ranks (each slice is summed once, in rank order, and broadcast).  max_ctas: 0 = default (32). */
#define PCRL_P2P_CHANNELS 16
int64_t pcrl_p2p_flag_bytes(void);
int64_t pcrl_p2p_state_bytes(void);
int pcrl_p2p_allreduce(const uint64_t* bufs_dev, const uint64_t* flags_dev, int rank, int world, int64_t off, int64_t n,
                       int channel, int32_t* state_dev, int max_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCRL_H_ */
