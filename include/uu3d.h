/* uu3d.h — C ABI of the B200-native uplift-and-upsample transformer hot path.
 *
 * The reference (goldbricklemon/uplift-upsample-3dhpe) has no FFI of its own: the path sits
 * behind three Python-level contracts.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference root):
 *
 *   uu_create / uu_destroy      build_uplift_upsample_transformer(config)
 *                               common/net/uplift_upsample_transformer_constructor.py:14-50
 *   uu_set_weight / uu_get_weight / uu_weight_info
 *                               weight_io.load_weights_with_callback (by group name, then position)
 *                               common/utils/weight_io.py:76-263 ; model.get_weights/set_weights train.py:400
 *   uu_forward / uu_forward_host
 *                               model([x2d, stride_mask], training=False) -> (full, central)
 *                               common/net/uplift_upsample_transformer.py:388-421 ; caller eval.py:63-71
 *   uu_train_config / uu_train_forward_backward / uu_grad_buffer / uu_adamw_step ; uu_train_step (all of it in one call)
 *                               train_step(): loss, gradients, AdamW                        train.py:464-506, :403-415
 *   uu_optimizer_state          the optimizer part of tf.train.Checkpoint (Adam moments, EMA copy)           train.py:417-430
 *   uu_comm_unique_id / uu_comm_init / uu_comm_destroy / uu_allreduce_gradients
 *                               the gradient exchange of the reference's tf.distribute strategy (train.py:464-506 runs
 *                               under strategy.run): NCCL sum all-reduce over NVLink
 *   uu_stride_mask              stride-mask rule of the data generators
 *                               common/dataset/uplifiting_dataset.py:377-394
 *
 * Conventions: every function returns 0 on success and non-zero on failure, never throws;
 * uu_last_error() returns a thread-local description of the last failure.  The caller owns all
 * I/O buffers; the library owns weights, workspace and optimizer state.  Device entry points are
 * asynchronous and stream-ordered on `stream` (a cudaStream_t passed as void*, NULL = default
 * stream).  One model instance per device; calls on one instance are serialised by the caller.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef UU3D_H
#define UU3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UU_MAX_STRIDED 8

typedef struct uu_model uu_model;

/* Hyper-parameters consumed by the constructor (config/ *.json keys in comments). */
typedef struct uu_spec {
  int32_t n_tok;                 /* SEQUENCE_LENGTH (71 for "N=351", 41 for "N=81") */
  int32_t n_joints;              /* NUM_KEYPOINTS */
  int32_t d_spatial;             /* SPATIAL_EMBED_DIM */
  int32_t d_temporal;            /* TEMPORAL_EMBED_DIM */
  int32_t spatial_depth;         /* SPATIAL_TRANSFORMER_BLOCKS */
  int32_t temporal_depth;        /* TEMPORAL_TRANSFORMER_BLOCKS */
  int32_t num_heads;             /* NUM_HEADS */
  int32_t h_spatial;             /* int(SPATIAL_EMBED_DIM * MLP_RATIO) */
  int32_t h_temporal;            /* int(TEMPORAL_EMBED_DIM * MLP_RATIO) */
  int32_t n_strided;             /* len(STRIDES) */
  int32_t strides[UU_MAX_STRIDED];
  int32_t pad_left[UU_MAX_STRIDED];   /* PADDINGS[i][0] (null => 1) */
  int32_t pad_right[UU_MAX_STRIDED];  /* PADDINGS[i][1] (null => 1) */
  int32_t has_strided_input;     /* constructor.py:16-21 */
  int32_t first_strided_token_attention_layer; /* FIRST_STRIDED_TOKEN_ATTENTION_LAYER */
  int32_t full_output;           /* not USE_REFINE */
} uu_spec;

/* Arithmetic of the dense contractions. */
enum { UU_PRECISION_FP32 = 0,    /* CUDA-core fp32 end to end, <= 1e-4 abs of the reference's float32 outputs */
       UU_PRECISION_BF16 = 1,    /* bf16 operands on tensor cores (tcgen05 / mma.sync), fp32 accumulation; LayerNorm,
                                    softmax and residual adds in fp32 registers, activations stored as bf16 */
       UU_PRECISION_TF32 = 2 };  /* the fp32 schedule (fp32 activations, LayerNorm, softmax, spatial stage) with the large
                                    GEMMs on tcgen05 kind::tf32 (TF32 products, fp32 accumulation): the intermediate point
                                    of the accuracy / throughput curve, what TensorFlow >= 2.4 computes on Ampere+ GPUs */

const char* uu_last_error(void);
int uu_version(void);

int uu_create(const uu_spec* spec, int device, uu_model** out);
int uu_destroy(uu_model* m);
int uu_set_precision(uu_model* m, int precision);
int uu_get_precision(const uu_model* m);

/* Weight inventory in Keras .h5 order: groups in `layer_names` order, tensors by position. */
int uu_weight_count(const uu_model* m);
int64_t uu_param_count(const uu_model* m);
int uu_weight_info(const uu_model* m, int flat_index, char* group, int group_cap, int* index_in_group,
                   int64_t shape[4], int* rank);
/* host fp32 buffers; shape/rank are checked against the model (weight_io.py:219-232) */
int uu_set_weight(uu_model* m, const char* group, int index, const float* host, const int64_t* shape, int rank);
int uu_get_weight(uu_model* m, const char* group, int index, float* host, int64_t capacity);

/* Forward pass on device buffers.
 *   x2d    : fp32 (B, n_tok, n_joints, 2) row-major.  Frames with mask == 0 are never read
 *            (equivalent to the caller-side mask multiply of eval.py:67).
 *   mask   : uint8 (B, n_tok), non-zero on tokens that carry a 2-D pose; may be NULL when the
 *            model has no strided input.
 *   full   : fp32 (B, n_tok, n_joints, 3) or NULL;  central : fp32 (B, n_joints, 3).
 * On a capturable stream (not NULL / legacy) the bf16 schedule is captured into a CUDA graph the second time a
 * (B, buffers, stream) combination is seen and replayed afterwards (UU_GRAPH=0 disables); results are identical. */
int uu_forward(uu_model* m, const float* x2d, const uint8_t* mask, int B, float* full, float* central, void* stream);
/* Same call with HOST buffers (pinned recommended): H2D copy (for large bf16 batches in chunks on a second stream,
 * overlapped with the spatial kernel of the previous chunk), forward, D2H copy, stream sync.
 * full == NULL: the full-sequence head still runs on the device (as in the reference) but is not copied back. */
int uu_forward_host(uu_model* m, const float* x2d, const uint8_t* mask, int B, float* full, float* central);

/* Sliding windows cut ON THE DEVICE from one video (replaces the host-side window
 * materialisation of common/dataset/uplifiting_dataset.py:341-394 + eval.py:63-71 — a 71x input blow-up).
 * video2d: (T, n_joints, 2) float; centers: B int32 frame indices in [0, T).  Window b, token k reads frame
 * (k - n_tok/2) * s_out + centers[b]; frames outside the video are padded by repeating the first / last in-range
 * strided sample (pad_copy = 1, PADDING_TYPE "copy") or with zeros (pad_copy = 0).  The stride mask is the globally
 * aligned one of the eval generator, (frame mod s_in == 0) with floor-mod, built on the device; tokens it rejects are
 * never read.  Outputs as uu_forward.  The _host variant takes host pointers for every buffer. */
int uu_forward_video(uu_model* m, const float* video2d, int T, const int32_t* centers, int B, int s_out, int s_in,
                     int pad_copy, float* full, float* central, void* stream);
int uu_forward_video_host(uu_model* m, const float* video2d, int T, const int32_t* centers, int B, int s_out, int s_in,
                          int pad_copy, float* full, float* central);
/* Test-time flip augmentation (eval.py:154-180): both outputs become
 * (f(x) + unflip(f(flip(x)))) / 2, flip = negate x and gather joints by `order` (config AUGM_FLIP_KEYPOINT_ORDER,
 * n == n_joints entries).  The flip is applied where the spatial kernel reads the key-points; nothing is copied. */
int uu_set_flip_order(uu_model* m, const int32_t* order, int n);
int uu_forward_tta(uu_model* m, const float* x2d, const uint8_t* mask, int B, float* full, float* central, void* stream);
int uu_forward_video_tta(uu_model* m, const float* video2d, int T, const int32_t* centers, int B, int s_out, int s_in,
                         int pad_copy, float* full, float* central, void* stream);
/* Key-frame interpolation (common/dataset/action_wise_eval.py:76-100): pred/out are
 * (n, values_per_frame) device arrays in evaluation order, frame_indices the video frame of each row (videos
 * concatenated; a video ends where the index does not increase).  Rows whose index is a multiple of keyframe_stride
 * are copied, rows between two key frames of a video are linear in list position, rows after the last key frame of
 * a video copy it. */
int uu_op_keyframe_interp(const float* pred, const int32_t* frame_indices, int n, int keyframe_stride, int values_per_frame,
                          float* out, void* stream);
/* Training-data synthesis on the device (uplifiting_dataset.py:669-761 tf_world_to_cam_and_2d):
 * seq3d (B, points_per_sample, 3) world coordinates, cams (B, 18) = unit quaternion (w, x, y, z), translation (3),
 * intrinsics (resolution 2, focal length 2, centre 2, radial 3, tangential 2).  cam3d (B, P, 3) and / or p2d (B, P, 2)
 * receive the camera-space poses and their Human3.6M projection (x/z, y/z clamped to [-1, 1] as in the reference). */
int uu_op_world_to_cam_and_2d(const float* seq3d, const float* cams, int B, int points_per_sample, float* cam3d, float* p2d,
                              void* stream);
/* Evaluation metrics on the device (common/dataset/metrics.py:13-81): root-aligned MPJPE and
 * N-MPJPE (root alignment + optimal per-pose scale).  pred (n, n_joints, 3) and gt (n, n_joints, 4 = x, y, z, valid) are
 * device arrays; jpe / njpe (optional, device, (n, n_joints)) receive the per-joint errors with -1 at invalid joints
 * (the reference's normalize=False form); result_host[3] = {mpjpe, nmpjpe, number of valid joints}.  Synchronous.
 * (P-MPJPE needs a per-pose SVD, metrics.py:84-117, and stays on the host.) */
int uu_op_pose_metrics(const float* pred, const float* gt, int n, int n_joints, int root, float* jpe, float* njpe,
                       double* result_host, void* stream);
/* The gather alone: src (B*n_tok int32 source frame, -1 = zeros), mask (B*n_tok), and, when x2d != NULL, the
 * materialised (B, n_tok, n_joints, 2) windows exactly as the reference generator yields them (unmasked). */
int uu_op_window_gather(const float* video2d, int T, const int32_t* centers, int B, int n_tok, int n_joints, int s_out,
                        int s_in, int pad_copy, int32_t* src, uint8_t* mask, float* x2d, void* stream);

/* Number of kernels of this library launched by the most recent forward on `m`. */
int uu_last_launch_count(const uu_model* m);

/* Per-kernel-kind device timing of the most recent forward: with profiling on, every launch is
 * bracketed by CUDA events on its own stream; uu_get_profile sums them by kind (ms) after the
 * forward.  Used by bench.py for the roofline of the dominant kernel; off by default. */
enum { UU_KIND_GATHER = 0, UU_KIND_SPATIAL, UU_KIND_TOKEN_FILL, UU_KIND_LAYERNORM, UU_KIND_ATTENTION,
       UU_KIND_GEMM_TC, UU_KIND_GEMM_F32, UU_KIND_CAST, UU_KIND_COUNT };
int uu_set_profiling(uu_model* m, int on);
int uu_get_profile(uu_model* m, float* ms_by_kind, int32_t* launches_by_kind, int n_kinds);

/* ---- training step (train.py:464-506, optimizer train.py:403-415) -------------------------------
 * uu_train_config      : BATCH_SIZE (the GLOBAL constant the losses are divided by), ROOT_KEYTPOINT,
 *                        LOSS_WEIGHT_CENTER / _SEQUENCE, DROP_PATH_RATE[3]; droppath_mode 0 = off, 1 = on
 *                        (counter-based RNG keyed by seed and step).
 * uu_train_forward_backward : forward (training=True), loss, gradients of every trainable tensor into the
 *                        flat gradient buffer.  gt3d: fp32 (B, n_tok, n_joints, 3) absolute key-points (root-
 *                        centring happens inside, train.py:467).  loss_dev: device pointer to one float.
 *                        Data-parallel: each rank passes its B_local windows; gradients combine by SUM.
 * uu_grad_buffer       : flat fp32 gradient buffer (same layout as the parameters) for the all-reduce.
 * uu_adamw_step        : tfa.optimizers.AdamW update with host-evaluated schedule values lr_t, wd_t and
 *                        Adam step t = iterations + 1; ema_decay < 0 disables the EMA copy. */
int uu_train_config(uu_model* m, int global_batch, int root_keypoint, float w_center, float w_sequence,
                    const float* drop_path_rate3, int droppath_mode, uint64_t seed);
int uu_train_forward_backward(uu_model* m, const float* x2d, const uint8_t* mask, const float* gt3d, int B,
                              int64_t step, float* loss_dev, void* stream);
/* Random token masking of the temporal input (net:287-311, :336-338; TOKEN_MASK_RATE > 0 with LEARNABLE_MASKED_TOKEN
 * false: masked value 0).  Training only; the central token is never masked; drawn per step from the counter RNG seeded
 * in uu_train_config.  uu_get_token_mask copies the 0 / 1 keep factors [B * n_tok] of the last step to the host. */
int uu_train_set_token_masking(uu_model* m, float rate);
int uu_get_token_mask(uu_model* m, float* host, int64_t capacity);
/* Arithmetic of the training GEMMs.  0 (default): fp32 on CUDA cores, gradients within 2e-3 of fp32 autograd.
 * 1: forward, dgrad and wgrad GEMMs of the temporal / strided blocks on tcgen05 kind::tf32 (fp32 data, TF32 products, fp32
 *    accumulation) -- what TensorFlow 2.4 itself does on Ampere-or-newer GPUs unless
 *    tf.config.experimental.enable_tensor_float_32_execution(False) is called -- and their attention on mma.sync with bf16
 *    hi + lo operand planes (16 mantissa bits per operand); the spatial blocks and everything that is not a GEMM stay fp32. */
int uu_train_set_math(uu_model* m, int mode);
int uu_grad_buffer(uu_model* m, float** dev_ptr, int64_t* n_floats);
/* Optimizer state for checkpoint / resume (the reference checkpoints its optimizer with tf.train.Checkpoint, train.py:417-
 * 430): device pointer to the flat fp32 buffer `which` = 0 (Adam first moment), 1 (Adam second moment) or 2 (EMA copy of the
 * weights), same layout as the parameters / uu_grad_buffer.  allocate != 0 creates the buffer when it does not exist yet
 * (moments zero-filled, EMA as a copy of the weights); otherwise *dev_ptr is NULL for a buffer that was never created.  The
 * step counter lives on the host (uu_train_step's `step`, uu_adamw_step's `t`). */
int uu_optimizer_state(uu_model* m, int which, int allocate, float** dev_ptr, int64_t* n_floats);
int uu_get_grad(uu_model* m, const char* group, int index, float* host, int64_t capacity);
/* Per-sample stochastic-depth factors (mask / keep_prob) drawn by the last step for `branch` 0 (attention residual) or 1
 * (MLP residual) of block `block` in stage 0 (spatial, B * n_tok samples), 1 (temporal) or 2 (strided, B samples): the
 * reference calls its DropPath layer once per branch with independent draws (vision_transformer.py:185-190). */
int uu_get_droppath_scale(uu_model* m, int stage, int block, int branch, float* host, int64_t capacity, float* keep_prob);
int uu_adamw_step(uu_model* m, float lr_t, float wd_t, float beta1, float beta2, float epsilon, int64_t t,
                  float ema_decay, void* stream);
int uu_get_ema_weight(uu_model* m, const char* group, int index, float* host, int64_t capacity);

/* ---- data-parallel training: one process (or thread) per GPU, one model per process ------------------------
 * uu_comm_unique_id : rank 0 obtains a 128-byte NCCL id and sends it to the other ranks by any means.
 * uu_comm_init      : every rank creates its communicator (collective call); world == 1 is a no-op.  NCCL is resolved at
 *                     run time (the libnccl.so.2 already loaded into the process, else the system one).
 * uu_allreduce_gradients : sum all-reduce of the whole flat gradient buffer, ordered after the work on `stream`.
 * uu_train_step     : forward + backward of the B local windows, gradient all-reduce issued in three buckets on a side
 *                     stream as the backward pass finishes them (strided blocks + heads, temporal blocks, the rest) so the
 *                     exchange overlaps the remaining backward work, all-reduced loss in loss_dev (device float), then the
 *                     fused AdamW / EMA update with the host-evaluated lr_t / wd_t (Adam t = step + 1).  Without a
 *                     communicator it is a plain local step. */
int uu_comm_unique_id(void* id_out_128_bytes, int capacity);
int uu_comm_init(uu_model* m, const void* unique_id_128_bytes, int rank, int world);
int uu_comm_destroy(uu_model* m);
int uu_comm_world_size(const uu_model* m);
int uu_allreduce_gradients(uu_model* m, void* stream);
int uu_train_step(uu_model* m, const float* x2d, const uint8_t* mask, const float* gt3d, int B, int64_t step, float lr_t,
                  float wd_t, float beta1, float beta2, float epsilon, float ema_decay, float* loss_dev, void* stream);

/* Stride-mask rule (host, integer, bit-exact): mask[n] = ((n - n_tok/2)*s_out + shift) floor-mod s_in == 0.
 * shift = centre frame index (eval, global alignment) or rand_shift*s_out (training). */
int uu_stride_mask(int n_tok, int s_out, int s_in, int64_t shift, uint8_t* mask_out);

/* ---- single-kernel entry points (device pointers) used by the parity tests ------------------- */
int uu_op_build_gather(const uint8_t* mask, int B, int n_tok, int32_t* scratch /*B+1*/, int32_t* list, int32_t* count,
                       void* stream);
int uu_op_token_fill(const uint8_t* mask, int rows, int n_tok, int d, const float* token, const float* pe, float* x,
                     void* stream);
int uu_op_layernorm(float* x, int rows, int d, const float* gamma, const float* beta, float eps, const float* table,
                    int period, void* y, int y_bf16, void* stream);
int uu_op_attention(const void* qkv, int is_bf16, int B, int S, int heads, int dh, const uint8_t* keep_mask,
                    int mask_stride, void* out, void* stream);
/* The tcgen05 / TMEM attention kernel of the bf16 schedule (vit:99-130): q | k | v rows bf16 (B * S, 1152) fed by TMA,
 * scores and probabilities in tensor memory, P . V with P read from tensor memory; 8 heads of dimension 48, S <= 80. */
int uu_op_attention_tc5(const void* qkv, int B, int S, const uint8_t* keep_mask, int mask_stride, void* out, void* stream);
/* Attention of the training step (vit:99-130 and its gradient) on mma.sync TF32 with fp32 rows as the tape holds them:
 * qkv (B * S, 3 * heads * dh), keep_mask as uu_op_attention.  dO == NULL: forward into out (B * S, heads * dh);
 * otherwise backward: dqkv (B * S, 3 * heads * dh) from dO (B * S, heads * dh).  nsplit 2 = bf16 hi + lo operand planes on
 * m16n8k16 (16 mantissa bits per operand; what uu_train_step uses), 3 = error-compensated TF32 (hi / lo operand split,
 * fp32-grade), 1 = plain TF32.  S <= 80, dh in {32, 48, 64}. */
int uu_op_attention_train(const float* qkv, const float* dO, int B, int S, int heads, int dh, const uint8_t* keep_mask,
                          int mask_stride, float* out, float* dqkv, int nsplit, void* stream);
/* C = act(A @ W + bias) (+ res); flags: 1 = ReLU, 2 = residual.  fp32 CUDA-core GEMM, W is (K, N). */
/* K2 alone: the fused spatial transformer (S1-S3 + spatial_norm, net:313-330) of a loaded model on the frames the
 * mask keeps (mask NULL: all B*n_tok frames).  precision fp32 -> out is float, bf16 -> out is bf16; out holds
 * [n_valid, 17*32] compact rows in gather-list order; n_valid_out (host) receives the row count. */
int uu_op_spatial(uu_model* m, const float* x2d, const uint8_t* mask, int B, void* out, int32_t* n_valid_out,
                  void* stream);
int uu_op_gemm_f32(const float* A, int64_t lda, const float* W, int M, int N, int K, const float* bias, int flags,
                   const float* res, int64_t ldr, float* C, int64_t ldc, void* stream);
/* tcgen05 GEMM: A bf16 (M, K) row-major, Wt bf16 (N_pad, K) = W^T with N_pad % 64 == 0. */
int uu_op_gemm_bf16(const void* A, int64_t lda, int M, int K, const void* Wt, int N_pad, int N, const float* bias,
                    int flags, const float* res, int64_t ldr, void* C, int c_bf16, int64_t ldc, void* stream);

/* tcgen05 kind::tf32 GEMM of the training step (uu_train_set_math 1): fp32 A (M, K) and Bt (N, K) = B^T, both row-major
 * with pitches lda / ldb (multiples of 4), N % 64 == 0; C fp32 = A . B (+ bias) (ReLU: flags & 1) (+ res: flags & 2). */
int uu_op_gemm_tf32(const float* A, int64_t lda, int M, int K, const float* Bt, int64_t ldb, int N, const float* bias,
                    int flags, const float* res, int64_t ldr, float* C, int64_t ldc, void* stream);
/* tcgen05 kind::tf32 weight gradient of the training step: dW (Kd, Nd) (+)= X^T dY with X (R, Kd) and dY (R, Nd) fp32
 * row-major (pitches ldx / ldy): the contraction runs over the row index, both operands are read MN-major through TMA,
 * split-K partials are summed in a fixed order (deterministic).  R >= 256, Kd % 32 == 0 (>= 128), Nd % 64 == 0.
 * Synchronous (temporary scratch). */
int uu_op_wgrad_tf32(const float* X, int64_t ldx, const float* dY, int64_t ldy, int64_t R, int Kd, int Nd, float* dW,
                     int accumulate, void* stream);
/* The folded epilogues of the bf16 schedule in isolation ("LayerNorm and residual folding"); they
 * replace LayerNormalization + Dense (vit:168-171, :183-195) and Dense + residual add.  Synchronous (temporaries).
 *   out[rows, N] (bf16) = act( LN(x; gamma, beta, eps) . W + bias ),  x bf16 [rows, d], W fp32 (d, N) on the device */
int uu_op_ln_gemm_bf16(const void* x, int rows, int d, const float* gamma, const float* beta, float eps, const float* W,
                       const float* bias, int N, int relu, void* out, void* stream);
/* The fused MLP of a temporal block (vit:190-195): x (bf16 [rows, 384], in place) += fc2(ReLU(fc1(LN(x)))) in one tcgen05
 * kernel; W1 (384, h) and W2 (h, 384) fp32 on the device, h % 64 == 0.  Synchronous (temporaries). */
int uu_op_mlp_bf16(void* x, int rows, int d, int h, const float* gamma, const float* beta, float eps, const float* W1,
                   const float* b1, const float* W2, const float* b2, float* stats_out, void* stream);
/*   resid != 0: x[rows, d] (bf16, in place) += A . W + bias          A bf16 [rows, K], W fp32 (K, d)
 *   resid == 0: x = A . W + bias + table[row % period]               table fp32 [period, d]
 *   stats_out (optional) [rows, d / 64, 2] fp32 = (sum, sum of squares) of the result per 64-column slot */
int uu_op_resid_gemm_bf16(const void* A, int rows, int K, const float* W, const float* bias, int d, void* x, int resid,
                          const float* table, int period, float* stats_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UU3D_H */
