/* rba_b200 — C ABI of the B200-native RbA hot path (Mask2Former forward + Rejected-by-All score).
 *
 * Drop-in boundary.  Plain pointers and sizes only; no torch types.
 * Every entry point:
 *   - takes BORROWED device pointers (the caller allocates inputs, outputs and keeps them alive),
 *   - launches on the given cudaStream_t (passed as void*; NULL = legacy default stream),
 *   - never synchronises the device, never frees caller memory,
 *   - returns 0 on success, non-zero on error; rba_last_error() gives the message
 *     (thread-local).  There is NO CPU fallback: without a CUDA device every compute entry fails.
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   rba_msda_forward          <- MultiScaleDeformableAttention.ms_deform_attn_forward
 *   rba_msda_backward         <- MultiScaleDeformableAttention.ms_deform_attn_backward
 *                                (mask2former/modeling/pixel_decoder/ops/src/vision.cpp:18-21,
 *                                 src/ms_deform_attn.h:25-45, src/cuda/ms_deform_attn_cuda.cu:25-84)
 *   rba_score_fused           <- F.interpolate x4 + MaskFormer.semantic_inference + get_RbA
 *                                (mask2former/maskformer_model.py:294-299,381-386; evaluate_ood.py:143-150)
 *   rba_model_* / rba_forward <- MaskFormer.forward eval branch behind META_ARCH_REGISTRY "MaskFormer"
 *                                (mask2former/maskformer_model.py:23,227-356), D2SwinTransformer
 *                                (modeling/backbone/swin.py:686), MSDeformAttnPixelDecoder
 *                                (modeling/pixel_decoder/msdeformattn.py:173), MultiScaleMaskedTransformerDecoder
 *                                (modeling/transformer_decoder/mask2former_transformer_decoder.py:232)
 *   rba_outlier_loss          <- SetCriterion.outlier_loss (mask2former/modeling/criterion.py:435-553), fwd + bwd
 *   rba_k_*                   per-kernel entry points (same kernels the engine launches), exported so the
 *                                parity tests can pin every stage against the oracle.
 *
 * GEMM operand format ("split planes"): a logical fp32 matrix X[rows, ld] feeding a GEMM is stored as two
 * bf16 planes hi = bf16(X), lo = bf16(X - hi), each [rows, ld] row-major.  hi+lo carries ~17 mantissa
 * bits; the tensor-core GEMM evaluates hi*hi + hi*lo + lo*hi with fp32 accumulation (bf16x3).
 */
#ifndef RBA_B200_H_
#define RBA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBA_OK 0
#define RBA_ERR_INVALID 1
#define RBA_ERR_CUDA 2
#define RBA_ERR_STATE 3

typedef struct rba_model rba_model; /* opaque engine handle */

/* ---- library ---- */
const char* rba_last_error(void);
int rba_version(void);
/* Number of kernels THIS library has launched in the calling process (all streams). */
int64_t rba_launch_count(void);

/* ---- model configuration (mirrors the yacs keys the reference reads) ---- */
typedef struct rba_config {
  int32_t embed_dim;        /* MODEL.SWIN.EMBED_DIM                    (mask2former/config.py:74-90) */
  int32_t depths[4];        /* MODEL.SWIN.DEPTHS */
  int32_t num_heads[4];     /* MODEL.SWIN.NUM_HEADS */
  int32_t window_size;      /* MODEL.SWIN.WINDOW_SIZE */
  int32_t conv_dim;         /* MODEL.SEM_SEG_HEAD.CONVS_DIM */
  int32_t mask_dim;         /* MODEL.SEM_SEG_HEAD.MASK_DIM */
  int32_t num_classes;      /* MODEL.SEM_SEG_HEAD.NUM_CLASSES */
  int32_t num_queries;      /* MODEL.MASK_FORMER.NUM_OBJECT_QUERIES */
  int32_t nheads;           /* MODEL.MASK_FORMER.NHEADS */
  int32_t dim_feedforward;  /* MODEL.MASK_FORMER.DIM_FEEDFORWARD */
  int32_t dec_layers;       /* MODEL.MASK_FORMER.DEC_LAYERS - 1 (mask2former_transformer_decoder.py:387-388) */
  int32_t enc_layers;       /* MODEL.SEM_SEG_HEAD.TRANSFORMER_ENC_LAYERS */
  int32_t enc_points;       /* MSDeformAttn n_points (4) */
  int32_t enc_ffn;          /* 1024, hard-coded at msdeformattn.py:315 */
  int32_t num_enc_levels;   /* len(DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES): 1 (res5) or 3 (res3..res5) */
  int32_t size_divisibility;/* MODEL.MASK_FORMER.SIZE_DIVISIBILITY */
  float pixel_mean[3];      /* MODEL.PIXEL_MEAN */
  float pixel_std[3];       /* MODEL.PIXEL_STD */
  int32_t backbone_type;    /* MODEL.BACKBONE.NAME: 0 = D2SwinTransformer, 1 = build_resnet_backbone (detectron2 bottleneck ResNet,
                               RESNETS.STRIDE_IN_1X1 false; the SWIN fields are then ignored) */
  int32_t resnet_depth;     /* MODEL.RESNETS.DEPTH: 50 or 101 */
} rba_config;

/* ---- engine ---- */
int rba_model_create(const rba_config* cfg, int device, rba_model** out);
void rba_model_destroy(rba_model* m);
/* Copies one tensor of the reference-layout state_dict (key = the reference's own key name, e.g.
 * "backbone.layers.0.blocks.0.attn.qkv.weight").  `data` is a HOST pointer to contiguous fp32. */
int rba_model_load_tensor(rba_model* m, const char* key, const float* data, const int64_t* shape, int ndim);
/* Checks that every tensor the configured architecture needs is present and builds derived device
 * buffers (bf16 split planes of the weights, re-laid-out conv filters, fused projection matrices). */
int rba_model_finalize(rba_model* m);
/* Options: "taps" (0/1: keep stage-boundary tensors of the next forwards for rba_model_get_tap),
 * "gemm_backend" (RBA_GEMM_FFMA / RBA_GEMM_TC, see below), "attn_backend" (2 = tcgen05 + TMA window attention
 * [default], 1 = mma.sync tensor-core kernel, 0 = fp32 CUDA-core kernel). */
int rba_model_set_option(rba_model* m, const char* name, int value);
/* Max batch / padded image size the workspace is sized for; (re)allocates device workspace. */
int rba_model_reserve(rba_model* m, int batch, int height, int width);

/* The workspace arena is re-allocated when a larger (batch, height, width) is seen.  The generation counter increments
 * on every re-allocation: a caller that captured rba_forward into a CUDA graph must re-capture when it changes (the
 * new forward runs in the new arena).  An arena that a capture ran in is NOT freed when outgrown (the old graph stays
 * replayable); rba_model_release_retired() frees such arenas after the caller has dropped the graphs (it
 * synchronises the device). */
int64_t rba_model_arena_generation(rba_model* m);
int rba_model_release_retired(rba_model* m);
/* Parity aid for the boolean attention masks of forward_prediction_heads (mask2former_transformer_decoder.py:483-486,
 * :433).  For prediction head `head` (0 = the head on the raw query features, i = after decoder layer i; only heads
 * < dec_layers feed a cross-attention): `dump` (device, B*Q*S_l uint8, or NULL) receives this engine's decisions
 * (1 = blocked, after the all-blocked-row reset) during the next forwards; `force` (device, same layout, or NULL) is
 * copied over them before the cross-attention reads them.  Lets a test run on the reference's own decisions and so
 * separate arithmetic parity from near-threshold decision flips.  Pointers are borrowed. */
int rba_model_debug_attn_mask(rba_model* m, int head, const uint8_t* force, uint8_t* dump);

#define RBA_IMG_U8 0   /* uint8 CHW, as produced by the reference's datasets (ToTensorV2) */
#define RBA_IMG_F32 1  /* float32 CHW */
/* MaskFormer.forward (eval) + get_RbA for a batch of B equally sized images (B,3,H,W), DEVICE pointers.
 * Outputs (device, each may be NULL): rba (B,H,W); sem_seg (B,K,H,W); pred_logits (B,Q,K+1);
 * pred_masks (B,Q,Hp/4,Wp/4) with Hp,Wp = H,W rounded up to size_divisibility. */
int rba_forward(rba_model* m, const void* images, int img_dtype, int B, int H, int W, float* rba,
                float* sem_seg, float* pred_logits, float* pred_masks, void* stream);
/* Same forward with the full output set.  ood_pred (B,2,H,W): the DenseHybrid head's logits resized to the image size
 * (`model(..., return_ood_pred=True)`, maskformer_model.py:303-305,350-351); needs the head's weights
 * (sem_seg_head.predictor.ood_pred.*, present when MODEL.MASK_FORMER.DENSE_HYBRID_LOSS is set) in the state_dict. */
typedef struct rba_outputs {
  float* rba;         /* (B,H,W) score selected by the "score_func" option */
  float* sem_seg;     /* (B,K or K+1,H,W) */
  float* pred_logits; /* (B,Q,K+1) */
  float* pred_masks;  /* (B,Q,Hp/4,Wp/4) */
  float* ood_pred;    /* (B,2,H,W) */
} rba_outputs;
int rba_forward_ex(rba_model* m, const void* images, int img_dtype, int B, int H, int W, const rba_outputs* out,
                   void* stream);
/* Copies a named stage-boundary tensor of the LAST forward into `dst` (device fp32, `capacity` floats):
 * "res2".."res5" (B,H_l*W_l,C_l token-major), "enc_out", "fpn_res4".. , "mask_feat_in", "dec_out".
 * Returns the element count through *count. Test/debug aid. */
int rba_model_get_tap(rba_model* m, const char* name, float* dst, int64_t capacity, int64_t* count, void* stream);

/* ---- fused score kernel ----
 * pred_masks (B,Q,h,w) low-res mask logits, pred_logits (B,Q,K+1).  Computes, for the padded frame
 * (4h,4w) cropped to (H,W):  sem_seg[b,c,y,x] = sum_q softmax(pred_logits[b,q,:])[c] * sigmoid(up4(pred_masks)[b,q,y,x])
 * (c < K) and rba[b,y,x] = -sum_c tanh(sem_seg[b,c,y,x]).  sem_seg may be NULL (not materialised). */
int rba_score_fused(const float* pred_masks, const float* pred_logits, int B, int Q, int K, int h, int w, int H,
                    int W, float* rba, float* sem_seg, void* stream);

/* ---- fused mask einsum + score (SURVEY §8d "Variant A": the kernel the engine runs for rba-only calls) ----
 * Same result as  pred_masks = einsum("bqc,bchw->bqhw", mask_embed, features) + bias[:, :, None, None]
 * (mask2former_transformer_decoder.py:479) followed by rba_score_fused, without materialising pred_masks.
 * mask_embed (B,Q,D) and features (B,h,w,D) [NHWC] are bf16 split planes (see above); bias (B,Q) fp32 or NULL;
 * pred_logits (B,Q,K+1).  Limits: Q <= 104, K + 1 <= 24, D a multiple of 64.
 * score_func selects the per-pixel reduction written to `score` (evaluate_ood.py:143-159):
 *   RBA_SCORE_RBA    -sum_c tanh(sem_seg[c])      (get_RbA)
 *   RBA_SCORE_ENERGY -logsumexp_c(sem_seg[c])     (get_energy, --score_func pebal)
 * include_void != 0 keeps the void column of the class softmax (semantic_inference_with_void,
 * maskformer_model.py:388-392): sem_seg then has K+1 planes and the score runs over K+1 classes.
 * sem_seg (B, K or K+1, H, W) may be NULL. */
#define RBA_SCORE_RBA 0
#define RBA_SCORE_ENERGY 1
/* engine option only ("score_func"): -logsumexp_c(sem_seg[c]) + log(softmax(ood_pred)[1] + 1e-9)
 * (get_densehybrid_score, evaluate_ood.py:161-173); the fused kernel runs RBA_SCORE_ENERGY and the head is added. */
#define RBA_SCORE_DENSEHYBRID 2
int rba_einsum_score_fused(const uint16_t* embed_hi, const uint16_t* embed_lo, const float* bias, const uint16_t* feat_hi,
                           const uint16_t* feat_lo, const float* pred_logits, int B, int Q, int K, int D, int h, int w,
                           int H, int W, int score_func, int include_void, float* score, float* sem_seg, void* stream);
/* Test / profiling hook.  RbA-only launches (sem_seg NULL, RBA_SCORE_RBA) run on the third-generation kernel
 * (csrc/score_fused3.cu: a thread owns a run of four output pixels, the class contraction stays in mma.sync registers);
 * variant 2 selects the second generation (score phase on tcgen05, csrc/score_fused2.cu), variant 1 the first (mma.sync
 * cells, csrc/score_fused.cu, which also serves every sem_seg / energy launch), 0 restores the default.  Process-wide; env
 * RBA_FS_VARIANT sets the initial value. */
int rba_k_set_fused_score_variant(int variant);

/* ---- input pipeline, device side: JPEG bitstreams -> planar RGB uint8 (n,3,H,W) on the device ----
 * Replaces the host-side decode + ToTensorV2 of the reference's dataset classes (support.py:73-81, datasets/*.py) for JPEG
 * inputs; the planes are what rba_model_forward reads (RBA_IMG_U8).  Decoder: nvJPEG, bound at run time (rba_jpeg_available()
 * is 0 and the calls fail with RBA_ERR_CUDA when libnvjpeg is absent).  data[i] are HOST pointers; stream-ordered. */
int rba_jpeg_available(void);
int rba_jpeg_info(const uint8_t* data, int64_t nbytes, int* height, int* width, int* channels);
int rba_jpeg_decode(const uint8_t* const* data, const int64_t* nbytes, int n, uint8_t* out, int H, int W, void* stream);

/* ---- streaming OoD metrics (replaces OODEvaluator.evaluate_ood / calculate_auroc, support.py:247-303, and the
 * per-image host round trip of compute_anomaly_scores, support.py:353-399) ----
 * A two-class histogram of order-preserving float keys (2 x 2^24 uint64 counters = rba_ood_hist_bytes() device bytes,
 * zero-initialised by the caller) accumulates (score, label) pixels; finalize sweeps it.  Labels: 1 = OoD (positive),
 * 0 = in-distribution, anything else ignored.  Result = sklearn roc_curve/auc/average_precision_score on scores
 * quantised to 2^-15 relative resolution.  out (device): auroc, aupr, fpr95, n_ood, n_ind (doubles). */
int64_t rba_ood_hist_bytes(void);
int64_t rba_ood_workspace_bytes(void);
int rba_ood_hist_update(const float* score, const void* label, int label_dtype_bytes, int64_t n, void* hist, void* stream);
int rba_ood_hist_finalize(const void* hist, void* workspace, double* out, void* stream);

/* ---- MSDeformAttn forward, same argument meaning as the reference FFI ----
 * value (B,S,M,D) fp32 (device); spatial_shapes (L,2) int64 (H_l,W_l) and level_start_index (L) int64 are HOST
 * arrays (shape metadata: the reference reads them on the host too, ms_deform_attn_cuda.cu:44-52 sizes);
 * sampling_loc (B,Lq,M,L,P,2) fp32 in [0,1] (x,y); attn_weight (B,Lq,M,L,P) fp32; out (B,Lq,M*D) fp32.
 * im2col_step is accepted for signature parity and validated like the reference
 * (batch % min(batch, im2col_step) == 0, ms_deform_attn_cuda.cu:55-57) but does not chunk the launch. */
int rba_msda_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                     const float* sampling_loc, const float* attn_weight, int B, int S, int M, int D, int Lq, int L,
                     int P, int im2col_step, float* out, void* stream);

/* Same op with spatial_shapes / level_start_index as DEVICE int64 arrays -- exactly the tensors the reference FFI
 * receives and its kernel reads on the device (ms_deform_im2col_cuda.cuh:242-250): no host copy, no synchronisation,
 * CUDA-graph capturable.  (The host-array variant above additionally validates sum(H*W) == S.) */
int rba_msda_forward_dev(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                         const float* sampling_loc, const float* attn_weight, int B, int S, int M, int D, int Lq, int L,
                         int P, int im2col_step, float* out, void* stream);

/* ---- MSDeformAttn backward: `ms_deform_attn_backward` of the same FFI (ops/src/vision.cpp:20,
 * ops/src/ms_deform_attn.h:44-66, ops/src/cuda/ms_deform_attn_cuda.cu:87-153) ----
 * grad_output (B,Lq,M*D) fp32.  Outputs (device): grad_value (B,S,M,D) -- zeroed inside, then accumulated with atomics --
 * grad_sampling_loc (B,Lq,M,L,P,2) and grad_attn_weight (B,Lq,M,L,P), both fully written.  Same validation as the
 * forward (batch % min(batch, im2col_step) == 0, ms_deform_attn_cuda.cu:117-119). */
int rba_msda_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                      const float* sampling_loc, const float* attn_weight, const float* grad_output, int B, int S, int M,
                      int D, int Lq, int L, int P, int im2col_step, float* grad_value, float* grad_sampling_loc,
                      float* grad_attn_weight, void* stream);

int rba_msda_backward_dev(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                          const float* sampling_loc, const float* attn_weight, const float* grad_output, int B, int S, int M,
                          int D, int Lq, int L, int P, int im2col_step, float* grad_value, float* grad_sampling_loc,
                          float* grad_attn_weight, void* stream);

/* ---- training-side RbA outlier loss, forward + backward in one call (SetCriterion.outlier_loss,
 * mask2former/modeling/criterion.py:435-553; OUTLIER_LOSS_FUNC squared_hinge) ----
 * pred_masks (B,Q,h,w), pred_logits (B,Q,K+1) fp32; outlier_masks (B,H,W) uint8 or int64 (1 = outlier, 0 = inlier,
 * anything else ignored); score_mode: 0..2 = OUTLIER_LOSS_TARGET "nls" with SCORE_NORM none / tanh / sigmoid, 3 = "energy".
 * loss: 1 float (device).  d_pred_masks (B,Q,h,w) and d_pred_logits (B,Q,K+1) receive d loss / d input (both or neither).
 * workspace: rba_outlier_loss_workspace_floats(...) floats, 16-byte aligned.  K <= 32. */
#define RBA_OL_NLS_NONE 0
#define RBA_OL_NLS_TANH 1
#define RBA_OL_NLS_SIGMOID 2
#define RBA_OL_ENERGY 3
int64_t rba_outlier_loss_workspace_floats(int B, int Q, int K, int h, int w);
int rba_outlier_loss(const float* pred_masks, const float* pred_logits, const void* outlier_masks, int label_dtype_bytes, int B,
                     int Q, int K, int h, int w, int H, int W, int score_mode, float inlier_upper_threshold,
                     float outlier_lower_threshold, float* loss, float* d_pred_masks, float* d_pred_logits, float* workspace,
                     void* stream);

/* ---- per-kernel entry points (device pointers) ---- */
/* ood_pred logits (B,h,w,2) -> bilinear align_corners=True resize to (H,W): ood_pred (B,2,H,W) and/or
 * score[b,y,x] += log(softmax(.)[1] + 1e-9) (maskformer_model.py:303-305, evaluate_ood.py:167-171). */
int rba_k_ood_pred_resize(const float* logits, int B, int h, int w, int H, int W, float* ood_pred, float* score,
                          void* stream);
/* fp32 [rows,cols] (row pitch ld floats) -> bf16 planes hi, lo [rows,cols] (pitch ldp elements). */
int rba_k_split(const float* x, int64_t rows, int cols, int64_t ld, uint16_t* hi, uint16_t* lo, int64_t ldp,
                void* stream);

#define RBA_ACT_NONE 0
#define RBA_ACT_RELU 1
#define RBA_ACT_GELU 2 /* exact erf GELU (nn.GELU(), swin.py:25) */

#define RBA_GEMM_FFMA 0 /* fp32 FMA pipe, operands reconstructed as hi+lo */
#define RBA_GEMM_TC 1   /* tcgen05 bf16x3 (hi*hi + hi*lo + lo*hi), fp32 accumulate in TMEM */

typedef struct rba_gemm_args {
  /* C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual).  A and W are split planes, K-contiguous. */
  const uint16_t* a_hi; const uint16_t* a_lo; int64_t lda;  /* [M,K] */
  const uint16_t* w_hi; const uint16_t* w_lo; int64_t ldw;  /* [N,K] */
  int32_t M, N, K;
  int32_t batch;                 /* >= 1; blockIdx.z */
  int64_t a_bstride, w_bstride;  /* elements between batches (0 = shared) */
  const float* bias;             /* NULL, [N] (per column) or [M] (per row, bias_per_row=1) */
  int32_t bias_per_row; int64_t bias_bstride;
  int32_t act;                   /* RBA_ACT_* applied after bias */
  const float* residual;         /* NULL or fp32 with the OUTPUT's indexing/pitch; added after act */
  float* c; int64_t ldc; int64_t c_bstride;  /* fp32 output or NULL */
  uint16_t* c_hi; uint16_t* c_lo; int64_t ldcp; int64_t cp_bstride; /* split output or NULL */
  /* Optional Swin window-reverse row map (swin.py:277-287): output row r of the windowed matrix goes to
   * token row map(r) of `c`/`residual`, rows that fall in the window padding are dropped. */
  int32_t swin_map; int32_t sw_H, sw_W, sw_ws, sw_shift;
  int32_t backend;               /* RBA_GEMM_* */
  /* Optional (tensor-core backend, split-plane output only): write the planes in the (window, part, head) tiled layout the
   * tcgen05 window attention reads with ONE bulk copy per operand tile -- the tcgen05 "core matrix" order
   *   [row / 144][col / C][(col % C) / 32][(col % 32) / 8][row % 144][col % 8]
   * (N = 3C: q | k | v blocks of `qkv_tile_heads` heads of 32 channels).  0 = plain row-major planes. */
  int32_t qkv_tile_heads;
} rba_gemm_args;
int rba_k_gemm(const rba_gemm_args* args, void* stream);

/* 3x3 stride-1 pad-1 convolution, NHWC planes in, fp32 NHWC out; filters as split planes [Cout][9*Cin]
 * with k = (ky*3+kx)*Cin + ci  (msdeformattn.py:281-290 output_conv, bias-free). */
int rba_k_conv3x3(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* w_hi, const uint16_t* w_lo, int B, int H,
                  int W, int Cin, int Cout, float* y, int backend, void* stream);

/* LayerNorm over the last dim with three row-gather modes; writes fp32 and/or split planes (either may be NULL).
 *  mode 0: plain rows.                      x [rows,C]
 *  mode 1: Swin window gather (swin.py:247-271): out row r (windowed, shifted, padded frame) = LN(x[token(r)]) or 0.
 *  mode 2: PatchMerging gather (swin.py:327-334): out [B*(H/2)*(W/2), 4C] = LN_4C(cat(x0,x1,x2,x3)). */
int rba_k_layernorm(const float* x, const float* gamma, const float* beta, int mode, int B, int H, int W, int C,
                    int ws, int shift, float eps, float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream);

/* Window attention core (swin.py:145-168): qkv fp32 [B*nW*ws*ws, 3C] -> out planes [rows, C].
 * bias_table [(2ws-1)^2, heads]; shift>0 adds the -100 region mask of swin.py:413-440 (computed analytically). */
int rba_k_window_attn(const float* qkv, const float* bias_table, int B, int H, int W, int C, int heads, int ws, int shift,
                      uint16_t* out_hi, uint16_t* out_lo, void* stream);

/* Tensor-core variant (mma.sync m16n8k16, bf16x3): qkv given as split planes [rows, 3C] (what the QKV GEMM writes). */
int rba_k_window_attn_planes(const uint16_t* qkv_hi, const uint16_t* qkv_lo, const float* bias_table, int B, int H, int W,
                             int C, int heads, int ws, int shift, uint16_t* out_hi, uint16_t* out_lo, void* stream);

/* tcgen05 + TMA variant (the engine's default): S = q k^T and O = P v as tcgen05.mma (bf16x3) with fp32 accumulators in
 * tensor memory, q / k / v tiles fetched by TMA straight from the planes, P handed to the second MMA through tensor memory.
 * `bias_prepared` is the per-head, log2(e)-scaled copy of relative_position_bias_table that
 * rba_k_window_attn_prepare_bias writes (rba_k_window_attn_bias_floats(heads) floats, once per block at load time).
 * shift must be 0 or ws / 2. */
int64_t rba_k_window_attn_bias_floats(int heads);
int rba_k_window_attn_prepare_bias(const float* bias_table, int heads, float* prepared, void* stream);
int rba_k_window_attn_tc(const uint16_t* qkv_hi, const uint16_t* qkv_lo, const float* bias_prepared, int B, int H, int W, int C,
                         int heads, int ws, int shift, uint16_t* out_hi, uint16_t* out_lo, void* stream);
/* Same kernel with q / k / v in the tiled layout of rba_gemm_args.qkv_tile_heads (what the engine runs: every operand tile
 * is 9216 contiguous bytes and arrives by one cp.async.bulk instead of 144 64-byte TMA box rows). */
int rba_k_window_attn_tc_tiled(const uint16_t* qkv_hi, const uint16_t* qkv_lo, const float* bias_prepared, int B, int H, int W,
                               int C, int heads, int ws, int shift, uint16_t* out_hi, uint16_t* out_lo, void* stream);

/* nn.MultiheadAttention core for the decoder (mask2former_transformer_decoder.py:52-53,110-113):
 * q [B,Lq,E], k,v [B,Lk,E] fp32 (already projected, q NOT yet scaled), mask (B,Lq,Lk) uint8 (1 = blocked, shared by
 * all heads) or NULL; out planes [B*Lq, E].
 * workspace: rba_k_mha_workspace_floats(...) floats (partials of the key splits). */
int64_t rba_k_mha_workspace_floats(int B, int Lq, int Lk, int heads);
int rba_k_mha(const float* q, const float* k, const float* v, const uint8_t* mask, int B, int Lq, int Lk, int E, int heads,
              uint16_t* out_hi, uint16_t* out_lo, float* workspace, void* stream);

/* GroupNorm(32) over token-major x [B, HW, C] (two deterministic passes) fused with what follows it in the pixel
 * decoder: y = GN(x) (+ bilinear_up(prev [B,hp,wp,C]) if prev) (relu if relu); writes fp32 and/or planes. */
int rba_k_groupnorm(const float* x, const float* gamma, const float* beta, int B, int H, int W, int C, int groups,
                    float eps, const float* prev, int hp, int wp, int relu, float* y, uint16_t* y_hi, uint16_t* y_lo,
                    double* workspace /* >= B*groups*2*chunks doubles, see rba_k_groupnorm_ws */, void* stream);
int64_t rba_k_groupnorm_ws(int B, int H, int W, int C, int groups);

/* Patch embedding (maskformer_model.py:255-257 + swin.py:479-495): normalise, zero-pad to (Hp,Wp), 4x4/4 conv, LN. */
int rba_k_patch_embed(const void* images, int img_dtype, int B, int H, int W, int Hp, int Wp, const float* mean,
                      const float* stdv, const float* conv_w /* [C,3,4,4] */, const float* conv_b, const float* gamma,
                      const float* beta, int C, float* tokens /* [B,(Hp/4)*(Wp/4),C] */, void* stream);

/* ---- ResNet backbone pieces (detectron2 build_resnet_backbone; the 1x1 / 3x3 bottleneck convolutions run on rba_k_gemm /
 * rba_k_conv3x3 with the batch norm folded into weights + bias) ----
 * stem: normalise + zero-pad to (Hp,Wp) (maskformer_model.py:255-257), 7x7 / stride 2 / pad 3 convolution with folded BN
 * (w [64][3*7*7], bias [64]) and ReLU -> out (B,Hp/2,Wp/2,64) NHWC fp32. */
int rba_k_stem_conv(const void* images, int img_dtype, int B, int H, int W, int Hp, int Wp, const float* mean,
                    const float* stdv, const float* w, const float* bias, float* out, void* stream);
/* 3x3 / stride 2 / pad 1 max-pool, NHWC fp32 in -> (B,H/2,W/2,C) fp32 and / or split planes (either may be NULL). */
int rba_k_maxpool3x3s2(const float* x, int B, int H, int W, int C, float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream);
/* y[b,oy,ox,:] = [relu](x[b,oy*stride,ox*stride,:] + bias): epilogue of the 3x3 convolutions (stride-2 ones are computed
 * at stride 1 and sub-sampled here), the ReLU after the residual add, the stride-2 gather of the projection shortcuts.
 * bias may be NULL; y (fp32) may alias x when stride == 1. */
int rba_k_bias_act_sub(const float* x, const float* bias, int B, int H, int W, int C, int stride, int relu, float* y,
                       uint16_t* y_hi, uint16_t* y_lo, void* stream);

/* Attention-mask of forward_prediction_heads (mask2former_transformer_decoder.py:483-486): bilinear resize of
 * mask logits (B,Q,h,w) to (th,tw), sigmoid < 0.5 -> 1, then rows that are entirely 1 are reset to 0 (:433).
 * out (B,Q,th*tw) uint8. */
int rba_k_attn_mask(const float* masks, int B, int Q, int h, int w, int th, int tw, uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RBA_B200_H_ */
