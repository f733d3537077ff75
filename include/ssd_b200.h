/*
 * ssd_b200.h -- C ABI of libssd_b200.so, the B200 (sm_100a) implementation of
 * the tf-ssd hot path.
 *
 * The reference (FurkanOM/tf-ssd) is pure Python/TensorFlow and has NO native
 * interface of its own; the boundary it exposes is a set of
 * Python call signatures.  Every entry point below therefore cites the
 * reference Python function (file:line under the reference tree) whose
 * arithmetic it replaces.  The Python modules under tf_ssd_b200/ keep the
 * reference's names/argument meanings and bind these symbols through ctypes
 * (tf_ssd_b200/_ffi.py); INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary.
 *   - every pointer named d_* is a DEVICE pointer owned by the caller
 *     (allocated e.g. as a torch tensor); h_* is a HOST pointer read
 *     synchronously before the call returns.  The library never frees or
 *     retains caller memory.
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed
 *     as void*); calls are re-entrant and safe to capture into a CUDA graph
 *     (no allocation, no synchronisation, no host<->device copy inside).
 *   - scratch memory is a caller-provided workspace; query its size with the
 *     matching ssd_*_workspace_bytes().
 *   - return value: 0 success; <0 argument error (SSD_ERR_*); >0 cudaError_t.
 *     ssd_last_error() returns a thread-local description of the last
 *     non-zero return.
 *   - boxes are normalised [y1, x1, y2, x2] float32; tensors are dense,
 *     row-major, in the shapes written next to each argument.
 */
#ifndef SSD_B200_H_
#define SSD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSD_B200_ABI_VERSION 1

#define SSD_OK               0
#define SSD_ERR_NULL        -1   /* a required pointer is NULL            */
#define SSD_ERR_SHAPE       -2   /* a size/shape argument is out of range  */
#define SSD_ERR_WORKSPACE   -3   /* workspace_bytes smaller than required  */
#define SSD_ERR_UNSUPPORTED -4   /* valid request this build cannot serve  */

#define SSD_MAX_FEATURE_MAPS 8
#define SSD_MAX_ASPECT_RATIOS 8

typedef void* ssd_stream_t;      /* cudaStream_t */

int         ssd_abi_version(void);
const char* ssd_last_error(void);
/* Device the library sees as current: fills name (<= name_len bytes), SM count
 * and compute capability major*10+minor.  Used by the loader to fail loudly on
 * a non-sm_100 device. */
int ssd_device_info(char* h_name, int name_len, int* h_sm_count, int* h_cc);

/* ---------------------------------------------------------------- priors -- */
/* utils/bbox_utils.py:131-214 generate_prior_boxes (+ :151-176 base boxes,
 * :131-148 scales).  h_aspect_ratios is the concatenation of the per-map
 * aspect-ratio lists (h_ar_counts[i] entries for map i); each map gets
 * h_ar_counts[i]+1 anchors per cell (the extra square box last).
 * d_out: [n_anchors,4].  n_anchors must equal sum(fm^2 * (count+1)). */
int ssd_prior_boxes(const int* h_fm_shapes, int n_maps,
                    const float* h_aspect_ratios, const int* h_ar_counts,
                    float* d_out, int n_anchors, ssd_stream_t stream);
int ssd_prior_box_count(const int* h_fm_shapes, int n_maps, const int* h_ar_counts);

/* ------------------------------------------------------------------- IoU -- */
/* utils/bbox_utils.py:24-55 generate_iou_map.
 * d_boxes: [N,4] (boxes_batched=0) or [B,N,4] (boxes_batched=1);
 * d_gt: [B,G,4]; d_out: [B,N,G].  The rank-2 mode of the reference
 * (transpose_perm=[1,0]) is B=1.  0/0 yields NaN like the reference. */
int ssd_iou_map(const float* d_boxes, const float* d_gt, int B, int N, int G,
                int boxes_batched, float* d_out, ssd_stream_t stream);

/* ------------------------------------------------------- target encoding -- */
/* utils/train_utils.py:102-136 calculate_actual_outputs, fused with
 * utils/bbox_utils.py:85-128 get_deltas_from_bboxes; the [B,N,G] IoU map is
 * never materialised.
 * d_priors [N,4]; d_gt_boxes [B,G,4]; d_gt_labels [B,G] int32 (-1 padding);
 * h_variances[4].  Outputs: d_deltas [B,N,4]; d_onehot [B,N,L] (may be NULL);
 * d_label [B,N] int32 and d_match [B,N] int32 = argmax GT index (either may be
 * NULL).  Positive iff max IoU > iou_threshold (strict); first max wins. */
int ssd_match_encode(const float* d_priors, const float* d_gt_boxes, const int32_t* d_gt_labels,
                     int B, int N, int G, int L, float iou_threshold, const float* h_variances,
                     float* d_deltas, float* d_onehot, int32_t* d_label, int32_t* d_match,
                     ssd_stream_t stream);

/* utils/bbox_utils.py:85-128 / :58-82 as stand-alone element-wise ops.
 * priors_batched: 0 -> d_priors [N,4] broadcast over B, 1 -> [B,N,4]. */
int ssd_encode_deltas(const float* d_priors, const float* d_boxes, int B, int N, int priors_batched,
                      float* d_deltas, ssd_stream_t stream);
int ssd_decode_boxes(const float* d_priors, const float* d_deltas, int B, int N, int priors_batched,
                     float* d_boxes, ssd_stream_t stream);

/* ------------------------------------------------------------------ loss -- */
/* ssd_loss.py:26-57 CustomLoss.loc_loss_fn + :59-91 CustomLoss.conf_loss_fn
 * (Huber on positives; categorical CE with hard-negative mining by rank).
 * d_actual_deltas/d_pred_deltas [B,N,4]; d_actual_labels [B,N,L] one-hot;
 * d_pred_labels [B,N,L]: probabilities (from_logits=0, the public signature;
 * Keras renormalise + clip 1e-7 path) or pre-softmax logits (from_logits=1,
 * what Keras substitutes inside model.fit).
 * Outputs: d_loc_loss [B], d_conf_loss [B] (per-image, like the reference).
 * Either half may be skipped by passing NULL for its inputs AND output.
 * The workspace keeps per-anchor state for ssd_loss_bwd. */
size_t ssd_loss_workspace_bytes(int B, int N, int L);
int ssd_loss_fwd(const float* d_actual_deltas, const float* d_pred_deltas,
                 const float* d_actual_labels, const float* d_pred_labels,
                 int B, int N, int L, float neg_pos_ratio, float loc_loss_alpha, int from_logits,
                 float* d_loc_loss, float* d_conf_loss,
                 void* d_workspace, size_t workspace_bytes, ssd_stream_t stream);
/* Gradient of  grad_scale * (sum_b loc[b] + sum_b conf[b])  w.r.t. pred_deltas
 * and the LOGITS (from_logits=1 forward required).  Keras' batch-mean
 * reduction (trainer.py:91-94) is grad_scale = 1/B.  Must follow ssd_loss_fwd
 * on the same workspace. */
int ssd_loss_bwd(const float* d_actual_deltas, const float* d_pred_deltas,
                 const float* d_actual_labels, const float* d_pred_logits,
                 int B, int N, int L, float loc_loss_alpha, float grad_scale,
                 float* d_grad_deltas, float* d_grad_logits,
                 const void* d_workspace, size_t workspace_bytes, ssd_stream_t stream);

/* --------------------------------------------------------- decode + NMS -- */
/* models/header.py:88 Activation("softmax") over the last axis: [rows,L]. */
int ssd_softmax(const float* d_logits, int64_t rows, int L, float* d_probs, ssd_stream_t stream);

/* models/decoder.py:60-93 SSDDecoder.call, fused: deltas*variances (:74),
 * get_bboxes_from_deltas (:75, utils/bbox_utils.py:58-82), background-row
 * suppression (:78-83) and tf.image.combined_non_max_suppression (:86-92 via
 * utils/bbox_utils.py:10-21) with max_output_size_per_class = max_total_size,
 * clip_boxes, pad_per_class=False.
 * d_priors [N,4]; d_pred_deltas [B,N,4]; d_pred_labels [B,N,L] probabilities
 * (from_logits=0) or logits (from_logits=1: softmax fused in).
 * Outputs (order of the reference's return, decoder.py:93):
 *   d_boxes [B,T,4] clipped to [0,1], zero padded; d_labels [B,T] float32
 *   class ids; d_scores [B,T]; d_valid [B] int32 number of detections, or -1
 *   if that image produced more candidates than max_candidates (overflow).
 * max_candidates: per-image capacity of the candidate list; 0 -> N (exact for
 * normalised probabilities with score_threshold >= 0.5). */
size_t ssd_decode_nms_workspace_bytes(int B, int N, int L, int max_total_size, int max_candidates);
int ssd_decode_nms(const float* d_priors, const float* d_pred_deltas, const float* d_pred_labels,
                   int B, int N, int L, const float* h_variances, int from_logits,
                   float score_threshold, float iou_threshold, int max_total_size, int max_candidates,
                   float* d_boxes, float* d_labels, float* d_scores, int32_t* d_valid,
                   void* d_workspace, size_t workspace_bytes, ssd_stream_t stream);

/* utils/bbox_utils.py:10-21 non_max_suppression -> the general
 * tf.image.combined_non_max_suppression (pad_per_class=False).
 * d_boxes [B,N,q,4] with q == 1 or q == L; d_scores [B,N,L].
 * Outputs as TensorFlow returns them: boxes [B,T,4], scores [B,T],
 * classes [B,T] float32, valid [B] int32 (-1 on candidate overflow).
 * max_candidates: 0 -> N*L (always sufficient). */
size_t ssd_combined_nms_workspace_bytes(int B, int N, int L, int max_output_size_per_class,
                                        int max_total_size, int max_candidates);
int ssd_combined_nms(const float* d_boxes, const float* d_scores, int B, int N, int q, int L,
                     int max_output_size_per_class, int max_total_size,
                     float iou_threshold, float score_threshold, int clip_boxes, int max_candidates,
                     float* d_out_boxes, float* d_out_scores, float* d_out_classes, int32_t* d_valid,
                     void* d_workspace, size_t workspace_bytes, ssd_stream_t stream);

/* ===================================================== network forward ==== */
/* The SSD forward pass (models/ssd_mobilenet_v2.py:15-47, models/ssd_vgg16.py:
 * 66-121, models/header.py:54-90) as NHWC fp16 tensors with fp32 accumulation.
 * The reference builds these layers from Keras (Conv2D / DepthwiseConv2D /
 * BatchNormalization / ReLU / MaxPool2D / L2Normalization); each entry below
 * replaces the Keras layer call sites cited.  BatchNormalization is folded
 * into the preceding kernel+bias by the host code (inference). */

#define SSD_ACT_NONE  0
#define SSD_ACT_RELU  1
#define SSD_ACT_RELU6 2

/* One convolution (Keras Conv2D call sites: models/ssd_vgg16.py:80-113,
 * models/ssd_mobilenet_v2.py:31-41, models/header.py:71-85, and the 1x1/3x3
 * convolutions inside keras_applications MobileNetV2 at ssd_mobilenet_v2.py:25).
 *   in      [B,H,W,Cin]   fp16 NHWC, Cin % 8 == 0
 *   weight  [Cout,KH,KW,Cin] fp16 ("OHWI": K-major for the tensor cores)
 *   bias    [Cout] fp32 (may be NULL)
 *   residual: optional fp16 tensor added AFTER bias+activation, addressed like
 *             output segment 0 (MobileNetV2 block_i_add).
 * Output pixel (b,oy,ox) = row m = (b*Ho+oy)*Wo+ox.  Channels [0,split) go to
 * out0, [split,Cout) to out1 (split == Cout: single output).  Element address
 * of segment s: out_s + b*img_stride_s + (oy*Wo+ox)*pix_stride_s + (c - begin_s)
 * in elements -- this is how the head convolutions write straight into the
 * concatenated [B,N_anchors,L] / [B,N_anchors,4] tensors (header.py:46-51).
 * out_f32: 0 -> fp16 outputs, 1 -> fp32 outputs. */
typedef struct ssd_conv_desc {
    const void*  in;
    const void*  weight;
    const float* bias;
    const void*  residual;
    void*        out0;
    void*        out1;
    int32_t B, H, W, Cin;
    int32_t Ho, Wo, Cout;
    int32_t KH, KW, stride, dilation, pad_top, pad_left;
    int32_t act;
    int32_t out_f32;
    int32_t split;
    int32_t reserved;
    int64_t img_stride0, pix_stride0;
    int64_t img_stride1, pix_stride1;
} ssd_conv_desc;

int ssd_conv2d(const ssd_conv_desc* h_desc, ssd_stream_t stream);

/* A chain of small convolutions as ONE launch: the tail of the SSD graphs -- models/ssd_mobilenet_v2.py:33-41
 * (extra2_1 ... extra4_2), models/ssd_vgg16.py:108-113 (conv9_1 ... conv11_2) -- plus the multibox heads
 * (models/header.py:68-85) of the feature maps that tail produces.  h_descs[i] are ordinary ssd_conv_desc; a layer may
 * read the out0 of an EARLIER layer of the chain.  h_phase[i] (non-decreasing) groups the layers: all inputs of a
 * layer must come from outside the chain or from a layer of a lower phase (e.g. phase k: extra layer k and the head of
 * the map produced in phase k-1).  A cluster of 8 CTAs owns a few images through the whole chain; every CTA computes a
 * share of each layer's output channels; a cluster barrier separates the phases (csrc/conv_chain.cu).
 * Constraints: KH == KW in {1,3}, dilation 1, stride 1|2, Cin % 128 == 0, no residual, one batch size, small maps
 * (ssd_conv_chain_supported tells; SSD_ERR_UNSUPPORTED otherwise). */
int ssd_conv_chain(const ssd_conv_desc* h_descs, const int32_t* h_phase, int n_layers, ssd_stream_t stream);
int ssd_conv_chain_supported(const ssd_conv_desc* h_descs, const int32_t* h_phase, int n_layers);
/* 0: unsupported; otherwise the number of waves of clusters the launch needs on the current device (1: all resident). */
int ssd_conv_chain_waves(const ssd_conv_desc* h_descs, const int32_t* h_phase, int n_layers);

/* Keras DepthwiseConv2D 3x3 (+ folded BN + ReLU6) inside MobileNetV2
 * (ssd_mobilenet_v2.py:25).  in [B,H,W,C] fp16, weight [3,3,C] fp16,
 * bias [C] fp32, out [B,Ho,Wo,C] fp16; C % 8 == 0. */
int ssd_depthwise3x3(const void* d_in, const void* d_weight, const float* d_bias, void* d_out,
                     int B, int H, int W, int C, int Ho, int Wo, int stride, int pad_top, int pad_left,
                     int act, ssd_stream_t stream);

/* The tail of a MobileNetV2 inverted-residual block (keras_applications mobilenet_v2._inverted_res_block under
 * models/ssd_mobilenet_v2.py:25) as ONE launch: DepthwiseConv2D 3x3 (+ folded BN + dw_act) -> 1x1 Conv2D (+ folded
 * BN + act, + residual).  The depthwise output is produced in shared memory as the tensor-core operand and never
 * written to global memory.  in [B,H,W,C] fp16, dw_weight [3,3,C] fp16, dw_bias [C] fp32 (may be NULL),
 * proj_weight [Cout,C] fp16, proj_bias [Cout] fp32 (may be NULL), residual / out [B,Ho,Wo,Cout] fp16.
 * C % 8 == 0, Cout % 8 == 0, Cout <= 256, stride 1 or 2; returns SSD_ERR_UNSUPPORTED otherwise. */
typedef struct ssd_dwproj_desc {
    const void* in; const void* dw_weight; const float* dw_bias;
    const void* proj_weight; const float* proj_bias; const void* residual; void* out;
    int32_t B, H, W, C, Ho, Wo, Cout;
    int32_t stride, pad_top, pad_left, dw_act, act;
    int32_t reserved;
} ssd_dwproj_desc;
int ssd_dwproj(const ssd_dwproj_desc* h_desc, ssd_stream_t stream);
/* 1 when ssd_dwproj can run this configuration (shape limits, alignment, a tile geometry that fits shared memory). */
int ssd_dwproj_supported(const ssd_dwproj_desc* h_desc);

/* A WHOLE MobileNetV2 inverted-residual block (keras_applications mobilenet_v2._inverted_res_block under
 * models/ssd_mobilenet_v2.py:25) as ONE launch: 1x1 expand Conv2D (+ folded BN + exp_act) -> DepthwiseConv2D 3x3
 * (+ folded BN + dw_act) -> 1x1 project Conv2D (+ folded BN + act, + residual).  Neither the expanded activation nor
 * the depthwise output is written to global memory: per output tile the input patch (tile + halo) is loaded once, the
 * expansion is computed on the tensor cores into TMEM 64 channels at a time, converted to a shared-memory patch, the
 * depthwise taps build the projection's operand tile in shared memory.
 * in [B,H,W,Cin] fp16, exp_weight [Cexp,Cin] fp16, dw_weight [3,3,Cexp] fp16, proj_weight [Cout,Cexp] fp16, biases
 * fp32 (may be NULL), residual / out [B,Ho,Wo,Cout] fp16.  Channels % 8 == 0, Cin <= 256, Cexp <= 1024, Cout <= 256,
 * stride 1 or 2; returns SSD_ERR_UNSUPPORTED otherwise (ssd_irblock_supported tells beforehand). */
typedef struct ssd_irblock_desc {
    const void* in; const void* exp_weight; const float* exp_bias;
    const void* dw_weight; const float* dw_bias;
    const void* proj_weight; const float* proj_bias; const void* residual; void* out;
    int32_t B, H, W, Cin, Cexp, Ho, Wo, Cout;
    int32_t stride, pad_top, pad_left, exp_act, dw_act, act;
    int32_t reserved0, reserved1;
} ssd_irblock_desc;
int ssd_irblock(const ssd_irblock_desc* h_desc, ssd_stream_t stream);
int ssd_irblock_supported(const ssd_irblock_desc* h_desc);
/* Programmatic dependent launch between consecutive kernels of a plan: -1 the default policy (on, unless SSD_B200_PDL=0 or a
 * CUDA injection library -- Nsight Compute, compute-sanitizer -- is attached), 0 off, 1 on.  Takes effect for launches (and
 * graph captures) made after the call; the training driver switches it off while it records its step (the many mid-size
 * kernels of a training step lose ~3 % when their successors' CTAs are scheduled early). */
int ssd_set_pdl(int mode);
/* ssd_irblock has two implementations: the tcgen05 / TMEM pipeline (any supported shape) and an mma.sync kernel with one
 * CTA per small output tile for MobileNetV2's large-map blocks 1-6 ((Cin, Cexp, Cout, stride) = (16,96,24,2), (24,144,24,1),
 * (24,144,32,2), (32,192,32,1), (32,192,64,2) with ReLU6 / ReLU6 / linear activations), plus a channel-grouped mma.sync
 * kernel for the small-map blocks 7-12, 14, 15 ((64,384,64,1), (64,384,96,1), (96,576,96,1), (160,960,160,1)).  Test hook:
 * -1 automatic (mma.sync for blocks 1-6, tcgen05 otherwise), 0 tcgen05 only, 1 every mma.sync variant that has an
 * instantiation, 2 the large-map variant only. */
int ssd_debug_irblock_mode(int mode);
/* Debug aid, not a reference interface: d_buf = device buffer of 5 x 512 uint64 that CTA 0 of later ssd_irblock
 * launches fills with per-role (globaltimer << 8 | tag) stamps (tools/trace_irblock.py); NULL switches it off. */
int ssd_irblock_trace(void* d_buf);
/* Debug aid: the tile geometry the planner picks for a block, without launching (h_out16 = bw, bh, bb, pw, ph, P, halves,
 * n_tiles, n_e, kc_in, wexp_stages, n_patch, out_bufs, quad, dw_R, smem_bytes). */
int ssd_irblock_plan(const ssd_irblock_desc* h_desc, int32_t* h_out16);
/* The same aid for ssd_dwproj (roles: 0 TMA, 1 MMA, 2 depthwise, 3 epilogue) and the per-image pass of ssd_decode_nms /
 * ssd_combined_nms (role 0: phase boundaries of image 0): d_buf = device buffer of 8 x 512 uint64
 * (tools/trace_kernel.py); NULL switches it off. */
int ssd_debug_trace(void* d_buf);
/* Test hook for ssd_conv2d's CTA-pair (cta_group::2, M = 256) kernel: -1 automatic (layers with at least one tile per
 * SM, full 128 / 256-wide N tiles, K >= 512), 0 never, 1 whenever the shape allows (small test shapes). */
int ssd_debug_pair_mode(int mode);

/* MobileNetV2's first three layers as ONE launch (models/ssd_mobilenet_v2.py:25 -> keras_applications Conv1_pad / Conv1 /
 * bn_Conv1 / Conv1_relu, expanded_conv_depthwise (+BN, ReLU6), expanded_conv_project (+BN)): the 3x3 stride-2 stem
 * straight from the NHWC image (float32 in [0,1], or the uint8 batch of utils/data_utils.py:33-37 with
 * convert_image_dtype fused in), the stride-1 SAME depthwise 3x3 of block 0 and its 1x1 projection.  Neither the stem
 * output [B,Hs,Ws,32] nor the depthwise output is written to global memory (one CTA per 30 x 10 output tile: staged image
 * patch -> mma.sync stem -> packed-half2 depthwise -> mma.sync projection -> coalesced stores).
 * stem_weight [32,3,3,3] fp16, dw_weight [3,3,32] fp16, proj_weight [Cout,32] fp16, biases fp32 (may be NULL),
 * out [B,Hs,Ws,Cout] fp16 with Hs, Ws = the stem's output size.  Cmid == 32, Cout in {8,16,24,32}; returns
 * SSD_ERR_UNSUPPORTED otherwise (ssd_stem_dwproj_supported tells beforehand). */
typedef struct ssd_stem_dwproj_desc {
    const void* image; const void* stem_weight; const float* stem_bias;
    const void* dw_weight; const float* dw_bias; const void* proj_weight; const float* proj_bias; void* out;
    int32_t image_u8;                 /* 0: float32 image, 1: uint8 image */
    int32_t B, H, W, Hs, Ws, Cmid, Cout;
    int32_t pad_top, pad_left, stem_act, dw_act, act;
    int32_t reserved;
} ssd_stem_dwproj_desc;
int ssd_stem_dwproj(const ssd_stem_dwproj_desc* h_desc, ssd_stream_t stream);
int ssd_stem_dwproj_supported(const ssd_stem_dwproj_desc* h_desc);

/* MobileNetV2 stem: keras_applications Conv1_pad + Conv1 (3x3, stride 2, Cin = 3) + bn_Conv1 +
 * Conv1_relu (models/ssd_mobilenet_v2.py:25), computed straight from the fp32 NHWC image
 * [B,H,W,3] (the fp32->fp16 input rounding of the pipeline is fused in).
 * weight [Cout,3,3,3] fp16 (BN folded), bias [Cout] fp32, out [B,Ho,Wo,Cout] fp16; Cout == 32. */
int ssd_stem_conv3x3s2(const float* d_img, const void* d_weight, const float* d_bias, void* d_out,
                       int B, int H, int W, int Cout, int Ho, int Wo, int pad_top, int pad_left, int act,
                       ssd_stream_t stream);
/* The same layer fed with the uint8 NHWC batch [B,H,W,3] the reference's input pipeline holds BEFORE
 * tf.image.convert_image_dtype (utils/data_utils.py:33-37): float32(u8) * float32(1/255) is fused into the load, so
 * the result is bit-identical to ssd_stem_conv3x3s2 on the converted float32 image, at a quarter of the bytes. */
int ssd_stem_conv3x3s2_u8(const void* d_img_u8, const void* d_weight, const float* d_bias, void* d_out,
                          int B, int H, int W, int Cout, int Ho, int Wo, int pad_top, int pad_left, int act,
                          ssd_stream_t stream);

/* The same first-layer kernel with the stride as a parameter: stride 2 / Cout 32 (MobileNetV2 Conv1, as above) and
 * stride 1 / Cout 64 -- VGG16's conv1_1 (models/ssd_vgg16.py:80: Conv2D(64, 3x3, SAME, ReLU) on the image).  K = 27 is
 * too shallow for the tensor-map path (nine 64-channel k-blocks for 3 real channels); here the layer is bound by
 * writing its output. */
int ssd_stem_conv3x3(const float* d_img, const void* d_weight, const float* d_bias, void* d_out,
                     int B, int H, int W, int Cout, int Ho, int Wo, int stride, int pad_top, int pad_left, int act,
                     ssd_stream_t stream);
int ssd_stem_conv3x3_u8(const void* d_img_u8, const void* d_weight, const float* d_bias, void* d_out,
                        int B, int H, int W, int Cout, int Ho, int Wo, int stride, int pad_top, int pad_left, int act,
                        ssd_stream_t stream);
/* Training-plan form: the image as fp16 NHWC with 8 channels (ssd_image_to_f16c8's output, which the first layer's
 * filter gradient reads) and the weights in the zero-padded OHWI [Cout,3,3,8] layout of ssd_conv2d / ssd_conv2d_wgrad,
 * so the forward of conv1_1 / Conv1 shares its variables with the backward kernels. */
int ssd_stem_conv3x3_f16c8(const void* d_img_f16c8, const void* d_weight_ohwi8, const float* d_bias, void* d_out,
                           int B, int H, int W, int Cout, int Ho, int Wo, int stride, int pad_top, int pad_left, int act,
                           ssd_stream_t stream);

/* Device-side input pipeline.  utils/data_utils.py:33-37: tf.image.convert_image_dtype(uint8 ->
 * float32) + tf.image.resize(img, (out_h, out_w)) (bilinear, half-pixel centres), optionally followed by
 * augmentation.py:119-139 flip_left_right, in one pass: d_img_u8 [H,W,3] uint8 -> d_out [out_h,out_w,3] float32
 * (one slot of the NHWC batch).  ssd_flip_boxes mirrors n boxes [y1,x1,y2,x2] in place (all-zero padding stays). */
int ssd_preprocess_image(const void* d_img_u8, int H, int W, float* d_out, int out_h, int out_w, int flip,
                         ssd_stream_t stream);
int ssd_flip_boxes(float* d_boxes, int n, ssd_stream_t stream);

/* Training-time augmentation of a whole batch.  Replaces augmentation.py:16-33 `apply` and the
 * operations it chains per example -- patch (:205-234: optional expand_image :164-202 on a canvas filled with the
 * image's per-channel mean, crop window, tf.image.resize back to H x W), flip_horizontally (:119-139),
 * random_brightness / _contrast / _hue / _saturation (:67-116), tf.clip_by_value(img, 0, 1) -- and the box updates
 * (utils/bbox_utils.py:217-233 renormalize_bboxes_with_min_max, the flip).  Every random decision is an input:
 * d_plans holds B plans of 16 32-bit words
 *   [0] flags: bit0 patch, bit1 expand, bit2 flip, bit3 brightness, bit4 contrast, bit5 hue, bit6 saturation,
 *       bit7 skip the final clip (the reference's single operations do not clip, only `apply` does, :32)
 *   [1..4] pad_top, pad_left, canvas_h, canvas_w (int32; H, W and zero pads without expand)
 *   [5..8] crop y0, x0, height, width on the canvas (int32; tf.image.sample_distorted_bounding_box's window)
 *   [9..12] brightness delta, contrast factor, hue delta, saturation factor (float32)   [13..15] reserved
 * d_images [B,H,W,3] float32 (already resized, utils/data_utils.py:36-37) -> d_out [B,out_h,out_w,3] (must not
 * alias).  `patch` resizes the window to out_h x out_w = H x W (:231); out = canvas size with the window = whole
 * canvas gives expand_image's own output; images without a patch need out_h x out_w == H x W.
 * d_boxes [B,G,4] is updated in place (all-zero padding rows stay zero; may be NULL when G == 0). */
size_t ssd_augment_workspace_bytes(int B, int H, int W, int out_h, int out_w);
int ssd_augment_batch(const float* d_images, float* d_out, float* d_boxes, int B, int H, int W, int out_h, int out_w, int G,
                      const void* d_plans, void* d_workspace, size_t workspace_bytes, ssd_stream_t stream);

/* fp32 NHWC image [B,H,W,3] (utils/data_utils.py:36 convert_image_dtype output)
 * -> fp16 NHWC with the channel dimension zero-padded to 8. */
int ssd_image_to_f16c8(const float* d_img, void* d_out, int64_t n_pixels, ssd_stream_t stream);
/* uint8 NHWC image -> the same fp16 c8 layout, convert_image_dtype (utils/data_utils.py:36) fused in. */
int ssd_image_u8_to_f16c8(const void* d_img_u8, void* d_out, int64_t n_pixels, ssd_stream_t stream);

/* Keras MaxPool2D(padding="same") (models/ssd_vgg16.py:82-101): window k,
 * stride s, TensorFlow SAME padding (odd pixel after, -inf padded). fp16 NHWC. */
int ssd_maxpool(const void* d_in, void* d_out, int B, int H, int W, int C, int Ho, int Wo,
                int k, int stride, int pad_top, int pad_left, ssd_stream_t stream);

/* L2Normalization.call (models/ssd_vgg16.py:63): x * rsqrt(max(sum_c x^2, 1e-12))
 * * scale[c].  in/out [rows,C] fp16, scale [C] fp32. */
int ssd_l2norm(const void* d_in, const float* d_scale, void* d_out, int64_t rows, int C, ssd_stream_t stream);

/* ============================================================ training ==== */
/* The reference trains through Keras: model.compile(Adam(1e-3), loss=[loc, conf]) and model.fit
 * (trainer.py:86-127) -- forward, loss, backward and the optimizer step are all inside TensorFlow.
 * The entries below are the pieces of that step that are not already covered by ssd_conv2d /
 * ssd_loss_bwd.  Gradients w.r.t. activations are fp16 NHWC tensors (loss-scaled by the caller),
 * gradients w.r.t. variables are fp32 and ACCUMULATE into their buffers (zero them per step).
 *
 * Data gradient of a convolution: for stride 1 it is ssd_conv2d itself applied to dY with the
 * filter produced by ssd_filter_flip_transpose and pad' = (k-1)*dilation - pad; for stride s the
 * caller first spreads dY with ssd_upsample_zero. */

/* Filter gradient of the Keras Conv2D described by *h_desc (same geometry fields as the forward
 * call; out0/out1/weight/bias are ignored):  dW[co][ky][kx][ci] += sum dY[b,oy,ox,co] * X[b,iy,ix,ci].
 * d_dy [B*Ho*Wo, ldy] fp16 (ldy >= Cout, multiple of 8); d_dw [Cout,KH,KW,Cin] fp32. */
int ssd_conv2d_wgrad(const ssd_conv_desc* h_desc, const void* d_dy, int ldy, float* d_dw, ssd_stream_t stream);

/* d_dy[i] = y[i] > 0 ? d_dy[i] : 0  (ReLU / the lower knee of ReLU6), fp16, n % 8 == 0. */
int ssd_relu_bwd(void* d_dy, const void* d_y, int64_t n, ssd_stream_t stream);
/* d_db[c] += sum over rows of d_dy[row][c], c < C; d_dy [rows, ld] fp16. */
int ssd_bias_grad(const void* d_dy, float* d_db, int64_t rows, int ld, int C, ssd_stream_t stream);
/* d_wt[ci][KH-1-ky][KW-1-kx][co] = d_w[co][ky][kx][ci], co padded with zeros to ldo (fp16). */
int ssd_filter_flip_transpose(const void* d_w, void* d_wt, int Cout, int KH, int KW, int Cin, int ldo,
                              ssd_stream_t stream);
/* d_out[b][oy*s][ox*s][:] = d_in[b][oy][ox][:]; every other element of d_out must already be zero. */
int ssd_upsample_zero(const void* d_in, void* d_out, int B, int Ho, int Wo, int C, int Hu, int Wu, int s,
                      ssd_stream_t stream);
/* Keras MaxPool2D backward (models/ssd_vgg16.py:82-101): the first maximum of each window (scan
 * order) receives the window's gradient.  accumulate != 0 adds to d_dx. */
int ssd_maxpool_bwd(const void* d_x, const void* d_y, const void* d_dy, void* d_dx, int B, int H, int W, int C,
                    int Ho, int Wo, int k, int stride, int pad_top, int pad_left, int accumulate, ssd_stream_t stream);
/* L2Normalization backward (models/ssd_vgg16.py:63): d_dx [rows,C] fp16, d_dscale[C] += ... (fp32). */
int ssd_l2norm_bwd(const void* d_x, const float* d_scale, const void* d_dy, void* d_dx, float* d_dscale,
                   int64_t rows, int C, int accumulate, ssd_stream_t stream);
/* Gradient of one head convolution's output from the concatenated loss gradients (inverse of the
 * head's two-segment store, models/header.py:46-51): d_dy [B, HW, ld] fp16 = [g_logits | g_deltas | 0]. */
int ssd_head_grad_gather(const float* d_g_logits, const float* d_g_deltas, void* d_dy, int B, int N, int L,
                         int anchor_offset, int HW, int A, int ld, ssd_stream_t stream);
/* Keras Adam (trainer.py:92; beta1 0.9, beta2 0.999, epsilon 1e-7), fused with the loss-scale removal
 * (inv_scale), the l2 kernel regulariser gradient (l2 = 2*5e-4 for ssd_vgg16.py:76 kernels, else 0), the
 * fp16 working-copy refresh (d_w16 may be NULL) and sum(w^2) of the pre-update weights (d_sumsq may be NULL).
 * lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t). */
int ssd_adam_step(float* d_w, float* d_m, float* d_v, const float* d_grad, void* d_w16, int64_t n, float lr_t,
                  float beta1, float beta2, float eps, float inv_scale, float l2, float* d_sumsq, ssd_stream_t stream);

/* Multi-tensor form of ssd_adam_step: ONE launch for all variables.  d_vars is a DEVICE array of n_vars
 * descriptors; sum(w^2) is accumulated only for descriptors with l2 != 0; max_n = the largest n. */
typedef struct ssd_adam_var {
    float* w; float* m; float* v; const float* grad; void* w16;   /* w16 may be NULL */
    int64_t n;
    float l2;
    float reserved;
} ssd_adam_var;
int ssd_adam_step_multi(const ssd_adam_var* d_vars, int n_vars, int64_t max_n, float lr_t, float beta1, float beta2,
                        float eps, float inv_scale, float* d_sumsq, ssd_stream_t stream);
/* Mixed-precision guard (no counterpart in the reference, which trains in fp32 -- trainer.py:91-94): activation gradients
 * are fp16 and loss-scaled, so one overflow would feed Inf / NaN into every variable.  ssd_grad_nonfinite_multi sets
 * d_flag[0] = 1 when any gradient of the table is non-finite (0 otherwise); the guarded Adam skips the whole update when
 * d_guard[0] != 0 and counts the skipped step in d_guard[1] (d_guard = int[2]; NULL = plain ssd_adam_step_multi). */
int ssd_grad_nonfinite_multi(const ssd_adam_var* d_vars, int n_vars, int64_t max_n, int* d_flag, ssd_stream_t stream);
int ssd_adam_step_multi_guarded(const ssd_adam_var* d_vars, int n_vars, int64_t max_n, float lr_t, float beta1, float beta2,
                                float eps, float inv_scale, float* d_sumsq, int* d_guard, ssd_stream_t stream);

/* ---- MobileNetV2 training (keras_applications MobileNetV2 under models/ssd_mobilenet_v2.py:25) ----
 * keras.layers.BatchNormalization(epsilon=1e-3, momentum=0.999) in TRAINING mode over the rows of an
 * [M = B*H*W, C] fp16 NHWC activation (C % 8 == 0, C <= 2048): batch mean / biased variance per channel,
 * y = act(gamma * (x - mean) * rstd + beta) (+ d_res, the block_i_add shortcut; may be NULL).
 * d_save [2*C] fp32 receives mean | rstd for the backward pass.  d_moving_mean / d_moving_var (may be
 * NULL) are updated as m*momentum + batch*(1-momentum), the variance with Bessel's correction like the
 * fused Keras kernel.  Deterministic (fixed-order partial sums in the workspace). */
size_t ssd_bn_workspace_bytes(int C);
int ssd_bn_train_fwd(const void* d_x, const float* d_gamma, const float* d_beta, float* d_moving_mean,
                     float* d_moving_var, int64_t M, int C, float eps, float momentum, int act,
                     const void* d_res, void* d_y, float* d_save, void* d_workspace, size_t workspace_bytes,
                     ssd_stream_t stream);
/* Backward of the above: d_dy is the gradient w.r.t. y.  d_dx [M,C] fp16 is overwritten; d_dres (may be
 * NULL) receives (accumulate_res ? += : =) d_dy; d_dgamma / d_dbeta [C] fp32 ACCUMULATE. */
int ssd_bn_train_bwd(const void* d_x, const void* d_dy, const float* d_gamma, const float* d_beta,
                     const float* d_save, int64_t M, int C, int act, void* d_dx, void* d_dres,
                     int accumulate_res, float* d_dgamma, float* d_dbeta, void* d_workspace,
                     size_t workspace_bytes, ssd_stream_t stream);
/* Gradients of Keras DepthwiseConv2D 3x3 (same geometry arguments as ssd_depthwise3x3):
 * data gradient d_dx [B,H,W,C] fp16 (accumulate != 0 adds), filter gradient d_dw [3,3,C] fp32 ACCUMULATES. */
int ssd_depthwise3x3_dgrad(const void* d_dy, const void* d_weight, void* d_dx, int B, int H, int W, int C,
                           int Ho, int Wo, int stride, int pad_top, int pad_left, int accumulate,
                           ssd_stream_t stream);
int ssd_depthwise3x3_wgrad(const void* d_x, const void* d_dy, float* d_dw, int B, int H, int W, int C, int Ho,
                           int Wo, int stride, int pad_top, int pad_left, ssd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SSD_B200_H_ */
