/*
 * unit_b200 -- C ABI of the B200-native UniT RoI stage (libunit_b200.so, sm_100a).
 *
 * Every entry point replaces one interface the reference (ubc-vision/UniT) reaches through Detectron2 /
 * torchvision on its RoI path; the reference file:line that calls it is cited per function (paths relative to
 * the reference root; [D2] = Detectron2, [TV] = torchvision, both un-vendored dependencies).
 *
 * Conventions
 *   - plain pointers + sizes; no torch types.  All pointers are DEVICE pointers unless the name ends in _host.
 *   - the caller owns every buffer including workspaces; kernels never allocate, free or retain pointers.
 *   - all work is enqueued on `stream` (a cudaStream_t); no host synchronisation, no default-stream use.
 *   - return 0 on success, a negative UNIT_E* code on failure; unit_last_error() gives the thread-local text.
 *   - boxes are xyxy fp32; class / index outputs are int64 where the reference's are (PyTorch LongTensor).
 *   - variable-length outputs are written into caller-sized maximum buffers with a device-side count.
 */
#ifndef UNIT_B200_H_
#define UNIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* unit_stream_t; /* cudaStream_t */

enum {
  UNIT_OK = 0,
  UNIT_EINVAL = -1,    /* bad shape / dtype / alignment / null pointer */
  UNIT_ECUDA = -2,     /* launch or runtime failure; text has cudaGetErrorString */
  UNIT_EWORKSPACE = -3 /* workspace too small */
};

enum { UNIT_F32 = 0, UNIT_BF16 = 1 };

int unit_version(void);
const char* unit_last_error(void);
/* sha256 of the sources (csrc/, this header, compile flags) the library was built from; the Python loader refuses a
 * library whose digest differs from the tree it sits in. */
const char* unit_source_digest(void);
/* Number of kernels this library has launched in this process (bench.py's `gpu_launches`). */
unsigned long long unit_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * ROIAlign forward / backward.
 * Replaces [D2] ROIPooler -> ROIAlign(aligned=True) == [TV] torch.ops.torchvision.roi_align /
 * _roi_align_backward; reference call sites modeling/roi_heads/roi_heads.py:356,364,499,511,598,610,708,715,
 * 729,829,843,911 (`self.box_pooler(features, [x.proposal_boxes ...])`).
 *   feat      [N,C,H,W]  NCHW, dtype f32 or bf16          rois [R,5] fp32 (batch_idx, x1, y1, x2, y2)
 *   out       [R,C,PH,PW] same dtype as feat
 *   rois_sorted != 0: rois are grouped by ascending batch index (what ROIPooler produces); enables the
 *   slab-resident kernels.  workspace: unit_roi_align_workspace_bytes(N, C, H, W, R, dtype) bytes.
 * Backward writes grad_feat [N,C,H,W] completely (no pre-zeroing needed by the caller).
 */
size_t unit_roi_align_workspace_bytes(int N, int C, int H, int W, int R, int dtype);
/* [D2] poolers.convert_boxes_to_pooler_format (inside ROIPooler.forward): boxes [R,4] concatenated in image order +
 * int32 prefix offsets [n_img+1] -> rois [R,5] = (image index, x1, y1, x2, y2). */
int unit_boxes_to_rois(const float* boxes, const int* offsets, int n_img, int R, float* rois, unit_stream_t stream);
int unit_roi_align_fwd(const void* feat, const float* rois, void* out, int N, int C, int H, int W, int R, int PH,
                       int PW, float spatial_scale, int sampling_ratio, int aligned, int dtype, int rois_sorted,
                       void* workspace, size_t workspace_bytes, unit_stream_t stream);
int unit_roi_align_bwd(const void* grad_out, const float* rois, void* grad_feat, int N, int C, int H, int W, int R,
                       int PH, int PW, float spatial_scale, int sampling_ratio, int aligned, int dtype,
                       int rois_sorted, void* workspace, size_t workspace_bytes, unit_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * pairwise IoU.  Replaces [D2] structures.pairwise_iou (weak_detector_fast_rcnn.py:327-329,411 and inside
 * [D2] ROIHeads.label_and_sample_proposals reached at roi_heads.py:459,563,794,925).
 *   iou[g*P + p] = inter > 0 ? inter / ((area1 + area2) - inter) : 0, each op rounded to fp32 (no FMA).
 */
int unit_pairwise_iou(const float* boxes1, const float* boxes2, float* iou, int G, int P, unit_stream_t stream);

/* Matcher on a given [G,P] quality matrix.  Replaces modeling/matcher.py:54-119 (UniT, 3 outputs) and [D2]
 * Matcher (2 outputs: pass matched_vals = NULL).  thresholds_host: the T user thresholds (ascending, without the
 * +-inf sentinels); labels_host: T+1 labels in {-1,0,1}.  G == 0 -> matches 0, labels labels_host[0], vals 0.
 * workspace: G floats (only read when allow_low_quality_matches). */
int unit_matcher(const float* iou, int G, int P, const float* thresholds_host, const int* labels_host, int T,
                 int allow_low_quality_matches, int64_t* matches, int8_t* match_labels, float* matched_vals,
                 void* workspace, size_t workspace_bytes, unit_stream_t stream);

/* Fused pairwise_iou + Matcher for every image of a batch in one launch (no [G,P] matrix in HBM).
 * gt_offsets / prop_offsets: int32 [n_img+1] prefix sums (device).  Outputs are indexed by global proposal row;
 * matches are image-local GT indices. */
int unit_iou_match(const float* gt_boxes, const int* gt_offsets, const float* prop_boxes, const int* prop_offsets,
                   int n_img, int P_total, const float* thresholds_host, const int* labels_host, int T,
                   int64_t* matches, int8_t* match_labels, float* matched_vals, unit_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * fg/bg labelling + sampling.  Replaces [D2] ROIHeads._sample_proposals + sampling.subsample_labels
 * (roi_heads.py:415 and via label_and_sample_proposals) and weak_detector_fast_rcnn.py:308-318.
 * Step 1 (unit_label_proposals): prop_classes = gt_classes[matches]; label 0 -> num_classes; label -1 -> -1;
 *   no GT -> num_classes.  pos_idx / neg_idx: per image, ascending image-local indices of foreground / background
 *   proposals, stored from prop_offsets[i]; counts[2*i], counts[2*i+1] = #pos, #neg.
 * Step 2 (unit_sample_gather): the host draws randperm(#pos), randperm(#neg) (the reference's two draws per image)
 *   and passes them concatenated; *_sel_offsets are int32 [n_img+1] prefix sums of the TAKEN counts
 *   (num_pos = min(#pos, int(batch*frac)), num_neg = min(#neg, batch-num_pos)), perm_*_offsets the prefix sums of
 *   the FULL permutation lengths.  Output row j of image i: j < num_pos ? pos[perm_pos[j]] : neg[perm_neg[j-num_pos]].
 */
int unit_label_proposals(const int64_t* matches, const int8_t* match_labels, const int64_t* gt_classes,
                         const int* gt_offsets, const int* prop_offsets, int n_img, int P_total, int num_classes,
                         int64_t* prop_classes, int64_t* pos_idx, int64_t* neg_idx, int* counts,
                         unit_stream_t stream);
int unit_sample_gather(const int64_t* pos_idx, const int64_t* neg_idx, const int64_t* perm_pos,
                       const int* perm_pos_offsets, const int64_t* perm_neg, const int* perm_neg_offsets,
                       const int* pos_sel_offsets, const int* neg_sel_offsets, const int* prop_offsets,
                       const int* gt_offsets, int n_img, int S_total, const float* prop_boxes,
                       const int64_t* prop_classes, const int64_t* matches, const float* gt_boxes,
                       int64_t* sampled_idx, float* out_boxes, int64_t* out_classes, int64_t* out_matched,
                       float* out_gt_boxes, const float* prop_field, float* out_field, unit_stream_t stream);
/* prop_field / out_field (both may be NULL): one pass-through float field of the proposals (objectness_logits),
 * concatenated like prop_boxes, gathered by the same launch: out_field[j] = prop_field[global row of sample j]. */

/* [D2] proposal_utils.add_ground_truth_to_proposals (called at roi_heads.py:459 through label_and_sample_proposals)
 * for n_img images in one launch per 32 images: out rows of image i = its prop_counts[i] proposals, then its
 * gt_counts[i] ground-truth boxes with objectness logit gt_logit; images back to back.  prop_boxes / prop_logits /
 * gt_boxes / prop_counts / gt_counts are HOST arrays of n_img device pointers / counts. */
int unit_append_gt(const float* const* prop_boxes, const float* const* prop_logits, const float* const* gt_boxes,
                   const int* prop_counts, const int* gt_counts, int n_img, float gt_logit, float* out_boxes,
                   float* out_logits, unit_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * softmax + Box2BoxTransform.apply_deltas.  Replaces [D2] FastRCNNOutputLayers.predict_probs / predict_boxes
 * (fast_rcnn.py:456,459; meta_arch/rcnn.py:525) and weak_detector_fast_rcnn.py:270-287.
 *   probs [R,K1] (nullable), boxes [R,4*KB] (nullable).  weights = (wx,wy,ww,wh); dw,dh clamped to scale_clamp. */
int unit_softmax_decode(const float* scores, const float* deltas, const float* proposals, float* probs, float* boxes,
                        int R, int K1, int KB, float wx, float wy, float ww, float wh, float scale_clamp,
                        unit_stream_t stream);
/* Box2BoxTransform.get_deltas (fast_rcnn.py:69-71 via FastRCNNOutputs.box_reg_loss). */
int unit_box_get_deltas(const float* src, const float* tgt, float* deltas, int R, float wx, float wy, float ww,
                        float wh, unit_stream_t stream);

/* Fused Fast R-CNN loss + gradients.  Replaces [D2] FastRCNNOutputs.losses as called at fast_rcnn.py:438-445:
 * loss_cls = mean softmax cross-entropy over the R sampled RoIs; loss_box_reg = sum over foreground RoIs of
 * smooth_l1(pred_deltas[gt class] - get_deltas(proposal, gt_box)) / R.  losses [2] = (loss_cls, loss_box_reg);
 * d_scores [R,K+1], d_deltas [R,4K] = their gradients (for upstream gradient 1).  workspace: 2*R floats. */
int unit_fastrcnn_loss(const float* scores, const float* deltas, const float* proposals, const float* gt_boxes,
                       const int64_t* gt_classes, int R, int K, float wx, float wy, float ww, float wh,
                       float smooth_l1_beta, float* losses, float* d_scores, float* d_deltas, void* workspace,
                       size_t workspace_bytes, unit_stream_t stream);
/* Same, with both gradients written into ONE packed buffer d_packed[R, ld_packed] = [d_scores (K+1) | d_deltas (4K) |
 * zeros]: the layout unit_predictor_wgrad consumes (ld_packed >= 128 there).  Here losses has THREE elements:
 * (loss_cls, loss_box_reg, loss_cls + loss_box_reg) -- the total the trainer logs, without another launch. */
int unit_fastrcnn_loss_packed(const float* scores, const float* deltas, const float* proposals, const float* gt_boxes,
                              const int64_t* gt_classes, int R, int K, float wx, float wy, float ww, float wh,
                              float smooth_l1_beta, float* losses, float* d_packed, int ld_packed, void* workspace,
                              size_t workspace_bytes, unit_stream_t stream);
/* Unreduced (per-RoI) losses and their per-row gradients for FastRCNNOutputsReduction / NLL / Regression
 * (modeling/roi_heads/fast_rcnn.py:24-130, weak_detector_fast_rcnn.py:23-37): row_ce [R] (nll != 0: scores are
 * log-probabilities and row_ce = -scores[r][cls]); row_box [R,4] smooth-L1 per coordinate of the gt class's deltas
 * (0 for background rows); d_scores [R,K+1] and d_box [R,4] their derivatives.  deltas may be NULL (class loss only). */
int unit_fastrcnn_row_losses(const float* scores, const float* deltas, const float* proposals, const float* gt_boxes,
                             const int64_t* gt_classes, int R, int K, float wx, float wy, float ww, float wh,
                             float smooth_l1_beta, int nll, float* row_ce, float* row_box, float* d_scores, float* d_box,
                             unit_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * fast_rcnn_inference for a batch of images in two launches.  Replaces [D2] fast_rcnn_inference(_single_image)
 * (fast_rcnn.py:461-468, weak_detector_fast_rcnn.py:299-306, meta_arch/rcnn.py:526): drop non-finite rows, drop
 * the background column, clip to the image, score > thresh, row-major (roi, class) candidates, class-wise batched
 * NMS, top-k.
 *   boxes [R,4*KB], probs [R,K+1], roi_offsets int32 [n_img+1], image_hw fp32 [n_img,2] (h,w), all device.
 * unit_detect_filter writes per-image candidate segments starting at roi_offsets[i]*K:
 *   cand_boxes [R*K,4], cand_scores [R*K], cand_roi/cand_cls int32 [R*K], cand_counts int32 [n_img].
 *   workspace (optional, >= 8*R + 256 bytes; the NMS workspace can be passed): lets the filter spread over the GPU.
 * unit_detect_nms consumes them and writes det_* [n_img, topk(...)] + det_counts [n_img].
 *   nms_mode: 0 = class-wise on the raw boxes (torchvision _batched_nms_vanilla), 1 = coordinate trick
 *   (boxes + cls*(max+1), torchvision _batched_nms_coordinate_trick), 2 = follow torchvision's CUDA rule
 *   (coordinate trick iff 4*Nc <= 100000), 3 = torchvision's CPU rule (4*Nc <= 4000).
 */
int unit_detect_filter(const float* boxes, const float* probs, const int* roi_offsets, const float* image_hw,
                       int n_img, int R, int K, int KB, float score_thresh, float* cand_boxes, float* cand_scores,
                       int* cand_roi, int* cand_cls, int* cand_counts, void* workspace, size_t workspace_bytes,
                       unit_stream_t stream);
size_t unit_nms_workspace_bytes(int n_seg, int total_candidates);
int unit_detect_nms(const float* cand_boxes, const float* cand_scores, const int* cand_roi, const int* cand_cls,
                    const int* cand_counts, const int* roi_offsets, int n_img, int R, int K, float nms_thresh,
                    int nms_mode,
                    int topk, float* det_boxes, float* det_scores, int64_t* det_classes, int64_t* det_roi,
                    int* det_counts, void* workspace, size_t workspace_bytes, unit_stream_t stream);
/* torchvision.ops.batched_nms / nms drop-in ([D2] layers.batched_nms, imported at fast_rcnn.py:9).
 *   idxs int64 [N] (NULL -> plain nms).  keep int64 [N] sorted by descending score (ties: lower index first),
 *   keep_count int32 [1].  max_keep < 0 -> all. */
int unit_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int N, float iou_thresh,
                     int nms_mode, int max_keep, int64_t* keep, int* keep_count, void* workspace,
                     size_t workspace_bytes, unit_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Lingual similarity.  Replaces fast_rcnn.py:376-382 `get_similarity` (+ the softmax of roi_heads.py:272).
 *   emb [V,D]; indexer int64 [K]; base int64 [B]; novel int64 [Nn]; raw / soft: [Nn,B] (either nullable). */
int unit_lingual_similarity(const float* emb, const int64_t* indexer, const int64_t* base, const int64_t* novel,
                            int D, int B, int Nn, float* raw, float* soft, unit_stream_t stream);

/* Visual similarity + combination + base->novel transfer, fused per RoI.
 * Replaces roi_heads.py:245-336 `get_similarity_matrices` (Sum combination; class-level terms are pre-reduced by
 * the host into static_*) and fast_rcnn.py:403-426 / 503-528 (transfer, + weak scores, + fine-tune terms).
 *   vis_logits [R,K+1]   mean OICR logits of the box-head features (NULL when no head uses 'visual')
 *   static_cls/bbox/seg [Nn,B]  sum of the class-level terms already multiplied by their 1/len(terms) weight
 *   wv_*            weight of the visual term for that head (0 = absent);  norm_* = divide by clamp(sum,1e-9)
 *   class_kind int32 [K]: -1 neither, 0..B-1 -> base slot + 0, 1000000+n -> novel slot n
 *   delta_scores [R,K+1], proposal_deltas [R,4K] in; weak_scores [R,K+1], ft_scores, ft_deltas nullable
 *   out_scores [R,K+1], out_bbox [R,4K]; out_s_cls / out_s_bbox / out_s_seg [R,Nn,B] nullable
 *   do_transfer == 0 -> scores/bbox pass through (training of the Base predictor, fast_rcnn.py:401)
 *   novel_neg_inf != 0 -> novel logits := -inf (fast_rcnn.py:427-428)
 */
typedef struct {
  int R, K, B, Nn;
  float vis_threshold;
  float wv_cls, wv_bbox, wv_seg;
  int norm_cls, norm_bbox, norm_seg;
  int do_transfer, novel_neg_inf;
  int static_per_roi; /* bit h set: static_<head h> is [R,Nn,B] (an explicit per-RoI similarity) instead of [Nn,B] */
  /* row strides (in floats) of delta_scores / proposal_deltas / ft_scores / ft_deltas, so that column blocks of one
   * packed GEMM output can be passed without a copy; 0 = dense (K+1 resp. 4K) */
  int ld_delta_scores, ld_proposal_deltas, ld_ft_scores, ld_ft_deltas;
  /* same for vis_logits / weak_scores (0 = dense K+1) */
  int ld_vis_logits, ld_weak_scores;
} unit_transfer_params;
int unit_similarity_transfer(const unit_transfer_params* p, const float* vis_logits, const float* static_cls,
                             const float* static_bbox, const float* static_seg, const int* base, const int* novel,
                             const int* class_kind, const float* delta_scores, const float* proposal_deltas,
                             const float* weak_scores, const float* ft_scores, const float* ft_deltas,
                             float* out_scores, float* out_bbox, float* out_s_cls, float* out_s_bbox,
                             float* out_s_seg, unit_stream_t stream);
/* Gradient of the transfer w.r.t. delta_scores / proposal_deltas for fixed similarity (S is built under frozen
 * weights in every shipped fine-tune YAML).  detach_transfer != 0 reproduces the .detach() of fast_rcnn.py:566,575. */
int unit_similarity_transfer_bwd(const unit_transfer_params* p, const float* s_cls, const float* s_bbox,
                                 const int* base, const int* novel, const int* class_kind, const float* g_scores,
                                 const float* g_bbox, int detach_transfer, float* g_delta_scores,
                                 float* g_proposal_deltas, unit_stream_t stream);

/* Gradient of the same fused step w.r.t. the visual logits (the mean OICR logits the visual similarity is built from).
 * The reference's get_similarity_matrices (modeling/roi_heads/roi_heads.py:245-257, called at :618) is not under
 * no_grad, so with a trainable box head the fine-tune loss reaches box_features through softmax -> renormalise ->
 * threshold -> S -> bmm (fast_rcnn.py:503-516).  Inputs are the forward's; g_vis_logits is [R,K+1]. */
int unit_similarity_transfer_bwd_vis(const unit_transfer_params* p, const float* vis_logits, const float* static_cls,
                                     const float* static_bbox, const int* base, const int* novel,
                                     const float* delta_scores, const float* proposal_deltas, const float* g_scores,
                                     const float* g_bbox, float* g_vis_logits, unit_stream_t stream);

/* Predictor GEMM on tcgen05 tensor cores: y[M,N] = x[M,K] . w[N,K]^T + bias[N], fp32 in / fp32 out, TF32 multiply
 * with fp32 accumulation in TMEM (TMA-fed, split-K with a deterministic reduction).  Replaces the packed nn.Linear
 * calls of fast_rcnn.py:386-392,488-489 and weak_detector_fast_rcnn.py:172-175.  K % 4 == 0; bias nullable. */
size_t unit_predictor_gemm_workspace_bytes(int M, int N, int K);
int unit_predictor_gemm(const float* x, const float* w, const float* bias, float* y, int M, int N, int K,
                        void* workspace, size_t workspace_bytes, unit_stream_t stream);
/* Two problems sharing M and K in ONE launch (+ one reduce launch): y1[M,ldy1] = x1 . w1^T + b1 and y2[M,ldy2] = x2 . w2^T
 * + b2 (N2 == 0: only the first).  Columns n >= N of a padded output row are written as zeros.  This is the packed
 * [delta | bbox | ft | mean-OICR] product on the box-head features together with the mean-OICR product on the weak
 * branch's features (fast_rcnn.py:486-494, roi_heads.py:252). */
size_t unit_predictor_gemm2_workspace_bytes(int M, int N1, int N2, int K);
int unit_predictor_gemm2(const float* x1, const float* w1, const float* b1, float* y1, int N1, int ldy1,
                         const float* x2, const float* w2, const float* b2, float* y2, int N2, int ldy2, int M, int K,
                         void* workspace, size_t workspace_bytes, unit_stream_t stream);

/* Weight / bias gradients of the trainable predictor columns on tcgen05 (TF32, fp32 accumulation in TMEM):
 *   dW[N,K] = gy[R,N]^T . x[R,K],  db[N] = column sums of gy      (N <= 128; gy rows padded with zeros to ldg >= 128)
 * Both operands are read as they lie in memory (MN-major UMMA operands: no transposed copies).  Gradient rows
 * [seg_rows[s], seg_rows[s+1]) go to w_dst[s] ([rows,K] row-major) / b_dst[s], multiplied by the device scalar
 * *seg_scale[s] when given (the upstream dL/dloss), overwritten or accumulated in place -- i.e. straight into the
 * parameters' .grad views of the flat all-reduce bucket.  Replaces autograd's Linear backward for cls_score_ft /
 * bbox_pred_ft (the trainable parameters of fast_rcnn.py:477-482, 527-528). */
size_t unit_predictor_wgrad_workspace_bytes(int R, int K);
int unit_predictor_wgrad(const float* gy, int ldg, const float* x, int R, int N, int K, int nseg, const int* seg_rows,
                         float* const* w_dst, float* const* b_dst, const float* const* seg_scale, int accumulate,
                         void* workspace, size_t workspace_bytes, unit_stream_t stream);


/* ---------------------------------------------------------------------------------------------------------
 * Mask transfer + class select + sigmoid.  Replaces mask_head.py:16-37 / 72-94 and [D2] mask_rcnn_inference.
 *   logits [D,K,M,M]; s_seg [D,Nn,B] (s_is_2d: [Nn,B]); x_delta [D,K,M,M] nullable; pred_classes int64 [D].
 *   out_logits [D,K,M,M] nullable (full transferred tensor); out_probs [D,1,M,M] nullable. */
int unit_mask_transfer(const float* logits, const float* s_seg, int s_is_2d, const int* base, const int* novel,
                       const int* class_kind, const float* x_delta, const int64_t* pred_classes, float* out_logits,
                       float* out_probs, int D, int K, int B, int Nn, int MM, unit_stream_t stream);
/* paste_masks_in_image.  Replaces [D2] layers.mask_ops.paste_masks_in_image via detector_postprocess
 * (meta_arch/rcnn.py:423).  masks [D,M,M] probabilities, boxes [D,4]; out uint8/bool [D,img_h,img_w] = bilinear
 * sample (align_corners=False, zero padding) >= threshold. */
int unit_mask_paste(const float* masks, const float* boxes, int D, int M, int img_h, int img_w, float threshold,
                    uint8_t* out, unit_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Weak-image training losses of the base stage.
 * unit_mil_loss replaces WeakDetectorOutputsBase.losses, weak_detector_fast_rcnn.py:189-214:
 *   per image i (rows img_offsets[i]..img_offsets[i+1]): x = softmax(cls_logits, classes) * softmax(det_logits,
 *   proposals of the image); class_vector[i] = sum_r x; loss = multiplier * mean BCE(clamp(class_vector, 1e-6,
 *   1 - 1e-6), gt_vector).  cls_logits / det_logits [R,K] (already divided by their temperatures), gt_vector [n_img,K]
 *   in {0,1}.  Writes mil_scores [R,K] (= x, the first OICR supervision), class_vector [n_img,K], loss [1] and the
 *   gradients d_cls, d_det [R,K] of loss (for upstream gradient 1).  max_rows = the largest per-image row count (a host
 *   hint that sizes the grid: one CTA per 128 rows of an image; R is always a valid value).  workspace:
 *   unit_mil_loss_workspace_bytes(n_img, max_rows, K).
 * unit_oicr_targets replaces compute_loss_inputs / get_proposal_clusters / label_and_sample_proposals
 *   (:353-408, :308-351): for every class present in gt_vector (ascending == torch.unique order) pick the
 *   image's not-yet-picked proposal with the largest probs[r, c] (first one on ties; picked rows count as 0
 *   afterwards), match every proposal to those pseudo boxes with pairwise_iou + the UniT Matcher
 *   (thresholds_host / labels_host as in unit_matcher), labels = class of the matched box | K (background) | -1
 *   (ignore), weights = score of the matched box, 0 where the matched IoU < bg_threshold (if bg_threshold > 0).
 *   probs [P_total, ld] row-major with ld >= K columns (ld = K for mil_scores, K+1 for a softmax of OICR scores).
 *   pgt_index int64 [n_img,K]: image-local index of the proposal picked for class c, -1 if absent; pgt_scores
 *   [n_img,K] its score.  Images without any present class get labels K and weights 0.
 * unit_weighted_ce_loss replaces weighted_softmax_with_loss (:220-227): loss [1] = mean_r(weights[r] *
 *   cross_entropy(scores[r], labels[r])), d_scores [R,K1] its gradient.  workspace: R floats. */
size_t unit_mil_loss_workspace_bytes(int n_img, int max_rows, int K);
int unit_mil_loss(const float* cls_logits, const float* det_logits, const int* img_offsets, const float* gt_vector,
                  int n_img, int R, int max_rows, int K, float multiplier, float* mil_scores, float* class_vector,
                  float* loss, float* d_cls, float* d_det, void* workspace, size_t workspace_bytes,
                  unit_stream_t stream);
int unit_oicr_targets(const float* probs, int ld, const float* prop_boxes, const int* prop_offsets,
                      const float* gt_vector, int n_img, int P_total, int K, const float* thresholds_host,
                      const int* labels_host, int T, float bg_threshold, int64_t* labels, float* weights,
                      int64_t* pgt_index, float* pgt_scores, unit_stream_t stream);
int unit_weighted_ce_loss(const float* scores, const int64_t* labels, const float* weights, int R, int K1,
                          float* loss, float* d_scores, void* workspace, size_t workspace_bytes,
                          unit_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIT_B200_H_ */
