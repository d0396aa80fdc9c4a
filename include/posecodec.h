/*
 * posecodec.h -- C ABI of libposecodec.so, the sm_100a heatmap codec.
 *
 * This is the drop-in boundary for the heatmap-codec hot path of
 * mindspore-lab/mindpose.  Every entry point names the reference interface it
 * replaces (paths relative to the reference tree).  The reference has no FFI of
 * its own (it is pure Python); the binding a maintainer would add is the ctypes
 * stub shown in INTEGRATION.md, and mindpose_b200/_lib.py is that stub.
 *
 * Conventions
 *  - All pointers named d_* are DEVICE pointers (contiguous, dense, row-major);
 *    pointers named h_* are HOST pointers.  The library never frees or retains
 *    caller memory and never synchronises the stream it is given.  Exceptions:
 *    the *_host entry points own a context with scratch buffers and return after
 *    the results are in the host buffers; pc_bottomup_decode takes a few KB of
 *    stream-ordered scratch (cudaMallocAsync / cudaFreeAsync on the caller's
 *    stream, from the device's default pool).
 *  - Tensors: heatmaps float32 NCHW; images uint8 HWC; keypoints float32
 *    [N, K, 3] = (x, y, visibility).
 *  - Sizes follow the reference's config convention where noted ([w, h]).
 *  - Return value: PC_OK (0) or a negative pc_status; pc_last_error() returns a
 *    thread-local, human readable message for the last failure on this thread.
 *    Argument errors map to Python ValueError (the reference raises ValueError
 *    for bad configuration), CUDA failures to RuntimeError.
 *  - Re-entrant: no global mutable state except one diagnostics counter per
 *    device (pc_bottomup_decode_stats); one stream per call.
 *  - stream is a cudaStream_t passed as void* (0 = legacy default stream).
 */
#ifndef POSECODEC_H_
#define POSECODEC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PC_VERSION 103 /* 0.1.1 */

typedef enum pc_status {
  PC_OK = 0,
  PC_ERR_INVALID_ARGUMENT = -1, /* -> ValueError */
  PC_ERR_UNSUPPORTED = -2,      /* -> ValueError (shape/config outside the kernels' range) */
  PC_ERR_CUDA = -3,             /* -> RuntimeError */
  PC_ERR_NO_DEVICE = -4         /* -> RuntimeError: no sm_100 device / driver */
} pc_status;

#define PC_MAX_JOINTS 64
#define PC_MAX_DARK_KERNEL 17 /* kernel_size <= 17 (sigma = 3 recipe) */
#define PC_MAX_GROUPS 128     /* default people-per-image capacity of the grouping kernels
                               * (pc_group_params.max_groups raises it) */
#define PC_MAX_DETECTIONS 64  /* max_num of the bottom-up decode / grouping (the reference's
                               * configs use 30; 33..64 takes the generic decode kernel) */
#define PC_MAX_SCALES 4       /* heat-map resolutions of the bottom-up target encoder */
#define PC_NMS_MAX_PEOPLE 1024 /* people per image pc_oks_nms can hold */

/* ---- library ----------------------------------------------------------- */

int pc_version(void);
const char* pc_last_error(void);
/* Fills sm_count / compute capability of `device`; PC_ERR_NO_DEVICE if absent. */
int pc_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ---- A1: TopDownBoxToCenterScale._xywh2cs ------------------------------
 * mindpose/data/transform/topdown_transform.py:131-154 (eval branch).
 * d_boxes f32 [N,4] (x,y,w,h) -> d_center f32 [N,2], d_scale f32 [N,2]. */
typedef struct pc_box_params {
  int32_t image_w, image_h; /* dataset_setting.image_size = [w, h] */
  float pixel_std;          /* 200 */
  float scale_padding;      /* 1.25 */
} pc_box_params;
int pc_box_to_center_scale(const float* d_boxes, float* d_center, float* d_scale,
                           const pc_box_params* params, int64_t n, void* stream);

/* ---- A2/A3: get_affine_transform / get_warp_matrix ---------------------
 * mindpose/data/transform/utils.py:44-98 and :158-190, as called from
 * TopDownAffine._affine / _udp_affine (topdown_transform.py:203-261).
 * d_center, d_scale f32 [N,2]; d_rot f32 [N] degrees (NULL = 0).
 * d_fwd  f64 [N,6]: forward 2x3 matrix (source image -> crop), the matrix the
 *        reference hands to cv2.warpAffine (UDP: its float32 value, widened).
 * d_inv  f64 [N,6]: its inverse in cv::warpAffine's op order (crop -> source).
 * Either output may be NULL. */
typedef struct pc_affine_params {
  int32_t image_w, image_h;
  float pixel_std;
  int32_t use_udp;
} pc_affine_params;
int pc_affine_matrices(const float* d_center, const float* d_scale, const float* d_rot,
                       double* d_fwd, double* d_inv, const pc_affine_params* params,
                       int64_t n, void* stream);
/* cv2.getAffineTransform(src, dst) op for op, for point triples the CALLER built (float32
 * [N,3,2] each, as the reference stores them: utils.py:81-96), plus the inverse.  With a
 * rotation the reference's three points depend on numpy's own sin / cos and on the dtype the
 * rotation arrives in (float32 from the dataset pipeline, transform.py:66-79); the per-sample
 * transforms evaluate those few scalars with numpy on the host and hand the points here, so
 * that rotated crops are bit-exact too (pc_affine_matrices uses the device's sin / cos:
 * identical at rot = 0, <= 1e-9 relative otherwise). */
int pc_affine_from_points(const float* d_src_points, const float* d_dst_points, double* d_fwd,
                          double* d_inv, int64_t n, void* stream);
/* Invert caller-supplied forward matrices (e.g. from cv2.getAffineTransform). */
int pc_invert_affine(const double* d_fwd, double* d_inv, int64_t n, void* stream);

/* ---- A4: cv2.warpAffine(image, M, (w,h), INTER_LINEAR) -----------------
 * topdown_transform.py:217-222 / :248-253.  OpenCV-exact fixed-point bilinear,
 * BORDER_CONSTANT 0, uint8, C channels interleaved (C in 1..4).
 * Crop i reads the dense HWC image at d_src + d_src_offset[i] (bytes) of size
 * d_src_hw[i] = (rows, cols); several crops may share one source image.
 * d_inv f64 [N,6] from pc_affine_matrices / pc_invert_affine.
 * d_dst u8 [N, dst_h, dst_w, C].
 * Memory the call touches besides its arguments: the aligned 32-bit words (quad kernel) or
 * 16-byte units (band kernel: its bulk copies are 16-byte aligned) that contain the first and
 * the last byte of a source row may be read up to 15 bytes beyond the row; d_src must therefore
 * lie in an allocation whose start and size are multiples of 16 bytes (every CUDA allocator's
 * are), and those bytes are never used.  3-channel crops of a width that is a multiple of 32
 * take a list of (1 + n * ceil(dst_h / 16)) ints from the library's stream-ordered pool on
 * `stream` for the duration of the call (capturable into a CUDA graph). */
typedef struct pc_warp_params {
  int32_t dst_w, dst_h, channels;
} pc_warp_params;
int pc_warp_affine_u8(const uint8_t* d_src, const int64_t* d_src_offset,
                      const int32_t* d_src_hw, const double* d_inv, uint8_t* d_dst,
                      const pc_warp_params* params, int64_t n, void* stream);

/* ---- N2: warp fused with vision.Normalize + HWC2CHW ----------------------
 * The step after the warp in the reference's pipeline
 * (mindpose/data/data_factory.py:127-138): Normalize(mean * 255, std * 255), then
 * HWC -> CHW.  Same sources / matrices as pc_warp_affine_u8; the uint8 crop is not
 * materialised: d_dst f32 [N, 3, dst_h, dst_w] = (crop - mean[c]) / std[c].
 * mean / std are the values handed to vision.Normalize (already multiplied by 255).
 * 3-channel images, dst_w % 4 == 0. */
typedef struct pc_warp_norm_params {
  int32_t dst_w, dst_h, channels;
  float mean[4], std[4];
} pc_warp_norm_params;
int pc_warp_affine_u8_norm_chw(const uint8_t* d_src, const int64_t* d_src_offset,
                               const int32_t* d_src_hw, const double* d_inv, float* d_dst,
                               const pc_warp_norm_params* params, int64_t n, void* stream);

/* ---- bottom-up evaluation preprocessing: rescale + pad + mask -------------
 * BottomUpRescale.transform (mindpose/data/transform/bottomup_transform.py:170-209:
 * cv2.resize(image, target, interpolation=cv2.INTER_LINEAR)) followed by
 * BottomUpPad.transform (:610-648: zeros right of / below the image up to max_image_size,
 * mask 1 on the image and 0 on the padding) -- the validation transforms of the shipped
 * HigherHRNet recipe -- for a batch of images in one pass.  Sources as in pc_warp_affine_u8
 * (one base pointer, a byte offset and (height, width) per image, HWC uint8, 3 channels);
 * d_dst_wh i32 [N,2] = the (width, height) BottomUpRescale._get_new_size gives each image
 * (host arithmetic: the Python mirror computes it); d_dst u8 [N, canvas_h, canvas_w, 3];
 * d_mask u8 [N, canvas_h, canvas_w] or NULL.  A target larger than the canvas is cut at the
 * canvas (the reference asserts; the Python mirror raises before the call).  Arithmetic:
 * OpenCV's 8-bit bilinear resize, bit for bit (oracle/resize.py), including the 2 x 2 area
 * path it takes when the source is exactly twice the target. */
int pc_rescale_pad_u8(const uint8_t* d_src, const int64_t* d_src_offset,
                      const int32_t* d_src_hw, const int32_t* d_dst_wh, uint8_t* d_dst,
                      uint8_t* d_mask, int32_t canvas_w, int32_t canvas_h, int32_t channels,
                      int64_t n, void* stream);

/* The same with the pipeline's next step fused, as pc_warp_affine_u8_norm_chw does for the crop
 * warp: vision.Normalize(mean * 255, std * 255) + HWC2CHW (mindpose/data/data_factory.py:127-138).
 * d_dst f32 [N, 3, canvas_h, canvas_w] = (pixel - mean[c]) / std[c]; the uint8 canvas is not
 * materialised and the zero padding becomes (0 - mean[c]) / std[c], which is what Normalize makes
 * of it in the reference's pipeline.  params: dst_w / dst_h = the canvas, mean / std as handed to
 * vision.Normalize.  Any canvas width. */
int pc_rescale_pad_u8_norm_chw(const uint8_t* d_src, const int64_t* d_src_offset,
                               const int32_t* d_src_hw, const int32_t* d_dst_wh, float* d_dst,
                               uint8_t* d_mask, const pc_warp_norm_params* params, int64_t n,
                               void* stream);

/* Keypoint half of TopDownAffine (topdown_transform.py:224-231 / :255-259):
 * in place on d_keypoints f32 [N,K,3]; standard path moves joints with
 * visibility > 0 only, UDP moves all joints (matrix taken as float32). */
int pc_affine_joints(float* d_keypoints, const double* d_fwd, int32_t num_joints,
                     int32_t use_udp, int64_t n, void* stream);

/* ---- A5/A6: TopDownGenerateTarget._encoding / _udp_encoding ------------
 * topdown_transform.py:324-375 / :377-430.
 * d_keypoints f32 [N,K,3] (crop coordinates) ->
 * d_target f32 [N,K,H,W], d_target_weight f32 [N,K]. */
typedef struct pc_encode_params {
  int32_t num_joints;
  int32_t image_w, image_h;     /* image_size  = [w, h] */
  int32_t heatmap_w, heatmap_h; /* heatmap_size = [w, h] */
  float sigma;                  /* 3*sigma must be an integer <= 15 */
  int32_t use_udp;
  int32_t use_joint_weights;    /* use_different_joint_weights */
  float joint_weights[PC_MAX_JOINTS];
} pc_encode_params;
int pc_topdown_encode(const float* d_keypoints, float* d_target, float* d_target_weight,
                      const pc_encode_params* params, int64_t n, void* stream);

/* ---- A8-A13: TopDownHeatMapDecoder.construct (+ flip-test averaging) ----
 * mindpose/models/decoders/top_down_decoder.py:72-215 and the post-network
 * half of _MultiRunNet.construct (engine/inferencer/topdown_inferencer.py:165-187).
 * d_heatmap f32 [N,K,H,W]; d_flipped f32 [N,K,H,W] or NULL (required iff
 * flip_test); d_center, d_scale f32 [N,2]; d_score f32 [N].
 * -> d_all_preds f32 [N,K,3] (x, y, maxval), d_all_boxes f32 [N,6].
 * One pass: each heatmap element is read from HBM exactly once. */
typedef struct pc_topdown_decode_params {
  int32_t num_joints, height, width;
  float pixel_std;
  int32_t to_original;
  int32_t shift_coordinate;
  int32_t use_udp;
  int32_t dark_udp_refine;
  int32_t kernel_size;
  int32_t flip_test;     /* average with d_flipped (channel permute + reversed x) */
  int32_t shift_heatmap; /* eval_setting.shift_heatmap */
  int32_t flip_index[PC_MAX_JOINTS];
  /* Optional explicit blur kernel (decoder.gaussian_kernel, row-major
   * kernel_size^2 floats).  dark_kernel_set = 0: built as
   * _create_gaussian_kernel does (top_down_decoder.py:207-215). */
  int32_t dark_kernel_set;
  float dark_kernel[PC_MAX_DARK_KERNEL * PC_MAX_DARK_KERNEL];
} pc_topdown_decode_params;
int pc_topdown_decode(const float* d_heatmap, const float* d_flipped, const float* d_center,
                      const float* d_scale, const float* d_score, float* d_all_preds,
                      float* d_all_boxes, const pc_topdown_decode_params* params, int64_t n,
                      void* stream);

/* ---- N1: BottomUpGenerateTarget._encoding (+ pad_to_same) ----------------
 * mindpose/data/transform/bottomup_transform.py:504-598 and
 * mindpose/data/transform/utils.py:213-232.
 * d_keypoints f32 [N, S, M, K, 3]: for every scale s the M people's joints in
 * heat-map pixels OF THAT SCALE (x, y, visibility); people are padded to M per
 * batch with visibility 0.
 * -> d_target f32 [N, S, K, Hmax, Wmax] (each scale zero-padded at the bottom /
 *    right to the largest map), d_tag_ind i32 [N, S, max_num, K, 2] with
 *    (mu_y * W_s + mu_x, 1) per visible joint, or [N, S, max_num, 2] when
 *    tag_per_joint = 0.  num_people > max_num is an argument error (the reference
 *    raises ValueError). */
typedef struct pc_bottomup_encode_params {
  int32_t num_joints, num_scales, num_people, max_num;
  int32_t heatmap_w[PC_MAX_SCALES], heatmap_h[PC_MAX_SCALES]; /* heatmap_sizes = [[w, h], ...] */
  float sigma;                                                /* 3*sigma integer <= 15 */
  int32_t tag_per_joint;
} pc_bottomup_encode_params;
int pc_bottomup_encode(const float* d_keypoints, float* d_target, int32_t* d_tag_ind,
                       const pc_bottomup_encode_params* params, int64_t n, void* stream);

/* ---- A14-A17: BottomUpHeatMapAEDecoder.construct ------------------------
 * mindpose/models/decoders/bottom_up_decoder.py:67-203.
 * num_stages = 2, with_ae_loss = [True, False] (the HigherHRNet recipe):
 * d_out0 f32 [N, C0, H0, W0] (heat | tag) with C0 = 2K (tag_per_joint, a tag plane per joint)
 * or K + 1 (tag_per_joint = 0: every joint reads the one shared tag plane, the reference's
 * broadcast at bottom_up_decoder.py:159-160), d_out1 f32 [N, K, H1, W1] with H1 = 2*H0,
 * W1 = 2*W0;  num_stages = 1: d_out0 f32 [N, C0, H1, W1], d_out1 = NULL.
 * d_mask u8 [N, Hm, Wm] (non-zero = valid).
 * -> d_val_k f32 [N,K,M], d_tag_k f32 [N,K,M,1], d_ind_k f32 [N,K,M,2] (x, y).
 * Optional (may be NULL): d_heatmap_raw f32 [N,K,H1,W1] (aggregated, masked,
 * pre-NMS), d_tagging f32 [N,K,H1,W1,1] ([N,1,H1,W1,1] with tag_per_joint = 0). */
typedef struct pc_bottomup_decode_params {
  int32_t num_joints;
  int32_t num_stages;
  int32_t h0, w0; /* stage-0 map size (ignored when num_stages == 1) */
  int32_t h1, w1; /* output / highest-resolution map size */
  int32_t mask_h, mask_w;
  int32_t use_nms, nms_kernel;
  int32_t max_num;          /* M <= PC_MAX_DETECTIONS (33..64: generic kernel) */
  int32_t shift_coordinate; /* A17, reproduced with the reference's pairing: entry t of the
                             * top M gets the +-0.25 offset of the t-th position in
                             * row-major order (bottom_up_decoder.py:195-201) */
  int32_t tag_per_joint;    /* 1: a tag plane per joint; 0: one plane shared by all joints */
} pc_bottomup_decode_params;
int pc_bottomup_decode(const float* d_out0, const float* d_out1, const uint8_t* d_mask,
                       float* d_val_k, float* d_tag_k, float* d_ind_k, float* d_heatmap_raw,
                       float* d_tagging, const pc_bottomup_decode_params* params, int64_t n,
                       void* stream);
/* Diagnostics: number of (image, joint) planes whose per-lane top-3 pass could not prove
 * the top M and that were re-scanned by the exact pass, on the current device, since the
 * last reset.  Synchronises the device.  Results are exact either way; a high count only
 * costs time (planes with fewer than M positive local maxima always take the exact pass). */
int pc_bottomup_decode_stats(int64_t* exact_pass_planes, int reset);

/* ---- A18/A19: match_by_tag + instance score + transform_keypoints -------
 * mindpose/utils/match.py:14-116, engine/inferencer/bottomup_inferencer.py:
 * 153-156 and data/transform/utils.py:235-274.
 * d_val_k [N,K,M], d_tag_k [N,K,M,1], d_ind_k [N,K,M,2] ->
 * d_ans f32 [N, G, K, 4] (x, y, val, tag; insertion order), d_num_groups i32 [N],
 * d_scores f32 [N, G], with G = max_groups (0 = PC_MAX_GROUPS) the people an image can hold.
 * The reference is unbounded (match.py:63-113); at most num_joints * max_num groups can form
 * (every detection its own), so max_groups = num_joints * max_num never overflows (510 for
 * the HigherHRNet recipe; the shared memory of the kernel bounds G at about 900 for K = 17).
 * d_num_groups[i] = -1 flags an image whose group count exceeded G: call again with more. */
typedef struct pc_group_params {
  int32_t num_joints, max_num;
  float vis_thr, tag_thr;
  int32_t ignore_too_much, use_rounded_norm;
  int32_t joint_order[PC_MAX_JOINTS];
  int32_t max_groups; /* G; 0 = PC_MAX_GROUPS */
} pc_group_params;
int pc_group_by_tag(const float* d_val_k, const float* d_tag_k, const float* d_ind_k,
                    float* d_ans, int32_t* d_num_groups, float* d_scores,
                    const pc_group_params* params, int64_t n, void* stream);
/* Back-projection of grouped people, in place on d_ans f32 [N, G, K, 4] (utils.py:235-274).
 * d_center, d_scale f64 [N,2]; d_heatmap_wh f64 [N,2] (= image_shape / downsample_scale);
 * max_groups = G as in pc_group_params (0 = PC_MAX_GROUPS). */
int pc_transform_keypoints(float* d_ans, const int32_t* d_num_groups, const double* d_center,
                           const double* d_scale, const double* d_heatmap_wh, float pixel_std,
                           int32_t num_joints, int32_t max_groups, int64_t n, void* stream);

/* ---- N3: BottomUpHeatMapAEInferencer._refine_missing --------------------
 * mindpose/engine/inferencer/bottomup_inferencer.py:189-249, for every person of
 * every image.  Runs on the grouped people in heat-map coordinates, i.e. after
 * pc_group_by_tag and before pc_transform_keypoints (bottomup_inferencer.py:158-166).
 * d_heatmap f32 [N,K,H,W] and d_tagging f32 [N,K,H,W,1] are the heatmap_raw /
 * tagging_heatmap outputs of pc_bottomup_decode; d_ans f32 [N, G, K, 4]
 * is updated in place (x, y, val of joints with val == 0); d_mean_tag f32
 * [N, G] is caller-provided scratch; G = max_groups (0 = PC_MAX_GROUPS). */
typedef struct pc_refine_params {
  int32_t num_joints, height, width;
  int32_t max_groups;
} pc_refine_params;
int pc_refine_missing(const float* d_heatmap, const float* d_tagging, float* d_ans,
                      const int32_t* d_num_groups, float* d_mean_tag,
                      const pc_refine_params* params, int64_t n, void* stream);

/* ---- N4: OKS rescoring + oks_nms / soft_oks_nms -------------------------
 * mindpose/engine/evaluator/topdown_evaluator.py:93-121 (rescoring loop, NMS call) and
 * mindpose/utils/nms.py:7-190 (oks_iou, oks_nms, _rescore, soft_oks_nms), for every image
 * of an evaluation in one launch.
 * People are grouped by image: image i owns rows [d_image_offset[i], d_image_offset[i+1])
 * of d_kpts f32 [P,K,3] (x, y, score), d_area f32 [P] and d_score f32 [P], already sorted
 * and de-duplicated by bbox_id (_sort_and_unique_bboxes, host list handling).
 * rescore != 0: d_score holds the box scores on entry and the rescored scores on return
 * (mean of the joint scores > rescore_vis_thr, times the box score).
 * -> d_keep i32 [P]: entries [offset[i], offset[i] + d_num_keep[i]) are the kept people of
 * image i as indices LOCAL to the image, in keep order; the rest is -1.  d_num_keep i32 [I];
 * -1 flags an image with more than max_people_per_image people.
 * Equal scores: the reference's argsort is unstable; here (score desc, position desc),
 * i.e. a stable ascending sort reversed. */
typedef struct pc_oks_nms_params {
  int32_t num_joints;
  int32_t rescore;         /* apply the evaluator's rescoring first */
  int32_t use_nms;         /* 0: keep everybody (evaluation config use_nms: False) */
  int32_t soft;            /* 0: oks_nms, 1: soft_oks_nms (gaussian) */
  int32_t max_dets;        /* soft_oks_nms max_dets (reference default 20) */
  int32_t use_iou_vis_thr; /* oks_iou's vis_thr is not None */
  float rescore_vis_thr;   /* evaluation config vis_thr */
  float oks_thr;           /* evaluation config oks_thr */
  float iou_vis_thr;       /* oks_iou vis_thr: joints of the DETECTION above it (nms.py:64) */
  int32_t max_people_per_image; /* >= max_i(offset[i+1] - offset[i]); <= PC_NMS_MAX_PEOPLE */
  double sigmas[PC_MAX_JOINTS];
  double rescore_vis_thr_f64; /* pc_oks_nms_f64 only: the thresholds as Python floats */
  double iou_vis_thr_f64;
} pc_oks_nms_params;
int pc_oks_nms(const float* d_kpts, const float* d_area, float* d_score,
               const int32_t* d_image_offset, int32_t* d_keep, int32_t* d_num_keep,
               const pc_oks_nms_params* params, int64_t num_images, void* stream);
/* The same for records that hold Python floats, which is what the reference's inferencer
 * emits (`pred.tolist()`, `box.tolist()`: mindpose/engine/inferencer/topdown_inferencer.py:
 * 135-140): numpy then runs the rescoring, dx**2 + dy**2, the areas and the sort keys in
 * float64 (the OKS values stay float32, nms.py:56).  d_kpts f64 [P,K,3], d_area f64 [P],
 * d_score f64 [P]; thresholds from the *_f64 fields (oks_thr stays float32: it is only
 * compared with float32 OKS values). */
int pc_oks_nms_f64(const double* d_kpts, const double* d_area, double* d_score,
                   const int32_t* d_image_offset, int32_t* d_keep, int32_t* d_num_keep,
                   const pc_oks_nms_params* params, int64_t num_images, void* stream);

/* ---- E: all-gather of the decoded keypoints over peer memory --------------
 * Not in the reference (it evaluates on rank 0, mindpose/callbacks/eval_callback.py:
 * 142-145); SURVEY.md section 8(e).  Packs this rank's d_preds f32 [n,K,3] + d_boxes f32
 * [n,6] into rows [row_offset, row_offset + n) of the gathered table f32 [total, K*3+6] OF
 * EVERY RANK in one kernel: through d_multicast_table, the NVSwitch multicast mapping of
 * the table (one multimem.st reaches all replicas), or, when that is NULL, through
 * h_peer_tables, a HOST array of num_peers device pointers -- the peer-mapped address of
 * the table on each rank, this rank's own included.  The tables are symmetric-memory
 * allocations owned by the caller, who orders the ranks with a barrier afterwards
 * (mindpose_b200/dist.py::PeerGather). */
int pc_scatter_results(const float* d_preds, const float* d_boxes, void* const* h_peer_tables,
                       int32_t num_peers, void* d_multicast_table, int64_t row_offset,
                       int32_t num_joints, int64_t n, void* stream);
/* The same with the ordering folded in: after its stores the kernel's last CTA increments
 * the step number *d_step (one u32 in this rank's memory, zero-initialised by the caller) and
 * writes it (release, system scope) into word my_rank of the flag array of every rank --
 * h_peer_flags is a HOST array of num_flag_peers device pointers to u32 [num_flag_peers]
 * arrays in symmetric memory, this rank's own included; d_counter is one zero-initialised
 * i32 in this rank's memory (the kernel leaves it at zero).  The step number lives in device
 * memory so that a captured CUDA graph of the step signals a new number at every replay.
 * A rank reads the gathered table of its scatter number s after pc_wait_peer_flags(its own
 * flag array, num_peers, d_step, lag) with lag = (scatters issued so far) - s: one tiny
 * kernel that returns once every source rank has published a step >= *d_step - lag (compared
 * modulo 2^32; a peer that never arrives traps the kernel after about 13 s instead of hanging
 * the GPU).  A table may be rewritten by a peer as soon as that peer has seen this rank's flag
 * for a LATER step, so callers alternate three tables when the wait is deferred by one step. */
int pc_scatter_results_signal(const float* d_preds, const float* d_boxes,
                              void* const* h_peer_tables, int32_t num_peers,
                              void* d_multicast_table, int64_t row_offset, int32_t num_joints,
                              int64_t n, void* const* h_peer_flags, int32_t num_flag_peers,
                              int32_t my_rank, uint32_t* d_step, int32_t* d_counter,
                              void* stream);
int pc_wait_peer_flags(const uint32_t* d_flags, int32_t num_peers, const uint32_t* d_step,
                       uint32_t lag, void* stream);

/* The decode and the all-gather as ONE kernel: pc_topdown_decode whose result stores ALSO go,
 * value by value as they are produced, into rows [row_offset, row_offset + n) of the gathered
 * table of every rank (multicast or peer-mapped, as in pc_scatter_results), and whose last
 * warp publishes the step like pc_scatter_results_signal does.  No scatter kernel, no second
 * pass over the results; the exchange overlaps the decode.  All fields as in
 * pc_scatter_results_signal.  d_all_preds / d_all_boxes still receive the local results. */
typedef struct pc_gather_target {
  void* const* h_peer_tables; /* HOST array of num_peers device pointers (NULL with multicast) */
  int32_t num_peers;
  void* d_multicast_table;    /* or NULL */
  int64_t row_offset;
  void* const* h_peer_flags;  /* HOST array of num_flag_peers device pointers */
  int32_t num_flag_peers, my_rank;
  uint32_t* d_step;
  int32_t* d_counter;
} pc_gather_target;
int pc_topdown_decode_gather(const float* d_heatmap, const float* d_flipped,
                             const float* d_center, const float* d_scale, const float* d_score,
                             float* d_all_preds, float* d_all_boxes,
                             const pc_topdown_decode_params* params, int64_t n,
                             const pc_gather_target* gather, void* stream);

/* ---- host-buffer front end (what the e2e number is measured through) ----
 * Same decode as pc_topdown_decode but every pointer is a HOST pointer.  The
 * context owns device scratch and two streams; crops are streamed through in
 * chunks so the host->device copy of chunk i+1 overlaps the kernel of chunk i.
 * A context serves one call at a time (one per host thread; any number of
 * contexts may exist); the *_host calls return after their results are in the
 * host buffers. */
typedef struct pc_ctx pc_ctx;
int pc_ctx_create(int device, int64_t scratch_bytes, pc_ctx** out);
int pc_ctx_destroy(pc_ctx* ctx);
/* Bytes the last *_host call on this context moved host->device and device->host
 * (what bench.py reports as h2d_bytes_per_step / d2h_bytes_per_step). */
int pc_ctx_last_transfer_bytes(const pc_ctx* ctx, int64_t* h2d_bytes, int64_t* d2h_bytes);
int pc_topdown_decode_host(pc_ctx* ctx, const float* h_heatmap, const float* h_flipped,
                           const float* h_center, const float* h_scale, const float* h_score,
                           float* h_all_preds, float* h_all_boxes,
                           const pc_topdown_decode_params* params, int64_t n);

/* Host-buffer crop warp: TopDownBoxToCenterScale + TopDownAffine for N crops, one
 * dense source image per crop, all of one size (h_images u8 [N, src_h, src_w, C]).
 * h_boxes f32 [N,4] (x,y,w,h); h_rot f32 [N] or NULL.
 * -> h_crops u8 [N, image_h, image_w, C]; optional h_center / h_scale f32 [N,2]. */
/* upload: which bytes of each source image cross PCIe, and how.  The warp only samples the
 * padded box of a crop (about 40 % of a 480x640 image on the bench data), so the front end
 * can upload just that rectangle (computed on the host from the box, padded by 4 pixels) into
 * the full-image layout the kernel reads: the crops are the same bits, the bytes outside the
 * rectangle are never read. */
#define PC_UPLOAD_FULL 0 /* whole source images, one contiguous copy per chunk */
#define PC_UPLOAD_ROI 1  /* the sampled rectangle of each image, one strided DMA copy per crop */
/* The rectangles (rows widened to 64-byte boundaries) are fetched by ONE kernel per chunk that
 * reads the caller's PINNED host buffer through its device mapping (16-byte loads over PCIe)
 * and writes the scratch: no per-crop copy-engine launch (4.6 us each, measured).  Needs
 * page-locked h_images (cudaHostAlloc / cudaHostRegister / torch pin_memory) and a row pitch
 * and image size that are multiples of 16 bytes; otherwise PC_UPLOAD_ROI is used. */
#define PC_UPLOAD_ROI_KERNEL 2
typedef struct pc_affine_host_params {
  int32_t src_h, src_w, channels;
  int32_t image_w, image_h; /* crop size, dataset_setting.image_size = [w, h] */
  float pixel_std, scale_padding;
  int32_t use_udp;
  int32_t upload; /* PC_UPLOAD_* */
} pc_affine_host_params;
/* Host-only helper (no CUDA call): the rectangle [x0, x1) x [y0, y1) of source pixels that
 * PC_UPLOAD_ROI uploads for one crop -> h_rect int32 [4] = (x0, y0, x1, y1), clipped to the
 * image (empty when x1 <= x0 or y1 <= y0).  h_box f32 [4] = (x, y, w, h); rot in degrees.
 * Returns PC_ERR_UNSUPPORTED for a box that is not finite or is degenerate (the front end
 * then uploads the whole image). */
int pc_crop_source_rect(const float* h_box, float rot, const pc_affine_host_params* params,
                        int32_t* h_rect);
int pc_topdown_affine_host(pc_ctx* ctx, const uint8_t* h_images, const float* h_boxes,
                           const float* h_rot, uint8_t* h_crops, float* h_center,
                           float* h_scale, const pc_affine_host_params* params, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* POSECODEC_H_ */
