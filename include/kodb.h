/*
 * kodb.h — C ABI of libkodb.so: the B200-native (sm_100a) YOLOv5 detection tail for
 * craston/object_detection_cib (`kod`).
 *
 * The reference has no FFI layer: its seams are Python names selected through Hydra `_target_`
 * strings.  Every entry point below cites the reference interface it replaces; the
 * Python drop-ins in `object_detection_cib_b200/` bind these symbols with ctypes.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the parameter name ends in `_h` (host).
 *  - The caller owns every buffer (inputs, outputs, workspaces); the library never allocates or frees
 *    device memory and keeps no global state.  Workspaces must be zero-filled once after allocation
 *    (by the caller or with kodb_workspace_reset); kernels leave them in a reusable state.  After a failed
 *    launch (KODB_ERR_LAUNCH, a CUDA error on the stream) call kodb_workspace_reset before reusing one.
 *  - Work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises the host.
 *  - Return value: 0 = OK, negative = error (KODB_ERR_*); nothing throws across the ABI.
 *    `kodb_error_string(code)` gives a static description.
 *  - Head tensors are the reference's planar layout: logically (B, A, H, W, P), physically
 *    [B][A][P][H][W] with the H*W plane contiguous (kod/nn/heads/yolov5.py:29-32 `Rearrange`);
 *    strides are given in ELEMENTS for the batch, anchor and channel dimensions.
 */
#ifndef KODB_H_
#define KODB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KODB_NUM_LEVELS 3   /* ll, ml, hl  (Yolov5NetworkResult, kod/nn/networks/yolov5.py:34-37) */
#define KODB_MAX_ANCHORS 8  /* num_anchors_per_cell (configs/nn/networks/yv5.yaml:2 uses 3) */
#define KODB_MAX_CLASSES 128
#define KODB_NUM_IOU_THRS 4 /* ious=[0.3,0.5,0.75,0.9]  (pycoco_map_eval.py:45-48) */

#define KODB_OK 0
#define KODB_ERR_INVALID_ARGUMENT (-1)
#define KODB_ERR_CAPACITY (-2)
#define KODB_ERR_LAUNCH (-3)
#define KODB_ERR_UNSUPPORTED (-4)

/* IoUType (kod/core/bbox/iou.py:9-14) */
#define KODB_IOU_IOU 0
#define KODB_IOU_GIOU 1
#define KODB_IOU_DIOU 2
#define KODB_IOU_CIOU 3

/* dtype of the ground-truth boxes (DetectionTarget.boxes, kod/data/detection.py:24-26) */
#define KODB_BOXES_F32 0
#define KODB_BOXES_F64 1

/* One pyramid level: AnchorBoxInfo(stride, boxes_wh) (kod/core/anchors/info.py:9-11) plus the grid
 * derived from the image shape exactly as the reference derives it. */
typedef struct kodb_level_cfg {
    int32_t stride;       /* pixels per cell */
    int32_t H, W;         /* grid: img_h // stride, img_w // stride  (yv5.py:265-268, layers.py:24-27) */
    int32_t num_anchors;  /* A */
    float anchor_px[KODB_MAX_ANCHORS][2];   /* (w,h) pixels, as fp32      (layers.py:47-48) */
    float anchor_grid[KODB_MAX_ANCHORS][2]; /* fp32(w * 1 / stride), ...  (yv5.py:227-239) */
    float fm_w, fm_h;     /* fp32(img_w / stride), fp32(img_h / stride)   (yv5.py:184-191) */
    float obj_ratio;      /* lambda_{ll,ml,hl}_objectness                 (loss.py:192-209) */
} kodb_level_cfg;

/* One level of DetectionHeadResult(box, obj, cls) (kod/nn/heads/types.py:8-11). */
typedef struct kodb_head_level {
    const float* box; /* (B,A,H,W,4) view */
    const float* obj; /* (B,A,H,W,1) view */
    const float* cls; /* (B,A,H,W,C) view */
    int64_t box_sb, box_sa, box_sp; /* element strides: batch, anchor, channel */
    int64_t obj_sb, obj_sa;
    int64_t cls_sb, cls_sa, cls_sp;
} kodb_head_level;

/* Gradient (or any writable) planes with the same logical shape as kodb_head_level. */
typedef struct kodb_grad_level {
    float* box;
    float* obj;
    float* cls;
    int64_t box_sb, box_sa, box_sp;
    int64_t obj_sb, obj_sa;
    int64_t cls_sb, cls_sa, cls_sp;
} kodb_grad_level;

/* Flat ground truth of one batch: the reference's tuple of B DetectionTarget(boxes, labels)
 * (kod/lightning/data_module.py:82-87) concatenated in image order. */
typedef struct kodb_targets {
    const void* boxes;          /* (T,4) xyxy pixels, fp32 or fp64 (boxes_dtype) */
    const int64_t* labels;      /* (T,) */
    const int32_t* img_offsets; /* (B+1,) prefix of per-image counts */
    int32_t boxes_dtype;        /* KODB_BOXES_F32 / KODB_BOXES_F64 */
    int32_t num_targets;        /* T */
    int32_t batch;              /* B */
    int32_t max_per_image;      /* max_b (img_offsets[b+1]-img_offsets[b]), known on the host; the matching stage
                                   keeps up to 512 boxes per image in shared memory and reads them from global
                                   memory beyond that (slower, same results) */
} kodb_targets;

/* Yolov5LossParams (loss.py:25-43) + IoUCalculator(iou_type, eps) (iou.py:249-268). */
typedef struct kodb_loss_cfg {
    float lambda_classification, lambda_localization, lambda_objectness;
    int32_t iou_type;  /* KODB_IOU_* */
    float eps;
    int32_t num_classes;
    int32_t img_w, img_h; /* image_feature_shape (loss.py:231-233 scales obj by (img_w/640)^2) */
    float assign_threshold; /* Yolov5LabelAssigner.threshold (configs/assigners/yv5.yaml:7) */
} kodb_loss_cfg;

/* Per-level outputs of the assigner: AssignmentTargetInfo (yv5.py:31-36), capacity `cap` rows. */
typedef struct kodb_assign_out {
    int64_t* samples;
    int64_t* anchors;
    int64_t* grid_y;
    int64_t* grid_x;
    int64_t* labels;
    float* gt_boxes;  /* (cap,4)  (cx-gi, cy-gj, w, h) grid units */
    float* anchor_wh; /* (cap,2)  anchor / stride */
} kodb_assign_out;

int kodb_version(void);
const char* kodb_error_string(int code);

/* ---- train tail --------------------------------------------------------------------------------
 * kodb_assign: Yolov5LabelAssigner.__call__ (kod/core/label_assignment/yv5.py:298-319), all three
 * levels, rows in the reference's order (offset-major, anchor, target).  counts[l] receives N_l; rows
 * beyond `cap` are counted but not written (the caller sizes cap = 5*A*T, the exact upper bound). */
int kodb_assign(const kodb_level_cfg* levels_h, const kodb_targets* targets_h, float threshold,
                const kodb_assign_out* out_h /* [3] */, int32_t cap, int32_t* counts /* dev [3] */,
                void* stream);

/* kodb_train_tail: Yolov5Loss.forward (kod/lightning/experiments/yv5_baseline/loss.py:166-248)
 * fused with its backward.  Computes
 *     losses[0..2] = (localization, objectness, classification)   (scaled, loss.py:240-248)
 * and, when `grads_h` is non-NULL, d(u0*loc + u1*obj + u2*cls)/d(head logits) for all nine head
 * tensors (every element of the grad planes is written).  u = `upstream_dev` (device float[3]) if
 * non-NULL else `upstream_h`.  If `skip_if_equal` != 0 and upstream_dev == upstream_h exactly, the
 * kernel exits without touching anything (used by autograd backward when the forward already
 * produced the gradients for the hinted upstream, e.g. training_step's batch_size factor,
 * exp.py:124-130).  n_assigned (device int32[3], may be NULL) receives N_l.
 * workspace: kodb_train_workspace_bytes(batch, max boxes per image, max anchors per cell over the levels, C) bytes,
 * zero-initialised once.  There is no limit on the number of boxes per image: the per-(image, level) row records
 * live in shared memory while they fit (~58 boxes per image with 3 anchors) and in the workspace beyond that
 * (slower, same results).
 * A target label outside [0, C) (IndexError in the reference, loss.py:145-152) turns the three losses into NaN and
 * raises bit 0 of the status word (kodb_train_status); the host wrappers reject such labels before the launch when the
 * labels are still on the host. */
int64_t kodb_train_workspace_bytes(int32_t batch, int32_t max_targets_per_image, int32_t max_anchors_per_cell,
                                   int32_t num_classes);
/* Status bits of the most recent kodb_train_tail launch that used `workspace` (synchronises `stream`):
 * bit 0 = label out of range, bit 1 = a bounded in-kernel wait expired (results of that launch are invalid). */
int kodb_train_status(const void* workspace, int32_t* status_h, void* stream);
/* Asynchronous zero-fill of any kodb workspace on `stream`: first use, or recovery after a failed launch. */
int kodb_workspace_reset(void* workspace, int64_t workspace_bytes, void* stream);
int kodb_train_tail(const kodb_head_level* heads_h /* [3] */, const kodb_level_cfg* levels_h /* [3] */,
                    const kodb_targets* targets_h, const kodb_loss_cfg* cfg_h,
                    const float* pos_weight /* dev [C] or NULL: Yolov5Loss.weights, loss.py:58-63 */,
                    const float* upstream_h /* host [3] */, const float* upstream_dev /* dev [3] or NULL */,
                    int32_t skip_if_equal, const kodb_grad_level* grads_h /* [3] or NULL */,
                    float* losses /* dev [3] */, int32_t* n_assigned /* dev [3] or NULL */,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* kodb_iou_aligned: IoUCalculator.__call__(boxes1, boxes2) (kod/core/bbox/iou.py:249-268 -> compute_iou :77-95,
 * compute_giou :136-161, compute_diou :164-197, compute_ciou :200-246) on two aligned (n, 4) xyxy fp32 arrays.
 * out (n,) receives the metric (may be NULL when only gradients are wanted).  If grad1 / grad2 are non-NULL they
 * receive upstream[i] * d(metric_i)/d(boxes1_i) / d(boxes2_i) as (n, 4) (upstream NULL = ones): the gradient autograd
 * derives from the reference's op chain (CIoU's alpha is a constant, iou.py:241-242).  16-byte aligned pointers. */
int kodb_iou_aligned(const float* boxes1, const float* boxes2, int64_t n, int32_t iou_type, float eps, float* out,
                     const float* upstream, float* grad1, float* grad2, void* stream);

/* One pyramid level of the Detect head (kod/nn/heads/yolov5.py:134-178): the feature map and the parameters of its
 * three 1x1 convolutions, and the three planar output tensors. */
typedef struct kodb_head_conv_level {
    const float* x;       /* (B, in_channels, H, W) contiguous fp32 */
    const float* weight;  /* (A*(5+C), in_channels): box rows, objectness rows, class rows -- the three conv weights
                             back to back (yolov5.py:21-26, :104-109, :55-60) */
    const float* bias_box; /* (A*4,) */
    const float* bias_obj; /* (A,)   */
    const float* bias_cls; /* (A*C,) */
    float* out_box;       /* (B, A*4, H, W): viewed (B,A,H,W,4) by Rearrange("b (a p) h w -> b a h w p") */
    float* out_obj;       /* (B, A,   H, W) */
    float* out_cls;       /* (B, A*C, H, W) */
    int32_t in_channels, H, W;
} kodb_head_conv_level;

/* kodb_head_forward: Yolov5Head.forward (kod/nn/heads/yolov5.py:170-178) of up to three pyramid levels in one launch:
 * per level the box / objectness / class 1x1 convolutions as ONE tf32 tensor-core GEMM that reads the feature map once
 * and writes the three planar tensors.  fp32 in, tf32 multiply, fp32 accumulate + bias (the precision class of torch's
 * default convolution policy, torch.backends.cudnn.allow_tf32).  A level with in_channels % 32 != 0, H*W % 4 != 0 (the
 * 13x13 grid of a 416 input), A*(5+C) > 256 or a base address that is not 16-byte aligned runs on a plain SIMT kernel
 * (fp32 multiply-add) inside the same call -- slower per byte, but those levels are small. */
int kodb_head_forward(const kodb_head_conv_level* levels_h /* [num_levels], 1..3 */, int32_t num_levels, int32_t batch,
                      int32_t num_anchors, int32_t num_classes, void* stream);

/* kodb_head_detections: the eval side of the same GEMM -- get_detections(Yolov5Head.forward(x)) (exp.py:70-102 on top of
 * yolov5.py:170-178) in ONE launch: the epilogue decodes the accumulators (sigmoid / grid / anchor, layers.py:55-89) and
 * writes the (B, AC, 5+C) detection rows (rows ll|ml|hl, r = (a*H+y)*W+x; columns x1,y1,x2,y2,sigmoid(obj),sigmoid(cls));
 * the planar logits are never written (out_* of the levels are ignored).  Precision: tf32 logits like kodb_head_forward; the
 * epilogue's sigmoid is ex2.approx / rcp.approx (~1e-6 relative, far below the tf32 error), so probabilities agree with
 * kodb_decode(kodb_head_forward(x)) to ~3e-6 relative and box corners to ~1e-3 px.  level_cfg_h[l] gives stride / anchors / grid of
 * levels_h[l].  Tensor-core path only: KODB_ERR_UNSUPPORTED for a level outside it (see kodb_head_forward; callers then
 * run kodb_head_forward + kodb_decode). */
int kodb_head_detections(const kodb_head_conv_level* levels_h /* [num_levels] */, const kodb_level_cfg* level_cfg_h,
                         int32_t num_levels, int32_t batch, int32_t num_anchors, int32_t num_classes, float* detections,
                         void* stream);

/* kodb_sgd_step: one optimizer step of torch.optim.SGD(momentum, nesterov) over SmartOptimizer's parameter groups
 * (kod/nn/optim/smart.py:11-60) with the per-group lr / momentum of the warm-up updater
 * (kod/lightning/experiments/yv5_baseline/warmup.py:11-58).  params / grads / momentum are flat fp32 arenas holding
 * the groups back to back; seg_begin_h (host, num_groups + 1 float offsets, multiples of 4) delimits them; hyper is a
 * DEVICE array [num_groups][4] = (lr, momentum, weight_decay, dampening), so the launch never synchronises and can
 * be captured into a CUDA graph.  momentum must be zero before the first step (== torch's buf = clone(d_p)).
 * zero_grads != 0 also writes grads = 0 (replaces optimizer.zero_grad()). */
int kodb_sgd_step(float* params, float* grads, float* momentum, const int64_t* seg_begin_h, int32_t num_groups,
                  const float* hyper, int32_t nesterov, int32_t zero_grads, void* stream);

/* ---- eval tail ---------------------------------------------------------------------------------
 * kodb_decode: DefaultYolov5Experiment.get_detections (exp.py:70-102): (B, AC, 5+C) contiguous fp32,
 * rows ll|ml|hl, r = (a*H+y)*W+x, columns x1,y1,x2,y2,sigmoid(obj),sigmoid(cls).  Inputs untouched. */
int kodb_decode(const kodb_head_level* heads_h, const kodb_level_cfg* levels_h, int32_t batch,
                int32_t num_classes, float* detections, void* stream);

/* NMS configuration: arguments + module constants of non_max_suppression (kod/core/nms.py:9-27). */
typedef struct kodb_nms_cfg {
    float conf_thres;   /* compared in fp32 (torch casts the python scalar to the tensor dtype) */
    double iou_thres;   /* torchvision CPU kernel compares the fp32 ratio with this double */
    float max_wh;       /* 4096 */
    int32_t max_det;    /* 300; at most 512 */
    int32_t max_nms;    /* 30000 */
    int32_t num_classes;
    uint32_t class_mask[KODB_MAX_CLASSES / 32]; /* `classes` filter; all ones = None */
    int32_t stage_mask; /* 0 or 3 = both kernels; 1 = candidate kernel only; 2 = select/NMS kernel only
                           (lets a profiler put an event between the two launches of one call) */
} kodb_nms_cfg;

/* Outputs of the NMS stage; dets rows beyond counts[b] are zero. */
typedef struct kodb_nms_out {
    float* dets;        /* (B, max_det, 6): x1,y1,x2,y2,conf,cls  (un-offset candidates, nms.py:73) */
    int64_t* keep_idx;  /* (B, max_det): index into the candidate list handed to torchvision.ops.nms */
    int32_t* counts;    /* (B,) */
    int32_t* num_candidates; /* (B,) candidates before the max_nms cut (diagnostic) */
    int64_t* keep_key;  /* (B, max_det) or NULL: row * num_classes + class of every kept detection, i.e. WHICH
                           (anchor-cell, class) pair of the (B, AC, 5+C) matrix it is (nms.py:46-48); -1 padding.
                           Unlike keep_idx it does not shift when an unrelated candidate flips at the conf edge. */
} kodb_nms_out;

/* Ground truth + outputs of the per-class IoU matching (pycoco_map_eval.py:50-106 -> COCOeval.evaluateImg).
 * If `gt_h` is NULL the matching stage is skipped. */
typedef struct kodb_match_out {
    float* rec_score;  /* (B, max_det) */
    int32_t* rec_cls;  /* (B, max_det) */
    uint8_t* rec_tp;   /* (B, max_det) bit t = true positive at iou_thrs[t] */
    int64_t* gt_count; /* (C,) accumulated (+=) */
    double iou_thrs[KODB_NUM_IOU_THRS];
} kodb_match_out;

/* Eval workspace: zero-filled once after allocation; every call leaves it reusable by later calls WITH THE SAME
 * `batch` (the zero-on-entry counters / histograms are laid out by batch size).  Re-zero it, or use another
 * buffer, before a call with a different batch size. */
int64_t kodb_eval_workspace_bytes(int32_t batch, int64_t rows_per_image, int32_t num_classes);

/* kodb_nms: non_max_suppression(detections, conf_thres, nms_thres, classes) (kod/core/nms.py:9-75) on a
 * decoded (B, N, 5+C) tensor. */
int kodb_nms(const float* detections, int32_t batch, int64_t rows, const kodb_nms_cfg* cfg_h,
             const kodb_nms_out* out_h, const kodb_targets* gt_h, const kodb_match_out* match_h,
             void* workspace, int64_t workspace_bytes, void* stream);

/* kodb_eval_tail: validation_step's get_detections -> non_max_suppression (exp.py:140-154) fused: the
 * (B, AC, 5+C) tensor is never materialised; result == kodb_nms(kodb_decode(heads)). */
int kodb_eval_tail(const kodb_head_level* heads_h, const kodb_level_cfg* levels_h, int32_t batch,
                   const kodb_nms_cfg* cfg_h, const kodb_nms_out* out_h, const kodb_targets* gt_h,
                   const kodb_match_out* match_h, void* workspace, int64_t workspace_bytes, void* stream);

/* kodb_map_match: the matching stage alone on NMS output (used when detections come from elsewhere). */
int kodb_map_match(const float* dets, const int32_t* counts, int32_t batch, int32_t max_det,
                   int32_t num_classes, const kodb_targets* gt_h, const kodb_match_out* match_h,
                   void* stream);

/* kodb_ap_accumulate: COCOeval.accumulate + the evaluator's report (pycoco_map_eval.py:108-125).
 * Records are (num_images, max_det) slots, valid where slot < counts[image]; image order = order of
 * add_predictions.  ap (device, fp64 [KODB_NUM_IOU_THRS][C]); NaN for classes without ground truth. */
int64_t kodb_ap_workspace_bytes(int64_t num_images, int32_t max_det, int32_t num_classes);
int kodb_ap_accumulate(const float* rec_score, const int32_t* rec_cls, const uint8_t* rec_tp,
                       const int32_t* counts, int64_t num_images, int32_t max_det, int32_t num_classes,
                       const int64_t* gt_count, double* ap, void* workspace, int64_t workspace_bytes,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KODB_H_ */
