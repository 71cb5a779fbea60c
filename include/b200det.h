/*
 * b200det.h -- C ABI of libb200det.so: the B200 (sm_100a) detection box-ops hot path.
 *
 * The reference (kostas1515/object_detectors) has no FFI: its boundary for this path is a set
 * of Python call signatures.  Each entry point below names the reference call it replaces
 * (paths relative to the reference root); the Python shims in object_detectors_b200/ keep the
 * reference signatures and call these through ctypes.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.  `stream` is a cudaStream_t passed as
 *    void* (NULL = legacy default stream).  All pointers are DEVICE pointers unless the
 *    parameter name ends in `_host`.
 *  - fp32 tensors are contiguous, int32 / int64 as stated.  Inputs are never written.
 *  - every call is asynchronous on `stream`, allocates nothing and is CUDA-graph capturable
 *    (except the *_host entry points, which synchronise the stream before returning).
 *    Scratch memory comes from the caller: query the size with the matching *_workspace_bytes.
 *  - return value: 0 = B200_OK, negative = error (b200_error_string).  No exceptions cross
 *    the boundary.  Data-dependent conditions (candidate slab overflow) are reported through
 *    a device-side status word the caller reads when convenient.
 *  - there is no CPU fallback: without a CUDA device every compute entry point returns
 *    B200_ERR_CUDA.
 */
#ifndef B200DET_H_
#define B200DET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

#define B200_MAX_SCALES 4
#define B200_MAX_ANCHORS 8
#define B200_MAX_RANKS 16    /* ranks of one NVLink domain served by the one-sided exchange */

enum {
    B200_OK = 0,
    B200_ERR_INVALID = -1,  /* bad argument (null pointer, size out of range, unsupported shape) */
    B200_ERR_CUDA = -2,     /* CUDA runtime error (launch failure, no device) */
    B200_ERR_WORKSPACE = -3 /* workspace too small / misaligned */
};

/* NMS flavours -- one arithmetic order per reference implementation */
enum {
    /* helper.nms_majority (yolo/utilities/helper.py:280-382): class-agnostic, removes
     * !(IoU < thr) with thr compared in fp32, IoU = inter/((area_j-inter)+area_i); kept box is
     * relabelled to the majority class of the boxes it removed with IoU > thr when those hold
     * more than one distinct class. */
    B200_NMS_MAJORITY = 0,
    /* torchvision.ops.nms (call sites yolo/benchmark.py:100, yolo/utilities/telemetry.py:207):
     * class-agnostic, suppress iff (double)(inter/(area_i+area_j-inter)) > thr. */
    B200_NMS_TV = 1,
    /* torchvision.ops.batched_nms, "vanilla" strategy (tvision/rpn.py:272, roi_heads.py:771,
     * retinanet.py:463, ssd.py:423): as B200_NMS_TV but only equal labels interact. */
    B200_NMS_TV_CLASS = 2,
    /* torchvision.ops.batched_nms, "coordinate trick" strategy: boxes are shifted by
     * label*(max_coordinate+1) in fp32 and suppressed class-agnostically. */
    B200_NMS_TV_TRICK = 3,
    /* torchvision.ops.batched_nms as shipped: per segment, the coordinate trick when boxes.numel() (4n) is
     * at most the switch limit, the per-class strategy above it (boxes.py; limit 100 000 on CUDA, 4 000 on
     * CPU; b200_set_batched_nms_auto_limit). */
    B200_NMS_TV_AUTO = 4
};
int b200_set_batched_nms_auto_limit(int64_t numel);

/* pairwise IoU flavours of helper.bbox_iou (yolo/utilities/helper.py:221-277) + torchvision */
enum {
    B200_IOU = 0,     /* helper.bbox_iou iou_type 0 */
    B200_GIOU = 1,    /* iou_type 1 (reference default, hydra/yolo/head.yaml:6) */
    B200_DIOU = 2,    /* iou_type 2 */
    B200_CIOU = 3,    /* iou_type 3 */
    B200_IOU_TV = 4   /* torchvision.ops.box_iou: union = (area1 + area2) - inter */
};

/* Geometry of one multi-scale YOLO head set, i.e. the state YOLOForw.forward rebuilds on every
 * call (yolo/nets/yolo_forw.py:93-119).  Head s is NCHW fp32 [batch, A*(5+C), grid[s], grid[s]];
 * flat anchor index inside a scale is n = (h*W + w)*A + a, scales concatenated in order. */
typedef struct b200_yolo_layout {
    int32_t num_scales;   /* 1..B200_MAX_SCALES */
    int32_t num_anchors;  /* A per scale, 1..B200_MAX_ANCHORS */
    int32_t num_classes;  /* C >= 1 */
    int32_t batch;        /* B >= 1 */
    int32_t softmax;      /* 1: cls = softmax(idf*t) (class_loss==1, default); 0: sigmoid(idf*t) */
    float img_size;       /* YOLOForw.img_size as fp32 */
    int32_t grid[B200_MAX_SCALES];
    /* fp32( fp32(a_px / (img_size/grid)) / grid ) per anchor: the reference's `anchor_w`,
     * `anchor_h` columns of cxypwh (yolo_forw.py:99,108-113), prepared by the host shim with the
     * reference's rounding (python double division, cast, fp32 division). */
    float anchor_rel[B200_MAX_SCALES][B200_MAX_ANCHORS][2];
} b200_yolo_layout;

int b200_abi_version(void);
const char* b200_error_string(int code);
/* number of SMs / device name of the current device (diagnostics for bench.py) */
int b200_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ------------------------------------------------------------------------------------------
 * YOLO decode
 * ---------------------------------------------------------------------------------------- */

/* Replaces YOLOForw.forward(input, targets=None) (yolo/nets/yolo_forw.py:81-119,163-176).
 * out: [B, N, 5+C] fp32, rows xc,yc,w,h (pixels), objectness, class probabilities.
 * idf: [C] fp32 class scale (`idf_logits`) or NULL for 1. */
int b200_yolo_decode_dense(const b200_yolo_layout* layout, const float* const* heads,
                           const float* idf, float* out, void* stream);

size_t b200_yolo_workspace_bytes(const b200_yolo_layout* layout, int32_t capacity);

/* Replaces the decode -> get_abs_coord -> score -> mask -> per-image gather sequence of
 * test_one_epoch (yolo/procedures/test_one_epoch.py:22-28,35) without materialising [B,N,5+C].
 * Per image b, candidates with conf*max_c(cls) > conf_thr in ascending anchor index:
 *   cand_box    [B, capacity, 4] fp32 x1,y1,x2,y2
 *   cand_score  [B, capacity]    fp32
 *   cand_label  [B, capacity]    int32 argmax class (first maximum)
 *   cand_anchor [B, capacity]    int32 flat anchor index n
 *   cand_count  [B]              int32 TRUE number of candidates (may exceed capacity: then
 *                                only the `capacity` lowest-slot ones are stored and status |= 1)
 * status: int32 device word, OR-ed with 1 on overflow (caller zeroes it). */
int b200_yolo_decode_filter(const b200_yolo_layout* layout, const float* const* heads,
                            const float* idf, float conf_thr, int32_t capacity,
                            float* cand_box, float* cand_score, int32_t* cand_label,
                            int32_t* cand_anchor, int32_t* cand_count, int32_t* status,
                            void* workspace, size_t workspace_bytes, void* stream);

/* The whole eval post-process of test_one_epoch.py:22-36 in one call: decode + filter +
 * compaction + NMS.  nms_mode B200_NMS_MAJORITY reproduces helper.nms_majority (the active
 * reference path), B200_NMS_TV / _TV_CLASS the torchvision variants (test_one_epoch.py:30,
 * benchmark.py:94-101).  Outputs per image, kept detections in descending score:
 *   det       [B, max_det, 6] fp32 x1,y1,x2,y2,score,label (label after majority relabel)
 *   det_keep  [B, max_det]    int32 index of the kept box in the image's candidate list
 *                             (ascending-anchor order == the reference's pred_conf order; may be NULL)
 *   det_anchor[B, max_det]    int32 flat anchor index of the kept box (may be NULL)
 *   det_count [B]             int32 number kept (clipped to max_det, status |= 2 if clipped)
 *   cand_count[B]             int32 number of candidates (may be NULL) */
int b200_yolo_postprocess(const b200_yolo_layout* layout, const float* const* heads,
                          const float* idf, float conf_thr, double nms_thr, int32_t nms_mode,
                          int32_t capacity, int32_t max_det, float* det, int32_t* det_keep,
                          int32_t* det_anchor, int32_t* det_count, int32_t* cand_count,
                          int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* The same call split at its one internal dependency, for callers that pipeline batches over several
 * streams (bench.py): phase 1 leaves the unordered candidate slab in `workspace`, phase 2 consumes it.
 * Phase 2 may run on another stream once phase 1 has completed (event); a workspace must not be handed to
 * phase 1 again before its phase 2 has finished.  b200_yolo_postprocess == decode + nms on one stream. */
int b200_yolo_postprocess_decode(const b200_yolo_layout* layout, const float* const* heads,
                                 const float* idf, float conf_thr, int32_t capacity, int32_t* status,
                                 void* workspace, size_t workspace_bytes, void* stream);
int b200_yolo_postprocess_nms(const b200_yolo_layout* layout, double nms_thr, int32_t nms_mode,
                              int32_t capacity, int32_t max_det, float* det, int32_t* det_keep,
                              int32_t* det_anchor, int32_t* det_count, int32_t* cand_count,
                              int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* End-to-end variant with HOST buffers: copies the heads host->device (pinned or pageable),
 * runs b200_yolo_postprocess and copies det / det_count back.  Device staging comes from an
 * internal pool sized on first use; synchronises before returning.
 * heads_host[s] : [B, A*(5+C), grid, grid] fp32 on the host
 * det_host      : [B, max_det, 6], det_count_host: [B] */
int b200_yolo_postprocess_host(const b200_yolo_layout* layout, const float* const* heads_host,
                               const float* idf_host, float conf_thr, double nms_thr,
                               int32_t nms_mode, int32_t capacity, int32_t max_det,
                               float* det_host, int32_t* det_keep_host, int32_t* det_count_host,
                               int32_t* status_host);

/* Profiling hook (bench.py): two cudaEvent_t (as void*, NULL to disable) that the next
 * b200_yolo_postprocess / b200_yolo_decode_filter calls record on their stream immediately
 * before and after the fused decode+filter kernel. */
int b200_debug_set_decode_events(void* ev_begin, void* ev_end);

/* Profiling hook: three more cudaEvent_t (NULL to disable) recorded by the next NMS launches after the
 * plan, pairs and resolve kernels; with b200_debug_set_decode_events this gives a per-step timeline. */
int b200_debug_set_timeline(void* after_plan, void* after_pairs, void* after_resolve);
/* Profiling hook: device buffer int64[segments, 16] that the resolve kernel fills with clock64 stamps at its
 * phase boundaries (A stage, B fixed point, C vote, D order, E emit, end) + n and K; NULL to disable. */
int b200_debug_set_resolve_prof(void* buf);
/* Same for the per-level select of the RPN / RetinaNet filter: int64[levels * batch, 8] globaltimer stamps of each
 * level's CTA (start, -, staged, selected, sorted, done) + level and slice; NULL to disable. */
int b200_debug_set_rpn_prof(void* buf);
/* Tuning hook: segment size above which a SERIAL post-process call (b200_yolo_postprocess: decode and NMS of one batch
 * on one stream) hands a segment to the 1024-thread single-launch NMS kernel instead of the three-launch path
 * (default and maximum 1500; lower values measured worse on the C2 workload: both paths then run back to back;
 * pipelined calls always split at 1500 and use the 256-thread instantiation). */
int b200_debug_set_serial_split(int boxes);
/* NMS kernel path (process-wide): 1 = the general three-launch path (plan / pairs / resolve: spatially pruned tile
 * pairs, small CTAs that co-reside with the streaming decode kernel), 0 = segments of <= 4096 boxes take the
 * single-launch path (nms_fused.cu: no work queue, no cross-kernel dependencies), -1 (default) = by workload:
 * candidate slabs of the YOLO post-process -> general for segments above 1500 boxes plus a 256-thread single-launch
 * kernel for the rest, array inputs (nms / batched_nms, RPN, ROI heads) -> single-launch, 2 = the 256-thread
 * single-launch kernel for every segment (measurement only).  All produce identical results (the tests run the NMS
 * cases through the general and the single-launch path). */
int b200_debug_set_nms_path(int general);
/* Tuning hook: launch shape of the NMS resolve CTAs (threads: multiple of 32 in 64..1024, dynamic shared memory
 * in KB 16..200; out-of-range values keep the current setting).  Default 1024 threads, 112 KB. */
int b200_debug_set_resolve(int threads, int smem_kb);

/* Variants of the fused decode+filter kernel.  All produce identical candidates; they differ in how the
 * head tensors are fetched (process-wide setting, default B200_DECODE_RING):
 *   RING    persistent CTAs; every warp streams 64-cell tiles through its own shared-memory stages with
 *           2-D tensor-map TMA (cp.async.bulk.tensor) + mbarrier and sweeps only the cells whose
 *           objectness can still pass.  Every byte of the head tensors is read exactly once, whatever
 *           the input looks like: the variant the HBM roofline fraction is quoted for.
 *   STREAM  register path; every byte is read once with coalesced 128-bit loads.
 *   GATED   register path; the objectness plane is read for every cell, class and box planes only by
 *           lanes that hold a cell whose objectness can still pass the threshold (score <= conf).
 *           DRAM traffic is input dependent (176 MB of the 495 MB batch on the benchmark input). */
enum { B200_DECODE_GATED = 0, B200_DECODE_STREAM = 1, B200_DECODE_RING = 3 };   /* 2: retired prototype */
int b200_set_decode_variant(int variant);
/* Tuning hook of the RING variant (values <= 0 keep the current setting): warps per CTA (1..8),
 * shared-memory stages per warp (warps x stages <= 32), persistent CTAs per SM (1..4). */
int b200_debug_set_ring(int warps, int stages_per_warp, int ctas_per_sm);

/* Replaces the inference branch of the legacy per-head layer YOLOLoss.forward(input, targets=None)
 * (yolo/nets/yolo_loss.py:34-105; callers yolo/benchmark.py:63, telemetry.py:46-92).
 * head: [B, A*(5+C), in_h, in_w] fp32; out: [B, A*in_h*in_w, 5+C] with rows ordered (a, h, w);
 * stride_* = fp32(img_size / in_*); anchors_scaled: DEVICE [A][2] fp32 = anchor_px / stride (:40). */
int b200_yolo_legacy_decode(const float* head, int32_t batch, int32_t num_anchors, int32_t num_classes,
                            int32_t in_h, int32_t in_w, float stride_w, float stride_h,
                            const float* anchors_scaled, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * ROI-head box post-process (Faster R-CNN family)
 * ---------------------------------------------------------------------------------------- */

/* Replaces RoIHeads.postprocess_detections (torchvision_models/tvision/roi_heads.py:715-781) for a batch:
 * scores = softmax (activation 0) / gombit (1) / sigmoid (2) of tfidf*class_logits, per-class box decode
 * with BoxCoder weights `weights_host` (x,y,w,h), clip to the image, background column dropped,
 * score > score_thr, w,h >= min_size, batched_nms per class (nms_mode B200_NMS_TV_CLASS or _TV_TRICK),
 * first max_det by score.
 *   class_logits [R, C], box_regression [R, 4C], proposals [R, 4] (images concatenated),
 *   row_offsets [B+1] device int32, image_hw [B, 2] device fp32 (height, width), tfidf [C] or NULL.
 * Outputs per image: det [B, max_det, 6] = x1,y1,x2,y2,score,label; det_keep [B, max_det] (index into the
 * image's filtered candidate list, may be NULL); det_count [B]; cand_count [B] (may be NULL).
 * capacity: candidate slab rows per image (status |= 1 on overflow). */
size_t b200_roi_workspace_bytes(int32_t batch, int32_t capacity);
int b200_roi_postprocess(const float* class_logits, const float* box_regression, const float* proposals,
                         const int32_t* row_offsets, int32_t batch, int32_t total_rows, int32_t num_classes,
                         const float* image_hw, const float* tfidf, int32_t activation,
                         const float* weights_host, float xform_clip, float score_thr, float min_size,
                         double nms_thr, int32_t nms_mode, int32_t capacity, int32_t max_det, float* det,
                         int32_t* det_keep, int32_t* det_count, int32_t* cand_count, int32_t* status,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * NMS on caller-provided boxes, batched over segments (images, or image x level)
 * ---------------------------------------------------------------------------------------- */

/* max_segment: host-known upper bound of any segment length (0 = total_boxes).  The scratch holds a
 * suppression bitmask of num_segments * max_segment * ceil(max_segment/64) * 8 bytes. */
size_t b200_nms_workspace_bytes(int64_t total_boxes, int32_t num_segments, int32_t max_segment);

/* Replaces helper.nms_majority (helper.py:280), torchvision.ops.nms / batched_nms.
 *   boxes  [T,4] fp32 xyxy; scores [T] fp32; labels [T] int32 (required unless mode==TV)
 *   seg_offsets [S+1] int32: segment s owns rows seg_offsets[s] .. seg_offsets[s+1]
 *   keep      [T] int64: for segment s the kept row indices RELATIVE to the segment start, in
 *             descending score (ties: lower index first), written at keep[seg_offsets[s] + k]
 *   keep_count[S] int32
 *   labels_out[T] int32 or NULL: (MAJORITY) label of the k-th kept box after relabelling,
 *             written at labels_out[seg_offsets[s] + k]
 *   iou_thr is a double: MAJORITY rounds it to fp32 like the reference's tensor compare,
 *   the TV modes compare (double)iou > iou_thr like torchvision. */
int b200_nms(const float* boxes, const float* scores, const int32_t* labels,
             const int32_t* seg_offsets, int32_t num_segments, int64_t total_boxes,
             int32_t max_segment, double iou_thr, int32_t mode, int64_t* keep, int32_t* keep_count, int32_t* labels_out,
             void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * pairwise IoU and IoU-based target matching
 * ---------------------------------------------------------------------------------------- */

/* Replaces helper.bbox_iou(bb1[M,1,4], bb2[1,N,4], iou_type, xcycwh) (helper.py:221-277) and
 * torchvision.ops.box_iou (kind B200_IOU_TV, xcycwh must be 0).  out: [M,N] fp32. */
int b200_box_iou(const float* boxes1, int32_t m, const float* boxes2, int32_t n, int32_t kind,
                 int32_t xcycwh, float* out, void* stream);

/* element-wise (paired) version: boxes1[K,4] vs boxes2[K,4] -> out[K] (yolo_forw.py:125 shape) */
int b200_box_iou_paired(const float* boxes1, const float* boxes2, int32_t k, int32_t kind,
                        int32_t xcycwh, float* out, void* stream);
/* Backward of b200_box_iou_paired for kinds IoU..CIoU (the loss-side call yolo_forw.py:125,143-146 runs under
 * autograd): grad_boxes{1,2}[k,4] = grad_out[k] * d out / d boxes{1,2}, matching torch autograd on the reference's
 * expression (ties of min/max split evenly, clamp passes at 0, CIoU alpha constant).  Either output may be NULL. */
int b200_box_iou_paired_backward(const float* boxes1, const float* boxes2, const float* grad_out, int32_t k,
                                 int32_t kind, int32_t xcycwh, float* grad_boxes1, float* grad_boxes2, void* stream);

/* Replaces the IoU + reductions of YOLOForw.get_target (yolo/nets/yolo_forw.py:183-201) for a
 * whole batch: gt [B, max_gt, 4] fp32 relative xc,yc,w,h (rows >= gt_count[b] ignored),
 * anchors [N,4] = cxypwh.  Outputs: best_anchor [B, max_gt] int64 (first argmax_n of the IoU
 * row), noobj [B,N] uint8 = all_m(iou < ignore_thr) with matched anchors cleared.
 * workspace: b200_iou_match_workspace_bytes(B, max_gt). */
size_t b200_iou_match_workspace_bytes(int32_t batch, int32_t max_gt);
int b200_iou_match(const float* gt, const int32_t* gt_count, int32_t batch, int32_t max_gt,
                   const float* anchors, int32_t n, int32_t kind, float ignore_thr,
                   int64_t* best_anchor, uint8_t* noobj, void* workspace, size_t workspace_bytes,
                   void* stream);

/* ------------------------------------------------------------------------------------------
 * RPN proposal filter
 * ---------------------------------------------------------------------------------------- */

size_t b200_rpn_workspace_bytes(int32_t batch, int32_t total_anchors, int32_t num_levels,
                                int32_t pre_nms_top_n);

/* Replaces BoxCoder.decode + RegionProposalNetwork.filter_proposals
 * (torchvision_models/tvision/rpn.py:215-280,355; _utils.py:186-223): per-level top-k on raw
 * objectness, decode of the selected anchors only, sigmoid, clip, small-box and score filters,
 * per-level NMS, first post_nms_top_n by score.  nms_mode selects which torchvision
 * batched_nms strategy is reproduced: B200_NMS_TV_CLASS ("vanilla", what torchvision runs for
 * > 4000 coordinates on CPU / > 100000 on CUDA) or B200_NMS_TV_TRICK (coordinate trick).
 *   objectness [B, total] fp32, deltas [B, total, 4] fp32, anchors [total, 4] fp32 xyxy,
 *   level_sizes [L] int32 (host), image_hw [B,2] fp32 (device; h,w)
 *   out_boxes [B, post_nms_top_n, 4], out_scores [B, post_nms_top_n], out_index [B, post] int32
 *   (flat anchor index, may be NULL), out_count [B] int32 */
int b200_rpn_filter(const float* objectness, const float* deltas, const float* anchors,
                    int32_t batch, int32_t total_anchors, const int32_t* level_sizes_host,
                    int32_t num_levels, const float* image_hw, int32_t pre_nms_top_n,
                    int32_t post_nms_top_n, double nms_thr, float score_thr, float min_size,
                    int32_t nms_mode, float* out_boxes, float* out_scores, int32_t* out_index, int32_t* out_count,
                    void* workspace, size_t workspace_bytes, void* stream);

/* RegionProposalNetwork._get_top_n_idx(objectness, num_anchors_per_level) (rpn.py:215-228): per image and level the
 * indices of the min(pre_nms_top_n, n_l) largest raw objectness logits, in descending order (Tensor.topk order for
 * tie-free input; equal logits: lower index first), offset by the level start and concatenated over the levels:
 * out_index [B, sum_l min(pre_nms_top_n, n_l)] int64.  workspace: b200_rpn_top_n_idx_workspace_bytes, 256 B aligned. */
size_t b200_rpn_top_n_idx_workspace_bytes(int32_t batch, int32_t total_anchors, int32_t num_levels, int32_t pre_nms_top_n);
int b200_rpn_top_n_idx(const float* objectness, int32_t batch, int32_t total_anchors, const int32_t* level_sizes_host,
                       int32_t num_levels, int32_t pre_nms_top_n, int64_t* out_index, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Same filter on boxes that are already decoded: the exact signature-level replacement of
 * RegionProposalNetwork.filter_proposals(proposals, objectness, image_shapes, num_anchors_per_level)
 * (rpn.py:230).  proposals [B, total, 4] fp32 xyxy. */
int b200_rpn_filter_proposals(const float* objectness, const float* proposals, int32_t batch,
                              int32_t total_anchors, const int32_t* level_sizes_host, int32_t num_levels,
                              const float* image_hw, int32_t pre_nms_top_n, int32_t post_nms_top_n,
                              double nms_thr, float score_thr, float min_size, int32_t nms_mode,
                              float* out_boxes, float* out_scores, int32_t* out_index, int32_t* out_count,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * element-wise pieces of the drop-in surface
 * ---------------------------------------------------------------------------------------- */

/* helper.get_abs_coord (yolo/utilities/helper.py:203-217): [n,4] xc,yc,w,h -> x1,y1,x2,y2. */
int b200_abs_coord(const float* boxes, int64_t n, float* out, void* stream);

/* BoxCoder.decode_single (torchvision_models/tvision/_utils.py:186-223): rel_codes [n, 4*k],
 * boxes [n,4] -> out [n, 4*k]; weights_host = (wx, wy, ww, wh); xform_clip = log(1000/16). */
int b200_boxcoder_decode(const float* rel_codes, const float* boxes, int64_t n, int32_t boxes_per_row,
                         const float* weights_host, float xform_clip, float* out, void* stream);
/* Replaces encode_boxes / BoxCoder.encode_single (torchvision_models/tvision/_utils.py:80-125,160-166):
 * out[n,4] = (wx*(gx-ex)/ew, wy*(gy-ey)/eh, ww*log(gw/ew), wh*log(gh/eh)) for matched pairs
 * reference_boxes[n,4], proposals[n,4] (xyxy).  weights_host: 4 host floats (x, y, w, h). */
int b200_boxcoder_encode(const float* reference_boxes, const float* proposals, int64_t n, const float* weights_host,
                         float* out, void* stream);

/* Matcher.__call__ (tvision/_utils.py:271-344) on a dense [M,N] quality matrix: matches [N] int64 =
 * argmax over M (first maximum), -1 below low_thr, -2 between the thresholds; with
 * allow_low_quality every prediction tying a ground truth's best quality gets its argmax back.
 * workspace: N * 8 bytes (only read when allow_low_quality). */
int b200_matcher(const float* quality, int32_t m, int32_t n, float high_thr, float low_thr,
                 int32_t allow_low_quality, int64_t* matches, void* workspace, size_t workspace_bytes,
                 void* stream);

/* RetinaNet.postprocess_detections (retinanet.py:414-472) for a whole batch: per image and level the
 * min(topk_candidates, .) best of sigmoid(tfidf[c] * logit) > score_thr over the flattened [anchors_l, C] scores (one
 * sliced select per level, only the survivors are decoded with BoxCoder(1,1,1,1) and clipped), then ONE class-aware
 * batched_nms per image and the first detections_per_img by score.
 *   cls_logits [B, sumA, C], bbox_regression [B, sumA, 4], anchors [sumA, 4] (shared by the images),
 *   level_anchors_host [L] anchors per level, tfidf [C] or NULL, image_hw [B,2];
 *   out_boxes [B, D, 4], out_scores [B, D], out_labels [B, D] int32 (class index, 0-based as the reference), out_count [B].
 * nms_mode: B200_NMS_TV_CLASS / _TV_TRICK / _TV_AUTO (torchvision's batched_nms strategies). */
size_t b200_retinanet_workspace_bytes(int32_t batch, int32_t total_anchors, int32_t num_classes, int32_t num_levels,
                                      int32_t topk_candidates);
int b200_retinanet_postprocess(const float* cls_logits, const float* bbox_regression, const float* anchors, int32_t batch,
                               int32_t total_anchors, int32_t num_classes, const int32_t* level_anchors_host,
                               int32_t num_levels, const float* tfidf, const float* image_hw, int32_t topk_candidates,
                               float score_thr, double nms_thr, int32_t nms_mode, int32_t detections_per_img,
                               float* out_boxes, float* out_scores, int32_t* out_labels, int32_t* out_count,
                               void* workspace, size_t workspace_bytes, void* stream);

/* SSD.postprocess_detections (ssd.py:386-430) for a whole batch: scores = softmax(tfidf * cls_logits) over all classes,
 * ONE box per anchor (BoxCoder.decode_single with `weights_host`, clipped to the image), per foreground class the
 * candidates with score > score_thr capped at topk_per_class by score (ssd.py:404-409), class-aware batched_nms, the
 * first max_det by score.  cls_logits [R, C], bbox_regression [R, 4], anchors [R, 4] (rows of image b =
 * [row_offsets[b], row_offsets[b+1])); det [B, max_det, 6] = x1,y1,x2,y2,score,label; workspace / capacity / status as
 * b200_roi_postprocess (overflow of the candidate slab is reported in status bit 0, cand_count holds the true counts). */
int b200_ssd_postprocess(const float* cls_logits, const float* bbox_regression, const float* anchors,
                         const int32_t* row_offsets, int32_t batch, int32_t total_rows, int32_t num_classes,
                         const float* image_hw, const float* tfidf, const float* weights_host, float xform_clip,
                         float score_thr, int32_t topk_per_class, double nms_thr, int32_t nms_mode, int32_t capacity,
                         int32_t max_det, float* det, int32_t* det_count, int32_t* cand_count, int32_t* status,
                         void* workspace, size_t workspace_bytes, void* stream);

/* box_iou + Matcher fused (rpn.py:192-193, roi_heads.py:633-634, retinanet.py:409-410, ssd.py:371-372):
 *   matches = Matcher(high, low, allow_low_quality)(box_iou(gt_boxes, boxes))        (_utils.py:271-344)
 * without ever writing the [M, N] quality matrix; bit-identical to that composition (torchvision's IoU arithmetic,
 * first maximum on ties, BELOW_LOW = -1 / BETWEEN = -2, low-quality restore incl. ties).  ssd != 0 adds SSDMatcher's
 * override matches[argmax_n q[m, :]] = m (:347-361; later m wins a shared box).  matched_vals [N] (nullable) receives
 * q.max(dim=0).  gt_boxes [M,4], boxes [N,4] xyxy fp32, 16 B aligned; matches [N] int64. */
size_t b200_match_boxes_workspace_bytes(int32_t m, int32_t n);
int b200_match_boxes(const float* gt_boxes, int32_t m, const float* boxes, int32_t n, float high_thr, float low_thr,
                     int32_t allow_low_quality, int32_t ssd, int64_t* matches, float* matched_vals, void* workspace,
                     size_t workspace_bytes, void* stream);

/* SSDMatcher.__call__(match_quality_matrix) on a materialised matrix = b200_matcher(high = low = threshold, no
 * low-quality restore) followed by this override (_utils.py:355-359).  workspace: 8 bytes per ground truth. */
int b200_matcher_ssd_override(const float* quality, int32_t m, int32_t n, int64_t* matches, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * result emission (the step right after the path)
 * ---------------------------------------------------------------------------------------- */

/* yolo/procedures/test_one_epoch.py:41-66 + yolo/utilities/helper.py:16-24 for a whole batch in one kernel:
 * det [B,max_det,6] / det_count [B] (b200_yolo_postprocess) -> packed records in image order,
 *   records [sum K, 6] = x, y, w, h (pixels of the original image: coord / inp_dim * size), area = w*h, score
 *   category [sum K]   = class_map[label] (COCO: the 80 -> 91 table) or label + 1 when class_map == NULL
 *   image    [sum K]   = image_id of the owning image;   total[0] = sum K
 * img_hw [B,2] = (height, width) of the original images, image_id [B] int64.  strict_reference != 0 reproduces the
 * reference's list semantics: images without detections are dropped before the list is matched with `targets` by
 * position, so the k-th non-empty image takes size and id of targets[k].  Outputs sized for B*max_det rows. */
int b200_emit_results(const float* det, const int32_t* det_count, int32_t batch, int32_t max_det,
                      const float* img_hw, const int64_t* image_id, float inp_dim, const int32_t* class_map,
                      int32_t num_map, int32_t strict_reference, float* records, int32_t* category,
                      int64_t* image, int32_t* total, void* stream);

/* torchvision.ops.boxes.clip_boxes_to_image (reference tvision/boxes.py; callers rpn.py:260, roi_heads.py:746,
 * retinanet.py:452, ssd.py:397): boxes [n, 4] xyxy (16-byte aligned), x clamped to [0, width], y to [0, height];
 * out may alias boxes. */
int b200_clip_boxes_to_image(const float* boxes, int64_t n, float height, float width, float* out, void* stream);

/* torchvision.ops.boxes.remove_small_boxes (callers rpn.py:263, roi_heads.py:767):
 * keep [<= n] int64 = ascending indices of the boxes with (x2 - x1) >= min_size and (y2 - y1) >= min_size,
 * count[0] = how many. */
int b200_remove_small_boxes(const float* boxes, int32_t n, float min_size, int64_t* keep, int32_t* count, void* stream);

/* ------------------------------------------------------------------------------------------
 * the path's one exchange step: all-gather of the variable-length kept-detection lists
 * (replaces the per-rank pickle files + barrier of yolo/procedures/eval_results.py:12-31 /
 * yolo/main.py:102-105 and `utils.all_gather`, torchvision_models/detection/utils.py:75-115)
 * ---------------------------------------------------------------------------------------- */

/* Message of one rank and step: [B * (1 + max_det*6)] fp32, per image the count (int bit pattern) followed by
 * the kept rows [x1,y1,x2,y2,score,label] in descending score. */

/* Packs [B,max_det,6] detections + counts into one contiguous fixed-capacity message (rows past the count are
 * zero), the send buffer of an all-gather. */
int b200_pack_detections(const float* det, const int32_t* det_count, int32_t batch,
                         int32_t max_det, float* message, void* stream);

/* NCCL form of the exchange: pack into `message`, then ncclAllGather(message -> gathered[world][message]) on
 * `stream`.  `nccl_comm` is the caller's ncclComm_t; ncclAllGather is resolved at run time from the libnccl.so.2
 * the process has loaded.  Rank-major result, identical on every rank. */
int b200_allgather_dets(const float* det, const int32_t* det_count, int32_t batch, int32_t max_det,
                        float* message, float* gathered, void* nccl_comm, void* stream);

/* One-sided form (NVLink / NVSwitch peer memory, one process per GPU on one node): every rank owns a device
 * buffer data[slots][world][message] that its peers map through CUDA IPC.  `push` stores this rank's kept lists of
 * its next step straight into every peer's buffer (only the counts and the rows that exist travel) and releases a
 * per-(slot, rank) flag there; `wait` completes on the stream when all ranks' messages of the next un-waited step
 * have arrived and acknowledges the slot to the peers (flow control: a rank can run at most `slots` steps ahead of
 * the slowest consumer).  No collective kernel, no rendezvous between pushes; step numbers live on the device, so
 * both calls are CUDA-graph capturable.  Pushes of one rank must be stream-ordered among themselves, and so must
 * waits.  Setup: create -> handle (64 bytes, exchange them by any means) -> connect(peer, handle) for every peer. */
typedef struct b200_exchange b200_exchange;
int b200_exchange_create(int32_t rank, int32_t world, int32_t batch, int32_t max_det, int32_t slots,
                         b200_exchange** out);
int b200_exchange_handle(b200_exchange* x, void* handle64);
int b200_exchange_connect(b200_exchange* x, int32_t peer, const void* handle64);
int b200_exchange_push(b200_exchange* x, const float* det, const int32_t* det_count, void* stream);
int b200_exchange_wait(b200_exchange* x, void* stream);
/* device pointer to the message rank `src_rank` pushed for `step`: valid from the completion of that step's wait
 * until the NEXT b200_exchange_wait runs on the stream (readers enqueue their work between the two waits) */
const float* b200_exchange_message(b200_exchange* x, int64_t step, int32_t src_rank);
/* copies the gathered messages of `step` ([world][message], rank-major) into a caller buffer on `stream` */
int b200_exchange_read(b200_exchange* x, int64_t step, float* gathered, void* stream);
/* host-synchronous: steps pushed / waited so far */
int b200_exchange_steps(b200_exchange* x, int64_t* pushed, int64_t* waited);
int b200_exchange_destroy(b200_exchange* x);

#ifdef __cplusplus
}
#endif
#endif /* B200DET_H_ */
