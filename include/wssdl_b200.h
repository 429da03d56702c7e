/*
 * wssdl_b200.h -- C ABI of libwssdl_b200.so: the B200 (sm_100a) implementation of the
 * detector hot path of syshin1014/wssdl_bus.
 *
 * Every entry point replaces one native/Cython interface of the reference (cited per
 * function, paths relative to the reference's code/lib).  Conventions for all DEVICE
 * entry points:
 *   - pointers are device pointers owned by the caller (no hidden allocation);
 *   - work is enqueued on `stream` and the call returns without synchronising;
 *   - the return value is WSSDL_OK or a negative WSSDL_E* code / positive cudaError_t;
 *     nothing ever calls exit() (the reference does: roi_pooling_op_gpu.cu.cc:102-107);
 *   - scratch memory is passed in as `workspace` sized by the matching *_workspace_bytes.
 * The *_host entry points take HOST pointers, are synchronous and manage their own
 * device scratch, like the reference's `_nms` (nms/gpu_nms.hpp:1-2).
 *
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef WSSDL_B200_H_
#define WSSDL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cudaStream_t without including cuda_runtime.h in client code */
typedef struct CUstream_st* wssdl_stream_t;

enum {
  WSSDL_OK = 0,
  WSSDL_EINVAL = -1,     /* bad argument (negative size, null pointer, ...)            */
  WSSDL_EWORKSPACE = -2, /* workspace too small                                       */
  WSSDL_EALIGN = -3,     /* pointer not aligned as documented                         */
  WSSDL_ELIMIT = -4,     /* size beyond what the kernels support                      */
  WSSDL_EZERODIV = -5    /* host NMS: a box pair with zero union (reference raises    */
                         /* ZeroDivisionError, nms/cpu_nms.c:2480-2483)               */
};

int wssdl_version(void);                    /* 100 * major + minor */
const char* wssdl_error_string(int code);   /* static string, also for cudaError_t codes */

/* Process-wide tuning switches, for experiments and for tests that must reach one particular
 * kernel.  Their defaults are read from the environment ONCE, at the first call into the
 * library (the variable named beside each key); no entry point reads the environment per
 * launch.  Results never depend on them: every kernel variant produces the same bytes.
 * wssdl_set_tuning returns WSSDL_OK or WSSDL_EINVAL (unknown key). */
enum {
  WSSDL_TUNE_ROI_FWD_KERNEL = 0,    /* 0 by shape, 1 direct, 2 tiled, 3 band, 4 sorted bins
                                       (WSSDL_ROI_FWD_KERNEL=direct|tiled|band|sorted)          */
  WSSDL_TUNE_ROI_FWD_SLICES = 1,    /* sorted bins: 32-channel slices per CTA, 0 = by shape
                                       (WSSDL_ROI_FWD_SLICES)                                    */
  WSSDL_TUNE_ROI_FWD_CHUNKS = 2,    /* sorted bins: RoI chunks per image, 0 = by shape
                                       (WSSDL_ROI_FWD_CHUNKS)                                    */
  WSSDL_TUNE_NMS_SWEEP_CLUSTER = 3, /* -1 by size, 0 single-CTA sweep, 1 cluster sweep
                                       (WSSDL_NMS_SWEEP_CLUSTER)                                 */
  WSSDL_TUNE_PROPOSALS_CLUSTER = 4, /* -1 (or 1) by shape: the largest cluster of 2 / 4 / 8 CTAs
                                       per image whose B clusters are resident at once, else one
                                       CTA per image; 0 one CTA per image; 2, 4, 8 that size
                                       (WSSDL_PROPOSALS_CLUSTER)                                 */
  WSSDL_TUNE_ROI_FWD_THREADS = 5,   /* sorted bins: threads per pooling CTA, 0 = default,
                                       1024 (one CTA per SM) or 512 (two)  (WSSDL_ROI_FWD_THREADS) */
  WSSDL_TUNE_PDL = 6,               /* programmatic dependent launch between the kernels of one
                                       call (pre-pass -> pooling, proposals -> pre-pass): 1 on
                                       (default), 0 off */
  WSSDL_TUNE_ROI_FWD_BALANCED = 7,  /* sorted bins: row bands that own equal shares of the map's
                                       rows (taller bands): -1 when the grid is long (default),
                                       0 shortest bands, 1 on                                    */
  WSSDL_TUNE_COUNT = 8
};
int wssdl_set_tuning(int key, int value);
int wssdl_get_tuning(int key);

/* ---------------------------------------------------------------- RoI max pooling
 * Replaces ROIPoolForwardLaucher / ROIPoolBackwardLaucher
 * (roi_pooling_layer/roi_pooling_op_gpu.h:18-27) and the CPU op bodies
 * RoiPoolOp<CPUDevice>::Compute / RoiPoolGradOp<CPUDevice>::Compute
 * (roi_pooling_layer/roi_pooling_op.cc:88-204, :333-466).
 *
 * bottom      [B,H,W,C] f32 NHWC        rois   [R,5] f32 (batch, x1, y1, x2, y2) image px
 * top, argmax [R,PH,PW,C] f32 / i32;    argmax = per-image flat index (h*W+w)*C+c, -1 if
 *             the bin is empty; argmax may be NULL (roi_pooling_op_gpu.cu.cc:82-83).
 * bin_mode    WSSDL_BIN_CPU_TRUNC reproduces the CPU op (roi_pooling_op.cc:167-170, the
 *             parity target: int cast BEFORE floor/ceil, bins never overlap);
 *             WSSDL_BIN_GPU_CEIL reproduces the CUDA op (roi_pooling_op_gpu.cu.cc:51-58).
 * A RoI whose batch index is outside [0,B) yields top=0/argmax=-1 (the reference reads
 * out of bounds).  16-byte aligned pointers and C%4==0 select the vectorised kernels;
 * anything else runs the scalar variant of the direct kernel.
 * workspace   optional scratch of wssdl_roi_pool_fwd_workspace_bytes(B, R, PH, PW) bytes
 *             (device, 16-byte aligned; may be NULL).  It holds the RoIs grouped by image
 *             (batches of more than 4096 RoIs: two small kernels on the same stream) and the
 *             sorted-bins kernel's class-sorted bin records (8 bytes per output bin, written
 *             by its sort pre-pass on the same stream).  Without it small batches run the
 *             band / tiled kernel and big ones the direct kernel.  Results are identical
 *             either way.
 * Four kernels give the same bytes and are picked by shape (csrc/roi_pool.cu,
 * csrc/roi_pool_bins.cu): "sorted" (32-channel slice of overlapping row bands of the map
 * staged in shared memory by one TMA tensor copy; the bins a CTA owns are counting-sorted by
 * size class so that every warp runs uniform loops; the detector's 7x7 shapes), "direct" (one
 * CTA per output row reading L2; big bins), "tiled" (16-channel slice of the whole map;
 * C % 32 != 0), "band" (the sorted kernel's predecessor: same staging, one thread per column
 * of bins; kept as the cross-check).  WSSDL_TUNE_ROI_FWD_KERNEL forces one where the shape
 * allows.
 */
enum { WSSDL_BIN_CPU_TRUNC = 0, WSSDL_BIN_GPU_CEIL = 1 };

size_t wssdl_roi_pool_fwd_workspace_bytes(int B, int R, int PH, int PW);

/* Host-only query (no CUDA call): which forward kernel a call of this shape takes and its
 * launch geometry, for tests and tuning.  force: 0 = by shape, 1 = direct, 2 = tiled,
 * 3 = band, 4 = sorted bins.  `out` holds 10 ints:
 *   out[0] kernel (0 direct, 1 tiled, 2 band, 3 sorted bins)
 *   out[1] row bands per image          out[2] rows per band
 *   out[3] first-row distance of two bands
 *   out[4] ranges per band (band / sorted: CTAs per band and slice) or RoI chunks per image (tiled)
 *   out[5] RoIs whose bin edges are resident at a time (sorted bins: in the sort pre-pass)
 *   out[6] dynamic shared memory bytes of the pooling kernel
 *   out[7] 1 if the RoI lists are built in-kernel (no counting-sort launches)
 *   out[8] sorted bins: 32-channel slices one CTA pools one after the other
 *   out[9] sorted bins: threads per pooling CTA
 * Assumes aligned pointers. */
int wssdl_roi_pool_fwd_plan(int B, int H, int W, int C, int R, int PH, int PW,
                            int with_workspace, int force, int* out);

int wssdl_roi_pool_fwd(const float* bottom, const float* rois, int B, int H, int W, int C,
                       int R, int PH, int PW, float spatial_scale, int bin_mode,
                       float* top, int* argmax, void* workspace, size_t workspace_bytes,
                       wssdl_stream_t stream);

/* bwd_mode WSSDL_BWD_ATOMIC: zero-fill + scatter through argmax with fp32 atomics
 *            (fast; summation order free => equal to the reference within 1e-5 rel;
 *            global fp32 atomics flush subnormals);
 *          WSSDL_BWD_GATHER: deterministic gather, one CTA per input cell, additions in
 *            the reference's order (roi, ph, pw ascending) => bit-exact.
 * Both apply the reference's in-RoI and feasible-bin tests (roi_pooling_op.cc:415-445), so
 * the result equals the reference's for ANY argmax tensor, not only one produced by the
 * forward pass (e.g. malformed RoIs with x2<x1 contribute nothing).
 */
enum { WSSDL_BWD_ATOMIC = 0, WSSDL_BWD_GATHER = 1 };

int wssdl_roi_pool_bwd(const float* top_diff, const int* argmax, const float* rois, int B,
                       int H, int W, int C, int R, int PH, int PW, float spatial_scale,
                       int bwd_mode, float* bottom_diff, wssdl_stream_t stream);

/* ---------------------------------------------------------------- greedy NMS
 * Replaces cpu_nms (nms/cpu_nms.pyx:17-68), utils.cython_nms.nms / nms_new
 * (utils/nms.pyx:17-68, :70-123), gpu_nms + _nms (nms/gpu_nms.pyx:16-31,
 * nms/nms_kernel.cu:91-144) and py_cpu_nms (nms/py_cpu_nms.py:10-38).
 *
 * dets [N,dets_stride] f32 rows (x1,y1,x2,y2,score,...), dets_stride >= 5.
 * Sorts on the device by (score desc, index desc) -- the order of
 * `scores.argsort(kind='stable')[::-1]`; the reference uses numpy's default unstable
 * argsort, so tie order is only defined for unique scores -- then runs the 64-bit bitmask
 * pass over the upper triangle and an on-device sweep.  keep[0..*num_keep) receives
 * indices into the ORIGINAL array in descending-score order, exactly `order[keep]`.
 *
 * mode   WSSDL_NMS_GE_F64: suppress iff (double)iou_f32 >= thresh   (cpu_nms.pyx:65 with a
 *                          Python-float thresh: cpu_nms.c:2492-2495)
 *        WSSDL_NMS_GT_F32: suppress iff iou_f32 > (float)thresh     (nms_kernel.cu:71,
 *                          py_cpu_nms.py:35)
 *        | WSSDL_NMS_CONTAIN: additionally suppress when inter/area_i > 0.95 or
 *                          inter/area_j > 0.95                       (nms_new, nms.pyx:117-120)
 * max_keep  stop after this many kept boxes (<=0: no limit); keep must hold
 *           min(N, max_keep>0 ? max_keep : N) ints.
 * status    device int[2] (may be NULL): [0] = 1 if some pair had a zero union.
 */
enum { WSSDL_NMS_GE_F64 = 0, WSSDL_NMS_GT_F32 = 1, WSSDL_NMS_CONTAIN = 4 };

size_t wssdl_nms_workspace_bytes(int N);

int wssdl_nms(const float* dets, int N, int dets_stride, double thresh, int mode,
              int max_keep, int* keep, int* num_keep, int* status, void* workspace,
              size_t workspace_bytes, wssdl_stream_t stream);

/* Host-pointer twins.  wssdl_gpu_nms_host has the argument list of the reference's
 * `_nms` (nms/gpu_nms.hpp:1-2: boxes already sorted by descending score, `>` against a
 * float threshold, keep_out = positions in the sorted array) and can be bound in its
 * place; wssdl_nms_host is the cpu_nms-shaped call (unsorted dets, double threshold,
 * `>=`, keep_out = original indices). */
int wssdl_gpu_nms_host(int* keep_out, int* num_out, const float* boxes_host, int boxes_num,
                       int boxes_dim, float nms_overlap_thresh, int device_id);
int wssdl_nms_host(int* keep_out, int* num_out, const float* dets_host, int N,
                   int dets_stride, double thresh, int mode, int max_keep, int device_id);

/* ---------------------------------------------------------------- IoU matrices
 * Replaces bbox_overlaps (utils/bbox.pyx:15-55) and bbox_overlaps_ui
 * (utils/bbox_ui.pyx:12-46).  boxes [N,4], query [K,4] -> out [N,K], row major.
 * kind WSSDL_IOU: intersection over union (+1 convention); WSSDL_IOU_UI: intersection
 * over area(boxes[n]).  The f64 entry evaluates the reference's expression tree with one
 * rounding per operation (no FMA) and is bit-exact; the f32 entry is the fast variant.
 */
enum { WSSDL_IOU = 0, WSSDL_IOU_UI = 1 };

int wssdl_bbox_overlaps_f64(const double* boxes, int N, const double* query, int K, int kind,
                            double* out, wssdl_stream_t stream);
int wssdl_bbox_overlaps_f32(const float* boxes, int N, const float* query, int K, int kind,
                            float* out, wssdl_stream_t stream);

/* ---------------------------------------------------------------- box transforms
 * Replaces bbox_transform_inv / clip_boxes / bbox_transform
 * (fast_rcnn/bbox_transform.py:30-61, :63-77, :10-28), fp32.
 * boxes [N,4]; deltas/out [N,4*k] (class-major groups of 4, :41-44).
 * exp() is evaluated in fp64 and rounded once to fp32 (correctly rounded); numpy's SIMD
 * expf is within 1 ulp of that, hence the 1e-5 relative tolerance on decoded boxes.
 */
int wssdl_bbox_transform_inv(const float* boxes, const float* deltas, int N, int k,
                             float* out, wssdl_stream_t stream);
int wssdl_clip_boxes(float* boxes, int N, int k, float im_h, float im_w,
                     wssdl_stream_t stream);
int wssdl_bbox_transform(const float* ex_rois, const float* gt_rois, int N, float* targets,
                         wssdl_stream_t stream);

/* ---------------------------------------------------------------- RPN proposals
 * Replaces proposal_layer (rpn_msr/proposal_layer_tf_bus.py:19-148): anchors generated on
 * the fly (generate_anchors.py:37-97 + shifts :55-71), bbox_transform_inv, clip_boxes,
 * _filter_boxes (:151-156), score sort + pre-NMS top-N (:129-133), NMS (:138), post-NMS
 * top-N (:139-146), batched over images: one CTA per image, everything in shared memory
 * (batches that leave SMs idle: a thread-block cluster of 2 / 4 / 8 CTAs per image splits both
 * stages of the keep-list NMS rounds and exchanges alive bits and column masks through
 * distributed shared memory; same results; WSSDL_TUNE_PROPOSALS_CLUSTER overrides the choice).
 *
 * cls_prob  [B,H,W,2A] f32 NHWC (fg score of anchor a = channel A+a, :86)
 * bbox_pred [B,H,W,4A] f32 NHWC (deltas of anchor a = channels 4a..4a+3, :106)
 * im_info   [B,info_stride] f32 rows (im_h, im_w, im_scale, ...)
 * base_anchors [A,4] f32, HOST pointer (integer valued; generate_anchors output, A <= 32;
 *           it is copied into the kernel's parameter block)
 * rois      [B*post_nms_topN,5] f32: image b owns rows [b*post, b*post+counts[b]);
 *           rows (b, x1,y1,x2,y2); unused rows are zero-filled
 * scores    [B*post_nms_topN] f32 (may be NULL), anchor_idx [B*post_nms_topN] i32 (may be
 *           NULL): the (h,w,a) anchor each RoI came from
 * counts    [B] i32
 * decoded   [B,H*W*A,4] f32 (may be NULL): every anchor decoded+clipped, row order (h,w,a)
 *           -- the intermediate of :116-119, for inspection and parity tests
 * nms_mode  WSSDL_NMS_GE_F64 (cpu_nms, what nms_wrapper.nms runs with cfg.USE_GPU_NMS = False,
 *           the reference's setting) or WSSDL_NMS_GT_F32 (gpu_nms / py_cpu_nms: `>` against the
 *           float threshold, what it runs with cfg.USE_GPU_NMS = True)
 * pre_nms_topN <= 0 means "no truncation" (:130); the candidates are sorted lazily, M =
 * max(1024, pow2ceil(2*post_nms_topN)) at a time (fewer when shared memory is short), so any
 * pre_nms_topN fits.
 * Limits (WSSDL_ELIMIT otherwise): H*W*A <= 32768 anchors per image, 0 < post_nms_topN <=
 * 4096 (the callers map the reference's "post_nms_topN <= 0: no truncation", :139, onto
 * min(pre_nms_topN, H*W*A) when that fits), and the per-image state must fit one SM's shared
 * memory: 8*1024 + 4*H*W*A + 20*post_nms_topN + 18 KB <= 226 KB (17100 anchors with 6000->300,
 * 12000->2000 or no truncation -> 4096 fit).
 */
size_t wssdl_proposals_workspace_bytes(int B, int H, int W, int A, int pre_nms_topN,
                                       int post_nms_topN);

int wssdl_proposals(const float* cls_prob, const float* bbox_pred, const float* im_info,
                    int info_stride, int B, int H, int W, int A, const float* base_anchors,
                    int feat_stride, int pre_nms_topN, int post_nms_topN, double nms_thresh,
                    int nms_mode, float min_size, float* rois, float* scores, int* anchor_idx,
                    int* counts,
                    float* decoded, void* workspace, size_t workspace_bytes,
                    wssdl_stream_t stream);

/* ---------------------------------------------------------------- anchor labelling
 * Replaces the deterministic part of anchor_target_layer[_joint]
 * (rpn_msr/anchor_target_layer_tf_bus.py:410-509): inside filter, fp64 IoU against the
 * foreground GT rows, uni-directional overlap against the background GT rows, labels.
 * Batched over images; the npr.choice subsampling (:512-527) stays on the host.
 *
 * gt_boxes [B,max_gt,5] f32 (x1,y1,x2,y2,cls), fg rows (cls != 0) first (:434-436);
 * num_gt [B] i32; base_anchors [A,4] f32 HOST pointer.
 * dataset_mode 0: 'SNUBH' (:430-468, explicit background boxes label anchors 0 when their
 *                 uni-directional overlap >= positive_overlap);
 *              1: 'SNUBH_FG' (:470-477, fg rows only) ; 2: other datasets (all rows are fg);
 *                 modes 1,2 use the classic rules with negative_overlap (:497-509).
 * labels [B,H*W*A] f32 in (h,w,a) order: 1 fg, 0 bg, -1 don't care (anchors not inside
 * the image are -1, as after _unmap :568); argmax_gt [B,H*W*A] i32 = row of the best fg GT
 * (first maximum, -1 outside); max_overlap [B,H*W*A] f64 (may be NULL).
 * An image without fg rows yields no positives (the reference raises in argmax there).
 * Limits: A <= 32, max_gt <= 64, B <= 65535.
 */
size_t wssdl_anchor_labels_workspace_bytes(int B, int H, int W, int A, int max_gt);

int wssdl_anchor_labels(const float* gt_boxes, const int* num_gt, int max_gt,
                        const float* im_info, int info_stride, int B, int H, int W, int A,
                        const float* base_anchors, int feat_stride, int dataset_mode,
                        double positive_overlap, double negative_overlap,
                        int clobber_positives, float* labels, int* argmax_gt,
                        double* max_overlap, void* workspace, size_t workspace_bytes,
                        wssdl_stream_t stream);

/* ---------------------------------------------------------------- anchor targets (sampled part)
 * Replaces the rest of anchor_target_layer[_joint] (rpn_msr/anchor_target_layer_tf_bus.py:
 * 512-611): fg / bg subsampling, regression targets (:533 -> fast_rcnn/bbox_transform.py:10-28
 * with the reference's dtypes: float64 anchors, float32 GT rows, one cast to float32), inside /
 * outside weights, `_unmap` and the final layouts, for the whole batch in one launch fed by
 * wssdl_anchor_labels -- nothing returns to the host in between.
 *
 * labels_pre, argmax_gt  [B_supervised, H*W*A] as written by wssdl_anchor_labels
 * gt_boxes   [B_supervised, max_gt, 5] f32 (the same tensor)
 * B_total >= B_supervised: the images behind the supervised ones are weakly supervised and get
 *            labels -1, zero targets and weights (:613-626)
 * num_fg = int(RPN_FG_FRACTION * RPN_BATCHSIZE), batchsize = RPN_BATCHSIZE (:513, :523)
 * sample_mode WSSDL_SAMPLE_RANKS: the caller drew with the host RNG (the reference's stream);
 *            `ranks` (device) holds, image after image, the ranks of the fg anchors to disable and
 *            then those of the bg anchors (rank = position among the image's anchors of that label
 *            in ascending index order -- what npr.choice(len, size, replace=False) returns),
 *            rank_off [2*B_supervised+1] (device) their offsets; the population sizes come from
 *            wssdl_anchor_label_counts (counts [B,2] = fg, bg);
 *            WSSDL_SAMPLE_PHILOX: drawn on the device, no host trip: anchor i of image b has the
 *            key philox4x32-10(counter (i, b, which, 0), key seed).x, which = 0 fg / 1 bg; the
 *            anchors with the smallest (key, i) are disabled.
 * inside_weights [4] f32 HOST (RPN_BBOX_INSIDE_WEIGHTS); positive_weight = RPN_POSITIVE_WEIGHT
 *            (< 0: uniform 1 / #examples, :541-545)
 * labels_out [B_total,1,A*H,W] f32; targets, inside, outside [B_total,4A,H,W] f32 (:568-598)
 * final_counts [B_supervised,2] i32 (may be NULL): fg, bg after subsampling
 * Limits: A <= 32, H*W*A <= 32768.
 */
enum { WSSDL_SAMPLE_RANKS = 0, WSSDL_SAMPLE_PHILOX = 1 };

int wssdl_anchor_label_counts(const float* labels, int B, int NA, int* counts,
                              wssdl_stream_t stream);

int wssdl_anchor_targets(const float* labels_pre, const int* argmax_gt, const float* gt_boxes,
                         int max_gt, int B_supervised, int B_total, int H, int W, int A,
                         const float* base_anchors, int feat_stride, int num_fg, int batchsize,
                         int sample_mode, const int* ranks, const int* rank_off,
                         unsigned long long seed, const float* inside_weights,
                         double positive_weight, float* labels_out, float* targets, float* inside,
                         float* outside, int* final_counts, wssdl_stream_t stream);

/* ---------------------------------------------------------------- the hot path as one call
 * proposal_layer -> roi_pool for a batch of images, the composition the reference runs per
 * image inside one sess.run (networks/VGGnet_test_bus.py:57-62; rpn_msr/proposal_layer_tf_bus.py
 * :19-148 feeding roi_pooling_op.cc:89-224).  Same arguments and results as wssdl_proposals
 * followed by wssdl_roi_pool_fwd on its RoI blob (feat is the [B,H,W,C] map the RPN head ran on),
 * with two differences: the blob's unused rows (row >= counts[b] of image b's post_nms_topN rows)
 * carry batch index -1 and pool to zeros / argmax -1, and the pooling skips its RoI grouping
 * pre-pass because the blob is image-major by construction.  post_nms_topN must be > 0.
 * rois_ready_event: NULL, or a cudaEvent_t recorded on `stream` between the two stages, when
 * rois / scores / counts are final: a consumer on another stream (the all-gather of the
 * detections) can start while the pooling runs.
 * top [B*post_nms_topN,PH,PW,C], argmax the same shape or NULL. */
size_t wssdl_hot_path_fwd_workspace_bytes(int B, int post_nms_topN, int PH, int PW);

int wssdl_hot_path_fwd(const float* feat, const float* cls_prob, const float* bbox_pred,
                       const float* im_info, int info_stride, int B, int H, int W, int C, int A,
                       const float* base_anchors, int feat_stride, int pre_nms_topN,
                       int post_nms_topN, double nms_thresh, int nms_mode, float min_size, int PH,
                       int PW, float spatial_scale, int bin_mode, float* rois, float* scores,
                       int* counts, float* top, int* argmax, void* workspace,
                       size_t workspace_bytes, wssdl_stream_t stream, void* rois_ready_event);

/* Stage 1 of the hot path on its own (the proposals with the blob convention above: unused rows
 * carry batch index -1; always one CTA per image, i.e. least SM time rather than least latency),
 * for callers that pipeline the two stages over consecutive batches on two streams; stage 2 is
 * wssdl_roi_pool_fwd_grouped on the blob. */
int wssdl_hot_path_proposals(const float* cls_prob, const float* bbox_pred, const float* im_info,
                             int info_stride, int B, int H, int W, int A,
                             const float* base_anchors, int feat_stride, int pre_nms_topN,
                             int post_nms_topN, double nms_thresh, int nms_mode, float min_size,
                             float* rois, float* scores, int* counts, wssdl_stream_t stream);

/* wssdl_roi_pool_fwd for image-major RoIs: row r belongs to image r / roi_stride (R = B *
 * roi_stride); rows whose batch index is not their block's image pool to zeros / -1.  The layout
 * wssdl_proposals writes; saves the RoI grouping pre-pass of the sorted-bins kernel. */
int wssdl_roi_pool_fwd_grouped(const float* bottom, const float* rois, int roi_stride, int B,
                               int H, int W, int C, int PH, int PW, float spatial_scale,
                               int bin_mode, float* top, int* argmax, void* workspace,
                               size_t workspace_bytes, wssdl_stream_t stream);

/* ---------------------------------------------------------------- proposal targets, on device
 * Replaces proposal_target_layer / proposal_target_layer_joint for the supervised images
 * (rpn_msr/proposal_target_layer_tf_bus.py:15-184) and _sample_rois (:228-280),
 * _compute_targets (:213-226), _get_bbox_regression_labels (:187-210).  Two launches:
 *
 * wssdl_roi_match   per image i < B_supervised: the candidates are the rows of `rois` whose batch
 *            column is i, in their original order, followed -- when add_gt -- by the image's fg GT
 *            rows (the leading rows of gt_boxes[i,:num_gt[i]] with class != 0, :38-50); fp64 IoU as
 *            utils/bbox.pyx:15-55, max / first argmax (:233-235), fg candidacy max >= fg_thresh
 *            (:239), bg candidacy bg_thresh_lo <= max < bg_thresh_hi (:252-253).
 *            counts [B_supervised,4] i32 = (candidates, fg candidates, bg candidates, fg GT rows).
 *            The candidate table stays in `workspace` for wssdl_roi_targets.
 * wssdl_roi_targets  the selected candidates into output rows: RoI (GT candidates as (i, box)),
 *            label (class of the assigned GT; 0 from row fg_this on, :264-266), fp32 bbox_transform
 *            targets, normalised in fp64 when the two double[4] arrays are given (:221-224),
 *            expanded to [.,4*num_classes] with inside / outside (= inside > 0, :84) weights.
 *   WSSDL_SAMPLE_RANKS   sel holds, per image, the fg then the bg selection as RANKS among the
 *            fg / bg candidates, in selection order -- what npr.choice(n, size=k, replace=False)
 *            returns for a population of that size; sel_off [2*B_supervised+1] delimits them and
 *            row_off [B_supervised+1] is the first output row of every image (no padding).
 *   WSSDL_SAMPLE_PHILOX  fg_this = min(fg_rois_per_image, n_fg), bg_this = min(rois_per_image -
 *            fg_this, n_bg) (:243, :256-258); the fg_this (bg_this) candidates with the smallest
 *            (philox4x32-10(counter (rank, i, 2 fg / 3 bg, 0), key seed).x, rank), in that order.
 *            Image i owns output rows [i*rois_per_image, (i+1)*rois_per_image); rows past
 *            fg_this + bg_this are zero; out_counts [B_supervised,2] = (fg_this, bg_this).
 * Limits: max_gt <= 64; with Philox rois_per_image + R + max_gt ints of shared memory <= 200 KB,
 * else WSSDL_ELIMIT. */
size_t wssdl_roi_targets_workspace_bytes(int B_supervised, int R, int max_gt);

int wssdl_roi_match(const float* rois, int R, const float* gt_boxes, const int* num_gt,
                    int max_gt, int B_supervised, int add_gt, double fg_thresh,
                    double bg_thresh_hi, double bg_thresh_lo, void* workspace,
                    size_t workspace_bytes, int* counts, wssdl_stream_t stream);

int wssdl_roi_targets(const float* rois, int R, const float* gt_boxes, int max_gt,
                      int B_supervised, int num_classes, const void* workspace, const int* counts,
                      int sample_mode, const int* sel, const int* sel_off, const int* row_off,
                      int fg_rois_per_image, int rois_per_image, unsigned long long seed,
                      const double* normalize_means, const double* normalize_stds,
                      const float* inside_weights, float* out_rois, float* out_labels,
                      float* out_targets, float* out_inside, float* out_outside, int* out_counts,
                      wssdl_stream_t stream);

/* ---------------------------------------------------------------- detection post-processing
 * Replaces the tail of im_detect (fast_rcnn/test_bus.py:207-223: rois / im_scale,
 * bbox_transform_inv per class, _clip_boxes :124-134) and the per-image body of test_net
 * (fast_rcnn/test_bus.py:360-401; twin fast_rcnn/train_bus.py:453-514): per class j >= 1
 * keep scores > score_thresh, greedy NMS (utils.cython_nms.nms: >= against a double
 * threshold), optional class-agnostic NMS over the survivors (:371-386), cap at
 * max_per_image over all classes (:394-401: keep scores >= the max_per_image-th largest).
 * Batched: one CTA per image, nothing leaves the device.
 *
 * rois       [B*roi_stride,5] f32 (batch, x1,y1,x2,y2) in the SCALED frame; image b owns rows
 *            [b*roi_stride, b*roi_stride + roi_counts[b]) (roi_counts NULL: all roi_stride)
 *            -- the layout wssdl_proposals writes
 * scores     [B*roi_stride,K] f32 class probabilities (column 0 = background)
 * bbox_pred  [B*roi_stride,4K] f32 class-major deltas
 * im_meta    [B,3] f32 rows (im_h, im_w of the UNSCALED image = im.shape[:2], im_scale)
 * dets       [B,K,roi_stride,5] f32 (x1,y1,x2,y2,score), class j of image b in
 *            descending-score order, det_counts [B,K] i32 valid rows (class 0: always 0);
 *            unused rows are zero-filled
 * pred_boxes [B*roi_stride,4K] f32 (may be NULL): the regressed + clipped boxes of :222-223
 * status     int[1] (may be NULL): 1 if some visited pair had a zero union
 * Limits: roi_stride <= 1024, K <= 64, (K-1)*roi_stride <= 1024 when cls_agnostic and <= 65536
 * otherwise (e.g. 21 classes x 300 RoIs with max_per_image = 100).
 */
int wssdl_detect_postprocess(const float* rois, const int* roi_counts, int roi_stride,
                             const float* scores, const float* bbox_pred, const float* im_meta,
                             int B, int K, float score_thresh, double nms_thresh,
                             int max_per_image, int cls_agnostic, float* dets, int* det_counts,
                             float* pred_boxes, int* status, wssdl_stream_t stream);

/* ---------------------------------------------------------------- evaluation: detection matching
 * Replaces the per-detection loop and the CorLoc pass of voc_eval_bus
 * (datasets/voc_eval_bus.py:206-247, :161-204) on the detection blob of
 * wssdl_detect_postprocess (after the all-gather): every detection of class j >= 1 is scored
 * against its image's ground-truth boxes of that class in fp64 (reference operation order,
 * bit-exact overlaps), best overlap = first maximum, then TP / FP / difficult / duplicate rules
 * in descending-score order per image, which is the order the reference's global walk visits
 * each image's detections in.  The argsort/cumsum/AP bookkeeping stays on the host.
 *
 * dets [B,K,S,5], det_counts [B,K]: as written by wssdl_detect_postprocess
 * gt_boxes [B,G,5] f32 (x1,y1,x2,y2,cls) in the detections' frame, num_gt [B] i32,
 * difficult [B,G] u8 (may be NULL = none difficult); G <= 64
 * tp, fp, fp_froc [B,K,S] u8: flags per detection (rows >= count and class 0 are not touched /
 *            zeroed; allocate zero-filled); fp_froc = score >= score_thresh and overlap <= ovthresh
 * img_stats [B,K,2] i32: [0] image has GT of the class (counts toward ni), [1] some detection
 *            with score >= score_thresh overlaps a GT by more than ovthresh (counts toward nok)
 * npos [K] i32: non-difficult GT boxes per class (:135)
 */
int wssdl_eval_match(const float* dets, const int* det_counts, int B, int K, int S,
                     const float* gt_boxes, const int* num_gt, const unsigned char* difficult,
                     int G, double ovthresh, float score_thresh, unsigned char* tp,
                     unsigned char* fp, unsigned char* fp_froc, int* img_stats, int* npos,
                     wssdl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WSSDL_B200_H_ */
