// The detector's hot path as ONE call: proposal_layer -> roi_pool for a batch of images
// (VGGnet_test_bus.py:57-62 composes the two ops per image inside one sess.run).
//
// What the fused entry buys over calling wssdl_proposals and wssdl_roi_pool_fwd back to back:
// the RoI blob it hands to the pooling is image-major with a fixed stride by construction, so the
// sorted-bins pre-pass needs no per-image RoI lists (two launches and their gaps less), and the
// unused rows behind an image's RoIs carry batch index -1: every pooling kernel writes zeros / -1
// for them instead of pooling a dummy box.
#include "common.cuh"

extern "C" size_t wssdl_hot_path_fwd_workspace_bytes(int B, int post_nms_topN, int PH, int PW) {
  if (B < 0 || post_nms_topN <= 0) return 256;
  return wssdl_roi_pool_fwd_workspace_bytes(B, B * post_nms_topN, PH, PW) + 256;
}

extern "C" int wssdl_hot_path_fwd(const float* feat, const float* cls_prob, const float* bbox_pred,
                                  const float* im_info, int info_stride, int B, int H, int W,
                                  int C, int A, const float* base_anchors, int feat_stride,
                                  int pre_nms_topN, int post_nms_topN, double nms_thresh,
                                  int nms_mode, float min_size, int PH, int PW,
                                  float spatial_scale, int bin_mode, float* rois, float* scores,
                                  int* counts, float* top, int* argmax, void* workspace,
                                  size_t workspace_bytes, wssdl_stream_t stream,
                                  void* rois_ready_event) {
  if (B < 0 || post_nms_topN <= 0 || (long long)B * post_nms_topN >= (1ll << 31)) return WSSDL_EINVAL;
  int rc = wssdl_proposals_impl(cls_prob, bbox_pred, im_info, info_stride, B, H, W, A, base_anchors,
                                feat_stride, pre_nms_topN, post_nms_topN, nms_thresh, nms_mode,
                                min_size, rois, scores, nullptr, counts, nullptr, stream, -1.0f, 0);
  if (rc != WSSDL_OK) return rc;
  if (rois_ready_event)   // rois / scores / counts are final here: a consumer on another stream
    WSSDL_RETURN_IF_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(rois_ready_event), to_cuda(stream)));
  return wssdl_roi_pool_fwd_impl(feat, rois, B, H, W, C, B * post_nms_topN, PH, PW, spatial_scale,
                                 bin_mode, top, argmax, workspace, workspace_bytes, stream,
                                 post_nms_topN);
}

// Stage 1 of the hot path on its own, for callers that pipeline the two stages over consecutive
// batches on two streams (pipeline.PipelinedHotPath: the proposals of batch k+1 run while batch k
// is pooled): wssdl_proposals with the hot path's blob convention (unused rows carry batch index
// -1) and always one CTA per image: overlapped with another batch's pooling it should take the
// least SM time, not the least latency.  Stage 2 is wssdl_roi_pool_fwd_grouped on the blob.
extern "C" int wssdl_hot_path_proposals(const float* cls_prob, const float* bbox_pred,
                                        const float* im_info, int info_stride, int B, int H, int W,
                                        int A, const float* base_anchors, int feat_stride,
                                        int pre_nms_topN, int post_nms_topN, double nms_thresh,
                                        int nms_mode, float min_size, float* rois, float* scores,
                                        int* counts, wssdl_stream_t stream) {
  if (post_nms_topN <= 0) return WSSDL_EINVAL;
  return wssdl_proposals_impl(cls_prob, bbox_pred, im_info, info_stride, B, H, W, A, base_anchors,
                              feat_stride, pre_nms_topN, post_nms_topN, nms_thresh, nms_mode,
                              min_size, rois, scores, nullptr, counts, nullptr, stream, -1.0f, 1);
}
