/*
 * mrcnn_b200.h — C ABI of libmrcnn_b200.so: the B200-native (sm_100a) RoI hot path of Mask R-CNN.
 *
 * This is the drop-in boundary for the native layer of delldu/MaskRCNN:
 *     c++ext/maskrcnn/csrc/vision.cpp:11-15   pybind module `maskrcnn._C` { nms, crop_forward, crop_backward }
 *     c++ext/maskrcnn/csrc/nms.h:15-30        at::Tensor nms(const at::Tensor& dets, float threshold)
 *     c++ext/maskrcnn/csrc/crop.h:14-34       void crop_forward(image, boxes, box_index, extrapolation_value,
 *                                                               crop_height, crop_width, crops&)
 *     c++ext/maskrcnn/csrc/crop.h:36-53       void crop_backward(grads, boxes, box_index, grads_image&)
 * plus fused entry points for the Python-level callers of those ops (model.py:276-393 roi_align,
 * :1307-1382 rpn_refine, :1389-1487 mrn_refine), which the reference runs as dozens of small launches.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers + sizes + a CUDA stream (cudaStream_t passed as void*); no torch types.
 *   - every function returns MRCNN_OK (0) or a negative MRCNN_E_* code; mrcnn_last_error() gives the text.
 *     Nothing ever calls exit() (the reference does: cpu/crop_cpu.cpp:47-50).
 *   - all work is enqueued on `stream`; no call synchronises the device or copies to the host.
 *     Data-dependent sizes (number of kept boxes) are written to device counters supplied by the caller.
 *   - the caller owns every buffer, including workspaces (size them with the *_workspace_bytes calls).
 *   - fp32 boxes are (y1, x1, y2, x2); box_index is int32; keep indices are int64 — as in the reference.
 *   - there is NO CPU implementation behind this ABI: host pointers are rejected (MRCNN_E_NOT_DEVICE_PTR).
 */
#ifndef MRCNN_B200_H_
#define MRCNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRCNN_ABI_VERSION 6

#define MRCNN_OK 0
#define MRCNN_E_INVALID_ARG (-1)    /* bad size / null pointer / unsupported combination            */
#define MRCNN_E_NOT_DEVICE_PTR (-2) /* a data pointer is not device memory (no CPU fallback exists) */
#define MRCNN_E_WORKSPACE (-3)      /* workspace too small                                          */
#define MRCNN_E_CUDA (-4)           /* a CUDA runtime call / launch failed                          */
#define MRCNN_E_BOX_INDEX (-5)      /* reported by mrcnn_poll_device_errors: box_index out of range */
#define MRCNN_E_CLASS_ID (-6)       /* reported by mrcnn_poll_device_errors: class id out of range  */

/* Memory layout of 4-D tensors.  Logical shape is always [N, C, H, W] as in the reference;
 * NHWC means the same tensor stored channels-last (torch.channels_last), which is what the
 * 128-bit channel-vectorised kernels want. */
#define MRCNN_NCHW 0
#define MRCNN_NHWC 1

typedef void* mrcnn_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define MRCNN_API __attribute__((visibility("default")))
#else
#define MRCNN_API
#endif

MRCNN_API int mrcnn_abi_version(void);
MRCNN_API const char* mrcnn_last_error(void); /* thread-local, valid until the next failing call on this thread */

/* Kernels flag recoverable data errors (box_index out of range: the reference exit(-1)s on CPU,
 * cpu/crop_cpu.cpp:47-50, and silently skips on CUDA, cuda/crop_cuda.cu:41-44) in a device word;
 * offending boxes produce extrapolation_value / contribute no gradient.  This call synchronises
 * `stream`, returns MRCNN_E_BOX_INDEX (or MRCNN_E_CLASS_ID: mrcnn_full_masks saw a class id outside [0, NC)) if the
 * flag was raised since the last poll, and clears it. */
MRCNN_API int mrcnn_poll_device_errors(mrcnn_stream_t stream);

/* ---- crop_and_resize (replaces crop_forward / crop_backward, crop.h:14-53) -------------------- */

/* image [B,C,H,W] (image_layout), boxes [N,4] normalised, box_index [N] in [0,B),
 * crops [N,C,crop_h,crop_w] (crops_layout) — every element is written (no pre-zeroing needed). */
MRCNN_API int mrcnn_crop_forward(const float* image, int B, int C, int H, int W, int image_layout,
                       const float* boxes, const int32_t* box_index, int N,
                       float extrapolation_value, int crop_h, int crop_w,
                       float* crops, int crops_layout, mrcnn_stream_t stream);

/* grads [N,C,crop_h,crop_w] (grads_layout) scatter-added into grads_image [B,C,H,W] (image_layout).
 * zero_fill != 0: grads_image is cleared first, as crop_cpu.cpp:197 does.  No gradient w.r.t. boxes. */
MRCNN_API int mrcnn_crop_backward(const float* grads, int grads_layout, const float* boxes,
                        const int32_t* box_index, int N, int crop_h, int crop_w,
                        float* grads_image, int B, int C, int H, int W, int image_layout,
                        int zero_fill, mrcnn_stream_t stream);

/* ---- PyramidROIAlign (replaces model.py:276-393 roi_align and its autograd backward) ---------- */

/* fm[l] = pyramid level P(l+2), l = 0..3: [B,C,H[l],W[l]] in fm_layout.  boxes [N,4] normalised;
 * box_index [N] = image of each box, or NULL for all-zero (the reference is batch-1, model.py:369).
 * image_area = float(image_height * image_width) (model.py:331).  out [N,C,pool,pool] (out_layout) in
 * INPUT box order.  levels_out: optional int32 [N] receiving the assigned level (2..5). */
MRCNN_API int mrcnn_pyramid_roi_align_forward(const float* const fm[4], const int H[4], const int W[4],
                                    int B, int C, int fm_layout,
                                    const float* boxes, const int32_t* box_index, int N, int pool,
                                    float image_area, float* out, int out_layout,
                                    int32_t* levels_out, mrcnn_stream_t stream);

/* Both heads' forward in ONE launch: the same RoIs pooled at 7x7 (box head, model.py:778) and 14x14 (mask head, model.py:889)
 * from a channels-last pyramid into channels-last crops out7 [N,C,7,7] / out14 [N,C,14,14] (C % 4 == 0, everything 16-byte
 * aligned).  Each CTA computes a RoI's 14x14 bins and then its 7x7 bins, whose taps lie inside the footprint the first pass has
 * just read: the second head costs its output bytes and almost no DRAM reads.  Bit-identical to two
 * mrcnn_pyramid_roi_align_forward calls (pool 7 and pool 14, extrapolation value 0). */
MRCNN_API int mrcnn_pyramid_roi_align_forward_pair(const float* const fm[4], const int H[4], const int W[4],
                                    int B, int C, const float* boxes, const int32_t* box_index, int N,
                                    float image_area, float* out7, float* out14, mrcnn_stream_t stream);

/* Adjoint of the forward: grads [N,C,pool,pool] (grads_layout) accumulated into gfm[l] [B,C,H[l],W[l]]
 * (gfm_layout), which is cleared first when zero_fill != 0 (what CropFunction.backward does per level,
 * c++ext/maskrcnn/__init__.py:52).  No gradient w.r.t. boxes (model.py:358).
 * algo:
 *   MRCNN_BWD_GATHER   row-owner gather: every run of 8 pixels of every gradient-map row is owned by one warp that
 *                      sums, in registers, the bins reaching it and writes each pixel exactly once - no atomics on
 *                      gradient data, no zero-fill pass, no read-modify-write of the pyramid.  Needs C % 4 == 0,
 *                      N > 0, N * pool^2 * C < 2^31, image_offsets_host == NULL and a 256-byte aligned workspace of
 *                      mrcnn_pyramid_roi_align_backward_workspace_bytes_ex() bytes.  Either layout of grads and gfm.
 *                      Summation order: a unit's bins are queued in the order the planning threads reach it (atomic
 *                      cursors), so two PLANS of the same boxes may add a pixel's terms in different orders and differ
 *                      in the last bits (tests bound it at 1e-5 relative); replaying one plan is bit-reproducible, and
 *                      mrcnn_set_deterministic(1) makes every plan order its items (bit-identical gradients).
 *                      The scatter's reductions are unordered by nature.
 *   MRCNN_BWD_SCATTER  clear, then scatter with column-aggregated 128-bit vector reductions
 *                      (red.global.add.v4.f32) for a channels-last gfm, scalar atomics for an NCHW gfm.
 *                      workspace may be NULL.
 *   MRCNN_BWD_AUTO     GATHER when its requirements are met (it is the faster one),
 *                      SCATTER otherwise.
 * image_offsets_host: optional HOST array of B+1 ints (scatter only): boxes of image i are rows
 * [off[i], off[i+1]), box_index is ignored, and the call clears + scatters image by image. */
#define MRCNN_BWD_AUTO 0
#define MRCNN_BWD_GATHER 1
#define MRCNN_BWD_SCATTER 2
/* Deterministic gather plans (process-wide; default off, or MRCNN_DETERMINISTIC=1 in the environment): every plan ends with a
 * pass that orders each unit's items by (gradient offset, column), so that two plans of the same boxes sum in the same order and
 * MRCNN_BWD_GATHER gradients are bit-reproducible across calls.  Set it BEFORE querying workspace sizes: the plan needs a second
 * item array (a plan given the smaller workspace fails with MRCNN_E_WORKSPACE).  No counterpart in the reference (its CPU loop has
 * one fixed order, its CUDA atomicAdd path none). */
MRCNN_API int mrcnn_set_deterministic(int on);
MRCNN_API size_t mrcnn_pyramid_roi_align_backward_workspace_bytes(const int H[4], const int W[4], int B, int N,
                                                                  int pool);
/* The same with the upstream-gradient layout taken into account: MRCNN_BWD_GATHER also serves NCHW gradient maps (every unit of
 * 8 pixels is one aligned 32-byte sector per channel plane, written once) and NCHW upstream gradients (what torch's conv
 * backward hands to an unmodified model.py), which it first transposes to [N][pool^2][C] in the tail of the workspace:
 * N * C * pool^2 * 4 more bytes.  For channels-last gradients it equals mrcnn_pyramid_roi_align_backward_workspace_bytes(). */
MRCNN_API size_t mrcnn_pyramid_roi_align_backward_workspace_bytes_ex(const int H[4], const int W[4], int B, int C, int N,
                                                                     int pool, int grads_layout);
MRCNN_API int mrcnn_pyramid_roi_align_backward(const float* grads, int grads_layout,
                                     const int H[4], const int W[4], int B, int C,
                                     const float* boxes, const int32_t* box_index, int N, int pool,
                                     float image_area, float* const gfm[4], int gfm_layout,
                                     int zero_fill, const int32_t* image_offsets_host, int algo,
                                     void* workspace, size_t workspace_bytes,
                                     mrcnn_stream_t stream);

/* The gather backward in two halves.  Its work-item queues (three small launches, ~60 us at 8192 RoIs) depend on the boxes,
 * pool, C and the pyramid geometry only - not on the gradients - so a training step can build them on a side stream
 * while the forward runs and keep the critical path of the backward to the one gather launch:
 *   _plan     fills `workspace` (mrcnn_pyramid_roi_align_backward_workspace_bytes(), 256-byte aligned) from the boxes;
 *   _planned  the gather itself: grads [N,C,pool,pool] and gfm channels-last, the SAME H, W, B, C, N, pool as the plan
 *             (not checked: the plan lives in device memory) and the workspace the plan filled, which it only reads -
 *             a plan can serve any number of backward calls.  zero_fill == 0 adds to what gfm holds.
 * Same requirements and same results as MRCNN_BWD_GATHER. */
MRCNN_API int mrcnn_pyramid_roi_align_backward_plan(const int H[4], const int W[4], int B, int C, const float* boxes,
                                                    const int32_t* box_index, int N, int pool, float image_area,
                                                    void* workspace, size_t workspace_bytes, mrcnn_stream_t stream);
MRCNN_API int mrcnn_pyramid_roi_align_backward_planned(const float* grads, const int H[4], const int W[4], int B, int C, int N,
                                                       int pool, float* const gfm[4], int zero_fill, const void* workspace,
                                                       size_t workspace_bytes, mrcnn_stream_t stream);

/* Both heads at once.  In training the reference pools the SAME RoIs twice from the same pyramid - 7x7 for the box head
 * (model.py:778) and 14x14 for the mask head (model.py:889) - and autograd then adds the two gradient pyramids that
 * CropFunction.backward returned per level.  This entry point takes both upstream gradients (channels-last,
 * grads_a [N,C,pool_a,pool_a], grads_b [N,C,pool_b,pool_b]) and writes their SUM into gfm[l] (channels-last) in one
 * row-owner gather: one pyramid write instead of two plus an add.  zero_fill == 0 adds to what gfm holds. */
MRCNN_API size_t mrcnn_pyramid_roi_align_backward_pair_workspace_bytes(const int H[4], const int W[4], int B, int N,
                                                                       int pool_a, int pool_b);
MRCNN_API int mrcnn_pyramid_roi_align_backward_pair(const float* grads_a, int pool_a, const float* grads_b, int pool_b,
                                                    const int H[4], const int W[4], int B, int C,
                                                    const float* boxes, const int32_t* box_index, int N,
                                                    float image_area, float* const gfm[4], int zero_fill,
                                                    void* workspace, size_t workspace_bytes, mrcnn_stream_t stream);

/* ---- NMS (replaces nms(), nms.h:15-30 -> cpu/nms_cpu.cpp:11-70) -------------------------------- */

/* dets [N,5] = (y1,x1,y2,x2,score), any order.  Suppress iff IoU >= threshold (the CPU rule,
 * nms_cpu.cpp:65 — the reference's own CUDA kernel uses '>').  keep_out: int64 [N], receives the
 * ASCENDING original indices of the survivors in its first *count_out entries (device int32).
 * From 449 boxes on the survivors are found by a grid-wide fixed-point iteration (one cooperative launch: it needs the
 * device to itself for a few microseconds at a time, like any cooperative kernel; it can be captured in a CUDA graph);
 * below that by a single-CTA sweep.  Both give the greedy result of nms_cpu.cpp bit for bit. */
MRCNN_API size_t mrcnn_nms_workspace_bytes(int N);
MRCNN_API int mrcnn_nms(const float* dets, int N, float threshold, int64_t* keep_out, int32_t* count_out,
              void* workspace, size_t workspace_bytes, mrcnn_stream_t stream);

/* ---- proposal layer (replaces MaskRCNN.rpn_refine, model.py:1307-1382; batched) ---------------- */

/* rpn_class [B,A,2] (bg,fg), rpn_bbox [B,A,4], anchors [A,4] px, std[4] = RPN_BBOX_STD_DEV (host ptr).
 * Top-pre_nms by fg score -> decode -> clip to [0,height]x[0,width] -> NMS(threshold) -> first post_nms
 * -> normalise.  rois_out [B,post_nms,4] (rows >= count are zero), counts_out int32 [B]. */
MRCNN_API size_t mrcnn_proposal_workspace_bytes(int B, int A, int pre_nms);
MRCNN_API int mrcnn_proposal_layer(const float* rpn_class, const float* rpn_bbox, const float* anchors,
                         int B, int A, int pre_nms, int post_nms, float nms_threshold,
                         const float* std4_host, float height, float width,
                         float* rois_out, int32_t* counts_out,
                         void* workspace, size_t workspace_bytes, mrcnn_stream_t stream);

/* NMS inside the proposal layer (process-wide switch; results are identical):
 *   MRCNN_PROPOSAL_NMS_MASK  batched 64x64 IoU-bitmask tiles (upper triangle) + one single-CTA sweep per image;
 *   MRCNN_PROPOSAL_NMS_LAZY  one 8-CTA cluster per image, boxes taken 64 at a time in score order and compared only with
 *                            the survivors found so far, stopping at the post_nms-th survivor: 64 * sum(survivors so
 *                            far) IoU tests instead of n^2 / 2; no N x N mask, no workspace traffic;
 *   MRCNN_PROPOSAL_NMS_HYBRID the first 1.25 post_nms boxes of every image resolved AT ONCE by a grid-wide fixed-point
 *                            iteration (lower-triangle IoU tiles, one CTA per 64 boxes, all images in one cooperative
 *                            launch), the lazy kernel only for images that still lack survivors after that prefix;
 *   MRCNN_PROPOSAL_NMS_AUTO  HYBRID when post_nms <= 2048, MASK otherwise (default). */
#define MRCNN_PROPOSAL_NMS_AUTO 0
#define MRCNN_PROPOSAL_NMS_MASK 1
#define MRCNN_PROPOSAL_NMS_LAZY 2
#define MRCNN_PROPOSAL_NMS_HYBRID 3
MRCNN_API int mrcnn_set_proposal_nms(int algo);

/* The same with the foreground probabilities alone, fg_scores [B,A] (what mrcnn_rpn_pack writes as fg_out): the layer's
 * only pass over the scores reads half the bytes. */
MRCNN_API int mrcnn_proposal_layer_fg(const float* fg_scores, const float* rpn_bbox, const float* anchors,
                            int B, int A, int pre_nms, int post_nms, float nms_threshold,
                            const float* std4_host, float height, float width,
                            float* rois_out, int32_t* counts_out,
                            void* workspace, size_t workspace_bytes, mrcnn_stream_t stream);

/* ---- RPN head output plumbing (replaces the permute / view / softmax of RPN.forward, model.py:624-641, and the three
 *      torch.cat of MaskRCNN.rpn_detect, model.py:1294-1304) ---------------------------------------------------- */

/* logits[l] [B,2K,H[l],W[l]] and bbox[l] [B,4K,H[l],W[l]] = conv_class / conv_bbox outputs of pyramid level l (HOST arrays
 * of `levels` device pointers; layout MRCNN_NCHW or MRCNN_NHWC for all of them), K = anchors per location.  Writes, in
 * the reference's anchor order (level-major, then y, x, anchor) with A = K * sum(H*W):
 *   logits_out [B,A,2] (rpn_class_logits), class_out [B,A,2] = softmax over (bg, fg) (rpn_class), bbox_out [B,A,4]
 *   (rpn_bbox), fg_out [B,A] = class_out[:,:,1] (input of mrcnn_proposal_layer_fg).  Any output may be NULL. */
MRCNN_API int mrcnn_rpn_pack(const float* const* logits, const float* const* bbox, const int* H, const int* W, int levels,
                             int B, int anchors_per_location, int layout, float* logits_out, float* class_out,
                             float* bbox_out, float* fg_out, mrcnn_stream_t stream);

/* Adjoint of the layout part of mrcnn_rpn_pack, for training: grad_logits [B,A,2] / grad_bbox [B,A,4] (either may be NULL
 * together with its destinations) written back as g_logits[l] [B,2K,H[l],W[l]] / g_bbox[l] [B,4K,H[l],W[l]] (HOST arrays of
 * device pointers, `layout` as in the forward).  The probabilities carry no gradient: they only feed the proposal layer. */
MRCNN_API int mrcnn_rpn_unpack(const float* grad_logits, const float* grad_bbox, const int* H, const int* W, int levels, int B,
                               int anchors_per_location, int layout, float* const* g_logits, float* const* g_bbox,
                               mrcnn_stream_t stream);

/* ---- detection layer (replaces MaskRCNN.mrn_refine, model.py:1389-1487; batched) --------------- */

/* rois [B,N,4] normalised, probs [B,N,NC], deltas [B,N,NC,4], windows [B,4] px (device).
 * min_confidence <= 0 disables the score filter (model.py:1441).  dets_out [B,max_inst,6] =
 * (y1,x1,y2,x2,score,class) score-descending, zero padded; counts_out int32 [B] (0 = the reference's
 * (None,None,None)); index_out optional int32 [B,max_inst] = source RoI of each detection. */
MRCNN_API size_t mrcnn_detection_workspace_bytes(int B, int N);
/* NMS inside the detection layer, same constants and meaning as mrcnn_set_proposal_nms (results are identical): LAZY =
 * chunks of 64 boxes in score order against the same-class survivors so far, stopping at the max_inst-th survivor (no
 * N x N mask); MASK = class-aware suppression words + sweep; AUTO (default) = LAZY when max_inst <= 1024.  With MASK
 * and N > 1024 the workspace above is required; LAZY never needs it. */
MRCNN_API int mrcnn_set_detection_nms(int algo);
MRCNN_API int mrcnn_detection_layer(const float* rois, const float* probs, const float* deltas,
                          const float* windows, int B, int N, int NC,
                          float min_confidence, float nms_threshold, int max_inst,
                          const float* std4_host, float height, float width,
                          float* dets_out, int32_t* counts_out, int32_t* index_out,
                          void* workspace, size_t workspace_bytes, mrcnn_stream_t stream);

/* ---- detection layer fused with what follows it (BASELINE configs[4]: 64 images sharded over the GPUs of one box) ----------
 * Replaces, on top of mrcnn_detection_layer: `mrn_rois = mrn_boxes.float() * 1.0 / h` (model.py:1188), the box_ind the mask
 * head's roi_align needs, and - there is no counterpart in the single-GPU reference - the all-gather of the detections over
 * the ranks, as ONE kernel: every CTA (= image) stores its packed row [max_inst * 6 detections, count] straight into the
 * exchange buffer of every rank through peer-mapped pointers (NVLink / NVSwitch stores), the rank's last CTA raises the rank's
 * flag on every rank (release at system scope).  mrcnn_detection_collect is the consuming half: per image, wait for the owner
 * rank's flag (acquire load from local memory) and copy the row into [total, max_inst, 6] / [total], image order = rank order
 * (contiguous balanced shards: the first total % world ranks own one image more).
 *
 * peer_bufs_host: HOST array of `world` device pointers, buffer of rank r mapped into this process (own buffer at [rank]); each
 * holds mrcnn_detection_exchange_bytes() bytes, zero-initialised once: [2 parities][total][max_inst * 6 + 1] floats, then [world]
 * uint32 flags.  state: two int32 of LOCAL device memory, zero-initialised once ([0] epoch, [1] CTA counter).  Successive
 * exchanges alternate parity, so a rank may run one exchange ahead of its peers; the results of exchange k must be consumed
 * (stream order is enough) before this rank issues exchange k + 2.  world == 0: no exchange (peer_bufs_host / state may be
 * NULL).  mask_boxes [B * max_inst, 4] / mask_box_ind [B * max_inst] (= (ind_offset + image) % ind_mod) may be NULL.
 * Every rank must issue the same sequence of exchanges; a rank that stops leaves the others spinning in the collect kernel. */
MRCNN_API size_t mrcnn_detection_exchange_bytes(int world, int total_images, int max_inst);
MRCNN_API int mrcnn_detection_layer_exchange(const float* rois, const float* probs, const float* deltas,
                          const float* windows, int B, int N, int NC,
                          float min_confidence, float nms_threshold, int max_inst,
                          const float* std4_host, float height, float width,
                          float* dets_out, int32_t* counts_out,
                          float* mask_boxes, int32_t* mask_box_ind, int ind_offset, int ind_mod,
                          void* const* peer_bufs_host, int world, int rank, int image_offset, int total_images,
                          int32_t* state, void* workspace, size_t workspace_bytes, mrcnn_stream_t stream);
MRCNN_API int mrcnn_detection_collect(const void* local_buf, int world, int total_images, int max_inst,
                          const int32_t* state, float* dets_all, int32_t* counts_all, mrcnn_stream_t stream);

/* ---- detection-target layer (replaces mrn_samples, model.py:396-576; data.boxes_overlaps data.py:151-189;
 *      data.boxes_deltas data.py:103-121) ------------------------------------------------------------------- */

/* Step 1, batched: rois [B,N,4] and gt_boxes [B,G,4] normalised (y1,x1,y2,x2), gt_class_ids int32 [B,G] (< 0 = COCO
 * crowd, 0 = padding; when an image has a crowd row only rows with class > 0 are matched, model.py:436-449).
 * Outputs per image: pos_idx / neg_idx int32 [B,N] = ASCENDING indices of the proposals with max IoU >= 0.5 /
 * < 0.5 and not on a crowd (torch.nonzero order, model.py:459, :513-516); assign int32 [B,N] = gt row of the max IoU
 * (first maximum); iou_max optional fp32 [B,N]; counts int32 [B,2] = {positives, negatives}. */
MRCNN_API int mrcnn_target_classify(const float* rois, const float* gt_boxes, const int32_t* gt_class_ids,
                                    int B, int N, int G, int32_t* pos_idx, int32_t* neg_idx, int32_t* assign,
                                    float* iou_max, int32_t* counts, mrcnn_stream_t stream);

/* Step 2 (optional, replaces the two torch.randperm draws, model.py:468 and :520, without a host round trip):
 * perm_pos[b] = stable argsort(keys_pos[b, :P_b]), perm_neg likewise; take int32 [B,2] = {min(P, pos_cap),
 * min(Q, neg_table[kept positives])}.  neg_table: DEVICE int32 [pos_cap + 1], built by the caller as
 * int(p / ratio - p) in double precision (model.py:518-519).  N <= 8192. */
MRCNN_API int mrcnn_target_select(const int32_t* counts, const float* keys_pos, const float* keys_neg,
                                  const int32_t* neg_table, int B, int N, int pos_cap, int32_t* perm_pos,
                                  int32_t* perm_neg, int32_t* take, mrcnn_stream_t stream);

/* Step 3: row t < take[b][0] is positive number t: proposal pos_idx[b][perm_pos[b][t]] (perm NULL = identity), its gt
 * class, (boxes_deltas / std4) and the mask_h x mask_w target cropped from gt_masks [B,G,H,W] (fp32) and rounded half
 * to even (model.py:474-507); the next take[b][1] rows are negatives (class 0, zero deltas / masks, model.py:525-541);
 * rows up to T are zero padding.  Outputs: rois_out [B,T,4], class_out int32 [B,T], deltas_out [B,T,4],
 * masks_out [B,T,mask_h,mask_w]. */
MRCNN_API int mrcnn_target_emit(const float* rois, const float* gt_boxes, const int32_t* gt_class_ids,
                                const float* gt_masks, int B, int N, int G, int H, int W, const int32_t* pos_idx,
                                const int32_t* neg_idx, const int32_t* perm_pos, const int32_t* perm_neg,
                                const int32_t* take, const int32_t* assign, const float* std4_host, int mask_h,
                                int mask_w, int T, float* rois_out, int32_t* class_out, float* deltas_out,
                                float* masks_out, mrcnn_stream_t stream);

/* ---- RPN anchor matching (replaces data.rpn_samples, data.py:449-591) ------------------------------------------ */

/* Steps 1-3 (data.py:495-535): anchors float64 [A,4] px, gt_boxes int32 [G,4] px, gt_class_ids int32 [G] (< 0 crowd).
 * match int32 [A]: -1 (max IoU < 0.3 and not on a crowd), +1 (max IoU >= 0.7, or the best anchor of a gt box), 0
 * neutral - BEFORE subsampling; argmax int32 [A] = gt row of the max IoU (first maximum, np.argmax).  IoU is fp32 as in
 * data.boxes_overlaps (data.py:151-189).  workspace: mrcnn_rpn_match_workspace_bytes(G), 8-byte aligned. */
MRCNN_API size_t mrcnn_rpn_match_workspace_bytes(int G);
MRCNN_API int mrcnn_rpn_match(const double* anchors, int A, const int32_t* gt_boxes, const int32_t* gt_class_ids, int G,
                              int32_t* match, int32_t* argmax, void* workspace, size_t workspace_bytes,
                              mrcnn_stream_t stream);

/* np.where(values == target)[0] on the device: ids_out int32 [n] receives the ascending indices, count_out (device
 * int32) their number.  workspace: mrcnn_compact_equal_workspace_bytes(n). */
MRCNN_API size_t mrcnn_compact_equal_workspace_bytes(int n);
MRCNN_API int mrcnn_compact_equal(const int32_t* values, int n, int32_t target, int32_t* ids_out, int32_t* count_out,
                                  void* workspace, size_t workspace_bytes, mrcnn_stream_t stream);

/* values[ids[perm[t]]] = fill for t < count: resets the anchors np.random.choice picked (data.py:543-553). */
MRCNN_API int mrcnn_scatter_fill(int32_t* values, const int32_t* ids, const int32_t* perm, int count, int32_t fill,
                                 mrcnn_stream_t stream);

/* data.py:557-589: float64 deltas / std for the positive anchors ids[0..*count) (ascending), rows up to T zero;
 * rpn_bbox_out float64 [T,4], 32-byte aligned. */
MRCNN_API int mrcnn_rpn_deltas(const double* anchors, const int32_t* gt_boxes, const int32_t* argmax, const int32_t* ids,
                               const int32_t* count, int T, const double* std4_host, double* rpn_bbox_out,
                               mrcnn_stream_t stream);

/* ---- mask paste-back (replaces data.full_masks, data.py:287-314) ------------------------------------------------ */

/* class_ids int64 [D], boxes [D,4] px (y1,x1,y2,x2; the detection layer's rounded boxes), masks [D,NC,mask_h,mask_w]
 * (the mask head's sigmoid output, model.py:1188) -> out uint8 [D,H,W], 1 where the detection's mask covers the pixel.
 * Per detection, exactly what the reference computes through PIL: (mask[class] * 255.0) truncated to 8 bits, resized to
 * (int(y2 - y1), int(x2 - x1)) with Pillow's 8-bit bilinear resample (22-bit fixed-point weights, horizontal pass into an
 * 8-bit intermediate, then vertical), pasted at (int(y1), int(x1)) clipped to the image, thresholded '> 127'.  Every
 * byte of `out` is written once (no pre-zeroing).  An empty box gives an empty mask (PIL raises ValueError there; zero
 * padded detection rows take this path).  mask_h, mask_w <= 64.  Any D: batches are just more rows.  workspace: 256-byte
 * aligned, mrcnn_full_masks_workspace_bytes() bytes (per detection: resampling taps + the 8-bit horizontal pass). */
MRCNN_API size_t mrcnn_full_masks_workspace_bytes(int D, int mask_h, int mask_w, int H, int W);
MRCNN_API int mrcnn_full_masks(const int64_t* class_ids, const float* boxes, const float* masks, int D, int NC, int mask_h,
                               int mask_w, int H, int W, uint8_t* out, void* workspace, size_t workspace_bytes,
                               mrcnn_stream_t stream);

/* ---- masks back to the original frame (replaces data.decode_masks, data.py:265-284) ---------------------------- */

/* masks [D,H,W], one byte per pixel: src_is_bool != 0 -> torch.bool bytes, non-zero = 255 (PIL mode '1' -> 'L',
 * data.py:272), else 'L' pixel values kept as they are.  The window rows [top, top + crop_h) x columns
 * [left, left + crop_w) of every mask (torchvision CenterCrop, data.py:273-274: the caller computes the origin
 * int(round((H - crop_h) / 2.0)) with the host language's round-half-even) are resized to out_h x out_w
 * (data.py:276-278) with Pillow's 8-bit bilinear resample - the same two-pass fixed-point arithmetic as
 * mrcnn_full_masks - into out uint8 [D,out_h,out_w]; NOT thresholded, like the reference.  Upscales and downscales.
 * workspace: 256-byte aligned, mrcnn_decode_masks_workspace_bytes() bytes (the taps of one column / row set). */
MRCNN_API size_t mrcnn_decode_masks_workspace_bytes(int crop_h, int crop_w, int out_h, int out_w);
MRCNN_API int mrcnn_decode_masks(const uint8_t* masks, int src_is_bool, int D, int H, int W, int top, int left, int crop_h,
                                 int crop_w, int out_h, int out_w, uint8_t* out, void* workspace, size_t workspace_bytes,
                                 mrcnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MRCNN_B200_H_ */
