/*
 * y3d.h -- C ABI of liby3d_b200.so: the B200 (sm_100a) implementation of the YOLOv10 / YOLOv10-3D
 * detection-head hot path (head decode, NMS-free top-k, dual task-aligned assignment, loss partials).
 *
 * The reference (baldhat/yolov10-3D) is pure Python/PyTorch and has no FFI for this path; each entry point
 * below replaces one reference *Python* function (cited per function, paths relative to the reference repo).
 * The host-side mirror of those Python signatures lives in yolov10-3d_b200/ and binds this ABI with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer named d_* / without prefix in a "device" section is a DEVICE pointer; lvl_* geometry arrays
 *    and the arrays of per-level pointers are small HOST arrays read during the call;
 *  - all buffers (inputs, outputs, workspace) are owned by the caller; nothing is allocated, freed or retained;
 *  - every function only enqueues work on `stream` (a cudaStream_t passed as void*; NULL = default stream),
 *    never synchronises the device and keeps no global mutable state (re-entrant, one stream per thread is fine);
 *  - return value: 0 = success; negative = Y3D_E* argument error (nothing was enqueued); positive = cudaError_t
 *    of a failed launch; y3d_strerror() names both;
 *  - fp32 only (the reference path is fp32; AMP callers up-cast), strides are in ELEMENTS;
 *  - tie-breaking everywhere: lowest index first;
 *  - workspace: query y3d_workspace_bytes(); contents need no initialisation; 256-byte alignment required.
 */
#ifndef Y3D_H_
#define Y3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Y3D_MAX_LEVELS 4
#define Y3D_MAX_TOPK 32      /* TAL top-k per GT (reference uses 10 / 13 / 8 / 1) */
#define Y3D_MAX_DET 1024     /* max_det of the NMS-free top-k (reference: 300 / 50) */

enum {
    Y3D_OK = 0,
    Y3D_EINVAL = -1,      /* bad shape / null pointer / bad flag */
    Y3D_EUNSUPPORTED = -2, /* config outside the compiled limits (levels, topk, max_det, reg_max ...) */
    Y3D_EALIGN = -3,      /* pointer / stride alignment the kernels need is not met */
    Y3D_EWORKSPACE = -4   /* workspace missing or too small */
};

enum { /* `stage` of y3d_workspace_bytes */
    Y3D_STAGE_POSTPROCESS = 1,
    Y3D_STAGE_TAL_ASSIGN = 2,
    Y3D_STAGE_V8_LOSS = 3,
    Y3D_STAGE_TAL_ASSIGN3D = 4,
    Y3D_STAGE_DECODE_TOPK = 5,
    Y3D_STAGE_DD_LOSS = 6
};

const char *y3d_strerror(int rc);
int y3d_abi_version(void);

/* bytes of scratch a stage needs for the given sizes (unused dimensions may be 0) */
size_t y3d_workspace_bytes(int stage, int B, int A, int nc, int M, int k, int D);

/* ---------------------------------------------------------------------------------------------------------
 * Head geometry.  The head emits one tensor per level, [B, C, h_l, w_l] (ultralytics/nn/modules/head.py:80-85).
 * Level l, image b, channel c, cell i (= y*w_l + x) is at lvl_ptr[l][b*lvl_sB[l] + c*lvl_sC[l] + i].
 * A concatenated x_cat [B, C, A] (head.py:56) is described with lvl_ptr[l] = x_cat + start_l, sB = C*A, sC = A.
 * Anchors are never read: anchor (x+0.5, y+0.5) and stride are recomputed from the cell index
 * (make_anchors, ultralytics/utils/tal.py:300-312).
 * ------------------------------------------------------------------------------------------------------- */

/* Detect.inference (head.py:53-79) = DFL softmax-integral (block.py:59-62) + dist2bbox (tal.py:315-325) * stride,
 * cat with sigmoid(cls).  y: [B, 4+nc, A] contiguous; xywh=1 normal path, xywh=0 the export path (head.py:107). */
int y3d_decode2d(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC, const int *lvl_hw,
                 const float *lvl_stride, int nl, int B, int nc, int reg_max, int xywh, float *y, void *stream);

/* ops.v10postprocess (ultralytics/utils/ops.py:852-865) when scores_first=0, nreg=4: preds[...,:4] boxes,
 * preds[...,4:] scores;  ops.v10_3Dpostprocess (ops.py:867-880) when scores_first=1, nreg=35.
 * preds element (b,a,ch) at preds[b*sB + a*sA + ch*sC] -- the reference passes a permuted view, so strides are
 * honoured instead of forcing a copy.  Outputs: reg [B,D,nreg], scores [B,D], labels [B,D] int64 (reference dtype),
 * anchor_idx [B,D] int32 (optional, may be NULL: the anchor each detection came from).  Requires 1 <= D <= min(A, Y3D_MAX_DET). */
int y3d_postprocess(const float *preds, int64_t sB, int64_t sA, int64_t sC, int B, int A, int nc, int nreg,
                    int scores_first, int D, float *reg, float *scores, int64_t *labels, int32_t *anchor_idx,
                    void *ws, size_t ws_bytes, void *stream);

/* v10Detect.forward export branch (head.py:526-531): decode (xyxy) + v10postprocess fused; the class scores of
 * all anchors are never written and the box channels are only decoded for the D winners.
 * out: [B, D, 6] = x1 y1 x2 y2 score label(as float); anchor_idx optional [B,D] int32. */
int y3d_decode_topk2d(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC, const int *lvl_hw,
                      const float *lvl_stride, int nl, int B, int nc, int reg_max, int xywh, int D, float *out,
                      int32_t *anchor_idx, void *ws, size_t ws_bytes, void *stream);

/* Image-sharded v10Detect export branch (BASELINE.json configs[3]): y3d_decode_topk2d on this
 * rank's B images, with the gather of the detections fused into its last kernels -- the box-decode kernel stores every
 * finished [D,6] row straight into every peer's gather buffer over NVLink (peer memory), value and call number in one
 * 64-bit word, so neither a fence nor a flag follows the data; a small kernel (one CTA per remote image) unpacks the
 * peers' words of this call into the result as they arrive.
 *  peer_bufs: HOST array of `world` device pointers, buffer r being rank r's (peer-mapped, e.g. symmetric memory), each
 *  y3d_gather_buffer_bytes(world, B, D) bytes, zero-filled before the first call: [2][world*B][D][6] floats (the result)
 *  followed by [2][world*B][6 D] 64-bit staging words.  After the call (stream order) rank `rank`'s buffer holds, at parity
 *  seq & 1, the detections of all ranks in rank order -- equal to all_gather of the ranks' y3d_decode_topk2d outputs;
 *  valid for consumers enqueued before this rank's next call.  seq: 1, 2, ... the same on every rank, +1 per call
 *  (must NOT be replayed from a CUDA graph; the low 32 bits must not be 0); every rank must make every call; a
 *  rank waits up to Y3D_XRANK_TIMEOUT_S seconds (default 600) and then sets *status (optional DEVICE int) to 1.
 *  Reference consumer: ultralytics/models/yolov10/val.py:10-24 (per-rank postprocess; DDP validation gathers). */
size_t y3d_gather_buffer_bytes(int world, int n_local, int D);
int y3d_decode_topk2d_sharded(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC, const int *lvl_hw,
                              const float *lvl_stride, int nl, int B, int nc, int reg_max, int xywh, int D, int rank,
                              int world, void *const *peer_bufs, unsigned long long seq, int *status, void *ws,
                              size_t ws_bytes, void *stream);

/* TaskAlignedAssigner.forward (ultralytics/utils/tal.py:44-94; get_pos_mask :96, get_box_metrics :108,
 * select_topk_candidates :133, select_highest_overlaps :237, get_targets :169) with CIoU from
 * ultralytics/utils/metrics.py:78-134.  M >= 1 (the M == 0 early-out of tal.py:68-76 is host-side glue).
 *  pd_scores [B,A,nc] (already sigmoid; element strides ss_B, ss_A, ss_C), pd_bboxes [B,A,4] xyxy px contiguous,
 *  anc_points [A,2] px, gt_labels [B,M], gt_bboxes [B,M,4] xyxy px, mask_gt [B,M] (0/1 floats).
 *  lvl_hw / lvl_stride (HOST, may be NULL): when given, promise that anc_points == make_anchors(levels)*stride so
 *  the candidate search walks each GT's rectangle instead of all A anchors (same results, less work).
 * outputs (reference dtypes): target_labels [B,A] int64, target_bboxes [B,A,4], target_scores [B,A,nc],
 *  fg_mask [B,A] uint8 (0/1; torch.bool storage), target_gt_idx [B,A] int64.  Any of target_labels / target_bboxes /
 *  target_scores may be NULL to skip that write. */
int y3d_tal_assign(const float *pd_scores, int64_t ss_B, int64_t ss_A, int64_t ss_C, const float *pd_bboxes,
                   const float *anc_points, const float *gt_labels, const float *gt_bboxes, const float *mask_gt,
                   int B, int A, int nc, int M, int topk, float alpha, float beta, float eps, const int *lvl_hw,
                   const float *lvl_stride, int nl, int64_t *target_labels, float *target_bboxes,
                   float *target_scores, uint8_t *fg_mask, int64_t *target_gt_idx, void *ws, size_t ws_bytes,
                   void *stream);

/* v8DetectionLoss.preprocess (loss.py:180-195, with the xywh2xyxy of ops.py:403-422) and DDDetectionLoss.preprocess
 * (loss.py:795-810): ragged ground truth -> the padded tensor the loss entry points take.
 *  batch_idx [N] (image index as float, like the dataloader's), cls [N], bboxes [N,4] xywh normalised to [0,1],
 *  extra [N, n_extra] (optional columns copied behind the box, e.g. the twelve 3D columns) -- all DEVICE, fp32.
 *  out [B, M, 5 + n_extra] = cls, xyxy in pixels (img_w, img_h = feature size * stride, loss.py:219), extra; rows of an
 *  image keep their order of appearance, rows beyond M are dropped, unused rows are zero.  counts [B] int32 (required):
 *  receives the number of rows of every image (M must be >= its maximum for nothing to be dropped; the Python mirror
 *  takes it from the host copy of batch_idx the dataloader already has, so no device sync is needed). */
int y3d_pack_targets(const float *batch_idx, const float *cls, const float *bboxes, const float *extra, int n_extra,
                     int N, int B, int M, float img_w, float img_h, float *out, int *counts, void *stream);

/* v8DetectionLoss.__call__ forward (ultralytics/utils/loss.py:206-257 with bbox_decode :197, BboxLoss :82-113)
 * for ONE branch, fused: head levels in, three loss items out; pd_scores / target_scores are never materialised.
 *  gt [B,M,5] = cls, xyxy px (the output of v8DetectionLoss.preprocess, loss.py:180-195; zero rows = padding), M >= 0.
 *  gains = hyp.box, hyp.cls, hyp.dfl.  loss_items: DEVICE float[4] = box, cls, dfl (after gains), and
 *  target_scores_sum (max(sum,1), loss.py:240).  partials (optional DEVICE double[4]): un-normalised
 *  sum (1-ciou)*w, sum bce, sum dfl*w, sum target_scores -- what a multi-GPU caller all-reduces before
 *  normalising; when `normalise` is 0 loss_items is left untouched.
 *  loss_total (optional DEVICE float[1], written when normalising): total_scale * (box + cls + dfl), summed over the
 *  branches -- with total_scale = batch size the reference's `loss.sum() * batch_size` (loss.py:257, 736), so that a
 *  forward-only caller needs no further kernel.
 *  v10DetectLoss (loss.py:727-737) = this with topk=10 on one2many + topk=1 on one2one.
 *  dbg_fg_mask [B,A] u8 / dbg_target_gt_idx [B,A] i32 (optional, NULL in production): the assignment the fused
 *  path used, for parity tests.
 *  prof_events (optional HOST array of 4 cudaEvent_t, NULL in production): recorded on `stream` before the first
 *  operation and after each kernel -- [0] start, [1] head streaming pass, [2] per-GT top-k + claims,
 *  [3] conflict resolution + foreground loss terms + final reduction -- so a benchmark can time each kernel inside
 *  the fused call. */
int y3d_v8_loss_fwd(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC, const int *lvl_hw,
                    const float *lvl_stride, int nl, int B, int nc, int reg_max, const float *gt, int M, int topk,
                    float gain_box, float gain_cls, float gain_dfl, int normalise, float *loss_items,
                    double *partials, float total_scale, float *loss_total, uint8_t *dbg_fg_mask,
                    int32_t *dbg_target_gt_idx, void *const *prof_events, void *ws, size_t ws_bytes, void *stream);

/* v10DetectLoss.__call__ forward (loss.py:727-737): BOTH branches of the consistent dual assignment in the same
 * launches -- one2many (top-k topk_o2m = 10) and one2one (top-k topk_o2o = 1) share the level geometry, the GT set
 * and the gains.  loss_items: DEVICE float[8] = (box, cls, dfl, target_scores_sum) x (one2many, one2one);
 * partials: DEVICE double[8] likewise; dbg_fg_mask [2,B,A] / dbg_target_gt_idx [2,B,A].  Other arguments as in
 * y3d_v8_loss_fwd.  total loss of the reference = (items[0..2] + items[4..6]).sum() * B. */
int y3d_v10_loss_fwd(const float *const *o2m_ptr, const int64_t *o2m_sB, const int64_t *o2m_sC,
                     const float *const *o2o_ptr, const int64_t *o2o_sB, const int64_t *o2o_sC, const int *lvl_hw,
                     const float *lvl_stride, int nl, int B, int nc, int reg_max, const float *gt, int M,
                     int topk_o2m, int topk_o2o, float gain_box, float gain_cls, float gain_dfl, int normalise,
                     float *loss_items, double *partials, float total_scale, float *loss_total, uint8_t *dbg_fg_mask,
                     int32_t *dbg_target_gt_idx, void *const *prof_events, void *ws, size_t ws_bytes, void *stream);

/* y3d_v10_loss_fwd on this rank's image shard of a batch that is spread over `world` GPUs of one node (SURVEY.md
 * section 8e), with the one cross-rank step fused into the last kernel: the CTA that finishes last pushes the rank's
 * 8 partial sums straight into every peer's exchange buffer over NVLink (peer memory), waits for the peers' flags,
 * sums in rank order and normalises by the batch-global target_scores_sum (loss.py:240) -- no separate collective
 * launch.  loss_items (required): DEVICE float[8], identical on every rank and equal to what one process computes on
 * the whole batch; partials (optional): DEVICE double[8], the sums over all ranks.
 *  peer_bufs_dev: DEVICE array of `world` pointers, entry r = rank r's exchange buffer (y3d_xrank_buffer_bytes(world)
 *  bytes, zero-initialised once, mapped into every peer: e.g. torch.distributed._symmetric_memory); seq: call counter,
 *  1, 2, 3, ... -- the same on every rank; status (optional): DEVICE int, 1 when a peer did not arrive within ~8 s
 *  (the items are NaN then).  world == 1 degenerates to y3d_v10_loss_fwd.  Every rank must make the call. */
int y3d_v10_loss_fwd_sharded(const float *const *o2m_ptr, const int64_t *o2m_sB, const int64_t *o2m_sC,
                             const float *const *o2o_ptr, const int64_t *o2o_sB, const int64_t *o2o_sC,
                             const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                             const float *gt, int M, int topk_o2m, int topk_o2o, float gain_box, float gain_cls,
                             float gain_dfl, float *loss_items, double *partials, float total_scale,
                             float *loss_total, int rank, int world, void *const *peer_bufs_dev,
                             unsigned long long seq, int defer, int *status, void *const *prof_events, void *ws,
                             size_t ws_bytes, void *stream);

/* The collecting half of the exchange on its own, for y3d_v10_loss_fwd_sharded(..., defer = 1): that call only POSTS
 * the rank's partial sums into the peers' buffers (loss_items / loss_total / partials may then be NULL and are not
 * written); this one waits for the peers' sums of call `seq`, adds them in rank order and writes loss_items [4 n],
 * loss_total (optional) and global_partials (optional) exactly as the undeferred call would have.  Launched later --
 * e.g. on a side stream behind an event, or right before the backward call -- the NVLink round trip and the wait for the
 * slowest rank overlap whatever ran in between.  Flow control is the caller's: a rank may post call j + 2 only after its
 * own resolve of call j + 1 has completed (two parity slots); dist.v10_loss_sharded(defer=True) does this with events.
 * peer_bufs: HOST array as in y3d_loss_allreduce_finalize. */
int y3d_loss_exchange_resolve(int n_branch, int rank, int world, void *const *peer_bufs, unsigned long long seq,
                              float gain_box, float gain_cls, float gain_dfl, float total_scale, float *loss_items,
                              float *loss_total, double *global_partials, int *status, void *stream);

/* Backward of y3d_v8_loss_fwd / y3d_v10_loss_fwd: what autograd produces in the reference for loss.py:206-257
 * (BCEWithLogits :240, BboxLoss.forward :82-96, _df_loss :99-113, bbox_decode :197-204; CIoU metrics.py:78-134 with
 * alpha under no_grad :128-129; the assigner, tal.py:44, is @torch.no_grad and therefore a constant).
 *  Call right after the forward call with the SAME geometry, gt, M, topk and the SAME, untouched workspace (the
 *  forward leaves, in the claim word of every foreground anchor, its assigned GT and alignment weight).
 *  grad_ptr / grad_sB / grad_sC: one gradient tensor per level, indexed like the head tensors; every element is
 *  written (box rows of background anchors are zero).  loss_items: DEVICE float[4 * n_branch] of the forward pass
 *  (after y3d_v8_loss_finalize in the multi-GPU case: the batch-global target_scores_sum must be in [4z+3]).
 *  grad_items: DEVICE float[3 * n_branch] = d total / d (box, cls, dfl) item of each branch; the reference's
 *  total = sum(items) * batch_size gives batch_size for each. */
int y3d_v8_loss_bwd(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC, float *const *grad_ptr,
                    const int64_t *grad_sB, const int64_t *grad_sC, const int *lvl_hw, const float *lvl_stride, int nl,
                    int B, int nc, int reg_max, const float *gt, int M, int topk, float gain_box, float gain_cls,
                    float gain_dfl, const float *loss_items, const float *grad_items, const void *ws, size_t ws_bytes,
                    void *stream);
int y3d_v10_loss_bwd(const float *const *o2m_ptr, const int64_t *o2m_sB, const int64_t *o2m_sC, float *const *o2m_grad,
                     const int64_t *o2m_gsB, const int64_t *o2m_gsC, const float *const *o2o_ptr, const int64_t *o2o_sB,
                     const int64_t *o2o_sC, float *const *o2o_grad, const int64_t *o2o_gsB, const int64_t *o2o_gsC,
                     const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max, const float *gt,
                     int M, int topk_o2m, int topk_o2o, float gain_box, float gain_cls, float gain_dfl,
                     const float *loss_items, const float *grad_items, const void *ws, size_t ws_bytes, void *stream);

/* The one cross-rank step of the image-sharded loss as a single kernel over NVLink peer memory:
 * all-reduce(sum) of the per-rank partials + the normalisation of y3d_v8_loss_finalize.
 *  peer_bufs: HOST array of `world` DEVICE pointers, peer_bufs[r] = rank r's exchange buffer of
 *  y3d_xrank_buffer_bytes(world) bytes, zero-initialised once, addressable from this device (symmetric / peer memory,
 *  e.g. torch.distributed._symmetric_memory); seq: call counter, 1, 2, 3, ... identical on all ranks; every rank must
 *  make the same sequence of calls.  partials: this rank's DEVICE double[4 * n_branch]; loss_items: DEVICE
 *  float[4 * n_branch] (identical on every rank afterwards); global_partials (optional): the summed partials;
 *  status (optional DEVICE int): 0, or 1 when a peer did not arrive within ~8 s (items are NaN then). */
size_t y3d_xrank_buffer_bytes(int world);
int y3d_loss_allreduce_finalize(const double *partials, int n_branch, int rank, int world, void *const *peer_bufs,
                                unsigned long long seq, float gain_box, float gain_cls, float gain_dfl,
                                float *loss_items, double *global_partials, int *status, void *stream);

/* v8DetectionLoss.bbox_decode (loss.py:197-204) + the permute/sigmoid of loss.py:214,232: head levels ->
 * pd_bboxes [B,A,4] xyxy in GRID units (caller multiplies by stride, loss.py:233) and, optionally (may be NULL),
 * pd_scores [B,A,nc] = sigmoid(class logits).  Same device arithmetic as the fused loss uses internally. */
int y3d_train_decode(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC, const int *lvl_hw,
                     const float *lvl_stride, int nl, int B, int nc, int reg_max, float *pd_bboxes, float *pd_scores,
                     void *stream);

/* turns all-reduced partials (see above) into loss items: DEVICE double[4*n_branch] -> DEVICE float[4*n_branch] */
int y3d_v8_loss_finalize(const double *partials, int n_branch, float gain_box, float gain_cls, float gain_dfl,
                         float *loss_items, void *stream);

/* v10Detect3d.decode (head.py:755-764): [B, nc+35, A] levels -> y [B, nc+35, A] contiguous
 * (cls logits | bbox xyxy px | center3d px | s3d | hd(24) | dep | dep_un). */
int y3d_decode3d(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC, const int *lvl_hw,
                 const float *lvl_stride, int nl, int B, int nc, float *y, void *stream);

/* Sparse evaluation of the 3D head at inference (v10Detect3d.inference_forward_feat, head.py:689-716): the class head
 * runs densely, every other head only on K = max_det candidate cells per image and level.
 *  y3d_select_candidates  select_candidates (head.py:681-687) + unravel_index (:652-657): cls = one level's class logits
 *      [B, nc, H, W] (element strides sB, sC; rows contiguous) -> idx [B, K, 2] int64 = (row, col) of the K cells with
 *      the largest max_c logit, descending, lowest flat index first among equals.  1 <= K <= min(H*W, Y3D_MAX_DET).
 *  y3d_extract_patches    extract_patches (head.py:659-679): x [B, C, H, W] -> out [B*K, C, P, P], the zero-padded P x P
 *      (P odd, the reference uses 5) neighbourhood of every candidate.
 *  y3d_scatter_candidates head.py:709-714: vals [B*K, Cout] (the 1x1 output of a head evaluated on the patches) ->
 *      out [B, Cout, H, W], zero except out[b, :, row_k, col_k] = vals[b*K + k, :]. */
int y3d_select_candidates(const float *cls, int64_t sB, int64_t sC, int B, int nc, int H, int W, int K, int64_t *idx,
                          void *stream);
int y3d_extract_patches(const float *x, int64_t sB, int64_t sC, const int64_t *idx, int B, int C, int H, int W, int K,
                        int P, float *out, void *stream);
int y3d_scatter_candidates(const float *vals, const int64_t *idx, int B, int Cout, int H, int W, int K, float *out,
                           void *stream);

/* rotate_iou_gpu_eval (ultralytics/data/datasets/kitti_eval.py:309-344; kernel :263-301, device functions :60-260):
 * overlap of rotated bird's-eye-view boxes for the KITTI evaluator (bev_box_overlap :458, box3d_overlap :503).
 *  boxes [N,5], query_boxes [K,5] = (cx, cy, dx, dy, angle), DEVICE fp32; iou [N,K] fp32.
 *  criterion -1: IoU; 0: intersection / area(query); 1: intersection / area(box); 2: intersection area. */
int y3d_rotate_iou_eval(const float *boxes, int N, const float *query_boxes, int K, int criterion, float *iou,
                        void *stream);

/* KITTIDataset.decode_preds (ultralytics/data/datasets/kitti.py:519-576; bin2angle decode_helper.py:12,
 * img_to_rect kitti_utils.py:241, alpha2ry :311, affine_transform :467), undo_augment=True.
 *  dets [B,D,37] fp32 = bbox(4) c3d(2) s3d(3) hd(24) dep un score(logit) label (yolov10_3D/val.py:46-47);
 *  calib [B,6] = cu cv fu fv tx ty, inv_affine [B,2,3], ratio [B,2], cls_mean_size [nc,3] -- all float64 like the
 *  reference's numpy math.  rows [B,D,14] float64 = cls alpha x1 y1 x2 y2 h w l x y z ry score; valid [B,D] uint8
 *  = !(score < threshold). */
int y3d_decode_preds3d(const float *dets, int B, int D, int nc, const double *calib, const double *inv_affine,
                       const double *ratio, const double *cls_mean_size, double threshold, double *rows,
                       uint8_t *valid, void *stream);

/* TaskAlignedAssigner3d.forward (ultralytics/utils/tal.py:391-452; keypoints utils/keypoint_utils.py:11-118).
 *  pd_scores [B,A,nc] contiguous (sigmoid), pd_bboxes [B,A,4] px, pd_3d [B,A,31], anc_points [A,2] px,
 *  stride [A], gts [B,M,17] packed (label bbox4 c2d2 s2d2 c3d2 s3d3 depth hbin hres), mask_gt [B,M],
 *  calibs [B,6], mean_sizes [nc,3].  flags: bit0 use_2d, bit1 use_3d, bit2 kps l2 metric, bit3 constrain_anchors.
 * outputs: target_labels [B,A] i64, target_scores [B,A,nc], target_vals [B,A,12] (c2d2 s2d2 c3d2 s3d3 dep hbin hres),
 *  fg_mask [B,A] u8, target_gt_idx [B,A] i64, pd_keypoints [B,A,8,3], gt_keypoints [B,M,8,3]. */
int y3d_tal_assign3d(const float *pd_scores, const float *pd_bboxes, const float *pd_3d, const float *anc_points,
                     const float *stride, const float *gts, const float *mask_gt, const float *calibs,
                     const float *mean_sizes, int B, int A, int nc, int M, int topk, float alpha, float beta,
                     float gamma, float eps, int flags, const int *lvl_hw, const float *lvl_stride, int nl,
                     int64_t *target_labels, float *target_scores, float *target_vals, uint8_t *fg_mask,
                     int64_t *target_gt_idx, float *pd_keypoints, float *gt_keypoints, void *ws, size_t ws_bytes,
                     void *stream);

/* DDDetectionLoss.__call__ forward (ultralytics/utils/loss.py:821-900; bbox_decode :812-819, compute_box2d_loss
 * :913-926, compute_box3d_loss :928-963, laplacian_aleatoric_uncertainty_loss_new :1112-1119, compute_heading_loss
 * :1122-1136) for ONE branch of DetectLoss3d (loss.py:741-771), fused with TaskAlignedAssigner3d (tal.py:391-700).
 *  head levels [B, nc+35, h, w] = cls(nc) | o2d(2) s2d(2) o3d(2) s3d(3) hd(24) dep dep_un;
 *  gts [B,M,17] = label, bbox xyxy px, center_2d(2), size_2d(2), center_3d(2), size_3d(3), depth, heading_bin,
 *  heading_res -- the output of DDDetectionLoss.preprocess (loss.py:795-810), zero rows = padding; calibs [B,6],
 *  mean_sizes [nc,3] (DEVICE).  flags as y3d_tal_assign3d.  gains: HOST float[6] = hyp.loss2d, cls, depth, offset3d,
 *  size3d, heading.  loss_items: DEVICE float[8] = the six loss items, target_scores_sum, number of foreground anchors
 *  (all zero when M == 0, as the reference's early return).  partials: DEVICE double[11] (required) = un-normalised
 *  sums: softplus, |offset2d|, |size2d|, depth, |offset3d|, |size3d|, heading CE, heading L1, x[label]*t,
 *  sum target_scores, n_fg -- what a multi-GPU caller all-reduces before y3d_dd_loss_finalize.
 *  dbg_target_gt_idx (optional) [B,A] int32: the assigned GT of every anchor, -1 = background.
 *  The distillation / foreground-depth-map terms of the fork (loss.py:792,890-895) are out of scope. */
int y3d_dd_loss_fwd(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC, const int *lvl_hw,
                    const float *lvl_stride, int nl, int B, int nc, const float *gts, int M, const float *calibs,
                    const float *mean_sizes, int topk, float alpha, float beta, float gamma, int flags,
                    const float *gains, int normalise, float *loss_items, double *partials,
                    int32_t *dbg_target_gt_idx, void *ws, size_t ws_bytes, void *stream);
/* Both branches of DetectLoss3d (loss.py:750-771) in the SAME four launches (streaming pass, top-k, conflict resolution,
 * foreground terms; the branch is a grid dimension): one2many with topk_o2m, one2one with topk_o2o.  Arguments as in
 * y3d_dd_loss_fwd; loss_items: DEVICE float[2][8], partials: DEVICE double[2][11], dbg_target_gt_idx [2,B,A] (optional);
 * ws: 2 x y3d_workspace_bytes(Y3D_STAGE_DD_LOSS, ...) bytes -- branch z owns the z-th half, which is what
 * y3d_dd_loss_bwd takes for that branch. */
int y3d_dd_loss_dual_fwd(const float *const *o2m_ptr, const int64_t *o2m_sB, const int64_t *o2m_sC,
                         const float *const *o2o_ptr, const int64_t *o2o_sB, const int64_t *o2o_sC, const int *lvl_hw,
                         const float *lvl_stride, int nl, int B, int nc, const float *gts, int M, const float *calibs,
                         const float *mean_sizes, int topk_o2m, int topk_o2o, float alpha, float beta, float gamma,
                         int flags, const float *gains, int normalise, float *loss_items, double *partials,
                         int32_t *dbg_target_gt_idx, void *ws, size_t ws_bytes, void *stream);
int y3d_dd_loss_finalize(const double *partials, int M, const float *gains, float *loss_items, void *stream);
/* Backward of y3d_dd_loss_fwd (autograd of loss.py:879-888 through compute_box2d_loss, compute_box3d_loss, the
 * Laplacian depth term and compute_heading_loss): call with the same geometry / gts / M and the untouched workspace of
 * the forward call.  grad_* index the gradient tensors like the head tensors; every element is written.
 * loss_items: DEVICE float[8] of the forward pass; grad_items: DEVICE float[6] = d total / d item. */
int y3d_dd_loss_bwd(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC, float *const *grad_ptr,
                    const int64_t *grad_sB, const int64_t *grad_sC, const int *lvl_hw, const float *lvl_stride, int nl,
                    int B, int nc, const float *gts, int M, const float *gains, const float *loss_items,
                    const float *grad_items, const void *ws, size_t ws_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* Y3D_H_ */
