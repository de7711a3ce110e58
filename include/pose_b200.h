/*
 * pose_b200.h -- C ABI of the B200-native heatmap hot path (render -> loss -> decode).
 *
 * Drop-in boundary for the hot path of myungsanglee/PyTorch-Pose-Estimation.  The reference has
 * no FFI of its own (it is pure Python); each entry point below names the reference function it
 * replaces (file:line in the reference tree).  The Python side (pose_b200/_cabi.py) binds these
 * with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions for every compute entry point:
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - asynchronous on `stream` (a cudaStream_t passed as void*); no allocation, no host
 *     synchronisation; thread-safe, one library shared by all ranks.  The only state kept between calls is a
 *     mutex-protected cache of launch configurations (occupancy, shared-memory opt-in) per (kernel, device) and a
 *     launch counter: a steady-state call makes no driver query and reads no environment variable;
 *   - the caller owns all memory, including workspaces (sizes from the *_workspace_bytes calls);
 *   - returns 0 on success, <0 for a bad argument (POSE_E*), >0 a cudaError_t from a launch;
 *     pose_b200_last_error() gives a thread-local message for the last non-zero return;
 *   - tensors are dense, C-contiguous, fp32 unless stated; maps are [N][K][H][W].
 */
#ifndef POSE_B200_H
#define POSE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define POSE_OK 0
#define POSE_EINVAL (-1)      /* bad shape / null pointer / unsupported parameter */
#define POSE_EALIGN (-2)      /* pointer not aligned as required                   */
#define POSE_EWORKSPACE (-3)  /* workspace too small                               */

/* kp_dtype values */
#define POSE_KP_F32 0
#define POSE_KP_F64 1

/* flags for pose_sbp_fused / pose_spm_fused */
#define POSE_F_GRAD 1u          /* write dlogits                                   */
#define POSE_F_TARGET_OUT 2u    /* also materialise the rendered target (render mode only) */
#define POSE_F_DECODE 4u        /* also decode joints from sigmoid(logits) in the same pass */
#define POSE_F_TMA 8u           /* stage the maps through shared memory with bulk async copies (render mode, 16-byte aligned maps): the fast path */
#define POSE_F_SIGMOID_CUDA 16u /* POSE_F_DECODE: rank with POSE_SIGMOID_ATEN_CUDA instead of POSE_SIGMOID_ATEN_CPU */
#define POSE_F_HEAD_LOGITS_OUT 32u   /* pose_sbp_head_fused: also write the logits (tests / debugging: the point of the call is not to) */
#define POSE_F_HEAD_NO_RESIDUAL 64u  /* pose_sbp_head_fused: skip the feature residual (plain TF32 features: 2^-11 relative error; diagnostics) */

/* Which torch.sigmoid the decoders reproduce BIT FOR BIT when they rank near-equal logits and report the confidence.
 * "First row-major index of the largest sigmoid value" (nms_sbp utils/sbp_utils.py:73-78, nms_spm utils/spm_utils.py:112-115)
 * depends on it: fp32 sigmoid is many-to-one and implementations one ulp apart merge different neighbours.
 *   ATEN_CPU : 1/(1+Sleef_expf_u10(-x)) -- the reference run on CPU tensors (the golden vectors)
 *   ATEN_CUDA: 1/(1+expf(-x)), libdevice expf, IEEE divide -- the reference run on CUDA tensors (the Lightning modules) */
#define POSE_SIGMOID_ATEN_CPU 0
#define POSE_SIGMOID_ATEN_CUDA 1

typedef void* pose_stream_t;
struct pose_exchange;

/* ---- library ------------------------------------------------------------------------------- */
int pose_b200_version(void);
/* sha256 (hex) of the sources + compiler flags this binary was built from; the Python loader refuses a library whose
 * hash differs from the tree it sits in (a stale binary must not pass for a fresh one) */
const char* pose_b200_source_hash(void);
const char* pose_b200_last_error(void);
/* kernels launched by this library since load (statistics only; relaxed atomic) */
unsigned long long pose_b200_launch_count(void);

/* ---- Gaussian template (host) -- SBPHeatmapGenerator.__init__ utils/sbp_utils.py:21-31,
 *      SPMHeatmapGenerator.__init__ utils/spm_utils.py:17-27.
 * Writes n*n fp32 values (n = number of samples of arange(0, 6*sigma+3)) row-major to out_host,
 * computed in double and rounded once to fp32.  Returns n, or POSE_EINVAL if capacity < n*n. */
int pose_gauss_template_host(double sigma, float* out_host, int capacity);
/* The same template in the layout pose_sbp_fused reads: n+1 rows of n+8 floats -- 4 zeros, the n values, 4 zeros -- and
 * an all-zero last row (a thread then fetches the targets of 4 consecutive pixels with two clamped indices, no range
 * branches).  Returns n, or POSE_EINVAL if capacity < (n+1)*(n+8). */
int pose_gauss_template_padded_host(double sigma, float* out_host, int capacity);

/* ---- SBP render -- SBPHeatmapGenerator.__call__ utils/sbp_utils.py:33-53, batched.
 * kp [N][K][2] (x, y) in heat-map pixels, fp32 or fp64 (kp_dtype); x<0 or y<0 = invisible.
 * lut: device copy of the n*n template.  target [N][K][H][W] is fully overwritten. */
int pose_sbp_render(const void* kp, int kp_dtype, float* target, int N, int K, int H, int W,
                    double sigma, const float* lut, int lut_n, pose_stream_t stream);

/* ---- SBP loss (+grad, +render, +decode) in one pass --
 *      SBPLoss.forward/encode_target models/loss/sbp_loss.py:20-66 (+ autograd backward),
 *      SBPHeatmapGenerator.__call__ utils/sbp_utils.py:33-53 when kp != NULL,
 *      DecodeSBP.forward/nms_sbp utils/sbp_utils.py:56-118 when POSE_F_DECODE.
 * Exactly one of `target_in` (dense target, [N][K][H][W]) or `kp` (keypoints, rendered in
 * registers) must be non-NULL.  With kp, `lut` is the device copy of the PADDED template
 * (pose_gauss_template_padded_host: (lut_n+1) x (lut_n+8) floats).
 *   loss      = (lambda_pos * S_pos + lambda_neg * S_neg) * inv_norm,    inv_norm = 1/(2*K*B_global)
 *   dlogits   = dloss/dlogits (scaled by inv_norm; written iff POSE_F_GRAD)
 *   loss_out  [1] fp32; loss_num_out [2] fp64 = (S_pos, S_neg) un-normalised (may be NULL)
 *   joints    [N][K][3] fp32 (x*scale, y*scale, conf) / (-scale,-scale,-1)   iff POSE_F_DECODE
 *   bbox [N][4] fp64 + packed_out [N][3K+1] (both or neither; needs POSE_F_DECODE): the epilogue launch also
 *             back-projects the joints exactly like pose_sbp_backproject (SBPmAPCOCO.update_state :141-163)
 *   exchange (may be NULL; needs bbox): multi-GPU -- see pose_exchange_t below; packed_out may then be NULL
 * Two launches: the fused streaming kernel (one CTA per heat map), then one epilogue grid (fixed-order two-level loss
 * reduction + back-projection) issued with programmatic dependent launch.
 * workspace: pose_sbp_fused_workspace_bytes(N, K) -- one fp64 (S_pos, S_neg) pair per heat map + the reduction's slice
 * sums and counter; contents need no initialisation. */
unsigned long long pose_sbp_fused_workspace_bytes(int N, int K);
int pose_sbp_fused(const float* logits, const float* target_in,
                   const void* kp, int kp_dtype, double sigma, const float* lut, int lut_n,
                   float* dlogits, float* target_out,
                   float* loss_out, double* loss_num_out,
                   float* joints, float conf_threshold, float coord_scale,
                   int N, int K, int H, int W,
                   float lambda_pos, float lambda_neg, double inv_norm,
                   unsigned flags,
                   const double* bbox, float* packed_out, int input_h, int input_w,
                   const struct pose_exchange* exchange,
                   void* workspace, unsigned long long workspace_bytes,
                   pose_stream_t stream);

/* ---- multi-GPU exchange over peer-mapped (symmetric) memory -- replaces NCCL for the per-step all-gather.
 * Every rank owns one exchange buffer of pose_exchange_layout() bytes, zero-initialised, mapped into every other
 * rank's address space (e.g. torch.distributed._symmetric_memory); peer_base[r] is rank r's buffer as seen from THIS
 * process.  Passing the descriptor to pose_sbp_fused makes the epilogue store rows / loss numerators / ids into every
 * rank's receive region and raise a per-rank flag; pose_exchange_finish waits for all flags of the step and reduces
 * the gathered numerators (rank order) into the global-batch loss.  Receive regions form a ring of
 * POSE_EXCHANGE_SLOTS steps; the rows ([world*batch_local][row_stride] fp32, the first 3K+1 of each valid) and ids
 * of completed step c are at pose_exchange_t.off_rows/off_ids[c % POSE_EXCHANGE_SLOTS].  With defer = 1 there is
 * no per-step finish launch at all: a step is published by the next step's fused kernel and completed inside the next
 * step's epilogue (ranks may drift up to two steps apart); pose_exchange_flush completes the last one. */
#define POSE_MAX_PEERS 16
#define POSE_EXCHANGE_SLOTS 4
typedef struct pose_exchange {
    int world, rank;
    int batch_local, num_keypoints;
    int row_stride;                           /* floats per exchanged row: 3K+1 rounded up to a multiple of 4 (set by _layout) */
    int defer;                                /* 0: lock-step, pose_exchange_finish(s) completes step s;
                                                 1: in-band -- the fused kernel of step s+1 publishes step s, the epilogue of
                                                    step s+1 completes it (writes its loss to loss_prev); no finish launch */
    void* peer_base[POSE_MAX_PEERS];
    unsigned long long off_ctrl, off_flags, off_rows[POSE_EXCHANGE_SLOTS], off_nums[POSE_EXCHANGE_SLOTS], off_ids[POSE_EXCHANGE_SLOTS];
    const long long* ids_local;               /* device [batch_local][2] (image_id, category_id) of this rank */
    void* multicast_base;                     /* NVLS multicast alias of the buffer on all ranks (multimem.st), or NULL */
    float* loss_prev;                         /* defer = 1: device scalar receiving the global loss of the previous step */
} pose_exchange_t;
/* fills the off_* fields of *x from world / batch_local / num_keypoints and returns the buffer size in bytes */
unsigned long long pose_exchange_layout(pose_exchange_t* x);
int pose_exchange_finish(const pose_exchange_t* x, double w0, double w1, double inv_norm, float* loss_out,
                         pose_stream_t stream);
int pose_exchange_flush(const pose_exchange_t* x, double w0, double w1, double inv_norm, float* loss_out,
                        pose_stream_t stream);

/* ---- fixed-order reduction of n (a, b) fp64 pairs, `stride` doubles apart: loss = (w0*A + w1*B)*inv_norm.
 *      Used on the all-gathered per-rank numerators (multi-GPU global loss); one CTA, deterministic. */
int pose_loss_reduce(const double* pairs, int n, long long stride, double w0, double w1, double inv_norm,
                     float* loss_out, double* num_out, pose_stream_t stream);

/* ---- in-place dlogits *= *grad_output (device scalar), skipped in-kernel when it is 1.0f --
 *      autograd backward of the 0-dim loss (module/sbp_detector.py:24-28 -> loss.backward()). */
int pose_scale_grad(float* dlogits, const float* grad_output, unsigned long long n, pose_stream_t stream);

/* ---- SBP decode -- nms_sbp utils/sbp_utils.py:56-82 + DecodeSBP.forward :103-118, batched.
 * x [N][K][H][W]; joints [N][K][3].  apply_sigmoid = DecodeSBP.pred.  Both coordinates are
 * multiplied by coord_scale (= input_w / W, :116), undetected rows are (-1,-1,-1)*scale on x,y.
 * sigmoid_ref (apply_sigmoid only): POSE_SIGMOID_ATEN_CPU / _CUDA -- argmax indices and confidences are bit-identical to
 * the reference evaluated with that torch.sigmoid.  refine: bit 0 (POSE_DEC_REFINE) adds the quarter-pixel shift (NOT in the
 * reference; off by default); bit 1 (POSE_DEC_NO_TMA) selects the register-staged kernel instead of the bulk-async (TMA)
 * staged one, which is used whenever the maps are 16-byte aligned (same results bit for bit; a GPU test compares them). */
#define POSE_DEC_REFINE 1
#define POSE_DEC_NO_TMA 2
int pose_sbp_decode(const float* x, float* joints, int N, int K, int H, int W,
                    float conf_threshold, int apply_sigmoid, float coord_scale,
                    int refine, int sigmoid_ref, pose_stream_t stream);

/* ---- SBP decode with flip-test averaging -- NOT in the reference (PARITY UNPINNED, opt-in).
 * x, x_flip [N][K][H][W]: the maps of the image and of its horizontal mirror.  flip_perm [K] int32 (device): channel
 * of x_flip holding joint k's mirrored map (left/right swap).  The decoded map is
 *   heat[n][k][h][w] = 0.5 * (act(x[n][k][h][w]) + act(x_flip[n][flip_perm[k]][h][W-1-w]))
 * evaluated on the fly; everything else (threshold, first-index tie-break, output rows, refine) as pose_sbp_decode. */
int pose_sbp_decode_flip(const float* x, const float* x_flip, const int* flip_perm, float* joints,
                         int N, int K, int H, int W, float conf_threshold, int apply_sigmoid,
                         float coord_scale, int refine, pose_stream_t stream);

/* ---- back-projection + COCO row fields -- SBPmAPCOCO.update_state utils/sbp_utils.py:141-163,
 *      SBPmAPPIS.update_state utils/sbp_pis_utils.py:23-45.
 * joints [N][K][3] (input-size scale, from decode); bbox [N][4] fp64 (x, y, w, h).
 * packed_out [N][3K+1]: K rows of (x_img, y_img, 1) for detected joints / (0,0,0) otherwise, followed by
 * the score = left-to-right fp32 sum of detected confidences / K. */
int pose_sbp_backproject(const float* joints, const double* bbox, float* packed_out,
                         int N, int K, int input_h, int input_w, pose_stream_t stream);

/* ---- SPM render -- SPMHeatmapGenerator/MaskGenerator/DisplacementGenerator
 *      utils/spm_utils.py:16-95 + concat dataset/spm_coco_dataset.py:77-86, batched.
 * centers [N][Pmax][2] int64, joints [N][Pmax][K][2] int64, counts [N] int32 (persons per image).
 * target [N][1+2K][R][R] fully overwritten.
 * workspace (may be NULL): pose_spm_fused_workspace_bytes(N, K, R) bytes, 16-byte aligned -- with it (and Pmax <= 64) the
 * target is written in ONE pass by the render-only form of the fused path (a per-image geometry pre-pass + a pure write
 * stream); without it the older zero-fill + patch pair runs (any number of persons).  Same bits either way. */
int pose_spm_render(const long long* centers, const long long* joints, const int* counts,
                    float* target, int N, int Pmax, int K, int R, double sigma,
                    const float* lut, int lut_n, void* workspace, unsigned long long workspace_bytes,
                    pose_stream_t stream);

/* ---- SPM loss fwd(+bwd) -- SPMLoss.forward models/loss/spm_loss.py:23-105.
 * logits/target/dlogits [N][1+2K][R][R]; loss = (lambda_root*S_root + lambda_disp*S_disp)*inv_norm,
 * inv_norm = 1/B_global.  loss_num_out [2] fp64 = (S_root, S_disp).  Three launches: the root mask of every image as bits
 * (the displacement planes need m = t0 > 0 of their pixel), one CTA per 16 KB unit of one plane, the fixed-order reduction of
 * the per-unit loss pairs; the workspace holds the pairs (16 B per unit) and the mask bits (R*R/8 B per image). */
unsigned long long pose_spm_loss_workspace_bytes(int N, int K, int R);
int pose_spm_loss(const float* logits, const float* target, float* dlogits,
                  float* loss_out, double* loss_num_out, int N, int K, int R,
                  float lambda_root, float lambda_disp, double inv_norm, int write_grad,
                  void* workspace, unsigned long long workspace_bytes, pose_stream_t stream);

/* ---- SPM render + loss (+grad) in one pass -- SPMHeatmapGenerator/MaskGenerator/DisplacementGenerator
 *      utils/spm_utils.py:16-95 evaluated in registers and fed into SPMLoss.forward models/loss/spm_loss.py:23-105
 *      (+ autograd backward): the target never touches HBM.  Persons as for pose_spm_render (Pmax <= 64);
 *      logits/dlogits/target_out [N][1+2K][R][R].  flags: POSE_F_GRAD (write dlogits), POSE_F_TARGET_OUT (also
 *      materialise the rendered target, bit-identical to pose_spm_render).  loss = (lambda_root*S_root +
 *      lambda_disp*S_disp)*inv_norm, inv_norm = 1/B_global; loss_num_out [2] fp64 = (S_root, S_disp).
 *      Three launches: a geometry pre-pass (one CTA per image: person boxes, covered-quad bits, a byte of geometry per
 *      pixel -> workspace), the streaming kernel (one CTA per 16 KB of a channel plane, logits staged with bulk async
 *      copies) and the deterministic two-level loss reduction.
 *      workspace: pose_spm_fused_workspace_bytes(N, K, R) (loss pairs + ~20 KB of geometry per image at R = 128). */
unsigned long long pose_spm_fused_workspace_bytes(int N, int K, int R);
int pose_spm_fused(const float* logits, const long long* centers, const long long* joints, const int* counts,
                   float* dlogits, float* target_out, float* loss_out, double* loss_num_out,
                   int N, int Pmax, int K, int R, double sigma, const float* lut, int lut_n,
                   float lambda_root, float lambda_disp, double inv_norm, unsigned flags,
                   void* workspace, unsigned long long workspace_bytes, pose_stream_t stream);

/* ---- SPM decode -- nms_spm :98-161, get_spm_keypoints :164-200, DecodeSPM.forward :225-250.
 * x [N][1+2K][R][R]; roots [N][Pmax][3], kps [N][Pmax][K][3], counts [N] (roots found, capped at
 * Pmax; total found incl. overflow in counts_total if non-NULL).  Ties between equal root
 * confidences break row-major (the reference's order is undefined there).  apply_act = DecodeSPM.pred; the root
 * confidences -- which decide the threshold test and the greedy order -- are then the reference's torch.sigmoid bit for
 * bit (sigmoid_ref: POSE_SIGMOID_ATEN_CPU / _CUDA).  No workspace. */
int pose_spm_decode(const float* x, float* roots, float* kps, int* counts, int* counts_total,
                    int N, int Pmax, int K, int R, float conf_threshold, double dist_threshold,
                    int apply_act, int sigmoid_ref, float input_size, pose_stream_t stream);

/* ---- SPM image-size rescale -- SPMmAPCOCO.update_state utils/spm_utils.py:302-304.
 * kps [N][Pmax][K][3] at input scale (from pose_spm_decode), counts [N], image_w / image_h [N] int64 ->
 * out [N][Pmax][K][3]: x * fp32(w / input_size), y * fp32(h / input_size), conf; rows >= counts[i] are not written.
 * out may alias kps. */
int pose_spm_rescale(const float* kps, const int* counts, const long long* image_w, const long long* image_h,
                     float* out, int N, int Pmax, int K, float input_size, pose_stream_t stream);

/* ---- SPM joint gather -- get_spm_keypoints utils/spm_utils.py:164-200 (no rescale).
 * roots [n][3] (x, y, conf) in map pixels, disp [2K][R][R] (already activated) -> kps [n][K][3]. */
int pose_spm_gather(const float* roots, const float* disp, float* kps, int n_roots, int K, int R,
                    double dist_threshold, pose_stream_t stream);

/* ---- SPM joint gather with hierarchical displacement chaining -- NOT in the reference (get_spm_keypoints reads every joint at
 *      the root pixel, utils/spm_utils.py:187-189); SURVEY 8 f-4, opt-in, parity unpinned.  parent [K] int32 on the device:
 *      parent[k] = the joint k's displacement is relative to (-1 = the root).  Joint k is read at its parent's decoded position
 *      (truncated to a pixel) and added to it with the arithmetic of get_spm_keypoints; an absent / off-map parent makes every
 *      descendant (0,0,0); chains longer than 16 are rejected as absent.  All parent[k] = -1: bit-identical to pose_spm_gather. */
int pose_spm_gather_chain(const float* roots, const float* disp, const int* parent, float* kps, int n_roots, int K, int R,
                          double dist_threshold, pose_stream_t stream);

/* ---- OKS / AP -- the keypoint evaluation behind SBPmAPCOCO.result utils/sbp_utils.py:166-189, SPMmAPCOCO.result
 *      utils/spm_utils.py:325-351 and SBPmAPPIS.result (utils/sbp_pis_utils.py), which the reference delegates to
 *      pycocotools COCOeval(gt, dt, "keypoints").evaluate()/.accumulate() (third party, unpinned: PARITY UNPINNED).
 * A group q = cat * n_images + img.  Detections [D] are sorted by (group, -score), at most max_det per group; ground
 * truths [G] sorted by group in annotation order.  det_off / gt_off [Q+1] int32 prefix offsets; pair_off [Q+1] int64
 * offsets of each group's D_q x G_q block in oks_out (row = detection).  All arithmetic is fp64.
 *   det_kp [D][K][3], gt_kp [G][K][3] (x, y, v); gt_bbox [G][4] (x, y, w, h); gt_area [G]; sigmas [K] (K <= 32).
 * pose_oks_matrix  = COCOeval.computeOks (+ the detection areas of COCO.loadRes -> det_area_out [D]). */
int pose_oks_matrix(const double* det_kp, const double* gt_kp, const double* gt_bbox, const double* gt_area,
                    const int* det_off, const int* gt_off, const long long* pair_off, const double* sigmas,
                    double* oks_out, double* det_area_out, int Q, int D, int G, long long n_pairs, int K,
                    pose_stream_t stream);

/* pose_oks_match = COCOeval.evaluateImg for every (group, area range, OKS threshold).
 *   gt_flags [G]: bit 0 = ignore (crowd or no labelled joint), bit 1 = iscrowd; area_rng [A][2]; iou_thrs [T].
 *   dt_match [A][T][D] int32 (global ground-truth index + 1, 0 = unmatched); dt_ignore [A][T][D]; gt_ignore_out [A][G].
 * workspace: pose_oks_match_workspace_bytes(A, T, G); contents need no initialisation. */
unsigned long long pose_oks_match_workspace_bytes(int A, int T, int G);
int pose_oks_match(const double* oks, const long long* pair_off, const int* det_off, const int* gt_off,
                   const double* det_area, const double* gt_area, const unsigned char* gt_flags,
                   const double* area_rng, const double* iou_thrs, int Q, int A, int T, int D, int G,
                   int* dt_match, unsigned char* dt_ignore, unsigned char* gt_ignore_out,
                   void* workspace, unsigned long long workspace_bytes, pose_stream_t stream);

/* pose_ap_accumulate = COCOeval.accumulate (single max_det).
 *   order [D] int64: detection indices sorted by (category, -score), stable, so that order[cat_det_off[c] ..
 *   cat_det_off[c+1]) lists category c's detections by descending score; cat_det_off / cat_gt_off [C+1]; rec_thrs [R].
 *   precision [T][R][C][A], recall [T][C][A] fp64, -1 where a category has no counted ground truth.
 * workspace: pose_ap_accumulate_workspace_bytes(A, T, D), 8-byte aligned. */
unsigned long long pose_ap_accumulate_workspace_bytes(int A, int T, int D);
int pose_ap_accumulate(const long long* order, const int* dt_match, const unsigned char* dt_ignore,
                       const unsigned char* gt_ignore, const int* cat_det_off, const int* cat_gt_off,
                       const double* rec_thrs, int C, int A, int T, int R, int D, int G,
                       double* precision, double* recall, void* workspace, unsigned long long workspace_bytes,
                       pose_stream_t stream);

/* ---- head fusion -- the detector's last layer, nn.Conv2d(512, num_keypoints, 1, 1, bias=False)
 *      models/detector/sbp.py:35-37,47, fused with what consumes its output: SBPLoss.forward models/loss/sbp_loss.py:20-66
 *      (+ autograd backward w.r.t. the logits), SBPHeatmapGenerator.__call__ utils/sbp_utils.py:33-53 (target rendered from
 *      kp), DecodeSBP.forward / nms_sbp utils/sbp_utils.py:56-118 and SBPmAPCOCO.update_state :141-163 -- the logits
 *      [N][K][H][W] never exist in HBM.
 *   features [N][C][H][W] fp32 NCHW (the head's input), weight [K][C] fp32 (conv weight [K,C,1,1]), kp / lut as in
 *   pose_sbp_fused except that `lut` is the UNPADDED n*n template (pose_gauss_template_host).  kp may be NULL with
 *   POSE_F_DECODE alone: the inference form (head -> DecodeSBP, inference_sbp.py:57-58,73), no target involved.
 *   The contraction runs on the tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, features through TMA tensor
 *   maps) with both operands split into a tf32 head and a residual, so the logits carry fp32-level error (csrc/head_kernels.cuh).
 *   Outputs as pose_sbp_fused (dlogits iff POSE_F_GRAD: the caller's autograd needs it for dW and dX; joints / packed_out iff
 *   POSE_F_DECODE); logits_out only with POSE_F_HEAD_LOGITS_OUT.  Decode: argmax in LOGIT space, first index among equal
 *   logits (all logits >= 17.5 tie: their sigmoid is 1.0f), confidence = the reference sigmoid of the maximum.
 *   Constraints (POSE_EINVAL otherwise): C % 32 == 0 and 48*C*4 bytes + 4 tiles must fit in shared memory (C <= 640),
 *   (H*W) % 128 == 0, K <= 17, template side <= 24.
 *   tuning: 0 = defaults; bits 0-7 feature stages in shared memory; bit 24 selects the variant that keeps the feature
 *   residuals in shared memory instead of tensor memory (bits 8-15: its residual stages) -- slower, kept for comparison.
 *   Three launches: weight split, the fused kernel (persistent, one CTA per SM, whole images), the epilogue of pose_sbp_fused.
 * workspace: pose_sbp_head_workspace_bytes(N, K, C), 256-byte aligned. */
unsigned long long pose_sbp_head_workspace_bytes(int N, int K, int C);
int pose_sbp_head_fused(const float* features, const float* weight, const void* kp, int kp_dtype, double sigma,
                        const float* lut, int lut_n, float* dlogits, float* logits_out, float* loss_out,
                        double* loss_num_out, float* joints, float conf_threshold, float coord_scale, int N, int C,
                        int K, int H, int W, float lambda_pos, float lambda_neg, double inv_norm, unsigned flags,
                        const double* bbox, float* packed_out, int input_h, int input_w, int tuning,
                        void* workspace, unsigned long long workspace_bytes, pose_stream_t stream);

/* ---- diagnostics of the reference sigmoids (csrc/common.cuh).
 * pose_sigmoid_ref_eval: y[i] = the device restatement of torch.sigmoid (sigmoid_ref) at x[i]; tests compare it with
 *   torch.sigmoid on CPU / CUDA tensors bit for bit.
 * pose_sigmoid_window_check: for every 61st fp32 m in (-80, inf], counts inputs below the candidate window of m whose
 *   reference sigmoid is not strictly below that of m (must be 0: the decoders only rank elements inside the window).
 *   violations_out [1] uint64 on device, zeroed by the call. */
int pose_sigmoid_ref_eval(const float* x, float* y, unsigned long long n, int sigmoid_ref, pose_stream_t stream);
int pose_sigmoid_window_check(unsigned long long* violations_out, int sigmoid_ref, pose_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POSE_B200_H */
