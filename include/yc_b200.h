/*
 * yc_b200.h -- C ABI of the B200-native detection post-backbone path.
 *
 * Drop-in boundary for xin-pu/yolo-continuous (reference paths relative to its
 * checkout).  The reference has no FFI of its own (100 % Python); these entry points
 * are what a binding for the path would call.  Every function takes plain pointers and
 * sizes, launches on the caller's stream, never synchronises the host unless stated,
 * and returns 0 on success or a negative yc_status; yc_last_error() gives the text.
 * All pointers are DEVICE pointers unless the name ends in _host.
 */
#ifndef YC_B200_H
#define YC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define YC_API __attribute__((visibility("default")))
#else
#define YC_API
#endif

#define YC_MAX_LEVELS 4
#define YC_MAX_ANCHORS 4

typedef void *yc_stream_t; /* cudaStream_t */

enum yc_status {
    YC_OK = 0,
    YC_ERR_INVALID = -1,     /* bad argument */
    YC_ERR_UNSUPPORTED = -2, /* shape / dtype not supported by this build */
    YC_ERR_CUDA = -3,        /* CUDA runtime / driver error */
    YC_ERR_WORKSPACE = -4    /* workspace too small */
};

enum yc_dtype { YC_F32 = 0, YC_BF16 = 1 };

/* decode performed by the head epilogue */
enum yc_head_kind {
    YC_HEAD_IDETECT = 0, /* nets/idetect.py:40-43 (also IAuxDetect lead branch, nets/iaux_detect.py:44-47) */
    YC_HEAD_IBIN = 1,    /* nets/ibin.py:56-72 + losses/sigmoid_bin.py:49-63 */
    YC_HEAD_RAW = 2      /* conv (+implicit) + permute only: training output, IAuxDetect m2 branch
                            (nets/iaux_detect.py:37-38) */
};

/* which kernel family runs the 1x1 conv */
enum yc_head_path {
    YC_PATH_AUTO = 0,    /* tcgen05 when the shape allows it, else generic */
    YC_PATH_TCGEN05 = 1, /* sm_100a tcgen05/TMEM/TMA kernels (bf16 maps: one bf16 MMA per k-step; float32 maps: three fp16
                            MMAs per k-step on a hi/lo split); YC_ERR_UNSUPPORTED if the shape does not fit */
    YC_PATH_GENERIC = 2  /* any-shape FFMA kernel (exact binary32 accumulation) */
};

YC_API const char *yc_last_error(void);
YC_API int yc_version(void);
/* 0 when device `dev` is an sm_100 part this library can run on. */
YC_API int yc_device_check(int dev);

/* ------------------------------------------------------------------------------------------
 * Head parameters.  Replaces the parameter handling of IDetect/IAuxDetect/IBin
 * (nets/idetect.py:21-24, nets/common.py:416-439): ImplicitA is folded into the bias
 * (b' = b + W.ia, binary64 accumulation), ImplicitM stays a per-channel epilogue scale.
 * yc_head_pack writes one blob per level:
 *   [bias2 f32 Npad][scale f32 Npad][scale_split f32 Npad][(scale,bias2) f32x2 Npad][(scale_split,bias2) f32x2 Npad]
 *   [w32 f32 N*K][w_hi_t f16 K*(na*npad_g)][w_lo_t likewise][w_bf16 Npad*K]   (Npad = N rounded up to 16, npad_g = N/na
 *   rounded up to 16: in the transposed copies the channels of anchor a start at column a*npad_g)
 * w_hi_t + w_lo_t = W[c,:] * 2^shift(c), split into two fp16 numbers and stored transposed: the operands of the
 * float32-grade tensor-core path (three fp16 MMAs per k-step, see csrc/yc_head_sm100_split.cu); scale_split undoes the
 * row scaling and the activation scaling.
 * W [N,K] f32 row-major (Conv2d weight [N,K,1,1]); bias [N] or NULL; ia [K] or NULL; im [N] or NULL.
 */
YC_API size_t yc_head_pack_bytes(int N, int K);
YC_API int yc_head_pack(const float *W, const float *bias, const float *ia, const float *im, int N, int K, int na,
                 void *blob, yc_stream_t stream);   /* na = anchors per level (row c of W = anchor c / (N/na)) */

typedef struct yc_head_level {
    const void *x;     /* [bs, K, H, W] NCHW feature map, dtype = yc_head_desc.x_dtype */
    const void *blob;  /* from yc_head_pack */
    float *raw;        /* optional [bs, na, H, W, no] pre-sigmoid output (forward()'s list `x`), or NULL */
    int32_t K, H, W;
    float stride;                         /* IDetect.stride[i] */
    float anchor_wh[YC_MAX_ANCHORS * 2];  /* anchor_grid[i] in pixels, nets/idetect.py:18-20 */
    float stride_y;                       /* multiplier of the y coordinate; 0 = same as `stride`.  The plain Detect
                                             head + decode_box (detect.py:29-87, normalised boxes) is this decode with
                                             stride = 1/W, stride_y = 1/H and anchor_wh = anchor / image_size */
} yc_head_level;

typedef struct yc_head_desc {
    int32_t kind;      /* yc_head_kind */
    int32_t path;      /* yc_head_path */
    int32_t x_dtype;   /* yc_dtype of the feature maps: YC_F32 -> float32-grade result (1e-5 parity: fp16 hi/lo split of
                          both operands on the tensor cores, |x| < 2^20; or exact FFMA on the generic path);
                          YC_BF16 -> bf16 MMA (1e-3 parity) */
    int32_t nl, na, no; /* levels, anchors per level, conv outputs per anchor (85; IBin 127) */
    int32_t bin_count;  /* IBin only (21) */
    int32_t bs;
    yc_head_level level[YC_MAX_LEVELS];
    float *z;           /* [bs, sum(na*H*W), no_out] decoded rows (no_out = no, IBin: no - 2*(bin_count+1) + 2),
                           NULL for YC_HEAD_RAW */
    const float *bins;  /* IBin: device [bin_count] (SigmoidBin.bins buffer) */
    int32_t x_channels_last; /* 1: the feature maps are [bs, H, W, K] (torch channels_last, e.g. the output of a re-parameterised
                                RepConv, nets/common.py:561-614, run channels-last): the head GEMM reads them as a K-major
                                operand.  bf16 maps on the tcgen05 path only (YC_ERR_UNSUPPORTED otherwise). */
} yc_head_desc;

/* IDetect/IAuxDetect/IBin.forward (eval and train) for all levels in one call.
 * Replaces nets/idetect.py:26-45, nets/iaux_detect.py:27-49, nets/ibin.py:35-74. */
YC_API int yc_head_forward(const yc_head_desc *desc, yc_stream_t stream);

/* Variant A decode: detect.decode_box, detect.py:29-87, one level.
 * conv [bs, na*no, ny, nx] f32 -> out [bs, na*ny*nx, no] normalised xywh + sigmoid scores. */
YC_API int yc_decode_box(const float *conv, int bs, int na, int no, int ny, int nx, const float *anchor_wh_scaled_host,
                  float *out, yc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Threshold + compaction + batched per-class NMS (+ letterbox undo).
 * Replaces detect.non_max_suppression, detect.py:90-144 (utils/bbox.py:121-198) and
 * torchvision.ops.nms at detect.py:133, for a whole batch without host round trips.
 */
typedef struct yc_nms_params {
    int32_t bs, rows, row_stride; /* pred [bs, rows, row_stride] f32, row = cx,cy,w,h,obj,cls[nc] */
    int32_t nc;
    float conf_thres;             /* compared in binary32, as torch does (detect.py:111) */
    double nms_thres;             /* compared in binary64 against the binary32 IoU, as torchvision does */
    int32_t write_corners;        /* 1: overwrite pred[..., :4] with x1,y1,x2,y2 (detect.py:98-103) */
    int32_t correct_boxes;        /* 1: apply yolo_correct_boxes (detect.py:147-165): rows become y1,x1,y2,x2 px */
    int32_t letterbox;
    int32_t input_h, input_w;
    const int32_t *image_hw;      /* device [bs,2] original image h,w (or [1,2] with image_hw_stride 0) */
    int32_t image_hw_stride;      /* 2 or 0 */
    float box_div_w, box_div_h;   /* when > 0: cx,w /= box_div_w and cy,h /= box_div_h (IEEE division) before the
                                     corners are formed -- turns IDetect's input-pixel boxes into the normalised
                                     boxes yolo_correct_boxes expects; 0 = off */
} yc_nms_params;

YC_API size_t yc_nms_workspace_bytes(int bs, int rows, int nc);
/* out_rows [bs*rows, 7] capacity (dense: image b's detections start at out_offsets[b]),
 * out_idx [bs*rows] original row index of each kept detection, out_counts [bs], out_offsets [bs+1]. */
YC_API int yc_nms_batched(float *pred, const yc_nms_params *p, void *workspace, size_t workspace_bytes, float *out_rows,
                   int32_t *out_idx, int32_t *out_counts, int32_t *out_offsets, yc_stream_t stream);

/* Fully fused step: head conv -> (epilogue) sigmoid/decode of box+objectness, class max, obj*cls >= conf,
 * candidate emission -> per-class NMS -> letterbox undo.  z is never materialised: HBM sees the feature
 * maps once and the detections.  Same results as yc_head_forward followed by yc_nms_batched (same arithmetic,
 * same tie rules).  Replaces the sequence reference detect.py:227-234 for I*Detect heads.  Needs the tcgen05
 * path for every level (bf16 maps); YC_ERR_UNSUPPORTED otherwise.  p->rows / nc must match the head. */
YC_API int yc_detect_fused(const yc_head_desc *desc, const yc_nms_params *p, void *workspace, size_t workspace_bytes,
                    float *out_rows, int32_t *out_idx, int32_t *out_counts, int32_t *out_offsets, yc_stream_t stream);

/* The two halves of yc_detect_fused, for callers that want to time or overlap them separately:
 * yc_detect_fused_head clears the counters and runs the head kernel, which leaves the candidates of every image
 * in `workspace`; yc_nms_from_candidates runs the per-class NMS + gather on them. */
YC_API int yc_detect_fused_head(const yc_head_desc *desc, const yc_nms_params *p, void *workspace, size_t workspace_bytes,
                         yc_stream_t stream);
YC_API int yc_nms_from_candidates(const yc_nms_params *p, void *workspace, size_t workspace_bytes, float *out_rows,
                           int32_t *out_idx, int32_t *out_counts, int32_t *out_offsets, yc_stream_t stream);
/* Pipelined form (NMS kernels of batch i on a second stream next to the head kernel of batch i+1, one workspace per
 * batch in flight): yc_nms_workspace_reset clears the counters -- enqueue it behind yc_nms_from_candidates on the
 * NMS stream -- and yc_detect_fused_head_noreset is yc_detect_fused_head for a workspace that is already clear, so
 * that nothing but the head kernel runs on the head stream. */
/* The persistent head kernels normally take every SM (one CTA each).  yc_reserve_sms(n) makes them leave n SMs free,
 * e.g. one for the single-CTA NCCL all-gather of the previous batch's detections, which cannot share an SM with a
 * head CTA and would otherwise delay the CTA it displaces by its whole duration.  Returns the previous value. */
YC_API int yc_reserve_sms(int n);
/* Device-to-device copy on `stream` (the per-step staging of a rank's detection message: one driver call instead of a
 * framework tensor copy on the host's critical path). */
YC_API int yc_copy_async(void *dst, const void *src, size_t bytes, yc_stream_t stream);
YC_API int yc_nms_workspace_reset(const yc_nms_params *p, void *workspace, size_t workspace_bytes, yc_stream_t stream);
YC_API int yc_detect_fused_head_noreset(const yc_head_desc *desc, const yc_nms_params *p, void *workspace,
                                 size_t workspace_bytes, yc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Detection exchange between the GPUs of one box (images shard across ranks, the only exchange is
 * the gather of the variable-length detection lists; the reference has no counterpart).  No collective library and no
 * host work per step: yc_xchg_push launches a kernel that stores this rank's message (the [counts | offsets] header and
 * the first rows of the detection list, exactly the buffer prefix yc_detect_fused leaves) into a slot of EVERY rank's
 * receive buffer over NVLink peer mappings and raises a flag there; yc_xchg_wait launches a kernel that waits until the
 * messages of all ranks with this rank's next sequence number have arrived.  Sequence numbers live on the device: both
 * launches can be captured in a CUDA graph and replayed.  Credits: a slot is reused only after the receiver has waited
 * for a later message (calling wait for sequence number j declares everything before j consumed).
 * Set-up: every rank calls yc_xchg_alloc (cudaMalloc + IPC handle), the 64-byte handles are exchanged by the host
 * (any transport), every rank opens the others' with yc_xchg_open and uploads the `world` base pointers (its own buffer at
 * index `rank`) as a device array `peers_dev`.
 * Receive buffer: message of rank r with sequence number q at  buf + ((q % slots) * world + r) * msg_bytes.
 */
YC_API size_t yc_xchg_bytes(int world, int slots, size_t msg_bytes);
YC_API int yc_xchg_alloc(int world, int slots, size_t msg_bytes, void **buf, uint8_t *handle64);
YC_API int yc_xchg_open(const uint8_t *handle64, void **peer_buf);
YC_API int yc_xchg_close(void *peer_buf);
YC_API int yc_xchg_free(void *buf);
/* msg: device [hdr_ints int32 (counts bs | offsets bs+1 | pad to a multiple of 4)][rows x 7 f32]; at most max_rows rows move. */
YC_API int yc_xchg_push(const void *msg, int hdr_ints, int bs, int max_rows, void *const *peers_dev, int world, int rank,
                 int slots, size_t msg_bytes, yc_stream_t stream);
/* lag: the kernel returns at once unless this rank has itself pushed message (next wait sequence number + lag) already;
 * "push(i); wait(lag = 1)" per step waits for step i - 1 and never spins in the steady state; lag = 0 waits for the next
 * message; lag = -1 waits for every message this rank has pushed so far (end of a stream of batches). */
YC_API int yc_xchg_wait(void *const *peers_dev, int world, int rank, int slots, size_t msg_bytes, int lag, yc_stream_t stream);
/* out4 (host): next push sequence number, next wait sequence number, internal, error (1 = a wait timed out). Synchronises. */
YC_API int yc_xchg_state(const void *buf, int world, int slots, size_t msg_bytes, uint32_t *out4, yc_stream_t stream);


/* torchvision.ops.nms drop-in for one box set (detect.py:133): boxes [n,4] xyxy, scores [n].
 * keep [n] receives kept indices in score order, *keep_count_dev their number. workspace from
 * yc_nms_workspace_bytes(1, n, 1). */
YC_API int yc_nms_single(const float *boxes, const float *scores, int n, double thr, void *workspace, size_t workspace_bytes,
                  int32_t *keep, int32_t *keep_count_dev, yc_stream_t stream);

/* utils/bbox.py:62-72 box_iou and :29-59 cvt_bbox on device. */
/* ------------------------------------------------------------------------------------------
 * The steps either side of the model.
 * yc_letterbox_batch replaces prepare_test_image (detect.py:16-26) + LetterBox.__call__ without scale_fill
 * (image_enhance/letter_box.py:27-60) for a batch: cv2.resize INTER_LINEAR (8-bit fixed point, bit exact),
 * constant border `pad_value` (114), /255, HWC -> CHW, channel order kept (BGR as cv2.imread gives it).
 * `imgs` is a DEVICE array of bs descriptors; src images are uint8 HWC with 3 channels on the device.
 * rs_w/rs_h = int(round(w*r)), int(round(h*r)); top/left = int(round(dh-0.1)), int(round(dw-0.1)) as the reference
 * computes them on the host.  out: [bs, 3, out_h, out_w] of `out_dtype` (YC_F32 / YC_BF16).
 */
typedef struct yc_letterbox_image {
    const uint8_t *src;
    int32_t src_h, src_w, src_pitch; /* pitch in bytes */
    int32_t rs_h, rs_w;              /* size after the resize */
    int32_t top, left;               /* where the resized image starts in the output */
    int32_t pad_value;               /* 114 */
} yc_letterbox_image;
YC_API int yc_letterbox_batch(const yc_letterbox_image *imgs, int bs, int out_h, int out_w, int out_dtype, void *out,
                       yc_stream_t stream);

/* Formatting loop of predict (detect.py:236-258) for every detection of a batch: rows [total,7]
 * (y1,x1,y2,x2,obj,class_conf,class) and offsets [bs+1] as yc_nms_batched / yc_detect_fused leave them ->
 * box_xyxy int32 [total,4] = (max(0,floor x1), max(0,floor y1), min(W,floor x2), min(H,floor y2)),
 * conf = obj*class_conf, label = int(class).  image_hw as in yc_nms_params. */
YC_API int yc_format_detections(const float *rows, const int32_t *offsets, int bs, const int32_t *image_hw,
                         int image_hw_stride, int32_t *box_xyxy, float *conf, int32_t *label, yc_stream_t stream);

YC_API int yc_box_iou(const float *b1, int n, const float *b2, int m, float *out, yc_stream_t stream);
/* Matching step of a batched on-device evaluator (the reference has none -- the IoU is its
 * box_iou, utils/bbox.py:62-72).  det_rows [total,7] + det_offsets [bs+1] as yc_nms_batched / yc_detect_fused leave them
 * (per image: class ascending, score descending); gt_boxes [n_gt,4] in the detections' coordinate convention, gt_labels
 * [n_gt] int32, gt_offsets [bs+1] (at most 2048 boxes per image are considered).  For every IoU threshold t a detection
 * takes the unmatched ground-truth box of its class with the highest IoU >= thr: tp [n_thr, total] uint8. */
YC_API int yc_match_detections(const float *det_rows, const int32_t *det_offsets, int bs, int total, const float *gt_boxes,
                        const int32_t *gt_labels, const int32_t *gt_offsets, const float *iou_thrs, int n_thr, uint8_t *tp,
                        yc_stream_t stream);
YC_API int yc_cvt_bbox(const float *in, int n, int flag, float *out, yc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* YC_B200_H */
