// yc_abi.cu -- C-ABI plumbing: error text, argument validation, dispatch between the tcgen05 head
// kernel and the any-shape path, and the small box utilities of utils/bbox.py.
#include <stdarg.h>
#include <string.h>

#include "yc_common.cuh"
#include "yc_nms.cuh"

namespace yc {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return YC_ERR_CUDA;
}

int launch_head_generic(const yc_head_desc *d, int rows_total, const int *row_off, unsigned level_mask,
                        cudaStream_t stream);
// Runs the levels that fit the tcgen05 kernel and reports the others in *left_mask; returns
// YC_ERR_UNSUPPORTED (reason in yc_last_error) when the head as a whole does not fit.
int launch_head_tcgen05(const yc_head_desc *d, int rows_total, const int *row_off, unsigned *left_mask,
                        const FusedDetect *fused, cudaStream_t stream);

// utils/bbox.py:62-72: area(a), area(b), clamped intersection, inter / (area_a + area_b - inter), binary32, no FMA
__device__ __forceinline__ float box_iou_f(const float4 a, const float4 b)
{
    const float a1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float a2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(a1, a2), inter));
}

__global__ void box_iou_kernel(const float4 *__restrict__ b1, int n, const float4 *__restrict__ b2, int m,
                               float *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * m) return;
    out[i] = box_iou_f(b1[i / m], b2[i % m]);
}

// Batched evaluator, matching step (the reference has no evaluator -- the IoU is its
// utils/bbox.py:62-72): one warp per (image, IoU threshold).  The detections of an image are visited in the order the NMS
// leaves them (class ascending, score descending within a class: the order greedy matching needs); each takes the
// not yet matched ground-truth box of its class with the highest IoU >= thr (ties: the first).  tp[t][d] = 1 if matched.
constexpr int EVAL_MAX_GT = 2048;   // ground-truth boxes per image (matched flags live in shared memory)
__global__ void __launch_bounds__(32) match_kernel(const float *__restrict__ det_rows, const int *__restrict__ det_off,
                                                   const float4 *__restrict__ gt_box, const int *__restrict__ gt_cls,
                                                   const int *__restrict__ gt_off, const float *__restrict__ thrs, int total,
                                                   unsigned char *__restrict__ tp)
{
    __shared__ unsigned int taken[EVAL_MAX_GT / 32];
    const int b = blockIdx.x, t = blockIdx.y, lane = threadIdx.x;
    const int d0 = det_off[b], d1 = det_off[b + 1], g0 = gt_off[b], ng = min(gt_off[b + 1] - g0, EVAL_MAX_GT);
    const float thr = thrs[t];
    for (int i = lane; i < EVAL_MAX_GT / 32; i += 32) taken[i] = 0u;
    __syncwarp();
    for (int d = d0; d < d1; ++d) {
        const float *r = det_rows + (size_t)d * 7;
        const float4 db = make_float4(r[0], r[1], r[2], r[3]);
        const int cls = (int)r[6];
        float best = -1.0f;
        int bi = -1;
        for (int g = lane; g < ng; g += 32) {
            if (gt_cls[g0 + g] != cls || (taken[g >> 5] >> (g & 31) & 1u)) continue;
            const float iou = box_iou_f(db, gt_box[g0 + g]);
            if (iou >= thr && iou > best) { best = iou; bi = g; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // highest IoU wins, ties go to the lower ground-truth index
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (ov > best || (ov == best && (bi < 0 || oi < bi)))) { best = ov; bi = oi; }
        }
        if (lane == 0) {
            tp[(size_t)t * total + d] = bi >= 0 ? 1 : 0;
            if (bi >= 0) taken[bi >> 5] |= 1u << (bi & 31);
        }
        __syncwarp();
    }
}

// utils/bbox.py:29-59 (tensor branch: the result starts as a clone of the input)
__global__ void cvt_bbox_kernel(const float4 *__restrict__ in, int n, int flag, float4 *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = in[i];
    float4 r = b;
    switch (flag) {
    case 0: case 2: r.y = b.z; r.z = b.y; break;
    case 1:
        r.z = __fsub_rn(b.y, b.x); r.w = __fsub_rn(b.w, b.z);
        r.x = __fadd_rn(b.x, __fmul_rn(r.z, 0.5f)); r.y = __fadd_rn(b.z, __fmul_rn(r.w, 0.5f)); break;
    case 3:
        r.z = __fsub_rn(b.z, b.x); r.w = __fsub_rn(b.w, b.y);
        r.x = __fadd_rn(b.x, __fmul_rn(r.z, 0.5f)); r.y = __fadd_rn(b.y, __fmul_rn(r.w, 0.5f)); break;
    case 4:
        r.x = __fsub_rn(b.x, __fmul_rn(b.z, 0.5f)); r.y = __fadd_rn(b.x, __fmul_rn(b.z, 0.5f));
        r.z = __fsub_rn(b.y, __fmul_rn(b.w, 0.5f)); r.w = __fadd_rn(b.y, __fmul_rn(b.w, 0.5f)); break;
    case 5:
        r.x = __fsub_rn(b.x, __fmul_rn(b.z, 0.5f)); r.y = __fsub_rn(b.y, __fmul_rn(b.w, 0.5f));
        r.z = __fadd_rn(b.x, __fmul_rn(b.z, 0.5f)); r.w = __fadd_rn(b.y, __fmul_rn(b.w, 0.5f)); break;
    }
    out[i] = r;
}

} // namespace yc

using namespace yc;

extern "C" const char *yc_last_error(void) { return g_err; }
extern "C" int yc_version(void) { return 100; }

extern "C" int yc_device_check(int dev)
{
    cudaDeviceProp prop;
    YC_CUDA(cudaGetDeviceProperties(&prop, dev));
    YC_REQUIRE(prop.major == 10, YC_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", dev,
               prop.major, prop.minor);
    return YC_OK;
}

static int validate_head(const yc_head_desc *d, bool need_z, int *row_off, int *rows_total_out)
{
    YC_REQUIRE(d, YC_ERR_INVALID, "yc_head_forward: null descriptor");
    YC_REQUIRE(d->nl >= 1 && d->nl <= YC_MAX_LEVELS && d->na >= 1 && d->na <= YC_MAX_ANCHORS && d->no >= 5 && d->bs >= 1,
               YC_ERR_INVALID, "yc_head_forward: bad nl=%d na=%d no=%d bs=%d", d->nl, d->na, d->no, d->bs);
    YC_REQUIRE(d->kind >= YC_HEAD_IDETECT && d->kind <= YC_HEAD_RAW, YC_ERR_INVALID, "yc_head_forward: bad kind %d",
               d->kind);
    YC_REQUIRE(d->x_dtype == YC_F32 || d->x_dtype == YC_BF16, YC_ERR_INVALID, "yc_head_forward: bad x_dtype %d",
               d->x_dtype);
    YC_REQUIRE(d->kind == YC_HEAD_RAW || d->z || !need_z, YC_ERR_INVALID, "yc_head_forward: z is null");
    if (d->kind == YC_HEAD_IBIN) {
        YC_REQUIRE(d->bins && d->bin_count >= 1 && d->no > 2 * (d->bin_count + 1) + 3, YC_ERR_INVALID,
                   "yc_head_forward: IBin needs bins and no > 2*(bin_count+1)+3");
    }
    int rows_total = 0;
    for (int i = 0; i < d->nl; ++i) {
        const yc_head_level &lv = d->level[i];
        YC_REQUIRE(lv.x && lv.blob && lv.K > 0 && lv.H > 0 && lv.W > 0, YC_ERR_INVALID,
                   "yc_head_forward: level %d has a null pointer or empty shape", i);
        YC_REQUIRE(d->kind != YC_HEAD_RAW || lv.raw, YC_ERR_INVALID, "yc_head_forward: YC_HEAD_RAW needs raw buffers");
        row_off[i] = rows_total;
        rows_total += d->na * lv.H * lv.W;
    }
    YC_REQUIRE((size_t)d->bs * rows_total < ((size_t)1 << 31), YC_ERR_UNSUPPORTED, "yc_head_forward: bs*rows >= 2^31");
    *rows_total_out = rows_total;
    return YC_OK;
}

extern "C" int yc_head_forward(const yc_head_desc *d, yc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int row_off[YC_MAX_LEVELS], rows_total = 0;
    const int vrc = validate_head(d, true, row_off, &rows_total);
    if (vrc != YC_OK) return vrc;
    const unsigned all = (1u << d->nl) - 1u;
    if (d->path == YC_PATH_GENERIC) return launch_head_generic(d, rows_total, row_off, all, stream);
    unsigned left = all;
    const int rc = launch_head_tcgen05(d, rows_total, row_off, &left, nullptr, stream);
    if (rc == YC_ERR_UNSUPPORTED && d->path == YC_PATH_AUTO) return launch_head_generic(d, rows_total, row_off, all, stream);
    if (rc != YC_OK) return rc;
    if (left) {
        YC_REQUIRE(d->path == YC_PATH_AUTO, YC_ERR_UNSUPPORTED, "yc_head_forward: levels 0x%x do not fit the tcgen05 kernel: %s",
                   left, g_err);
        return launch_head_generic(d, rows_total, row_off, left, stream);
    }
    return YC_OK;
}

extern "C" int yc_box_iou(const float *b1, int n, const float *b2, int m, float *out, yc_stream_t stream)
{
    if (n <= 0 || m <= 0) return YC_OK;
    YC_REQUIRE(b1 && b2 && out, YC_ERR_INVALID, "yc_box_iou: null argument");
    YC_REQUIRE((((uintptr_t)b1 | (uintptr_t)b2) & 15) == 0, YC_ERR_INVALID, "yc_box_iou: boxes must be 16-byte aligned");
    const size_t total = (size_t)n * m;
    box_iou_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4 *)b1, n,
                                                                                      (const float4 *)b2, m, out);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

extern "C" int yc_cvt_bbox(const float *in, int n, int flag, float *out, yc_stream_t stream)
{
    YC_REQUIRE(flag >= 0 && flag <= 5, YC_ERR_INVALID, "yc_cvt_bbox: bad flag %d", flag);
    if (n <= 0) return YC_OK;
    YC_REQUIRE(in && out, YC_ERR_INVALID, "yc_cvt_bbox: null argument");
    YC_REQUIRE((((uintptr_t)in | (uintptr_t)out) & 15) == 0, YC_ERR_INVALID, "yc_cvt_bbox: boxes must be 16-byte aligned");
    cvt_bbox_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4 *)in, n, flag, (float4 *)out);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

static int fused_setup(const yc_head_desc *d, const yc_nms_params *p, void *workspace, size_t workspace_bytes,
                       int *row_off, int *rows_total, FusedDetect *f)
{
    const int vrc = validate_head(d, false, row_off, rows_total);
    if (vrc != YC_OK) return vrc;
    YC_REQUIRE(p && workspace, YC_ERR_INVALID, "yc_detect_fused: null argument");
    YC_REQUIRE(d->kind == YC_HEAD_IDETECT || d->kind == YC_HEAD_IBIN, YC_ERR_UNSUPPORTED, "yc_detect_fused: IDetect / IBin decode only");
    const int nc_head = d->kind == YC_HEAD_IBIN ? d->no - 2 * (d->bin_count + 1) - 3 : d->no - 5;
    YC_REQUIRE(p->bs == d->bs && p->rows == *rows_total && p->nc == nc_head && p->nc > 0, YC_ERR_INVALID,
               "yc_detect_fused: nms params (bs=%d rows=%d nc=%d) do not match the head (bs=%d rows=%d no=%d)", p->bs,
               p->rows, p->nc, d->bs, *rows_total, d->no);
    YC_REQUIRE(!p->correct_boxes || p->image_hw, YC_ERR_INVALID, "yc_detect_fused: correct_boxes needs image_hw");
    YC_REQUIRE(p->bs <= 65535, YC_ERR_UNSUPPORTED, "yc_detect_fused: bs > 65535");
    void *base = (void *)round_up_sz((size_t)workspace, 256);
    f->ws = carve(base, p->bs, p->rows, p->nc);
    YC_REQUIRE(f->ws.total_bytes + ((char *)base - (char *)workspace) <= workspace_bytes, YC_ERR_WORKSPACE,
               "yc_detect_fused: workspace %zu < %zu", workspace_bytes, f->ws.total_bytes + 256);
    f->conf = p->conf_thres; f->div_w = p->box_div_w; f->div_h = p->box_div_h; f->nc = p->nc;
    return YC_OK;
}

namespace yc { int set_reserved_sms(int n); }

extern "C" int yc_copy_async(void *dst, const void *src, size_t bytes, yc_stream_t stream)
{
    YC_REQUIRE(dst && src, YC_ERR_INVALID, "yc_copy_async: null argument");
    YC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return YC_OK;
}

extern "C" int yc_reserve_sms(int n)
{
    return yc::set_reserved_sms(n);
}

extern "C" int yc_nms_workspace_reset(const yc_nms_params *p, void *workspace, size_t workspace_bytes, yc_stream_t stream_)
{
    YC_REQUIRE(p && workspace, YC_ERR_INVALID, "yc_nms_workspace_reset: null argument");
    YC_REQUIRE(p->bs > 0 && p->rows > 0 && p->nc > 0, YC_ERR_INVALID, "yc_nms_workspace_reset: bad shape");
    void *base = (void *)round_up_sz((size_t)workspace, 256);
    NmsWs ws = carve(base, p->bs, p->rows, p->nc);
    YC_REQUIRE(ws.total_bytes + ((char *)base - (char *)workspace) <= workspace_bytes, YC_ERR_WORKSPACE,
               "yc_nms_workspace_reset: workspace %zu < %zu", workspace_bytes, ws.total_bytes + 256);
    YC_CUDA(cudaMemsetAsync(ws.counters, 0, ws.counters_bytes, (cudaStream_t)stream_));
    return YC_OK;
}

static int fused_head(const yc_head_desc *d, const yc_nms_params *p, void *workspace, size_t workspace_bytes, bool reset,
                      yc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int row_off[YC_MAX_LEVELS], rows_total = 0;
    FusedDetect f;
    const int src = fused_setup(d, p, workspace, workspace_bytes, row_off, &rows_total, &f);
    if (src != YC_OK) return src;
    // check the shape before touching the workspace, so that an unsupported call has no side effects
    if (reset) YC_CUDA(cudaMemsetAsync(f.ws.counters, 0, f.ws.counters_bytes, stream));
    unsigned left = 0;
    const int rc = launch_head_tcgen05(d, rows_total, row_off, &left, &f, stream);
    if (rc != YC_OK) return rc;
    YC_REQUIRE(left == 0, YC_ERR_UNSUPPORTED, "yc_detect_fused: levels 0x%x do not fit the tcgen05 kernel: %s", left, g_err);
    return YC_OK;
}

extern "C" int yc_detect_fused_head(const yc_head_desc *d, const yc_nms_params *p, void *workspace,
                                    size_t workspace_bytes, yc_stream_t stream)
{
    return fused_head(d, p, workspace, workspace_bytes, true, stream);
}

extern "C" int yc_detect_fused_head_noreset(const yc_head_desc *d, const yc_nms_params *p, void *workspace,
                                            size_t workspace_bytes, yc_stream_t stream)
{
    return fused_head(d, p, workspace, workspace_bytes, false, stream);
}

extern "C" int yc_nms_from_candidates(const yc_nms_params *p, void *workspace, size_t workspace_bytes, float *out_rows,
                                      int32_t *out_idx, int32_t *out_counts, int32_t *out_offsets, yc_stream_t stream_)
{
    YC_REQUIRE(p && workspace && out_rows && out_idx && out_counts && out_offsets, YC_ERR_INVALID,
               "yc_nms_from_candidates: null argument");
    YC_REQUIRE(p->bs > 0 && p->rows > 0 && p->nc > 0 && p->bs <= 65535, YC_ERR_INVALID, "yc_nms_from_candidates: bad shape");
    YC_REQUIRE(!p->correct_boxes || p->image_hw, YC_ERR_INVALID, "yc_nms_from_candidates: correct_boxes needs image_hw");
    void *base = (void *)round_up_sz((size_t)workspace, 256);
    NmsWs ws = carve(base, p->bs, p->rows, p->nc);
    YC_REQUIRE(ws.total_bytes + ((char *)base - (char *)workspace) <= workspace_bytes, YC_ERR_WORKSPACE,
               "yc_nms_from_candidates: workspace %zu < %zu", workspace_bytes, ws.total_bytes + 256);
    return launch_nms_tail(p, ws, out_rows, out_idx, out_counts, out_offsets, (cudaStream_t)stream_);
}

extern "C" int yc_detect_fused(const yc_head_desc *d, const yc_nms_params *p, void *workspace, size_t workspace_bytes,
                               float *out_rows, int32_t *out_idx, int32_t *out_counts, int32_t *out_offsets,
                               yc_stream_t stream)
{
    YC_REQUIRE(out_rows && out_idx && out_counts && out_offsets, YC_ERR_INVALID, "yc_detect_fused: null argument");
    const int rc = yc_detect_fused_head(d, p, workspace, workspace_bytes, stream);
    if (rc != YC_OK) return rc;
    return yc_nms_from_candidates(p, workspace, workspace_bytes, out_rows, out_idx, out_counts, out_offsets, stream);
}

extern "C" int yc_match_detections(const float *det_rows, const int32_t *det_offsets, int bs, int total, const float *gt_boxes,
                                   const int32_t *gt_labels, const int32_t *gt_offsets, const float *iou_thrs, int n_thr,
                                   uint8_t *tp, yc_stream_t stream)
{
    if (bs <= 0 || n_thr <= 0 || total <= 0) return YC_OK;
    YC_REQUIRE(det_rows && det_offsets && gt_boxes && gt_labels && gt_offsets && iou_thrs && tp, YC_ERR_INVALID,
               "yc_match_detections: null argument");
    YC_REQUIRE(((uintptr_t)gt_boxes & 15) == 0, YC_ERR_INVALID, "yc_match_detections: gt_boxes must be 16-byte aligned");
    YC_REQUIRE(bs <= 65535 && n_thr <= 65535, YC_ERR_UNSUPPORTED, "yc_match_detections: grid too large");
    match_kernel<<<dim3(bs, n_thr), 32, 0, (cudaStream_t)stream>>>(det_rows, det_offsets, (const float4 *)gt_boxes, gt_labels,
                                                                  gt_offsets, iou_thrs, total, tp);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}
