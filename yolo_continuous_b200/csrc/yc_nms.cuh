// yc_nms.cuh -- workspace layout and candidate emission shared by the stand-alone threshold kernel
// (yc_postproc.cu) and the fused head epilogue (yc_head_sm100.cu).
#pragma once
#include "yc_common.cuh"

namespace yc {

struct NmsWs {
    float4 *box;                    // [bs*rows] corners, indexed by original row
    float2 *oc;                     // [bs*rows] (obj, class_conf)
    unsigned long long *key_unsorted; // [bs*rows] candidates in arrival order, per image
    int *cls_unsorted;              // [bs*rows]
    unsigned long long *key_bucket; // [bs*rows] candidates grouped by class, then sorted in place
    int *kept_row;                  // [bs*rows] kept original rows, per segment
    float4 *kept_box;               // [bs*rows] spill of kept boxes beyond KEPT_SMEM
    int *counters;                  // start of the zero-initialised region
    int *cand_count;                // [bs]
    int *hist;                      // [bs*nc] candidates per (image, class)
    int *cursor;                    // [bs*nc]
    int *kept_count;                // [bs*nc]
    int *img_total;                 // [bs] kept detections of an image + 1 once known (0 = not yet): chained scan
    int *ticket;                    // [1] finish_kernel: CTAs take their work item in the order they start running
    int *seg_off;                   // [bs*nc]
    int *kept_off;                  // [bs*nc]
    size_t counters_bytes;
    size_t total_bytes;
};

static inline NmsWs carve(void *base, int bs, int rows, int nc)
{
    NmsWs w;
    char *p = (char *)base;
    const size_t n = (size_t)bs * rows, s = (size_t)bs * nc;
    auto take = [&](size_t bytes) { char *q = p; p += round_up_sz(bytes, 256); return q; };
    w.box = (float4 *)take(n * sizeof(float4));
    w.kept_box = (float4 *)take(n * sizeof(float4));
    w.key_unsorted = (unsigned long long *)take(n * 8);
    w.key_bucket = (unsigned long long *)take(n * 8);
    w.oc = (float2 *)take(n * sizeof(float2));
    w.cls_unsorted = (int *)take(n * 4);
    w.kept_row = (int *)take(n * 4);
    char *c0 = p;
    w.cand_count = (int *)take((size_t)bs * 4);
    w.hist = (int *)take(s * 4);
    w.cursor = (int *)take(s * 4);
    w.kept_count = (int *)take(s * 4);
    w.img_total = (int *)take((size_t)bs * 4);
    w.ticket = (int *)take(4);
    w.counters = (int *)c0;
    w.counters_bytes = (size_t)(p - c0);
    w.seg_off = (int *)take(s * 4);
    w.kept_off = (int *)take(s * 4);
    w.total_bytes = (size_t)(p - (char *)base);
    return w;
}

// score -> 64-bit key whose ascending order is (score descending, original row ascending),
// i.e. the order of a stable descending sort (torchvision nms; detect.py:133).
__device__ __forceinline__ unsigned long long make_key(float score, int row)
{
    unsigned int u = __float_as_uint(score);
    if (score == 0.0f) u = 0u; // -0 == +0 for the reference's sort
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)(~u) << 32) | (unsigned int)row;
}


// One candidate row that passed obj*cls >= conf (reference detect.py:111-121): record its corners,
// (obj, class_conf), sort key and class, and count it.  Called with the full warp converged; `pass`
// selects the lanes that emit.  Slots are claimed with one atomicAdd per warp (ballot + popc).
__device__ __forceinline__ void emit_candidates(bool pass, int b, int r, int rows, int nc, float x1, float y1, float x2,
                                                float y2, float obj, float conf_cls, float score, int cls,
                                                const NmsWs &ws)
{
    const unsigned ballot = __ballot_sync(0xffffffffu, pass);
    if (!ballot) return;
    const int lane = threadIdx.x & 31, leader = __ffs(ballot) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&ws.cand_count[b], __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pass) {
        const size_t ib = (size_t)b * rows;
        const int slot = base + __popc(ballot & ((1u << lane) - 1u));
        ws.box[ib + r] = make_float4(x1, y1, x2, y2);
        ws.oc[ib + r] = make_float2(obj, conf_cls);
        ws.key_unsorted[ib + slot] = make_key(score, r);
        ws.cls_unsorted[ib + slot] = cls;
        atomicAdd(&ws.hist[(size_t)b * nc + cls], 1);
    }
}

// single-lane form of emit_candidates (used where one lane owns a finished candidate)
__device__ __forceinline__ void emit_one(int b, int r, int rows, int nc, float x1, float y1, float x2, float y2, float obj,
                                         float conf_cls, float score, int cls, const NmsWs &ws)
{
    const size_t ib = (size_t)b * rows;
    const int slot = atomicAdd(&ws.cand_count[b], 1);
    ws.box[ib + r] = make_float4(x1, y1, x2, y2);
    ws.oc[ib + r] = make_float2(obj, conf_cls);
    ws.key_unsorted[ib + slot] = make_key(score, r);
    ws.cls_unsorted[ib + slot] = cls;
    atomicAdd(&ws.hist[(size_t)b * nc + cls], 1);
}

// candidate into a slot reserved beforehand (the fused head epilogue reserves one slot per objectness survivor with a
// single atomicAdd per warp, issued before the class scan so that its latency is hidden)
__device__ __forceinline__ void emit_at(int slot, int b, int r, int rows, int nc, float x1, float y1, float x2, float y2,
                                        float obj, float conf_cls, float score, int cls, const NmsWs &ws)
{
    const size_t ib = (size_t)b * rows;
    ws.box[ib + r] = make_float4(x1, y1, x2, y2);
    ws.oc[ib + r] = make_float2(obj, conf_cls);
    ws.key_unsorted[ib + slot] = make_key(score, r);
    ws.cls_unsorted[ib + slot] = cls;
    atomicAdd(&ws.hist[(size_t)b * nc + cls], 1);
}
// a reserved slot whose row failed the final threshold: skipped by bucket_kernel
__device__ __forceinline__ void emit_hole(int slot, int b, int rows, const NmsWs &ws)
{
    ws.cls_unsorted[(size_t)b * rows + slot] = -1;
}

// xywh (optionally divided by the input size) -> corners, in the reference's operation order
// (detect.py:98-103): x1 = cx - w/2 ...
__device__ __forceinline__ void xywh_to_corners(float cx, float cy, float bw, float bh, float div_w, float div_h, float &x1,
                                                float &y1, float &x2, float &y2)
{
    if (div_w > 0.f) { cx = __fdiv_rn(cx, div_w); bw = __fdiv_rn(bw, div_w); }
    if (div_h > 0.f) { cy = __fdiv_rn(cy, div_h); bh = __fdiv_rn(bh, div_h); }
    const float hw = __fmul_rn(bw, 0.5f), hh = __fmul_rn(bh, 0.5f);
    x1 = __fsub_rn(cx, hw); y1 = __fsub_rn(cy, hh);
    x2 = __fadd_rn(cx, hw); y2 = __fadd_rn(cy, hh);
}

// set by yc_detect_fused: the head epilogue emits NMS candidates instead of writing z
struct FusedDetect {
    float conf, div_w, div_h;
    int nc;
    NmsWs ws;
};

// kernels after compaction (segment offsets -> bucket scatter -> per-segment NMS -> scan -> gather)
int launch_nms_tail(const yc_nms_params *p, const NmsWs &ws, float *out_rows, int *out_idx, int *out_counts,
                    int *out_offsets, cudaStream_t stream);

} // namespace yc
