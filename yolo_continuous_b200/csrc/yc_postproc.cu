// yc_postproc.cu -- confidence threshold + stream compaction, class bucketing, per-(image,class)
// sort + blocked greedy NMS, gather (+ letterbox undo).  Replaces detect.non_max_suppression
// (reference detect.py:90-144) and torchvision.ops.nms (detect.py:133) for a whole batch with
// no host round trip.  All comparisons that decide membership are evaluated exactly as the
// reference evaluates them (binary32, round-to-nearest, no FMA contraction).
#include "yc_common.cuh"
#include "yc_nms.cuh"

namespace yc {

constexpr int TC_ROWS = 128;    // rows per CTA in the threshold/compaction kernel
constexpr int NMS_NT = 128;     // threads per NMS CTA == sorted boxes per chunk
constexpr int SORT_SMEM = 1024; // segments up to this size are sorted in shared memory
constexpr int KEPT_SMEM = 256;  // kept boxes cached in shared memory per segment

// ---- mbarrier / bulk-copy helpers (TMA 1D) -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- K1: corners, class max, threshold, compaction (detect.py:98-116) ------------------------
// One CTA owns TC_ROWS consecutive rows of one image.  The row tile ([128][row_stride] floats,
// one contiguous span of pred) is staged in shared memory with a single bulk copy when it is
// 16-byte aligned, so HBM sees full-line requests; each thread then owns one row
// (stride-row_stride reads are bank-conflict free for odd row_stride such as 85).
__global__ void __launch_bounds__(TC_ROWS) threshold_compact_kernel(float *__restrict__ pred, int rows, int row_stride,
                                                                    int nc, float conf, int write_corners, float div_w,
                                                                    float div_h, NmsWs ws)
{
    extern __shared__ __align__(128) float tile[];
    __shared__ __align__(8) uint64_t bar;
    const int b = blockIdx.y, r0 = blockIdx.x * TC_ROWS, tid = threadIdx.x;
    const int m = min(TC_ROWS, rows - r0);
    float *src = pred + ((size_t)b * rows + r0) * row_stride;
    const uint32_t bytes = (uint32_t)m * row_stride * 4u;
    const bool bulk = (((uintptr_t)src | bytes) & 15u) == 0;
    if (bulk) {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar, bytes);
            bulk_g2s(tile, src, bytes, &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        for (int i = tid; i < m * row_stride; i += TC_ROWS) tile[i] = src[i];
        __syncthreads();
    }
    const int r = r0 + tid;
    bool pass = false;
    int best = 0;
    float x1 = 0, y1 = 0, x2 = 0, y2 = 0, obj = 0, bv = 0, score = 0;
    if (tid < m) {
        const float *q = tile + tid * row_stride;
        xywh_to_corners(q[0], q[1], q[2], q[3], div_w, div_h, x1, y1, x2, y2);
        obj = q[4];
        bv = q[5];
#pragma unroll 8
        for (int c = 1; c < nc; ++c) {
            const float v = q[5 + c];
            if (v > bv) { bv = v; best = c; } // first maximum, as torch.max
        }
        score = __fmul_rn(obj, bv);
        pass = score >= conf;
        if (write_corners) {
            float *o = src + (size_t)tid * row_stride;
            o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
        }
    }
    emit_candidates(pass, b, r, rows, nc, x1, y1, x2, y2, obj, bv, score, best, ws);
}

// ---- K2: class-histogram scan + bucket scatter --------------------------------------------------------
// grid (ceil(rows / BUCKET_SLOTS), bs).  Every CTA of an image rescans the image's class histogram (nc
// values, trivial) into shared memory, CTA 0 publishes it as seg_off, and each CTA moves its slice of the
// image's candidate keys into their class buckets.
constexpr int BUCKET_SLOTS = 4096;
constexpr int BUCKET_SMEM_NC = 1024;

__global__ void __launch_bounds__(256) bucket_kernel(int rows, int nc, NmsWs ws)
{
    __shared__ int s_off[BUCKET_SMEM_NC];
    __shared__ int warp_tot[8];
    __shared__ int carry_s;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = ws.cand_count[b];
    const int s0 = blockIdx.x * BUCKET_SLOTS;
    if (blockIdx.x != 0 && s0 >= n) return;
    const size_t sb = (size_t)b * nc;
    const bool first = blockIdx.x == 0;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < nc; c0 += 256) {
        const int c = c0 + tid;
        const int v = c < nc ? ws.hist[sb + c] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        int pre = carry_s;
        for (int w = 0; w < wid; ++w) pre += warp_tot[w];
        if (c < nc) {
            if (c < BUCKET_SMEM_NC) s_off[c] = pre + inc - v;
            if (first || nc > BUCKET_SMEM_NC) ws.seg_off[sb + c] = pre + inc - v; // same value from every CTA
        }
        __syncthreads();
        if (tid == 255) carry_s = pre + inc;
        __syncthreads();
    }
    const size_t ib = (size_t)b * rows;
    const int s1 = min(n, s0 + BUCKET_SLOTS);
    for (int slot = s0 + tid; slot < s1; slot += 256) {
        const int cls = ws.cls_unsorted[ib + slot];
        if (cls < 0) continue; // slot reserved by the fused head epilogue for a row that failed obj * cls >= conf
        const int off = cls < BUCKET_SMEM_NC ? s_off[cls] : ws.seg_off[sb + cls];
        const int pos = off + atomicAdd(&ws.cursor[sb + cls], 1);
        ws.key_bucket[ib + pos] = ws.key_unsorted[ib + slot];
    }
}

// IoU test of torchvision's CPU nms kernel (third-party; call site detect.py:133):
// inter / (area_a + area_b - inter) > thr, binary32 round-to-nearest throughout, the quotient
// compared against thr_f = largest binary32 <= the binary64 threshold (equivalent to the
// reference's promotion of the quotient to binary64).  NaN compares false.
__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, const float thr_f)
{
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    // disjoint boxes (the vast majority of pairs): the quotient is +-0 or NaN, never above a non-negative threshold -- skip
    // the extents, the two areas and the IEEE division.  w = max(0, xx2 - xx1) is 0 exactly when !(xx2 > xx1): the
    // difference of two distinct floats is never 0 (gradual underflow), and NaN compares false / max(0, NaN) = 0
    if (!(xx2 > xx1 && yy2 > yy1) && thr_f >= 0.0f) return false;
    const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    const float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float uni = __fsub_rn(__fadd_rn(aa, ab), inter);
    return __fdiv_rn(inter, uni) > thr_f;
}

// ---- K4: one CTA per (image, class) segment: sort, then blocked greedy suppression -----------------
// Sort: bitonic network with every comparator ascending, so partners beyond n can simply be
// skipped (equivalent to +inf padding) -- works for any n, in shared memory up to SORT_SMEM keys
// and in place in global memory above that.
// Greedy: sorted boxes are consumed in chunks of NMS_NT.  A chunk is first tested against every box
// kept so far (one thread per box), then a NMS_NT x NMS_NT upper-triangular IoU bitmask is built in
// shared memory and warp 0 walks it: the lane owning the current 64-bit word finds the next
// surviving box with ffs, all lanes OR that box's mask row into their word.  This is exactly the
// sequential rule "keep i; suppress every later j with IoU(i,j) > thr" of the reference.
struct NmsSmem {
    unsigned long long skeys[SORT_SMEM];
    float4 sbox[NMS_NT];
    float4 skept[KEPT_SMEM];
    unsigned long long smask[NMS_NT][2];    // IoU bits of row i against later boxes of the chunk (OR of two threads' parts)
    unsigned char sdeadb[NMS_NT];           // box i of the chunk is suppressed by a box kept earlier
    unsigned int sdead[NMS_NT / 32];
    unsigned int sany[NMS_NT / 32];
    unsigned long long skeptw[2];
    int s_nkept;
};

// whole-CTA path for one segment of n >= 2 boxes (all NMS_NT threads call it)
__device__ void nms_segment_cta(int b, int c, int n, int rows, int nc, float thr_f, const NmsWs &ws, NmsSmem &sm)
{
    unsigned long long *skeys = sm.skeys;
    float4 *sbox = sm.sbox, *skept = sm.skept;
    unsigned long long(*smask)[2] = sm.smask;
    unsigned char *sdeadb = sm.sdeadb;
    unsigned int *sdead = sm.sdead, *sany = sm.sany;
    unsigned long long *skeptw = sm.skeptw;
    int &s_nkept = sm.s_nkept;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t seg = (size_t)b * nc + c;
    const size_t ib = (size_t)b * rows;
    const int off = ws.seg_off[seg];
    unsigned long long *gkeys = ws.key_bucket + ib + off;
    int *kept_row = ws.kept_row + ib + off;
    float4 *kept_box = ws.kept_box + ib + off;

    unsigned long long *K = gkeys;
    if (n <= SORT_SMEM) {
        for (int i = tid; i < n; i += NMS_NT) skeys[i] = gkeys[i];
        K = skeys;
    }
    __syncthreads();
    // (a thread takes comparator t of a step directly -- lower index i = t with a 0 bit inserted at the stride's position --
    // instead of walking all elements and skipping the upper halves of the pairs: half the instructions)
    int half_pairs = 1;
    while (2 * half_pairs < n) half_pairs <<= 1;   // comparators per step of the padded network
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        const int h = k >> 1;
        for (int t = tid; t < half_pairs; t += NMS_NT) {
            const int i = ((t & ~(h - 1)) << 1) | (t & (h - 1)), l = i ^ (k - 1);
            if (l < n) {
                const unsigned long long a = K[i], d = K[l];
                if (a > d) { K[i] = d; K[l] = a; }
            }
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = tid; t < half_pairs; t += NMS_NT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                if (l < n) {
                    const unsigned long long a = K[i], d = K[l];
                    if (a > d) { K[i] = d; K[l] = a; }
                }
            }
            __syncthreads();
        }
    }

    if (tid == 0) s_nkept = 0;
    __syncthreads();

    for (int base = 0; base < n; base += NMS_NT) {
        const int m = min(NMS_NT, n - base);
        float4 bx = make_float4(0, 0, 0, 0);
        if (tid < m) {
            const int row = (int)(K[base + tid] & 0xffffffffull);
            bx = ws.box[ib + row];
            sbox[tid] = bx;
        }
        const int nk = s_nkept;
        sdeadb[tid] = tid >= m;
        smask[tid][0] = 0;
        smask[tid][1] = 0;
        __syncthreads();
        {   // chunk against the kept list: NMS_NT / 2^ceil(log2 m) threads share a box and stride over the kept boxes (a
            // short last chunk against a long kept list is otherwise a handful of lanes walking it alone)
            int mp = 1;
            while (mp < m) mp <<= 1;
            const int tpb = NMS_NT / mp, bi = tid / tpb, sub = tid - bi * tpb;
            if (bi < m) {
                const float4 bb = sbox[bi];
                bool hit = false;
                const int nks = min(nk, KEPT_SMEM);   // kept boxes cached in shared memory, the rest in global memory
                int k = sub;
                for (; k < nks && !hit; k += tpb) hit = iou_gt(skept[k], bb, thr_f);
                for (; k < nk && !hit; k += tpb) hit = iou_gt(kept_box[k], bb, thr_f);
                if (hit) sdeadb[bi] = 1;
            }
        }
        __syncthreads();
        const bool dead = sdeadb[tid] != 0;
        // upper-triangular IoU bits of the chunk, balanced: rows lo = min(t, m-1-t) and hi = m-1-lo have m - 1 columns
        // between them; thread lo takes the first m / 2 columns of its row, thread hi the rest of row lo and all of its own
        // row (the parts are OR-ed into smask atomically).  Every thread does ~m / 2 tests (the plain triangle: thread 0 does
        // m - 1, the last one none).
        unsigned long long m0 = 0, m1 = 0, p0 = 0, p1 = 0;
        if (tid < m) {
            const int part = m - 1 - tid, lo = min(tid, part), half = m >> 1;
            if (!dead) {   // own row: columns (tid, m), or the first `half` of them for the lower thread of a pair
                const int jend = (tid == lo && tid != part) ? tid + 1 + half : m;
                for (int j = tid + 1; j < jend; ++j) {
                    if (iou_gt(bx, sbox[j], thr_f)) {
                        if (j < 64) m0 |= 1ull << j;
                        else m1 |= 1ull << (j - 64);
                    }
                }
            }
            if (tid != lo) {   // upper thread of a pair: the rest of row lo
                if (!sdeadb[lo]) {
                    const float4 bl = sbox[lo];
                    for (int j = lo + 1 + half; j < m; ++j) {
                        if (iou_gt(bl, sbox[j], thr_f)) {
                            if (j < 64) p0 |= 1ull << j;
                            else p1 |= 1ull << (j - 64);
                        }
                    }
                }
                if (p0) atomicOr(&smask[lo][0], p0);
                if (p1) atomicOr(&smask[lo][1], p1);
            }
            if (m0) atomicOr(&smask[tid][0], m0);
            if (m1) atomicOr(&smask[tid][1], m1);
        }
        const unsigned d = __ballot_sync(0xffffffffu, dead);
        const unsigned any_mask = __ballot_sync(0xffffffffu, (m0 | m1 | p0 | p1) != 0ull);
        if (lane == 0) { sdead[wid] = d; sany[wid] = any_mask; }
        __syncthreads();
        // kept set of this chunk as a 128-bit map (keptw): without intra-chunk overlaps it is simply the boxes
        // that survived the kept-list test; otherwise warp 0 walks the bitmask (greedy rule, in score order)
        if (wid == 0) {
            unsigned long long remv = 0;
            if (lane < 2) remv = (unsigned long long)sdead[2 * lane] | ((unsigned long long)sdead[2 * lane + 1] << 32);
            unsigned long long kept = 0;
            const bool overlaps = (sany[0] | sany[1] | sany[2] | sany[3]) != 0u;
            if (!overlaps) {
                kept = ~remv;
            } else {
                for (int w = 0; w < 2; ++w) {
                    while (true) {
                        const unsigned long long cur = __shfl_sync(0xffffffffu, remv, w);
                        const unsigned long long alive = ~cur;
                        if (!alive) break;
                        const int i = __ffsll((long long)alive) - 1;
                        if (lane < 2) remv |= smask[w * 64 + i][lane];
                        if (lane == w) { remv |= 1ull << i; kept |= 1ull << i; }
                    }
                }
            }
            if (lane < 2) skeptw[lane] = kept;
        }
        __syncthreads();
        {   // every thread appends its own box if kept: position = kept so far + kept boxes before it in the chunk
            const unsigned long long k0 = skeptw[0], k1 = skeptw[1];
            const bool mine = tid < 64 ? (k0 >> tid) & 1ull : (k1 >> (tid - 64)) & 1ull;
            if (mine) {
                const int before = tid < 64 ? __popcll(k0 & ((1ull << tid) - 1ull))
                                            : __popcll(k0) + __popcll(k1 & ((1ull << (tid - 64)) - 1ull));
                const int pos = nk + before;
                kept_row[pos] = (int)(K[base + tid] & 0xffffffffull);
                if (pos < KEPT_SMEM) skept[pos] = bx;
                else kept_box[pos] = bx;
            }
            if (tid == 0) s_nkept = nk + __popcll(k0) + __popcll(k1);
        }
        __syncthreads();
    }
    if (tid == 0) ws.kept_count[seg] = s_nkept;
    __syncthreads(); // shared state is reused by the next segment
}


// warp path for a segment of 2..32 boxes: shuffle bitonic sort of the keys (one per lane, +inf padding),
// then the sequential greedy rule with the current box broadcast from its lane
__device__ __forceinline__ void nms_segment_warp(int b, int c, int n, int rows, int nc, float thr_f, const NmsWs &ws)
{
    const int lane = threadIdx.x & 31;
    const size_t seg = (size_t)b * nc + c, ib = (size_t)b * rows;
    const int off = ws.seg_off[seg];
    unsigned long long key = lane < n ? ws.key_bucket[ib + off + lane] : ~0ull;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, j);
            const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
            key = take_min ? (other < key ? other : key) : (other > key ? other : key);
        }
    }
    const int row = (int)(key & 0xffffffffull);
    float4 bx = make_float4(0, 0, 0, 0);
    if (lane < n) bx = ws.box[ib + row];
    bool alive = lane < n;
    for (int i = 0; i < n - 1; ++i) {
        const unsigned mask = __ballot_sync(0xffffffffu, alive);
        if (!(mask >> i & 1u)) continue;
        float4 bi;
        bi.x = __shfl_sync(0xffffffffu, bx.x, i); bi.y = __shfl_sync(0xffffffffu, bx.y, i);
        bi.z = __shfl_sync(0xffffffffu, bx.z, i); bi.w = __shfl_sync(0xffffffffu, bx.w, i);
        if (lane > i && alive && iou_gt(bi, bx, thr_f)) alive = false;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, alive);
    if (alive) ws.kept_row[ib + off + __popc(mask & ((1u << lane) - 1u))] = row;
    if (lane == 0) ws.kept_count[seg] = __popc(mask);
}

// grid (ceil(nc / spc), bs), 4 warps: warp w < spc owns class blockIdx.x*spc + w (spc = 4 when there are many
// segments, 1 when the grid would otherwise not fill the GPU).  Segments of up to 32 boxes (the
// common case at detection thresholds) are finished by their warp; larger ones are then processed one
// after another by the whole CTA.
__global__ void __launch_bounds__(NMS_NT) nms_segment_kernel(int rows, int nc, float thr_f, int spc, NmsWs ws)
{
    __shared__ NmsSmem sm;
    __shared__ int big_n[NMS_NT / 32];
    const int b = blockIdx.y, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * spc + wid; // spc = segments (warps in use) per CTA: 4, or 1 for small grids
    int n = 0;
    if (wid < spc && c < nc) n = ws.hist[(size_t)b * nc + c];
    if (n == 1) {
        if (lane == 0) {
            const size_t ib = (size_t)b * rows;
            const int off = ws.seg_off[(size_t)b * nc + c];
            ws.kept_row[ib + off] = (int)(ws.key_bucket[ib + off] & 0xffffffffull);
            ws.kept_count[(size_t)b * nc + c] = 1;
        }
    } else if (n >= 2 && n <= 32) {
        nms_segment_warp(b, c, n, rows, nc, thr_f, ws);
    }
    if (lane == 0) big_n[wid] = n > 32 ? n : 0;
    __syncthreads();
    for (int w = 0; w < spc; ++w) {
        const int nb = big_n[w]; // uniform across the CTA
        if (nb) nms_segment_cta(b, blockIdx.x * spc + w, nb, rows, nc, thr_f, ws, sm);
    }
}

// ---- K5: per-image finish: kept-count scan, chained scan across images, output assembly -----------------
// grid = bs * slices CTAs.  A CTA takes its work item (image b, slice y) from a ticket counter, i.e. in the order in
// which CTAs actually start running; slice 0 of image b scans the image's kept counts and publishes the image total
// (img_total[b] = total + 1).  Every CTA obtains its output offset by summing the totals of all earlier images,
// waiting for the few that are not published yet: those belong to LOWER tickets, whose CTAs are already running, so the
// wait cannot starve whatever order the hardware dispatches CTAs in (and whatever else occupies the SMs).  It then
// writes the image's rows (reference detect.py:121,137) and optionally undoes the letterbox:
// yolo_correct_boxes (detect.py:140-142,147-165) is evaluated with numpy's promotion rules: binary64 when
// letterbox_image is set (except box_hw *= scale, rounded back to binary32), binary32 until the final
// multiply by the image shape otherwise; the result is stored as binary32.
struct CorrectParams {
    int enabled, letterbox, in_h, in_w;
    const int *image_hw;
    int image_hw_stride;
};

constexpr int FINISH_NT = 256;   // 256 x 47 registers: fits next to a resident head CTA (see TC_MAX_REGS)
constexpr int FINISH_SMEM_NC = 1024;

__global__ void __launch_bounds__(FINISH_NT) finish_kernel(int bs, int slices, int rows, int nc, NmsWs ws, int *__restrict__ out_counts,
                                                            int *__restrict__ out_offsets, float *__restrict__ out_rows,
                                                            int *__restrict__ out_idx, CorrectParams cp)
{
    __shared__ int s_red[FINISH_NT / 32];
    __shared__ int s_total, s_base;
    __shared__ int s_koff[FINISH_SMEM_NC + 1], s_soff[FINISH_SMEM_NC];
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(ws.ticket, 1);
    __syncthreads();
    const int b = s_ticket / slices, slice = s_ticket - b * slices;
    const size_t sb = (size_t)b * nc;
    const bool flat = nc <= FINISH_SMEM_NC;
    // exclusive scan of kept_count[b][:] -> kept_off (warp 0, 32 classes per step)
    if (wid == 0) {
        int running = 0;
        for (int c0 = 0; c0 < nc; c0 += 32) {
            const int c = c0 + lane;
            const int v = c < nc ? ws.kept_count[sb + c] : 0;
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (c < nc) {
                ws.kept_off[sb + c] = running + inc - v;
                if (flat) { s_koff[c] = running + inc - v; s_soff[c] = ws.seg_off[sb + c]; }
            }
            running += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) {
            s_total = running;
            if (flat) s_koff[nc] = running;
            if (slice == 0) atomicExch(&ws.img_total[b], running + 1); // publish
        }
    }
    // chained scan: sum of the totals of images 0..b-1
    int part = 0;
    for (int p = tid; p < b; p += FINISH_NT) {
        int v;
        while ((v = atomicAdd(&ws.img_total[p], 0)) == 0) __nanosleep(40);
        part += v - 1;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) s_red[wid] = part;
    __syncthreads();
    if (tid == 0) {
        int base = 0;
        for (int w = 0; w < FINISH_NT / 32; ++w) base += s_red[w];
        s_base = base;
        if (slice == 0) {
            out_counts[b] = s_total;
            out_offsets[b] = base;
            if (b == bs - 1) out_offsets[bs] = base + s_total;
        }
    }
    __syncthreads(); // also makes warp 0's kept_off writes visible to the block
    const int img_base = s_base;
    if (s_total == 0 || slice * FINISH_NT >= s_total) return;

    double off[2] = {0, 0}, scl[2] = {1, 1}, ims[2] = {1, 1};
    if (cp.enabled) {
        const int *hw = cp.image_hw + (size_t)b * cp.image_hw_stride;
        ims[0] = (double)hw[0]; ims[1] = (double)hw[1];
        if (cp.letterbox) {
            const double ins[2] = {(double)cp.in_h, (double)cp.in_w};
            const double r = fmin(ins[0] / ims[0], ins[1] / ims[1]);
            for (int d = 0; d < 2; ++d) {
                const double ns = rint(ims[d] * r);
                off[d] = (ins[d] - ns) / 2.0 / ins[d];
                scl[d] = ins[d] / ns;
            }
        }
    }
    const size_t ib = (size_t)b * rows;
    const int total = s_total;
    // one thread per output row: class found by binary search over the kept offsets (flat mode), or one warp per
    // class when the class table does not fit shared memory
    auto write_row = [&](int c, int row, size_t oidx) {
        const float4 bx = ws.box[ib + row];
        const float2 oc = ws.oc[ib + row];
        float o0 = bx.x, o1 = bx.y, o2 = bx.z, o3 = bx.w;
        if (cp.enabled) {
            const float yx[2] = {__fmul_rn(__fadd_rn(bx.y, bx.w), 0.5f), __fmul_rn(__fadd_rn(bx.x, bx.z), 0.5f)};
            const float hw[2] = {__fsub_rn(bx.w, bx.y), __fsub_rn(bx.z, bx.x)};
            float mn[2], mx[2];
            if (cp.letterbox) {
                for (int d = 0; d < 2; ++d) {
                    const double v = __dmul_rn(__dsub_rn((double)yx[d], off[d]), scl[d]);
                    const float h32 = (float)__dmul_rn((double)hw[d], scl[d]);
                    const double half = (double)__fmul_rn(h32, 0.5f);
                    mn[d] = (float)__dmul_rn(__dsub_rn(v, half), ims[d]);
                    mx[d] = (float)__dmul_rn(__dadd_rn(v, half), ims[d]);
                }
            } else {
                for (int d = 0; d < 2; ++d) {
                    const float half = __fmul_rn(hw[d], 0.5f);
                    mn[d] = (float)__dmul_rn((double)__fsub_rn(yx[d], half), ims[d]);
                    mx[d] = (float)__dmul_rn((double)__fadd_rn(yx[d], half), ims[d]);
                }
            }
            o0 = mn[0]; o1 = mn[1]; o2 = mx[0]; o3 = mx[1];
        }
        float *o = out_rows + oidx * 7;
        o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3;
        o[4] = oc.x; o[5] = oc.y; o[6] = (float)c;
        out_idx[oidx] = row;
    };
    // `slices` CTAs share the rows of an image (one per image at detection thresholds; several when most of the
    // 25 200 rows of a low-threshold image survive): slice y takes the k with (k / FINISH_NT) % slices == y
    if (flat) {
        for (int k = slice * FINISH_NT + tid; k < total; k += FINISH_NT * slices) {
            int lo = 0, hi = nc; // last c with s_koff[c] <= k (empty classes share their successor's offset)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_koff[mid] <= k) lo = mid;
                else hi = mid;
            }
            write_row(lo, ws.kept_row[ib + s_soff[lo] + (k - s_koff[lo])], (size_t)img_base + k);
        }
    } else {
        for (int c = wid + slice * (FINISH_NT / 32); c < nc; c += (FINISH_NT / 32) * slices) {
            const int nk = ws.kept_count[sb + c];
            if (nk == 0) continue;
            const int *kept_row = ws.kept_row + ib + ws.seg_off[sb + c];
            const size_t obase = (size_t)img_base + ws.kept_off[sb + c];
            for (int k = lane; k < nk; k += 32) write_row(c, kept_row[k], obase + k);
        }
    }
}

// ---- torchvision.ops.nms drop-in for one box set ---------------------------------------------------
__global__ void single_prepare_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, int n, NmsWs ws)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { ws.hist[0] = n; ws.seg_off[0] = 0; }
    if (i >= n) return;
    ws.box[i] = make_float4(boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]);
    ws.key_bucket[i] = make_key(scores[i], i);
}

__global__ void single_finish_kernel(NmsWs ws, int *keep, int *keep_count)
{
    const int nk = ws.kept_count[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += gridDim.x * blockDim.x) keep[i] = ws.kept_row[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) *keep_count = nk;
}

static float thr_to_f32_floor(double thr)
{
    float tf = (float)thr;
    if ((double)tf > thr) tf = nextafterf(tf, -INFINITY);
    return tf;
}

int launch_nms_tail(const yc_nms_params *p, const NmsWs &ws, float *out_rows, int *out_idx, int *out_counts,
                    int *out_offsets, cudaStream_t stream)
{
    bucket_kernel<<<dim3((p->rows + BUCKET_SLOTS - 1) / BUCKET_SLOTS, p->bs), 256, 0, stream>>>(p->rows, p->nc, ws);
    const int spc = (long long)p->nc * p->bs >= 8 * 148 ? NMS_NT / 32 : 1;
    nms_segment_kernel<<<dim3((p->nc + spc - 1) / spc, p->bs), NMS_NT, 0, stream>>>(p->rows, p->nc,
                                                                                     thr_to_f32_floor(p->nms_thres), spc, ws);
    CorrectParams cp{p->correct_boxes, p->letterbox, p->input_h, p->input_w, (const int *)p->image_hw, p->image_hw_stride};
    // slices per image: enough CTAs to cover the GPU when the batch alone does not
    int slices = 1;
    if (p->bs < 32)
        while (slices < 16 && (long long)p->bs * slices < 148 && (long long)slices * FINISH_NT * 4 < p->rows) slices *= 2;
    finish_kernel<<<p->bs * slices, FINISH_NT, 0, stream>>>(p->bs, slices, p->rows, p->nc, ws, out_counts, out_offsets,
                                                            out_rows, out_idx, cp);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

} // namespace yc

using namespace yc;

extern "C" size_t yc_nms_workspace_bytes(int bs, int rows, int nc)
{
    if (bs <= 0 || rows <= 0 || nc <= 0) return 0;
    return carve(nullptr, bs, rows, nc).total_bytes + 256;
}

extern "C" int yc_nms_batched(float *pred, const yc_nms_params *p, void *workspace, size_t workspace_bytes,
                              float *out_rows, int32_t *out_idx, int32_t *out_counts, int32_t *out_offsets,
                              yc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    YC_REQUIRE(p && pred && workspace && out_rows && out_idx && out_counts && out_offsets, YC_ERR_INVALID,
               "yc_nms_batched: null argument");
    YC_REQUIRE(p->bs > 0 && p->rows > 0 && p->nc > 0 && p->row_stride >= 5 + p->nc, YC_ERR_INVALID,
               "yc_nms_batched: bad shape bs=%d rows=%d nc=%d row_stride=%d", p->bs, p->rows, p->nc, p->row_stride);
    YC_REQUIRE((size_t)p->bs * p->rows < (size_t)1 << 31, YC_ERR_UNSUPPORTED, "yc_nms_batched: bs*rows >= 2^31");
    YC_REQUIRE(p->bs <= 65535, YC_ERR_UNSUPPORTED, "yc_nms_batched: bs > 65535");
    YC_REQUIRE(!p->correct_boxes || p->image_hw, YC_ERR_INVALID, "yc_nms_batched: correct_boxes needs image_hw");
    void *base = (void *)round_up_sz((size_t)workspace, 256);
    NmsWs ws = carve(base, p->bs, p->rows, p->nc);
    YC_REQUIRE(ws.total_bytes + ((char *)base - (char *)workspace) <= workspace_bytes, YC_ERR_WORKSPACE,
               "yc_nms_batched: workspace %zu < %zu", workspace_bytes, ws.total_bytes + 256);
    const size_t tile_bytes = (size_t)TC_ROWS * p->row_stride * 4;
    YC_REQUIRE(tile_bytes <= 200 * 1024, YC_ERR_UNSUPPORTED, "yc_nms_batched: row_stride %d too large", p->row_stride);

    YC_CUDA(cudaMemsetAsync(ws.counters, 0, ws.counters_bytes, stream));
    if (tile_bytes > 48 * 1024)
        YC_CUDA(cudaFuncSetAttribute(threshold_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tile_bytes));
    dim3 g1((p->rows + TC_ROWS - 1) / TC_ROWS, p->bs);
    threshold_compact_kernel<<<g1, TC_ROWS, tile_bytes, stream>>>(pred, p->rows, p->row_stride, p->nc, p->conf_thres,
                                                                  p->write_corners, p->box_div_w, p->box_div_h, ws);
    return launch_nms_tail(p, ws, out_rows, out_idx, out_counts, out_offsets, stream);
}

extern "C" int yc_nms_single(const float *boxes, const float *scores, int n, double thr, void *workspace,
                             size_t workspace_bytes, int32_t *keep, int32_t *keep_count_dev, yc_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    YC_REQUIRE(keep_count_dev && workspace, YC_ERR_INVALID, "yc_nms_single: null argument");
    if (n <= 0) {
        YC_CUDA(cudaMemsetAsync(keep_count_dev, 0, sizeof(int), stream));
        return YC_OK;
    }
    YC_REQUIRE(boxes && scores && keep, YC_ERR_INVALID, "yc_nms_single: null argument");
    void *base = (void *)round_up_sz((size_t)workspace, 256);
    NmsWs ws = carve(base, 1, n, 1);
    YC_REQUIRE(ws.total_bytes + ((char *)base - (char *)workspace) <= workspace_bytes, YC_ERR_WORKSPACE,
               "yc_nms_single: workspace %zu < %zu", workspace_bytes, ws.total_bytes + 256);
    YC_CUDA(cudaMemsetAsync(ws.counters, 0, ws.counters_bytes, stream));
    single_prepare_kernel<<<(n + 255) / 256, 256, 0, stream>>>(boxes, scores, n, ws);
    nms_segment_kernel<<<dim3(1, 1), NMS_NT, 0, stream>>>(n, 1, thr_to_f32_floor(thr), 1, ws);
    single_finish_kernel<<<min((n + 255) / 256, 64), 256, 0, stream>>>(ws, keep, keep_count_dev);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}
