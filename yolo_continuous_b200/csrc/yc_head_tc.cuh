// yc_head_tc.cuh -- parameter blocks, tile scheduler and epilogue device functions shared by the 1-CTA
// (yc_head_sm100.cu) and the 2-CTA / cta_group::2 (yc_head_sm100_2cta.cu) head kernels.
#pragma once
#include <stdlib.h>

#include "yc_common.cuh"
#include "yc_nms.cuh"
#include "yc_sm100.cuh"

#ifdef YC_EXPERIMENT_NO_SB
#define YC_SB(p) make_float2(1.0f, 0.0f)
#else
#define YC_SB(p) __ldg(p)
#endif

namespace yc {

using namespace sm100;

constexpr int TC_BM = 128;       // pixels per tile (UMMA M)
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_MAX_N = 256;
// k per pipeline stage is a template parameter BK (64 or 128):
//   A stage = two {64 px, BK k} bf16 boxes (BK*128 B each); B stage = BK/64 boxes {64 k, Npad} (32 KB each)
constexpr int TC_B_BOX_BYTES = TC_MAX_N * 64 * 2;    // 32 KB
constexpr int TC_TMEM_COLS = 512;
#ifndef T2_BK
#define T2_BK 64                 // k per feature-map stage of the CTA-pair kernel (yc_head_sm100_2cta.cu)
#endif
constexpr int TC_NON_EPI_THREADS = 128;
// Register cap of the head kernels: 512 threads x 96 leave a quarter of the SM's register file (and ~19 KB of its
// shared memory) to the NMS kernels of the previous batch, which run next to the head kernel on a second stream.
#ifndef TC_MAX_REGS
#define TC_MAX_REGS 96
#endif

struct TcLevel {
    const float2 *sb;   // (scale, bias2) per column
    float *raw;         // [bs, na, HW, no] or null
    int n_groups;       // anchor groups per pixel tile: 1 (all anchors in one 256-column MMA tile) or na (IBin: one
                        // 128-column MMA tile per anchor); tiles of a level are ordered group major
    int bmap0;          // first weight tensor map of this level (one per group)
    int K, HW, nx;
    int tiles_per_img;  // ceil(HW / 128)
    int boxes_per_img;  // ceil(HW / 64): 64-pixel TMA boxes per image (CTA-pair kernel: a tile is 4 consecutive boxes of the
                        // level's box list, across image boundaries, so a 400-pixel P5 map wastes 12 % of a tile, not 28 %)
    int n_boxes;        // bs * boxes_per_img
    int tile_begin;     // first tile id of this level in schedule order
    int chunk_tiles;    // n_groups > 1: pixel tiles per chunk of the tile order (see tile_coord_w)
    int row_off;        // first z row of this level
    float stride, stride_y;
    float anchor_wh[YC_MAX_ANCHORS * 2];
};

struct TcParams {
    TcLevel lv[YC_MAX_LEVELS]; // in schedule order (largest K first)
    int n_lv;
    int total_tiles;
    int bs, na, no, npad;      // na = anchors per MMA tile, no = accumulator columns per anchor
    int npad_g;                // split kernel: columns per anchor in the transposed weight copies (round_up(no, 16))
    int epi_warps;             // epilogue warps: 4 per anchor of the tile; IBin (one anchor per tile, 127 sigmoids per row):
                               // 12 = 3 per TMEM lane quadrant, each taking a third of the columns
    int na_real;               // anchors of the head (raw map indexing)
    int no_out;                // columns of a z row (no; IBin: nc + 5)
    int ibin, bin_count;       // IBin decode (nets/ibin.py:56-72)
    float bin_step;
    const float *bins;         // device [bin_count] (SigmoidBin.bins)
    int rows_total;
    int write_z;               // 0 for YC_HEAD_RAW
    float *z;
    uint32_t idesc;
    uint32_t b_box_bytes;      // npad * 64 * 2: bytes one weight box brings
    uint32_t b_slot_bytes;     // room one weight box takes in a stage of the 1-CTA kernel (b_box_bytes rounded up to 1 KB)
    uint32_t slab_bytes;       // per epilogue warp: 32*no*4 (z slab) or TC_QUEUE_ROWS*nc*4 (fused survivor queue)
    int debug;                 // YC_TC_DEBUG bits (timing experiments only): 1 skip epilogue work, 2 skip MMA issue, 4 skip TMA,
                               // 128 z epilogue without the TMEM reads / decode (stores only), 256 without the stores
    int stages;                // depth of the smem ring
    int a_hint;                // 1 = feature-map loads carry the L2 evict-first policy (CTA-pair kernel; YC_TC_AHINT overrides)
    int half_off;              // > 0: z / raw rows are produced by halves (store_rows_half<half_off>), see there
    int tab_entries;           // (scale, bias) pairs staged in shared memory: n_lv * na_real * no (half-row epilogue)
    int a_kmajor;              // feature maps are channels-last: A is a K-major operand
    // fused mode (yc_detect_fused): the epilogue thresholds and emits NMS candidates, z is never written
    int fused;
    int nc;
    float conf, div_w, div_h;
    NmsWs ws;
};

struct TcMaps {
    CUtensorMap a[YC_MAX_LEVELS];
    CUtensorMap b[YC_MAX_LEVELS * YC_MAX_ANCHORS];
};

struct TileCoord { int lv, b, p0, g; };

// tile id -> (level, image, first pixel, anchor group); `width` = pixels per tile (128, or 256 for a CTA pair)
__host__ __device__ __forceinline__ TileCoord tile_coord_w(const TcParams &P, int t, int width)
{
    int l = 0;
#pragma unroll
    for (int i = 1; i < YC_MAX_LEVELS; ++i)
        if (i < P.n_lv && t >= P.lv[i].tile_begin) l = i;
    int r = t - P.lv[l].tile_begin;
    TileCoord c;
    c.lv = l;
    c.g = 0;
    if (P.lv[l].n_groups > 1) {
        // One anchor group per tile (IBin: 127 of 128 columns per anchor).  Group major inside CHUNKS of pixel tiles: a
        // chunk's pixel tiles are walked once per anchor group, so a CTA keeps a group's weights for several tiles, and a
        // chunk's feature maps (sized to stay in L2) come from HBM once instead of once per group.
        const int n_px = P.bs * P.lv[l].tiles_per_img, ct = P.lv[l].chunk_tiles;
        const int ch = r / (ct * P.lv[l].n_groups);
        r -= ch * ct * P.lv[l].n_groups;
        const int cs = min(ct, n_px - ch * ct);   // pixel tiles of this chunk (the last one may be short)
        c.g = r / cs;
        r = ch * ct + r - c.g * cs;
    }
    c.b = r / P.lv[l].tiles_per_img;
    c.p0 = (r - c.b * P.lv[l].tiles_per_img) * width;
    return c;
}

__device__ __forceinline__ TileCoord tile_coord(const TcParams &P, int t) { return tile_coord_w(P, t, TC_BM); }

// CTA-pair kernels: tile id -> (level, first 64-pixel box of the tile in the level's flat box list)
struct BoxTile { int lv, j0; };
__device__ __forceinline__ BoxTile box_tile(const TcParams &P, int t)
{
    int l = 0;
#pragma unroll
    for (int i = 1; i < YC_MAX_LEVELS; ++i)
        if (i < P.n_lv && t >= P.lv[i].tile_begin) l = i;
    BoxTile c;
    c.lv = l;
    c.j0 = (t - P.lv[l].tile_begin) * 4;
    return c;
}
// box j of a level -> (image, first pixel); image = bs (out of range: TMA zero-fills, the epilogue skips) past the end
__device__ __forceinline__ void box_coord(const TcLevel &L, int bs, int j, int &b, int &p0)
{
    if (j >= L.n_boxes) { b = bs; p0 = 0; return; }
    b = j / L.boxes_per_img;
    p0 = (j - b * L.boxes_per_img) * 64;
}

// fp32-grade mode (yc_head_sm100_split.cu): a tile owns two accumulators, the main one (x_hi * w_hi) and, TC_SPLIT_CORR
// columns further, the correction (x_hi * w_lo + x_lo * w_hi); the epilogue adds them in IEEE binary32.
constexpr int TC_SPLIT_CORR = 128;

// this thread's W accumulator values starting at `taddr` (SPLIT: main + correction)
template <int W, bool SPLIT>
__device__ __forceinline__ void ld_acc(uint32_t taddr, uint32_t *v)
{
    TmemLd<W>::ld(taddr, v);
    if (SPLIT) {
        uint32_t u[W];
        TmemLd<W>::ld(taddr + (uint32_t)TC_SPLIT_CORR, u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < W; ++j) v[j] = __float_as_uint(__fadd_rn(__uint_as_float(v[j]), __uint_as_float(u[j])));
    } else {
        tmem_ld_wait();
    }
}

// Epilogue for W consecutive accumulator columns [c0, c0+W) of this thread's row.
//   RAW:   slab[o] = t                     (pre-sigmoid map, forward()'s list `x`)
//   !RAW:  slab[o] = decode(sigmoid(t))    (z row; o<2 xy, o<4 wh: nets/idetect.py:40-42)
template <int W, bool RAW, bool SPLIT = false>
__device__ __forceinline__ void epi_chunk(uint32_t taddr, int c0, const float2 *__restrict__ sb, float *__restrict__ srow,
                                          float gx, float gy, float stride, float stride_y, float aw, float ah)
{
    uint32_t v[W];
    ld_acc<W, SPLIT>(taddr + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const float2 s_b = YC_SB(sb + c0 + j); // same address for the whole warp: one broadcast load
        const float t = fmaf(__uint_as_float(v[j]), s_b.x, s_b.y);
        float r;
        if (RAW) {
            r = t;
        } else {
            r = sigmoidf_fast(t);
            const int o = c0 + j;
            if (o == 0) r = decode_xy(r, gx, stride);
            else if (o == 1) r = decode_xy(r, gy, stride_y);
            else if (o == 2) r = decode_wh(r, aw);
            else if (o == 3) r = decode_wh(r, ah);
        }
        srow[c0 + j] = r;
    }
}

template <bool RAW, bool SPLIT = false>
__device__ __forceinline__ void epi_row(uint32_t taddr, int no, const float2 *__restrict__ sb, float *__restrict__ srow,
                                        float gx, float gy, float stride, float stride_y, float aw, float ah)
{
    int c0 = 0;
    for (; c0 + 16 <= no; c0 += 16) epi_chunk<16, RAW, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, aw, ah);
    const int rem = no - c0;
    if (rem & 8) { epi_chunk<8, RAW, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, aw, ah); c0 += 8; }
    if (rem & 4) { epi_chunk<4, RAW, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, aw, ah); c0 += 4; }
    if (rem & 2) { epi_chunk<2, RAW, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, aw, ah); c0 += 2; }
    if (rem & 1) { epi_chunk<1, RAW, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, aw, ah); }
}


// IBin z row (reference nets/ibin.py:56-72, losses/sigmoid_bin.py:49-63) from the 127 accumulator columns of one
// (pixel, anchor): [x, y | w: reg + bins | h: reg + bins | obj | cls] -> [x, y, w, h, obj, cls]; argmax over the
// sigmoided bins takes the first maximum.  Same operation order as ibin_decode_kernel (generic path).
template <int W, bool SPLIT = false>
__device__ __forceinline__ void ibin_chunk(uint32_t taddr, int c0, const float2 *__restrict__ sb, float *__restrict__ srow,
                                           float gx, float gy, float stride, float stride_y, int len, float &reg_w, float &reg_h, float &best_w,
                                           float &best_h, int &idx_w, int &idx_h)
{
    uint32_t v[W];
    ld_acc<W, SPLIT>(taddr + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const int o = c0 + j;
        const float2 s_b = YC_SB(sb + o);
        const float sg = sigmoidf_fast(fmaf(__uint_as_float(v[j]), s_b.x, s_b.y));
        if (o < 2) {
            srow[o] = o == 0 ? decode_xy(sg, gx, stride) : decode_xy(sg, gy, stride_y);
        } else if (o < 2 + 2 * len) {
            if (o < 2 + len) {
                const int k = o - 2;
                if (k == 0) reg_w = sg;
                else if (sg > best_w) { best_w = sg; idx_w = k - 1; }
            } else {
                const int k = o - 2 - len;
                if (k == 0) reg_h = sg;
                else if (sg > best_h) { best_h = sg; idx_h = k - 1; }
            }
        } else {
            srow[o - 2 * len + 2] = sg;
        }
    }
}

// sigmoid of the accumulator columns [cb, ce) (objectness / classes) into srow[o - shift]: no per-column case analysis
template <int W, bool SPLIT = false>
__device__ __forceinline__ void sig_chunk(uint32_t taddr, int c0, const float2 *__restrict__ sb, float *__restrict__ srow, int shift)
{
    uint32_t v[W];
    ld_acc<W, SPLIT>(taddr + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const float2 s_b = YC_SB(sb + c0 + j);
        srow[c0 + j - shift] = sigmoidf_fast(fmaf(__uint_as_float(v[j]), s_b.x, s_b.y));
    }
}
// (SPLIT reads two accumulators per column: 16-column chunks keep the register count of the 32-column ones)
template <bool SPLIT = false>
__device__ __forceinline__ void epi_range_sig(uint32_t taddr, int cb, int ce, const float2 *__restrict__ sb, float *__restrict__ srow,
                                              int shift)
{
    constexpr int MW = SPLIT ? 16 : 32;
    int c0 = cb;
    for (; c0 + MW <= ce; c0 += MW) sig_chunk<MW, SPLIT>(taddr, c0, sb, srow, shift);
    const int rem = ce - c0;
    if (!SPLIT && (rem & 16)) { sig_chunk<16, SPLIT>(taddr, c0, sb, srow, shift); c0 += 16; }
    if (rem & 8) { sig_chunk<8, SPLIT>(taddr, c0, sb, srow, shift); c0 += 8; }
    if (rem & 4) { sig_chunk<4, SPLIT>(taddr, c0, sb, srow, shift); c0 += 4; }
    if (rem & 2) { sig_chunk<2, SPLIT>(taddr, c0, sb, srow, shift); c0 += 2; }
    if (rem & 1) { sig_chunk<1, SPLIT>(taddr, c0, sb, srow, shift); }
}

// columns [cb, ce) of the IBin row's box part; final_w / final_h: this range held all of the w / h block, finish it
template <bool SPLIT = false>
__device__ __forceinline__ void epi_range_ibin(uint32_t taddr, int cb, int ce, bool final_w, bool final_h, const float2 *__restrict__ sb,
                                               float *__restrict__ srow, float gx, float gy, float stride, float stride_y,
                                               float aw, float ah, const TcParams &P)
{
    const int len = P.bin_count + 1;
    float reg_w = 0.f, reg_h = 0.f, best_w = -1.f, best_h = -1.f;
    int idx_w = 0, idx_h = 0;
    constexpr int MW = SPLIT ? 16 : 32;
    int c0 = cb;
    for (; c0 + MW <= ce; c0 += MW) ibin_chunk<MW, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, len, reg_w, reg_h, best_w, best_h, idx_w, idx_h);
    const int rem = ce - c0;
    if (!SPLIT && (rem & 16)) { ibin_chunk<16, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, len, reg_w, reg_h, best_w, best_h, idx_w, idx_h); c0 += 16; }
    if (rem & 8) { ibin_chunk<8, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, len, reg_w, reg_h, best_w, best_h, idx_w, idx_h); c0 += 8; }
    if (rem & 4) { ibin_chunk<4, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, len, reg_w, reg_h, best_w, best_h, idx_w, idx_h); c0 += 4; }
    if (rem & 2) { ibin_chunk<2, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, len, reg_w, reg_h, best_w, best_h, idx_w, idx_h); c0 += 2; }
    if (rem & 1) { ibin_chunk<1, SPLIT>(taddr, c0, sb, srow, gx, gy, stride, stride_y, len, reg_w, reg_h, best_w, best_h, idx_w, idx_h); }
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        if (!(d == 0 ? final_w : final_h)) continue;
        float r = __fmul_rn(d == 0 ? reg_w : reg_h, 2.0f);
        r = __fadd_rn(r, -1.0f);
        r = __fmul_rn(r, P.bin_step);
        float res = __fadd_rn(r, __ldg(P.bins + (d == 0 ? idx_w : idx_h)));
        res = fminf(fmaxf(res, 0.0f), 4.0f);
        srow[2 + d] = __fmul_rn(res, d == 0 ? aw : ah);
    }
}

// raw columns [cb, ce) of this thread's row into the slab row
template <bool SPLIT = false>
__device__ __forceinline__ void epi_range_raw(uint32_t taddr, int cb, int ce, const float2 *__restrict__ sb, float *__restrict__ srow)
{
    constexpr int MW = SPLIT ? 16 : 32;
    int c0 = cb;
    for (; c0 + MW <= ce; c0 += MW) epi_chunk<MW, true, SPLIT>(taddr, c0, sb, srow, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
    const int rem = ce - c0;
    if (!SPLIT && (rem & 16)) { epi_chunk<16, true, SPLIT>(taddr, c0, sb, srow, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f); c0 += 16; }
    if (rem & 8) { epi_chunk<8, true, SPLIT>(taddr, c0, sb, srow, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f); c0 += 8; }
    if (rem & 4) { epi_chunk<4, true, SPLIT>(taddr, c0, sb, srow, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f); c0 += 4; }
    if (rem & 2) { epi_chunk<2, true, SPLIT>(taddr, c0, sb, srow, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f); c0 += 2; }
    if (rem & 1) { epi_chunk<1, true, SPLIT>(taddr, c0, sb, srow, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f); }
}


// ---- fused epilogue (S3): class max / threshold / candidate emission straight from TMEM -----------------
// logits of W consecutive class columns starting at accumulator column c (class index c - 5)
template <int W, bool EXACT>
__device__ __forceinline__ void cls_chunk(uint32_t taddr, int c, const float2 *__restrict__ sb, float &bestv, int &besti,
                                          int cbase = 5)
{
    uint32_t v[W];
    TmemLd<W>::ld(taddr + (uint32_t)c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const float2 s_b = YC_SB(sb + c + j);
        float t = fmaf(__uint_as_float(v[j]), s_b.x, s_b.y);
        if (EXACT) {
            t = sigmoidf_fast(t); // compare what z would hold (first maximum of the sigmoids, torch.max)
            if (t > bestv) { bestv = t; besti = c + j - cbase; }
        } else {
            bestv = fmaxf(bestv, t); // quick pass: only the largest class logit is needed
        }
    }
}

// classes live in the accumulator columns [cbase, no) (IDetect: 5; IBin: behind the two bin blocks and the objectness)
template <bool EXACT>
__device__ __forceinline__ void cls_scan(uint32_t taddr, int no, const float2 *__restrict__ sb, float &bestv, int &besti,
                                         int cbase = 5)
{
    int c = cbase;
    for (; c + 16 <= no; c += 16) cls_chunk<16, EXACT>(taddr, c, sb, bestv, besti, cbase);
    const int rem = no - c;
    if (rem & 8) { cls_chunk<8, EXACT>(taddr, c, sb, bestv, besti, cbase); c += 8; }
    if (rem & 4) { cls_chunk<4, EXACT>(taddr, c, sb, bestv, besti, cbase); c += 4; }
    if (rem & 2) { cls_chunk<2, EXACT>(taddr, c, sb, bestv, besti, cbase); c += 2; }
    if (rem & 1) { cls_chunk<1, EXACT>(taddr, c, sb, bestv, besti, cbase); }
}


// survivors of the objectness filter park their raw class accumulators in the warp's shared-memory queue
template <int W>
__device__ __forceinline__ void queue_chunk(uint32_t taddr, int c, bool pass, float *__restrict__ qrow)
{
    uint32_t v[W];
    TmemLd<W>::ld(taddr + (uint32_t)c, v);
    tmem_ld_wait();
    if (pass) {
#pragma unroll
        for (int j = 0; j < W; ++j) qrow[c - 5 + j] = __uint_as_float(v[j]);
    }
}

__device__ __forceinline__ void queue_classes(uint32_t taddr, int no, bool pass, float *__restrict__ qrow)
{
    int c = 5;
    for (; c + 16 <= no; c += 16) queue_chunk<16>(taddr, c, pass, qrow);
    const int rem = no - c;
    if (rem & 8) { queue_chunk<8>(taddr, c, pass, qrow); c += 8; }
    if (rem & 4) { queue_chunk<4>(taddr, c, pass, qrow); c += 4; }
    if (rem & 2) { queue_chunk<2>(taddr, c, pass, qrow); c += 2; }
    if (rem & 1) { queue_chunk<1>(taddr, c, pass, qrow); }
}

// nc = 16 * NCH classes: all class accumulators of this thread's row into registers with ONE wait, so that the TMEM
// buffer can be handed back before anything is written to the queue
template <int NCH>
__device__ __forceinline__ void grab_classes(uint32_t taddr, bool pass, float *__restrict__ qrow, uint64_t *tempty, int lane,
                                             bool pair)
{
    // at most 48 accumulators in registers at a time (the kernel is capped at TC_MAX_REGS registers)
    // (as few, as wide loads as possible: the cost is per instruction)
    constexpr int H1 = NCH > 3 ? 3 : NCH, H2 = NCH - H1;
    uint32_t v[H1 * 16];
    if (H1 == 3) {
        TmemLd<32>::ld(taddr + 5u, v);
        TmemLd<16>::ld(taddr + 5u + 32u, v + 32);
    } else if (H1 == 2) {
        TmemLd<32>::ld(taddr + 5u, v);
    } else {
        TmemLd<16>::ld(taddr + 5u, v);
    }
    tmem_ld_wait();
    if (H2 > 0) {
        if (pass) {
#pragma unroll
            for (int j = 0; j < H1 * 16; ++j) qrow[j] = __uint_as_float(v[j]);
        }
        if (H2 == 3) {
            TmemLd<32>::ld(taddr + 5u + 16u * H1, v);
            TmemLd<16>::ld(taddr + 5u + 16u * H1 + 32u, v + 32);
        } else if (H2 == 2) {
            TmemLd<32>::ld(taddr + 5u + 16u * H1, v);
        } else {
            TmemLd<16>::ld(taddr + 5u + 16u * H1, v);
        }
        tmem_ld_wait();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (pair) mbar_arrive_leader(tempty);
        else mbar_arrive(tempty);
    }
    if (pass) {
#pragma unroll
        for (int j = 0; j < (H2 > 0 ? H2 : H1) * 16; ++j) qrow[(H2 > 0 ? H1 * 16 : 0) + j] = __uint_as_float(v[j]);
    }
    __syncwarp();
}

constexpr int TC_QUEUE_ROWS = 4; // survivors per warp and tile handled through the queue; more -> in-register scan

// warp-cooperative write of `nv` finished rows from the slab to global memory
__device__ __forceinline__ void slab_store(float *__restrict__ gdst, const float *__restrict__ slab, int nv, int no, int lane)
{
    const uint32_t bytes = (uint32_t)nv * no * 4u;
    fence_proxy_async_smem();
    __syncwarp();
    if ((((uintptr_t)gdst | bytes) & 15u) == 0) {
        if (lane == 0) {
            bulk_store(gdst, slab, bytes);
            bulk_commit();
        }
    } else { // unaligned span (odd shapes): plain coalesced stores
        for (int i = lane; i < nv * no; i += 32) gdst[i] = slab[i];
    }
}


// Non-fused epilogue of one epilogue warp for one tile: TMEM -> scale/bias -> sigmoid / decode (or raw, or IBin) -> slab in
// shared memory -> bulk store of the warp's rows (one contiguous span of z / of the raw map).  `e` = epilogue warp index,
// q = TMEM lane quadrant (warp % 4), g = anchor group of the tile, `tmem_tile` = first accumulator column of the tile's
// buffer.  The buffer is handed back through `tempty` as soon as this warp's TMEM reads are done.
// IBin (nets/ibin.py:56-72): 127 accumulator columns and as many sigmoids per row made the epilogue the bottleneck with
// one warp per quadrant (1.28 ms against 0.30 ms with the epilogue switched off), so the three warps of a quadrant split
// the columns [x y | w bins | h bins] [obj + a third of the classes] [rest] and fill ONE slab per quadrant, which the
// first of them stores (named barriers 1..4 and 5..8, 96 threads each).
template <bool SPLIT>
__device__ __forceinline__ void store_epilogue(const TcParams &P, const TcLevel &L, int b, int p0, int g, int e, int q, int lane,
                                               uint32_t tmem_tile, uint8_t *slabs, uint64_t *tempty, long long *pf = nullptr)
{
    // pf (timing experiments, YC_TC_DEBUG bit 8): cycles spent [0] waiting for the slab's previous store to be read out,
    // [1] in the TMEM reads + decode, [2] issuing the store, [3] IBin: in the named barriers
    long long c0_ = pf ? clock64() : 0;
    const int no = P.no, no_out = P.no_out;
    const int a = P.ibin ? 0 : e >> 2;          // anchor of the tile handled by this warp (IBin: one anchor per tile)
    const int prow0 = p0 + 32 * q;              // first pixel of this warp's 32 rows
    const int nv = min(32, L.HW - prow0);       // valid rows (<= 0: nothing to store)
    const int p = prow0 + lane;
    const float gx = (float)(p % L.nx), gy = (float)(p / L.nx);
    const int ar = g * P.na + a;                // anchor of the head this warp decodes
    const float aw = L.anchor_wh[2 * ar], ah = L.anchor_wh[2 * ar + 1];
    const float2 *sb = L.sb + ar * no;
    const uint32_t taddr = tmem_tile + ((uint32_t)(32 * q) << 16) + (uint32_t)(a * no);
    if (P.ibin) {
        const int part = e >> 2;
        const int len = P.bin_count + 1, s1 = 2 + 2 * len;       // s1: objectness column
        const int sc = s1 + (no - s1) / 3;                       // classes are shared 1/3 : 2/3 by parts 1 and 2
        float *zs = (float *)(slabs + (size_t)q * P.slab_bytes), *rs = zs + 32 * no_out;
        if (part == 0) {
            if (lane == 0) bulk_wait_read0();   // the previous stores from this quadrant's slab have been read out
            __syncwarp();
        }
        if (pf) { const long long c = clock64(); pf[0] += c - c0_; c0_ = c; }
        named_bar_sync(1 + q, 96);
        if (pf) { const long long c = clock64(); pf[3] += c - c0_; c0_ = c; }
        // part 0: [x y | w block]   part 1: [h block] + objectness and a third of the classes   part 2: the rest
        const int cb = part == 0 ? 0 : (part == 1 ? 2 + len : sc), ce = part == 0 ? 2 + len : (part == 1 ? sc : no);
        if (L.raw) epi_range_raw<SPLIT>(taddr, cb, ce, sb, rs + lane * no);
        if (P.write_z) {
            float *zrow = zs + lane * no_out;
            if (part == 0) epi_range_ibin<SPLIT>(taddr, 0, 2 + len, true, false, sb, zrow, gx, gy, L.stride, L.stride_y, aw, ah, P);
            if (part == 1) {
                epi_range_ibin<SPLIT>(taddr, 2 + len, s1, false, true, sb, zrow, gx, gy, L.stride, L.stride_y, aw, ah, P);
                epi_range_sig<SPLIT>(taddr, s1, sc, sb, zrow, 2 * len - 2);
            }
            if (part == 2) epi_range_sig<SPLIT>(taddr, sc, no, sb, zrow, 2 * len - 2);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);
        if (pf) { const long long c = clock64(); pf[1] += c - c0_; c0_ = c; }
        named_bar_sync(5 + q, 96);              // all three parts of the rows are in the slab
        if (pf) { const long long c = clock64(); pf[3] += c - c0_; c0_ = c; }
        if (part == 0 && nv > 0) {
            if (L.raw) slab_store(L.raw + (((size_t)b * P.na_real + ar) * L.HW + prow0) * no, rs, nv, no, lane);
            if (P.write_z)
                slab_store(P.z + ((size_t)b * P.rows_total + L.row_off + (size_t)ar * L.HW + prow0) * no_out, zs, nv,
                           no_out, lane);
        }
        if (pf) pf[2] += clock64() - c0_;
        return;
    }
    float *slab = (float *)(slabs + (size_t)e * P.slab_bytes);
    if (L.raw) {
        if (lane == 0) bulk_wait_read0(); // previous store from this slab has been read out
        __syncwarp();
        epi_row<true, SPLIT>(taddr, no, sb, slab + lane * no, gx, gy, L.stride, L.stride_y, aw, ah);
        if (nv > 0)
            slab_store(L.raw + (((size_t)b * P.na_real + ar) * L.HW + prow0) * no, slab, nv, no, lane);
    }
    if (P.write_z) {
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        if (pf) { const long long c = clock64(); pf[0] += c - c0_; c0_ = c; }
        if (!(P.debug & 128)) epi_row<false, SPLIT>(taddr, no, sb, slab + lane * no_out, gx, gy, L.stride, L.stride_y, aw, ah);
    }
    // all TMEM reads of this warp are done: hand the accumulator buffer back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty);
    if (pf) { const long long c = clock64(); pf[1] += c - c0_; c0_ = c; }
    if (P.write_z && nv > 0 && !(P.debug & 256))
        slab_store(P.z + ((size_t)b * P.rows_total + L.row_off + (size_t)ar * L.HW + prow0) * no_out, slab, nv,
                   no_out, lane);
    if (pf) pf[2] += clock64() - c0_;
}



// ---- z / raw epilogue by half rows ---------------------------------------------------------------------------------
// The whole-row form above gives a lane one row of the accumulator (tcgen05.ld .32x32b) and a warp a 32-row slab
// (32 * no * 4 bytes: 130 KB for the twelve epilogue warps of a COCO tile, which leaves the 1-CTA kernel two pipeline
// stages), reads TMEM in 16-column chunks with a wait after each, fetches (scale, bias) from global memory per column and
// writes the slab with generic stores.  Measured on the C2 batch (YC_TC_DEBUG bit 8): 81 cycles per column and warp, 60 %
// of what the two MUFU operations of a sigmoid allow, and the MMA warp waits for the 2-stage ring 68 % of the time.
// This form reads the accumulator with the .16x32bx2 shape: a pass covers 16 rows, lanes 0-15 own the columns [0, OFF) of
// their row and lanes 16-31 the columns [OFF, 2 OFF) (OFF ~ no / 2).  All of a thread's columns of a pass are fetched by
// a few wide loads behind ONE wait (at most 64 registers), so that the sigmoids of a pass are independent instructions
// and the TMEM buffer goes back to the MMA warp as soon as the second pass is in registers; the slab is 16 rows (half
// the shared memory: three stages instead of two); (scale, bias) come from a table in shared memory; the slab is
// written with st.shared.
constexpr int HG = 8;        // columns per group of independent sigmoids in store_rows_half (16: no change, 32: 5 % slower)
constexpr int HG_IBIN = 16;  // ... in store_rows_half_ibin (two warps per scheduler: 8 -> 16 gains 4 %, 32 loses it again)
template <int OFF>
__device__ __forceinline__ void ld_half_row(uint32_t ta, uint32_t *v)
{
    if constexpr ((OFF & 64) != 0) TmemLdHalf<64, OFF>::ld(ta, v);
    if constexpr ((OFF & 32) != 0) TmemLdHalf<32, OFF>::ld(ta + (OFF & 64), v + (OFF & 64));
    if constexpr ((OFF & 16) != 0) TmemLdHalf<16, OFF>::ld(ta + (OFF & 96), v + (OFF & 96));
    if constexpr ((OFF & 8) != 0) TmemLdHalf<8, OFF>::ld(ta + (OFF & 112), v + (OFF & 112));
    if constexpr ((OFF & 4) != 0) TmemLdHalf<4, OFF>::ld(ta + (OFF & 120), v + (OFF & 120));
    if constexpr ((OFF & 2) != 0) TmemLdHalf<2, OFF>::ld(ta + (OFF & 124), v + (OFF & 124));
    if constexpr ((OFF & 1) != 0) TmemLdHalf<1, OFF>::ld(ta + (OFF & 126), v + (OFF & 126));
}

// bulk store of `rows` finished rows (no floats each) of the 16-row slab; unaligned spans fall back to plain stores
__device__ __forceinline__ void half_slab_store(float *__restrict__ gdst, uint32_t slab_s, int rows, int no, int lane)
{
    const uint32_t bytes = (uint32_t)rows * no * 4u;
    fence_proxy_async_smem();
    __syncwarp();
    if (rows <= 0) return;
    if ((((uintptr_t)gdst | bytes) & 15u) == 0) {
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(slab_s), "r"(bytes) : "memory");
            bulk_commit();
        }
    } else {
        for (int i = lane; i < rows * no; i += 32) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(slab_s + 4u * i));
            gdst[i] = v;
        }
    }
}

// One warp, one tile: the 32 pixels [prow0, prow0 + 32) of image b (nv of them exist) for anchor `ar` of the head.
//   taddr   TMEM address of (first lane of the warp's quadrant, first accumulator column of the anchor)
//   tab_s   shared address of the anchor's (scale, bias) pairs;  slab_s  shared address of the warp's 16-row slab
//   PAIR    the TMEM-empty barrier lives in the pair's leader CTA
template <int OFF, bool PAIR>
__device__ __forceinline__ void store_rows_half(const TcParams &P, const TcLevel &L, int b, int prow0, int nv, int ar, uint32_t taddr,
                                                uint32_t tab_s, uint32_t slab_s, uint32_t dummy_s, uint64_t *tempty, int lane, long long *pf = nullptr)
{
    const int no = P.no;
    const int half = lane >> 4, r = lane & 15;
    const int cbase = half * OFF;                 // first column owned by this thread
    const float aw = L.anchor_wh[2 * ar], ah = L.anchor_wh[2 * ar + 1];
    const uint32_t trow = tab_s + (uint32_t)cbase * 8u;
    float *const srow = shared_f32(slab_s) + r * no + cbase;   // this thread's columns of its slab row
    float *const dummy = shared_f32(dummy_s);                 // where the columns that do not exist (lanes 16-31: next anchor) go
    const int ncols = half == 0 ? OFF : max(0, min(OFF, no - OFF));   // columns of this thread that exist
    long long c0_ = pf ? clock64() : 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        uint32_t v[OFF];
        ld_half_row<OFF>(taddr + ((uint32_t)(16 * pass) << 16), v);
        tmem_ld_wait();
        if (pass == 1) {   // all TMEM reads of this warp are done: hand the accumulator buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_leader(tempty);
                else mbar_arrive(tempty);
            }
        }
        if (pf) { const long long c = clock64(); pf[3] += c - c0_; c0_ = c; }
        const int rows = min(16, nv - 16 * pass);
        const int p = prow0 + 16 * pass + r;
        const size_t grow = (size_t)prow0 + 16 * pass;
        if (L.raw) {       // forward()'s list `x`: the pre-sigmoid maps
            if (lane == 0) bulk_wait_read0();   // the slab's previous store has been read out
            __syncwarp();
            if (pf) { const long long c = clock64(); pf[0] += c - c0_; c0_ = c; }
            // groups of HG columns: HG independent values, then their stores (through shared_f32 pointers, which the compiler
            // may schedule among the arithmetic; columns that do not exist go to the dummy word instead of a branch)
#pragma unroll
            for (int j0 = 0; j0 < OFF; j0 += HG) {
                float w[HG];
#pragma unroll
                for (int j = j0; j < j0 + HG && j < OFF; ++j) {
                    const float2 s_b = lds_f32x2(trow + 8u * j);
                    w[j - j0] = fmaf(__uint_as_float(v[j]), s_b.x, s_b.y);
                }
#pragma unroll
                for (int j = j0; j < j0 + HG && j < OFF; ++j)
                    *(j < ncols ? srow + j : dummy) = w[j - j0];
            }
            if (pf) { const long long c = clock64(); pf[1] += c - c0_; c0_ = c; }
            half_slab_store(L.raw + (((size_t)b * P.na_real + ar) * L.HW + grow) * no, slab_s, rows, no, lane);
            if (pf) { const long long c = clock64(); pf[2] += c - c0_; c0_ = c; }
        }
        if (P.write_z) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            if (pf) { const long long c = clock64(); pf[0] += c - c0_; c0_ = c; }
            const float gx = (float)(p % L.nx), gy = (float)(p / L.nx);
#pragma unroll
            for (int j0 = 0; j0 < OFF; j0 += HG) {
                float w[HG];
#pragma unroll
                for (int j = j0; j < j0 + HG && j < OFF; ++j) {
                    const float2 s_b = lds_f32x2(trow + 8u * j);
                    float sg = sigmoidf_fast(fmaf(__uint_as_float(v[j]), s_b.x, s_b.y));
                    if (j < 4 && half == 0) {   // box columns (nets/idetect.py:41-42); OFF >= 4, so they belong to the lower half
                        if (j == 0) sg = decode_xy(sg, gx, L.stride);
                        else if (j == 1) sg = decode_xy(sg, gy, L.stride_y);
                        else if (j == 2) sg = decode_wh(sg, aw);
                        else sg = decode_wh(sg, ah);
                    }
                    w[j - j0] = sg;
                }
#pragma unroll
                for (int j = j0; j < j0 + HG && j < OFF; ++j)
                    *(j < ncols ? srow + j : dummy) = w[j - j0];
            }
            if (pf) { const long long c = clock64(); pf[1] += c - c0_; c0_ = c; }
            half_slab_store(P.z + ((size_t)b * P.rows_total + L.row_off + (size_t)ar * L.HW + grow) * no, slab_s, rows, no, lane);
            if (pf) { const long long c = clock64(); pf[2] += c - c0_; c0_ = c; }
        }
    }
}

// IBin rows by halves (nets/ibin.py:56-72, losses/sigmoid_bin.py:49-63): one anchor per tile, the accumulator row is
// [x y | w: reg + bins | h: reg + bins | obj | cls] = no <= 128 columns, the z row [x y w h obj cls] = no - 2 LEN + 2.  EIGHT
// epilogue warps: warp (quadrant q, pass) owns the 16 rows 32 q + 16 pass .. of the tile; lanes 0-15 hold the columns
// [0, 64) of their row -- the whole box part and the objectness (2 + 2 LEN + 1 <= 64) -- and lanes 16-31 the columns
// [64, 128), classes only.  One 64-register load, one wait, the TMEM buffer goes back, then 64 independent sigmoids per
// thread; the bin arg-max (first maximum of the sigmoids, as torch.max) runs in registers of the lower half-warp.
// The whole-row form (three warps per quadrant sharing the columns, 16-column chunks, branches on the column's role with
// a run-time LEN) took 186-226 cycles per box column and warp (YC_TC_DEBUG bit 8).
// TAIL1: no >= 127, so only the very last column of the upper half (accumulator column 127) can be missing: the stores of all
// other columns need no dummy-word select
template <int LEN, bool PAIR, bool TAIL1>
__device__ __forceinline__ void store_rows_half_ibin(const TcParams &P, const TcLevel &L, int b, int prow0, int nv, int ar,
                                                     uint32_t taddr, uint32_t tab_s, uint32_t slab_s, uint32_t dummy_s, uint64_t *tempty, int lane)
{
    constexpr int OFF = 64, C_OBJ = 2 + 2 * LEN, SHIFT = 2 * LEN - 2;   // z column = accumulator column - SHIFT from C_OBJ on
    static_assert(C_OBJ < OFF, "box part and objectness must sit in the lower half");
    const int no = P.no, no_out = P.no_out;
    const int half = lane >> 4, r = lane & 15, cbase = half * OFF;
    const int ncols = half == 0 ? OFF : max(0, min(OFF, no - OFF));
    const uint32_t trow = tab_s + (uint32_t)cbase * 8u;
    float *const dummy = shared_f32(dummy_s);
    uint32_t v[OFF];
    TmemLdHalf<64, OFF>::ld(taddr, v);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (PAIR) mbar_arrive_leader(tempty);
        else mbar_arrive(tempty);
    }
    const int rows = min(16, nv);
    if (L.raw) {
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        float *const srow = shared_f32(slab_s) + r * no + cbase;
#pragma unroll
        for (int j0 = 0; j0 < OFF; j0 += HG_IBIN) {
            float w[HG_IBIN];
#pragma unroll
            for (int j = j0; j < j0 + HG_IBIN; ++j) {
                const float2 s_b = lds_f32x2(trow + 8u * j);
                w[j - j0] = fmaf(__uint_as_float(v[j]), s_b.x, s_b.y);
            }
#pragma unroll
            for (int j = j0; j < j0 + HG_IBIN; ++j)
                *(((TAIL1 && j < OFF - 1) || j < ncols) ? srow + j : dummy) = w[j - j0];
        }
        half_slab_store(L.raw + (((size_t)b * P.na_real + ar) * L.HW + prow0) * no, slab_s, rows, no, lane);
    }
    if (P.write_z) {
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        const int p = prow0 + r;
        const float gx = (float)(p % L.nx), gy = (float)(p / L.nx);
        float *const srow = shared_f32(slab_s) + r * no_out;   // this thread's z row
        float *const scol = srow + (cbase - SHIFT);            // [j]: z column of accumulator column cbase + j
        float reg_w = 0.f, reg_h = 0.f, best_w = -1.f, best_h = -1.f;
        int idx_w = 0, idx_h = 0;
#pragma unroll
        for (int j0 = 0; j0 < OFF; j0 += HG_IBIN) {
            float w[HG_IBIN];
#pragma unroll
            for (int j = j0; j < j0 + HG_IBIN; ++j) {
                const float2 s_b = lds_f32x2(trow + 8u * j);
                w[j - j0] = sigmoidf_rcp(fmaf(__uint_as_float(v[j]), s_b.x, s_b.y));
            }
#pragma unroll
            for (int j = j0; j < j0 + HG_IBIN; ++j) {
                const float sg = w[j - j0];
                if (j < C_OBJ) {
                    // lower half: box part (kept in registers); upper half: a class column
                    if (j == 2) reg_w = sg;
                    else if (j > 2 && j < 2 + LEN) { if (sg > best_w) { best_w = sg; idx_w = j - 3; } }
                    else if (j == 2 + LEN) reg_h = sg;
                    else if (j > 2 + LEN) { if (sg > best_h) { best_h = sg; idx_h = j - 3 - LEN; } }
                    if (j < 2) {   // lower half: decoded x / y at z column j; upper half: a class
                        const float d = j == 0 ? decode_xy(sg, gx, L.stride) : decode_xy(sg, gy, L.stride_y);
                        *(half == 0 ? srow + j : (j < ncols ? scol + j : dummy)) = half == 0 ? d : sg;
                    } else {
                        *(half != 0 && ((TAIL1 && j < OFF - 1) || j < ncols) ? scol + j : dummy) = sg;
                    }
                } else {
                    *(((TAIL1 && j < OFF - 1) || j < ncols) ? scol + j : dummy) = sg;
                }
            }
        }
        if (half == 0) {   // (reg * 2 - 1) * step + bins[argmax], clamped to [0, 4], times the anchor
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                float t = __fmul_rn(d == 0 ? reg_w : reg_h, 2.0f);
                t = __fadd_rn(t, -1.0f);
                t = __fmul_rn(t, P.bin_step);
                float res = __fadd_rn(t, __ldg(P.bins + (d == 0 ? idx_w : idx_h)));
                res = fminf(fmaxf(res, 0.0f), 4.0f);
                srow[2 + d] = __fmul_rn(res, L.anchor_wh[2 * ar + d]);
            }
        }
        half_slab_store(P.z + ((size_t)b * P.rows_total + L.row_off + (size_t)ar * L.HW + prow0) * no_out, slab_s, rows, no_out, lane);
    }
}

// Fused IBin epilogue by half rows (nets/ibin.py:56-74 + detect.py:108-121 in one pass).  Eight epilogue warps: warp
// (quadrant, pass) owns 16 rows.  Every warp first reads the objectness column of its quadrant (one 1-column load) and
// rejects on it; a warp that holds a survivor then fetches its 16 rows with ONE 64-register half-row load and hands the
// TMEM buffer back BEFORE it decodes anything -- the whole-row form (fused_epilogue_ibin) scans 80 class columns and
// decodes 46 box columns out of TMEM while it holds the buffer, and with ~0.5 % of the rows above the objectness
// threshold nearly every second tile has such a warp (measured: 392 us per 16 images at 1280x1280 against 160 us with
// the epilogue switched off).  Lanes 0-15 own the box part, the objectness and the first classes of their row, lanes 16-31
// the remaining classes of the same row; the class maximum is combined with one shuffle (first maximum, as torch.max).
// survivors of the objectness test among this warp's 16 rows, from the objectness accumulator of the quadrant's 32 rows
// (lane = row, already loaded): class scores are sigmoids (<= 1), so obj >= conf is necessary for obj * cls >= conf
template <int LEN>
__device__ __forceinline__ unsigned ibin_obj_survivors(uint32_t o_raw, uint32_t tab_s, int pass16, int nv, float conf, int lane)
{
    const float2 so = lds_f32x2(tab_s + 8u * (2 + 2 * LEN));
    const float obj32 = sigmoidf_rcp(fmaf(__uint_as_float(o_raw), so.x, so.y));
    const bool mine = (lane >> 4) == pass16 && (lane & 15) < nv;
    return __ballot_sync(0xffffffffu, mine && obj32 >= conf);
}

template <int LEN, bool PAIR>
__device__ __forceinline__ void fused_epilogue_ibin_half_tail(const TcParams &P, const TcLevel &L, int b, int prow0, int nv, int ar,
                                                              uint32_t tq, int pass16, uint32_t tab_s, uint32_t qrow_s, uint64_t *tempty,
                                                              int lane, unsigned surv);

template <int LEN, bool PAIR>
__device__ __forceinline__ void fused_epilogue_ibin_half(const TcParams &P, const TcLevel &L, int b, int prow0, int nv, int ar,
                                                         uint32_t tq, int pass16, uint32_t tab_s, uint32_t qrow_s, uint64_t *tempty,
                                                         int lane)
{
    // objectness of the quadrant's 32 rows (lane = row); this warp looks at its own 16
    uint32_t o1[1];
    TmemLd<1>::ld(tq + (uint32_t)(2 + 2 * LEN), o1);
    tmem_ld_wait();
    const unsigned surv = ibin_obj_survivors<LEN>(o1[0], tab_s, pass16, nv, P.conf, lane);
    fused_epilogue_ibin_half_tail<LEN, PAIR>(P, L, b, prow0, nv, ar, tq, pass16, tab_s, qrow_s, tempty, lane, surv);
}

// the rest of fused_epilogue_ibin_half once the survivors are known (the CTA-pair kernel probes the objectness of all
// anchors of a tile behind one wait and then comes here anchor by anchor).
// A survivor's row is decoded by the WHOLE warp: the two lanes that hold it park its 127 raw accumulators in the warp's
// row buffer, then the 32 lanes spread over the columns -- three classes, one w bin and one h bin each -- and the three
// arg-maxes are warp reductions (first maximum: larger value wins, ties go to the smaller index, as torch.max).  ~300
// instructions per survivor row; the per-thread form (every lane runs the 64 sigmoids of its half row and the bin scan
// for all 16 rows of the warp) took ~1800 at a third of an instruction per cycle, longer than a tile's MMAs, so that a
// warp with a survivor came late to the next tile: 224 -> 133 us per 16 images at 1280x1280 without that path.
template <int LEN, bool PAIR>
__device__ __forceinline__ void fused_epilogue_ibin_half_tail(const TcParams &P, const TcLevel &L, int b, int prow0, int nv, int ar,
                                                              uint32_t tq, int pass16, uint32_t tab_s, uint32_t qrow_s, uint64_t *tempty,
                                                              int lane, unsigned surv)
{
    constexpr int OFF = 64, C_OBJ = 2 + 2 * LEN, C_CLS = C_OBJ + 1;
    const int no = P.no, nc = no - C_CLS;
    const int half = lane >> 4, cbase = half * OFF;
    uint32_t v[OFF];
    if (surv) {
        TmemLdHalf<64, OFF>::ld(tq + ((uint32_t)(16 * pass16) << 16), v);
        tmem_ld_wait();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (PAIR) mbar_arrive_leader(tempty);
        else mbar_arrive(tempty);
    }
    if (!surv) return;
    float *const q = shared_f32(qrow_s);   // [128] raw accumulators of the row being decoded
    // (scale, bias) applied, sigmoid: column c of the parked row
    auto sig = [&](int c) {
        const float2 s_b = lds_f32x2(tab_s + 8u * c);
        return sigmoidf_rcp(fmaf(q[c], s_b.x, s_b.y));
    };
    unsigned rows = (surv >> (16 * pass16)) & 0xFFFFu;   // survivors among this warp's 16 rows
    while (rows) {
        const int rr = __ffs(rows) - 1;
        rows &= rows - 1;
        if ((lane & 15) == rr) {   // lanes rr (columns 0..63) and rr + 16 (columns 64..127)
#pragma unroll
            for (int j = 0; j < OFF; ++j) q[cbase + j] = __uint_as_float(v[j]);
        }
        __syncwarp();
        // class maximum: lane l scans the classes l, l + 32, l + 64 in ascending order
        float bv = -1.0f;
        int best = 0;
#pragma unroll
        for (int k3 = 0; k3 < 3; ++k3) {
            const int k = lane + 32 * k3;
            if (k < nc) {
                const float sg = sig(C_CLS + k);
                if (sg > bv) { bv = sg; best = k; }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, best, off);
            if (ov > bv || (ov == bv && oi < best)) { bv = ov; best = oi; }
        }
        const float obj = sig(C_OBJ);
        const float score = __fmul_rn(obj, bv);
        if (obj >= P.conf && score >= P.conf) {   // warp-uniform
            // bins: lane l < LEN - 1 holds bin l of the w block and of the h block
            float bw = -1.0f, bh = -1.0f;
            int iw = lane, ih = lane;
            if (lane < LEN - 1) {
                bw = sig(3 + lane);
                bh = sig(3 + LEN + lane);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ow = __shfl_xor_sync(0xffffffffu, bw, off), oh = __shfl_xor_sync(0xffffffffu, bh, off);
                const int jw = __shfl_xor_sync(0xffffffffu, iw, off), jh = __shfl_xor_sync(0xffffffffu, ih, off);
                if (ow > bw || (ow == bw && jw < iw)) { bw = ow; iw = jw; }
                if (oh > bh || (oh == bh && jh < ih)) { bh = oh; ih = jh; }
            }
            if (lane == 0) {
                const int p = prow0 + rr;
                const float cx = decode_xy(sig(0), (float)(p % L.nx), L.stride);
                const float cy = decode_xy(sig(1), (float)(p / L.nx), L.stride_y);
                float wh[2];
#pragma unroll
                for (int d = 0; d < 2; ++d) {   // (reg * 2 - 1) * step + bins[argmax], clamped to [0, 4], times the anchor
                    float t = __fmul_rn(sig(d == 0 ? 2 : 2 + LEN), 2.0f);
                    t = __fadd_rn(t, -1.0f);
                    t = __fmul_rn(t, P.bin_step);
                    float res = __fadd_rn(t, __ldg(P.bins + (d == 0 ? iw : ih)));
                    res = fminf(fmaxf(res, 0.0f), 4.0f);
                    wh[d] = __fmul_rn(res, L.anchor_wh[2 * ar + d]);
                }
                float x1, y1, x2, y2;
                xywh_to_corners(cx, cy, wh[0], wh[1], P.div_w, P.div_h, x1, y1, x2, y2);
                emit_one(b, L.row_off + ar * L.HW + p, P.rows_total, P.nc, x1, y1, x2, y2, obj, bv, score, best, P.ws);
            }
        }
        __syncwarp();   // the row buffer is free again
    }
}

// the OFF values instantiated (tcgen05.ld takes the half-split offset as an immediate)
__host__ __device__ inline int half_off_for(int no, int na_tile)
{
    const int need = (no + 1) / 2;
    const int cand[6] = {4, 8, 16, 32, 43, 64};
    for (int i = 0; i < 6; ++i)   // lanes 16-31 read the columns [OFF, 2 OFF) of the LAST anchor too: they must stay inside the buffer
        if (cand[i] >= need && cand[i] <= no && (na_tile - 1) * no + 2 * cand[i] <= TC_MAX_N) return cand[i];
    return 0;
}

template <bool PAIR>
__device__ __forceinline__ void store_rows_half_any(const TcParams &P, const TcLevel &L, int b, int prow0, int nv, int ar, uint32_t taddr,
                                                    uint32_t tab_s, uint32_t slab_s, uint32_t dummy_s, uint64_t *tempty, int lane, long long *pf = nullptr)
{
    switch (P.half_off) {
    case 4: store_rows_half<4, PAIR>(P, L, b, prow0, nv, ar, taddr, tab_s, slab_s, dummy_s, tempty, lane, pf); break;
    case 8: store_rows_half<8, PAIR>(P, L, b, prow0, nv, ar, taddr, tab_s, slab_s, dummy_s, tempty, lane, pf); break;
    case 16: store_rows_half<16, PAIR>(P, L, b, prow0, nv, ar, taddr, tab_s, slab_s, dummy_s, tempty, lane, pf); break;
    case 32: store_rows_half<32, PAIR>(P, L, b, prow0, nv, ar, taddr, tab_s, slab_s, dummy_s, tempty, lane, pf); break;
    case 43: store_rows_half<43, PAIR>(P, L, b, prow0, nv, ar, taddr, tab_s, slab_s, dummy_s, tempty, lane, pf); break;
    default: store_rows_half<64, PAIR>(P, L, b, prow0, nv, ar, taddr, tab_s, slab_s, dummy_s, tempty, lane, pf); break;
    }
}


// Fused epilogue of one warp for one tile (see yc_detect_fused): reads this warp's 32 rows of the accumulator
// (anchor `ar`), rejects on objectness, scans the classes of the survivors, emits NMS candidates and hands the
// TMEM buffer back through `tempty` (PAIR: the barrier lives in the pair's leader CTA).
// The (scale, bias) pairs of the 4 box columns and the objectness column are loaded by the caller BEFORE it waits
// for the accumulator (`sbv`), so that nothing but the TMEM load sits between "tile finished" and the early reject.
// `cls` holds the pairs of the classes lane, lane + 32, lane + 64 (the classes this lane scans for a queued survivor).
// With 227 KB of shared memory carved out there is next to no L1 left, so every __ldg here is an L2 round trip: the
// caller keeps the struct in registers and reloads it only when the level changes.
struct BoxSb { float2 v[5]; float2 cls[3]; };
__device__ __forceinline__ BoxSb load_box_sb(const float2 *__restrict__ sb, int lane, int nc)
{
    BoxSb r;
#pragma unroll
    for (int j = 0; j < 5; ++j) r.v[j] = YC_SB(sb + j);
#pragma unroll
    for (int k = 0; k < 3; ++k) r.cls[k] = lane + 32 * k < nc ? YC_SB(sb + 5 + lane + 32 * k) : make_float2(0.f, 0.f);
    return r;
}

template <bool PAIR>
__device__ __forceinline__ void fused_epilogue(const TcParams &P, const TcLevel &L, int b, int prow0, int nv, int ar,
                                               uint32_t taddr, float *slab, uint64_t *tempty, int lane, const BoxSb &sbv)
{
    const int no = P.no;
    const int p = prow0 + lane;
    const float aw = L.anchor_wh[2 * ar], ah = L.anchor_wh[2 * ar + 1];
    const float2 *sb = L.sb + ar * no;
    // box + objectness logits (columns 0..4), then the largest class logit
    // one 8-column load (3 columns more than needed) instead of a 4- and a 1-column load: next to running MMAs a
    // tcgen05.ld costs ~250 cycles per instruction, almost independent of its width
    uint32_t v8[8];
    TmemLd<8>::ld(taddr, v8);
    tmem_ld_wait();
    float tb[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) tb[j] = fmaf(__uint_as_float(v8[j]), sbv.v[j].x, sbv.v[j].y);
    // Early reject on objectness alone: class scores are sigmoids (<= 1), so obj >= conf is necessary
    // for obj*cls >= conf.  At detection thresholds >99% of rows stop here after 5 columns; only
    // warps holding a survivor scan the class columns (exactly as the z path would see them).
    const float obj = sigmoidf_fast(tb[4]);
    bool pass = lane < nv && obj >= P.conf;
    const unsigned surv = __ballot_sync(0xffffffffu, pass);
    const int n_surv = __popc(surv);
    if (n_surv > 0 && n_surv <= TC_QUEUE_ROWS) {
        // Few survivors (the common case): reserve their candidate slots with ONE atomicAdd (its round trip hides
        // behind what follows), copy their class accumulators to shared memory, hand the TMEM buffer back at once,
        // then scan the classes with the 32 lanes spread over the classes.
        int base = 0;
        if (lane == 0) base = atomicAdd(&P.ws.cand_count[b], n_surv);
        float *q = shared_f32(smem_addr(slab)); // per-warp queue [TC_QUEUE_ROWS][nc] (a pointer the compiler knows to be shared)
        const int nc = P.nc;
        float *qrow = q + __popc(surv & ((1u << lane) - 1u)) * nc;
        switch ((nc & 15) == 0 ? nc >> 4 : 0) {
        case 1: grab_classes<1>(taddr, pass, qrow, tempty, lane, PAIR); break;
        case 2: grab_classes<2>(taddr, pass, qrow, tempty, lane, PAIR); break;
        case 3: grab_classes<3>(taddr, pass, qrow, tempty, lane, PAIR); break;
        case 4: grab_classes<4>(taddr, pass, qrow, tempty, lane, PAIR); break;
        case 5: grab_classes<5>(taddr, pass, qrow, tempty, lane, PAIR); break;
        case 6: grab_classes<6>(taddr, pass, qrow, tempty, lane, PAIR); break;
        default:
            queue_classes(taddr, no, pass, qrow);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_leader(tempty);
                else mbar_arrive(tempty);
            }
        }
        unsigned left = surv;
        float my_bv = 0.0f;
        int my_best = 0, my_src = 0;
        for (int sidx = 0; sidx < n_surv; ++sidx) {
            const int src = __ffs(left) - 1;
            left &= left - 1;
            float bv = -1.0f;
            int best = 0;
            if (nc <= 96) { // ascending classes per lane: strict > keeps the first
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int c = lane + 32 * k;
                    if (c < nc) {
                        const float sg = sigmoidf_fast(fmaf(q[sidx * nc + c], sbv.cls[k].x, sbv.cls[k].y));
                        if (sg > bv) { bv = sg; best = c; }
                    }
                }
            } else {
                for (int c = lane; c < nc; c += 32) {
                    const float2 s_b = __ldg(sb + 5 + c);
                    const float sg = sigmoidf_fast(fmaf(q[sidx * nc + c], s_b.x, s_b.y));
                    if (sg > bv) { bv = sg; best = c; }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) { // larger value wins, ties go to the smaller class
                const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, best, off);
                if (ov > bv || (ov == bv && oi < best)) { bv = ov; best = oi; }
            }
            if (lane == sidx) { my_bv = bv; my_best = best; my_src = src; }
        }
        // lane i finishes survivor i: all lanes take part in the shuffles, lanes >= n_surv read lane 0
        const float o_s = __shfl_sync(0xffffffffu, obj, my_src);
        const float t0 = __shfl_sync(0xffffffffu, tb[0], my_src), t1 = __shfl_sync(0xffffffffu, tb[1], my_src);
        const float t2 = __shfl_sync(0xffffffffu, tb[2], my_src), t3 = __shfl_sync(0xffffffffu, tb[3], my_src);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < n_surv) {
            const float score = __fmul_rn(o_s, my_bv);
            if (score >= P.conf) {
                const int ps = prow0 + my_src;
                const float cx = decode_xy(sigmoidf_fast(t0), (float)(ps % L.nx), L.stride);
                const float cy = decode_xy(sigmoidf_fast(t1), (float)(ps / L.nx), L.stride_y);
                const float bw = decode_wh(sigmoidf_fast(t2), aw), bh = decode_wh(sigmoidf_fast(t3), ah);
                float x1, y1, x2, y2;
                xywh_to_corners(cx, cy, bw, bh, P.div_w, P.div_h, x1, y1, x2, y2);
                emit_at(base + lane, b, L.row_off + ar * L.HW + ps, P.rows_total, nc, x1, y1, x2, y2, o_s, my_bv, score,
                        my_best, P.ws);
            } else {
                emit_hole(base + lane, b, P.rows_total, P.ws);
            }
        }
        __syncwarp();
        return;
    }
    if (n_surv > 0) {
        // many survivors (low thresholds): class scan in registers, first maximum of the sigmoids
        float bv = -1.0f;
        int best = 0;
        cls_scan<true>(taddr, no, sb, bv, best);
        const float score = __fmul_rn(obj, bv);
        pass = pass && score >= P.conf;
        const float gx = (float)(p % L.nx), gy = (float)(p / L.nx);
        const float cx = decode_xy(sigmoidf_fast(tb[0]), gx, L.stride);
        const float cy = decode_xy(sigmoidf_fast(tb[1]), gy, L.stride_y);
        const float bw = decode_wh(sigmoidf_fast(tb[2]), aw), bh = decode_wh(sigmoidf_fast(tb[3]), ah);
        float x1, y1, x2, y2;
        xywh_to_corners(cx, cy, bw, bh, P.div_w, P.div_h, x1, y1, x2, y2);
        emit_candidates(pass, b, L.row_off + ar * L.HW + p, P.rows_total, P.nc, x1, y1, x2, y2, obj, bv,
                        score, best, P.ws);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
            if (PAIR) mbar_arrive_leader(tempty);
            else mbar_arrive(tempty);
        }
}

// Fused epilogue for the IBin head (nets/ibin.py:56-74 + detect.py:108-121 in one pass; one anchor per tile, one warp per
// TMEM lane quadrant): reject on the objectness column alone (it sits behind the two bin blocks); only a warp that holds a
// survivor reads the rest of its rows -- class scores (first maximum of the sigmoids, as the z path sees them), then the
// 2 + 2 * (bin_count + 1) box columns through the same decode as the z-writing epilogue (bin arg-max on the sigmoids)
// -- and emits NMS candidates.  The 127-sigmoid row decode therefore runs for the few warps with a survivor instead of for
// every row, and z is never written.  `scratch`: 4 floats per lane of this warp's shared-memory area.
__device__ __forceinline__ void fused_epilogue_ibin(const TcParams &P, const TcLevel &L, int b, int prow0, int nv, int ar,
                                                    uint32_t taddr, float *scratch, uint64_t *tempty, int lane)
{
    const int no = P.no, len = P.bin_count + 1, c_obj = 2 + 2 * len, c_cls = c_obj + 1;
    const int p = prow0 + lane;
    const float2 *sb = L.sb + ar * no;
    uint32_t v1[1];
    TmemLd<1>::ld(taddr + (uint32_t)c_obj, v1);
    tmem_ld_wait();
    const float2 so = __ldg(sb + c_obj);
    const float obj = sigmoidf_fast(fmaf(__uint_as_float(v1[0]), so.x, so.y));
    bool pass = lane < nv && obj >= P.conf;    // class scores are sigmoids (<= 1): obj >= conf is necessary
    if (__ballot_sync(0xffffffffu, pass)) {
        float bv = -1.0f;
        int best = 0;
        cls_scan<true>(taddr, no, sb, bv, best, c_cls);
        const float score = __fmul_rn(obj, bv);
        pass = pass && score >= P.conf;
        float *srow = scratch + lane * 4;
        epi_range_ibin<false>(taddr, 0, c_obj, true, true, sb, srow, (float)(p % L.nx), (float)(p / L.nx), L.stride, L.stride_y,
                              L.anchor_wh[2 * ar], L.anchor_wh[2 * ar + 1], P);
        float x1, y1, x2, y2;
        xywh_to_corners(srow[0], srow[1], srow[2], srow[3], P.div_w, P.div_h, x1, y1, x2, y2);
        emit_candidates(pass, b, L.row_off + ar * L.HW + p, P.rows_total, P.nc, x1, y1, x2, y2, obj, bv, score, best, P.ws);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty);
}


} // namespace yc
