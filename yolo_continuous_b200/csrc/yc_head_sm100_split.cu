// yc_head_sm100_split.cu -- float32 feature maps on the tensor cores at float32 grade (reference nets/idetect.py:31 is a
// float32 convolution; north_star asks for 1e-5 parity in that mode).
//
// Scheme.  Every operand is split into two fp16 numbers, v = hi + lo (22 significant bits), and the product is
//     x * w  ~=  x_hi*w_hi  +  (x_hi*w_lo + x_lo*w_hi)                      (the dropped x_lo*w_lo term is 2^-22 relative)
// evaluated as three kind::f16 tcgen05 MMAs per k-step with fp32 accumulation in TMEM.  The large term and the two small
// terms go to SEPARATE accumulators: the tensor core's fp32 accumulation is not IEEE round-to-nearest (measured drift
// 1.5e-5 on a K = 1024 logit when everything shares one accumulator), and its error scales with the
// magnitude of the running sum and the number of accumulation steps -- the correction accumulator stays 2^-11 of the
// main one, so only K/16 steps (not 3K/16) see the full magnitude.  The epilogue adds the two in IEEE binary32.
//   weights      split once by yc_head_pack (row c scaled by 2^shift(c) so that hi/lo stay normal fp16 numbers; stored
//                transposed [K][Npad] so that B is an MN-major operand like A); the epilogue scale is im * 2^-shift * 2^XSHIFT
//   activations  split IN THE KERNEL: TMA lands the float32 tile of NCHW pixels, four converter warps read it, scale by
//                2^-XSHIFT (range +-2^20; below |x| = 2 the low part is a subnormal fp16 with absolute error <= 2^-21, which
//                is what bounds the result: ~2e-7 on a logit for N(0,.02) weights at K = 1024) and write x_hi / x_lo over the
//                landing buffer in the MN-major SWIZZLE_128B layout the MMA reads.  HBM sees the float32 maps once:
//                11.47 MB/img, nothing else (S1 fp32: 20.04 MB/img with z).
// Tile = 128 pixels x ONE anchor (npad = round_up(no, 16) <= 128 columns): main + correction = 256 TMEM columns, double
// buffered; the three anchors of a pixel block are consecutive tiles, which run on neighbouring SMs at the same time, so
// the block's feature maps come from HBM once and from L2 twice.
// Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, then one or two groups of four converter warps, then 4
// epilogue warps (12 for IBin).
#include "yc_head_tc.cuh"

namespace yc {

constexpr int TS_BK = 32;                       // k per stage
constexpr int TS_A_BYTES = TC_BM * TS_BK * 4;   // 16 KB: float32 landing buffer == x_hi (8 KB) | x_lo (8 KB) after conversion
constexpr int TS_B_HALF = 128 * TS_BK * 2;      // 8 KB: w_hi (or w_lo) rows of up to 128 columns, two {64 n, 32 k} boxes
constexpr int TS_STAGE_BYTES = TS_A_BYTES + 2 * TS_B_HALF;   // 32 KB
constexpr int TS_MAX_STAGES = 6;
constexpr int TS_CONV_BAR_ID = 9;               // named barriers 9, 10 of the converter groups (1..8 belong to the IBin epilogue)

__device__ __forceinline__ uint32_t pack_half2(float a, float b)
{
    const __half2 h = __floats2half2_rn(a, b);
    return *(const uint32_t *)&h;
}

// NCG converter groups of four warps each take alternate k-blocks: one group converts a stage in ~600 cycles (8 LDS.128,
// ~160 ALU instructions and 8 STS.128 per thread behind a 128-thread barrier), more than the stage's six MMAs take
template <int NCG>
__global__ void __maxnreg__(96)
head_tcs_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams P)
{
    constexpr int TS_NON_EPI_WARPS = 4 + 4 * NCG;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *stage_base = smem;
    const int n_stages = P.stages;
    uint8_t *slabs = smem + n_stages * TS_STAGE_BYTES;
    const int n_epi_warps = P.epi_warps;
    uint64_t *bars = (uint64_t *)(slabs + (size_t)4 * P.slab_bytes);
    uint64_t *full_bar = bars;                          // TMA landed (A float32 + both weight halves)
    uint64_t *conv_bar = bars + TS_MAX_STAGES;          // converters wrote x_hi / x_lo
    uint64_t *empty_bar = bars + 2 * TS_MAX_STAGES;     // the stage's MMAs retired
    uint64_t *tfull_bar = bars + 3 * TS_MAX_STAGES;     // [2]
    uint64_t *tempty_bar = tfull_bar + 2;               // [2]
    uint32_t *tmem_ptr_smem = (uint32_t *)(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_groups = P.lv[0].n_groups;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < P.n_lv; ++i) {
            prefetch_tmap(&maps.a[i]);
            prefetch_tmap(&maps.b[2 * i]);
            prefetch_tmap(&maps.b[2 * i + 1]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < n_stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&conv_bar[i], 4);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], (uint32_t)n_epi_warps);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr_smem, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // tile id -> (level, image, first pixel, anchor group): groups innermost (see the header comment)
    auto coord = [&](int t) {
        int l = 0;
#pragma unroll
        for (int i = 1; i < YC_MAX_LEVELS; ++i)
            if (i < P.n_lv && t >= P.lv[i].tile_begin) l = i;
        int r = t - P.lv[l].tile_begin;
        TileCoord c;
        c.lv = l;
        c.g = r % n_groups;
        r /= n_groups;
        c.b = r / P.lv[l].tiles_per_img;
        c.p0 = (r - c.b * P.lv[l].tiles_per_img) * TC_BM;
        return c;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = (uint32_t)TS_A_BYTES + 4u * P.b_box_bytes;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
            const TileCoord tc = coord(t);
            const int nkb = (P.lv[tc.lv].K + TS_BK - 1) / TS_BK;
            const CUtensorMap *ma = &maps.a[tc.lv], *mh = &maps.b[2 * tc.lv], *ml = &maps.b[2 * tc.lv + 1];
            const int n0 = tc.g * P.npad_g;   // first weight column of this anchor in the padded transposed copies
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                if (elect_one()) {
                    uint8_t *sa = stage_base + stage * TS_STAGE_BYTES, *sb = sa + TS_A_BYTES;
                    const bool da = !(P.debug & 16), db = !(P.debug & 32);
                    if (da || db) mbar_arrive_expect_tx(&full_bar[stage], (da ? (uint32_t)TS_A_BYTES : 0u) + (db ? 4u * P.b_box_bytes : 0u));
                    else mbar_arrive(&full_bar[stage]);
                    if (da) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)   // four {32 px, 32 k} float32 boxes
                        tma_load_3d(sa + j * 4096, ma, &full_bar[stage], tc.p0 + 32 * j, kb * TS_BK, tc.b);
                    }
                    if (db) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) { // {64 n, 32 k} boxes of w_hi and w_lo
                        tma_load_2d(sb + j * 4096, mh, &full_bar[stage], n0 + 64 * j, kb * TS_BK);
                        tma_load_2d(sb + TS_B_HALF + j * 4096, ml, &full_bar[stage], n0 + 64 * j, kb * TS_BK);
                    }
                    }
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0, it = 0;
        uint32_t phase = 0;
        // A and B are both MN-major SWIZZLE_128B: 64-element chunks LBO = 32 k-rows * 128 B apart, 8-k groups SBO = 1024 B
        // apart; one k16 step = two 8-k groups = 2048 B
        const uint32_t s0 = smem_addr(stage_base);
        const uint64_t d_xhi = smem_desc(s0, 4096, 1024, SWZ_128B);
        const uint64_t d_xlo = smem_desc(s0 + 8192, 4096, 1024, SWZ_128B);
        const uint64_t d_whi = smem_desc(s0 + TS_A_BYTES, 4096, 1024, SWZ_128B);
        const uint64_t d_wlo = smem_desc(s0 + TS_A_BYTES + TS_B_HALF, 4096, 1024, SWZ_128B);
        const uint32_t idesc = P.idesc;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
            const TileCoord tc = coord(t);
            const int nkb = (P.lv[tc.lv].K + TS_BK - 1) / TS_BK;
            const int buf = it & 1;
            mbar_wait(&tempty_bar[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t t_main = tmem_base + (uint32_t)buf * TC_MAX_N, t_corr = t_main + TC_SPLIT_CORR;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[stage], phase);   // the weight halves (async proxy writes) are visible to this thread
                mbar_wait(&conv_bar[stage], phase);   // x_hi / x_lo are in place
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t so = (uint64_t)((uint32_t)(stage * TS_STAGE_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < TS_BK / 16; ++k) {
                        if (P.debug & 2) break;
                        const uint64_t ko = so + (uint64_t)((k * 2048) >> 4);
                        const uint32_t acc = (uint32_t)((kb | k) != 0);
                        mma_f16(t_corr, d_xlo + ko, d_whi + ko, idesc, acc);
                        mma_f16(t_corr, d_xhi + ko, d_wlo + ko, idesc, 1u);
                        mma_f16(t_main, d_xhi + ko, d_whi + ko, idesc, acc);
                    }
                    mma_commit(&empty_bar[stage]);
                    if (kb == nkb - 1) mma_commit(&tfull_bar[buf]);
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4 && warp < TS_NON_EPI_WARPS) {
        // ===================== converters: float32 landing tile -> x_hi | x_lo (fp16, in place) =====================
        // thread -> k-row (lane) of the 32-pixel box `pq` (warp): 8 x LDS.128 along the swizzled 128-byte row, then (after
        // all four warps have read everything: the outputs overwrite other threads' inputs) 4 + 4 x STS.128 into rows of
        // the two fp16 operands, same swizzle (16-byte chunk index ^ (row & 7)) as TMA would have written
        const int cg = (warp - 4) >> 2, pq = (warp - 4) & 3, k = lane;   // converter group, pixel quarter, k-row
        const uint32_t sw = (uint32_t)(k & 7);
        const float down = 1.0f / (float)(1 << YC_SPLIT_XSHIFT);
        int stage = 0, gk = 0;   // gk: k-blocks seen so far (group cg converts those with gk % NCG == cg)
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
            const TileCoord tc = coord(t);
            const int nkb = (P.lv[tc.lv].K + TS_BK - 1) / TS_BK;
            for (int kb = 0; kb < nkb; ++kb, ++gk) {
                if (NCG > 1 && gk % NCG != cg) {
                    if (++stage == n_stages) { stage = 0; phase ^= 1u; }
                    continue;
                }
                uint8_t *sa = stage_base + stage * TS_STAGE_BYTES;
                mbar_wait(&full_bar[stage], phase);
                if (P.debug & 64) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&conv_bar[stage]);
                    if (++stage == n_stages) { stage = 0; phase ^= 1u; }
                    continue;
                }
                float4 v[8];
                const uint8_t *src = sa + pq * 4096 + k * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = *(const float4 *)(src + (((uint32_t)c ^ sw) << 4));
                named_bar_sync(TS_CONV_BAR_ID + cg, 128);
                uint8_t *dhi = sa + (pq >> 1) * 4096 + k * 128, *dlo = dhi + 8192;
#pragma unroll
                for (int j = 0; j < 4; ++j) {   // 8 pixels per 16-byte chunk
                    const float x[8] = {v[2 * j].x * down, v[2 * j].y * down, v[2 * j].z * down, v[2 * j].w * down,
                                        v[2 * j + 1].x * down, v[2 * j + 1].y * down, v[2 * j + 1].z * down, v[2 * j + 1].w * down};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __half2 h = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
                        const float2 hf = __half22float2(h);
                        hi[e] = *(const uint32_t *)&h;
                        lo[e] = pack_half2(x[2 * e] - hf.x, x[2 * e + 1] - hf.y);   // exact differences, rounded once
                    }
                    const uint32_t off = (((uint32_t)((pq & 1) * 4 + j)) ^ sw) << 4;
                    *(uint4 *)(dhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *(uint4 *)(dlo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&conv_bar[stage]);
                if (++stage == n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= TS_NON_EPI_WARPS) {
        // ===================== epilogue (shared with the bf16 kernel; SPLIT adds the two accumulators) ==============
        const int e = warp - TS_NON_EPI_WARPS, q = warp & 3;
        int it = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
            const TileCoord tc = coord(t);
            const int buf = it & 1;
            mbar_wait(&tfull_bar[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            if (P.debug & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[buf]);
                continue;
            }
            store_epilogue<true>(P, P.lv[tc.lv], tc.b, tc.p0, tc.g, e, q, lane, tmem_base + (uint32_t)(buf * TC_MAX_N), slabs,
                                 &tempty_bar[buf]);
        }
        if (lane == 0) bulk_wait_all0();
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// host side; called by launch_head_tcgen05 (yc_head_sm100.cu) for float32 feature maps
int launch_head_split(const yc_head_desc *d, int rows_total, const int *row_off, unsigned *left_mask, void *enc_fn, int num_sms,
                      cudaStream_t stream)
{
    EncodeTiledFn enc = (EncodeTiledFn)enc_fn;
    const int N = d->na * d->no;
    const bool ibin = d->kind == YC_HEAD_IBIN;
    int dbg = 0;
    { const char *e = getenv("YC_TS_DEBUG"); dbg = e ? atoi(e) : 0; }
    const int npad_g = round_up(d->no, 16), wt = d->na * npad_g;   // layout of w_hi_t / w_lo_t (yc_head_pack)
    const int npad = (dbg & 8) ? 128 : npad_g;
    YC_REQUIRE(npad <= TC_SPLIT_CORR, YC_ERR_UNSUPPORTED, "tcgen05 fp32 head: %d outputs per anchor do not fit 128 accumulator columns",
               d->no);
    const int no_out = ibin ? d->no - 2 * (d->bin_count + 1) + 2 : d->no;
    bool any_raw = false;
    for (int i = 0; i < d->nl; ++i) any_raw = any_raw || d->level[i].raw != nullptr;
    // one slab per TMEM lane quadrant (one anchor per tile): z rows, and for IBin the raw rows beside them
    const uint32_t slab_bytes = ibin ? (uint32_t)round_up(32 * no_out * 4 + (any_raw ? 32 * d->no * 4 : 0), 16)
                                     : (uint32_t)round_up(32 * d->no * 4, 16);
    const int epi_warps = ibin ? 12 : 4;
    const int ncg = ibin ? 1 : 2;      // converter groups (IBin's 12 epilogue warps leave threads / registers for one)
    const size_t fixed = 1024 + (size_t)4 * slab_bytes + 256;
    int stages = TS_MAX_STAGES;
    while (stages > 2 && fixed + (size_t)stages * TS_STAGE_BYTES > 227 * 1024) --stages;
    const size_t smem_bytes = fixed + (size_t)stages * TS_STAGE_BYTES;
    YC_REQUIRE(smem_bytes <= 227 * 1024, YC_ERR_UNSUPPORTED, "tcgen05 fp32 head: needs %zu bytes of shared memory", smem_bytes);
    unsigned fit = 0;
    for (int i = 0; i < d->nl; ++i) {
        const yc_head_level &lv = d->level[i];
        const size_t HW = (size_t)lv.H * lv.W;
        const bool ok = (HW * 4) % 16 == 0 && ((uintptr_t)lv.x & 15) == 0 && (d->kind != YC_HEAD_RAW || lv.raw);
        if (ok) fit |= 1u << i;
        else set_error("level %d (K=%d, H*W=%zu) does not meet the TMA alignment rules", i, lv.K, HW);
    }
    *left_mask = ((1u << d->nl) - 1u) & ~fit;
    if (!fit) return YC_ERR_UNSUPPORTED;
    int order[YC_MAX_LEVELS], n = 0;
    for (int i = 0; i < d->nl; ++i)
        if (fit >> i & 1u) order[n++] = i;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (d->level[order[j]].K > d->level[order[i]].K) { int t = order[i]; order[i] = order[j]; order[j] = t; }

    TcMaps maps;
    TcParams P;
    memset(&P, 0, sizeof(P));
    P.n_lv = n;
    P.bs = d->bs; P.na = 1; P.no = d->no; P.npad = npad;
    P.epi_warps = epi_warps;
    P.na_real = d->na; P.no_out = no_out;
    P.rows_total = rows_total;
    P.npad_g = npad_g;
    P.write_z = d->kind != YC_HEAD_RAW ? 1 : 0;
    if (ibin) {
        P.ibin = 1; P.bin_count = d->bin_count;
        P.bin_step = (float)(4.0 / (double)d->bin_count);
        P.bins = d->bins;
    }
    P.z = d->z;
    P.idesc = instr_desc_f16(/*f16*/ 0, /*A MN-major*/ 1, /*B MN-major*/ 1, (uint32_t)TC_BM, (uint32_t)npad);
    P.b_box_bytes = 64 * TS_BK * 2;    // one {64 n, 32 k} fp16 box
    P.slab_bytes = slab_bytes;
    P.stages = stages;
    P.debug = dbg;
    int tiles = 0;
    for (int s = 0; s < n; ++s) {
        const int i = order[s];
        const yc_head_level &lv = d->level[i];
        const int HW = lv.H * lv.W;
        BlobView bv = blob_view(lv.blob, N, lv.K);
        TcLevel &L = P.lv[s];
        L.sb = bv.sb_split;
        L.raw = lv.raw;
        L.K = lv.K; L.HW = HW; L.nx = lv.W;
        L.tiles_per_img = (HW + TC_BM - 1) / TC_BM;
        L.n_groups = d->na;
        L.tile_begin = tiles;
        L.row_off = row_off[i];
        L.stride = lv.stride;
        L.stride_y = lv.stride_y > 0.f ? lv.stride_y : lv.stride;
        for (int j = 0; j < YC_MAX_ANCHORS * 2; ++j) L.anchor_wh[j] = lv.anchor_wh[j];
        tiles += d->bs * L.tiles_per_img * d->na;
        {   // A: X [bs, K, HW] float32, box {32 px, 32 k, 1}
            cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)lv.K, (cuuint64_t)d->bs};
            cuuint64_t gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)HW * lv.K * 4};
            cuuint32_t box[3] = {32, (cuuint32_t)TS_BK, 1}, est[3] = {1, 1, 1};
            CUresult r = enc(&maps.a[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)lv.x, gdim, gstr, box, est,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            YC_REQUIRE(r == CUDA_SUCCESS, YC_ERR_CUDA, "cuTensorMapEncodeTiled(A f32, level %d) failed: %d", i, (int)r);
        }
        for (int h = 0; h < 2; ++h) {   // B: w_hi_t / w_lo_t [K, na*npad_g] fp16 (weights transposed, per-anchor padded), box {64 n, 32 k}
            cuuint64_t gdim[2] = {(cuuint64_t)wt, (cuuint64_t)lv.K};
            cuuint64_t gstr[1] = {(cuuint64_t)wt * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)TS_BK}, est[2] = {1, 1};
            CUresult r = enc(&maps.b[2 * s + h], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)(h ? bv.w_lo_t : bv.w_hi_t), gdim,
                             gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            YC_REQUIRE(r == CUDA_SUCCESS, YC_ERR_CUDA, "cuTensorMapEncodeTiled(B f16, level %d) failed: %d", i, (int)r);
        }
    }
    P.total_tiles = tiles;
    const int grid = tiles < num_sms ? tiles : num_sms;
    const int threads = 32 * (4 + 4 * ncg + epi_warps);
    void (*kern)(const TcMaps, const TcParams) = ncg == 2 ? head_tcs_kernel<2> : head_tcs_kernel<1>;
    YC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kern<<<grid, threads, smem_bytes, stream>>>(maps, P);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

} // namespace yc
