// yc_head_sm100.cu -- the head's 1x1 output convolution as a persistent, warp-specialised
// tcgen05 / TMEM / TMA kernel for sm_100a, with ImplicitM scale + bias, sigmoid and the grid/anchor
// box decode fused into the epilogue (reference nets/idetect.py:31-45 in one pass).
//
// Formulation.  Per level the conv is D[p, c] = sum_k X[b, k, p] * W[c, k]  (p = pixel, c = a*no + o).
//   A operand = feature map, straight from NCHW: pixels are contiguous, so A is "MN-major"; a TMA box
//               {64 px, 64 k} lands as [64 k-rows][128 B] with the 128-byte swizzle -- the canonical
//               MN-major SWIZZLE_128B UMMA layout.  Two boxes make the 128-pixel M tile.
//   B operand = packed bf16 weights [Npad, K], K-major, TMA box {64 k, Npad} with the 128-byte swizzle.
//   D         = 128 lanes (pixels) x Npad fp32 columns in TMEM, double buffered (2 x 256 columns).
// With pixels on TMEM lanes, one epilogue thread owns one output row (b, a, y, x, 0..no): a warp's
// 32 rows are ONE contiguous span of z (32*no*4 bytes), staged in shared memory and written with a
// single bulk (TMA) store.  The permute(0,1,3,4,2) of the reference therefore costs nothing.
//
// Warp roles (128 + 128*na threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warp 3 idle, then 4 epilogue warps per anchor (warp%4 = TMEM lane quadrant).
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), static
// round-robin tile scheduler over all (level, image, 128-pixel block) tiles, heaviest level first.

#include "yc_head_tc.cuh"

namespace yc {

// DBG instantiations honour the YC_TC_DEBUG timing switches (bit 1 skip epilogue work, 2 skip MMA issue, 4 skip TMA,
// 8 print clock64 wait statistics of CTA 0); the production instantiation carries none of that code.
// AK: the feature maps are channels-last ([bs, H, W, K], e.g. the output of a channels-last RepConv): A is then a K-major
// operand, one {64 k, 128 px} TMA box per stage, instead of the MN-major operand NCHW maps give.
// FUSED: the fused step (epilogue emits NMS candidates; capped at TC_MAX_REGS registers so that the NMS kernels of the
// previous batch fit beside it); otherwise the z / raw writing forward, which keeps up to 64 accumulator values of a
// half row in registers next to the sigmoids in flight and gets the whole register file (512 threads x 128).
// IBIN: the IBin head; a separate instantiation so that its 64-register half-row loads do not weigh on the register
// allocation of the other heads' epilogues.
template <int BK, bool DBG, bool AK, bool FUSED, bool IBIN>
__global__ void __maxnreg__(FUSED ? TC_MAX_REGS : 128)
head_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams P)
{
    constexpr int TC_BK = BK;
    constexpr int TC_A_BYTES = TC_BM * BK * 2;
    // one weight box (64 k x Npad rows) takes P.b_slot_bytes of a stage: 32 KB for the 256-column tiles, 16 KB for IBin's
    // one-anchor tiles, which buys those a deeper ring
    const int TC_B_SLOT = (int)P.b_slot_bytes;
    const int TC_STAGE_BYTES = TC_A_BYTES + (BK / 64) * TC_B_SLOT;
    extern __shared__ uint8_t smem_raw[];
    // carve: [stages: A|B] (1024-aligned) [slabs] [barriers]
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *stage_base = smem;
    const int n_stages = P.stages;
    float *slabs = (float *)(smem + n_stages * TC_STAGE_BYTES);
    const int n_epi_warps = P.epi_warps;
    const int n_slabs = (IBIN && !P.half_off) ? 4 : n_epi_warps;   // whole-row IBin epilogue: one slab per TMEM lane quadrant, shared by its three warps
    float2 *sbtab = (float2 *)((uint8_t *)slabs + (size_t)n_slabs * P.slab_bytes);   // half-row epilogue: (scale, bias) of every level
    uint64_t *bars = (uint64_t *)(sbtab + P.tab_entries);
    uint64_t *full_bar = bars;                        // [TC_MAX_STAGES]
    uint64_t *empty_bar = bars + TC_MAX_STAGES;       // [TC_MAX_STAGES]
    uint64_t *tfull_bar = bars + 2 * TC_MAX_STAGES;   // [2]
    uint64_t *tempty_bar = bars + 2 * TC_MAX_STAGES + 2; // [2]
    uint32_t *tmem_ptr_smem = (uint32_t *)(bars + 2 * TC_MAX_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < P.n_lv; ++i) {
            prefetch_tmap(&maps.a[i]);
            for (int g = 0; g < P.lv[i].n_groups; ++g) prefetch_tmap(&maps.b[P.lv[i].bmap0 + g]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < n_stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], (uint32_t)n_epi_warps);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr_smem, TC_TMEM_COLS);
    if (P.tab_entries) {   // level s, column c of the head at [s * na_real * no + c]
        const int n_head = P.na_real * P.no;
        for (int i = threadIdx.x; i < P.tab_entries; i += blockDim.x) sbtab[i] = __ldg(P.lv[i / n_head].sb + i % n_head);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // Producer and MMA issuer: the WHOLE warp runs the loop (uniform control flow keeps barrier addresses and
    // descriptors in uniform registers) and one elected lane issues the TMA / tcgen05 instructions.  Issuing from
    // inside an `if (lane == 0)` region instead makes ptxas wrap every uniform-datapath instruction in an
    // ELECT / R2UR / BRA.U.ANY loop and rebuild both descriptors with ~15 dependent uniform ops per MMA: measured
    // ~165 cycles per issued MMA against the 128-cycle tensor pipe floor (tools/mma_probe.cu: 130 cycles with
    // descriptors formed by one 64-bit add).
    const bool skip_epi = DBG && (P.debug & 1), skip_mma = DBG && (P.debug & 2), skip_tma = DBG && (P.debug & 4);
    const bool prof = DBG && (P.debug & 8) && blockIdx.x == 0;
    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        // Weight residency: when a level has exactly n_stages k-blocks and a tile starts at ring slot 0,
        // k-block kb of W always lands in slot kb.  After one such tile the B halves of all slots hold
        // the whole W of that level, so the following tiles of the same level load only A (the MMA
        // thread is unaware: it reads B from the slot as usual; a slot's B half is only ever written
        // by this warp, after the slot's empty barrier).
        int resident_lv = -1;
        long long p_wait = 0, p_issue = 0, p_t0 = DBG ? clock64() : 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
            const TileCoord tc = tile_coord(P, t);
            const int nkb = (P.lv[tc.lv].K + TC_BK - 1) / TC_BK;
            const bool aligned = nkb == n_stages && stage == 0;
            const int wkey = tc.lv * YC_MAX_ANCHORS + tc.g; // identifies the weight tile
            const bool load_b = !(aligned && resident_lv == wkey);
            const CUtensorMap *ma = &maps.a[tc.lv], *mb = &maps.b[P.lv[tc.lv].bmap0 + tc.g];
            const uint32_t tx = (uint32_t)TC_A_BYTES + (load_b ? (BK / 64) * P.b_box_bytes : 0u);
            for (int kb = 0; kb < nkb; ++kb) {
                const long long w0 = prof ? clock64() : 0;
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                const long long w1 = prof ? clock64() : 0;
                if (prof) p_wait += w1 - w0;
                if (elect_one()) {
                    uint8_t *sa = stage_base + stage * TC_STAGE_BYTES, *sb = sa + TC_A_BYTES;
                    if (skip_tma) {
                        mbar_arrive(&full_bar[stage]);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], tx);
                        if (AK) {   // rows = pixels of the flat [bs*HW, K] view; a tile's tail rows may belong to the next image
#pragma unroll
                            for (int j = 0; j < BK / 64; ++j)
                                tma_load_2d(sa + j * (TC_A_BYTES / (BK / 64)), ma, &full_bar[stage], kb * TC_BK + j * 64,
                                            tc.b * P.lv[tc.lv].HW + tc.p0);
                        } else {
                            tma_load_3d(sa, ma, &full_bar[stage], tc.p0, kb * TC_BK, tc.b);
                            tma_load_3d(sa + TC_A_BYTES / 2, ma, &full_bar[stage], tc.p0 + 64, kb * TC_BK, tc.b);
                        }
                        if (load_b) {
#pragma unroll
                            for (int j = 0; j < BK / 64; ++j)
                                tma_load_2d(sb + j * TC_B_SLOT, mb, &full_bar[stage], kb * TC_BK + j * 64, 0);
                        }
                    }
                }
                __syncwarp();
                if (prof) p_issue += clock64() - w1;
                if (++stage == n_stages) { stage = 0; phase ^= 1u; }
            }
            resident_lv = aligned ? wkey : -1;
        }
        if (prof && lane == 0) printf("[yc prof] producer: total %lld cyc, waiting on empty %lld, issuing %lld\n", clock64() - p_t0, p_wait, p_issue);
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        long long m_wt = 0, m_wf = 0, m_is = 0, m_t0 = DBG ? clock64() : 0;
        int m_kb = 0;
        // A: MN-major SW128: 64-px chunks LBO = BK*128 B apart (one {64 px, BK k} box each), 8-k groups SBO = 1024 B
        //    apart; one k16 step = two 8-k groups = 2048 B
        // B: K-major SW128, one box per 64 k: 8-row groups SBO = 1024 B apart; k16 step = 32 B inside the 128-B row
        // The start-address field holds (address >> 4) in the low 14 bits: stage / k offsets are plain adds.
        const uint32_t s0 = smem_addr(stage_base);
        // (AK: A is K-major like B -- one {64 k, 128 px} box per 64 k: 8-row groups SBO = 1024 B apart, k16 step = 32 B)
        const uint64_t da0 = AK ? smem_desc(s0, 16, 1024, SWZ_128B) : smem_desc(s0, TC_A_BYTES / 2, 1024, SWZ_128B);
        const uint64_t db0 = smem_desc(s0 + TC_A_BYTES, 16, 1024, SWZ_128B);
        const uint32_t idesc = P.idesc;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
            const TileCoord tc = tile_coord(P, t);
            const int nkb = (P.lv[tc.lv].K + TC_BK - 1) / TC_BK;
            const int buf = it & 1;
            long long w0 = prof ? clock64() : 0;
            mbar_wait(&tempty_bar[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u); // epilogue drained this buffer
            if (prof) m_wt += clock64() - w0;
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)buf * TC_MAX_N;
            for (int kb = 0; kb < nkb; ++kb) {
                w0 = prof ? clock64() : 0;
                mbar_wait(&full_bar[stage], phase);
                const long long w1 = prof ? clock64() : 0;
                if (prof) { m_wf += w1 - w0; ++m_kb; }
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t so = (uint64_t)((uint32_t)(stage * TC_STAGE_BYTES) >> 4);
                    if (!skip_mma) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k)
                            mma_f16(tmem_d, da0 + so + (uint64_t)((AK ? (k / 4) * (TC_BM * 128) + (k % 4) * 32 : k * 2048) >> 4),
                                    db0 + so + (uint64_t)(((k / 4) * TC_B_SLOT + (k % 4) * 32) >> 4), idesc,
                                    (uint32_t)((kb | k) != 0));
                    }
                    mma_commit(&empty_bar[stage]); // frees the smem slot when these MMAs retire
                    if (kb == nkb - 1) mma_commit(&tfull_bar[buf]);
                }
                __syncwarp();
                if (prof) m_is += clock64() - w1;
                if (++stage == n_stages) { stage = 0; phase ^= 1u; }
            }
        }
        if (prof && lane == 0)
            printf("[yc prof] mma: total %lld cyc, %d tiles %d k-blocks, waiting on tmem-empty %lld, on smem-full %lld, issuing %lld\n",
                   clock64() - m_t0, it, m_kb, m_wt, m_wf, m_is);
    } else if (warp >= TC_NON_EPI_THREADS / 32) {
        // ===================== epilogue: TMEM -> sigmoid/decode -> slab -> bulk store =====================
        const int e = warp - TC_NON_EPI_THREADS / 32;
        const int q = warp & 3;     // TMEM lane quadrant this warp may read
        const int a = IBIN ? 0 : e >> 2;         // anchor of the tile handled by this warp (IBin: one anchor per tile)
        float *slab = (float *)((uint8_t *)slabs + (size_t)e * P.slab_bytes);
        int it = 0, cur_key = -1;
        BoxSb sbv;
        long long pf[4] = {0, 0, 0, 0}, pf_wait = 0, pf_t0 = DBG ? clock64() : 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
            const TileCoord tc = tile_coord(P, t);
            const TcLevel &L = P.lv[tc.lv];
            const int buf = it & 1;
            const int prow0 = tc.p0 + 32 * q;            // first pixel of this warp's 32 rows
            const int nv = min(32, L.HW - prow0);        // valid rows (<= 0: nothing to store)
            const int ar = tc.g * P.na + a;               // anchor of the head this warp decodes
            const float2 *sb = L.sb + ar * P.no;
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * TC_MAX_N + a * P.no);

            // fused mode: the (scale, bias) pairs this warp needs live in registers; they change with the level and,
            // when the anchors of a pixel block are separate tiles (na*no > 256 columns), with the anchor group
            if (FUSED && !IBIN && tc.lv * YC_MAX_ANCHORS + tc.g != cur_key) {
                sbv = load_box_sb(sb, lane, P.nc);
                cur_key = tc.lv * YC_MAX_ANCHORS + tc.g;
            }
            const long long w0 = prof ? clock64() : 0;
            mbar_wait(&tfull_bar[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            if (prof) pf_wait += clock64() - w0;
            if (skip_epi) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[buf]);
                continue;
            }
            if (FUSED) {
                if (IBIN && P.half_off) {   // eight warps: (quadrant, half of its rows)
                    const int pass16 = e >> 2;
                    fused_epilogue_ibin_half<22, false>(P, L, tc.b, prow0 + 16 * pass16, nv - 16 * pass16, ar, taddr, pass16,
                                                 smem_addr(sbtab + (tc.lv * P.na_real + ar) * P.no), smem_addr(slab), &tempty_bar[buf], lane);
                    continue;
                }
                if (IBIN) fused_epilogue_ibin(P, L, tc.b, prow0, nv, ar, taddr, slab, &tempty_bar[buf], lane);
                else fused_epilogue<false>(P, L, tc.b, prow0, nv, ar, taddr, slab, &tempty_bar[buf], lane, sbv);
                continue;
            }
            if (IBIN && P.half_off) {   // eight warps: (quadrant, half of its rows)
                const int pass = e >> 2;
                if (P.no >= 127)
                    store_rows_half_ibin<22, false, true>(P, L, tc.b, prow0 + 16 * pass, nv - 16 * pass, ar, taddr + ((uint32_t)(16 * pass) << 16),
                                                          smem_addr(sbtab + (tc.lv * P.na_real + ar) * P.no), smem_addr(slab),
                                                          smem_addr(bars + 16), &tempty_bar[buf], lane);
                else
                    store_rows_half_ibin<22, false, false>(P, L, tc.b, prow0 + 16 * pass, nv - 16 * pass, ar, taddr + ((uint32_t)(16 * pass) << 16),
                                                           smem_addr(sbtab + (tc.lv * P.na_real + ar) * P.no), smem_addr(slab),
                                                           smem_addr(bars + 16), &tempty_bar[buf], lane);
                continue;
            }
            if (!IBIN && P.half_off) {
                store_rows_half_any<false>(P, L, tc.b, prow0, nv, ar, taddr, smem_addr(sbtab + (tc.lv * P.na_real + ar) * P.no),
                                           smem_addr(slab), smem_addr(bars + 16), &tempty_bar[buf], lane, prof ? pf : nullptr);
                continue;
            }
            store_epilogue<false>(P, L, tc.b, tc.p0, tc.g, e, q, lane, tmem_base + (uint32_t)(buf * TC_MAX_N), (uint8_t *)slabs,
                                  &tempty_bar[buf], prof ? pf : nullptr);
        }
        if (prof && lane == 0)
            printf("[yc prof] epilogue warp %d: %d tiles, total %lld cyc: waiting tfull %lld, slab read-out %lld, tmem+decode %lld, "
                   "store issue %lld, named barriers %lld\n", e, it, clock64() - pf_t0, pf_wait, pf[0], pf[1], pf[2], pf[3]);
        if (lane == 0) bulk_wait_all0(); // global writes complete before the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// ---- host side ------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int g_num_sms = 0;
static int g_reserved_sms = 0;   // yc_reserve_sms: SMs the persistent head kernels leave alone

int set_reserved_sms(int n)
{
    const int old = g_reserved_sms;
    g_reserved_sms = n < 0 ? 0 : n;
    return old;
}

int launch_head_tc2(const TcMaps &maps, TcParams &P, int num_sms, cudaStream_t stream); // yc_head_sm100_2cta.cu
int launch_head_tc2i(const TcMaps &maps, TcParams &P, int num_sms, cudaStream_t stream); // yc_head_sm100_2cta_ibin.cu
int launch_head_split(const yc_head_desc *d, int rows_total, const int *row_off, unsigned *left_mask, void *enc_fn, int num_sms,
                      cudaStream_t stream);                                              // yc_head_sm100_split.cu

int launch_head_tcgen05(const yc_head_desc *d, int rows_total, const int *row_off, unsigned *left_mask,
                        const FusedDetect *fused, cudaStream_t stream)
{
    const int N = d->na * d->no;
    // MMA tile: all anchors side by side when they fit 256 accumulator columns (IDetect: 3 x 85), otherwise one
    // anchor per tile (IBin: 127 -> 128 columns), the anchors of a pixel block being consecutive tiles
    const bool ibin = d->kind == YC_HEAD_IBIN;
    const int na_tile = (!ibin && round_up(N, 16) <= TC_MAX_N) ? d->na : 1;
    const int n_groups = d->na / na_tile;
    const int npad = round_up(na_tile * d->no, 16);
    const int npad_total = round_up(N, 16);
    YC_REQUIRE(!(fused && d->x_dtype != YC_BF16), YC_ERR_UNSUPPORTED, "tcgen05 head: the fused step takes bf16 feature maps");
    EncodeTiledFn enc = encode_tiled();
    YC_REQUIRE(enc != nullptr, YC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (!g_num_sms) {
        int dev = 0;
        YC_CUDA(cudaGetDevice(&dev));
        YC_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    YC_REQUIRE(!(d->x_channels_last && d->x_dtype != YC_BF16), YC_ERR_UNSUPPORTED, "tcgen05 head: channels-last maps must be bf16");
    if (d->x_dtype == YC_F32)   // float32 maps: fp16 hi/lo split, three MMAs per k-step, float32-grade result
        return launch_head_split(d, rows_total, row_off, left_mask, (void *)enc, g_num_sms - g_reserved_sms > 0 ? g_num_sms - g_reserved_sms : 1,
                                 stream);
    YC_REQUIRE(npad <= TC_MAX_N && na_tile <= 3, YC_ERR_UNSUPPORTED,
               "tcgen05 head: %d anchors x %d outputs do not fit the 256-column accumulator tile", d->na, d->no);

    // per epilogue warp: z slab (32 rows) or, in the fused mode, the survivor queue (TC_QUEUE_ROWS x nc floats)
    const int no_out = ibin ? d->no - 2 * (d->bin_count + 1) + 2 : d->no;
    // IBin: one slab per quadrant holding the z rows and (if asked for) the raw rows of 32 pixels
    bool any_raw = false;
    for (int i = 0; i < d->nl; ++i) any_raw = any_raw || d->level[i].raw != nullptr;
    // z / raw rows by halves (store_rows_half): 16-row slabs and the (scale, bias) table in shared memory
    int half_off = (!fused && !ibin) ? half_off_for(d->no, na_tile) : 0;
    // IBin by halves (store_rows_half_ibin<22>): the reference's 21 bins, box part and objectness inside the lower 64 columns
    if (ibin && d->bin_count == 21 && d->no > 64 && d->no <= 128) half_off = 64;   // (the fused step too: fused_epilogue_ibin_half)
    { const char *e = getenv("YC_TC_HALF"); if (e && atoi(e) == 0) half_off = 0; }   // experiments: the whole-row epilogue
    const int tab_entries = half_off ? d->nl * N : 0;
    if (tab_entries * 8 > 16 * 1024) half_off = 0;
    const uint32_t slab_bytes = fused ? ((ibin && half_off) ? 512u : (uint32_t)round_up(TC_QUEUE_ROWS * (d->no - 5) * 4, 16))   // (fused IBin by halves: one 128-float row buffer per warp)
                                : ibin ? (half_off ? (uint32_t)round_up(16 * (any_raw ? d->no : no_out) * 4, 16)   // one slab, raw rows then z rows
                                                   : (uint32_t)round_up(32 * no_out * 4 + (any_raw ? 32 * d->no * 4 : 0), 16))
                                       : (uint32_t)round_up((half_off ? 16 : 32) * d->no * 4, 16);
    // fused IBin: one warp per quadrant (most rows stop at the objectness); by halves: two per quadrant; whole rows: three
    const int epi_warps = ibin ? (half_off ? 8 : fused ? 4 : 12) : 4 * na_tile;
    // K=64 per stage: 4 stages in the fused mode (no z slabs in shared memory), 2 next to the slabs.
    // (K=128 x 2 stages measured 6 us slower on the C2 batch.)
    int bk = 64;
    // CTA pairs (cta_group::2) for the fused step: see yc_head_sm100_2cta.cu (measured 2-4 % faster than the 1-CTA
    // kernel for batches of 32 images and more: 8 feature-map stages instead of 4).  YC_TC_2CTA=0 keeps the 1-CTA kernel.
    // (r03) the z-writing forward takes the pair kernel too when its rows go by halves: per tile the 1-CTA kernel streams
    // all of W from L2 (777 MB per C2 batch next to 367 MB of maps) through a ring that the slabs leave 2-3 stages deep
    bool pair = (fused != nullptr || half_off > 0) && n_groups == 1 && npad % 16 == 0;
    // IBin by half rows, NCHW maps: the CTA-pair kernel that reads the maps once per pixel tile for all anchors
    // (yc_head_sm100_2cta_ibin.cu; YC_TC_2CTA=0 keeps the 1-CTA kernel, one (pixel tile, anchor) per tile)
    const bool ibin_pair_ok = ibin && half_off > 0 && !d->x_channels_last && d->na <= 4;
    if (ibin_pair_ok) pair = true;
    { const char *e = getenv("YC_TC_2CTA"); if (e && atoi(e) == 0) pair = false; }
    if (pair) bk = T2_BK; // feature-map box height of the CTA-pair kernel
    const int tile_px = pair ? 2 * TC_BM : TC_BM;
    const uint32_t b_slot_bytes = (uint32_t)round_up(npad * 64 * 2, 1024);   // npad is a multiple of 16: 2 KB steps
    const size_t stage_bytes = (size_t)TC_BM * bk * 2 + (size_t)(bk / 64) * b_slot_bytes;
    const size_t fixed = 1024 + (size_t)(ibin ? (half_off ? 8 : 4) : 4 * na_tile) * slab_bytes + 256 +   // (fused IBin: 4 warps x one area each)
                         (half_off ? (size_t)tab_entries * 8 : 0);
    int stages = TC_MAX_STAGES;
    while (stages > 2 && fixed + (size_t)stages * stage_bytes > 227 * 1024) --stages;
    const size_t smem_bytes = fixed + (size_t)stages * stage_bytes;
    YC_REQUIRE(smem_bytes <= 227 * 1024, YC_ERR_UNSUPPORTED, "tcgen05 head: needs %zu bytes of shared memory", smem_bytes);
    // which levels fit: TMA needs 16-byte aligned bases and row pitches
    unsigned fit = 0;
    for (int i = 0; i < d->nl; ++i) {
        const yc_head_level &lv = d->level[i];
        const size_t HW = (size_t)lv.H * lv.W;
        const bool ok = (d->x_channels_last || (HW * 2) % 16 == 0) && ((size_t)lv.K * 2) % 16 == 0 && ((uintptr_t)lv.x & 15) == 0 &&
                        (d->kind != YC_HEAD_RAW || lv.raw) && !(fused && lv.raw);
        if (ok) fit |= 1u << i;
        else set_error("level %d (K=%d, H*W=%zu) does not meet the TMA alignment rules", i, lv.K, HW);
    }
    *left_mask = ((1u << d->nl) - 1u) & ~fit;
    if (!fit) return YC_ERR_UNSUPPORTED;

    // schedule order: largest K first (longest tiles first)
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
    P.bs = d->bs; P.na = na_tile; P.no = d->no; P.npad = npad;
    P.epi_warps = epi_warps;
    P.na_real = d->na; P.no_out = no_out;
    P.rows_total = rows_total;
    P.write_z = d->kind != YC_HEAD_RAW ? 1 : 0;
    if (ibin) {
        P.ibin = 1; P.bin_count = d->bin_count;
        P.bin_step = (float)(4.0 / (double)d->bin_count); // SigmoidBin(min=0, max=4), nets/ibin.py:16-17
        P.bins = d->bins;
    }
    P.z = d->z;
    P.idesc = instr_desc_f16(/*bf16*/ 1, /*A: MN-major (NCHW) or K-major (channels-last)*/ d->x_channels_last ? 0 : 1,
                             /*B K-major*/ 0, (uint32_t)tile_px, (uint32_t)npad);
    P.a_kmajor = d->x_channels_last ? 1 : 0;
    P.b_box_bytes = (uint32_t)(pair ? npad / 2 : npad) * 64 * 2;
    P.b_slot_bytes = b_slot_bytes;
    P.slab_bytes = slab_bytes;
    P.stages = stages;
    P.half_off = half_off;
    P.tab_entries = half_off ? n * N : 0;   // levels in schedule order (those that meet the TMA rules)
    { const char *e = getenv("YC_TC_DEBUG"); P.debug = e ? atoi(e) : 0; }
    // feature maps are read once: in the fused step their loads carry the L2 evict-first policy (A/B on one box: 79.0 / 79.3
    // -> 78.3 / 78.4 us per C2 batch; the z-writing forward measured no gain and keeps the default policy)
    P.a_hint = fused ? 1 : 0;
    { const char *e = getenv("YC_TC_AHINT"); if (e) P.a_hint = atoi(e); }
    if (fused) {
        P.fused = 1; P.nc = fused->nc; P.conf = fused->conf; P.div_w = fused->div_w; P.div_h = fused->div_h;
        P.ws = fused->ws;
        P.write_z = 0;
    }
    int tiles = 0;
    for (int s = 0; s < n; ++s) {
        const int i = order[s];
        const yc_head_level &lv = d->level[i];
        const int HW = lv.H * lv.W;
        BlobView bv = blob_view(lv.blob, N, lv.K);
        TcLevel &L = P.lv[s];
        L.sb = bv.sb;
        L.raw = lv.raw;
        L.K = lv.K; L.HW = HW; L.nx = lv.W;
        L.tiles_per_img = (HW + tile_px - 1) / tile_px;
        L.boxes_per_img = (HW + 63) / 64;
        L.n_boxes = d->bs * L.boxes_per_img;
        L.n_groups = n_groups;
        L.bmap0 = s * n_groups;
        L.tile_begin = tiles;
        {   // chunks of the tile order (n_groups > 1): a multiple of the grid, about 24 MB of feature maps, so that the
            // second and third anchor group's pass over a chunk reads it from L2
            const int sms = g_num_sms - g_reserved_sms > 0 ? g_num_sms - g_reserved_sms : 1;
            const size_t tile_bytes = (size_t)lv.K * TC_BM * 2;
            int m = (int)((size_t)(24u << 20) / ((size_t)sms * tile_bytes));
            { const char *e = getenv("YC_TC_CHUNK"); if (e) m = atoi(e); }   // experiments (a huge value: group major over the whole level)
            L.chunk_tiles = sms * (m < 1 ? 1 : m);
        }
        L.row_off = row_off[i];
        L.stride = lv.stride;
        L.stride_y = lv.stride_y > 0.f ? lv.stride_y : lv.stride;
        for (int j = 0; j < YC_MAX_ANCHORS * 2; ++j) L.anchor_wh[j] = lv.anchor_wh[j];
        tiles += pair ? (L.n_boxes + 3) / 4 : d->bs * L.tiles_per_img * n_groups;   // (pair + IBin: a tile holds all anchors)
        if (d->x_channels_last) {   // A: X [bs*HW, K] bf16 (channels-last), box {64 k, 128 px} (pair kernel: 64 px)
            cuuint64_t gdim[2] = {(cuuint64_t)lv.K, (cuuint64_t)d->bs * HW};
            cuuint64_t gstr[1] = {(cuuint64_t)lv.K * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)(pair ? 64 : TC_BM)}, est[2] = {1, 1};
            CUresult r = enc(&maps.a[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)lv.x, gdim, gstr, box, est,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            YC_REQUIRE(r == CUDA_SUCCESS, YC_ERR_CUDA, "cuTensorMapEncodeTiled(A channels-last, level %d) failed: %d", i, (int)r);
        } else {   // A: X [bs, K, HW] bf16, box {64 px, 64 k, 1}
            cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)lv.K, (cuuint64_t)d->bs};
            cuuint64_t gstr[2] = {(cuuint64_t)HW * 2, (cuuint64_t)HW * lv.K * 2};
            cuuint32_t box[3] = {64, (cuuint32_t)bk, 1}, est[3] = {1, 1, 1};
            CUresult r = enc(&maps.a[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)lv.x, gdim, gstr, box, est,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            YC_REQUIRE(r == CUDA_SUCCESS, YC_ERR_CUDA, "cuTensorMapEncodeTiled(A, level %d) failed: %d", i, (int)r);
        }
        for (int g = 0; g < n_groups; ++g) {   // B: rows [g*na_tile*no, ...) of W [Npad_total, K] bf16, box {64 k, npad}
            const int row0 = g * na_tile * d->no;  // rows past Npad_total are zero-filled by TMA
            cuuint64_t gdim[2] = {(cuuint64_t)lv.K, (cuuint64_t)(npad_total - row0)};
            cuuint64_t gstr[1] = {(cuuint64_t)lv.K * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)(pair ? npad / 2 : npad)}, est[2] = {1, 1};
            CUresult r = enc(&maps.b[s * n_groups + g], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                             (void *)(bv.w_bf + (size_t)row0 * lv.K), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            YC_REQUIRE(r == CUDA_SUCCESS, YC_ERR_CUDA, "cuTensorMapEncodeTiled(B, level %d) failed: %d", i, (int)r);
        }
    }
    P.total_tiles = tiles;

    if (pair && ibin) return launch_head_tc2i(maps, P, g_num_sms - g_reserved_sms > 1 ? g_num_sms - g_reserved_sms : 2, stream);
    if (pair) return launch_head_tc2(maps, P, g_num_sms - g_reserved_sms > 1 ? g_num_sms - g_reserved_sms : 2, stream);
    const int sms = g_num_sms - g_reserved_sms > 0 ? g_num_sms - g_reserved_sms : 1;
    const int grid = tiles < sms ? tiles : sms;
    const int threads = TC_NON_EPI_THREADS + 32 * epi_warps;
    // (the YC_TC_DEBUG instantiations exist for NCHW maps only)
    void (*kern)(const TcMaps, const TcParams);
    const bool dbg = P.debug && !P.a_kmajor;
    if (ibin) {
        if (P.fused) kern = P.a_kmajor ? head_tc_kernel<64, false, true, true, true> : dbg ? head_tc_kernel<64, true, false, true, true> : head_tc_kernel<64, false, false, true, true>;
        else kern = P.a_kmajor ? head_tc_kernel<64, false, true, false, true> : dbg ? head_tc_kernel<64, true, false, false, true> : head_tc_kernel<64, false, false, false, true>;
    } else {
        if (P.fused) kern = P.a_kmajor ? head_tc_kernel<64, false, true, true, false> : dbg ? head_tc_kernel<64, true, false, true, false> : head_tc_kernel<64, false, false, true, false>;
        else kern = P.a_kmajor ? head_tc_kernel<64, false, true, false, false> : dbg ? head_tc_kernel<64, true, false, false, false> : head_tc_kernel<64, false, false, false, false>;
    }
    YC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kern<<<grid, threads, smem_bytes, stream>>>(maps, P);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

} // namespace yc
