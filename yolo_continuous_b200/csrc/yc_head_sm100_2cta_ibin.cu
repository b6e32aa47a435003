// yc_head_sm100_2cta_ibin.cu -- the IBin head (nets/ibin.py: 127 accumulator columns per anchor) on CTA pairs, feature maps
// read ONCE per pixel tile for all anchors.
//
// Why not the tile order of the other kernels.  One IBin anchor fills a 128-column accumulator, so the 1-CTA kernel runs
// one (pixel tile, anchor) per tile and walks the pixel tiles once per anchor: the maps go through L2 -> SM three times
// (HBM once, thanks to the chunked order) and the weights of an anchor are reloaded whenever the group changes.  Measured
// on the C5 batch (16 images at 1280x1280, YC_TC_DEBUG): loads alone 152 us, loads + MMAs 160-178 us, whichever kernel
// issues the MMAs (1-CTA M = N = 128 at 134 cycles, or cta_group::2 M = 256, N = 128 at 83: tools/mma_probe) -- the L2 -> SM
// path is the limit, not the tensor pipe.  Here a tile is 256 pixels (four 64-pixel boxes, two per CTA) x ALL anchors:
//   every k-block of the maps is loaded once and multiplied by the k-block of each anchor's weights (G MMAs groups);
//   a CTA holds half of each anchor's weight rows: 64 rows x 64 k = 8 KB per (anchor, k-block); twelve such slots keep
//     all of W for K <= 256 (P3, 3/4 of the pixels) resident, deeper levels stream through the same slots from L2;
//   TMEM is a ring of four 128-column accumulators: anchor g of the pair's it-th tile uses slot (G it + g) mod 4, so with
//     G = 3 the first anchor of the next tile starts while the epilogue still drains the last two of this one.
// Roles per CTA: warp 0 feature-map producer, warp 2 TMEM allocator + weight producer, warps 1 and 3 MMA issuers (leader
// CTA, alternate k-blocks, ordered by a `turn` counter as in yc_head_sm100_2cta.cu), eight epilogue warps: (TMEM lane
// quadrant, half of its rows), the half-row IBin epilogues of yc_head_tc.cuh (z / raw writing, or the fused step).
#include "yc_head_tc.cuh"

namespace yc {

constexpr int TI_A_BYTES = TC_BM * 64 * 2;   // this CTA's 128 pixels x 64 k (two {64 px, 64 k} boxes)
constexpr int TI_W_SLOT = 64 * 64 * 2;       // this CTA's 64 weight rows of one anchor x 64 k
constexpr int TI_W_SLOTS = 12;
constexpr int TI_MAX_A_STAGES = 8;
constexpr int TI_T_SLOTS = 4;                // TMEM accumulators (128 columns each)
constexpr int TI_ACC_COLS = 128;

template <bool DBG, bool FUSED>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(FUSED ? TC_MAX_REGS : 128)
head_tc2i_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams P)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int na_st = P.stages, G = P.na_real;
    uint8_t *a_ring = smem;
    uint8_t *w_slots = smem + na_st * TI_A_BYTES;
    uint8_t *slabs = w_slots + TI_W_SLOTS * TI_W_SLOT;
    const int n_epi_warps = P.epi_warps;   // 8
    float2 *sbtab = (float2 *)(slabs + (size_t)n_epi_warps * P.slab_bytes);
    uint64_t *bars = (uint64_t *)(sbtab + P.tab_entries);
    uint64_t *a_full = bars, *a_empty = a_full + TI_MAX_A_STAGES;
    uint64_t *w_full = a_empty + TI_MAX_A_STAGES, *w_empty = w_full + TI_W_SLOTS;
    uint64_t *tfull = w_empty + TI_W_SLOTS, *tempty = tfull + TI_T_SLOTS;
    uint32_t *tmem_ptr = (uint32_t *)(tempty + TI_T_SLOTS);   // [0] TMEM base, [1] `turn`, [2] scratch word of the epilogues

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader
    const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < P.n_lv; ++i) {
            prefetch_tmap(&maps.a[i]);
            for (int g = 0; g < G; ++g) prefetch_tmap(&maps.b[P.lv[i].bmap0 + g]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < na_st; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < TI_W_SLOTS; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
        }
        for (int i = 0; i < TI_T_SLOTS; ++i) {
            mbar_init(&tfull[i], 2);                                   // one commit from each MMA warp
            mbar_init(&tempty[i], (uint32_t)(2 * n_epi_warps));        // epilogue warps of both CTAs
        }
        *(volatile int *)(tmem_ptr + 1) = 0;
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_ptr, TC_TMEM_COLS);
    {   // (scale, bias) of level s, column c of the head at [s * na_real * no + c]
        const int n_head = P.na_real * P.no;
        for (int i = threadIdx.x; i < P.tab_entries; i += blockDim.x) sbtab[i] = __ldg(P.lv[i / n_head].sb + i % n_head);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync(); // both CTAs' barriers are initialised before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // Every role walks the same deterministic tile sequence t = pair, pair + n_pairs, ... (whole-warp loops, one elected
    // issuing lane).  A level whose weights fit the slots (G * k-blocks <= 12) keeps them: only the first of the pair's
    // consecutive tiles of that level loads them, and the last one hands the slots back.
    const bool skip_epi = DBG && (P.debug & 1), skip_mma = DBG && (P.debug & 2), skip_tma = DBG && (P.debug & 4);
    if (warp == 0) {
        // ===================== feature-map (A) producer, both CTAs =====================
        int sa = 0;
        uint32_t pa = 0;
        for (int t = pair; t < P.total_tiles; t += n_pairs) {
            const BoxTile tc = box_tile(P, t);
            const int nkb = (P.lv[tc.lv].K + 63) / 64;
            int b0, p0, b1, p1;   // this CTA's two 64-pixel boxes (possibly of different images)
            box_coord(P.lv[tc.lv], P.bs, tc.j0 + 2 * (int)rank, b0, p0);
            box_coord(P.lv[tc.lv], P.bs, tc.j0 + 2 * (int)rank + 1, b1, p1);
            const CUtensorMap *ma = &maps.a[tc.lv];
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&a_empty[sa], pa ^ 1u);
                if (elect_one()) {
                    uint8_t *dst = a_ring + sa * TI_A_BYTES;
                    if (skip_tma) {
                        if (rank == 0) mbar_arrive(&a_full[sa]);
                    } else {
                        if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2u * (uint32_t)TI_A_BYTES);
                        tma_load_3d_pair(dst, ma, &a_full[sa], p0, kb * 64, b0);
                        tma_load_3d_pair(dst + TI_A_BYTES / 2, ma, &a_full[sa], p1, kb * 64, b1);
                    }
                }
                __syncwarp();
                if (++sa == na_st) { sa = 0; pa ^= 1u; }
            }
        }
    } else if (warp == 2) {
        // ===================== weight producer, both CTAs: own half of every anchor's rows =====================
        int resident = -1;
        uint32_t pbits = 0; // bit s = parity of the next load into weight slot s
        for (int t = pair; t < P.total_tiles; t += n_pairs) {
            const BoxTile tc = box_tile(P, t);
            const TcLevel &L = P.lv[tc.lv];
            const int nkb = (L.K + 63) / 64;
            const bool fits = nkb * G <= TI_W_SLOTS;
            const bool load_w = !(fits && resident == tc.lv);
            resident = fits ? tc.lv : -1;
            if (!load_w) continue;
            for (int kb = 0; kb < nkb; ++kb) {
                for (int g = 0; g < G; ++g) {
                    const int s = (kb * G + g) % TI_W_SLOTS;
                    mbar_wait(&w_empty[s], ((pbits >> s) & 1u) ^ 1u);
                    pbits ^= 1u << s;
                    if (elect_one()) {
                        if (skip_tma) {
                            if (rank == 0) mbar_arrive(&w_full[s]);
                        } else {
                            if (rank == 0) mbar_arrive_expect_tx(&w_full[s], 2u * P.b_box_bytes);
                            tma_load_2d_pair(w_slots + s * TI_W_SLOT, &maps.b[L.bmap0 + g], &w_full[s], kb * 64, (int)rank * 64);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ===================== MMA issuers (leader CTA): warp 1 the even k-blocks, warp 3 the odd ones =====================
        if (rank == 0) {
            const int me = warp == 3 ? 1 : 0;
            int sa = 0, it = 0, resident = -1, gk = 0;   // gk: k-blocks of this pair so far
            uint32_t pa = 0, pbits = 0;
            // A: MN-major SW128, two {64 px, 64 k} boxes 8 KB apart, k16 step = 2048 B;  B: K-major SW128, k16 step = 32 B
            const uint64_t da0 = smem_desc(smem_addr(a_ring), TI_A_BYTES / 2, 1024, SWZ_128B);
            const uint64_t db0 = smem_desc(smem_addr(w_slots), 16, 1024, SWZ_128B);
            const uint32_t idesc = P.idesc;
            volatile int *turn = (volatile int *)(tmem_ptr + 1);
            for (int t = pair; t < P.total_tiles; t += n_pairs, ++it) {
                const BoxTile tc = box_tile(P, t);
                const int nkb = (P.lv[tc.lv].K + 63) / 64;
                const bool fits = nkb * G <= TI_W_SLOTS;
                const bool load_w = !(fits && resident == tc.lv);
                resident = fits ? tc.lv : -1;
                // resident weights go back to the producer with the LAST tile of the pair that reads them
                bool release_w = true;
                if (fits && t + n_pairs < P.total_tiles) release_w = box_tile(P, t + n_pairs).lv != tc.lv;
                for (int kb = 0; kb < nkb; ++kb, ++gk) {
                    if ((gk & 1) == me) {
                        if (kb == 0) {   // the tile's accumulators: anchor g in slot (G it + g) mod 4, drained by both CTAs
                            for (int g = 0; g < G; ++g) {
                                const int u = G * it + g;
                                mbar_wait(&tempty[u & 3], ((uint32_t)(u >> 2) & 1u) ^ 1u);
                            }
                        }
                        if (load_w)
                            for (int g = 0; g < G; ++g) {
                                const int s = (kb * G + g) % TI_W_SLOTS;
                                mbar_wait(&w_full[s], (pbits >> s) & 1u);
                            }
                        mbar_wait(&a_full[sa], pa);
                        tc_fence_after();
                        const uint64_t da = da0 + (uint64_t)((uint32_t)(sa * TI_A_BYTES) >> 4);
                        while (*turn != gk) { }
                        tc_fence_after();   // the other issuer's MMAs (ordered before its `turn` store) precede ours
                        if (elect_one()) {
                            for (int g = 0; g < G; ++g) {
                                const int s = (kb * G + g) % TI_W_SLOTS;
                                const uint32_t tmem_d = tmem_base + (uint32_t)(((G * it + g) & 3) * TI_ACC_COLS);
                                const uint64_t db = db0 + (uint64_t)((uint32_t)(s * TI_W_SLOT) >> 4);
                                if (!skip_mma) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        mma_f16_pair(tmem_d, da + (uint64_t)((k * 2048) >> 4), db + (uint64_t)((k * 32) >> 4), idesc,
                                                     (uint32_t)((kb | k) != 0));
                                }
                            }
                            tc_fence_before();
                            __threadfence_block();
                            *turn = gk + 1;
                            mma_commit_pair(&a_empty[sa]);
                            if (release_w)
                                for (int g = 0; g < G; ++g) mma_commit_pair(&w_empty[(kb * G + g) % TI_W_SLOTS]);
                            if (kb == nkb - 1)   // this warp's share of the tile is issued (the other warp commits after the loop)
                                for (int g = 0; g < G; ++g) mma_commit_pair(&tfull[(G * it + g) & 3]);
                        }
                        __syncwarp();
                    }
                    if (load_w)
                        for (int g = 0; g < G; ++g) pbits ^= 1u << ((kb * G + g) % TI_W_SLOTS);
                    if (++sa == na_st) { sa = 0; pa ^= 1u; }
                }
                // every MMA of the tile precedes the later of the two commits (a warp without a k-block in this tile still
                // counts); the warp that issued the last k-block has committed there
                if (((gk - 1) & 1) != me) {
                    if (elect_one())
                        for (int g = 0; g < G; ++g) mma_commit_pair(&tfull[(G * it + g) & 3]);
                    __syncwarp();
                }
            }
        }
    } else if (warp >= TC_NON_EPI_THREADS / 32) {
        // ===================== epilogue (both CTAs): eight warps, (quadrant, half of its rows), anchor after anchor =========
        const int e = warp - TC_NON_EPI_THREADS / 32;
        const int q = warp & 3, pass16 = e >> 2;
        const uint32_t slab_s = smem_addr(slabs + (size_t)e * P.slab_bytes), dummy_s = smem_addr(tmem_ptr + 2);
        int it = 0;
        for (int t = pair; t < P.total_tiles; t += n_pairs, ++it) {
            const BoxTile tc = box_tile(P, t);
            const TcLevel &L = P.lv[tc.lv];
            int img, pbox;   // this warp's quadrant is half of one 64-pixel box
            box_coord(L, P.bs, tc.j0 + 2 * (int)rank + (q >> 1), img, pbox);
            const int prow0 = pbox + 32 * (q & 1) + 16 * pass16;
            const int nv = (img < P.bs ? min(32, L.HW - (pbox + 32 * (q & 1))) : 0) - 16 * pass16;
            if (FUSED && !skip_epi) {
                // the accumulators of a tile's anchors complete together (k-block major MMA order): probe the objectness of
                // all of them behind ONE wait, so that the anchors without a survivor go back to the MMA warps at once
                uint32_t o[TI_T_SLOTS];
                unsigned surv[TI_T_SLOTS];
#pragma unroll
                for (int g = 0; g < TI_T_SLOTS; ++g) {
                    if (g < G) {
                        const int u = G * it + g;
                        mbar_wait(&tfull[u & 3], (uint32_t)(u >> 2) & 1u);
                    }
                }
                tc_fence_after();
#pragma unroll
                for (int g = 0; g < TI_T_SLOTS; ++g)
                    if (g < G)
                        TmemLd<1>::ld(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(((G * it + g) & 3) * TI_ACC_COLS + 2 + 2 * 22), &o[g]);
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < TI_T_SLOTS; ++g)
                    if (g < G)
                        surv[g] = ibin_obj_survivors<22>(o[g], smem_addr(sbtab + (tc.lv * P.na_real + g) * P.no), pass16, nv, P.conf, lane);
                if (DBG && (P.debug & 16)) {   // timing experiments: no survivor path
#pragma unroll
                    for (int g = 0; g < TI_T_SLOTS; ++g) surv[g] = 0;
                }
                // anchors without a survivor first (their accumulators go back at once), then the others one by one: each
                // fetches its rows, hands the accumulator back and only then decodes
#pragma unroll
                for (int round = 0; round < 2; ++round) {
#pragma unroll
                    for (int g = 0; g < TI_T_SLOTS; ++g)
                        if (g < G && (surv[g] != 0) == (round == 1)) {
                            const int slot = (G * it + g) & 3;
                            fused_epilogue_ibin_half_tail<22, true>(P, L, img, prow0, nv, g, tmem_base + ((uint32_t)(32 * q) << 16) +
                                                                    (uint32_t)(slot * TI_ACC_COLS), pass16,
                                                                    smem_addr(sbtab + (tc.lv * P.na_real + g) * P.no), slab_s, &tempty[slot],
                                                                    lane, surv[g]);
                        }
                }
                continue;
            }
            for (int g = 0; g < G; ++g) {
                const int u = G * it + g, slot = u & 3;
                mbar_wait(&tfull[slot], (uint32_t)(u >> 2) & 1u);
                tc_fence_after();
                if (skip_epi) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&tempty[slot]);
                    continue;
                }
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(slot * TI_ACC_COLS);
                const uint32_t tab_s = smem_addr(sbtab + (tc.lv * P.na_real + g) * P.no);
                if (FUSED)
                    fused_epilogue_ibin_half<22, true>(P, L, img, prow0, nv, g, taddr, pass16, tab_s, slab_s, &tempty[slot], lane);
                else if (P.no >= 127)
                    store_rows_half_ibin<22, true, true>(P, L, img, prow0, nv, g, taddr + ((uint32_t)(16 * pass16) << 16), tab_s, slab_s,
                                                         dummy_s, &tempty[slot], lane);
                else
                    store_rows_half_ibin<22, true, false>(P, L, img, prow0, nv, g, taddr + ((uint32_t)(16 * pass16) << 16), tab_s, slab_s,
                                                          dummy_s, &tempty[slot], lane);
            }
        }
        if (!FUSED && lane == 0) bulk_wait_all0();   // global writes complete before the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync(); // the peer's shared memory / TMEM stay alive until the leader's MMAs are done with them
    tc_fence_after();
    if (warp == 2) tmem_dealloc_pair(tmem_base, TC_TMEM_COLS);
}

// host side: descriptors prepared by launch_head_tcgen05 (yc_head_sm100.cu); P.total_tiles counts 256-pixel tiles
int launch_head_tc2i(const TcMaps &maps, TcParams &P, int num_sms, cudaStream_t stream)
{
    const size_t fixed = 1024 + (size_t)P.epi_warps * P.slab_bytes + 512 + (size_t)TI_W_SLOTS * TI_W_SLOT + (size_t)P.tab_entries * 8;
    // fused step: at most 214 KB, the rest of the SM's shared memory stays free for the NMS kernels of the previous batch
    const size_t cap = (P.fused ? 214 : 227) * 1024;
    int stages = TI_MAX_A_STAGES;
    while (stages > 2 && fixed + (size_t)stages * TI_A_BYTES > cap) --stages;
    const size_t smem_bytes = fixed + (size_t)stages * TI_A_BYTES;
    YC_REQUIRE(smem_bytes <= 227 * 1024, YC_ERR_UNSUPPORTED, "2-CTA IBin head: needs %zu bytes of shared memory", smem_bytes);
    P.stages = stages;
    void (*kern)(const TcMaps, const TcParams);
    if (P.fused) kern = P.debug ? head_tc2i_kernel<true, true> : head_tc2i_kernel<false, true>;
    else kern = P.debug ? head_tc2i_kernel<true, false> : head_tc2i_kernel<false, false>;
    YC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int pairs = num_sms / 2;
    if (P.total_tiles < pairs) pairs = P.total_tiles;
    kern<<<2 * pairs, TC_NON_EPI_THREADS + 32 * P.epi_warps, smem_bytes, stream>>>(maps, P);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

} // namespace yc
