// yc_head_sm100_2cta.cu -- CTA-pair (tcgen05 cta_group::2) version of the fused head kernel.
//
// Why a pair: with one CTA the 128x256x16 MMA reads A (4 KB) and all of B (8 KB) from its own shared memory,
// 96 B/clk against a 128 B/clk port that TMA is also filling -- measured ~165 cycles per MMA instead of 128
// (profiles/README.md).  In a pair, each CTA keeps its own 128 pixels of A and only HALF of the weight rows;
// one cta_group::2 MMA (M = 256 pixels, N = 256) issued by the leader reads A and B from both CTAs and writes
// each CTA's 128 x 256 fp32 accumulator into that CTA's TMEM.  Per CTA and MMA: 4 KB + 4 KB.
// Half a weight k-block being 16 KB, eight weight slots hold all of W for K <= 512 (P3 and P4 of the COCO
// head), so those levels stream only feature maps; the feature-map ring is separate and as deep as fits.
//
// Roles per CTA: warp 0 TMA producer (own A tile + own half of B, bytes accounted on the LEADER's barriers),
// warp 1 MMA issuer (leader only), warp 2 TMEM allocator, 12 epilogue warps (fused epilogue of yc_head_tc.cuh).
// Barriers: a_full/b_full live in the leader (both producers' TMA complete there); a_empty/b_empty/tfull are
// signalled in both CTAs by multicast tcgen05.commit; tempty lives in the leader and counts the epilogue warps
// of both CTAs.
#include "yc_head_tc.cuh"

namespace yc {

#ifndef T2_BK
#define T2_BK 128                              // k per stage: the multicast commits cost ~250 cycles per stage
#endif
constexpr int T2_A_BYTES = TC_BM * T2_BK * 2;  // this CTA's 128 pixels x BK k (two {64 px, BK k} boxes)
constexpr int T2_B_BOX = 128 * 64 * 2;         // 16 KB: this CTA's (up to) 128 weight rows x 64 k
constexpr int T2_B_BYTES = (T2_BK / 64) * T2_B_BOX;
#ifndef T2_B_SLOTS_K
#define T2_B_SLOTS_K 256                       // weight slots cover K <= this (resident weights)
#endif
constexpr int T2_B_SLOTS = T2_B_SLOTS_K / T2_BK;
constexpr int T2_MAX_A_STAGES = 8;

struct T2Ring { // carved identically in both CTAs
    uint8_t *a_ring, *b_slots;
    float *queues;
    uint64_t *a_full, *a_empty, *b_full, *b_empty, *tfull, *tempty;
    uint32_t *tmem_ptr;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_NON_EPI_THREADS + 128 * 3, 1)
head_tc2_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams P)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int na_st = P.stages;
    T2Ring R;
    R.a_ring = smem;
    R.b_slots = smem + na_st * T2_A_BYTES;
    R.queues = (float *)(R.b_slots + T2_B_SLOTS * T2_B_BYTES);
    const int n_epi_warps = 4 * P.na;
    uint64_t *bars = (uint64_t *)((uint8_t *)R.queues + (size_t)n_epi_warps * P.slab_bytes);
    R.a_full = bars;
    R.a_empty = R.a_full + T2_MAX_A_STAGES;
    R.b_full = R.a_empty + T2_MAX_A_STAGES;
    R.b_empty = R.b_full + T2_B_SLOTS;
    R.tfull = R.b_empty + T2_B_SLOTS;
    R.tempty = R.tfull + 2;
    R.tmem_ptr = (uint32_t *)(R.tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader
    const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < P.n_lv; ++i) {
            prefetch_tmap(&maps.a[i]);
            prefetch_tmap(&maps.b[P.lv[i].bmap0]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < na_st; ++i) {
            mbar_init(&R.a_full[i], 1);
            mbar_init(&R.a_empty[i], 1);
        }
        for (int i = 0; i < T2_B_SLOTS; ++i) {
            mbar_init(&R.b_full[i], 1);
            mbar_init(&R.b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&R.tfull[i], 1);
            mbar_init(&R.tempty[i], (uint32_t)(2 * n_epi_warps)); // epilogue warps of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair(R.tmem_ptr, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync(); // both CTAs' barriers are initialised before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = *R.tmem_ptr;

    // Weight slot of k-block kb is kb % 8 in every tile; a slot's parity bit flips on every reload, so the
    // slots need no common ring position.  A level with <= 8 k-blocks keeps its whole W (this CTA's half) in
    // the slots: consecutive tiles with the same weight tile skip the weight loads altogether.  Producer (both
    // CTAs) and MMA issuer derive `load_b` from the same deterministic tile sequence.
    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int sa = 0, it = 0, resident = -1;
            uint32_t pa = 0, pbits = 0; // pbits: bit s = parity of the next load into weight slot s
            const bool prof = (P.debug & 8) && blockIdx.x < 2;
            long long w_a = 0, w_b = 0, t0 = clock64();
            for (int t = pair; t < P.total_tiles; t += n_pairs, ++it) {
                const TileCoord tc = tile_coord_w(P, t, 2 * TC_BM);
                const TcLevel &L = P.lv[tc.lv];
                const int nkb = (L.K + T2_BK - 1) / T2_BK;
                const int wkey = tc.lv * YC_MAX_ANCHORS + tc.g;
                const bool load_b = !(nkb <= T2_B_SLOTS && resident == wkey);
                const int p_own = tc.p0 + TC_BM * (int)rank;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (load_b) {
                        const int s = kb % T2_B_SLOTS;
                        long long c0 = prof ? clock64() : 0;
                        mbar_wait(&R.b_empty[s], ((pbits >> s) & 1u) ^ 1u);
                        if (prof) w_b += clock64() - c0;
                        if (P.debug & 4) {
                            if (rank == 0) mbar_arrive(&R.b_full[s]);
                        } else {
                            if (rank == 0) mbar_arrive_expect_tx(&R.b_full[s], 2u * (T2_BK / 64) * P.b_box_bytes);
#pragma unroll
                            for (int j = 0; j < T2_BK / 64; ++j)
                                tma_load_2d_pair(R.b_slots + s * T2_B_BYTES + j * T2_B_BOX, &maps.b[L.bmap0 + tc.g], &R.b_full[s],
                                                 kb * T2_BK + j * 64, (int)rank * (P.npad / 2));
                        }
                        pbits ^= 1u << s;
                    }
                    long long c1 = prof ? clock64() : 0;
                    mbar_wait(&R.a_empty[sa], pa ^ 1u);
                    if (prof) w_a += clock64() - c1;
                    uint8_t *dst = R.a_ring + sa * T2_A_BYTES;
                    if (P.debug & 4) {
                        if (rank == 0) mbar_arrive(&R.a_full[sa]);
                    } else {
                        if (rank == 0) mbar_arrive_expect_tx(&R.a_full[sa], 2u * (uint32_t)T2_A_BYTES);
                        tma_load_3d_pair(dst, &maps.a[tc.lv], &R.a_full[sa], p_own, kb * T2_BK, tc.b);
                        tma_load_3d_pair(dst + T2_A_BYTES / 2, &maps.a[tc.lv], &R.a_full[sa], p_own + 64, kb * T2_BK, tc.b);
                    }
                    if (++sa == na_st) { sa = 0; pa ^= 1u; }
                }
                resident = nkb <= T2_B_SLOTS ? wkey : -1;
            }
            if (prof)
                printf("[yc prof2] producer rank %u: total %lld cyc, waiting a_empty %lld, b_empty %lld\n", rank,
                       clock64() - t0, w_a, w_b);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA, one lane) =====================
        if (lane == 0 && rank == 0) {
            int sa = 0, it = 0, resident = -1;
            uint32_t pa = 0, pbits = 0;
            const bool prof = (P.debug & 8) && blockIdx.x == 0;
            long long w_t = 0, w_a = 0, w_b = 0, t0 = clock64();
            int n_kb = 0;
            for (int t = pair; t < P.total_tiles; t += n_pairs, ++it) {
                const TileCoord tc = tile_coord_w(P, t, 2 * TC_BM);
                const TcLevel &L = P.lv[tc.lv];
                const int nkb = (L.K + T2_BK - 1) / T2_BK;
                const int wkey = tc.lv * YC_MAX_ANCHORS + tc.g;
                const bool load_b = !(nkb <= T2_B_SLOTS && resident == wkey);
                const int buf = it & 1;
                long long c0 = prof ? clock64() : 0;
                mbar_wait(&R.tempty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u); // both CTAs' epilogues drained the buffer
                if (prof) w_t += clock64() - c0;
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * TC_MAX_N;
                for (int kb = 0; kb < nkb; ++kb) {
                    const int s = kb % T2_B_SLOTS;
                    if (load_b) {
                        c0 = prof ? clock64() : 0;
                        mbar_wait(&R.b_full[s], (pbits >> s) & 1u);
                        if (prof) w_b += clock64() - c0;
                        pbits ^= 1u << s;
                    }
                    c0 = prof ? clock64() : 0;
                    mbar_wait(&R.a_full[sa], pa);
                    if (prof) { w_a += clock64() - c0; ++n_kb; }
                    tc_fence_after();
                    const uint32_t aaddr = smem_addr(R.a_ring + sa * T2_A_BYTES);
                    const uint32_t baddr = smem_addr(R.b_slots + s * T2_B_BYTES);
#pragma unroll
                    for (int k = 0; k < T2_BK / 16; ++k) {
                        const uint64_t da = smem_desc(aaddr + k * 2048, T2_A_BYTES / 2, 1024, SWZ_128B);
                        const uint64_t db = smem_desc(baddr + (k / 4) * T2_B_BOX + (k % 4) * 32, 16, 1024, SWZ_128B);
                        if (!(P.debug & 2)) mma_f16_pair(tmem_d, da, db, P.idesc, (uint32_t)((kb | k) != 0));
                    }
                    mma_commit_pair(&R.a_empty[sa]);
                    if (++sa == na_st) { sa = 0; pa ^= 1u; }
                    if (load_b) mma_commit_pair(&R.b_empty[s]);
                    if (kb == nkb - 1) mma_commit_pair(&R.tfull[buf]);
                }
                resident = nkb <= T2_B_SLOTS ? wkey : -1;
            }
            if (prof)
                printf("[yc prof2] mma: total %lld cyc, %d tiles %d k-blocks, waiting tmem-empty %lld, a_full %lld, b_full %lld\n",
                       clock64() - t0, it, n_kb, w_t, w_a, w_b);
        }
    } else if (warp >= TC_NON_EPI_THREADS / 32) {
        // ===================== fused epilogue (both CTAs) =====================
        const int e = warp - TC_NON_EPI_THREADS / 32;
        const int q = warp & 3, a = e >> 2;
        float *slab = (float *)((uint8_t *)R.queues + (size_t)e * P.slab_bytes);
        int it = 0;
        for (int t = pair; t < P.total_tiles; t += n_pairs, ++it) {
            const TileCoord tc = tile_coord_w(P, t, 2 * TC_BM);
            const TcLevel &L = P.lv[tc.lv];
            const int buf = it & 1;
            const int prow0 = tc.p0 + TC_BM * (int)rank + 32 * q;
            const int nv = min(32, L.HW - prow0);
            const int ar = tc.g * P.na + a;
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * TC_MAX_N + a * P.no);
            mbar_wait(&R.tfull[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            if (P.debug & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&R.tempty[buf]);
                continue;
            }
            fused_epilogue<true>(P, L, tc.b, prow0, nv, ar, taddr, slab, &R.tempty[buf], lane);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync(); // the peer's shared memory / TMEM stay alive until the leader's MMAs are done with them
    tc_fence_after();
    if (warp == 2) tmem_dealloc_pair(tmem_base, TC_TMEM_COLS);
}

// host side: same descriptors as the 1-CTA kernel; only the weight box (half the rows) and the grid differ
int launch_head_tc2(const TcMaps &maps, TcParams &P, int num_sms, cudaStream_t stream)
{
    const size_t fixed = 1024 + (size_t)4 * P.na * P.slab_bytes + 512 + (size_t)T2_B_SLOTS * T2_B_BYTES;
    int stages = T2_MAX_A_STAGES;
    while (stages > 2 && fixed + (size_t)stages * T2_A_BYTES > 227 * 1024) --stages;
    const size_t smem_bytes = fixed + (size_t)stages * T2_A_BYTES;
    YC_REQUIRE(smem_bytes <= 227 * 1024, YC_ERR_UNSUPPORTED, "2-CTA head: needs %zu bytes of shared memory", smem_bytes);
    P.stages = stages;
    YC_CUDA(cudaFuncSetAttribute(head_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int pairs = num_sms / 2;
    if (P.total_tiles < pairs) pairs = P.total_tiles;
    head_tc2_kernel<<<2 * pairs, TC_NON_EPI_THREADS + 128 * P.na, smem_bytes, stream>>>(maps, P);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

} // namespace yc
