// yc_head_sm100_2cta.cu -- CTA-pair (tcgen05 cta_group::2) version of the fused head kernel.
//
// Why a pair: with one CTA the 128x256x16 MMA reads A (4 KB) and all of B (8 KB) from its own shared memory,
// 96 B/clk against a 128 B/clk port that TMA is also filling -- measured ~165 cycles per MMA instead of 128
// (profiles/README.md).  In a pair, each CTA keeps its own 128 pixels of A and only HALF of the weight rows;
// one cta_group::2 MMA (M = 256 pixels, N = 256) issued by the leader reads A and B from both CTAs and writes
// each CTA's 128 x 256 fp32 accumulator into that CTA's TMEM.  Per CTA and MMA: 4 KB + 4 KB.
// Half a weight k-block (64 k) being 16 KB, four weight slots hold all of W for K <= 256 (P3 of the COCO head, 3/4 of
// the tiles), so that level streams only feature maps; deeper levels stream their weights through the same slots
// from L2.  The feature-map ring is separate and as deep as fits (8 stages): the step is latency bound on feature-map
// bytes in flight, and this is what the pair buys over the 1-CTA kernel (B 128 KB resident + 4 stages there).
//
// Roles per CTA: warp 0 feature-map producer (own A tile), warp 2 TMEM allocator then weight producer (own half of
// the weight rows), bytes accounted on the LEADER's barriers; warps 1 and 3 MMA issuers (leader only, alternating
// tiles / TMEM buffers); 12 epilogue warps (fused epilogue of yc_head_tc.cuh).
// Barriers: a_full/b_full live in the leader (both producers' TMA complete there); a_empty/b_empty/tfull are
// signalled in both CTAs by multicast tcgen05.commit; tempty lives in the leader and counts the epilogue warps
// of both CTAs.
#include "yc_head_tc.cuh"

namespace yc {

constexpr int T2_A_BYTES = TC_BM * T2_BK * 2;  // this CTA's 128 pixels x BK k (two {64 px, BK k} boxes)
constexpr int T2_B_BOX = 128 * 64 * 2;         // 16 KB: this CTA's (up to) 128 weight rows x 64 k
constexpr int T2_B_BYTES = (T2_BK / 64) * T2_B_BOX;
#ifndef T2_B_SLOTS_K
#define T2_B_SLOTS_K 256                       // weight slots cover K <= this (resident weights): 64 KB per CTA, which
                                               // leaves 8 feature-map stages (128 KB in flight per CTA); with 512
                                               // (P4 resident too, 4-5 stages) the kernel measured 4 % slower
#endif
constexpr int T2_B_SLOTS = T2_B_SLOTS_K / T2_BK;
constexpr int T2_MAX_A_STAGES = 8;

struct T2Ring { // carved identically in both CTAs
    uint8_t *a_ring, *b_slots;
    float *queues;
    uint64_t *a_full, *a_empty, *b_full, *b_empty, *tfull, *tempty;
    uint32_t *tmem_ptr;
};

// advance a ring position by n stages
__device__ __forceinline__ void ring_advance(int &s, uint32_t &parity, int n, int depth)
{
    s += n;
    while (s >= depth) { s -= depth; parity ^= 1u; }
}

// AK: channels-last feature maps (A K-major): each 64-pixel box is one {64 k, 64 px} TMA box of the flat [bs*HW, K] view
// FUSED: see head_tc_kernel (register cap of the fused step; the z-writing forward gets 128 registers)
template <bool DBG, bool AK, bool FUSED>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(FUSED ? TC_MAX_REGS : 128)
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
    float2 *sbtab = (float2 *)((uint8_t *)R.queues + (size_t)n_epi_warps * P.slab_bytes);   // z-writing mode: (scale, bias) table
    uint64_t *bars = (uint64_t *)(sbtab + P.tab_entries);
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
            mbar_init(&R.tfull[i], 2);   // one commit from each MMA warp
            mbar_init(&R.tempty[i], (uint32_t)(2 * n_epi_warps)); // epilogue warps of both CTAs
        }
        *(volatile int *)(R.tmem_ptr + 1) = 0;   // `turn` counter of the two MMA warps
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair(R.tmem_ptr, TC_TMEM_COLS);
    if (P.tab_entries) {   // level s, column c of the head at [s * na_real * no + c]
        const int n_head = P.na_real * P.no;
        for (int i = threadIdx.x; i < P.tab_entries; i += blockDim.x) sbtab[i] = __ldg(P.lv[i / n_head].sb + i % n_head);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync(); // both CTAs' barriers are initialised before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem_base = *R.tmem_ptr;

    // Every role walks the same deterministic tile sequence t = pair, pair + n_pairs, ... and derives from it the
    // ring positions and `load_b` (whether the weight tile has to be (re)loaded: a level whose k-blocks all fit the
    // weight slots keeps its W resident, so only the first of consecutive tiles with the same weight tile loads it).
    // All role loops are whole-warp loops with one elected issuing lane (see yc_head_sm100.cu for why).
    const bool skip_epi = DBG && (P.debug & 1), skip_mma = DBG && (P.debug & 2), skip_tma = DBG && (P.debug & 4);
    if (warp == 0) {
        // ===================== feature-map (A) producer, both CTAs =====================
        int sa = 0;
        uint32_t pa = 0;
        const uint64_t pol = l2_policy_evict_first();
        for (int t = pair; t < P.total_tiles; t += n_pairs) {
            const BoxTile tc = box_tile(P, t);
            const int nkb = (P.lv[tc.lv].K + T2_BK - 1) / T2_BK;
            int b0, p0, b1, p1;   // this CTA's two 64-pixel boxes (possibly of different images)
            box_coord(P.lv[tc.lv], P.bs, tc.j0 + 2 * (int)rank, b0, p0);
            box_coord(P.lv[tc.lv], P.bs, tc.j0 + 2 * (int)rank + 1, b1, p1);
            const CUtensorMap *ma = &maps.a[tc.lv];
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&R.a_empty[sa], pa ^ 1u);
                if (elect_one()) {
                    uint8_t *dst = R.a_ring + sa * T2_A_BYTES;
                    if (skip_tma) {
                        if (rank == 0) mbar_arrive(&R.a_full[sa]);
                    } else {
                        if (rank == 0) mbar_arrive_expect_tx(&R.a_full[sa], 2u * (uint32_t)T2_A_BYTES);
                        if (AK) {   // image index bs (past the end) -> rows past the tensor: TMA zero-fills
                            const int HW = P.lv[tc.lv].HW;
                            tma_load_2d_pair(dst, ma, &R.a_full[sa], kb * T2_BK, b0 * HW + p0);
                            tma_load_2d_pair(dst + T2_A_BYTES / 2, ma, &R.a_full[sa], kb * T2_BK, b1 * HW + p1);
                        } else if (P.a_hint) {
                            tma_load_3d_pair_hint(dst, ma, &R.a_full[sa], p0, kb * T2_BK, b0, pol);
                            tma_load_3d_pair_hint(dst + T2_A_BYTES / 2, ma, &R.a_full[sa], p1, kb * T2_BK, b1, pol);
                        } else {
                            tma_load_3d_pair(dst, ma, &R.a_full[sa], p0, kb * T2_BK, b0);
                            tma_load_3d_pair(dst + T2_A_BYTES / 2, ma, &R.a_full[sa], p1, kb * T2_BK, b1);
                        }
                    }
                }
                __syncwarp();
                if (++sa == na_st) { sa = 0; pa ^= 1u; }
            }
        }
    } else if (warp == 2) {
        // ===================== weight (B) producer, both CTAs: own half of the weight rows =====================
        int resident = -1;
        uint32_t pbits = 0; // bit s = parity of the next load into weight slot s
        for (int t = pair; t < P.total_tiles; t += n_pairs) {
            const BoxTile tc = box_tile(P, t);
            const TcLevel &L = P.lv[tc.lv];
            const int nkb = (L.K + T2_BK - 1) / T2_BK;
            const int wkey = tc.lv;   // the fused mode has one anchor group: the level identifies the weight tile
            const bool load_b = !(nkb <= T2_B_SLOTS && resident == wkey);
            resident = nkb <= T2_B_SLOTS ? wkey : -1;
            if (!load_b) continue;
            const CUtensorMap *mb = &maps.b[L.bmap0];
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % T2_B_SLOTS;
                mbar_wait(&R.b_empty[s], ((pbits >> s) & 1u) ^ 1u);
                pbits ^= 1u << s;
                if (elect_one()) {
                    if (skip_tma) {
                        if (rank == 0) mbar_arrive(&R.b_full[s]);
                    } else {
                        if (rank == 0) mbar_arrive_expect_tx(&R.b_full[s], 2u * (T2_BK / 64) * P.b_box_bytes);
#pragma unroll
                        for (int j = 0; j < T2_BK / 64; ++j)
                            tma_load_2d_pair(R.b_slots + s * T2_B_BYTES + j * T2_B_BOX, mb, &R.b_full[s], kb * T2_BK + j * 64,
                                             (int)rank * (P.npad / 2));
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ===================== MMA issuers (leader CTA): warp 1 issues the even k-blocks of every tile, warp 3 the
        // odd ones.  One stage hand-over (barrier poll, fence, election, commits) costs an issuing warp ~300 cycles
        // in which the tensor pipe, whose queue is only ~2 instructions deep, runs dry (measured: MMA-only time =
        // 128 cycles x MMAs + 300 cycles x stages); with two warps the hand-over of one hides behind the MMAs of the
        // other.  Both accumulate into the same TMEM tile and the k-blocks must be added in a fixed order (fp32
        // accumulation is order dependent: results stay bit-identical to the 1-CTA kernel and run to run), so the
        // warps pass a `turn` counter (k-blocks issued so far) through shared memory: everything but the MMA issue
        // itself is done before a warp's turn comes.  Each warp commits its own MMAs (tfull counts 2).
        if (rank == 0) {
            const int me = warp == 3 ? 1 : 0;
            int sa = 0, it = 0, resident = -1, g = 0;   // g: global k-block counter of the pair
            uint32_t pa = 0, pbits = 0;
            const uint64_t da0 = AK ? smem_desc(smem_addr(R.a_ring), 16, 1024, SWZ_128B)
                                    : smem_desc(smem_addr(R.a_ring), T2_A_BYTES / 2, 1024, SWZ_128B);
            const uint64_t db0 = smem_desc(smem_addr(R.b_slots), 16, 1024, SWZ_128B);
            const uint32_t idesc = P.idesc;
            volatile int *turn = (volatile int *)(R.tmem_ptr + 1);
            for (int t = pair; t < P.total_tiles; t += n_pairs, ++it) {
                const BoxTile tc = box_tile(P, t);
                const int nkb = (P.lv[tc.lv].K + T2_BK - 1) / T2_BK;
                const int wkey = tc.lv;
                const bool load_b = !(nkb <= T2_B_SLOTS && resident == wkey);
                resident = nkb <= T2_B_SLOTS ? wkey : -1;
                // A weight slot is handed back to the producer by the LAST tile that reads it: a resident weight tile
                // (all k-blocks in the slots) is read by every following tile of the same level without a reload, so
                // the release waits for the tile after which the producer loads again (the tile sequence is
                // deterministic: look one tile ahead).  Streaming levels reuse the slots inside a tile and release
                // per k-block.  One release per load keeps the producer's phase bookkeeping exact.
                bool release_b = true;
                if (nkb <= T2_B_SLOTS && t + n_pairs < P.total_tiles) release_b = box_tile(P, t + n_pairs).lv != wkey;
                const int buf = it & 1;
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * TC_MAX_N;
                // both CTAs' epilogues drained the buffer (also orders this warp's tfull arrival after the previous
                // phase of tfull[buf], which the epilogues waited for)
                mbar_wait(&R.tempty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = kb % T2_B_SLOTS;
                    if ((g & 1) == me) {
                        if (load_b) mbar_wait(&R.b_full[s], (pbits >> s) & 1u);
                        mbar_wait(&R.a_full[sa], pa);
                        tc_fence_after();
                        const uint64_t da = da0 + (uint64_t)((uint32_t)(sa * T2_A_BYTES) >> 4);
                        const uint64_t db = db0 + (uint64_t)((uint32_t)(s * T2_B_BYTES) >> 4);
                        while (*turn != g) { }
                        tc_fence_after();   // the other issuer's MMAs (ordered before its `turn` store) precede ours
                        if (elect_one()) {
                            if (!skip_mma) {
#pragma unroll
                                for (int k = 0; k < T2_BK / 16; ++k)
                                    mma_f16_pair(tmem_d, da + (uint64_t)((AK ? k * 32 : k * 2048) >> 4),
                                                 db + (uint64_t)(((k / 4) * T2_B_BOX + (k % 4) * 32) >> 4), idesc,
                                                 (uint32_t)((kb | k) != 0));
                            }
                            tc_fence_before();  // tcgen05 memory model: cross-thread MMA -> MMA order needs the fence pair
                            __threadfence_block();
                            *turn = g + 1;
                            mma_commit_pair(&R.a_empty[sa]);
                            if (release_b) mma_commit_pair(&R.b_empty[s]);
                        }
                        __syncwarp();
                    }
                    if (load_b) pbits ^= 1u << s;
                    if (++sa == na_st) { sa = 0; pa ^= 1u; }
                }
                // every MMA of the tile precedes the later of the two commits (a warp without a k-block in this
                // tile still counts)
                if (elect_one()) mma_commit_pair(&R.tfull[buf]);
                __syncwarp();
            }
        }
    } else if (warp >= TC_NON_EPI_THREADS / 32) {
        // ===================== fused epilogue (both CTAs) =====================
        const int e = warp - TC_NON_EPI_THREADS / 32;
        const int q = warp & 3, a = e >> 2;
        float *slab = (float *)((uint8_t *)R.queues + (size_t)e * P.slab_bytes);
        int it = 0, cur_lv = -1;
        BoxSb sbv;
        const bool eprof = DBG && (P.debug & 8) && blockIdx.x == 0;
        long long e_wait = 0, e_work = 0, e_max = 0;
        int e_slow = 0;
        for (int t = pair; t < P.total_tiles; t += n_pairs, ++it) {
            const BoxTile tc = box_tile(P, t);
            const TcLevel &L = P.lv[tc.lv];
            const int buf = it & 1;
            int img, pbox;   // this warp's 32 rows are half of one 64-pixel box
            box_coord(L, P.bs, tc.j0 + 2 * (int)rank + (q >> 1), img, pbox);
            const int prow0 = pbox + 32 * (q & 1);
            const int nv = img < P.bs ? min(32, L.HW - prow0) : 0;
            const int ar = a;
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * TC_MAX_N + a * P.no);
            if (FUSED && tc.lv != cur_lv) { // one anchor group per tile, so `ar` is fixed for this warp
                sbv = load_box_sb(L.sb + ar * P.no, lane, P.nc);
                cur_lv = tc.lv;
            }
            const long long e0 = eprof ? clock64() : 0;
            mbar_wait(&R.tfull[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const long long e1 = eprof ? clock64() : 0;
            if (skip_epi) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&R.tempty[buf]);
                continue;
            }
            if (FUSED)
                fused_epilogue<true>(P, L, img, prow0, nv, ar, taddr, slab, &R.tempty[buf], lane, sbv);
            else   // z / raw maps (the drop-in forward): rows by halves through the warp's 16-row slab
                store_rows_half_any<true>(P, L, img, prow0, nv, ar, taddr, smem_addr(sbtab + (tc.lv * P.na_real + ar) * P.no),
                                          smem_addr(slab), smem_addr(bars + 40), &R.tempty[buf], lane);
            if (eprof) {
                const long long e2 = clock64();
                e_wait += e1 - e0;
                e_work += e2 - e1;
                if (e2 - e1 > e_max) e_max = e2 - e1;
                if (e2 - e1 > 600) ++e_slow;
            }
        }
        if (eprof && lane == 0)
            printf("[yc prof2] epilogue warp %d: %d tiles, waiting tfull %lld, working %lld (max %lld per tile, %d tiles > 600)\n",
                   e, it, e_wait, e_work, e_max, e_slow);
        if (!FUSED && lane == 0) bulk_wait_all0();   // global writes complete before the CTA exits
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
    const size_t fixed = 1024 + (size_t)4 * P.na * P.slab_bytes + 512 + (size_t)T2_B_SLOTS * T2_B_BYTES + (size_t)P.tab_entries * 8;
    // fused step: at most 214 KB -- the rest of the SM's 228 KB stays free for the NMS kernels of the previous batch, which
    // run next to this kernel on a second stream (largest of them: 17.5 KB + 1 KB reserved per CTA)
    const size_t cap = (P.fused ? 214 : 227) * 1024;
    int stages = T2_MAX_A_STAGES;
    while (stages > 2 && fixed + (size_t)stages * T2_A_BYTES > cap) --stages;
    const size_t smem_bytes = fixed + (size_t)stages * T2_A_BYTES;
    YC_REQUIRE(smem_bytes <= 227 * 1024, YC_ERR_UNSUPPORTED, "2-CTA head: needs %zu bytes of shared memory", smem_bytes);
    P.stages = stages;
    void (*kern)(const TcMaps, const TcParams);
    if (P.fused)
        kern = P.a_kmajor ? (P.debug ? head_tc2_kernel<true, true, true> : head_tc2_kernel<false, true, true>)
                          : (P.debug ? head_tc2_kernel<true, false, true> : head_tc2_kernel<false, false, true>);
    else
        kern = P.a_kmajor ? (P.debug ? head_tc2_kernel<true, true, false> : head_tc2_kernel<false, true, false>)
                          : (P.debug ? head_tc2_kernel<true, false, false> : head_tc2_kernel<false, false, false>);
    YC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int pairs = num_sms / 2;
    if (P.total_tiles < pairs) pairs = P.total_tiles;
    kern<<<2 * pairs, TC_NON_EPI_THREADS + 128 * P.na, smem_bytes, stream>>>(maps, P);
    YC_CUDA(cudaGetLastError());
    return YC_OK;
}

} // namespace yc
